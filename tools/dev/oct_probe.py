# octomap65536 (BASELINE configs[2]) device-resident timing for the current env settings (TP_WARP_FORM etc.)
import os, sys, json, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import trajectory_planner_b200 as tp
import bench
ap = argparse.ArgumentParser(); ap.add_argument("--total", type=int, default=65536); ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
class A: octomap_total = a.total; steps = a.steps; warmup = 1
torch.cuda.set_device(0)
line = bench.run_octomap(A, tp, torch, dist, 0, 1, 0, quiet=True, total=a.total, K=a.steps, W=1)
print(json.dumps({k: line[k] for k in ("value", "ms_per_step", "gpu_launches")}), {m: (v["success_rate"], v["lbfgs_iters_per_solve"], v["astar_expansions_per_solve"]) for m, v in line["config"]["per_map"].items()})
