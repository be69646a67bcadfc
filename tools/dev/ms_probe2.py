# why are the first poly calls slow inside bench.py?  octomap step (small) -> [cpu baseline] -> extra_minsnap
import os, sys, time, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, bench
import trajectory_planner_b200 as tp
args = argparse.Namespace(octomap_total=8192, steps=2, warmup=1, mode="fast")
mode = sys.argv[1]
if mode != "none":
    bench.run_octomap(args, tp, torch, None, 0, 1, 0, quiet=True, total=8192, K=2, W=1)
if mode == "cpu":
    bench.octomap_cpu_baseline(16, os.cpu_count() or 1)
r = bench.extra_minsnap(tp, 0, 34.8)
print(mode, "solve wall", r["solve_only"]["wall_ms"], "loop wall", r["loop"]["wall_ms"], "kernel", r["loop"]["kernel_ms"], flush=True)
