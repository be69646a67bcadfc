# ncu target: one warp-form solve launch with B trajectories (default 148: one worker per SM, the latency picture)
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import trajectory_planner_b200 as tp, bench
pmap = tp.OccMap.from_tpm(bench.MAP_TPM); eng = tp.Engine(0); eng.set_map(pmap); p = tp.default_params()
B = int(os.environ.get("PROBE_B", "148"))
off, ctrl = bench.make_workload(tp, pmap, eng.query_points, B, bench.SEED, p)
out, res = eng.make_plan_batch(p, off, ctrl)
print('ok', (res['status'] == 1).mean(), int(res['lbfgs_iters'].sum()))
