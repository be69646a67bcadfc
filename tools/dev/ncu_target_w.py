# ncu target: one warp-form solve launch with B trajectories (default 148: one worker per SM, the latency picture)
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import trajectory_planner_b200 as tp, bench
pmap = tp.OccMap.from_tpm(bench.MAP_TPM); eng = tp.Engine(0); eng.set_map(pmap); p = tp.default_params()
B = int(os.environ.get("PROBE_B", "148"))
off, ctrl = bench.make_workload(tp, pmap, eng.query_points, 4096 if os.environ.get("PROBE_NMAX") else B, bench.SEED, p)
NMAX = int(os.environ.get("PROBE_NMAX", "0"))
if NMAX:   # keep only trajectories with at most NMAX control points, repeated up to B: one size class = ONE launch at full residency
    keep = [b for b in range(len(off) - 1) if off[b + 1] - off[b] <= NMAX]
    keep = (keep * (B // len(keep) + 1))[:B]
    chunks = [ctrl[off[b]:off[b + 1]] for b in keep]
    off = np.concatenate([[0], np.cumsum([len(c) for c in chunks])]).astype(np.int32)
    ctrl = np.concatenate(chunks, 0)
out, res = eng.make_plan_batch(p, off, ctrl)
print('ok', (res['status'] == 1).mean(), int(res['lbfgs_iters'].sum()))
