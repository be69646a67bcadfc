# ncu target: one makePlan batch on the maze raster (BASELINE configs[2]'s A*-bound half), B trajectories
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import trajectory_planner_b200 as tp, bench
name = os.environ.get("PROBE_MAP", "maze"); B = int(os.environ.get("PROBE_B", "1024"))
m = tp.OccMap.from_tpm(os.path.join(bench.ROOT, "data", "maps", name + ".tpm")); info = m.info()
e = tp.Engine(0); e.set_map(m); p = tp.default_params()
S, G = bench.octomap_pairs(m.grid("inflated"), info, int(B * 1.3) + 64, np.random.default_rng(bench.SEED + 7))
off, ctrl, valid = e.frontend_batch(p, S, G)
keep = np.flatnonzero((valid != 0) & (np.diff(off) >= 7))[:B]
chunks = [ctrl[off[b]:off[b + 1]] for b in keep]
o2 = np.concatenate([[0], np.cumsum([len(c) for c in chunks])]).astype(np.int32); c2 = np.concatenate(chunks, 0)
out, res = e.make_plan_batch(p, o2, c2)
print('ok', len(keep), (res['status'] == 1).mean(), int(res['lbfgs_iters'].sum()), int(res['astar_expansions'].sum()), int(res['astar_searches'].sum()))
