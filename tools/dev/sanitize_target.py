# small end-to-end run of the round-2 kernels for compute-sanitizer --tool memcheck
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import trajectory_planner_b200 as tp, bench
pmap = tp.OccMap.from_tpm(bench.MAP_TPM); eng = tp.Engine(0); eng.set_map(pmap); p = tp.default_params()
off, ctrl = bench.make_workload(tp, pmap, eng.query_points, 640, bench.SEED, p)   # > 444 workers: park lists + k_phase_a
out, res = eng.make_plan_batch(p, off, ctrl)
print("makePlan", (res["status"] == 1).mean())
s = eng.sample_batch(off[:17], ctrl[:off[16]], dt=0.05)
print("sample", len(s["t"]))
fm = tp.OccMap.from_tpm(os.path.join(ROOT, "data", "maps", "field.tpm")); e2 = tp.Engine(0); e2.set_map(fm)
pp = tp.default_poly_params(); pp.max_iter = 4
pt = tp.PolyTraj(e2, pp)
rng = np.random.default_rng(1)
paths = []
for K1 in (2, 3, 5, 9, 14, 20, 40, 64):
    ang, st = rng.uniform(0, 2 * np.pi, K1 - 1), rng.uniform(1, 3, K1 - 1)
    xy = np.vstack([[0, 0], np.cumsum(np.column_stack([st * np.cos(ang), st * np.sin(ang)]), 0)])
    paths.append(np.column_stack([xy, np.full(K1, 1.0)]))
sols, st = pt.solve_batch(paths); print("minsnap", st)
r = pt.make_plan_batch(paths[:6]); print("loop", [x["iters"] for x in r])
r = pt.make_plan_corridor_batch(paths[:5]); print("corridor", [x["iters"] for x in r], [list(x["status"]) for x in r])
r = tp.PolyTraj(eng, pp).make_plan_corridor_batch(paths[:3], occmap=True); print("occmap", [x["iters"] for x in r])
