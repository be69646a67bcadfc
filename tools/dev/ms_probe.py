import os, sys, time, numpy as np
sys.path.insert(0, "/root/repo")
import trajectory_planner_b200 as tp
ROOT="/root/repo"
m = tp.OccMap.from_tpm(os.path.join(ROOT, "data", "maps", "field.tpm"))
e = tp.Engine(0); e.set_map(m)
p = tp.default_poly_params(); pt = tp.PolyTraj(e, p)
rng = np.random.default_rng(1)
paths = []
for b in range(16384):
    k = rng.integers(8, 21)
    q = np.cumsum(rng.uniform(-2, 2, (k, 3)), 0); q[:, 2] = 1.0
    paths.append(q)
for n in (256, 16384, 16384, 16384):
    t0 = time.perf_counter(); pt.solve_batch(paths[:n]); print("solve", n, round(1e3 * (time.perf_counter() - t0), 1), "ms", flush=True)
t0 = time.perf_counter(); off, wp = pt._flat(paths); print("flat", round(1e3 * (time.perf_counter() - t0), 1), "ms")
for n in (16384, 16384):
    t0 = time.perf_counter(); pt.make_plan_batch(paths[:n]); print("loop", n, round(1e3 * (time.perf_counter() - t0), 1), "ms", flush=True)
