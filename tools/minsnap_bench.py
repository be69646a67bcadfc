#!/usr/bin/env python
"""Min-snap batch (BASELINE.json configs[3]): 16,384 waypoint paths (8-20 waypoints, random walk with 1-4 m steps at
z = 1 inside the known-free region of field.bt), params of cfg/planner_interactive.yaml with mode = true.
Times (a) the batched exact KKT solve alone and (b) the whole solve -> sample -> collision check -> insert-waypoint loop
through the host-memory C ABI.  Prints one JSON line."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trajectory_planner_b200 as tp

B = int(os.environ.get("MS_B", "16384"))
m = tp.OccMap.from_tpm(os.path.join(ROOT, "data", "maps", "field.tpm"))
info = m.info()
e = tp.Engine(0)
e.set_map(m)
p = tp.default_poly_params()
p.max_iter = int(os.environ.get("MS_MAX_ITER", "20"))
pt = tp.PolyTraj(e, p)
rng = np.random.default_rng(20261018)
occ, known = m.grid("occupied"), m.grid("known")
kz = int(np.floor((1.0 - info["origin"][2]) / info["res"]))
free = np.argwhere((known[:, :, kz] != 0) & (occ[:, :, kz] == 0))
org, res = np.array(info["origin"]), info["res"]
# vectorised random walks: candidate steps are accepted when the box check at the new waypoint passes (GPU query)
paths = [None] * B
cur = org + (np.column_stack([free[rng.integers(len(free), size=B)], np.full(B, kz)]) + 0.5) * res
cur[:, 2] = 1.0
nw = rng.integers(8, 21, size=B)
pts = [[c] for c in cur]
for step in range(19):
    need = np.array([len(pts[b]) < nw[b] for b in range(B)])
    pending = np.nonzero(need)[0]
    for _try in range(30):
        if len(pending) == 0:
            break
        ang, st = rng.uniform(0, 2 * np.pi, len(pending)), rng.uniform(1, 4, len(pending))
        last = np.array([pts[b][-1] for b in pending])
        cand = last + np.column_stack([st * np.cos(ang), st * np.sin(ang), np.zeros(len(pending))])
        hit = pt.box_collision(cand)
        for q, b in enumerate(pending):
            if not hit[q]:
                pts[b].append(cand[q])
        pending = pending[hit != 0]
paths = [np.array(q) for q in pts if len(q) >= 2]
B = len(paths)
nwp = np.array([len(q) for q in paths])
# (a) solve only
pt.solve_batch(paths[:256])
t0 = time.perf_counter(); sols, status = pt.solve_batch(paths); t_solve = time.perf_counter() - t0
e.profile_enable(True); e.profile_get()
sols, status = pt.solve_batch(paths)
prof = e.profile_get()
# (b) the loop
t0 = time.perf_counter(); res_ = pt.make_plan_batch(paths); t_loop = time.perf_counter() - t0
valid = np.array([r["valid"] for r in res_]); iters = np.array([r["iters"] for r in res_])
out = dict(metric="min-snap polyTrajOctomap batch", batch=B, waypoints=dict(min=int(nwp.min()), mean=float(nwp.mean()), max=int(nwp.max())),
           solve_only=dict(solves_per_s=B / t_solve, wall_ms=1e3 * t_solve, kernel_ms=prof["ms"]["minsnap_solve"], singular=int((status != 0).sum())),
           loop=dict(paths_per_s=B / t_loop, wall_ms=1e3 * t_loop, valid_rate=float(valid.mean()), iters_mean=float(iters.mean()), iters_max=int(iters.max()),
                     max_iter=int(p.max_iter)))
print(json.dumps(out))
