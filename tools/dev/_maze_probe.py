import os, sys, numpy as np
sys.path.insert(0,'/root/repo')
import trajectory_planner_b200 as tp
name = os.environ.get("PROBE_MAP", "maze")
m = tp.OccMap.from_tpm('/root/repo/data/maps/%s.tpm' % name); info = m.info()
e = tp.Engine(0); e.set_map(m); p = tp.default_params()
inf = m.grid("inflated"); kz = int(np.floor((1.0 - info["origin"][2]) / info["res"]))
free = np.argwhere(inf[:, :, kz] == 0); rng = np.random.default_rng(3); org, res = np.array(info["origin"]), info["res"]
n = 64
a, b = free[rng.integers(len(free), size=n)], free[rng.integers(len(free), size=n)]
S = org + (np.column_stack([a[:, 0], a[:, 1], np.full(n, kz)]) + 0.5) * res; G = org + (np.column_stack([b[:, 0], b[:, 1], np.full(n, kz)]) + 0.5) * res
S[:, 2] = G[:, 2] = 1.0
off, ctrl, valid = e.frontend_batch(p, S, G)
keep = [i for i in range(n) if valid[i] and off[i+1]-off[i] >= 7][:int(os.environ.get("PROBE_B","8"))]
o2 = np.concatenate([[0], np.cumsum([off[i+1]-off[i] for i in keep])]).astype(np.int32); c2 = np.concatenate([ctrl[off[i]:off[i+1]] for i in keep])
out, r = e.make_plan_batch(p, o2, c2)
print('status', r['status'], 'exp', r['astar_expansions'], 'searches', r['astar_searches'], 'iters', r['lbfgs_iters'])
