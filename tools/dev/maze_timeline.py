import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import trajectory_planner_b200 as tp
name = os.environ.get("PROBE_MAP", "maze"); B = int(os.environ.get("PROBE_B", "4096"))
m = tp.OccMap.from_tpm(os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'data', 'maps', '%s.tpm' % name)); info = m.info()
e = tp.Engine(0); e.set_map(m); p = tp.default_params()
inf = m.grid("inflated"); kz = int(np.floor((1.0 - info["origin"][2]) / info["res"]))
free = np.argwhere(inf[:, :, kz] == 0); rng = np.random.default_rng(20261018); org, res = np.array(info["origin"]), info["res"]
n = int(B * 1.6)
a, b = free[rng.integers(len(free), size=n)], free[rng.integers(len(free), size=n)]
S = org + (np.column_stack([a[:, 0], a[:, 1], np.full(n, kz)]) + 0.5) * res; G = org + (np.column_stack([b[:, 0], b[:, 1], np.full(n, kz)]) + 0.5) * res
S[:, 2] = G[:, 2] = 1.0
d = np.linalg.norm(S - G, axis=1); S, G = S[(d >= 2) & (d <= 20)], G[(d >= 2) & (d <= 20)]
off, ctrl, valid = e.frontend_batch(p, S, G)
keep = [i for i in range(len(S)) if valid[i] and off[i+1]-off[i] >= 7][:B]
o2 = np.concatenate([[0], np.cumsum([off[i+1]-off[i] for i in keep])]).astype(np.int32); c2 = np.concatenate([ctrl[off[i]:off[i+1]] for i in keep])
e.make_plan_batch(p, o2, c2)
os.environ["TP_TIMELINE"] = "/tmp/tl.bin"
out, r = e.make_plan_batch(p, o2, c2)
tl = np.fromfile("/tmp/tl.bin", dtype=np.int64).reshape(-1, 4)
t0 = tl[:,0].min(); st = (tl[:,0]-t0)/1e6; en = (tl[:,1]-t0)/1e6; dur = en - st
print("makespan %.1f ms, sum of durations %.0f ms, success %.3f" % (en.max(), dur.sum(), np.mean(r['status']==1)))
for i in np.argsort(-dur)[:10]:
    print("traj %d N %d start %.1f end %.1f dur %.1f status %d iters %d exp %d searches %d rounds %d" % (i, o2[i+1]-o2[i], st[i], en[i], dur[i], r['status'][i], r['lbfgs_iters'][i], r['astar_expansions'][i], r['astar_searches'][i], r['outer_rounds'][i]))
h = np.histogram(dur, bins=[0,1,2,5,10,20,50,100,200,500,1000])
print("duration histogram (ms):", list(zip(h[1][:-1], h[0])))
top = np.argsort(-dur)[:3]
np.savez("/tmp/slow_maze.npz", **{"t%d" % i: c2[o2[i]:o2[i+1]] for i in top})
