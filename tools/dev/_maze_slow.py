import os, sys, time, numpy as np
sys.path.insert(0,'/root/repo')
import trajectory_planner_b200 as tp
name = os.environ.get("PROBE_MAP", "maze")
m = tp.OccMap.from_tpm('/root/repo/data/maps/%s.tpm' % name)
e = tp.Engine(0); e.set_map(m); p = tp.default_params()
z = np.load("/tmp/slow_maze.npz")
for k in list(z.keys())[:int(os.environ.get("PROBE_N", "1"))]:
    c = z[k]; o = np.array([0, len(c)], np.int32)
    t0 = time.perf_counter(); out, r = e.make_plan_batch(p, o, c); dt = time.perf_counter() - t0
    print(k, 'N', len(c), 'ms %.1f' % (dt*1e3), 'status', r['status'], 'iters', r['lbfgs_iters'], 'exp', r['astar_expansions'], 'searches', r['astar_searches'], flush=True)
