import os, sys, time, numpy as np
sys.path.insert(0,'/root/repo')
import trajectory_planner_b200 as tp, bench
pmap = tp.OccMap.from_tpm(bench.MAP_TPM); eng = tp.Engine(0); eng.set_map(pmap); p = tp.default_params()
z = np.load('/root/repo/tools/_slow_trajs.npz')
for k in ('t3909',):
    c = z[k]; o = np.array([0, len(c)], np.int32)
    eng.make_plan_batch(p, o, c)
    t0 = time.perf_counter(); out, res = eng.make_plan_batch(p, o, c); dt = time.perf_counter() - t0
    print(k, 'ms %.1f' % (dt*1e3), res['status'], res['lbfgs_iters'], res['astar_expansions'], res['astar_searches'])
