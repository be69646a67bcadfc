import os, sys, numpy as np
sys.path.insert(0,'/root/repo')
import trajectory_planner_b200 as tp, bench
pmap = tp.OccMap.from_tpm(bench.MAP_TPM); eng = tp.Engine(0); eng.set_map(pmap); p = tp.default_params()
off, ctrl = bench.make_workload(tp, pmap, eng.query_points, 256, bench.SEED, p)
N = np.diff(off)
g = eng.init_guides_batch(p, off, ctrl)
for b in [int(np.argmax(N)), int(np.argsort(N)[len(N)//2]), int(np.argmin(N))]:
    o1 = np.array([0, N[b]], np.int32); c = ctrl[off[b]:off[b+1]]
    gg = (np.array([0, len(g[b]['cp'])], np.int32), g[b]['cp'], g[b]['p'], g[b]['v'])
    print('N', N[b], 'pairs', len(g[b]['cp']))
    eng.optimize_batch(p, o1, c, gg)
