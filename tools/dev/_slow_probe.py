# single-trajectory probe of the batch's slowest solves (TP_ASTAR_TIMING / TP_LBFGS_TIMING builds print per-search cycles)
import os, sys, time, numpy as np
sys.path.insert(0,'/root/repo')
import trajectory_planner_b200 as tp, bench
pmap = tp.OccMap.from_tpm(bench.MAP_TPM); eng = tp.Engine(0); eng.set_map(pmap); p = tp.default_params()
off, ctrl = bench.make_workload(tp, pmap, eng.query_points, 4096, bench.SEED + int(os.environ.get("PROBE_SEED_OFFSET", "0")), p)
for b in [int(x) for x in os.environ.get("PROBE_IDS", "102,1983,3909,1045").split(",")]:
    c = ctrl[off[b]:off[b+1]]; o = np.array([0, len(c)], np.int32)
    eng.make_plan_batch(p, o, c)
    t0 = time.perf_counter(); out, res = eng.make_plan_batch(p, o, c); dt = time.perf_counter() - t0
    print('traj', b, 'N', len(c), 'ms %.1f' % (dt*1e3), 'status', res['status'], 'iters', res['lbfgs_iters'], 'exp', res['astar_expansions'], 'searches', res['astar_searches'], flush=True)
