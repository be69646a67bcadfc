import sys, numpy as np
sys.path.insert(0,'/root/repo')
import trajectory_planner_b200 as tp, bench
pmap = tp.OccMap.from_tpm(bench.MAP_TPM); eng = tp.Engine(0); eng.set_map(pmap); p = tp.default_params()
for so in (0, 3):
    off, ctrl = bench.make_workload(tp, pmap, eng.query_points, 4096, bench.SEED + so, p)
    out, r = eng.make_plan_batch(p, off, ctrl)
    np_ = r['n_guide_pairs']; it = r['lbfgs_iters']
    print('seed+%d pairs pct50/90/99/max' % so, np.percentile(np_, [50, 90, 99]), np_.max(), 'frac>64 %.4f' % np.mean(np_ > 64), 'iters share of traj with >64 pairs %.3f' % (it[np_ > 64].sum() / it.sum()), 'frac>48 %.4f' % np.mean(np_ > 48))
    top = np.argsort(-it)[:8]
    print('  top iters:', [(int(it[i]), int(np_[i]), int(r['outer_rounds'][i])) for i in top])
