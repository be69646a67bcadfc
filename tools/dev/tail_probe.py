# who finishes last in one bench batch (seed offset from argv): start / end / duration per trajectory from TP_TIMELINE
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import trajectory_planner_b200 as tp, bench
off_ = int(sys.argv[1]) if len(sys.argv) > 1 else 2
pmap = tp.OccMap.from_tpm(bench.MAP_TPM); eng = tp.Engine(0); eng.set_map(pmap); p = tp.default_params()
offsets, ctrl = bench.make_workload(tp, pmap, eng.query_points, 4096, bench.SEED + off_, p)
eng.make_plan_batch(p, offsets, ctrl)
os.environ["TP_TIMELINE"] = "/tmp/tl.bin"
out, r = eng.make_plan_batch(p, offsets, ctrl)
tl = np.fromfile("/tmp/tl.bin", dtype=np.int64).reshape(-1, 4)
t0 = tl[:, 0].min(); st = (tl[:, 0] - t0) / 1e6; en = (tl[:, 1] - t0) / 1e6
sb = np.where(tl[:, 2] > 1000000, (tl[:, 2] - t0) / 1e6, st)   # phase-B start of resumed trajectories
N = np.diff(offsets)
print("makespan %.1f ms" % en.max())
act = [(en > t).sum() for t in np.arange(0, en.max(), 4.0)]
print("active trajectories (not finished) every 4 ms:", act)
for i in np.argsort(-en)[:12]:
    print("traj %d N %d phaseA-start %.1f phaseB-start %.1f busy %.1f end %.1f pairs %d status %d iters %d evals %d exp %d searches %d rounds %d fail %d" % (i, N[i], st[i], sb[i], en[i] - sb[i], en[i], r['n_guide_pairs'][i], r['status'][i], r['lbfgs_iters'][i], r['lbfgs_evals'][i], r['astar_expansions'][i], r['astar_searches'][i], r['outer_rounds'][i], r['fail_count'][i]))
work = r['lbfgs_evals'] * N
print("corr(end time, evals*N) among last 5%:", np.corrcoef(en[np.argsort(-en)[:200]], work[np.argsort(-en)[:200]])[0, 1])
top = np.argsort(-work)[:12]
long_ = np.argsort(-(en - sb))[:15]
print("longest phase B turns:", [(int(i), int(N[i]), round(float(sb[i]), 1), round(float(en[i] - sb[i]), 1), int(r['n_guide_pairs'][i]), int(r['status'][i])) for i in long_])
print("largest evals*N:", [(int(i), int(N[i]), int(r['lbfgs_evals'][i]), round(float(en[i]), 1)) for i in top])
