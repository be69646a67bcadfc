import os, sys, numpy as np
sys.path.insert(0,'/root/repo')
import trajectory_planner_b200 as tp, bench
pmap = tp.OccMap.from_tpm(bench.MAP_TPM); eng = tp.Engine(0); eng.set_map(pmap); p = tp.default_params()
B = int(os.environ.get("PROBE_B", "4096"))
off, ctrl = bench.make_workload(tp, pmap, eng.query_points, B, bench.SEED + int(os.environ.get("PROBE_SEED_OFFSET", "0")), p)
eng.make_plan_batch(p, off, ctrl)
os.environ["TP_TIMELINE"] = "/tmp/tl.bin"
out, res = eng.make_plan_batch(p, off, ctrl)
tl = np.fromfile("/tmp/tl.bin", dtype=np.int64).reshape(-1, 4)
t0 = tl[:,0].min(); st = (tl[:,0]-t0)/1e6; en = (tl[:,1]-t0)/1e6; dur = en-st
N = np.diff(off)
print("makespan %.1f ms; block duration ms: mean %.2f p50 %.2f p95 %.2f p99 %.2f max %.2f; sum %.0f ms" % (en.max(), dur.mean(), np.median(dur), np.percentile(dur,95), np.percentile(dur,99), dur.max(), dur.sum()))
# concurrency over time
ts = np.linspace(0, en.max(), 29)
for a, b_ in zip(ts[:-1], ts[1:]):
    mid = 0.5*(a+b_); act = ((st <= mid) & (en > mid)).sum()
    print("t=%6.1f ms active blocks %4d" % (mid, act))
order = np.argsort(-en)[:8]
for i in order: print("late finisher: traj %d N %d start %.1f end %.1f dur %.1f iters %d exp %d rounds %d" % (i, N[i], st[i], en[i], dur[i], tl[i,3] & 0xffffffff, tl[i,3] >> 32, res['outer_rounds'][i]))
it = (tl[:,3] & 0xffffffff).astype(float)
print("us per iteration (dur/iters): p50 %.2f p95 %.2f" % (np.median(dur*1e3/np.maximum(it,1)), np.percentile(dur*1e3/np.maximum(it,1),95)))
