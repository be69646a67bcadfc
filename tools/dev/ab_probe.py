# A/B of two builds of the library (TP_B200_LIB chosen by the caller): 4,096-batch wall time (min / median of 8) and the
# sum + median of 64 single solves, same workload as bench.py
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import trajectory_planner_b200 as tp, bench
pmap = tp.OccMap.from_tpm(bench.MAP_TPM); eng = tp.Engine(0); eng.set_map(pmap); p = tp.default_params()
off, ctrl = bench.committed_workload(tp, pmap, p, 4096)
for _ in range(3): eng.make_plan_batch(p, off, ctrl)
w = []
for _ in range(8):
    t = time.perf_counter(); out, res = eng.make_plan_batch(p, off, ctrl); w.append(1e3 * (time.perf_counter() - t))
Ns = np.diff(off); lat = []
for b in range(64):
    o1 = np.array([0, Ns[b]], np.int32); c1 = ctrl[off[b]:off[b + 1]]
    eng.make_plan_batch(p, o1, c1)
    t = time.perf_counter(); eng.make_plan_batch(p, o1, c1); lat.append(1e3 * (time.perf_counter() - t))
print("%s: batch min %.2f median %.2f ms | 64 single solves: sum %.2f ms, p50 %.3f ms | checksum %.12e" % (
    os.environ.get("TP_B200_LIB", "default"), min(w), np.median(w), sum(lat), np.median(lat), float(np.abs(out).sum())))
