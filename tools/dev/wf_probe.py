# warp-form kernel probe: batch time, per-iteration time and (with the TP_WF_TIMING build loaded through TP_B200_LIB)
# per-phase cycle totals, for several batch sizes
import os, sys, time, ctypes as C, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import trajectory_planner_b200 as tp, bench
from trajectory_planner_b200 import _capi
pmap = tp.OccMap.from_tpm(bench.MAP_TPM); eng = tp.Engine(0); eng.set_map(pmap); p = tp.default_params()
off_all, ctrl_all = bench.make_workload(tp, pmap, eng.query_points, 4096, bench.SEED, p)
NMAX = int(os.environ.get("PROBE_NMAX", "0"))
if NMAX:   # keep only trajectories with at most NMAX control points, repeated up to 4096
    keep = [b for b in range(4096) if off_all[b + 1] - off_all[b] <= NMAX]
    keep = (keep * (4096 // len(keep) + 1))[:4096]
    chunks = [ctrl_all[off_all[b]:off_all[b + 1]] for b in keep]
    off_all = np.concatenate([[0], np.cumsum([len(c) for c in chunks])]).astype(np.int32)
    ctrl_all = np.concatenate(chunks, 0)
    print("filtered to N <=", NMAX, "mean N", np.diff(off_all).mean())
L = _capi.load()
has = hasattr(L, "tp_debug_wf_phase_get")
out = (C.c_ulonglong * 8)()
for B in [int(x) for x in os.environ.get("PROBE_BS", "1,148,592,1184,4096").split(",")]:
    off = off_all[:B + 1]; ctrl = ctrl_all[:off[B]]
    eng.make_plan_batch(p, off, ctrl)
    if has: L.tp_debug_wf_phase_get(out)
    t = time.perf_counter(); o, r = eng.make_plan_batch(p, off, ctrl); dt = time.perf_counter() - t
    its = int(r["lbfgs_iters"].sum())
    msg = "B %5d: %.2f ms, %d iterations, %.2f us/iteration/batch, max iters of one trajectory %d" % (B, dt * 1e3, its, dt * 1e6 / its, r["lbfgs_iters"].max())
    if has:
        L.tp_debug_wf_phase_get(out)
        v = np.array(list(out), float); it = v[5]
        msg += " | cycles/iter: total %.0f eval %.0f (%.0f/eval) gram %.0f coeffs %.0f direction %.0f other %.0f" % (
            v[0] / it, v[1] / it, v[1] / v[6], v[2] / it, v[3] / it, v[4] / it, (v[0] - v[1] - v[2] - v[3] - v[4]) / it)
    print(msg, flush=True)
