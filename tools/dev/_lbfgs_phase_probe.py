# whole-batch L-BFGS phase totals from the TP_LBFGS_TIMING build (csrc/build_timing.sh with TIMING_LEVEL=1)
import os, sys, ctypes as C, numpy as np
sys.path.insert(0,'/root/repo')
import trajectory_planner_b200 as tp, bench
from trajectory_planner_b200 import _capi
pmap = tp.OccMap.from_tpm(bench.MAP_TPM); eng = tp.Engine(0); eng.set_map(pmap); p = tp.default_params()
B = int(os.environ.get("PROBE_B", "4096"))
off, ctrl = bench.make_workload(tp, pmap, eng.query_points, B, bench.SEED, p)
L = _capi.load()
out = (C.c_ulonglong * 8)()
import contextlib
with open(os.devnull, 'w') as dn:
    eng.make_plan_batch(p, off, ctrl)
    L.tp_debug_phase_get(out)
    eng.make_plan_batch(p, off, ctrl)
    L.tp_debug_phase_get(out)
v = np.array(list(out), float)
it = v[5]
print("iterations %.0f evals %.0f | cycles/iter: total %.0f eval %.0f gram %.0f coeffs %.0f direction %.0f other %.0f" % (
    it, v[6], v[0]/it, v[1]/it, v[2]/it, v[3]/it, v[4]/it, (v[0]-v[1]-v[2]-v[3]-v[4])/it))
