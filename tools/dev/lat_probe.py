# single-solve latency breakdown: wall vs kernel time vs launches, a few problems of the bench batch
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import trajectory_planner_b200 as tp
import bench
pmap = tp.OccMap.from_tpm(bench.MAP_TPM)
eng = tp.Engine(0)
eng.set_map(pmap)
p = tp.default_params()
offsets, ctrl = bench.committed_workload(tp, pmap, p, 4096)
Ns = np.diff(offsets)
for b in (0, 1, 2, 3, 5, 8):
    o1 = np.array([0, Ns[b]], np.int32)
    c1 = ctrl[offsets[b]:offsets[b + 1]]
    for _ in range(5):
        eng.make_plan_batch(p, o1, c1)
    eng.profile_enable(True); eng.profile_get()
    l0 = eng.launch_count
    lat = []
    for _ in range(20):
        t0 = time.perf_counter(); out, res = eng.make_plan_batch(p, o1, c1); lat.append(1e3 * (time.perf_counter() - t0))
    pr = eng.profile_get(); eng.profile_enable(False)
    print(f"traj {b} N {Ns[b]} iters {res['lbfgs_iters'][0]} exp {res['astar_expansions'][0]} rounds {res['outer_rounds'][0]}: wall p50 {np.median(lat):.3f} ms, "
          f"kernel solve {pr['ms']['solve']/20:.3f} ms, all kernels {sum(pr['ms'].values())/20:.3f} ms, launches/call {(eng.launch_count-l0)/20:.1f}, "
          f"us/iter {1e3*pr['ms']['solve']/20/max(res['lbfgs_iters'][0],1):.2f}")
