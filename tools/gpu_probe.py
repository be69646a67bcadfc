#!/usr/bin/env python
"""GPU probe (development aid): determinism of the fast mode, per-kernel time, A* cost per expansion."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import trajectory_planner_b200 as tp
sys.argv = sys.argv[:1] + [a for a in sys.argv[1:]]
import bench

B = int(os.environ.get("PROBE_B", "4096"))
pmap = tp.OccMap.from_tpm(bench.MAP_TPM)
eng = tp.Engine(0); eng.set_map(pmap)
p = tp.default_params()
off, ctrl = bench.make_workload(tp, pmap, eng.query_points, B, bench.SEED, p)
for strict in (0, 1):
    p.strict_order = strict
    outs = []
    for rep in range(3):
        eng.profile_enable(True); eng.profile_get()
        t0 = time.perf_counter()
        out, res = eng.make_plan_batch(p, off, ctrl)
        dt = time.perf_counter() - t0
        prof = eng.profile_get()
        outs.append((out.copy(), res.copy()))
        print(f"strict={strict} rep={rep} {B/dt:.0f} solves/s wall; ms {({k: round(v,2) for k,v in prof['ms'].items()})} "
              f"launches {prof['launches']['lbfgs']} success {np.mean(res['status']==1):.4f} "
              f"iters {res['lbfgs_iters'].mean():.1f} exp mean {res['astar_expansions'].mean():.1f} max {res['astar_expansions'].max()} "
              f"rounds max {res['outer_rounds'].max()}")
    same = all(np.array_equal(outs[0][0], o[0]) and np.array_equal(outs[0][1], o[1]) for o in outs[1:])
    print(f"strict={strict}: run-to-run bit-identical: {same}")
# A* cost per expansion: searches taken from the collision segments of the batch
p.strict_order = 0
segs = eng.find_collision_seg_batch(p, off, ctrl)
starts, ends = [], []
for b in range(B):
    c = ctrl[off[b]:off[b + 1]]
    for s0, s1 in segs[b]:
        starts.append(c[s0]); ends.append(c[s1])
starts, ends = np.array(starts), np.array(ends)
print("A* searches", len(starts))
for S in (1, 32, 1024, len(starts)):
    eng.astar_batch(p, starts[:S], ends[:S])
    eng.profile_enable(False)
    t0 = time.perf_counter()
    paths, ex = eng.astar_batch(p, starts[:S], ends[:S])
    dt = time.perf_counter() - t0
    print(f"A* S={S}: {dt*1e3:.2f} ms wall, expansions total {ex.sum()} max {ex.max()}, {dt*1e9/max(ex.sum(),1):.1f} ns/exp aggregate, "
          f"{dt*1e9/max(ex.max(),1):.1f} ns per expansion of the longest")
t0 = time.perf_counter(); g = eng.init_guides_batch(p, off, ctrl); dt = time.perf_counter() - t0
print(f"init_guides_batch wall {dt*1e3:.1f} ms")
