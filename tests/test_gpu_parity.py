"""GPU parity: every CUDA entry point of the C ABI against the CPU oracle on identical seeded inputs.

Tolerances are the north-star's: voxel indices / occupancy / collision decisions bit-exact; FP64
cost and gradient per evaluation <= 1e-10 relative; optimised control points <= 1e-6 m at equal
iteration counts.  `strict_order` runs additionally have to reproduce the CPU iterate to rounding
(the only operations not shared bit-for-bit with glibc are pow(x,3) and atan2).
"""
import numpy as np
import pytest

from helpers import flat_guides, rel_err, traj

pytestmark = pytest.mark.gpu


def test_query_points_lines_unknown_bit_exact(engine, sq_omap):
    rng = np.random.default_rng(7)
    n = 200000
    xyz = np.column_stack([rng.uniform(-21, 21, n), rng.uniform(-21, 21, n), rng.uniform(-0.5, 3.5, n)])
    # exact voxel-face coordinates and far-away / non-finite points
    xyz[:1000] = np.round(xyz[:1000], 1)
    xyz[1000] = [1e300, 0, 1]
    xyz[1001] = [np.nan, 0, 1]
    xyz[1002] = [-20.0, -20.0, -0.1]
    xyz[1003] = [20.0, 20.0, 2.9]
    assert np.array_equal(engine.query_points(xyz), sq_omap.query(xyz))
    assert np.array_equal(engine.query_unknown(xyz), sq_omap.query_unknown(xyz))
    a = xyz[:50000]
    b = a + rng.normal(0, 0.4, a.shape)
    assert np.array_equal(engine.query_lines(a, b), sq_omap.query_lines(a, b))
    assert engine.query_points(np.zeros((0, 3))).size == 0


def _oracle_guides(orc, sq_omap, problems, B):
    """Oracle makePlan steps 1-3 per trajectory -> flat guide list + per-trajectory planners."""
    per, planners, oks = [], [], []
    for b in range(B):
        pl = orc.Planner(sq_omap)
        pl.set_ctrl(traj(problems, b))
        oks.append(pl.init_guides())
        per.append(pl.get_guides())
        planners.append(pl)
    return per, planners, oks


@pytest.mark.parametrize("strict", [0, 1])
def test_cost_and_gradient_parity(tp, engine, orc, sq_omap, problems, strict):
    B = 48
    off = problems["offsets"][:B + 1]
    per, planners, _ = _oracle_guides(orc, sq_omap, problems, B)
    rng = np.random.default_rng(3)
    ctrl = problems["ctrl"][:off[B]].copy()
    ctrl += rng.normal(0, 0.15, ctrl.shape)  # move off the seed so every cost branch is exercised
    weights = np.column_stack([rng.choice([1.0, 2.0, 8.0], B), np.ones(B)])
    p = tp.default_params()
    p.strict_order = strict
    f, g = engine.cost_batch(p, off, ctrl, flat_guides(per), weights)
    worst_f, worst_g = 0.0, 0.0
    for b in range(B):
        c = ctrl[off[b]:off[b + 1]]
        planners[b].set_weights(weights[b, 0], weights[b, 1])
        fo, go, _ = planners[b].cost(c[3:-3].ravel())
        gs = g[3 * (off[b] - 6 * b): 3 * (off[b + 1] - 6 * (b + 1))]
        worst_f = max(worst_f, abs(f[b] - fo) / max(abs(fo), 1e-300))
        worst_g = max(worst_g, np.max(np.abs(gs - go)) / max(np.max(np.abs(go)), 1e-300))
    print(f"strict={strict} worst rel err: cost {worst_f:.3e} grad {worst_g:.3e}")
    assert worst_f <= 1e-10 and worst_g <= 1e-10
    if strict:
        assert worst_f <= 1e-14 and worst_g <= 1e-14


def test_cost_plan_in_z_and_dynamic_obstacles(tp, engine, orc, sq_omap, problems):
    """Height barrier (x-row quirk), uncertain-aware factor and the dynamic-obstacle term."""
    B = 8
    off = problems["offsets"][:B + 1]
    p = tp.default_params()
    p.plan_in_z = 1
    p.uncertain_factor = 2.0
    po = sq_omap.lib.default_params()
    po.plan_in_z = 1
    po.uncertain_factor = 2.0
    rng = np.random.default_rng(5)
    ctrl = problems["ctrl"][:off[B]].copy()
    ctrl += rng.normal(0, 0.2, ctrl.shape)
    per = []
    pls = []
    for b in range(B):
        pl = orc.Planner(sq_omap, po)
        pl.set_ctrl(traj(problems, b))
        pl.init_guides()
        per.append(pl.get_guides())
        pls.append(pl)
    f, g = engine.cost_batch(p, off, ctrl, flat_guides(per))
    for b in range(B):
        fo, go, _ = pls[b].cost(ctrl[off[b]:off[b + 1]][3:-3].ravel())
        gs = g[3 * (off[b] - 6 * b): 3 * (off[b + 1] - 6 * (b + 1))]
        assert abs(f[b] - fo) <= 1e-10 * abs(fo)
        assert np.max(np.abs(gs - go)) <= 1e-10 * np.max(np.abs(go))


@pytest.mark.parametrize("strict", [0, 1])
def test_optimize_parity(tp, engine, orc, sq_omap, problems, strict):
    """One optimize() (fused cost + L-BFGS kernel) vs lbfgs_optimize on the CPU."""
    B = 48
    off = problems["offsets"][:B + 1]
    per, planners, _ = _oracle_guides(orc, sq_omap, problems, B)
    p = tp.default_params()
    p.strict_order = strict
    ctrl_out, res, xf = engine.optimize_batch(p, off, problems["ctrl"][:off[B]], flat_guides(per))
    same_iters, within, worst_same, exact = 0, 0, 0.0, 0
    for b in range(B):
        o = planners[b].optimize()
        co = planners[b].get_ctrl()
        cg = ctrl_out[off[b]:off[b + 1]]
        d = float(np.max(np.abs(cg - co)))
        if res["iters"][b] == o["iters"] and res["evals"][b] == o["evals"] and res["ret"][b] == o["ret"]:
            same_iters += 1
            worst_same = max(worst_same, d)
            within += d <= 1e-6
            exact += d == 0.0
    print(f"strict={strict}: equal (ret,iters,evals) {same_iters}/{B}, of those within 1e-6 m: {within}, "
          f"bit-identical: {exact}, worst {worst_same:.3e}")
    assert same_iters >= int(0.75 * B)
    assert within == same_iters
    if strict:
        assert same_iters == B and worst_same <= 1e-9


def test_has_collision_and_segments_bit_exact(tp, engine, orc, sq_omap, problems):
    B = len(problems["offsets"]) - 1
    off = problems["offsets"]
    p = tp.default_params()
    hit = engine.has_collision_batch(p, off, problems["ctrl"])
    segs = engine.find_collision_seg_batch(p, off, problems["ctrl"])
    nhit = 0
    for b in range(B):
        pl = orc.Planner(sq_omap)
        pl.set_ctrl(traj(problems, b))
        assert bool(hit[b]) == pl.has_collision()
        assert np.array_equal(segs[b], pl.find_collision_seg()), b
        nhit += int(hit[b])
    assert 0 < nhit  # the workload does cross obstacles


def test_astar_paths_bit_exact(tp, engine, orc, sq_omap, problems):
    B = len(problems["offsets"]) - 1
    p = tp.default_params()
    starts, ends = [], []
    for b in range(B):
        pl = orc.Planner(sq_omap)
        c = traj(problems, b)
        pl.set_ctrl(c)
        for s0, s1 in pl.find_collision_seg():
            starts.append(c[s0])
            ends.append(c[s1])
    # plus pairs that fail (goal buried in an obstacle / out of the pool)
    starts.append([0.0, 0.0, 1.0]); ends.append([30.0, 0.0, 1.0])
    starts.append([-8.0, -8.0, 1.0]); ends.append([-7.0, -7.6, 1.0])
    paths, ex = engine.astar_batch(p, starts, ends)
    pl = orc.Planner(sq_omap)
    nfound = 0
    for s in range(len(starts)):
        po, eo = pl.astar(starts[s], ends[s])
        assert (po is None) == (paths[s] is None), s
        assert ex[s] == eo, (s, ex[s], eo)
        if po is not None:
            nfound += 1
            assert np.array_equal(po, paths[s]), s
    assert nfound > 10


def test_init_guides_parity(tp, engine, orc, sq_omap, problems):
    """findCollisionSeg + pathSearch + assignGuidePointsSemiCircle (makePlan steps 1-3)."""
    B = len(problems["offsets"]) - 1
    p = tp.default_params()
    out = engine.init_guides_batch(p, problems["offsets"], problems["ctrl"])
    per, planners, oks = _oracle_guides(orc, sq_omap, problems, B)
    npairs = 0
    for b in range(B):
        assert out[b]["ok"] == oks[b], b
        if not oks[b]:
            continue
        assert np.array_equal(out[b]["segs"], planners[b].get_segs()), b
        cp, gp, gv = per[b]
        assert np.array_equal(out[b]["cp"], cp), b
        npairs += len(cp)
        if len(cp):
            assert np.max(np.abs(out[b]["p"] - gp)) <= 1e-9, b   # atan2 is not glibc's
            assert np.max(np.abs(out[b]["v"] - gv)) <= 1e-9, b
    assert npairs > 50


@pytest.mark.parametrize("strict", [0, 1])
def test_make_plan_batch_parity(tp, engine, orc, sq_omap, problems, strict):
    """The batched entry point vs the oracle's makePlan on the same control points."""
    off = problems["offsets"]
    B = len(off) - 1
    p = tp.default_params()
    p.strict_order = strict
    out, res = engine.make_plan_batch(p, off, problems["ctrl"])
    po = sq_omap.lib.default_params()
    ok_o, out_o, st_o = orc.make_plan_batch(sq_omap, po, off, problems["ctrl"], nthreads=4)
    status_o = np.where(st_o["success"] == 1, 1, 0)
    status_g = np.where(res["status"] == 1, 1, 0)
    same_flow = 0
    within = 0
    worst = 0.0
    lf_worst = 0.0
    for b in range(B):
        keys = ["outer_rounds", "fail_count", "lbfgs_runs", "lbfgs_iters", "lbfgs_evals", "astar_searches",
                "astar_expansions", "n_guide_pairs"]
        if status_o[b] == status_g[b] and all(res[k][b] == st_o[k][b] for k in keys):
            same_flow += 1
            d = float(np.max(np.abs(out[off[b]:off[b + 1]] - out_o[off[b]:off[b + 1]])))
            worst = max(worst, d)
            within += d <= 1e-6
            if status_g[b] == 1:
                lf_worst = max(lf_worst, abs(res["linear_factor"][b] - st_o["linear_factor"][b]))
    agree = int(np.sum(status_o == status_g))
    print(f"strict={strict}: status agree {agree}/{B}; identical control flow {same_flow}/{B}; of those within "
          f"1e-6 m: {within} (worst {worst:.3e}); linear factor worst {lf_worst:.3e}; "
          f"success rate gpu {status_g.mean():.3f} oracle {status_o.mean():.3f}")
    assert agree >= int(0.9 * B)
    assert same_flow >= int(0.6 * B)
    assert within == same_flow
    assert lf_worst <= 1e-6
    if strict:
        assert same_flow >= int(0.95 * B)


def test_bspline_traj_class_drop_in(tp, engine, orc, sq_omap, problems):
    """The reference-shaped planner object: updatePath / makePlan / getPose / isCurrTrajValid."""
    bt = tp.BsplineTraj(engine)
    ok = bt.updatePathFromStartGoal(problems["starts"][0], problems["goals"][0])
    assert ok
    c0 = bt.getControlPoints()
    assert c0.shape[0] == 3
    pl = orc.Planner(sq_omap)
    pl.set_ctrl(c0.T)
    ok_o, st = pl.make_plan()
    assert bt.makePlan() == ok_o
    if ok_o:
        assert bt.isCurrTrajValid()
        assert abs(bt.getLinearFactor() - st["linear_factor"]) < 1e-6
        x, y, z, yaw = bt.getPose(0.5 * bt.getDuration())
        ref = orc.bspline_at(bt.getControlPoints().T, [0.5 * bt.getDuration()])[0]
        assert np.allclose([x, y, z], ref, atol=1e-12)
