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
        pl = orc.Planner(sq_omap)
        pl.set_ctrl(c)          # perturbed end points too
        pl.add_guides(*per[b])
        pl.set_weights(weights[b, 0], weights[b, 1])
        fo, go, _ = pl.cost(c[3:-3].ravel())
        gs = g[3 * (off[b] - 6 * b): 3 * (off[b + 1] - 6 * (b + 1))]
        worst_f = max(worst_f, abs(f[b] - fo) / max(abs(fo), 1e-300))
        worst_g = max(worst_g, np.max(np.abs(gs - go)) / max(np.max(np.abs(go)), 1e-300))
        if not strict:
            # the default (warp-form) evaluation is pinned BIT FOR BIT to its CPU restatement (oracle/wform_port.hpp)
            fw, gw = pl.cost_wform(c[3:-3].ravel())
            assert f[b] == fw and np.array_equal(gs, gw), b
    print(f"strict={strict} worst rel err: cost {worst_f:.3e} grad {worst_g:.3e}")
    assert worst_f <= 1e-10 and worst_g <= 1e-10
    if strict:
        assert worst_f <= 1e-14 and worst_g <= 1e-14


@pytest.mark.parametrize("strict", [0, 1])
def test_cost_plan_in_z_and_dynamic_obstacles(tp, engine, orc, sq_omap, problems, strict):
    """Height barrier (x-row quirk), uncertain-aware factor and getDynamicObstacleCost (bsplineTraj.cpp:1001-1064) per
    evaluation, in both reduction orders, with per-trajectory weightDistance_ / weightDynamicObstacle_ overrides:
    <= 1e-10 relative vs the reference-order oracle (strict: <= 1e-14), and the default mode bit-identical to its own
    restatement.  Then one optimize() with the obstacles: bit-identical in both modes."""
    B = 8
    off = problems["offsets"][:B + 1]
    p = tp.default_params()
    p.plan_in_z = 1
    p.uncertain_factor = 2.0
    p.strict_order = strict
    po = sq_omap.lib.default_params()
    po.plan_in_z = 1
    po.uncertain_factor = 2.0
    po.fast_order = 0 if strict else 4
    rng = np.random.default_rng(5)
    ctrl = problems["ctrl"][:off[B]].copy()
    ctrl += rng.normal(0, 0.2, ctrl.shape)
    per = []
    for b in range(B):
        pl = orc.Planner(sq_omap, po)
        pl.set_ctrl(traj(problems, b))
        pl.init_guides()
        per.append(pl.get_guides())
    # obstacles next to the first trajectories' mid points (inside their distance threshold) and one far away
    mids = np.array([traj(problems, b)[len(traj(problems, b)) // 2] for b in range(3)])
    dyn = (np.vstack([mids + [0.3, -0.2, 0.0], [[30.0, 30.0, 1.0]]]),
           np.array([[0.5, 0.2, 0.0], [-0.3, 0.4, 0.0], [0.0, 0.0, 0.0], [0.1, 0.1, 0.0]]),
           np.array([[0.6, 0.8, 1.7], [0.5, 0.5, 1.7], [1.0, 0.4, 1.7], [0.5, 0.5, 1.7]]))
    weights = np.column_stack([rng.choice([1.0, 4.0], B), rng.choice([1.0, 2.0, 8.0], B)])
    f, g = engine.cost_batch(p, off, ctrl, flat_guides(per), weights, dyn=dyn)
    dyn_active = 0
    for b in range(B):
        c = ctrl[off[b]:off[b + 1]]
        pl = orc.Planner(sq_omap, po)
        pl.set_ctrl(c)
        pl.set_dyn(*dyn)
        pl.add_guides(*per[b])
        pl.set_weights(weights[b, 0], weights[b, 1])
        fo, go, terms = pl.cost(c[3:-3].ravel())
        dyn_active += terms[3] > 0
        gs = g[3 * (off[b] - 6 * b): 3 * (off[b + 1] - 6 * (b + 1))]
        tol = 1e-14 if strict else 1e-10
        assert abs(f[b] - fo) <= tol * abs(fo), b
        assert np.max(np.abs(gs - go)) <= tol * np.max(np.abs(go)), b
        if not strict:
            fw, gw = pl.cost_wform(c[3:-3].ravel())
            assert f[b] == fw and np.array_equal(gs, gw), b
    assert dyn_active >= 3   # the dynamic-obstacle term is really exercised
    # one optimize() with the obstacles present
    co, res, xf = engine.optimize_batch(p, off, ctrl, flat_guides(per), weights, dyn=dyn)
    for b in range(B):
        pl = orc.Planner(sq_omap, po)
        pl.set_ctrl(ctrl[off[b]:off[b + 1]])
        pl.set_dyn(*dyn)
        pl.add_guides(*per[b])
        pl.set_weights(weights[b, 0], weights[b, 1])
        o = pl.optimize()
        if strict:
            # the reference-order oracle computes the obstacle radius with pow(x, 0.5); the device with sqrt: identical
            # flow is the rule, not a guarantee
            if (res["iters"][b], res["evals"][b], res["ret"][b]) != (o["iters"], o["evals"], o["ret"]):
                continue
            assert np.max(np.abs(pl.get_ctrl() - co[off[b]:off[b + 1]])) <= 1e-9, b
        else:
            assert (res["iters"][b], res["evals"][b], res["ret"][b], res["fx"][b]) == (o["iters"], o["evals"], o["ret"], o["fx"]), b
            assert np.array_equal(pl.get_ctrl(), co[off[b]:off[b + 1]]), b


@pytest.mark.parametrize("strict", [0, 1])
def test_optimize_parity(tp, engine, orc, sq_omap, problems, strict):
    """One optimize() (fused cost + L-BFGS kernel) vs the CPU, identical inputs, BIT FOR BIT in both modes.

    strict_order=1 reproduces lbfgs_optimize in the reference's serial summation order.  The default mode (warp form,
    the benchmarked kernel) reproduces oracle/wform_port.hpp — the CPU restatement of its own arithmetic (gather-form
    cost with FMAs, Gram-form direction, butterfly / reduce-scatter sums): same return code, iteration and evaluation
    counts, final cost and control points, to the last bit.  How far that arithmetic is from the reference-order
    iterate is quantified on the CPU (tests/test_oracle_cpu.py::test_fast_order_vs_reference_order_study)."""
    B = len(problems["offsets"]) - 1
    off = problems["offsets"]
    po = sq_omap.lib.default_params()
    po.fast_order = 0 if strict else 4
    per, planners = [], []
    for b in range(B):
        pl = orc.Planner(sq_omap, po)
        pl.set_ctrl(traj(problems, b))
        pl.init_guides()
        per.append(pl.get_guides())
        planners.append(pl)
    p = tp.default_params()
    p.strict_order = strict
    ctrl_out, res, xf = engine.optimize_batch(p, off, problems["ctrl"], flat_guides(per))
    same, exact, worst = 0, 0, 0.0
    iters = 0
    for b in range(B):
        o = planners[b].optimize()
        co = planners[b].get_ctrl()
        cg = ctrl_out[off[b]:off[b + 1]]
        d = float(np.max(np.abs(cg - co)))
        worst = max(worst, d)
        iters += o["iters"]
        if res["iters"][b] == o["iters"] and res["evals"][b] == o["evals"] and res["ret"][b] == o["ret"] and res["fx"][b] == o["fx"]:
            same += 1
            exact += d == 0.0
        xs = xf[3 * (off[b] - 6 * b): 3 * (off[b + 1] - 6 * (b + 1))].reshape(-1, 3)
        assert np.array_equal(xs, o["x"]), b          # the solver's own x (reverted on a failed line search)
        if res["ret"][b] >= 0 or res["ret"][b] == -1004:
            assert np.array_equal(xs, cg[3:-3])
    print(f"strict={strict}: equal (ret,iters,evals,fx) {same}/{B}; bit-identical control points {exact}/{B}; worst |diff| "
          f"{worst:.3e}; {iters} L-BFGS iterations compared")
    assert same == B and exact == B


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
    # after one optimize() most trajectories are collision free: the decision must still agree
    per, planners, _ = _oracle_guides(orc, sq_omap, problems, B)
    moved = np.concatenate([(planners[b].optimize(), planners[b].get_ctrl())[1] for b in range(B)], 0)
    hit2 = engine.has_collision_batch(p, off, moved)
    segs2 = engine.find_collision_seg_batch(p, off, moved)
    for b in range(B):
        assert bool(hit2[b]) == planners[b].has_collision(), b
        assert np.array_equal(segs2[b], planners[b].find_collision_seg()), b


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
    # plus pairs that fail (goal buried in an obstacle / out of the pool) and a trivial one
    starts.append([0.0, 0.0, 1.0]); ends.append([30.0, 0.0, 1.0])
    starts.append([-8.0, -8.0, 1.0]); ends.append([-7.0, -7.6, 1.0])
    starts.append([0.0, 0.0, 1.0]); ends.append([0.02, 0.01, 1.0])
    starts.append([0.0, 0.0, 2.5]); ends.append([1.0, 0.5, 1.0])   # start above the height band
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


@pytest.mark.parametrize("soft", [0, 1])
def test_init_guides_parity(tp, engine, orc, sq_omap, problems, soft):
    """findCollisionSeg + pathSearch + shortcut + assignGuidePointsSemiCircle (makePlan steps 1-3).
    soft=0: oracle with std::atan2 (reference-faithful) -> guide points agree to 1e-9;
    soft=1: oracle with the same deterministic atan2 the device uses -> bit-identical."""
    B = len(problems["offsets"]) - 1
    p = tp.default_params()
    out = engine.init_guides_batch(p, problems["offsets"], problems["ctrl"])
    po = sq_omap.lib.default_params()
    po.soft_atan2 = soft
    npairs = 0
    for b in range(B):
        pl = orc.Planner(sq_omap, po)
        pl.set_ctrl(traj(problems, b))
        ok = pl.init_guides()
        assert out[b]["ok"] == ok, b
        if not ok:
            continue
        assert np.array_equal(out[b]["segs"], pl.get_segs()), b
        cp, gp, gv = pl.get_guides()
        assert np.array_equal(out[b]["cp"], cp), b
        npairs += len(cp)
        if len(cp):
            if soft:
                assert np.array_equal(out[b]["p"], gp) and np.array_equal(out[b]["v"], gv), b
            else:
                assert np.max(np.abs(out[b]["p"] - gp)) <= 1e-9, b
                assert np.max(np.abs(out[b]["v"] - gv)) <= 1e-9, b
    assert npairs > 50


def _plan_compare(off, out, res, out_o, st_o):
    B = len(off) - 1
    keys = ["outer_rounds", "fail_count", "lbfgs_runs", "lbfgs_iters", "lbfgs_evals", "astar_searches",
            "astar_expansions", "n_guide_pairs", "last_lbfgs_ret"]
    status_o = np.where(st_o["success"] == 1, 1, 0)
    status_g = np.where(res["status"] == 1, 1, 0)
    r = dict(agree=int(np.sum(status_o == status_g)), same_flow=0, within=0, exact=0, worst=0.0, lf_worst=0.0,
             succ_g=float(status_g.mean()), succ_o=float(status_o.mean()))
    for b in range(B):
        if status_o[b] == status_g[b] and all(res[k][b] == st_o[k][b] for k in keys):
            r["same_flow"] += 1
            d = float(np.max(np.abs(out[off[b]:off[b + 1]] - out_o[off[b]:off[b + 1]])))
            r["worst"] = max(r["worst"], d)
            r["within"] += d <= 1e-6
            r["exact"] += d == 0.0
            if status_g[b] == 1:
                r["lf_worst"] = max(r["lf_worst"], abs(res["linear_factor"][b] - st_o["linear_factor"][b]))
    return r


def test_make_plan_batch_strict_is_bit_faithful(tp, engine, orc, sq_omap, problems):
    """The batched entry point in strict order vs the oracle's makePlan with the shared deterministic
    atan2: whole-solve control flow (rounds, fail counts, A* expansions, L-BFGS iterations) and final
    control points must be identical."""
    off = problems["offsets"]
    B = len(off) - 1
    p = tp.default_params()
    p.strict_order = 1
    out, res = engine.make_plan_batch(p, off, problems["ctrl"])
    po = sq_omap.lib.default_params()
    po.soft_atan2 = 1
    _, out_o, st_o = orc.make_plan_batch(sq_omap, po, off, problems["ctrl"], nthreads=4)
    r = _plan_compare(off, out, res, out_o, st_o)
    print("strict vs oracle(soft atan2):", r)
    assert r["agree"] == B and r["same_flow"] == B
    assert r["within"] == B and r["worst"] <= 1e-9 and r["lf_worst"] <= 1e-9
    assert (res["status"] == 1).sum() > 0.8 * B


def test_make_plan_batch_default_mode_is_bit_faithful_to_its_oracle(tp, engine, orc, sq_omap, problems):
    """The batched entry point in its DEFAULT mode (warp-form kernel, what bench.py measures) vs the oracle's makePlan
    running the restated warp-form arithmetic (fast_order=1) and the shared deterministic atan2: identical control
    flow (rounds, fail counts, A* expansions, L-BFGS iterations / evaluations) and bit-identical control points."""
    off = problems["offsets"]
    B = len(off) - 1
    p = tp.default_params()
    assert p.strict_order == 0
    out, res = engine.make_plan_batch(p, off, problems["ctrl"])
    po = sq_omap.lib.default_params()
    po.soft_atan2 = 1
    po.fast_order = 4
    _, out_o, st_o = orc.make_plan_batch(sq_omap, po, off, problems["ctrl"], nthreads=4)
    r = _plan_compare(off, out, res, out_o, st_o)
    print("default mode vs oracle(fast_order, soft atan2):", r)
    assert r["agree"] == B and r["same_flow"] == B and r["exact"] == B
    assert np.array_equal(res["final_cost"], st_o["final_cost"])
    assert r["lf_worst"] <= 1e-12
    assert (res["status"] == 1).sum() > 0.8 * B


@pytest.mark.parametrize("strict", [0, 1])
def test_make_plan_batch_vs_reference_faithful_oracle(tp, engine, orc, sq_omap, problems, strict):
    """Same, against the oracle with std::atan2 (the reference's libm).  ULP-level differences in guide
    points (and, for strict=0, in reduction order) are amplified by the solve, so identical control flow
    holds only for a fraction; the solve QUALITY must be statistically the same."""
    off = problems["offsets"]
    B = len(off) - 1
    p = tp.default_params()
    p.strict_order = strict
    out, res = engine.make_plan_batch(p, off, problems["ctrl"])
    po = sq_omap.lib.default_params()
    _, out_o, st_o = orc.make_plan_batch(sq_omap, po, off, problems["ctrl"], nthreads=4)
    r = _plan_compare(off, out, res, out_o, st_o)
    print(f"strict={strict} vs oracle(std::atan2):", r)
    assert r["agree"] >= int(0.9 * B)
    assert abs(r["succ_g"] - r["succ_o"]) <= 0.06
    assert r["same_flow"] >= int((0.5 if strict else 0.15) * B)
    # every successful GPU trajectory is collision free under the oracle's own check
    for b in range(B):
        if res["status"][b] == 1:
            pl = orc.Planner(sq_omap)
            pl.set_ctrl(out[off[b]:off[b + 1]])
            assert not pl.has_collision(), b


def test_make_plan_with_dynamic_obstacles(tp, engine, orc, sq_omap, problems):
    off = problems["offsets"][:17]
    B = 16
    ctrl = problems["ctrl"][:off[B]]
    mid = np.array([traj(problems, b)[len(traj(problems, b)) // 2] for b in range(2)])
    dyn = (mid + [0.3, 0.3, 0.0], np.array([[0.2, 0.1, 0.0], [-0.1, 0.2, 0.0]]), np.array([[0.5, 0.5, 1.7], [0.4, 0.6, 1.7]]))
    p = tp.default_params()
    p.strict_order = 1
    out, res = engine.make_plan_batch(p, off, ctrl, dyn)
    po = sq_omap.lib.default_params()
    po.soft_atan2 = 1
    same = 0
    for b in range(B):
        pl = orc.Planner(sq_omap, po)
        pl.set_ctrl(ctrl[off[b]:off[b + 1]])
        pl.set_dyn(*dyn)
        ok, st = pl.make_plan()
        if ok == (res["status"][b] == 1) and st["lbfgs_iters"] == res["lbfgs_iters"][b]:
            same += 1
            assert np.max(np.abs(pl.get_ctrl() - out[off[b]:off[b + 1]])) <= 1e-9, b
    print("dynamic obstacles: identical flow", same, "/", B)
    assert same >= B - 2   # pow(x,0.5) vs sqrt and the obstacle-size sqrt are the only unshared ops


def test_bspline_traj_class_drop_in(tp, engine, orc, sq_omap, problems):
    """The reference-shaped planner object: updatePath / makePlan / getPose / isCurrTrajValid."""
    p = tp.default_params()
    p.strict_order = 1
    bt = tp.BsplineTraj(engine, p)
    ok = bt.updatePathFromStartGoal(problems["starts"][0], problems["goals"][0])
    assert ok
    c0 = bt.getControlPoints()
    assert c0.shape[0] == 3
    po = sq_omap.lib.default_params()
    po.soft_atan2 = 1
    pl = orc.Planner(sq_omap, po)
    pl.set_ctrl(c0.T)
    ok_o, st = pl.make_plan()
    assert bt.makePlan() == ok_o
    assert np.array_equal(bt.getControlPoints().T, pl.get_ctrl())
    if ok_o:
        assert bt.isCurrTrajValid()
        assert abs(bt.getLinearFactor() - st["linear_factor"]) < 1e-12
        x, y, z, yaw = bt.getPose(0.5 * bt.getDuration())
        ref = orc.bspline_at(bt.getControlPoints().T, [0.5 * bt.getDuration()])[0]
        assert np.array_equal([x, y, z], ref)
        # a failed replan (here: too few control points) leaves the committed trajectory, its duration and its factor in
        # place (bspline_ is only replaced by a successful makePlan, bsplineTraj.cpp:376-377)
        dur, lf, committed = bt.getDuration(), bt.getLinearFactor(), bt.getTrajectoryControlPoints()
        bt.updateControlPoints(bt.getControlPoints().T[:5])
        assert not bt.makePlan()
        assert bt.getDuration() == dur and bt.getLinearFactor() == lf and np.array_equal(bt.getTrajectoryControlPoints(), committed)
        assert np.array_equal(bt.getPose(0.5 * dur)[:3], [x, y, z])


def test_batched_pose_sampling_matches_host_and_oracle(tp, engine, orc, problems):
    """tp_vigo_sample_batch (bsplineTraj::getPose / evalTraj / evalTrajToMsg for a whole batch): position, velocity and
    acceleration bit-identical to the host de Boor (tp_bspline_eval) and to the oracle's bspline::at; yaw within 4 ulp
    of libm's atan2; ragged and empty sample lists, out-of-range times (clamped as bspline.cpp:33-35)."""
    off, ctrl = problems["offsets"], problems["ctrl"]
    B = 24
    ts = 0.2
    rng = np.random.default_rng(7)
    times = []
    for b in range(B):
        dur = (off[b + 1] - off[b] - 3) * ts
        k = int(rng.integers(0, 40)) if b % 5 else 0   # some trajectories have no samples at all
        tt = rng.uniform(-0.5, dur + 0.5, k)            # some before 0 / after the end
        if k >= 3:
            tt[0], tt[1], tt[2] = 0.0, dur, 3 * ts     # end points and a knot
        times.append(tt)
    out = engine.sample_batch(off[:B + 1], ctrl[:off[B]], times=times, ts=ts, vel=True, acc=True, yaw=True)
    n = 0
    for b in range(B):
        c = ctrl[off[b]:off[b + 1]]
        sl = slice(out["t_offsets"][b], out["t_offsets"][b + 1])
        if len(times[b]) == 0:
            assert sl.start == sl.stop
            continue
        assert np.array_equal(out["pos"][sl], tp.bspline_eval(c, times[b], ts, 0))
        assert np.array_equal(out["vel"][sl], tp.bspline_eval(c, times[b], ts, 1))
        assert np.array_equal(out["acc"][sl], tp.bspline_eval(c, times[b], ts, 2))
        assert np.array_equal(out["pos"][sl], orc.bspline_at(c, times[b], 3, ts))
        assert np.array_equal(out["vel"][sl], orc.bspline_deriv_at(c, times[b], 1, ts))
        v = out["vel"][sl]
        ref = np.arctan2(v[:, 1], v[:, 0])
        assert np.all(np.abs(out["yaw"][sl] - ref) <= 4 * np.spacing(np.abs(ref)) + 1e-300)
        n += len(times[b])
    # evalTraj's accumulated grid for every trajectory of the batch, positions only
    g = engine.sample_batch(off, ctrl, dt=0.05, ts=ts, vel=False, yaw=False)
    b = len(off) - 2
    sl = slice(g["t_offsets"][b], g["t_offsets"][b + 1])
    assert np.array_equal(g["pos"][sl], tp.bspline_eval(ctrl[off[b]:off[b + 1]], g["t"][sl], ts, 0))
    print(f"batched sampling: {n} ragged samples + {len(g['t'])} grid samples bit-identical to the host de Boor")


# ------------------------------------------------------------------------------------ committed golden vectors
# tests/golden/vigo_golden.npz was generated by tools/make_golden.py through oracle/_ref, i.e. with the
# REFERENCE'S OWN solver/lbfgs.hpp as the L-BFGS iterate.  The CUDA path is compared with the fixtures
# directly (no oracle call): this is the parity pin that travels to the GPU box.
import os

_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vigo_golden.npz")


def test_golden_queries_astar_segments(tp, engine):
    g = np.load(_GOLD)
    assert np.array_equal(engine.query_points(g["q_xyz"]), g["q_hit"])
    assert np.array_equal(engine.query_unknown(g["q_xyz"]), g["q_unknown"])
    assert np.array_equal(engine.query_lines(g["q_xyz"], g["q_b"]), g["q_line"])
    p = tp.default_params()
    paths, ex = engine.astar_batch(p, g["astar_starts"], g["astar_ends"])
    pos = 0
    for s in range(len(ex)):
        assert ex[s] == g["astar_exp"][s]
        n = g["astar_len"][s]
        assert (paths[s] is None) == (n < 0)
        if n >= 0:
            assert np.array_equal(paths[s], g["astar_paths"][pos:pos + n])
            pos += n
    off, ctrl = g["offsets"], g["ctrl"]
    assert np.array_equal(engine.has_collision_batch(p, off, ctrl), g["has_collision"])
    segs = engine.find_collision_seg_batch(p, off, ctrl)
    assert np.array_equal([len(s) for s in segs], g["nseg"])
    assert np.array_equal(np.concatenate([np.asarray(s).reshape(-1, 2) for s in segs], 0), g["segs"])


def test_golden_cost_optimize_makeplan_strict_bit_exact(tp, engine):
    """strict_order CUDA run vs the golden vectors produced with the reference's lbfgs.hpp and the shared
    deterministic atan2: cost/gradient to 1e-14 (pow(x,3) is the only op not shared bit-for-bit),
    optimize() and the whole makePlan bit-identical (control points, return codes, evaluation counts)."""
    g = np.load(_GOLD)
    off, ctrl, pert = g["offsets"], g["ctrl"], g["ctrl_perturbed"]
    B = len(off) - 1
    guides = (g["soft1_g_off"], g["soft1_g_cp"], g["soft1_g_p"], g["soft1_g_v"])
    p = tp.default_params()
    p.strict_order = 1
    f, grad = engine.cost_batch(p, off, pert, guides)
    assert np.max(np.abs(f - g["soft1_cost"]) / np.abs(g["soft1_cost"])) <= 1e-14
    assert np.max(np.abs(grad - g["soft1_grad"])) <= 1e-14 * np.max(np.abs(g["soft1_grad"]))
    co, res, _ = engine.optimize_batch(p, off, ctrl, guides)
    st = g["soft1_opt_stats"]
    assert np.array_equal(res["ret"], st[:, 0].astype(int)) and np.array_equal(res["evals"], st[:, 2].astype(int))
    assert np.array_equal(res["fx"], st[:, 3])
    assert np.array_equal(co, g["soft1_opt_ctrl"])
    out, r = engine.make_plan_batch(p, off, ctrl)
    gs = g["soft1_plan_stats"]
    assert np.array_equal((r["status"] == 1).astype(int), gs["success"])
    for k in ["outer_rounds", "fail_count", "lbfgs_runs", "lbfgs_evals", "astar_searches", "astar_expansions",
              "n_guide_pairs", "last_lbfgs_ret"]:
        assert np.array_equal(r[k], gs[k]), k
    assert np.array_equal(out, g["soft1_plan_ctrl"])
    ok = gs["success"] == 1
    assert np.max(np.abs(r["linear_factor"][ok] - gs["linear_factor"][ok])) <= 1e-12


def test_golden_fast_mode_per_evaluation_and_quality(tp, engine):
    """Default (tree-reduction) mode vs the golden vectors: per-evaluation cost/gradient <= 1e-10 relative
    (north-star tolerance) and the reference-faithful (libm atan2) makePlan outcome statistically equal."""
    g = np.load(_GOLD)
    off, ctrl, pert = g["offsets"], g["ctrl"], g["ctrl_perturbed"]
    guides = (g["soft0_g_off"], g["soft0_g_cp"], g["soft0_g_p"], g["soft0_g_v"])
    p = tp.default_params()
    f, grad = engine.cost_batch(p, off, pert, guides)
    assert np.max(np.abs(f - g["soft0_cost"]) / np.abs(g["soft0_cost"])) <= 1e-10
    assert np.max(np.abs(grad - g["soft0_grad"])) <= 1e-10 * np.max(np.abs(g["soft0_grad"]))
    out, r = engine.make_plan_batch(p, off, ctrl)
    gs = g["soft0_plan_stats"]
    assert np.mean((r["status"] == 1).astype(int) == gs["success"]) >= 0.9
    hit = engine.has_collision_batch(p, off, out)
    assert not np.any(hit[r["status"] == 1])


def test_astar_exhaustive_failure_and_long_detour(tp, orc):
    """A* on a synthetic map: (1) goal inside a sealed box -> the reference pops the start's whole connected
    component and fails; the CUDA path detects unreachability with a flood fill and must report the SAME
    failure and the SAME expansion count; (2) a long detour (> the flood-fill trigger) that IS reachable:
    the parked heap is restored and the path stays bit-identical."""
    res, origin, dims, inf = 0.1, (-8.0, -8.0, -0.1), (160, 160, 30), (4, 4, 2)
    m = tp.OccMap(res, origin, dims, inf)
    om = orc.Map(res, origin, dims, inf)
    pts = []
    zs = np.arange(0.05, 2.9, 0.1)
    def wall(x0, y0, x1, y1):
        n = int(max(abs(x1 - x0), abs(y1 - y0)) / 0.05) + 1
        for t in np.linspace(0, 1, n):
            for z in zs:
                pts.append([x0 + t * (x1 - x0), y0 + t * (y1 - y0), z])
    # sealed box around (2.5, 0): interior 3 m wide
    wall(1.0, -1.5, 4.0, -1.5); wall(4.0, -1.5, 4.0, 1.5); wall(4.0, 1.5, 1.0, 1.5); wall(1.0, 1.5, 1.0, -1.5)
    # a long wall with a gap far away -> long detour
    wall(-3.0, -4.5, -3.0, 3.0)
    pts = np.array(pts)
    m.add_points(pts)
    om.add_points(pts)
    e = tp.Engine(0)
    e.set_map(m)
    p = tp.default_params()
    starts = np.array([[-1.0, 0.0, 1.0], [-1.5, -1.0, 1.0], [0.0, 0.3, 1.0], [2.5, 0.2, 1.0]])
    ends = np.array([[2.5, 0.0, 1.0], [-4.5, -1.0, 1.0], [2.6, -0.4, 1.0], [2.2, -0.6, 1.0]])
    paths, ex = e.astar_batch(p, starts, ends)
    pl = orc.Planner(om)
    big = 0
    for s in range(len(starts)):
        po, eo = pl.astar(starts[s], ends[s])
        assert ex[s] == eo, (s, ex[s], eo)
        assert (po is None) == (paths[s] is None), s
        if po is not None:
            assert np.array_equal(po, paths[s]), s
        big += eo > 4096
    assert paths[0] is None and ex[0] > 5000      # exhaustive failure, count from the flood fill
    assert paths[1] is not None and ex[1] > 4096  # reachable long detour across the trigger
    assert big >= 3 and paths[2] is None and paths[3] is not None
    e.close()


def test_collision_sweep_box_map_bit_exact(tp, orc):
    """BASELINE configs[4] in small: 1 M samples against the box.bt raster, decisions bit-exact vs the oracle
    (uniform points over the map's bounding box plus a margin outside it, and trajectory-coherent samples)."""
    from helpers import oracle_map_from
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    m = tp.OccMap.from_tpm(os.path.join(root, "data", "maps", "box.tpm"))
    om = oracle_map_from(orc, m)
    e = tp.Engine(0)
    e.set_map(m)
    info = m.info()
    lo = np.array(info["origin"]) - 0.5
    hi = np.array(info["origin"]) + np.array(info["dims"]) * info["res"] + 0.5
    rng = np.random.default_rng(3)
    q = rng.uniform(lo, hi, (1 << 20, 3))
    p0 = rng.uniform(lo, hi, (256, 1, 3))
    d = rng.normal(size=(256, 1, 3)); d /= np.linalg.norm(d, axis=2, keepdims=True)
    coh = (p0 + d * (0.025 * np.arange(512))[None, :, None]).reshape(-1, 3)
    for pts in (q, coh):
        got = e.query_points(pts)
        assert np.array_equal(got, om.query(pts))
        assert 0.0 < got.mean() < 1.0
    e.close()


@pytest.mark.parametrize("name", ["maze", "tunnel"])
def test_make_plan_on_octomap_rasters_strict_bit_faithful(tp, orc, name):
    """BASELINE configs[2] in small: ViGO solves on maze.bt / tunnel.bt rasterised to the occMap contract (native
    0.1 m resolution, inflation (4,4,2)); start/goal over free inflated cells of the z in [0.7,1.3] slab, 2-20 m
    apart.  strict_order: whole-solve control flow and control points identical to the oracle."""
    from helpers import oracle_map_from
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    m = tp.OccMap.from_tpm(os.path.join(root, "data", "maps", name + ".tpm"))
    om = oracle_map_from(orc, m)
    info = m.info()
    inf = m.grid("inflated")
    kz = int(np.floor((1.0 - info["origin"][2]) / info["res"]))
    free = np.argwhere(inf[:, :, kz] == 0)
    rng = np.random.default_rng(17)
    S, G = [], []
    while len(S) < 48:
        a, b = free[rng.integers(len(free))], free[rng.integers(len(free))]
        pa = np.array(info["origin"]) + (np.array([a[0], a[1], kz]) + 0.5) * info["res"]
        pb = np.array(info["origin"]) + (np.array([b[0], b[1], kz]) + 0.5) * info["res"]
        pa[2] = pb[2] = 1.0
        if 2.0 <= np.linalg.norm(pa - pb) <= 20.0:
            S.append(pa); G.append(pb)
    p = tp.default_params()
    p.strict_order = 1
    off, ctrl, valid = tp.frontend_batch(m, p, np.array(S), np.array(G))
    keep = [b for b in range(len(S)) if valid[b] and off[b + 1] - off[b] >= 7][:32]
    assert len(keep) >= 16
    o2 = np.concatenate([[0], np.cumsum([off[b + 1] - off[b] for b in keep])]).astype(np.int32)
    c2 = np.concatenate([ctrl[off[b]:off[b + 1]] for b in keep], 0)
    e = tp.Engine(0)
    e.set_map(m)
    out, res = e.make_plan_batch(p, o2, c2)
    po = om.lib.default_params()
    po.soft_atan2 = 1
    _, out_o, st_o = orc.make_plan_batch(om, po, o2, c2, nthreads=4)
    r = _plan_compare(o2, out, res, out_o, st_o)
    print(name, r)
    assert r["agree"] == len(keep) and r["same_flow"] == len(keep) and r["exact"] == len(keep)
    # default mode: bit-identical to the oracle running the warp-form arithmetic; successes collision free
    p.strict_order = 0
    out2, res2 = e.make_plan_batch(p, o2, c2)
    po.fast_order = 4
    _, out_f, st_f = orc.make_plan_batch(om, po, o2, c2, nthreads=4)
    rf = _plan_compare(o2, out2, res2, out_f, st_f)
    print(name, "default mode vs oracle(fast_order):", rf)
    assert rf["agree"] == len(keep) and rf["same_flow"] == len(keep) and rf["exact"] == len(keep)
    for b in range(len(keep)):
        if res2["status"][b] == 1:
            pl = orc.Planner(om)
            pl.set_ctrl(out2[o2[b]:o2[b + 1]])
            assert not pl.has_collision(), b
    assert abs((res2["status"] == 1).mean() - st_o["success"].mean()) <= 0.15
    e.close()


@pytest.mark.gpu
def test_device_front_end_matches_host_and_oracle(tp, engine, orc, sq_map, sq_omap):
    """(start, goal) pairs -> initial control points on the device (tp_vigo_frontend_batch_device, one warp per pair)
    vs the host C++ front end (same operation order: bit-identical up to pow() roundings) and the numpy restatement
    of src/bspline_node.cpp:332-371 / bsplineTraj.cpp:207-323 / bspline.cpp:74-138 (oracle/frontend_np.py, 1e-9)."""
    from helpers import random_pairs
    from oracle import frontend_np
    rng = np.random.default_rng(7)
    p = tp.default_params()
    S, G = random_pairs(sq_omap, 400, rng, min_dist=0.05)
    # edge cases: goal inside an obstacle, start == goal, very short and very long paths (max_path_length cut)
    occ = np.argwhere(sq_map.grid("inflated") != 0)
    info = sq_map.info()
    bad_goal = np.array(info["origin"]) + (occ[len(occ) // 2] + 0.5) * info["res"]
    S = np.vstack([S, S[:1], S[1:2], [[-9.4, -9.4, 1.0]], [[-9.4, 9.4, 1.0]]])
    G = np.vstack([G, bad_goal[None], S[1:2], [[9.4, 9.4, 1.0]], [[9.4, -9.4, 1.0]]])
    off_h, ctrl_h, valid_h = tp.frontend_batch(sq_map, p, S, G)
    off_d, ctrl_d, valid_d = engine.frontend_batch(p, S, G)
    assert np.array_equal(valid_h, valid_d)
    assert np.array_equal(off_h, off_d)
    assert valid_d[-4] == 0 and valid_d[-3] == 0
    d = np.abs(ctrl_h - ctrl_d)
    print(f"device vs host front end: {len(S)} pairs, {int(valid_d.sum())} valid, {off_d[-1]} control points, "
          f"max |diff| {d.max():.3e}, bit-identical {float(np.mean(d == 0)):.4f}")
    assert d.max() <= 1e-12
    for b in list(range(0, 40)) + [len(S) - 2, len(S) - 1]:
        want = frontend_np.start_goal_to_ctrl(S[b], G[b], sq_omap)
        got = ctrl_d[off_d[b]:off_d[b + 1]]
        if want is None:
            assert len(got) == 0
        else:
            assert got.shape == want.shape and np.max(np.abs(got - want)) <= 1e-9
    # the device front end feeds the solve: same makePlan result as from the host front end
    keep = [b for b in range(len(S)) if valid_d[b] and off_d[b + 1] - off_d[b] >= 7][:64]
    offs = np.concatenate([[0], np.cumsum([off_d[b + 1] - off_d[b] for b in keep])]).astype(np.int32)
    cd = np.concatenate([ctrl_d[off_d[b]:off_d[b + 1]] for b in keep])
    ch = np.concatenate([ctrl_h[off_h[b]:off_h[b + 1]] for b in keep])
    p.strict_order = 1
    out_d, res_d = engine.make_plan_batch(p, offs, cd)
    out_h, res_h = engine.make_plan_batch(p, offs, ch)
    same = [b for b in range(len(keep)) if np.array_equal(cd[offs[b]:offs[b + 1]], ch[offs[b]:offs[b + 1]])]
    assert len(same) >= 0.9 * len(keep)
    for b in same:
        assert np.array_equal(out_d[offs[b]:offs[b + 1]], out_h[offs[b]:offs[b + 1]]) and res_d["status"][b] == res_h["status"][b]


_TMEM_SCRIPT = r'''
import sys, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import trajectory_planner_b200 as tp
from helpers import flat_guides
g = np.load(sys.argv[2])
off, ctrl = g["offsets"], g["ctrl"]
guides = (g["soft0_g_off"], g["soft0_g_cp"], g["soft0_g_p"], g["soft0_g_v"])
eng = tp.Engine(0); eng.set_map(tp.OccMap.from_tpm(sys.argv[1] + "/data/maps/square_static.tpm"))
p = tp.default_params()                      # fast order -> the tensor-memory kernel (TP_LBFGS_TMEM=1)
ps = tp.default_params(); ps.strict_order = 1  # bit-faithful kernel, not affected by the variable
c_t, r_t, _ = eng.optimize_batch(p, off, ctrl, guides)
c_s, r_s, _ = eng.optimize_batch(ps, off, ctrl, guides)
rel = np.abs(r_t["fx"] - r_s["fx"]) / np.abs(r_s["fx"])
same = np.mean((r_t["iters"] == r_s["iters"]) & (r_t["evals"] == r_s["evals"]))
print("optimize: final-cost rel diff median %.2e max %.2e, same (iters, evals) %.2f" % (np.median(rel), rel.max(), same))
assert np.median(rel) <= 1e-3 and rel.max() <= 0.25 and same >= 0.2
out, r = eng.make_plan_batch(p, off, ctrl)
gs = g["soft0_plan_stats"]
agree = np.mean((r["status"] == 1).astype(int) == gs["success"])
print("makePlan: success agreement with the golden run %.3f" % agree)
assert agree >= 0.9
assert not np.any(eng.has_collision_batch(p, off, out)[r["status"] == 1])
'''


@pytest.mark.gpu
def test_tensor_memory_history_variant_subprocess():
    """The L-BFGS kernel with its history in tensor memory (TP_LBFGS_TMEM=1: tcgen05.alloc / st / ld, reduce-scatter
    Gram update) is selected once per process, so it runs in a child: one optimize() against the bit-faithful kernel
    (same statistics as the default fast kernel must meet) and the whole makePlan against the golden outcome."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, TP_LBFGS_TMEM="1")
    r = subprocess.run([sys.executable, "-c", _TMEM_SCRIPT, root, _GOLD], env=env, capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr[-2000:])
    assert r.returncode == 0


_NOPARK_SCRIPT = r'''
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
import trajectory_planner_b200 as tp
z = np.load(sys.argv[2])
eng = tp.Engine(0); eng.set_map(tp.OccMap.from_tpm(sys.argv[1] + "/data/maps/" + (sys.argv[4] if len(sys.argv) > 4 else "square_static") + ".tpm"))
outs = {}
for strict in (0, 1):
    p = tp.default_params(); p.strict_order = strict
    out, res = eng.make_plan_batch(p, z["offsets"], z["ctrl"])
    outs["out%d" % strict] = out; outs["res%d" % strict] = res
np.savez(sys.argv[3], **outs)
'''


@pytest.mark.gpu
def test_park_and_resume_scheduling_changes_no_result(tp, engine, sq_map, sq_omap, tmp_path):
    """k_solve parks trajectories after makePlan steps 1-3 and resumes them later once the batch outnumbers the
    resident workers (1 400 trajectories here; the other tests stay below that).  Scheduling must not change a single
    bit of any result, in either reduction order: compare with a child process that runs with parking disabled
    (TP_PARK_THRESH=-1)."""
    import subprocess, sys
    from helpers import make_problems
    pr = make_problems(tp, sq_map, sq_omap, 1400, seed=77)
    inp, outp = str(tmp_path / "in.npz"), str(tmp_path / "out.npz")
    np.savez(inp, offsets=pr["offsets"], ctrl=pr["ctrl"])
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", _NOPARK_SCRIPT, root, inp, outp], env=dict(os.environ, TP_PARK_THRESH="-1"),
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    ref = np.load(outp)
    for strict in (0, 1):
        p = tp.default_params()
        p.strict_order = strict
        out, res = engine.make_plan_batch(p, pr["offsets"], pr["ctrl"])
        assert np.array_equal(out, ref["out%d" % strict]), strict
        for f in ("status", "lbfgs_iters", "lbfgs_evals", "astar_expansions", "outer_rounds", "final_cost", "linear_factor"):
            assert np.array_equal(res[f], ref["res%d" % strict][f]), (strict, f)
        print(f"strict={strict}: 1400 trajectories bit-identical with and without parking; success rate {np.mean(res['status'] == 1):.3f}")


@pytest.mark.gpu
def test_phase_a_kernel_and_schedule_change_no_result_on_maze(tp, orc, tmp_path):
    """The A*-bound raster: 1 200 trajectories on maze.bt (long searches, flood fills, most goals unreachable) through
    k_phase_a + team blocks with the bucket-ordered phase B, against (a) a child process with parking disabled (everything
    inside the team kernel) — bit-identical in every field, and (b) the CPU oracle of the default mode on a sample."""
    import subprocess, sys
    import bench
    from helpers import oracle_map_from
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    m = tp.OccMap.from_tpm(os.path.join(root, "data", "maps", "maze.tpm"))
    e = tp.Engine(0)
    e.set_map(m)
    p = tp.default_params()
    S, G = bench.octomap_pairs(m.grid("inflated"), m.info(), 1700, np.random.default_rng(5))
    off, ctrl, valid = e.frontend_batch(p, S, G)
    keep = np.flatnonzero((valid != 0) & (np.diff(off) >= 7))[:1200]
    chunks = [ctrl[off[b]:off[b + 1]] for b in keep]
    o2 = np.concatenate([[0], np.cumsum([len(c) for c in chunks])]).astype(np.int32)
    c2 = np.concatenate(chunks, 0)
    inp, outp = str(tmp_path / "in.npz"), str(tmp_path / "out.npz")
    np.savez(inp, offsets=o2, ctrl=c2)
    r = subprocess.run([sys.executable, "-c", _NOPARK_SCRIPT, root, inp, outp, "maze"], env=dict(os.environ, TP_PARK_THRESH="-1"),
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    ref = np.load(outp)
    out, res = e.make_plan_batch(p, o2, c2)
    assert np.array_equal(out, ref["out0"])
    for f in ("status", "lbfgs_iters", "lbfgs_evals", "astar_searches", "astar_expansions", "outer_rounds", "fail_count", "final_cost"):
        assert np.array_equal(res[f], ref["res0"][f]), f
    # the CPU oracle of the default mode on the first 48 trajectories
    om = oracle_map_from(orc, m)
    po = om.lib.default_params()
    po.soft_atan2 = 1
    po.fast_order = 4
    n = 48
    _, out_o, st_o = orc.make_plan_batch(om, po, o2[:n + 1], c2[:o2[n]], nthreads=os.cpu_count() or 4)
    assert np.array_equal(out[:o2[n]], out_o)
    assert np.array_equal(res["astar_expansions"][:n], st_o["astar_expansions"]) and np.array_equal(res["lbfgs_iters"][:n], st_o["lbfgs_iters"])
    print(f"maze, 1200 trajectories: bit-identical with and without k_phase_a / parking; {n} of them bit-identical to the CPU oracle; "
          f"success {np.mean(res['status'] == 1):.3f}, A* expansions per solve {res['astar_expansions'].mean():.0f}")
    e.close()


@pytest.mark.gpu
def test_make_plan_batch_edge_cases(tp, engine, orc, sq_map, sq_omap):
    """Empty batch; ragged batch mixing the smallest legal trajectory (7 control points = one free point), an illegal
    one (6), lengths at the size-class boundaries (40/41, 64/65, 104/105) and long ones (150, 200 control points: the
    largest class, history rows in shared memory up to the 227 KB limit); a trajectory beyond the kernel's capacity
    must fail LOUDLY (error code + message), never silently."""
    from trajectory_planner_b200 import TpError
    p = tp.default_params()
    p.strict_order = 1
    out, res = engine.make_plan_batch(p, np.zeros(1, np.int32), np.zeros((0, 3)))
    assert out.shape == (0, 3) and len(res) == 0
    # straight paths across the map (through obstacles), resampled to the wanted number of control points
    rng = np.random.default_rng(5)
    lens = [7, 6, 8, 40, 41, 64, 65, 104, 105, 150, 200, 12, 7]
    chunks = []
    for n in lens:
        a = np.array([-9.0 + rng.uniform(0, 1), -9.0 + rng.uniform(0, 18), 1.0])
        d = np.array([1.0, rng.uniform(-0.3, 0.3), 0.0])
        d /= np.linalg.norm(d)
        step = min(0.25, 17.0 / max(n - 1, 1))
        chunks.append(a[None, :] + np.arange(n)[:, None] * step * d[None, :])
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    ctrl = np.concatenate(chunks, 0)
    out, res = engine.make_plan_batch(p, off, ctrl)
    assert res["status"][1] == tp.TP_STATUS_INVALID
    assert np.array_equal(out[off[1]:off[2]], ctrl[off[1]:off[2]])
    legal = [b for b in range(len(lens)) if lens[b] >= 7]
    o2 = np.concatenate([[0], np.cumsum([lens[b] for b in legal])]).astype(np.int32)
    c2 = np.concatenate([chunks[b] for b in legal], 0)
    out2, res2 = engine.make_plan_batch(p, o2, c2)
    po = sq_omap.lib.default_params()
    po.soft_atan2 = 1
    _, out_o, st_o = orc.make_plan_batch(sq_omap, po, o2, c2, nthreads=4)
    r = _plan_compare(o2, out2, res2, out_o, st_o)
    print("edge cases, strict vs oracle:", r, "statuses", res2["status"].tolist())
    assert r["agree"] == len(legal) and r["same_flow"] == len(legal) and r["exact"] == len(legal)
    # the same trajectories give the same results inside the ragged batch that also holds the illegal one
    for i, b in enumerate(legal):
        assert np.array_equal(out[off[b]:off[b + 1]], out2[o2[i]:o2[i + 1]]) and res["status"][b] == res2["status"][i]
    # default (fast) order on the same batch: runs, and successful trajectories are collision free
    p.strict_order = 0
    out3, res3 = engine.make_plan_batch(p, o2, c2)
    hit = engine.has_collision_batch(p, o2, out3)
    assert not np.any(hit[res3["status"] == 1])
    po.fast_order = 4
    _, out_f, st_f = orc.make_plan_batch(sq_omap, po, o2, c2, nthreads=4)
    rf = _plan_compare(o2, out3, res3, out_f, st_f)
    print("edge cases, default mode vs oracle(fast_order):", rf)
    assert rf["agree"] == len(legal) and rf["same_flow"] == len(legal) and rf["exact"] == len(legal)
    # beyond capacity: 400 control points need more shared memory than a block can have
    big = np.array([-9.5, 0.0, 1.0])[None, :] + np.arange(400)[:, None] * np.array([0.04, 0.0, 0.0])[None, :]
    with pytest.raises(TpError) as ei:
        engine.make_plan_batch(p, np.array([0, 400], np.int32), big)
    assert "shared memory" in str(ei.value)


@pytest.mark.gpu
def test_multi_engine_entry_matches_single_engine(tp, engine, sq_map, sq_omap):
    """tp_vigo_make_plan_batch_multi (one host thread per engine, chunks from a shared cursor): two engines
    — here both on GPU 0, on a multi-GPU box one per device — must return, for every trajectory, exactly what a single
    engine returns for the whole batch, in both reduction orders, whichever engine solved it."""
    from helpers import make_problems
    pr = make_problems(tp, sq_map, sq_omap, 700, seed=321)
    e2 = tp.Engine(0)
    e2.set_map(sq_map)
    ndev = tp._capi.load().tp_device_count()
    e3 = None
    if ndev > 1:   # a second GPU when the box has one
        e3 = tp.Engine(1)
        e3.set_map(sq_map)
    engines = [engine, e3 if e3 is not None else e2]
    for strict in (0, 1):
        p = tp.default_params()
        p.strict_order = strict
        ref_out, ref_res = engine.make_plan_batch(p, pr["offsets"], pr["ctrl"])
        out, res, who = tp.make_plan_batch_multi(engines, p, pr["offsets"], pr["ctrl"], chunk=128)
        assert np.array_equal(out, ref_out)
        for f in ("status", "lbfgs_iters", "lbfgs_evals", "astar_expansions", "outer_rounds", "final_cost", "linear_factor"):
            assert np.array_equal(res[f], ref_res[f]), (strict, f)
        assert set(np.unique(who)) == {0, 1}, "both engines must have taken chunks"
        out2, res2, _ = tp.make_plan_batch_multi(engines, p, pr["offsets"], pr["ctrl"])   # library-chosen chunk size
        assert np.array_equal(out2, ref_out)
    # an empty batch and a bad argument
    o, r, _ = tp.make_plan_batch_multi(engines, tp.default_params(), np.zeros(1, np.int32), np.zeros((0, 3)))
    assert o.shape == (0, 3) and len(r) == 0
    e2.close()
    if e3 is not None:
        e3.close()
