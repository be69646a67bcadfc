"""CPU tests of the ORACLE (test infrastructure): pinned to the reference's own solver/lbfgs.hpp, to the
committed golden vectors, and to analytic known answers.  The reference holds no golden vectors for this
path, so the pins are: (1) oracle/_ref = the restatement linked against the reference's
lbfgs.hpp, (2) tests/golden/vigo_golden.npz generated from (1) by tools/make_golden.py, (3) KATs."""
import os

import numpy as np
import pytest

from helpers import make_problems, oracle_map_from, traj

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "vigo_golden.npz")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "liborc_ref.so")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_oracle_reproduces_golden_vectors(orc, sq_omap, gold):
    """The oracle with ITS OWN L-BFGS port reproduces, bit for bit, vectors generated through the
    reference's lbfgs.hpp: cost/gradient, one optimize(), collision decisions, the whole makePlan."""
    off, ctrl, pert = gold["offsets"], gold["ctrl"], gold["ctrl_perturbed"]
    B = len(off) - 1
    for soft in (0, 1):
        tag = f"soft{soft}"
        po = sq_omap.lib.default_params()
        po.soft_atan2 = soft
        g_off = gold[f"{tag}_g_off"]
        gpos = 0
        for b in range(B):
            c = ctrl[off[b]:off[b + 1]]
            pl = orc.Planner(sq_omap, po)
            pl.set_ctrl(c)
            assert int(pl.has_collision()) == gold["has_collision"][b]
            pl.init_guides()
            cp, gp, gv = pl.get_guides()
            sl = slice(g_off[b], g_off[b + 1])
            assert np.array_equal(cp, gold[f"{tag}_g_cp"][sl])
            assert np.array_equal(gp, gold[f"{tag}_g_p"][sl]) and np.array_equal(gv, gold[f"{tag}_g_v"][sl])
            pc = pert[off[b]:off[b + 1]]
            pl2 = orc.Planner(sq_omap, po)
            pl2.set_ctrl(pc)
            pl2.add_guides(cp, gp, gv)
            f, g, _ = pl2.cost(pc[3:-3].ravel())
            assert f == gold[f"{tag}_cost"][b]
            assert np.array_equal(g, gold[f"{tag}_grad"][gpos:gpos + len(g)])
            gpos += len(g)
            o = pl.optimize()
            # (the reference's lbfgs_optimize does not report its iteration count: golden iters = -1)
            assert [o["ret"], o["evals"]] == list(gold[f"{tag}_opt_stats"][b, [0, 2]].astype(int))
            assert o["fx"] == gold[f"{tag}_opt_stats"][b, 3]
            assert np.array_equal(pl.get_ctrl(), gold[f"{tag}_opt_ctrl"][off[b]:off[b + 1]])
        _, out, st = orc.make_plan_batch(sq_omap, po, off, ctrl, nthreads=4)
        assert np.array_equal(out, gold[f"{tag}_plan_ctrl"])
        gs = gold[f"{tag}_plan_stats"]
        okm = gs["success"] == 1
        for k in st.dtype.names:
            if k != "lbfgs_iters":
                sel = okm if k == "linear_factor" else slice(None)
                assert np.array_equal(st[k][sel], gs[k][sel]), k


def test_oracle_map_queries_and_astar_golden(orc, sq_omap, gold):
    assert np.array_equal(sq_omap.query(gold["q_xyz"]), gold["q_hit"])
    assert np.array_equal(sq_omap.query_unknown(gold["q_xyz"]), gold["q_unknown"])
    assert np.array_equal(sq_omap.query_lines(gold["q_xyz"], gold["q_b"]), gold["q_line"])
    pl = orc.Planner(sq_omap)
    pos = 0
    for s, e, n, ex in zip(gold["astar_starts"], gold["astar_ends"], gold["astar_len"], gold["astar_exp"]):
        p, x = pl.astar(s, e)
        assert x == ex
        assert (p is None) == (n < 0)
        if p is not None:
            assert np.array_equal(p, gold["astar_paths"][pos:pos + n])
            pos += n


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built (needs /root/reference)")
def test_lbfgs_port_is_pinned_to_reference_header(orc, sq_map):
    """lbfgs_port.hpp vs the reference's solver/lbfgs.hpp (oracle/_ref, use_ref_lbfgs=1) on the ViGO cost
    function: same return code, iteration / evaluation counts and bit-identical iterates, over 96 seeded
    problems and over whole makePlan runs (which chain ~5 optimize() calls with escalating weights)."""
    import trajectory_planner_b200 as tp
    omr = oracle_map_from(orc, sq_map, ref=True)
    omp = oracle_map_from(orc, sq_map, ref=False)
    assert omr.lib.L.orc_is_ref_build() == 1 and omp.lib.L.orc_is_ref_build() == 0
    pr = make_problems(tp, sq_map, omp, 96, seed=4242)
    off = pr["offsets"]
    pr_ = omr.lib.default_params()
    pr_.use_ref_lbfgs = 1
    pp = omp.lib.default_params()
    rets = set()
    for b in range(len(off) - 1):
        a = orc.Planner(omr, pr_)
        c = orc.Planner(omp, pp)
        for pl in (a, c):
            pl.set_ctrl(traj(pr, b))
            pl.init_guides()
        oa, oc = a.optimize(), c.optimize()
        # (iteration count is not reported by the reference's lbfgs_optimize)
        assert (oa["ret"], oa["evals"], oa["fx"]) == (oc["ret"], oc["evals"], oc["fx"]), (b, oa, oc)
        assert np.array_equal(oa["x"], oc["x"])
        assert np.array_equal(a.get_ctrl(), c.get_ctrl())
        rets.add(oa["ret"])
    assert len(rets) >= 2  # convergence and at least one other exit are exercised
    _, out_r, st_r = orc.make_plan_batch(omr, pr_, off, pr["ctrl"], nthreads=4)
    _, out_p, st_p = orc.make_plan_batch(omp, pp, off, pr["ctrl"], nthreads=4)
    assert np.array_equal(out_r, out_p)
    ok = st_r["success"] == 1
    for k in st_r.dtype.names:
        if k == "lbfgs_iters":
            continue   # not reported by the reference's lbfgs_optimize
        sel = ok if k == "linear_factor" else slice(None)   # only defined when makePlan succeeded
        assert np.array_equal(st_r[k][sel], st_p[k][sel]), k


# ------------------------------------------------------------------------------------ known answers
def _planner(orc, omap, ctrl, **kw):
    po = omap.lib.default_params()
    for k, v in kw.items():
        setattr(po, k, v)
    pl = orc.Planner(omap, po)
    pl.set_ctrl(np.asarray(ctrl, float))
    return pl


def test_kat_collinear_control_points_have_zero_cost(orc, sq_omap):
    """Equally spaced collinear control points at 0.1 m spacing: jerk = 0, |v| = 0.5 <= 1, a = 0 ->
    every term is exactly 0 and so is the gradient (bsplineTraj.cpp:934-999)."""
    c = np.array([[0.125 * i, 0.0625 * i, 1.0] for i in range(20)])   # dyadic spacing: exact arithmetic
    pl = _planner(orc, sq_omap, c)
    f, g, terms = pl.cost(c[3:-3].ravel())
    assert f == 0.0 and np.all(g == 0.0) and np.all(np.asarray(terms) == 0.0)
    c = np.array([[0.1 * i, 0.05 * i, 1.0] for i in range(20)])        # inexact spacing: rounding only
    f, g, _ = _planner(orc, sq_omap, c).cost(c[3:-3].ravel())
    assert 0.0 <= f <= 1e-28 and np.max(np.abs(g)) <= 1e-13


def test_kat_feasibility_and_smoothness_hand_computed(orc, sq_omap):
    """One displaced control point, hand-evaluated with the reference's formulas:
    smoothness S = sum |c[i+3]-3c[i+2]+3c[i+1]-c[i]|^2; feasibility with maxVel=maxAcc=1 HARD-CODED
    (bsplineTraj.cpp:955-956), k = 1/0.2^2 = 25."""
    N = 12
    c = np.array([[0.1 * i, 0.0, 1.0] for i in range(N)])
    c[6, 1] = 0.3
    pl = _planner(orc, sq_omap, c)
    f, g, terms = pl.cost(c[3:-3].ravel())
    S = sum(np.sum((c[i + 3] - 3 * c[i + 2] + 3 * c[i + 1] - c[i]) ** 2) for i in range(N - 3))
    k = 25.0
    F = 0.0
    for i in range(N - 1):
        v = (c[i + 1] - c[i]) / 0.2
        F += np.sum(np.where(v > 1, (v - 1) ** 2 * k, 0) + np.where(v < -1, (v + 1) ** 2 * k, 0))
    for i in range(N - 2):
        a = (c[i + 2] - 2 * c[i + 1] + c[i]) * k
        F += np.sum(np.where(a > 1, (a - 1) ** 2, 0) + np.where(a < -1, (a + 1) ** 2, 0))
    assert F > 0 and S > 0
    assert abs(terms[1] - S) <= 1e-12 * S and abs(terms[2] - F) <= 1e-12 * F
    assert terms[0] == 0.0 and abs(f - (S + F)) <= 1e-12 * f


def test_kat_distance_cost_pieces(orc, sq_omap):
    """The three pieces of the distance barrier (bsplineTraj.cpp:852-894) for one guide pair, d_thr=0.5."""
    N = 10
    base = np.array([[0.1 * i, 0.0, 1.0] for i in range(N)])
    for dist, expect in [(1.2, (1.2 - 0.5 - 0.0) ** 3 if False else None), (0.3, 0.2 ** 3), (-0.2, None)]:
        c = base.copy()
        pl = _planner(orc, sq_omap, c)
        gp = c[4] - np.array([0.0, dist, 0.0])     # (c - p) . v = dist with v = +y
        pl.add_guides(np.array([4], np.int32), gp[None], np.array([[0.0, 1.0, 0.0]]))
        f, g, terms = pl.cost(c[3:-3].ravel())
        e = 0.5 - dist
        if e <= -0.5:
            want = (-e) ** 3
        elif 0 < e <= 0.5:
            want = e ** 3
        elif e >= 0.5:
            want = 3 * 0.5 * e * e - 3 * 0.25 * e + 0.125
        else:
            want = 0.0
        assert abs(terms[0] - want) <= 1e-15 + 1e-13 * abs(want), (dist, terms[0], want)


@pytest.mark.parametrize("plan_in_z", [0, 1])
def test_gradient_matches_finite_differences(orc, sq_omap, problems, plan_in_z):
    """Central differences of the oracle cost vs its analytic gradient.  With plan_in_z=1 the reference
    writes the height-barrier gradient into the X row (bsplineTraj.cpp:904-925): that term is therefore
    NOT a true gradient and is excluded by comparing only where the barrier is inactive (z = 1.0 sits
    inside [0.7+0.2, 1.3-0.2] -> active cubic piece; so we just check plan_in_z=0 strictly and the
    y/z rows for plan_in_z=1)."""
    rng = np.random.default_rng(1)
    for b in range(4):
        c = traj(problems, b) + rng.normal(0, 0.1, traj(problems, b).shape)
        pl0 = _planner(orc, sq_omap, traj(problems, b), plan_in_z=plan_in_z)
        pl0.init_guides()
        cp, gp, gv = pl0.get_guides()
        pl = _planner(orc, sq_omap, c, plan_in_z=plan_in_z)
        pl.add_guides(cp, gp, gv)
        x = c[3:-3].ravel().copy()
        f, g, _ = pl.cost(x)
        h = 1e-6
        idx = rng.choice(len(x), size=min(24, len(x)), replace=False)
        for i in idx:
            if plan_in_z and i % 3 != 1:
                continue
            xp, xm = x.copy(), x.copy()
            xp[i] += h
            xm[i] -= h
            fd = (pl.cost(xp)[0] - pl.cost(xm)[0]) / (2 * h)
            assert abs(fd - g[i]) <= 1e-4 * max(1.0, abs(g[i])), (b, i, fd, g[i])


def test_de_boor_matches_scipy_bspline(orc):
    from scipy.interpolate import BSpline
    rng = np.random.default_rng(2)
    N, ts = 23, 0.2
    c = rng.normal(0, 1, (N, 3))
    knots = (np.arange(N + 4) - 3) * ts                # bspline.cpp:19-28
    t = np.linspace(0, (N - 3) * ts, 301)
    want = BSpline(knots, c, 3)(t)
    got = orc.bspline_at(c, t)
    assert np.max(np.abs(got - want)) <= 1e-12
    wantd = BSpline(knots, c, 3).derivative(1)(t[:-1])
    gotd = orc.bspline_deriv_at(c, t[:-1], 1)
    assert np.max(np.abs(gotd - wantd)) <= 1e-10


# ------------------------------------------------------------------------------------ the benchmarked arithmetic
# oracle/wform_port.hpp restates the product's default kernel (lean team form).  The GPU tests pin the kernel to it bit
# for bit; the tests below quantify, on the CPU, how that arithmetic relates to the REFERENCE-ORDER iterate.

def test_kernel_schedule_of_the_recurrences_equals_the_stepwise_form(orc):
    """The kernel runs the two 16-step triangular recurrences of the coefficient-space two-loop four steps per round of
    shuffles (tp_lbfgs_warp.cuh: wf_coeffs); oracle/wform_port.hpp restates both that schedule (coeffs_blocked) and the
    step-by-step form (coeffs, the one every solve of the restatement uses).  On 200,000 random Gram blocks — masked as the
    Gram update leaves them or dense, with signed zeros, 1e300 / 1e-300, infinities and NaNs mixed in — the coefficients
    and g.d must be equal bit for bit (two NaNs count as equal)."""
    bad = sum(orc.lib().L.orc_wform_coeffs_schedules_check(seed, 50000) for seed in range(4))
    assert bad == 0


def test_fast_order_evaluation_and_direction_vs_reference_order(orc, sq_map, sq_omap, tp):
    """(1) cost and gradient of the lean form vs the reference-order restatement: <= 1e-10 relative per evaluation
    (north-star tolerance; measured ~1e-15).  (2) Every L-BFGS iteration of a lean-form solve: the Gram-form direction vs
    the reference's two-loop recursion (lbfgs.hpp:1293-1316, serial order) on the SAME (g, S, Y): <= 1e-10 relative
    (measured ~1e-14).  Both team sizes (1 and 4 warps)."""
    pr = make_problems(tp, sq_map, sq_omap, 48, seed=5)
    rng = np.random.default_rng(11)
    for nw in (4, 1):
        po = sq_omap.lib.default_params()
        po.fast_order = nw
        worst_f = worst_g = worst_d = 0.0
        iters = 0
        for b in range(48):
            pl = orc.Planner(sq_omap, po)
            c = traj(pr, b) + rng.normal(0, 0.1, traj(pr, b).shape)
            pl.set_ctrl(traj(pr, b))
            pl.init_guides()
            guides = pl.get_guides()
            pl2 = orc.Planner(sq_omap, po)
            pl2.set_ctrl(c)
            pl2.add_guides(*guides)
            x = c[3:-3].ravel()
            f0, g0, _ = pl2.cost(x)
            f1, g1 = pl2.cost_wform(x)
            worst_f = max(worst_f, abs(f0 - f1) / abs(f0))
            worst_g = max(worst_g, float(np.max(np.abs(g0 - g1)) / np.max(np.abs(g0))))
            w, st = pl.wform_direction_check()
            worst_d = max(worst_d, w)
            iters += st["iters"]
        print(f"team of {nw} warp(s): evaluation rel diff cost {worst_f:.2e} grad {worst_g:.2e}; direction vs two-loop "
              f"{worst_d:.2e} over {iters} iterations")
        assert worst_f <= 1e-10 and worst_g <= 1e-10
        assert worst_d <= 1e-10 and iters > 2000


def test_fast_order_vs_reference_order_study(orc, sq_map, sq_omap, tp):
    """4 096 seeded problems, whole makePlan on the CPU in both arithmetics (same deterministic atan2): the lean form is
    the same algorithm with different roundings, and the solve amplifies roundings (jerk Hessian condition ~1e9, loose
    g_epsilon) — any two reference builds differ the same way (libm vs soft atan2 flips 1/3 of the flows).  Asserted:
    equal success rate within 1 %, equal final-cost distribution, every lean-form success collision free; recorded: the
    fraction with identical control flow and, of those, within 1e-6 m."""
    B = 4096
    pr = make_problems(tp, sq_map, sq_omap, B, seed=20261018)
    off, ctrl = pr["offsets"], pr["ctrl"]
    po = sq_omap.lib.default_params()
    po.soft_atan2 = 1
    nthreads = os.cpu_count() or 4
    _, out_r, st_r = orc.make_plan_batch(sq_omap, po, off, ctrl, nthreads=nthreads)
    po.fast_order = 4
    _, out_f, st_f = orc.make_plan_batch(sq_omap, po, off, ctrl, nthreads=nthreads)
    keys = ["outer_rounds", "fail_count", "lbfgs_runs", "lbfgs_iters", "lbfgs_evals", "astar_searches", "astar_expansions"]
    same = np.ones(B, bool)
    for k in keys + ["success"]:
        same &= st_r[k] == st_f[k]
    d = np.array([np.abs(out_r[off[b]:off[b + 1]] - out_f[off[b]:off[b + 1]]).max() for b in range(B)])
    sr, sf = st_r["success"].mean(), st_f["success"].mean()
    both = (st_r["success"] == 1) & (st_f["success"] == 1)
    fr, ff = st_r["final_cost"][both], st_f["final_cost"][both]
    print(f"reference order: success {sr:.4f}, iterations/solve {st_r['lbfgs_iters'].mean():.1f}; lean form: success {sf:.4f}, "
          f"iterations/solve {st_f['lbfgs_iters'].mean():.1f}; identical control flow {same.mean():.3f}, of those within 1e-6 m "
          f"{(d[same] <= 1e-6).mean():.3f} (worst {d[same].max():.2e}); final cost median ref {np.median(fr):.4f} lean {np.median(ff):.4f}, "
          f"median |rel diff| {np.median(np.abs(fr - ff) / fr):.2e}")
    assert abs(sr - sf) <= 0.01
    assert abs(st_r["lbfgs_iters"].mean() - st_f["lbfgs_iters"].mean()) <= 0.03 * st_r["lbfgs_iters"].mean()
    assert abs(np.median(fr) - np.median(ff)) <= 1e-3 * np.median(fr)
    assert np.median(np.abs(fr - ff) / fr) <= 1e-3
    assert same.mean() >= 0.15
    hit = 0
    for b in np.flatnonzero(st_f["success"] == 1)[:512]:
        pl = orc.Planner(sq_omap)
        pl.set_ctrl(out_f[off[b]:off[b + 1]])
        hit += pl.has_collision()
    assert hit == 0


MS_GOLD = os.path.join(ROOT, "tests", "golden", "minsnap_osqp_golden.npz")


def minsnap_golden_cases(corridor=False):
    """(path, bc[4,3], times, coef[3, 8K], status[3], corridor sizes) of tests/golden/minsnap_osqp_golden.npz: solutions of the
    REFERENCE's own libosqp.so with the reference's settings (tools/make_minsnap_golden.py, oracle/osqp_ref.c)."""
    g = np.load(MS_GOLD)
    for k in range(int(g["count"])):
        r = g[f"corridor_{k}"]
        if (len(r) > 0) != corridor:
            continue
        yield g[f"path_{k}"], g[f"bc_{k}"], g[f"times_{k}"], g[f"coef_{k}"], g[f"status_{k}"], r


# OSQP stops at eps_abs = eps_rel = 1e-3 on its (scaled) residuals: its coefficients reproduce the exact minimiser's
# trajectory to about 1e-3 of the path's extent.  Tolerance of the pin: 2e-3 x path length, median over the cases 1e-3 m.
MS_TOL_REL = 2e-3


def test_minsnap_kkt_oracle_is_pinned_to_the_reference_osqp():
    """The exact KKT oracle (oracle/frontend_np.minsnap_solve) against the solutions the reference's own prebuilt
    libosqp.so returns for the same QPs (polyTrajSolver.cpp:162-223, 870-879), at that solver's accuracy."""
    from oracle import frontend_np as F
    from oracle import polytraj_np as PN
    diffs = []
    for path, bc, times, coef, status, _ in minsnap_golden_cases():
        assert np.all(status == 1)   # OSQP_SOLVED
        c2, t2 = F.minsnap_solve(path, 1.0, bc[0], bc[1], bc[2], bc[3])
        assert np.max(np.abs(t2 - times)) <= 1e-12
        a, _ = PN.get_trajectory(coef, times, path[-1], 0.05)
        b, _ = PN.get_trajectory(c2, t2, path[-1], 0.05)
        d = float(np.max(np.abs(a - b)))
        assert d <= MS_TOL_REL * times[-1], (len(path), d)
        diffs.append(d)
        # the objectives agree (OSQP's point is only approximately feasible, so its objective may sit slightly BELOW the minimum)
        K = len(path) - 1
        P = F.minsnap_P(K)

        def obj(c):
            x = c.copy()
            for s in range(K):
                x[s * 8:(s + 1) * 8] *= (times[s + 1] - times[s]) ** np.arange(8)
            return 0.5 * x @ P @ x
        for ax in range(3):
            assert abs(obj(c2[ax]) - obj(coef[ax])) <= 5e-2 * obj(coef[ax]) + 1e-9
    print("KKT oracle vs reference OSQP: %d cases, median %.2e m, worst %.2e m" % (len(diffs), np.median(diffs), max(diffs)))
    assert np.median(diffs) <= 1e-3


def test_corridor_qp_oracle_certificate_and_reference_osqp():
    """The corridor-constrained min-snap QP (polyTrajSolver.cpp:555-579, 813-840, 960-1012) solved by the oracle's
    interior-point method: (1) optimality CERTIFICATE — stationarity, primal feasibility, sign and complementarity of the
    multipliers — so the solution is the unique minimiser independently of the algorithm; (2) against the REFERENCE's
    own libosqp.so on the same QPs (golden): OSQP (eps 1e-3) violates the corridor by millimetres and sits within
    centimetres of the exact minimiser, with an objective that is not above it."""
    from oracle import frontend_np as F
    from oracle import polytraj_np as PN
    n_cases = 0
    for path, bc, times, coef, status, r in minsnap_golden_cases(corridor=True):
        K = len(path) - 1
        c2, t2, st, raw = PN.corridor_solve(path, r, 8.0, bc=bc)
        assert np.all(st == 0) and np.max(np.abs(t2 - times)) <= 1e-12
        P = F.minsnap_P(K)
        Aeq, b = F.minsnap_Ab(path, times, bc[0], bc[1], bc[2], bc[3])
        Ac, lo, hi = PN.corridor_rows(path, times, r, 8.0)
        a, _ = PN.get_trajectory(coef, times, path[-1], 0.05)
        bb, _ = PN.get_trajectory(c2, t2, path[-1], 0.05)
        for ax in range(3):
            c, y, mu = raw[ax]["c"], raw[ax]["y"], raw[ax]["mu"]
            w = Ac @ c
            scale = 1.0 + np.max(np.abs(P @ c))
            assert np.max(np.abs(P @ c + Aeq.T @ y + Ac.T @ mu)) <= 1e-6 * scale          # stationarity
            assert np.max(np.abs(Aeq @ c - b[:, ax])) <= 1e-8                                # equality rows
            assert np.max(w - hi[:, ax]) <= 1e-9 and np.max(lo[:, ax] - w) <= 1e-9           # inside the corridor
            assert np.all((mu <= 1e-5) | (hi[:, ax] - w <= 1e-4)) and np.all((mu >= -1e-5) | (w - lo[:, ax] <= 1e-4))   # signs
            assert np.max(np.abs(mu) * np.minimum(hi[:, ax] - w, w - lo[:, ax])) <= 1e-6     # complementarity
            # the reference's OSQP answer for this axis
            x = coef[ax].copy()
            for s in range(K):
                x[s * 8:(s + 1) * 8] *= (times[s + 1] - times[s]) ** np.arange(8)
            wo = Ac @ x
            viol = max(np.max(wo - hi[:, ax]), np.max(lo[:, ax] - wo), 0.0)
            if status[ax] == 1:   # OSQP_SOLVED (2 = its iteration limit)
                assert viol <= 2e-2 and np.max(np.abs(a[:, ax] - bb[:, ax])) <= 5e-2
                assert 0.5 * x @ P @ x <= 0.5 * c @ P @ c * (1 + 1e-9)   # an infeasible point can only be cheaper
        n_cases += 1
    assert n_cases >= 6
