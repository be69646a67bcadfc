"""GPU parity of the secondary path: batched min-snap solve, trajectory sampling, box collision check and the
insert-waypoint loop (polyTrajSolver / polyTrajOctomap) against the numpy oracle.
Tolerances (north star): min-snap coefficients 1e-9 relative; occupancy / collision decisions bit-exact on identical
positions; sampled positions 1e-12 (CUDA pow vs glibc pow)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def field(tp):
    from oracle import polytraj_np as PT
    m = tp.OccMap.from_tpm(os.path.join(ROOT, "data", "maps", "field.tpm"))
    info = m.info()
    g = PT.Grid3(info["res"], info["origin"], m.grid("occupied"), m.grid("known"))
    e = tp.Engine(0)
    e.set_map(m)
    yield m, g, e
    e.close()


def _random_paths(g, info, rng, counts, clear=True):
    from oracle import polytraj_np as PT
    kz = int(np.floor((1.0 - info["origin"][2]) / info["res"]))
    known_free = np.argwhere((g.known[:, :, kz] != 0) & (g.occ[:, :, kz] == 0))
    box, mr = (0.4, 0.4, 0.2), 0.2
    paths = []
    for nw in counts:
        while True:
            a = known_free[rng.integers(len(known_free))]
            p0 = np.array(info["origin"]) + (np.array([a[0], a[1], kz]) + 0.5) * info["res"]
            p0[2] = 1.0
            if PT.check_collision(g, p0, box, mr):
                continue
            p = [p0]
            ok = True
            for _ in range(nw - 1):
                for _t in range(60):
                    ang, st = rng.uniform(0, 2 * np.pi), rng.uniform(1, 4)
                    q = p[-1] + np.array([st * np.cos(ang), st * np.sin(ang), 0.0])
                    seg_ok = not PT.check_collision(g, q, box, mr)
                    if seg_ok and clear:
                        for s in np.linspace(0, 1, int(st / 0.2) + 2):
                            if PT.check_collision(g, p[-1] + s * (q - p[-1]), box, mr):
                                seg_ok = False
                                break
                    if seg_ok:
                        p.append(q)
                        break
                else:
                    ok = False
                    break
            if ok:
                paths.append(np.array(p))
                break
    return paths


def test_minsnap_solve_parity(tp, field):
    from oracle import frontend_np as F
    m, g, e = field
    rng = np.random.default_rng(1)
    paths = []
    for K in (1, 2, 3, 5, 7, 9, 12, 15, 19, 19, 8, 8):
        wp = [np.array([0.0, 0.0, 1.0])]
        for _ in range(K):
            ang, st = rng.uniform(0, 2 * np.pi), rng.uniform(1, 4)
            wp.append(wp[-1] + np.array([st * np.cos(ang), st * np.sin(ang), rng.uniform(-0.2, 0.2)]))
        paths.append(np.array(wp))
    paths.append(np.array([[0, 0, 1.0], [1, 1, 1.0], [2, 0, 1.0], [4, 10, 1.0]]))   # src/test/waypoint.yaml:2-5
    pt = tp.PolyTraj(e)
    sols, status = pt.solve_batch(paths)
    assert np.all(status == 0)
    worst = 0.0
    for p, (coef, times) in zip(paths, sols):
        co, to = F.minsnap_solve(p, 1.0)
        assert np.max(np.abs(times - to)) <= 1e-12
        for a in range(3):
            worst = max(worst, np.max(np.abs(coef[a] - co[a])) / max(np.max(np.abs(co[a])), 1e-300))
    print("min-snap coefficient worst relative error", worst)
    assert worst <= 1e-9
    # non-rest boundary conditions (updateInitVel / updateInitAcc)
    bc = np.zeros((len(paths), 12))
    bc[:, 0:3] = [0.5, -0.2, 0.0]
    bc[:, 6:9] = [0.1, 0.3, 0.0]
    sols2, st2 = pt.solve_batch(paths, bc)
    co, _ = F.minsnap_solve(paths[5], 1.0, v0=bc[5, 0:3], a0=bc[5, 6:9])
    assert np.max(np.abs(sols2[5][0] - co)) <= 1e-9 * np.max(np.abs(co))
    # degenerate inputs: too few waypoints
    s3, st3 = pt.solve_batch([np.array([[0, 0, 1.0]]), paths[2]])
    assert st3[0] == -3 and st3[1] == 0
    # the band solver's limits: the longest path it accepts (63 segments, n = 882), one more (rejected), and the lower
    # continuity degrees (fewer junction rows: the band layout changes with cont)
    long_paths = []
    for K in (40, 63, 64):
        wp = [np.array([0.0, 0.0, 1.0])]
        for _ in range(K):
            ang, st = rng.uniform(0, 2 * np.pi), rng.uniform(1, 3)
            wp.append(wp[-1] + np.array([st * np.cos(ang), st * np.sin(ang), rng.uniform(-0.1, 0.1)]))
        long_paths.append(np.array(wp))
    with pytest.raises(tp.TpError):   # 64 segments: the whole call fails loudly (TP_ERR_CAPACITY), nothing is solved silently
        pt.solve_batch(long_paths)
    sl, stl = pt.solve_batch(long_paths[:2])
    assert list(stl) == [0, 0]
    for p_, (coef, times) in zip(long_paths[:2], sl[:2]):
        co, to = F.minsnap_solve(p_, 1.0)
        # n = 560 / 882 with a KKT condition number beyond 1e10: BOTH solutions carry ~1e-7 of noise (the oracle's lstsq returns
        # 1e-7 where the rest-to-rest boundary rows force exact zeros), so the comparison is at 1e-6 here, 1e-9 at the benchmark's sizes
        assert np.max(np.abs(coef - co)) <= 1e-6 * np.max(np.abs(co))
    for cont in (2, 3):
        pc = tp.default_poly_params()
        pc.cont = cont
        sc, stc = tp.PolyTraj(e, pc).solve_batch(paths[:8])
        assert np.all(stc == 0)
        for p_, (coef, times) in zip(paths[:8], sc):
            co, to = F.minsnap_solve(p_, 1.0, cont=cont)
            assert np.max(np.abs(coef - co)) <= 1e-9 * np.max(np.abs(co)), cont


def test_minsnap_kernel_against_reference_osqp_golden(tp, field):
    """k_minsnap_solve against the solutions of the REFERENCE's own libosqp.so (tests/golden/minsnap_osqp_golden.npz,
    tools/make_minsnap_golden.py) at that solver's accuracy (eps 1e-3: 2e-3 x path length on sampled positions)."""
    from oracle import polytraj_np as PN
    from test_oracle_cpu import minsnap_golden_cases, MS_TOL_REL
    m, g, e = field
    cases = list(minsnap_golden_cases())
    pt = tp.PolyTraj(e)
    bc = np.array([c[1].reshape(12) for c in cases])
    sols, status = pt.solve_batch([c[0] for c in cases], bc)
    assert np.all(status == 0)
    diffs = []
    for (path, _, times, coef, _, _), (cg, tg) in zip(cases, sols):
        assert np.max(np.abs(tg - times)) <= 1e-12
        a, _ = PN.get_trajectory(coef, times, path[-1], 0.05)
        b, _ = PN.get_trajectory(cg, tg, path[-1], 0.05)
        d = float(np.max(np.abs(a - b)))
        assert d <= MS_TOL_REL * times[-1], (len(path), d)
        diffs.append(d)
    print("k_minsnap_solve vs reference OSQP: %d cases, median %.2e m, worst %.2e m" % (len(diffs), np.median(diffs), max(diffs)))
    assert np.median(diffs) <= 1e-3


def test_box_collision_bit_exact_and_samples(tp, field):
    from oracle import polytraj_np as PT
    m, g, e = field
    info = m.info()
    pt = tp.PolyTraj(e)
    rng = np.random.default_rng(2)
    lo = np.array(g.bbmin) - 0.6
    hi = np.array(g.bbmax) + 0.6
    xyz = rng.uniform(lo, hi, (20000, 3))
    xyz[:2000] = np.round(xyz[:2000], 1)       # voxel faces
    got = pt.box_collision(xyz)
    want = np.array([PT.check_collision(g, p, (0.4, 0.4, 0.2), 0.2) for p in xyz], np.uint8)
    assert np.array_equal(got, want)
    assert 0.05 < got.mean() < 0.999
    # sampling + per-sample decisions + colliding segments on real solves
    paths = _random_paths(g, info, rng, [8, 12, 20, 9], clear=False)
    sols, _ = pt.solve_batch(paths)
    valid, segs, ns, samples, shit = pt.check_batch(paths, sols, want_samples=True)
    for b, p in enumerate(paths):
        traj, _ = PT.get_trajectory(sols[b][0], sols[b][1], p[-1], 0.1)
        assert ns[b] == len(traj)
        assert np.max(np.abs(samples[b] - traj)) <= 1e-12 * max(1.0, np.max(np.abs(traj)))
        # decisions are checked bit-exactly on the DEVICE's own positions (pow differs in the last ulp)
        has, sg, flags = PT.check_collision_traj(g, samples[b], sols[b][1], 0.1, (0.4, 0.4, 0.2), 0.2)
        assert np.array_equal(shit[b].astype(bool), flags)
        assert bool(valid[b]) == (not has) and list(segs[b]) == sg


def test_make_plan_adding_waypoint_loop(tp, field):
    from oracle import polytraj_np as PT
    m, g, e = field
    info = m.info()
    rng = np.random.default_rng(5)
    paths = _random_paths(g, info, rng, [8, 10, 12, 9, 8, 14])
    p = tp.default_poly_params()
    p.max_iter = 6
    p.max_waypoints = 48
    pt = tp.PolyTraj(e, p)
    res = pt.make_plan_batch(paths)
    nvalid = 0
    for b, path in enumerate(paths):
        o = PT.make_plan_adding_waypoint(path, g, max_iter=6, max_waypoints=48)
        assert res[b]["valid"] == o["valid"], b
        assert res[b]["iters"] == o["iters"], b
        assert res[b]["path"].shape == o["path"].shape and np.array_equal(res[b]["path"], o["path"]), b
        assert np.max(np.abs(res[b]["coef"] - o["coef"])) <= 1e-9 * np.max(np.abs(o["coef"])), b
        nvalid += res[b]["valid"]
    assert nvalid >= 1


def test_corridor_qp_against_oracle_and_reference_osqp(tp, field):
    """tp_corridor_solve_batch (corridor-constrained min-snap, polyTrajSolver.cpp:555-579, 813-840, 960-1012) on the golden
    corridor cases: against the oracle's interior-point solution (same algorithm, both stopped at mu < 1e-10 and a relative residual of 1e-6: sampled positions within 2e-3 m (the barrier
    terms reach 1e19 in the last iterations: the two runs stop at slightly different iterates; OSQP's own answers are 1e-2 m off), objective
    within 1e-4 relative, inside the corridor to 1e-9 m), and against the reference's own libosqp.so answers at OSQP's accuracy."""
    from oracle import frontend_np as F
    from oracle import polytraj_np as PN
    from test_oracle_cpu import minsnap_golden_cases
    m, g, e = field
    cases = list(minsnap_golden_cases(corridor=True))
    pt = tp.PolyTraj(e)
    bc = np.array([c[1].reshape(12) for c in cases])
    sols, status = pt.corridor_solve_batch([c[0] for c in cases], [c[5] for c in cases], 8.0, bc)
    assert np.all(status == 0), status
    worst, worst_osqp, worst_obj = 0.0, 0.0, 0.0
    for (path, bcv, times, coef, st_osqp, r), (cg, tg) in zip(cases, sols):
        K = len(path) - 1
        co, to, so, _ = PN.corridor_solve(path, r, 8.0, bc=bcv)
        assert np.all(so == 0) and np.max(np.abs(tg - to)) <= 1e-12
        a, _ = PN.get_trajectory(cg, tg, path[-1], 0.05)
        b, _ = PN.get_trajectory(co, to, path[-1], 0.05)
        worst = max(worst, float(np.max(np.abs(a - b))))
        P = F.minsnap_P(K)
        Ac, lo, hi = PN.corridor_rows(path, times, r, 8.0)
        ref, _ = PN.get_trajectory(coef, times, path[-1], 0.05)
        for ax in range(3):
            x, xo = cg[ax].copy(), co[ax].copy()
            for s in range(K):
                sc = (times[s + 1] - times[s]) ** np.arange(8)
                x[s * 8:(s + 1) * 8] *= sc
                xo[s * 8:(s + 1) * 8] *= sc
            w = Ac @ x
            assert np.max(w - hi[:, ax]) <= 1e-9 and np.max(lo[:, ax] - w) <= 1e-9
            worst_obj = max(worst_obj, abs(x @ P @ x - xo @ P @ xo) / max(xo @ P @ xo, 1.0))
            if st_osqp[ax] == 1:
                worst_osqp = max(worst_osqp, float(np.max(np.abs(a[:, ax] - ref[:, ax]))))
    print("corridor QP: worst |pos| vs the oracle %.2e m (objective %.2e relative), vs the reference OSQP (solved axes) %.2e m" % (worst, worst_obj, worst_osqp))
    assert worst <= 2e-3 and worst_obj <= 1e-4 and worst_osqp <= 5e-2
    # an infeasible corridor (2 mm around a sharp corner) is reported, not returned as a solution
    path = np.array([[0, 0, 1.0], [2, 0, 1.0], [2, 2, 1.0], [0, 2, 1.0]])
    _, st = pt.corridor_solve_batch([path], [np.full(3, 0.002)], 8.0)
    co, to, so, _ = PN.corridor_solve(path, np.full(3, 0.002), 8.0)
    assert np.array_equal(st[0] != 0, so != 0)


def test_make_plan_corridor_constraint_loop(tp, field):
    """polyTrajOctomap::makePlanCorridorConstraint (polyTrajOctomap.cpp:388-530) on field.bt: valid flags, iteration counts and
    final corridor radii identical to the oracle's loop; trajectories within 2e-3 m; valid ones re-checked collision free."""
    from oracle import polytraj_np as PT
    m, g, e = field
    rng = np.random.default_rng(11)
    paths = _random_paths(g, m.info(), rng, [3, 4, 5, 6, 8, 10, 4, 5], clear=False)
    p = tp.default_poly_params()
    p.max_iter = 6
    pt = tp.PolyTraj(e, p)
    out = pt.make_plan_corridor_batch(paths, 0.5, 0.8, 8.0)
    n_valid = n_strict = 0
    for q, r in zip(paths, out):
        o = PT.make_plan_corridor(q, g, 0.5, 0.8, 8.0, max_iter=6)
        assert r["valid"] == o["valid"]
        if np.any(r["status"] != 0) or np.any(o["status"] != 0):
            # a corridor shrunk to the edge of feasibility: whether the last QP still converges within 60 interior-point
            # iterations is decided in the last bits, on either side; such a path is invalid in both
            assert not r["valid"]
            continue
        n_strict += 1
        assert r["iters"] == o["iters"], (r["iters"], o["iters"])
        assert np.allclose(r["r"], o["r"], rtol=0, atol=1e-15)
        a, _ = PT.get_trajectory(r["coef"], r["times"], q[-1], 0.1)
        b, _ = PT.get_trajectory(o["coef"], o["times"], q[-1], 0.1)
        assert np.max(np.abs(a - b)) <= 2e-3
        if r["valid"]:
            n_valid += 1
            has, _, _ = PT.check_collision_traj(g, a, r["times"], 0.1, (0.4, 0.4, 0.2), 0.2)
            assert not has
    assert n_strict >= len(paths) // 2
    print(f"corridor loop: {n_valid}/{len(paths)} valid, {n_strict} compared iteration by iteration, iterations {[r['iters'] for r in out]}")
    # single-waypoint path (polyTrajOctomap.cpp:390-395)
    one = pt.make_plan_corridor_batch([paths[0][:1]])
    assert one[0]["valid"]


def test_poly_traj_occmap_plan(tp, engine, orc, sq_omap):
    """polyTrajOccMap::makePlan (polyTrajOccMap.cpp:326-399) on the ViGO occupancy map: corridor loop with the reference's
    'inflated AND unknown' collision test, boundary velocities honoured; and the corridorConstraint = false mode (one
    equality-only solve, always valid) equal to the plain min-snap solve."""
    from oracle import polytraj_np as PT
    rng = np.random.default_rng(3)
    paths, bcs = [], []
    for k in range(6):
        K1 = int(rng.integers(3, 8))
        p0 = rng.uniform(-8, 8, 2)
        ang, st = rng.uniform(0, 2 * np.pi, K1 - 1), rng.uniform(1, 4, K1 - 1)
        xy = np.vstack([p0, p0 + np.cumsum(np.column_stack([st * np.cos(ang), st * np.sin(ang)]), 0)])
        if k == 5:
            xy[-1] = [30.0, 30.0]   # leaves the map: outside counts as inflated-occupied and unknown -> a collision
        paths.append(np.column_stack([xy, np.full(K1, 1.0)]))
        bc = np.zeros(12)
        bc[0:3] = rng.uniform(-0.5, 0.5, 3)   # initial velocity
        bcs.append(bc)
    # short hops in free space: feasible inside the initial 0.5 m corridors (rest-to-rest segments lag the constant-speed
    # corridor centres by ~0.1 x length) and collision free
    paths.append(np.array([[-9, 0, 1.0], [-7.5, 0.3, 1.0], [-6, -0.2, 1.0], [-4.5, 0, 1.0]]))
    bcs.append(np.zeros(12))
    p = tp.default_poly_params()
    p.max_iter = 5
    pt = tp.PolyTraj(engine, p)
    out = pt.make_plan_corridor_batch(paths, 0.5, 0.8, 8.0, bc=np.array(bcs), occmap=True)
    n_cmp = 0
    for q, bc, r in zip(paths, bcs, out):
        o = PT.make_plan_corridor(q, None, 0.5, 0.8, 8.0, max_iter=5, occmap=sq_omap, bc=bc.reshape(4, 3))
        assert r["valid"] == o["valid"]
        if np.any(r["status"] != 0) or np.any(o["status"] != 0):
            continue
        n_cmp += 1
        assert r["iters"] == o["iters"] and np.allclose(r["r"], o["r"], rtol=0, atol=1e-15)
        a, _ = PT.get_trajectory(r["coef"], r["times"], q[-1], 0.1)
        b, _ = PT.get_trajectory(o["coef"], o["times"], q[-1], 0.1)
        assert np.max(np.abs(a - b)) <= 2e-3
    assert n_cmp >= 3 and not out[5]["valid"] and out[6]["valid"] and out[6]["iters"] == 1
    plain = pt.make_plan_corridor_batch(paths, bc=np.array(bcs), occmap=True, corridor_constraint=False)
    sols, st = pt.solve_batch(paths, np.array(bcs))
    for r, (c, t) in zip(plain, sols):
        assert r["valid"] and r["iters"] == 1 and np.max(np.abs(r["coef"] - c)) <= 1e-9 * np.max(np.abs(c))
    print(f"polyTrajOccMap: {n_cmp} paths compared iteration by iteration, valid {[r['valid'] for r in out]}, iterations {[r['iters'] for r in out]}")
