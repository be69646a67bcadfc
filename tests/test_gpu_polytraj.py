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
