"""Golden vectors for the min-snap path, produced by the REFERENCE's own QP solver.

Runs the reference's prebuilt libosqp.so (through oracle/_ref/libosqp_ref.so, built by `make -C oracle ref` from
oracle/osqp_ref.c against the reference's osqp headers) on the QPs polyTrajSolver builds (constructP / constructA /
constructBound, polyTrajSolver.cpp:241-847, restated in oracle/frontend_np.py and oracle/polytraj_np.py) with the
reference's settings (OsqpEigen defaults, verbosity off: polyTrajSolver.cpp:162-223), de-normalises the solution as
solveX does (:870-879) and writes tests/golden/minsnap_osqp_golden.npz.  Only runs in the build container (needs
/root/reference); the npz travels.  The golden solutions are OSQP's eps = 1e-3 ADMM answers: the tests compare the
exact KKT oracle and the CUDA kernels with them at that solver's accuracy (tolerances in the tests).

    python tools/make_minsnap_golden.py
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import frontend_np as F   # noqa: E402
from oracle import polytraj_np as PN  # noqa: E402

LIB = os.path.join(ROOT, "oracle", "_ref", "libosqp_ref.so")


def osqp_ref(P, q, A, l, u, mode=0):
    L = ctypes.CDLL(LIB)
    L.osqp_ref_solve.restype = ctypes.c_longlong
    n, m = len(q), len(l)
    x, y, info = np.zeros(n), np.zeros(m), np.zeros(4)
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (P, q, A, l, u)]
    c = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    st = L.osqp_ref_solve(ctypes.c_longlong(n), ctypes.c_longlong(m), *[c(a) for a in arrs], ctypes.c_int(mode), c(x), c(y), c(info))
    return int(st), x, y, info


def denorm(x, times, n=8):
    c = x.copy()
    for s in range(len(times) - 1):
        dt = times[s + 1] - times[s]
        for d in range(n):
            c[s * n + d] /= dt ** d
    return c


def main():
    rng = np.random.default_rng(20261019)
    recs = {}
    k = 0
    for K1 in [2, 3, 4, 5, 6, 8, 10, 12, 14, 16, 18, 20] * 2:
        # a random walk with 1-4 m steps (the bench's waypoint generator without the map), z = 1 +- 0.3
        ang = rng.uniform(0, 2 * np.pi, K1 - 1)
        st = rng.uniform(1, 4, K1 - 1)
        path = np.zeros((K1, 3))
        path[:, 2] = 1.0 + rng.uniform(-0.3, 0.3, K1)
        path[1:, 0] = np.cumsum(st * np.cos(ang))
        path[1:, 1] = np.cumsum(st * np.sin(ang))
        bc = np.zeros((4, 3))
        if k % 3 == 1:   # moving start (polyTrajSolver::updateInitVel / updateInitAcc)
            bc[0] = rng.uniform(-1, 1, 3)
            bc[2] = rng.uniform(-0.5, 0.5, 3)
        corridor = (k % 4 == 2) and K1 >= 3
        K = K1 - 1
        seglen = np.linalg.norm(np.diff(path, axis=0), axis=1)
        times = np.concatenate([[0.0], np.cumsum(seglen / 1.0)])
        P = F.minsnap_P(K)
        A, b = F.minsnap_Ab(path, times, bc[0], bc[1], bc[2], bc[3])
        lo, hi = b.copy(), b.copy()
        r = None
        if corridor:
            r = np.full(K, 0.5) * (0.8 ** rng.integers(0, 3, K))
            Ac, lc, uc = PN.corridor_rows(path, times, r, 8.0)
            A = np.vstack([A, Ac])
            lo = np.vstack([lo, lc])
            hi = np.vstack([hi, uc])
        coef = np.zeros((3, 8 * K))
        status = np.zeros(3, int)
        iters = np.zeros(3, int)
        for ax in range(3):
            stv, x, _, info = osqp_ref(P, np.zeros(8 * K), A, lo[:, ax], hi[:, ax], 0)
            coef[ax] = denorm(x, times)
            status[ax] = stv
            iters[ax] = int(info[0])
        recs[f"path_{k}"] = path
        recs[f"bc_{k}"] = bc
        recs[f"times_{k}"] = times
        recs[f"coef_{k}"] = coef
        recs[f"status_{k}"] = status
        recs[f"iters_{k}"] = iters
        recs[f"corridor_{k}"] = np.zeros(0) if r is None else r
        print(k, "K", K, "corridor" if corridor else "", "status", status, "iters", iters)
        k += 1
    recs["count"] = np.array(k)
    out = os.path.join(ROOT, "tests", "golden", "minsnap_osqp_golden.npz")
    np.savez_compressed(out, **recs)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
