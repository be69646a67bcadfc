"""ORACLE — TEST INFRASTRUCTURE ONLY.

numpy restatement of the *front end* of the reference's ViGO demo node: start/goal -> seed
min-snap path -> resampled input path -> least-squares uniform cubic B-spline control points.
It pins the product's C++ front end (trajectory_planner_b200/csrc/tp_frontend.cpp) in tests.

Follows: src/bspline_node.cpp:317-371 (pipeline), polyTrajSolver.cpp:125-138 (time
allocation), :241-271 (P), :314-584 (A, equality rows), :587-813 (bounds), :870-879 (de-
normalisation), :1051-1071 (getPos), polyTrajOccMap.cpp:434-446 (getTrajectory),
bsplineTraj.cpp:207-245 (inputPathCheck), :247-288 (fillPath), :290-323 (updatePath),
:754-793 (adjustPathLengthDirect; its function-static prevPathLength is an explicit 0.0 here),
bspline.cpp:74-138 (parameterizeToBspline).

The reference solves the seed QP with OSQP (eps 1e-3, nondeterministic adaptive rho); this
restatement solves the same equality-constrained QP exactly through its KKT system (SURVEY.md
§8c), so it is the *intended* solution, not OSQP's 1e-3 approximation of it.
"""
import numpy as np


def minsnap_P(K, deg=7, diff=4):
    n = deg + 1
    P = np.zeros((K * n, K * n))
    for s in range(K):
        for i in range(diff, n):
            for j in range(diff, n):
                f = 1.0
                for d in range(diff):
                    f *= float((i - d) * (j - d))
                f /= float(i + j - 2 * diff + 1)
                P[s * n + i, s * n + j] = f
    return P


def _deriv_row(t, order, n):
    row = np.zeros(n)
    for d in range(order, n):
        c = 1.0
        for k in range(order):
            c *= (d - k)
        row[d] = c * (t ** (d - order))
    return row


def minsnap_Ab(path, times, v0, v1, a0, a1, deg=7, cont=4):
    """Equality rows in the reference's order (constructA / constructBound).  path [K+1,3]."""
    path = np.asarray(path, float)
    K = len(path) - 1
    n = deg + 1
    dts = np.diff(times)
    rows, rhs = [], []

    def add(row, b):
        rows.append(row)
        rhs.append(np.asarray(b, float))

    def seg_row(seg, t, order, scale=1.0):
        r = np.zeros(K * n)
        r[seg * n:(seg + 1) * n] = _deriv_row(t, order, n) * scale
        return r

    add(seg_row(0, 0.0, 0), path[0])
    add(seg_row(K - 1, 1.0, 0), path[-1])
    for i in range(K - 1):
        add(seg_row(i, 1.0, 0), path[i + 1])
    for i in range(K - 1):
        add(seg_row(i, 1.0, 0) - seg_row(i + 1, 0.0, 0), np.zeros(3))
    add(seg_row(0, 0.0, 1), v0)
    add(seg_row(K - 1, 1.0, 1), v1)
    for i in range(K - 1):
        add(seg_row(i, 1.0, 1, dts[i + 1]) - seg_row(i + 1, 0.0, 1, dts[i]), np.zeros(3))
    add(seg_row(0, 0.0, 2), a0)
    add(seg_row(K - 1, 1.0, 2), a1)
    for i in range(K - 1):
        add(seg_row(i, 1.0, 2, dts[i + 1] ** 2) - seg_row(i + 1, 0.0, 2, dts[i] ** 2), np.zeros(3))
    for order in (3, 4):
        if cont >= order:
            for i in range(K - 1):
                add(seg_row(i, 1.0, order, dts[i + 1] ** order) - seg_row(i + 1, 0.0, order, dts[i] ** order),
                    np.zeros(3))
    return np.array(rows), np.array(rhs)


def minsnap_solve(path, desired_vel, v0=(0, 0, 0), v1=(0, 0, 0), a0=(0, 0, 0), a1=(0, 0, 0), deg=7, diff=4, cont=4):
    """Exact KKT solution of the reference QP.  -> (coef [3, K*(deg+1)] in real time, times)."""
    path = np.asarray(path, float)
    K = len(path) - 1
    n = deg + 1
    seglen = np.linalg.norm(np.diff(path, axis=0), axis=1)
    times = np.concatenate([[0.0], np.cumsum(seglen / desired_vel)])
    P = minsnap_P(K, deg, diff)
    A, b = minsnap_Ab(path, times, v0, v1, a0, a1, deg, cont)
    m = len(A)
    KKT = np.block([[P, A.T], [A, np.zeros((m, m))]])
    coef = np.zeros((3, K * n))
    for ax in range(3):
        sol = np.linalg.lstsq(KKT, np.concatenate([np.zeros(K * n), b[:, ax]]), rcond=None)[0]
        c = sol[:K * n].copy()
        for s in range(K):
            dt = times[s + 1] - times[s]
            for d in range(n):
                c[s * n + d] /= dt ** d
        coef[ax] = c
    return coef, times


def poly_pos(coef, times, t, deg=7):
    n = deg + 1
    for i in range(len(times) - 1):
        if times[i] <= t <= times[i + 1]:
            tt = t - times[i]
            p = np.zeros(3)
            for d in range(n):
                p += coef[:, i * n + d] * tt ** d
            return p
    return np.zeros(3)


def poly_trajectory(coef, times, dt):
    out = []
    t = 0.0
    while t <= times[-1]:
        out.append(poly_pos(coef, times, t))
        t += dt
    return np.array(out)


def adjust_path_length_direct(path, omap, max_path_length, prev_path_length=0.0):
    path = [np.asarray(p, float) for p in path]
    out = []
    exceed = False
    min_len = 0.0
    p_start = path[0]
    for i in range(len(path) - 1):
        p1, p2 = path[i], path[i + 1]
        total = np.linalg.norm(p2 - p_start)
        if total >= max(prev_path_length, max_path_length):
            exceed = True
        out.append(p1)
        if exceed:
            free = not bool(omap.query_lines(p1, p2)[0])
            if free and min_len >= 1.5:
                out.append(p2)
                return np.array(out)
        if bool(omap.query_lines(p1, p2)[0]):
            min_len = 0.0
        else:
            min_len += np.linalg.norm(p2 - p1)
    out.append(path[-1])
    return np.array(out)


def input_path_check(path, omap, ctrl_pt_dist=0.25, max_path_length=20.0):
    if len(path) == 0:
        return True, path
    adj = adjust_path_length_direct(path, omap, max_path_length)
    d = np.linalg.norm(np.diff(adj, axis=0), axis=1)
    if np.any(d > ctrl_pt_dist * 1.5):
        return False, None
    keep = [adj[0]]
    prev = adj[0]
    for p in adj[1:]:
        if np.linalg.norm(p - prev) >= ctrl_pt_dist * 0.8:
            keep.append(p)
            prev = p
    keep.append(keep[-1])
    return True, np.array(keep)


def fill_path(path):
    path = np.asarray(path, float)
    if len(path) <= 1:
        return None
    if len(path) == 2:
        ps, pf = path
        return np.array([ps, (pf - ps) / 3.0 + ps, 2.0 * (pf - ps) / 3.0 + ps, pf])
    if len(path) == 3:
        ps, pm, pf = path
        return np.array([ps, (ps + pm) / 2.0, pm, (pm + pf) / 2.0, pf])
    return path


def parameterize_to_bspline(ts, points, start_end):
    points = np.asarray(points, float)
    K = len(points)
    A = np.zeros((K + 4, K + 2))
    for i in range(K):
        A[i, i:i + 3] = np.array([1, 4, 1]) / 6.0
    A[K, 0:3] = np.array([-1, 0, 1]) / 2.0 / ts
    A[K + 1, K - 1:K + 2] = np.array([-1, 0, 1]) / 2.0 / ts
    A[K + 2, 0:3] = np.array([1, -2, 1]) / ts / ts
    A[K + 3, K - 1:K + 2] = np.array([1, -2, 1]) / ts / ts
    b = np.concatenate([points, np.asarray(start_end, float).reshape(4, 3)], 0)
    return np.linalg.lstsq(A, b, rcond=None)[0]  # [K+2, 3]


def start_goal_to_ctrl(start, goal, omap, max_vel=2.0, ctrl_pt_dist=0.25, ctrl_pt_ts=0.2, max_path_length=20.0,
                       max_tries=60):
    """-> control points [N,3] or None (goal occupied).  src/bspline_node.cpp:332-371."""
    start = np.asarray(start, float)
    goal = np.asarray(goal, float)
    coef, times = minsnap_solve([start, goal], max_vel)
    dt = ctrl_pt_dist / max_vel
    adjusted = None
    for _ in range(max_tries):  # the node's 0.05 s wall-clock loop, made a bounded loop
        traj = poly_trajectory(coef, times, dt)
        ok, adjusted_try = input_path_check(traj, omap, ctrl_pt_dist, max_path_length)
        if adjusted_try is not None:
            adjusted = adjusted_try
        if ok:
            break
        dt *= 0.8
    if adjusted is None:
        return None
    if omap.query(adjusted[-1])[0]:
        return None
    inp = adjust_path_length_direct(adjusted, omap, max_path_length)
    if len(inp) < 4:
        inp = fill_path(adjusted)
        if inp is None:
            return None
    zeros = np.zeros((4, 3))
    return parameterize_to_bspline(ctrl_pt_ts, inp, zeros)
