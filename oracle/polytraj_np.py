"""ORACLE — TEST INFRASTRUCTURE ONLY.

numpy restatement of the reference's min-snap "solve -> sample -> collision check -> insert waypoint" path
(polyTrajOctomap::makePlanAddingWaypoint and what it calls).  Follows:
  polyTrajSolver.cpp:125-138 (avgTimeAllocation), :241-271 (P), :314-584 (A), :587-813 (bounds), :870-879
  (de-normalisation), :1026-1049 (getPose: position = sum_d c_d * pow(t, d)), :1125-1137 (getTrajectory:
  t = 0; t < T; t += delT, then the last waypoint appended);
  polyTrajOctomap.cpp:178-186 (insertWaypoint, highest segment first), :259-321 (loop), :547-569 (checkCollision:
  a box of (xNum+1)(yNum+1)(zNum+1) points, xNum = int((xmax-xmin)/mapRes) from FLOAT point coordinates),
  :571-590 (checkCollisionPoint: outside the metric bounding box / unknown / occupied), :634-656
  (checkCollisionTraj: colliding sample -> first segment with knot[i] <= t <= knot[i+1], t accumulated).
The QP is solved exactly through its KKT system (frontend_np.minsnap_solve) instead of OSQP's 1e-3 ADMM answer
(SURVEY.md §8c).  Reference defect made explicit: the reference never refreshes the solver's path after
insertWaypoint (polyTrajOctomap.cpp:287-289 commented out), so its loop re-solves the ORIGINAL path until
maxIter; this restatement implements the evident intent (re-updatePath after every insertion).
The octree (external octomap, unpinned) is replaced by this repo's grid contract: cell = floor((p - origin)/res),
unknown = known bit clear or outside the grid.
"""
import numpy as np

from . import frontend_np


def solve(path, desired_vel=1.0, cont=4):
    return frontend_np.minsnap_solve(path, desired_vel, cont=cont)


def get_trajectory(coef, times, last_waypoint, delT):
    """polyTrajSolver::getTrajectory -> (positions [S,3], t values [S-1] of the polynomial samples)."""
    out, ts = [], []
    t = 0.0
    end = times[-1]
    n = coef.shape[1] // (len(times) - 1)
    while t < end:
        p = np.zeros(3)
        for i in range(len(times) - 1):
            if times[i] <= t <= times[i + 1]:
                tt = t - times[i]
                for d in range(n):
                    p += coef[:, i * n + d] * np.power(tt, d)   # x += c * pow(t, d), d ascending
                break
        out.append(p)
        ts.append(t)
        t += delT
    out.append(np.asarray(last_waypoint, float))
    return np.array(out), np.array(ts)


def corridor_rows(path, times, corridor_size, corridor_res, deg=7):
    """Corridor inequality rows of polyTrajSolver (updateCorridorParam :985-1012, constructA :555-579, constructBound
    :813-840): per segment i with size r_i != 0, samples t = 0; t <= 1; t += 1/ceil(duration_i * corridorRes) (accumulated,
    normalised segment time), row = [t^d] on the segment's coefficients, bounds pMid(t) -+ r_i with pMid the linear
    interpolation of the segment's waypoints.  -> (A [m, K*(deg+1)], lower [m, 3], upper [m, 3]); rows in (segment,
    ascending t) order — the reference iterates an unordered_map, the QP does not depend on the row order."""
    path = np.asarray(path, float)
    K = len(path) - 1
    n = deg + 1
    rows, lo, hi = [], [], []
    for i in range(K):
        r = float(corridor_size[i])
        if r == 0.0:
            continue
        duration = times[i + 1] - times[i]
        num = int(np.ceil(duration * corridor_res))
        dt = 1.0 / num
        t = 0.0
        while t <= 1.0:
            row = np.zeros(K * n)
            row[i * n:(i + 1) * n] = [np.power(t, d) for d in range(n)]
            mid = path[i] + (path[i + 1] - path[i]) * (t - 0.0) / (1.0 - 0.0)
            rows.append(row)
            lo.append(mid - r)
            hi.append(mid + r)
            t += dt
    return np.array(rows).reshape(-1, K * n), np.array(lo).reshape(-1, 3), np.array(hi).reshape(-1, 3)


IPM_MAX_IT = 60
IPM_MU_TOL = 1e-10
IPM_RES_TOL = 1e-6


def corridor_qp_ipm(P, Aeq, b, Ac, lo, hi):
    """min 1/2 c'Pc  s.t.  Aeq c = b,  lo <= Ac c <= hi  — the QP polyTrajSolver hands to OSQP when corridor constraints
    are set (polyTrajSolver.cpp:162-223, 555-579, 813-840) — solved to convergence by a Mehrotra predictor-corrector
    interior-point method (the reference's OSQP stops at eps 1e-3).  The SAME algorithm, start point and update formulas
    as the device kernel (trajectory_planner_b200/csrc/tp_corridor.cuh), with dense numpy solves:
      slacks su = hi - Ac c, sl = Ac c - lo, multipliers lu, ll > 0;  start: the equality-only minimiser,
      su / sl = max(slack, 0.01 (hi - lo)), lu = ll = 1;  Newton system (P + Ac' S Ac) dc + Aeq' dy = -rd - Ac' g,
      Aeq dc = -rp with S = lu/su + ll/sl;  step lengths 0.995 of the distance to the boundary, primal and dual separately;
      stop when mu < 1e-10 and the residuals are below 1e-6 (1 + |Pc|_inf); 60 iterations without that -> status 1
      (infeasible or not converged).
    -> (c, y, mu_ineq = lu - ll, status, iterations)."""
    n, m, mc = P.shape[0], Aeq.shape[0], len(lo)
    K0 = np.block([[P, Aeq.T], [Aeq, np.zeros((m, m))]])
    sol = np.linalg.solve(K0, np.concatenate([np.zeros(n), b]))
    c, y = sol[:n].copy(), sol[n:].copy()
    if mc == 0:
        return c, y, np.zeros(0), 0, 0
    w = Ac @ c
    su = np.maximum(hi - w, 1e-2 * (hi - lo))
    sl = np.maximum(w - lo, 1e-2 * (hi - lo))
    lu, ll = np.ones(mc), np.ones(mc)

    def maxstep(v, dv):
        neg = dv < 0
        return min(1.0, float(np.min(-v[neg] / dv[neg]))) if np.any(neg) else 1.0

    for it in range(IPM_MAX_IT):
        w = Ac @ c
        Pc = P @ c
        rd = Pc + Aeq.T @ y + Ac.T @ (lu - ll)
        rp = Aeq @ c - b
        ru = w + su - hi
        rl = w - sl - lo
        mu = (su @ lu + sl @ ll) / (2 * mc)
        res = max(np.max(np.abs(rd)), np.max(np.abs(rp)), np.max(np.abs(ru)), np.max(np.abs(rl)))
        if mu < IPM_MU_TOL and res < IPM_RES_TOL * (1 + np.max(np.abs(Pc))):
            return c, y, lu - ll, 0, it
        Sig = lu / su + ll / sl
        Kmat = np.block([[P + Ac.T @ (Sig[:, None] * Ac), Aeq.T], [Aeq, np.zeros((m, m))]])

        def step(rcu, rcl):
            g = (-rcu + lu * ru) / su + (rcl + ll * rl) / sl
            rhs = np.concatenate([-rd - Ac.T @ g, -rp])
            try:
                s = np.linalg.solve(Kmat, rhs)
            except np.linalg.LinAlgError:   # barrier terms of 1e19 next to entries of order 1: numerically singular in the last
                s = np.linalg.lstsq(Kmat, rhs, rcond=None)[0]   # iterations (the device guards the vanished pivot instead)
            dc, dy = s[:n], s[n:]
            Adc = Ac @ dc
            dsu, dsl = -ru - Adc, rl + Adc
            return dc, dy, dsu, dsl, (-rcu - lu * dsu) / su, (-rcl - ll * dsl) / sl

        dc, dy, dsu, dsl, dlu, dll = step(su * lu, sl * ll)
        ap = min(maxstep(su, dsu), maxstep(sl, dsl))
        ad = min(maxstep(lu, dlu), maxstep(ll, dll))
        mu_aff = ((su + ap * dsu) @ (lu + ad * dlu) + (sl + ap * dsl) @ (ll + ad * dll)) / (2 * mc)
        sigma = (mu_aff / mu) ** 3
        dc, dy, dsu, dsl, dlu, dll = step(su * lu + dsu * dlu - sigma * mu, sl * ll + dsl * dll - sigma * mu)
        ap = min(0.995 * min(maxstep(su, dsu), maxstep(sl, dsl)), 1.0)
        ad = min(0.995 * min(maxstep(lu, dlu), maxstep(ll, dll)), 1.0)
        c = c + ap * dc
        su = su + ap * dsu
        sl = sl + ap * dsl
        y = y + ad * dy
        lu = lu + ad * dlu
        ll = ll + ad * dll
        if not (np.all(np.isfinite(c)) and mu < 1e30):
            break
    return c, y, lu - ll, 1, IPM_MAX_IT


def corridor_solve(path, corridor_size, corridor_res=8.0, desired_vel=1.0, cont=4, bc=None):
    """polyTrajSolver::solve with setCorridorConstraint(corridorSizeVec, corridorRes) -> (coef [3, 8K] real time, times,
    status[3], per-axis dict(c, y, mu) in normalised time for the optimality certificate)."""
    path = np.asarray(path, float)
    K = len(path) - 1
    bc = np.zeros((4, 3)) if bc is None else np.asarray(bc, float).reshape(4, 3)
    seglen = np.linalg.norm(np.diff(path, axis=0), axis=1)
    times = np.concatenate([[0.0], np.cumsum(seglen / desired_vel)])
    P = frontend_np.minsnap_P(K)
    Aeq, b = frontend_np.minsnap_Ab(path, times, bc[0], bc[1], bc[2], bc[3], cont=cont)
    Ac, lo, hi = corridor_rows(path, times, corridor_size, corridor_res)
    coef = np.zeros((3, 8 * K))
    status = np.zeros(3, int)
    raw = []
    for ax in range(3):
        c, y, mu, st, _ = corridor_qp_ipm(P, Aeq, b[:, ax], Ac, lo[:, ax], hi[:, ax])
        status[ax] = st
        raw.append(dict(c=c, y=y, mu=mu))
        cc = c.copy()
        for s in range(K):
            cc[s * 8:(s + 1) * 8] /= (times[s + 1] - times[s]) ** np.arange(8)
        coef[ax] = cc
    return coef, times, status, raw


def check_collision_traj_occmap(omap, traj, times, delT):
    """polyTrajOccMap::checkCollisionTraj (polyTrajOccMap.cpp:523-546) on the oracle's occMap: a sample collides when it is
    inflated-occupied AND unknown (the reference's conjunction)."""
    traj = np.asarray(traj, float)
    flags = (omap.query(traj) != 0) & (omap.query_unknown(traj) != 0)
    seg = set()
    t = 0.0
    for c in flags:
        if c:
            for i in range(len(times) - 1):
                if times[i] <= t <= times[i + 1]:
                    seg.add(i)
                    break
        t += delT
    return bool(flags.any()), sorted(seg), flags


def make_plan_corridor(path, grid, init_r=0.5, fs=0.8, corridor_res=8.0, desired_vel=1.0, delT=0.1, box=(0.4, 0.4, 0.2),
                       map_res=0.2, max_iter=100, cont=4, occmap=None, bc=None):
    """polyTrajOctomap::makePlanCorridorConstraint (polyTrajOctomap.cpp:388-459): corridors of radius init_r around every
    path segment, the radius of a colliding segment shrunk by fs, until collision free or maxIter.  An infeasible
    (or unconverged) QP ends the loop with valid = False: the reference keeps its previous solution in that case
    (solveX returns early, polyTrajSolver.cpp:871), which can never become valid.  -> dict(valid, coef, times, iters, r)."""
    path = np.asarray(path, float)
    r = np.full(len(path) - 1, float(init_r))
    it = 0
    valid = False
    coef = times = None
    while not valid:
        c2, times, status, _ = corridor_solve(path, r, corridor_res, desired_vel, cont, bc)
        if np.any(status != 0):
            if coef is None:
                coef = c2
            it += 1
            break
        coef = c2
        traj, _ = get_trajectory(coef, times, path[-1], delT)
        if occmap is not None:   # polyTrajOccMap::makePlan (polyTrajOccMap.cpp:326-399) on the ViGO occupancy map
            has, segs, _ = check_collision_traj_occmap(occmap, traj, times, delT)
        else:
            has, segs, _ = check_collision_traj(grid, traj, times, delT, box, map_res)
        valid = not has
        if not valid:
            for s in segs:
                r[s] *= fs
        it += 1
        if it > max_iter:
            break
    return dict(valid=valid, coef=coef, times=times, iters=it, r=r, status=status)


class Grid3:
    """occupied / known grids + metric bounding box of the known cells (octomap getMetricMin/Max stand-in)."""

    def __init__(self, res, origin, occ, known, bbmin=None, bbmax=None):
        self.res, self.origin, self.occ, self.known = float(res), np.asarray(origin, float), occ, known
        idx = np.argwhere(known != 0)
        self.bbmin = self.origin + idx.min(0) * self.res if bbmin is None else np.asarray(bbmin, float)
        self.bbmax = self.origin + (idx.max(0) + 1) * self.res if bbmax is None else np.asarray(bbmax, float)

    def collision_point(self, p):
        p = np.asarray(p, float)
        if np.any(p < self.bbmin) or np.any(p > self.bbmax):
            return True
        idx = np.floor((p - self.origin) / self.res)
        if np.any(idx < 0) or np.any(idx >= np.array(self.occ.shape)):
            return True
        i, j, k = idx.astype(int)
        if not self.known[i, j, k]:
            return True
        return bool(self.occ[i, j, k])


def check_collision(grid, p, box, map_res):
    """polyTrajOctomap::checkCollision: p -> float (pose2Octomap), corners in double, box points back to float."""
    pf = np.asarray(p, dtype=np.float32).astype(np.float64)
    lo = pf - np.asarray(box, float) / 2
    hi = pf + np.asarray(box, float) / 2
    num = ((hi - lo) / map_res).astype(int)
    for a in range(num[0] + 1):
        for b in range(num[1] + 1):
            for c in range(num[2] + 1):
                q = np.array([lo[0] + a * map_res, lo[1] + b * map_res, lo[2] + c * map_res], dtype=np.float32).astype(np.float64)
                if grid.collision_point(q):
                    return True
    return False


def check_collision_traj(grid, traj, times, delT, box, map_res):
    """-> (has_collision, sorted colliding segments, per-sample flags)."""
    seg = set()
    flags = []
    t = 0.0
    has = False
    for p in traj:
        c = check_collision(grid, p, box, map_res)
        flags.append(c)
        if c:
            has = True
            for i in range(len(times) - 1):
                if times[i] <= t <= times[i + 1]:
                    seg.add(i)
                    break
        t += delT
    return has, sorted(seg), np.array(flags, bool)


def insert_waypoint(path, segs):
    path = [np.asarray(p, float) for p in path]
    for idx in sorted(segs, reverse=True):
        path.insert(idx + 1, (path[idx] + path[idx + 1]) / 2)
    return np.array(path)


def make_plan_adding_waypoint(path, grid, desired_vel=1.0, delT=0.1, box=(0.4, 0.4, 0.2), map_res=0.2, max_iter=100, cont=4,
                              max_waypoints=256):
    """-> dict(valid, path, coef, times, iters)."""
    path = np.asarray(path, float)
    it = 0
    valid = False
    coef = times = None
    while not valid:
        coef, times = solve(path, desired_vel, cont)
        traj, _ = get_trajectory(coef, times, path[-1], delT)
        has, segs, _ = check_collision_traj(grid, traj, times, delT, box, map_res)
        valid = not has
        if not valid:
            if len(path) + len(segs) > max_waypoints:
                break
            path = insert_waypoint(path, segs)
        it += 1
        if it > max_iter:
            break
    return dict(valid=valid, path=path, coef=coef, times=times, iters=it)
