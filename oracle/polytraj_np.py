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
