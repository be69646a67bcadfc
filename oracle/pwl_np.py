"""ORACLE — TEST INFRASTRUCTURE ONLY.

Python restatement of the reference's piecewise-linear fallback trajectory pwlTraj (piecewiseLinearTraj.cpp) and of
polyTrajSolver::getPose (polyTrajSolver.cpp:1026-1049).  Follows:
  piecewiseLinearTraj.cpp:31-46 (updatePath: heading of each segment, the last point repeats the previous heading),
  :82-119 (avgTimeAllocation: rotation period then forward period per waypoint, desiredVel 1.0 / desiredAngularVel 0.5,
  piecewiseLinearTraj.h:20-21), :171-196 (makePlan: t = 0; t < T; t += delT, then the pose at T), :199-268 (getPose),
  utils.h:19 (PI_const = 3.1415926), :43-52 (quaternion_from_rpy wraps yaw > PI_const by 2 PI_const), :69-82 (distances).
The yaw the reference carries through a tf2 quaternion round trip is carried as the angle itself.
"""
import math

import numpy as np

PI_CONST = 3.1415926


def yaw_distance(y1, y2):
    d = abs(y2 - y1)
    if d > PI_CONST:
        d = 2 * PI_CONST - d
    return d


def _wrap(yaw):
    return yaw - 2 * PI_CONST if yaw > PI_CONST else yaw


def plan(path, yaw=None, desired_vel=1.0, desired_ang_vel=0.5):
    """-> (yaw[K], knots).  yaw=None: useYaw = false."""
    path = np.asarray(path, float)
    K = len(path)
    use_yaw = yaw is not None
    if use_yaw:
        yw = [float(v) for v in yaw]
    else:
        yw = [0.0] * K
        y = 0.0
        for i in range(K - 1):
            y = math.atan2(path[i + 1][1] - path[i][1], path[i + 1][0] - path[i][0])
            yw[i] = y
        yw[K - 1] = y
    total = 0.0
    knots = []
    for i in range(K - 1):
        if i != 0:
            total += yaw_distance(yw[i - 1], yw[i]) / desired_ang_vel
        else:
            total += 0.0
        knots.append(total)
        d = math.sqrt(math.pow(path[i][0] - path[i + 1][0], 2) + math.pow(path[i][1] - path[i + 1][1], 2) + math.pow(path[i][2] - path[i + 1][2], 2))
        total += d / desired_vel
        knots.append(total)
    if use_yaw:
        total += yaw_distance(yw[K - 2], yw[K - 1]) / desired_ang_vel
        knots.append(total)
    return np.array(yw), np.array(knots)


def get_pose(path, yaw, knots, t):
    path = np.asarray(path, float)
    K = len(path)
    if t >= knots[-1]:
        return np.array([path[K - 1][0], path[K - 1][1], path[K - 1][2], _wrap(yaw[K - 1])])
    out = np.zeros(4)
    for i in range(len(knots) - 1):
        t0, t1 = knots[i], knots[i + 1]
        if t0 <= t <= t1:
            if i % 2 == 1:
                pi = (i - 1) // 2
                yd = yaw[pi + 1] - yaw[pi]
                direction, yda = 1.0, abs(yd)
                if yda <= PI_CONST and yd >= 0:
                    direction = 1.0
                elif yda <= PI_CONST and yd < 0:
                    direction = -1.0
                elif yda > PI_CONST and yd >= 0:
                    direction = -1.0
                    yda = 2 * PI_CONST - yda
                elif yda > PI_CONST and yd < 0:
                    direction = 1.0
                    yda = 2 * PI_CONST - yda
                out[:3] = path[pi + 1]
                out[3] = _wrap(yaw[pi] + direction * (t - t0) / (t1 - t0) * yda)
            else:
                pi = i // 2
                a, b = path[pi], path[pi + 1]
                if t1 - t0 < 1e-3:
                    out[:3] = a
                else:
                    for k in range(3):
                        out[k] = a[k] + (t - t0) * (b[k] - a[k]) / (t1 - t0)
                out[3] = _wrap(yaw[pi])
            break
    return out


def make_plan(path, delT, yaw=None):
    yw, knots = plan(path, yaw)
    traj = []
    t = 0.0
    while t < knots[-1]:
        traj.append(get_pose(path, yw, knots, t))
        t += delT
    traj.append(get_pose(path, yw, knots, knots[-1]))
    return np.array(traj), yw, knots


def poly_get_pose(coef, times, t):
    """polyTrajSolver::getPose: coef [3, 8K] real-time coefficients -> (x, y, z, yaw)."""
    K = len(times) - 1
    for i in range(K):
        if times[i] <= t <= times[i + 1]:
            tt = t - times[i]
            x = y = z = 0.0
            for d in range(8):
                pw = math.pow(tt, d)
                x += coef[0][8 * i + d] * pw
                y += coef[1][8 * i + d] * pw
                z += coef[2][8 * i + d] * pw
            if tt == 0:
                tt = 0.01
            dx = dy = 0.0
            for d in range(8):
                pw = math.pow(tt, d - 1)
                dx += d * coef[0][8 * i + d] * pw
                dy += d * coef[1][8 * i + d] * pw
            return np.array([x, y, z, math.atan2(dy, dx)])
    return np.zeros(4)
