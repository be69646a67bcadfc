// Batched pose-at-time / trajectory sampling: what bsplineTraj::getPose (bsplineTraj.cpp:1402-1419),
// evalTraj (:1438-1447) and evalTrajToMsg (:1502-1518) do for ONE trajectory and one time on the host, for the samples of a
// whole batch in one launch: position = bspline_.at(t), velocity / acceleration = the derivative splines' at(t)
// (bspline.cpp:32-72), yaw = atan2(v_y, v_x).  One thread per sample; a sample reads the 4 control points of its knot
// span through the read-only path (the batch's control points are L2 resident) and forms the derivative splines'
// control points q_i = p (c_(i+1) - c_i) / (u_(i+p+1) - u_(i+1)) on the fly with the expressions of tp_bspline_eval
// (tp_frontend.cpp), so positions / velocities / accelerations are bit-identical to the host entry.
#pragma once
#include "tp_device.cuh"

struct GlobalCP {   // control point i of the cubic spline
  const double* c;
  __device__ __forceinline__ D3 operator()(int i) const { return d3(__ldg(c + 3 * i), __ldg(c + 3 * i + 1), __ldg(c + 3 * i + 2)); }
};
struct GlobalVelCP {   // control point i of the velocity spline (degree 2)
  const double* c;
  double ts;
  __device__ __forceinline__ D3 operator()(int i) const {
    const double den = (double)(i + 3 + 1 - 3) * ts - (double)(i + 1 - 3) * ts;
    const GlobalCP g{c};
    const D3 a = g(i), b = g(i + 1);
    return d3((3.0 * (b.x - a.x)) / den, (3.0 * (b.y - a.y)) / den, (3.0 * (b.z - a.z)) / den);
  }
};
struct GlobalAccCP {   // control point i of the acceleration spline (degree 1)
  const double* c;
  double ts;
  __device__ __forceinline__ D3 operator()(int i) const {
    const double den = (double)(i + 2 + 1 - 2) * ts - (double)(i + 1 - 2) * ts;
    const GlobalVelCP g{c, ts};
    const D3 a = g(i), b = g(i + 1);
    return d3((2.0 * (b.x - a.x)) / den, (2.0 * (b.y - a.y)) / den, (2.0 * (b.z - a.z)) / den);
  }
};

// samples of trajectory b: t_offsets[b] .. t_offsets[b+1]-1 (ragged); outputs indexed by sample (pos / vel / acc: 3 doubles)
__global__ void __launch_bounds__(256) k_sample_traj(int B, const int* __restrict__ offsets, const double* __restrict__ ctrl,
                                                     const int* __restrict__ t_offsets, const double* __restrict__ t, double ts,
                                                     double* __restrict__ pos, double* __restrict__ vel, double* __restrict__ acc,
                                                     double* __restrict__ yaw) {
  const long total = t_offsets[B];
  for (long s = blockIdx.x * (long)blockDim.x + threadIdx.x; s < total; s += (long)gridDim.x * blockDim.x) {
    int lo = 0, hi = B;   // the trajectory that owns sample s: the last b with t_offsets[b] <= s
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(t_offsets + mid) <= s) lo = mid; else hi = mid;
    }
    const int o = __ldg(offsets + lo), N = __ldg(offsets + lo + 1) - o;
    const double* c = ctrl + 3 * (size_t)o;
    const double tt = t[s];
    if (N < 4) continue;
    const D3 p = bspline_at(GlobalCP{c}, N, 3, ts, tt);
    pos[3 * s] = p.x; pos[3 * s + 1] = p.y; pos[3 * s + 2] = p.z;
    if (vel || yaw) {
      const D3 v = bspline_at(GlobalVelCP{c, ts}, N - 1, 2, ts, tt);
      if (vel) { vel[3 * s] = v.x; vel[3 * s + 1] = v.y; vel[3 * s + 2] = v.z; }
      if (yaw) yaw[s] = tp_atan2(v.y, v.x);
    }
    if (acc) {
      const D3 a = bspline_at(GlobalAccCP{c, ts}, N - 2, 1, ts, tt);
      acc[3 * s] = a.x; acc[3 * s + 1] = a.y; acc[3 * s + 2] = a.z;
    }
  }
}
