// Fused ViGO cost/gradient evaluation + L-BFGS (More-Thuente) iterate, one thread block per
// trajectory, all solver state in shared memory.
//
// Replaces, on the GPU: bsplineTraj::optimize / costFunction / getDistanceCost / getSmoothnessCost /
// getFeasibilityCost / getDynamicObstacleCost (bsplineTraj.cpp:687-718, 802-1064) and
// lbfgs::lbfgs_optimize / line_search_morethuente / update_trial_interval
// (solver/lbfgs.hpp:1024-1349, 716-937, 506-714).
//
// Design (B200): n = 3(N-6) <= ~300 FP64 unknowns per problem -> a latency-bound, FP64-pipe
// workload with no dense contraction (no tensor cores).  Every vector element e is OWNED by one
// thread (e = tid + k*THREADS) for the whole solve, so x/g/d/xp/gp and the 2m history rows need
// no barriers between vector updates; barriers are only paid for (a) the cost evaluation's
// neighbour reads of control points and (b) one per scalar reduction.  The gradient is computed
// in GATHER form (each element sums its <= 4 jerk + 2 velocity + 3 acceleration stencil terms in
// the reference's accumulation order) so it is atomic-free and bit-reproduces the CPU's scatter
// loops.  Line-search scalars are block-uniform and recomputed redundantly in registers.
// Reductions: fixed xor-shuffle tree + fixed-order cross-warp sum (deterministic); with
// `strict_order` thread 0 redoes every sum serially in the reference's order, which reproduces
// the CPU iterate bit for bit (parity/debug mode).
#pragma once
#include "tp_vigo.cuh"

struct EvalCtx {
  int N, n;
  double* cp;  // shared: 3N control points; the optimised x aliases cp + 9
  const GuidePair* pairs;
  const int* head;  // per control point of this trajectory: first pair or -1
  double w_dist, w_dyn;
  int n_dyn;
  const double* dyn_pos;
  const double* dyn_vel;
  const double* dyn_size;
};

struct Red {
  double* buf;  // shared [2][TP_LB_WARPS][4]
  int flip;
};

// ---------------------------------------------------------------------------- reductions
template <int NV>
__device__ __forceinline__ void block_sum(Red& R, double (&v)[NV], int tid) {
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  double* b = R.buf + R.flip * (TP_LB_WARPS * 4);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) b[warp * 4 + i] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double s = b[i];
#pragma unroll
    for (int w = 1; w < TP_LB_WARPS; ++w) s += b[w * 4 + i];
    v[i] = s;
  }
  R.flip ^= 1;
}

// thread-0 result broadcast (strict mode)
template <int NV>
__device__ __forceinline__ void block_bcast(Red& R, double (&v)[NV], int tid) {
  double* b = R.buf + R.flip * (TP_LB_WARPS * 4);
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) b[i] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = b[i];
  R.flip ^= 1;
}

// serial dot, lbfgs.hpp:453-461
__device__ __forceinline__ double serial_dot(const double* x, const double* y, int n) {
  double s = 0.;
  for (int i = 0; i < n; ++i) s += x[i] * y[i];
  return s;
}

// up to 3 dot products in one reduction
template <bool STRICT, int NV>
__device__ __forceinline__ void block_dots(Red& R, const double* const (&A)[NV], const double* const (&Bv)[NV], int n,
                                           double (&out)[NV], int tid) {
  if (STRICT) {
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int i = 0; i < NV; ++i) out[i] = serial_dot(A[i], Bv[i], n);
    }
    block_bcast<NV>(R, out, tid);
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) out[i] = 0.0;
    for (int e = tid; e < n; e += TP_LB_THREADS) {
#pragma unroll
      for (int i = 0; i < NV; ++i) out[i] += A[i][e] * Bv[i][e];
    }
    block_sum<NV>(R, out, tid);
  }
}

// ---------------------------------------------------------------------------- cost pieces
// one (control point c, axis a) slice of getDistanceCost's pair loop, bsplineTraj.cpp:839-895
__device__ __forceinline__ void distance_terms(const VigoConst& C, const EvalCtx& E, int c, int a, double& grad,
                                               double& cost) {
  const double dth = C.p.dthresh;
  const double cx = E.cp[3 * c], cy = E.cp[3 * c + 1], cz = E.cp[3 * c + 2];
  for (int gi = E.head[c]; gi >= 0;) {
    const GuidePair& pr = E.pairs[gi];
    const double vx = pr.v[0], vy = pr.v[1], vz = pr.v[2];
    const double dist = ((cx - pr.p[0]) * vx + (cy - pr.p[1]) * vy) + (cz - pr.p[2]) * vz;
    const double distErr = dth - dist;
    const double va = a == 0 ? vx : (a == 1 ? vy : vz);
    const bool unk = pr.unknown != 0;
    double costTemp, gt;
    bool active = true;
    if (distErr <= -1.0 * dth) {
      costTemp = cube_cr(-distErr);
      gt = (3.0 * ((-distErr) * (-distErr))) * va;
    } else if (distErr > 0 && distErr <= dth) {
      costTemp = cube_cr(distErr);
      gt = (-3.0 * (distErr * distErr)) * va;
      if (unk) {
        costTemp *= C.p.uncertain_factor;
        gt *= C.p.uncertain_factor;
      }
    } else if (distErr >= dth) {
      costTemp = C.dist_a * (distErr * distErr) + C.dist_b * distErr + C.dist_c;
      gt = (-(2 * C.dist_a * distErr + C.dist_b)) * va;
      if (unk) {
        costTemp *= C.p.uncertain_factor;
        gt *= C.p.uncertain_factor;
      }
    } else {
      active = false;
      costTemp = 0;
      gt = 0;
    }
    if (active) {
      if (!C.p.plan_in_z && a == 2) gt = 0.0;
      grad += gt;
      cost += costTemp;
    }
    gi = pr.next;
  }
  if (C.p.plan_in_z) {
    // height barrier; its gradient lands on the X row (reference quirk, bsplineTraj.cpp:897-930)
    const double hth = 0.2;
    const double hmin = cz - C.p.min_height, hmax = cz - C.p.max_height;
    const double ua = a == 0 ? 1.0 : 0.0;
    if (hmin < 0) {
      const double e = hth - hmin;
      cost += C.h_a * (e * e) + C.h_b * e + C.h_c;
      grad += (-(2 * C.h_a * e + C.h_b)) * -ua;
    } else if (hmin >= 0 && hmax < hth) {
      const double e = hth - hmin;
      cost += cube_cr(e);
      grad += (-3.0 * (e * e)) * -ua;
    }
    if (hmax > 0) {
      const double e = hth + hmax;
      cost += C.h_a * (e * e) + C.h_b * e + C.h_c;
      grad += (-(2 * C.h_a * e + C.h_b)) * ua;
    } else if (hmax <= 0 && hmax >= -hth) {
      const double e = hth + hmax;
      cost += cube_cr(e);
      grad += (-3.0 * (e * e)) * ua;
    }
  }
}

// (c, a) slice of getDynamicObstacleCost, bsplineTraj.cpp:1010-1040
__device__ __forceinline__ void dynamic_terms(const VigoConst& C, const EvalCtx& E, int c, int a, double& grad,
                                              double& cost) {
  const double dd = C.p.dthresh_dyn;
  const double cx = E.cp[3 * c], cy = E.cp[3 * c + 1];
  for (int j = 0; j < E.n_dyn; ++j) {
    const double hx = E.dyn_size[3 * j] / 2, hy = E.dyn_size[3 * j + 1] / 2;
    const double size = sqrt(hx * hx + hy * hy);
    for (int n = 0; n <= C.pred_num; n += 2) {
      const double tt = (double)n * C.p.ts;
      const double ox = E.dyn_pos[3 * j] + tt * E.dyn_vel[3 * j];
      const double oy = E.dyn_pos[3 * j + 1] + tt * E.dyn_vel[3 * j + 1];
      const double distThresh = (1 - (double)(C.pred_num != 0 ? n / C.pred_num : 0) * 0.2) * dd;
      const double dx = cx - ox, dy = cy - oy;
      const double dn = sqrt((dx * dx + dy * dy) + 0.0 * 0.0);
      const double dist = dn - size;
      const double distErr = distThresh - dist;
      const double ga = a == 0 ? dx / dn : (a == 1 ? dy / dn : 0.0 / dn);
      if (distErr <= 0) {
      } else if (distErr > 0 && distErr <= distThresh) {
        cost += cube_cr(distErr);
        grad += (-3.0 * (distErr * distErr)) * ga;
      } else if (distErr >= distThresh) {
        cost += (C.dyn_a * (distErr * distErr) + C.dyn_b * distErr + C.dyn_c);
        grad += (-(2 * C.dyn_a * distErr + C.dyn_b)) * ga;
      }
    }
  }
}

// Gradient of the optimised elements owned by this thread (gather form) + this thread's share
// of the weighted cost.  g (shared, n) receives costFunction's gradient (bsplineTraj.cpp:817-819).
// `d` (may be null): search direction; `dg_part` accumulates this thread's share of g.d over the
// elements it has just WRITTEN (g[i] is produced by thread (i+9) mod THREADS here, not by the thread
// that owns index i elsewhere, so the dot must not re-read g before the next barrier).
__device__ __forceinline__ double eval_partial(const VigoConst& C, const EvalCtx& E, double* g, const double* d,
                                               double& dg_part, int tid) {
  const int N = E.N;
  const double* cp = E.cp;
  const double cts = C.p.ctrl_pt_ts;
  const double k2 = C.ts_inv_sqr;
  double sD = 0, sS = 0, sF = 0, sO = 0;
  for (int e = tid; e < 3 * N; e += TP_LB_THREADS) {
    const int c = e / 3, a = e - 3 * c;
    // ---- this element's own cost terms (every (i,a) term is owned by exactly one thread)
    if (c <= N - 4) {  // jerk_i, i = c  (getSmoothnessCost, :938-940)
      const double j = cp[e + 9] - 3 * cp[e + 6] + 3 * cp[e + 3] - cp[e];
      sS += j * j;
    }
    if (c <= N - 2) {  // v_i (getFeasibilityCost, :961-975)
      const double v = (cp[e + 3] - cp[e]) / cts;
      if (v > 1.0) sF += ((v - 1.0) * (v - 1.0)) * k2;
      else if (v < -1.0) sF += ((v + 1.0) * (v + 1.0)) * k2;
    }
    if (c <= N - 3) {  // a_i (:979-995)
      const double ac = (cp[e + 6] - 2 * cp[e + 3] + cp[e]) * k2;
      if (ac > 1.0) sF += (ac - 1.0) * (ac - 1.0);
      else if (ac < -1.0) sF += (ac + 1.0) * (ac + 1.0);
    }
    if (c < TP_DEGREE || c > N - TP_DEGREE - 1) continue;
    // ---- gradient of an optimised element: gather in the reference's accumulation order
    double gs = 0.0;
    {
      const double j0 = cp[e] - 3 * cp[e - 3] + 3 * cp[e - 6] - cp[e - 9];      // i = c-3 -> col(i+3) += gt
      const double j1 = cp[e + 3] - 3 * cp[e] + 3 * cp[e - 3] - cp[e - 6];      // i = c-2 -> col(i+2) += -3 gt
      const double j2 = cp[e + 6] - 3 * cp[e + 3] + 3 * cp[e] - cp[e - 3];      // i = c-1 -> col(i+1) += 3 gt
      const double j3 = cp[e + 9] - 3 * cp[e + 6] + 3 * cp[e + 3] - cp[e];      // i = c   -> col(i)   += -gt
      gs += 2.0 * j0;
      gs += -3.0 * (2.0 * j1);
      gs += 3.0 * (2.0 * j2);
      gs += -(2.0 * j3);
    }
    double gf = 0.0;
    {
      const double vm = (cp[e] - cp[e - 3]) / cts;  // i = c-1, writes col(i+1)
      if (vm > 1.0) gf += 2 * (vm - 1.0) / cts * k2;
      else if (vm < -1.0) gf += 2 * (vm + 1.0) / cts * k2;
      const double v0 = (cp[e + 3] - cp[e]) / cts;  // i = c, writes col(i)
      if (v0 > 1.0) gf += -2 * (v0 - 1.0) / cts * k2;
      else if (v0 < -1.0) gf += -2 * (v0 + 1.0) / cts * k2;
      const double a0 = (cp[e] - 2 * cp[e - 3] + cp[e - 6]) * k2;  // i = c-2, col(i+2)
      if (a0 > 1.0) gf += 2 * (a0 - 1.0) * k2;
      else if (a0 < -1.0) gf += 2 * (a0 + 1.0) * k2;
      const double a1 = (cp[e + 3] - 2 * cp[e] + cp[e - 3]) * k2;  // i = c-1, col(i+1)
      if (a1 > 1.0) gf += -4 * (a1 - 1.0) * k2;
      else if (a1 < -1.0) gf += -4 * (a1 + 1.0) * k2;
      const double a2 = (cp[e + 6] - 2 * cp[e + 3] + cp[e]) * k2;  // i = c, col(i)
      if (a2 > 1.0) gf += 2 * (a2 - 1.0) * k2;
      else if (a2 < -1.0) gf += 2 * (a2 + 1.0) * k2;
    }
    double gd = 0.0, cD = 0.0;
    distance_terms(C, E, c, a, gd, cD);
    if (a == 0) sD += cD;
    double go = 0.0;
    if (E.n_dyn > 0) {
      double cO = 0.0;
      dynamic_terms(C, E, c, a, go, cO);
      if (a == 0) sO += cO;
    }
    const double gv = E.w_dist * gd + C.p.w_smooth * gs + C.p.w_feas * gf + E.w_dyn * go;
    g[e - 3 * TP_DEGREE] = gv;
    if (d) dg_part += gv * d[e - 3 * TP_DEGREE];
  }
  return E.w_dist * sD + C.p.w_smooth * sS + C.p.w_feas * sF + E.w_dyn * sO;
}

// thread 0, strict mode: the four cost sums in the reference's serial order
__device__ double serial_cost(const VigoConst& C, const EvalCtx& E) {
  const int N = E.N;
  const double* cp = E.cp;
  double D = 0.0, S = 0.0, F = 0.0, O = 0.0;
  for (int i = TP_DEGREE; i <= N - TP_DEGREE - 1; ++i) {
    double gdummy = 0.0;
    // distance_terms adds the pair costs of control point i in list order, then the height terms
    distance_terms(C, E, i, 0, gdummy, D);
  }
  for (int i = 0; i < N - TP_DEGREE; ++i) {
    const double jx = cp[3 * i + 9] - 3 * cp[3 * i + 6] + 3 * cp[3 * i + 3] - cp[3 * i];
    const double jy = cp[3 * i + 10] - 3 * cp[3 * i + 7] + 3 * cp[3 * i + 4] - cp[3 * i + 1];
    const double jz = cp[3 * i + 11] - 3 * cp[3 * i + 8] + 3 * cp[3 * i + 5] - cp[3 * i + 2];
    S += (jx * jx + jy * jy) + jz * jz;
  }
  const double cts = C.p.ctrl_pt_ts, k2 = C.ts_inv_sqr;
  for (int i = 0; i < N - 1; ++i)
    for (int j = 0; j < 3; ++j) {
      const double v = (cp[3 * (i + 1) + j] - cp[3 * i + j]) / cts;
      if (v > 1.0) F += ((v - 1.0) * (v - 1.0)) * k2;
      else if (v < -1.0) F += ((v + 1.0) * (v + 1.0)) * k2;
    }
  for (int i = 0; i < N - 2; ++i)
    for (int j = 0; j < 3; ++j) {
      const double ac = (cp[3 * (i + 2) + j] - 2 * cp[3 * (i + 1) + j] + cp[3 * i + j]) * k2;
      if (ac > 1.0) F += (ac - 1.0) * (ac - 1.0);
      else if (ac < -1.0) F += (ac + 1.0) * (ac + 1.0);
    }
  if (E.n_dyn > 0)
    for (int i = TP_DEGREE; i <= N - TP_DEGREE - 1; ++i) {
      double gdummy = 0.0;
      dynamic_terms(C, E, i, 0, gdummy, O);
    }
  return E.w_dist * D + C.p.w_smooth * S + C.p.w_feas * F + E.w_dyn * O;
}

// costFunction at the control points currently in E.cp: g <- gradient, returns {f, g.d}
// (d may be null -> second value 0).  Entered with E.cp visible to all threads.
template <bool STRICT>
__device__ __forceinline__ void eval_cost(const VigoConst& C, const EvalCtx& E, Red& R, double* g, const double* d,
                                          double& f, double& dg, int tid) {
  double dgp = 0.0;
  double part = eval_partial(C, E, g, d, dgp, tid);
  double v[2];
  if (STRICT) {
    __syncthreads();
    if (tid == 0) {
      v[0] = serial_cost(C, E);
      v[1] = d ? serial_dot(g, d, E.n) : 0.0;
    }
    block_bcast<2>(R, v, tid);
  } else {
    v[0] = part;
    v[1] = dgp;
    block_sum<2>(R, v, tid);
  }
  f = v[0];
  dg = v[1];
}

// ---------------------------------------------------------------------------- More-Thuente helpers
// minimiser formulas of lbfgs.hpp:308-391 (same expressions, same order)
__device__ __forceinline__ double mt_cubic(double u, double fu, double du, double v, double fv, double dv) {
  const double d = v - u;
  const double theta = (fu - fv) * 3 / d + du + dv;
  double p = fabs(theta), q = fabs(du), r = fabs(dv);
  double s = p >= q ? p : q;
  s = s >= r ? s : r;
  const double a = theta / s;
  double gamm = s * sqrt(a * a - (du / s) * (dv / s));
  if (v < u) gamm = -gamm;
  p = gamm - du + theta;
  q = gamm - du + gamm + dv;
  r = p / q;
  return u + r * d;
}
__device__ __forceinline__ double mt_cubic2(double u, double fu, double du, double v, double fv, double dv,
                                            double xmin, double xmax) {
  const double d = v - u;
  const double theta = (fu - fv) * 3 / d + du + dv;
  double p = fabs(theta), q = fabs(du), r = fabs(dv);
  double s = p >= q ? p : q;
  s = s >= r ? s : r;
  const double a = theta / s;
  double gamm = a * a - (du / s) * (dv / s);
  gamm = gamm > 0 ? s * sqrt(gamm) : 0;
  if (u < v) gamm = -gamm;
  p = gamm - dv + theta;
  q = gamm - dv + gamm + du;
  r = p / q;
  if (r < 0. && gamm != 0.) return v - r * d;
  if (a < 0) return xmax;
  return xmin;
}
__device__ __forceinline__ double mt_quad(double u, double fu, double du, double v, double fv) {
  const double a = v - u;
  return u + du / ((fu - fv) / a + du) / 2 * a;
}
__device__ __forceinline__ double mt_quad2(double u, double du, double v, double dv) {
  const double a = u - v;
  return v + dv / (dv - du) * a;
}

#define LB_OUTOFINTERVAL (-1010)
#define LB_INCORRECT_TMINMAX (-1009)
#define LB_ROUNDING (-1008)
#define LB_MINSTEP (-1007)
#define LB_MAXSTEP (-1006)
#define LB_MAXLINESEARCH (-1005)
#define LB_MAXITER (-1004)
#define LB_WIDTHTOOSMALL (-1003)
#define LB_INVALIDPARAMS (-1002)
#define LB_INCREASEGRADIENT (-1001)
#define LB_INVALID_N (-1021)

// update_trial_interval, lbfgs.hpp:506-714
__device__ __forceinline__ int mt_update(double& x, double& fx, double& dx, double& y, double& fy, double& dy,
                                         double& t, double& ft, double& dt, double tmin, double tmax, int& brackt) {
  int bound;
  const int dsign = dt * (dx / fabs(dx)) < 0.;
  double mc, mq, newt;
  if (brackt) {
    if (t <= (x <= y ? x : y) || (x >= y ? x : y) <= t) return LB_OUTOFINTERVAL;
    if (0. <= dx * (t - x)) return LB_INCREASEGRADIENT;
    if (tmax < tmin) return LB_INCORRECT_TMINMAX;
  }
  if (fx < ft) {
    brackt = 1;
    bound = 1;
    mc = mt_cubic(x, fx, dx, t, ft, dt);
    mq = mt_quad(x, fx, dx, t, ft);
    newt = (fabs(mc - x) < fabs(mq - x)) ? mc : mc + 0.5 * (mq - mc);
  } else if (dsign) {
    brackt = 1;
    bound = 0;
    mc = mt_cubic(x, fx, dx, t, ft, dt);
    mq = mt_quad2(x, dx, t, dt);
    newt = (fabs(mc - t) > fabs(mq - t)) ? mc : mq;
  } else if (fabs(dt) < fabs(dx)) {
    bound = 1;
    mc = mt_cubic2(x, fx, dx, t, ft, dt, tmin, tmax);
    mq = mt_quad2(x, dx, t, dt);
    if (brackt) newt = (fabs(t - mc) < fabs(t - mq)) ? mc : mq;
    else newt = (fabs(t - mc) > fabs(t - mq)) ? mc : mq;
  } else {
    bound = 0;
    if (brackt) newt = mt_cubic(t, ft, dt, y, fy, dy);
    else if (x < t) newt = tmax;
    else newt = tmin;
  }
  if (fx < ft) {
    y = t; fy = ft; dy = dt;
  } else {
    if (dsign) { y = x; fy = fx; dy = dx; }
    x = t; fx = ft; dx = dt;
  }
  if (tmax < newt) newt = tmax;
  if (newt < tmin) newt = tmin;
  if (brackt && bound) {
    mq = x + 0.66 * (y - x);
    if (x < y) { if (mq < newt) newt = mq; }
    else { if (newt < mq) newt = mq; }
  }
  t = newt;
  return 0;
}

// shared-memory footprint (doubles) of one solve
__host__ __device__ inline size_t lbfgs_smem_doubles(int N, int m) {
  const int n = 3 * (N - 2 * TP_DEGREE);
  const int nn = n > 0 ? n : 0;
  return (size_t)3 * N + 4 * (size_t)nn + 2 * (size_t)m * nn + 2 * (size_t)m + 2 * TP_LB_WARPS * 4 + 8;
}

// One optimize() (bsplineTraj.cpp:687-718) on the control points in E.cp (shared).  On return
// E.cp holds the LAST EVALUATED point, as the reference's cost callback leaves
// optData_.controlPoints (:803); xfinal (global, may be null) receives the solver's own x.
template <bool STRICT>
__device__ void lbfgs_run(const VigoConst& C, EvalCtx& E, double* sm, tp_lbfgs_result& out, double* xfinal, int tid) {
  const int n = E.n, m = C.p.lbfgs_m;
  double* x = E.cp + 3 * TP_DEGREE;
  double* g = sm;
  double* xp = g + n;
  double* gp = xp + n;
  double* d = gp + n;
  double* S = d + n;
  double* Y = S + (size_t)m * n;
  double* alpha = Y + (size_t)m * n;
  double* YS = alpha + m;
  Red R;
  R.buf = YS + m;
  R.flip = 0;

  const double min_step = 1e-20, max_step = 1e20, ftol = 1e-4, gtol = 0.9, xtol = 1e-16;  // lbfgs.hpp:942-954
  const int max_ls = C.p.lbfgs_max_linesearch;
  int evals = 0, k = 0, ret, bsum = 0;
  double fx, dgdummy;
  __syncthreads();
  eval_cost<STRICT>(C, E, R, g, nullptr, fx, dgdummy, tid);
  ++evals;
  for (int e = tid; e < n; e += TP_LB_THREADS) d[e] = -g[e];
  double nr[2];
  {
    const double* const A2[2] = {x, g};
    block_dots<STRICT, 2>(R, A2, A2, n, nr, tid);
  }
  double xnorm = sqrt(nr[0]), gnorm = sqrt(nr[1]);
  if (xnorm < 1.0) xnorm = 1.0;
  if (gnorm / xnorm <= C.p.lbfgs_g_eps) {
    ret = 2;  // LBFGS_ALREADY_MINIMIZED
  } else {
    // step = 1/sqrt(d.d): d = -g so d.d == g.g bit for bit
    double step = 1.0 / sqrt(nr[1]);
    int end = 0;
    k = 1;
    for (;;) {
      for (int e = tid; e < n; e += TP_LB_THREADS) {
        xp[e] = x[e];
        gp[e] = g[e];
      }
      // ---------------- line_search_morethuente (lbfgs.hpp:716-937)
      int ls;
      {
        int count = 0, brackt = 0, stage1 = 1, uinfo = 0;
        double dg, stx, fxx, dgx, sty, fy, dgy, finit, ftest1, dginit, dgtest, width, prev_width, stmin, stmax;
        double stp = step;
        if (stp <= 0.) {
          ls = LB_INVALIDPARAMS;
        } else {
          double v1[1];
          {
            const double* const A1[1] = {g};
            const double* const B1[1] = {d};
            block_dots<STRICT, 1>(R, A1, B1, n, v1, tid);
          }
          dginit = v1[0];
          if (0 < dginit) {
            ls = LB_INCREASEGRADIENT;
          } else {
            finit = fx;
            dgtest = ftol * dginit;
            width = max_step - min_step;
            prev_width = 2.0 * width;
            stx = sty = 0.;
            fxx = fy = finit;
            dgx = dgy = dginit;
            for (;;) {
              if (brackt) {
                stmin = stx <= sty ? stx : sty;
                stmax = stx >= sty ? stx : sty;
              } else {
                stmin = stx;
                stmax = stp + 4.0 * (stp - stx);
              }
              if (stp < min_step) stp = min_step;
              if (max_step < stp) stp = max_step;
              if ((brackt && ((stp <= stmin || stmax <= stp) || max_ls <= count + 1 || uinfo != 0)) ||
                  (brackt && (stmax - stmin <= xtol * stmax)))
                stp = stx;
              // x = xp + stp*s  (veccpy + vecadd, lbfgs.hpp:824-825)
              for (int e = tid; e < n; e += TP_LB_THREADS) {
                double t = xp[e];
                t += stp * d[e];
                x[e] = t;
              }
              __syncthreads();
              eval_cost<STRICT>(C, E, R, g, d, fx, dg, tid);
              ++evals;
              ftest1 = finit + stp * dgtest;
              ++count;
              if (brackt && ((stp <= stmin || stmax <= stp) || uinfo != 0)) { ls = LB_ROUNDING; break; }
              if (stp == max_step && fx <= ftest1 && dg <= dgtest) { ls = LB_MAXSTEP; break; }
              if (stp == min_step && (ftest1 < fx || dgtest <= dg)) { ls = LB_MINSTEP; break; }
              if (brackt && (stmax - stmin) <= xtol * stmax) { ls = LB_WIDTHTOOSMALL; break; }
              if (max_ls <= count) { ls = LB_MAXLINESEARCH; break; }
              if (fx <= ftest1 && fabs(dg) <= gtol * (-dginit)) { ls = count; break; }
              if (stage1 && fx <= ftest1 && (ftol <= gtol ? ftol : gtol) * dginit <= dg) stage1 = 0;
              if (stage1 && ftest1 < fx && fx <= fxx) {
                double fm = fx - stp * dgtest;
                double fxm = fxx - stx * dgtest;
                double fym = fy - sty * dgtest;
                double dgm = dg - dgtest;
                double dgxm = dgx - dgtest;
                double dgym = dgy - dgtest;
                uinfo = mt_update(stx, fxm, dgxm, sty, fym, dgym, stp, fm, dgm, stmin, stmax, brackt);
                fxx = fxm + stx * dgtest;
                fy = fym + sty * dgtest;
                dgx = dgxm + dgtest;
                dgy = dgym + dgtest;
              } else {
                double ft = fx, dt = dg;
                uinfo = mt_update(stx, fxx, dgx, sty, fy, dgy, stp, ft, dt, stmin, stmax, brackt);
              }
              if (brackt) {
                if (0.66 * prev_width <= fabs(sty - stx)) stp = stx + 0.5 * (sty - stx);
                prev_width = width;
                width = fabs(sty - stx);
              }
            }
          }
        }
        step = stp;
      }
      if (ls < 0) {
        // revert the SOLVER's x,g (lbfgs.hpp:1189-1197); the control points keep the last trial
        if (xfinal)
          for (int e = tid; e < n; e += TP_LB_THREADS) xfinal[e] = xp[e];
        xfinal = nullptr;
        ret = ls;
        break;
      }
      {
        const double* const A2[2] = {x, g};
        block_dots<STRICT, 2>(R, A2, A2, n, nr, tid);
      }
      xnorm = sqrt(nr[0]);
      gnorm = sqrt(nr[1]);
      if (xnorm < 1.0) xnorm = 1.0;
      if (gnorm / xnorm <= C.p.lbfgs_g_eps) { ret = 0; break; }
      if (C.p.lbfgs_max_iter != 0 && C.p.lbfgs_max_iter < k + 1) { ret = LB_MAXITER; break; }
      double* s = S + (size_t)end * n;
      double* y = Y + (size_t)end * n;
      for (int e = tid; e < n; e += TP_LB_THREADS) {
        s[e] = x[e] - xp[e];
        y[e] = g[e] - gp[e];
      }
      double yv[2];
      {
        const double* const A2[2] = {y, y};
        const double* const B2[2] = {s, y};
        block_dots<STRICT, 2>(R, A2, B2, n, yv, tid);
      }
      const double ys = yv[0], yy = yv[1];
      if (tid == 0) YS[end] = ys;  // read back only by this thread's registers below (broadcast via ysj)
      const int bound = (m <= k) ? m : k;
      bsum += bound;
      ++k;
      end = (end + 1) % m;
      for (int e = tid; e < n; e += TP_LB_THREADS) d[e] = -g[e];
      // two-loop recursion (lbfgs.hpp:1293-1316); ys_j / alpha_j kept in shared, written by all
      // threads with identical values would race benignly — instead thread 0 writes, all read
      // after the next barrier inside block_dots.
      int j = end;
      __syncthreads();
      for (int i = 0; i < bound; ++i) {
        j = (j + m - 1) % m;
        double a1[1];
        {
          const double* const A1[1] = {S + (size_t)j * n};
          const double* const B1[1] = {d};
          block_dots<STRICT, 1>(R, A1, B1, n, a1, tid);
        }
        const double al = a1[0] / YS[j];
        if (tid == 0) alpha[j] = al;
        const double cneg = -al;
        const double* yj = Y + (size_t)j * n;
        for (int e = tid; e < n; e += TP_LB_THREADS) d[e] += cneg * yj[e];
      }
      const double sc = ys / yy;
      for (int e = tid; e < n; e += TP_LB_THREADS) d[e] *= sc;
      __syncthreads();  // alpha[] written by thread 0 above must be visible
      for (int i = 0; i < bound; ++i) {
        double b1[1];
        {
          const double* const A1[1] = {Y + (size_t)j * n};
          const double* const B1[1] = {d};
          block_dots<STRICT, 1>(R, A1, B1, n, b1, tid);
        }
        const double beta = b1[0] / YS[j];
        const double cc = alpha[j] - beta;
        const double* sj = S + (size_t)j * n;
        for (int e = tid; e < n; e += TP_LB_THREADS) d[e] += cc * sj[e];
        j = (j + 1) % m;
      }
      step = 1.0;
    }
  }
  if (xfinal)
    for (int e = tid; e < n; e += TP_LB_THREADS) xfinal[e] = x[e];
  out.ret = ret;
  out.iters = k;
  out.evals = evals;
  out.reserved = bsum;  // sum of two-loop depths (flop accounting)
  out.fx = fx;
  __syncthreads();
}
