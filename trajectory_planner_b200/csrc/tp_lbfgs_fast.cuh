// Throughput form of the fused ViGO cost + L-BFGS kernel ("vector-free" L-BFGS on a Gram matrix).
//
// Same mathematics as lbfgs::lbfgs_optimize (solver/lbfgs.hpp:1024-1349) — m = 16 pairs, the
// More-Thuente line search of tp_lbfgs.cuh verbatim, the same convergence / iteration tests — but
// the two-loop recursion (lbfgs.hpp:1293-1316) is evaluated in COEFFICIENT SPACE:
//
//   d = sum_j a_j s_j + sum_j b_j y_j + c g,   (a, b, c) from the 16x16 Gram blocks
//   SY[i][j] = s_i.y_j,  YY[i][j] = y_i.y_j,  Sg[i] = s_i.g,  Yg[i] = y_i.g.
//
// Per iteration only ONE row/column of SY and YY and the two vectors Sg, Yg change: 5 x 16 dot
// products of length n that are independent of each other.  They are computed as one
// [S;Y] (32 x n) times [s_new y_new g] (n x 3) product on the FP64 tensor-core path
// (mma.sync m8n8k4.f64: the k-reduction happens inside the MMA, no shuffle trees), one 8-row tile
// per warp.  The 2 x 16 sequential block reductions of the classic form (one barrier + shuffle
// tree each) collapse into: one MMA pass, one barrier, a 16-step lane-parallel triangular solve in
// warp 0, one barrier, one fused multiply-add pass forming d.  Shared memory holds x/g/xp/gp/d,
// the 2 x 16 history rows and the Gram blocks; every vector element is owned by one thread.
//
// Rounding differs from the serial reference sums (as it already does for any parallel
// reduction); `strict_order` selects the classic kernel of tp_lbfgs.cuh, which reproduces the CPU
// iterate bit for bit.
#pragma once
#include "tp_lbfgs.cuh"

#define VF_M 16          // history pairs (bsplineTraj.cpp:697); other values use the classic kernel
#define VF_GS 17         // Gram row stride (doubles): conflict-free row reads across lanes
#define VF_PAIRS_SM 64   // guide pairs staged in shared memory (more -> read through L2)
#define VF_PAIRS_TM 160  // same, tm layout (the history is not in shared memory: room for more)
// One Gram buffer, indexed by AGE (0 = newest pair): SY[i][a] = s_i.y_a, YY[i][a] = y_i.y_a (row stride VF_GS),
// Sg[i] = s_i.g, Yg[i] = y_i.g, inv[i] = 1/(s_i.y_i), tb = scratch.  Every iteration the blocks shift by one
// age (old[i][a] -> new[i+1][a+1]) and the newest pair fills row / column 0.
#define VF_G_SY 0
#define VF_G_YY (VF_M * VF_GS)
#define VF_G_SG (2 * VF_M * VF_GS)
#define VF_G_YG (VF_G_SG + VF_M)
#define VF_G_INV (VF_G_YG + VF_M)
#define VF_G_TB (VF_G_INV + VF_M)
#define VF_GRAM (VF_G_TB + VF_M)

// History in TENSOR MEMORY (tm layout): the 2 x 16 history values of a vector element are only ever combined with
// other values of the SAME element (Gram products, direction), so every thread keeps the history of the elements it
// owns in its own TMEM lane: element q of thread t = columns [64 q, 64 q + 64) of lane t (16 s then 16 y doubles, by
// slot).  VF_TM_Q elements per thread fit the block's VF_TM_COLS columns (4 blocks x 128 = the SM's 512 columns);
// the elements beyond 128 * VF_TM_Q keep their history columns in shared memory ("overflow").
#define VF_TM_COLS 128
#define VF_TM_Q 2
#define VF_TM_ELEMS (VF_TM_Q * TP_LB_THREADS)

struct VfLayout {  // offsets in doubles from the start of dynamic shared memory
  int ns;          // smem layout: padded history row stride, ns % 16 == 4 (conflict-free MMA fragment loads);
                   // tm layout: row stride of the overflow history (elements e >= VF_TM_ELEMS), 0 if none
  int cp, g, xp, gp, d, S, Y, gram, ca, cb, sc, red, gred, pair, pstart, total;
  int pairs_cap;   // guide pairs the `pair` region holds
};
__host__ __device__ inline VfLayout vf_layout(int N, bool tm = false) {
  VfLayout L;
  const int n = 3 * (N - 2 * TP_DEGREE);
  const int nn = n > 0 ? n : 0;
  int ns = (nn + 15) / 16 * 16 + 4;
  int gs = ns;
  if (tm) {
    const int novf = nn + 3 * TP_DEGREE - VF_TM_ELEMS;
    ns = novf > 0 ? novf + (novf & 1) : 0;
    gs = nn + (nn & 1);
  }
  L.ns = ns;
  int o = 0;
  L.cp = o; o += 3 * N + (N & 1);
  L.g = o; o += gs;
  L.xp = o; o += nn + (nn & 1);
  L.gp = o; o += nn + (nn & 1);
  L.d = o; o += nn + (nn & 1);
  L.S = o; o += VF_M * ns;
  L.Y = o; o += VF_M * ns;
  L.gram = o; o += 2 * VF_GRAM;   // two age-ordered Gram buffers (ping-pong), see VF_G_*
  L.ca = o; o += VF_M;
  L.cb = o; o += VF_M;
  L.sc = o; o += 8;
  L.red = o; o += 2 * TP_LB_WARPS * 4;
  L.gred = o; o += tm ? TP_LB_WARPS * (2 * VF_M + 2) : 0;   // per-warp partial Gram sums (tm layout)
  L.pairs_cap = tm ? VF_PAIRS_TM : VF_PAIRS_SM;
  L.pair = o; o += L.pairs_cap * 7;
  L.pstart = o; o += (N + 2 + 1) / 2;   // (N + 1) ints
  L.total = o;
  return L;
}
__host__ __device__ inline size_t vf_smem_bytes(int N, bool tm = false) { return (size_t)vf_layout(N, tm).total * 8; }

struct VfCtx {
  int N, n;
  double* sm;          // dynamic shared memory base
  VfLayout L;
  const GuidePair* pairs;   // global list of this trajectory
  const int* head;          // global per-control-point list heads
  int n_pairs;
  int serial_warp;     // which warp runs the serial coefficient phase (rotated per block: warp w lives on SM sub-partition w)
  uint32_t tbase;      // tm layout: TMEM address of this block's columns (lane 0)
  bool pairs_in_sm;
  double w_dist, w_dyn;
  int n_dyn;
  const double* dyn_pos;
  const double* dyn_vel;
  const double* dyn_size;
};

// one (pair, control point) term of getDistanceCost, bsplineTraj.cpp:839-895 (same expressions
// as distance_terms in tp_lbfgs.cuh)
__device__ __forceinline__ void vf_pair_term(const VigoConst& C, double cx, double cy, double cz, double px, double py,
                                             double pz, double vx, double vy, double vz, bool unk, int a, double& grad,
                                             double& cost) {
  const double dth = C.p.dthresh;
  const double dist = fma(cx - px, vx, fma(cy - py, vy, (cz - pz) * vz));
  const double e = dth - dist;
  double va = a == 0 ? vx : (a == 1 ? vy : vz);
  if (!C.p.plan_in_z && a == 2) va = 0.0;
  double costTemp, gt;
  if (e <= -dth) {                       // far beyond the plane: (-e)^3, not scaled by the uncertain factor (:852-861)
    costTemp = -(e * e) * e;
    gt = 3.0 * (e * e);
    unk = false;
  } else if (e > 0 && e <= dth) {        // (:862-878)
    costTemp = (e * e) * e;
    gt = -3.0 * (e * e);
  } else if (e >= dth) {                 // (:879-894)
    costTemp = fma(fma(C.dist_a, e, C.dist_b), e, C.dist_c);
    gt = -fma(2.0 * C.dist_a, e, C.dist_b);
  } else {
    return;
  }
  if (unk) { costTemp *= C.p.uncertain_factor; gt *= C.p.uncertain_factor; }
  grad = fma(gt, va, grad);
  cost += costTemp;
}

// height barrier of getDistanceCost (bsplineTraj.cpp:897-930); gradient lands on the X row (quirk)
__device__ __forceinline__ void vf_height_term(const VigoConst& C, double cz, int a, double& grad, double& cost) {
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

// stage this trajectory's guide pairs as a CSR list (by control point, append order) in shared memory
__device__ void vf_stage_pairs(VfCtx& V, int tid) {
  int* pstart = reinterpret_cast<int*>(V.sm + V.L.pstart);
  double* ps = V.sm + V.L.pair;
  V.pairs_in_sm = V.n_pairs <= V.L.pairs_cap;
  if (!V.pairs_in_sm) return;   // uniform
  // counts per control point (thread per control point walks its list)
  for (int c = tid; c < V.N; c += TP_LB_THREADS) {
    int cnt = 0;
    for (int gi = V.head[c]; gi >= 0; gi = V.pairs[gi].next) ++cnt;
    pstart[c + 1] = cnt;
  }
  if (tid == 0) pstart[0] = 0;
  __syncthreads();
  if (tid == 0)
    for (int c = 0; c < V.N; ++c) pstart[c + 1] += pstart[c];
  __syncthreads();
  for (int c = tid; c < V.N; c += TP_LB_THREADS) {
    int w = pstart[c];
    for (int gi = V.head[c]; gi >= 0;) {
      const GuidePair& pr = V.pairs[gi];
      double* q = ps + 7 * w;
      q[0] = pr.p[0]; q[1] = pr.p[1]; q[2] = pr.p[2];
      q[3] = pr.v[0]; q[4] = pr.v[1]; q[5] = pr.v[2];
      q[6] = pr.unknown ? 1.0 : 0.0;
      ++w;
      gi = pr.next;
    }
  }
  __syncthreads();
}

// signed excess over the +-1 box of getFeasibilityCost (bsplineTraj.cpp:955-956: maxVel = maxAcc = 1.0 hard-coded)
__device__ __forceinline__ double vf_excess(double v) { return fmax(v - 1.0, 0.0) + fmin(v + 1.0, 0.0); }

// costFunction (bsplineTraj.cpp:802-821) at the control points in shared memory: writes the gradient
// of the elements this thread owns, returns this thread's shares of {f, g.d, g.g, x.x}.
// Throughput form: every (control point, axis) element reads its 7-point stencil once, divisions by
// the constant control-point timestep become multiplications by its reciprocal, the +-1 feasibility
// branches become max/min, products feed FMAs.  Same terms as getSmoothnessCost / getFeasibilityCost /
// getDistanceCost / getDynamicObstacleCost; only the rounding order differs from the serial reference.
__device__ __forceinline__ void vf_eval_partial(const VigoConst& C, const VfCtx& V, bool with_d, double (&v)[4], int tid) {
  const int N = V.N;
  const double* cp = V.sm + V.L.cp;
  double* g = V.sm + V.L.g;
  const double* d = V.sm + V.L.d;
  const int* pstart = reinterpret_cast<const int*>(V.sm + V.L.pstart);
  const double* ps = V.sm + V.L.pair;
  const double icts = 1.0 / C.p.ctrl_pt_ts;
  const double k2 = C.ts_inv_sqr;
  const double gv_c = 2.0 * icts * k2;   // d/dc of (v -+ 1)^2 k2
  double sD = 0, sS = 0, sF = 0, sO = 0, dg = 0, gg = 0, xx = 0;
  for (int e = tid; e < 3 * N; e += TP_LB_THREADS) {
    const int c = e / 3, a = e - 3 * c;
    if (c < TP_DEGREE || c > N - TP_DEGREE - 1) {
      // fixed control points: only the cost terms they own (their stencils reach optimised points)
      if (c <= N - 4) {
        const double j = cp[e + 9] - 3 * cp[e + 6] + 3 * cp[e + 3] - cp[e];
        sS = fma(j, j, sS);
      }
      if (c <= N - 2) {
        const double ev = vf_excess((cp[e + 3] - cp[e]) * icts);
        sF = fma(ev * ev, k2, sF);
      }
      if (c <= N - 3) {
        const double ea = vf_excess((cp[e + 6] - 2 * cp[e + 3] + cp[e]) * k2);
        sF = fma(ea, ea, sF);
      }
      continue;
    }
    const double pm3 = cp[e - 9], pm2 = cp[e - 6], pm1 = cp[e - 3], p0 = cp[e], p1 = cp[e + 3], p2 = cp[e + 6], p3 = cp[e + 9];
    // jerk terms i = c-3 .. c (getSmoothnessCost, :938-947)
    const double j0 = (p0 - pm3) - 3.0 * (pm1 - pm2);
    const double j1 = (p1 - pm2) - 3.0 * (p0 - pm1);
    const double j2 = (p2 - pm1) - 3.0 * (p1 - p0);
    const double j3 = (p3 - p0) - 3.0 * (p2 - p1);
    sS = fma(j3, j3, sS);
    const double gs = 2.0 * ((j0 - j3) + 3.0 * (j2 - j1));
    // feasibility (getFeasibilityCost, :961-995)
    const double evm = vf_excess((p0 - pm1) * icts), ev0 = vf_excess((p1 - p0) * icts);
    const double ea0 = vf_excess((p0 - 2.0 * pm1 + pm2) * k2), ea1 = vf_excess((p1 - 2.0 * p0 + pm1) * k2),
                 ea2 = vf_excess((p2 - 2.0 * p1 + p0) * k2);
    sF = fma(ev0 * ev0, k2, sF);
    sF = fma(ea2, ea2, sF);
    const double gf = gv_c * (evm - ev0) + 2.0 * k2 * ((ea0 + ea2) - 2.0 * ea1);
    // distance to the guide planes (getDistanceCost, :839-930)
    double gd = 0.0, cD = 0.0;
    {
      const double cx = cp[3 * c], cy = cp[3 * c + 1], cz = cp[3 * c + 2];
      if (V.pairs_in_sm) {
        for (int q = pstart[c]; q < pstart[c + 1]; ++q) {
          const double* pr = ps + 7 * q;
          vf_pair_term(C, cx, cy, cz, pr[0], pr[1], pr[2], pr[3], pr[4], pr[5], pr[6] != 0.0, a, gd, cD);
        }
      } else {
        for (int gi = V.head[c]; gi >= 0;) {
          const GuidePair& pr = V.pairs[gi];
          vf_pair_term(C, cx, cy, cz, pr.p[0], pr.p[1], pr.p[2], pr.v[0], pr.v[1], pr.v[2], pr.unknown != 0, a, gd, cD);
          gi = pr.next;
        }
      }
      if (C.p.plan_in_z) vf_height_term(C, cz, a, gd, cD);
    }
    if (a == 0) sD += cD;
    double go = 0.0;
    if (V.n_dyn > 0) {
      EvalCtx E;
      E.N = N; E.n = V.n; E.cp = const_cast<double*>(cp); E.pairs = V.pairs; E.head = V.head;
      E.w_dist = V.w_dist; E.w_dyn = V.w_dyn; E.n_dyn = V.n_dyn; E.dyn_pos = V.dyn_pos; E.dyn_vel = V.dyn_vel;
      E.dyn_size = V.dyn_size;
      double cO = 0.0;
      dynamic_terms(C, E, c, a, go, cO);
      if (a == 0) sO += cO;
    }
    const double gv = fma(V.w_dist, gd, fma(C.p.w_smooth, gs, fma(C.p.w_feas, gf, V.w_dyn * go)));
    const int i = e - 3 * TP_DEGREE;
    g[i] = gv;
    gg = fma(gv, gv, gg);
    xx = fma(p0, p0, xx);
    if (with_d) dg = fma(gv, d[i], dg);
  }
  v[0] = V.w_dist * sD + C.p.w_smooth * sS + C.p.w_feas * sF + V.w_dyn * sO;
  v[1] = dg;
  v[2] = gg;
  v[3] = xx;
}

#ifdef TP_LBFGS_TIMING
__device__ long long g_tp_t_partial = 0, g_tp_t_sum = 0;
__device__ unsigned long long g_tp_phase[8];   // whole-batch phase totals: total, eval, gram, coeffs, direction, iterations, evals
#endif
__device__ __forceinline__ void vf_eval(const VigoConst& C, const VfCtx& V, Red& R, bool with_d, double& f, double& dg,
                                        double& gg, double& xx, int tid) {
  double v[4];
#ifdef TP_LBFGS_TIMING
  long long t0 = clock64();
#endif
  vf_eval_partial(C, V, with_d, v, tid);
#ifdef TP_LBFGS_TIMING
  long long t1 = clock64();
#endif
  block_sum<4>(R, v, tid);
#ifdef TP_LBFGS_TIMING
  if (tid == 0) { g_tp_t_partial += t1 - t0; g_tp_t_sum += clock64() - t1; }
#endif
  f = v[0]; dg = v[1]; gg = v[2]; xx = v[3];
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Gram update after the pair (s_new, y_new) was stored in slot `slot`: one 8-row tile of [S;Y] per
// warp against the columns [s_new, y_new, g].  Entered after a barrier that made the new rows and
// g visible; the caller issues the barrier that publishes the Gram entries.
__device__ __forceinline__ void vf_gram_update(const VfCtx& V, int slot, int gbuf, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  const int ns = V.L.ns, n = V.n;
  const double* S = V.sm + V.L.S;
  const double* Y = V.sm + V.L.Y;
  const double* g = V.sm + V.L.g;
  const int row = lane >> 2, kq = lane & 3;
  const int j = 8 * (warp & 1) + row;                              // history slot of this lane's A row
  const double* arow = (warp < 2 ? S : Y) + (size_t)j * ns;
  const int col = lane >> 2;                                        // B column of this lane
  const double* bvec = col == 0 ? S + (size_t)slot * ns : (col == 1 ? Y + (size_t)slot * ns : g);
  const bool bvalid = col < 3;
  double c00 = 0, c01 = 0, c10 = 0, c11 = 0, c20 = 0, c21 = 0, c30 = 0, c31 = 0;
  // history rows and g are zero-padded up to ns >= roundup16(n) + 4: no tail guards; four independent
  // accumulator pairs break the MMA dependency chain
  const int nk = (n + 15) & ~15;
  for (int k0 = 0; k0 < nk; k0 += 16) {
    const double a0 = arow[k0 + kq], a1 = arow[k0 + 4 + kq], a2 = arow[k0 + 8 + kq], a3 = arow[k0 + 12 + kq];
    double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
    if (bvalid) { b0 = bvec[k0 + kq]; b1 = bvec[k0 + 4 + kq]; b2 = bvec[k0 + 8 + kq]; b3 = bvec[k0 + 12 + kq]; }
    dmma884(c00, c01, a0, b0);
    dmma884(c10, c11, a1, b1);
    dmma884(c20, c21, a2, b2);
    dmma884(c30, c31, a3, b3);
  }
  const double r0 = (c00 + c10) + (c20 + c30);   // C[row][2*kq]
  const double r1 = (c01 + c11) + (c21 + c31);   // C[row][2*kq + 1]
  double* G = V.sm + V.L.gram + (size_t)gbuf * VF_GRAM;
  const int age = (slot - j) & (VF_M - 1);   // age of pair j once `slot` is the newest
  if (warp < 2) {
    // rows s_j: col0 = s_j.s_new (unused), col1 = s_j.y_new, col2 = s_j.g
    if (kq == 0) { if (age != 0) G[VF_G_SY + age * VF_GS + 0] = r1; }
    else if (kq == 1) G[VF_G_SG + age] = r0;
  } else {
    // rows y_j: col0 = y_j.s_new, col1 = y_j.y_new, col2 = y_j.g
    if (kq == 0) {
      G[VF_G_SY + 0 * VF_GS + age] = r0;
      G[VF_G_YY + 0 * VF_GS + age] = r1;
      G[VF_G_YY + age * VF_GS + 0] = r1;
    } else if (kq == 1) G[VF_G_YG + age] = r0;
  }
}

// shift the age-ordered Gram blocks by one pair: new[i+1][a+1] = old[i][a], new inv[a+1] = old inv[a]
__device__ __forceinline__ void vf_gram_shift(const VfCtx& V, int gbuf_new, int tid) {
  const double* O = V.sm + V.L.gram + (size_t)(gbuf_new ^ 1) * VF_GRAM;
  double* G = V.sm + V.L.gram + (size_t)gbuf_new * VF_GRAM;
  // both blocks are contiguous [16][17] arrays: cell c = i*17 + a moves to c + 18; walk the 2 x 272-cell range
  for (int c = tid; c < 2 * VF_M * VF_GS; c += TP_LB_THREADS) {
    const int cc = c >= VF_M * VF_GS ? c - VF_M * VF_GS : c;
    const int i = cc / VF_GS, a = cc - i * VF_GS;
    if (i < VF_M - 1 && a < VF_M - 1) G[c + VF_GS + 1] = O[c];
  }
  if (tid < VF_M - 1) G[VF_G_INV + tid + 1] = O[VF_G_INV + tid];
}

// Two-loop recursion in coefficient space (lbfgs.hpp:1293-1316) on the age-ordered Gram buffer `G`
// (age i = the i-th newest pair; entries of pairs not stored yet are zero).  One warp, lane i <-> pair of age i.
// Both triangular recurrences are ROLLED loops (the SM's instruction cache, not its pipes, is what the
// fused kernel runs out of — profiles/r01c): per step the finished value is broadcast through shared
// memory and every lane applies one FMA to its own running sum.
// `newest` = slot of the newest pair, gg = g.g.  Writes ca[slot], cb[slot] (coefficients of
// s_slot, y_slot), sc[0] = coefficient of g, sc[1] = g.d (the line search's dginit).
// (noinline functions take plain pointers / scalars: a struct passed by reference would live in local memory)
__device__ __noinline__ void vf_coeffs(double* G, double* ca, double* cb, double* sc, int newest, int bound, double gg, int lane) {
  const double* SY = G + VF_G_SY;
  const double* YY = G + VF_G_YY;
  double* tb = G + VF_G_TB;          // broadcast scratch (16 doubles)
  const unsigned FULL = 0xffffffffu;
  const int i = lane & (VF_M - 1);
  // 1/ys and 1/yy of the newest pair (two lanes divide concurrently)
  const double dd = lane == 0 ? SY[0] : YY[0];
  const double rc = 1.0 / dd;
  const double inv0 = __shfl_sync(FULL, rc, 0);
  const double gamma = SY[0] * __shfl_sync(FULL, rc, 1);   // ys/yy (lbfgs.hpp:1305)
  if (lane == 0) G[VF_G_INV] = inv0;
  const double inv_i = i == 0 ? inv0 : G[VF_G_INV + i];     // 0 for pairs not stored yet
  const double sg_i = G[VF_G_SG + i], yg_i = G[VF_G_YG + i];
  // Both recurrences run on ROW-SCALED quantities (r_i / ys_i, (y_i.d) / ys_i) so that the value a step broadcasts is
  // a lane's running value itself, and on matrix rows preloaded into registers (masked to the strict triangle): per
  // step ONE register shuffle + ONE FMA on the dependent chain.  All 16 steps always run: entries of pairs not
  // stored yet are zero (so are their 1/ys), which makes their steps no-ops.
  (void)bound;
  double lo[VF_M];   // lo[a] = (s_i.y_a) / ys_i for a < i (pair i older than pair a), else 0
#pragma unroll
  for (int a = 0; a < VF_M; ++a) lo[a] = i > a ? SY[i * VF_GS + a] * inv_i : 0.0;
  // ---- first loop, newest -> oldest: alpha_a = (s_a.q)/ys_a; q -= alpha_a y_a  (q starts at -g)
  double rr = -sg_i * inv_i;
  double al = 0.0;
#pragma unroll
  for (int a = 0; a < VF_M; ++a) {
    const double ala = __shfl_sync(FULL, rr, a);
    if (i == a) al = ala;
    rr = fma(-ala, lo[a], rr);
  }
  // ---- acc_i = y_i.(gamma q) = -gamma (y_i.g + sum_a alpha_a y_i.y_a)
  if (lane < VF_M) tb[i] = al;
  double up[VF_M];   // up[j] = (s_j.y_i) / ys_i for j > i, else 0
#pragma unroll
  for (int j = 0; j < VF_M; ++j) up[j] = i < j ? SY[j * VF_GS + i] * inv_i : 0.0;
  __syncwarp();
  double t0 = yg_i, t1 = 0.0, t2 = 0.0, t3 = 0.0;
#pragma unroll
  for (int a = 0; a < VF_M; a += 4) {
    t0 = fma(tb[a], YY[i * VF_GS + a], t0);
    t1 = fma(tb[a + 1], YY[i * VF_GS + a + 1], t1);
    t2 = fma(tb[a + 2], YY[i * VF_GS + a + 2], t2);
    t3 = fma(tb[a + 3], YY[i * VF_GS + a + 3], t3);
  }
  // ---- second loop, oldest -> newest: beta_j = (y_j.d)/ys_j; d += (alpha_j - beta_j) s_j; accp_i = (y_i.d)/ys_i
  double accp = (-gamma * ((t0 + t1) + (t2 + t3))) * inv_i;
#pragma unroll
  for (int j = VF_M - 1; j >= 0; --j) {
    const double bj = __shfl_sync(FULL, accp, j);   // final for pair j: later steps only touch i < j
    const double u = fma(tb[j], up[j], accp);        // off the dependent chain
    accp = fma(-bj, up[j], u);
  }
  const double aa = al - accp;
  const bool mine = lane < VF_M && i < bound;
  const double bb = -gamma * al;
  if (mine) {
    const int si = (newest - i) & (VF_M - 1);
    ca[si] = aa;
    cb[si] = bb;
  }
  // g.d = sum aa_i s_i.g + sum b_i y_i.g - gamma g.g
  double part = mine ? fma(aa, sg_i, bb * yg_i) : 0.0;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
  if (lane == 0) {
    sc[0] = -gamma;
    sc[1] = part - gamma * gg;
  }
}

// ---------------------------------------------------------------------------- history in tensor memory
// tcgen05.ld / st with the 32x32b shape: thread t of warp w reads / writes consecutive 32-bit columns of TMEM lane
// 32 (w % 4) + t — a lane-private scratchpad next to the register file (measured: tools/microbench/tmem_rt.cu).
__device__ __forceinline__ void tm_st1(uint32_t taddr, double v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(__double2loint(v)), "r"(__double2hiint(v)) : "memory");
}
__device__ __forceinline__ void tm_st8(uint32_t taddr, const double (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
               "r"(__double2loint(v[0])), "r"(__double2hiint(v[0])), "r"(__double2loint(v[1])), "r"(__double2hiint(v[1])),
               "r"(__double2loint(v[2])), "r"(__double2hiint(v[2])), "r"(__double2loint(v[3])), "r"(__double2hiint(v[3])),
               "r"(__double2loint(v[4])), "r"(__double2hiint(v[4])), "r"(__double2loint(v[5])), "r"(__double2hiint(v[5])),
               "r"(__double2loint(v[6])), "r"(__double2hiint(v[6])), "r"(__double2loint(v[7])), "r"(__double2hiint(v[7]))
               : "memory");
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 16 columns = 8 doubles; the wait takes the destination registers as in/out operands so that no use can be
// scheduled above it
__device__ __forceinline__ void tm_ld8(uint32_t taddr, double (&v)[8]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])::"memory");
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = __hiloint2double((int)r[2 * k + 1], (int)r[2 * k]);
}
// two 16-column loads in flight, one wait
__device__ __forceinline__ void tm_ld8x2(uint32_t ta, uint32_t tb, double (&va)[8], double (&vb)[8]) {
  uint32_t r[16], q[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(ta));
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
                 "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
               : "r"(tb));
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(q[0]),
                 "+r"(q[1]), "+r"(q[2]), "+r"(q[3]), "+r"(q[4]), "+r"(q[5]), "+r"(q[6]), "+r"(q[7]), "+r"(q[8]), "+r"(q[9]),
                 "+r"(q[10]), "+r"(q[11]), "+r"(q[12]), "+r"(q[13]), "+r"(q[14]), "+r"(q[15])::"memory");
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    va[k] = __hiloint2double((int)r[2 * k + 1], (int)r[2 * k]);
    vb[k] = __hiloint2double((int)q[2 * k + 1], (int)q[2 * k]);
  }
}
// allocate this block's VF_TM_COLS columns (warp 0), publish the address through `slot` (shared); returns the
// address with the calling warp's lane quarter.  Blocks until columns are free: every holder finishes on its own.
__device__ __forceinline__ uint32_t tm_block_alloc(uint32_t* slot, int tid) {
  if ((tid >> 5) == 0) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(slot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sa), "r"(VF_TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  return *slot + ((uint32_t)(((tid >> 5) & 3) * 32) << 16);
}
__device__ __forceinline__ void tm_block_free(const uint32_t* slot, int tid) {
  __syncthreads();
  if ((tid >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*slot), "r"(VF_TM_COLS) : "memory");
}

// number of owned-element rounds of a block: element e = tid + 128 q of the control-point array (e - 9 = index in x)
__device__ __forceinline__ int vf_rounds(int n) { return (n + 3 * TP_DEGREE + TP_LB_THREADS - 1) / TP_LB_THREADS; }

// 2 x 8 history values of owned element q: slots 8 ha .. 8 ha + 7 of S (ya = 0) or Y (ya = 1), and the same for b
__device__ __forceinline__ void vf_hist_ld8x2(const VfCtx& V, int q, int ya, int ha, int yb, int hb, int tid, double (&va)[8],
                                              double (&vb)[8]) {
  if (q < VF_TM_Q) {
    tm_ld8x2(V.tbase + (uint32_t)(64 * q + 32 * ya + 16 * ha), V.tbase + (uint32_t)(64 * q + 32 * yb + 16 * hb), va, vb);
  } else {
    const int o = tid + TP_LB_THREADS * (q - VF_TM_Q);
    const bool ok = o < V.L.ns;
    const double* ra = V.sm + (ya ? V.L.Y : V.L.S) + (size_t)(8 * ha) * V.L.ns + o;
    const double* rb = V.sm + (yb ? V.L.Y : V.L.S) + (size_t)(8 * hb) * V.L.ns + o;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      va[k] = ok ? ra[(size_t)k * V.L.ns] : 0.0;
      vb[k] = ok ? rb[(size_t)k * V.L.ns] : 0.0;
    }
  }
}
__device__ __forceinline__ void vf_hist_st(const VfCtx& V, int q, int slot, int tid, double s, double y) {
  if (q < VF_TM_Q) {
    tm_st1(V.tbase + (uint32_t)(64 * q + 2 * slot), s);
    tm_st1(V.tbase + (uint32_t)(64 * q + 32 + 2 * slot), y);
    tm_wait_st();
  } else {
    const int o = tid + TP_LB_THREADS * (q - VF_TM_Q);
    if (o < V.L.ns) {
      V.sm[V.L.S + (size_t)slot * V.L.ns + o] = s;
      V.sm[V.L.Y + (size_t)slot * V.L.ns + o] = y;
    }
  }
}
__device__ __forceinline__ void vf_hist_zero(const VfCtx& V, int tid) {
  const double z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int c = 0; c < VF_TM_COLS; c += 16) tm_st8(V.tbase + c, z);
  tm_wait_st();
  for (int e = tid; e < 2 * VF_M * V.L.ns; e += TP_LB_THREADS) V.sm[V.L.S + e] = 0.0;
}

// Gram update, tm layout, part 1 (all threads): store the new pair (s, y) = (x - xp, g - gp) in slot `slot`, then the
// per-warp partial sums of   s_j.g, y_j.g  (j = 0..15, the new pair included)  and  s_new.y_new, y_new.y_new.
// Only 2 x 16 + 2 fresh dot products per iteration: the other Gram entries of the new pair follow from
// y_new = g - g_prev by linearity (vf_gram_finish_tm).
// The 32 sums over the warp's lanes are one recursive-halving reduce-scatter (31 double shuffles, 5 dependent steps,
// fixed order => deterministic): lanes 0-15 end with s_lane.g, lanes 16-31 with y_(lane-16).g.
__device__ __noinline__ void vf_gram_partial_tm(double* sm_, int N_, uint32_t tbase, int slot, int tid) {
  VfCtx V;   // only the fields the history accessors read
  V.sm = sm_; V.N = N_; V.n = 3 * (N_ - 2 * TP_DEGREE); V.L = vf_layout(N_, true); V.tbase = tbase;
  const unsigned FULL = 0xffffffffu;
  const int lane = tid & 31, warp = tid >> 5, n = V.n, Q = vf_rounds(n);
  const double* cp = V.sm + V.L.cp;
  const double* g = V.sm + V.L.g;
  const double* xp = V.sm + V.L.xp;
  const double* gp = V.sm + V.L.gp;
  double* gred = V.sm + V.L.gred + warp * (2 * VF_M + 2);
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
  double e0 = 0.0, e1 = 0.0;
  double acc[16];   // after the xor-16 exchange: lanes 0-15 hold the S-type sums, lanes 16-31 the Y-type sums
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.0;
  for (int q = 0; q < Q; ++q) {
    const int e = tid + TP_LB_THREADS * q, i = e - 3 * TP_DEGREE;
    const bool ok = i >= 0 && i < n;
    const double gi = ok ? g[i] : 0.0;
    const double s = ok ? cp[e] - xp[i] : 0.0;
    const double y = ok ? gi - gp[i] : 0.0;
    vf_hist_st(V, q, slot, tid, s, y);
    e0 = fma(s, y, e0);
    e1 = fma(y, y, e1);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double vs[8], vy[8];
      vf_hist_ld8x2(V, q, 0, h, 1, h, tid, vs, vy);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const double ps = vs[k] * gi, py = vy[k] * gi;
        const double send = b4 ? ps : py, keep = b4 ? py : ps;
        acc[8 * h + k] += keep + __shfl_xor_sync(FULL, send, 16);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double send = b3 ? acc[k] : acc[k + 8], keep = b3 ? acc[k + 8] : acc[k];
    acc[k] = keep + __shfl_xor_sync(FULL, send, 8);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double send = b2 ? acc[k] : acc[k + 4], keep = b2 ? acc[k + 4] : acc[k];
    acc[k] = keep + __shfl_xor_sync(FULL, send, 4);
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const double send = b1 ? acc[k] : acc[k + 2], keep = b1 ? acc[k + 2] : acc[k];
    acc[k] = keep + __shfl_xor_sync(FULL, send, 2);
  }
  {
    const double send = b0 ? acc[0] : acc[1], keep = b0 ? acc[1] : acc[0];
    acc[0] = keep + __shfl_xor_sync(FULL, send, 1);
  }
  e0 = warp_sum(e0);
  e1 = warp_sum(e1);
  gred[lane] = acc[0];
  if (lane == 0) {
    gred[2 * VF_M] = e0;
    gred[2 * VF_M + 1] = e1;
  }
}

// Gram update, tm layout, part 2 (the serial warp, after a barrier): combine the warps' partial sums in a fixed
// order and fill the newest pair's row / column of the age-ordered blocks `G` (O = the previous iteration's buffer):
//   Sg[a] = s_a.g, Yg[a] = y_a.g                                  fresh
//   SY[a][0] = s_a.y_new = s_a.g - s_a.g_prev = Sg[a] - O.Sg[a-1]   (a >= 1: pair a was pair a-1 one iteration ago)
//   YY[a][0] = YY[0][a] = y_a.y_new = Yg[a] - O.Yg[a-1]
//   SY[0][0] = s_new.y_new, YY[0][0] = y_new.y_new                fresh
__device__ __forceinline__ void vf_gram_finish_tm(const VfCtx& V, double* G, const double* O, int slot, int lane) {
  const double* gred = V.sm + V.L.gred;
  const int str = 2 * VF_M + 2;
  const double tot = ((gred[lane] + gred[str + lane]) + gred[2 * str + lane]) + gred[3 * str + lane];
  const int j = lane & (VF_M - 1);
  const int age = (slot - j) & (VF_M - 1);
  if (lane < VF_M) {
    G[VF_G_SG + age] = tot;
    if (age != 0) G[VF_G_SY + age * VF_GS] = tot - O[VF_G_SG + age - 1];
    else G[VF_G_SY] = ((gred[2 * VF_M] + gred[str + 2 * VF_M]) + gred[2 * str + 2 * VF_M]) + gred[3 * str + 2 * VF_M];
  } else {
    G[VF_G_YG + age] = tot;
    if (age != 0) {
      const double v = tot - O[VF_G_YG + age - 1];
      G[VF_G_YY + age] = v;
      G[VF_G_YY + age * VF_GS] = v;
    } else {
      G[VF_G_YY] = ((gred[2 * VF_M + 1] + gred[str + 2 * VF_M + 1]) + gred[2 * str + 2 * VF_M + 1]) + gred[3 * str + 2 * VF_M + 1];
    }
  }
  __syncwarp();
}

// d = cg g + sum_j ca[j] s_j + cb[j] y_j over all 16 slots (coefficients of slots not stored yet are zero)
__device__ __noinline__ void vf_direction_tm(double* sm_, int N_, uint32_t tbase, double cg, int tid) {
  VfCtx V;
  V.sm = sm_; V.N = N_; V.n = 3 * (N_ - 2 * TP_DEGREE); V.L = vf_layout(N_, true); V.tbase = tbase;
  const int n = V.n, Q = vf_rounds(n);
  const double* g = V.sm + V.L.g;
  const double* ca = V.sm + V.L.ca;
  const double* cb = V.sm + V.L.cb;
  double* d = V.sm + V.L.d;
  for (int q = 0; q < Q; ++q) {
    const int i = tid + TP_LB_THREADS * q - 3 * TP_DEGREE;
    const bool ok = i >= 0 && i < n;
    double a0 = ok ? cg * g[i] : 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    double va[8], vb[8];
    vf_hist_ld8x2(V, q, 0, 0, 0, 1, tid, va, vb);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      a0 = fma(ca[k], va[k], a0);
      a1 = fma(ca[8 + k], vb[k], a1);
    }
    vf_hist_ld8x2(V, q, 1, 0, 1, 1, tid, va, vb);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      a2 = fma(cb[k], va[k], a2);
      a3 = fma(cb[8 + k], vb[k], a3);
    }
    if (ok) d[i] = (a0 + a1) + (a2 + a3);
  }
}

// One optimize() (bsplineTraj.cpp:687-718), throughput form.  V.sm + L.cp holds the control points on
// entry; on return it holds the LAST EVALUATED point (bsplineTraj.cpp:803) and xfinal (global, may be
// null) the solver's own x.
template <bool TM>
__device__ void lbfgs_run_fast(const VigoConst& C, VfCtx& V, tp_lbfgs_result& out, double* xfinal, int tid) {
  const int n = V.n, N = V.N, ns = V.L.ns;
  double* sm = V.sm;
  double* cp = sm + V.L.cp;
  double* x = cp + 3 * TP_DEGREE;
  double* g = sm + V.L.g;
  double* xp = sm + V.L.xp;
  double* gp = sm + V.L.gp;
  double* d = sm + V.L.d;
  double* S = sm + V.L.S;
  double* Y = sm + V.L.Y;
  const double* ca = sm + V.L.ca;
  const double* cb = sm + V.L.cb;
  const double* sc = sm + V.L.sc;
  Red R;
  R.buf = sm + V.L.red;
  R.flip = 0;
  const double min_step = 1e-20, max_step = 1e20, ftol = 1e-4, gtol = 0.9, xtol = 1e-16;  // lbfgs.hpp:942-954
  const int max_ls = C.p.lbfgs_max_linesearch;
  const double geps2 = C.p.lbfgs_g_eps * C.p.lbfgs_g_eps;
  // zero the history rows, the Gram blocks and the padding of g (garbage would poison the MMA tiles)
  if (TM) {
    vf_hist_zero(V, tid);
  } else {
    for (int e = tid; e < 2 * VF_M * ns; e += TP_LB_THREADS) S[e] = 0.0;   // S and Y are contiguous
    for (int e = n + tid; e < ns; e += TP_LB_THREADS) g[e] = 0.0;
  }
  for (int e = tid; e < 2 * VF_GRAM + 2 * VF_M + 8; e += TP_LB_THREADS) (sm + V.L.gram)[e] = 0.0;   // gram, ca, cb, sc
  vf_stage_pairs(V, tid);
  __syncthreads();
  int evals = 0, k = 0, ret, bsum = 0;
  double fx, dgd, gg, xx;
#ifdef TP_LBFGS_TIMING
  long long tE = 0, tG = 0, tC = 0, tD = 0, tT0 = clock64(), tq0, tq1;
#define LT0 tq0 = clock64();
#define LT(acc) { tq1 = clock64(); acc += tq1 - tq0; tq0 = tq1; }
#else
#define LT0
#define LT(acc)
#endif
  vf_eval(C, V, R, false, fx, dgd, gg, xx, tid);
  ++evals;
  // owned elements: i = e - 9 for e = tid + q*THREADS in [9, 9 + n)
#define VF_OWNED(i) for (int i = tid - 3 * TP_DEGREE; i < n; i += TP_LB_THREADS) if (i >= 0)
  VF_OWNED(i) d[i] = -g[i];
  double xnorm = sqrt(xx), gnorm = sqrt(gg);
  if (xnorm < 1.0) xnorm = 1.0;
  if (gnorm / xnorm <= C.p.lbfgs_g_eps) {
    ret = 2;  // LBFGS_ALREADY_MINIMIZED
  } else {
    double step = 1.0 / sqrt(gg);   // 1/||d||, d = -g
    double dginit_next = -gg;       // g.d for d = -g
    int end = 0, gbuf = 0;
    k = 1;
    for (;;) {
      VF_OWNED(i) { xp[i] = x[i]; gp[i] = g[i]; }
      // ---------------- line_search_morethuente (lbfgs.hpp:716-937), same scalar logic as tp_lbfgs.cuh
      int ls;
      {
        int count = 0, brackt = 0, stage1 = 1, uinfo = 0;
        double dg, stx, fxx, dgx, sty, fy, dgy, finit, ftest1, dginit, dgtest, width, prev_width, stmin, stmax;
        double stp = step;
        if (stp <= 0.) {
          ls = LB_INVALIDPARAMS;
        } else {
          dginit = dginit_next;
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
              VF_OWNED(i) x[i] = xp[i] + stp * d[i];
              LT0
              __syncthreads();
              vf_eval(C, V, R, true, fx, dg, gg, xx, tid);
              LT(tE)
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
        if (xfinal) VF_OWNED(i) xfinal[i] = xp[i];
        xfinal = nullptr;
        ret = ls;
        break;
      }
      // ||g|| / max(1, ||x||) <= g_epsilon (lbfgs.hpp:1218-1225) without the two square roots and the division
      if (gg <= geps2 * (xx < 1.0 ? 1.0 : xx)) { ret = 0; break; }
      if (C.p.lbfgs_max_iter != 0 && C.p.lbfgs_max_iter < k + 1) { ret = LB_MAXITER; break; }
      // ---------------- new pair into slot `end`, Gram update, coefficients, direction
      const int bound = (VF_M <= k) ? VF_M : k;
      bsum += bound;
      LT0
      gbuf ^= 1;
      if (TM) {
        vf_gram_shift(V, gbuf, tid);
        vf_gram_partial_tm(sm, N, V.tbase, end, tid);
        __syncthreads();
        LT(tG)
        if ((tid >> 5) == V.serial_warp) {
          double* G = sm + V.L.gram + (size_t)gbuf * VF_GRAM;
          vf_gram_finish_tm(V, G, sm + V.L.gram + (size_t)(gbuf ^ 1) * VF_GRAM, end, tid & 31);
          vf_coeffs(G, sm + V.L.ca, sm + V.L.cb, sm + V.L.sc, end, bound, gg, tid & 31);
        }
        __syncthreads();
        LT(tC)
        dginit_next = sc[1];
        vf_direction_tm(sm, N, V.tbase, sc[0], tid);
      } else {
        double* s = S + (size_t)end * ns;
        double* y = Y + (size_t)end * ns;
        VF_OWNED(i) { s[i] = x[i] - xp[i]; y[i] = g[i] - gp[i]; }
        vf_gram_shift(V, gbuf, tid);
        __syncthreads();
        vf_gram_update(V, end, gbuf, tid);
        __syncthreads();
        LT(tG)
        if ((tid >> 5) == V.serial_warp)
          vf_coeffs(sm + V.L.gram + (size_t)gbuf * VF_GRAM, sm + V.L.ca, sm + V.L.cb, sm + V.L.sc, end, bound, gg, tid & 31);
        __syncthreads();
        LT(tC)
        const double cg = sc[0];
        dginit_next = sc[1];
        VF_OWNED(i) {
          double a0 = cg * g[i], a1 = 0.0;
          const double* sp = S + i;
          const double* yp = Y + i;
#pragma unroll 4
          for (int j = 0; j < bound; ++j) {
            a0 = fma(ca[j], sp[0], a0);
            a1 = fma(cb[j], yp[0], a1);
            sp += ns;
            yp += ns;
          }
          d[i] = a0 + a1;
        }
      }
      LT(tD)
      ++k;
      end = (end + 1) & (VF_M - 1);
      step = 1.0;
    }
  }
  if (xfinal) VF_OWNED(i) xfinal[i] = x[i];
#ifdef TP_LBFGS_TIMING
  if (tid == 0 && k > 50 && TP_LBFGS_TIMING > 1)
    printf("[lbfgs] N %d k %d evals %d cycles/iter: total %.0f eval %.0f gram %.0f coeffs %.0f dform %.0f | per eval: partial %.0f blocksum %.0f\n", N, k, evals,
           (double)(clock64() - tT0) / k, (double)tE / k, (double)tG / k, (double)tC / k, (double)tD / k,
           (double)g_tp_t_partial / evals, (double)g_tp_t_sum / evals);
  if (tid == 0) { g_tp_t_partial = 0; g_tp_t_sum = 0; }
  if (tid == 0) {
    atomicAdd(&g_tp_phase[0], (unsigned long long)(clock64() - tT0));
    atomicAdd(&g_tp_phase[1], (unsigned long long)tE);
    atomicAdd(&g_tp_phase[2], (unsigned long long)tG);
    atomicAdd(&g_tp_phase[3], (unsigned long long)tC);
    atomicAdd(&g_tp_phase[4], (unsigned long long)tD);
    atomicAdd(&g_tp_phase[5], (unsigned long long)k);
    atomicAdd(&g_tp_phase[6], (unsigned long long)evals);
  }
#endif
  out.ret = ret;
  out.iters = k;
  out.evals = evals;
  out.reserved = bsum;
  out.fx = fx;
  (void)N;
  __syncthreads();
#undef VF_OWNED
}
