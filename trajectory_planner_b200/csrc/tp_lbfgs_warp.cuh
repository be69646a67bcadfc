// Lean ("team") form of the fused ViGO cost + L-BFGS kernel: a team of NW warps (1 or 4) owns one trajectory.
//
// Same mathematics as lbfgs::lbfgs_optimize (solver/lbfgs.hpp:1024-1349) with m = 16 pairs, the More-Thuente line
// search of tp_lbfgs.cuh verbatim and the reference's convergence / iteration tests.  The block form of round 1
// (tp_lbfgs_fast.cuh, four warps per trajectory) spent 61 % of its samples at block barriers; here nothing ever
// waits on another warp: every reduction is a shuffle tree, every hand-over a __syncwarp, and an SM runs as many
// independent solves as its shared memory holds.
//
// THE ARITHMETIC BELOW IS A SPECIFICATION.  oracle/wform_port.hpp restates it operation for operation on the
// CPU (32 emulated lanes) and tests/ assert that this kernel and that restatement agree BIT FOR BIT (return code,
// iteration / evaluation counts, control points).  Any change to an expression, an accumulation order or a
// reduction tree here must be mirrored there.  The rules that make that possible:
//   * compiled with --fmad=false: a fused multiply-add happens exactly where fma() is written;
//   * optimised element i (x[i], i = 0..n-1) is owned by thread i % (32 NW); a thread visits its elements in
//     increasing order and keeps private running sums;
//   * sums over a warp's lanes are fixed shuffle trees (wf_sum4, the 16-lane butterfly of wf_coeffs); the warps'
//     results are combined in warp order, ((w0 + w1) + w2) + w3; the Gram dot products are serial per lane over the
//     warp's contiguous chunk of elements with four interleaved partial sums (i mod 4), combined (a0 + a1) + (a2 + a3);
//   * divisions and square roots are IEEE (FP64 always is on the device).
//
// Direction: the two-loop recursion (lbfgs.hpp:1293-1316) is evaluated in COEFFICIENT SPACE on Gram blocks of
//   s_i.y_j,  y_i.y_j,  s_i.g,  y_i.g          (slot-indexed, circular)
//   d = sum_j ca_j s_j + sum_j cb_j y_j + cg g.
// Per iteration only the column of the newest pair and s_i.g / y_i.g change: 4 x 16 dot products of length n.  Lane l
// owns history ROW l (s_0..s_15, y_0..y_15) for this step and runs its two dot products (with y_new and with g)
// serially over the elements — 64 dot products, no cross-lane reduction at all — instead of 2 x 16 SEQUENTIAL
// reductions; the two 16-step triangular recurrences run on register shuffles, four steps per round (wf_coeffs).
// The history is stored element-major (wf_layout), and the direction pass also opens the next iteration (wf_direction).
// WHY A TEAM: one warp issues an instruction every ~4 cycles (dependent FP64 / shared-memory chains), and shared memory
// (2 x 16 history rows per trajectory) limits an SM to ~5 trajectories — with one warp each, most issue slots idle
// (measured: 14.5 k cycles per iteration).  Four warps share the element-parallel phases (evaluation, Gram dot
// products, direction); the serial phases were made cheap enough not to dominate (~1.1 k of ~7.5 k cycles per iteration).
// INSTRUCTION COUNT, LATENCY PER INSTRUCTION and REGISTERS are the constraints (measured): a warp
// issues one instruction per ~9 cycles, so an iteration costs what it issues; compiled for three teams per SM (168
// registers) the kernel spilled 36 loads per warp and iteration and ran slower than two spill-free teams (TP_TEAM_BLOCKS).  Hence: constants and the per-solve context live in shared
// memory (a `const VigoConst&` reaching a noinline function turns every access into a generic global load), every
// rare path is out of line, small vector loops stay rolled.
#pragma once
#include "tp_lbfgs.cuh"
#include "tp_lbfgs_fast.cuh"   // vf_height_term (same height-barrier terms)

#define WF_M 16            // history pairs (bsplineTraj.cpp:697)
#define WF_GS 18           // Gram row stride (even: rows are read with 128-bit loads)
#define WF_PAIRS_SM 64     // guide pairs staged in shared memory (more -> read through L2)
// Gram region (doubles).  Lane s <-> SLOT s.
//   A [s][c] = (s_s . y_c) / ys_s   if pair s is OLDER than pair c, else 0     (first loop: row-scaled)
//   Bt[c][s] = (s_s . y_c) / ys_c   if pair s is OLDER than pair c, else 0     (second loop: column-scaled, transposed)
//   YY[s][c] = y_s . y_c
#define WF_G_A 0
#define WF_G_BT (WF_M * WF_GS)
#define WF_G_YY (2 * WF_M * WF_GS)
#define WF_G_SG (3 * WF_M * WF_GS)
#define WF_G_YG (WF_G_SG + WF_M)
#define WF_G_INV (WF_G_YG + WF_M)
#define WF_G_TB (WF_G_INV + WF_M)
#define WF_G_YS (WF_G_TB + WF_M)      // ys, yy, 1/yy of the newest pair
#define WF_GRAM (WF_G_YS + 4)
#define WF_CTX 32          // doubles reserved for WfShared in front of the control points

// per-solve context + evaluation constants, in shared memory in front of the control points
struct WfShared {
  double icts, k2, gv_c, ga_c;            // 1/ctrl_pt_ts, 1/ctrl_pt_ts^2, gradient factors of the feasibility terms
  double dth, dist_a, dist_b, dist_c;     // distance_threshold and the quadratic piece of getDistanceCost (:835)
  double unc;                             // uncertain_aware_factor
  double w_dist, w_smooth, w_feas, w_dyn;
  double f_const;                         // w_feas x the feasibility terms made of fixed control points only
  const GuidePair* pairs;                 // global list of this trajectory
  const int* head;                        // global per-control-point list heads
  const double* dyn_pos;
  const double* dyn_vel;
  const double* dyn_size;
  int N, n, n_pairs, pairs_in_sm, plan_in_z, n_dyn;
};
static_assert(sizeof(WfShared) <= WF_CTX * 8, "WfShared outgrew its shared-memory slot");

#define WF_HS 34           // history: doubles per ELEMENT (32 rows + 2 of padding, see wf_layout)
struct WfLayout {   // offsets in doubles from the worker's solver base (context, then control points)
  int cp, g, xp, gp, d, H, gram, ca, cb, sc, red, gpart, pair, pstart, total;
};
#define WF_NW_MAX 4
__host__ __device__ inline WfLayout wf_layout(int N) {
  WfLayout L;
  const int n = 3 * (N - 2 * TP_DEGREE);
  const int nn = n > 0 ? n : 0;
  const int ne = (nn + 3) & ~3;   // vectors are zero-padded to a multiple of 4 (the row dot products run 4 elements a step)
  int o = WF_CTX;
  L.cp = o; o += 3 * N + (N & 1);
  L.g = o; o += ne;
  L.xp = o; o += ne;
  L.gp = o; o += ne;
  L.d = o; o += ne;
  // History, ELEMENT-major: H[i * WF_HS + r], r = 0..15 <-> s_slot, r = 16..31 <-> y_slot.  The Gram dot products (lane
  // <-> row, serial over the elements) read 32 consecutive doubles per element; the direction (thread <-> element) reads
  // its element's 32 rows with sixteen 128-bit loads at constant offsets - no per-row address arithmetic - and the
  // stride of 34 doubles (272 B = 16 mod 128) spreads the eight lanes of a quarter-warp over all 32 banks.
  L.H = o; o += ne * WF_HS;       // (o is even here: 16-byte aligned rows)
  L.gram = o; o += WF_GRAM;
  L.ca = o; o += WF_M;
  L.cb = o; o += WF_M;
  L.sc = o; o += 8;
  L.red = o; o += WF_NW_MAX * 4;           // per-warp results of the evaluation's four sums
  L.gpart = o; o += WF_NW_MAX * 32 * 2;    // per-warp partial Gram dot products
  L.pair = o; o += WF_PAIRS_SM * 7;
  L.pstart = o; o += (N + 2) / 2;   // (N + 1) ints
  L.total = o;
  return L;
}

#define WF_FULL 0xffffffffu
#ifdef WF_COEFFS_ROLLED
#define WF_COEFFS_LOOP _Pragma("unroll 1")
#else
#define WF_COEFFS_LOOP _Pragma("unroll")
#endif
// the noinline pieces below get plain pointers: tell the compiler they point into shared memory (LDS / STS instead of
// generic loads and 64-bit address arithmetic)
#define WF_ASSUME_SHARED(p) __builtin_assume(__isShared(p))
// team-wide synchronisation: the team is the whole thread block when NW > 1
template <int NW>
__device__ __forceinline__ void wf_sync() {
  if (NW == 1) __syncwarp();
  else __syncthreads();
}

// Sums of four per-lane values over the warp, every lane receives all four totals.  Recursive halving on the first
// two levels (xor 16: a lane whose bit 4 is clear keeps v0, v1 and adds the partner's, a lane whose bit is set keeps
// v2, v3; xor 8: bit 3 clear keeps the first of the two, set the second), then a butterfly 4, 2, 1 inside each group of
// eight lanes (group k = lanes 8k..8k+7 holds the total of value k), then four broadcasts: 10 shuffles instead of 20.
__device__ __forceinline__ void wf_sum4(double (&v)[4], int lane) {
  const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
  const double k0 = up16 ? v[2] : v[0], k1 = up16 ? v[3] : v[1];
  const double s0 = up16 ? v[0] : v[2], s1 = up16 ? v[1] : v[3];
  const double a0 = k0 + __shfl_xor_sync(WF_FULL, s0, 16);
  const double a1 = k1 + __shfl_xor_sync(WF_FULL, s1, 16);
  double b = (up8 ? a1 : a0) + __shfl_xor_sync(WF_FULL, up8 ? a0 : a1, 8);
  b = b + __shfl_xor_sync(WF_FULL, b, 4);
  b = b + __shfl_xor_sync(WF_FULL, b, 2);
  b = b + __shfl_xor_sync(WF_FULL, b, 1);
  v[0] = __shfl_sync(WF_FULL, b, 0);
  v[1] = __shfl_sync(WF_FULL, b, 8);
  v[2] = __shfl_sync(WF_FULL, b, 16);
  v[3] = __shfl_sync(WF_FULL, b, 24);
}

// signed excess over the +-1 box of getFeasibilityCost (bsplineTraj.cpp:955-956: maxVel = maxAcc = 1.0 hard-coded):
// v - 1 above the box, v + 1 below it, +0 inside
__device__ __forceinline__ double wf_excess(double v) {
  const double e = fabs(v) - 1.0;
  return e > 0.0 ? copysign(e, v) : 0.0;
}

// The feasibility terms of getFeasibilityCost that involve FIXED control points only (velocity terms 0, 1, N-3, N-2,
// acceleration terms 0, N-3): constant during one optimize().  Per axis the six terms are accumulated in that order;
// the axes are combined (x + y) + z.  Returns w_feas x the sum to every lane.
__device__ __noinline__ double wf_const_terms(const double* cp, int N, double icts, double k2, double w_feas, int lane) {
  WF_ASSUME_SHARED(cp);
  double t = 0.0;
  if (lane < 3) {
    const int a = lane;
    const int vi[4] = {0, 1, N - 3, N - 2};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = 3 * vi[q] + a;
      const double ev = wf_excess((cp[e + 3] - cp[e]) * icts);
      t = fma(ev * ev, k2, t);
    }
    const int ai[2] = {0, N - 3};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int e = 3 * ai[q] + a;
      const double ea = wf_excess((cp[e + 6] - 2 * cp[e + 3] + cp[e]) * k2);
      t = fma(ea, ea, t);
    }
  }
  const double tx = __shfl_sync(WF_FULL, t, 0), ty = __shfl_sync(WF_FULL, t, 1), tz = __shfl_sync(WF_FULL, t, 2);
  return w_feas * ((tx + ty) + tz);
}

// Per-solve set-up (once per optimize()): the context in front of the control points, the guide pairs staged as a CSR
// list (by control point, append order) in shared memory, the constant cost terms.  `base` = solver base of the worker.
template <int NW>
__device__ __noinline__ void wf_setup(const VigoConst& C, double* base, int N, const GuidePair* pairs, const int* head,
                                      int n_pairs, double w_dist, double w_dyn, int n_dyn, const double* dyn_pos,
                                      const double* dyn_vel, const double* dyn_size, int tid) {
  WF_ASSUME_SHARED(base);
  constexpr int P = 32 * NW;
  const int lane = tid;
  const WfLayout L = wf_layout(N);
  WfShared* X = reinterpret_cast<WfShared*>(base);
  const double icts = 1.0 / C.p.ctrl_pt_ts, k2 = C.ts_inv_sqr;
  const int in_sm = n_pairs <= WF_PAIRS_SM;
  wf_sync<NW>();
  if (lane == 0) {
    X->icts = icts; X->k2 = k2;
    X->gv_c = 2.0 * icts * k2;   // d/dc of (v -+ 1)^2 k2
    X->ga_c = 2.0 * k2;          // d/dc of (a -+ 1)^2
    X->dth = C.p.dthresh; X->dist_a = C.dist_a; X->dist_b = C.dist_b; X->dist_c = C.dist_c;
    X->unc = C.p.uncertain_factor;
    X->w_dist = w_dist; X->w_smooth = C.p.w_smooth; X->w_feas = C.p.w_feas; X->w_dyn = w_dyn;
    X->pairs = pairs; X->head = head; X->dyn_pos = dyn_pos; X->dyn_vel = dyn_vel; X->dyn_size = dyn_size;
    X->N = N; X->n = 3 * (N - 2 * TP_DEGREE); X->n_pairs = n_pairs; X->pairs_in_sm = in_sm;
    X->plan_in_z = C.p.plan_in_z; X->n_dyn = n_dyn;
  }
  // zero g / xp / gp / d (their padding must be zero), the history rows, the Gram blocks and the coefficients
  // (g .. sc are contiguous in the layout)
#pragma unroll 2
  for (int e = lane; e < L.pair - L.g; e += P) (base + L.g)[e] = 0.0;
  if (tid < 32) {
    const double fc = wf_const_terms(base + L.cp, N, icts, k2, C.p.w_feas, lane);
    if (lane == 0) X->f_const = fc;
  }
  if (in_sm) {
    int* pstart = reinterpret_cast<int*>(base + L.pstart);
    double* ps = base + L.pair;
    for (int c = lane; c < N; c += P) {
      int cnt = 0;
      for (int gi = head[c]; gi >= 0; gi = pairs[gi].next) ++cnt;
      pstart[c + 1] = cnt;
    }
    if (lane == 0) pstart[0] = 0;
    wf_sync<NW>();
    if (lane == 0)
      for (int c = 0; c < N; ++c) pstart[c + 1] += pstart[c];
    wf_sync<NW>();
    for (int c = lane; c < N; c += P) {
      int w = pstart[c];
      for (int gi = head[c]; gi >= 0;) {
        const GuidePair& pr = pairs[gi];
        double* q = ps + 7 * w;
        q[0] = pr.p[0]; q[1] = pr.p[1]; q[2] = pr.p[2];
        q[3] = pr.v[0]; q[4] = pr.v[1]; q[5] = pr.v[2];
        q[6] = pr.unknown ? 1.0 : 0.0;
        ++w;
        gi = pr.next;
      }
    }
  }
  wf_sync<NW>();
}

// one (pair, control point, axis) term of getDistanceCost, bsplineTraj.cpp:839-895 (same expressions as vf_pair_term,
// constants from the shared-memory context)
__device__ __forceinline__ void wf_pair_term(double dth, double qa, double qb, double qc, double unc, bool plan_in_z,
                                             double cx, double cy, double cz, const double* pr, int a, double& grad, double& cost) {
  const double vx = pr[3], vy = pr[4], vz = pr[5];
  bool unk = pr[6] != 0.0;
  const double dist = fma(cx - pr[0], vx, fma(cy - pr[1], vy, (cz - pr[2]) * vz));
  const double e = dth - dist;
  double va = a == 0 ? vx : (a == 1 ? vy : vz);
  if (!plan_in_z && a == 2) va = 0.0;
  double costTemp, gt;
  if (e <= -dth) {                       // far beyond the plane: (-e)^3, not scaled by the uncertain factor (:852-861)
    costTemp = -(e * e) * e;
    gt = 3.0 * (e * e);
    unk = false;
  } else if (e > 0 && e <= dth) {        // (:862-878)
    costTemp = (e * e) * e;
    gt = -3.0 * (e * e);
  } else if (e >= dth) {                 // (:879-894)
    costTemp = fma(fma(qa, e, qb), e, qc);
    gt = -fma(2.0 * qa, e, qb);
  } else {
    return;
  }
  if (unk) { costTemp *= unc; gt *= unc; }
  grad = fma(gt, va, grad);
  cost += costTemp;
}

// rarely-taken cost terms of one (control point, axis), out of line (code size): guide pairs that did not fit the
// shared-memory stage (read from the global linked list), the height barrier, dynamic obstacles.  Results are ADDED by
// the caller: r[0] gradient / r[1] cost of the distance term, r[2] / r[3] of the dynamic-obstacle term.
__device__ __noinline__ void wf_rare_terms(const VigoConst& C, const double* base, int c, int a, double (&r)[4]) {
  WF_ASSUME_SHARED(base);
  const WfShared* X = reinterpret_cast<const WfShared*>(base);
  const double* cp = base + WF_CTX;
  double gd = 0.0, cD = 0.0, go = 0.0, cO = 0.0;
  if (!X->pairs_in_sm) {
    const double cx = cp[3 * c], cy = cp[3 * c + 1], cz = cp[3 * c + 2];
    for (int gi = X->head[c]; gi >= 0;) {
      const GuidePair& pr = X->pairs[gi];
      const double q[7] = {pr.p[0], pr.p[1], pr.p[2], pr.v[0], pr.v[1], pr.v[2], pr.unknown ? 1.0 : 0.0};
      wf_pair_term(X->dth, X->dist_a, X->dist_b, X->dist_c, X->unc, X->plan_in_z != 0, cx, cy, cz, q, a, gd, cD);
      gi = pr.next;
    }
  }
  if (X->plan_in_z) vf_height_term(C, cp[3 * c + 2], a, gd, cD);
  if (X->n_dyn > 0) {
    EvalCtx E;
    E.N = X->N; E.n = X->n; E.cp = const_cast<double*>(cp); E.pairs = X->pairs; E.head = X->head;
    E.w_dist = X->w_dist; E.w_dyn = X->w_dyn; E.n_dyn = X->n_dyn; E.dyn_pos = X->dyn_pos; E.dyn_vel = X->dyn_vel;
    E.dyn_size = X->dyn_size;
    dynamic_terms(C, E, c, a, go, cO);
  }
  r[0] = gd; r[1] = cD; r[2] = go; r[3] = cO;
}

#ifdef TP_WF_TIMING
__device__ unsigned long long g_wf_phase[8];   // whole-batch cycle totals: total, eval, gram, coeffs, direction, iterations, evals
#endif

// costFunction (bsplineTraj.cpp:802-821) at the control points in shared memory: writes the gradient, returns
// {f, g.d, g.g, x.x} to every lane.  Same terms as getSmoothnessCost / getFeasibilityCost / getDistanceCost /
// getDynamicObstacleCost.  Throughput form: the 7-point stencil of an element is read once and turned into first /
// second / third differences (velocity, acceleration and jerk share them), divisions by the constant control-point
// timestep become multiplications by its reciprocal, the +-1 feasibility branches become a clamp.  Every cost term is
// owned by exactly one element: element (c, a) owns jerk / velocity / acceleration term c; the elements of the first
// optimised point c = 3 additionally own the terms 0..2 that reach into the fixed points; terms made of fixed points
// only are X->f_const.
struct WfEv { double v[4]; };   // returned BY VALUE (registers); an array by reference would travel through the stack
template <int NW>
__device__ __noinline__ WfEv wf_eval(const VigoConst& C, double* base, int with_d, int tid) {
  WF_ASSUME_SHARED(base);
  double out[4];
  constexpr int P = 32 * NW;
  const int lane = tid & 31;
  const WfShared* X = reinterpret_cast<const WfShared*>(base);
  const int N = X->N, n = X->n;
  const WfLayout L = wf_layout(N);
  const double* cp = base + L.cp;
  double* g = base + L.g;
  const double* d = base + L.d;
  const int* pstart = reinterpret_cast<const int*>(base + L.pstart);
  const double* ps = base + L.pair;
  const double icts = X->icts, k2 = X->k2, gv_c = X->gv_c, ga_c = X->ga_c;
  const double w_dist = X->w_dist, w_smooth = X->w_smooth, w_feas = X->w_feas, w_dyn = X->w_dyn;
  const bool in_sm = X->pairs_in_sm != 0, plan_in_z = X->plan_in_z != 0;
  const bool rare = !in_sm || plan_in_z || X->n_dyn > 0;
  double sD = 0, sS = 0, sF = 0, sO = 0, dg = 0, gg = 0, xx = 0;
#pragma unroll 1
  for (int i = tid; i < n; i += P) {
    const int e = i + 3 * TP_DEGREE;
    const int c = e / 3, a = e - 3 * c;
    const double pm3 = cp[e - 9], pm2 = cp[e - 6], pm1 = cp[e - 3], p0 = cp[e], p1 = cp[e + 3], p2 = cp[e + 6], p3 = cp[e + 9];
    // first differences d_k = p_(k+1) - p_k, second a_k = d_(k+1) - d_k (acceleration term c+k uses points c+k..c+k+2),
    // third j_k = a_(k+1) - a_k (jerk term c+k uses points c+k..c+k+3)
    const double dm3 = pm2 - pm3, dm2 = pm1 - pm2, dm1 = p0 - pm1, d0 = p1 - p0, d1 = p2 - p1, d2 = p3 - p2;
    const double am3 = dm2 - dm3, am2 = dm1 - dm2, am1 = d0 - dm1, a0 = d1 - d0, a1 = d2 - d1;
    const double jm3 = am2 - am3, jm2 = am1 - am2, jm1 = a0 - am1, j0 = a1 - a0;
    // gradient of the jerk terms c-3 .. c (getSmoothnessCost, :938-947): d/dp0 = 2 (jm3 - 3 jm2 + 3 jm1 - j0)
    const double gs = 2.0 * ((jm3 - j0) + 3.0 * (jm1 - jm2));
    // feasibility (getFeasibilityCost, :961-995)
    const double evm = wf_excess(dm1 * icts), ev0 = wf_excess(d0 * icts);
    const double eam2 = wf_excess(am2 * k2), eam1 = wf_excess(am1 * k2), ea0 = wf_excess(a0 * k2);
    const double gf = fma(gv_c, evm - ev0, ga_c * ((eam2 + ea0) - 2.0 * eam1));
    // cost terms owned by this element
    sS = fma(j0, j0, sS);
    sF = fma(ev0 * ev0, k2, sF);
    sF = fma(ea0, ea0, sF);
    if (c == TP_DEGREE) {
      sS = fma(jm3, jm3, sS);
      sS = fma(jm2, jm2, sS);
      sS = fma(jm1, jm1, sS);
      sF = fma(evm * evm, k2, sF);
      sF = fma(eam2, eam2, sF);
      sF = fma(eam1, eam1, sF);
    }
    // distance to the guide planes (getDistanceCost, :839-930)
    double gd = 0.0, cD = 0.0, go = 0.0, cO = 0.0;
    if (in_sm) {
      const int q0 = pstart[c], q1 = pstart[c + 1];
      if (q0 < q1) {
        const double cx = cp[3 * c], cy = cp[3 * c + 1], cz = cp[3 * c + 2];
        for (int q = q0; q < q1; ++q)
          wf_pair_term(X->dth, X->dist_a, X->dist_b, X->dist_c, X->unc, plan_in_z, cx, cy, cz, ps + 7 * q, a, gd, cD);
      }
    }
    if (rare) {
      double r[4];
      wf_rare_terms(C, base, c, a, r);
      gd += r[0]; cD += r[1]; go += r[2]; cO += r[3];
    }
    if (a == 0) { sD += cD; sO += cO; }
    const double gv = fma(w_dist, gd, fma(w_smooth, gs, fma(w_feas, gf, w_dyn * go)));
    g[i] = gv;
    gg = fma(gv, gv, gg);
    xx = fma(p0, p0, xx);
    if (with_d) dg = fma(gv, d[i], dg);
  }
  out[0] = w_dist * sD + w_smooth * sS + w_feas * sF + w_dyn * sO;
  out[1] = dg;
  out[2] = gg;
  out[3] = xx;
  wf_sum4(out, lane);
  if (NW > 1) {
    // the warps' results, combined in warp order
    double* rb = base + L.red;
    if (lane == 0) {
      double* r = rb + 4 * (tid >> 5);
      r[0] = out[0]; r[1] = out[1]; r[2] = out[2]; r[3] = out[3];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      double v = rb[k];
#pragma unroll
      for (int w = 1; w < NW; ++w) v = v + rb[4 * w + k];
      out[k] = v;
    }
  }
  WfEv r;
  r.v[0] = out[0] + X->f_const; r.v[1] = out[1]; r.v[2] = out[2]; r.v[3] = out[3];
  return r;
}

// New pair (s, y) = (x - xp, g - gp) into slot `slot` (y also into the gp vector, which is dead from here on), then the
// fresh Gram entries.  Lane l owns history row l (rows 0..15 = s_0..s_15, rows 16..31 = y_0..y_15; rows of pairs not
// stored yet are zero) and runs its two dot products (with y_new, with g) serially over the zero-padded elements, four
// interleaved partial sums each.  Lane l < 16 (row s_l): A[l][slot] = (s_l.y_new)/ys_l, Bt[slot][l] = (s_l.y_new)/ys_new,
// Sg[l] = s_l.g; the new pair's own row / column entries (it is older than nobody) are zeroed; ys_new = s_new.y_new and
// its reciprocal are published.  Lane 16 + j (row y_j): YY[j][slot] = YY[slot][j] = y_j.y_new, Yg[j] = y_j.g.
// With NW > 1 every warp covers a contiguous chunk of the elements for all 32 rows; only the serial warp returns with
// the Gram blocks written (the others return after handing in their partial sums).
template <int NW>
__device__ __noinline__ void wf_gram_update(double* base, int N_, int slot, int serial_warp, int tid) {
  WF_ASSUME_SHARED(base);
  constexpr int P = 32 * NW;
  const int lane = tid & 31, warp = tid >> 5;
  const WfLayout L = wf_layout(N_);
  const int n = 3 * (N_ - 2 * TP_DEGREE);
  const double* x = base + L.cp + 3 * TP_DEGREE;
  const double* g = base + L.g;
  const double* xp = base + L.xp;
  double* gp = base + L.gp;
  double* H = base + L.H;
  double* G = base + L.gram;
#pragma unroll 1
  for (int i = tid; i < n; i += P) {
    const double yi = g[i] - gp[i];
    H[i * WF_HS + slot] = x[i] - xp[i];
    H[i * WF_HS + WF_M + slot] = yi;
    gp[i] = yi;
  }
  wf_sync<NW>();
  // this warp's contiguous chunk of the zero-padded elements (a multiple of 4 elements)
  const double* row = H + lane;
  const int n4 = (n + 3) & ~3;
  const int chunk = ((n4 / 4 + NW - 1) / NW) * 4;
  const int i0 = warp * chunk;
  const int i1 = i0 + chunk < n4 ? i0 + chunk : n4;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
#pragma unroll 1
  for (int i = i0; i < i1; i += 4) {
    const double2 y01 = *reinterpret_cast<const double2*>(gp + i), y23 = *reinterpret_cast<const double2*>(gp + i + 2);
    const double2 g01 = *reinterpret_cast<const double2*>(g + i), g23 = *reinterpret_cast<const double2*>(g + i + 2);
    const double r0 = row[i * WF_HS], r1 = row[(i + 1) * WF_HS], r2 = row[(i + 2) * WF_HS], r3 = row[(i + 3) * WF_HS];
    a0 = fma(r0, y01.x, a0); b0 = fma(r0, g01.x, b0);
    a1 = fma(r1, y01.y, a1); b1 = fma(r1, g01.y, b1);
    a2 = fma(r2, y23.x, a2); b2 = fma(r2, g23.x, b2);
    a3 = fma(r3, y23.y, a3); b3 = fma(r3, g23.y, b3);
  }
  double dy = (a0 + a1) + (a2 + a3), dgv = (b0 + b1) + (b2 + b3);
  if (NW > 1) {
    // the warps' partial dot products, combined in warp order by the serial warp
    double* gpart = base + L.gpart;
    *reinterpret_cast<double2*>(gpart + 2 * (warp * 32 + lane)) = make_double2(dy, dgv);
    __syncthreads();
    if (warp != serial_warp) return;
    double2 v = *reinterpret_cast<const double2*>(gpart + 2 * lane);
#pragma unroll
    for (int w = 1; w < NW; ++w) {
      const double2 q = *reinterpret_cast<const double2*>(gpart + 2 * (w * 32 + lane));
      v.x = v.x + q.x;
      v.y = v.y + q.y;
    }
    dy = v.x; dgv = v.y;
  }
  // ys = s_new.y_new sits in lane `slot`, yy = y_new.y_new in lane 16 + slot; their reciprocals are computed once
  const double dsel = __shfl_sync(WF_FULL, dy, lane < WF_M ? slot : WF_M + slot);   // lanes < 16: ys, lanes >= 16: yy
  const double rc = 1.0 / dsel;
  if (lane < WF_M) {
    const double inv_new = rc;                                  // 1 / ys_new
    const bool self = lane == slot;
    const double inv_l = self ? inv_new : G[WF_G_INV + lane];
    G[WF_G_A + slot * WF_GS + lane] = 0.0;
    G[WF_G_BT + lane * WF_GS + slot] = 0.0;
    // (the only cell both groups of stores touch is the diagonal, which gets 0 from either)
    G[WF_G_A + lane * WF_GS + slot] = self ? 0.0 : dy * inv_l;
    G[WF_G_BT + slot * WF_GS + lane] = self ? 0.0 : dy * inv_new;
    G[WF_G_SG + lane] = dgv;
    if (self) {
      G[WF_G_INV + slot] = inv_new;
      G[WF_G_YS] = dy;
    }
  } else {
    const int j = lane - WF_M;
    G[WF_G_YY + j * WF_GS + slot] = dy;
    G[WF_G_YY + slot * WF_GS + j] = dy;
    G[WF_G_YG + j] = dgv;
    if (j == slot) {
      G[WF_G_YS + 1] = dy;
      G[WF_G_YS + 2] = rc;                                      // 1 / yy_new
    }
  }
  __syncwarp();
}

// Two-loop recursion in coefficient space (lbfgs.hpp:1293-1316).  Lane s (and lane s + 16, redundantly) <-> the pair in
// SLOT s; its age is (newest - s) & 15 (0 = newest).  Entries of pairs not stored yet are zero and so is their 1/ys,
// which makes their steps no-ops.  Both triangular recurrences run on scaled quantities (r_s / ys_s, (y_s.d) / ys_s):
// per step one register shuffle and one FMA on the dependent chain, the matrix entry comes from the pre-scaled,
// pre-masked blocks A / Bt (wf_gram_update) through one shared-memory load with no arithmetic behind it.
//   first loop, newest -> oldest (step t handles slot st = newest - t): alpha_st = r_st (final by then);
//       every lane: r_s = fma(-alpha_st, A[s][st], r_s)
//   middle: t_s = yg_s + sum_c alpha_c YY[s][c]  (slot order, four interleaved sums c mod 4);  b_s = (-gamma t_s) / ys_s
//   second loop, oldest -> newest (slot st = newest - t, t = 15..0): c_st = alpha_st - b_st (final by then);
//       every lane: b_s = fma(c_st, Bt[s][st], b_s)
//   ca[s] = alpha_s - b_s, cb[s] = -gamma alpha_s, cg = -gamma, g.d = sum_s fma(ca_s, sg_s, cb_s yg_s) - gamma g.g
// Writes ca[], cb[] (coefficients of s_slot, y_slot), sc[0] = coefficient of g, sc[1] = g.d.
__device__ __noinline__ void wf_coeffs(double* G, double* ca, double* cb, double* sc, int newest, double gg, int lane) {
  WF_ASSUME_SHARED(G);
  WF_ASSUME_SHARED(ca);
  WF_ASSUME_SHARED(cb);
  WF_ASSUME_SHARED(sc);
  const int s = lane & (WF_M - 1);
  double* tb = G + WF_G_TB;
  const double ys = G[WF_G_YS], inv_yy = G[WF_G_YS + 2];
  const double gamma = ys * inv_yy;   // ys/yy (lbfgs.hpp:1305)
  const double inv_s = G[WF_G_INV + s];
  const double sg_s = G[WF_G_SG + s], yg_s = G[WF_G_YG + s];
  const double* arow = G + WF_G_A + s * WF_GS;
  const double* brow = G + WF_G_BT + s * WF_GS;
  const double* yyrow = G + WF_G_YY + s * WF_GS;
  // ---- first loop
  double rr = -sg_s * inv_s;
#ifdef WF_COEFFS_STEPWISE
#pragma unroll
  for (int t = 0; t < WF_M; ++t) {
    const int st = (newest - t) & (WF_M - 1);
    const double ala = __shfl_sync(WF_FULL, rr, st);
    rr = fma(-ala, arow[st], rr);
  }
#else
  // Four steps per round of shuffles: the four rows whose alphas become final in steps 4k..4k+3 are broadcast at once
  // and every lane redoes their in-block updates (six FMAs, the operands and the order the owning lanes use), so the
  // dependent chain per four steps is ONE shuffle round trip + 4 FMAs instead of four round trips.  Every row still
  // receives its 16 FMAs in step order with the same operands: the results are bit-identical to the step-by-step loop.
  const double* Ab = G + WF_G_A;
WF_COEFFS_LOOP
  for (int k = 0; k < WF_M / 4; ++k) {
    const int s0 = (newest - 4 * k) & (WF_M - 1), s1 = (s0 - 1) & (WF_M - 1), s2 = (s0 - 2) & (WF_M - 1), s3 = (s0 - 3) & (WF_M - 1);
    const double a0 = __shfl_sync(WF_FULL, rr, s0);
    double a1 = __shfl_sync(WF_FULL, rr, s1), a2 = __shfl_sync(WF_FULL, rr, s2), a3 = __shfl_sync(WF_FULL, rr, s3);
    const double m0 = arow[s0], m1 = arow[s1], m2 = arow[s2], m3 = arow[s3];
    const double e10 = Ab[s1 * WF_GS + s0];
    const double e20 = Ab[s2 * WF_GS + s0], e21 = Ab[s2 * WF_GS + s1];
    const double e30 = Ab[s3 * WF_GS + s0], e31 = Ab[s3 * WF_GS + s1], e32 = Ab[s3 * WF_GS + s2];
    a1 = fma(-a0, e10, a1);
    a2 = fma(-a0, e20, a2);
    a3 = fma(-a0, e30, a3);
    rr = fma(-a0, m0, rr);
    a2 = fma(-a1, e21, a2);
    a3 = fma(-a1, e31, a3);
    rr = fma(-a1, m1, rr);
    a3 = fma(-a2, e32, a3);
    rr = fma(-a2, m2, rr);
    rr = fma(-a3, m3, rr);
  }
#endif
  const double al = rr;
  // ---- middle
  __syncwarp();
  if (lane < WF_M) tb[s] = al;
  __syncwarp();
  double t0 = yg_s, t1 = 0.0, t2 = 0.0, t3 = 0.0;
#pragma unroll
  for (int c = 0; c < WF_M; c += 4) {
    const double2 al01 = *reinterpret_cast<const double2*>(tb + c), al23 = *reinterpret_cast<const double2*>(tb + c + 2);
    const double2 yy01 = *reinterpret_cast<const double2*>(yyrow + c), yy23 = *reinterpret_cast<const double2*>(yyrow + c + 2);
    t0 = fma(al01.x, yy01.x, t0);
    t1 = fma(al01.y, yy01.y, t1);
    t2 = fma(al23.x, yy23.x, t2);
    t3 = fma(al23.y, yy23.y, t3);
  }
  // ---- second loop
  double bacc = (-gamma * ((t0 + t1) + (t2 + t3))) * inv_s;
#ifdef WF_COEFFS_STEPWISE
#pragma unroll
  for (int t = WF_M - 1; t >= 0; --t) {
    const int st = (newest - t) & (WF_M - 1);
    const double cst = __shfl_sync(WF_FULL, al - bacc, st);
    bacc = fma(cst, brow[st], bacc);
  }
#else
  // same four-step blocking, oldest block first: b of the block's four rows by shuffle, their alphas from tb[]
  const double* Bb = G + WF_G_BT;
WF_COEFFS_LOOP
  for (int k = 0; k < WF_M / 4; ++k) {
    const int s0 = (newest - (WF_M - 1) + 4 * k) & (WF_M - 1), s1 = (s0 + 1) & (WF_M - 1), s2 = (s0 + 2) & (WF_M - 1), s3 = (s0 + 3) & (WF_M - 1);
    const double b0 = __shfl_sync(WF_FULL, bacc, s0);
    double b1 = __shfl_sync(WF_FULL, bacc, s1), b2 = __shfl_sync(WF_FULL, bacc, s2), b3 = __shfl_sync(WF_FULL, bacc, s3);
    const double l0 = tb[s0], l1 = tb[s1], l2 = tb[s2], l3 = tb[s3];
    const double m0 = brow[s0], m1 = brow[s1], m2 = brow[s2], m3 = brow[s3];
    const double e10 = Bb[s1 * WF_GS + s0];
    const double e20 = Bb[s2 * WF_GS + s0], e21 = Bb[s2 * WF_GS + s1];
    const double e30 = Bb[s3 * WF_GS + s0], e31 = Bb[s3 * WF_GS + s1], e32 = Bb[s3 * WF_GS + s2];
    const double c0 = l0 - b0;
    b1 = fma(c0, e10, b1);
    b2 = fma(c0, e20, b2);
    b3 = fma(c0, e30, b3);
    bacc = fma(c0, m0, bacc);
    const double c1 = l1 - b1;
    b2 = fma(c1, e21, b2);
    b3 = fma(c1, e31, b3);
    bacc = fma(c1, m1, bacc);
    const double c2 = l2 - b2;
    b3 = fma(c2, e32, b3);
    bacc = fma(c2, m2, bacc);
    const double c3 = l3 - b3;
    bacc = fma(c3, m3, bacc);
  }
#endif
  const double aa = al - bacc;
  const double bb = -gamma * al;
  if (lane < WF_M) {
    ca[s] = aa;
    cb[s] = bb;
  }
  // g.d = sum aa_s s_s.g + sum bb_s y_s.g - gamma g.g
  double part = fma(aa, sg_s, bb * yg_s);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) part = part + __shfl_xor_sync(WF_FULL, part, o);
  if (lane == 0) {
    sc[0] = -gamma;
    sc[1] = part - gamma * gg;
  }
  __syncwarp();
}

// d = cg g + sum_j ca[j] s_j + cb[j] y_j over all 16 slots (coefficients of slots not stored yet are zero), four
// running sums per element: s_0..7, s_8..15, y_0..7, y_8..15, combined (a0 + a1) + (a2 + a3).
// The same pass opens the next iteration for its element (it holds d_i and g_i in registers and owns x_i): xp = x, gp = g
// and the FIRST trial point of the next line search, x = xp + 1.0 d (lbfgs.hpp:1333 sets step = 1.0; the clamps of
// line_search_morethuente leave a first trial of 1.0 alone) — the caller skips its own copy and update loops.
template <int NW>
__device__ __noinline__ void wf_direction(double* base, int N_, int tid) {
  WF_ASSUME_SHARED(base);
  constexpr int P = 32 * NW;
  const WfLayout L = wf_layout(N_);
  const int n = 3 * (N_ - 2 * TP_DEGREE);
  const double* g = base + L.g;
  const double* H = base + L.H;
  double* d = base + L.d;
  double* x = base + L.cp + 3 * TP_DEGREE;
  double* xp = base + L.xp;
  double* gp = base + L.gp;
  double ca[WF_M], cb[WF_M];
#pragma unroll
  for (int j = 0; j < WF_M; j += 2) {
    const double2 a = *reinterpret_cast<const double2*>(base + L.ca + j), b = *reinterpret_cast<const double2*>(base + L.cb + j);
    ca[j] = a.x; ca[j + 1] = a.y; cb[j] = b.x; cb[j + 1] = b.y;
  }
  const double cg = base[L.sc];
#pragma unroll 1
  for (int i = tid; i < n; i += P) {
    const double gi = g[i], xi = x[i];
    double a0 = cg * gi, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const double2* hp = reinterpret_cast<const double2*>(H + i * WF_HS);
#pragma unroll
    for (int j = 0; j < WF_M / 4; ++j) {
      const double2 s0 = hp[j], s1 = hp[j + 4], y0 = hp[j + 8], y1 = hp[j + 12];
      a0 = fma(ca[2 * j], s0.x, a0);
      a1 = fma(ca[2 * j + 8], s1.x, a1);
      a2 = fma(cb[2 * j], y0.x, a2);
      a3 = fma(cb[2 * j + 8], y1.x, a3);
      a0 = fma(ca[2 * j + 1], s0.y, a0);
      a1 = fma(ca[2 * j + 9], s1.y, a1);
      a2 = fma(cb[2 * j + 1], y0.y, a2);
      a3 = fma(cb[2 * j + 9], y1.y, a3);
    }
    const double dv = (a0 + a1) + (a2 + a3);
    d[i] = dv;
    xp[i] = xi;
    gp[i] = gi;
    x[i] = xi + dv;   // = xp + 1.0 * d bit for bit
  }
}

// state of one More-Thuente line search (lbfgs.hpp:716-937) that its rarely-taken tail needs
struct WfLs {
  double stx, fxx, dgx, sty, fy, dgy, stp, stmin, stmax, width, prev_width, dgtest;
  int brackt, stage1, uinfo;
};
// The tail of a line-search trial that was NOT accepted (lbfgs.hpp:866-934): stage switch, update_trial_interval on the
// function or the modified function, bisection safeguard.  Out of line: most iterations accept their first trial.
__device__ __noinline__ void wf_ls_tail(WfLs& T, double fx, double dg, double ftest1, double dginit) {
  const double ftol = 1e-4, gtol = 0.9;
  if (T.stage1 && fx <= ftest1 && (ftol <= gtol ? ftol : gtol) * dginit <= dg) T.stage1 = 0;
  if (T.stage1 && ftest1 < fx && fx <= T.fxx) {
    double fm = fx - T.stp * T.dgtest;
    double fxm = T.fxx - T.stx * T.dgtest;
    double fym = T.fy - T.sty * T.dgtest;
    double dgm = dg - T.dgtest;
    double dgxm = T.dgx - T.dgtest;
    double dgym = T.dgy - T.dgtest;
    T.uinfo = mt_update(T.stx, fxm, dgxm, T.sty, fym, dgym, T.stp, fm, dgm, T.stmin, T.stmax, T.brackt);
    T.fxx = fxm + T.stx * T.dgtest;
    T.fy = fym + T.sty * T.dgtest;
    T.dgx = dgxm + T.dgtest;
    T.dgy = dgym + T.dgtest;
  } else {
    double ft = fx, dt = dg;
    T.uinfo = mt_update(T.stx, T.fxx, T.dgx, T.sty, T.fy, T.dgy, T.stp, ft, dt, T.stmin, T.stmax, T.brackt);
  }
  if (T.brackt) {
    if (0.66 * T.prev_width <= fabs(T.sty - T.stx)) T.stp = T.stx + 0.5 * (T.sty - T.stx);
    T.prev_width = T.width;
    T.width = fabs(T.sty - T.stx);
  }
}

// One optimize() (bsplineTraj.cpp:687-718), warp form.  `base` = the worker's solver base after wf_setup; base + L.cp
// holds the control points on entry; on return it holds the LAST EVALUATED point (bsplineTraj.cpp:803) and xfinal
// (global, may be null) the solver's own x.
template <int NW>
__device__ __noinline__ void lbfgs_run_team(const VigoConst& C, double* base, int N, int serial_warp, tp_lbfgs_result& out,
                                            double* xfinal, int tid) {
  WF_ASSUME_SHARED(base);
  constexpr int P = 32 * NW;
  const int lane = tid & 31;
  const WfLayout L = wf_layout(N);
  const int n = 3 * (N - 2 * TP_DEGREE);
  double* cp = base + L.cp;
  double* x = cp + 3 * TP_DEGREE;
  double* g = base + L.g;
  double* xp = base + L.xp;
  double* gp = base + L.gp;
  double* d = base + L.d;
  const double* sc = base + L.sc;
  const double min_step = 1e-20, max_step = 1e20, ftol = 1e-4, gtol = 0.9, xtol = 1e-16;  // lbfgs.hpp:942-954
  const int max_ls = C.p.lbfgs_max_linesearch, max_iter = C.p.lbfgs_max_iter;
  const double g_eps = C.p.lbfgs_g_eps, geps2 = g_eps * g_eps;
#define WF_OWNED(i) _Pragma("unroll 1") for (int i = tid; i < n; i += P)
  wf_sync<NW>();
  int evals = 0, k = 0, ret, bsum = 0;
  double fx, gg, xx;
#ifdef TP_WF_TIMING
  long long tE = 0, tG = 0, tC = 0, tD = 0, tT0 = clock64(), tq0, tq1;
#define WT0 tq0 = clock64();
#define WT(acc) { tq1 = clock64(); acc += tq1 - tq0; tq0 = tq1; }
#else
#define WT0
#define WT(acc)
#endif
  {
    const WfEv ev = wf_eval<NW>(C, base, 0, tid);
    fx = ev.v[0]; gg = ev.v[2]; xx = ev.v[3];
  }
  ++evals;
  WF_OWNED(i) d[i] = -g[i];
  double xnorm = sqrt(xx), gnorm = sqrt(gg);
  if (xnorm < 1.0) xnorm = 1.0;
  if (gnorm / xnorm <= g_eps) {
    ret = 2;  // LBFGS_ALREADY_MINIMIZED
  } else {
    double step = 1.0 / sqrt(gg);   // 1/||d||, d = -g
    double dginit = -gg;            // g.d for d = -g
    int end = 0;
    k = 1;
    for (;;) {
      // from the second iteration on wf_direction has already saved xp / gp and advanced x by the unit first trial step
      const bool pre = k > 1;
      if (!pre) WF_OWNED(i) { xp[i] = x[i]; gp[i] = g[i]; }
      // ---------------- line_search_morethuente (lbfgs.hpp:716-937), same scalar logic as tp_lbfgs.cuh
      int ls;
      if (step <= 0. || 0 < dginit) {
        ls = step <= 0. ? LB_INVALIDPARAMS : LB_INCREASEGRADIENT;
        if (pre) WF_OWNED(i) x[i] = xp[i];   // no trial is evaluated: the control points stay the last evaluated point
      } else {
        int count = 0, brackt = 0, stage1 = 1, uinfo = 0;
        const double finit = fx;
        double dg, stp = step;
        const double dgtest = ftol * dginit;
        double width = max_step - min_step;
        double prev_width = 2.0 * width;
        double stx = 0., sty = 0.;
        double fxx = finit, fy = finit;
        double dgx = dginit, dgy = dginit;
        double stmin, stmax;
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
          if (NW == 1) __syncwarp();   // every lane is done reading the control points of the previous evaluation
                                       // (NW > 1: the barrier inside the evaluation's reduction already says so)
          if (!(pre && count == 0 && stp == 1.0)) WF_OWNED(i) x[i] = xp[i] + stp * d[i];
          WT0
          wf_sync<NW>();
          {
            const WfEv ev = wf_eval<NW>(C, base, 1, tid);
            fx = ev.v[0]; dg = ev.v[1]; gg = ev.v[2]; xx = ev.v[3];
          }
          WT(tE)
          ++evals;
          const double ftest1 = finit + stp * dgtest;
          ++count;
          if (brackt && ((stp <= stmin || stmax <= stp) || uinfo != 0)) { ls = LB_ROUNDING; break; }
          if (stp == max_step && fx <= ftest1 && dg <= dgtest) { ls = LB_MAXSTEP; break; }
          if (stp == min_step && (ftest1 < fx || dgtest <= dg)) { ls = LB_MINSTEP; break; }
          if (brackt && (stmax - stmin) <= xtol * stmax) { ls = LB_WIDTHTOOSMALL; break; }
          if (max_ls <= count) { ls = LB_MAXLINESEARCH; break; }
          if (fx <= ftest1 && fabs(dg) <= gtol * (-dginit)) { ls = count; break; }
          // not accepted: the rarely-taken tail, out of line (its state travels through a struct only here)
          WfLs T;
          T.stx = stx; T.fxx = fxx; T.dgx = dgx; T.sty = sty; T.fy = fy; T.dgy = dgy; T.stp = stp; T.stmin = stmin;
          T.stmax = stmax; T.width = width; T.prev_width = prev_width; T.dgtest = dgtest;
          T.brackt = brackt; T.stage1 = stage1; T.uinfo = uinfo;
          wf_ls_tail(T, fx, dg, ftest1, dginit);
          stx = T.stx; fxx = T.fxx; dgx = T.dgx; sty = T.sty; fy = T.fy; dgy = T.dgy; stp = T.stp;
          width = T.width; prev_width = T.prev_width; brackt = T.brackt; stage1 = T.stage1; uinfo = T.uinfo;
        }
        step = stp;
      }
      if (ls < 0) {
        if (xfinal) WF_OWNED(i) xfinal[i] = xp[i];
        xfinal = nullptr;
        ret = ls;
        break;
      }
      // ||g|| / max(1, ||x||) <= g_epsilon (lbfgs.hpp:1218-1225) without the two square roots and the division
      if (gg <= geps2 * (xx < 1.0 ? 1.0 : xx)) { ret = 0; break; }
      if (max_iter != 0 && max_iter < k + 1) { ret = LB_MAXITER; break; }
      // ---------------- new pair into slot `end`, Gram update, coefficients, direction
      bsum += (WF_M <= k) ? WF_M : k;
      WT0
      wf_gram_update<NW>(base, N, end, serial_warp, tid);
      WT(tG)
      if (NW == 1 || (tid >> 5) == serial_warp) wf_coeffs(base + L.gram, base + L.ca, base + L.cb, base + L.sc, end, gg, lane);
      wf_sync<NW>();
      WT(tC)
      dginit = sc[1];
      wf_direction<NW>(base, N, tid);
      WT(tD)
      ++k;
      end = (end + 1) & (WF_M - 1);
      step = 1.0;
    }
  }
  if (xfinal) WF_OWNED(i) xfinal[i] = x[i];
#ifdef TP_WF_TIMING
  if (tid == 0) {
    atomicAdd(&g_wf_phase[0], (unsigned long long)(clock64() - tT0));
    atomicAdd(&g_wf_phase[1], (unsigned long long)tE);
    atomicAdd(&g_wf_phase[2], (unsigned long long)tG);
    atomicAdd(&g_wf_phase[3], (unsigned long long)tC);
    atomicAdd(&g_wf_phase[4], (unsigned long long)tD);
    atomicAdd(&g_wf_phase[5], (unsigned long long)k);
    atomicAdd(&g_wf_phase[6], (unsigned long long)evals);
  }
#endif
  out.ret = ret;
  out.iters = k;
  out.evals = evals;
  out.reserved = bsum;
  out.fx = fx;
  wf_sync<NW>();
#undef WF_OWNED
}
