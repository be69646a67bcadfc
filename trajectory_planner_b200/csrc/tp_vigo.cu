// ViGO batch engine: kernels, per-GPU engine state and the C ABI entry points of include/tp_b200.h.
// sm_100a only; compiled with --fmad=false (see tp_device.cuh).  No CPU fallback: every compute
// entry point needs a CUDA device.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "tp_lbfgs.cuh"
#include "tp_lbfgs_fast.cuh"
#include "tp_lbfgs_warp.cuh"
#include "tp_frontend.cuh"
#include "tp_map.h"
#include "tp_outer.cuh"

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t _e = (call);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      tp_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));     \
      return TP_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

// =========================================================================== kernels
// ---- map queries (occMap::isInflatedOccupied / isUnknown / isInflatedOccupiedLine)
// One thread per query; xyz is read as three coalesced FP64 streams per warp (24 B/query) and the
// map word through the read-only path (one 32 B sector per random gather).
// Four consecutive queries per thread and iteration: 96 B of coordinates as six 16-byte streaming loads (all issued
// before the first use), four independent map gathers, four flags as one 4-byte store (the arrays are 16- / 4-byte
// aligned at query indices that are multiples of four).
__device__ __forceinline__ unsigned dq_flag(const DevMap& map, const D3& p, int unknown) {
  return unknown ? (dm_unknown(map, p) ? 1u : 0u) : (dm_inflated(map, p) ? 1u : 0u);
}
template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_query_points(DevMap map, long n, const double* __restrict__ xyz, uint8_t* __restrict__ out, int unknown) {
  const long stride = (long)gridDim.x * blockDim.x;
  const long nquad = n >> 2;
  const bool aligned = ((reinterpret_cast<uintptr_t>(xyz) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 3) == 0);
  if (aligned) {
    const double2* __restrict__ v = reinterpret_cast<const double2*>(xyz);
    for (long j = blockIdx.x * (long)blockDim.x + threadIdx.x; j < nquad; j += stride) {
      const double2 a = __ldcs(v + 6 * j), b = __ldcs(v + 6 * j + 1), c = __ldcs(v + 6 * j + 2), d = __ldcs(v + 6 * j + 3),
                    e = __ldcs(v + 6 * j + 4), f = __ldcs(v + 6 * j + 5);
      // the four cell addresses first, then the four map gathers back to back (one L2 round trip per iteration instead of
      // four dependent ones), then the flags; a point outside the grid reads word 0 and is forced to "occupied / unknown"
      const double px[4] = {a.x, b.y, d.x, e.y}, py[4] = {a.y, c.x, d.y, f.x}, pz[4] = {b.x, c.y, e.x, f.y};
      size_t w[4];
      int bit[4];
      bool in[4];
      const uint32_t* __restrict__ grid = unknown ? map.known : map.inflated;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int ix, iy, iz;
        in[q] = dm_index(map, px[q], py[q], pz[q], ix, iy, iz);
        w[q] = in[q] ? ((size_t)ix * map.dim[1] + iy) * map.wz + (iz >> 5) : 0;
        bit[q] = in[q] ? (iz & 31) : 0;
      }
      uint32_t word[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) word[q] = __ldg(grid + w[q]);
      unsigned h = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const unsigned set = (word[q] >> bit[q]) & 1u;
        const unsigned flag = in[q] ? (unknown ? (set ^ 1u) : set) : 1u;
        h |= flag << (8 * q);
      }
      __stcs(reinterpret_cast<unsigned*>(out) + j, h);
    }
  }
  // the last n % 4 queries, or everything when the buffers are not aligned
  for (long i = (aligned ? 4 * nquad : 0) + blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += stride) {
    const D3 p = d3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    out[i] = (uint8_t)dq_flag(map, p, unknown);
  }
}
__global__ void k_query_lines(DevMap map, long n, const double* __restrict__ a, const double* __restrict__ b,
                              uint8_t* __restrict__ out) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    out[i] = dm_line(map, d3(a[3 * i], a[3 * i + 1], a[3 * i + 2]), d3(b[3 * i], b[3 * i + 1], b[3 * i + 2])) ? 1 : 0;
}

// ---- batch state set-up
__global__ void k_init_states(BatchView bv, VigoConst C) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bv.B) return;
  TrajState& st = bv.st[b];
  st.N = bv.off[b + 1] - bv.off[b];
  st.off = bv.off[b];
  st.status = TS_ACTIVE;
  st.has_col = 0;
  st.fail_count = 0;
  st.round = 0;
  st.nseg = 0;
  st.n_pairs = 0;
  st.err = 0;
  st.lbfgs_runs = st.lbfgs_iters = st.lbfgs_evals = st.last_ret = 0;
  st.astar_searches = st.astar_expansions = 0;
  st.astar_unreach = 0;
  st.vclock = 0;
  st.w_dist = C.p.w_distance;
  st.w_dyn = C.p.w_dyn;
  st.final_cost = 0;
  st.linear_factor = 1.0;
}
__global__ void k_fill_int(int* p, long n, int v) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_iota(int* p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

// ---- costFunction (test/parity entry): one block per trajectory
template <bool STRICT>
__global__ void __launch_bounds__(TP_LB_THREADS) k_cost(BatchView bv, VigoConst C, double* f_out, double* grad_out) {
  extern __shared__ double sm[];
  const int b = blockIdx.x, tid = threadIdx.x;
  const TrajState& st = bv.st[b];
  const int N = st.N, n = 3 * (N - 2 * TP_DEGREE);
  if (n <= 0) {
    if (tid == 0) f_out[b] = 0.0;
    return;
  }
  if (!STRICT) {
    // the throughput form of the evaluation (tp_lbfgs_fast.cuh) — what the default solve kernel runs
    VfCtx V;
    V.N = N; V.n = n; V.sm = sm; V.L = vf_layout(N);
    V.pairs = bv.pairs + (size_t)b * C.gcap; V.head = bv.cp_head + st.off; V.n_pairs = st.n_pairs; V.pairs_in_sm = false;
    V.serial_warp = 0;
    V.w_dist = st.w_dist; V.w_dyn = st.w_dyn;
    V.n_dyn = bv.n_dyn; V.dyn_pos = bv.dyn_pos; V.dyn_vel = bv.dyn_vel; V.dyn_size = bv.dyn_size;
    for (int e = tid; e < 3 * N; e += TP_LB_THREADS) sm[V.L.cp + e] = bv.ctrl[3 * (size_t)st.off + e];
    __syncthreads();
    vf_stage_pairs(V, tid);
    __syncthreads();
    Red R2;
    R2.buf = sm + V.L.red;
    R2.flip = 0;
    double f, dg, gg, xx;
    vf_eval(C, V, R2, false, f, dg, gg, xx, tid);
    if (tid == 0) f_out[b] = f;
    double* go = grad_out + 3 * ((size_t)st.off - (size_t)2 * TP_DEGREE * b);
    for (int e = tid; e < n; e += TP_LB_THREADS) go[e] = sm[V.L.g + e];
    return;
  }
  double* cp = sm;
  double* g = cp + 3 * N;
  Red R;
  R.buf = g + n;
  R.flip = 0;
  for (int e = tid; e < 3 * N; e += TP_LB_THREADS) cp[e] = bv.ctrl[3 * (size_t)st.off + e];
  EvalCtx E;
  E.N = N; E.n = n; E.cp = cp;
  E.pairs = bv.pairs + (size_t)b * C.gcap;
  E.head = bv.cp_head + st.off;
  E.w_dist = st.w_dist; E.w_dyn = st.w_dyn;
  E.n_dyn = bv.n_dyn; E.dyn_pos = bv.dyn_pos; E.dyn_vel = bv.dyn_vel; E.dyn_size = bv.dyn_size;
  __syncthreads();
  double f, dg;
  eval_cost<STRICT>(C, E, R, g, nullptr, f, dg, tid);
  if (tid == 0) f_out[b] = f;
  double* go = grad_out + 3 * ((size_t)st.off - (size_t)2 * TP_DEGREE * b);
  for (int e = tid; e < n; e += TP_LB_THREADS) go[e] = g[e];
}

// ---- costFunction in the team form (4 warps), the default evaluation
__global__ void __launch_bounds__(TP_LB_THREADS) k_cost_t(BatchView bv, VigoConst C, double* f_out, double* grad_out) {
  extern __shared__ double sm[];
  const int b = blockIdx.x, tid = threadIdx.x;
  const TrajState& st = bv.st[b];
  const int N = st.N, n = 3 * (N - 2 * TP_DEGREE);
  if (n <= 0) {
    if (tid == 0) f_out[b] = 0.0;
    return;
  }
  const WfLayout L = wf_layout(N);
  for (int e = tid; e < 3 * N; e += TP_LB_THREADS) sm[L.cp + e] = bv.ctrl[3 * (size_t)st.off + e];
  __syncthreads();
  wf_setup<TP_LB_WARPS>(C, sm, N, bv.pairs + (size_t)b * C.gcap, bv.cp_head + st.off, st.n_pairs, st.w_dist, st.w_dyn, bv.n_dyn,
                        bv.dyn_pos, bv.dyn_vel, bv.dyn_size, tid);
  const WfEv evr = wf_eval<TP_LB_WARPS>(C, sm, 0, tid);
  const double* ev = evr.v;
  __syncthreads();
  if (tid == 0) f_out[b] = ev[0];
  double* go = grad_out + 3 * ((size_t)st.off - (size_t)2 * TP_DEGREE * b);
  for (int e = tid; e < n; e += TP_LB_THREADS) go[e] = sm[L.g + e];
}

// ---- optimize(): fused cost + L-BFGS, one block per ACTIVE trajectory.
// MODE 0: classic two-loop, tree reductions; 1: classic, serial-order (bit-faithful); 2: vector-free (tp_lbfgs_fast.cuh)
template <int MODE>
__global__ void __launch_bounds__(TP_LB_THREADS, 4) k_lbfgs(BatchView bv, VigoConst C, const int* __restrict__ active,
                                                         const int* __restrict__ n_active, tp_lbfgs_result* res_out,
                                                         double* xfinal_out, double* counters) {
  extern __shared__ double sm[];
  if ((int)blockIdx.x >= *n_active) return;
  const int b = active[blockIdx.x], tid = threadIdx.x;
  TrajState& st = bv.st[b];
  const int N = st.N, n = 3 * (N - 2 * TP_DEGREE);
  // tensor-memory address slot: one word past this trajectory's layout (the launch reserves 16 B beyond the largest)
  uint32_t* tm_slot = reinterpret_cast<uint32_t*>(sm + vf_layout(N < 2 * TP_DEGREE ? 2 * TP_DEGREE : N, true).total);
  uint32_t tbase = 0;
  if (MODE == 3) tbase = tm_block_alloc(tm_slot, tid);
  tp_lbfgs_result r;
  if (n <= 0) {
    r.ret = LB_INVALID_N; r.iters = 0; r.evals = 0; r.reserved = 0; r.fx = 0;
  } else {
    double* cp = MODE == 4 ? sm + WF_CTX : sm;
    double* gctrl = bv.ctrl + 3 * (size_t)st.off;
    for (int e = tid; e < 3 * N; e += TP_LB_THREADS) cp[e] = gctrl[e];
    EvalCtx E;
    E.N = N; E.n = n; E.cp = cp;
    E.pairs = bv.pairs + (size_t)b * C.gcap;
    E.head = bv.cp_head + st.off;
    E.w_dist = st.w_dist; E.w_dyn = st.w_dyn;
    E.n_dyn = bv.n_dyn; E.dyn_pos = bv.dyn_pos; E.dyn_vel = bv.dyn_vel; E.dyn_size = bv.dyn_size;
    double* xf = xfinal_out ? xfinal_out + 3 * ((size_t)st.off - (size_t)2 * TP_DEGREE * b) : nullptr;
    if (MODE == 4) {
      __syncthreads();
      wf_setup<TP_LB_WARPS>(C, sm, N, E.pairs, E.head, st.n_pairs, E.w_dist, E.w_dyn, E.n_dyn, E.dyn_pos, E.dyn_vel, E.dyn_size, tid);
      lbfgs_run_team<TP_LB_WARPS>(C, sm, N, 0, r, xf, tid);
    } else if (MODE >= 2) {
      VfCtx V;
      V.N = N; V.n = n; V.sm = sm; V.L = vf_layout(N, MODE == 3);
      V.pairs = E.pairs; V.head = E.head; V.n_pairs = st.n_pairs; V.pairs_in_sm = false; V.serial_warp = 0;
      V.w_dist = E.w_dist; V.w_dyn = E.w_dyn;
      V.n_dyn = E.n_dyn; V.dyn_pos = E.dyn_pos; V.dyn_vel = E.dyn_vel; V.dyn_size = E.dyn_size;
      V.tbase = tbase;
      lbfgs_run_fast<MODE == 3>(C, V, r, xf, tid);
    } else {
      lbfgs_run<MODE == 1>(C, E, cp + 3 * N, r, xf, tid);
    }
    // the control points keep the last evaluated point (bsplineTraj.cpp:803)
    for (int e = tid; e < n; e += TP_LB_THREADS) gctrl[3 * TP_DEGREE + e] = cp[3 * TP_DEGREE + e];
  }
  if (tid == 0) {
    st.lbfgs_runs += 1;
    st.lbfgs_iters += r.iters;
    st.lbfgs_evals += r.evals;
    st.last_ret = r.ret;
    st.final_cost = r.fx;
    if (n > 0) st.vclock += (long long)r.evals * (10LL * N + 2LL * n);
    if (res_out) res_out[b] = r;
    if (counters && n > 0) {
      // algorithmic FP64 flops: E(81N + 21G + 4n) + sum_k (8 b_k + 15) n
      const double fl = (double)r.evals * (81.0 * N + 21.0 * st.n_pairs + 4.0 * n) + (8.0 * r.reserved + 15.0 * r.iters) * n;
      atomicAdd(&counters[0], fl);
      atomicAdd(&counters[1], (double)r.iters);
      atomicAdd(&counters[2], (double)r.evals);
    }
  }
  if (MODE == 3) tm_block_free(tm_slot, tid);
}

// ---- hasCollisionTrajectory (+ hasDynamicCollisionTrajectory), block-wide on control points in shared memory
struct SmemCP {
  const double* p;
  __device__ __forceinline__ D3 operator()(int i) const { return d3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
};
// returns static | dynamic << 1 to every thread (contains a block barrier)
__device__ __forceinline__ int dev_has_collision(const double* cps, int N, const BatchView& bv, const VigoConst& C,
                                                 const DevMap& map, int tid) {
  SmemCP cp{cps};
  const double cts = C.p.ctrl_pt_ts;
  const double duration = (double)(N - TP_DEGREE) * cts;          // knots_(N), bspline.cpp:27
  const double limit = (1.0 - C.p.not_check_ratio) * duration;    // bsplineTraj.h:313
  int hit = 0, dyn = 0;
  for (int s = tid; s < C.n_t_check; s += TP_LB_THREADS) {
    const double t = bv.t_check[s];
    if (!(t <= duration)) break;  // the table is increasing
    const D3 p = bspline_at(cp, N, TP_DEGREE, cts, t);
    if (t <= limit && dm_inflated(map, p)) hit = 1;
    for (int j = 0; j < bv.n_dyn; ++j) {  // bsplineTraj.h:344-368 (samples of evalTraj(): t <= duration)
      const double size = fmin(bv.dyn_size[3 * j] / 2, bv.dyn_size[3 * j + 1] / 2);
      const double dx = p.x - bv.dyn_pos[3 * j], dy = p.y - bv.dyn_pos[3 * j + 1];
      const double dist = sqrt((dx * dx + dy * dy) + 0.0 * 0.0) - size;
      if (dist < 0) dyn = 1;
    }
  }
  return __syncthreads_or(hit | (dyn << 1));
}
__global__ void __launch_bounds__(TP_LB_THREADS) k_has_collision(BatchView bv, VigoConst C, DevMap map,
                                                                 const int* __restrict__ active,
                                                                 const int* __restrict__ n_active, uint8_t* hit_out,
                                                                 double* counters) {
  extern __shared__ double sm[];
  if ((int)blockIdx.x >= *n_active) return;
  const int b = active[blockIdx.x], tid = threadIdx.x;
  TrajState& st = bv.st[b];
  const int N = st.N;
  if (N < 4) {
    if (tid == 0) { st.has_col = 1; if (hit_out) hit_out[b] = 1; }
    return;
  }
  for (int e = tid; e < 3 * N; e += TP_LB_THREADS) sm[e] = bv.ctrl[3 * (size_t)st.off + e];
  __syncthreads();
  const int any = dev_has_collision(sm, N, bv, C, map, tid);
  if (tid == 0) {
    st.has_col = any;
    if (hit_out) hit_out[b] = (uint8_t)(any & 1);
    if (counters) atomicAdd(&counters[3], floor((double)(N - TP_DEGREE) * C.p.ctrl_pt_ts / C.check_ts) + 1.0);
  }
}

// ---- linearFeasibilityReparam (bsplineTraj.cpp:1116-1137), block-wide.  cp: control points (shared),
// q / r: scratch for the velocity / acceleration spline control points (3(N-1), 3(N-2) doubles, shared),
// red: 2 * TP_LB_WARPS doubles (shared).  Returns the factor to thread 0.
__device__ __forceinline__ double dev_reparam(const double* cp, double* q, double* r, double* red, int N, const BatchView& bv,
                                              const VigoConst& C, int tid) {
  const double ts = C.p.ctrl_pt_ts;
  for (int e = tid; e < 3 * (N - 1); e += TP_LB_THREADS) {
    const int i = e / 3;
    const double den = (double)(i + 3 + 1 - 3) * ts - (double)(i + 1 - 3) * ts;  // knots_(i+p+1) - knots_(i+1), p = 3
    q[e] = (3.0 * (cp[e + 3] - cp[e])) / den;
  }
  __syncthreads();
  for (int e = tid; e < 3 * (N - 2); e += TP_LB_THREADS) {
    const int i = e / 3;
    const double den = (double)(i + 2 + 1 - 2) * ts - (double)(i + 1 - 2) * ts;  // derivative spline: p = 2
    r[e] = (2.0 * (q[e + 3] - q[e])) / den;
  }
  __syncthreads();
  SmemCP vq{q}, ar{r};
  const double duration = (double)(N - TP_DEGREE) * ts;
  double mv = 0.0, ma = 0.0;
  for (int s = tid; s < C.n_t_reparam; s += TP_LB_THREADS) {
    const double t = bv.t_reparam[s];
    if (!(t < duration)) break;
    const double v = norm3(bspline_at(vq, N - 1, 2, ts, t));
    const double a = norm3(bspline_at(ar, N - 2, 1, ts, t));
    mv = fmax(mv, v);
    ma = fmax(ma, a);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mv = fmax(mv, __shfl_xor_sync(0xffffffffu, mv, o));
    ma = fmax(ma, __shfl_xor_sync(0xffffffffu, ma, o));
  }
  if ((tid & 31) == 0) { red[tid >> 5] = mv; red[TP_LB_WARPS + (tid >> 5)] = ma; }
  __syncthreads();
  double f = 1.0;
  if (tid == 0) {
    for (int w = 1; w < TP_LB_WARPS; ++w) { mv = fmax(mv, red[w]); ma = fmax(ma, red[TP_LB_WARPS + w]); }
    const double fv = C.p.max_vel / mv;
    const double fa = sqrt(C.p.max_acc / ma);
    f = fmin(fv, fa);
  }
  return f;
}
__global__ void __launch_bounds__(TP_LB_THREADS) k_reparam(BatchView bv, VigoConst C) {
  extern __shared__ double sm[];
  __shared__ double red[2 * TP_LB_WARPS];
  const int b = blockIdx.x, tid = threadIdx.x;
  TrajState& st = bv.st[b];
  if (st.status != TP_STATUS_SUCCESS) return;
  const int N = st.N;
  double* cp = sm;
  for (int e = tid; e < 3 * N; e += TP_LB_THREADS) cp[e] = bv.ctrl[3 * (size_t)st.off + e];
  __syncthreads();
  const double f = dev_reparam(cp, cp + 3 * N, cp + 3 * N + 3 * (N - 1), red, N, bv, C, tid);
  if (tid == 0) st.linear_factor = f;
}

// ---- outer loop pieces, one warp per trajectory
struct PlanSmem {
  AStarSmem as;
  uint8_t hit[TP_MAX_CTRL];
  uint8_t line[TP_MAX_CTRL];
  int segA[TP_MAX_SEG_HARD][2];
  int segB[TP_MAX_SEG_HARD][2];
  int prevSeg[TP_MAX_SEG_HARD][2];
};
__device__ __forceinline__ Worker make_worker(const VigoConst& C, const AStarPools& P, int w, int lane, AStarSmem* as) {
  Worker W;
  W.nodes = P.nodes + (size_t)w * (P.pool_nodes + 1);
  W.sm = as;
  W.heap_k_gl = P.heap_k + (size_t)w * C.heap_cap;
  W.heap_n_gl = P.heap_n + (size_t)w * C.heap_cap;
  W.path = P.paths + (size_t)w * C.path_cap * 3;
  W.sc = P.sc + (size_t)w * C.max_seg * TP_SC_CAP * 3;
  W.sc_len = P.sc_len + (size_t)w * C.max_seg;
  W.round_ptr = P.rounds + w;
  W.lane = lane;
  W.flood_trigger = TP_FLOOD_TRIGGER;
  W.goal_unreachable = 0;
  W.flood_wasted = 0;
  W.flood_stats = P.flood_stats;
  return W;
}

// makePlan steps 1-3 (bsplineTraj.cpp:341-352) for trajectory b; `bv.ctrl + 3*st.off` are its control points
// (global or shared).  Leaves st.status == TS_ACTIVE when the solve goes on.
__device__ __noinline__ void dev_plan_init(const BatchView& bv, const VigoConst& C, const DevMap& map, TrajState& st, Worker& W,
                              PlanSmem& S, int b, int lane) {
  int err = 0;
  if (st.N < 2 * TP_DEGREE + 1 || st.N > TP_MAX_CTRL) {
    if (lane == 0) st.status = TP_STATUS_INVALID;
    __syncwarp();
    return;
  }
  int nseg = find_collision_seg(map, C, bv, st, S.hit, S.line, S.segA, lane, err);
  const int npaths = path_search(map, C, bv, st, W, S.segA, nseg, err);
  if (lane == 0) {
    st.nseg = nseg;
    for (int i = 0; i < nseg; ++i) { st.seg[i][0] = S.segA[i][0]; st.seg[i][1] = S.segA[i][1]; }
    if (npaths < 0) st.status = TP_STATUS_FAIL_ASTAR;
    else assign_guides(map, C, bv, b, st, W, S.segA, nseg, npaths);
    st.err |= err;
  }
  __syncwarp();
}

// body of optimizeTrajectory's loop after the collision check (bsplineTraj.cpp:628-679).  Sets st.status to a
// final value or leaves TS_ACTIVE (another optimize() follows).
__device__ __noinline__ void dev_plan_step(const BatchView& bv, const VigoConst& C, const DevMap& map, TrajState& st, Worker& W,
                              PlanSmem& S, int b, int lane) {
  int err = 0;
  const int hasCol = st.has_col & 1, hasDyn = (st.has_col >> 1) & 1;
  if (!hasCol && !hasDyn) {  // :628-630 -> :682-684
    if (lane == 0) {
      st.w_dist = C.p.w_distance;
      st.w_dyn = C.p.w_dyn;
      st.status = TP_STATUS_SUCCESS;
    }
    __syncwarp();
    return;
  }
  if (st.round == 0 && lane == 0) st.vclock = 0;  // the reference starts its timer after the first optimize() (:618)
  __syncwarp();
  // deterministic replacements of the 0.03 s wall clock (:632-638): virtual clock + round cap
  if (st.round >= C.p.max_outer_rounds || (C.p.vclock_budget > 0 && st.vclock > (long long)C.p.vclock_budget)) {
    if (lane == 0) {
      st.w_dist = C.p.w_distance;
      st.w_dyn = C.p.w_dyn;
      st.status = TP_STATUS_FAIL_OPTIMIZE;
    }
    __syncwarp();
    return;
  }
  const int fail_in = st.fail_count;
  if (fail_in >= 4) {  // :640-648
    int nseg = find_collision_seg(map, C, bv, st, S.hit, S.line, S.segA, lane, err);
    const int np = path_search(map, C, bv, st, W, S.segA, nseg, err);
    if (np >= 0 && lane == 0) assign_guides(map, C, bv, b, st, W, S.segA, nseg, np);
    __syncwarp();
  }
  if (fail_in >= 8) {  // :650-654
    if (lane == 0) {
      st.round += 1;
      st.w_dist = C.p.w_distance;
      st.w_dyn = C.p.w_dyn;
      st.status = TP_STATUS_FAIL_OPTIMIZE;
      st.err |= err;
    }
    __syncwarp();
    return;
  }
  int add_fail = 0;
  if (hasCol) {
    // ---- isReguideRequired (:573-608)
    const int nprev = st.nseg;
    for (int i = lane; i < nprev; i += 32) { S.prevSeg[i][0] = st.seg[i][0]; S.prevSeg[i][1] = st.seg[i][1]; }
    __syncwarp();
    const int nnew = find_collision_seg(map, C, bv, st, S.hit, S.line, S.segA, lane, err);
    int nre = 0;
    if (lane == 0) {
      st.nseg = nnew;
      for (int i = 0; i < nnew; ++i) { st.seg[i][0] = S.segA[i][0]; st.seg[i][1] = S.segA[i][1]; }
      // compareCollisionSeg (bsplineTraj.h:379-403) + set<int> of segment indices
      unsigned long long need = 0ull;  // bit i: segment i needs a re-guide (max_seg <= 64)
      for (int sidx = 0; sidx < nnew; ++sidx) {
        const int s0 = S.segA[sidx][0], s1 = S.segA[sidx][1];
        auto visit = [&](int i) {
          const bool overl = index_in_seg(S.prevSeg, nprev, i);
          if (!overl || cp_requires_new_guide(C, bv, b, st, i)) {
            const int k = find_seg_index(S.segA, nnew, i);
            if (k >= 0) need |= (1ull << k);
          }
        };
        for (int i = s0 + 1; i <= s1 - 1; ++i) visit(i);
        if (s1 - s0 - 1 == 0)
          for (int i = s0; i <= s1; ++i) visit(i);
      }
      for (int k = 0; k < nnew; ++k)
        if ((need >> k) & 1ull) {
          S.segB[nre][0] = S.segA[k][0];
          S.segB[nre][1] = S.segA[k][1];
          ++nre;
        }
    }
    nre = __shfl_sync(0xffffffffu, nre, 0);
    __syncwarp();
    if (nre > 0) {
      int nre2 = nre;
      const int np = path_search(map, C, bv, st, W, S.segB, nre2, err);
      if (np >= 0) {
        if (lane == 0) assign_guides(map, C, bv, b, st, W, S.segB, nre2, np);
      } else
        add_fail = 1;
    } else
      add_fail = 1;
  }
  if (lane == 0) {
    st.round += 1;
    if (add_fail) {
      st.w_dist *= 2.0;
      st.fail_count += 1;
    }
    if (hasDyn) st.w_dyn *= 2.0;
    st.err |= err;
  }
  __syncwarp();
}

// standalone steps 1-3 (parity entry tp_vigo_init_guides_batch): one warp per worker, trajectories from a queue
__global__ void __launch_bounds__(32) k_plan_init(BatchView bv, VigoConst C, DevMap map, AStarPools P, int* queue,
                                                  int* active_out, int* n_active_out) {
  __shared__ PlanSmem S;
  const int lane = threadIdx.x;
  Worker W = make_worker(C, P, blockIdx.x, lane, &S.as);
  for (;;) {
    int b = 0;
    if (lane == 0) b = atomicAdd(queue, 1);
    b = __shfl_sync(0xffffffffu, b, 0);
    if (b >= bv.B) break;
    TrajState& st = bv.st[b];
    dev_plan_init(bv, C, map, st, W, S, b, lane);
    if (lane == 0 && st.status == TS_ACTIVE) active_out[atomicAdd(n_active_out, 1)] = b;
    __syncwarp();
  }
}

// ---- THE batched entry point's kernel: bsplineTraj::makePlan (bsplineTraj.cpp:333-385) for one trajectory
// per thread block, start to finish: segments -> A* detours -> guide points -> [fused cost+L-BFGS ->
// collision check -> re-guide / weight doubling]* -> time re-parameterisation.  Trajectories progress
// independently (no lock-step rounds, no host round trips): a long A* search or a 15-round solve delays
// only its own block while the hardware block scheduler back-fills the SM.  Warp 0 runs the serial
// outer-loop logic (lane-parallel map queries); all four warps run the solver, the collision check and
// the re-parameterisation.  Shared memory: control points + trajectory state persist; the solver state
// and the A* heap / tables alias each other (they are never live at the same time).
// Blocks are issued longest trajectory first (`order`); the launch is per size class so that the
// shared-memory footprint (and with it the blocks resident per SM) follows the trajectories' length.
// Parked trajectories of one batch (see k_solve): per size class a list of ids (filled with -1 before the launch),
// its tail (slots handed out), its head (slots claimed) and the number of trajectories that finished phase A.
#ifndef TP_TEAM_BLOCKS
// resident teams per SM the team-form kernel is compiled for.  2: 252 registers, no spills; 3: 168 registers (36 local-memory
// loads per warp and iteration: measured 1.7 % slower on the 4,096 batch, 7 % slower on a single solve); 4: 128 registers
#define TP_TEAM_BLOCKS 2
#endif
#define TP_PARK_BUCKETS 16
struct ParkQueue {
  int* list;       // [4 classes][TP_PARK_BUCKETS][stride]
  int* tail;       // [4][TP_PARK_BUCKETS]  slots handed out to parkers
  int* head;       // [4][TP_PARK_BUCKETS]  slots claimed by resumers
  int* started;    // [4]  trajectories of the class that are through phase A
  int stride;
  int thresh;      // A* expansions of makePlan steps 1-3 from which a trajectory counts as a hard start; < 0: parking disabled
  int score_mode;  // difficulty score of a parked trajectory: 0 guide pairs, 1 pairs x N / 32, 2 pairs x 8 + expansions / 64
  int b[3];        // bucket bounds on the score (descending)
  // time slicing (k_phase_a + k_solve<4>, resume = 2): a trajectory's FIRST turn is `slice` optimise / check / re-guide
  // rounds; if still active it goes to the back of its class's ring and its next turn runs to the end
  int* ring;       // [4][stride]  ids waiting for their next turn (-1 = empty slot)
  int* rtail;      // [4]  ring slots handed out
  int* rhead;      // [4]  ring slots claimed
  int* remaining;  // [4]  trajectories of the class that are not finished yet
  int slice;       // rounds of the first turn (<= 0: run to the end)
  int rev;         // one worker in `rev` scans the buckets easiest-first (0 = none)
};
// Bucket of a trajectory parked after makePlan steps 1-3, hardest-looking first.  0-5: by the length of its first A*
// searches (a trajectory whose first searches were long has long re-guide searches too — on the octomap rasters the
// longest trajectories are 100 k-expansion ones, 250 ms alone, and must start first), an unreachable goal counts as a
// long search; 6-15: by the difficulty score (guide pairs: the best cheap predictor of the remaining work, Spearman 0.8
// with it — tools/dev/predict_probe.py)
__device__ __forceinline__ int park_bucket(const ParkQueue& park, const TrajState& ps) {
  const int np = park.score_mode == 0 ? ps.n_pairs : (park.score_mode == 1 ? ps.n_pairs * ps.N / 32 : ps.n_pairs * 8 + ps.astar_expansions / 64);
  const int ex = ps.astar_expansions, t = park.thresh;
  if (ex >= 32 * t) return 0;
  if (ex >= 16 * t) return 1;
  if (ex >= 8 * t) return 2;
  if (ex >= 4 * t || ps.astar_unreach) return 3;
  if (ex >= 2 * t) return 4;
  if (ex >= t) return 5;
  return np >= 24 ? 6 : (np >= 16 ? 7 : (np >= 12 ? 8 : (np >= 8 ? 9 : (np >= 6 ? 10 : (np >= 4 ? 11 : (np >= 3 ? 12 : (np >= 2 ? 13 : (np >= 1 ? 14 : 15))))))));
}

struct SolveLayout {
  int st;      // TrajState (doubles offset)
  int vf;      // start of the solver region (cp first)
  int plan;    // PlanSmem (aliases the solver region after cp)
  int total;   // doubles
};
__host__ __device__ inline SolveLayout solve_layout(int N, int mode, int m) {
  SolveLayout L;
  int o = 0;
  L.st = o; o += (int)((sizeof(TrajState) + 7) / 8);
  L.vf = o;
  const int cp_d = 3 * N + (N & 1);
  // mode 4 (team form): the solver region starts with the WfShared context, then the control points
  const int solver = mode == 4 ? wf_layout(N).total : (mode >= 2 ? vf_layout(N, mode == 3).total : (int)lbfgs_smem_doubles(N, m) + (N & 1));
  L.plan = L.vf + cp_d + (mode == 4 ? WF_CTX : 0);
  const int plan_end = L.plan + (int)((sizeof(PlanSmem) + 7) / 8);
  const int rp_end = L.plan + 3 * (N - 1) + 3 * (N - 2) + 2 * TP_LB_WARPS + 4;
  int end = L.vf + solver;
  if (plan_end > end) end = plan_end;
  if (rp_end > end) end = rp_end;
  L.total = end;
  return L;
}

// Returns 1 when the trajectory was PARKED after makePlan steps 1-3 (park_thresh >= 0 and its first A* searches took
// fewer than park_thresh expansions): its state is back in HBM and a later solve_one(resume = 1) finishes it.
template <int MODE>
__device__ __forceinline__ int solve_one(const BatchView& bv, const VigoConst& C, const DevMap& map, const AStarPools& P,
                                         int b, int class_max_n, int s_slot, int sw, uint32_t tbase, double* counters,
                                         long long* timeline, int resume, int rounds_this_pass, int park_thresh, double* sm) {
  const int tid = threadIdx.x, lane = tid & 31;
  long long t_start = 0;
  if (timeline && tid == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_start));
  const SolveLayout SL = solve_layout(class_max_n, MODE, C.p.lbfgs_m);
  TrajState& st = *reinterpret_cast<TrajState*>(sm + SL.st);
  double* cp = sm + SL.vf + (MODE == 4 ? WF_CTX : 0);
  PlanSmem& PS = *reinterpret_cast<PlanSmem*>(sm + SL.plan);
  // ---- load state + control points
  {
    const int* src = reinterpret_cast<const int*>(&bv.st[b]);
    int* dst = reinterpret_cast<int*>(&st);
    for (int i = tid; i < (int)(sizeof(TrajState) / 4); i += TP_LB_THREADS) dst[i] = src[i];
  }
  __syncthreads();
  const int N = st.N, n = 3 * (N - 2 * TP_DEGREE);
  double* gctrl = bv.ctrl + 3 * (size_t)st.off;
  for (int e = tid; e < 3 * N; e += TP_LB_THREADS) cp[e] = gctrl[e];
  __syncthreads();
  const bool is_serial_warp = (tid >> 5) == sw;
  BatchView lbv = bv;                      // the outer-loop code reads control points through bv.ctrl + 3*off
  lbv.ctrl = cp - 3 * (size_t)st.off;
  Worker W = make_worker(C, P, s_slot, lane, &PS.as);
#ifdef TP_LBFGS_TIMING
  const long long pinit0 = clock64();
#endif
  // ---- makePlan steps 1-3 (first pass only; a resumed trajectory continues its optimise / check / re-guide loop:
  // every optimize() starts from a fresh L-BFGS state, so splitting a solve at a round boundary changes nothing)
  if (!resume) {
    if (is_serial_warp) dev_plan_init(lbv, C, map, st, W, PS, b, lane);
  } else if (st.astar_unreach) {
    W.flood_trigger = TP_FLOOD_TRIGGER_AGAIN;
  }
  __syncthreads();
  if (!resume && park_thresh >= 0 && is_serial_warp && lane == 0 && W.goal_unreachable) st.astar_unreach = 1;
  __syncthreads();
  if (!resume && park_thresh >= 0 && st.status == TS_ACTIVE) {
    // park it (state only — the control points have not changed): the worker goes on sweeping the batch through
    // makePlan steps 1-3, and whoever drains the parked lists resumes it at the optimise / check loop
    const int* src = reinterpret_cast<const int*>(&st);
    int* dst = reinterpret_cast<int*>(&bv.st[b]);
    for (int i = tid; i < (int)(sizeof(TrajState) / 4); i += TP_LB_THREADS) dst[i] = src[i];
    if (timeline && tid == 0) timeline[4 * (size_t)b] = t_start;
    __syncthreads();
    return 1;
  }
  double fl = 0.0, its = 0.0, evs = 0.0, smp = 0.0;
  int rounds_done = 0;
  // time slicing: a NEGATIVE round limit -r means "the first r rounds of this trajectory only" — the first turn is short
  // (most trajectories finish in it; the survivors reveal themselves early), every later turn runs to the end
  if (rounds_this_pass < 0) rounds_this_pass = st.lbfgs_runs < -rounds_this_pass ? -rounds_this_pass - st.lbfgs_runs : 0x7fffffff;
#ifdef TP_LBFGS_TIMING
  long long ph[4] = {0, 0, 0, 0}, pt0 = clock64(), pt1;
  const long long pstart = pinit0;
  ph[0] = pt0 - pinit0;
#define PH(i) { pt1 = clock64(); ph[i] += pt1 - pt0; pt0 = pt1; }
#else
#define PH(i)
#endif
  PH(0)
  while (st.status == TS_ACTIVE && rounds_done < rounds_this_pass) {
    ++rounds_done;
    // ---- optimize()
    tp_lbfgs_result r;
    if (n <= 0) {
      r.ret = LB_INVALID_N; r.iters = 0; r.evals = 0; r.reserved = 0; r.fx = 0;
    } else if (MODE == 4) {
      wf_setup<TP_LB_WARPS>(C, sm + SL.vf, N, bv.pairs + (size_t)b * C.gcap, bv.cp_head + st.off, st.n_pairs, st.w_dist, st.w_dyn,
                            bv.n_dyn, bv.dyn_pos, bv.dyn_vel, bv.dyn_size, tid);
      lbfgs_run_team<TP_LB_WARPS>(C, sm + SL.vf, N, sw, r, nullptr, tid);
    } else if (MODE >= 2) {
      VfCtx V;
      V.N = N; V.n = n; V.sm = cp; V.L = vf_layout(N, MODE == 3);
      V.tbase = tbase;
      V.pairs = bv.pairs + (size_t)b * C.gcap; V.head = bv.cp_head + st.off; V.n_pairs = st.n_pairs; V.pairs_in_sm = false;
      V.serial_warp = sw;
      V.w_dist = st.w_dist; V.w_dyn = st.w_dyn;
      V.n_dyn = bv.n_dyn; V.dyn_pos = bv.dyn_pos; V.dyn_vel = bv.dyn_vel; V.dyn_size = bv.dyn_size;
      lbfgs_run_fast<MODE == 3>(C, V, r, nullptr, tid);
    } else {
      EvalCtx E;
      E.N = N; E.n = n; E.cp = cp;
      E.pairs = bv.pairs + (size_t)b * C.gcap;
      E.head = bv.cp_head + st.off;
      E.w_dist = st.w_dist; E.w_dyn = st.w_dyn;
      E.n_dyn = bv.n_dyn; E.dyn_pos = bv.dyn_pos; E.dyn_vel = bv.dyn_vel; E.dyn_size = bv.dyn_size;
      lbfgs_run<MODE == 1>(C, E, cp + 3 * N + (N & 1), r, nullptr, tid);
    }
    if (tid == 0) {
      st.lbfgs_runs += 1;
      st.lbfgs_iters += r.iters;
      st.lbfgs_evals += r.evals;
      st.last_ret = r.ret;
      st.final_cost = r.fx;
      if (n > 0) {
        st.vclock += (long long)r.evals * (10LL * N + 2LL * n);
        // algorithmic FP64 flops: E(81N + 21G + 4n) + sum_k (8 b_k + 15) n
        fl += (double)r.evals * (81.0 * N + 21.0 * st.n_pairs + 4.0 * n) + (8.0 * r.reserved + 15.0 * r.iters) * n;
        its += r.iters;
        evs += r.evals;
      }
    }
    PH(1)
    // ---- hasCollisionTrajectory
    int any = 1;
    if (N >= 4) any = dev_has_collision(cp, N, bv, C, map, tid);
    if (tid == 0) {
      st.has_col = any;
      smp += floor((double)(N - TP_DEGREE) * C.p.ctrl_pt_ts / C.check_ts) + 1.0;
    }
    __syncthreads();
    PH(2)
    // ---- loop body: success / failure / re-guide / weight doubling
    if (is_serial_warp) {
      dev_plan_step(lbv, C, map, st, W, PS, b, lane);
      if (lane == 0 && W.goal_unreachable) st.astar_unreach = 1;
    }
    __syncthreads();
    PH(3)
  }
#ifdef TP_LBFGS_TIMING
  if (tid == 0)
    printf("[solve] b %d N %d iters %d exp %d rounds %d | kcycles: total %lld init %lld lbfgs %lld collision %lld step %lld\n", b, N,
           st.lbfgs_iters, st.astar_expansions, rounds_done, (clock64() - pstart) / 1000, ph[0] / 1000, ph[1] / 1000, ph[2] / 1000, ph[3] / 1000);
#endif
  // ---- linearFeasibilityReparam
  if (st.status == TP_STATUS_SUCCESS) {
    double* q = sm + SL.plan;
    double* rr = q + 3 * (N - 1);
    double* red = rr + 3 * (N - 2) + (N & 1);
    const double f = dev_reparam(cp, q, rr, red, N, bv, C, tid);
    if (tid == 0) st.linear_factor = f;
  }
  __syncthreads();
  // ---- write back
  for (int e = tid; e < 3 * N; e += TP_LB_THREADS) gctrl[e] = cp[e];
  {
    const int* src = reinterpret_cast<const int*>(&st);
    int* dst = reinterpret_cast<int*>(&bv.st[b]);
    for (int i = tid; i < (int)(sizeof(TrajState) / 4); i += TP_LB_THREADS) dst[i] = src[i];
  }
  if (tid == 0) {
    if (counters) {
      atomicAdd(&counters[0], fl);
      atomicAdd(&counters[1], its);
      atomicAdd(&counters[2], evs);
      atomicAdd(&counters[3], smp);
    }
    if (timeline) {   // development aid (TP_TIMELINE): start / end time [ns], SM id, iterations of every block
      long long t_end;
      unsigned smid;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_end));
      asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
      if (!resume) timeline[4 * (size_t)b] = t_start;
      timeline[4 * (size_t)b + 1] = t_end;
      timeline[4 * (size_t)b + 2] = resume ? t_start : (long long)smid;   // resumed: when its phase B turn began
      timeline[4 * (size_t)b + 3] = (long long)st.lbfgs_iters | ((long long)st.astar_expansions << 32);
    }
  }
  const int unfinished = st.status == TS_ACTIVE;   // the pass ended at its round limit (time slicing / two-pass scheduling)
  __syncthreads();
  return unfinished ? 2 : 0;
}

// Persistent workers: one launch per size class; a worker of class c owns shared memory for that class's longest
// trajectory and serves its own class first, then STEALS from the classes of shorter trajectories (they fit), so
// every worker stays busy until the whole batch is drained.  Each class's list is ordered hardest-first.
// cls_begin[5]: ranges of `order` per class (largest trajectories = class 0); cls_next[4]: atomic cursors.
template <int MODE>
__global__ void __launch_bounds__(TP_LB_THREADS, MODE == 3 ? 4 : (MODE == 4 ? TP_TEAM_BLOCKS : 3)) k_solve(const __grid_constant__ BatchView bv, const __grid_constant__ VigoConst C, const __grid_constant__ DevMap map, const __grid_constant__ AStarPools P,
                                                            const int* __restrict__ order, const int* __restrict__ cls_begin,
                                                            int* cls_next, int my_class, int class_max_n,
                                                            int* slot_flags, double* counters, long long* timeline,
                                                            int resume, int rounds_this_pass, ParkQueue park) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x;
  const SolveLayout SL = solve_layout(class_max_n, MODE, C.p.lbfgs_m);
  int* pick = reinterpret_cast<int*>(sm + SL.total);   // 3 ints past the layout (the launch reserves 16 B)
  uint32_t* tm_slot = reinterpret_cast<uint32_t*>(pick + 2);
  uint32_t tbase = 0;
  if (MODE == 3) tbase = tm_block_alloc(tm_slot, tid);   // the worker's L-BFGS history columns, held until it exits
  // ---- claim an A* node pool (one per resident worker) and a serial-warp role.  Warp w of every block lives on SM
  // sub-partition w, so the workers sharing an SM take DIFFERENT warps for their serial phases (A*, guide points, the
  // coefficient recurrences): a per-SM ticket rotates the role.
  if (tid == 0) {
    const int W_ = P.workers;
    int sidx = (int)(((unsigned)(blockIdx.x + 977 * my_class) * 2654435761u) % (unsigned)W_);
    while (atomicCAS(&slot_flags[sidx], 0, 1) != 0) sidx = sidx + 1 == W_ ? 0 : sidx + 1;
    unsigned smid;
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    const int ticket = atomicAdd(&slot_flags[W_ + (int)(smid & 255u)], 1);
    pick[1] = sidx | ((ticket & (TP_LB_WARPS - 1)) << 24);
  }
  __syncthreads();
  const int s_slot = pick[1] & 0xFFFFFF;
  const int sw = (pick[1] >> 24) & (TP_LB_WARPS - 1);
  // ---- phase A: sweep the class queues through makePlan steps 1-3 (segments, A*, guide points) and park every
  // trajectory that is still active, in a bucket that says how hard it looks NOW: long first searches / unreachable
  // goals (the batch's tail if they started late) first, then by the number of guide pairs.
  // ---- phase B (main queues drained): resume the parked trajectories, own class first, then the classes of shorter
  // ones, hardest bucket first.  Only slots a parker has reserved are claimed; once every trajectory of the class has
  // been through phase A the lists are final and an empty scan ends the class.
  int pb_class = -1;   // thread 0: >= 0 once this worker is in phase B
  const bool tail_worker = park.rev > 0 && (blockIdx.x % park.rev) == park.rev - 1;
  for (;;) {
    if (tid == 0) {
      int b = -1, cc = 0, res = resume;
      if (resume == 2) {
        // k_phase_a runs beside this kernel: prefer parked trajectories (phase B: first the ones that have not started,
        // hardest-looking first, then the ring of those waiting for their next turn); with none ready, help with phase A (so
        // the batch drains whoever is resident); leave once no trajectory of the classes served is unfinished
        for (;;) {
          bool all_done = true;
          for (int c = my_class; c < 4; ++c)   // read BEFORE the scans: 0 = nothing of that class can show up any more
            if (cls_begin[c + 1] > cls_begin[c] && *((volatile int*)&park.remaining[c]) > 0) all_done = false;
          // not started yet: BUCKET-major over the classes this worker can serve, so that a worker of a long-trajectory
          // class takes a hard-looking short trajectory before an easy-looking one of its own class (order within the
          // buckets: hardest-looking first, or easiest first for one worker in `park.rev`)
          for (int kk = 0; kk < TP_PARK_BUCKETS && b < 0; ++kk) {
            const int k = tail_worker ? TP_PARK_BUCKETS - 1 - kk : kk;
            for (int c = my_class; c < 4 && b < 0; ++c) {
              if (cls_begin[c + 1] == cls_begin[c]) continue;
              const int q = c * TP_PARK_BUCKETS + k;
              for (;;) {
                const int h = *((volatile int*)&park.head[q]);
                if (h >= *((volatile int*)&park.tail[q])) break;
                if (atomicCAS(&park.head[q], h, h + 1) != h) continue;
                while ((b = *((volatile int*)&park.list[(size_t)q * park.stride + h])) < 0) __nanosleep(100);
                break;
              }
              if (b >= 0) cc = c;
            }
          }
          for (int c = my_class; c < 4 && b < 0 && park.slice > 0; ++c) {   // next turn of a trajectory that has had one
            if (cls_begin[c + 1] == cls_begin[c]) continue;
            while (b < 0) {
              const int h = *((volatile int*)&park.rhead[c]);
              if (h >= *((volatile int*)&park.rtail[c])) break;
              if (atomicCAS(&park.rhead[c], h, h + 1) != h) continue;
              int* slot = &park.ring[(size_t)c * park.stride + (h % park.stride)];
              while ((b = atomicExch(slot, -1)) < 0) __nanosleep(100);
            }
            if (b >= 0) cc = c;
          }
          if (b >= 0) { __threadfence(); res = 1; break; }
          for (int c = my_class; c < 4 && b < 0; ++c) {
            if (cls_begin[c + 1] == cls_begin[c] || *((volatile int*)&cls_next[c]) >= cls_begin[c + 1] - cls_begin[c]) continue;
            const int i = cls_begin[c] + atomicAdd(&cls_next[c], 1);
            if (i < cls_begin[c + 1]) { b = order[i]; cc = c; }
          }
          if (b >= 0) { res = 0; break; }
          if (all_done) break;
          __nanosleep(500);
        }
      } else if (pb_class < 0) {
        for (int c = my_class; c < 4 && b < 0; ++c) {
          if (cls_begin[c + 1] == cls_begin[c]) continue;
          const int i = cls_begin[c] + atomicAdd(&cls_next[c], 1);
          if (i < cls_begin[c + 1]) { b = order[i]; cc = c; }
        }
        if (b < 0 && park.thresh >= 0) pb_class = my_class;
      }
      while (resume != 2 && b < 0 && pb_class >= 0 && pb_class < 4) {
        const int c = pb_class, csize = cls_begin[c + 1] - cls_begin[c];
        if (csize == 0) { ++pb_class; continue; }
        // buckets of the class, hardest-looking first; only slots that a parker has already reserved are claimed
        const bool final_pass = *((volatile int*)&park.started[c]) >= csize;   // read BEFORE the scan: lists are final
        for (int k = 0; k < TP_PARK_BUCKETS && b < 0; ++k) {
          const int q = c * TP_PARK_BUCKETS + k;
          for (;;) {
            const int h = *((volatile int*)&park.head[q]);
            if (h >= *((volatile int*)&park.tail[q])) break;
            if (atomicCAS(&park.head[q], h, h + 1) != h) continue;
            while ((b = *((volatile int*)&park.list[(size_t)q * park.stride + h])) < 0) __nanosleep(100);
            break;
          }
        }
        if (b >= 0) { __threadfence(); res = 1; cc = c; }
        else if (final_pass) ++pb_class;
        else __nanosleep(500);
      }
      pick[0] = b;
      pick[3] = cc | (res << 8);
    }
    __syncthreads();
    const int b = pick[0];
    if (b < 0) break;
    const int res = pick[3] >> 8;
    const int parked = solve_one<MODE>(bv, C, map, P, b, class_max_n, s_slot, sw, tbase, counters, timeline, res,
                                       (resume == 2 && res && park.slice > 0) ? -park.slice : rounds_this_pass, res ? -1 : park.thresh, sm);
    if (resume == 2 && tid == 0 && parked != 1) {
      const int cc = pick[3] & 255;
      if (parked == 2) {   // still active after its turn: to the back of the class's ring
        const int pos = atomicAdd(&park.rtail[cc], 1);
        int* slot = &park.ring[(size_t)cc * park.stride + (pos % park.stride)];
        __threadfence();                                  // the trajectory's state is in HBM before its id shows up
        while (atomicCAS(slot, -1, b) != -1) __nanosleep(100);
      } else {
        __threadfence();
        atomicSub(&park.remaining[cc], 1);
      }
    }
    if (park.thresh >= 0 && !res && tid == 0) {
      const int cc = pick[3] & 255;
      if (parked == 1) {
        // bucket by the guide pairs the first searches produced (the best cheap predictor of the remaining work)
        const TrajState* ps = reinterpret_cast<const TrajState*>(sm + SL.st);
        const int bk = park_bucket(park, *ps);
        const int q = cc * TP_PARK_BUCKETS + bk;
        const int pos = atomicAdd(&park.tail[q], 1);
        __threadfence();                                  // the trajectory's state is in HBM before its id shows up
        atomicExch(&park.list[(size_t)q * park.stride + pos], b);
      }
      __threadfence();
      atomicAdd(&park.started[cc], 1);                    // makePlan steps 1-3 of one more trajectory of class cc are done
    }
  }
  if (tid == 0) {
    __threadfence();
    atomicExch(&slot_flags[s_slot], 0);
  }
  if (MODE == 3) tm_block_free(tm_slot, tid);
}

// Phase A as its own launch (team-form batches): makePlan steps 1-3 (segments, A*, guide points) are serial work for ONE
// warp, and a team block (128 threads, 45-96 KB of shared memory, 168 registers) keeps three warps at a barrier while its
// serial warp searches — an SM ran three searches at a time (ncu on maze.bt: 75 % of all samples at that barrier).  This
// kernel runs the same dev_plan_init with one warp per worker and only the A* scratch in shared memory (30 KB: seven
// workers per SM), parks every trajectory that is still active into the SAME per-class lists k_solve's phase B drains
// (k_solve runs beside it with resume = 2: phase B only), and exits as the queues run dry, which hands the SM over to the
// team blocks.  Results do not depend on who ran phase A (same device functions, same order of operations per trajectory).
__global__ void __launch_bounds__(32) k_phase_a(const __grid_constant__ BatchView bv, const __grid_constant__ VigoConst C,
                                                const __grid_constant__ DevMap map, const __grid_constant__ AStarPools P,
                                                const int* __restrict__ order, const int* __restrict__ cls_begin, int* cls_next,
                                                int* slot_flags, long long* timeline, ParkQueue park) {
  __shared__ PlanSmem S;
  const int lane = threadIdx.x;
  int s_slot = 0;
  if (lane == 0) {
    const int W_ = P.workers;
    int sidx = (int)(((unsigned)(blockIdx.x + 7919) * 2654435761u) % (unsigned)W_);
    while (atomicCAS(&slot_flags[sidx], 0, 1) != 0) sidx = sidx + 1 == W_ ? 0 : sidx + 1;
    s_slot = sidx;
  }
  s_slot = __shfl_sync(0xffffffffu, s_slot, 0);
  Worker W = make_worker(C, P, s_slot, lane, &S.as);
  for (;;) {
    int b = -1, cc = 0;
    if (lane == 0)
      for (int c = 0; c < 4 && b < 0; ++c) {
        if (cls_begin[c + 1] == cls_begin[c]) continue;
        const int i = cls_begin[c] + atomicAdd(&cls_next[c], 1);
        if (i < cls_begin[c + 1]) { b = order[i]; cc = c; }
      }
    b = __shfl_sync(0xffffffffu, b, 0);
    if (b < 0) break;
    cc = __shfl_sync(0xffffffffu, cc, 0);
    long long t_start = 0;
    if (timeline && lane == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_start));
    TrajState& st = bv.st[b];
    W.flood_trigger = TP_FLOOD_TRIGGER;
    W.goal_unreachable = 0;
    W.flood_wasted = 0;
    dev_plan_init(bv, C, map, st, W, S, b, lane);
    if (lane == 0) {
      if (W.goal_unreachable) st.astar_unreach = 1;
      if (timeline) timeline[4 * (size_t)b] = t_start;
      if (st.status == TS_ACTIVE) {
        const int bk = park_bucket(park, st);
        const int q = cc * TP_PARK_BUCKETS + bk;
        const int pos = atomicAdd(&park.tail[q], 1);
        __threadfence();                                  // the trajectory's state is in HBM before its id shows up
        atomicExch(&park.list[(size_t)q * park.stride + pos], b);
      } else {   // finished in phase A (failed A*, already collision free, ...)
        if (timeline) {
          long long t_end;
          asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_end));
          timeline[4 * (size_t)b + 1] = t_end;
          timeline[4 * (size_t)b + 3] = (long long)st.astar_expansions << 32;
        }
        __threadfence();
        atomicSub(&park.remaining[cc], 1);
      }
      __threadfence();
      atomicAdd(&park.started[cc], 1);
    }
    __syncwarp();
  }
  if (lane == 0) {
    __threadfence();
    atomicExch(&slot_flags[s_slot], 0);
  }
}

#include "tp_solve_warp.cuh"

// standalone A* (parity entry): one warp per (start, end) pair
__global__ void __launch_bounds__(32) k_astar(VigoConst C, DevMap map, AStarPools P, int* queue, int S_,
                                              const double* __restrict__ starts, const double* __restrict__ ends,
                                              int* path_len, double* paths, int* expansions) {
  __shared__ AStarSmem as_sm;
  const int lane = threadIdx.x;
  Worker W = make_worker(C, P, blockIdx.x, lane, &as_sm);
  for (;;) {
    int s = 0;
    if (lane == 0) s = atomicAdd(queue, 1);
    s = __shfl_sync(0xffffffffu, s, 0);
    if (s >= S_) break;
    int ex = 0, err = 0;
    const int len = astar_search(map, C, W, d3(starts[3 * s], starts[3 * s + 1], starts[3 * s + 2]),
                                 d3(ends[3 * s], ends[3 * s + 1], ends[3 * s + 2]), ex, err);
    if (lane == 0) {
      path_len[s] = len;
      expansions[s] = ex;
    }
    if (len > 0)
      for (int e = lane; e < 3 * len; e += 32) paths[(size_t)s * C.path_cap * 3 + e] = W.path[e];
    __syncwarp();
  }
}

// standalone findCollisionSeg (parity entry)
__global__ void __launch_bounds__(32) k_find_seg(BatchView bv, VigoConst C, DevMap map, int* nseg_out, int* segs_out) {
  __shared__ PlanSmem S;
  const int b = blockIdx.x, lane = threadIdx.x;
  const TrajState& st = bv.st[b];
  int err = 0;
  int n = 0;
  if (st.N >= 2 * TP_DEGREE + 1 && st.N <= TP_MAX_CTRL) n = find_collision_seg(map, C, bv, st, S.hit, S.line, S.segA, lane, err);
  if (lane == 0) {
    nseg_out[b] = n;
    for (int i = 0; i < n; ++i) {
      segs_out[((size_t)b * C.max_seg + i) * 2] = S.segA[i][0];
      segs_out[((size_t)b * C.max_seg + i) * 2 + 1] = S.segA[i][1];
    }
  }
}

// difficulty proxy for the issue order: number of optimised control points inside inflated obstacles
__global__ void k_count_colliding(BatchView bv, DevMap map, int* out) {
  const int b = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= bv.B) return;
  const int o = bv.off[b], N = bv.off[b + 1] - o;
  int c = 0;
  for (int i = TP_DEGREE + lane; i <= N - TP_DEGREE - 1; i += 32)
    c += dm_inflated(map, bv.ctrl[3 * (size_t)(o + i)], bv.ctrl[3 * (size_t)(o + i) + 1], bv.ctrl[3 * (size_t)(o + i) + 2]) ? 1 : 0;
  for (int q = 16; q > 0; q >>= 1) c += __shfl_xor_sync(0xffffffffu, c, q);
  if (lane == 0) out[b] = c;
}

// difficulty key after the first round (see tp_vigo_make_plan_batch): -1 when the trajectory is finished
__global__ void k_difficulty(BatchView bv, int* out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bv.B) return;
  const TrajState& st = bv.st[b];
  out[b] = st.status == TS_ACTIVE ? st.lbfgs_iters + (3 * st.astar_expansions) / 10 + 5 * st.n_pairs : -1;
}

__global__ void k_collect_results(BatchView bv, tp_vigo_result* out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bv.B) return;
  const TrajState& st = bv.st[b];
  tp_vigo_result r;
  r.status = st.status == TS_ACTIVE ? TP_STATUS_FAIL_OPTIMIZE : st.status;
  if (st.err != 0 && r.status != TP_STATUS_SUCCESS) r.status = TP_STATUS_FAIL_CAPACITY;
  r.outer_rounds = st.round;
  r.fail_count = st.fail_count;
  r.lbfgs_runs = st.lbfgs_runs;
  r.lbfgs_iters = st.lbfgs_iters;
  r.lbfgs_evals = st.lbfgs_evals;
  r.astar_searches = st.astar_searches;
  r.astar_expansions = st.astar_expansions;
  r.n_guide_pairs = st.n_pairs;
  r.last_lbfgs_ret = st.last_ret;
  r.final_cost = st.final_cost;
  r.linear_factor = st.linear_factor;
  out[b] = r;
}

#include "tp_poly.cuh"
#include "tp_corridor.cuh"
#include "tp_sample.cuh"

// =========================================================================== engine
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return TP_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      tp_set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
      return TP_ERR_CUDA;
    }
    cap = want;
    return TP_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() { return (T*)p; }
};

struct tp_engine {
  int device = 0;
  tp_engine_cfg cfg;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  int64_t launches = 0;
  // map replica
  bool has_map = false;
  DevMap dmap;
  DevBuf map_infl, map_known, map_occ;
  double known_bbmin[3] = {0, 0, 0}, known_bbmax[3] = {0, 0, 0};   // metric bounding box of the known cells
  DevBuf poly_scratch, poly_tacc;
  int poly_tacc_n = 0;
  double poly_tacc_dt = -1;
  double map_res = 0;
  // tables (depend on params + map res)
  DevBuf t_check, t_reparam, a_line;
  int n_t_check = 0, n_t_reparam = 0, n_a_line = 0;
  double tab_check_ts = -1, tab_ts = -1, tab_res = -1;
  // A* pools
  AStarPools pools;
  DevBuf pool_nodes, pool_heaps, pool_heapn, pool_paths, pool_sc, pool_sclen, pool_rounds, pool_flags;
  cudaStream_t class_stream[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_stage = nullptr, ev_join[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool stage_busy = false;
  bool solve_attr_set = false;
  int pools_key[8] = {0};
  // batch buffers
  DevBuf off, ctrl, st, pairs, cp_head, cp_tail, active[2], counters, results, dyn, scratch_a, scratch_b, scratch_c, parkq;
  int* h_counters = nullptr;  // pinned
  // pinned staging for host-memory calls
  void* h_stage = nullptr;
  size_t h_stage_cap = 0;
  bool lbfgs_attr_set = false;
  int max_smem_optin = 0;
  bool fe_attr_set = false;
  // measurement
  bool profile = false;
  DevBuf dev_counters;  // 8 doubles
  struct ProfEntry { int kind; int n; cudaEvent_t a, b; };
  std::vector<ProfEntry> prof_entries;
  std::vector<cudaEvent_t> ev_pool;
  double query_points = 0;
  double* counters_ptr() { return profile ? dev_counters.as<double>() : nullptr; }
};

// RAII per-launch timer: records two events around a launch when profiling is on
struct ProfScope {
  tp_engine* e;
  cudaStream_t s;
  cudaEvent_t b = nullptr;
  int kind;
  ProfScope(tp_engine* e_, int kind_, cudaStream_t s_, int n_ = 0) : e(e_), s(s_), kind(kind_) {
    if (!e->profile) return;
    cudaEvent_t a;
    auto get = [&]() {
      cudaEvent_t ev;
      if (!e->ev_pool.empty()) { ev = e->ev_pool.back(); e->ev_pool.pop_back(); }
      else cudaEventCreate(&ev);
      return ev;
    };
    a = get();
    b = get();
    cudaEventRecord(a, s);
    e->prof_entries.push_back({kind, n_, a, b});
  }
  ~ProfScope() {
    if (b) cudaEventRecord(b, s);
  }
};

static int ensure_stage(tp_engine* e, size_t bytes) {
  if (bytes <= e->h_stage_cap) return TP_OK;
  if (e->h_stage) cudaFreeHost(e->h_stage);
  e->h_stage = nullptr;
  e->h_stage_cap = 0;
  size_t want = bytes + bytes / 4 + 4096;
  cudaError_t er = cudaMallocHost(&e->h_stage, want);
  if (er != cudaSuccess) {
    tp_set_error("cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(er));
    return TP_ERR_CUDA;
  }
  e->h_stage_cap = want;
  return TP_OK;
}

static void make_const(const tp_engine* e, const tp_vigo_params* p, VigoConst& C) {
  memset(&C, 0, sizeof(C));
  C.p = *p;
  const double dth = p->dthresh;
  C.dist_a = 3.0 * dth;                       // bsplineTraj.cpp:835
  C.dist_b = -3.0 * (dth * dth);
  C.dist_c = std::pow(dth, 3);
  const double hth = 0.2;                     // :836-837
  C.h_a = 3.0 * hth;
  C.h_b = -3 * (hth * hth);
  C.h_c = std::pow(hth, 3);
  const double dd = p->dthresh_dyn;           // :1009
  C.dyn_a = 3.0 * dd;
  C.dyn_b = -3 * (dd * dd);
  C.dyn_c = std::pow(dd, 3);
  C.ts_inv_sqr = 1 / (p->ctrl_pt_ts * p->ctrl_pt_ts);  // :959
  C.check_ts = e->map_res / p->max_vel / 2.0;           // bsplineTraj.h:312
  C.pred_num = (int)(p->pred_horizon / p->ts);          // bsplineTraj.cpp:1007
  for (int a = 0; a < 3; ++a) C.pool[a] = 2 * (int)(p->max_obstacle_size[a] / e->map_res);  // :191-193
  int kl = (int)((p->max_height - p->min_height) / e->map_res) + 3;
  if (kl > C.pool[2]) kl = C.pool[2];
  if (kl > 254) kl = 254;   // packed node ids keep the layer in 8 bits (255 = spare slot)
  if (kl < 1) kl = 1;
  C.pool_kl = kl;
  C.max_seg = e->cfg.max_segments;
  C.gcap = e->cfg.max_guide_pairs;
  C.path_cap = e->cfg.max_path_cells;
  size_t pool_nodes = (size_t)C.pool[0] * C.pool[1] * C.pool_kl;
  size_t hc = e->cfg.astar_heap_cap > 0 ? (size_t)e->cfg.astar_heap_cap : std::min(pool_nodes + 1, (size_t)1 << 20);
  C.heap_cap = (int)hc;
}

// sample-time tables: the reference accumulates t += dt in FP64; reproduce the same sums once
static int ensure_tables(tp_engine* e, VigoConst& C) {
  const int max_n = TP_MAX_CTRL;
  const double max_dur = (double)(max_n - TP_DEGREE) * C.p.ctrl_pt_ts;
  if (e->tab_check_ts != C.check_ts) {
    std::vector<double> t;
    for (double x = 0; x <= max_dur; x += C.check_ts) {
      t.push_back(x);
      if (t.size() > (size_t)4000000) break;
    }
    if (e->t_check.ensure(t.size() * 8) != TP_OK) return TP_ERR_CUDA;
    CK(cudaMemcpyAsync(e->t_check.p, t.data(), t.size() * 8, cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->n_t_check = (int)t.size();
    e->tab_check_ts = C.check_ts;
  }
  if (e->tab_ts != C.p.ts) {
    std::vector<double> t;
    for (double x = 0.0; x < max_dur; x += C.p.ts) {
      t.push_back(x);
      if (t.size() > (size_t)4000000) break;
    }
    if (e->t_reparam.ensure(t.size() * 8) != TP_OK) return TP_ERR_CUDA;
    CK(cudaMemcpyAsync(e->t_reparam.p, t.data(), t.size() * 8, cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->n_t_reparam = (int)t.size();
    e->tab_ts = C.p.ts;
  }
  if (e->tab_res != e->map_res) {
    std::vector<double> t;
    for (double a = 0.0; a <= 1.0; a += e->map_res) {  // bsplineTraj.h:197
      t.push_back(a);
      if (t.size() > (size_t)100000) break;
    }
    if (e->a_line.ensure(t.size() * 8) != TP_OK) return TP_ERR_CUDA;
    CK(cudaMemcpyAsync(e->a_line.p, t.data(), t.size() * 8, cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->n_a_line = (int)t.size();
    e->tab_res = e->map_res;
  }
  C.n_t_check = e->n_t_check;
  C.n_t_reparam = e->n_t_reparam;
  C.n_a_line = e->n_a_line;
  return TP_OK;
}

static int ensure_pools(tp_engine* e, const VigoConst& C) {
  const int key[8] = {C.pool[0], C.pool[1], C.pool_kl, C.heap_cap, C.path_cap, C.max_seg, 0, 0};
  if (memcmp(key, e->pools_key, sizeof(key)) == 0 && e->pools.workers > 0) return TP_OK;
  const size_t pool_nodes = (size_t)C.pool[0] * C.pool[1] * C.pool_kl;
  const size_t per_worker = (pool_nodes + 1) * sizeof(ANode) + (size_t)C.heap_cap * 12 + (size_t)C.path_cap * 24 +
                            (size_t)C.max_seg * TP_SC_CAP * 24 + (size_t)C.max_seg * 4 + 4;
  int workers = e->cfg.astar_workers;
  if (workers <= 0) {
    const double budget = e->cfg.astar_mem_gb * 1e9;
    size_t freeb = 0, totalb = 0;
    cudaMemGetInfo(&freeb, &totalb);
    double use = std::min(budget, 0.5 * (double)freeb);
    workers = (int)std::min<double>(use / (double)per_worker, (double)e->sm_count * 11);
    workers = (workers / e->sm_count) * e->sm_count;
    if (workers < e->sm_count) workers = std::max(1, (int)(use / (double)per_worker));
  }
  if (workers < 1) workers = 1;
  if (e->pool_nodes.ensure((size_t)workers * (pool_nodes + 1) * sizeof(ANode)) != TP_OK) return TP_ERR_CUDA;
  if (e->pool_heaps.ensure((size_t)workers * C.heap_cap * 8) != TP_OK) return TP_ERR_CUDA;
  if (e->pool_heapn.ensure((size_t)workers * C.heap_cap * 4) != TP_OK) return TP_ERR_CUDA;
  if (e->pool_paths.ensure((size_t)workers * C.path_cap * 24) != TP_OK) return TP_ERR_CUDA;
  if (e->pool_sc.ensure((size_t)workers * C.max_seg * TP_SC_CAP * 24) != TP_OK) return TP_ERR_CUDA;
  if (e->pool_sclen.ensure((size_t)workers * C.max_seg * 4) != TP_OK) return TP_ERR_CUDA;
  if (e->pool_rounds.ensure((size_t)workers * 4) != TP_OK) return TP_ERR_CUDA;
  if (e->pool_flags.ensure(((size_t)workers + 256 + 2) * 4) != TP_OK) return TP_ERR_CUDA;   // pool slots + per-SM tickets + flood statistics
  CK(cudaMemsetAsync(e->pool_flags.p, 0, ((size_t)workers + 256 + 2) * 4, e->stream));
  CK(cudaMemsetAsync(e->pool_nodes.p, 0, (size_t)workers * (pool_nodes + 1) * sizeof(ANode), e->stream));
  CK(cudaMemsetAsync(e->pool_rounds.p, 0, (size_t)workers * 4, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  e->pools.nodes = e->pool_nodes.as<ANode>();
  e->pools.heap_k = e->pool_heaps.as<double>();
  e->pools.heap_n = e->pool_heapn.as<uint32_t>();
  e->pools.paths = e->pool_paths.as<double>();
  e->pools.sc = e->pool_sc.as<double>();
  e->pools.sc_len = e->pool_sclen.as<int>();
  e->pools.rounds = e->pool_rounds.as<uint32_t>();
  e->pools.flood_stats = e->pool_flags.as<int>() + workers + 256;
  e->pools.pool_nodes = pool_nodes;
  e->pools.workers = workers;
  memcpy(e->pools_key, key, sizeof(key));
  return TP_OK;
}

static int check_params(const tp_engine* e, const tp_vigo_params* p) {
  if (!e || !p) { tp_set_error("null engine/params"); return TP_ERR_INVALID_ARG; }
  if (!e->has_map) { tp_set_error("engine has no map: call tp_engine_set_map first"); return TP_ERR_NO_MAP; }
  if (p->lbfgs_m < 1 || p->lbfgs_m > 64 || !(p->ctrl_pt_ts > 0) || !(p->max_vel > 0) || !(p->ts > 0) ||
      p->lbfgs_max_linesearch < 1) {
    tp_set_error("invalid ViGO parameters");
    return TP_ERR_INVALID_ARG;
  }
  for (int a = 0; a < 3; ++a) {
    const int pool = 2 * (int)(p->max_obstacle_size[a] / e->map_res);
    if (pool < 4 || pool > TP_AXIS_MAX) {
      tp_set_error("A* pool of %d cells along axis %d (2*int(max_obstacle_size/res)) is outside [4, %d]", pool, a, TP_AXIS_MAX);
      return TP_ERR_CAPACITY;
    }
  }
  return TP_OK;
}

// Batch set-up shared by all ViGO entry points: offsets/ctrl on the device, states initialised.
struct BatchSetup {
  BatchView bv;
  VigoConst C;
  int max_n = 0;
  long total = 0;
  std::vector<int> h_off;
};

static int setup_batch(tp_engine* e, const tp_vigo_params* p, int B, const int32_t* offsets, const double* ctrl, int mem,
                       cudaStream_t s, BatchSetup& bs, bool need_guides) {
  int rc = check_params(e, p);
  if (rc != TP_OK) return rc;
  if (B <= 0 || !offsets || !ctrl) { tp_set_error("empty batch or null pointers"); return TP_ERR_INVALID_ARG; }
  make_const(e, p, bs.C);
  rc = ensure_tables(e, bs.C);
  if (rc != TP_OK) return rc;
  bs.h_off.resize(B + 1);
  if (mem == TP_MEM_DEVICE) {
    CK(cudaMemcpyAsync(bs.h_off.data(), offsets, (size_t)(B + 1) * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  } else
    memcpy(bs.h_off.data(), offsets, (size_t)(B + 1) * 4);
  bs.max_n = 0;
  for (int b = 0; b < B; ++b) {
    const int n = bs.h_off[b + 1] - bs.h_off[b];
    if (n < 0) { tp_set_error("offsets must be non-decreasing"); return TP_ERR_INVALID_ARG; }
    if (n > TP_MAX_CTRL) { tp_set_error("trajectory %d has %d control points (max %d)", b, n, TP_MAX_CTRL); return TP_ERR_CAPACITY; }
    bs.max_n = std::max(bs.max_n, n);
  }
  bs.total = bs.h_off[B];
  const size_t tot = (size_t)std::max<long>(bs.total, 1);
  if (e->off.ensure((size_t)(B + 1) * 4) != TP_OK || e->st.ensure((size_t)B * sizeof(TrajState)) != TP_OK ||
      e->cp_head.ensure(tot * 4) != TP_OK || e->cp_tail.ensure(tot * 4) != TP_OK ||
      e->pairs.ensure((size_t)B * bs.C.gcap * sizeof(GuidePair)) != TP_OK || e->counters.ensure(64 * 4) != TP_OK ||
      e->active[0].ensure((size_t)B * 4) != TP_OK || e->active[1].ensure((size_t)B * 4) != TP_OK)
    return TP_ERR_CUDA;
  BatchView& bv = bs.bv;
  memset(&bv, 0, sizeof(bv));
  bv.B = B;
  bv.total_pts = (int)bs.total;
  if (mem == TP_MEM_DEVICE) {
    bv.off = offsets;
    bv.ctrl = const_cast<double*>(ctrl);
  } else {
    if (e->ctrl.ensure(tot * 24) != TP_OK) return TP_ERR_CUDA;
    CK(cudaMemcpyAsync(e->off.p, offsets, (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(e->ctrl.p, ctrl, (size_t)bs.total * 24, cudaMemcpyHostToDevice, s));
    bv.off = e->off.as<int>();
    bv.ctrl = e->ctrl.as<double>();
  }
  bv.st = e->st.as<TrajState>();
  bv.pairs = e->pairs.as<GuidePair>();
  bv.cp_head = e->cp_head.as<int>();
  bv.cp_tail = e->cp_tail.as<int>();
  bv.t_check = e->t_check.as<double>();
  bv.t_reparam = e->t_reparam.as<double>();
  bv.a_line = e->a_line.as<double>();
  k_init_states<<<(B + 127) / 128, 128, 0, s>>>(bv, bs.C);
  k_fill_int<<<std::min<long>((long)(tot + 255) / 256, 1184), 256, 0, s>>>(bv.cp_head, (long)tot, -1);
  k_fill_int<<<std::min<long>((long)(tot + 255) / 256, 1184), 256, 0, s>>>(bv.cp_tail, (long)tot, -1);
  e->launches += 3;
  (void)need_guides;
  return TP_OK;
}

// flat guide list (host memory) -> device linked lists; optional per-trajectory weights
static int upload_guides(tp_engine* e, BatchSetup& bs, const int32_t* g_offsets, const int32_t* g_cp, const double* g_p,
                         const double* g_v, const double* w_override, cudaStream_t s) {
  const int B = bs.bv.B;
  const int gcap = bs.C.gcap;
  std::vector<TrajState> hst((size_t)B);
  CK(cudaMemcpyAsync(hst.data(), bs.bv.st, (size_t)B * sizeof(TrajState), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  std::vector<int> head((size_t)std::max<long>(bs.total, 1), -1), tail(head.size(), -1);
  std::vector<GuidePair> pairs((size_t)B * gcap);
  memset(pairs.data(), 0, pairs.size() * sizeof(GuidePair));
  const tp_map* hm = nullptr;
  (void)hm;
  for (int b = 0; b < B; ++b) {
    if (w_override) {
      hst[b].w_dist = w_override[2 * b];
      hst[b].w_dyn = w_override[2 * b + 1];
    }
    if (!g_offsets) continue;
    const int g0 = g_offsets[b], g1 = g_offsets[b + 1];
    if (g1 - g0 > gcap) { tp_set_error("trajectory %d has %d guide pairs (engine max_guide_pairs %d)", b, g1 - g0, gcap); return TP_ERR_CAPACITY; }
    for (int g = g0; g < g1; ++g) {
      const int gi = g - g0;
      GuidePair& pr = pairs[(size_t)b * gcap + gi];
      for (int a = 0; a < 3; ++a) { pr.p[a] = g_p[3 * g + a]; pr.v[a] = g_v[3 * g + a]; }
      pr.next = -1;
      pr.unknown = -1;  // resolved on the device below
      const int c = g_cp[g];
      if (c < 0 || c >= hst[b].N) { tp_set_error("guide pair %d: control point %d out of range", g, c); return TP_ERR_INVALID_ARG; }
      const int ci = hst[b].off + c;
      if (tail[ci] < 0) head[ci] = gi; else pairs[(size_t)b * gcap + tail[ci]].next = gi;
      tail[ci] = gi;
    }
    hst[b].n_pairs = g1 - g0;
  }
  CK(cudaMemcpyAsync(bs.bv.st, hst.data(), (size_t)B * sizeof(TrajState), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(bs.bv.cp_head, head.data(), head.size() * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(bs.bv.cp_tail, tail.data(), tail.size() * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(bs.bv.pairs, pairs.data(), pairs.size() * sizeof(GuidePair), cudaMemcpyHostToDevice, s));
  CK(cudaStreamSynchronize(s));
  return TP_OK;
}

__global__ void k_resolve_unknown(BatchView bv, VigoConst C, DevMap map) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= (long)bv.B * C.gcap) return;
  const int b = (int)(i / C.gcap), gi = (int)(i % C.gcap);
  if (gi >= bv.st[b].n_pairs) return;
  GuidePair& pr = bv.pairs[i];
  pr.unknown = dm_unknown(map, d3(pr.p[0], pr.p[1], pr.p[2])) ? 1 : 0;
}

// which fused cost+L-BFGS kernel runs: 1 strict (bit-faithful); 2 vector-free with the history in shared memory +
// FP64 MMA Gram update (default); 3 vector-free with the history in tensor memory (TP_LBFGS_TMEM=1: 4 workers per SM
// instead of 3, measured slower per iteration so far); 0 classic two-loop with tree reductions (lbfgs_m != 16, or
// TP_LBFGS_CLASSIC=1)
// 4 = lean team form (tp_lbfgs_warp.cuh) with a team of four warps per trajectory, the default; 5 = the same arithmetic
// family with a team of ONE warp (TP_WARP_FORM=1: its results differ in rounding from mode 4's, both are pinned to the
// oracle); TP_LBFGS_BLOCK=1 keeps the round-1 forms 2 / 3 for A/B measurements
static int lbfgs_mode(const tp_vigo_params* p) {
  if (p->strict_order) return 1;
  static const bool classic = getenv("TP_LBFGS_CLASSIC") != nullptr;
  static const bool tmem_hist = getenv("TP_LBFGS_TMEM") != nullptr;
  static const bool block_form = getenv("TP_LBFGS_BLOCK") != nullptr || tmem_hist;
  static const bool warp_form = getenv("TP_WARP_FORM") != nullptr;
  if (p->lbfgs_m != VF_M || classic) return 0;
  return block_form ? (tmem_hist ? 3 : 2) : (warp_form ? 5 : 4);
}
static size_t lbfgs_smem_bytes(const tp_vigo_params* p, int max_n) {
  const int mode = lbfgs_mode(p);
  if (mode >= 4) return (size_t)wf_layout(max_n).total * 8;
  return mode >= 2 ? vf_smem_bytes(max_n, mode == 3) + 16 : lbfgs_smem_doubles(max_n, p->lbfgs_m) * 8;
}
template <class... A>
static void launch_lbfgs(int mode, int grid, size_t smem, cudaStream_t s, A... args) {
  if (mode == 5) k_lbfgs_w<<<grid, 32, smem, s>>>(args...);
  else if (mode == 4) k_lbfgs<4><<<grid, TP_LB_THREADS, smem, s>>>(args...);
  else if (mode == 1) k_lbfgs<1><<<grid, TP_LB_THREADS, smem, s>>>(args...);
  else if (mode == 2) k_lbfgs<2><<<grid, TP_LB_THREADS, smem, s>>>(args...);
  else if (mode == 3) k_lbfgs<3><<<grid, TP_LB_THREADS, smem, s>>>(args...);
  else k_lbfgs<0><<<grid, TP_LB_THREADS, smem, s>>>(args...);
}

static int set_lbfgs_smem(tp_engine* e, size_t bytes) {
  if ((int)bytes > e->max_smem_optin) {
    tp_set_error("L-BFGS state needs %zu B of shared memory (> %d B per block)", bytes, e->max_smem_optin);
    return TP_ERR_CAPACITY;
  }
  if (!e->lbfgs_attr_set) {
    CK(cudaFuncSetAttribute(k_lbfgs<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, e->max_smem_optin));
    CK(cudaFuncSetAttribute(k_lbfgs<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, e->max_smem_optin));
    CK(cudaFuncSetAttribute(k_lbfgs<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, e->max_smem_optin));
    CK(cudaFuncSetAttribute(k_lbfgs<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, e->max_smem_optin));
    CK(cudaFuncSetAttribute(k_cost<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, e->max_smem_optin));
    CK(cudaFuncSetAttribute(k_cost<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, e->max_smem_optin));
    CK(cudaFuncSetAttribute(k_lbfgs<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, e->max_smem_optin));
    CK(cudaFuncSetAttribute(k_cost_t, cudaFuncAttributeMaxDynamicSharedMemorySize, e->max_smem_optin));
    CK(cudaFuncSetAttribute(k_lbfgs_w, cudaFuncAttributeMaxDynamicSharedMemorySize, e->max_smem_optin));
    CK(cudaFuncSetAttribute(k_cost_w, cudaFuncAttributeMaxDynamicSharedMemorySize, e->max_smem_optin));
    e->lbfgs_attr_set = true;
  }
  return TP_OK;
}

// =========================================================================== C ABI
extern "C" {

int tp_version(void) { return 100; }
int tp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}
void tp_engine_default_cfg(tp_engine_cfg* c) {
  memset(c, 0, sizeof(*c));
  c->astar_workers = 0;
  c->max_segments = 32;
  c->max_guide_pairs = 256;
  c->astar_heap_cap = 0;
  c->max_path_cells = 4096;
  c->lbfgs_threads = TP_LB_THREADS;
  c->astar_mem_gb = 48.0;
}
void tp_vigo_default_params(tp_vigo_params* p) {
  // cfg/bspline_interactive/bspline_planner_param.yaml + desired_velocity/acceleration
  // (src/bspline_node.cpp:210-211,230-231)
  memset(p, 0, sizeof(*p));
  p->ts = 0.1; p->dthresh = 0.5; p->max_vel = 2.0; p->max_acc = 3.0;
  p->w_distance = 1.0; p->w_smooth = 1.0; p->w_feas = 1.0; p->w_dyn = 1.0;
  p->min_height = 0.7; p->max_height = 1.3; p->uncertain_factor = 1.0;
  p->pred_horizon = 2.0; p->dthresh_dyn = 0.5; p->max_path_length = 20.0;
  p->max_obstacle_size[0] = 5; p->max_obstacle_size[1] = 5; p->max_obstacle_size[2] = 3;
  p->ctrl_pt_dist = 0.25; p->ctrl_pt_ts = 0.2; p->not_check_ratio = 0.0;
  p->lbfgs_g_eps = 0.01; p->plan_in_z = 0; p->lbfgs_m = 16; p->lbfgs_max_iter = 200;
  p->lbfgs_max_linesearch = 40; p->max_outer_rounds = 24; p->astar_max_expansions = 200000;
  p->strict_order = 0;
  p->vclock_budget = 3000000;
}

tp_engine_t* tp_engine_create(int device, const tp_engine_cfg* cfg) {
  int n = 0;
  cudaError_t er = cudaGetDeviceCount(&n);
  if (er != cudaSuccess || n <= 0) {
    tp_set_error("no CUDA device available (%s); this engine has no CPU fallback",
                 er != cudaSuccess ? cudaGetErrorString(er) : "device count 0");
    return nullptr;
  }
  if (device < 0 || device >= n) {
    tp_set_error("device %d out of range (count %d)", device, n);
    return nullptr;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    tp_set_error("cudaSetDevice(%d) failed", device);
    return nullptr;
  }
  tp_engine* e = new tp_engine();
  e->device = device;
  if (cfg) e->cfg = *cfg; else tp_engine_default_cfg(&e->cfg);
  if (e->cfg.max_segments <= 0) e->cfg.max_segments = 32;
  if (e->cfg.max_segments > TP_MAX_SEG_HARD) e->cfg.max_segments = TP_MAX_SEG_HARD;
  if (e->cfg.max_guide_pairs <= 0) e->cfg.max_guide_pairs = 256;
  if (e->cfg.max_path_cells <= 0) e->cfg.max_path_cells = 4096;
  if (!(e->cfg.astar_mem_gb > 0)) e->cfg.astar_mem_gb = 48.0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMallocHost((void**)&e->h_counters, 64 * sizeof(int)) != cudaSuccess) {
    tp_set_error("engine initialisation failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete e;
    return nullptr;
  }
  e->sm_count = prop.multiProcessorCount;
  e->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  memset(&e->pools, 0, sizeof(e->pools));
  return e;
}
void tp_engine_destroy(tp_engine_t* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  DevBuf* bufs[] = {&e->map_occ, &e->poly_scratch, &e->poly_tacc, &e->map_infl, &e->map_known, &e->t_check, &e->t_reparam, &e->a_line, &e->pool_nodes, &e->pool_heaps, &e->pool_heapn,
                    &e->pool_paths, &e->pool_sc, &e->pool_sclen, &e->pool_rounds, &e->pool_flags, &e->off, &e->ctrl, &e->st, &e->pairs,
                    &e->cp_head, &e->cp_tail, &e->active[0], &e->active[1], &e->counters, &e->results, &e->dyn,
                    &e->scratch_a, &e->scratch_b, &e->scratch_c, &e->parkq};
  for (DevBuf* b : bufs) b->release();
  if (e->h_counters) cudaFreeHost(e->h_counters);
  if (e->h_stage) cudaFreeHost(e->h_stage);
  for (int i = 0; i < 8; ++i) {
    if (e->class_stream[i]) cudaStreamDestroy(e->class_stream[i]);
    if (e->ev_join[i]) cudaEventDestroy(e->ev_join[i]);
  }
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_stage) cudaEventDestroy(e->ev_stage);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}
int tp_engine_set_map(tp_engine_t* e, const tp_map_t* m) {
  if (!e || !m) return TP_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  std::vector<uint32_t> wi, wk;
  m->pack(2, wi);
  m->pack(1, wk);
  if (e->map_infl.ensure(wi.size() * 4) != TP_OK || e->map_known.ensure(wk.size() * 4) != TP_OK) return TP_ERR_CUDA;
  CK(cudaMemcpy(e->map_infl.p, wi.data(), wi.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(e->map_known.p, wk.data(), wk.size() * 4, cudaMemcpyHostToDevice));
  if (e->pools.workers > 0) CK(cudaMemset(e->pools.flood_stats, 0, 2 * sizeof(int)));   // the statistics belong to the map
  {
    std::vector<uint32_t> wo;
    m->pack(0, wo);
    if (e->map_occ.ensure(wo.size() * 4) != TP_OK) return TP_ERR_CUDA;
    CK(cudaMemcpy(e->map_occ.p, wo.data(), wo.size() * 4, cudaMemcpyHostToDevice));
    // bounding box of the known cells (the polyTraj collision contract's getMetricMin/Max)
    int lo[3] = {m->dims[0], m->dims[1], m->dims[2]}, hi[3] = {-1, -1, -1};
    for (int ix = 0; ix < m->dims[0]; ++ix)
      for (int iy = 0; iy < m->dims[1]; ++iy)
        for (int iz = 0; iz < m->dims[2]; ++iz)
          if (m->known[m->addr(ix, iy, iz)]) {
            const int v[3] = {ix, iy, iz};
            for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], v[a]); hi[a] = std::max(hi[a], v[a]); }
          }
    for (int a = 0; a < 3; ++a) {
      e->known_bbmin[a] = m->origin[a] + (hi[a] >= 0 ? lo[a] : 0) * m->res;
      e->known_bbmax[a] = m->origin[a] + (hi[a] >= 0 ? hi[a] + 1 : 0) * m->res;
    }
  }
  e->dmap.inflated = e->map_infl.as<uint32_t>();
  e->dmap.known = e->map_known.as<uint32_t>();
  e->dmap.res = m->res;
  e->dmap.inv_res = 1.0 / m->res;
  for (int a = 0; a < 3; ++a) { e->dmap.mn[a] = m->origin[a]; e->dmap.dim[a] = m->dims[a]; }
  e->dmap.wz = m->wz();
  e->map_res = m->res;
  e->has_map = true;
  return TP_OK;
}
int tp_engine_synchronize(tp_engine_t* e) {
  if (!e) return TP_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  return TP_OK;
}
int64_t tp_engine_launch_count(const tp_engine_t* e) { return e ? e->launches : 0; }
void* tp_engine_stream(tp_engine_t* e) { return e ? (void*)e->stream : nullptr; }

static cudaStream_t pick_stream(tp_engine* e, void* s) { return s ? (cudaStream_t)s : e->stream; }

static int query_common(tp_engine_t* e, int64_t n, const double* a, const double* b, uint8_t* out, int mem, void* stream,
                        int kind) {
  if (!e) return TP_ERR_INVALID_ARG;
  if (!e->has_map) { tp_set_error("engine has no map"); return TP_ERR_NO_MAP; }
  if (n < 0 || (n > 0 && (!a || !out))) return TP_ERR_INVALID_ARG;
  if (n == 0) return TP_OK;
  CK(cudaSetDevice(e->device));
  cudaStream_t s = pick_stream(e, stream);
  const double* da = a;
  const double* db = b;
  uint8_t* dout = out;
  if (mem == TP_MEM_HOST) {
    if (e->scratch_a.ensure((size_t)n * 24) != TP_OK || e->scratch_c.ensure((size_t)n) != TP_OK) return TP_ERR_CUDA;
    CK(cudaMemcpyAsync(e->scratch_a.p, a, (size_t)n * 24, cudaMemcpyHostToDevice, s));
    da = e->scratch_a.as<double>();
    if (kind == 2) {
      if (e->scratch_b.ensure((size_t)n * 24) != TP_OK) return TP_ERR_CUDA;
      CK(cudaMemcpyAsync(e->scratch_b.p, b, (size_t)n * 24, cudaMemcpyHostToDevice, s));
      db = e->scratch_b.as<double>();
    }
    dout = e->scratch_c.as<uint8_t>();
  }
  const int threads = 256;
  // point queries: four per thread and iteration, a grid of whole waves (8 blocks of 256 threads fit an SM)
  const long work = kind == 2 ? n : (n + 3) / 4;
  // resident blocks per SM the point-query kernel is compiled for (registers: 4 -> 62, 6 -> 40, 8 -> 32 with a small spill);
  // the grid is exactly one resident wave of the grid-stride loop
  static const int minb = getenv("TP_QUERY_MINB") ? atoi(getenv("TP_QUERY_MINB")) : 4;
  const long blocks = std::min<long>((work + threads - 1) / threads, (long)e->sm_count * (kind == 2 ? 8 : (minb >= 8 ? 8 : (minb >= 6 ? 6 : 4))));
  {
    ProfScope ps(e, 5, s);
    if (kind == 2) k_query_lines<<<(int)blocks, threads, 0, s>>>(e->dmap, (long)n, da, db, dout);
    else if (minb >= 8) k_query_points<8><<<(int)blocks, threads, 0, s>>>(e->dmap, (long)n, da, dout, kind);
    else if (minb >= 6) k_query_points<6><<<(int)blocks, threads, 0, s>>>(e->dmap, (long)n, da, dout, kind);
    else k_query_points<4><<<(int)blocks, threads, 0, s>>>(e->dmap, (long)n, da, dout, kind);
  }
  e->query_points += (double)n;
  e->launches += 1;
  CK(cudaGetLastError());
  if (mem == TP_MEM_HOST) {
    CK(cudaMemcpyAsync(out, dout, (size_t)n, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  return TP_OK;
}
int tp_query_points(tp_engine_t* e, int64_t n, const double* xyz, uint8_t* hit, int mem, void* stream) {
  return query_common(e, n, xyz, nullptr, hit, mem, stream, 0);
}
int tp_query_unknown(tp_engine_t* e, int64_t n, const double* xyz, uint8_t* unknown, int mem, void* stream) {
  return query_common(e, n, xyz, nullptr, unknown, mem, stream, 1);
}
int tp_query_lines(tp_engine_t* e, int64_t n, const double* a, const double* b, uint8_t* hit, int mem, void* stream) {
  if (n > 0 && !b) return TP_ERR_INVALID_ARG;
  return query_common(e, n, a, b, hit, mem, stream, 2);
}

// identity active list + its count in counters[0]
static int make_identity_active(tp_engine* e, int B, cudaStream_t s) {
  k_iota<<<(B + 255) / 256, 256, 0, s>>>(e->active[0].as<int>(), B);
  e->launches += 1;
  e->h_counters[0] = B;
  CK(cudaMemcpyAsync(e->counters.p, e->h_counters, 4, cudaMemcpyHostToDevice, s));
  return TP_OK;
}

// dynamic obstacles (always host pointers: a handful of values) -> the batch view
static int upload_dyn(tp_engine* e, BatchSetup& bs, int n_dyn, const double* dyn_pos, const double* dyn_vel, const double* dyn_size,
                      cudaStream_t s) {
  if (n_dyn < 0 || (n_dyn > 0 && (!dyn_pos || !dyn_vel || !dyn_size))) { tp_set_error("dynamic obstacles: bad arguments"); return TP_ERR_INVALID_ARG; }
  if (n_dyn == 0) return TP_OK;
  if (e->dyn.ensure((size_t)n_dyn * 72) != TP_OK) return TP_ERR_CUDA;
  double* d = e->dyn.as<double>();
  CK(cudaMemcpyAsync(d, dyn_pos, (size_t)n_dyn * 24, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d + 3 * (size_t)n_dyn, dyn_vel, (size_t)n_dyn * 24, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d + 6 * (size_t)n_dyn, dyn_size, (size_t)n_dyn * 24, cudaMemcpyHostToDevice, s));
  bs.bv.n_dyn = n_dyn;
  bs.bv.dyn_pos = d;
  bs.bv.dyn_vel = d + 3 * (size_t)n_dyn;
  bs.bv.dyn_size = d + 6 * (size_t)n_dyn;
  return TP_OK;
}

int tp_vigo_cost_batch(tp_engine_t* e, const tp_vigo_params* p, int32_t B, const int32_t* offsets, const double* ctrl,
                       const int32_t* g_offsets, const int32_t* g_cp, const double* g_p, const double* g_v,
                       const double* w_override, double* f, double* grad, int mem, void* stream) {
  return tp_vigo_cost_batch_dyn(e, p, B, offsets, ctrl, g_offsets, g_cp, g_p, g_v, w_override, 0, nullptr, nullptr, nullptr, f, grad,
                                mem, stream);
}

int tp_vigo_cost_batch_dyn(tp_engine_t* e, const tp_vigo_params* p, int32_t B, const int32_t* offsets, const double* ctrl,
                           const int32_t* g_offsets, const int32_t* g_cp, const double* g_p, const double* g_v,
                           const double* w_override, int32_t n_dyn, const double* dyn_pos, const double* dyn_vel,
                           const double* dyn_size, double* f, double* grad, int mem, void* stream) {
  if (mem != TP_MEM_HOST) { tp_set_error("tp_vigo_cost_batch: host memory only"); return TP_ERR_INVALID_ARG; }
  if (!e || !f || !grad) return TP_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  cudaStream_t s = pick_stream(e, stream);
  BatchSetup bs;
  int rc = setup_batch(e, p, B, offsets, ctrl, mem, s, bs, true);
  if (rc != TP_OK) return rc;
  rc = upload_dyn(e, bs, n_dyn, dyn_pos, dyn_vel, dyn_size, s);
  if (rc != TP_OK) return rc;
  rc = upload_guides(e, bs, g_offsets, g_cp, g_p, g_v, w_override, s);
  if (rc != TP_OK) return rc;
  k_resolve_unknown<<<(int)(((long)B * bs.C.gcap + 255) / 256), 256, 0, s>>>(bs.bv, bs.C, e->dmap);
  const size_t nvar = (size_t)std::max<long>(3 * (bs.total - 6L * B), 1);
  if (e->scratch_a.ensure((size_t)B * 8) != TP_OK || e->scratch_b.ensure(nvar * 8) != TP_OK) return TP_ERR_CUDA;
  const int cmode = lbfgs_mode(p);
  const size_t smem = p->strict_order ? ((size_t)3 * bs.max_n + 3 * (size_t)bs.max_n + 40) * 8
                                      : (cmode >= 4 ? (size_t)wf_layout(bs.max_n).total * 8 : vf_smem_bytes(bs.max_n));
  rc = set_lbfgs_smem(e, smem);
  if (rc != TP_OK) return rc;
  if (cmode == 4) k_cost_t<<<B, TP_LB_THREADS, smem, s>>>(bs.bv, bs.C, e->scratch_a.as<double>(), e->scratch_b.as<double>());
  else if (cmode == 5) k_cost_w<<<B, 32, smem, s>>>(bs.bv, bs.C, e->scratch_a.as<double>(), e->scratch_b.as<double>());
  else if (p->strict_order) k_cost<true><<<B, TP_LB_THREADS, smem, s>>>(bs.bv, bs.C, e->scratch_a.as<double>(), e->scratch_b.as<double>());
  else k_cost<false><<<B, TP_LB_THREADS, smem, s>>>(bs.bv, bs.C, e->scratch_a.as<double>(), e->scratch_b.as<double>());
  e->launches += 2;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(f, e->scratch_a.p, (size_t)B * 8, cudaMemcpyDeviceToHost, s));
  // ragged gradient: trajectory b's slice starts at 3*(off[b] - 6b); trajectories with N < 7 own nothing
  CK(cudaMemcpyAsync(grad, e->scratch_b.p, nvar * 8, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return TP_OK;
}

int tp_vigo_optimize_batch(tp_engine_t* e, const tp_vigo_params* p, int32_t B, const int32_t* offsets, double* ctrl,
                           const int32_t* g_offsets, const int32_t* g_cp, const double* g_p, const double* g_v,
                           const double* w_override, tp_lbfgs_result* res, double* x_final, int mem, void* stream) {
  return tp_vigo_optimize_batch_dyn(e, p, B, offsets, ctrl, g_offsets, g_cp, g_p, g_v, w_override, 0, nullptr, nullptr, nullptr, res,
                                    x_final, mem, stream);
}

int tp_vigo_optimize_batch_dyn(tp_engine_t* e, const tp_vigo_params* p, int32_t B, const int32_t* offsets, double* ctrl,
                               const int32_t* g_offsets, const int32_t* g_cp, const double* g_p, const double* g_v,
                               const double* w_override, int32_t n_dyn, const double* dyn_pos, const double* dyn_vel,
                               const double* dyn_size, tp_lbfgs_result* res, double* x_final, int mem, void* stream) {
  if (mem != TP_MEM_HOST) { tp_set_error("tp_vigo_optimize_batch: host memory only"); return TP_ERR_INVALID_ARG; }
  if (!e) return TP_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  cudaStream_t s = pick_stream(e, stream);
  BatchSetup bs;
  int rc = setup_batch(e, p, B, offsets, ctrl, mem, s, bs, true);
  if (rc != TP_OK) return rc;
  rc = upload_dyn(e, bs, n_dyn, dyn_pos, dyn_vel, dyn_size, s);
  if (rc != TP_OK) return rc;
  rc = upload_guides(e, bs, g_offsets, g_cp, g_p, g_v, w_override, s);
  if (rc != TP_OK) return rc;
  k_resolve_unknown<<<(int)(((long)B * bs.C.gcap + 255) / 256), 256, 0, s>>>(bs.bv, bs.C, e->dmap);
  const size_t nvar = (size_t)std::max<long>(3 * (bs.total - 6L * B), 1);
  if (e->results.ensure((size_t)B * sizeof(tp_lbfgs_result)) != TP_OK || e->scratch_b.ensure(nvar * 8) != TP_OK) return TP_ERR_CUDA;
  CK(cudaMemsetAsync(e->results.p, 0, (size_t)B * sizeof(tp_lbfgs_result), s));
  const size_t smem = lbfgs_smem_bytes(p, bs.max_n);
  rc = set_lbfgs_smem(e, smem);
  if (rc != TP_OK) return rc;
  rc = make_identity_active(e, B, s);
  if (rc != TP_OK) return rc;
  launch_lbfgs(lbfgs_mode(p), B, smem, s, bs.bv, bs.C, (const int*)e->active[0].as<int>(), (const int*)e->counters.as<int>(),
               e->results.as<tp_lbfgs_result>(), x_final ? e->scratch_b.as<double>() : (double*)nullptr, (double*)nullptr);
  e->launches += 2;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(ctrl, bs.bv.ctrl, (size_t)bs.total * 24, cudaMemcpyDeviceToHost, s));
  if (res) CK(cudaMemcpyAsync(res, e->results.p, (size_t)B * sizeof(tp_lbfgs_result), cudaMemcpyDeviceToHost, s));
  if (x_final) CK(cudaMemcpyAsync(x_final, e->scratch_b.p, nvar * 8, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return TP_OK;
}

int tp_vigo_has_collision_batch(tp_engine_t* e, const tp_vigo_params* p, int32_t B, const int32_t* offsets,
                                const double* ctrl, uint8_t* hit, int mem, void* stream) {
  if (mem != TP_MEM_HOST) { tp_set_error("tp_vigo_has_collision_batch: host memory only"); return TP_ERR_INVALID_ARG; }
  if (!e || !hit) return TP_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  cudaStream_t s = pick_stream(e, stream);
  BatchSetup bs;
  int rc = setup_batch(e, p, B, offsets, ctrl, mem, s, bs, false);
  if (rc != TP_OK) return rc;
  if (e->scratch_c.ensure((size_t)B) != TP_OK) return TP_ERR_CUDA;
  CK(cudaMemsetAsync(e->scratch_c.p, 0, (size_t)B, s));
  rc = make_identity_active(e, B, s);
  if (rc != TP_OK) return rc;
  k_has_collision<<<B, TP_LB_THREADS, (size_t)3 * bs.max_n * 8, s>>>(bs.bv, bs.C, e->dmap, e->active[0].as<int>(),
                                                                      e->counters.as<int>(), e->scratch_c.as<uint8_t>(), nullptr);
  e->launches += 1;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(hit, e->scratch_c.p, (size_t)B, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return TP_OK;
}

int tp_vigo_sample_batch(tp_engine_t* e, double ctrl_pt_ts, int32_t B, const int32_t* offsets, const double* ctrl,
                         const int32_t* t_offsets, const double* t, double* pos, double* vel, double* acc, double* yaw,
                         int mem, void* stream) {
  if (!e || B < 0 || !(ctrl_pt_ts > 0)) return TP_ERR_INVALID_ARG;
  if (B == 0) return TP_OK;
  if (!offsets || !ctrl || !t_offsets || !t || !pos) return TP_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  cudaStream_t s = pick_stream(e, stream);
  const int blocks_cap = e->sm_count * 8;
  if (mem == TP_MEM_DEVICE) {
    // sample count unknown on the host: a grid-stride launch sized for the machine
    k_sample_traj<<<blocks_cap, 256, 0, s>>>(B, offsets, ctrl, t_offsets, t, ctrl_pt_ts, pos, vel, acc, yaw);
    e->launches += 1;
    CK(cudaGetLastError());
    return TP_OK;
  }
  if (mem != TP_MEM_HOST) return TP_ERR_INVALID_ARG;
  if (offsets[0] != 0 || t_offsets[0] != 0) { tp_set_error("tp_vigo_sample_batch: offsets[0] and t_offsets[0] must be 0"); return TP_ERR_INVALID_ARG; }
  for (int b = 0; b < B; ++b)
    if (offsets[b + 1] < offsets[b] || t_offsets[b + 1] < t_offsets[b]) { tp_set_error("tp_vigo_sample_batch: offsets must be non-decreasing"); return TP_ERR_INVALID_ARG; }
  const size_t np = (size_t)offsets[B], nt = (size_t)t_offsets[B];
  if (nt == 0) return TP_OK;
  const size_t ob = ((size_t)(B + 1) * 4 + 15) & ~(size_t)15;
  // scratch_a: offsets | t_offsets | ctrl | t      scratch_b: pos | vel | acc | yaw
  if (e->scratch_a.ensure(2 * ob + (3 * np + nt) * 8) != TP_OK || e->scratch_b.ensure(10 * nt * 8) != TP_OK) return TP_ERR_CUDA;
  char* a = e->scratch_a.as<char>();
  int* d_off = (int*)a;
  int* d_toff = (int*)(a + ob);
  double* d_ctrl = (double*)(a + 2 * ob);
  double* d_t = d_ctrl + 3 * np;
  double* d_pos = e->scratch_b.as<double>();
  double* d_vel = d_pos + 3 * nt;
  double* d_acc = d_vel + 3 * nt;
  double* d_yaw = d_acc + 3 * nt;
  CK(cudaMemcpyAsync(d_off, offsets, (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_toff, t_offsets, (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_ctrl, ctrl, 3 * np * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_t, t, nt * 8, cudaMemcpyHostToDevice, s));
  const int blocks = (int)std::min<size_t>((nt + 255) / 256, (size_t)blocks_cap);
  k_sample_traj<<<blocks, 256, 0, s>>>(B, d_off, d_ctrl, d_toff, d_t, ctrl_pt_ts, d_pos, vel ? d_vel : nullptr, acc ? d_acc : nullptr,
                                       yaw ? d_yaw : nullptr);
  e->launches += 1;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(pos, d_pos, 3 * nt * 8, cudaMemcpyDeviceToHost, s));
  if (vel) CK(cudaMemcpyAsync(vel, d_vel, 3 * nt * 8, cudaMemcpyDeviceToHost, s));
  if (acc) CK(cudaMemcpyAsync(acc, d_acc, 3 * nt * 8, cudaMemcpyDeviceToHost, s));
  if (yaw) CK(cudaMemcpyAsync(yaw, d_yaw, nt * 8, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return TP_OK;
}

int tp_vigo_find_collision_seg_batch(tp_engine_t* e, const tp_vigo_params* p, int32_t B, const int32_t* offsets,
                                     const double* ctrl, int32_t* nseg, int32_t* segs, int mem, void* stream) {
  if (mem != TP_MEM_HOST) { tp_set_error("tp_vigo_find_collision_seg_batch: host memory only"); return TP_ERR_INVALID_ARG; }
  if (!e || !nseg || !segs) return TP_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  cudaStream_t s = pick_stream(e, stream);
  BatchSetup bs;
  int rc = setup_batch(e, p, B, offsets, ctrl, mem, s, bs, false);
  if (rc != TP_OK) return rc;
  const size_t nsegs = (size_t)B * bs.C.max_seg * 2;
  if (e->scratch_a.ensure((size_t)B * 4) != TP_OK || e->scratch_b.ensure(nsegs * 4) != TP_OK) return TP_ERR_CUDA;
  CK(cudaMemsetAsync(e->scratch_b.p, 0, nsegs * 4, s));
  k_find_seg<<<B, 32, 0, s>>>(bs.bv, bs.C, e->dmap, e->scratch_a.as<int>(), e->scratch_b.as<int>());
  e->launches += 1;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(nseg, e->scratch_a.p, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(segs, e->scratch_b.p, nsegs * 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return TP_OK;
}

int tp_astar_batch(tp_engine_t* e, const tp_vigo_params* p, int32_t S, const double* starts, const double* ends,
                   int32_t* path_len, double* paths, int32_t* expansions, int mem, void* stream) {
  if (mem != TP_MEM_HOST) { tp_set_error("tp_astar_batch: host memory only"); return TP_ERR_INVALID_ARG; }
  int rc = check_params(e, p);
  if (rc != TP_OK) return rc;
  if (S <= 0 || !starts || !ends || !path_len || !paths || !expansions) return TP_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  cudaStream_t s = pick_stream(e, stream);
  VigoConst C;
  make_const(e, p, C);
  rc = ensure_tables(e, C);
  if (rc != TP_OK) return rc;
  rc = ensure_pools(e, C);
  if (rc != TP_OK) return rc;
  const size_t pbytes = (size_t)S * C.path_cap * 24;
  if (e->scratch_a.ensure((size_t)S * 48) != TP_OK || e->scratch_b.ensure(pbytes) != TP_OK ||
      e->scratch_c.ensure((size_t)S * 8) != TP_OK || e->counters.ensure(64 * 4) != TP_OK)
    return TP_ERR_CUDA;
  double* dS = e->scratch_a.as<double>();
  double* dE = dS + 3 * (size_t)S;
  CK(cudaMemcpyAsync(dS, starts, (size_t)S * 24, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(dE, ends, (size_t)S * 24, cudaMemcpyHostToDevice, s));
  CK(cudaMemsetAsync(e->counters.p, 0, 64 * 4, s));
  int* dlen = e->scratch_c.as<int>();
  int* dexp = dlen + S;
  const int grid = std::min(e->pools.workers, S);
  k_astar<<<grid, 32, 0, s>>>(C, e->dmap, e->pools, e->counters.as<int>(), S, dS, dE, dlen, e->scratch_b.as<double>(), dexp);
  e->launches += 1;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(path_len, dlen, (size_t)S * 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(expansions, dexp, (size_t)S * 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(paths, e->scratch_b.p, pbytes, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return TP_OK;
}

// shared by init_guides and make_plan: makePlan steps 1-3 on the whole batch
static int run_plan_init(tp_engine* e, BatchSetup& bs, cudaStream_t s) {
  int rc = ensure_pools(e, bs.C);
  if (rc != TP_OK) return rc;
  CK(cudaMemsetAsync(e->counters.p, 0, 64 * 4, s));
  int* cnt = e->counters.as<int>();
  const int grid = std::min(e->pools.workers, bs.bv.B);
  {
    ProfScope ps(e, 3, s);
    k_plan_init<<<grid, 32, 0, s>>>(bs.bv, bs.C, e->dmap, e->pools, cnt + 1, e->active[0].as<int>(), cnt + 0);
  }
  e->launches += 1;
  CK(cudaGetLastError());
  return TP_OK;
}

int tp_vigo_init_guides_batch(tp_engine_t* e, const tp_vigo_params* p, int32_t B, const int32_t* offsets,
                              const double* ctrl, uint8_t* ok, int32_t* nseg, int32_t* segs, int32_t* g_count,
                              int32_t* g_cp, double* g_p, double* g_v) {
  if (!e || !ok || !nseg || !segs || !g_count || !g_cp || !g_p || !g_v) return TP_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  cudaStream_t s = e->stream;
  BatchSetup bs;
  int rc = setup_batch(e, p, B, offsets, ctrl, TP_MEM_HOST, s, bs, true);
  if (rc != TP_OK) return rc;
  rc = run_plan_init(e, bs, s);
  if (rc != TP_OK) return rc;
  const int gcap = bs.C.gcap;
  std::vector<TrajState> hst((size_t)B);
  std::vector<GuidePair> pairs((size_t)B * gcap);
  std::vector<int> head((size_t)std::max<long>(bs.total, 1));
  CK(cudaMemcpyAsync(hst.data(), bs.bv.st, (size_t)B * sizeof(TrajState), cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(pairs.data(), bs.bv.pairs, pairs.size() * sizeof(GuidePair), cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(head.data(), bs.bv.cp_head, head.size() * 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  for (int b = 0; b < B; ++b) {
    const TrajState& st = hst[b];
    ok[b] = st.status == TS_ACTIVE ? 1 : 0;
    nseg[b] = st.nseg;
    for (int i = 0; i < bs.C.max_seg; ++i) {
      segs[((size_t)b * bs.C.max_seg + i) * 2] = i < st.nseg ? st.seg[i][0] : 0;
      segs[((size_t)b * bs.C.max_seg + i) * 2 + 1] = i < st.nseg ? st.seg[i][1] : 0;
    }
    // flat list ordered by control point, append order within a control point
    int w = 0;
    for (int c = 0; c < st.N; ++c)
      for (int gi = head[st.off + c]; gi >= 0; gi = pairs[(size_t)b * gcap + gi].next) {
        const GuidePair& pr = pairs[(size_t)b * gcap + gi];
        g_cp[(size_t)b * gcap + w] = c;
        for (int a = 0; a < 3; ++a) {
          g_p[((size_t)b * gcap + w) * 3 + a] = pr.p[a];
          g_v[((size_t)b * gcap + w) * 3 + a] = pr.v[a];
        }
        ++w;
      }
    g_count[b] = w;
  }
  return TP_OK;
}

int tp_vigo_make_plan_batch(tp_engine_t* e, const tp_vigo_params* p, int32_t B, const int32_t* offsets,
                            const double* ctrl_in, double* ctrl_out, tp_vigo_result* results, int32_t n_dyn,
                            const double* dyn_pos, const double* dyn_vel, const double* dyn_size, int mem, void* stream) {
  if (!e) return TP_ERR_INVALID_ARG;
  if (B == 0) {   // an empty batch is a valid no-op (after the usual parameter / map checks)
    const int rc0 = check_params(e, p);
    return rc0;
  }
  if (!ctrl_out || !results) return TP_ERR_INVALID_ARG;
  if (n_dyn < 0 || (n_dyn > 0 && (!dyn_pos || !dyn_vel || !dyn_size))) return TP_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  cudaStream_t s = pick_stream(e, stream);
  BatchSetup bs;
  // device mode works in place on ctrl_out
  if (mem == TP_MEM_DEVICE && ctrl_in != ctrl_out) {
    std::vector<int> ho(B + 1);
    CK(cudaMemcpyAsync(ho.data(), offsets, (size_t)(B + 1) * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaMemcpyAsync(ctrl_out, ctrl_in, (size_t)ho[B] * 24, cudaMemcpyDeviceToDevice, s));
  }
  int rc = setup_batch(e, p, B, offsets, mem == TP_MEM_DEVICE ? ctrl_out : ctrl_in, mem, s, bs, true);
  if (rc != TP_OK) return rc;
  rc = upload_dyn(e, bs, n_dyn, dyn_pos, dyn_vel, dyn_size, s);
  if (rc != TP_OK) return rc;
  rc = ensure_pools(e, bs.C);
  if (rc != TP_OK) return rc;
  int mode = lbfgs_mode(p);
  {
    // Throughput-bound batches (several waves of trajectories per worker) do better with the history in tensor memory:
    // 4 workers per SM instead of 3 (76.7 k vs 72.0 k solves/s at 16 384, 73.8 k vs 70.0 k at 8 192); batches whose makespan is a few long
    // trajectories do better with the faster block of the shared-memory variant (67.7 k vs 65.2 k at 4 096).
    static const bool forced = getenv("TP_LBFGS_TMEM") != nullptr || getenv("TP_LBFGS_SMEM_HISTORY") != nullptr;
    static const int tm_batch = getenv("TP_TMEM_BATCH") ? atoi(getenv("TP_TMEM_BATCH")) : 8192;
    if (mode == 2 && !forced && B >= tm_batch) mode = 3;
  }
  if (!e->solve_attr_set) {
    CK(cudaFuncSetAttribute(k_solve<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, e->max_smem_optin));
    CK(cudaFuncSetAttribute(k_solve<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, e->max_smem_optin));
    CK(cudaFuncSetAttribute(k_solve<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, e->max_smem_optin));
    CK(cudaFuncSetAttribute(k_solve<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, e->max_smem_optin));
    CK(cudaFuncSetAttribute(k_solve<3>, cudaFuncAttributePreferredSharedMemoryCarveout, getenv("TP_CARVEOUT") ? atoi(getenv("TP_CARVEOUT")) : 70));   // 4 x 33.5 KB of shared memory: leave the rest to L1
    CK(cudaFuncSetAttribute(k_solve<0>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CK(cudaFuncSetAttribute(k_solve<1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CK(cudaFuncSetAttribute(k_solve<2>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CK(cudaFuncSetAttribute(k_solve<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, e->max_smem_optin));
    CK(cudaFuncSetAttribute(k_solve<4>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CK(cudaFuncSetAttribute(k_solve_w, cudaFuncAttributeMaxDynamicSharedMemorySize, e->max_smem_optin));
    CK(cudaFuncSetAttribute(k_solve_w, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    for (int i = 0; i < 8; ++i) {
      CK(cudaStreamCreateWithFlags(&e->class_stream[i], cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&e->ev_join[i], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&e->ev_stage, cudaEventDisableTiming));
    e->solve_attr_set = true;
  }
  // ---- issue order.  Size classes (by control-point count) so that shared memory per block — hence resident
  // blocks per SM — follows the trajectory length; one launch of persistent workers per non-empty class, on sibling
  // streams.  Inside a class the queue is ordered by the number of control points that start inside inflated
  // obstacles.  How hard a trajectory really is shows only in its first A* searches, so the workers themselves
  // re-order the batch: cheap starts are parked after makePlan steps 1-3 and resumed once the class queues are
  // drained, expensive starts run to the end first (k_solve, phases A / B).  TP_TWO_PASS=1 keeps the older
  // host-driven variant of the same idea (one round for everybody, sort by first-round work on the host, second
  // launch) for A/B measurements.
  // upper N bound of each class, largest first.  With the history in tensor memory the shared-memory footprint is
  // flat up to N = 88 (the A* scratch dominates) and grows only with the overflow elements beyond 2 per thread
  // (+12 KB at N = 104): one class serves every path the reference accepts (max_path_length 20 m => N <= 105).
  static const int class_lim_smem[4] = {TP_MAX_CTRL, 104, 64, 40};
  static const int class_lim_tm[4] = {TP_MAX_CTRL, 160, 104, 0};
  const int* class_lim = mode == 3 ? class_lim_tm : class_lim_smem;
  auto class_of = [&](int b) {
    const int n = bs.h_off[b + 1] - bs.h_off[b];
    return n > class_lim[1] ? 0 : (n > class_lim[2] ? 1 : (n > class_lim[3] ? 2 : 3));
  };
  long long* tl = nullptr;
  const char* tl_path = getenv("TP_TIMELINE");
  if (tl_path) {
    if (e->scratch_a.ensure((size_t)B * 32) != TP_OK) return TP_ERR_CUDA;
    tl = e->scratch_a.as<long long>();
  }
  std::vector<int> key((size_t)B), ids;
  // one pass: sort `ids` by (class, key desc, N desc), upload, launch per class (hardest eighth of each class first)
  auto run_pass = [&](int resume, int rounds) -> int {
    std::stable_sort(ids.begin(), ids.end(), [&](int x, int y) {
      const int cx = class_of(x), cy = class_of(y);
      if (cx != cy) return cx < cy;
      if (key[x] != key[y]) return key[x] > key[y];
      return (bs.h_off[x + 1] - bs.h_off[x]) > (bs.h_off[y + 1] - bs.h_off[y]);
    });
    const int M = (int)ids.size();
    if (M == 0) return TP_OK;
    // class ranges + work estimates
    int cb[5] = {0, 0, 0, 0, 0}, nmax[4] = {0, 0, 0, 0};
    double work[4] = {0, 0, 0, 0};
    for (int i = 0; i < M; ++i) {
      const int c = class_of(ids[i]), n = bs.h_off[ids[i] + 1] - bs.h_off[ids[i]];
      cb[c + 1] += 1;
      nmax[c] = std::max(nmax[c], n);
      work[c] += (double)n * (1.0 + 0.05 * std::max(key[ids[i]], 0));
    }
    for (int c = 0; c < 4; ++c) cb[c + 1] += cb[c];
    size_t smem[4] = {0, 0, 0, 0};
    for (int c = 0; c < 4; ++c)
      if (nmax[c] > 0) {
        smem[c] = mode == 5 ? (size_t)wsolve_layout(nmax[c]).total * 8 : (size_t)solve_layout(nmax[c], mode, p->lbfgs_m).total * 8 + 16;
        if ((int)smem[c] > e->max_smem_optin) {
          tp_set_error("a %d-control-point trajectory needs %zu B of shared memory (> %d B per block)", nmax[c], smem[c], e->max_smem_optin);
          return TP_ERR_CAPACITY;
        }
      }
    // worker mix per SM: a[c] persistent workers of class c (a worker serves its class and every class of shorter
    // trajectories); at most 4 workers (registers), shared memory permitting; minimise the estimated makespan
    // T = max_c (work of classes 0..c) / (workers of classes 0..c)
    int best[4] = {0, 0, 0, 0};
    double bestT = 1e300;
    int bestN = 0;
    // budget below the 228 KB of an SM: a mix that fits only to the last KB on paper (measured: 90 + 67 + 67 KB) leaves
    // its third launch waiting for a slot until the batch is over
    const size_t smem_sm = 222 * 1024;
    // warp form: a worker is one warp (its own block of 32 threads); registers allow up to 12 per SM
    static const int wf_max = getenv("TP_W_MAX") ? atoi(getenv("TP_W_MAX")) : 12;
    static const int team_env = getenv("TP_TEAM_MAX") ? atoi(getenv("TP_TEAM_MAX")) : 0;
    // resident blocks per SM = the kernels' __launch_bounds__ (team form: TP_TEAM_BLOCKS, tensor-memory form: 4, others: 3)
    const int maxw = mode == 5 ? wf_max : (mode == 4 ? (team_env > 0 ? team_env : TP_TEAM_BLOCKS) : (mode == 3 ? 4 : 3));
    int a[4];
    for (a[0] = 0; a[0] <= maxw; ++a[0])
      for (a[1] = 0; a[0] + a[1] <= maxw; ++a[1])
        for (a[2] = 0; a[0] + a[1] + a[2] <= maxw; ++a[2])
          for (a[3] = 0; a[0] + a[1] + a[2] + a[3] <= maxw; ++a[3]) {
            size_t used_sm = 0;
            bool ok = true;
            for (int c = 0; c < 4; ++c) {
              if (a[c] > 0 && nmax[c] == 0) ok = false;
              used_sm += (size_t)a[c] * (((smem[c] + 1023) & ~(size_t)1023) + 1024);   // 1 KB granularity + 1 KB per block
            }
            if (!ok || used_sm > smem_sm) continue;
            double T = 0, wsum = 0;
            int nsum = 0;
            for (int c = 0; c < 4; ++c) {
              wsum += work[c];
              nsum += a[c];
              if (work[c] > 0) {
                if (nsum == 0) { ok = false; break; }
                T = std::max(T, wsum / nsum);
              }
            }
            if (!ok || nsum == 0) continue;
            if (T < bestT * 0.999 || (T < bestT * 1.001 && nsum > bestN)) {
              bestT = T;
              bestN = nsum;
              for (int c = 0; c < 4; ++c) best[c] = a[c];
            }
          }
    if (bestN == 0) { tp_set_error("no feasible worker mix"); return TP_ERR_CAPACITY; }
    // Worker counts per class need not be whole numbers per SM: persistent workers pull from shared queues, so the grid
    // of class c is sized to the chip (work share of the classes it serves first, shared memory of all 148 SMs) and the
    // hardware block scheduler packs the three grids.  A mix like 92 + 67 + 46 KB leaves 17 KB of every SM idle; sized
    // to the chip the same batch keeps ~3.5 instead of 3 teams per SM.
    int grid_of[4] = {0, 0, 0, 0};
    {
      static const int frac_env = getenv("TP_MIX_FRAC") ? atoi(getenv("TP_MIX_FRAC")) : 0;   // measured: no gain while registers cap an SM at 3 teams
      double wtot = 0, unit = 0;
      for (int c = 0; c < 4; ++c) wtot += work[c];
      for (int c = 0; c < 4; ++c)
        if (work[c] > 0) unit += work[c] / wtot * (double)(((smem[c] + 1023) & ~(size_t)1023) + 1024);
      const double chip = (double)smem_sm * e->sm_count * 0.96;
      const int cap_blocks = e->sm_count * maxw;   // registers
      const double total = std::min(chip / std::max(unit, 1.0), (double)cap_blocks);
      for (int c = 0; c < 4; ++c) {
        grid_of[c] = best[c] * e->sm_count;
        if (frac_env && mode >= 4 && M > bestN * e->sm_count && wtot > 0)
          grid_of[c] = work[c] > 0 ? std::max(1, (int)(work[c] / wtot * total + 0.5)) : 0;
      }
    }
    if (e->stage_busy) CK(cudaEventSynchronize(e->ev_stage));   // the previous copy out of the staging buffer
    if (ensure_stage(e, (size_t)M * 4 + 128) != TP_OK) return TP_ERR_CUDA;
    memcpy(e->h_stage, ids.data(), (size_t)M * 4);
    int* h_cb = reinterpret_cast<int*>(static_cast<char*>(e->h_stage) + (size_t)M * 4);
    for (int c = 0; c < 5; ++c) h_cb[c] = cb[c];
    for (int c = 0; c < 4; ++c) h_cb[5 + c] = 0;
    for (int c = 0; c < 4; ++c) h_cb[9 + c] = cb[c + 1] - cb[c];
    int* d_cb = e->counters.as<int>() + 16;   // [16..20] class ranges, [21..24] cursors, [25..28] unfinished trajectories per class
    CK(cudaMemcpyAsync(e->active[0].p, e->h_stage, (size_t)M * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_cb, h_cb, 13 * 4, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(e->ev_stage, s));
    e->stage_busy = true;
    // parking (k_solve phases A / B): only when the batch outnumbers the resident workers — otherwise every trajectory
    // starts at once anyway — and only for whole solves
    ParkQueue pq;
    pq.list = nullptr; pq.tail = pq.head = pq.started = nullptr; pq.stride = M; pq.thresh = -1;
    pq.score_mode = 0; pq.b[0] = 16; pq.b[1] = 8; pq.b[2] = 2;
    pq.ring = nullptr; pq.rtail = pq.rhead = nullptr; pq.remaining = d_cb + 9; pq.slice = 0; pq.rev = 0;
    {
      static const int park_env = getenv("TP_PARK_THRESH") ? atoi(getenv("TP_PARK_THRESH")) : 1500;
      int workers_total = 0;
      for (int c = 0; c < 4; ++c) workers_total += grid_of[c];
      if (park_env >= 0 && !resume && rounds == 0x7fffffff && M > workers_total) {
        const int nq = 4 * TP_PARK_BUCKETS;
        if (e->parkq.ensure(((size_t)(nq + 4) * M + 2 * nq + 16) * 4) != TP_OK) return TP_ERR_CUDA;
        int* base = e->parkq.as<int>();
        CK(cudaMemsetAsync(base, 0, (size_t)(2 * nq + 16) * 4, s));
        CK(cudaMemsetAsync(base + 2 * nq + 16, 0xFF, (size_t)(nq + 4) * M * 4, s));
        pq.tail = base; pq.head = base + nq; pq.started = base + 2 * nq; pq.list = base + 2 * nq + 16;
        pq.rtail = pq.started + 4; pq.rhead = pq.started + 8; pq.ring = pq.list + (size_t)nq * M;
        // rounds of a trajectory's first turn (team form with k_phase_a; 0 = every trajectory runs to its end once started)
        static const int slice_env = getenv("TP_SLICE_ROUNDS") ? atoi(getenv("TP_SLICE_ROUNDS")) : 0;   // measured: 1..3 lengthen the batch by 5-10 % (profiles/r02_summary.md)
        pq.slice = slice_env;
        static const int rev_env = getenv("TP_PARK_REV") ? atoi(getenv("TP_PARK_REV")) : 0;
        pq.rev = rev_env;
        pq.thresh = park_env;
        pq.score_mode = getenv("TP_PARK_SCORE") ? atoi(getenv("TP_PARK_SCORE")) : 0;
        pq.b[0] = getenv("TP_PARK_B0") ? atoi(getenv("TP_PARK_B0")) : 16;
        pq.b[1] = getenv("TP_PARK_B1") ? atoi(getenv("TP_PARK_B1")) : 8;
        pq.b[2] = getenv("TP_PARK_B2") ? atoi(getenv("TP_PARK_B2")) : 2;
      }
    }
    CK(cudaEventRecord(e->ev_fork, s));
    // team form: makePlan steps 1-3 as their own launch of light one-warp workers (k_phase_a) beside the team blocks, which
    // then run phase B only (resume = 2).  Needs the park lists and enough A* node pools for both sets of workers.
    bool hybrid = false;
    {
      static const int phase_a_env = getenv("TP_PHASE_A") ? atoi(getenv("TP_PHASE_A")) : 7;   // light workers per SM; 0 = off
      int workers_total = 0;
      for (int c = 0; c < 4; ++c) workers_total += grid_of[c];
      const int ga = std::min(std::min(M, e->sm_count * phase_a_env), e->pools.workers - workers_total);
      if (phase_a_env > 0 && mode == 4 && pq.thresh >= 0 && ga >= e->sm_count) {
        hybrid = true;
        cudaStream_t as = e->class_stream[7];
        CK(cudaStreamWaitEvent(as, e->ev_fork, 0));
        {
          ProfScope ps(e, 3, as, ga);
          k_phase_a<<<ga, 32, 0, as>>>(bs.bv, bs.C, e->dmap, e->pools, e->active[0].as<int>(), d_cb, d_cb + 5, e->pool_flags.as<int>(), tl, pq);
        }
        CK(cudaGetLastError());
        CK(cudaEventRecord(e->ev_join[7], as));
        e->launches += 1;
      }
    }
    int used = 0;
    for (int c = 0; c < 4; ++c) {
      if (grid_of[c] == 0) continue;
      // no more workers than trajectories this class (and the ones it can steal) can feed
      const int feed = cb[4] - cb[c];
      const int grid = std::min(grid_of[c], std::max(feed, 1));
      cudaStream_t cs = e->class_stream[used];
      CK(cudaStreamWaitEvent(cs, e->ev_fork, 0));
      {
        ProfScope ps(e, 0, cs, grid);
        const int* ord = e->active[0].as<int>();
        if (mode == 5) k_solve_w<<<grid, 32, smem[c], cs>>>(bs.bv, bs.C, e->dmap, e->pools, ord, d_cb, d_cb + 5, c, nmax[c], (int)(smem[c] / 8), e->pool_flags.as<int>(), e->counters_ptr(), tl, resume, rounds, pq);
        else if (mode == 1) k_solve<1><<<grid, TP_LB_THREADS, smem[c], cs>>>(bs.bv, bs.C, e->dmap, e->pools, ord, d_cb, d_cb + 5, c, nmax[c], e->pool_flags.as<int>(), e->counters_ptr(), tl, resume, rounds, pq);
        else if (mode == 2) k_solve<2><<<grid, TP_LB_THREADS, smem[c], cs>>>(bs.bv, bs.C, e->dmap, e->pools, ord, d_cb, d_cb + 5, c, nmax[c], e->pool_flags.as<int>(), e->counters_ptr(), tl, resume, rounds, pq);
        else if (mode == 3) k_solve<3><<<grid, TP_LB_THREADS, smem[c], cs>>>(bs.bv, bs.C, e->dmap, e->pools, ord, d_cb, d_cb + 5, c, nmax[c], e->pool_flags.as<int>(), e->counters_ptr(), tl, resume, rounds, pq);
        else if (mode == 4) k_solve<4><<<grid, TP_LB_THREADS, smem[c], cs>>>(bs.bv, bs.C, e->dmap, e->pools, ord, d_cb, d_cb + 5, c, nmax[c], e->pool_flags.as<int>(), e->counters_ptr(), tl, hybrid ? 2 : resume, rounds, pq);
        else k_solve<0><<<grid, TP_LB_THREADS, smem[c], cs>>>(bs.bv, bs.C, e->dmap, e->pools, ord, d_cb, d_cb + 5, c, nmax[c], e->pool_flags.as<int>(), e->counters_ptr(), tl, resume, rounds, pq);
      }
      CK(cudaGetLastError());
      CK(cudaEventRecord(e->ev_join[used], cs));
      e->launches += 1;
      ++used;
    }
    for (int i = 0; i < used; ++i) CK(cudaStreamWaitEvent(s, e->ev_join[i], 0));
    if (hybrid) CK(cudaStreamWaitEvent(s, e->ev_join[7], 0));
    if (getenv("TP_PROF_DUMP")) fprintf(stderr, "[tp-mix] grids by class: %d %d %d %d; whole workers/SM by class: %d %d %d %d (smem %zu %zu %zu %zu B), class sizes %d %d %d %d\n",
                                        grid_of[0], grid_of[1], grid_of[2], grid_of[3], best[0], best[1], best[2], best[3], smem[0], smem[1], smem[2], smem[3], cb[1] - cb[0], cb[2] - cb[1],
                                        cb[3] - cb[2], cb[4] - cb[3]);
    return TP_OK;
  };
  const bool two_pass = getenv("TP_TWO_PASS") != nullptr && B >= 256;
  // pass 1 (or the only pass for small batches)
  k_count_colliding<<<(B + 3) / 4, 128, 0, s>>>(bs.bv, e->dmap, e->active[1].as<int>());
  e->launches += 1;
  CK(cudaMemcpyAsync(key.data(), e->active[1].p, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  ids.resize((size_t)B);
  for (int b = 0; b < B; ++b) ids[b] = b;
  rc = run_pass(0, two_pass ? 1 : 0x7fffffff);
  if (rc != TP_OK) return rc;
  if (two_pass) {
    k_difficulty<<<(B + 127) / 128, 128, 0, s>>>(bs.bv, e->active[1].as<int>());
    e->launches += 1;
    CK(cudaMemcpyAsync(key.data(), e->active[1].p, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    ids.clear();
    for (int b = 0; b < B; ++b)
      if (key[b] >= 0) ids.push_back(b);
    rc = run_pass(1, 0x7fffffff);
    if (rc != TP_OK) return rc;
  }
  if (tl) {
    std::vector<long long> h((size_t)B * 4);
    CK(cudaMemcpyAsync(h.data(), tl, (size_t)B * 32, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (FILE* f = fopen(tl_path, "wb")) {
      fwrite(h.data(), 8, h.size(), f);
      fclose(f);
    }
  }
  if (e->results.ensure((size_t)B * sizeof(tp_vigo_result)) != TP_OK) return TP_ERR_CUDA;
  tp_vigo_result* dres = mem == TP_MEM_DEVICE ? results : e->results.as<tp_vigo_result>();
  k_collect_results<<<(B + 127) / 128, 128, 0, s>>>(bs.bv, dres);
  e->launches += 1;
  CK(cudaGetLastError());
  if (mem == TP_MEM_HOST) {
    CK(cudaMemcpyAsync(ctrl_out, bs.bv.ctrl, (size_t)bs.total * 24, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(results, dres, (size_t)B * sizeof(tp_vigo_result), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  return TP_OK;
}

// ---- one host, several GPUs: independent trajectories, no exchange step.  One host thread per engine
// pulls contiguous chunks of the batch from a shared cursor (dynamic load balance: the engines of a box rarely finish
// equal shares at the same time — per-shard tails differ) and runs them through tp_vigo_make_plan_batch with host
// buffers; every engine holds its own replica of the map.  Results land at the trajectories' own positions, so the
// output is independent of which engine solved what (each trajectory's result does not depend on its batch).
int tp_vigo_make_plan_batch_multi(tp_engine_t* const* engines, int32_t n_engines, const tp_vigo_params* p, int32_t B,
                                  const int32_t* offsets, const double* ctrl_in, double* ctrl_out, tp_vigo_result* results,
                                  int32_t n_dyn, const double* dyn_pos, const double* dyn_vel, const double* dyn_size,
                                  int32_t chunk, int32_t* engine_of) {
  if (!engines || n_engines < 1 || !p) { tp_set_error("tp_vigo_make_plan_batch_multi: no engines / params"); return TP_ERR_INVALID_ARG; }
  for (int i = 0; i < n_engines; ++i)
    if (!engines[i]) { tp_set_error("tp_vigo_make_plan_batch_multi: engine %d is null", i); return TP_ERR_INVALID_ARG; }
  if (B == 0) return check_params(engines[0], p);
  if (B < 0 || !offsets || !ctrl_in || !ctrl_out || !results) return TP_ERR_INVALID_ARG;
  if (offsets[0] != 0) { tp_set_error("offsets[0] must be 0"); return TP_ERR_INVALID_ARG; }
  if (chunk <= 0) {
    // about four chunks per engine (balance) but not below 1,024 trajectories (a chunk's makespan ends in its tail)
    chunk = std::max(1024, (B + 4 * n_engines - 1) / (4 * n_engines));
  }
  const int n_chunks = (B + chunk - 1) / chunk;
  std::atomic<int> cursor(0), first_err(TP_OK);
  std::vector<std::string> err_msg((size_t)n_engines);
  auto worker = [&](int ei) {
    std::vector<int32_t> loc;
    for (;;) {
      const int c = cursor.fetch_add(1);
      if (c >= n_chunks || first_err.load() != TP_OK) break;
      const int b0 = c * chunk, b1 = std::min(B, b0 + chunk);
      loc.resize((size_t)(b1 - b0) + 1);
      for (int b = b0; b <= b1; ++b) loc[(size_t)(b - b0)] = offsets[b] - offsets[b0];
      const int rc = tp_vigo_make_plan_batch(engines[ei], p, b1 - b0, loc.data(), ctrl_in + 3 * (size_t)offsets[b0],
                                             ctrl_out + 3 * (size_t)offsets[b0], results + b0, n_dyn, dyn_pos, dyn_vel, dyn_size,
                                             TP_MEM_HOST, nullptr);
      if (rc != TP_OK) {
        int expect = TP_OK;
        if (first_err.compare_exchange_strong(expect, rc)) err_msg[(size_t)ei] = tp_last_error();
        break;
      }
      if (engine_of)
        for (int b = b0; b < b1; ++b) engine_of[b] = ei;
    }
  };
  std::vector<std::thread> th;
  for (int i = 1; i < n_engines; ++i) th.emplace_back(worker, i);
  worker(0);
  for (auto& t : th) t.join();
  if (first_err.load() != TP_OK) {
    for (const auto& m : err_msg)
      if (!m.empty()) tp_set_error("%s", m.c_str());
    return first_err.load();
  }
  return TP_OK;
}

// ---- front end on the device: (start, goal) pairs -> ragged initial control points
int64_t tp_vigo_frontend_batch_device(tp_engine_t* e, const tp_vigo_params* p, int32_t B, const double* starts, const double* goals,
                                      int32_t* offsets_out, double* ctrl_out, int64_t ctrl_cap, uint8_t* valid, int mem,
                                      void* stream) {
  if (!e || !p || B < 0 || !starts || !goals || !offsets_out || !ctrl_out) return TP_ERR_INVALID_ARG;
  int rc = check_params(e, p);
  if (rc != TP_OK) return rc;
  CK(cudaSetDevice(e->device));
  cudaStream_t s = pick_stream(e, stream);
  if (B == 0) {
    const int32_t zero = 0;
    if (mem == TP_MEM_HOST) offsets_out[0] = 0;
    else CK(cudaMemcpyAsync(offsets_out, &zero, 4, cudaMemcpyHostToDevice, s));
    return 0;
  }
  const size_t stride = (size_t)(FE_KMAX + 2) * 3;
  // scratch_a: [starts | goals] (host mode), scratch_b: per-problem control points, scratch_c: counts + offsets + valid
  if (e->scratch_b.ensure((size_t)B * stride * 8) != TP_OK || e->scratch_c.ensure((size_t)B * 9 + 64) != TP_OK) return TP_ERR_CUDA;
  const double *ds = starts, *dg = goals;
  if (mem == TP_MEM_HOST) {
    if (e->scratch_a.ensure((size_t)B * 48) != TP_OK) return TP_ERR_CUDA;
    CK(cudaMemcpyAsync(e->scratch_a.p, starts, (size_t)B * 24, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(e->scratch_a.as<double>() + 3 * (size_t)B, goals, (size_t)B * 24, cudaMemcpyHostToDevice, s));
    ds = e->scratch_a.as<double>();
    dg = ds + 3 * (size_t)B;
  }
  int* d_count = e->scratch_c.as<int>();
  int* d_off = d_count + B;
  unsigned char* d_valid = reinterpret_cast<unsigned char*>(d_off + B + 1);
  if (!e->fe_attr_set) {   // per engine = per device (function attributes are per context)
    CK(cudaFuncSetAttribute(k_frontend, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FeSmem)));
    e->fe_attr_set = true;
  }
  const int grid = std::min(B, e->sm_count * 5);
  {
    ProfScope ps(e, 3, s, grid);
    k_frontend<<<grid, 32, sizeof(FeSmem), s>>>(e->dmap, *p, B, ds, dg, e->scratch_b.as<double>(), d_count);
  }
  int* off_dst = mem == TP_MEM_DEVICE ? offsets_out : d_off;
  k_fe_scan<<<1, 1024, 0, s>>>(d_count, B, off_dst);
  e->launches += 2;
  CK(cudaGetLastError());
  int32_t total = 0;
  CK(cudaMemcpyAsync(&total, off_dst + B, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if ((int64_t)total > ctrl_cap) {
    tp_set_error("tp_vigo_frontend_batch_device: ctrl_cap %lld < %d control points", (long long)ctrl_cap, total);
    return TP_ERR_CAPACITY;
  }
  double* ctrl_dst = ctrl_out;
  if (mem == TP_MEM_HOST) {
    if (e->ctrl.ensure((size_t)std::max(total, 1) * 24) != TP_OK) return TP_ERR_CUDA;
    ctrl_dst = e->ctrl.as<double>();
  }
  unsigned char* valid_dst = mem == TP_MEM_DEVICE ? valid : d_valid;
  k_fe_gather<<<B, 64, 0, s>>>(e->scratch_b.as<double>(), off_dst, B, ctrl_dst, (long)ctrl_cap, valid_dst);
  e->launches += 1;
  CK(cudaGetLastError());
  if (mem == TP_MEM_HOST) {
    CK(cudaMemcpyAsync(offsets_out, d_off, (size_t)(B + 1) * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ctrl_out, ctrl_dst, (size_t)total * 24, cudaMemcpyDeviceToHost, s));
    if (valid) CK(cudaMemcpyAsync(valid, d_valid, (size_t)B, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  return total;
}

#ifdef TP_WF_TIMING
// development builds only (build_timing.sh with EXTRA=-DTP_WF_TIMING): whole-batch warp-form phase cycle totals, reset on read
extern "C" int tp_debug_wf_phase_get(unsigned long long* out) {
  unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cudaMemcpyFromSymbol(out, g_wf_phase, sizeof(z)) != cudaSuccess) return -1;
  if (cudaMemcpyToSymbol(g_wf_phase, z, sizeof(z)) != cudaSuccess) return -1;
  return 0;
}
#endif
#ifdef TP_LBFGS_TIMING
// development builds only (build_timing.sh): whole-batch L-BFGS phase cycle totals, reset on read
extern "C" int tp_debug_phase_get(unsigned long long* out) {
  unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cudaMemcpyFromSymbol(out, g_tp_phase, sizeof(z)) != cudaSuccess) return -1;
  cudaMemcpyToSymbol(g_tp_phase, z, sizeof(z));
  return 0;
}
#endif

// ---- measurement
__global__ void k_fp64_fma(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9 + 1.0, a1 = a0 + 0.1, a2 = a0 + 0.2, a3 = a0 + 0.3, a4 = a0 + 0.4, a5 = a0 + 0.5,
         a6 = a0 + 0.6, a7 = a0 + 0.7;
  const double m = 0.9999999, c = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
    a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}
__global__ void k_gather32(const uint4* __restrict__ buf, long n_sectors, long per_thread, unsigned long long* sink) {
  // every thread reads `per_thread` pseudo-random 32 B sectors (two 16 B loads of one sector)
  unsigned long long x = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
  unsigned long long acc = 0;
  for (long i = 0; i < per_thread; ++i) {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    const long sidx = (long)(x % (unsigned long long)n_sectors);
    const uint4 a = __ldg(&buf[2 * sidx]);
    const uint4 b = __ldg(&buf[2 * sidx + 1]);
    acc += a.x + b.w;
  }
  if (acc == 0xFFFFFFFFFFFFFFFFull) *sink = acc;
}

int tp_engine_profile_enable(tp_engine_t* e, int on) {
  if (!e) return TP_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  if (on) {
    if (e->dev_counters.ensure(8 * 8) != TP_OK) return TP_ERR_CUDA;
    CK(cudaMemset(e->dev_counters.p, 0, 64));
  }
  e->profile = on != 0;
  return TP_OK;
}
int tp_engine_profile_get(tp_engine_t* e, tp_profile* out) {
  if (!e || !out) return TP_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  CK(cudaDeviceSynchronize());
  memset(out, 0, sizeof(*out));
  for (auto& pe : e->prof_entries) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, pe.a, pe.b) == cudaSuccess && pe.kind >= 0 && pe.kind < TP_PROF_KINDS) {
      out->ms[pe.kind] += ms;
      out->launches[pe.kind] += 1;
      if (getenv("TP_PROF_DUMP")) fprintf(stderr, "[tp-prof] kind %d n %d ms %.4f\n", pe.kind, pe.n, ms);
    }
    e->ev_pool.push_back(pe.a);
    e->ev_pool.push_back(pe.b);
  }
  e->prof_entries.clear();
  if (e->dev_counters.p) {
    double h[8];
    CK(cudaMemcpy(h, e->dev_counters.p, 64, cudaMemcpyDeviceToHost));
    out->lbfgs_flops = h[0];
    out->lbfgs_iters = h[1];
    out->lbfgs_evals = h[2];
    out->check_samples = h[3];
    CK(cudaMemset(e->dev_counters.p, 0, 64));
  }
  out->query_points = e->query_points;
  e->query_points = 0;
  return TP_OK;
}
int tp_microbench_fp64(tp_engine_t* e, double* tflops) {
  if (!e || !tflops) return TP_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  const int blocks = e->sm_count * 8, threads = 256, iters = 1 << 15;
  if (e->scratch_a.ensure((size_t)blocks * threads * 8) != TP_OK) return TP_ERR_CUDA;
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(a, e->stream));
    k_fp64_fma<<<blocks, threads, 0, e->stream>>>(e->scratch_a.as<double>(), iters);
    CK(cudaEventRecord(b, e->stream));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    const double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
    if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  e->launches += 5;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  *tflops = best;
  return TP_OK;
}
int tp_microbench_gather(tp_engine_t* e, int64_t bytes, double* gbs) {
  if (!e || !gbs || bytes < 4096) return TP_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  const long n_sectors = bytes / 32;
  if (e->scratch_b.ensure((size_t)n_sectors * 32) != TP_OK || e->scratch_c.ensure(64) != TP_OK) return TP_ERR_CUDA;
  CK(cudaMemsetAsync(e->scratch_b.p, 1, (size_t)n_sectors * 32, e->stream));
  const int blocks = e->sm_count * 16, threads = 256;
  const long per_thread = 256;
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(a, e->stream));
    k_gather32<<<blocks, threads, 0, e->stream>>>(e->scratch_b.as<uint4>(), n_sectors, per_thread,
                                                   e->scratch_c.as<unsigned long long>());
    CK(cudaEventRecord(b, e->stream));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    const double by = 32.0 * (double)per_thread * blocks * threads;
    if (rep > 0) best = std::max(best, by / (ms * 1e-3) / 1e9);
  }
  e->launches += 5;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  *gbs = best;
  return TP_OK;
}


// ------------------------------------------------------------------------------------ min-snap / polyTraj (secondary path)
static PolyMap make_polymap(const tp_engine* e) {
  PolyMap pm;
  pm.occ = (const uint32_t*)e->map_occ.p;
  pm.known = (const uint32_t*)e->map_known.p;
  pm.res = e->dmap.res;
  for (int a = 0; a < 3; ++a) { pm.mn[a] = e->dmap.mn[a]; pm.dim[a] = e->dmap.dim[a]; pm.bbmin[a] = e->known_bbmin[a]; pm.bbmax[a] = e->known_bbmax[a]; }
  pm.wz = e->dmap.wz;
  return pm;
}
void tp_poly_default_params(tp_poly_params* p) {
  // cfg/planner_interactive.yaml + polyTrajOctomap.cpp:14-108 (mode = true: the re-insert loop)
  memset(p, 0, sizeof(*p));
  p->desired_vel = 1.0; p->delT = 0.1;
  p->box[0] = 0.4; p->box[1] = 0.4; p->box[2] = 0.2;
  p->map_res = 0.2; p->cont = 4; p->max_iter = 100; p->max_waypoints = 64;
}
static int poly_check_params(const tp_engine* e, const tp_poly_params* p, bool need_map) {
  if (!e || !p) return TP_ERR_INVALID_ARG;
  if (need_map && !e->has_map) { tp_set_error("engine has no map"); return TP_ERR_NO_MAP; }
  if (!(p->desired_vel > 0) || !(p->delT > 0) || !(p->map_res > 0) || p->cont < 2 || p->cont > 4) {
    tp_set_error("invalid polyTraj parameters");
    return TP_ERR_INVALID_ARG;
  }
  return TP_OK;
}
// device-side solve on device-resident waypoints; outputs device pointers
static int poly_solve_device(tp_engine* e, const tp_poly_params* p, int B, int max_k, const int* d_off, const double* d_wp,
                             const double* d_bc, double* d_coef, double* d_times, int* d_status, cudaStream_t s) {
  if (max_k > PL_MAX_SEG) { tp_set_error("a path has %d segments (max %d)", max_k, PL_MAX_SEG); return TP_ERR_CAPACITY; }
  // one WARP per problem (tp_band.cuh): four workers per block, up to 16 blocks per SM
  const size_t stride = poly_scratch_doubles(std::max(max_k, 1));
  const int grid = std::min((B + 3) / 4, e->sm_count * 16);
  if (e->poly_scratch.ensure((size_t)grid * 4 * stride * 8) != TP_OK || e->counters.ensure(64 * 4) != TP_OK)
    return TP_ERR_CUDA;
  CK(cudaMemsetAsync(e->counters.p, 0, 64 * 4, s));
  PolySolveArgs A;
  A.B = B; A.wp_off = d_off; A.wp = d_wp; A.bc = d_bc; A.desired_vel = p->desired_vel; A.cont = p->cont;
  A.coef = d_coef; A.times = d_times; A.status = d_status; A.scratch = e->poly_scratch.as<double>(); A.stride = stride;
  A.queue = e->counters.as<int>();
  {
    ProfScope ps(e, 6, s, B);
    k_minsnap_solve<<<grid, PL_THREADS, 0, s>>>(A);
  }
  e->launches += 1;
  CK(cudaGetLastError());
  return TP_OK;
}
static int poly_ensure_tacc(tp_engine* e, double delT, cudaStream_t s) {
  if (e->poly_tacc_dt == delT && e->poly_tacc_n > 0) return TP_OK;
  std::vector<double> t;
  double x = 0.0;
  for (int i = 0; i < 65536; ++i) { t.push_back(x); x += delT; }
  if (e->poly_tacc.ensure(t.size() * 8) != TP_OK) return TP_ERR_CUDA;
  CK(cudaMemcpyAsync(e->poly_tacc.p, t.data(), t.size() * 8, cudaMemcpyHostToDevice, s));
  CK(cudaStreamSynchronize(s));
  e->poly_tacc_n = (int)t.size();
  e->poly_tacc_dt = delT;
  return TP_OK;
}

int tp_minsnap_solve_batch(tp_engine_t* e, const tp_poly_params* p, int32_t B, const int32_t* wp_offsets, const double* waypoints,
                           const double* bc, double* coef, double* times, int32_t* status) {
  int rc = poly_check_params(e, p, false);
  if (rc != TP_OK) return rc;
  if (B <= 0 || !wp_offsets || !waypoints || !coef || !times || !status) return TP_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  cudaStream_t s = e->stream;
  const int total = wp_offsets[B];
  int max_k = 0;
  for (int b = 0; b < B; ++b) max_k = std::max(max_k, wp_offsets[b + 1] - wp_offsets[b] - 1);
  const size_t ncoef = (size_t)24 * std::max(total - B, 1);
  if (e->off.ensure((size_t)(B + 1) * 4) != TP_OK || e->ctrl.ensure((size_t)std::max(total, 1) * 24) != TP_OK ||
      e->scratch_a.ensure(ncoef * 8) != TP_OK || e->scratch_b.ensure((size_t)std::max(total, 1) * 8) != TP_OK ||
      e->scratch_c.ensure((size_t)B * 4 + (bc ? (size_t)B * 96 : 0) + 64) != TP_OK)
    return TP_ERR_CUDA;
  CK(cudaMemcpyAsync(e->off.p, wp_offsets, (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(e->ctrl.p, waypoints, (size_t)total * 24, cudaMemcpyHostToDevice, s));
  int* d_status = e->scratch_c.as<int>();
  double* d_bc = nullptr;
  if (bc) {
    d_bc = reinterpret_cast<double*>(e->scratch_c.as<char>() + (((size_t)B * 4 + 63) & ~(size_t)63));
    CK(cudaMemcpyAsync(d_bc, bc, (size_t)B * 96, cudaMemcpyHostToDevice, s));
  }
  rc = poly_solve_device(e, p, B, max_k, e->off.as<int>(), e->ctrl.as<double>(), d_bc, e->scratch_a.as<double>(), e->scratch_b.as<double>(), d_status, s);
  if (rc != TP_OK) return rc;
  CK(cudaMemcpyAsync(coef, e->scratch_a.p, (size_t)24 * (total - B) * 8, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(times, e->scratch_b.p, (size_t)total * 8, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(status, d_status, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return TP_OK;
}

int tp_poly_check_batch(tp_engine_t* e, const tp_poly_params* p, int32_t B, const int32_t* wp_offsets, const double* waypoints,
                        const double* coef, const double* times, uint8_t* valid, uint8_t* seg_hit, int32_t* n_samples,
                        double* samples, uint8_t* sample_hit, int32_t samp_cap) {
  int rc = poly_check_params(e, p, true);
  if (rc != TP_OK) return rc;
  if (B <= 0 || !wp_offsets || !waypoints || !coef || !times || !valid || !seg_hit || !n_samples) return TP_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  cudaStream_t s = e->stream;
  rc = poly_ensure_tacc(e, p->delT, s);
  if (rc != TP_OK) return rc;
  const int total = wp_offsets[B], nseg = total - B;
  const size_t ncoef = (size_t)24 * std::max(nseg, 1);
  const size_t samp_bytes = samples ? (size_t)B * samp_cap * 25 : 0;
  if (e->off.ensure((size_t)(B + 1) * 4) != TP_OK || e->ctrl.ensure((size_t)std::max(total, 1) * 24) != TP_OK ||
      e->scratch_a.ensure(ncoef * 8) != TP_OK || e->scratch_b.ensure((size_t)std::max(total, 1) * 8) != TP_OK ||
      e->scratch_c.ensure((size_t)B * 8 + (size_t)std::max(nseg, 1) + 64) != TP_OK || (samples && e->results.ensure(samp_bytes + 64) != TP_OK))
    return TP_ERR_CUDA;
  CK(cudaMemcpyAsync(e->off.p, wp_offsets, (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(e->ctrl.p, waypoints, (size_t)total * 24, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(e->scratch_a.p, coef, (size_t)24 * nseg * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(e->scratch_b.p, times, (size_t)total * 8, cudaMemcpyHostToDevice, s));
  PolyCheckArgs A;
  A.B = B; A.wp_off = e->off.as<int>(); A.wp = e->ctrl.as<double>(); A.coef = e->scratch_a.as<double>(); A.times = e->scratch_b.as<double>();
  A.t_acc = e->poly_tacc.as<double>(); A.n_t_acc = e->poly_tacc_n;
  for (int a = 0; a < 3; ++a) A.box[a] = p->box[a];
  A.map_res = p->map_res;
  A.n_samples = e->scratch_c.as<int>();
  A.valid = reinterpret_cast<uint8_t*>(A.n_samples + B);
  A.seg_hit = A.valid + B;
  A.samples = samples ? e->results.as<double>() : nullptr;
  A.sample_hit = samples ? reinterpret_cast<uint8_t*>(e->results.as<double>() + (size_t)B * samp_cap * 3) : nullptr;
  A.samp_cap = samp_cap;
  if (samples) CK(cudaMemsetAsync(e->results.p, 0, samp_bytes, s));
  {
    ProfScope ps(e, 7, s, B);
    k_poly_check<<<B, PL_THREADS, 0, s>>>(A, make_polymap(e));
  }
  e->launches += 1;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(valid, A.valid, (size_t)B, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(seg_hit, A.seg_hit, (size_t)nseg, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(n_samples, A.n_samples, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
  if (samples) {
    CK(cudaMemcpyAsync(samples, A.samples, (size_t)B * samp_cap * 24, cudaMemcpyDeviceToHost, s));
    if (sample_hit) CK(cudaMemcpyAsync(sample_hit, A.sample_hit, (size_t)B * samp_cap, cudaMemcpyDeviceToHost, s));
  }
  CK(cudaStreamSynchronize(s));
  return TP_OK;
}

int tp_poly_box_collision(tp_engine_t* e, const tp_poly_params* p, int64_t n, const double* xyz, uint8_t* hit) {
  int rc = poly_check_params(e, p, true);
  if (rc != TP_OK) return rc;
  if (n < 0 || (n > 0 && (!xyz || !hit))) return TP_ERR_INVALID_ARG;
  if (n == 0) return TP_OK;
  CK(cudaSetDevice(e->device));
  cudaStream_t s = e->stream;
  if (e->scratch_a.ensure((size_t)n * 24) != TP_OK || e->scratch_c.ensure((size_t)n) != TP_OK) return TP_ERR_CUDA;
  CK(cudaMemcpyAsync(e->scratch_a.p, xyz, (size_t)n * 24, cudaMemcpyHostToDevice, s));
  const long blocks = std::min<long>((n + 255) / 256, (long)e->sm_count * 8);
  k_poly_box_points<<<(int)blocks, 256, 0, s>>>(make_polymap(e), (long)n, e->scratch_a.as<double>(), p->box[0], p->box[1], p->box[2],
                                                 p->map_res, e->scratch_c.as<uint8_t>());
  e->launches += 1;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(hit, e->scratch_c.p, (size_t)n, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return TP_OK;
}

// polyTrajOctomap::makePlanAddingWaypoint (polyTrajOctomap.cpp:259-321) for B paths: solve -> sample -> box collision check ->
// insert a midpoint into every colliding segment (highest index first, :178-186) -> re-solve, until collision free,
// max_iter iterations (the reference's countIter > maxIter_ exit) or the waypoint cap.  The waypoint lists grow on the
// host between device passes; only the still-invalid paths are re-submitted.
int tp_polytraj_make_plan_batch(tp_engine_t* e, const tp_poly_params* p, int32_t B, const int32_t* wp_offsets, const double* waypoints,
                                int32_t* wp_offsets_out, double* waypoints_out, int64_t wp_cap, double* coef_out, double* times_out,
                                uint8_t* valid_out, int32_t* iters_out) {
  return tp_polytraj_make_plan_batch_bc(e, p, B, wp_offsets, waypoints, nullptr, wp_offsets_out, waypoints_out, wp_cap, coef_out,
                                        times_out, valid_out, iters_out);
}

int tp_polytraj_make_plan_batch_bc(tp_engine_t* e, const tp_poly_params* p, int32_t B, const int32_t* wp_offsets, const double* waypoints,
                                   const double* bc, int32_t* wp_offsets_out, double* waypoints_out, int64_t wp_cap, double* coef_out,
                                   double* times_out, uint8_t* valid_out, int32_t* iters_out) {
  int rc = poly_check_params(e, p, true);
  if (rc != TP_OK) return rc;
  if (B <= 0 || !wp_offsets || !waypoints || !wp_offsets_out || !waypoints_out || !coef_out || !times_out || !valid_out || !iters_out)
    return TP_ERR_INVALID_ARG;
  const int cap = std::min(p->max_waypoints > 1 ? p->max_waypoints : 64, PL_MAX_SEG + 1);
  for (int b = 0; b < B; ++b)
    if (wp_offsets[b + 1] - wp_offsets[b] > cap) { tp_set_error("path %d has more than %d waypoints", b, cap); return TP_ERR_CAPACITY; }
  CK(cudaSetDevice(e->device));
  cudaStream_t s = e->stream;
  rc = poly_ensure_tacc(e, p->delT, s);
  if (rc != TP_OK) return rc;
  const int total = wp_offsets[B];
  const size_t pstride = poly_scratch_doubles(cap - 1);
  const int grid = std::min(B, e->sm_count * 8);
  const size_t st_wp = (size_t)B * cap * 3, st_coef = (size_t)B * 24 * (cap - 1), st_t = (size_t)B * cap;
  // staging rows | packed outputs (sized for wp_cap)
  const size_t packed = (size_t)wp_cap * 3 + (size_t)wp_cap * 24 + (size_t)wp_cap;
  if (e->off.ensure((size_t)(B + 1) * 4) != TP_OK || e->ctrl.ensure((size_t)std::max(total, 1) * 24) != TP_OK ||
      e->poly_scratch.ensure((size_t)grid * pstride * 8) != TP_OK ||
      e->scratch_a.ensure((st_wp + st_coef + st_t) * 8) != TP_OK || e->scratch_b.ensure(packed * 8) != TP_OK ||
      e->scratch_c.ensure((size_t)B * 16 + 64) != TP_OK || e->counters.ensure(64 * 4) != TP_OK || (bc && e->dyn.ensure((size_t)B * 96) != TP_OK))
    return TP_ERR_CUDA;
  CK(cudaMemcpyAsync(e->off.p, wp_offsets, (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(e->ctrl.p, waypoints, (size_t)total * 24, cudaMemcpyHostToDevice, s));
  if (bc) CK(cudaMemcpyAsync(e->dyn.p, bc, (size_t)B * 96, cudaMemcpyHostToDevice, s));
  CK(cudaMemsetAsync(e->counters.p, 0, 64 * 4, s));
  PolyLoopArgs A;
  A.B = B; A.wp_off = e->off.as<int>(); A.wp = e->ctrl.as<double>(); A.bc = bc ? e->dyn.as<double>() : nullptr;
  A.desired_vel = p->desired_vel; A.cont = p->cont; A.max_iter = p->max_iter; A.cap = cap;
  A.t_acc = e->poly_tacc.as<double>(); A.n_t_acc = e->poly_tacc_n;
  for (int a = 0; a < 3; ++a) A.box[a] = p->box[a];
  A.map_res = p->map_res;
  A.wp_st = e->scratch_a.as<double>(); A.coef_st = A.wp_st + st_wp; A.times_st = A.coef_st + st_coef;
  A.n_wp = e->scratch_c.as<int>(); A.iters = A.n_wp + B;
  int* d_off_out = A.iters + B;            // [B + 1]
  A.valid = reinterpret_cast<uint8_t*>(d_off_out + B + 1);
  A.scratch = e->poly_scratch.as<double>(); A.stride = pstride; A.queue = e->counters.as<int>();
  {
    ProfScope ps(e, 6, s, B);
    k_polytraj_loop<<<grid, PL_THREADS, 0, s>>>(A, make_polymap(e));
  }
  k_fe_scan<<<1, 1024, 0, s>>>(A.n_wp, B, d_off_out);
  double* d_wp_out = e->scratch_b.as<double>();
  double* d_coef_out = d_wp_out + (size_t)wp_cap * 3;
  double* d_times_out = d_coef_out + (size_t)wp_cap * 24;
  k_polytraj_pack<<<B, 64, 0, s>>>(A, d_off_out, d_wp_out, d_coef_out, d_times_out, (long)wp_cap);
  e->launches += 3;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(wp_offsets_out, d_off_out, (size_t)(B + 1) * 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(valid_out, A.valid, (size_t)B, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(iters_out, A.iters, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  const int64_t tot = wp_offsets_out[B];
  if (tot > wp_cap) { tp_set_error("tp_polytraj_make_plan_batch: wp_cap too small (%lld waypoints)", (long long)tot); return TP_ERR_CAPACITY; }
  CK(cudaMemcpyAsync(waypoints_out, d_wp_out, (size_t)tot * 24, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(times_out, d_times_out, (size_t)tot * 8, cudaMemcpyDeviceToHost, s));
  if (tot > B) CK(cudaMemcpyAsync(coef_out, d_coef_out, (size_t)(tot - B) * 24 * 8, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return TP_OK;
}

}  // extern "C"

// corridor-constrained min-snap: one QP solve with caller-supplied radii (solve_only) or the whole
// makePlanCorridorConstraint loop
static int corridor_run(tp_engine_t* e, const tp_poly_params* p, int32_t B, const int32_t* wp_offsets, const double* waypoints,
                        const double* bc, const double* r_in, int solve_only, int occmap, int no_corridor, double init_r, double fs, double corridor_res,
                        double* coef_out, double* times_out, uint8_t* valid_out, int32_t* iters_out, double* r_out, int32_t* status_out) {
  int rc = poly_check_params(e, p, !solve_only);
  if (rc != TP_OK) return rc;
  if (B <= 0 || !wp_offsets || !waypoints || !coef_out || !times_out) return TP_ERR_INVALID_ARG;
  if ((!no_corridor && !(corridor_res > 0)) || (!solve_only && (!(init_r > 0) || !(fs > 0) || !(fs < 1)))) {
    tp_set_error("corridor: need corridor_res > 0, initial radius > 0 and a shrinking factor in (0, 1)");
    return TP_ERR_INVALID_ARG;
  }
  if (wp_offsets[0] != 0) { tp_set_error("corridor: wp_offsets[0] must be 0"); return TP_ERR_INVALID_ARG; }
  int kmax = 1;
  for (int b = 0; b < B; ++b) {
    const int nw = wp_offsets[b + 1] - wp_offsets[b];
    if (nw < 1 || nw > PL_MAX_SEG + 1) { tp_set_error("path %d has %d waypoints (1..%d supported)", b, nw, PL_MAX_SEG + 1); return TP_ERR_CAPACITY; }
    kmax = std::max(kmax, nw - 1);
  }
  CK(cudaSetDevice(e->device));
  cudaStream_t s = e->stream;
  if (!solve_only) {
    rc = poly_ensure_tacc(e, p->delT, s);
    if (rc != TP_OK) return rc;
  }
  const int total = wp_offsets[B], nseg = total - B;
  const size_t stride = corridor_scratch_doubles(kmax);
  const int grid = std::min(B, e->sm_count * 4);
  // scratch_b: coef | times | radii in | radii out      scratch_c: iters | status | valid
  const size_t nb = (size_t)24 * std::max(nseg, 1) + (size_t)total + 2 * (size_t)std::max(nseg, 1);
  if (e->off.ensure((size_t)(B + 1) * 4) != TP_OK || e->ctrl.ensure((size_t)total * 24) != TP_OK ||
      e->poly_scratch.ensure((size_t)grid * stride * 8) != TP_OK || e->scratch_b.ensure(nb * 8) != TP_OK ||
      e->scratch_c.ensure((size_t)B * 17 + 64) != TP_OK || e->counters.ensure(64 * 4) != TP_OK || (bc && e->dyn.ensure((size_t)B * 96) != TP_OK))
    return TP_ERR_CUDA;
  CK(cudaMemcpyAsync(e->off.p, wp_offsets, (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(e->ctrl.p, waypoints, (size_t)total * 24, cudaMemcpyHostToDevice, s));
  if (bc) CK(cudaMemcpyAsync(e->dyn.p, bc, (size_t)B * 96, cudaMemcpyHostToDevice, s));
  CK(cudaMemsetAsync(e->counters.p, 0, 64 * 4, s));
  CorridorArgs A;
  A.B = B; A.wp_off = e->off.as<int>(); A.wp = e->ctrl.as<double>(); A.bc = bc ? e->dyn.as<double>() : nullptr;
  A.desired_vel = p->desired_vel; A.cont = p->cont; A.max_iter = p->max_iter;
  A.init_r = init_r; A.fs = fs; A.corridor_res = corridor_res; A.solve_only = solve_only; A.occmap = occmap; A.no_corridor = no_corridor;
  A.t_acc = e->poly_tacc.as<double>(); A.n_t_acc = e->poly_tacc_n;
  for (int a = 0; a < 3; ++a) A.box[a] = p->box[a];
  A.map_res = p->map_res;
  A.coef = e->scratch_b.as<double>(); A.times = A.coef + (size_t)24 * std::max(nseg, 1);
  double* d_rin = A.times + total;
  A.r_out = d_rin + std::max(nseg, 1);
  A.r_in = nullptr;
  if (r_in && nseg > 0) {
    CK(cudaMemcpyAsync(d_rin, r_in, (size_t)nseg * 8, cudaMemcpyHostToDevice, s));
    A.r_in = d_rin;
  }
  A.iters = e->scratch_c.as<int>(); A.status = A.iters + B; A.valid = reinterpret_cast<uint8_t*>(A.status + 3 * B);
  A.scratch = e->poly_scratch.as<double>(); A.stride = stride; A.kmax = kmax; A.queue = e->counters.as<int>();
  CK(cudaMemsetAsync(A.coef, 0, (size_t)24 * std::max(nseg, 1) * 8, s));
  {
    ProfScope ps(e, 6, s, B);
    k_corridor_loop<<<grid, PL_THREADS, 0, s>>>(A, solve_only && !e->has_map ? PolyMap{} : make_polymap(e), e->dmap);
  }
  e->launches += 1;
  CK(cudaGetLastError());
  if (nseg > 0) CK(cudaMemcpyAsync(coef_out, A.coef, (size_t)24 * nseg * 8, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(times_out, A.times, (size_t)total * 8, cudaMemcpyDeviceToHost, s));
  if (valid_out) CK(cudaMemcpyAsync(valid_out, A.valid, (size_t)B, cudaMemcpyDeviceToHost, s));
  if (iters_out) CK(cudaMemcpyAsync(iters_out, A.iters, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
  if (status_out) CK(cudaMemcpyAsync(status_out, A.status, (size_t)B * 12, cudaMemcpyDeviceToHost, s));
  if (r_out && nseg > 0) CK(cudaMemcpyAsync(r_out, A.r_out, (size_t)nseg * 8, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return TP_OK;
}

extern "C" int tp_corridor_solve_batch(tp_engine_t* e, const tp_poly_params* p, int32_t B, const int32_t* wp_offsets, const double* waypoints,
                                       const double* bc, const double* corridor_size, double corridor_res, double* coef, double* times,
                                       int32_t* status) {
  if (!corridor_size || !status) return TP_ERR_INVALID_ARG;
  return corridor_run(e, p, B, wp_offsets, waypoints, bc, corridor_size, 1, 0, 0, 0.0, 0.0, corridor_res, coef, times, nullptr, nullptr, nullptr, status);
}

extern "C" int tp_polytraj_corridor_plan_batch(tp_engine_t* e, const tp_poly_params* p, int32_t B, const int32_t* wp_offsets,
                                               const double* waypoints, const double* bc, double init_r, double fs, double corridor_res,
                                               double* coef_out, double* times_out, uint8_t* valid_out, int32_t* iters_out,
                                               double* r_out, int32_t* status_out) {
  if (!valid_out || !iters_out) return TP_ERR_INVALID_ARG;
  return corridor_run(e, p, B, wp_offsets, waypoints, bc, nullptr, 0, 0, 0, init_r, fs, corridor_res, coef_out, times_out, valid_out, iters_out,
                      r_out, status_out);
}

extern "C" int tp_polytraj_occmap_plan_batch(tp_engine_t* e, const tp_poly_params* p, int32_t B, const int32_t* wp_offsets,
                                             const double* waypoints, const double* bc, int32_t corridor_constraint, double init_r, double fs,
                                             double corridor_res, double* coef_out, double* times_out, uint8_t* valid_out, int32_t* iters_out,
                                             double* r_out, int32_t* status_out) {
  if (!valid_out || !iters_out) return TP_ERR_INVALID_ARG;
  if (!corridor_constraint) { init_r = 1.0; fs = 0.5; }   // unused in that mode
  return corridor_run(e, p, B, wp_offsets, waypoints, bc, nullptr, 0, 1, corridor_constraint ? 0 : 1, init_r, fs, corridor_res, coef_out, times_out,
                      valid_out, iters_out, r_out, status_out);
}
