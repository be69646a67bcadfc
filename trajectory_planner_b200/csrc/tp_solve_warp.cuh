// bsplineTraj::makePlan (bsplineTraj.cpp:333-385) with ONE WARP per trajectory, start to finish: segments -> A*
// detours -> guide points -> [fused cost + L-BFGS (tp_lbfgs_warp.cuh) -> collision check -> re-guide / weight
// doubling]* -> time re-parameterisation.  A worker is a warp with its own slice of shared memory and its own A*
// node pool; workers never synchronise with each other (no block barrier anywhere), so an SM runs as many
// independent trajectories — and as many concurrent A* searches — as its shared memory holds.  Included by tp_vigo.cu
// after the outer-loop pieces (dev_plan_init / dev_plan_step / ParkQueue).
#pragma once
#include "tp_lbfgs_warp.cuh"

// ---- hasCollisionTrajectory (+ hasDynamicCollisionTrajectory), warp-wide; returns static | dynamic << 1 to every lane
__device__ __forceinline__ int dev_has_collision_w(const double* cps, int N, const BatchView& bv, const VigoConst& C,
                                                   const DevMap& map, int lane) {
  SmemCP cp{cps};
  const double cts = C.p.ctrl_pt_ts;
  const double duration = (double)(N - TP_DEGREE) * cts;          // knots_(N), bspline.cpp:27
  const double limit = (1.0 - C.p.not_check_ratio) * duration;    // bsplineTraj.h:313
  int hit = 0, dyn = 0;
  for (int s0 = 0; s0 < C.n_t_check; s0 += 32) {
    const int s = s0 + lane;
    bool live = false;
    if (s < C.n_t_check) {
      const double t = bv.t_check[s];
      if (t <= duration) {  // the table is increasing
        live = true;
        const D3 p = bspline_at(cp, N, TP_DEGREE, cts, t);
        if (t <= limit && dm_inflated(map, p)) hit = 1;
        for (int j = 0; j < bv.n_dyn; ++j) {  // bsplineTraj.h:344-368 (samples of evalTraj(): t <= duration)
          const double size = fmin(bv.dyn_size[3 * j] / 2, bv.dyn_size[3 * j + 1] / 2);
          const double dx = p.x - bv.dyn_pos[3 * j], dy = p.y - bv.dyn_pos[3 * j + 1];
          const double dist = sqrt((dx * dx + dy * dy) + 0.0 * 0.0) - size;
          if (dist < 0) dyn = 1;
        }
      }
    }
    // the reference returns at the first hit (bsplineTraj.h:319-322): later samples cannot change the answer, and
    // without dynamic obstacles there is nothing else to learn from them
    const unsigned anyhit = __ballot_sync(WF_FULL, hit != 0);
    if ((anyhit && bv.n_dyn == 0) || !__any_sync(WF_FULL, live)) break;
  }
  const int h = __any_sync(WF_FULL, hit != 0) ? 1 : 0;
  const int dn = __any_sync(WF_FULL, dyn != 0) ? 2 : 0;
  return h | dn;
}

// ---- linearFeasibilityReparam (bsplineTraj.cpp:1116-1137), warp-wide.  cp: control points (shared), q / r: scratch
// for the velocity / acceleration spline control points (3(N-1), 3(N-2) doubles, shared).  Returns the factor to all lanes.
__device__ __forceinline__ double dev_reparam_w(const double* cp, double* q, double* r, int N, const BatchView& bv,
                                                const VigoConst& C, int lane) {
  const double ts = C.p.ctrl_pt_ts;
  __syncwarp();
  for (int e = lane; e < 3 * (N - 1); e += 32) {
    const int i = e / 3;
    const double den = (double)(i + 3 + 1 - 3) * ts - (double)(i + 1 - 3) * ts;  // knots_(i+p+1) - knots_(i+1), p = 3
    q[e] = (3.0 * (cp[e + 3] - cp[e])) / den;
  }
  __syncwarp();
  for (int e = lane; e < 3 * (N - 2); e += 32) {
    const int i = e / 3;
    const double den = (double)(i + 2 + 1 - 2) * ts - (double)(i + 1 - 2) * ts;  // derivative spline: p = 2
    r[e] = (2.0 * (q[e + 3] - q[e])) / den;
  }
  __syncwarp();
  SmemCP vq{q}, ar{r};
  const double duration = (double)(N - TP_DEGREE) * ts;
  double mv = 0.0, ma = 0.0;
  for (int s = lane; s < C.n_t_reparam; s += 32) {
    const double t = bv.t_reparam[s];
    if (!(t < duration)) break;
    const double v = norm3(bspline_at(vq, N - 1, 2, ts, t));
    const double a = norm3(bspline_at(ar, N - 2, 1, ts, t));
    mv = fmax(mv, v);
    ma = fmax(ma, a);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mv = fmax(mv, __shfl_xor_sync(WF_FULL, mv, o));
    ma = fmax(ma, __shfl_xor_sync(WF_FULL, ma, o));
  }
  const double fv = C.p.max_vel / mv;
  const double fa = sqrt(C.p.max_acc / ma);
  return fmin(fv, fa);
}

struct WSolveLayout {
  int st;      // TrajState (doubles offset)
  int vf;      // start of the solver region (WfShared context, then the control points)
  int plan;    // PlanSmem (aliases the solver region after the control points)
  int total;   // doubles, even
};
__host__ __device__ inline WSolveLayout wsolve_layout(int N) {
  WSolveLayout L;
  int o = 0;
  L.st = o; o += (int)((sizeof(TrajState) + 15) / 16) * 2;
  L.vf = o;
  const int cp_d = 3 * N + (N & 1);
  L.plan = L.vf + WF_CTX + cp_d;
  const int plan_end = L.plan + (int)((sizeof(PlanSmem) + 7) / 8);
  const int rp_end = L.plan + 3 * (N - 1) + 3 * (N - 2) + 4;
  int end = L.vf + wf_layout(N).total;
  if (plan_end > end) end = plan_end;
  if (rp_end > end) end = rp_end;
  L.total = end + (end & 1);
  return L;
}

// warp-wide copy of a TrajState (ints)
__device__ __forceinline__ void w_copy_state(TrajState* dst, const TrajState* src, int lane) {
  const int* s = reinterpret_cast<const int*>(src);
  int* d = reinterpret_cast<int*>(dst);
  for (int i = lane; i < (int)(sizeof(TrajState) / 4); i += 32) d[i] = s[i];
}

// Returns 1 when the trajectory was PARKED after makePlan steps 1-3 (see k_solve): its state is back in HBM and a
// later solve_one_w(resume = 1) finishes it.
__device__ __forceinline__ int solve_one_w(const BatchView& bv, const VigoConst& C, const DevMap& map, const AStarPools& P, int b,
                                           int class_max_n, int s_slot, double* counters, long long* timeline, int resume,
                                           int rounds_this_pass, int park_thresh, double* sm, int lane) {
  long long t_start = 0;
  if (timeline && lane == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_start));
  const WSolveLayout SL = wsolve_layout(class_max_n);
  TrajState& st = *reinterpret_cast<TrajState*>(sm + SL.st);
  double* base = sm + SL.vf;
  double* cp = base + WF_CTX;
  PlanSmem& PS = *reinterpret_cast<PlanSmem*>(sm + SL.plan);
  __syncwarp();
  w_copy_state(&st, &bv.st[b], lane);
  __syncwarp();
  const int N = st.N, n = 3 * (N - 2 * TP_DEGREE);
  double* gctrl = bv.ctrl + 3 * (size_t)st.off;
  for (int e = lane; e < 3 * N; e += 32) cp[e] = gctrl[e];
  __syncwarp();
  BatchView lbv = bv;                      // the outer-loop code reads control points through bv.ctrl + 3*off
  lbv.ctrl = cp - 3 * (size_t)st.off;
  Worker W = make_worker(C, P, s_slot, lane, &PS.as);
  // ---- makePlan steps 1-3 (first pass only; a resumed trajectory continues its optimise / check / re-guide loop:
  // every optimize() starts from a fresh L-BFGS state, so splitting a solve at a round boundary changes nothing)
  if (!resume) {
    dev_plan_init(lbv, C, map, st, W, PS, b, lane);
    if (park_thresh >= 0 && lane == 0 && W.goal_unreachable) st.astar_unreach = 1;
    __syncwarp();
    if (park_thresh >= 0 && st.status == TS_ACTIVE) {
      // park it (state only — the control points have not changed)
      w_copy_state(&bv.st[b], &st, lane);
      if (timeline && lane == 0) timeline[4 * (size_t)b] = t_start;
      __syncwarp();
      return 1;
    }
  } else if (st.astar_unreach) {
    W.flood_trigger = TP_FLOOD_TRIGGER_AGAIN;
  }
  double fl = 0.0, its = 0.0, evs = 0.0, smp = 0.0;
  int rounds_done = 0;
  while (st.status == TS_ACTIVE && rounds_done < rounds_this_pass) {
    ++rounds_done;
    // ---- optimize()
    tp_lbfgs_result r;
    if (n <= 0) {
      r.ret = LB_INVALID_N; r.iters = 0; r.evals = 0; r.reserved = 0; r.fx = 0;
    } else {
      wf_setup<1>(C, base, N, bv.pairs + (size_t)b * C.gcap, bv.cp_head + st.off, st.n_pairs, st.w_dist, st.w_dyn, bv.n_dyn,
               bv.dyn_pos, bv.dyn_vel, bv.dyn_size, lane);
      lbfgs_run_team<1>(C, base, N, 0, r, nullptr, lane);
    }
    if (lane == 0) {
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
    __syncwarp();
    // ---- hasCollisionTrajectory
    int any = 1;
    if (N >= 4) any = dev_has_collision_w(cp, N, bv, C, map, lane);
    if (lane == 0) {
      st.has_col = any;
      smp += floor((double)(N - TP_DEGREE) * C.p.ctrl_pt_ts / C.check_ts) + 1.0;
    }
    __syncwarp();
    // ---- loop body: success / failure / re-guide / weight doubling
    dev_plan_step(lbv, C, map, st, W, PS, b, lane);
    if (lane == 0 && W.goal_unreachable) st.astar_unreach = 1;
    __syncwarp();
  }
  // ---- linearFeasibilityReparam
  if (st.status == TP_STATUS_SUCCESS) {
    double* q = sm + SL.plan;
    double* rr = q + 3 * (N - 1);
    const double f = dev_reparam_w(cp, q, rr, N, bv, C, lane);
    if (lane == 0) st.linear_factor = f;
  }
  __syncwarp();
  // ---- write back
  for (int e = lane; e < 3 * N; e += 32) gctrl[e] = cp[e];
  w_copy_state(&bv.st[b], &st, lane);
  if (lane == 0) {
    if (counters) {
      atomicAdd(&counters[0], fl);
      atomicAdd(&counters[1], its);
      atomicAdd(&counters[2], evs);
      atomicAdd(&counters[3], smp);
    }
    if (timeline) {   // development aid (TP_TIMELINE): start / end time [ns], SM id, iterations of every trajectory
      long long t_end;
      unsigned smid;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_end));
      asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
      if (!resume) timeline[4 * (size_t)b] = t_start;
      timeline[4 * (size_t)b + 1] = t_end;
      timeline[4 * (size_t)b + 2] = (long long)smid;
      timeline[4 * (size_t)b + 3] = (long long)st.lbfgs_iters | ((long long)st.astar_expansions << 32);
    }
  }
  __syncwarp();
  return 0;
}

// Persistent warp workers: one launch per size class; a worker of class c owns a shared-memory slice for that class's
// longest trajectory and serves its own class first, then STEALS from the classes of shorter trajectories.  Same
// two-phase schedule as k_solve (phase A: makePlan steps 1-3 and park; phase B: resume, hardest-looking first).
// A block is `blockDim.x / 32` independent workers (no block-level synchronisation at all).
__global__ void __launch_bounds__(128) k_solve_w(const __grid_constant__ BatchView bv, const __grid_constant__ VigoConst C,
                                                 const __grid_constant__ DevMap map, const __grid_constant__ AStarPools P,
                                                 const int* __restrict__ order, const int* __restrict__ cls_begin, int* cls_next,
                                                 int my_class, int class_max_n, int slice_doubles, int* slot_flags,
                                                 double* counters, long long* timeline, int resume, int rounds_this_pass,
                                                 ParkQueue park) {
  extern __shared__ double sm_all[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* sm = sm_all + (size_t)wib * slice_doubles;
  const WSolveLayout SL = wsolve_layout(class_max_n);
  // ---- claim an A* node pool (one per resident worker)
  int s_slot = 0;
  if (lane == 0) {
    const int W_ = P.workers;
    const unsigned wid = blockIdx.x * (blockDim.x >> 5) + wib;
    int sidx = (int)(((wid + 977u * (unsigned)my_class) * 2654435761u) % (unsigned)W_);
    while (atomicCAS(&slot_flags[sidx], 0, 1) != 0) sidx = sidx + 1 == W_ ? 0 : sidx + 1;
    s_slot = sidx;
  }
  s_slot = __shfl_sync(WF_FULL, s_slot, 0);
  int pb_class = -1;   // lane 0: >= 0 once this worker is in phase B
  for (;;) {
    int b = -1, cc = 0, res = resume;
    if (lane == 0) {
      if (pb_class < 0) {
        for (int c = my_class; c < 4 && b < 0; ++c) {
          if (cls_begin[c + 1] == cls_begin[c]) continue;
          const int i = cls_begin[c] + atomicAdd(&cls_next[c], 1);
          if (i < cls_begin[c + 1]) { b = order[i]; cc = c; }
        }
        if (b < 0 && park.thresh >= 0) pb_class = my_class;
      }
      while (b < 0 && pb_class >= 0 && pb_class < 4) {
        const int c = pb_class, csize = cls_begin[c + 1] - cls_begin[c];
        if (csize == 0) { ++pb_class; continue; }
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
    }
    b = __shfl_sync(WF_FULL, b, 0);
    if (b < 0) break;
    cc = __shfl_sync(WF_FULL, cc, 0);
    res = __shfl_sync(WF_FULL, res, 0);
    const int parked = solve_one_w(bv, C, map, P, b, class_max_n, s_slot, counters, timeline, res, rounds_this_pass,
                                   res ? -1 : park.thresh, sm, lane);
    if (park.thresh >= 0 && !res && lane == 0) {
      if (parked) {
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
    __syncwarp();
  }
  if (lane == 0) {
    __threadfence();
    atomicExch(&slot_flags[s_slot], 0);
  }
}

// ---- optimize() only (parity entry tp_vigo_optimize_batch, warp form): one warp per ACTIVE trajectory
__global__ void __launch_bounds__(32) k_lbfgs_w(BatchView bv, VigoConst C, const int* __restrict__ active,
                                                const int* __restrict__ n_active, tp_lbfgs_result* res_out, double* xfinal_out,
                                                double* counters) {
  extern __shared__ double sm[];
  if ((int)blockIdx.x >= *n_active) return;
  const int b = active[blockIdx.x], lane = threadIdx.x;
  TrajState& st = bv.st[b];
  const int N = st.N, n = 3 * (N - 2 * TP_DEGREE);
  tp_lbfgs_result r;
  if (n <= 0) {
    r.ret = LB_INVALID_N; r.iters = 0; r.evals = 0; r.reserved = 0; r.fx = 0;
  } else {
    double* cp = sm + WF_CTX;
    double* gctrl = bv.ctrl + 3 * (size_t)st.off;
    for (int e = lane; e < 3 * N; e += 32) cp[e] = gctrl[e];
    double* xf = xfinal_out ? xfinal_out + 3 * ((size_t)st.off - (size_t)2 * TP_DEGREE * b) : nullptr;
    wf_setup<1>(C, sm, N, bv.pairs + (size_t)b * C.gcap, bv.cp_head + st.off, st.n_pairs, st.w_dist, st.w_dyn, bv.n_dyn,
             bv.dyn_pos, bv.dyn_vel, bv.dyn_size, lane);
    lbfgs_run_team<1>(C, sm, N, 0, r, xf, lane);
    // the control points keep the last evaluated point (bsplineTraj.cpp:803)
    for (int e = lane; e < n; e += 32) gctrl[3 * TP_DEGREE + e] = cp[3 * TP_DEGREE + e];
  }
  if (lane == 0) {
    st.lbfgs_runs += 1;
    st.lbfgs_iters += r.iters;
    st.lbfgs_evals += r.evals;
    st.last_ret = r.ret;
    st.final_cost = r.fx;
    if (n > 0) st.vclock += (long long)r.evals * (10LL * N + 2LL * n);
    if (res_out) res_out[b] = r;
    if (counters && n > 0) {
      const double fl = (double)r.evals * (81.0 * N + 21.0 * st.n_pairs + 4.0 * n) + (8.0 * r.reserved + 15.0 * r.iters) * n;
      atomicAdd(&counters[0], fl);
      atomicAdd(&counters[1], (double)r.iters);
      atomicAdd(&counters[2], (double)r.evals);
    }
  }
}

// ---- costFunction (parity entry tp_vigo_cost_batch, warp form): one warp per trajectory
__global__ void __launch_bounds__(32) k_cost_w(BatchView bv, VigoConst C, double* f_out, double* grad_out) {
  extern __shared__ double sm[];
  const int b = blockIdx.x, lane = threadIdx.x;
  const TrajState& st = bv.st[b];
  const int N = st.N, n = 3 * (N - 2 * TP_DEGREE);
  if (n <= 0) {
    if (lane == 0) f_out[b] = 0.0;
    return;
  }
  const WfLayout L = wf_layout(N);
  for (int e = lane; e < 3 * N; e += 32) sm[L.cp + e] = bv.ctrl[3 * (size_t)st.off + e];
  __syncwarp();
  wf_setup<1>(C, sm, N, bv.pairs + (size_t)b * C.gcap, bv.cp_head + st.off, st.n_pairs, st.w_dist, st.w_dyn, bv.n_dyn, bv.dyn_pos,
           bv.dyn_vel, bv.dyn_size, lane);
  const WfEv evr = wf_eval<1>(C, sm, 0, lane);
  const double* ev = evr.v;
  __syncwarp();
  if (lane == 0) f_out[b] = ev[0];
  double* go = grad_out + 3 * ((size_t)st.off - (size_t)2 * TP_DEGREE * b);
  for (int e = lane; e < n; e += 32) go[e] = sm[L.g + e];
}
