// Corridor-constrained min-snap: polyTrajOctomap::makePlanCorridorConstraint (polyTrajOctomap.cpp:388-530)
// and the inequality QP polyTrajSolver builds for it (setCorridorConstraint / updateCorridorParam, polyTrajSolver.cpp:960-1012;
// corridor rows of constructA :555-579 and constructBound :813-840).  Included after tp_poly.cuh (same translation unit).
//
//   min 1/2 c'Pc   s.t.  Aeq c = b  (the min-snap equality rows),   mid_j - r_s <= [1 t_j .. t_j^7] c_s <= mid_j + r_s
//
// per axis, rows j = the corridor samples t = 0; t <= 1; t += 1/ceil(duration_s * corridorRes) of every segment s, mid_j the
// point of the straight segment at t_j.  The reference hands it to OSQP (eps 1e-3: its answers violate the corridor by
// millimetres, tests/golden/minsnap_osqp_golden.npz); here it is solved to convergence by a Mehrotra predictor-corrector
// interior-point method — the algorithm, start point and update formulas of oracle/polytraj_np.corridor_qp_ipm, which the
// tests compare against.  Each Newton system (P + Ac' S Ac) dc + Aeq' dy = ..., Aeq dc = ... has the sparsity of the
// equality-only KKT system (Ac' S Ac is block diagonal, 8 x 8 per segment): the same BAND matrix (tp_band.cuh), factorised
// by warp 0 of the block (one factorisation serves the predictor and the corrector); the per-row vector work runs on all
// four warps.
// One thread block per path runs the whole loop: solve -> sample -> box collision check -> shrink the corridors of the
// colliding segments by fs -> re-solve, until collision free, an infeasible QP, or maxIter.
#pragma once

#define CR_MAX_ROWS 4096      // corridor rows per path the solver accepts
#define CR_NVEC 14            // per-row work vectors
#define CR_IPM_MAX_IT 60
#define CR_IPM_MU_TOL 1e-10
#define CR_IPM_RES_TOL 1e-6

struct CorridorArgs {
  int B;
  const int* wp_off;       // [B+1]
  const double* wp;        // [3 * total]
  const double* bc;        // [B * 12] or null
  double desired_vel;
  int cont, max_iter;
  double init_r, fs, corridor_res;
  const double* r_in;      // solve-only mode: per-segment radii at wp_off[b] - b + s; else null
  int solve_only;          // 1: one QP solve with r_in, no collision loop
  int occmap;              // 1: polyTrajOccMap's collision test on the ViGO occupancy map instead of the octomap box check
  int no_corridor;         // 1: polyTrajOccMap::makePlan(corridorConstraint = false): the equality-only solve, always valid
  const double* t_acc;
  int n_t_acc;
  double box[3];
  double map_res;
  double* coef;            // out: 24 * (wp_off[b] - b) + a * 8 K + 8 s + d
  double* times;           // out: [total]
  uint8_t* valid;          // out [B]
  int* iters;              // out [B] outer iterations
  int* status;             // out [3 B] per axis: 0 converged, 1 infeasible / not converged, < 0 the solver's error
  double* r_out;           // out: final radii per segment (may be null)
  double* scratch;
  size_t stride;           // doubles per resident block: corridor_scratch_doubles(kmax)
  int kmax;                // segments of the longest path
  int* queue;
};

__host__ __device__ inline size_t corridor_scratch_doubles(int kmax) {
  const size_t n = (size_t)band_order(kmax, 4);
  return 2 * n * BD_W + 3 * n           // A0, A, R
         + 3 * n + 2 * n + 2 * n        // X[3], DX, RD, piv + kind (ints)
         + (size_t)kmax + 8             // dt
         + (size_t)CR_MAX_ROWS * (8 + 1 + 3 + 2 + CR_NVEC);   // Arow, seg, mid, lo / hi, work vectors
}

struct BlkRed {   // block-wide reductions (128 threads), result to every thread
  double* s;
  __device__ __forceinline__ double sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    return (s[0] + s[1]) + (s[2] + s[3]);
  }
  __device__ __forceinline__ double max(double v) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    return fmax(fmax(s[0], s[1]), fmax(s[2], s[3]));
  }
  __device__ __forceinline__ double min(double v) { return -max(-v); }
};

// band_factor / band_solve (one warp) inside a block: warp 0 works, the block waits
__device__ int blk_band_factor(double* A, int n, int* piv, bool guard) {
  __shared__ int f_rc;
  if (threadIdx.x < 32) {
    const int rc = band_factor(A, n, piv, threadIdx.x, guard);
    if (threadIdx.x == 0) f_rc = rc;
  }
  __syncthreads();
  const int rc = f_rc;
  __syncthreads();
  return rc;
}
__device__ void blk_band_solve(const double* A, int n, const int* piv, double* x) {
  if (threadIdx.x < 32) band_solve(A, n, piv, x, 1, threadIdx.x);
  __syncthreads();
}

struct CorridorWork {   // one path's scratch, carved from the resident block's slice
  double *A0, *A, *R, *X, *DX, *RD, *T1, *dt, *Arow, *mid, *lo, *hi, *V;
  int *piv, *kind, *seg;   // kind[i]: coefficient index 8 s + d of band row i, or -1 for a constraint row
  int n, nvar, mc, K, cont;
};

// One axis of the corridor QP by the whole block (see the header comment).  X (n) holds the equality-only solution [c; y]
// on entry and the minimiser on return.  Returns (to every thread) 0 converged / 1 infeasible or not converged.
__device__ int corridor_ipm(const CorridorWork& W, int ax, const int* rs /*shared: first row of each segment, [K+1]*/, double* s_red) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = W.n, mc = W.mc, K = W.K;
  const BandLayout L(K, W.cont);
  double* x = W.X + (size_t)ax * n;
  double *su = W.V, *sl = su + CR_MAX_ROWS, *lu = sl + CR_MAX_ROWS, *ll = lu + CR_MAX_ROWS, *ru = ll + CR_MAX_ROWS,
         *rl = ru + CR_MAX_ROWS, *dsu = rl + CR_MAX_ROWS, *dsl = dsu + CR_MAX_ROWS, *dlu = dsl + CR_MAX_ROWS,
         *dll = dlu + CR_MAX_ROWS, *gv = dll + CR_MAX_ROWS, *rcu = gv + CR_MAX_ROWS, *rcl = rcu + CR_MAX_ROWS, *sig = rcl + CR_MAX_ROWS;
  const double* lo = W.lo;
  const double* hi = W.hi;
  BlkRed red{s_red};
  auto row_dot = [&](int j, const double* v) {   // [1 t .. t^7] . v_segment
    const double* a = W.Arow + 8 * (size_t)j;
    const double* c = v + L.var(W.seg[j], 0);   // a segment's coefficients are contiguous in band order
    double acc = 0.0;
    for (int d = 0; d < PL_NC; ++d) acc += a[d] * c[d];
    return acc;
  };
  for (int j = tid; j < mc; j += PL_THREADS) {
    const double w = row_dot(j, x), span = hi[j] - lo[j];
    su[j] = fmax(hi[j] - w, 1e-2 * span);
    sl[j] = fmax(w - lo[j], 1e-2 * span);
    lu[j] = 1.0;
    ll[j] = 1.0;
  }
  __syncthreads();
  // one Newton step with complementarity targets rcu / rcl: DX = [dc; dy], dsu, dsl, dlu, dll
  auto newton = [&]() {
    for (int j = tid; j < mc; j += PL_THREADS)
      gv[j] = (-rcu[j] + lu[j] * ru[j]) / su[j] + (rcl[j] + ll[j] * rl[j]) / sl[j];
    __syncthreads();
    for (int i = tid; i < n; i += PL_THREADS) {
      double v = -W.RD[i];
      const int q = W.kind[i];
      if (q >= 0) {
        const int s = q / PL_NC, d = q - s * PL_NC;
        double acc = 0.0;
        for (int j = rs[s]; j < rs[s + 1]; ++j) acc += W.Arow[8 * (size_t)j + d] * gv[j];
        v -= acc;
      }
      W.DX[i] = v * W.T1[i];
    }
    __syncthreads();
    blk_band_solve(W.A, n, W.piv, W.DX);
    for (int i = tid; i < n; i += PL_THREADS) W.DX[i] *= W.T1[i];
    __syncthreads();
    for (int j = tid; j < mc; j += PL_THREADS) {
      const double adc = row_dot(j, W.DX);
      const double a = -ru[j] - adc, b = rl[j] + adc;
      dsu[j] = a;
      dsl[j] = b;
      dlu[j] = (-rcu[j] - lu[j] * a) / su[j];
      dll[j] = (-rcl[j] - ll[j] * b) / sl[j];
    }
    __syncthreads();
  };
  auto max_steps = [&](double& ap, double& ad) {   // largest steps in (0, 1] that keep slacks / multipliers non-negative
    double p = 1.0, q = 1.0;
    for (int j = tid; j < mc; j += PL_THREADS) {
      if (dsu[j] < 0) p = fmin(p, -su[j] / dsu[j]);
      if (dsl[j] < 0) p = fmin(p, -sl[j] / dsl[j]);
      if (dlu[j] < 0) q = fmin(q, -lu[j] / dlu[j]);
      if (dll[j] < 0) q = fmin(q, -ll[j] / dll[j]);
    }
    ap = red.min(p);
    ad = red.min(q);
  };
  int status = 1;
  for (int it = 0; it < CR_IPM_MAX_IT; ++it) {
    // RD = K0 x - [0; b] (+ Ac'(lu - ll) on the coefficient rows): dual residual rd and primal residual rp; |Pc|_inf
    double pcmax = 0.0, resmax = 0.0, comp = 0.0;
    for (int i = tid; i < n; i += PL_THREADS) {
      const int j0 = i - BD_KL > 0 ? i - BD_KL : 0, j1 = i + BD_KL < n - 1 ? i + BD_KL : n - 1;   // K0 itself has bandwidth 13
      double acc = 0.0;
      for (int j = j0; j <= j1; ++j) acc += W.A0[bd_idx(i, j)] * x[j];
      const int q = W.kind[i];
      double v;
      if (q >= 0) {
        const int s = q / PL_NC, d = q - s * PL_NC, v0 = L.var(s, 0);
        double pc = 0.0;
        for (int e = 0; e < PL_NC; ++e) pc += W.A0[bd_idx(i, v0 + e)] * x[v0 + e];
        pcmax = fmax(pcmax, fabs(pc));
        double g = 0.0;
        for (int j = rs[s]; j < rs[s + 1]; ++j) g += W.Arow[8 * (size_t)j + d] * (lu[j] - ll[j]);
        v = acc + g;
      } else {
        v = acc - W.R[3 * i + ax];
      }
      W.RD[i] = v;
      resmax = fmax(resmax, fabs(v));
    }
    for (int j = tid; j < mc; j += PL_THREADS) {
      const double w = row_dot(j, x);
      const double a = w + su[j] - hi[j], b = w - sl[j] - lo[j];
      ru[j] = a;
      rl[j] = b;
      resmax = fmax(resmax, fmax(fabs(a), fabs(b)));
      comp += su[j] * lu[j] + sl[j] * ll[j];
      sig[j] = lu[j] / su[j] + ll[j] / sl[j];
    }
    const double mu = red.sum(comp) / (2.0 * mc);
    const double res = red.max(resmax);
    const double pcm = red.max(pcmax);
    if (!(mu < 1e30) || !isfinite(res)) break;
    if (mu < CR_IPM_MU_TOL && res < CR_IPM_RES_TOL * (1.0 + pcm)) { status = 0; break; }
    // Newton matrix: K0 with Ac' S Ac added to the segments' 8 x 8 diagonal blocks
    for (size_t e = tid; e < (size_t)n * BD_W; e += PL_THREADS) W.A[e] = W.A0[e];
    __syncthreads();
    for (int e = tid; e < K * PL_NC * PL_NC; e += PL_THREADS) {
      const int s = e / (PL_NC * PL_NC), q = e - s * PL_NC * PL_NC, d1 = q / PL_NC, d2 = q - d1 * PL_NC;
      double acc = 0.0;
      for (int j = rs[s]; j < rs[s + 1]; ++j) acc += sig[j] * W.Arow[8 * (size_t)j + d1] * W.Arow[8 * (size_t)j + d2];
      W.A[bd_idx(L.var(s, d1), L.var(s, d2))] += acc;
    }
    __syncthreads();
    // symmetric equilibration D K D, D_i = 1 / sqrt(K_ii) on the coefficient rows whose diagonal carries a barrier term
    // (they reach 1e19 in the last iterations, beside constraint entries of order 1): the factorisation keeps its digits
    for (int i = tid; i < n; i += PL_THREADS) {
      const double dii = W.kind[i] >= 0 ? fabs(W.A[bd_idx(i, i)]) : 0.0;
      W.T1[i] = dii > 1.0 ? 1.0 / sqrt(dii) : 1.0;
    }
    __syncthreads();
    for (int e = tid; e < n * BD_W; e += PL_THREADS) {
      const int i = e / BD_W, j = i + (e - i * BD_W) - BD_KL;
      if (j >= 0 && j < n) W.A[e] *= W.T1[i] * W.T1[j];
    }
    __syncthreads();
    if (blk_band_factor(W.A, n, W.piv, true) != 0) break;
    // predictor
    for (int j = tid; j < mc; j += PL_THREADS) { rcu[j] = su[j] * lu[j]; rcl[j] = sl[j] * ll[j]; }
    __syncthreads();
    newton();
    double ap, ad;
    max_steps(ap, ad);
    double ca = 0.0;
    for (int j = tid; j < mc; j += PL_THREADS)
      ca += (su[j] + ap * dsu[j]) * (lu[j] + ad * dlu[j]) + (sl[j] + ap * dsl[j]) * (ll[j] + ad * dll[j]);
    const double mu_aff = red.sum(ca) / (2.0 * mc);
    const double r3 = mu_aff / mu, sigma = r3 * r3 * r3;
    // corrector
    for (int j = tid; j < mc; j += PL_THREADS) {
      rcu[j] = su[j] * lu[j] + dsu[j] * dlu[j] - sigma * mu;
      rcl[j] = sl[j] * ll[j] + dsl[j] * dll[j] - sigma * mu;
    }
    __syncthreads();
    newton();
    max_steps(ap, ad);
    ap = fmin(0.995 * ap, 1.0);
    ad = fmin(0.995 * ad, 1.0);
    for (int i = tid; i < n; i += PL_THREADS) x[i] += (W.kind[i] >= 0 ? ap : ad) * W.DX[i];
    for (int j = tid; j < mc; j += PL_THREADS) {
      su[j] += ap * dsu[j];
      sl[j] += ap * dsl[j];
      lu[j] += ad * dlu[j];
      ll[j] += ad * dll[j];
    }
    __syncthreads();
  }
  __syncthreads();
  return status;
}

__global__ void __launch_bounds__(PL_THREADS) k_corridor_loop(CorridorArgs A, PolyMap map, DevMap dmap) {
  __shared__ int s_b, s_go;
  __shared__ double s_dt[PL_MAX_SEG + 1], s_r[PL_MAX_SEG + 1], s_red[PL_THREADS / 32];
  __shared__ int s_rs[PL_MAX_SEG + 2];
  __shared__ uint8_t s_seg[PL_MAX_SEG + 1];
  const int tid = threadIdx.x;
  double* base = A.scratch + (size_t)blockIdx.x * A.stride;
  const size_t nm = (size_t)band_order(A.kmax, 4);
  CorridorWork W;
  W.A0 = base; W.A = W.A0 + nm * BD_W; W.R = W.A + nm * BD_W; W.X = W.R + 3 * nm; W.DX = W.X + 3 * nm; W.RD = W.DX + nm;
  W.piv = reinterpret_cast<int*>(W.RD + nm); W.kind = W.piv + nm; W.T1 = W.RD + 2 * nm; W.dt = W.RD + 3 * nm;
  W.Arow = W.dt + A.kmax + 8; W.seg = reinterpret_cast<int*>(W.Arow + 8 * (size_t)CR_MAX_ROWS); W.mid = W.Arow + 9 * (size_t)CR_MAX_ROWS;
  W.lo = W.mid + 3 * (size_t)CR_MAX_ROWS; W.hi = W.lo + CR_MAX_ROWS; W.V = W.hi + CR_MAX_ROWS;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_b = atomicAdd(A.queue, 1);
    __syncthreads();
    const int b = s_b;
    if (b >= A.B) break;
    const int w0 = A.wp_off[b], nwp = A.wp_off[b + 1] - w0, K = nwp - 1;
    const double* wp = A.wp + 3 * (size_t)w0;
    double* coef = A.coef + (size_t)3 * PL_NC * (w0 - b);
    double* times = A.times + w0;
    if (tid == 0) { A.valid[b] = 0; A.iters[b] = 0; A.status[3 * b] = A.status[3 * b + 1] = A.status[3 * b + 2] = 0; }
    if (nwp < 2) {   // single-point path (polyTrajOctomap.cpp:390-395)
      if (tid == 0) A.valid[b] = 1;
      continue;
    }
    if (K > PL_MAX_SEG) {
      if (tid == 0) A.status[3 * b] = A.status[3 * b + 1] = A.status[3 * b + 2] = -2;
      continue;
    }
    const BandLayout L(K, A.cont);
    const int n = L.n, nvar = PL_NC * K;
    W.n = n; W.nvar = nvar; W.K = K; W.cont = A.cont;
    if (tid < 32) band_build(wp, nwp, A.bc ? A.bc + 12 * (size_t)b : nullptr, A.desired_vel, A.cont, times, W.A0, W.R, W.dt, tid);
    for (int i = tid; i < n; i += PL_THREADS) W.kind[i] = -1;
    __syncthreads();
    for (int q = tid; q < nvar; q += PL_THREADS) W.kind[L.var(q / PL_NC, q % PL_NC)] = q;
    for (int s = tid; s < K; s += PL_THREADS) s_dt[s] = W.dt[s];
    __syncthreads();
    // ---- corridor rows (updateCorridorParam): per segment t = 0; t <= 1; t += 1 / ceil(duration * corridorRes)
    if (tid == 0) {
      int m = 0;
      for (int s = 0; s < K; ++s) {
        s_rs[s] = m;
        s_r[s] = A.r_in ? A.r_in[(w0 - b) + s] : A.init_r;
        const int num = (int)ceil(s_dt[s] * A.corridor_res);
        const double dt = 1.0 / num;
        for (double t = 0; t <= 1.0 && m <= CR_MAX_ROWS; t += dt) {
          if (m < CR_MAX_ROWS) {
            for (int d = 0; d < PL_NC; ++d) W.Arow[8 * (size_t)m + d] = pow(t, (double)d);
            W.seg[m] = s;
            for (int a = 0; a < 3; ++a) W.mid[3 * (size_t)m + a] = wp[3 * s + a] + (wp[3 * (s + 1) + a] - wp[3 * s + a]) * (t - 0.0) / (1.0 - 0.0);
          }
          ++m;
        }
      }
      s_rs[K] = m;
    }
    __syncthreads();
    const int mc = s_rs[K];
    if (mc > CR_MAX_ROWS) {
      if (tid == 0) A.status[3 * b] = A.status[3 * b + 1] = A.status[3 * b + 2] = -4;
      continue;
    }
    W.mc = mc;
    int it = 0;
    for (;;) {
      int bad = 0;
      for (int ax = 0; ax < 3; ++ax) {
        double* x = W.X + (size_t)ax * n;
        // start point: K0^-1 [0; b]  (K0 is re-factorised per axis: the interior-point iterations overwrite A)
        for (size_t e = tid; e < (size_t)n * BD_W; e += PL_THREADS) W.A[e] = W.A0[e];
        for (int i = tid; i < n; i += PL_THREADS) x[i] = W.R[3 * i + ax];
        __syncthreads();
        int st = blk_band_factor(W.A, n, W.piv, false) != 0 ? -1 : 0;
        if (st == 0) {
          blk_band_solve(W.A, n, W.piv, x);
          for (int j = tid; j < mc; j += PL_THREADS) {
            const double r = s_r[W.seg[j]], m = W.mid[3 * (size_t)j + ax];
            W.lo[j] = m - r;
            W.hi[j] = m + r;
          }
          __syncthreads();
          st = (mc > 0 && !A.no_corridor) ? corridor_ipm(W, ax, s_rs, s_red) : 0;
        }
        if (tid == 0) A.status[3 * b + ax] = st;
        if (st != 0) bad = 1;
        // de-normalise: c_d /= dt^d (solveX..Z, polyTrajSolver.cpp:874-878)
        for (int q = tid; q < nvar; q += PL_THREADS) {
          const int s = q / PL_NC, d = q - s * PL_NC;
          coef[(size_t)ax * nvar + q] = x[L.var(s, d)] / pow(s_dt[s], (double)d);
        }
        __syncthreads();
      }
      ++it;
      if (A.solve_only || bad) break;
      if (A.no_corridor) {   // polyTrajOccMap.cpp:370-374: no collision check in this mode
        if (tid == 0) A.valid[b] = 1;
        break;
      }
      int ntraj = 0;
      const int any = poly_check_one(map, wp, K, coef, times, A.t_acc, A.n_t_acc, A.box, A.map_res, s_seg, &ntraj, nullptr, nullptr, 0,
                                     A.occmap ? &dmap : nullptr);
      if (tid == 0) {
        int go = 0;
        if (!any) A.valid[b] = 1;
        else {
          for (int s = 0; s < K; ++s)
            if (s_seg[s]) s_r[s] = s_r[s] * A.fs;   // adjustCorridorSize (polyTrajOctomap.cpp:188-192), before the count check
          go = it <= A.max_iter;                    // ++countIter; if (countIter > maxIter_) break;  (:431-434)
        }
        s_go = go;
      }
      __syncthreads();
      if (!s_go) break;
    }
    if (tid == 0) A.iters[b] = it;
    if (A.r_out)
      for (int s = tid; s < K; s += PL_THREADS) A.r_out[(w0 - b) + s] = s_r[s];
  }
}
