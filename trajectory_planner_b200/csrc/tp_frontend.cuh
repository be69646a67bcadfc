// Device front end of the ViGO solve: B (start, goal) pairs -> initial control points, one warp
// per problem.  Same pipeline, same operation order as the host front end in tp_frontend.cpp, which restates
//   seed:   polyTrajOccMap::makePlan(false) for one segment, rest to rest (polyTrajOccMap.cpp:326-399; QP of
//           polyTrajSolver.cpp:241-271 / :314-584 / :587-813, time allocation :125-138, de-normalisation :870-879),
//           solved exactly through its 14 x 14 KKT system (Gaussian elimination with partial pivoting);
//   sample: polyTrajOccMap::getTrajectory(dt) with the accumulated t += dt (polyTrajOccMap.cpp:434-446) and
//           polyTrajSolver::getPos = sum_d c_d pow(t, d) (polyTrajSolver.cpp:1051-1071);
//   check:  bsplineTraj::inputPathCheck (bsplineTraj.cpp:207-245) inside the dt *= 0.8 loop of
//           src/bspline_node.cpp:355-366 (wall-clock cap -> 60 iterations);
//   update: bsplineTraj::updatePath (bsplineTraj.cpp:290-323): goal check, adjustPathLengthDirect (:754-793, the
//           function-static prevPathLength is 0 per problem), fillPath (:247-288), bspline::parameterizeToBspline
//           (bspline.cpp:74-138) as a Householder least-squares fit that only visits the structurally non-zero
//           entries of the (K+4) x (K+2) system (band of width 3 + 4 boundary rows): skipping exact zeros leaves every
//           rounding of the dense host loops unchanged, so the control points are bit-identical to the host's.
// pow(t, d) of glibc is correctly rounded (to < 1 ULP, in practice to nearest); the device reproduces it with a
// double-double product chain rounded once at the end.
#pragma once
#include "tp_device.cuh"
#include "../../include/tp_b200.h"

#define FE_SCAP 1024   // samples per resampling pass
#define FE_KMAX 160    // path points after inputPathCheck (control points = K + 2)

// t^d for small non-negative integer d, rounded once from a double-double value
__device__ __forceinline__ double fe_powi(double t, int d) {
  double hi = 1.0, lo = 0.0;
  for (int k = 0; k < d; ++k) {
    const double ph = hi * t;
    const double pe = __fma_rn(hi, t, -ph);
    const double pl = lo * t + pe;
    const double s = ph + pl;
    lo = pl - (s - ph);
    hi = s;
  }
  return hi;
}

struct FeSmem {
  double pts[FE_SCAP * 3];        // samples of the current pass; later: QR workspace
  double tt[FE_SCAP];             // accumulated sample times
  double fit[(FE_KMAX + 1) * 3];  // path handed to the fit
  double kkt[3][14 * 14 + 14];    // per-axis KKT system
  double coef[3][8];
  unsigned char occ[FE_SCAP];
  short keep[FE_SCAP];
  int n_s, L, K, flag;
};

// solve_dense of tp_frontend.cpp on a 14 x 14 row-major system (one thread)
__device__ inline bool fe_solve14(double* A, double* b) {
  const int n = 14;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    double best = fabs(A[c * n + c]);
    for (int r = c + 1; r < n; ++r)
      if (fabs(A[r * n + c]) > best) {
        best = fabs(A[r * n + c]);
        piv = r;
      }
    if (best < 1e-300) return false;
    if (piv != c) {
      for (int k = 0; k < n; ++k) {
        const double t = A[piv * n + k];
        A[piv * n + k] = A[c * n + k];
        A[c * n + k] = t;
      }
      const double t = b[piv];
      b[piv] = b[c];
      b[c] = t;
    }
    for (int r = c + 1; r < n; ++r) {
      const double f = A[r * n + c] / A[c * n + c];
      if (f == 0) continue;
      for (int k = c; k < n; ++k) A[r * n + k] -= f * A[c * n + k];
      b[r] -= f * b[c];
    }
  }
  for (int r = n - 1; r >= 0; --r) {
    double s = b[r];
    for (int k = r + 1; k < n; ++k) s -= A[r * n + k] * b[k];
    b[r] = s / A[r * n + r];
  }
  return true;
}

// adjustPathLengthDirect on `n` points (3 doubles each): the result is always a PREFIX of the path; returns its
// length.  Warp-collective; line queries only when the path reaches max_len at all.
__device__ inline int fe_adjust_len(const DevMap& map, const double* p, int n, double max_len, unsigned char* occ, int lane) {
  const unsigned FULL = 0xffffffffu;
  const double lim = max_len > 0.0 ? max_len : 0.0;
  const D3 p0 = d3(p[0], p[1], p[2]);
  int any = 0;
  for (int i = lane; i + 1 < n; i += 32)
    if (norm3(d3(p[3 * i + 3], p[3 * i + 4], p[3 * i + 5]) - p0) >= lim) any = 1;
  any = __any_sync(FULL, any);
  if (!any) return n;
  for (int i = lane; i + 1 < n; i += 32)
    occ[i] = dm_line(map, d3(p[3 * i], p[3 * i + 1], p[3 * i + 2]), d3(p[3 * i + 3], p[3 * i + 4], p[3 * i + 5])) ? 1 : 0;
  __syncwarp();
  int L = n;
  if (lane == 0) {
    bool exceed = false;
    double min_len = 0.0;
    for (int i = 0; i + 1 < n; ++i) {
      const D3 p1 = d3(p[3 * i], p[3 * i + 1], p[3 * i + 2]), p2 = d3(p[3 * i + 3], p[3 * i + 4], p[3 * i + 5]);
      if (norm3(p2 - p0) >= lim) exceed = true;
      if (exceed && !occ[i] && min_len >= 1.5) { L = i + 2; break; }
      if (occ[i]) min_len = 0.0;
      else min_len += norm3(p2 - p1);
    }
  }
  return __shfl_sync(FULL, L, 0);
}

// parameterizeToBspline (bspline.cpp:74-138) with zero start/end conditions on the K points in S.fit: control points
// (K + 2) x 3 to `out`.  Workspace W (in S.pts): band[K][3] (cols r..r+2 of row r), tail[K][3] (cols K-1..K+1 of row
// r where the band does not cover them), bot[4][K+2] (boundary rows), Bm[(K+4)][3], X[(K+2)][3].
struct FeQR {
  double *band, *tail, *bot, *Bm, *X;
  int K, cols;
  __device__ __forceinline__ double* at(int r, int k) const {   // structurally non-zero entries only
    if (r >= K) return bot + (size_t)(r - K) * cols + k;
    if (k >= r && k <= r + 2) return band + 3 * r + (k - r);
    return tail + 3 * r + (k - (K - 1));
  }
};
__device__ inline void fe_fit(FeSmem& S, double ts, double* out, int lane) {
  const int K = S.K, rows = K + 4, cols = K + 2;
  FeQR Q;
  Q.K = K; Q.cols = cols;
  Q.band = S.pts;
  Q.tail = Q.band + 3 * K;
  Q.bot = Q.tail + 3 * K;
  Q.Bm = Q.bot + 4 * cols;
  Q.X = Q.Bm + 3 * rows;
  const double pr[3] = {1, 4, 1}, vr[3] = {-1, 0, 1}, ar[3] = {1, -2, 1};
  for (int e = lane; e < 3 * K; e += 32) {
    Q.band[e] = (1 / 6.0) * pr[e % 3];
    Q.tail[e] = 0.0;
  }
  for (int e = lane; e < 4 * cols; e += 32) Q.bot[e] = 0.0;
  for (int e = lane; e < 3 * K; e += 32) Q.Bm[e] = S.fit[e];
  for (int e = lane; e < 12; e += 32) Q.Bm[3 * K + e] = 0.0;   // zero start / end velocity and acceleration
  __syncwarp();
  if (lane < 3) {
    const int k = lane;
    Q.bot[0 * cols + k] = (1 / 2.0 / ts) * vr[k];
    Q.bot[1 * cols + K - 1 + k] = (1 / 2.0 / ts) * vr[k];
    Q.bot[2 * cols + k] = (1 / ts / ts) * ar[k];
    Q.bot[3 * cols + K - 1 + k] = (1 / ts / ts) * ar[k];
  }
  __syncwarp();
  // Householder QR, column by column; the reflector of column c lives on row c and the boundary rows
  for (int c = 0; c < cols; ++c) {
    int rset[5], nr = 0;
    if (c < K) rset[nr++] = c;
    for (int r = (c < K ? K : c); r < rows; ++r) rset[nr++] = r;
    double v[5], nrm = 0.0;
    for (int i = 0; i < nr; ++i) {
      v[i] = *Q.at(rset[i], c);
      nrm += v[i] * v[i];
    }
    nrm = sqrt(nrm);
    if (nrm == 0.0) continue;   // warp-uniform
    const double alpha = v[0] > 0 ? -nrm : nrm;
    v[0] -= alpha;
    double vn = 0.0;
    for (int i = 0; i < nr; ++i) vn += v[i] * v[i];
    if (vn == 0.0) continue;
    __syncwarp();
    // columns with a non-zero entry on those rows: c, c+1, c+2 and K-1, K, K+1 (each once), then the 3 right-hand sides
    int kcol = -1;
    if (lane < 3) {
      if (c + lane < cols) kcol = c + lane;
    } else if (lane < 6) {
      const int k = K - 1 + (lane - 3);
      if (k > c + 2 && k < cols) kcol = k;
    }
    if (kcol >= 0) {
      double s = 0.0;
      for (int i = 0; i < nr; ++i) s += v[i] * *Q.at(rset[i], kcol);
      s = 2 * s / vn;
      for (int i = 0; i < nr; ++i) *Q.at(rset[i], kcol) -= s * v[i];
    } else if (lane >= 6 && lane < 9) {
      const int k = lane - 6;
      double s = 0.0;
      for (int i = 0; i < nr; ++i) s += v[i] * Q.Bm[3 * rset[i] + k];
      s = 2 * s / vn;
      for (int i = 0; i < nr; ++i) Q.Bm[3 * rset[i] + k] -= s * v[i];
    }
    __syncwarp();
  }
  // back substitution, one lane per right-hand side; row r of R: cols r+1, r+2 and K-1..K+1
  if (lane < 3) {
    const int k = lane;
    for (int r = cols - 1; r >= 0; --r) {
      double s = Q.Bm[3 * r + k];
      for (int j = r + 1; j < cols; ++j) {
        if (r < K && j > r + 2 && j < K - 1) { j = K - 2; continue; }   // structurally zero stretch
        s -= *Q.at(r, j) * Q.X[3 * j + k];
      }
      Q.X[3 * r + k] = s / *Q.at(r, r);
    }
  }
  __syncwarp();
  for (int e = lane; e < 3 * cols; e += 32) out[e] = Q.X[e];
  __syncwarp();
}

// one warp per (start, goal) pair; control points to ctrl_tmp[b][(FE_KMAX + 2) * 3], count[b] = control points (0 = the
// reference's updatePath returns false / caps exceeded)
__global__ void __launch_bounds__(32) k_frontend(DevMap map, tp_vigo_params p, int B, const double* __restrict__ starts,
                                                 const double* __restrict__ goals, double* ctrl_tmp, int* count) {
  extern __shared__ double fe_raw[];
  FeSmem& S = *reinterpret_cast<FeSmem*>(fe_raw);
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const D3 s = d3(starts[3 * b], starts[3 * b + 1], starts[3 * b + 2]);
    const D3 g = d3(goals[3 * b], goals[3 * b + 1], goals[3 * b + 2]);
    const double dist = norm3(g - s);
    bool ok = dist > 0;
    const double T = dist / p.max_vel;   // time allocation, polyTrajSolver.cpp:125-138
    // ---- seed: 14 x 14 KKT system per axis (lanes 0-2)
    if (ok) {
      int good = 1;
      if (lane < 3) {
        double* Kk = S.kkt[lane];
        double* bb = Kk + 196;
        for (int e = 0; e < 210; ++e) Kk[e] = 0.0;
        for (int i = 4; i < 8; ++i)
          for (int j = 4; j < 8; ++j) {
            double f = 1.0;
            for (int d = 0; d < 4; ++d) f *= (double)((i - d) * (j - d));
            f /= (double)(i + j - 7);
            Kk[i * 14 + j] = f;
          }
        for (int j = 0; j < 8; ++j) {
          double a[6] = {0, 0, 0, 0, 0, 0};
          a[0] = j == 0 ? 1.0 : 0.0;
          a[1] = 1.0;
          a[2] = j == 1 ? 1.0 : 0.0;
          a[3] = j >= 1 ? (double)j : 0.0;
          a[4] = j == 2 ? 2.0 : 0.0;
          a[5] = j >= 2 ? (double)(j * (j - 1)) : 0.0;
          for (int r = 0; r < 6; ++r) {
            Kk[(8 + r) * 14 + j] = a[r];
            Kk[j * 14 + 8 + r] = a[r];
          }
        }
        bb[8] = lane == 0 ? s.x : (lane == 1 ? s.y : s.z);
        bb[9] = lane == 0 ? g.x : (lane == 1 ? g.y : g.z);
        good = fe_solve14(Kk, bb) ? 1 : 0;
        for (int d = 0; d < 8; ++d) S.coef[lane][d] = bb[d] / fe_powi(T, d);
      }
      ok = __all_sync(FULL, good);
    }
    __syncwarp();
    // ---- resample until inputPathCheck is satisfied (src/bspline_node.cpp:355-366)
    int K = 0;
    if (ok) {
      double dt = p.ctrl_pt_dist / p.max_vel;   // bsplineTraj::getInitTs
      bool have = false;
      for (int it = 0; it < 60 && !have; ++it, dt *= 0.8) {
        if (lane == 0) {
          int n = 0;
          for (double t = 0; t <= T; t += dt) {
            if (n < FE_SCAP) S.tt[n] = t;
            if (++n > FE_SCAP) break;
          }
          S.n_s = n;
        }
        __syncwarp();
        const int n = S.n_s;
        if (n > FE_SCAP) break;   // finer than the sample buffer: give up (never reached with sane parameters)
        for (int i = lane; i < n; i += 32) {
          const double t = S.tt[i];
          double x = 0, y = 0, z = 0;
          for (int d = 0; d < 8; ++d) {
            const double pw = fe_powi(t, d);
            x += S.coef[0][d] * pw;
            y += S.coef[1][d] * pw;
            z += S.coef[2][d] * pw;
          }
          S.pts[3 * i] = x; S.pts[3 * i + 1] = y; S.pts[3 * i + 2] = z;
        }
        __syncwarp();
        if (n == 0) continue;
        const int L = fe_adjust_len(map, S.pts, n, p.max_path_length, S.occ, lane);
        int bad = 0;
        for (int i = lane; i + 1 < L; i += 32)
          if (norm3(d3(S.pts[3 * i], S.pts[3 * i + 1], S.pts[3 * i + 2]) - d3(S.pts[3 * i + 3], S.pts[3 * i + 4], S.pts[3 * i + 5])) >
              p.ctrl_pt_dist * 1.5)
            bad = 1;
        if (__any_sync(FULL, bad)) continue;
        // thinning: keep a point when it is >= 0.8 ctrl_pt_dist from the last kept one, then repeat the last
        if (lane == 0) {
          int m = 0;
          D3 prev = d3(S.pts[0], S.pts[1], S.pts[2]);
          for (int i = 0; i < L; ++i) {
            const D3 q = d3(S.pts[3 * i], S.pts[3 * i + 1], S.pts[3 * i + 2]);
            if (i == 0 || norm3(q - prev) >= p.ctrl_pt_dist * 0.8) {
              if (m < FE_KMAX) S.keep[m] = (short)i;
              ++m;
              prev = q;
            }
          }
          S.K = m;
        }
        __syncwarp();
        K = S.K;
        have = true;
      }
      if (!have || K == 0 || K + 1 > FE_KMAX) ok = false;
    }
    // `best` = kept points + the last one again
    if (ok) {
      for (int i = lane; i < K; i += 32) {
        const int src = S.keep[i];
        S.fit[3 * i] = S.pts[3 * src]; S.fit[3 * i + 1] = S.pts[3 * src + 1]; S.fit[3 * i + 2] = S.pts[3 * src + 2];
      }
      __syncwarp();
      if (lane < 3) S.fit[3 * K + lane] = S.fit[3 * (K - 1) + lane];
      __syncwarp();
      K += 1;
      // updatePath: goal inside an inflated obstacle -> false (:291-295)
      if (dm_inflated(map, d3(S.fit[3 * (K - 1)], S.fit[3 * (K - 1) + 1], S.fit[3 * (K - 1) + 2]))) ok = false;
    }
    int Kin = 0;
    if (ok) {
      Kin = fe_adjust_len(map, S.fit, K, p.max_path_length, S.occ, lane);
      if (Kin < 4) {   // fillPath works on the whole path (:304-312)
        if (K <= 1) ok = false;
        else if (K == 2) {
          if (lane < 3) {
            const double a0 = S.fit[lane], a1 = S.fit[3 + lane], d = a1 - a0;
            S.fit[3 + lane] = d / 3.0 + a0;
            S.fit[6 + lane] = 2.0 * d / 3.0 + a0;
            S.fit[9 + lane] = a1;
          }
          Kin = 4;
        } else if (K == 3) {
          if (lane < 3) {
            const double a0 = S.fit[lane], a1 = S.fit[3 + lane], a2 = S.fit[6 + lane];
            S.fit[3 + lane] = (a0 + a1) / 2.0;
            S.fit[6 + lane] = a1;
            S.fit[9 + lane] = (a1 + a2) / 2.0;
            S.fit[12 + lane] = a2;
          }
          Kin = 5;
        } else {
          Kin = K;
        }
        __syncwarp();
      }
    }
    int N = 0;
    if (ok) {
      S.K = Kin;
      __syncwarp();
      fe_fit(S, p.ctrl_pt_ts, ctrl_tmp + (size_t)b * (FE_KMAX + 2) * 3, lane);
      N = Kin + 2;
    }
    if (lane == 0) count[b] = N;
    __syncwarp();
  }
}

// exclusive scan of count[B] into offsets[B + 1] (one block)
__global__ void __launch_bounds__(1024) k_fe_scan(const int* __restrict__ count, int B, int* offsets) {
  __shared__ int part[1024];
  const int tid = threadIdx.x, per = (B + 1023) / 1024;
  const int lo = tid * per, hi = min(B, lo + per);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += count[i];
  part[tid] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int v = tid >= o ? part[tid - o] : 0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  int run = part[tid] - s;
  for (int i = lo; i < hi; ++i) {
    offsets[i] = run;
    run += count[i];
  }
  if (tid == 1023) offsets[B] = part[1023];
}
__global__ void k_fe_gather(const double* __restrict__ ctrl_tmp, const int* __restrict__ offsets, int B, double* ctrl, long cap,
                            unsigned char* valid) {
  const int b = blockIdx.x;
  if (b >= B) return;
  const int o = offsets[b], n = offsets[b + 1] - o;
  if (valid && threadIdx.x == 0) valid[b] = n > 0 ? 1 : 0;
  if ((long)o + n > cap) return;
  for (int e = threadIdx.x; e < 3 * n; e += blockDim.x) ctrl[3 * (size_t)o + e] = ctrl_tmp[(size_t)b * (FE_KMAX + 2) * 3 + e];
}
