// Banded KKT solver of the min-snap path: ONE WARP per problem.
//
// The QP of polyTrajSolver (P: polyTrajSolver.cpp:241-271, equality rows :314-584, bounds :587-813) couples a segment's 8
// coefficients only with the constraints at its two ends.  Ordered segment by segment —
//     [start rows: pos, vel, acc] [c_0] [junction 0 rows] [c_1] [junction 1 rows] ... [c_(K-1)] [end rows: pos, vel, acc]
// with junction i = { waypoint position of segment i at 1;  continuity of position, velocity, acceleration (, jerk, snap)
// between segments i and i + 1 } — the KKT matrix [P A^T; A 0] is BANDED with half-bandwidth 13 (8 + the 6 junction rows
// - 1).  Round 1 factorised it as a dense n x n matrix (n = 14 K: n^3 / 3 flops, 6 MB of scratch at K = 63, one thread
// block per problem, three block barriers per pivot); here it is an LAPACK-style band LU with partial pivoting (dgbtf2:
// pivot among the 14 rows of the band, multipliers stored in place and not swapped, U grows to bandwidth 26) in band
// storage (40 doubles per row: 58 KB at K = 13) — n * 13 * 26 multiply-adds per factorisation, no block barrier at all.
// The equations and unknowns are those of the reference; only their order differs (the order the reference hands its rows
// to OSQP in is irrelevant to the solution).
#pragma once

#define BD_KL 13                       // lower bandwidth (= upper bandwidth of the matrix itself)
#define BD_KU 26                       // upper bandwidth of U after partial pivoting (kl + ku)
#define BD_W (BD_KL + BD_KU + 1)       // doubles stored per row: columns i - 13 .. i + 26
#define BD_FULL 0xffffffffu

__host__ __device__ inline size_t bd_idx(int i, int j) { return (size_t)i * BD_W + (size_t)(j - i + BD_KL); }

struct BandLayout {
  int K, J, n;   // segments, rows per junction (cont + 2), order of the system
  __host__ __device__ BandLayout(int K_, int cont) : K(K_), J(cont + 2), n(6 + 8 * K_ + (K_ - 1) * (cont + 2)) {}
  __host__ __device__ int var(int s, int d) const { return 3 + s * (8 + J) + d; }                 // coefficient d of segment s
  __host__ __device__ int start_row(int q) const { return q; }                                    // q = derivative order 0..2 at the start
  __host__ __device__ int junc_row(int i, int q) const { return 3 + i * (8 + J) + 8 + q; }        // q: 0 waypoint, 1 pos, 2 vel, 3 acc, 4 jerk, 5 snap
  __host__ __device__ int end_row(int q) const { return 3 + (K - 1) * (8 + J) + 8 + q; }          // q = derivative order 0..2 at the end
};
__host__ __device__ inline int band_order(int K, int cont) { return 6 + 8 * K + (K - 1) * (cont + 2); }
// doubles of scratch one problem needs: the band matrix, the right-hand sides / solution (n x 3) and the pivots
__host__ __device__ inline size_t band_scratch_doubles(int K, int cont) {
  const size_t n = (size_t)band_order(K, cont);
  return n * BD_W + 3 * n + (n + 1) / 2 + 2;
}

// Time allocation (avgTimeAllocation, polyTrajSolver.cpp:125-138) + assembly, by one warp: A (band storage, zeroed here),
// R (n x 3) = right-hand sides of the three axes, dt[K] = segment durations, times[nwp] = knots.
__device__ inline void band_build(const double* wp, int nwp, const double* bc, double desired_vel, int cont, double* times,
                                  double* A, double* R, double* dt, int lane) {
  const int K = nwp - 1;
  const BandLayout L(K, cont);
  const int n = L.n;
  if (lane == 0) {
    double tt = 0.0;
    times[0] = 0.0;
    for (int i = 1; i < nwp; ++i) {
      const double dx = wp[3 * i] - wp[3 * i - 3], dy = wp[3 * i + 1] - wp[3 * i - 2], dz = wp[3 * i + 2] - wp[3 * i - 1];
      const double dur = sqrt(dx * dx + dy * dy + dz * dz) / desired_vel;
      dt[i - 1] = dur;
      tt += dur;
      times[i] = tt;
    }
  }
  for (size_t e = lane; e < (size_t)n * BD_W; e += 32) A[e] = 0.0;
  for (int e = lane; e < 3 * n; e += 32) R[e] = 0.0;
  __syncwarp();
  // P: snap Gram matrix on normalised time (constructP)
  for (int e = lane; e < K * 16; e += 32) {
    const int s = e / 16, i = 4 + (e % 16) / 4, j = 4 + (e % 4);
    double f = 1.0;
    for (int d = 0; d < 4; ++d) f *= (double)((i - d) * (j - d));
    f /= (double)(i + j - 7);
    A[bd_idx(L.var(s, i), L.var(s, j))] = f;
  }
  // derivative row of segment `seg` at normalised time 0 (at1 = false) or 1: value c_d t^(d - order) * scale, also transposed
  auto put = [&](int row, int seg, bool at1, int order, double scale, double sign) {
    for (int d = order; d < 8; ++d) {
      if (!at1 && d != order) break;
      double c = 1.0;
      for (int k = 0; k < order; ++k) c *= (double)(d - k);
      const double v = sign * c * scale;
      const int col = L.var(seg, d);
      A[bd_idx(row, col)] += v;
      A[bd_idx(col, row)] += v;
    }
  };
  // one lane per junction; lane 0 also writes the start rows, the lane of the last junction (or lane 0) the end rows
  for (int i = lane; i < K - 1; i += 32) {
    int r = L.junc_row(i, 0);
    put(r, i, true, 0, 1.0, 1.0);
    for (int a = 0; a < 3; ++a) R[3 * r + a] = wp[3 * (i + 1) + a];
    r = L.junc_row(i, 1);
    put(r, i, true, 0, 1.0, 1.0);
    put(r, i + 1, false, 0, 1.0, -1.0);
    for (int order = 1; order <= cont; ++order) {
      double sl = 1.0, sr = 1.0;
      for (int k = 0; k < order; ++k) { sl *= dt[i + 1]; sr *= dt[i]; }
      r = L.junc_row(i, 1 + order);
      put(r, i, true, order, sl, 1.0);
      put(r, i + 1, false, order, sr, -1.0);
    }
  }
  if (lane == 0) {
    for (int q = 0; q < 3; ++q) {
      put(L.start_row(q), 0, false, q, 1.0, 1.0);
      put(L.end_row(q), K - 1, true, q, 1.0, 1.0);
    }
    for (int a = 0; a < 3; ++a) {
      R[3 * L.start_row(0) + a] = wp[a];
      R[3 * L.end_row(0) + a] = wp[3 * K + a];
      if (bc) {   // v0, v1, a0, a1
        R[3 * L.start_row(1) + a] = bc[a];
        R[3 * L.end_row(1) + a] = bc[3 + a];
        R[3 * L.start_row(2) + a] = bc[6 + a];
        R[3 * L.end_row(2) + a] = bc[9 + a];
      }
    }
  }
  __syncwarp();
}

// PA = LU of the band matrix in place, by one warp (dgbtf2: multipliers below the diagonal, not swapped).  Returns (to
// every lane) 0 or -1 (singular).  guard = true (interior-point Newton systems, whose barrier terms reach 1e19 next to
// entries of order 1 in the last iterations): a column that has cancelled to exactly zero gets the pivot 1e300 instead —
// the unknown of that column is left unchanged by this solve (x_k = b_k / 1e300 = 0); the iteration's own residual test,
// computed from the unfactorised matrix, decides what the step was worth.
__device__ inline int band_factor(double* A, int n, int* piv, int lane, bool guard = false) {
  for (int k = 0; k < n; ++k) {
    // pivot: largest |A[r][k]| among rows k .. k + KL
    int r = k + lane;
    double best = (lane <= BD_KL && r < n) ? fabs(A[bd_idx(r, k)]) : -1.0;
    int bi = r;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {   // the candidates sit in lanes 0..13
      const double ov = __shfl_down_sync(BD_FULL, best, o);
      const int oi = __shfl_down_sync(BD_FULL, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    best = __shfl_sync(BD_FULL, best, 0);
    const int p = __shfl_sync(BD_FULL, bi, 0);
    if (!(best > 1e-300)) {
      if (!guard || best != best) return -1;
      if (lane == 0) { piv[k] = k; A[bd_idx(k, k)] = 1e300; }
      __syncwarp();
      continue;   // nothing to eliminate: the column is zero below the diagonal
    }
    if (lane == 0) piv[k] = p;
    const int jmax = k + BD_KU < n - 1 ? k + BD_KU : n - 1;
    if (p != k) {
      const int j = k + lane;
      if (j <= jmax) {
        const double a = A[bd_idx(k, j)];
        A[bd_idx(k, j)] = A[bd_idx(p, j)];
        A[bd_idx(p, j)] = a;
      }
      __syncwarp();
    }
    const double inv = 1.0 / A[bd_idx(k, k)];
    const int imax = k + BD_KL < n - 1 ? k + BD_KL : n - 1;
    const int nr = imax - k, nc = jmax - k;   // rows / columns to update
    for (int e = lane; e < nr * nc; e += 32) {
      const int i = k + 1 + e / nc, j = k + 1 + e % nc;
      const double mik = A[bd_idx(i, k)];
      if (mik != 0.0) A[bd_idx(i, j)] -= (mik * inv) * A[bd_idx(k, j)];
    }
    __syncwarp();
    if (lane < nr) A[bd_idx(k + 1 + lane, k)] *= inv;
    __syncwarp();
  }
  return 0;
}

// X (n x nrhs, row-major) <- A^-1 X with the factors of band_factor, by one warp: column-oriented forward and backward
// substitution (no reductions: step k scatters x_k into the <= 13 / <= 26 rows of the band).
__device__ inline void band_solve(const double* A, int n, const int* piv, double* X, int nrhs, int lane) {
  for (int k = 0; k < n; ++k) {
    const int p = piv[k];
    if (p != k && lane < nrhs) {
      const double a = X[(size_t)k * nrhs + lane];
      X[(size_t)k * nrhs + lane] = X[(size_t)p * nrhs + lane];
      X[(size_t)p * nrhs + lane] = a;
    }
    __syncwarp();
    const int nr = (k + BD_KL < n - 1 ? k + BD_KL : n - 1) - k;
    for (int e = lane; e < nr * nrhs; e += 32) {
      const int i = k + 1 + e / nrhs, c = e % nrhs;
      const double l = A[bd_idx(i, k)];
      if (l != 0.0) X[(size_t)i * nrhs + c] -= l * X[(size_t)k * nrhs + c];
    }
    __syncwarp();
  }
  for (int i = n - 1; i >= 0; --i) {
    if (lane < nrhs) X[(size_t)i * nrhs + lane] /= A[bd_idx(i, i)];
    __syncwarp();
    const int r0 = i - BD_KU > 0 ? i - BD_KU : 0, nr = i - r0;
    for (int e = lane; e < nr * nrhs; e += 32) {
      const int r = r0 + e / nrhs, c = e % nrhs;
      X[(size_t)r * nrhs + c] -= A[bd_idx(r, i)] * X[(size_t)i * nrhs + c];
    }
    __syncwarp();
  }
}

// One min-snap solve by ONE WARP: waypoints wp[nwp] (+ boundary conditions bc[12] = v0, v1, a0, a1 or null) -> knots
// times[nwp], coefficients coef[3][8 K] (axis-major, real time: de-normalised as polyTrajSolver.cpp:874-878).  `scratch` =
// band_scratch_doubles(K, cont) doubles, dt = K doubles (shared or global).  Returns (to every lane) 0 ok, -1 singular KKT,
// -2 too many segments, -3 fewer than two waypoints.
__device__ inline int band_minsnap_solve(const double* wp, int nwp, const double* bc, double desired_vel, int cont, double* coef,
                                         double* times, double* scratch, double* dt, int max_seg, int lane) {
  const int K = nwp - 1;
  if (K < 1 || K > max_seg) return K < 1 ? -3 : -2;
  const BandLayout L(K, cont);
  const int n = L.n;
  double* A = scratch;
  double* R = A + (size_t)n * BD_W;
  int* piv = reinterpret_cast<int*>(R + 3 * (size_t)n);
  band_build(wp, nwp, bc, desired_vel, cont, times, A, R, dt, lane);
  if (band_factor(A, n, piv, lane) != 0) return -1;
  band_solve(A, n, piv, R, 3, lane);
  const int nvar = 8 * K;
  for (int e = lane; e < 3 * nvar; e += 32) {
    const int a = e / nvar, q = e - a * nvar, s = q / 8, d = q - s * 8;
    coef[(size_t)a * nvar + q] = R[3 * (size_t)L.var(s, d) + a] / pow(dt[s], (double)d);
  }
  __syncwarp();
  return 0;
}
