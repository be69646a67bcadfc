// Secondary path: batched min-snap polynomial solve + trajectory sampling + box collision check + the
// "insert a midpoint waypoint into every colliding segment and re-solve" loop of
// polyTrajOctomap::makePlanAddingWaypoint (polyTrajOctomap.cpp:259-321).  Included at the end of tp_vigo.cu
// (same translation unit: shares tp_engine).
//
//  k_minsnap_solve : the QP of polyTrajSolver (P: polyTrajSolver.cpp:241-271, equality rows :314-584, bounds
//                    :587-813, time allocation :125-138, de-normalisation :870-879) solved EXACTLY through its
//                    KKT system [P A^T; A 0][x; lambda] = [0; b] (the reference hands it to OSQP, eps 1e-3): ordered
//                    segment by segment the system is banded (half-bandwidth 13), and ONE WARP per problem runs a band
//                    LU with partial pivoting (tp_band.cuh); the three axes share the factorisation.
//  k_poly_check    : polyTrajSolver::getTrajectory (:1125-1137, accumulated t += delT, position = sum c_d pow(t,d))
//                    + polyTrajOctomap::checkCollisionTraj / checkCollision / checkCollisionPoint
//                    (polyTrajOctomap.cpp:634-656, 547-590): per sample a box of points, each "collision" when
//                    outside the known bounding box, unknown, or occupied; colliding samples mark their segment.
#pragma once
#include "tp_band.cuh"

#define PL_THREADS 128
#define PL_DEG 7
#define PL_NC (PL_DEG + 1)
#define PL_MAX_SEG 63          // segments per path the solver accepts (n = 14 K <= 882)

struct PolyMap {   // the 3-state grid of the polyTraj collision contract
  const uint32_t* __restrict__ occ;
  const uint32_t* __restrict__ known;
  double res;
  double mn[3];
  int dim[3];
  int wz;
  double bbmin[3], bbmax[3];   // metric bounding box of the known cells (octomap getMetricMin/Max stand-in)
};

// polyTrajOctomap::checkCollisionPoint with ignoreUnknown = false
__device__ __forceinline__ bool pm_collision_point(const PolyMap& m, double x, double y, double z) {
  if (x < m.bbmin[0] || x > m.bbmax[0] || y < m.bbmin[1] || y > m.bbmax[1] || z < m.bbmin[2] || z > m.bbmax[2]) return true;
  const double fx = floor((x - m.mn[0]) / m.res), fy = floor((y - m.mn[1]) / m.res), fz = floor((z - m.mn[2]) / m.res);
  if (!(fx >= 0.0 && fx < (double)m.dim[0] && fy >= 0.0 && fy < (double)m.dim[1] && fz >= 0.0 && fz < (double)m.dim[2])) return true;
  const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
  const size_t w = ((size_t)ix * m.dim[1] + iy) * m.wz + (iz >> 5);
  if (!((__ldg(&m.known[w]) >> (iz & 31)) & 1u)) return true;
  return (__ldg(&m.occ[w]) >> (iz & 31)) & 1u;
}
// polyTrajOctomap::checkCollision: p is converted to float (pose2Octomap), the box corners are doubles computed
// from the float coordinates, every box point goes back through float (octomap::point3d)
__device__ __forceinline__ bool pm_collision_box(const PolyMap& m, double px, double py, double pz, const double box[3],
                                                 double map_res) {
  const double fx = (double)(float)px, fy = (double)(float)py, fz = (double)(float)pz;
  const double xmin = fx - box[0] / 2, xmax = fx + box[0] / 2;
  const double ymin = fy - box[1] / 2, ymax = fy + box[1] / 2;
  const double zmin = fz - box[2] / 2, zmax = fz + box[2] / 2;
  const int xNum = (int)((xmax - xmin) / map_res), yNum = (int)((ymax - ymin) / map_res), zNum = (int)((zmax - zmin) / map_res);
  for (int a = 0; a <= xNum; ++a)
    for (int b = 0; b <= yNum; ++b)
      for (int c = 0; c <= zNum; ++c) {
        const double qx = (double)(float)(xmin + a * map_res), qy = (double)(float)(ymin + b * map_res),
                     qz = (double)(float)(zmin + c * map_res);
        if (pm_collision_point(m, qx, qy, qz)) return true;
      }
  return false;
}

// ------------------------------------------------------------------------------------------- solve
struct PolySolveArgs {
  int B;
  const int* wp_off;      // [B+1] waypoint offsets
  const double* wp;       // [3 * total waypoints]
  const double* bc;       // [B * 12] v0, v1, a0, a1 per problem (may be null -> zeros)
  double desired_vel;
  int cont;               // continuity degree 2..4
  double* coef;           // out: problem b, axis a, segment s, power d at 24*(wp_off[b]-b) + a*8*K + 8*s + d
  double* times;          // out: [total waypoints] time knots
  int* status;            // out: 0 ok, -1 singular, -2 too many segments
  double* scratch;        // [workers * stride]
  size_t stride;          // doubles per worker: poly_scratch_doubles(longest path)
  int* queue;
};

// doubles of scratch one min-snap solve needs for paths of up to kmax segments (band matrix, right-hand sides, pivots, dt)
__host__ __device__ inline size_t poly_scratch_doubles(int kmax) { return band_scratch_doubles(kmax, 4) + (size_t)kmax + 8; }

// One min-snap solve inside a block-per-path kernel: warp 0 runs the banded solver (tp_band.cuh), the block waits.
// waypoints wp[nwp] (+ boundary conditions bc[12] = v0, v1, a0, a1 or null) -> knots times[nwp], coefficients coef[3][8 K]
// (axis-major).  `scratch` = poly_scratch_doubles(K) doubles.  Returns (to every thread) 0 ok, -1 singular KKT, -2 too many
// segments, -3 fewer than two waypoints.
__device__ int poly_solve_one(const double* wp, int nwp, const double* bc, double desired_vel, int cont, double* coef,
                              double* times, double* scratch) {
  __shared__ int s_rc;
  if (threadIdx.x < 32) {
    const int K = nwp - 1;
    double* dt = scratch + ((K >= 1 && K <= PL_MAX_SEG) ? band_scratch_doubles(K, cont) : 0);
    const int rc = band_minsnap_solve(wp, nwp, bc, desired_vel, cont, coef, times, scratch, dt, PL_MAX_SEG, threadIdx.x);
    if (threadIdx.x == 0) s_rc = rc;
  }
  __syncthreads();
  const int rc = s_rc;
  __syncthreads();
  return rc;
}

__global__ void __launch_bounds__(PL_THREADS) k_minsnap_solve(PolySolveArgs A) {
  // ONE WARP per problem: four independent workers per block, problems from a shared queue, no block barrier
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* scratch = A.scratch + ((size_t)blockIdx.x * (PL_THREADS / 32) + warp) * A.stride;
  for (;;) {
    int b = 0;
    if (lane == 0) b = atomicAdd(A.queue, 1);
    b = __shfl_sync(0xffffffffu, b, 0);
    if (b >= A.B) break;
    const int w0 = A.wp_off[b], nwp = A.wp_off[b + 1] - w0, K = nwp - 1;
    double* dt = scratch + ((K >= 1 && K <= PL_MAX_SEG) ? band_scratch_doubles(K, A.cont) : 0);
    const int st = band_minsnap_solve(A.wp + 3 * (size_t)w0, nwp, A.bc ? A.bc + 12 * (size_t)b : nullptr, A.desired_vel, A.cont,
                                      A.coef + (size_t)3 * PL_NC * (w0 - b), A.times + w0, scratch, dt, PL_MAX_SEG, lane);
    if (lane == 0) A.status[b] = st;
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------- sample + check
struct PolyCheckArgs {
  int B;
  const int* wp_off;
  const double* wp;
  const double* coef;
  const double* times;
  const double* t_acc;     // accumulated sample times: t_acc[0] = 0, t_acc[s+1] = t_acc[s] + delT
  int n_t_acc;
  double box[3];
  double map_res;
  uint8_t* valid;          // [B] 1 = no colliding sample
  uint8_t* seg_hit;        // [total segments] at wp_off[b] - b + i
  int* n_samples;          // [B] trajectory entries (polynomial samples + the appended last waypoint)
  double* samples;         // optional [B * samp_cap * 3] positions (parity entry), else null
  uint8_t* sample_hit;     // optional [B * samp_cap]
  int samp_cap;
};

__device__ __forceinline__ void poly_eval(const double* coef, const double* knots, int K, double t, double out[3]) {
  // polyTrajSolver::getPose: first segment with knot[i] <= t <= knot[i+1]; position = sum_d c_d * pow(t - knot[i], d)
  out[0] = out[1] = out[2] = 0.0;
  for (int i = 0; i < K; ++i) {
    if (t >= knots[i] && t <= knots[i + 1]) {
      const double tt = t - knots[i];
      for (int a = 0; a < 3; ++a) {
        const double* c = coef + (size_t)a * PL_NC * K + PL_NC * i;
        double x = 0.0;
        for (int d = 0; d < PL_NC; ++d) x += c[d] * pow(tt, (double)d);
        out[a] = x;
      }
      return;
    }
  }
}

// getTrajectory + checkCollisionTraj for one path by the whole block: samples t_acc[s] < T plus the appended last
// waypoint; marks seg_hit[K]; returns (to every thread) whether any sample collides; *n_traj = trajectory entries.
__device__ int poly_check_one(const PolyMap& map, const double* wp, int K, const double* coef, const double* knots,
                              const double* t_acc, int n_t_acc, const double box[3], double map_res, uint8_t* seg_hit,
                              int* n_traj, double* samples, uint8_t* sample_hit, int samp_cap, const DevMap* occmap = nullptr) {
  const int tid = threadIdx.x;
  const double T = knots[K];
  for (int i = tid; i < K; i += PL_THREADS) seg_hit[i] = 0;
  __syncthreads();
  // number of polynomial samples: t_acc[s] < T
  int lo = 0, hi = n_t_acc;   // first s with t_acc[s] >= T
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (t_acc[mid] < T) lo = mid + 1; else hi = mid;
  }
  const int npoly = lo, ntraj = npoly + 1;
  int any = 0;
  for (int s = tid; s < ntraj; s += PL_THREADS) {
    double p[3];
    if (s < npoly) poly_eval(coef, knots, K, t_acc[s], p);
    else { p[0] = wp[3 * (size_t)K]; p[1] = wp[3 * (size_t)K + 1]; p[2] = wp[3 * (size_t)K + 2]; }
    // polyTrajOctomap: the collision box on the 3-state grid; polyTrajOccMap (occmap != null): the sample itself, colliding when
    // it is inflated-occupied AND unknown (polyTrajOccMap.cpp:531 — the reference's conjunction, preserved)
    const bool hit = occmap ? (dm_inflated(*occmap, d3(p[0], p[1], p[2])) && dm_unknown(*occmap, d3(p[0], p[1], p[2])))
                            : pm_collision_box(map, p[0], p[1], p[2], box, map_res);
    if (samples && s < samp_cap) {
      double* o = samples + (size_t)s * 3;
      o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
      sample_hit[s] = hit ? 1 : 0;
    }
    if (hit) {
      any = 1;
      const double t = s < n_t_acc ? t_acc[s] : 1e300;   // checkCollisionTraj's own accumulated t (:639,654)
      for (int i = 0; i < K; ++i)
        if (t >= knots[i] && t <= knots[i + 1]) { seg_hit[i] = 1; break; }
    }
  }
  any = __syncthreads_or(any);
  *n_traj = ntraj;
  return any;
}

__global__ void __launch_bounds__(PL_THREADS) k_poly_check(PolyCheckArgs A, PolyMap map) {
  const int b = blockIdx.x, tid = threadIdx.x;
  const int w0 = A.wp_off[b], nwp = A.wp_off[b + 1] - w0, K = nwp - 1;
  if (K < 1) {
    if (tid == 0) { A.valid[b] = 1; A.n_samples[b] = 0; }
    return;
  }
  int ntraj = 0;
  const int any = poly_check_one(map, A.wp + 3 * (size_t)w0, K, A.coef + (size_t)3 * PL_NC * (w0 - b), A.times + w0, A.t_acc, A.n_t_acc,
                                 A.box, A.map_res, A.seg_hit + (w0 - b), &ntraj, A.samples ? A.samples + (size_t)b * A.samp_cap * 3 : nullptr,
                                 A.sample_hit ? A.sample_hit + (size_t)b * A.samp_cap : nullptr, A.samp_cap);
  if (tid == 0) { A.valid[b] = any ? 0 : 1; A.n_samples[b] = ntraj; }
}

// ------------------------------------------------------------------------------------------- the whole loop on the device
// polyTrajOctomap::makePlanAddingWaypoint (polyTrajOctomap.cpp:259-321) for one path per thread block, start to finish:
// solve -> sample + box check -> insert a midpoint waypoint into every colliding segment (highest index first, :179-185)
// -> re-solve, until the trajectory is collision free, maxIter is exceeded or the waypoint cap is reached.  No host round
// trip per iteration (round 1: one solve launch + one check launch + two copies + a host-side insertion per iteration).
// Outputs go to fixed-stride staging rows (cap waypoints per path); a scan + gather pass packs them.
struct PolyLoopArgs {
  int B;
  const int* wp_off;       // [B+1]
  const double* wp;        // [3 * total]
  const double* bc;        // [B * 12] or null
  double desired_vel;
  int cont, max_iter, cap; // cap = waypoints per path the staging rows hold (<= PL_MAX_SEG + 1)
  const double* t_acc;
  int n_t_acc;
  double box[3];
  double map_res;
  double* wp_st;           // [B][cap][3]
  double* coef_st;         // [B][3 * 8 * (cap - 1)]
  double* times_st;        // [B][cap]
  int* n_wp;               // [B]
  uint8_t* valid;          // [B]
  int* iters;              // [B]
  double* scratch;         // [grid * stride]
  size_t stride;           // doubles per block: poly_scratch_doubles(cap - 1)
  int* queue;
};

__global__ void __launch_bounds__(PL_THREADS) k_polytraj_loop(PolyLoopArgs A, PolyMap map) {
  __shared__ int s_b, s_n, s_go, s_it;
  __shared__ uint8_t s_seg[PL_MAX_SEG + 1];
  const int tid = threadIdx.x;
  double* scratch = A.scratch + (size_t)blockIdx.x * A.stride;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_b = atomicAdd(A.queue, 1);
    __syncthreads();
    const int b = s_b;
    if (b >= A.B) break;
    const int w0 = A.wp_off[b], nwp0 = A.wp_off[b + 1] - w0;
    double* wp = A.wp_st + (size_t)b * A.cap * 3;
    double* coef = A.coef_st + (size_t)b * 3 * PL_NC * (A.cap - 1);
    double* times = A.times_st + (size_t)b * A.cap;
    for (int e = tid; e < 3 * nwp0; e += PL_THREADS) wp[e] = A.wp[3 * (size_t)w0 + e];
    if (tid == 0) { s_n = nwp0; A.valid[b] = 0; A.iters[b] = 0; }
    __syncthreads();
    if (nwp0 < 2) {   // single-point path (:262-266)
      if (tid == 0) { A.valid[b] = 1; A.n_wp[b] = nwp0; }
      continue;
    }
    int it = 0;
    for (;;) {
      const int nwp = s_n, K = nwp - 1;
      const int st = poly_solve_one(wp, nwp, A.bc ? A.bc + 12 * (size_t)b : nullptr, A.desired_vel, A.cont, coef, times, scratch);
      ++it;
      int any = 1, ntraj = 0;
      if (st == 0) any = poly_check_one(map, wp, K, coef, times, A.t_acc, A.n_t_acc, A.box, A.map_res, s_seg, &ntraj, nullptr, nullptr, 0);
      else {
        for (int i = tid; i < K; i += PL_THREADS) s_seg[i] = 0;   // an unsolved path has no trajectory to check
        __syncthreads();
      }
      if (tid == 0) {
        int go = 0;
        s_it = it;
        if (st == 0 && !any) A.valid[b] = 1;
        else if (it <= A.max_iter) {   // ++countIter; if (countIter > maxIter_) break;  (:302-305)
          int add = 0;
          for (int i = 0; i < K; ++i) add += s_seg[i];
          if (add == 0) {
            // a colliding sample outside every knot interval (or an unsolvable path): nothing to insert, so every further
            // iteration would re-solve the same path until maxIter — jump there
            s_it = A.max_iter + 1;
          } else if (nwp + add <= A.cap) {
            // insertWaypoint: midpoints of the colliding segments, highest index first
            int n = nwp;
            for (int i = K - 1; i >= 0; --i)
              if (s_seg[i]) {
                for (int q = n - 1; q > i; --q)
                  for (int a = 0; a < 3; ++a) wp[3 * (q + 1) + a] = wp[3 * q + a];
                for (int a = 0; a < 3; ++a) wp[3 * (i + 1) + a] = (wp[3 * i + a] + wp[3 * (i + 2) + a]) / 2;
                ++n;
              }
            s_n = n;
            go = 1;
          }
        }
        s_go = go;
      }
      __syncthreads();
      if (!s_go) break;
    }
    if (tid == 0) { A.n_wp[b] = s_n; A.iters[b] = s_it; }
  }
}

// pack the staging rows: waypoints / knots by waypoint offset, coefficients by segment offset
__global__ void k_polytraj_pack(PolyLoopArgs A, const int* __restrict__ off_out, double* wp_out, double* coef_out, double* times_out,
                                long wp_cap) {
  const int b = blockIdx.x;
  if (b >= A.B) return;
  const int o = off_out[b], n = off_out[b + 1] - o;
  if ((long)o + n > wp_cap) return;
  const double* wp = A.wp_st + (size_t)b * A.cap * 3;
  const double* coef = A.coef_st + (size_t)b * 3 * PL_NC * (A.cap - 1);
  const double* times = A.times_st + (size_t)b * A.cap;
  for (int e = threadIdx.x; e < 3 * n; e += blockDim.x) wp_out[3 * (size_t)o + e] = wp[e];
  if (n >= 2) {
    for (int e = threadIdx.x; e < n; e += blockDim.x) times_out[o + e] = times[e];
    for (int e = threadIdx.x; e < 3 * PL_NC * (n - 1); e += blockDim.x) coef_out[(size_t)3 * PL_NC * (o - b) + e] = coef[e];
  }
}

// box collision check on caller-supplied positions (parity entry: decisions are bit-exact functions of the position)
__global__ void k_poly_box_points(PolyMap map, long n, const double* __restrict__ xyz, double bx, double by, double bz,
                                  double map_res, uint8_t* __restrict__ out) {
  const double box[3] = {bx, by, bz};
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    out[i] = pm_collision_box(map, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], box, map_res) ? 1 : 0;
}
