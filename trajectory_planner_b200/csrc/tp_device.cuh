// Device-side building blocks shared by the ViGO kernels: bit-packed map gathers, de Boor
// evaluation, deterministic reductions.  Everything here is compiled with --fmad=false: voxel
// indices and collision decisions have to match the CPU path bit for bit (the reference is built
// without FMA contraction, CMakeLists.txt:6), so the operation order below is the reference's.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define TP_DEGREE 3  // bsplineDegree, bsplineTraj.h:19

// Flat, read-only, bit-packed occupancy grid in HBM: z fastest, `wz` uint32 words per (x,y)
// column (one word holds a whole column when nz <= 32, e.g. the 400x400x30 grid of the
// reference's occupancy_map.yaml -> 640 KB, L2 resident).
struct DevMap {
  const uint32_t* __restrict__ inflated;
  const uint32_t* __restrict__ known;
  double res;
  double inv_res;   // 1 / res: candidate quotient, confirmed or replaced by the true division (dm_cell)
  double mn[3];
  int dim[3];
  int wz;
};

struct D3 {
  double x, y, z;
};
__device__ __forceinline__ D3 d3(double x, double y, double z) {
  D3 r;
  r.x = x; r.y = y; r.z = z;
  return r;
}
__device__ __forceinline__ D3 operator+(const D3& a, const D3& b) { return d3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ D3 operator-(const D3& a, const D3& b) { return d3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ D3 operator*(double s, const D3& a) { return d3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ D3 operator*(const D3& a, double s) { return d3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ D3 operator/(const D3& a, double s) { return d3(a.x / s, a.y / s, a.z / s); }
// fixed-size reductions in the order Eigen 3.3 produces for Vector3d: (x+y)+z
__device__ __forceinline__ double dot3(const D3& a, const D3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ double norm3(const D3& a) { return sqrt(dot3(a, a)); }
__device__ __forceinline__ D3 cross3(const D3& a, const D3& b) {
  return d3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// floor(d / res) with the TRUE division's result (bit-exact voxel index) at the price of a multiplication: d * (1/res)
// is within 2 ulp of d / res, so the two floors can differ only when the quotient is within ~1e-12 of an integer;
// anything closer than 1e-6 to an integer (and every non-finite / huge value) takes the division.
__device__ __forceinline__ double dm_cell(double d, double res, double inv_res) {
  const double t = d * inv_res;
  if (!(fabs(t - rint(t)) >= 1e-6)) return floor(d / res);
  return floor(t);
}
// posToIndex = floor((p - origin)/res) (true division); false when outside the grid
__device__ __forceinline__ bool dm_index(const DevMap& m, double x, double y, double z, int& ix, int& iy, int& iz) {
  const double fx = dm_cell(x - m.mn[0], m.res, m.inv_res);
  const double fy = dm_cell(y - m.mn[1], m.res, m.inv_res);
  const double fz = dm_cell(z - m.mn[2], m.res, m.inv_res);
  // the comparisons are false for NaN -> outside
  const bool in = (fx >= 0.0) && (fx < (double)m.dim[0]) && (fy >= 0.0) && (fy < (double)m.dim[1]) && (fz >= 0.0) &&
                  (fz < (double)m.dim[2]);
  if (in) {
    ix = (int)fx;
    iy = (int)fy;
    iz = (int)fz;
  }
  return in;
}
// occMap::isInflatedOccupied: outside counts as occupied
__device__ __forceinline__ bool dm_inflated(const DevMap& m, double x, double y, double z) {
  int ix, iy, iz;
  if (!dm_index(m, x, y, z, ix, iy, iz)) return true;
  const uint32_t w = __ldg(&m.inflated[((size_t)ix * m.dim[1] + iy) * m.wz + (iz >> 5)]);
  return (w >> (iz & 31)) & 1u;
}
__device__ __forceinline__ bool dm_inflated(const DevMap& m, const D3& p) { return dm_inflated(m, p.x, p.y, p.z); }
// occMap::isUnknown: outside counts as unknown
__device__ __forceinline__ bool dm_unknown(const DevMap& m, const D3& p) {
  int ix, iy, iz;
  if (!dm_index(m, p.x, p.y, p.z, ix, iy, iz)) return true;
  const uint32_t w = __ldg(&m.known[((size_t)ix * m.dim[1] + iy) * m.wz + (iz >> 5)]);
  return !((w >> (iz & 31)) & 1u);
}
// occMap::isInflatedOccupiedLine (contract in include/tp_b200.h)
__device__ __forceinline__ bool dm_line(const DevMap& m, const D3& a, const D3& b) {
  if (dm_inflated(m, a) || dm_inflated(m, b)) return true;
  const D3 diff = b - a;
  const double dist = norm3(diff);
  const D3 unit = diff / dist;
  const int steps = (int)(dist / m.res);
  const D3 inc = unit * m.res;
  for (int i = 1; i < steps; ++i) {
    const D3 pc = a + (double)i * inc;
    if (dm_inflated(m, pc)) return true;
  }
  return false;
}

// bspline::at (bspline.cpp:32-58) for the uniform knot vector u_i = (i - degree)*ts of
// bspline.cpp:19-28, control points read through `cp` (3 doubles per point).  `n` = number of
// control points of THIS spline, `degree` its degree.  Same clamp, span search result and
// de Boor operation order as the reference.
template <class CP>
__device__ __forceinline__ D3 bspline_at(const CP& cp, int n, int degree, double ts, double t) {
  const double duration = (double)(n - degree) * ts;
  const double tb = fmin(fmax(0.0, t), duration);
  // smallest k >= degree with u_{k+1} >= tb  (linear scan in the reference, :37-42)
  int k = degree + (int)(tb / ts) - 1;
  if (k < degree) k = degree;
  while ((double)(k + 1 - degree) * ts < tb) ++k;
  while (k > degree && (double)(k - degree) * ts >= tb) --k;
  D3 d[4];
#pragma unroll
  for (int i = 0; i <= TP_DEGREE; ++i)
    if (i <= degree) d[i] = cp(k - degree + i);
  for (int r = 1; r <= degree; ++r) {
    for (int i = degree; i >= r; --i) {
      const double ka = (double)(i + k - degree - degree) * ts;
      const double kb = (double)(i + 1 + k - r - degree) * ts;
      const double alpha = (tb - ka) / (kb - ka);
      d[i] = (1 - alpha) * d[i - 1] + alpha * d[i];
    }
  }
  return d[degree];
}

// ------------------------------------------------------------------ reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// pow(x,3) of glibc is correctly rounded in all but ~1e-4 of cases; reproduce that with an
// error-free product (the explicit fma is exempt from --fmad=false)
__device__ __forceinline__ double cube_cr(double x) {
  const double hi = x * x;
  const double lo = __fma_rn(x, x, -hi);
  const double h2 = hi * x;
  const double e2 = __fma_rn(hi, x, -h2);
  const double l2 = lo * x + e2;
  return h2 + l2;
}

// Deterministic atan2 for the guide-point search (utils.h:84-86 calls std::atan2).  libm's atan2
// is neither bit-portable across platforms nor available on the device, and ULP-level differences
// in guide points are amplified by the ill-conditioned solve; this routine uses only IEEE
// +,-,*,/ in a fixed order (no FMA), so the CPU restatement in the oracle's `soft_atan2` mode
// reproduces it bit for bit.  Accuracy ~1-2 ULP.  Reduction: q = min/max in [0,1];
// q > tan(pi/8): atan q = pi/4 + atan((q-1)/(q+1)); series in z = t^2, |t| <= 0.4143, 24 terms.
__host__ __device__ inline double tp_atan_series(double t) {
  const double z = t * t;
  double s = 1.0 / 47.0;
  for (int k = 22; k >= 0; --k) s = 1.0 / (double)(2 * k + 1) - z * s;
  return t * s;
}
__host__ __device__ inline double tp_atan2(double y, double x) {
  const double PI = 3.14159265358979323846, PI_2 = 1.57079632679489661923, PI_4 = 0.78539816339744830962;
  const double ay = y < 0 ? -y : y, ax = x < 0 ? -x : x;
  if (ax == 0.0 && ay == 0.0) return 0.0;
  const bool swap = ay > ax;
  const double q = swap ? ax / ay : ay / ax;
  double r;
  if (q > 0.41421356237309503) r = PI_4 + tp_atan_series((q - 1.0) / (q + 1.0));
  else r = tp_atan_series(q);
  if (swap) r = PI_2 - r;
  if (x < 0) r = PI - r;
  if (y < 0) r = -r;
  return r;
}
