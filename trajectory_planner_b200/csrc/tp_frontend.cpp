// Host-side front end of the ViGO solve: what src/bspline_node.cpp:332-371 does between the two
// RViz clicks and bsplineTraj::makePlan().  Stays on the host (SURVEY.md §8f-2 marks a device
// version as a later row); it produces the control points that are the input of the GPU path.
//
//   seed:   1-segment min-snap polynomial start->goal, rest to rest
//           (polyTrajOccMap::makePlan(false), polyTrajOccMap.cpp:326-399; QP of
//           polyTrajSolver.cpp:241-271 (P), :314-584 (A), :587-813 (bounds), time allocation
//           :125-138, de-normalisation :870-879).  The reference hands the QP to OSQP
//           (eps 1e-3); we solve the same equality-constrained QP exactly via its KKT system.
//   sample: polyTrajOccMap::getTrajectory(dt) (polyTrajOccMap.cpp:434-446) + polyTrajSolver::getPos
//           (polyTrajSolver.cpp:1051-1071)
//   check:  bsplineTraj::inputPathCheck (bsplineTraj.cpp:207-245), dt *= 0.8 loop of
//           src/bspline_node.cpp:355-366 (its 0.05 s wall clock -> a 60-iteration cap)
//   update: bsplineTraj::updatePath (bsplineTraj.cpp:290-323): goal check, adjustPathLengthDirect
//           (:754-793; the function-static prevPathLength is an explicit 0.0 per problem), fillPath
//           (:247-288), bspline::parameterizeToBspline (bspline.cpp:74-138).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/tp_b200.h"
#include "tp_map.h"

namespace {

struct P3 {
  double x, y, z;
};
inline P3 sub(const P3& a, const P3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline double norm(const P3& a) { return std::sqrt((a.x * a.x + a.y * a.y) + a.z * a.z); }

// dense solve A x = b (n x n, row-major) with partial pivoting; returns false if singular
bool solve_dense(std::vector<double>& A, std::vector<double>& b, int n) {
  for (int c = 0; c < n; ++c) {
    int piv = c;
    double best = std::fabs(A[(size_t)c * n + c]);
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(A[(size_t)r * n + c]) > best) {
        best = std::fabs(A[(size_t)r * n + c]);
        piv = r;
      }
    if (best < 1e-300) return false;
    if (piv != c) {
      for (int k = 0; k < n; ++k) std::swap(A[(size_t)piv * n + k], A[(size_t)c * n + k]);
      std::swap(b[piv], b[c]);
    }
    for (int r = c + 1; r < n; ++r) {
      double f = A[(size_t)r * n + c] / A[(size_t)c * n + c];
      if (f == 0) continue;
      for (int k = c; k < n; ++k) A[(size_t)r * n + k] -= f * A[(size_t)c * n + k];
      b[r] -= f * b[c];
    }
  }
  for (int r = n - 1; r >= 0; --r) {
    double s = b[r];
    for (int k = r + 1; k < n; ++k) s -= A[(size_t)r * n + k] * b[k];
    b[r] = s / A[(size_t)r * n + r];
  }
  return true;
}

// one-segment min-snap, degree 7, normalised time; rows: p(0), p(1), v(0), v(1), a(0), a(1)
// (constructA order for K = 1).  Endpoint velocity/acceleration bounds are NOT scaled by the
// duration in the reference (constructBound); with the rest-to-rest conditions used here that is
// immaterial.  coef[3][8] returned in real time (c_d /= T^d).
bool seed_minsnap(const P3& s, const P3& g, double desired_vel, double coef[3][8], double& T) {
  const int n = 8, m = 6, diff = 4;
  T = norm(sub(g, s)) / desired_vel;
  double P[8][8];
  std::memset(P, 0, sizeof(P));
  for (int i = diff; i < n; ++i)
    for (int j = diff; j < n; ++j) {
      double f = 1.0;
      for (int d = 0; d < diff; ++d) f *= (double)((i - d) * (j - d));
      f /= (double)(i + j - diff * 2 + 1);
      P[i][j] = f;
    }
  double A[6][8];
  std::memset(A, 0, sizeof(A));
  A[0][0] = 1.0;
  for (int d = 0; d < n; ++d) A[1][d] = 1.0;
  A[2][1] = 1.0;
  for (int d = 1; d < n; ++d) A[3][d] = (double)d;
  A[4][2] = 2.0;
  for (int d = 2; d < n; ++d) A[5][d] = (double)(d * (d - 1));
  const double rhs[3][6] = {{s.x, g.x, 0, 0, 0, 0}, {s.y, g.y, 0, 0, 0, 0}, {s.z, g.z, 0, 0, 0, 0}};
  for (int ax = 0; ax < 3; ++ax) {
    const int N = n + m;
    std::vector<double> K((size_t)N * N, 0.0), b(N, 0.0);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) K[(size_t)i * N + j] = P[i][j];
    for (int r = 0; r < m; ++r)
      for (int j = 0; j < n; ++j) {
        K[(size_t)(n + r) * N + j] = A[r][j];
        K[(size_t)j * N + n + r] = A[r][j];
      }
    for (int r = 0; r < m; ++r) b[n + r] = rhs[ax][r];
    if (!solve_dense(K, b, N)) return false;
    for (int d = 0; d < n; ++d) coef[ax][d] = b[d] / std::pow(T, d);
  }
  return true;
}

P3 poly_pos(const double coef[3][8], double t) {
  double x = 0, y = 0, z = 0;
  for (int d = 0; d < 8; ++d) {
    const double pw = std::pow(t, d);
    x += coef[0][d] * pw;
    y += coef[1][d] * pw;
    z += coef[2][d] * pw;
  }
  return {x, y, z};
}

bool line_occ(const tp_map* m, const P3& a, const P3& b) {
  const double pa[3] = {a.x, a.y, a.z}, pb[3] = {b.x, b.y, b.z};
  return m->is_inflated_occupied_line(pa, pb);
}

// bsplineTraj::adjustPathLengthDirect with prevPathLength = 0
void adjust_path_length_direct(const tp_map* m, const std::vector<P3>& path, double max_len, std::vector<P3>& out) {
  out.clear();
  bool exceed = false;
  double min_len = 0.0;
  const P3 p0 = path[0];
  for (size_t i = 0; i + 1 < path.size(); ++i) {
    const P3 p1 = path[i], p2 = path[i + 1];
    const double total = norm(sub(p2, p0));
    if (total >= std::max(0.0, max_len)) exceed = true;
    out.push_back(p1);
    if (exceed) {
      const bool free_ = !line_occ(m, p1, p2);
      if (free_ && min_len >= 1.5) {
        out.push_back(p2);
        return;
      }
    }
    if (line_occ(m, p1, p2))
      min_len = 0.0;
    else
      min_len += norm(sub(p2, p1));
  }
  out.push_back(path.back());
}

// bsplineTraj::inputPathCheck
bool input_path_check(const tp_map* m, const std::vector<P3>& path, double cp_dist, double max_len,
                      std::vector<P3>& adjusted) {
  if (path.empty()) return true;
  std::vector<P3> adj;
  adjust_path_length_direct(m, path, max_len, adj);
  for (size_t i = 0; i + 1 < adj.size(); ++i)
    if (norm(sub(adj[i], adj[i + 1])) > cp_dist * 1.5) return false;
  adjusted.clear();
  P3 prev = adj[0];
  for (size_t i = 0; i < adj.size(); ++i) {
    if (i == 0) {
      adjusted.push_back(adj[i]);
      prev = adj[i];
    } else if (norm(sub(adj[i], prev)) >= cp_dist * 0.8) {
      adjusted.push_back(adj[i]);
      prev = adj[i];
    }
  }
  adjusted.push_back(adjusted.back());
  return true;
}

bool fill_path(const std::vector<P3>& p, std::vector<P3>& out) {
  if (p.size() <= 1) return false;
  if (p.size() == 2) {
    const P3 d = sub(p[1], p[0]);
    out = {p[0],
           {d.x / 3.0 + p[0].x, d.y / 3.0 + p[0].y, d.z / 3.0 + p[0].z},
           {2.0 * d.x / 3.0 + p[0].x, 2.0 * d.y / 3.0 + p[0].y, 2.0 * d.z / 3.0 + p[0].z},
           p[1]};
  } else if (p.size() == 3) {
    out = {p[0],
           {(p[0].x + p[1].x) / 2.0, (p[0].y + p[1].y) / 2.0, (p[0].z + p[1].z) / 2.0},
           p[1],
           {(p[1].x + p[2].x) / 2.0, (p[1].y + p[2].y) / 2.0, (p[1].z + p[2].z) / 2.0},
           p[2]};
  } else
    out = p;
  return true;
}

// least squares min |A x - b| for 3 right-hand sides via Householder QR (A: rows x cols, row-major)
void lstsq3(std::vector<double>& A, std::vector<double>& B, int rows, int cols, std::vector<double>& X) {
  for (int c = 0; c < cols; ++c) {
    double nrm = 0;
    for (int r = c; r < rows; ++r) nrm += A[(size_t)r * cols + c] * A[(size_t)r * cols + c];
    nrm = std::sqrt(nrm);
    if (nrm == 0) continue;
    const double alpha = A[(size_t)c * cols + c] > 0 ? -nrm : nrm;
    std::vector<double> v(rows - c);
    for (int r = c; r < rows; ++r) v[r - c] = A[(size_t)r * cols + c];
    v[0] -= alpha;
    double vn = 0;
    for (double t : v) vn += t * t;
    if (vn == 0) continue;
    for (int k = c; k < cols; ++k) {
      double s = 0;
      for (int r = c; r < rows; ++r) s += v[r - c] * A[(size_t)r * cols + k];
      s = 2 * s / vn;
      for (int r = c; r < rows; ++r) A[(size_t)r * cols + k] -= s * v[r - c];
    }
    for (int k = 0; k < 3; ++k) {
      double s = 0;
      for (int r = c; r < rows; ++r) s += v[r - c] * B[(size_t)r * 3 + k];
      s = 2 * s / vn;
      for (int r = c; r < rows; ++r) B[(size_t)r * 3 + k] -= s * v[r - c];
    }
  }
  X.assign((size_t)cols * 3, 0.0);
  for (int k = 0; k < 3; ++k)
    for (int r = cols - 1; r >= 0; --r) {
      double s = B[(size_t)r * 3 + k];
      for (int j = r + 1; j < cols; ++j) s -= A[(size_t)r * cols + j] * X[(size_t)j * 3 + k];
      X[(size_t)r * 3 + k] = s / A[(size_t)r * cols + r];
    }
}

// bspline::parameterizeToBspline
void fit_bspline(double ts, const std::vector<P3>& pts, const double se[12], std::vector<double>& ctrl) {
  const int K = (int)pts.size(), rows = K + 4, cols = K + 2;
  std::vector<double> A((size_t)rows * cols, 0.0), B((size_t)rows * 3, 0.0);
  const double pr[3] = {1, 4, 1}, vr[3] = {-1, 0, 1}, ar[3] = {1, -2, 1};
  for (int i = 0; i < K; ++i)
    for (int k = 0; k < 3; ++k) A[(size_t)i * cols + i + k] = (1 / 6.0) * pr[k];
  for (int k = 0; k < 3; ++k) {
    A[(size_t)K * cols + k] = (1 / 2.0 / ts) * vr[k];
    A[(size_t)(K + 1) * cols + K - 1 + k] = (1 / 2.0 / ts) * vr[k];
    A[(size_t)(K + 2) * cols + k] = (1 / ts / ts) * ar[k];
    A[(size_t)(K + 3) * cols + K - 1 + k] = (1 / ts / ts) * ar[k];
  }
  for (int i = 0; i < K; ++i) {
    B[(size_t)i * 3] = pts[i].x;
    B[(size_t)i * 3 + 1] = pts[i].y;
    B[(size_t)i * 3 + 2] = pts[i].z;
  }
  for (int i = 0; i < 4; ++i)
    for (int k = 0; k < 3; ++k) B[(size_t)(K + i) * 3 + k] = se[3 * i + k];
  lstsq3(A, B, rows, cols, ctrl);
}

// bspline::at for a degree-p uniform spline given as packed control points
void bspline_at(const std::vector<double>& cp, int n, int degree, double ts, double t, double out[3]) {
  const double duration = (double)((n - 1 + degree + 1 + 1) - degree - 1 - degree) * ts;
  const double tb = std::min(std::max(0.0, t), duration);
  int k = degree;
  while ((double)(k + 1 - degree) * ts < tb) ++k;
  double d[4][3];
  for (int i = 0; i <= degree; ++i)
    for (int a = 0; a < 3; ++a) d[i][a] = cp[(size_t)3 * (k - degree + i) + a];
  for (int r = 1; r <= degree; ++r)
    for (int i = degree; i >= r; --i) {
      const double ka = (double)(i + k - degree - degree) * ts;
      const double kb = (double)(i + 1 + k - r - degree) * ts;
      const double alpha = (tb - ka) / (kb - ka);
      for (int a = 0; a < 3; ++a) d[i][a] = (1 - alpha) * d[i - 1][a] + alpha * d[i][a];
    }
  for (int a = 0; a < 3; ++a) out[a] = d[degree][a];
}

}  // namespace

extern "C" {

int tp_bspline_fit(double ts, int32_t K, const double* points, const double* start_end4, double* ctrl_out) {
  if (!(ts > 0) || K < 4 || !points || !start_end4 || !ctrl_out) {
    tp_set_error("tp_bspline_fit: need ts > 0, >= 4 points and 4 boundary vectors (bspline.cpp:78-93)");
    return TP_ERR_INVALID_ARG;
  }
  std::vector<P3> pts(K);
  for (int i = 0; i < K; ++i) pts[i] = {points[3 * i], points[3 * i + 1], points[3 * i + 2]};
  std::vector<double> c;
  fit_bspline(ts, pts, start_end4, c);
  std::memcpy(ctrl_out, c.data(), sizeof(double) * 3 * (K + 2));
  return TP_OK;
}

int tp_bspline_eval(int32_t N, const double* ctrl, double ts, int32_t deriv, int32_t nt, const double* t, double* out) {
  if (N < 4 || !ctrl || !(ts > 0) || deriv < 0 || deriv > 2 || (nt > 0 && (!t || !out))) return TP_ERR_INVALID_ARG;
  std::vector<double> cp(ctrl, ctrl + 3 * (size_t)N);
  int n = N, degree = 3;
  for (int k = 0; k < deriv; ++k) {  // bspline::getDerivative
    std::vector<double> q((size_t)3 * (n - 1));
    for (int i = 0; i < n - 1; ++i) {
      const double den = (double)(i + degree + 1 - degree) * ts - (double)(i + 1 - degree) * ts;
      for (int a = 0; a < 3; ++a) q[3 * i + a] = ((double)degree * (cp[3 * (i + 1) + a] - cp[3 * i + a])) / den;
    }
    cp.swap(q);
    --n;
    --degree;
  }
  for (int i = 0; i < nt; ++i) bspline_at(cp, n, degree, ts, t[i], out + 3 * i);
  return TP_OK;
}


// ------------------------------------------------------------------------------------------------------------------
// pwlTraj (piecewiseLinearTraj.cpp): the fallback polyTrajOctomap / polyTrajOccMap return when no valid polynomial
// trajectory was found (polyTrajOctomap.cpp:309-317).  Serial host logic in the reference and here (a handful of
// waypoints); the yaw that the reference carries through a quaternion is carried as the angle, with
// quaternion_from_rpy's wrap (utils.h:45-47: yaw > PI_const -> yaw - 2 PI_const) applied where the reference converts.
static const double kPIc = 3.1415926;   // utils.h:19
static double pwl_yaw_distance(double y1, double y2) {   // utils.h:74-82
  double delta = std::fabs(y2 - y1);
  if (delta > kPIc) delta = 2 * kPIc - delta;
  return delta;
}
static double pwl_wrap(double yaw) { return yaw > kPIc ? yaw - 2 * kPIc : yaw; }

int tp_pwl_plan(int32_t K, const double* path, const double* yaw_in, double desired_vel, double desired_ang_vel, double* yaw_out,
                double* times_out) {
  if (K < 2 || !path || !yaw_out || !times_out || !(desired_vel > 0) || !(desired_ang_vel > 0)) {
    tp_set_error("tp_pwl_plan: need >= 2 waypoints and positive velocities");
    return TP_ERR_INVALID_ARG;
  }
  const bool use_yaw = yaw_in != nullptr;
  if (use_yaw) {
    for (int i = 0; i < K; ++i) yaw_out[i] = yaw_in[i];
  } else {   // pwlTraj::updatePath (:31-46): heading of each segment, the last point repeats the previous heading
    double yaw = 0.0;
    for (int i = 0; i < K - 1; ++i) {
      yaw = std::atan2(path[3 * (i + 1) + 1] - path[3 * i + 1], path[3 * (i + 1)] - path[3 * i]);
      yaw_out[i] = yaw;
    }
    yaw_out[K - 1] = yaw;
  }
  // avgTimeAllocation (:82-119): per waypoint a rotation period (zero for the first) then a forward period
  double total = 0.0;
  int n = 0;
  for (int i = 0; i < K - 1; ++i) {
    if (i != 0) total += pwl_yaw_distance(yaw_out[i - 1], yaw_out[i]) / desired_ang_vel;
    else total += 0.0;
    times_out[n++] = total;
    const double dx = path[3 * i] - path[3 * (i + 1)], dy = path[3 * i + 1] - path[3 * (i + 1) + 1], dz = path[3 * i + 2] - path[3 * (i + 1) + 2];
    total += std::sqrt(std::pow(dx, 2) + std::pow(dy, 2) + std::pow(dz, 2)) / desired_vel;
    times_out[n++] = total;
  }
  if (use_yaw) {
    total += pwl_yaw_distance(yaw_out[K - 2], yaw_out[K - 1]) / desired_ang_vel;
    times_out[n++] = total;
  }
  return n;
}

int tp_pwl_eval(int32_t K, const double* path, const double* yaw, int32_t n_times, const double* times, int32_t nt, const double* t,
                double* out) {
  if (K < 2 || !path || !yaw || n_times < 2 || !times || nt < 0 || (nt > 0 && (!t || !out))) return TP_ERR_INVALID_ARG;
  for (int q = 0; q < nt; ++q) {   // pwlTraj::getPose (:199-268)
    const double tq = t[q];
    double* o = out + 4 * q;
    o[0] = o[1] = o[2] = o[3] = 0.0;
    if (tq >= times[n_times - 1]) {
      o[0] = path[3 * (K - 1)]; o[1] = path[3 * (K - 1) + 1]; o[2] = path[3 * (K - 1) + 2];
      o[3] = pwl_wrap(yaw[K - 1]);
      continue;
    }
    for (int i = 0; i < n_times - 1; ++i) {
      const double t0 = times[i], t1 = times[i + 1];
      if (!(tq >= t0 && tq <= t1)) continue;
      if (i % 2 == 1) {   // rotation period: at waypoint pointIdx + 1, turning from yaw[pointIdx] to yaw[pointIdx + 1]
        const int pi = (i - 1) / 2;
        const double yd = yaw[pi + 1] - yaw[pi];
        double dir = 1.0, yda = std::fabs(yd);
        if (yda <= kPIc && yd >= 0) dir = 1.0;
        else if (yda <= kPIc && yd < 0) dir = -1.0;
        else if (yda > kPIc && yd >= 0) { dir = -1.0; yda = 2 * kPIc - yda; }
        else if (yda > kPIc && yd < 0) { dir = 1.0; yda = 2 * kPIc - yda; }
        o[0] = path[3 * (pi + 1)]; o[1] = path[3 * (pi + 1) + 1]; o[2] = path[3 * (pi + 1) + 2];
        o[3] = pwl_wrap(yaw[pi] + dir * (tq - t0) / (t1 - t0) * yda);
      } else {            // forward period from waypoint pointIdx to pointIdx + 1
        const int pi = i / 2;
        const double* a = path + 3 * pi;
        const double* b = path + 3 * (pi + 1);
        if (t1 - t0 < 1e-3) {
          o[0] = a[0]; o[1] = a[1]; o[2] = a[2];
        } else {
          o[0] = a[0] + (tq - t0) * (b[0] - a[0]) / (t1 - t0);
          o[1] = a[1] + (tq - t0) * (b[1] - a[1]) / (t1 - t0);
          o[2] = a[2] + (tq - t0) * (b[2] - a[2]) / (t1 - t0);
        }
        o[3] = pwl_wrap(yaw[pi]);
      }
      break;
    }
  }
  return TP_OK;
}

int tp_poly_eval(int32_t K, const double* coef, const double* times, int32_t nt, const double* t, double* out) {
  if (K < 1 || !coef || !times || nt < 0 || (nt > 0 && (!t || !out))) return TP_ERR_INVALID_ARG;
  for (int q = 0; q < nt; ++q) {   // polyTrajSolver::getPose (polyTrajSolver.cpp:1026-1049)
    double tq = t[q];
    double* o = out + 4 * q;
    o[0] = o[1] = o[2] = o[3] = 0.0;
    for (int i = 0; i < K; ++i) {
      if (!(tq >= times[i] && tq <= times[i + 1])) continue;
      tq = tq - times[i];
      const double* cx = coef + 8 * i;
      const double* cy = coef + 8 * (size_t)K + 8 * i;
      const double* cz = coef + 16 * (size_t)K + 8 * i;
      double x = 0, y = 0, z = 0;
      for (int d = 0; d < 8; ++d) {
        const double pw = std::pow(tq, d);
        x += cx[d] * pw; y += cy[d] * pw; z += cz[d] * pw;
      }
      if (tq == 0) tq = 0.01;
      double dx = 0, dy = 0;
      for (int d = 0; d < 8; ++d) {
        const double pw = std::pow(tq, d - 1);
        dx += d * cx[d] * pw; dy += d * cy[d] * pw;
      }
      o[0] = x; o[1] = y; o[2] = z; o[3] = std::atan2(dy, dx);
      break;
    }
  }
  return TP_OK;
}

int64_t tp_vigo_frontend_batch(const tp_map_t* m, const tp_vigo_params* p, int32_t B, const double* starts,
                               const double* goals, int32_t* offsets_out, double* ctrl_out, int64_t ctrl_cap,
                               uint8_t* valid) {
  if (!m || !p || B < 0 || !starts || !goals || !offsets_out || !ctrl_out) return TP_ERR_INVALID_ARG;
  int64_t total = 0;
  offsets_out[0] = 0;
  const double zeros[12] = {0};
  for (int b = 0; b < B; ++b) {
    const P3 s = {starts[3 * b], starts[3 * b + 1], starts[3 * b + 2]};
    const P3 g = {goals[3 * b], goals[3 * b + 1], goals[3 * b + 2]};
    bool ok = true;
    std::vector<double> ctrl;
    double coef[3][8], T = 0;
    if (!(norm(sub(g, s)) > 0) || !seed_minsnap(s, g, p->max_vel, coef, T)) ok = false;
    std::vector<P3> adjusted, best;
    if (ok) {
      double dt = p->ctrl_pt_dist / p->max_vel;  // bsplineTraj::getInitTs
      bool have = false;
      for (int it = 0; it < 60; ++it) {
        std::vector<P3> traj;
        for (double t = 0; t <= T; t += dt) traj.push_back(poly_pos(coef, t));
        std::vector<P3> adj;
        const bool sat = input_path_check(m, traj, p->ctrl_pt_dist, p->max_path_length, adj);
        if (sat) {
          best = adj;
          have = true;
          break;
        }
        dt *= 0.8;
      }
      if (!have || best.empty()) ok = false;
    }
    if (ok) {
      const P3 goal = best.back();
      if (m->is_inflated_occupied(goal.x, goal.y, goal.z)) ok = false;  // updatePath :291-295
    }
    if (ok) {
      std::vector<P3> inp;
      adjust_path_length_direct(m, best, p->max_path_length, inp);
      if (inp.size() < 4) ok = fill_path(best, inp);
      if (ok) fit_bspline(p->ctrl_pt_ts, inp, zeros, ctrl);
    }
    const int64_t n = ok ? (int64_t)ctrl.size() / 3 : 0;
    if (total + n > ctrl_cap) {
      tp_set_error("tp_vigo_frontend_batch: ctrl_cap too small");
      return TP_ERR_CAPACITY;
    }
    if (n) std::memcpy(ctrl_out + 3 * total, ctrl.data(), sizeof(double) * 3 * n);
    total += n;
    offsets_out[b + 1] = (int32_t)total;
    if (valid) valid[b] = ok ? 1 : 0;
  }
  return total;
}

// bsplineTraj::inputPathCheck (bsplineTraj.cpp:207-245) for one path
int tp_vigo_input_path_check(const tp_map_t* m, const tp_vigo_params* p, int32_t K, const double* path, double* adjusted,
                             int32_t cap, int32_t* n_adjusted) {
  if (!m || !p || K < 0 || (K > 0 && !path) || !n_adjusted) return TP_ERR_INVALID_ARG;
  std::vector<P3> in((size_t)K), adj;
  for (int i = 0; i < K; ++i) in[i] = {path[3 * i], path[3 * i + 1], path[3 * i + 2]};
  const bool sat = input_path_check(m, in, p->ctrl_pt_dist, p->max_path_length, adj);
  *n_adjusted = (int32_t)adj.size();
  if (adjusted) {
    if ((int)adj.size() > cap) {
      tp_set_error("tp_vigo_input_path_check: capacity %d < %zu points", cap, adj.size());
      return TP_ERR_CAPACITY;
    }
    for (size_t i = 0; i < adj.size(); ++i) { adjusted[3 * i] = adj[i].x; adjusted[3 * i + 1] = adj[i].y; adjusted[3 * i + 2] = adj[i].z; }
  }
  return sat ? 1 : 0;
}

// bsplineTraj::updatePath (bsplineTraj.cpp:290-323) for one path: goal check, adjustPathLengthDirect, fillPath,
// parameterizeToBspline.  Returns the number of control points written, 0 when the reference returns false.
int tp_vigo_update_path(const tp_map_t* m, const tp_vigo_params* p, int32_t K, const double* path, const double* start_end4,
                        double* ctrl_out, int32_t cap) {
  if (!m || !p || K <= 0 || !path || !ctrl_out) return TP_ERR_INVALID_ARG;
  std::vector<P3> in((size_t)K), inp;
  for (int i = 0; i < K; ++i) in[i] = {path[3 * i], path[3 * i + 1], path[3 * i + 2]};
  const P3 goal = in.back();
  if (m->is_inflated_occupied(goal.x, goal.y, goal.z)) return 0;   // :291-295
  adjust_path_length_direct(m, in, p->max_path_length, inp);
  if (inp.size() < 4 && !fill_path(in, inp)) return 0;
  const double zeros[12] = {0};
  std::vector<double> ctrl;
  fit_bspline(p->ctrl_pt_ts, inp, start_end4 ? start_end4 : zeros, ctrl);
  const int n = (int)ctrl.size() / 3;
  if (n > cap) {
    tp_set_error("tp_vigo_update_path: capacity %d < %d control points", cap, n);
    return TP_ERR_CAPACITY;
  }
  std::memcpy(ctrl_out, ctrl.data(), sizeof(double) * 3 * n);
  return n;
}

}  // extern "C"
