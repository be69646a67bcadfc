// ORACLE — TEST INFRASTRUCTURE ONLY (see tp_oracle.hpp for scope and parity-pinning status).
//
// CPU restatement of the reference ViGO solve.  Short names for citations:
//   bT.cpp = include/trajectory_planner/bsplineTraj.cpp     bT.h = .../bsplineTraj.h
//   bs.cpp = include/trajectory_planner/bspline.cpp
//   astar.cpp / astar.h = include/trajectory_planner/path_search/astarOcc.{cpp,h}
//   lbfgs.hpp = include/trajectory_planner/solver/lbfgs.hpp
//   utils.h = include/trajectory_planner/utils.h
#include "tp_oracle.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <queue>
#include <set>
#include <thread>
#include <utility>

#include "lbfgs_port.hpp"
#include "wform_port.hpp"
#ifdef TP_ORACLE_REF_LBFGS
// The reference's own header, straight from /root/reference (never copied into this repo).
#include "trajectory_planner/solver/lbfgs.hpp"
#endif

namespace orc {

// ------------------------------------------------------------------ small vector helpers
static inline V3 operator+(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator*(double s, const V3& a) { return {s * a.x, s * a.y, s * a.z}; }
static inline V3 operator*(const V3& a, double s) { return {a.x * s, a.y * s, a.z * s}; }
static inline V3 operator/(const V3& a, double s) { return {a.x / s, a.y / s, a.z / s}; }
static inline double dot3(const V3& a, const V3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline double sqnorm3(const V3& a) { return dot3(a, a); }
static inline double norm3(const V3& a) { return std::sqrt(sqnorm3(a)); }
static inline V3 cross3(const V3& a, const V3& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// utils.h:19 (sic: truncated pi) and utils.h:84-86
static const double PI_const = 3.1415926;
// Restatement of the product's deterministic atan2 (trajectory_planner_b200/csrc/tp_device.cuh,
// tp_atan2): IEEE +,-,*,/ only, fixed order, so both sides agree bit for bit.  Used only when
// VigoParams::soft_atan2 is set; the reference-faithful default is std::atan2 (utils.h:84-86).
static inline double soft_atan_series(double t) {
  const double z = t * t;
  double s = 1.0 / 47.0;
  for (int k = 22; k >= 0; --k) s = 1.0 / (double)(2 * k + 1) - z * s;
  return t * s;
}
static inline double soft_atan2(double y, double x) {
  const double PI = 3.14159265358979323846, PI_2 = 1.57079632679489661923, PI_4 = 0.78539816339744830962;
  const double ay = y < 0 ? -y : y, ax = x < 0 ? -x : x;
  if (ax == 0.0 && ay == 0.0) return 0.0;
  const bool swap = ay > ax;
  const double q = swap ? ax / ay : ay / ax;
  double r;
  if (q > 0.41421356237309503) r = PI_4 + soft_atan_series((q - 1.0) / (q + 1.0));
  else r = soft_atan_series(q);
  if (swap) r = PI_2 - r;
  if (x < 0) r = PI - r;
  if (y < 0) r = -r;
  return r;
}
static inline double angleBetweenVectors(const V3& a, const V3& b, bool soft) {
  const double y = norm3(cross3(a, b)), x = dot3(a, b);
  return soft ? soft_atan2(y, x) : std::atan2(y, x);
}

// ------------------------------------------------------------------ map contract
void OccMap::init(double res_, const double mn_[3], const int dim_[3], const int inf_[3]) {
  res = res_;
  for (int a = 0; a < 3; ++a) { mn[a] = mn_[a]; dim[a] = dim_[a]; inf[a] = inf_[a]; }
  size_t n = (size_t)dim[0] * dim[1] * dim[2];
  occ.assign(n, 0);
  known.assign(n, 0);
  inflated.assign(n, 0);
}
// posToIndex = floor((pos - mapSizeMin)/res) per axis; inside iff 0 <= idx < dim
bool OccMap::index_of(const V3& p, int idx[3]) const {
  const double q[3] = {p.x, p.y, p.z};
  bool in = true;
  for (int a = 0; a < 3; ++a) {
    double f = std::floor((q[a] - mn[a]) / res);
    // guard the int conversion for far-away / non-finite coordinates
    if (!(f >= -1.0)) { idx[a] = -1; in = false; }
    else if (f > 2147483000.0) { idx[a] = 2147483000; in = false; }
    else {
      idx[a] = (int)f;
      if (idx[a] < 0 || idx[a] >= dim[a]) in = false;
    }
  }
  return in;
}
void OccMap::add_occupied_cell(int ix, int iy, int iz) {
  if (ix < 0 || iy < 0 || iz < 0 || ix >= dim[0] || iy >= dim[1] || iz >= dim[2]) return;
  occ[addr(ix, iy, iz)] = 1;
  known[addr(ix, iy, iz)] = 1;
  for (int dx = -inf[0]; dx <= inf[0]; ++dx)
    for (int dy = -inf[1]; dy <= inf[1]; ++dy)
      for (int dz = -inf[2]; dz <= inf[2]; ++dz) {
        int jx = ix + dx, jy = iy + dy, jz = iz + dz;
        if (jx < 0 || jy < 0 || jz < 0 || jx >= dim[0] || jy >= dim[1] || jz >= dim[2]) continue;
        inflated[addr(jx, jy, jz)] = 1;
      }
}
void OccMap::add_free_cell(int ix, int iy, int iz) {
  if (ix < 0 || iy < 0 || iz < 0 || ix >= dim[0] || iy >= dim[1] || iz >= dim[2]) return;
  known[addr(ix, iy, iz)] = 1;
}
void OccMap::add_occupied_point(const V3& p) {
  int idx[3];
  if (!index_of(p, idx)) return;
  add_occupied_cell(idx[0], idx[1], idx[2]);
}
bool OccMap::isInflatedOccupied(const V3& p) const {
  int idx[3];
  if (!index_of(p, idx)) return true;  // outside the map counts as occupied
  return inflated[addr(idx[0], idx[1], idx[2])] != 0;
}
bool OccMap::isUnknown(const V3& p) const {
  int idx[3];
  if (!index_of(p, idx)) return true;
  return known[addr(idx[0], idx[1], idx[2])] == 0;
}
bool OccMap::isInflatedOccupiedLine(const V3& a, const V3& b) const {
  if (isInflatedOccupied(a) || isInflatedOccupied(b)) return true;
  V3 diff = b - a;
  double dist = norm3(diff);
  V3 unit = diff / dist;
  int stepNum = (int)(dist / res);
  V3 inc = unit * res;
  for (int i = 1; i < stepNum; ++i) {
    V3 pc = a + (double)i * inc;
    if (isInflatedOccupied(pc)) return true;
  }
  return false;
}

// ------------------------------------------------------------------ uniform B-spline (bs.cpp)
struct Bspline {
  int degree = 3;
  int n = 0;                 // number of control points
  double ts = 0.2;
  std::vector<V3> cp;
  std::vector<double> knots;
  double duration = 0;
  Bspline() {}
  Bspline(int deg, const std::vector<V3>& c, double ts_) : degree(deg), n((int)c.size()), ts(ts_), cp(c) {
    initKnots();
  }
  // bs.cpp:19-28
  void initKnots() {
    int knotsNum = n - 1 + degree + 1 + 1;
    knots.resize(knotsNum);
    for (int i = 0; i < knotsNum; ++i) knots[i] = (i - degree) * ts;
    duration = knots[knotsNum - degree - 1];
  }
  // bs.cpp:32-58 (de Boor)
  V3 at(double t) const {
    double tb = std::min(std::max(0.0, t), duration);
    int k = degree;
    while (true) {
      if (knots[k + 1] >= tb) break;
      ++k;
    }
    V3 d[8];
    for (int i = 0; i <= degree; ++i) d[i] = cp[k - degree + i];
    for (int r = 1; r <= degree; ++r) {
      for (int i = degree; i >= r; --i) {
        double alpha = (tb - knots[i + k - degree]) / (knots[i + 1 + k - r] - knots[i + k - degree]);
        d[i] = (1 - alpha) * d[i - 1] + alpha * d[i];
      }
    }
    return d[degree];
  }
  // bs.cpp:64-72
  Bspline derivative() const {
    std::vector<V3> q(n - 1);
    for (int i = 0; i < n - 1; ++i)
      q[i] = ((double)degree * (cp[i + 1] - cp[i])) / (knots[i + degree + 1] - knots[i + 1]);
    return Bspline(degree - 1, q, ts);
  }
};

// ------------------------------------------------------------------ A* (astar.cpp / astar.h)
struct GridNode {
  int rounds = 0;
  int state = 3;          // UNDEFINED (astar.h:17-26)
  int idx[3] = {0, 0, 0};
  double g = 0, f = 0;    // "inf = 1 >> 20" == 0 (astar.h:13,29)
  GridNode* cameFrom = nullptr;
};
struct NodeCmp {  // astar.h:33-38
  bool operator()(GridNode* a, GridNode* b) const { return a->f > b->f; }
};
struct AStar {
  const OccMap* map = nullptr;
  int POOL[3] = {0, 0, 0}, CENTER[3] = {0, 0, 0};
  double minH = 0, maxH = 3;
  std::vector<GridNode> pool;
  int rounds = 0;
  double step = 0.1, inv_step = 10;
  V3 center{0, 0, 0};
  const double tie_breaker = 1.0 + 1.0 / 10000;  // astar.h:65
  std::vector<GridNode*> gridPath;
  int max_expansions = 0;
  long total_expansions = 0;
  int last_expansions = 0;

  // astar.cpp:16-37
  void init(const OccMap* m, const int pool_size[3], double minHeight, double maxHeight) {
    map = m;
    for (int a = 0; a < 3; ++a) { POOL[a] = pool_size[a]; CENTER[a] = pool_size[a] / 2; }
    minH = minHeight;
    maxH = maxHeight;
    pool.assign((size_t)POOL[0] * POOL[1] * POOL[2], GridNode());
    rounds = 0;
  }
  GridNode* node(int i, int j, int k) { return &pool[((size_t)i * POOL[1] + j) * POOL[2] + k]; }
  // astar.h:89-92
  V3 Index2Coord(const int idx[3]) const {
    return {(double)(idx[0] - CENTER[0]) * step + center.x, (double)(idx[1] - CENTER[1]) * step + center.y,
            (double)(idx[2] - CENTER[2]) * step + center.z};
  }
  // astar.h:94-105 (cast<int>() truncates toward zero)
  bool Coord2Index(const V3& pt, int idx[3]) const {
    idx[0] = (int)((pt.x - center.x) * inv_step + 0.5) + CENTER[0];
    idx[1] = (int)((pt.y - center.y) * inv_step + 0.5) + CENTER[1];
    idx[2] = (int)((pt.z - center.z) * inv_step + 0.5) + CENTER[2];
    if (idx[0] < 0 || idx[0] >= POOL[0] || idx[1] < 0 || idx[1] >= POOL[1] || idx[2] < 0 || idx[2] >= POOL[2])
      return false;
    return true;
  }
  bool checkOccupancy(const V3& p) const { return map->isInflatedOccupied(p); }  // astar.h:58
  // astar.cpp:39-63 (override-style ifs kept)
  double getDiagHeu(const GridNode* n1, const GridNode* n2) const {
    double dx = std::abs(n1->idx[0] - n2->idx[0]);
    double dy = std::abs(n1->idx[1] - n2->idx[1]);
    double dz = std::abs(n1->idx[2] - n2->idx[2]);
    double h = 0.0;
    int diag = (int)std::min(std::min(dx, dy), dz);
    dx -= diag;
    dy -= diag;
    dz -= diag;
    if (dx == 0) h = 1.0 * std::sqrt(3.0) * diag + std::sqrt(2.0) * std::min(dy, dz) + 1.0 * std::abs(dy - dz);
    if (dy == 0) h = 1.0 * std::sqrt(3.0) * diag + std::sqrt(2.0) * std::min(dx, dz) + 1.0 * std::abs(dx - dz);
    if (dz == 0) h = 1.0 * std::sqrt(3.0) * diag + std::sqrt(2.0) * std::min(dx, dy) + 1.0 * std::abs(dx - dy);
    return h;
  }
  double getHeu(const GridNode* a, const GridNode* b) const { return tie_breaker * getDiagHeu(a, b); }  // astar.h:84-87
  // astar.cpp:90-117
  bool adjustStartEnd(V3 start_pt, V3 end_pt, int start_idx[3], int end_idx[3]) {
    if (!Coord2Index(start_pt, start_idx) || !Coord2Index(end_pt, end_idx)) return false;
    if (checkOccupancy(Index2Coord(start_idx))) {
      do {
        V3 d = start_pt - end_pt;
        start_pt = (d / norm3(d)) * step + start_pt;
        if (!Coord2Index(start_pt, start_idx)) return false;
      } while (checkOccupancy(Index2Coord(start_idx)));
    }
    if (checkOccupancy(Index2Coord(end_idx))) {
      do {
        V3 d = end_pt - start_pt;
        end_pt = (d / norm3(d)) * step + end_pt;
        if (!Coord2Index(end_pt, end_idx)) return false;
      } while (checkOccupancy(Index2Coord(end_idx)));
    }
    return true;
  }
  // astar.cpp:119-244; the 0.2 s wall-clock exit (:231) is replaced by max_expansions
  bool search(double step_size, const V3& start_pt, const V3& end_pt) {
    ++rounds;
    last_expansions = 0;
    step = step_size;
    inv_step = 1 / step_size;
    center = (start_pt + end_pt) / 2;
    int start_idx[3], end_idx[3];
    if (!adjustStartEnd(start_pt, end_pt, start_idx, end_idx)) return false;
    GridNode* startPtr = node(start_idx[0], start_idx[1], start_idx[2]);
    GridNode* endPtr = node(end_idx[0], end_idx[1], end_idx[2]);
    std::priority_queue<GridNode*, std::vector<GridNode*>, NodeCmp> openSet;
    std::memcpy(startPtr->idx, start_idx, sizeof(start_idx));
    std::memcpy(endPtr->idx, end_idx, sizeof(end_idx));  // (set before use by getHeu; same values as :152)
    startPtr->rounds = rounds;
    startPtr->g = 0;
    startPtr->f = getHeu(startPtr, endPtr);
    startPtr->state = 1;
    startPtr->cameFrom = nullptr;
    openSet.push(startPtr);
    int num_iter = 0;
    while (!openSet.empty()) {
      num_iter++;
      GridNode* current = openSet.top();
      openSet.pop();
      if (current->idx[0] == endPtr->idx[0] && current->idx[1] == endPtr->idx[1] && current->idx[2] == endPtr->idx[2]) {
        gridPath.clear();
        gridPath.push_back(current);
        while (current->cameFrom != nullptr) {
          current = current->cameFrom;
          gridPath.push_back(current);
        }
        last_expansions = num_iter;
        total_expansions += num_iter;
        return true;
      }
      current->state = 2;
      for (int dx = -1; dx <= 1; dx++)
        for (int dy = -1; dy <= 1; dy++)
          for (int dz = -1; dz <= 1; dz++) {
            if (dx == 0 && dy == 0 && dz == 0) continue;
            int nb[3] = {current->idx[0] + dx, current->idx[1] + dy, current->idx[2] + dz};
            if (nb[0] < 1 || nb[0] >= POOL[0] - 1 || nb[1] < 1 || nb[1] >= POOL[1] - 1 || nb[2] < 1 || nb[2] >= POOL[2] - 1)
              continue;
            GridNode* nbp = node(nb[0], nb[1], nb[2]);
            std::memcpy(nbp->idx, nb, sizeof(nb));
            bool explored = nbp->rounds == rounds;
            if (explored && nbp->state == 2) continue;
            nbp->rounds = rounds;
            V3 c = Index2Coord(nbp->idx);
            if (c.z > maxH || c.z < minH) continue;
            if (checkOccupancy(c)) continue;
            double static_cost = std::sqrt((double)(dx * dx + dy * dy + dz * dz));
            double tentative = current->g + static_cost;
            if (!explored) {
              nbp->state = 1;
              nbp->cameFrom = current;
              nbp->g = tentative;
              nbp->f = tentative + getHeu(nbp, endPtr);
              openSet.push(nbp);
            } else if (tentative < nbp->g) {
              nbp->cameFrom = current;
              nbp->g = tentative;
              nbp->f = tentative + getHeu(nbp, endPtr);
            }
          }
      if (max_expansions > 0 && num_iter >= max_expansions) {
        last_expansions = num_iter;
        total_expansions += num_iter;
        return false;
      }
    }
    last_expansions = num_iter;
    total_expansions += num_iter;
    return false;
  }
  // astar.cpp:246-254
  std::vector<V3> getPath() const {
    std::vector<V3> path;
    for (auto p : gridPath) path.push_back(Index2Coord(p->idx));
    std::reverse(path.begin(), path.end());
    return path;
  }
};

// ------------------------------------------------------------------ the planner (bT.cpp / bT.h)
static const int bsplineDegree = 3;  // bT.h:19

struct Planner {
  const OccMap* map;
  VigoParams P;
  int N = 0;
  std::vector<double> ctrl;  // 3 x N column-major == optData_.controlPoints (bT.h:22)
  std::vector<std::vector<V3>> guideP, guideV;  // bT.h:23-24
  std::vector<V3> dynPos, dynVel, dynSize;      // bT.h:26-28
  std::vector<std::pair<int, int>> collisionSeg;
  std::vector<std::vector<V3>> astarPaths;
  double wDist, wDyn;  // weightDistance_, weightDynamicObstacle_ (mutated by the outer loop)
  AStar astar;
  PlanStats stats;
  double linearFactor = 1.0;
  std::vector<double> gD, gS, gF, gO;  // the four 3xN gradient matrices of bT.cpp:807-810

  Planner(const OccMap* m, const VigoParams& p) : map(m), P(p) {
    wDist = P.w_distance;
    wDyn = P.w_dyn;
    // bT.cpp:187-195 (setMap): pool = 2*int(max_obstacle_size/res) per axis
    int pool[3];
    for (int a = 0; a < 3; ++a) pool[a] = 2 * (int)(P.max_obstacle_size[a] / map->res);
    astar.init(map, pool, P.min_height, P.max_height);
    astar.max_expansions = P.astar_max_expansions;
    std::memset(&stats, 0, sizeof(stats));
  }
  V3 col(int i) const { return {ctrl[3 * i], ctrl[3 * i + 1], ctrl[3 * i + 2]}; }
  std::vector<V3> cols() const {
    std::vector<V3> c(N);
    for (int i = 0; i < N; ++i) c[i] = col(i);
    return c;
  }
  // bT.cpp:393-401 + the tail of updatePath :315-319
  void setControlPoints(const double* c, int n) {
    N = n;
    ctrl.assign(c, c + 3 * n);
    guideP.assign(N, {});
    guideV.assign(N, {});
    dynPos.clear(); dynVel.clear(); dynSize.clear();
    collisionSeg.clear();
    astarPaths.clear();
    wDist = P.w_distance;
    wDyn = P.w_dyn;
    std::memset(&stats, 0, sizeof(stats));
  }

  // ---- cost terms -------------------------------------------------------------
  // bT.cpp:823-932
  void getDistanceCost(double& cost, std::vector<double>& grad) const {
    cost = 0.0;
    double costTemp;
    V3 gradientTemp;
    const double dth = P.dthresh;
    double a = 3.0 * dth, b = -3.0 * (dth * dth), c = std::pow(dth, 3);
    const double hth = 0.2;
    double ah = 3.0 * hth, bh = -3 * (hth * hth), ch = std::pow(hth, 3);
    for (int i = bsplineDegree; i <= N - bsplineDegree - 1; ++i) {
      for (size_t j = 0; j < guideP[i].size(); ++j) {
        double dist = dot3(col(i) - guideP[i][j], guideV[i][j]);
        bool unknownGuidePoint = map->isUnknown(guideP[i][j]);
        double distErr = dth - dist;
        V3 g = guideV[i][j];
        if (distErr <= -1.0 * dth) {
          costTemp = std::pow(-distErr, 3);
          gradientTemp = (3.0 * ((-distErr) * (-distErr))) * g;
          if (!P.plan_in_z) gradientTemp.z = 0.0;
          cost += costTemp;
          grad[3 * i] += gradientTemp.x; grad[3 * i + 1] += gradientTemp.y; grad[3 * i + 2] += gradientTemp.z;
        } else if (distErr > 0 && distErr <= dth) {
          costTemp = std::pow(distErr, 3);
          gradientTemp = (-3.0 * (distErr * distErr)) * g;
          if (unknownGuidePoint) {
            costTemp *= P.uncertain_factor;
            gradientTemp = gradientTemp * P.uncertain_factor;
          }
          if (!P.plan_in_z) gradientTemp.z = 0.0;
          cost += costTemp;
          grad[3 * i] += gradientTemp.x; grad[3 * i + 1] += gradientTemp.y; grad[3 * i + 2] += gradientTemp.z;
        } else if (distErr >= dth) {
          costTemp = a * (distErr * distErr) + b * distErr + c;
          gradientTemp = (-(2 * a * distErr + b)) * g;
          if (unknownGuidePoint) {
            costTemp *= P.uncertain_factor;
            gradientTemp = gradientTemp * P.uncertain_factor;
          }
          if (!P.plan_in_z) gradientTemp.z = 0.0;
          cost += costTemp;
          grad[3 * i] += gradientTemp.x; grad[3 * i + 1] += gradientTemp.y; grad[3 * i + 2] += gradientTemp.z;
        }
      }
      if (P.plan_in_z) {
        // bT.cpp:897-930: height barrier whose gradient lands on the X row (reference quirk)
        double hmin = ctrl[3 * i + 2] - P.min_height;
        double hmax = ctrl[3 * i + 2] - P.max_height;
        if (hmin < 0) {
          double e = hth - hmin;
          cost += ah * (e * e) + bh * e + ch;
          grad[3 * i] += (-(2 * ah * e + bh)) * -1.0;
          grad[3 * i + 1] += (-(2 * ah * e + bh)) * 0.0;
          grad[3 * i + 2] += (-(2 * ah * e + bh)) * 0.0;
        } else if (hmin >= 0 && hmax < hth) {
          double e = hth - hmin;
          cost += std::pow(e, 3);
          grad[3 * i] += (-3.0 * (e * e)) * -1.0;
          grad[3 * i + 1] += (-3.0 * (e * e)) * 0.0;
          grad[3 * i + 2] += (-3.0 * (e * e)) * 0.0;
        }
        if (hmax > 0) {
          double e = hth + hmax;
          cost += ah * (e * e) + bh * e + ch;
          grad[3 * i] += (-(2 * ah * e + bh)) * 1.0;
          grad[3 * i + 1] += (-(2 * ah * e + bh)) * 0.0;
          grad[3 * i + 2] += (-(2 * ah * e + bh)) * 0.0;
        } else if (hmax <= 0 && hmax >= -hth) {
          double e = hth + hmax;
          cost += std::pow(e, 3);
          grad[3 * i] += (-3.0 * (e * e)) * 1.0;
          grad[3 * i + 1] += (-3.0 * (e * e)) * 0.0;
          grad[3 * i + 2] += (-3.0 * (e * e)) * 0.0;
        }
      }
    }
  }
  // bT.cpp:934-950
  void getSmoothnessCost(double& cost, std::vector<double>& grad) const {
    cost = 0.0;
    for (int i = 0; i < N - bsplineDegree; ++i) {
      V3 jerk = col(i + 3) - 3 * col(i + 2) + 3 * col(i + 1) - col(i);
      cost += sqnorm3(jerk);
      V3 gt = 2.0 * jerk;
      const double* g = &gt.x;
      for (int a = 0; a < 3; ++a) {
        grad[3 * i + a] += -g[a];
        grad[3 * (i + 1) + a] += 3.0 * g[a];
        grad[3 * (i + 2) + a] += -3.0 * g[a];
        grad[3 * (i + 3) + a] += g[a];
      }
    }
  }
  // bT.cpp:952-999  (maxVel = maxAcc = 1.0 hard-coded there, :955-956)
  void getFeasibilityCost(double& cost, std::vector<double>& grad) const {
    cost = 0.0;
    const double maxVel = 1.0, maxAcc = 1.0;
    const double cts = P.ctrl_pt_ts;
    const double tsInvSqr = 1 / (cts * cts);
    for (int i = 0; i < N - 1; ++i) {
      for (int j = 0; j < 3; ++j) {
        double vi = (ctrl[3 * (i + 1) + j] - ctrl[3 * i + j]) / cts;
        if (vi > maxVel) {
          cost += ((vi - maxVel) * (vi - maxVel)) * tsInvSqr;
          grad[3 * i + j] += -2 * (vi - maxVel) / cts * tsInvSqr;
          grad[3 * (i + 1) + j] += 2 * (vi - maxVel) / cts * tsInvSqr;
        } else if (vi < -maxVel) {
          cost += ((vi + maxVel) * (vi + maxVel)) * tsInvSqr;
          grad[3 * i + j] += -2 * (vi + maxVel) / cts * tsInvSqr;
          grad[3 * (i + 1) + j] += 2 * (vi + maxVel) / cts * tsInvSqr;
        }
      }
    }
    for (int i = 0; i < N - 2; ++i) {
      for (int j = 0; j < 3; ++j) {
        double ai = (ctrl[3 * (i + 2) + j] - 2 * ctrl[3 * (i + 1) + j] + ctrl[3 * i + j]) * tsInvSqr;
        if (ai > maxAcc) {
          cost += (ai - maxAcc) * (ai - maxAcc);
          grad[3 * i + j] += 2 * (ai - maxAcc) * tsInvSqr;
          grad[3 * (i + 1) + j] += -4 * (ai - maxAcc) * tsInvSqr;
          grad[3 * (i + 2) + j] += 2 * (ai - maxAcc) * tsInvSqr;
        } else if (ai < -maxAcc) {
          cost += (ai + maxAcc) * (ai + maxAcc);
          grad[3 * i + j] += 2 * (ai + maxAcc) * tsInvSqr;
          grad[3 * (i + 1) + j] += -4 * (ai + maxAcc) * tsInvSqr;
          grad[3 * (i + 2) + j] += 2 * (ai + maxAcc) * tsInvSqr;
        }
      }
    }
  }
  // bT.cpp:1001-1064
  void getDynamicObstacleCost(double& cost, std::vector<double>& grad) const {
    cost = 0;
    if (dynPos.size() == 0) return;
    const int skipFactor = 2;
    int predictionNum = (int)(P.pred_horizon / P.ts);
    const double dd = P.dthresh_dyn;
    double a = 3.0 * dd, b = -3 * (dd * dd), c = std::pow(dd, 3);
    for (int i = bsplineDegree; i <= N - bsplineDegree - 1; ++i) {
      V3 cpt = col(i);
      for (size_t j = 0; j < dynPos.size(); ++j) {
        double hx = dynSize[j].x / 2, hy = dynSize[j].y / 2;
        double size = std::pow(hx * hx + hy * hy, 0.5);  // bT.cpp:1013 (pow(.,0.5), kept literal)
        V3 ov = dynVel[j];
        for (int n = 0; n <= predictionNum; n += skipFactor) {
          V3 op = dynPos[j] + (double)(n * P.ts) * ov;
          // double(n/predictionNum): INTEGER division (reference quirk, bT.cpp:1020)
          double distThresh = (1 - (double)(predictionNum != 0 ? n / predictionNum : 0) * 0.2) * dd;
          V3 diff = cpt - op;
          diff.z = 0.0;
          double dn = norm3(diff);
          double dist = dn - size;
          double distErr = distThresh - dist;
          V3 g = diff / dn;
          if (distErr <= 0) {
          } else if (distErr > 0 && distErr <= distThresh) {
            cost += std::pow(distErr, 3);
            V3 t = (-3.0 * (distErr * distErr)) * g;
            grad[3 * i] += t.x; grad[3 * i + 1] += t.y; grad[3 * i + 2] += t.z;
          } else if (distErr >= distThresh) {
            cost += (a * (distErr * distErr) + b * distErr + c);
            V3 t = (-(2 * a * distErr + b)) * g;
            grad[3 * i] += t.x; grad[3 * i + 1] += t.y; grad[3 * i + 2] += t.z;
          }
        }
      }
    }
  }
  // bT.cpp:802-821
  double terms[4] = {0, 0, 0, 0};
  double costFunction(const double* x, double* grad, int n) {
    std::memcpy(ctrl.data() + 3 * bsplineDegree, x, n * sizeof(double));
    gD.assign(3 * N, 0.0); gS.assign(3 * N, 0.0); gF.assign(3 * N, 0.0); gO.assign(3 * N, 0.0);
    double D, S, F, O;
    getDistanceCost(D, gD);
    getSmoothnessCost(S, gS);
    getFeasibilityCost(F, gF);
    getDynamicObstacleCost(O, gO);
    terms[0] = D; terms[1] = S; terms[2] = F; terms[3] = O;
    double total = wDist * D + P.w_smooth * S + P.w_feas * F + wDyn * O;
    for (int e = 0; e < n; ++e) {
      int k = 3 * bsplineDegree + e;
      grad[e] = wDist * gD[k] + P.w_smooth * gS[k] + P.w_feas * gF[k] + wDyn * gO[k];
    }
    return total;
  }
  double operator()(const double* x, double* g, int n) { return costFunction(x, g, n); }
#ifdef TP_ORACLE_REF_LBFGS
  int ref_evals = 0;
  static double refCallback(void* inst, const double* x, double* g, const int n) {
    Planner* p = reinterpret_cast<Planner*>(inst);
    ++p->ref_evals;
    return p->costFunction(x, g, n);
  }
#endif
  // bT.cpp:687-718.  Result lives in ctrl as last written by the cost callback (:803); x is discarded.
  LbfgsStats optimize() {
    int n = 3 * (N - 2 * bsplineDegree);
    std::vector<double> x(ctrl.begin() + 3 * bsplineDegree, ctrl.begin() + 3 * bsplineDegree + n);
    LbfgsStats st;
    if (P.fast_order) {
      // the product's default arithmetic (warp form, oracle/wform_port.hpp): operates in place on ctrl
      wform::Problem wp;
      fillWformProblem(wp);
      wform::Solver ws(wp, (P.fast_order >= 1 && P.fast_order <= 4 ? P.fast_order : 4));
      ws.trace = wtrace;
      st = ws.run(x.data());
    } else
#ifdef TP_ORACLE_REF_LBFGS
    if (P.use_ref_lbfgs) {
      lbfgs::lbfgs_parameter_t sp;
      lbfgs::lbfgs_load_default_parameters(&sp);
      sp.mem_size = P.lbfgs_m;
      sp.max_iterations = P.lbfgs_max_iter;
      sp.g_epsilon = P.lbfgs_g_eps;
      sp.max_linesearch = P.lbfgs_max_linesearch;
      double fx = 0;
      ref_evals = 0;
      st.ret = lbfgs::lbfgs_optimize(n, x.data(), &fx, Planner::refCallback, NULL, NULL, this, &sp);
      st.fx = fx;
      st.evals = ref_evals;
      st.iters = -1;  // not observable through the reference API
    } else
#endif
    {
      LbfgsParams lp;
      lp.m = P.lbfgs_m;
      lp.max_iter = P.lbfgs_max_iter;
      lp.g_eps = P.lbfgs_g_eps;
      lp.max_linesearch = P.lbfgs_max_linesearch;
      Lbfgs<Planner> solver(*this, lp);
      st = solver.optimize(n, x.data());
    }
    vclock += (long long)st.evals * (10LL * N + 2LL * n);
    stats.lbfgs_runs += 1;
    stats.lbfgs_iters += st.iters > 0 ? st.iters : 0;
    stats.lbfgs_evals += st.evals;
    stats.last_lbfgs_ret = st.ret;
    stats.final_cost = st.fx;
    last_x = x;
    return st;
  }
  wform::Trace* wtrace = nullptr;
  // inputs of the warp-form solve, constants derived as the product's make_const (tp_vigo.cu) derives them
  void fillWformProblem(wform::Problem& wp) {
    wp.N = N;
    wp.cp = ctrl.data();
    wp.pstart.assign(N + 1, 0);
    wp.pairs.clear();
    for (int i = 0; i < N; ++i) {
      for (size_t j = 0; j < guideP[i].size(); ++j) {
        const V3& p = guideP[i][j];
        const V3& v = guideV[i][j];
        const double row[7] = {p.x, p.y, p.z, v.x, v.y, v.z, map->isUnknown(p) ? 1.0 : 0.0};
        wp.pairs.insert(wp.pairs.end(), row, row + 7);
      }
      wp.pstart[i + 1] = (int)(wp.pairs.size() / 7);
    }
    wp.ctrl_pt_ts = P.ctrl_pt_ts; wp.ts = P.ts; wp.dthresh = P.dthresh; wp.dthresh_dyn = P.dthresh_dyn;
    wp.w_dist = wDist; wp.w_smooth = P.w_smooth; wp.w_feas = P.w_feas; wp.w_dyn = wDyn;
    wp.uncertain_factor = P.uncertain_factor; wp.min_height = P.min_height; wp.max_height = P.max_height;
    wp.plan_in_z = P.plan_in_z;
    wp.pred_num = (int)(P.pred_horizon / P.ts);
    const double dth = P.dthresh, hth = 0.2, dd = P.dthresh_dyn;
    wp.dist_a = 3.0 * dth; wp.dist_b = -3.0 * (dth * dth); wp.dist_c = std::pow(dth, 3);
    wp.h_a = 3.0 * hth; wp.h_b = -3 * (hth * hth); wp.h_c = std::pow(hth, 3);
    wp.dyn_a = 3.0 * dd; wp.dyn_b = -3 * (dd * dd); wp.dyn_c = std::pow(dd, 3);
    wp.ts_inv_sqr = 1 / (P.ctrl_pt_ts * P.ctrl_pt_ts);
    for (size_t j = 0; j < dynPos.size(); ++j) {
      const double a[3] = {dynPos[j].x, dynPos[j].y, dynPos[j].z}, b[3] = {dynVel[j].x, dynVel[j].y, dynVel[j].z},
                   c[3] = {dynSize[j].x, dynSize[j].y, dynSize[j].z};
      wp.dyn_pos.insert(wp.dyn_pos.end(), a, a + 3);
      wp.dyn_vel.insert(wp.dyn_vel.end(), b, b + 3);
      wp.dyn_size.insert(wp.dyn_size.end(), c, c + 3);
    }
    wp.g_eps = P.lbfgs_g_eps; wp.max_iter = P.lbfgs_max_iter; wp.max_linesearch = P.lbfgs_max_linesearch;
  }
  // costFunction in the warp form's arithmetic (one evaluation at the current control points)
  double costFunctionWform(double* grad, int n) {
    wform::Problem wp;
    fillWformProblem(wp);
    wform::Solver ws(wp, (P.fast_order >= 1 && P.fast_order <= 4 ? P.fast_order : 4));
    double f, dg, gg, xx;
    ws.f_const = ws.const_terms();
    ws.eval(false, f, dg, gg, xx);
    std::memcpy(grad, ws.g.data(), sizeof(double) * n);
    return f;
  }
  std::vector<double> last_x;
  long long vclock = 0;  // virtual clock, 10 ns units (see VigoParams::vclock_budget)

  // ---- collision logic ----------------------------------------------------------
  // bT.h:196-204: a accumulates by res (0, .1, .2, ... — serial adds, kept)
  bool checkCollisionLine(const V3& p1, const V3& p2) const {
    for (double a = 0.0; a <= 1.0; a += map->res) {
      V3 pMid = a * p1 + (1 - a) * p2;
      if (map->isInflatedOccupied(pMid)) return true;
    }
    return false;
  }
  // bT.h:206-240
  void shortcutPath(const std::vector<V3>& path, std::vector<V3>& sc) const {
    sc.clear();
    size_t ptr1 = 0, ptr2 = 2;
    sc.push_back(path[ptr1]);
    if (path.size() == 1) return;
    if (path.size() == 2) { sc.push_back(path[1]); return; }
    while (true) {
      if (ptr2 > path.size() - 1) break;
      V3 p1 = path[ptr1], p2 = path[ptr2];
      if (!checkCollisionLine(p1, p2)) {
        if (ptr2 >= path.size() - 1) { sc.push_back(p2); break; }
        ++ptr2;
      } else {
        sc.push_back(path[ptr2 - 1]);
        ptr1 = ptr2 - 1;
        ptr2 = ptr1 + 2;
      }
    }
  }
  // bT.h:251-304
  bool findGuidePointSemiCircle(int cpIdx, const std::pair<int, int>& seg, const std::vector<V3>& path, V3& guidePoint) const {
    double minAngle = PI_const * 0.0 / 4.0;
    double maxAngle = PI_const * 4.0 / 4.0;
    int numCp = seg.second - seg.first - 1;
    double targetAngle;
    V3 psudo;
    if (numCp != 0) {
      int order = cpIdx - seg.first;
      targetAngle = (cpIdx - seg.first) * PI_const / (numCp + 2);
      targetAngle = std::min(std::max(minAngle, targetAngle), maxAngle);
      double ratio = double(order) / double(numCp + 1.0);
      psudo = ratio * (path.back() - path[0]) + path[0];
    } else {
      targetAngle = PI_const / 2.0;
      psudo = (path[0] + path.back()) / 2.0;
    }
    V3 direction = path[0] - psudo;
    for (size_t i = 0; i + 1 < path.size(); ++i) {
      V3 wpCurr = path[i], wpNext = path[i + 1];
      double angleCurr = angleBetweenVectors(direction, wpCurr - psudo, P.soft_atan2 != 0);
      double angleNext = angleBetweenVectors(direction, wpNext - psudo, P.soft_atan2 != 0);
      if (targetAngle >= angleCurr && targetAngle <= angleNext) {
        double prevAngleDiff = 0.0;
        V3 prevTemp{0, 0, 0};
        for (double a = 1.0; a >= 0.0; a -= 0.1) {
          V3 temp = a * wpCurr + (1 - a) * wpNext;
          double tempAngle = angleBetweenVectors(direction, temp - psudo, P.soft_atan2 != 0);
          double angleDiff = tempAngle - targetAngle;
          if (angleDiff == 0) { guidePoint = temp; return true; }
          if (angleDiff * prevAngleDiff < 0) {
            double total = std::abs(angleDiff) + std::abs(prevAngleDiff);
            guidePoint = std::abs(prevAngleDiff) / total * (temp - prevTemp) + prevTemp;
            return true;
          }
          prevAngleDiff = angleDiff;
          prevTemp = temp;
        }
      }
    }
    return false;
  }
  // bT.cpp:517-571
  void assignGuidePointsSemiCircle(const std::vector<std::vector<V3>>& paths, const std::vector<std::pair<int, int>>& segs) {
    std::vector<std::vector<V3>> sc(paths.size());
    for (size_t i = 0; i < paths.size(); ++i) shortcutPath(paths[i], sc[i]);
    // H4: the reference leaves guidePoint uninitialised when the search fails; we start at 0.
    V3 guidePoint{0, 0, 0}, guideDirection;
    for (size_t i = 0; i < segs.size(); ++i) {
      if (i >= sc.size()) break;  // (merge quirk can leave fewer segs than paths, never more)
      std::pair<int, int> seg = segs[i];
      const std::vector<V3>& path = sc[i];
      for (int c = seg.first + 1; c < seg.second; ++c) {
        findGuidePointSemiCircle(c, seg, path, guidePoint);
        guideP[c].push_back(guidePoint);
        V3 d = guidePoint - col(c);
        guideDirection = d / norm3(d);
        guideV[c].push_back(guideDirection);
        ++stats.n_guide_pairs;
      }
      bool lineCollision = (seg.second - seg.first - 1 == 0);
      if (lineCollision) {
        int forwardIdx = 1;
        findGuidePointSemiCircle(seg.first, seg, path, guidePoint);
        V3 mid = (col(seg.first) + col(seg.second)) / 2.0;
        V3 d = guidePoint - mid;
        guideDirection = d / norm3(d);
        for (int c = seg.first - forwardIdx; c <= seg.second + forwardIdx; ++c) {
          if (c >= bsplineDegree && c <= N - bsplineDegree - 1) {
            guideP[c].push_back(guidePoint);
            guideV[c].push_back(guideDirection);
            ++stats.n_guide_pairs;
          }
        }
      }
    }
  }
  // bT.cpp:403-445
  void findCollisionSeg(std::vector<std::pair<int, int>>& out) const {
    out.clear();
    bool prev = false;
    int endIdx = (int)((N - bsplineDegree - 1) - P.not_check_ratio * (N - 2 * bsplineDegree));
    int s = bsplineDegree, e = bsplineDegree;
    for (int i = bsplineDegree; i <= endIdx; ++i) {
      V3 p = col(i);
      bool hit = map->isInflatedOccupied(p);
      if (hit != prev) {
        if (hit) s = i - 1;
        else { e = i; out.push_back({s, e}); }
      }
      if (hit && i == endIdx - 1) {  // corner case fires one index early (reference quirk)
        e = N - 1;
        out.push_back({s, e});
      }
      if (i != bsplineDegree) {
        if (!prev && !hit) {
          if (map->isInflatedOccupiedLine(col(i - 1), p)) out.push_back({i - 1, i});
        }
      }
      prev = hit;
    }
  }
  // bT.cpp:447-514 (incl. the merge quirk :496-511: unmerged segments are dropped)
  bool pathSearch(std::vector<std::pair<int, int>>& segs, std::vector<std::vector<V3>>& paths) {
    paths.clear();
    std::vector<int> mergeIndices;
    int num = (int)segs.size();
    for (int i = 0; i < num; ++i) {
      std::pair<int, int> seg = segs[i];
      V3 pStart = col(seg.first), pEnd = col(seg.second);
      ++stats.astar_searches;
      bool ok = astar.search(map->res, pStart, pEnd);
      stats.astar_expansions += astar.last_expansions;
      vclock += 30LL * astar.last_expansions;
      if (ok) {
        std::vector<V3> sp = astar.getPath();
        sp[0] = pStart;
        sp.push_back(pEnd);
        paths.push_back(sp);
      } else {
        if (i + 1 < num) {
          std::pair<int, int> nextSeg = segs[i + 1];
          if (nextSeg.first - seg.second <= 2) {
            V3 pS = col(seg.first), pE = col(nextSeg.second);
            ++stats.astar_searches;
            bool ok2 = astar.search(map->res, pS, pE);
            stats.astar_expansions += astar.last_expansions;
            vclock += 30LL * astar.last_expansions;
            if (ok2) {
              std::vector<V3> sp = astar.getPath();
              sp[0] = pS;
              sp.push_back(pE);
              paths.push_back(sp);
              mergeIndices.push_back(i);
              ++i;
              continue;
            }
          }
        }
        return false;
      }
    }
    if ((int)mergeIndices.size() != 0) {
      int midx = 0;
      std::vector<std::pair<int, int>> tmp;
      for (int i = 0; i < num; ++i) {
        if (midx < (int)mergeIndices.size() && i == mergeIndices[midx]) {
          tmp.push_back({segs[i].first, segs[i + 1].second});
          ++i;
          ++midx;
        }
        // else: the reference pushes into `collisionSeg` itself (:507) and then overwrites it —
        // the unmerged segments are lost.
      }
      segs = tmp;
    }
    return true;
  }
  // bT.h:307-325
  bool hasCollisionTrajectory() const {
    Bspline traj(bsplineDegree, cols(), P.ctrl_pt_ts);
    double ts = map->res / P.max_vel / 2.0;
    for (double t = 0; t <= (1.0 - P.not_check_ratio) * traj.duration; t += ts) {
      if (map->isInflatedOccupied(traj.at(t))) return true;
    }
    return false;
  }
  // bT.cpp:1433-1447
  std::vector<V3> evalTraj(double dt) const {
    std::vector<V3> out;
    Bspline traj(bsplineDegree, cols(), P.ctrl_pt_ts);
    for (double t = 0; t <= traj.duration; t += dt) out.push_back(traj.at(t));
    return out;
  }
  // bT.h:344-368
  bool hasDynamicCollisionTrajectory() const {
    std::vector<V3> traj = evalTraj(map->res / P.max_vel / 2.0);
    for (const V3& p : traj) {
      for (size_t i = 0; i < dynPos.size(); ++i) {
        double size = std::min(dynSize[i].x / 2, dynSize[i].y / 2);
        V3 diff = p - dynPos[i];
        diff.z = 0.0;
        double dist = norm3(diff) - size;
        if (dist < 0) return true;
      }
    }
    return false;
  }
  // bT.h:370-429
  static bool indexInSeg(const std::vector<std::pair<int, int>>& segs, int idx) {
    for (auto s : segs) if (idx >= s.first && idx <= s.second) return true;
    return false;
  }
  static int findSegIndex(const std::vector<std::pair<int, int>>& segs, int idx) {
    int c = 0;
    for (auto s : segs) { if (idx >= s.first && idx <= s.second) return c; ++c; }
    return -1;
  }
  bool isControlPointRequireNewGuide(int c) const {
    V3 cp = col(c);
    for (size_t i = 0; i < guideP[c].size(); ++i) {
      double dist = dot3(cp - guideP[c][i], guideV[c][i]);
      double distErr = P.dthresh - dist;
      if (distErr > 0) return false;
    }
    return true;
  }
  // bT.cpp:573-608
  bool isReguideRequired(std::vector<std::pair<int, int>>& reguideSeg) {
    std::vector<std::pair<int, int>> prevSeg = collisionSeg;
    findCollisionSeg(collisionSeg);
    std::vector<int> newPts, overlapped;
    for (auto ns : collisionSeg) {  // compareCollisionSeg, bT.h:379-403
      for (int i = ns.first + 1; i <= ns.second - 1; ++i) {
        if (indexInSeg(prevSeg, i)) overlapped.push_back(i); else newPts.push_back(i);
      }
      bool lineCollision = (ns.second - ns.first - 1 == 0);
      if (lineCollision) {
        for (int i = ns.first; i <= ns.second; ++i) {
          if (indexInSeg(prevSeg, i)) overlapped.push_back(i); else newPts.push_back(i);
        }
      }
    }
    std::set<int> segIdx;
    for (int i : newPts) segIdx.insert(findSegIndex(collisionSeg, i));
    for (int i : overlapped) {
      if (isControlPointRequireNewGuide(i)) segIdx.insert(findSegIndex(collisionSeg, i));
    }
    if (segIdx.size() == 0) return false;
    for (int i : segIdx) reguideSeg.push_back(collisionSeg[i]);
    return true;
  }
  // bT.cpp:611-685; the 0.03 s wall-clock exit (:633) is replaced by max_outer_rounds
  bool optimizeTrajectory() {
    optimize();
    double w0 = wDist, wd0 = wDyn;
    int failCount = 0;
    std::vector<std::vector<V3>> tmpPaths;
    int round = 0;
    vclock = 0;  // ros::Time startTime = ros::Time::now();  (bT.cpp:618)
    while (true) {
      bool hasCol = hasCollisionTrajectory();
      bool hasDyn = dynPos.size() != 0 ? hasDynamicCollisionTrajectory() : false;
      if (!hasCol && !hasDyn) break;
      if (round >= P.max_outer_rounds || (P.vclock_budget > 0 && vclock > (long long)P.vclock_budget)) {
        wDist = w0; wDyn = wd0;
        stats.outer_rounds = round; stats.fail_count = failCount;
        return false;
      }
      ++round;
      if (failCount >= 4) {
        std::vector<std::pair<int, int>> seg;
        findCollisionSeg(seg);
        if (pathSearch(seg, tmpPaths)) {
          astarPaths = tmpPaths;
          assignGuidePointsSemiCircle(tmpPaths, seg);
        }
      }
      if (failCount >= 8) {
        wDist = w0; wDyn = wd0;
        stats.outer_rounds = round; stats.fail_count = failCount;
        return false;
      }
      if (hasCol) {
        std::vector<std::pair<int, int>> reguideSeg;
        if (isReguideRequired(reguideSeg)) {
          if (pathSearch(reguideSeg, tmpPaths)) {
            astarPaths = tmpPaths;
            assignGuidePointsSemiCircle(tmpPaths, reguideSeg);
          } else {
            wDist *= 2.0;
            ++failCount;
          }
        } else {
          wDist *= 2.0;
          ++failCount;
        }
      }
      if (hasDyn) wDyn *= 2.0;
      optimize();
    }
    wDist = w0; wDyn = wd0;
    stats.outer_rounds = round; stats.fail_count = failCount;
    return true;
  }
  // bT.cpp:1116-1137
  void linearFeasibilityReparam() {
    double maxV = 0.0, maxA = 0.0;
    Bspline traj(bsplineDegree, cols(), P.ctrl_pt_ts);
    Bspline vel = traj.derivative();
    Bspline acc = vel.derivative();
    for (double t = 0.0; t < traj.duration; t += P.ts) {
      double v = norm3(vel.at(t)), a = norm3(acc.at(t));
      if (v > maxV) maxV = v;
      if (a > maxA) maxA = a;
    }
    double fv = P.max_vel / maxV;
    double fa = std::sqrt(P.max_acc / maxA);
    linearFactor = std::min(fv, fa);
  }
  // bT.cpp:333-385
  bool makePlan() {
    stats.linear_factor = linearFactor;
    findCollisionSeg(collisionSeg);
    if (!pathSearch(collisionSeg, astarPaths)) { stats.success = 0; return false; }
    assignGuidePointsSemiCircle(astarPaths, collisionSeg);
    bool ok = optimizeTrajectory();
    if (!ok) { stats.success = 0; return false; }
    linearFeasibilityReparam();
    stats.linear_factor = linearFactor;
    stats.success = 1;
    return true;
  }
};

}  // namespace orc

// =========================================================================== C API (ctypes)
using namespace orc;
extern "C" {

int orc_is_ref_build() {
#ifdef TP_ORACLE_REF_LBFGS
  return 1;
#else
  return 0;
#endif
}

void orc_default_params(VigoParams* p) {
  // cfg/bspline_interactive/bspline_planner_param.yaml + src/bspline_node.cpp:230-231
  std::memset(p, 0, sizeof(*p));
  p->ts = 0.1; p->dthresh = 0.5; p->max_vel = 2.0; p->max_acc = 3.0;
  p->w_distance = 1.0; p->w_smooth = 1.0; p->w_feas = 1.0; p->w_dyn = 1.0;
  p->min_height = 0.7; p->max_height = 1.3; p->uncertain_factor = 1.0;
  p->pred_horizon = 2.0; p->dthresh_dyn = 0.5; p->max_path_length = 20.0;
  p->max_obstacle_size[0] = 5; p->max_obstacle_size[1] = 5; p->max_obstacle_size[2] = 3;
  p->ctrl_pt_dist = 0.25; p->ctrl_pt_ts = 0.2; p->not_check_ratio = 0.0;
  p->lbfgs_g_eps = 0.01; p->plan_in_z = 0; p->lbfgs_m = 16; p->lbfgs_max_iter = 200;
  p->lbfgs_max_linesearch = 40; p->max_outer_rounds = 24; p->astar_max_expansions = 200000;
  p->use_ref_lbfgs = 0;
  p->soft_atan2 = 0;
  p->vclock_budget = 3000000;
}
int orc_sizeof_params() { return (int)sizeof(VigoParams); }
int orc_sizeof_stats() { return (int)sizeof(PlanStats); }

// ---- map
void* orc_map_create(double res, const double* mn, const int* dim, const int* inf) {
  OccMap* m = new OccMap();
  m->init(res, mn, dim, inf);
  return m;
}
void orc_map_free(void* m) { delete (OccMap*)m; }
void orc_map_add_points(void* m_, const double* xyz, long n) {
  OccMap* m = (OccMap*)m_;
  for (long i = 0; i < n; ++i) m->add_occupied_point({xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]});
}
void orc_map_add_cells(void* m_, const int* ijk, long n, int occupied) {
  OccMap* m = (OccMap*)m_;
  for (long i = 0; i < n; ++i) {
    if (occupied) m->add_occupied_cell(ijk[3 * i], ijk[3 * i + 1], ijk[3 * i + 2]);
    else m->add_free_cell(ijk[3 * i], ijk[3 * i + 1], ijk[3 * i + 2]);
  }
}
void orc_map_get_grids(void* m_, uint8_t* occ, uint8_t* known, uint8_t* inflated) {
  OccMap* m = (OccMap*)m_;
  size_t n = m->occ.size();
  if (occ) std::memcpy(occ, m->occ.data(), n);
  if (known) std::memcpy(known, m->known.data(), n);
  if (inflated) std::memcpy(inflated, m->inflated.data(), n);
}
void orc_map_set_grids(void* m_, const uint8_t* occ, const uint8_t* known, const uint8_t* inflated) {
  OccMap* m = (OccMap*)m_;
  size_t n = m->occ.size();
  if (occ) std::memcpy(m->occ.data(), occ, n);
  if (known) std::memcpy(m->known.data(), known, n);
  if (inflated) std::memcpy(m->inflated.data(), inflated, n);
}
void orc_query_points(void* m_, const double* xyz, long n, uint8_t* hit) {
  OccMap* m = (OccMap*)m_;
  for (long i = 0; i < n; ++i) hit[i] = m->isInflatedOccupied({xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]});
}
void orc_query_unknown(void* m_, const double* xyz, long n, uint8_t* out) {
  OccMap* m = (OccMap*)m_;
  for (long i = 0; i < n; ++i) out[i] = m->isUnknown({xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]});
}
void orc_query_lines(void* m_, const double* a, const double* b, long n, uint8_t* hit) {
  OccMap* m = (OccMap*)m_;
  for (long i = 0; i < n; ++i)
    hit[i] = m->isInflatedOccupiedLine({a[3 * i], a[3 * i + 1], a[3 * i + 2]}, {b[3 * i], b[3 * i + 1], b[3 * i + 2]});
}
void orc_point_indices(void* m_, const double* xyz, long n, int* idx, uint8_t* inside) {
  OccMap* m = (OccMap*)m_;
  for (long i = 0; i < n; ++i) inside[i] = m->index_of({xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, idx + 3 * i);
}

// ---- spline
void orc_bspline_at(const double* ctrl, int N, int degree, double ts, const double* t, int nt, double* out) {
  std::vector<V3> c(N);
  for (int i = 0; i < N; ++i) c[i] = {ctrl[3 * i], ctrl[3 * i + 1], ctrl[3 * i + 2]};
  Bspline s(degree, c, ts);
  for (int i = 0; i < nt; ++i) {
    V3 p = s.at(t[i]);
    out[3 * i] = p.x; out[3 * i + 1] = p.y; out[3 * i + 2] = p.z;
  }
}
// d-th derivative spline evaluated at t (d = 1, 2)
void orc_bspline_deriv_at(const double* ctrl, int N, double ts, int d, const double* t, int nt, double* out) {
  std::vector<V3> c(N);
  for (int i = 0; i < N; ++i) c[i] = {ctrl[3 * i], ctrl[3 * i + 1], ctrl[3 * i + 2]};
  Bspline s(bsplineDegree, c, ts);
  for (int k = 0; k < d; ++k) s = s.derivative();
  for (int i = 0; i < nt; ++i) {
    V3 p = s.at(t[i]);
    out[3 * i] = p.x; out[3 * i + 1] = p.y; out[3 * i + 2] = p.z;
  }
}

// ---- planner
void* orc_planner_create(void* map, const VigoParams* p) { return new Planner((OccMap*)map, *p); }
void orc_planner_free(void* pl) { delete (Planner*)pl; }
void orc_planner_set_params(void* pl_, const VigoParams* p) {
  Planner* pl = (Planner*)pl_;
  pl->P = *p; pl->wDist = p->w_distance; pl->wDyn = p->w_dyn;
  pl->astar.max_expansions = p->astar_max_expansions;
}
void orc_planner_set_ctrl(void* pl, const double* ctrl, int N) { ((Planner*)pl)->setControlPoints(ctrl, N); }
void orc_planner_get_ctrl(void* pl_, double* out) {
  Planner* pl = (Planner*)pl_;
  std::memcpy(out, pl->ctrl.data(), sizeof(double) * 3 * pl->N);
}
void orc_planner_add_guides(void* pl_, const int* cp, const double* p, const double* v, int G) {
  Planner* pl = (Planner*)pl_;
  for (int g = 0; g < G; ++g) {
    pl->guideP[cp[g]].push_back({p[3 * g], p[3 * g + 1], p[3 * g + 2]});
    pl->guideV[cp[g]].push_back({v[3 * g], v[3 * g + 1], v[3 * g + 2]});
  }
}
int orc_planner_get_guides(void* pl_, int* cp, double* p, double* v, int cap) {
  Planner* pl = (Planner*)pl_;
  int g = 0;
  for (int i = 0; i < pl->N; ++i)
    for (size_t j = 0; j < pl->guideP[i].size(); ++j) {
      if (g < cap) {
        cp[g] = i;
        p[3 * g] = pl->guideP[i][j].x; p[3 * g + 1] = pl->guideP[i][j].y; p[3 * g + 2] = pl->guideP[i][j].z;
        v[3 * g] = pl->guideV[i][j].x; v[3 * g + 1] = pl->guideV[i][j].y; v[3 * g + 2] = pl->guideV[i][j].z;
      }
      ++g;
    }
  return g;
}
void orc_planner_set_dyn(void* pl_, const double* pos, const double* vel, const double* size, int M) {
  Planner* pl = (Planner*)pl_;
  pl->dynPos.clear(); pl->dynVel.clear(); pl->dynSize.clear();
  for (int i = 0; i < M; ++i) {
    pl->dynPos.push_back({pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]});
    pl->dynVel.push_back({vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]});
    pl->dynSize.push_back({size[3 * i], size[3 * i + 1], size[3 * i + 2]});
  }
}
void orc_planner_set_weights(void* pl_, double wDist, double wDyn) {
  Planner* pl = (Planner*)pl_;
  pl->wDist = wDist; pl->wDyn = wDyn;
}
double orc_planner_cost(void* pl_, const double* x, double* grad, int n, double* terms4) {
  Planner* pl = (Planner*)pl_;
  double f = pl->costFunction(x, grad, n);
  if (terms4) std::memcpy(terms4, pl->terms, sizeof(double) * 4);
  return f;
}
// one optimize(); out4 = {ret, iters, evals, _}, fx; x_final = solver's x (ctrl keeps last evaluated)
int orc_planner_optimize(void* pl_, int* out4, double* fx, double* x_final) {
  Planner* pl = (Planner*)pl_;
  LbfgsStats st = pl->optimize();
  out4[0] = st.ret; out4[1] = st.iters; out4[2] = st.evals; out4[3] = 0;
  *fx = st.fx;
  if (x_final) std::memcpy(x_final, pl->last_x.data(), sizeof(double) * pl->last_x.size());
  return st.ret;
}
// costFunction in the warp form's arithmetic at x (the benchmarked kernel's evaluation)
double orc_planner_cost_wform(void* pl_, const double* x, double* grad, int n) {
  Planner* pl = (Planner*)pl_;
  std::memcpy(pl->ctrl.data() + 3 * bsplineDegree, x, n * sizeof(double));
  return pl->costFunctionWform(grad, n);
}
// The kernel's schedule of the two triangular recurrences (four steps per round of shuffles, wform::Solver::coeffs_blocked)
// against the step-by-step form (coeffs) on `cases` random Gram blocks: masked as wf_gram_update leaves them (0..16 stored
// pairs, random newest slot) or dense, with signed zeros, huge / tiny magnitudes, infinities and NaNs mixed in.  Returns
// the number of cases whose ca / cb / sc differ in any bit (NaN == NaN).
int orc_wform_coeffs_schedules_check(int seed, int cases) {
  using namespace wform;
  std::vector<double> cpbuf(3 * 8, 0.0);
  Problem wp;
  wp.N = 8;
  wp.cp = cpbuf.data();
  Solver ws(wp, 4);
  unsigned long long st = 0x9E3779B97F4A7C15ull * (unsigned long long)(seed + 1);
  auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
  auto uni = [&]() { return (double)(rnd() >> 11) / 9007199254740992.0; };
  auto val = [&](int special) {
    double v = (uni() - 0.5) * 2.0 * std::pow(10.0, (int)(rnd() % 7) - 3);
    if (special) {
      switch (rnd() % 24) {
        case 0: v = 0.0; break;
        case 1: v = -0.0; break;
        case 2: v = 1e300; break;
        case 3: v = -1e-300; break;
        case 4: v = INFINITY; break;
        case 5: v = NAN; break;
        default: break;
      }
    }
    return v;
  };
  int bad = 0;
  for (int cs = 0; cs < cases; ++cs) {
    const int newest = (int)(rnd() % M), count = (int)(rnd() % (M + 1)), dense = (cs % 5) == 4, special = (cs % 3) == 2;
    std::fill(ws.G.begin(), ws.G.end(), 0.0);
    auto age = [&](int s_) { return (newest - s_) & (M - 1); };
    for (int s_ = 0; s_ < M; ++s_) {
      const bool stored_s = dense || age(s_) < count;
      if (stored_s) { ws.Sg()[s_] = val(special); ws.Yg()[s_] = val(special); ws.INV()[s_] = val(special); }
      for (int c = 0; c < M; ++c) {
        const bool stored_c = dense || age(c) < count;
        if (!(stored_s && stored_c)) continue;
        ws.YY()[s_ * GS + c] = val(special);
        if (dense || age(s_) > age(c)) ws.A()[s_ * GS + c] = val(special);      // pair s older than pair c
        if (dense || age(c) > age(s_)) ws.Bt()[s_ * GS + c] = val(special);     // Bt[row][col]: pair col older than pair row
      }
    }
    ws.YS()[0] = val(special); ws.YS()[1] = val(special); ws.YS()[2] = val(special);
    const double gg = std::fabs(val(0));
    double r1[2 * M + 2], r2[2 * M + 2];
    ws.coeffs(newest, gg);
    std::memcpy(r1, ws.ca.data(), M * 8); std::memcpy(r1 + M, ws.cb.data(), M * 8); r1[2 * M] = ws.sc[0]; r1[2 * M + 1] = ws.sc[1];
    ws.coeffs_blocked(newest, gg);
    std::memcpy(r2, ws.ca.data(), M * 8); std::memcpy(r2 + M, ws.cb.data(), M * 8); r2[2 * M] = ws.sc[0]; r2[2 * M + 1] = ws.sc[1];
    bool same = true;   // bit for bit; two NaNs count as equal (IEEE 754 leaves a NaN's sign and payload to the implementation)
    for (int q = 0; q < 2 * M + 2; ++q)
      if (std::memcmp(&r1[q], &r2[q], 8) != 0 && !(std::isnan(r1[q]) && std::isnan(r2[q]))) same = false;
    if (!same) ++bad;
  }
  return bad;
}
// One warp-form optimize() with a per-iteration direction check: for the (g, S, Y) the warp form had at every iteration,
// recompute the direction with the reference's two-loop recursion in its serial order (lbfgs.hpp:1293-1316) and return
// the largest ||d_gram - d_twoloop|| / ||d_twoloop||; out3 = {ret, iters, evals}.
double orc_planner_wform_direction_check(void* pl_, int* out3) {
  Planner* pl = (Planner*)pl_;
  const int keep = pl->P.fast_order;
  if (!keep) pl->P.fast_order = 4;
  wform::Trace tr;
  pl->wtrace = &tr;
  LbfgsStats st = pl->optimize();
  pl->wtrace = nullptr;
  pl->P.fast_order = keep;
  out3[0] = st.ret; out3[1] = st.iters; out3[2] = st.evals;
  const int m = 16;
  double worst = 0.0;
  for (size_t it = 0; it < tr.d.size(); ++it) {
    const int n = (int)tr.g[it].size();
    const std::vector<double>& S = tr.S[it];
    const std::vector<double>& Y = tr.Y[it];
    std::vector<double> d(n), alpha(m, 0.0);
    for (int i = 0; i < n; ++i) d[i] = -tr.g[it][i];
    const int newest = tr.end[it], bound = tr.bound[it];
    const double ys = detail::dot(&Y[(size_t)newest * n], &S[(size_t)newest * n], n);
    const double yy = detail::dot(&Y[(size_t)newest * n], &Y[(size_t)newest * n], n);
    int j = (newest + 1) % m;
    for (int i = 0; i < bound; ++i) {
      j = (j + m - 1) % m;
      alpha[j] = detail::dot(&S[(size_t)j * n], d.data(), n);
      alpha[j] /= detail::dot(&Y[(size_t)j * n], &S[(size_t)j * n], n);
      const double c = -alpha[j];
      for (int e = 0; e < n; ++e) d[e] += c * Y[(size_t)j * n + e];
    }
    const double sc = ys / yy;
    for (int e = 0; e < n; ++e) d[e] *= sc;
    for (int i = 0; i < bound; ++i) {
      double beta = detail::dot(&Y[(size_t)j * n], d.data(), n);
      beta /= detail::dot(&Y[(size_t)j * n], &S[(size_t)j * n], n);
      const double c = alpha[j] - beta;
      for (int e = 0; e < n; ++e) d[e] += c * S[(size_t)j * n + e];
      j = (j + 1) % m;
    }
    double num = 0.0, den = 0.0;
    for (int e = 0; e < n; ++e) { const double df = tr.d[it][e] - d[e]; num += df * df; den += d[e] * d[e]; }
    const double rel = std::sqrt(num) / std::sqrt(den);
    if (rel > worst) worst = rel;
  }
  return worst;
}
int orc_planner_find_collision_seg(void* pl_, int* segs, int cap) {
  Planner* pl = (Planner*)pl_;
  std::vector<std::pair<int, int>> s;
  pl->findCollisionSeg(s);
  for (size_t i = 0; i < s.size() && (int)i < cap; ++i) { segs[2 * i] = s[i].first; segs[2 * i + 1] = s[i].second; }
  return (int)s.size();
}
int orc_planner_has_collision(void* pl) { return ((Planner*)pl)->hasCollisionTrajectory() ? 1 : 0; }
// initial phase of makePlan only: findCollisionSeg + pathSearch + assignGuide (bT.cpp:341-352)
int orc_planner_init_guides(void* pl_) {
  Planner* pl = (Planner*)pl_;
  pl->findCollisionSeg(pl->collisionSeg);
  if (!pl->pathSearch(pl->collisionSeg, pl->astarPaths)) return 0;
  pl->assignGuidePointsSemiCircle(pl->astarPaths, pl->collisionSeg);
  return 1;
}
int orc_planner_get_segs(void* pl_, int* segs, int cap) {
  Planner* pl = (Planner*)pl_;
  for (size_t i = 0; i < pl->collisionSeg.size() && (int)i < cap; ++i) {
    segs[2 * i] = pl->collisionSeg[i].first; segs[2 * i + 1] = pl->collisionSeg[i].second;
  }
  return (int)pl->collisionSeg.size();
}
// path k of astarPaths_ -> xyz; returns its length (or -1)
int orc_planner_get_astar_path(void* pl_, int k, double* xyz, int cap) {
  Planner* pl = (Planner*)pl_;
  if (k < 0 || k >= (int)pl->astarPaths.size()) return -1;
  const auto& p = pl->astarPaths[k];
  for (size_t i = 0; i < p.size() && (int)i < cap; ++i) { xyz[3 * i] = p[i].x; xyz[3 * i + 1] = p[i].y; xyz[3 * i + 2] = p[i].z; }
  return (int)p.size();
}
int orc_planner_make_plan(void* pl_, PlanStats* st) {
  Planner* pl = (Planner*)pl_;
  bool ok = pl->makePlan();
  if (st) *st = pl->stats;
  return ok ? 1 : 0;
}
double orc_soft_atan2(double y, double x) { return soft_atan2(y, x); }
double orc_planner_linear_factor(void* pl_) {
  Planner* pl = (Planner*)pl_;
  pl->linearFeasibilityReparam();
  return pl->linearFactor;
}
// A* alone: returns number of path points (cell centres start->goal, before the caller's
// start/end overwrite of bT.cpp:457-458) or -1 on failure
int orc_astar(void* pl_, const double* s, const double* e, double* xyz, int cap, int* expansions) {
  Planner* pl = (Planner*)pl_;
  bool ok = pl->astar.search(pl->map->res, {s[0], s[1], s[2]}, {e[0], e[1], e[2]});
  if (expansions) *expansions = pl->astar.last_expansions;
  if (!ok) return -1;
  std::vector<V3> p = pl->astar.getPath();
  for (size_t i = 0; i < p.size() && (int)i < cap; ++i) { xyz[3 * i] = p[i].x; xyz[3 * i + 1] = p[i].y; xyz[3 * i + 2] = p[i].z; }
  return (int)p.size();
}
int orc_shortcut(void* pl_, const double* path, int n, double* out, int cap) {
  Planner* pl = (Planner*)pl_;
  std::vector<V3> p(n), sc;
  for (int i = 0; i < n; ++i) p[i] = {path[3 * i], path[3 * i + 1], path[3 * i + 2]};
  pl->shortcutPath(p, sc);
  for (size_t i = 0; i < sc.size() && (int)i < cap; ++i) { out[3 * i] = sc[i].x; out[3 * i + 1] = sc[i].y; out[3 * i + 2] = sc[i].z; }
  return (int)sc.size();
}

// ---- batch makePlan over `nthreads` host threads (one problem at a time per thread): the CPU
// baseline of bench.py.  offsets[B+1] index control points (not doubles).  Returns solved count.
int orc_make_plan_batch(void* map, const VigoParams* params, int B, const int* offsets, const double* ctrl_in,
                        double* ctrl_out, PlanStats* stats, int nthreads, double* per_problem_ms) {
  if (nthreads < 1) nthreads = 1;
  std::atomic<int> next(0), okc(0);
  auto worker = [&]() {
    Planner pl((OccMap*)map, *params);
    for (;;) {
      int b = next.fetch_add(1);
      if (b >= B) break;
      int N = offsets[b + 1] - offsets[b];
      auto t0 = std::chrono::steady_clock::now();
      pl.setControlPoints(ctrl_in + 3 * (size_t)offsets[b], N);
      bool ok = pl.makePlan();
      auto t1 = std::chrono::steady_clock::now();
      if (per_problem_ms) per_problem_ms[b] = std::chrono::duration<double, std::milli>(t1 - t0).count();
      if (ok) okc.fetch_add(1);
      if (ctrl_out) std::memcpy(ctrl_out + 3 * (size_t)offsets[b], pl.ctrl.data(), sizeof(double) * 3 * N);
      if (stats) stats[b] = pl.stats;
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) th.emplace_back(worker);
  for (auto& t : th) t.join();
  return okc.load();
}

}  // extern "C"
