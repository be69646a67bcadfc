// Drop-in shim: `trajPlanner::bsplineTraj` with the reference's method names
// (include/trajectory_planner/bsplineTraj.h:87-181), forwarding to the C ABI of tp_b200.h.
//
// ROS-free: in a catkin workspace the maintainer swaps the two light stand-in types below for the real
// ones (`Eigen::Vector3d`, `nav_msgs::Path`) — the conversion is a loop over points — and replaces the
// `mapManager::occMap` argument of setMap by the tp_map_t the node builds from the same prebuilt PCD
// (tp_map_load_pcd) or voxel list (tp_map_add_cells).  Header-only; link libtp_b200.so.
//
// Error behaviour follows the reference: bool returns + a message on stdout; no exceptions.
#pragma once
#include <array>
#include <cmath>
#include <cstdio>
#include <memory>
#include <vector>

#include "../tp_b200.h"

namespace trajPlanner {

using Vec3 = std::array<double, 3>;          // stand-in for Eigen::Vector3d
using Path = std::vector<Vec3>;              // stand-in for nav_msgs::Path (positions only)
struct Pose { double x, y, z, yaw; };        // stand-in for geometry_msgs::PoseStamped (position + yaw)

// One engine per GPU, shared by all planners of the process (and by the batched entry point).
class engineB200 {
 public:
  explicit engineB200(int device = 0) : e_(tp_engine_create(device, nullptr)) {
    if (!e_) std::printf("[BsplineTraj]: %s\n", tp_last_error());
  }
  ~engineB200() { if (e_) tp_engine_destroy(e_); }
  engineB200(const engineB200&) = delete;
  engineB200& operator=(const engineB200&) = delete;
  tp_engine_t* get() const { return e_; }
  bool ok() const { return e_ != nullptr; }
  // The new batched entry point: makePlan() for B trajectories that share map and parameters.
  // offsets[B+1]; ctrl = 3 x sum(N) column-major (== optData_.controlPoints per trajectory), updated in place.
  bool makePlanBatch(const tp_vigo_params& p, const std::vector<int32_t>& offsets, std::vector<double>& ctrl,
                     std::vector<tp_vigo_result>& results) {
    const int32_t B = (int32_t)offsets.size() - 1;
    results.resize((size_t)B);
    return tp_vigo_make_plan_batch(e_, &p, B, offsets.data(), ctrl.data(), ctrl.data(), results.data(), 0, nullptr,
                                   nullptr, nullptr, TP_MEM_HOST, nullptr) == TP_OK;
  }
  // Batched pose-at-time queries (getPose / evalTraj for a whole batch on the GPU): trajectory b is sampled at the times
  // t[tOffsets[b] .. tOffsets[b+1]-1]; pos / vel are 3 doubles per sample, yaw one (vel / yaw may be null pointers).
  bool samplePoseBatch(double ctrlPtTs, const std::vector<int32_t>& offsets, const std::vector<double>& ctrl,
                       const std::vector<int32_t>& tOffsets, const std::vector<double>& t, std::vector<double>& pos,
                       std::vector<double>* vel, std::vector<double>* yaw) {
    const int32_t B = (int32_t)offsets.size() - 1;
    pos.assign(3 * t.size(), 0.0);
    if (vel) vel->assign(3 * t.size(), 0.0);
    if (yaw) yaw->assign(t.size(), 0.0);
    return tp_vigo_sample_batch(e_, ctrlPtTs, B, offsets.data(), ctrl.data(), tOffsets.data(), t.data(), pos.data(),
                                vel ? vel->data() : nullptr, nullptr, yaw ? yaw->data() : nullptr, TP_MEM_HOST, nullptr) == TP_OK;
  }
  // One batch on several engines of this host (one per GPU, the map set on each): independent trajectories, no exchange
  // step; one host thread per engine inside the library.  engineOf (may be null) reports who solved what.
  static bool makePlanBatchMulti(const std::vector<engineB200*>& engines, const tp_vigo_params& p, const std::vector<int32_t>& offsets,
                                 std::vector<double>& ctrl, std::vector<tp_vigo_result>& results, std::vector<int32_t>* engineOf = nullptr) {
    const int32_t B = (int32_t)offsets.size() - 1;
    results.resize((size_t)B);
    if (engineOf) engineOf->assign((size_t)B, -1);
    std::vector<tp_engine_t*> h;
    for (engineB200* e : engines) h.push_back(e->get());
    return tp_vigo_make_plan_batch_multi(h.data(), (int32_t)h.size(), &p, B, offsets.data(), ctrl.data(), ctrl.data(), results.data(), 0,
                                         nullptr, nullptr, nullptr, 0, engineOf ? engineOf->data() : nullptr) == TP_OK;
  }
 private:
  tp_engine_t* e_;
};

class bsplineTraj {
 public:
  bsplineTraj() { tp_vigo_default_params(&p_); }
  explicit bsplineTraj(const std::shared_ptr<engineB200>& eng) : eng_(eng) { tp_vigo_default_params(&p_); }
  void init(const std::shared_ptr<engineB200>& eng) { eng_ = eng; }
  tp_vigo_params& params() { return p_; }    // the 16 rosparam keys of initParam (bsplineTraj.cpp:24-172)

  void setMap(const tp_map_t* map) {          // bsplineTraj.cpp:187-195
    map_ = map;
    if (eng_ && eng_->ok() && tp_engine_set_map(eng_->get(), map) != TP_OK) std::printf("[BsplineTraj]: %s\n", tp_last_error());
  }
  void updateMaxVel(double v) { p_.max_vel = v; }   // :197-200
  void updateMaxAcc(double a) { p_.max_acc = a; }   // :202-205

  bool inputPathCheck(const Path& path, Path& adjusted, double /*dt*/, double& /*finalTime*/) {   // :207-245
    std::vector<double> adj(3 * (4 * path.size() + 4096));
    int32_t n = 0;
    const int rc = tp_vigo_input_path_check(map_, &p_, (int32_t)path.size(), flat(path), adj.data(), (int32_t)(adj.size() / 3), &n);
    adjusted.assign((size_t)(n > 0 ? n : 0), Vec3{});
    for (int i = 0; i < n; ++i) adjusted[i] = {adj[3 * i], adj[3 * i + 1], adj[3 * i + 2]};
    return rc == 1;
  }
  bool updatePath(const Path& path, const std::vector<Vec3>& startEndConditions) {   // :290-323
    if (startEndConditions.size() != 4 || !map_) return false;
    double se[12];
    for (int i = 0; i < 4; ++i) for (int a = 0; a < 3; ++a) se[3 * i + a] = startEndConditions[i][a];
    ctrl_.assign(3 * (path.size() + 1024), 0.0);
    const int n = tp_vigo_update_path(map_, &p_, (int32_t)path.size(), flat(path), se, ctrl_.data(), (int32_t)(ctrl_.size() / 3));
    if (n == 0) std::printf("[BsplineTraj]: Invalid goal position!\n");
    if (n <= 0) { init_ = false; return false; }
    ctrl_.resize(3 * (size_t)n);
    dyn_.clear();                              // clear() inside updatePath (:310,393)
    init_ = true;
    return true;
  }
  void updateDynamicObstacles(const std::vector<Vec3>& pos, const std::vector<Vec3>& vel, const std::vector<Vec3>& size) {  // :326-330
    dyn_.clear();
    for (const auto* v : {&pos, &vel, &size}) for (const Vec3& q : *v) dyn_.insert(dyn_.end(), q.begin(), q.end());
    n_dyn_ = (int)pos.size();
  }
  bool makePlan() {                            // :333-385
    if (!init_ || !eng_ || !eng_->ok()) return false;
    const int32_t N = (int32_t)(ctrl_.size() / 3);
    const int32_t off[2] = {0, N};
    const double* d = dyn_.empty() ? nullptr : dyn_.data();
    const int rc = tp_vigo_make_plan_batch(eng_->get(), &p_, 1, off, ctrl_.data(), ctrl_.data(), &res_, dyn_.empty() ? 0 : n_dyn_,
                                           d, d ? d + 3 * n_dyn_ : nullptr, d ? d + 6 * n_dyn_ : nullptr, TP_MEM_HOST, nullptr);
    if (rc != TP_OK) { std::printf("[BsplineTraj]: %s\n", tp_last_error()); return false; }
    if (res_.status == TP_STATUS_FAIL_ASTAR) std::printf("[BsplineTraj]: Fail because of A* failure.\n");
    if (res_.status != TP_STATUS_SUCCESS) return false;
    // the committed trajectory (bspline_, bsplineTraj.cpp:376-377) and its re-parameterisation factor are replaced only
    // by a SUCCESSFUL plan: after a failed replan the node keeps tracking the previous trajectory, as in the reference
    traj_ = ctrl_;
    linearFactor_ = res_.linear_factor;
    return true;
  }
  bool makePlan(Path& trajectory, bool /*yaw*/ = true) {
    if (!makePlan()) return false;
    trajectory.clear();
    for (double t = 0; t * linearFactor_ <= getDuration(); t += p_.ts) { const Pose q = getPose(t * linearFactor_, false); trajectory.push_back({q.x, q.y, q.z}); }
    return true;
  }
  // ---- queries (host side, bsplineTraj.cpp:1139-1145, 1402-1419, bsplineTraj.h:151-181)
  Pose getPose(double t, bool yaw = true) const {   // evaluates the COMMITTED trajectory (bspline_)
    double p[3] = {0, 0, 0}, v[3] = {1, 0, 0};
    const int32_t N = (int32_t)(traj_.size() / 3);
    if (N < 4) return {0, 0, 0, 0};
    tp_bspline_eval(N, traj_.data(), p_.ctrl_pt_ts, 0, 1, &t, p);
    if (yaw) tp_bspline_eval(N, traj_.data(), p_.ctrl_pt_ts, 1, 1, &t, v);
    return {p[0], p[1], p[2], yaw ? std::atan2(v[1], v[0]) : 0.0};
  }
  double getDuration() const { return traj_.size() < 12 ? 0.0 : ((double)(traj_.size() / 3) - 3.0) * p_.ctrl_pt_ts; }
  const std::vector<double>& getTrajectoryControlPoints() const { return traj_; }   // bspline_'s control points
  double getTimestep() const { return p_.ts; }
  double getLinearFactor() const { return linearFactor_; }
  double getLinearReparamTime(double t) const { return linearFactor_ * t; }
  double getInitTs() const { return p_.ctrl_pt_dist / p_.max_vel; }
  double getControlPointTs() const { return p_.ctrl_pt_ts; }
  double getControlPointDist() const { return p_.ctrl_pt_dist; }
  const std::vector<double>& getControlPoints() const { return ctrl_; }   // optData_.controlPoints, 3 x N column-major
  const tp_vigo_result& lastResult() const { return res_; }
  bool isCurrTrajValid() {
    if (!init_ || !eng_ || !eng_->ok()) return false;
    const int32_t off[2] = {0, (int32_t)(ctrl_.size() / 3)};
    uint8_t hit = 1;
    if (tp_vigo_has_collision_batch(eng_->get(), &p_, 1, off, ctrl_.data(), &hit, TP_MEM_HOST, nullptr) != TP_OK) return false;
    return hit == 0;
  }

 private:
  static const double* flat(const Path& p) { return p.empty() ? nullptr : p[0].data(); }   // std::array is contiguous
  std::shared_ptr<engineB200> eng_;
  const tp_map_t* map_ = nullptr;
  tp_vigo_params p_;
  tp_vigo_result res_{};
  std::vector<double> ctrl_, dyn_;   // working control points (optData_.controlPoints), dynamic obstacles
  std::vector<double> traj_;         // control points of the committed trajectory (bspline_)
  int n_dyn_ = 0;
  bool init_ = false;
  double linearFactor_ = 1.0;
};

}  // namespace trajPlanner
