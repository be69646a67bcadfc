// Drop-in shim: `trajPlanner::polyTrajOctomap` (include/trajectory_planner/polyTrajOctomap.h:21-137) and the fallback
// `trajPlanner::pwlTraj` (piecewiseLinearTraj.h) with the reference's method names, forwarding to the C ABI of tp_b200.h.
//
// ROS-free: `pose` is utils.h:20-41's struct; nav_msgs::Path overloads become std::vector<pose>; the octomap the
// reference fetches from octomap_server (updateMap, polyTrajOctomap.cpp:129-147) is the tp_map_t the node builds from the
// same .bt file (tp_map_load_bt) and hands over with setMap.  Visualisation publishers are not part of the path.
// Header-only; link libtp_b200.so.  Error behaviour follows the reference: messages on stdout, no exceptions.
#pragma once
#include <cmath>
#include <cstdio>
#include <memory>
#include <set>
#include <vector>

#include "../tp_b200.h"
#include "bsplineTraj_b200.hpp"   // engineB200

namespace trajPlanner {

struct pose {   // utils.h:20-41
  double x = 0, y = 0, z = 0, yaw = 0;
  pose() = default;
  pose(double _x, double _y, double _z) : x(_x), y(_y), z(_z), yaw(0) {}
  pose(double _x, double _y, double _z, double _yaw) : x(_x), y(_y), z(_z), yaw(_yaw) {}
};

// piecewise linear trajectory (piecewiseLinearTraj.cpp): rotate in place, then move, per waypoint
class pwlTraj {
 public:
  void updatePath(const std::vector<pose>& path, bool useYaw = false) { updatePath(path, desiredVel_, useYaw); }   // :31-46
  void updatePath(const std::vector<pose>& path, double desiredVel, bool useYaw = false) {                        // :66-81
    path_ = path;
    const int K = (int)path.size();
    std::vector<double> xyz((size_t)3 * K), yin((size_t)K), yout((size_t)K), times((size_t)2 * K + 1);
    for (int i = 0; i < K; ++i) { xyz[3 * i] = path[i].x; xyz[3 * i + 1] = path[i].y; xyz[3 * i + 2] = path[i].z; yin[i] = path[i].yaw; }
    const int n = tp_pwl_plan(K, xyz.data(), useYaw ? yin.data() : nullptr, desiredVel, desiredAngularVel_, yout.data(), times.data());
    desiredTime_.clear();
    if (n < 0) { std::printf("[pwlTraj]: %s\n", tp_last_error()); return; }
    desiredTime_.assign(times.begin(), times.begin() + n);
    for (int i = 0; i < K; ++i) path_[i].yaw = yout[i];
  }
  void makePlan(std::vector<pose>& trajectory, double delT) {   // :171-196
    trajectory.clear();
    if (desiredTime_.empty()) return;
    for (double t = 0; t < desiredTime_.back(); t += delT) trajectory.push_back(getPose(t));
    trajectory.push_back(getPose(desiredTime_.back()));
  }
  pose getPose(double t) const {   // :199-268
    pose p;
    const int K = (int)path_.size();
    if (K < 2 || desiredTime_.size() < 2) return K ? path_.back() : p;
    std::vector<double> xyz((size_t)3 * K), yaw((size_t)K);
    for (int i = 0; i < K; ++i) { xyz[3 * i] = path_[i].x; xyz[3 * i + 1] = path_[i].y; xyz[3 * i + 2] = path_[i].z; yaw[i] = path_[i].yaw; }
    double o[4];
    tp_pwl_eval(K, xyz.data(), yaw.data(), (int32_t)desiredTime_.size(), desiredTime_.data(), 1, &t, o);
    return pose(o[0], o[1], o[2], o[3]);
  }
  std::vector<double> getTimeKnot() const { return desiredTime_; }
  double getDuration() const { return desiredTime_.empty() ? -1.0 : desiredTime_.back(); }   // :274-281
  double getDesiredVel() const { return desiredVel_; }
  double getDesiredAngularVel() const { return desiredAngularVel_; }
  pose getFirstPose() const { return path_.empty() ? pose() : path_[0]; }

 private:
  std::vector<pose> path_;
  std::vector<double> desiredTime_;
  double desiredVel_ = 1.0, desiredAngularVel_ = 0.5;   // piecewiseLinearTraj.h:20-21
};

class polyTrajOctomap {
 public:
  polyTrajOctomap() { tp_poly_default_params(&p_); }
  explicit polyTrajOctomap(const std::shared_ptr<engineB200>& eng) : eng_(eng) { tp_poly_default_params(&p_); }
  void init(const std::shared_ptr<engineB200>& eng) { eng_ = eng; }
  // the rosparam keys of the constructor (polyTrajOctomap.cpp:14-127): desired_velocity, sample_delta_time,
  // collision_box, map_resolution, continuity_degree, maximum_iteration_num
  tp_poly_params& params() { return p_; }
  void setMode(bool addingWaypoint) { mode_ = addingWaypoint; }   // `mode`: true adding waypoint, false corridor constraint

  void updateMap(const tp_map_t* map) {   // replaces the octomap_server call of updateMap (:129-147)
    map_ = map;
    if (eng_ && eng_->ok() && tp_engine_set_map(eng_->get(), map) != TP_OK) std::printf("[Trajectory Planner INFO]: %s\n", tp_last_error());
  }
  void updatePath(const std::vector<pose>& path) { path_ = path; }   // :174-176
  void updateInitVel(double vx, double vy, double vz) { initVel_[0] = vx; initVel_[1] = vy; initVel_[2] = vz; }   // :193-204
  void updateInitAcc(double ax, double ay, double az) { initAcc_[0] = ax; initAcc_[1] = ay; initAcc_[2] = az; }   // :206-218
  void setDefaultInit() { updateInitVel(0, 0, 0); updateInitAcc(0, 0, 0); }                                       // :220-224

  void makePlan() { std::vector<pose> t; makePlan(t, p_.delT); }                                                  // :227-241
  void makePlan(std::vector<pose>& trajectory, double delT = 0.1) {                                               // :249-261
    if (mode_) makePlanAddingWaypoint(trajectory, delT);
    else makePlanCorridorConstraint(trajectory, delT);
  }
  void makePlanAddingWaypoint() { std::vector<pose> t; makePlanAddingWaypoint(t, p_.delT); }
  // polyTrajOctomap.cpp:323-385.  Like the reference, the boundary velocity / acceleration are reset to zero here
  // (setDefaultInit, :336): updateInitVel / updateInitAcc do not reach the solver through makePlan.
  void makePlanAddingWaypoint(std::vector<pose>& trajectory, double delT) { plan(trajectory, delT, nullptr, 0.0, 0.0); }
  void makePlanCorridorConstraint() { std::vector<pose> t; makePlanCorridorConstraint(t, p_.delT); }
  // polyTrajOctomap.cpp:461-530: corridor constraints of initial radius initR_, shrunk by fs_ on colliding segments
  void makePlanCorridorConstraint(std::vector<pose>& trajectory, double delT) { plan(trajectory, delT, &corridorRes_, initR_, fs_); }

  // ---- collision checking on the engine's map (polyTrajOctomap.cpp:547-656)
  bool checkCollision(const pose& p) {   // the collision box around p (:547-569)
    if (!ready()) return true;
    const double q[3] = {p.x, p.y, p.z};
    uint8_t hit = 1;
    if (tp_poly_box_collision(eng_->get(), &p_, 1, q, &hit) != TP_OK) std::printf("[Trajectory Planner INFO]: %s\n", tp_last_error());
    return hit != 0;
  }
  bool checkCollisionTraj(const std::vector<pose>& trajectory, std::vector<int>& collisionIdx) {   // :620-632
    collisionIdx.clear();
    if (!ready() || trajectory.empty()) return false;
    std::vector<double> q((size_t)3 * trajectory.size());
    for (size_t i = 0; i < trajectory.size(); ++i) { q[3 * i] = trajectory[i].x; q[3 * i + 1] = trajectory[i].y; q[3 * i + 2] = trajectory[i].z; }
    std::vector<uint8_t> hit(trajectory.size(), 1);
    if (tp_poly_box_collision(eng_->get(), &p_, (int64_t)trajectory.size(), q.data(), hit.data()) != TP_OK)
      std::printf("[Trajectory Planner INFO]: %s\n", tp_last_error());
    for (size_t i = 0; i < hit.size(); ++i) if (hit[i]) collisionIdx.push_back((int)i);
    return !collisionIdx.empty();
  }

  // ---- pose at time (:658-679), duration (:681-691)
  pose getPose(double t) {
    if (t > getDuration()) t = getDuration();
    if (path_.size() == 1) return path_[0];
    if (findValidTraj_) {
      double o[4] = {0, 0, 0, 0};
      tp_poly_eval((int32_t)times_.size() - 1, coef_.data(), times_.data(), 1, &t, o);
      return pose(o[0], o[1], o[2], o[3]);
    }
    return pwl_.getPose(t);
  }
  double getDuration() const {
    if (path_.size() == 1) return 0.0;
    if (findValidTraj_) return times_.empty() ? 0.0 : times_.back();
    return pwl_.getDuration();
  }
  double getDegree() const { return 7; }
  double getDiffDegree() const { return 4; }
  double getContinuityDegree() const { return p_.cont; }
  double getDesiredVel() const { return p_.desired_vel; }
  double getInitialRadius() const { return initR_; }
  double getShrinkFactor() const { return fs_; }
  void setCorridor(double initR, double fs, double corridorRes) { initR_ = initR; fs_ = fs; corridorRes_ = corridorRes; }
  bool foundValidTraj() const { return findValidTraj_; }
  int lastIterations() const { return iters_; }
  const std::vector<pose>& getPath() const { return path_; }                 // with the inserted waypoints
  const std::vector<double>& getCoefficients() const { return coef_; }       // axis a, segment s, power d at a*8K + 8s + d
  const std::vector<double>& getTimeKnot() const { return times_; }

 private:
  bool ready() const { return eng_ && eng_->ok() && map_; }
  void plan(std::vector<pose>& trajectory, double delT, const double* corridorRes, double initR, double fs) {
    findValidTraj_ = false;
    trajectory.clear();
    if (path_.size() == 1) { trajectory = path_; findValidTraj_ = true; return; }
    if (path_.size() < 2 || !ready()) { std::printf("[Trajectory Planner INFO]: no path / engine / map.\n"); return; }
    setDefaultInit();
    const int32_t K1 = (int32_t)path_.size();
    std::vector<double> wp((size_t)3 * K1);
    for (int i = 0; i < K1; ++i) { wp[3 * i] = path_[i].x; wp[3 * i + 1] = path_[i].y; wp[3 * i + 2] = path_[i].z; }
    const int32_t off[2] = {0, K1};
    const int64_t cap = 64;
    int32_t off_o[2] = {0, 0};
    std::vector<double> wp_o((size_t)3 * cap), coef((size_t)24 * cap), times((size_t)cap);
    uint8_t valid = 0;
    tp_poly_params q = p_;
    q.delT = delT;
    int rc;
    if (corridorRes)
      rc = tp_polytraj_corridor_plan_batch(eng_->get(), &q, 1, off, wp.data(), nullptr, initR, fs, *corridorRes, coef.data(), times.data(),
                                           &valid, &iters_, nullptr, nullptr);
    else
      rc = tp_polytraj_make_plan_batch_bc(eng_->get(), &q, 1, off, wp.data(), nullptr, off_o, wp_o.data(), cap, coef.data(), times.data(),
                                          &valid, &iters_);
    if (rc != TP_OK) { std::printf("[Trajectory Planner INFO]: %s\n", tp_last_error()); return; }
    if (!corridorRes) {   // the path with the inserted waypoints (insertWaypoint, :178-186)
      path_.clear();
      for (int i = off_o[0]; i < off_o[1]; ++i) path_.push_back(pose(wp_o[3 * i], wp_o[3 * i + 1], wp_o[3 * i + 2]));
    }
    const int K = (int)path_.size() - 1;
    coef_.assign(coef.begin(), coef.begin() + (size_t)24 * K);
    times_.assign(times.begin(), times.begin() + K + 1);
    if (valid) {
      std::printf("[Trajectory Planner INFO]: Found valid trajectory!\n");
      findValidTraj_ = true;
      // polyTrajSolver::getTrajectory (polyTrajSolver.cpp:1125-1137)
      for (double t = 0; t < times_.back(); t += delT) {
        double o[4];
        tp_poly_eval(K, coef_.data(), times_.data(), 1, &t, o);
        trajectory.push_back(pose(o[0], o[1], o[2], o[3]));
      }
      trajectory.push_back(path_.back());
    } else {
      std::printf("[Trajectory Planner INFO]: Not found. Return the best. Please consider piecewise linear trajectory!!\n");
      pwl_ = pwlTraj();
      pwl_.updatePath(path_);
      pwl_.makePlan(trajectory, delT);
    }
  }

  std::shared_ptr<engineB200> eng_;
  const tp_map_t* map_ = nullptr;
  tp_poly_params p_;
  bool mode_ = true;
  double initR_ = 0.5, fs_ = 0.8, corridorRes_ = 8.0;   // initial_radius, shrinking_factor, corridor_res
  double initVel_[3] = {0, 0, 0}, initAcc_[3] = {0, 0, 0};
  std::vector<pose> path_;
  std::vector<double> coef_, times_;
  pwlTraj pwl_;
  bool findValidTraj_ = false;
  int32_t iters_ = 0;
};

}  // namespace trajPlanner
