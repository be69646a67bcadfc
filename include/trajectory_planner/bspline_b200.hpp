// Drop-in shim: `trajPlanner::bspline` (include/trajectory_planner/bspline.h:22-35) — the value type the reference's
// nodes get from bsplineTraj::getTrajectory(): at(t), getDuration(), getDerivative(), getControlPoints() and the
// static parameterizeToBspline — over the C ABI of tp_b200.h (tp_bspline_eval / tp_bspline_fit; pose-at-time queries
// stay on the host).  ROS / Eigen free: Vec3 stands in for Eigen::Vector3d, control points are the
// 3 x N column-major array of Eigen::MatrixXd.  Header-only; link libtp_b200.so.
#pragma once
#include <array>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../tp_b200.h"

namespace trajPlanner {

using Vec3 = std::array<double, 3>;

class bspline {
 public:
  bspline() = default;
  // degree must be 3 for a spline built from control points (bsplineDegree, bsplineTraj.h:19); lower degrees only arise
  // from getDerivative()
  bspline(int degree, const std::vector<double>& controlPoints3xN, double ts) : base_(controlPoints3xN), ts_(ts), deriv_(3 - degree) {
    if (degree != 3) { std::printf("[bspline]: only cubic splines can be built from control points\n"); deriv_ = 0; }
  }
  Vec3 at(double t) const {                      // bspline.cpp:32-58
    Vec3 out{0, 0, 0};
    const int32_t N = (int32_t)(base_.size() / 3);
    if (N >= 4) tp_bspline_eval(N, base_.data(), ts_, deriv_, 1, &t, out.data());
    return out;
  }
  double getDuration() const {                   // knots_(N) of THIS spline: (n - degree) ts, bspline.cpp:27
    const int n = (int)(base_.size() / 3) - deriv_, degree = 3 - deriv_;
    return n > degree ? (double)(n - degree) * ts_ : 0.0;
  }
  bspline getDerivative() const {                // bspline.cpp:64-72 (velocity, then acceleration spline)
    bspline d = *this;
    if (deriv_ >= 2) { std::printf("[bspline]: derivative order above 2 is not supported\n"); return d; }
    d.deriv_ = deriv_ + 1;
    return d;
  }
  int getDegree() const { return 3 - deriv_; }
  // control points of THIS spline (for a derivative: q_i = p (c_(i+1) - c_i) / (u_(i+p+1) - u_(i+1)), bspline.cpp:66-70)
  std::vector<double> getControlPoints() const {
    std::vector<double> cp = base_;
    int n = (int)(cp.size() / 3), degree = 3;
    for (int k = 0; k < deriv_; ++k) {
      std::vector<double> q((size_t)3 * (n - 1));
      for (int i = 0; i < n - 1; ++i) {
        const double den = (double)(i + degree + 1 - degree) * ts_ - (double)(i + 1 - degree) * ts_;
        for (int a = 0; a < 3; ++a) q[3 * i + a] = ((double)degree * (cp[3 * (i + 1) + a] - cp[3 * i + a])) / den;
      }
      cp.swap(q);
      --n;
      --degree;
    }
    return cp;
  }
  // bspline.cpp:74-138.  Like the reference, malformed input is fatal (exit(0), bspline.cpp:80-91).
  static void parameterizeToBspline(double ts, const std::vector<Vec3>& points, const std::vector<Vec3>& startEndConditions,
                                    std::vector<double>& controlPoints3xN) {
    if (ts <= 0 || points.size() < 4 || startEndConditions.size() != 4) {
      std::printf("[B-spline]:Invalid input.\n");
      std::exit(0);
    }
    controlPoints3xN.assign(3 * (points.size() + 2), 0.0);
    if (tp_bspline_fit(ts, (int32_t)points.size(), points[0].data(), startEndConditions[0].data(), controlPoints3xN.data()) != TP_OK) {
      std::printf("[B-spline]: %s\n", tp_last_error());
      std::exit(0);
    }
  }

 private:
  std::vector<double> base_;   // control points of the cubic spline this one derives from
  double ts_ = 0.1;
  int deriv_ = 0;
};

}  // namespace trajPlanner
