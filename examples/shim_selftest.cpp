// Runs the three drop-in shims (bsplineTraj, bspline, polyTrajOctomap + pwlTraj) on a GPU and prints what a node would
// read back, one "key value..." line each; tests/test_gpu_shims.py compares the lines with the Python API on the same
// inputs.  usage: shim_selftest <square.tpm> <field.tpm>
#include <cstdio>
#include <cstdlib>

#include "../include/trajectory_planner/bsplineTraj_b200.hpp"
#include "../include/trajectory_planner/bspline_b200.hpp"
#include "../include/trajectory_planner/polyTrajOctomap_b200.hpp"

using namespace trajPlanner;

static void print_vec(const char* key, const std::vector<double>& v) {
  std::printf("%s", key);
  for (double x : v) std::printf(" %.17g", x);
  std::printf("\n");
}

int main(int argc, char** argv) {
  if (argc < 3) return 64;
  auto eng = std::make_shared<engineB200>(0);
  if (!eng->ok()) return 3;
  // ---------------------------------------------------------------- bsplineTraj on the square map
  const int32_t inflate[3] = {4, 4, 2};
  tp_map_t* sq = tp_map_load_tpm(argv[1], inflate);
  if (!sq) { std::printf("map: %s\n", tp_last_error()); return 2; }
  bsplineTraj planner(eng);
  planner.setMap(sq);
  planner.updateMaxVel(2.0);
  planner.updateMaxAcc(3.0);
  planner.params().strict_order = 1;
  Path path;
  for (int i = 0; i <= 80; ++i) path.push_back({-6.0 + 0.15 * i, -6.0 + 0.15 * i, 1.0});
  const std::vector<Vec3> se(4, Vec3{0, 0, 0});
  if (!planner.updatePath(path, se)) return 4;
  print_vec("ctrl_in", planner.getControlPoints());
  const std::vector<double> ctrl0 = planner.getControlPoints();
  const bool ok = planner.makePlan();
  std::printf("makeplan %d %d %d %d\n", (int)ok, planner.lastResult().lbfgs_iters, planner.lastResult().astar_expansions, planner.lastResult().outer_rounds);
  print_vec("ctrl_out", planner.getControlPoints());
  print_vec("traj", planner.getTrajectoryControlPoints());
  std::printf("duration %.17g linear_factor %.17g valid %d\n", planner.getDuration(), planner.getLinearFactor(), (int)planner.isCurrTrajValid());
  const Pose q = planner.getPose(0.37 * planner.getDuration());
  std::printf("pose %.17g %.17g %.17g %.17g\n", q.x, q.y, q.z, q.yaw);
  {   // batched entry points of the engine object: pose sampling and the multi-engine batch (two engines on this GPU)
    const std::vector<double>& c = planner.getTrajectoryControlPoints();
    const std::vector<int32_t> off = {0, (int32_t)(c.size() / 3)}, toff = {0, 3};
    const std::vector<double> tt = {0.0, 0.37 * planner.getDuration(), planner.getDuration()};
    std::vector<double> pos, vel, yaw;
    const bool sok = eng->samplePoseBatch(planner.getControlPointTs(), off, c, toff, tt, pos, &vel, &yaw);
    std::printf("sample %d %.17g %.17g %.17g %.17g\n", (int)sok, pos[3], pos[4], pos[5], yaw[1]);
    auto eng2 = std::make_shared<engineB200>(0);
    tp_engine_set_map(eng2->get(), sq);
    std::vector<int32_t> boff = {0};
    std::vector<double> bc;
    for (int r = 0; r < 6; ++r) {   // six copies of the unoptimised control points
      bc.insert(bc.end(), ctrl0.begin(), ctrl0.end());
      boff.push_back((int32_t)(bc.size() / 3));
    }
    std::vector<tp_vigo_result> br;
    std::vector<int32_t> who;
    const bool mok = engineB200::makePlanBatchMulti({eng.get(), eng2.get()}, planner.params(), boff, bc, br, &who);
    bool same = mok;
    for (int r = 0; r < 6 && same; ++r)
      for (size_t i = 0; i < ctrl0.size(); ++i)
        if (bc[r * ctrl0.size() + i] != planner.getControlPoints()[i]) { same = false; break; }
    std::printf("multi %d same_as_single %d engines_used %d %d\n", (int)mok, (int)same, (int)who.front(), (int)who.back());
  }
  // a failed replan (start inside an obstacle column is rejected by updatePath -> the committed trajectory must survive)
  const std::vector<double> committed = planner.getTrajectoryControlPoints();
  Path bad;
  for (int i = 0; i <= 10; ++i) bad.push_back({100.0 + i, 100.0, 1.0});   // outside the map
  const bool up = planner.updatePath(bad, se);
  const bool ok2 = up && planner.makePlan();
  std::printf("replan %d %d committed_kept %d\n", (int)up, (int)ok2, (int)(planner.getTrajectoryControlPoints() == committed));
  // ---------------------------------------------------------------- bspline value type
  bspline sp(3, committed, planner.getControlPointTs());
  const Vec3 a = sp.at(1.234), v = sp.getDerivative().at(1.234), acc = sp.getDerivative().getDerivative().at(1.234);
  std::printf("bspline %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", a[0], a[1], a[2], v[0], v[1], v[2], acc[0], acc[1], acc[2],
              sp.getDuration());
  // ---------------------------------------------------------------- polyTrajOctomap on field.bt
  const int32_t none[3] = {0, 0, 0};
  tp_map_t* field = tp_map_load_tpm(argv[2], none);
  if (!field) { std::printf("map: %s\n", tp_last_error()); return 2; }
  polyTrajOctomap poly(eng);
  poly.updateMap(field);
  poly.params().max_iter = 8;
  std::vector<pose> wps;
  for (int i = 3; i + 2 < argc; i += 3) wps.push_back(pose(std::atof(argv[i]), std::atof(argv[i + 1]), std::atof(argv[i + 2])));
  for (int mode = 1; mode >= 0; --mode) {
    poly.setMode(mode == 1);
    poly.updatePath(wps);
    std::vector<pose> traj;
    poly.makePlan(traj, 0.1);
    std::printf("poly mode %d valid %d iters %d waypoints %d samples %d duration %.17g\n", mode, (int)poly.foundValidTraj(), poly.lastIterations(),
                (int)poly.getPath().size(), (int)traj.size(), poly.getDuration());
    print_vec(mode ? "poly_coef_1" : "poly_coef_0", poly.getCoefficients());
    const pose pp = poly.getPose(0.5 * poly.getDuration());
    std::printf("poly_pose %d %.17g %.17g %.17g %.17g\n", mode, pp.x, pp.y, pp.z, pp.yaw);
    std::vector<int> idx;
    std::printf("poly_traj_collides %d %d\n", mode, (int)poly.checkCollisionTraj(traj, idx));
  }
  tp_map_destroy(sq);
  tp_map_destroy(field);
  return 0;
}
