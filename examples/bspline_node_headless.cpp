// Headless equivalent of the reference's src/bspline_node.cpp:239-386 main loop body (no ROS / RViz): one fixed
// start / goal on the prebuilt map, seed path -> inputPathCheck loop -> updatePath -> makePlan -> pose queries,
// through the drop-in shim.  Build: see INTEGRATION.md.  Needs a GPU to run (the engine has no CPU fallback).
#include <cstdio>

#include "../include/trajectory_planner/bsplineTraj_b200.hpp"

int main(int argc, char** argv) {
  const char* map_path = argc > 1 ? argv[1] : "data/maps/square_static.tpm";
  const int32_t inflate[3] = {4, 4, 2};   // ceil(robot_size [0.8,0.8,0.3] / (2 * 0.1)), occupancy_map.yaml:9,36
  tp_map_t* map = tp_map_load_tpm(map_path, inflate);
  if (!map) { std::printf("map: %s\n", tp_last_error()); return 2; }
  auto eng = std::make_shared<trajPlanner::engineB200>(0);
  if (!eng->ok()) return 3;   // no GPU -> no planner (by design)
  trajPlanner::bsplineTraj planner(eng);
  planner.setMap(map);
  planner.updateMaxVel(2.0);   // desired_velocity / desired_acceleration (src/bspline_node.cpp:230-231)
  planner.updateMaxAcc(3.0);
  // seed path: here simply the straight segment sampled every 0.2 m (the node uses a 1-segment min-snap polynomial)
  trajPlanner::Path path;
  for (int i = 0; i <= 80; ++i) path.push_back({-6.0 + 0.15 * i, -6.0 + 0.15 * i, 1.0});
  trajPlanner::Path adjusted;
  double finalTime = 0;
  planner.inputPathCheck(path, adjusted, planner.getInitTs(), finalTime);
  const std::vector<trajPlanner::Vec3> startEnd(4, trajPlanner::Vec3{0, 0, 0});
  if (!planner.updatePath(adjusted.empty() ? path : adjusted, startEnd)) return 4;
  const bool ok = planner.makePlan();
  const tp_vigo_result& r = planner.lastResult();
  std::printf("makePlan %s: %d rounds, %d L-BFGS iterations, %d A* expansions, duration %.2f s, linear factor %.3f\n",
              ok ? "ok" : "failed", r.outer_rounds, r.lbfgs_iters, r.astar_expansions, planner.getDuration(), planner.getLinearFactor());
  if (ok) {
    const trajPlanner::Pose q = planner.getPose(0.5 * planner.getDuration());
    std::printf("pose at T/2: %.3f %.3f %.3f yaw %.3f; valid %d\n", q.x, q.y, q.z, q.yaw, (int)planner.isCurrTrajValid());
  }
  tp_map_destroy(map);
  return ok ? 0 : 1;
}
