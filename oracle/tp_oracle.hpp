// ORACLE — TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this.  The product (trajectory_planner_b200/)
// never includes, links or calls it.
//
// Single-threaded, Eigen-free, ROS-free CPU restatement of the reference's ViGO B-spline
// solve path (hanyujin02/trajectory_planner), function by function, FP64, no FMA contraction
// (build with -ffp-contract=off; the reference's CMakeLists.txt:6 is -O3 with no -march, i.e.
// no FMA on x86-64).  Every function cites the reference file:line it follows.
//
// PARITY PINNING STATUS:
//  * The reference has no golden vectors / KATs / assertions for this path (SURVEY.md §4,
//    §8c) and its whole cannot be compiled here (needs ROS, Eigen, PCL, map_manager).
//  * The one piece that does compile standalone — solver/lbfgs.hpp — is pulled in directly
//    from /root/reference by the `oracle/_ref` build (-DTP_ORACLE_REF_LBFGS) and this
//    oracle's own L-BFGS restatement (lbfgs_port.hpp) is pinned to it bit-for-bit by
//    tests/test_oracle_cpu.py::test_lbfgs_port_is_pinned_to_reference_header.  => L-BFGS iterate: PINNED to reference code.
//  * Everything else (costs, de Boor, collision logic, A*, guide points): restated from the
//    reference source by reading it; "parity unpinned" by reference-run outputs.  It is pinned
//    by analytic known-answer tests and finite-difference gradient checks in tests/.
//  * mapManager::occMap is an external package absent from /root/reference (unpinned
//    version).  The map contract implemented here (OccMap below) is this repo's definition
//    (SURVEY.md §8c), shared by oracle and product.
//  * Eigen is absent; fixed-size 3-vector reductions use the order Eigen 3.3's
//    LinearVectorizedTraversal/CompleteUnrolling produces with 2-wide packets: (x+y)+z.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace orc {

struct V3 {
  double x, y, z;
};

// ---------------------------------------------------------------- map contract
// Stand-in for mapManager::occMap (external; call sites: bsplineTraj.h:197,199,312,319,332,
// bsplineTraj.cpp:292,412,435,736-783,841, astarOcc.h:58).
struct OccMap {
  double res;
  double mn[3];       // mapSizeMin
  int dim[3];         // cells per axis
  int inf[3];         // inflation half-widths in cells
  std::vector<uint8_t> occ;       // raw occupied (== "known" for a prebuilt PCD map)
  std::vector<uint8_t> known;     // known (free or occupied)
  std::vector<uint8_t> inflated;  // inflated occupancy
  void init(double res_, const double mn_[3], const int dim_[3], const int inf_[3]);
  inline size_t addr(int ix, int iy, int iz) const {
    return ((size_t)ix * dim[1] + iy) * dim[2] + iz;  // z fastest
  }
  bool index_of(const V3& p, int idx[3]) const;  // false if outside
  void add_occupied_point(const V3& p);          // prebuilt-map insertion + inflation
  void add_occupied_cell(int ix, int iy, int iz);
  void add_free_cell(int ix, int iy, int iz);
  bool isInflatedOccupied(const V3& p) const;
  bool isInflatedOccupiedLine(const V3& a, const V3& b) const;
  bool isUnknown(const V3& p) const;
};

// ---------------------------------------------------------------- parameters
// bsplineTraj.cpp:24-172 (rosparam keys bspline_traj/*), bsplineTraj.h:46-47,58.
// Layout is mirrored by tests/ via ctypes; keep in sync with oracle/oracle.py.
struct VigoParams {
  double ts;               // bspline_traj/timestep
  double dthresh;          // distance_threshold
  double max_vel, max_acc;
  double w_distance, w_smooth, w_feas, w_dyn;
  double min_height, max_height;
  double uncertain_factor;
  double pred_horizon;
  double dthresh_dyn;
  double max_path_length;
  double max_obstacle_size[3];
  double ctrl_pt_dist;     // controlPointDistance_ 0.25
  double ctrl_pt_ts;       // controlPointsTs_ 0.2
  double not_check_ratio;  // notCheckRatio_ 0.0
  double lbfgs_g_eps;      // 0.01
  int plan_in_z;
  int lbfgs_m;             // 16
  int lbfgs_max_iter;      // 200
  int lbfgs_max_linesearch;  // 40
  // deterministic replacements for the reference's wall-clock exits
  int max_outer_rounds;      // replaces the 0.03 s limit of bsplineTraj.cpp:633
  int astar_max_expansions;  // replaces the 0.2 s limit of astarOcc.cpp:231
  int use_ref_lbfgs;         // (only honoured by the _ref build) 1 = reference header
  int soft_atan2;            // 1: deterministic software atan2 (bit-matches the GPU); 0: std::atan2 (reference)
  // Deterministic stand-in for the 0.03 s wall clock of bsplineTraj.cpp:618,632-638: a virtual clock in
  // 10 ns units, started where the reference starts its timer (after the first optimize()), advanced by
  // evals*(10 N + 2 n) per optimize() and 30 per A* expansion (a cost model of the reference's CPU
  // path: ~0.1 us per control point per cost evaluation incl. its share of the two-loop recursion,
  // ~0.3 us per expansion), checked where the reference checks its clock.  0 disables it.
  int vclock_budget;         // default 3,000,000 (= 30 ms)
  int fast_order;            // 4 / 1: optimize() runs the product's lean-form arithmetic with a team of 4 warps (the
                             // product's default) / 1 warp (oracle/wform_port.hpp); 0: reference order
};

struct PlanStats {
  int success;
  int outer_rounds;
  int fail_count;
  int lbfgs_runs;
  int lbfgs_iters;
  int lbfgs_evals;
  int astar_searches;
  int astar_expansions;
  int n_guide_pairs;
  int last_lbfgs_ret;
  double final_cost;
  double linear_factor;
};

}  // namespace orc
