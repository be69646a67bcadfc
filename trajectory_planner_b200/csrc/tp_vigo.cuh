// Internal device/host structures of the ViGO batch engine (not part of the C ABI).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/tp_b200.h"
#include "tp_device.cuh"

#define TP_MAX_CTRL 512       // control points per trajectory the kernels accept
#define TP_MAX_SEG_HARD 64    // hard cap on collision segments per trajectory
#define TP_SC_CAP 64          // points per shortcut path kept for guide assignment
#ifndef TP_LB_THREADS
#define TP_LB_THREADS 128
#endif
#define TP_LB_WARPS (TP_LB_THREADS / 32)

// trajectory status while the outer loop runs (final values are TP_STATUS_*)
#define TS_ACTIVE 100

struct GuidePair {  // one (guide point, guide direction) pair; singly linked per control point
  double p[3];
  double v[3];
  int next;      // next pair of the same control point (append order), -1 = end
  int unknown;   // map_->isUnknown(p) hoisted out of the cost loop (bsplineTraj.cpp:841)
};

struct TrajState {
  int N;            // control points
  int off;          // first control point in the packed ctrl array
  int status;       // TS_ACTIVE or TP_STATUS_*
  int has_col;      // result of the last collision check (static | dynamic<<1)
  int fail_count;
  int round;
  int nseg;
  int n_pairs;
  int err;          // capacity overflow flags
  int lbfgs_runs, lbfgs_iters, lbfgs_evals, last_ret;
  int astar_searches, astar_expansions;
  int pad;
  int astar_unreach;        // a search of this trajectory proved its goal unreachable (later searches check early)
  int pad2;
  long long vclock;         // virtual clock (tp_vigo_params::vclock_budget)
  double w_dist, w_dyn;     // weightDistance_, weightDynamicObstacle_ (mutated by the outer loop)
  double final_cost, linear_factor;
  int seg[TP_MAX_SEG_HARD][2];
};

// constants derived once on the host with the reference's own expressions
struct VigoConst {
  tp_vigo_params p;
  double dist_a, dist_b, dist_c;  // bsplineTraj.cpp:835
  double h_a, h_b, h_c;           // :837
  double dyn_a, dyn_b, dyn_c;     // :1009
  double ts_inv_sqr;              // :959
  double check_ts;                // bsplineTraj.h:312
  int pred_num;                   // bsplineTraj.cpp:1007
  int pool[3];                    // bsplineTraj.cpp:191-193
  int pool_kl;                    // stored z layers per A* pool (height band, see tp_vigo.cu)
  int heap_cap, path_cap, gcap, max_seg;
  int n_t_check, n_t_reparam, n_a_line;
};

struct ANode {  // one A* grid node (astarOcc.h:16-31), 32 B = one sector; first 16 B = one LDG.128
  uint32_t stamp_state;  // (round << 2) | state
  uint32_t parent;       // packed id of cameFrom
  double g;
  uint32_t heap_pos;     // position of this node's entry in the open-set heap (valid while OPEN)
  uint32_t pad0;
  uint64_t pad1;
};
// Open-set heap: keys (fScore) and node ids in two dense shared-memory arrays, so that a sift level
// costs one LDS pair + one compare; entries beyond TP_HEAP_SMEM spill to HBM.  In-place key updates
// (astarOcc.cpp:223-228) patch the cached key through the node's heap_pos, which every sift keeps
// current with fire-and-forget stores and the neighbour evaluation loads speculatively.
#define TP_HEAP_SMEM 2048
#define TP_AXIS_MAX 256   // A* pool cells per axis (2*int(max_obstacle_size/res)) the tables accept
struct AStarSmem {
  double hk[TP_HEAP_SMEM];
  uint32_t hn[TP_HEAP_SMEM];
  // per-search lookup tables: map cell index of pool index i/j/k along each axis (-1 = outside the
  // map), computed once with the reference's own FP expressions (Index2Coord -> posToIndex)
  short tx[TP_AXIS_MAX], ty[TP_AXIS_MAX], tz[TP_AXIS_MAX];
  unsigned char band[TP_AXIS_MAX];  // cell centre z inside [min_height, max_height] (astarOcc.cpp:202)
  // staging of one expansion's accepted neighbours for the serial pass
  double st_f[32], st_g[32];
  uint32_t st_id[32], st_pos[32];
};

struct BatchView {
  int B;
  int total_pts;
  const int* off;       // [B+1]
  double* ctrl;         // [3 * total]
  TrajState* st;        // [B]
  GuidePair* pairs;     // [B * gcap]
  int* cp_head;         // [total]
  int* cp_tail;         // [total]
  int n_dyn;
  const double* dyn_pos;
  const double* dyn_vel;
  const double* dyn_size;
  const double* t_check;    // accumulated sample times of hasCollisionTrajectory (t += check_ts)
  const double* t_reparam;  // accumulated sample times of linearFeasibilityReparam (t += ts)
  const double* a_line;     // accumulated a values of checkCollisionLine (a += res)
};

struct AStarPools {
  ANode* nodes;        // [workers * (pool_nodes + 1)]
  double* heap_k;      // [workers * heap_cap]   spill area for heap entries >= TP_HEAP_SMEM
  uint32_t* heap_n;    // [workers * heap_cap]
  double* paths;       // [workers * path_cap * 3]
  double* sc;          // [workers * max_seg * TP_SC_CAP * 3]   shortcut paths
  int* sc_len;         // [workers * max_seg]
  uint32_t* rounds;    // [workers]
  int* flood_stats;    // [2] floods that proved the goal unreachable / that found it reachable (first checks only)
  size_t pool_nodes;
  int workers;
};
