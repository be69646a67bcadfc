/* tp_b200.h — C ABI of the B200-native batched trajectory-optimisation engine.
 *
 * This is the drop-in boundary for the ViGO B-spline solve path of hanyujin02/trajectory_planner.
 * The reference has no FFI / plugin interface: the boundary is the public C++ surface of
 * trajPlanner::bsplineTraj (include/trajectory_planner/bsplineTraj.h:87-181), trajPlanner::bspline
 * (bspline.h:22-35) and the four mapManager::occMap calls that path makes.  Each entry point below
 * cites the reference member it replaces; INTEGRATION.md shows the shim a maintainer adds on the
 * reference side (a ROS-free `trajPlanner::bsplineTraj` whose methods forward here).
 *
 * Conventions: plain pointers and sizes, caller-owned buffers, `int` return codes (0 = TP_OK,
 * negative = error, see tp_last_error()), no exceptions cross the boundary.  An engine handle is
 * bound to ONE GPU; use one engine per GPU (one process per GPU under torchrun, or one host thread
 * per engine).  Calls on one engine are serialised by the caller.  There is NO CPU fallback:
 * every compute entry point fails with TP_ERR_NO_DEVICE when no CUDA device is usable.
 *
 * Memory spaces: every batch entry point takes `mem` = TP_MEM_HOST (pointers are host memory; the
 * call does H2D, compute, D2H and returns when results are in the host buffers) or TP_MEM_DEVICE
 * (pointers are device memory on the engine's GPU; the call enqueues on `stream` and returns
 * without synchronising unless noted).
 */
#ifndef TP_B200_H
#define TP_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TP_OK 0
#define TP_ERR_INVALID_ARG (-1)
#define TP_ERR_NO_DEVICE (-2)
#define TP_ERR_CUDA (-3)
#define TP_ERR_IO (-4)
#define TP_ERR_NO_MAP (-5)
#define TP_ERR_CAPACITY (-6)

#define TP_MEM_HOST 0
#define TP_MEM_DEVICE 1

/* per-trajectory solve status (tp_vigo_result.status) */
#define TP_STATUS_SUCCESS 1        /* makePlan() returned true                                  */
#define TP_STATUS_FAIL_ASTAR 0     /* "Fail because of A* failure"      (bsplineTraj.cpp:345-349) */
#define TP_STATUS_FAIL_OPTIMIZE (-1) /* optimizeTrajectory() false: failCount>=8 or round cap (:633,:650) */
#define TP_STATUS_FAIL_CAPACITY (-2) /* an engine capacity (segments / guide pairs / A* heap) overflowed */
#define TP_STATUS_INVALID (-3)     /* fewer than 7 control points etc.                          */

typedef struct tp_map tp_map_t;       /* host-side occupancy map (the occMap contract)          */
typedef struct tp_engine tp_engine_t; /* one GPU: device map replica, streams, scratch          */

/* ViGO parameters: the 16 rosparam keys of bsplineTraj::initParam (bsplineTraj.cpp:24-172), the
 * compile-time constants of bsplineTraj.h:46-47,58, the L-BFGS settings of bsplineTraj.cpp:695-699
 * and deterministic caps that replace the reference's wall-clock exits. */
typedef struct tp_vigo_params {
  double ts;                   /* bspline_traj/timestep                                  0.1 */
  double dthresh;              /* bspline_traj/distance_threshold                        0.5 */
  double max_vel, max_acc;     /* updateMaxVel / updateMaxAcc (bsplineTraj.cpp:197-205)       */
  double w_distance, w_smooth, w_feas, w_dyn; /* weight_* keys                               */
  double min_height, max_height;
  double uncertain_factor;     /* uncertain_aware_factor                                     */
  double pred_horizon;         /* prediction_horizon                                         */
  double dthresh_dyn;          /* distance_threshold_dynamic                                 */
  double max_path_length;
  double max_obstacle_size[3];
  double ctrl_pt_dist;         /* controlPointDistance_ (bsplineTraj.h:46)              0.25 */
  double ctrl_pt_ts;           /* controlPointsTs_      (bsplineTraj.h:47)               0.2 */
  double not_check_ratio;      /* notCheckRatio_        (bsplineTraj.h:58)               0.0 */
  double lbfgs_g_eps;          /* solverParams.g_epsilon (bsplineTraj.cpp:699)          0.01 */
  int32_t plan_in_z;           /* plan_in_z_axis                                             */
  int32_t lbfgs_m;             /* mem_size (:697)                                         16 */
  int32_t lbfgs_max_iter;      /* max_iterations (:698)                                  200 */
  int32_t lbfgs_max_linesearch;/* lbfgs.hpp:948                                           40 */
  int32_t max_outer_rounds;    /* replaces the 0.03 s exit of bsplineTraj.cpp:633         24 */
  int32_t astar_max_expansions;/* replaces the 0.2 s exit of astarOcc.cpp:231         200000 */
  int32_t strict_order;        /* 1: serial-order reductions (bit-reproduces the CPU sums)   */
  int32_t vclock_budget;       /* deterministic stand-in for the 0.03 s wall clock of bsplineTraj.cpp:618,
                                  632-638: virtual clock in 10 ns units, started after the first optimize(),
                                  += evals*(10 N + 2 n) per optimize(), += 30 per A* expansion, checked where
                                  the reference checks its timer; 0 = off                    3000000 */
} tp_vigo_params;

/* per-trajectory outcome of tp_vigo_make_plan_batch */
typedef struct tp_vigo_result {
  int32_t status;          /* TP_STATUS_*                                                    */
  int32_t outer_rounds;    /* iterations of the optimise/check/re-guide loop (:619-681)      */
  int32_t fail_count;      /* failCount of bsplineTraj.cpp:615                               */
  int32_t lbfgs_runs;      /* number of optimize() calls                                     */
  int32_t lbfgs_iters;     /* total L-BFGS iterations (k of lbfgs.hpp:1165,1287)             */
  int32_t lbfgs_evals;     /* total cost-function evaluations                                */
  int32_t astar_searches;
  int32_t astar_expansions;
  int32_t n_guide_pairs;
  int32_t last_lbfgs_ret;  /* lbfgs status code of the last run (lbfgs.hpp:20-80 numbering)  */
  double final_cost;       /* fx of the last run                                             */
  double linear_factor;    /* linearFeasibilityReparam() (bsplineTraj.cpp:1116-1137)         */
} tp_vigo_result;

typedef struct tp_lbfgs_result {
  int32_t ret, iters, evals, reserved;
  double fx;
} tp_lbfgs_result;

typedef struct tp_map_info {
  double res;
  double origin[3];
  int32_t dims[3];
  int32_t inflate[3];
  int64_t n_occupied, n_inflated, n_known;
  int64_t packed_bytes; /* bytes of one bit-packed grid as laid out in HBM */
} tp_map_info;

typedef struct tp_engine_cfg {
  int32_t astar_workers;      /* concurrent A* searches (0 = auto from astar_mem_gb)          */
  int32_t max_segments;       /* collision segments per trajectory (default 32)              */
  int32_t max_guide_pairs;    /* guide (point,direction) pairs per trajectory (default 256)  */
  int32_t astar_heap_cap;     /* open-set capacity per search (0 = auto)                      */
  int32_t max_path_cells;     /* A* path length cap (default 4096)                            */
  int32_t lbfgs_threads;      /* threads per trajectory block: 64 | 128 (default 128)        */
  double astar_mem_gb;        /* HBM budget for A* node pools (default 48)                   */
  int32_t reserved[4];
} tp_engine_cfg;

const char* tp_last_error(void);
int tp_version(void);
int tp_device_count(void);

/* ------------------------------------------------------------------------------------ maps
 * The occMap contract (mapManager::occMap is an external, unpinned package; this header is the
 * definition both oracle and product share — SURVEY.md §8c):
 *   index = floor((p - origin)/res) per axis; inside iff 0 <= index < dims;
 *   isInflatedOccupied(p): outside -> true, else inflated bit;   (bsplineTraj.h:199,319; astarOcc.h:58)
 *   isUnknown(p):          outside -> true, else !known bit;     (bsplineTraj.cpp:841)
 *   isInflatedOccupiedLine(a,b): either endpoint hit, else samples a + i*res*unit(b-a),
 *                                i = 1 .. int(|b-a|/res)-1       (bsplineTraj.cpp:435)
 *   prebuilt insertion: a point marks its cell occupied+known and sets the inflated bit in the
 *   index box +-inflate[] around it.                                                        */
tp_map_t* tp_map_create(double res, const double origin[3], const int32_t dims[3], const int32_t inflate[3]);
void tp_map_destroy(tp_map_t* m);
int tp_map_add_points(tp_map_t* m, const double* xyz, int64_t n);
int tp_map_add_cells(tp_map_t* m, const int32_t* ijk, int64_t n, int occupied);
/* ASCII .pcd (the reference's map/square_static_map.pcd) through the prebuilt insertion */
int tp_map_load_pcd(tp_map_t* m, const char* path);
/* OctoMap .bt (the map/ directory .bt files): occupied leaves -> occupied cells, free leaves -> known cells; leaf keys
 * are mapped to cells by their centre.  The tree's res must equal the map's res. */
int tp_map_load_bt(tp_map_t* m, const char* path);
/* this repo's compact raster format (.tpm: header + bit-packed occupied/known grids) */
int tp_map_save_tpm(const tp_map_t* m, const char* path);
tp_map_t* tp_map_load_tpm(const char* path, const int32_t inflate[3]);
int tp_map_info_get(const tp_map_t* m, tp_map_info* info);
/* which: 0 occupied, 1 known, 2 inflated; out = dims[0]*dims[1]*dims[2] bytes, z fastest */
int tp_map_get_grid(const tp_map_t* m, int which, uint8_t* out);
/* .bt helper: metric bounding box of all known leaves (octomap getMetricMin/Max; polyTrajOctomap.cpp:573-574) */
int tp_bt_bbox(const char* path, double* res, double mn[3], double mx[3], int occupied_only);

/* ------------------------------------------------------------------------------------ engine */
void tp_engine_default_cfg(tp_engine_cfg* cfg);
tp_engine_t* tp_engine_create(int device, const tp_engine_cfg* cfg);
void tp_engine_destroy(tp_engine_t* e);
/* replicate the bit-packed map into this GPU's HBM (bsplineTraj::setMap, bsplineTraj.cpp:187) */
int tp_engine_set_map(tp_engine_t* e, const tp_map_t* m);
int tp_engine_synchronize(tp_engine_t* e);
/* number of kernels this engine has launched since creation (bench.py's gpu_launches) */
int64_t tp_engine_launch_count(const tp_engine_t* e);
/* the CUDA stream the engine enqueues on when `stream` arguments are NULL */
void* tp_engine_stream(tp_engine_t* e);

void tp_vigo_default_params(tp_vigo_params* p);

/* ------------------------------------------------------------------------------------ measurement
 * Per-kernel device timing with CUDA events on the launching stream (bench.py's roofline legs).
 * Kinds: 0 the per-trajectory solve kernel of tp_vigo_make_plan_batch (one entry per size-class launch; the
 * launches of one call overlap), 1 trajectory collision check, 2 (unused), 3 initial segments+A*+guides
 * (tp_vigo_init_guides_batch), 4 re-parameterisation, 5 map queries. */
#define TP_PROF_KINDS 8
typedef struct tp_profile {
  double ms[TP_PROF_KINDS];         /* summed device time of the launches of each kind            */
  int64_t launches[TP_PROF_KINDS];
  double lbfgs_flops;               /* algorithmic FP64 flops executed by the optimize kernel      */
  double lbfgs_iters, lbfgs_evals;  /* summed over all trajectories                               */
  double check_samples;             /* de Boor samples evaluated + map-queried by the check kernel */
  double query_points;              /* points answered by tp_query_* kernels                       */
} tp_profile;
int tp_engine_profile_enable(tp_engine_t* e, int on);
/* synchronises the engine's stream(s), returns the totals since the last call and resets them */
int tp_engine_profile_get(tp_engine_t* e, tp_profile* out);
/* micro-benchmarks for the roofline denominators MEASURED_PEAKS.json does not carry:
 * dependent-free FP64 FMA throughput (TFLOP/s, 2 flops per FMA) and random 32 B-sector gathers
 * from a `bytes`-sized buffer (GB/s of sectors; bytes <= ~100 MB stays L2 resident). */
int tp_microbench_fp64(tp_engine_t* e, double* tflops);
int tp_microbench_gather(tp_engine_t* e, int64_t bytes, double* gbs);

/* ------------------------------------------------------------------------------------ map queries
 * occMap::isInflatedOccupied / isUnknown / isInflatedOccupiedLine over n points (xyz = n x 3 FP64). */
int tp_query_points(tp_engine_t* e, int64_t n, const double* xyz, uint8_t* hit, int mem, void* stream);
int tp_query_unknown(tp_engine_t* e, int64_t n, const double* xyz, uint8_t* unknown, int mem, void* stream);
int tp_query_lines(tp_engine_t* e, int64_t n, const double* a, const double* b, uint8_t* hit, int mem, void* stream);

/* ------------------------------------------------------------------------------------ ViGO pieces
 * Batches are ragged: trajectory b owns control points offsets[b] .. offsets[b+1]-1 of the
 * 3 x sum(N) column-major FP64 array `ctrl` (xyz contiguous per point == optData_.controlPoints,
 * bsplineTraj.h:22).  Guide pairs are passed as a flat list sorted by trajectory:
 * g_offsets[B+1], g_cp[G] (control point index inside its trajectory), g_p[3G], g_v[3G]; within
 * one control point their order is the order of optData_.guidePoints[i] (bsplineTraj.h:23-24). */

/* bsplineTraj::costFunction (bsplineTraj.cpp:802-821) at the given control points:
 * f[B], grad[3 x sum(N-6)] (gradient of the optimised columns 3..N-4, ragged by offsets[b]-6b).
 * w_override (may be NULL): per-trajectory {weightDistance_, weightDynamicObstacle_}.        */
int tp_vigo_cost_batch(tp_engine_t* e, const tp_vigo_params* p, int32_t B, const int32_t* offsets,
                       const double* ctrl, const int32_t* g_offsets, const int32_t* g_cp, const double* g_p,
                       const double* g_v, const double* w_override, double* f, double* grad, int mem, void* stream);

/* Same with dynamic obstacles (bsplineTraj::updateDynamicObstacles, bsplineTraj.cpp:387-391; the term is
 * getDynamicObstacleCost, :1001-1064): n_dyn obstacles, pos / vel / size 3 doubles each (host pointers). */
int tp_vigo_cost_batch_dyn(tp_engine_t* e, const tp_vigo_params* p, int32_t B, const int32_t* offsets,
                           const double* ctrl, const int32_t* g_offsets, const int32_t* g_cp, const double* g_p,
                           const double* g_v, const double* w_override, int32_t n_dyn, const double* dyn_pos,
                           const double* dyn_vel, const double* dyn_size, double* f, double* grad, int mem, void* stream);

/* bsplineTraj::optimize (bsplineTraj.cpp:687-718): one fused cost+L-BFGS run per trajectory.
 * ctrl is updated in place (it keeps the last evaluated point, as the reference's callback
 * leaves it); x_final (may be NULL, 3 x sum(N-6)) receives the solver's own x.              */
int tp_vigo_optimize_batch(tp_engine_t* e, const tp_vigo_params* p, int32_t B, const int32_t* offsets, double* ctrl,
                           const int32_t* g_offsets, const int32_t* g_cp, const double* g_p, const double* g_v,
                           const double* w_override, tp_lbfgs_result* res, double* x_final, int mem, void* stream);

int tp_vigo_optimize_batch_dyn(tp_engine_t* e, const tp_vigo_params* p, int32_t B, const int32_t* offsets, double* ctrl,
                               const int32_t* g_offsets, const int32_t* g_cp, const double* g_p, const double* g_v,
                               const double* w_override, int32_t n_dyn, const double* dyn_pos, const double* dyn_vel,
                               const double* dyn_size, tp_lbfgs_result* res, double* x_final, int mem, void* stream);

/* bsplineTraj::hasCollisionTrajectory (bsplineTraj.h:307-325): hit[B] */
int tp_vigo_has_collision_batch(tp_engine_t* e, const tp_vigo_params* p, int32_t B, const int32_t* offsets,
                                const double* ctrl, uint8_t* hit, int mem, void* stream);

/* bsplineTraj::findCollisionSeg (bsplineTraj.cpp:403-445): nseg[B], segs[B x max_segments x 2] */
int tp_vigo_find_collision_seg_batch(tp_engine_t* e, const tp_vigo_params* p, int32_t B, const int32_t* offsets,
                                     const double* ctrl, int32_t* nseg, int32_t* segs, int mem, void* stream);

/* AStar::AstarSearch + getPath (astarOcc.cpp:119-254) for S independent (start,end) pairs:
 * path_len[S] (-1 = failure), paths[S x max_path_cells x 3], expansions[S] */
int tp_astar_batch(tp_engine_t* e, const tp_vigo_params* p, int32_t S, const double* starts, const double* ends,
                   int32_t* path_len, double* paths, int32_t* expansions, int mem, void* stream);

/* first three steps of makePlan (bsplineTraj.cpp:341-352): findCollisionSeg + pathSearch +
 * assignGuidePointsSemiCircle.  Outputs: ok[B], nseg/segs (collisionSeg_ after the merge step),
 * guide pairs in the flat layout above (g_offsets[B+1], capacity g_cap per trajectory =
 * engine max_guide_pairs, rows b*g_cap ..). Host memory only. */
int tp_vigo_init_guides_batch(tp_engine_t* e, const tp_vigo_params* p, int32_t B, const int32_t* offsets,
                              const double* ctrl, uint8_t* ok, int32_t* nseg, int32_t* segs, int32_t* g_count,
                              int32_t* g_cp, double* g_p, double* g_v);

/* ------------------------------------------------------------------------------------ the batched entry point
 * bsplineTraj::makePlan (bsplineTraj.cpp:333-385) for B independent problems that share the map
 * and parameters: findCollisionSeg -> A* detours -> guide points -> [L-BFGS -> collision check ->
 * re-guide / weight doubling]* -> linear time re-parameterisation.  ctrl_in / ctrl_out may alias.
 * dyn_* (may be NULL / 0): dynamic obstacles shared by the whole batch
 * (bsplineTraj::updateDynamicObstacles, bsplineTraj.cpp:326-330).
 * mem = TP_MEM_DEVICE: offsets / ctrl_in / ctrl_out / results are device pointers (dyn_* stay host pointers).  NOTED
 * EXCEPTION to "returns without synchronising": the call synchronises `stream` twice while it sets the batch up (it reads
 * the offsets back to size the shared-memory classes, and one 4-byte key per trajectory to order the class queues); the
 * solve kernels themselves are only enqueued, results are complete when `stream` has drained. */
int tp_vigo_make_plan_batch(tp_engine_t* e, const tp_vigo_params* p, int32_t B, const int32_t* offsets,
                            const double* ctrl_in, double* ctrl_out, tp_vigo_result* results, int32_t n_dyn,
                            const double* dyn_pos, const double* dyn_vel, const double* dyn_size, int mem, void* stream);

/* The same batch on SEVERAL engines of one host (one engine per GPU of the box; the map is replicated with
 * tp_engine_set_map on each): independent trajectories, no exchange step.  One host thread per
 * engine pulls contiguous chunks of `chunk` trajectories (<= 0: chosen by the library) from a shared cursor and runs
 * them through tp_vigo_make_plan_batch; host buffers only.  Results are written at the trajectories' own positions
 * and do not depend on the assignment; engine_of (may be NULL, [B]) reports which engine solved each trajectory.
 * There is no counterpart in the reference (one bsplineTraj per process, bsplineTraj.cpp:333). */
int tp_vigo_make_plan_batch_multi(tp_engine_t* const* engines, int32_t n_engines, const tp_vigo_params* p, int32_t B,
                                  const int32_t* offsets, const double* ctrl_in, double* ctrl_out, tp_vigo_result* results,
                                  int32_t n_dyn, const double* dyn_pos, const double* dyn_vel, const double* dyn_size,
                                  int32_t chunk, int32_t* engine_of);

/* ------------------------------------------------------------------------------------ front end (host side)
 * What src/bspline_node.cpp:332-371 does before makePlan, for B (start, goal) pairs: seed 1-segment
 * min-snap polynomial (exact KKT solve of the QP polyTrajSolver.cpp builds) -> resample until
 * consecutive points are <= 1.5 * ctrl_pt_dist apart (bsplineTraj::inputPathCheck, :207-245) ->
 * updatePath (:290-323: goal check, adjustPathLengthDirect, fillPath, parameterizeToBspline).
 * offsets_out[B+1]; ctrl_out capacity = ctrl_cap points; valid[b] = 0 when updatePath would
 * return false (that trajectory gets zero control points).  Returns total points or <0. */
int64_t tp_vigo_frontend_batch(const tp_map_t* m, const tp_vigo_params* p, int32_t B, const double* starts,
                               const double* goals, int32_t* offsets_out, double* ctrl_out, int64_t ctrl_cap,
                               uint8_t* valid);
/* The same front end as ONE CUDA kernel (one warp per pair): seed KKT solve, resampling loop,
 * inputPathCheck, updatePath and the least-squares B-spline fit (bspline.cpp:74-138) on the device, so that a batch of
 * raw (start, goal) pairs never leaves HBM between the front end and tp_vigo_make_plan_batch (mem = TP_MEM_DEVICE:
 * every pointer is device memory; TP_MEM_HOST: host buffers, copies inside).  Same operation order as the host
 * version: control points are bit-identical wherever glibc's pow() is correctly rounded.  Paths that need more than
 * 1024 samples per resampling pass or more than 160 points after inputPathCheck are reported invalid.
 * Returns total control points or <0. */
int64_t tp_vigo_frontend_batch_device(tp_engine_t* e, const tp_vigo_params* p, int32_t B, const double* starts,
                                      const double* goals, int32_t* offsets_out, double* ctrl_out, int64_t ctrl_cap,
                                      uint8_t* valid, int mem, void* stream);
/* bsplineTraj::inputPathCheck (bsplineTraj.cpp:207-245): returns 1 when consecutive points are already
 * <= 1.5 * ctrl_pt_dist apart, 0 otherwise; `adjusted` (may be NULL) receives the adjusted path. */
int tp_vigo_input_path_check(const tp_map_t* m, const tp_vigo_params* p, int32_t K, const double* path, double* adjusted,
                             int32_t cap, int32_t* n_adjusted);
/* bsplineTraj::updatePath (bsplineTraj.cpp:290-323) for one path of K points (+ v0,v1,a0,a1 or NULL for rest):
 * goal check, adjustPathLengthDirect, fillPath, parameterizeToBspline.  Returns the number of control points
 * written to ctrl_out, 0 when the reference's updatePath returns false, < 0 on error. */
int tp_vigo_update_path(const tp_map_t* m, const tp_vigo_params* p, int32_t K, const double* path, const double* start_end4,
                        double* ctrl_out, int32_t cap);
/* bspline::parameterizeToBspline (bspline.cpp:74-138): K points (+ v0,v1,a0,a1) -> K+2 control points */
int tp_bspline_fit(double ts, int32_t K, const double* points, const double* start_end4, double* ctrl_out);
/* bspline::at / getDerivative().at (bspline.cpp:32-72), host side: one trajectory, nt times */
int tp_bspline_eval(int32_t N, const double* ctrl, double ts, int32_t deriv, int32_t nt, const double* t, double* out);
/* The same for a whole batch on the device: bsplineTraj::getPose (bsplineTraj.cpp:1402-1419) /
 * evalTraj (:1438-1447) / evalTrajToMsg (:1502-1518) for B trajectories in one launch.  Trajectory b (control points
 * offsets[b] .. offsets[b+1]-1 of ctrl) is sampled at the times t[t_offsets[b] .. t_offsets[b+1]-1]; outputs are
 * indexed by sample: pos[3 T] (bspline_.at(t)), and when non-NULL vel[3 T], acc[3 T] (the derivative splines' at(t)) and
 * yaw[T] = atan2(v_y, v_x) (tp_atan2 of tp_device.cuh: IEEE +,-,*,/ only, within 2 ulp of libm's).  pos / vel / acc are
 * bit-identical to tp_bspline_eval.  mem = TP_MEM_DEVICE: all pointers are device pointers, the launch is asynchronous
 * on `stream`. */
int tp_vigo_sample_batch(tp_engine_t* e, double ctrl_pt_ts, int32_t B, const int32_t* offsets, const double* ctrl,
                         const int32_t* t_offsets, const double* t, double* pos, double* vel, double* acc, double* yaw,
                         int mem, void* stream);

/* ------------------------------------------------------------------------------------ min-snap / polyTraj (secondary path)
 * polyTrajSolver's QP (polyTrajSolver.cpp:241-904; degree 7, differential degree 4) solved exactly through its KKT
 * system, and polyTrajOctomap's collision-check-and-insert-waypoint loop (polyTrajOctomap.cpp:259-321, 547-656) on the
 * engine's map with this contract: a point is "collision" when it lies outside the metric bounding box of the known
 * cells, in an unknown cell, or in an occupied cell (raw occupancy, not inflated); a sample collides when any point
 * of its collision box does (polyTrajOctomap.cpp:547-569, float coordinates as octomap::point3d).
 * Paths are ragged: path b owns waypoints wp_offsets[b] .. wp_offsets[b+1]-1 (K_b = n_b - 1 segments).
 * coef layout: problem b, axis a, segment s, power d at 24*(wp_offsets[b] - b) + a*8*K_b + 8*s + d (real time,
 * de-normalised as polyTrajSolver.cpp:874-878); times = the time knots (desiredTime_). */
typedef struct tp_poly_params {
  double desired_vel;      /* desired_velocity                         1.0 */
  double delT;             /* sample_delta_time                        0.1 */
  double box[3];           /* collision_box                [0.4,0.4,0.2] */
  double map_res;          /* map_resolution (box stride)              0.2 */
  int32_t cont;            /* continuity_degree (2..4)                   4 */
  int32_t max_iter;        /* maximum_iteration_num                    100 */
  int32_t max_waypoints;   /* cap on a path's waypoints while inserting (<= 64) */
  int32_t reserved;
} tp_poly_params;
void tp_poly_default_params(tp_poly_params* p);
/* polyTrajSolver::solve (polyTrajSolver.cpp:849-904). bc (may be NULL = rest): per problem v0, v1, a0, a1 (12 doubles).
 * status[b]: 0 ok, -1 singular KKT, -2 too many segments, -3 fewer than 2 waypoints.  Host memory. */
int tp_minsnap_solve_batch(tp_engine_t* e, const tp_poly_params* p, int32_t B, const int32_t* wp_offsets, const double* waypoints,
                           const double* bc, double* coef, double* times, int32_t* status);
/* getTrajectory + checkCollisionTraj (polyTrajSolver.cpp:1125-1137, polyTrajOctomap.cpp:634-656): valid[B],
 * seg_hit[sum K] (segment i of path b at wp_offsets[b]-b+i), n_samples[B]; optional positions / per-sample flags
 * (samples[B*samp_cap*3], sample_hit[B*samp_cap]; pass NULL, NULL, 0 to skip). */
int tp_poly_check_batch(tp_engine_t* e, const tp_poly_params* p, int32_t B, const int32_t* wp_offsets, const double* waypoints,
                        const double* coef, const double* times, uint8_t* valid, uint8_t* seg_hit, int32_t* n_samples,
                        double* samples, uint8_t* sample_hit, int32_t samp_cap);
/* polyTrajOctomap::checkCollision (polyTrajOctomap.cpp:547-569) on n caller-supplied positions */
int tp_poly_box_collision(tp_engine_t* e, const tp_poly_params* p, int64_t n, const double* xyz, uint8_t* hit);
/* polyTrajOctomap::makePlanAddingWaypoint (polyTrajOctomap.cpp:259-321) for B paths.  The reference never refreshes the
 * solver's path after insertWaypoint (:287-289 commented out) and so re-solves the original path until maxIter; this
 * entry point implements the evident intent (re-solve the path WITH the inserted waypoints).  Outputs: the final
 * waypoint lists (wp_offsets_out[B+1], waypoints_out with room for wp_cap points), their coefficients / knots in
 * the layouts above (sized for wp_cap), valid[B], iters[B]. */
int tp_polytraj_make_plan_batch(tp_engine_t* e, const tp_poly_params* p, int32_t B, const int32_t* wp_offsets, const double* waypoints,
                                int32_t* wp_offsets_out, double* waypoints_out, int64_t wp_cap, double* coef_out, double* times_out,
                                uint8_t* valid_out, int32_t* iters_out);
/* Same with per-path boundary conditions bc[B][12] = v0, v1, a0, a1 (polyTrajOctomap::updateInitVel / updateInitAcc,
 * polyTrajOctomap.cpp:193-224; NULL = rest-to-rest).  The whole loop runs on the device, one thread block per path. */
int tp_polytraj_make_plan_batch_bc(tp_engine_t* e, const tp_poly_params* p, int32_t B, const int32_t* wp_offsets, const double* waypoints,
                                   const double* bc, int32_t* wp_offsets_out, double* waypoints_out, int64_t wp_cap, double* coef_out,
                                   double* times_out, uint8_t* valid_out, int32_t* iters_out);


/* ---- corridor constraints (polyTrajOctomap `mode: false`, the reference's default config cfg/planner_interactive.yaml:31)
 * polyTrajSolver::solve with setCorridorConstraint(corridorSizeVec, corridorRes) (polyTrajSolver.cpp:960-1012, rows of
 * constructA :555-579 / constructBound :813-840): per axis  min 1/2 c'Pc  s.t. the min-snap equality rows and
 * mid_j - r_s <= p_s(t_j) <= mid_j + r_s at the corridor samples t = 0; t <= 1; t += 1/ceil(duration_s corridorRes) of every
 * segment s (mid_j on the straight segment).  Solved to convergence by an interior-point method (the reference's OSQP
 * stops at eps 1e-3).  corridor_size[sum K]: radius of segment s of path b at wp_offsets[b] - b + s (0 is not supported:
 * the reference skips such segments).  status[3 B] per axis: 0 converged, 1 infeasible or not converged, -1 singular,
 * -2 / -3 too many / too few waypoints, -4 more than 4096 corridor rows.  coef / times in the layouts above. */
int tp_corridor_solve_batch(tp_engine_t* e, const tp_poly_params* p, int32_t B, const int32_t* wp_offsets, const double* waypoints,
                            const double* bc, const double* corridor_size, double corridor_res, double* coef, double* times,
                            int32_t* status);
/* polyTrajOctomap::makePlanCorridorConstraint (polyTrajOctomap.cpp:388-530) for B paths, the whole loop on the device (one
 * thread block per path): corridors of radius init_r (initial_radius), solve, sample + box collision check, radius of the
 * colliding segments *= fs (shrinking_factor), until collision free or maxIter.  An infeasible / unconverged QP ends a
 * path's loop with valid = 0 (the reference keeps its previous solution there, polyTrajSolver.cpp:871, which can never
 * become valid).  The waypoints do not change; r_out (may be NULL, [sum K]) = final radii, status_out (may be NULL, [3 B]). */
int tp_polytraj_corridor_plan_batch(tp_engine_t* e, const tp_poly_params* p, int32_t B, const int32_t* wp_offsets,
                                    const double* waypoints, const double* bc, double init_r, double fs, double corridor_res,
                                    double* coef_out, double* times_out, uint8_t* valid_out, int32_t* iters_out, double* r_out,
                                    int32_t* status_out);
/* polyTrajOccMap::makePlan(trajectory, corridorConstraint) (polyTrajOccMap.cpp:326-399): the same loop on the ViGO
 * occupancy map (mapManager::occMap): a trajectory sample collides when it is inflated-occupied AND unknown
 * (checkCollisionTraj, :523-546 — the reference's conjunction, preserved); bc = initVel, endVel, initAcc, endAcc are honoured
 * (:336-341).  corridor_constraint = 0: one equality-only solve, valid = 1 without a collision check (:370-374). */
int tp_polytraj_occmap_plan_batch(tp_engine_t* e, const tp_poly_params* p, int32_t B, const int32_t* wp_offsets,
                                  const double* waypoints, const double* bc, int32_t corridor_constraint, double init_r, double fs,
                                  double corridor_res, double* coef_out, double* times_out, uint8_t* valid_out, int32_t* iters_out,
                                  double* r_out, int32_t* status_out);

/* polyTrajSolver::getPose (polyTrajSolver.cpp:1026-1049) on one solution in the coef / times layout above (K segments):
 * out[4 nt] = x, y, z, yaw = atan2(dy, dx), with the reference's t == 0 -> 0.01 substitution for the heading; a time
 * outside every knot interval yields zeros (the reference returns a default-constructed pose).  Host side. */
int tp_poly_eval(int32_t K, const double* coef, const double* times, int32_t nt, const double* t, double* out);

/* ------------------------------------------------------------------------------------ pwlTraj (fallback trajectory)
 * piecewiseLinearTraj.cpp: what polyTrajOctomap / polyTrajOccMap return when no valid polynomial trajectory was found
 * (polyTrajOctomap.cpp:309-317).  Host side (serial, a handful of waypoints), no engine needed.
 * tp_pwl_plan = pwlTraj::updatePath + avgTimeAllocation (:31-46, :82-119): yaw_in NULL (useYaw = false) -> headings from
 * the segments; yaw_out[K]; times_out (room for 2 (K-1) + 1 knots: rotation and forward periods alternate); returns the
 * number of knots or a negative error.  tp_pwl_eval = pwlTraj::getPose (:199-268): out[4 nt] = x, y, z, yaw. */
int tp_pwl_plan(int32_t K, const double* path, const double* yaw_in, double desired_vel, double desired_ang_vel, double* yaw_out,
                double* times_out);
int tp_pwl_eval(int32_t K, const double* path, const double* yaw, int32_t n_times, const double* times, int32_t nt, const double* t,
                double* out);

#ifdef __cplusplus
}
#endif
#endif /* TP_B200_H */
