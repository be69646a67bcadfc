"""ctypes binding of the C ABI in include/tp_b200.h (libtp_b200.so, built in-tree by
trajectory_planner_b200/csrc/Makefile).  This is the only way Python reaches the engine: the
hot path is the CUDA library, there is no Python/CPU fallback."""
import ctypes as C
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TP_B200_LIB") or os.path.join(PKG_DIR, "libtp_b200.so")   # TP_B200_LIB: development builds (timing probes)
CSRC_DIR = os.path.join(PKG_DIR, "csrc")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)

TP_OK = 0
TP_MEM_HOST, TP_MEM_DEVICE = 0, 1
TP_STATUS_SUCCESS, TP_STATUS_FAIL_ASTAR, TP_STATUS_FAIL_OPTIMIZE = 1, 0, -1
TP_STATUS_FAIL_CAPACITY, TP_STATUS_INVALID = -2, -3


class VigoParams(C.Structure):
    _fields_ = [
        ("ts", C.c_double), ("dthresh", C.c_double), ("max_vel", C.c_double), ("max_acc", C.c_double),
        ("w_distance", C.c_double), ("w_smooth", C.c_double), ("w_feas", C.c_double), ("w_dyn", C.c_double),
        ("min_height", C.c_double), ("max_height", C.c_double), ("uncertain_factor", C.c_double),
        ("pred_horizon", C.c_double), ("dthresh_dyn", C.c_double), ("max_path_length", C.c_double),
        ("max_obstacle_size", C.c_double * 3), ("ctrl_pt_dist", C.c_double), ("ctrl_pt_ts", C.c_double),
        ("not_check_ratio", C.c_double), ("lbfgs_g_eps", C.c_double),
        ("plan_in_z", C.c_int32), ("lbfgs_m", C.c_int32), ("lbfgs_max_iter", C.c_int32),
        ("lbfgs_max_linesearch", C.c_int32), ("max_outer_rounds", C.c_int32), ("astar_max_expansions", C.c_int32),
        ("strict_order", C.c_int32), ("vclock_budget", C.c_int32),
    ]


class EngineCfg(C.Structure):
    _fields_ = [
        ("astar_workers", C.c_int32), ("max_segments", C.c_int32), ("max_guide_pairs", C.c_int32),
        ("astar_heap_cap", C.c_int32), ("max_path_cells", C.c_int32), ("lbfgs_threads", C.c_int32),
        ("astar_mem_gb", C.c_double), ("reserved", C.c_int32 * 4),
    ]


class PolyParams(C.Structure):
    _fields_ = [("desired_vel", C.c_double), ("delT", C.c_double), ("box", C.c_double * 3), ("map_res", C.c_double),
                ("cont", C.c_int32), ("max_iter", C.c_int32), ("max_waypoints", C.c_int32), ("reserved", C.c_int32)]


class MapInfo(C.Structure):
    _fields_ = [
        ("res", C.c_double), ("origin", C.c_double * 3), ("dims", C.c_int32 * 3), ("inflate", C.c_int32 * 3),
        ("n_occupied", C.c_int64), ("n_inflated", C.c_int64), ("n_known", C.c_int64), ("packed_bytes", C.c_int64),
    ]


class Profile(C.Structure):
    _fields_ = [("ms", C.c_double * 8), ("launches", C.c_int64 * 8), ("lbfgs_flops", C.c_double),
                ("lbfgs_iters", C.c_double), ("lbfgs_evals", C.c_double), ("check_samples", C.c_double),
                ("query_points", C.c_double)]


PROF_KINDS = ["solve", "collision_check", "plan_step", "plan_init", "reparam", "map_query", "minsnap_solve", "poly_check"]

RESULT_DTYPE = np.dtype([
    ("status", "i4"), ("outer_rounds", "i4"), ("fail_count", "i4"), ("lbfgs_runs", "i4"), ("lbfgs_iters", "i4"),
    ("lbfgs_evals", "i4"), ("astar_searches", "i4"), ("astar_expansions", "i4"), ("n_guide_pairs", "i4"),
    ("last_lbfgs_ret", "i4"), ("final_cost", "f8"), ("linear_factor", "f8")])
LBFGS_DTYPE = np.dtype([("ret", "i4"), ("iters", "i4"), ("evals", "i4"), ("reserved", "i4"), ("fx", "f8")])

# every symbol include/tp_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "tp_last_error", "tp_version", "tp_device_count", "tp_map_create", "tp_map_destroy", "tp_map_add_points",
    "tp_map_add_cells", "tp_map_load_pcd", "tp_map_load_bt", "tp_map_save_tpm", "tp_map_load_tpm", "tp_map_info_get",
    "tp_map_get_grid", "tp_bt_bbox", "tp_engine_default_cfg", "tp_engine_create", "tp_engine_destroy",
    "tp_engine_set_map", "tp_engine_synchronize", "tp_engine_launch_count", "tp_engine_stream",
    "tp_vigo_default_params", "tp_query_points", "tp_query_unknown", "tp_query_lines", "tp_vigo_cost_batch", "tp_vigo_cost_batch_dyn", "tp_vigo_optimize_batch_dyn",
    "tp_vigo_optimize_batch", "tp_vigo_has_collision_batch", "tp_vigo_sample_batch", "tp_vigo_find_collision_seg_batch", "tp_astar_batch",
    "tp_vigo_init_guides_batch", "tp_vigo_make_plan_batch", "tp_vigo_make_plan_batch_multi", "tp_vigo_frontend_batch", "tp_vigo_frontend_batch_device", "tp_vigo_input_path_check", "tp_vigo_update_path", "tp_bspline_fit",
    "tp_bspline_eval", "tp_engine_profile_enable", "tp_engine_profile_get", "tp_microbench_fp64",
    "tp_microbench_gather", "tp_poly_default_params", "tp_minsnap_solve_batch", "tp_poly_check_batch",
    "tp_poly_box_collision", "tp_polytraj_make_plan_batch", "tp_polytraj_make_plan_batch_bc", "tp_poly_eval", "tp_corridor_solve_batch", "tp_polytraj_corridor_plan_batch", "tp_polytraj_occmap_plan_batch", "tp_pwl_plan", "tp_pwl_eval",
]


def build(verbose=False):
    """Compile libtp_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC_DIR, "-j", "4"]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `make -C {CSRC_DIR}` (or __graft_entry__.build()). "
            "trajectory_planner_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.tp_last_error.restype = C.c_char_p
    L.tp_map_create.restype = vp
    L.tp_map_create.argtypes = [C.c_double, _dp, _ip, _ip]
    L.tp_map_destroy.argtypes = [vp]
    L.tp_map_add_points.argtypes = [vp, _dp, C.c_int64]
    L.tp_map_add_cells.argtypes = [vp, _ip, C.c_int64, C.c_int]
    L.tp_map_load_pcd.argtypes = [vp, C.c_char_p]
    L.tp_map_load_bt.argtypes = [vp, C.c_char_p]
    L.tp_map_save_tpm.argtypes = [vp, C.c_char_p]
    L.tp_map_load_tpm.restype = vp
    L.tp_map_load_tpm.argtypes = [C.c_char_p, _ip]
    L.tp_map_info_get.argtypes = [vp, C.POINTER(MapInfo)]
    L.tp_map_get_grid.argtypes = [vp, C.c_int, _u8p]
    L.tp_bt_bbox.argtypes = [C.c_char_p, _dp, _dp, _dp, C.c_int]
    L.tp_engine_default_cfg.argtypes = [C.POINTER(EngineCfg)]
    L.tp_engine_create.restype = vp
    L.tp_engine_create.argtypes = [C.c_int, C.POINTER(EngineCfg)]
    L.tp_engine_destroy.argtypes = [vp]
    L.tp_engine_set_map.argtypes = [vp, vp]
    L.tp_engine_synchronize.argtypes = [vp]
    L.tp_engine_launch_count.restype = C.c_int64
    L.tp_engine_launch_count.argtypes = [vp]
    L.tp_engine_stream.restype = vp
    L.tp_engine_stream.argtypes = [vp]
    L.tp_vigo_default_params.argtypes = [C.POINTER(VigoParams)]
    L.tp_query_points.argtypes = [vp, C.c_int64, vp, vp, C.c_int, vp]
    L.tp_query_unknown.argtypes = [vp, C.c_int64, vp, vp, C.c_int, vp]
    L.tp_query_lines.argtypes = [vp, C.c_int64, vp, vp, vp, C.c_int, vp]
    PP = C.POINTER(VigoParams)
    L.tp_vigo_cost_batch.argtypes = [vp, PP, C.c_int32, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, vp]
    L.tp_vigo_optimize_batch.argtypes = [vp, PP, C.c_int32, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, vp]
    L.tp_vigo_make_plan_batch_multi.argtypes = [vp, C.c_int32, PP, C.c_int32, vp, vp, vp, vp, C.c_int32, vp, vp, vp, C.c_int32, vp]
    L.tp_vigo_cost_batch_dyn.argtypes = [vp, PP, C.c_int32, vp, vp, vp, vp, vp, vp, vp, C.c_int32, vp, vp, vp, vp, vp, C.c_int, vp]
    L.tp_vigo_optimize_batch_dyn.argtypes = [vp, PP, C.c_int32, vp, vp, vp, vp, vp, vp, vp, C.c_int32, vp, vp, vp, vp, vp, C.c_int, vp]
    L.tp_vigo_sample_batch.argtypes = [vp, C.c_double, C.c_int32, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, vp]
    L.tp_vigo_has_collision_batch.argtypes = [vp, PP, C.c_int32, vp, vp, vp, C.c_int, vp]
    L.tp_vigo_find_collision_seg_batch.argtypes = [vp, PP, C.c_int32, vp, vp, vp, vp, C.c_int, vp]
    L.tp_astar_batch.argtypes = [vp, PP, C.c_int32, vp, vp, vp, vp, vp, C.c_int, vp]
    L.tp_vigo_init_guides_batch.argtypes = [vp, PP, C.c_int32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.tp_vigo_make_plan_batch.argtypes = [vp, PP, C.c_int32, vp, vp, vp, vp, C.c_int32, vp, vp, vp, C.c_int, vp]
    L.tp_vigo_frontend_batch.restype = C.c_int64
    L.tp_vigo_frontend_batch.argtypes = [vp, PP, C.c_int32, vp, vp, vp, vp, C.c_int64, vp]
    L.tp_vigo_frontend_batch_device.restype = C.c_int64
    L.tp_vigo_frontend_batch_device.argtypes = [vp, PP, C.c_int32, vp, vp, vp, vp, C.c_int64, vp, C.c_int, vp]
    L.tp_vigo_input_path_check.argtypes = [vp, PP, C.c_int32, vp, vp, C.c_int32, vp]
    L.tp_vigo_update_path.argtypes = [vp, PP, C.c_int32, vp, vp, vp, C.c_int32]
    L.tp_bspline_fit.argtypes = [C.c_double, C.c_int32, vp, vp, vp]
    L.tp_bspline_eval.argtypes = [C.c_int32, vp, C.c_double, C.c_int32, C.c_int32, vp, vp]
    QP = C.POINTER(PolyParams)
    L.tp_poly_default_params.argtypes = [QP]
    L.tp_minsnap_solve_batch.argtypes = [vp, QP, C.c_int32, vp, vp, vp, vp, vp, vp]
    L.tp_poly_check_batch.argtypes = [vp, QP, C.c_int32, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int32]
    L.tp_poly_box_collision.argtypes = [vp, QP, C.c_int64, vp, vp]
    L.tp_polytraj_make_plan_batch.argtypes = [vp, QP, C.c_int32, vp, vp, vp, vp, C.c_int64, vp, vp, vp, vp]
    L.tp_polytraj_make_plan_batch_bc.argtypes = [vp, QP, C.c_int32, vp, vp, vp, vp, vp, C.c_int64, vp, vp, vp, vp]
    L.tp_corridor_solve_batch.argtypes = [vp, QP, C.c_int32, vp, vp, vp, vp, C.c_double, vp, vp, vp]
    L.tp_polytraj_corridor_plan_batch.argtypes = [vp, QP, C.c_int32, vp, vp, vp, C.c_double, C.c_double, C.c_double, vp, vp, vp, vp, vp, vp]
    L.tp_polytraj_occmap_plan_batch.argtypes = [vp, QP, C.c_int32, vp, vp, vp, C.c_int32, C.c_double, C.c_double, C.c_double, vp, vp, vp, vp, vp, vp]
    L.tp_poly_eval.argtypes = [C.c_int32, vp, vp, C.c_int32, vp, vp]
    L.tp_pwl_plan.argtypes = [C.c_int32, vp, vp, C.c_double, C.c_double, vp, vp]
    L.tp_pwl_eval.argtypes = [C.c_int32, vp, vp, C.c_int32, vp, C.c_int32, vp, vp]
    L.tp_engine_profile_enable.argtypes = [vp, C.c_int]
    L.tp_engine_profile_get.argtypes = [vp, C.POINTER(Profile)]
    L.tp_microbench_fp64.argtypes = [vp, _dp]
    L.tp_microbench_gather.argtypes = [vp, C.c_int64, _dp]
    _lib = L
    return L


class TpError(RuntimeError):
    pass


def check(rc, what=""):
    if rc is None or (isinstance(rc, int) and rc < 0):
        msg = load().tp_last_error()
        raise TpError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")
    return rc


def ptr(a):
    """void* of a numpy array (None -> NULL)."""
    return None if a is None else a.ctypes.data_as(C.c_void_p)
