"""trajectory_planner_b200 — B200-native batched engine for the ViGO B-spline solve path of
hanyujin02/trajectory_planner.  The compute path is the in-tree CUDA library
libtp_b200.so (sm_100a) behind the C ABI of include/tp_b200.h; importing this package without
that library built raises — there is no CPU fallback."""
from ._capi import (TP_STATUS_SUCCESS, TP_STATUS_FAIL_ASTAR, TP_STATUS_FAIL_OPTIMIZE, TP_STATUS_FAIL_CAPACITY,
                    TP_STATUS_INVALID, RESULT_DTYPE, LBFGS_DTYPE, VigoParams, EngineCfg, TpError, build, load)
from .api import (OccMap, Engine, BsplineTraj, PolyTraj, default_poly_params, default_params, frontend_batch, bspline_fit, bspline_eval,
                  make_plan_batch_multi)

__all__ = ["OccMap", "Engine", "BsplineTraj", "PolyTraj", "default_poly_params", "default_params", "frontend_batch", "bspline_fit", "bspline_eval", "make_plan_batch_multi",
           "VigoParams", "EngineCfg", "TpError", "build", "load", "RESULT_DTYPE", "LBFGS_DTYPE"]
