"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes front for oracle/liborc.so (and oracle/_ref).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_u8p = C.POINTER(C.c_uint8)


class VigoParams(C.Structure):
    _fields_ = [
        ("ts", C.c_double), ("dthresh", C.c_double), ("max_vel", C.c_double), ("max_acc", C.c_double),
        ("w_distance", C.c_double), ("w_smooth", C.c_double), ("w_feas", C.c_double), ("w_dyn", C.c_double),
        ("min_height", C.c_double), ("max_height", C.c_double), ("uncertain_factor", C.c_double),
        ("pred_horizon", C.c_double), ("dthresh_dyn", C.c_double), ("max_path_length", C.c_double),
        ("max_obstacle_size", C.c_double * 3), ("ctrl_pt_dist", C.c_double), ("ctrl_pt_ts", C.c_double),
        ("not_check_ratio", C.c_double), ("lbfgs_g_eps", C.c_double),
        ("plan_in_z", C.c_int), ("lbfgs_m", C.c_int), ("lbfgs_max_iter", C.c_int),
        ("lbfgs_max_linesearch", C.c_int), ("max_outer_rounds", C.c_int), ("astar_max_expansions", C.c_int),
        ("use_ref_lbfgs", C.c_int), ("soft_atan2", C.c_int), ("vclock_budget", C.c_int), ("fast_order", C.c_int),
    ]


class PlanStats(C.Structure):
    _fields_ = [
        ("success", C.c_int), ("outer_rounds", C.c_int), ("fail_count", C.c_int), ("lbfgs_runs", C.c_int),
        ("lbfgs_iters", C.c_int), ("lbfgs_evals", C.c_int), ("astar_searches", C.c_int),
        ("astar_expansions", C.c_int), ("n_guide_pairs", C.c_int), ("last_lbfgs_ret", C.c_int),
        ("final_cost", C.c_double), ("linear_factor", C.c_double),
    ]


STATS_DTYPE = np.dtype([
    ("success", "i4"), ("outer_rounds", "i4"), ("fail_count", "i4"), ("lbfgs_runs", "i4"),
    ("lbfgs_iters", "i4"), ("lbfgs_evals", "i4"), ("astar_searches", "i4"), ("astar_expansions", "i4"),
    ("n_guide_pairs", "i4"), ("last_lbfgs_ret", "i4"), ("final_cost", "f8"), ("linear_factor", "f8")])


def build(ref=True):
    """(Re)build liborc.so and, when /root/reference is present, _ref/liborc_ref.so."""
    subprocess.check_call(["make", "-s", "-C", HERE, "all" if ref else os.path.join(HERE, "liborc.so")])


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t=_dp):
    return a.ctypes.data_as(t)


class Lib:
    def __init__(self, ref=False):
        path = os.path.join(HERE, "_ref", "liborc_ref.so") if ref else os.path.join(HERE, "liborc.so")
        if not os.path.exists(path):
            build(ref=ref)
        self.path = path
        L = self.L = C.CDLL(path)
        L.orc_map_create.restype = C.c_void_p
        L.orc_map_create.argtypes = [C.c_double, _dp, _ip, _ip]
        L.orc_map_free.argtypes = [C.c_void_p]
        L.orc_map_add_points.argtypes = [C.c_void_p, _dp, C.c_long]
        L.orc_map_add_cells.argtypes = [C.c_void_p, _ip, C.c_long, C.c_int]
        L.orc_map_get_grids.argtypes = [C.c_void_p, _u8p, _u8p, _u8p]
        L.orc_map_set_grids.argtypes = [C.c_void_p, _u8p, _u8p, _u8p]
        L.orc_query_points.argtypes = [C.c_void_p, _dp, C.c_long, _u8p]
        L.orc_query_unknown.argtypes = [C.c_void_p, _dp, C.c_long, _u8p]
        L.orc_query_lines.argtypes = [C.c_void_p, _dp, _dp, C.c_long, _u8p]
        L.orc_point_indices.argtypes = [C.c_void_p, _dp, C.c_long, _ip, _u8p]
        L.orc_bspline_at.argtypes = [_dp, C.c_int, C.c_int, C.c_double, _dp, C.c_int, _dp]
        L.orc_bspline_deriv_at.argtypes = [_dp, C.c_int, C.c_double, C.c_int, _dp, C.c_int, _dp]
        L.orc_planner_create.restype = C.c_void_p
        L.orc_planner_create.argtypes = [C.c_void_p, C.POINTER(VigoParams)]
        L.orc_planner_free.argtypes = [C.c_void_p]
        L.orc_planner_set_params.argtypes = [C.c_void_p, C.POINTER(VigoParams)]
        L.orc_planner_set_ctrl.argtypes = [C.c_void_p, _dp, C.c_int]
        L.orc_planner_get_ctrl.argtypes = [C.c_void_p, _dp]
        L.orc_planner_add_guides.argtypes = [C.c_void_p, _ip, _dp, _dp, C.c_int]
        L.orc_planner_get_guides.argtypes = [C.c_void_p, _ip, _dp, _dp, C.c_int]
        L.orc_planner_set_dyn.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_int]
        L.orc_planner_set_weights.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.orc_planner_cost.restype = C.c_double
        L.orc_planner_cost.argtypes = [C.c_void_p, _dp, _dp, C.c_int, _dp]
        L.orc_planner_optimize.argtypes = [C.c_void_p, _ip, _dp, _dp]
        L.orc_planner_cost_wform.restype = C.c_double
        L.orc_planner_cost_wform.argtypes = [C.c_void_p, _dp, _dp, C.c_int]
        L.orc_planner_wform_direction_check.restype = C.c_double
        L.orc_planner_wform_direction_check.argtypes = [C.c_void_p, _ip]
        L.orc_wform_coeffs_schedules_check.restype = C.c_int
        L.orc_wform_coeffs_schedules_check.argtypes = [C.c_int, C.c_int]
        L.orc_planner_find_collision_seg.argtypes = [C.c_void_p, _ip, C.c_int]
        L.orc_planner_has_collision.argtypes = [C.c_void_p]
        L.orc_planner_init_guides.argtypes = [C.c_void_p]
        L.orc_planner_get_segs.argtypes = [C.c_void_p, _ip, C.c_int]
        L.orc_planner_get_astar_path.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int]
        L.orc_planner_make_plan.argtypes = [C.c_void_p, C.POINTER(PlanStats)]
        L.orc_planner_linear_factor.restype = C.c_double
        L.orc_planner_linear_factor.argtypes = [C.c_void_p]
        L.orc_astar.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_int, _ip]
        L.orc_shortcut.argtypes = [C.c_void_p, _dp, C.c_int, _dp, C.c_int]
        L.orc_make_plan_batch.argtypes = [C.c_void_p, C.POINTER(VigoParams), C.c_int, _ip, _dp, _dp,
                                          C.c_void_p, C.c_int, _dp]
        L.orc_default_params.argtypes = [C.POINTER(VigoParams)]
        L.orc_soft_atan2.restype = C.c_double
        L.orc_soft_atan2.argtypes = [C.c_double, C.c_double]
        assert L.orc_sizeof_params() == C.sizeof(VigoParams)
        assert L.orc_sizeof_stats() == C.sizeof(PlanStats) == STATS_DTYPE.itemsize

    def default_params(self):
        p = VigoParams()
        self.L.orc_default_params(C.byref(p))
        return p


_libs = {}


def lib(ref=False):
    if ref not in _libs:
        _libs[ref] = Lib(ref)
    return _libs[ref]


class Map:
    """Oracle occupancy map (the occMap contract of tp_oracle.hpp)."""

    def __init__(self, res, mn, dim, inf, ref=False):
        self.lib = lib(ref)
        self.res = float(res)
        self.mn = _d(mn)
        self.dim = np.ascontiguousarray(dim, dtype=np.int32)
        self.inf = np.ascontiguousarray(inf, dtype=np.int32)
        self.h = self.lib.L.orc_map_create(self.res, _p(self.mn), _p(self.dim, _ip), _p(self.inf, _ip))

    def __del__(self):
        try:
            self.lib.L.orc_map_free(self.h)
        except Exception:
            pass

    def add_points(self, xyz):
        xyz = _d(xyz).reshape(-1, 3)
        self.lib.L.orc_map_add_points(self.h, _p(xyz), len(xyz))

    def add_cells(self, ijk, occupied=True):
        ijk = np.ascontiguousarray(ijk, dtype=np.int32).reshape(-1, 3)
        self.lib.L.orc_map_add_cells(self.h, _p(ijk, _ip), len(ijk), 1 if occupied else 0)

    def grids(self):
        n = int(np.prod(self.dim))
        o, k, i = (np.zeros(n, np.uint8) for _ in range(3))
        self.lib.L.orc_map_get_grids(self.h, _p(o, _u8p), _p(k, _u8p), _p(i, _u8p))
        shp = tuple(int(v) for v in self.dim)
        return o.reshape(shp), k.reshape(shp), i.reshape(shp)

    def set_grids(self, occ, known, inflated):
        a = [np.ascontiguousarray(g, dtype=np.uint8).ravel() for g in (occ, known, inflated)]
        self.lib.L.orc_map_set_grids(self.h, _p(a[0], _u8p), _p(a[1], _u8p), _p(a[2], _u8p))

    def query(self, xyz):
        xyz = _d(xyz).reshape(-1, 3)
        out = np.zeros(len(xyz), np.uint8)
        self.lib.L.orc_query_points(self.h, _p(xyz), len(xyz), _p(out, _u8p))
        return out

    def query_unknown(self, xyz):
        xyz = _d(xyz).reshape(-1, 3)
        out = np.zeros(len(xyz), np.uint8)
        self.lib.L.orc_query_unknown(self.h, _p(xyz), len(xyz), _p(out, _u8p))
        return out

    def query_lines(self, a, b):
        a = _d(a).reshape(-1, 3)
        b = _d(b).reshape(-1, 3)
        out = np.zeros(len(a), np.uint8)
        self.lib.L.orc_query_lines(self.h, _p(a), _p(b), len(a), _p(out, _u8p))
        return out

    def indices(self, xyz):
        xyz = _d(xyz).reshape(-1, 3)
        idx = np.zeros((len(xyz), 3), np.int32)
        ins = np.zeros(len(xyz), np.uint8)
        self.lib.L.orc_point_indices(self.h, _p(xyz), len(xyz), _p(idx, _ip), _p(ins, _u8p))
        return idx, ins


class Planner:
    """One reference-style bsplineTraj instance (control points in, makePlan out)."""

    def __init__(self, omap, params=None):
        self.map = omap
        self.lib = omap.lib
        self.params = params if params is not None else self.lib.default_params()
        self.h = self.lib.L.orc_planner_create(omap.h, C.byref(self.params))
        self.N = 0

    def __del__(self):
        try:
            self.lib.L.orc_planner_free(self.h)
        except Exception:
            pass

    def set_params(self, p):
        self.params = p
        self.lib.L.orc_planner_set_params(self.h, C.byref(p))

    def set_ctrl(self, ctrl):
        c = _d(ctrl).reshape(-1, 3)
        self.N = len(c)
        self.lib.L.orc_planner_set_ctrl(self.h, _p(c), self.N)

    def get_ctrl(self):
        out = np.zeros((self.N, 3))
        self.lib.L.orc_planner_get_ctrl(self.h, _p(out))
        return out

    def add_guides(self, cp, p, v):
        cp = np.ascontiguousarray(cp, dtype=np.int32)
        p = _d(p).reshape(-1, 3)
        v = _d(v).reshape(-1, 3)
        self.lib.L.orc_planner_add_guides(self.h, _p(cp, _ip), _p(p), _p(v), len(cp))

    def get_guides(self, cap=65536):
        cp = np.zeros(cap, np.int32)
        p = np.zeros((cap, 3))
        v = np.zeros((cap, 3))
        g = self.lib.L.orc_planner_get_guides(self.h, _p(cp, _ip), _p(p), _p(v), cap)
        return cp[:g].copy(), p[:g].copy(), v[:g].copy()

    def set_dyn(self, pos, vel, size):
        pos, vel, size = (_d(a).reshape(-1, 3) for a in (pos, vel, size))
        self.lib.L.orc_planner_set_dyn(self.h, _p(pos), _p(vel), _p(size), len(pos))

    def set_weights(self, wd, wdyn):
        self.lib.L.orc_planner_set_weights(self.h, wd, wdyn)

    def cost(self, x):
        x = _d(x).ravel()
        g = np.zeros_like(x)
        terms = np.zeros(4)
        f = self.lib.L.orc_planner_cost(self.h, _p(x), _p(g), len(x), _p(terms))
        return f, g, terms

    def cost_wform(self, x):
        """costFunction in the product's warp-form arithmetic (oracle/wform_port.hpp)."""
        x = _d(x).ravel()
        g = np.zeros_like(x)
        f = self.lib.L.orc_planner_cost_wform(self.h, _p(x), _p(g), len(x))
        return f, g

    def wform_direction_check(self):
        """One warp-form optimize(); returns (worst relative difference between the Gram-form direction and the
        reference's two-loop recursion on the same (g, S, Y), stats)."""
        out = np.zeros(3, np.int32)
        worst = self.lib.L.orc_planner_wform_direction_check(self.h, _p(out, _ip))
        return worst, dict(ret=int(out[0]), iters=int(out[1]), evals=int(out[2]))

    def optimize(self):
        out = np.zeros(4, np.int32)
        fx = C.c_double(0)
        n = 3 * (self.N - 6)
        xf = np.zeros(max(n, 1))
        self.lib.L.orc_planner_optimize(self.h, _p(out, _ip), C.byref(fx), _p(xf))
        return dict(ret=int(out[0]), iters=int(out[1]), evals=int(out[2]), fx=fx.value, x=xf[:n].reshape(-1, 3))

    def find_collision_seg(self):
        s = np.zeros(2 * 256, np.int32)
        n = self.lib.L.orc_planner_find_collision_seg(self.h, _p(s, _ip), 256)
        return s[:2 * n].reshape(-1, 2).copy()

    def has_collision(self):
        return bool(self.lib.L.orc_planner_has_collision(self.h))

    def init_guides(self):
        return bool(self.lib.L.orc_planner_init_guides(self.h))

    def get_segs(self):
        s = np.zeros(2 * 256, np.int32)
        n = self.lib.L.orc_planner_get_segs(self.h, _p(s, _ip), 256)
        return s[:2 * n].reshape(-1, 2).copy()

    def get_astar_path(self, k, cap=8192):
        xyz = np.zeros((cap, 3))
        n = self.lib.L.orc_planner_get_astar_path(self.h, k, _p(xyz), cap)
        return None if n < 0 else xyz[:n].copy()

    def make_plan(self):
        st = PlanStats()
        ok = self.lib.L.orc_planner_make_plan(self.h, C.byref(st))
        return bool(ok), {f[0]: getattr(st, f[0]) for f in PlanStats._fields_}

    def linear_factor(self):
        return self.lib.L.orc_planner_linear_factor(self.h)

    def astar(self, s, e, cap=16384):
        s = _d(s)
        e = _d(e)
        xyz = np.zeros((cap, 3))
        ex = C.c_int(0)
        n = self.lib.L.orc_astar(self.h, _p(s), _p(e), _p(xyz), cap, C.byref(ex))
        return (None if n < 0 else xyz[:n].copy()), ex.value

    def shortcut(self, path, cap=16384):
        path = _d(path).reshape(-1, 3)
        out = np.zeros((cap, 3))
        n = self.lib.L.orc_shortcut(self.h, _p(path), len(path), _p(out), cap)
        return out[:n].copy()


def bspline_at(ctrl, t, degree=3, ts=0.2):
    ctrl = _d(ctrl).reshape(-1, 3)
    t = _d(np.atleast_1d(t))
    out = np.zeros((len(t), 3))
    lib().L.orc_bspline_at(_p(ctrl), len(ctrl), degree, ts, _p(t), len(t), _p(out))
    return out


def bspline_deriv_at(ctrl, t, d=1, ts=0.2):
    ctrl = _d(ctrl).reshape(-1, 3)
    t = _d(np.atleast_1d(t))
    out = np.zeros((len(t), 3))
    lib().L.orc_bspline_deriv_at(_p(ctrl), len(ctrl), ts, d, _p(t), len(t), _p(out))
    return out


def make_plan_batch(omap, params, offsets, ctrl_in, nthreads=1, want_ms=False):
    """CPU baseline: reference-style makePlan over a batch, one problem per host thread."""
    offsets = np.ascontiguousarray(offsets, dtype=np.int32)
    ctrl_in = _d(ctrl_in).reshape(-1, 3)
    B = len(offsets) - 1
    out = np.zeros_like(ctrl_in)
    stats = np.zeros(B, STATS_DTYPE)
    ms = np.zeros(B)
    ok = omap.lib.L.orc_make_plan_batch(omap.h, C.byref(params), B, _p(offsets, _ip), _p(ctrl_in), _p(out),
                                        stats.ctypes.data_as(C.c_void_p), int(nthreads), _p(ms))
    return (ok, out, stats, ms) if want_ms else (ok, out, stats)
