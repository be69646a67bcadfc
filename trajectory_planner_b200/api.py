"""Host-side mirror of the reference interface for the ViGO B-spline path, on top of the C ABI.

`OccMap` stands in for mapManager::occMap (prebuilt-map loading + the query contract), `Engine`
owns one GPU, and `BsplineTraj` keeps the method names of trajPlanner::bsplineTraj
(include/trajectory_planner/bsplineTraj.h:87-181): setMap / updateMaxVel / updateMaxAcc /
updatePath / updateDynamicObstacles / makePlan / getControlPoints / getPose / getDuration /
getLinearFactor / isCurrTrajValid, each a batch-of-one call into the same CUDA kernels the
batched entry point (`Engine.make_plan_batch`) uses.  Nothing here computes on the CPU except the
front end (path -> control points) and pose-at-time queries, which the design keeps on the host.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import (EngineCfg, MapInfo, VigoParams, RESULT_DTYPE, LBFGS_DTYPE, TP_MEM_DEVICE, TP_MEM_HOST, check, ptr)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def default_params():
    p = VigoParams()
    _capi.load().tp_vigo_default_params(C.byref(p))
    return p


class OccMap:
    """Host occupancy map honouring the occMap contract of include/tp_b200.h."""

    def __init__(self, res=0.1, origin=(-20.0, -20.0, -0.1), dims=(400, 400, 30), inflate=(4, 4, 2), _handle=None):
        self.L = _capi.load()
        if _handle is not None:
            self.h = _handle
        else:
            o, d, i = _f64(origin), _i32(dims), _i32(inflate)
            self.h = self.L.tp_map_create(float(res), o.ctypes.data_as(_capi._dp), d.ctypes.data_as(_capi._ip),
                                          i.ctypes.data_as(_capi._ip))
            if not self.h:
                check(-1, "tp_map_create")

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.tp_map_destroy(self.h)
                self.h = None
        except Exception:
            pass

    @staticmethod
    def inflate_cells(robot_size=(0.8, 0.8, 0.3), res=0.1):
        """ceil(robot_size / (2 res)) per axis (cfg/bspline_interactive/occupancy_map.yaml:9,36)."""
        return tuple(int(np.ceil(s / (2.0 * res))) for s in robot_size)

    @classmethod
    def from_pcd(cls, path, res=0.1, map_size=(40.0, 40.0, 3.0), ground_height=-0.1, robot_size=(0.8, 0.8, 0.3)):
        """The reference's prebuilt-map set-up (occupancy_map.yaml:36-38,50)."""
        origin = (-map_size[0] / 2.0, -map_size[1] / 2.0, ground_height)
        dims = tuple(int(np.ceil(s / res)) for s in map_size)
        m = cls(res, origin, dims, cls.inflate_cells(robot_size, res))
        check(m.L.tp_map_load_pcd(m.h, str(path).encode()), "tp_map_load_pcd")
        return m

    @classmethod
    def from_bt(cls, path, robot_size=(0.8, 0.8, 0.3), pad=1.0, occupied_only_bbox=True):
        """Rasterise an OctoMap .bt at its native resolution into the occMap contract."""
        L = _capi.load()
        res = C.c_double(0)
        mn, mx = np.zeros(3), np.zeros(3)
        check(L.tp_bt_bbox(str(path).encode(), C.byref(res), mn.ctypes.data_as(_capi._dp), mx.ctypes.data_as(_capi._dp),
                           1 if occupied_only_bbox else 0), "tp_bt_bbox")
        r = res.value
        lo = np.floor((mn - pad) / r) * r
        dims = np.ceil((mx + pad - lo) / r).astype(int)
        m = cls(r, tuple(lo), tuple(int(v) for v in dims), cls.inflate_cells(robot_size, r))
        check(L.tp_map_load_bt(m.h, str(path).encode()), "tp_map_load_bt")
        return m

    @classmethod
    def from_tpm(cls, path, robot_size=(0.8, 0.8, 0.3), inflate=None):
        L = _capi.load()
        if inflate is None:
            # res is in the file: peek at it
            with open(path, "rb") as f:
                hdr = f.read(16)
            res = float(np.frombuffer(hdr[8:16], dtype=np.float64)[0])
            inflate = cls.inflate_cells(robot_size, res)
        inf = _i32(inflate)
        h = L.tp_map_load_tpm(str(path).encode(), inf.ctypes.data_as(_capi._ip))
        if not h:
            check(-1, "tp_map_load_tpm")
        return cls(_handle=h)

    def save_tpm(self, path):
        check(self.L.tp_map_save_tpm(self.h, str(path).encode()), "tp_map_save_tpm")

    def add_points(self, xyz):
        xyz = _f64(xyz).reshape(-1, 3)
        check(self.L.tp_map_add_points(self.h, xyz.ctypes.data_as(_capi._dp), len(xyz)), "tp_map_add_points")

    def add_cells(self, ijk, occupied=True):
        ijk = _i32(ijk).reshape(-1, 3)
        check(self.L.tp_map_add_cells(self.h, ijk.ctypes.data_as(_capi._ip), len(ijk), 1 if occupied else 0),
              "tp_map_add_cells")

    def info(self):
        mi = MapInfo()
        check(self.L.tp_map_info_get(self.h, C.byref(mi)), "tp_map_info_get")
        return dict(res=mi.res, origin=tuple(mi.origin), dims=tuple(mi.dims), inflate=tuple(mi.inflate),
                    n_occupied=mi.n_occupied, n_inflated=mi.n_inflated, n_known=mi.n_known,
                    packed_bytes=mi.packed_bytes)

    def grid(self, which="inflated"):
        w = {"occupied": 0, "known": 1, "inflated": 2}[which]
        d = self.info()["dims"]
        out = np.zeros(int(np.prod(d)), np.uint8)
        check(self.L.tp_map_get_grid(self.h, w, out.ctypes.data_as(_capi._u8p)), "tp_map_get_grid")
        return out.reshape(d)


def frontend_batch(omap, params, starts, goals):
    """start/goal pairs -> (offsets[B+1], ctrl[sum N, 3], valid[B]) — src/bspline_node.cpp:332-371."""
    L = _capi.load()
    starts, goals = _f64(starts).reshape(-1, 3), _f64(goals).reshape(-1, 3)
    B = len(starts)
    cap = 1024 * max(B, 1)
    off = np.zeros(B + 1, np.int32)
    ctrl = np.zeros((cap, 3))
    valid = np.zeros(B, np.uint8)
    n = L.tp_vigo_frontend_batch(omap.h, C.byref(params), B, ptr(starts), ptr(goals), ptr(off), ptr(ctrl), cap, ptr(valid))
    check(int(n), "tp_vigo_frontend_batch")
    return off, ctrl[:n].copy(), valid


def bspline_fit(ts, points, start_end=None):
    pts = _f64(points).reshape(-1, 3)
    se = _f64(np.zeros((4, 3)) if start_end is None else start_end).reshape(4, 3)
    out = np.zeros((len(pts) + 2, 3))
    check(_capi.load().tp_bspline_fit(float(ts), len(pts), ptr(pts), ptr(se), ptr(out)), "tp_bspline_fit")
    return out


def bspline_eval(ctrl, t, ts=0.2, deriv=0):
    ctrl = _f64(ctrl).reshape(-1, 3)
    t = _f64(np.atleast_1d(t))
    out = np.zeros((len(t), 3))
    check(_capi.load().tp_bspline_eval(len(ctrl), ptr(ctrl), float(ts), int(deriv), len(t), ptr(t), ptr(out)),
          "tp_bspline_eval")
    return out


class Engine:
    """One GPU: bit-packed map replica in HBM, A* node pools, streams and scratch."""

    def __init__(self, device=0, cfg=None, **cfg_kw):
        self.L = _capi.load()
        c = EngineCfg()
        self.L.tp_engine_default_cfg(C.byref(c))
        if cfg is not None:
            c = cfg
        for k, v in cfg_kw.items():
            setattr(c, k, v)
        self.cfg = c
        self.h = self.L.tp_engine_create(int(device), C.byref(c))
        if not self.h:
            check(-1, "tp_engine_create")
        self.device = int(device)
        self.map = None

    def close(self):
        if getattr(self, "h", None):
            self.L.tp_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_map(self, omap):
        check(self.L.tp_engine_set_map(self.h, omap.h), "tp_engine_set_map")
        self.map = omap

    def synchronize(self):
        check(self.L.tp_engine_synchronize(self.h), "tp_engine_synchronize")

    @property
    def launch_count(self):
        return int(self.L.tp_engine_launch_count(self.h))

    @property
    def stream(self):
        return self.L.tp_engine_stream(self.h)

    # ---- measurement
    def profile_enable(self, on=True):
        check(self.L.tp_engine_profile_enable(self.h, 1 if on else 0), "tp_engine_profile_enable")

    def profile_get(self):
        pr = _capi.Profile()
        check(self.L.tp_engine_profile_get(self.h, C.byref(pr)), "tp_engine_profile_get")
        out = dict(lbfgs_flops=pr.lbfgs_flops, lbfgs_iters=pr.lbfgs_iters, lbfgs_evals=pr.lbfgs_evals,
                   check_samples=pr.check_samples, query_points=pr.query_points, ms={}, launches={})
        for i, k in enumerate(_capi.PROF_KINDS):
            out["ms"][k] = pr.ms[i]
            out["launches"][k] = int(pr.launches[i])
        return out

    def microbench_fp64(self):
        v = C.c_double(0)
        check(self.L.tp_microbench_fp64(self.h, C.byref(v)), "tp_microbench_fp64")
        return v.value

    def microbench_gather(self, nbytes):
        v = C.c_double(0)
        check(self.L.tp_microbench_gather(self.h, int(nbytes), C.byref(v)), "tp_microbench_gather")
        return v.value

    # ---- map queries
    def query_points(self, xyz):
        xyz = _f64(xyz).reshape(-1, 3)
        out = np.zeros(len(xyz), np.uint8)
        check(self.L.tp_query_points(self.h, len(xyz), ptr(xyz), ptr(out), TP_MEM_HOST, None), "tp_query_points")
        return out

    def query_unknown(self, xyz):
        xyz = _f64(xyz).reshape(-1, 3)
        out = np.zeros(len(xyz), np.uint8)
        check(self.L.tp_query_unknown(self.h, len(xyz), ptr(xyz), ptr(out), TP_MEM_HOST, None), "tp_query_unknown")
        return out

    def query_lines(self, a, b):
        a, b = _f64(a).reshape(-1, 3), _f64(b).reshape(-1, 3)
        out = np.zeros(len(a), np.uint8)
        check(self.L.tp_query_lines(self.h, len(a), ptr(a), ptr(b), ptr(out), TP_MEM_HOST, None), "tp_query_lines")
        return out

    def query_points_device(self, n, xyz_ptr, out_ptr, stream=None):
        """Device pointers (e.g. torch tensors' data_ptr()); enqueues one kernel on `stream`."""
        check(self.L.tp_query_points(self.h, int(n), C.c_void_p(xyz_ptr), C.c_void_p(out_ptr), TP_MEM_DEVICE,
                                     C.c_void_p(stream) if stream else None), "tp_query_points")

    # ---- ViGO pieces
    @staticmethod
    def _guides(guides, B):
        if guides is None:
            return None, None, None, None
        g_off, g_cp, g_p, g_v = guides
        return _i32(g_off), _i32(g_cp), _f64(g_p).reshape(-1, 3), _f64(g_v).reshape(-1, 3)

    @staticmethod
    def _dyn(dyn):
        if dyn is None:
            return 0, None, None, None
        pos, vel, size = (_f64(a).reshape(-1, 3) for a in dyn)
        return len(pos), pos, vel, size

    def cost_batch(self, params, offsets, ctrl, guides=None, weights=None, dyn=None):
        """costFunction per trajectory; dyn = (pos, vel, size) of the dynamic obstacles (updateDynamicObstacles)."""
        offsets, ctrl = _i32(offsets), _f64(ctrl).reshape(-1, 3)
        B = len(offsets) - 1
        g_off, g_cp, g_p, g_v = self._guides(guides, B)
        w = None if weights is None else _f64(weights).reshape(B, 2)
        nd, dp, dv, ds = self._dyn(dyn)
        f = np.zeros(B)
        nvar = max(3 * (len(ctrl) - 6 * B), 1)
        grad = np.zeros(nvar)
        check(self.L.tp_vigo_cost_batch_dyn(self.h, C.byref(params), B, ptr(offsets), ptr(ctrl), ptr(g_off), ptr(g_cp),
                                            ptr(g_p), ptr(g_v), ptr(w), nd, ptr(dp), ptr(dv), ptr(ds), ptr(f), ptr(grad),
                                            TP_MEM_HOST, None), "tp_vigo_cost_batch_dyn")
        return f, grad

    def optimize_batch(self, params, offsets, ctrl, guides=None, weights=None, want_x=True, dyn=None):
        offsets, ctrl = _i32(offsets), _f64(ctrl).reshape(-1, 3).copy()
        B = len(offsets) - 1
        g_off, g_cp, g_p, g_v = self._guides(guides, B)
        w = None if weights is None else _f64(weights).reshape(B, 2)
        res = np.zeros(B, LBFGS_DTYPE)
        nvar = max(3 * (len(ctrl) - 6 * B), 1)
        xf = np.zeros(nvar) if want_x else None
        nd, dp, dv, ds = self._dyn(dyn)
        check(self.L.tp_vigo_optimize_batch_dyn(self.h, C.byref(params), B, ptr(offsets), ptr(ctrl), ptr(g_off), ptr(g_cp),
                                                ptr(g_p), ptr(g_v), ptr(w), nd, ptr(dp), ptr(dv), ptr(ds), ptr(res), ptr(xf),
                                                TP_MEM_HOST, None), "tp_vigo_optimize_batch_dyn")
        return ctrl, res, xf

    def sample_batch(self, offsets, ctrl, times=None, dt=None, ts=0.2, vel=True, acc=False, yaw=True):
        """Batched pose-at-time queries (tp_vigo_sample_batch): `times` = list of per-trajectory time arrays, or `dt` for
        evalTraj's accumulated grid t = 0; t <= duration; t += dt (bsplineTraj.cpp:1438-1447) per trajectory.
        -> dict(t_offsets, t, pos [T,3], vel, acc, yaw) (None for the outputs not asked for)."""
        offsets, ctrl = _i32(offsets), _f64(ctrl).reshape(-1, 3)
        B = len(offsets) - 1
        if times is None:
            times = []
            for b in range(B):
                dur = (float(offsets[b + 1] - offsets[b]) - 3.0) * ts
                tt, t = [], 0.0
                while t <= dur:
                    tt.append(t)
                    t += dt
                times.append(tt)
        t_off = np.zeros(B + 1, np.int32)
        t_off[1:] = np.cumsum([len(x) for x in times])
        t = _f64(np.concatenate([np.asarray(x, float).ravel() for x in times]) if B else np.zeros(0))
        T = len(t)
        pos = np.zeros((T, 3))
        v = np.zeros((T, 3)) if vel else None
        a = np.zeros((T, 3)) if acc else None
        y = np.zeros(T) if yaw else None
        check(self.L.tp_vigo_sample_batch(self.h, float(ts), B, ptr(offsets), ptr(ctrl), ptr(t_off), ptr(t), ptr(pos), ptr(v), ptr(a),
                                          ptr(y), TP_MEM_HOST, None), "tp_vigo_sample_batch")
        return dict(t_offsets=t_off, t=t, pos=pos, vel=v, acc=a, yaw=y)

    def has_collision_batch(self, params, offsets, ctrl):
        offsets, ctrl = _i32(offsets), _f64(ctrl).reshape(-1, 3)
        B = len(offsets) - 1
        hit = np.zeros(B, np.uint8)
        check(self.L.tp_vigo_has_collision_batch(self.h, C.byref(params), B, ptr(offsets), ptr(ctrl), ptr(hit),
                                                 TP_MEM_HOST, None), "tp_vigo_has_collision_batch")
        return hit

    def find_collision_seg_batch(self, params, offsets, ctrl):
        offsets, ctrl = _i32(offsets), _f64(ctrl).reshape(-1, 3)
        B = len(offsets) - 1
        ms = self.cfg.max_segments
        nseg = np.zeros(B, np.int32)
        segs = np.zeros((B, ms, 2), np.int32)
        check(self.L.tp_vigo_find_collision_seg_batch(self.h, C.byref(params), B, ptr(offsets), ptr(ctrl), ptr(nseg),
                                                      ptr(segs), TP_MEM_HOST, None), "tp_vigo_find_collision_seg_batch")
        return [segs[b, :nseg[b]].copy() for b in range(B)]

    def astar_batch(self, params, starts, ends):
        starts, ends = _f64(starts).reshape(-1, 3), _f64(ends).reshape(-1, 3)
        S = len(starts)
        pc = self.cfg.max_path_cells
        plen = np.zeros(S, np.int32)
        paths = np.zeros((S, pc, 3))
        ex = np.zeros(S, np.int32)
        check(self.L.tp_astar_batch(self.h, C.byref(params), S, ptr(starts), ptr(ends), ptr(plen), ptr(paths), ptr(ex),
                                    TP_MEM_HOST, None), "tp_astar_batch")
        return [None if plen[s] < 0 else paths[s, :plen[s]].copy() for s in range(S)], ex

    def init_guides_batch(self, params, offsets, ctrl):
        offsets, ctrl = _i32(offsets), _f64(ctrl).reshape(-1, 3)
        B = len(offsets) - 1
        ms, gc = self.cfg.max_segments, self.cfg.max_guide_pairs
        ok = np.zeros(B, np.uint8)
        nseg = np.zeros(B, np.int32)
        segs = np.zeros((B, ms, 2), np.int32)
        gcount = np.zeros(B, np.int32)
        g_cp = np.zeros((B, gc), np.int32)
        g_p = np.zeros((B, gc, 3))
        g_v = np.zeros((B, gc, 3))
        check(self.L.tp_vigo_init_guides_batch(self.h, C.byref(params), B, ptr(offsets), ptr(ctrl), ptr(ok), ptr(nseg),
                                               ptr(segs), ptr(gcount), ptr(g_cp), ptr(g_p), ptr(g_v)),
              "tp_vigo_init_guides_batch")
        out = []
        for b in range(B):
            n = gcount[b]
            out.append(dict(ok=bool(ok[b]), segs=segs[b, :nseg[b]].copy(), cp=g_cp[b, :n].copy(), p=g_p[b, :n].copy(),
                            v=g_v[b, :n].copy()))
        return out

    # ---- front end on the device: (start, goal) pairs -> initial control points (src/bspline_node.cpp:332-371)
    def frontend_batch(self, params, starts, goals):
        """Same contract as the module-level (host) frontend_batch, computed by one CUDA kernel on this engine's map."""
        starts, goals = _f64(starts).reshape(-1, 3), _f64(goals).reshape(-1, 3)
        B = len(starts)
        cap = 162 * max(B, 1)
        off = np.zeros(B + 1, np.int32)
        ctrl = np.zeros((cap, 3))
        valid = np.zeros(B, np.uint8)
        n = self.L.tp_vigo_frontend_batch_device(self.h, C.byref(params), B, ptr(starts), ptr(goals), ptr(off), ptr(ctrl), cap,
                                                 ptr(valid), TP_MEM_HOST, None)
        check(int(n), "tp_vigo_frontend_batch_device")
        return off, ctrl[:n].copy(), valid

    # ---- the batched entry point
    def make_plan_batch(self, params, offsets, ctrl, dyn=None):
        """Host buffers in, host buffers out (H2D + solve + D2H inside the call)."""
        offsets, ctrl = _i32(offsets), _f64(ctrl).reshape(-1, 3)
        B = len(offsets) - 1
        out = np.empty_like(ctrl)
        res = np.zeros(B, RESULT_DTYPE)
        n_dyn, dp, dv, ds = 0, None, None, None
        if dyn is not None:
            dp, dv, ds = (_f64(a).reshape(-1, 3) for a in dyn)
            n_dyn = len(dp)
        check(self.L.tp_vigo_make_plan_batch(self.h, C.byref(params), B, ptr(offsets), ptr(ctrl), ptr(out), ptr(res),
                                             n_dyn, ptr(dp), ptr(dv), ptr(ds), TP_MEM_HOST, None),
              "tp_vigo_make_plan_batch")
        return out, res

    def make_plan_batch_device(self, params, B, offsets_ptr, ctrl_in_ptr, ctrl_out_ptr, results_ptr, stream=None):
        """Device pointers (inputs already resident in HBM); results stay on the device."""
        check(self.L.tp_vigo_make_plan_batch(self.h, C.byref(params), int(B), C.c_void_p(offsets_ptr),
                                             C.c_void_p(ctrl_in_ptr), C.c_void_p(ctrl_out_ptr), C.c_void_p(results_ptr),
                                             0, None, None, None, TP_MEM_DEVICE, C.c_void_p(stream) if stream else None),
              "tp_vigo_make_plan_batch")


def make_plan_batch_multi(engines, params, offsets, ctrl, dyn=None, chunk=0):
    """One batch on several engines of this host (one per GPU, the map set on each): tp_vigo_make_plan_batch_multi.
    -> (ctrl_out, results, engine_of)."""
    L = _capi.load()
    offsets, ctrl = _i32(offsets), _f64(ctrl).reshape(-1, 3)
    B = len(offsets) - 1
    out = np.empty_like(ctrl)
    res = np.zeros(B, RESULT_DTYPE)
    who = np.full(B, -1, np.int32)
    n_dyn, dp, dv, ds = 0, None, None, None
    if dyn is not None:
        dp, dv, ds = (_f64(a).reshape(-1, 3) for a in dyn)
        n_dyn = len(dp)
    arr = (C.c_void_p * len(engines))(*[e.h for e in engines])
    check(L.tp_vigo_make_plan_batch_multi(arr, len(engines), C.byref(params), B, ptr(offsets), ptr(ctrl), ptr(out), ptr(res),
                                          n_dyn, ptr(dp), ptr(dv), ptr(ds), int(chunk), ptr(who)), "tp_vigo_make_plan_batch_multi")
    return out, res, who


def default_poly_params():
    p = _capi.PolyParams()
    _capi.load().tp_poly_default_params(C.byref(p))
    return p


def _poly_split(wp_off, coef, times):
    """flat coef/times -> per-path (coef[3, 8K], times[K+1])."""
    out = []
    for b in range(len(wp_off) - 1):
        K = wp_off[b + 1] - wp_off[b] - 1
        base = 24 * (wp_off[b] - b)
        out.append((coef[base:base + 24 * K].reshape(3, 8 * K) if K > 0 else np.zeros((3, 0)), times[wp_off[b]:wp_off[b + 1]]))
    return out


class PolyTraj:
    """Batched min-snap path of trajPlanner::polyTrajOctomap / polyTrajSolver on one engine (secondary path)."""

    def __init__(self, engine, params=None):
        self.engine = engine
        self.params = params if params is not None else default_poly_params()

    @staticmethod
    def _flat(paths):
        off = np.concatenate([[0], np.cumsum([len(p) for p in paths])]).astype(np.int32)
        wp = np.concatenate([_f64(p).reshape(-1, 3) for p in paths], 0) if len(paths) else np.zeros((0, 3))
        return off, np.ascontiguousarray(wp)

    def solve_batch(self, paths, bc=None):
        """polyTrajSolver::solve for a list of waypoint arrays -> list of (coef[3, 8K], times[K+1]), status[B]."""
        off, wp = self._flat(paths)
        B = len(paths)
        coef = np.zeros(24 * max(int(off[-1]) - B, 1))
        times = np.zeros(max(int(off[-1]), 1))
        status = np.zeros(B, np.int32)
        bcf = None if bc is None else _f64(bc).reshape(B, 12)
        check(self.engine.L.tp_minsnap_solve_batch(self.engine.h, C.byref(self.params), B, ptr(off), ptr(wp), ptr(bcf), ptr(coef),
                                                   ptr(times), ptr(status)), "tp_minsnap_solve_batch")
        return _poly_split(off, coef, times), status

    def check_batch(self, paths, sols, want_samples=False, samp_cap=2048):
        off, wp = self._flat(paths)
        B = len(paths)
        coef = np.concatenate([s[0].ravel() for s in sols]) if B else np.zeros(0)
        times = np.concatenate([s[1] for s in sols]) if B else np.zeros(0)
        valid = np.zeros(B, np.uint8)
        seg = np.zeros(max(int(off[-1]) - B, 1), np.uint8)
        ns = np.zeros(B, np.int32)
        samples = np.zeros((B, samp_cap, 3)) if want_samples else None
        shit = np.zeros((B, samp_cap), np.uint8) if want_samples else None
        check(self.engine.L.tp_poly_check_batch(self.engine.h, C.byref(self.params), B, ptr(off), ptr(wp), ptr(_f64(coef)), ptr(_f64(times)),
                                                ptr(valid), ptr(seg), ptr(ns), ptr(samples), ptr(shit), samp_cap if want_samples else 0),
              "tp_poly_check_batch")
        segs = [np.nonzero(seg[off[b] - b: off[b + 1] - b - 1])[0] for b in range(B)]
        if want_samples:
            return valid, segs, ns, [samples[b, :ns[b]] for b in range(B)], [shit[b, :ns[b]] for b in range(B)]
        return valid, segs, ns

    def box_collision(self, xyz):
        xyz = _f64(xyz).reshape(-1, 3)
        out = np.zeros(len(xyz), np.uint8)
        check(self.engine.L.tp_poly_box_collision(self.engine.h, C.byref(self.params), len(xyz), ptr(xyz), ptr(out)), "tp_poly_box_collision")
        return out

    def corridor_solve_batch(self, paths, corridor_size, corridor_res=8.0, bc=None):
        """polyTrajSolver::solve with corridor constraints (tp_corridor_solve_batch): corridor_size = list of per-segment radii.
        -> ([(coef [3, 8K], times)], status [B, 3])."""
        off, wp = self._flat(paths)
        B = len(paths)
        r = _f64(np.concatenate([np.asarray(x, float).ravel() for x in corridor_size]))
        bcv = None if bc is None else _f64(bc).reshape(B, 12)
        coef = np.zeros(max(len(wp) - B, 1) * 24)
        times = np.zeros(len(wp))
        status = np.zeros((B, 3), np.int32)
        check(self.engine.L.tp_corridor_solve_batch(self.engine.h, C.byref(self.params), B, ptr(off), ptr(wp), ptr(bcv), ptr(r),
                                                    float(corridor_res), ptr(coef), ptr(times), ptr(status)), "tp_corridor_solve_batch")
        return _poly_split(off, coef, times), status

    def make_plan_corridor_batch(self, paths, init_r=0.5, fs=0.8, corridor_res=8.0, bc=None, occmap=False, corridor_constraint=True):
        """polyTrajOctomap::makePlanCorridorConstraint for a list of waypoint arrays (occmap=True: polyTrajOccMap::makePlan on
        the ViGO occupancy map, optionally without corridor constraints) -> list of dict(valid, iters, coef, times, r, status)."""
        off, wp = self._flat(paths)
        B = len(paths)
        bcv = None if bc is None else _f64(bc).reshape(B, 12)
        nseg = max(len(wp) - B, 1)
        coef = np.zeros(nseg * 24)
        times = np.zeros(len(wp))
        valid = np.zeros(B, np.uint8)
        iters = np.zeros(B, np.int32)
        r = np.zeros(nseg)
        status = np.zeros((B, 3), np.int32)
        if occmap:
            check(self.engine.L.tp_polytraj_occmap_plan_batch(self.engine.h, C.byref(self.params), B, ptr(off), ptr(wp), ptr(bcv),
                                                              1 if corridor_constraint else 0, float(init_r), float(fs), float(corridor_res),
                                                              ptr(coef), ptr(times), ptr(valid), ptr(iters), ptr(r), ptr(status)),
                  "tp_polytraj_occmap_plan_batch")
        else:
            check(self.engine.L.tp_polytraj_corridor_plan_batch(self.engine.h, C.byref(self.params), B, ptr(off), ptr(wp), ptr(bcv), float(init_r),
                                                                float(fs), float(corridor_res), ptr(coef), ptr(times), ptr(valid), ptr(iters),
                                                                ptr(r), ptr(status)), "tp_polytraj_corridor_plan_batch")
        sols = _poly_split(off, coef, times)
        return [dict(valid=bool(valid[b]), iters=int(iters[b]), coef=sols[b][0], times=sols[b][1], r=r[off[b] - b:off[b + 1] - b - 1].copy(),
                     status=status[b].copy()) for b in range(B)]

    def make_plan_batch(self, paths, bc=None):
        """polyTrajOctomap::makePlanAddingWaypoint for a list of waypoint arrays (bc: optional [B, 12] v0, v1, a0, a1) ->
        list of dict(valid, iters, path, coef, times)."""
        off, wp = self._flat(paths)
        B = len(paths)
        bcv = None if bc is None else _f64(bc).reshape(B, 12)
        cap = B * 64
        off_o = np.zeros(B + 1, np.int32)
        wp_o = np.zeros((cap, 3))
        coef = np.zeros(24 * cap)
        times = np.zeros(cap)
        valid = np.zeros(B, np.uint8)
        iters = np.zeros(B, np.int32)
        check(self.engine.L.tp_polytraj_make_plan_batch_bc(self.engine.h, C.byref(self.params), B, ptr(off), ptr(wp), ptr(bcv), ptr(off_o),
                                                           ptr(wp_o), cap, ptr(coef), ptr(times), ptr(valid), ptr(iters)),
              "tp_polytraj_make_plan_batch_bc")
        sols = _poly_split(off_o, coef, times)
        return [dict(valid=bool(valid[b]), iters=int(iters[b]), path=wp_o[off_o[b]:off_o[b + 1]].copy(), coef=sols[b][0], times=sols[b][1])
                for b in range(B)]


class BsplineTraj:
    """Method-for-method stand-in for trajPlanner::bsplineTraj (bsplineTraj.h:87-181) for one
    trajectory: the planner state lives on the host, every heavy step runs on the GPU engine."""

    def __init__(self, engine, params=None):
        self.engine = engine
        self.params = params if params is not None else default_params()
        self.ctrl = None          # optData_.controlPoints: the working points
        self.traj = None          # bspline_: the committed trajectory, replaced only by a successful makePlan (bsplineTraj.cpp:376-377)
        self.init_ = False
        self.linear_factor = 1.0
        self.dyn = None
        self.last_result = None

    def setMap(self, omap):
        self.engine.set_map(omap)

    def updateMaxVel(self, v):
        self.params.max_vel = float(v)

    def updateMaxAcc(self, a):
        self.params.max_acc = float(a)

    def getInitTs(self):
        return self.params.ctrl_pt_dist / self.params.max_vel

    def getControlPointTs(self):
        return self.params.ctrl_pt_ts

    def getControlPointDist(self):
        return self.params.ctrl_pt_dist

    def updateControlPoints(self, ctrl):
        """Directly set optData_.controlPoints (what updatePath leaves behind, bsplineTraj.cpp:315-319)."""
        self.ctrl = _f64(ctrl).reshape(-1, 3).copy()
        self.dyn = None
        self.init_ = True
        return True

    def updatePathFromStartGoal(self, start, goal):
        """src/bspline_node.cpp:332-371: seed min-snap path, inputPathCheck loop, updatePath."""
        off, ctrl, valid = frontend_batch(self.engine.map, self.params, [start], [goal])
        if not valid[0]:
            return False
        return self.updateControlPoints(ctrl)

    def inputPathCheck(self, path):
        """bsplineTraj::inputPathCheck (bsplineTraj.cpp:207-245) -> (satisfied, adjusted_path)."""
        path = _f64(path).reshape(-1, 3)
        adj = np.zeros((max(4 * len(path), 64) + 4096, 3))
        n = C.c_int32(0)
        rc = check(_capi.load().tp_vigo_input_path_check(self.engine.map.h, C.byref(self.params), len(path), ptr(path),
                                                         ptr(adj), len(adj), C.byref(n)), "tp_vigo_input_path_check")
        return bool(rc), adj[:n.value].copy()

    def updatePath(self, path, start_end_conditions=None):
        """bsplineTraj::updatePath (bsplineTraj.cpp:290-323): goal check, adjustPathLengthDirect, fillPath,
        parameterizeToBspline — the reference's host-side code path, kept on the host."""
        path = _f64(path).reshape(-1, 3)
        if self.engine.map is None:
            raise _capi.TpError("setMap first")
        se = None if start_end_conditions is None else _f64(start_end_conditions).reshape(4, 3)
        out = np.zeros((len(path) + 1024, 3))
        n = check(_capi.load().tp_vigo_update_path(self.engine.map.h, C.byref(self.params), len(path), ptr(path), ptr(se),
                                                   ptr(out), len(out)), "tp_vigo_update_path")
        if n <= 0:
            return False
        return self.updateControlPoints(out[:n])

    def updateDynamicObstacles(self, pos, vel, size):
        self.dyn = (pos, vel, size)

    def makePlan(self):
        if not self.init_:
            return False
        off = np.array([0, len(self.ctrl)], np.int32)
        out, res = self.engine.make_plan_batch(self.params, off, self.ctrl, self.dyn)
        self.last_result = res[0]
        self.ctrl = out
        if res[0]["status"] == _capi.TP_STATUS_SUCCESS:
            self.traj = out.copy()   # a failed replan leaves the previous trajectory (and its factor) in place, as the reference does
            self.linear_factor = float(res[0]["linear_factor"])
            return True
        return False

    def getControlPoints(self):
        return self.ctrl.T.copy()  # 3 x N like the reference's Eigen::MatrixXd

    def getTrajectoryControlPoints(self):
        return None if self.traj is None else self.traj.T.copy()

    def getDuration(self):
        return 0.0 if self.traj is None else (len(self.traj) - 3) * self.params.ctrl_pt_ts

    def getTimestep(self):
        return self.params.ts

    def getLinearFactor(self):
        return self.linear_factor

    def getLinearReparamTime(self, t):
        return self.linear_factor * t

    def getPose(self, t, yaw=True):
        """-> (x, y, z, yaw) of the COMMITTED trajectory — bsplineTraj.cpp:1402-1419 (one trajectory, one time: on the host;
        batches go through Engine.sample_batch)."""
        if self.traj is None:
            return 0.0, 0.0, 0.0, 0.0
        p = bspline_eval(self.traj, [t], self.params.ctrl_pt_ts, 0)[0]
        if not yaw:
            return p[0], p[1], p[2], 0.0
        v = bspline_eval(self.traj, [t], self.params.ctrl_pt_ts, 1)[0]
        return p[0], p[1], p[2], float(np.arctan2(v[1], v[0]))

    def isCurrTrajValid(self):
        if not self.init_:
            return False
        off = np.array([0, len(self.ctrl)], np.int32)
        return not bool(self.engine.has_collision_batch(self.params, off, self.ctrl)[0])
