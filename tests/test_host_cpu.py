"""CPU tests of the PRODUCT's host side: the C-ABI library loads and exports every declared symbol,
the map loaders / front end / B-spline helpers (host code by design) agree with the numpy oracle, the
engine refuses to run without a GPU, and the multi-GPU sharding logic works under gloo (world_size 2)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_MAPS = "/root/reference/map"


def test_c_abi_exports_every_declared_symbol(tp):
    import ctypes
    hdr = open(os.path.join(ROOT, "include", "tp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(tp_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 35
    L = ctypes.CDLL(tp._capi.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing
    # and the binding's own list agrees with the header
    assert set(tp._capi.SYMBOLS) <= names


def test_engine_fails_loudly_without_gpu(tp):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert tp.load().tp_device_count() == 0
    with pytest.raises(tp.TpError) as ei:
        tp.Engine(0)
    assert "no CPU fallback" in str(ei.value)


def test_product_never_imports_the_oracle():
    for dp, _, fs in os.walk(os.path.join(ROOT, "trajectory_planner_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle|#include\s+\"[^\"]*oracle", src, flags=re.M), f


def test_tpm_roundtrip_and_map_contract(tp, sq_map, sq_omap, tmp_path):
    info = sq_map.info()
    assert info["dims"] == (400, 400, 30) and info["inflate"] == (4, 4, 2) and abs(info["res"] - 0.1) < 1e-15
    assert info["n_occupied"] == 102844   # every PCD point of square_static_map.pcd lands in its own voxel
    occ_o, known_o, inf_o = sq_omap.grids()
    assert np.array_equal(sq_map.grid("occupied"), occ_o)
    assert np.array_equal(sq_map.grid("inflated"), inf_o)   # product inflation == oracle inflation
    p = tmp_path / "m.tpm"
    sq_map.save_tpm(str(p))
    m2 = tp.OccMap.from_tpm(str(p))
    assert np.array_equal(m2.grid("inflated"), sq_map.grid("inflated"))
    assert np.array_equal(m2.grid("known"), sq_map.grid("known"))


def test_add_points_inflation_matches_oracle(tp, orc):
    rng = np.random.default_rng(0)
    m = tp.OccMap(0.1, (-2.0, -2.0, -0.1), (40, 40, 30), (4, 4, 2))
    om = orc.Map(0.1, (-2.0, -2.0, -0.1), (40, 40, 30), (4, 4, 2))
    pts = np.column_stack([rng.uniform(-2.5, 2.5, 300), rng.uniform(-2.5, 2.5, 300), rng.uniform(-0.3, 3.2, 300)])
    pts[:10] = np.round(pts[:10], 1)   # voxel faces
    m.add_points(pts)
    om.add_points(pts)
    o, k, i = om.grids()
    assert np.array_equal(m.grid("occupied"), o) and np.array_equal(m.grid("known"), k)
    assert np.array_equal(m.grid("inflated"), i)


@pytest.mark.skipif(not os.path.isdir(REF_MAPS), reason="reference maps not present on this box")
def test_loaders_against_numpy_oracle_on_reference_maps(tp, sq_map):
    from oracle import maps_np
    pts = maps_np.read_pcd_ascii(os.path.join(REF_MAPS, "square_static_map.pcd"))
    assert len(pts) == 102844
    m = tp.OccMap.from_pcd(os.path.join(REF_MAPS, "square_static_map.pcd"))
    assert np.array_equal(m.grid("occupied"), sq_map.grid("occupied"))   # committed raster == fresh load
    idx = np.floor((pts - np.array(m.info()["origin"])) / 0.1).astype(int)
    g = np.zeros((400, 400, 30), np.uint8)
    g[idx[:, 0], idx[:, 1], idx[:, 2]] = 1
    assert np.array_equal(g, m.grid("occupied"))
    # .bt: product loader vs the numpy parser, via the committed rasters
    for name in ("box", "maze"):
        mb = tp.OccMap.from_bt(os.path.join(REF_MAPS, name + ".bt"))
        mt = tp.OccMap.from_tpm(os.path.join(ROOT, "data", "maps", name + ".tpm"))
        assert mb.info()["dims"] == mt.info()["dims"]
        assert np.array_equal(mb.grid("occupied"), mt.grid("occupied"))
        res, leaves = maps_np.read_bt(os.path.join(REF_MAPS, name + ".bt"))
        cells = maps_np.bt_to_cells(leaves)
        n_occ = sum(len(c) for c, occ in cells if occ) if isinstance(cells, list) else None
        if n_occ is not None:
            assert n_occ == mb.info()["n_occupied"]


def test_bspline_fit_fixture_and_eval(tp, orc):
    """src/test/testBsplineFit.cpp:5-21 — 10 collinear points (0, 0.4 i, 1), zero boundary conditions,
    ts = 0.1.  parameterizeToBspline is a LEAST-SQUARES fit of K+4 equations (K samples + v0,v1,a0,a1)
    in K+2 unknowns (bspline.cpp:95-131), so the rest-to-rest conditions pull the spline off the
    constant-speed samples; the fit must be the least-squares optimum of that system."""
    pts = np.array([[0.0, 0.4 * i, 1.0] for i in range(10)])
    ts = 0.1
    c = tp.bspline_fit(ts, pts)
    assert c.shape == (12, 3)
    t = np.arange(10) * ts
    # x and z are reproduced exactly (they are constant); y is a compromise with the boundary rows
    ev = tp.bspline_eval(c, t, ts)
    assert np.max(np.abs(ev[:, [0, 2]] - pts[:, [0, 2]])) <= 1e-9
    assert np.max(np.abs(ev[:, 1] - pts[:, 1])) < 0.2
    # stationarity of the least-squares residual: A^T (A c - b) = 0 with the rows of bspline.cpp:95-127
    K = 10
    A = np.zeros((K + 4, K + 2))
    for i in range(K):
        A[i, i:i + 3] = np.array([1.0, 4.0, 1.0]) / 6.0
    A[K, 0:3] = np.array([-1.0, 0.0, 1.0]) / (2 * ts)
    A[K + 1, K - 1:K + 2] = np.array([-1.0, 0.0, 1.0]) / (2 * ts)
    A[K + 2, 0:3] = np.array([1.0, -2.0, 1.0]) / ts ** 2
    A[K + 3, K - 1:K + 2] = np.array([1.0, -2.0, 1.0]) / ts ** 2
    b = np.vstack([pts, np.zeros((4, 3))])
    assert np.max(np.abs(A.T @ (A @ c - b))) <= 1e-8
    # host eval == oracle de Boor, bit for bit
    tt = np.linspace(0, 9 * ts, 57)
    assert np.array_equal(tp.bspline_eval(c, tt, ts), orc.bspline_at(c, tt, ts=ts))
    # numpy restatement of parameterizeToBspline agrees to least-squares accuracy
    from oracle import frontend_np
    c_np = frontend_np.parameterize_to_bspline(ts, pts, np.zeros((4, 3)))
    assert np.max(np.abs(c - c_np)) <= 1e-9


def test_frontend_matches_numpy_restatement(tp, sq_map, sq_omap):
    """start/goal -> min-snap seed -> resample -> updatePath -> control points (src/bspline_node.cpp:332-371)."""
    from oracle import frontend_np
    from helpers import random_pairs
    rng = np.random.default_rng(5)
    S, G = random_pairs(sq_omap, 24, rng)
    p = tp.default_params()
    off, ctrl, valid = tp.frontend_batch(sq_map, p, S, G)
    checked = 0
    for b in range(len(S)):
        want = frontend_np.start_goal_to_ctrl(S[b], G[b], sq_omap)
        got = ctrl[off[b]:off[b + 1]]
        if want is None:
            assert not valid[b]
            continue
        assert valid[b] and got.shape == want.shape, (b, got.shape, want.shape)
        assert np.max(np.abs(got - want)) <= 1e-8
        checked += 1
    assert checked >= 20
    # goal inside an obstacle -> updatePath returns false (bsplineTraj.cpp:292-295)
    occ = np.argwhere(sq_map.grid("inflated")[:, :, 11] != 0)[0]
    bad = np.array(sq_map.info()["origin"]) + (np.array([occ[0], occ[1], 11]) + 0.5) * 0.1
    _, _, v2 = tp.frontend_batch(sq_map, p, S[:1], bad[None])
    assert v2[0] == 0


def test_shard_bounds_partition():
    from trajectory_planner_b200.sharding import shard_bounds, shard_batch
    for B in (0, 1, 7, 4096, 65536):
        for w in (1, 2, 3, 8):
            bd = shard_bounds(B, w)
            assert bd[0][0] == 0 and bd[-1][1] == B
            assert all(bd[i][1] == bd[i + 1][0] for i in range(w - 1))
            assert max(e - s for s, e in bd) <= -(-B // w)
    off = np.array([0, 7, 19, 30, 41, 60], np.int32)
    ctrl = np.arange(60 * 3, dtype=float).reshape(60, 3)
    parts = [shard_batch(off, ctrl, r, 2) for r in range(2)]
    assert np.array_equal(np.concatenate([p[1] for p in parts]), ctrl)
    assert parts[1][0][0] == 0 and parts[1][2] == (3, 5)


_WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from trajectory_planner_b200.sharding import shard_batch, gather_batch
from trajectory_planner_b200 import RESULT_DTYPE
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(3)
N = rng.integers(7, 40, 37)
off = np.concatenate([[0], np.cumsum(N)]).astype(np.int32)
ctrl = rng.normal(size=(off[-1], 3))
loc_off, loc_ctrl, (b0, b1) = shard_batch(off, ctrl, rank, world)
# stand-in for the per-rank solve (no GPU here): a deterministic function of the shard
res = np.zeros(b1 - b0, RESULT_DTYPE)
res["status"] = 1
res["lbfgs_iters"] = np.diff(loc_off)
res["final_cost"] = [loc_ctrl[loc_off[i]:loc_off[i + 1]].sum() for i in range(b1 - b0)]
out = gather_batch(loc_ctrl * 2.0, res, off)
full_ctrl, full_res = out
assert np.array_equal(full_ctrl, ctrl * 2.0)
assert np.array_equal(full_res["lbfgs_iters"], N)
assert np.allclose(full_res["final_cost"], [ctrl[off[i]:off[i + 1]].sum() for i in range(len(N))])
g0 = gather_batch(loc_ctrl, res, off, dst=0)
assert (g0 is None) == (rank != 0)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_sharded_gather_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def test_reference_arm_of_bench_runs_on_cpu():
    """bench.py --impl reference times the CPU path (oracle) and prints the contract's JSON line."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--ref-sample", "64"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "solves/s" and line["value"] > 0
    assert line["cpu_baseline"]["cores"] >= 1 and line["e2e"]["h2d_bytes_per_step"] == 0


def test_cpp_shim_compiles_and_fails_loudly_without_gpu(tp, tmp_path):
    """The reference-shaped C++ shim (include/trajectory_planner/bsplineTraj_b200.hpp) builds as C++14 against
    the C ABI, and the headless node exits with the 'no engine' code when there is no GPU."""
    exe = tmp_path / "node"
    r = subprocess.run(["g++", "-std=c++14", "-O1", "-Wall", "-Werror", os.path.join(ROOT, "examples", "bspline_node_headless.cpp"),
                        "-o", str(exe), "-L" + os.path.dirname(tp._capi.LIB_PATH), "-ltp_b200",
                        "-Wl,-rpath," + os.path.dirname(tp._capi.LIB_PATH)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    import torch
    if not torch.cuda.is_available():
        r2 = subprocess.run([str(exe), os.path.join(ROOT, "data", "maps", "square_static.tpm")], capture_output=True, text=True, timeout=120)
        assert r2.returncode == 3 and "no CPU fallback" in r2.stdout


def test_update_path_host_entry(tp, sq_map, sq_omap):
    """tp_vigo_update_path == the numpy restatement of bsplineTraj::updatePath on a resampled path."""
    from oracle import frontend_np
    p = tp.default_params()
    from helpers import random_pairs
    S, G = random_pairs(sq_omap, 1, np.random.default_rng(9), min_dist=6.0)
    path = S[0] + np.linspace(0, 1, int(np.linalg.norm(G[0] - S[0]) / 0.2) + 1)[:, None] * (G[0] - S[0])
    assert sq_omap.query(path[-1:])[0] == 0
    import ctypes as C
    out = np.zeros((len(path) + 1024, 3))
    n = tp.load().tp_vigo_update_path(sq_map.h, C.byref(p), len(path), tp._capi.ptr(path), None, tp._capi.ptr(out), len(out))
    assert n > 4
    inp = frontend_np.adjust_path_length_direct(list(path), sq_omap, p.max_path_length)
    want = frontend_np.parameterize_to_bspline(p.ctrl_pt_ts, np.array(inp), np.zeros((4, 3)))
    assert want.shape[0] == n and np.max(np.abs(out[:n] - want)) <= 1e-8


def test_pwl_fallback_and_poly_pose_match_the_restatement(tp):
    """pwlTraj (piecewiseLinearTraj.cpp) and polyTrajSolver::getPose through the C ABI (host side) against the Python
    restatement: knots, headings and poses, with and without caller-supplied yaw, large turns (> PI_const) included."""
    import ctypes as C
    from oracle import pwl_np as PW
    L = tp.load()
    rng = np.random.default_rng(3)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    worst = 0.0
    for trial in range(12):
        K = int(rng.integers(2, 9))
        path = np.ascontiguousarray(np.cumsum(rng.uniform(-3, 3, (K, 3)), 0))
        if trial == 3:
            path[2] = path[1]   # a zero-length segment (forward period shorter than 1e-3 s)
        yaw_in = np.ascontiguousarray(rng.uniform(-3.1, 3.1, K)) if trial % 2 else None
        yaw = np.zeros(K)
        knots = np.zeros(2 * K + 1)
        n = L.tp_pwl_plan(K, P(path), P(yaw_in) if yaw_in is not None else None, 1.0, 0.5, P(yaw), P(knots))
        yo, ko = PW.plan(path, yaw_in)
        assert n == len(ko) and np.array_equal(yaw, yo) and np.array_equal(knots[:n], ko)
        t = np.ascontiguousarray(np.concatenate([np.linspace(-0.1, ko[-1] + 0.2, 57), ko]))
        out = np.zeros((len(t), 4))
        assert L.tp_pwl_eval(K, P(path), P(yaw), n, P(knots), len(t), P(t), P(out)) == 0
        ref = np.array([PW.get_pose(path, yo, ko, float(x)) for x in t])
        worst = max(worst, float(np.max(np.abs(out - ref))))
    assert worst == 0.0
    # polyTrajSolver::getPose on an exact min-snap solution
    from oracle import frontend_np as F
    path = np.array([[0, 0, 1.0], [1, 1, 1.0], [2, 0, 1.0], [4, 10, 1.0]])   # src/test/waypoint.yaml:2-5
    coef, times = F.minsnap_solve(path, 1.0)
    cf = np.ascontiguousarray(coef.reshape(-1))
    t = np.ascontiguousarray(np.concatenate([np.linspace(0, times[-1], 41), times, [times[-1] + 1.0]]))
    out = np.zeros((len(t), 4))
    assert L.tp_poly_eval(len(times) - 1, P(cf), P(np.ascontiguousarray(times)), len(t), P(t), P(out)) == 0
    ref = np.array([PW.poly_get_pose(coef, times, float(x)) for x in t])
    assert np.max(np.abs(out - ref)) <= 1e-12
    assert np.all(out[-1] == 0)   # outside every knot interval: the default pose
