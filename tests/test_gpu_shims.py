"""The drop-in C++ shims (include/trajectory_planner/*_b200.hpp: bsplineTraj, bspline, polyTrajOctomap + pwlTraj) actually
RUN on the GPU: examples/shim_selftest.cpp is compiled against libtp_b200.so, executed, and everything it prints is compared
with the Python API (same C ABI underneath) on the same inputs — bit for bit where both sides call the same entry point."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _parse(text):
    out = {}
    for line in text.splitlines():
        k, _, rest = line.partition(" ")
        out.setdefault(k, []).append(rest)
    return out


def test_cpp_shims_run_on_the_gpu_and_match_the_python_api(tp, engine, sq_map, tmp_path):
    exe = tmp_path / "shim_selftest"
    libdir = os.path.dirname(tp._capi.LIB_PATH)
    r = subprocess.run(["g++", "-std=c++14", "-O1", "-Wall", "-Werror", os.path.join(ROOT, "examples", "shim_selftest.cpp"), "-o", str(exe),
                        "-L" + libdir, "-ltp_b200", "-Wl,-rpath," + libdir], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    field_tpm = os.path.join(ROOT, "data", "maps", "field.tpm")
    fm = tp.OccMap.from_tpm(field_tpm)
    info = fm.info()
    # waypoints inside field.bt's known region at z = 1
    lo = np.array(info["origin"]) + 0.25 * np.array(info["dims"]) * info["res"]
    hi = np.array(info["origin"]) + 0.75 * np.array(info["dims"]) * info["res"]
    rng = np.random.default_rng(5)
    wps = np.column_stack([np.linspace(lo[0], hi[0], 5), rng.uniform(lo[1], hi[1], 5), np.full(5, 1.0)])
    args = [str(exe), os.path.join(ROOT, "data", "maps", "square_static.tpm"), field_tpm] + ["%.17g" % v for v in wps.ravel()]
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    o = _parse(r.stdout)
    vec = lambda k: np.array([float(x) for x in o[k][0].split()])
    # ---- bsplineTraj
    p = tp.default_params()
    p.max_vel, p.max_acc, p.strict_order = 2.0, 3.0, 1
    bt = tp.BsplineTraj(engine, p)
    path = np.array([[-6.0 + 0.15 * i, -6.0 + 0.15 * i, 1.0] for i in range(81)])
    assert bt.updatePath(path, np.zeros((4, 3)))
    assert np.array_equal(vec("ctrl_in"), bt.getControlPoints().T.ravel())
    ok = bt.makePlan()
    mk = [int(x) for x in o["makeplan"][0].split()]
    assert mk[0] == int(ok)
    assert np.array_equal(vec("ctrl_out"), bt.getControlPoints().T.ravel())
    if ok:
        assert np.array_equal(vec("traj"), vec("ctrl_out"))
        d = o["duration"][0].split()
        assert float(d[0]) == bt.getDuration() and float(d[2]) == bt.getLinearFactor() and int(d[4]) == 1
        x, y, z, yaw = bt.getPose(0.37 * bt.getDuration())
        assert np.allclose([float(v) for v in o["pose"][0].split()], [x, y, z, yaw], rtol=0, atol=1e-15)
        sm = [float(v) for v in o["sample"][0].split()]
        assert sm[0] == 1 and np.allclose(sm[1:4], [x, y, z], rtol=0, atol=0) and abs(sm[4] - yaw) <= 1e-14
        mu = o["multi"][0].split()
        assert mu[0] == "1" and mu[2] == "1"   # six copies solved on two engines == the single solve, bit for bit
    rp = o["replan"][0].split()
    assert rp[0] == "0" and rp[1] == "0" and rp[3] == "1"   # rejected path: the committed trajectory survives
    # ---- bspline value type
    c = vec("traj").reshape(-1, 3)
    b = np.array([float(v) for v in o["bspline"][0].split()])
    ts = p.ctrl_pt_ts
    assert np.array_equal(b[0:3], tp.bspline_eval(c, [1.234], ts, 0)[0])
    assert np.array_equal(b[3:6], tp.bspline_eval(c, [1.234], ts, 1)[0])
    assert np.array_equal(b[6:9], tp.bspline_eval(c, [1.234], ts, 2)[0])
    assert b[9] == (len(c) - 3) * ts
    # ---- polyTrajOctomap: adding-waypoint mode, then corridor mode
    e2 = tp.Engine(0)
    e2.set_map(fm)
    pp = tp.default_poly_params()
    pp.max_iter = 8
    pt = tp.PolyTraj(e2, pp)
    r1 = pt.make_plan_batch([wps])[0]
    l1 = o["poly"][0].split()
    assert int(l1[3]) == int(r1["valid"]) and int(l1[5]) == r1["iters"] and int(l1[7]) == len(r1["path"])
    assert np.array_equal(vec("poly_coef_1"), r1["coef"].ravel())
    r0 = pt.make_plan_corridor_batch([wps], 0.5, 0.8, 8.0)[0]
    l0 = o["poly"][1].split()
    assert int(l0[3]) == int(r0["valid"]) and int(l0[5]) == r0["iters"]
    assert np.array_equal(vec("poly_coef_0"), r0["coef"].ravel())
    for mode, res in ((1, r1), (0, r0)):
        line = [x for x in o["poly_pose"] if x.startswith(str(mode))][0].split()
        coll = [x for x in o["poly_traj_collides"] if x.startswith(str(mode))][0].split()
        if res["valid"]:
            assert coll[1] == "0"   # a valid plan's sampled trajectory passes the shim's own collision check
            from oracle import pwl_np as PW
            ref = PW.poly_get_pose(res["coef"], res["times"], 0.5 * res["times"][-1])
            assert np.allclose([float(v) for v in line[1:]], ref, rtol=0, atol=1e-9)
    e2.close()
    print("shim selftest:", {k: v[0][:60] for k, v in o.items() if k in ("makeplan", "replan")}, o["poly"])
