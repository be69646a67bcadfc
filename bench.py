#!/usr/bin/env python
"""bench.py — ViGO B-spline solves/sec on B200 (BASELINE.json metric), one process per GPU.

  python bench.py --gpus N --steps K --warmup W          (N>1: launched by torch.distributed.run)
  python bench.py --impl reference --gpus N ...          (reference arm: the CPU path on host cores)

A "step" is one pass of the hot path — bsplineTraj::makePlan for a whole batch — over one synthetic
batch.  Default workload: 4,096 random start/goal pairs on the square_static map (BASELINE.json
configs[1]; the pairs of rank 0 are committed in data/workloads/square4096_pairs.npz and shared by both
arms); with N GPUs every rank solves its own 4,096-problem batch (weak scaling, no collective on the
solve path).  `--workload octomap65536` is BASELINE.json configs[2]: 65,536 solves on the maze.bt and
tunnel.bt rasters (half each), sharded over the N ranks in contiguous ranges (strong scaling).
The default N = 1 run also reports configs[2..4] in `extras` (each with its roofline and CPU baseline).

`value`   : solves/s, inputs already resident in HBM, CUDA-event timed on the launching stream.
`e2e`     : same metric through the C ABI with HOST (pinned) buffers: H2D + solve + D2H per step.
`roofline`: the fused cost+L-BFGS kernel (dominant) against the FP64 FMA peak measured live.
`cpu_baseline`: the CPU oracle (reference-order restatement; oracle/_ref links the reference's own
            lbfgs.hpp when it was built) on the host cores over a bounded sample of the same batch.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
MAP_TPM = os.path.join(ROOT, "data", "maps", "square_static.tpm")
SEED = 20261018


def random_pairs(query, B, rng, lo=-9.5, hi=9.5, z=1.0, min_dist=2.0):
    """Seeded start/goal pairs, both ends free in the inflated grid, >= 2 m apart."""
    S, G = [], []
    while len(S) < B:
        n = 2 * (B - len(S)) + 64
        s = np.column_stack([rng.uniform(lo, hi, (n, 2)), np.full(n, z)])
        g = np.column_stack([rng.uniform(lo, hi, (n, 2)), np.full(n, z)])
        ok = (query(s) == 0) & (query(g) == 0) & (np.linalg.norm(g - s, axis=1) >= min_dist)
        S.extend(s[ok])
        G.extend(g[ok])
    return np.array(S[:B]), np.array(G[:B])


def make_workload(tp, pmap, query, B, seed, params, want_pairs=False):
    rng = np.random.default_rng(seed)
    offs, chunks, SS, GG = [0], [], [], []
    while len(offs) - 1 < B:
        need = B - (len(offs) - 1)
        S, G = random_pairs(query, int(need * 1.15) + 16, rng)
        off, ctrl, valid = tp.frontend_batch(pmap, params, S, G)
        for b in range(len(S)):
            if valid[b] and off[b + 1] - off[b] >= 7 and len(offs) - 1 < B:
                chunks.append(ctrl[off[b]:off[b + 1]])
                offs.append(offs[-1] + len(chunks[-1]))
                SS.append(S[b])
                GG.append(G[b])
    if want_pairs:
        return np.array(SS), np.array(GG), np.array(offs, np.int32), np.concatenate(chunks, 0)
    return np.array(offs, np.int32), np.concatenate(chunks, 0)


PAIRS_NPZ = os.path.join(ROOT, "data", "workloads", "square4096_pairs.npz")


def committed_workload(tp, pmap, params, B):
    """The committed (start, goal) pairs of the N = 1 workload -> control points through the product's host front end."""
    z = np.load(PAIRS_NPZ)
    S, G = z["starts"][:B], z["goals"][:B]
    off, ctrl, valid = tp.frontend_batch(pmap, params, S, G)
    if not (np.all(valid[:B] != 0) and np.array_equal(np.diff(off), z["n_ctrl"][:B])):
        raise SystemExit("data/workloads/square4096_pairs.npz does not match the front end: rerun tools/make_workload_pairs.py")
    return off.astype(np.int32), ctrl


def octomap_pairs(inflated, info, n, rng):
    """BASELINE.json configs[2]: start / goal uniform over free inflated cells of the z = 1.0 slab, 2-20 m apart."""
    kz = int(np.floor((1.0 - info["origin"][2]) / info["res"]))
    free = np.argwhere(inflated[:, :, kz] == 0)
    org, res = np.array(info["origin"]), info["res"]
    S, G = np.zeros((0, 3)), np.zeros((0, 3))
    while len(S) < n:
        m = int((n - len(S)) * 1.6) + 64
        a, b = free[rng.integers(len(free), size=m)], free[rng.integers(len(free), size=m)]
        s = org + (np.column_stack([a[:, 0], a[:, 1], np.full(m, kz)]) + 0.5) * res
        g = org + (np.column_stack([b[:, 0], b[:, 1], np.full(m, kz)]) + 0.5) * res
        s[:, 2] = g[:, 2] = 1.0
        d = np.linalg.norm(s - g, axis=1)
        ok = (d >= 2) & (d <= 20)
        S, G = np.vstack([S, s[ok]]), np.vstack([G, g[ok]])
    return S[:n], G[:n]


def oracle_map(O, pmap):
    info = pmap.info()
    om = O.Map(info["res"], info["origin"], info["dims"], info["inflate"])
    occ = pmap.grid("occupied")
    om.add_cells(np.argwhere(occ != 0), occupied=True)
    return om


def cpu_baseline(tp, pmap, offsets, ctrl, sample, threads):
    """The CPU path on `threads` host cores over the first `sample` problems.  -> (solves/s, info)."""
    from oracle import oracle as O
    om = oracle_map(O, pmap)
    p = O.lib().default_params()
    ref_built = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "liborc_ref.so"))
    n = min(sample, len(offsets) - 1)
    off = offsets[:n + 1]
    c = ctrl[:off[n]]
    O.make_plan_batch(om, p, off[:min(n, 64) + 1], c[:off[min(n, 64)]], nthreads=threads)  # warm-up
    t0 = time.perf_counter()
    ok, _, st, ms = O.make_plan_batch(om, p, off, c, nthreads=threads, want_ms=True)
    dt = time.perf_counter() - t0
    info = dict(kind="port", cores=threads,
                sample=f"first {n} problems of the batch, {threads} host threads, one problem per thread at a time; "
                       f"oracle = reference-order CPU restatement (L-BFGS port pinned bit-for-bit to the reference's "
                       f"lbfgs.hpp{' via oracle/_ref' if ref_built else ''}); single-solve p50 "
                       f"{np.median(ms):.3f} ms p95 {np.percentile(ms, 95):.3f} ms; success {ok}/{n}",
                p50_ms=float(np.median(ms)), p95_ms=float(np.percentile(ms, 95)),
                solves_per_s_per_core=float(n / dt / threads))
    return n / dt, info


class ClockSampler:
    def __init__(self, gpu_index):
        self.path = f"/tmp/tp_clocks_{os.getpid()}.csv"
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx),
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nme, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if sm:
            out = dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def oracle_map_np(O, tpm_path, inflate=(4, 4, 2)):
    """Oracle map straight from a .tpm raster with the numpy reader (no product library involved)."""
    from oracle import maps_np
    z = maps_np.read_tpm(tpm_path)
    om = O.Map(z["res"], z["origin"], z["dims"], inflate)
    om.add_cells(np.argwhere(z["occupied"] != 0), occupied=True)
    free = np.argwhere((z["known"] != 0) & (z["occupied"] == 0))
    if len(free):
        om.add_cells(free, occupied=False)
    return om, z


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path (oracle: reference-order restatement whose L-BFGS
    is pinned to the reference's own lbfgs.hpp) on all host cores, same config: every step solves the FIRST `ref_sample`
    problems of the committed 4,096 batch.  Inputs are built with numpy only (oracle/maps_np.py, oracle/frontend_np.py):
    this arm never loads the product library."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O, frontend_np
    om, _ = oracle_map_np(O, MAP_TPM)
    threads = os.cpu_count() or 1
    sample = int(args.ref_sample)
    z = np.load(PAIRS_NPZ)
    chunks = []
    for b in range(sample):
        c = frontend_np.start_goal_to_ctrl(z["starts"][b], z["goals"][b], om)
        if c is None or len(c) != z["n_ctrl"][b]:
            raise SystemExit(f"numpy front end disagrees with the committed workload at pair {b}")
        chunks.append(c)
    offsets = np.concatenate([[0], np.cumsum([len(c) for c in chunks])]).astype(np.int32)
    ctrl = np.concatenate(chunks, 0)
    po = O.lib().default_params()
    for _ in range(max(args.warmup, 1)):
        O.make_plan_batch(om, po, offsets[:65], ctrl[:offsets[64]], nthreads=threads)
    t0 = time.perf_counter()
    mss = []
    for _ in range(args.steps):
        ok, _, st, ms = O.make_plan_batch(om, po, offsets, ctrl, nthreads=threads, want_ms=True)
        mss.append(ms)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    ms = np.concatenate(mss)
    ref_built = bool(O.lib(ref=False).L.orc_is_ref_build()) if hasattr(O.lib().L, "orc_is_ref_build") else False
    line = dict(metric="ViGO B-spline solves/sec", value=value, unit="solves/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * dt / args.steps, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f64", data="synthetic", impl="reference",
                config=dict(workload="batch of 4,096 ViGO solves, random start/goal pairs on square_static_map.pcd "
                                     "(0.1 m voxels, 400x400x30 grid), per GPU", batch_per_gpu=4096, map="square_static",
                            seed=SEED, same_config=True,
                            sample=f"each step = the first {sample} problems of that batch (data/workloads/square4096_pairs.npz), "
                                   "control points from the numpy front end (<= 1e-9 m from the product's)"),
                cpu_baseline=dict(value=value, unit="solves/s", cores=threads, kind="port",
                                  sample=f"first {sample} problems of the 4,096 batch x {args.steps} steps, {threads} host threads, one "
                                         f"problem per thread at a time; p50 {np.median(ms):.3f} ms/solve; success {int(ok)}/{sample}; "
                                         "oracle = reference-order CPU restatement, L-BFGS port pinned bit-for-bit to the reference's lbfgs.hpp"),
                e2e=dict(value=value, unit="solves/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                single_solve_p50_ms=float(np.median(ms)), product_library_loaded=("trajectory_planner_b200" in sys.modules))
    print(json.dumps(line), flush=True)


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


class OctomapShard:
    """One rank's share of BASELINE.json configs[2] on one map: seeded (start, goal) pairs over the whole batch (every
    rank draws the same list), the rank's contiguous range through the DEVICE front end, control points resident."""

    def __init__(self, tp, torch, name, total, rank, world, local, seed):
        from trajectory_planner_b200 import sharding
        self.tp, self.torch, self.name = tp, torch, name
        self.pmap = tp.OccMap.from_tpm(os.path.join(ROOT, "data", "maps", name + ".tpm"))
        self.eng = tp.Engine(local)
        self.eng.set_map(self.pmap)
        self.p = tp.default_params()
        info = self.pmap.info()
        rng = np.random.default_rng(seed)
        # over-draw so that enough pairs survive the front end's validity checks; the kept list is the same on every rank
        S, G = octomap_pairs(self.pmap.grid("inflated"), info, int(total * 1.25) + 256, rng)
        off, ctrl, valid = self.eng.frontend_batch(self.p, S, G)
        keep = np.flatnonzero((valid != 0) & (np.diff(off) >= 7))[:total]
        if len(keep) < total:
            raise SystemExit(f"{name}: only {len(keep)} valid pairs of {total}")
        b0, b1 = sharding.shard_bounds(total, world)[rank]
        mine = keep[b0:b1]
        chunks = [ctrl[off[b]:off[b + 1]] for b in mine]
        self.offsets = np.concatenate([[0], np.cumsum([len(c) for c in chunks])]).astype(np.int32)
        self.ctrl = np.concatenate(chunks, 0)
        self.B = len(mine)
        dev = torch.device("cuda", local)
        self.d_off = torch.from_numpy(self.offsets).to(dev)
        self.d_in = torch.from_numpy(self.ctrl).to(dev)
        self.d_out = torch.empty_like(self.d_in)
        self.d_res = torch.empty(self.B * tp.RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)

    def step(self, stream):
        self.eng.make_plan_batch_device(self.p, self.B, self.d_off.data_ptr(), self.d_in.data_ptr(), self.d_out.data_ptr(),
                                        self.d_res.data_ptr(), stream)

    def results(self):
        return np.frombuffer(self.d_res.cpu().numpy().tobytes(), dtype=self.tp.RESULT_DTYPE)


def octomap_cpu_baseline(sample_per_map, threads):
    """The CPU path on the first problems of both rasters (numpy front end, oracle maps from the numpy .tpm reader)."""
    from oracle import oracle as O, frontend_np
    n, dt, okc = 0, 0.0, 0
    for name in ("maze", "tunnel"):
        om, z = oracle_map_np(O, os.path.join(ROOT, "data", "maps", name + ".tpm"))
        _, _, infl = om.grids()
        info = dict(origin=z["origin"], res=z["res"], dims=z["dims"])
        S, G = octomap_pairs(infl, info, 4 * sample_per_map, np.random.default_rng(SEED + 11))
        chunks = []
        for b in range(len(S)):
            c = frontend_np.start_goal_to_ctrl(S[b], G[b], om)
            if c is not None and len(c) >= 7:
                chunks.append(c)
            if len(chunks) == sample_per_map:
                break
        offsets = np.concatenate([[0], np.cumsum([len(c) for c in chunks])]).astype(np.int32)
        po = O.lib().default_params()
        t0 = time.perf_counter()
        ok, _, st = O.make_plan_batch(om, po, offsets, np.concatenate(chunks, 0), nthreads=threads)
        dt += time.perf_counter() - t0
        n += len(chunks)
        okc += int(ok)
    return dict(value=n / dt, unit="solves/s", cores=threads, kind="port",
                sample=f"{n} problems ({sample_per_map} per map, own seeded pairs of the same distribution), {threads} host threads; success {okc}/{n}",
                solves_per_s_per_core=n / dt / threads)


def run_octomap(args, tp, torch, dist, rank, world, local, quiet=False, total=None, K=None, W=None):
    """BASELINE.json configs[2]: `total` ViGO solves (half on maze.bt, half on tunnel.bt rasters), sharded over the ranks in
    contiguous ranges, strong scaling.  A step = every rank solves its maze range and its tunnel range (two engines, two
    streams, concurrently)."""
    total = int(total or args.octomap_total)
    K = K or max(args.steps, 1)
    W = W if W is not None else max(args.warmup, 3)
    dev = torch.device("cuda", local)
    shards = [OctomapShard(tp, torch, name, total // 2, rank, world, local, SEED + 7 + i) for i, name in enumerate(("maze", "tunnel"))]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the two rasters' batches are independent: each runs on its own stream (its own engine), so one raster's tail of long
    # A* searches overlaps the other's bulk; the step is timed on the launching stream around both
    side = torch.cuda.Stream(device=dev)
    ev_go, ev_done = torch.cuda.Event(), torch.cuda.Event()

    def step_all():
        ev_go.record(tstream)
        side.wait_event(ev_go)
        shards[0].step(stream)
        shards[1].step(side.cuda_stream)
        ev_done.record(side)
        tstream.wait_event(ev_done)

    for _ in range(W):
        step_all()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    for sh in shards:
        sh.eng.profile_enable(True)
        sh.eng.profile_get()
    l0 = sum(sh.eng.launch_count for sh in shards)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    for k in range(K):
        flush.zero_()
        ev[k][0].record()
        step_all()
        ev[k][1].record()
    barrier()
    total_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    profs = [sh.eng.profile_get() for sh in shards]
    launches = sum(sh.eng.launch_count for sh in shards) - l0
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = total * K / (total_ms * 1e-3)
    # e2e: host buffers through the C ABI
    t0 = time.perf_counter()
    for sh in shards:
        sh.eng.make_plan_batch(sh.p, sh.offsets, sh.ctrl)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    if rank != 0:
        if world > 1 and not quiet:
            dist.barrier()
            dist.destroy_process_group()
        return None
    res = [sh.results() for sh in shards]
    fp64_peak = shards[0].eng.microbench_fp64()
    flops = sum(pr["lbfgs_flops"] for pr in profs) / K
    step_ms = total_ms / K
    pts = sum(int(sh.offsets[-1]) for sh in shards)
    nb = sum(sh.B for sh in shards)
    per_map = {sh.name: dict(batch_this_rank=sh.B, success_rate=float(np.mean(r["status"] == 1)),
                             control_points_mean=float(np.diff(sh.offsets).mean()),
                             lbfgs_iters_per_solve=float(r["lbfgs_iters"].mean()),
                             astar_expansions_per_solve=float(r["astar_expansions"].mean()),
                             astar_expansions_max=int(r["astar_expansions"].max())) for sh, r in zip(shards, res)}
    ach = flops / (step_ms * 1e-3) / 1e12
    line = dict(metric="ViGO B-spline solves/sec", value=value, unit="solves/s", n_gpus=world, steps=K, warmup=W,
                ms_per_step=step_ms, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=f"batch of {total:,} ViGO solves on maze.bt and tunnel.bt rasterised to voxel grids (half each), "
                                     f"sharded across {world} B200 in contiguous ranges", total=total, seed=SEED + 7,
                            reduction_order="fast", l2="flushed between timed iterations (256 MB write)", per_map=per_map),
                clocks=clocks, gpu_launches=int(launches),
                e2e=dict(value=total / e2e_s, unit="solves/s", h2d_bytes_per_step=int(pts * 24 + (nb + 2) * 4),
                         d2h_bytes_per_step=int(pts * 24 + nb * tp.RESULT_DTYPE.itemsize)),
                roofline=dict(bound="fp64", achieved=ach, peak=fp64_peak, unit="TFLOP/s", frac=ach / fp64_peak if fp64_peak else None,
                              traffic=None, kernel="k_solve<team> (fused cost + L-BFGS share; these rasters are A*-bound: see "
                                                   "astar_expansions_per_solve)",
                              peak_source="measured live: dependent-free FP64 FMA micro-benchmark", flops_per_step=flops))
    for sh in shards:
        sh.eng.close()
    if not quiet:
        threads = os.cpu_count() or 1
        line["cpu_baseline"] = octomap_cpu_baseline(48, threads)
        print(json.dumps(line), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
    return line


def extra_minsnap(tp, eng_dev, fp64_peak, B=16384):
    """BASELINE.json configs[3]: 16,384 waypoint paths (8-20 waypoints, random walk with 1-4 m steps at z = 1 inside the
    known-free region of field.bt), cfg/planner_interactive.yaml with mode = true: (a) the batched exact KKT solve,
    (b) the whole solve -> sample -> collision check -> insert-waypoint loop."""
    m = tp.OccMap.from_tpm(os.path.join(ROOT, "data", "maps", "field.tpm"))
    info = m.info()
    e = tp.Engine(eng_dev)
    e.set_map(m)
    p = tp.default_poly_params()
    p.max_iter = 20
    pt = tp.PolyTraj(e, p)
    rng = np.random.default_rng(SEED)
    occ, known = m.grid("occupied"), m.grid("known")
    kz = int(np.floor((1.0 - info["origin"][2]) / info["res"]))
    free = np.argwhere((known[:, :, kz] != 0) & (occ[:, :, kz] == 0))
    org, res = np.array(info["origin"]), info["res"]
    cur = org + (np.column_stack([free[rng.integers(len(free), size=B)], np.full(B, kz)]) + 0.5) * res
    cur[:, 2] = 1.0
    nw = rng.integers(8, 21, size=B)
    P = np.zeros((B, 20, 3))
    P[:, 0] = cur
    cnt = np.ones(B, int)
    for _try in range(19 * 12):   # vectorised random walks; a step is accepted when the box check at the new waypoint passes
        pending = np.flatnonzero(cnt < nw)
        if len(pending) == 0:
            break
        ang, st = rng.uniform(0, 2 * np.pi, len(pending)), rng.uniform(1, 4, len(pending))
        cand = P[pending, cnt[pending] - 1] + np.column_stack([st * np.cos(ang), st * np.sin(ang), np.zeros(len(pending))])
        ok = pt.box_collision(cand) == 0
        P[pending[ok], cnt[pending[ok]]] = cand[ok]
        cnt[pending[ok]] += 1
    paths = [P[b, :cnt[b]].copy() for b in range(B) if cnt[b] >= 2]
    B = len(paths)
    nwp = np.array([len(q) for q in paths])
    pt.solve_batch(paths[:256])
    # one untimed full-size pass of both entries (as the warm-up steps of the main metric): the first full-size call after
    # the octomap engines were closed was measured 0.5-1 s slower than the same call in a fresh process
    pt.solve_batch(paths)
    pt.make_plan_batch(paths)
    e.profile_enable(True)
    e.profile_get()
    t0 = time.perf_counter()
    sols, status = pt.solve_batch(paths)
    t_solve = time.perf_counter() - t0
    prof = e.profile_get()
    t0 = time.perf_counter()
    res_ = pt.make_plan_batch(paths)
    t_loop = time.perf_counter() - t0
    loop_kms = e.profile_get()["ms"]["minsnap_solve"]
    valid = np.array([r["valid"] for r in res_])
    iters = np.array([r["iters"] for r in res_])
    # the reference's DEFAULT mode (cfg/planner_interactive.yaml:31 `mode: false`): corridor constraints, interior-point QP
    nc = min(B, 2048)
    pc = tp.default_poly_params()
    pc.max_iter = 8
    ptc = tp.PolyTraj(e, pc)
    ptc.make_plan_corridor_batch(paths[:64])
    e.profile_get()
    t0 = time.perf_counter()
    res_c = ptc.make_plan_corridor_batch(paths[:nc], 0.5, 0.8, 8.0)
    t_corr = time.perf_counter() - t0
    corr_kms = e.profile_get()["ms"]["minsnap_solve"]
    e.profile_enable(False)
    K = nwp - 1
    flops = float(np.sum(2.0 * (14 * K) * 22 ** 2 + 3 * 4.0 * (14 * K) * 22))   # banded KKT factor + three right-hand sides
    kms = prof["ms"]["minsnap_solve"]
    ach = flops / (kms * 1e-3) / 1e12 if kms > 0 else None
    # CPU baseline: the numpy restatement of the loop (exact KKT solve per iteration), bounded sample, one core
    from oracle import polytraj_np as PN
    g3 = PN.Grid3(info["res"], info["origin"], occ, known)
    ns, t_cpu = 0, 0.0
    for q in paths[:48]:
        t0 = time.perf_counter()
        try:
            PN.make_plan_adding_waypoint(q, g3, max_iter=int(p.max_iter))
        except np.linalg.LinAlgError:   # numpy's lstsq gives up on a few degenerate KKT systems: not counted
            continue
        t_cpu += time.perf_counter() - t0
        ns += 1
        if ns == 24:
            break
    e.close()
    return dict(workload="min-snap polyTrajOctomap batch: 16,384 waypoint paths (8-20 waypoints each) on field.bt with "
                         "collision-check-and-reinsert loop", batch=B,
                waypoints=dict(min=int(nwp.min()), mean=float(nwp.mean()), max=int(nwp.max())),
                solve_only=dict(value=B / (kms * 1e-3) if kms > 0 else None, unit="solves/s (kernel)", kernel_ms=kms,
                                wall_ms=1e3 * t_solve, singular=int((status != 0).sum())),
                loop=dict(value=B / t_loop, unit="paths/s", wall_ms=1e3 * t_loop, kernel_ms=loop_kms, valid_rate=float(valid.mean()),
                          iters_mean=float(iters.mean()), max_iter=int(p.max_iter)),
                corridor_loop=dict(value=nc / t_corr, unit="paths/s", paths=nc, wall_ms=1e3 * t_corr, kernel_ms=corr_kms,
                                   valid_rate=float(np.mean([r["valid"] for r in res_c])),
                                   infeasible_rate=float(np.mean([bool(np.any(r["status"] != 0)) for r in res_c])),
                                   iters_mean=float(np.mean([r["iters"] for r in res_c])), max_iter=int(pc.max_iter),
                                   note="polyTrajOctomap::makePlanCorridorConstraint (initial_radius 0.5, shrinking_factor 0.8, corridor_res 8): "
                                        "Mehrotra interior-point QP per axis on the band KKT matrix, whole loop on the device"),
                roofline=dict(bound="fp64", achieved=ach, peak=fp64_peak, unit="TFLOP/s", frac=(ach / fp64_peak) if ach and fp64_peak else None,
                              traffic=None, kernel="k_minsnap_solve",
                              flops_model="banded KKT: 2 (14K) 22^2 + 12 (14K) 22 per path (SURVEY.md 8d); the kernel runs a band LU with partial pivoting "
                                          "(half-bandwidth 13, U bandwidth 26), one warp per path"),
                cpu_baseline=dict(value=ns / t_cpu, unit="paths/s", cores=1, kind="port",
                                  sample=f"{ns} of the first paths, numpy restatement of the loop (oracle/polytraj_np.py), one core"))


def extra_sweep(tp, torch, dev, eng_dev, flush, hbm):
    """BASELINE.json configs[4]: 1 M trajectory samples against box.bt (+ 16 M, + a synthetic 1e8-voxel grid): achieved
    GB/s against the HBM roofline (stream side: 24 B xyz + 1 B flag per query come from / go to HBM) and against the
    measured L2 random-sector gather rate (map side: one 32 B sector per query, L2 resident)."""
    eng = tp.Engine(eng_dev)
    tstream = torch.cuda.current_stream(dev)
    stream = tstream.cuda_stream
    cases = []

    def timed(nq, q, hit, reps=5):
        for _ in range(3):
            eng.query_points_device(nq, q.data_ptr(), hit.data_ptr(), stream)
        torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.query_points_device(nq, q.data_ptr(), hit.data_ptr(), stream)
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        return float(np.median(ms))

    l2_gather = eng.microbench_gather(12_500_000)

    def case(name, pmap, nq, coherent=False):
        eng.set_map(pmap)
        info = pmap.info()
        lo = np.array(info["origin"])
        hi = lo + np.array(info["dims"]) * info["res"]
        g = torch.Generator(device=dev)
        g.manual_seed(1)
        if coherent:   # trajectories of 512 consecutive samples, 2.5 cm apart along random directions
            nt = nq // 512
            p0 = torch.rand((nt, 1, 3), generator=g, device=dev, dtype=torch.float64) * torch.tensor(hi - lo, device=dev) + torch.tensor(lo, device=dev)
            d = torch.randn((nt, 1, 3), generator=g, device=dev, dtype=torch.float64)
            d = d / d.norm(dim=2, keepdim=True)
            q = (p0 + d * (0.025 * torch.arange(512, device=dev, dtype=torch.float64))[None, :, None]).reshape(-1, 3).contiguous()
        else:
            q = torch.rand((nq, 3), generator=g, device=dev, dtype=torch.float64) * torch.tensor(hi - lo, device=dev) + torch.tensor(lo, device=dev)
        hit = torch.empty(len(q), dtype=torch.uint8, device=dev)
        ms = timed(len(q), q, hit)
        n = len(q)
        cases.append(dict(map=name, cells=int(np.prod(info["dims"])), packed_MB=info["packed_bytes"] / 1e6, queries=n,
                          pattern="trajectory-coherent" if coherent else "uniform", ms=ms, gqueries_per_s=n / ms / 1e6,
                          algorithmic_gbs_57B=n * 57.0 / (ms * 1e-3) / 1e9,
                          hbm_stream_gbs_25B=n * 25.0 / (ms * 1e-3) / 1e9, frac_hbm=n * 25.0 / (ms * 1e-3) / 1e9 / hbm,
                          l2_sector_gbs_32B=n * 32.0 / (ms * 1e-3) / 1e9, frac_l2_gather=n * 32.0 / (ms * 1e-3) / 1e9 / l2_gather,
                          hit_rate=float(hit.float().mean().item())))
        return q

    box = tp.OccMap.from_tpm(os.path.join(ROOT, "data", "maps", "box.tpm"))
    q1m = case("box.bt", box, 1 << 20)
    case("box.bt", box, 1 << 20, coherent=True)
    case("box.bt", box, 16 << 20)
    # CPU baseline: the oracle's point query on one host core over the same 1 M points
    from oracle import oracle as O
    om, _ = oracle_map_np(O, os.path.join(ROOT, "data", "maps", "box.tpm"))
    qh = q1m.cpu().numpy()
    t0 = time.perf_counter()
    om.query(qh)
    t_cpu = time.perf_counter() - t0
    rng = np.random.default_rng(0)
    dims = (1000, 1000, 100)
    m = tp.OccMap(0.1, (0.0, 0.0, 0.0), dims, (0, 0, 0))
    nbox = int(0.1 * np.prod(dims) / (20 * 20 * 10))
    ii, jj, kk = np.meshgrid(np.arange(20), np.arange(20), np.arange(10), indexing="ij")
    cells = []
    for _ in range(nbox):
        c = rng.integers(0, np.array(dims) - np.array([20, 20, 10]))
        cells.append(np.stack([ii.ravel() + c[0], jj.ravel() + c[1], kk.ravel() + c[2]], 1))
    m.add_cells(np.concatenate(cells).astype(np.int32))
    case("synthetic 1e8 voxels, 10 % boxes (12.5 MB bit-packed: L2 resident)", m, 16 << 20)
    eng.close()
    head = cases[0]
    return dict(workload="collision-query sweep: 1M trajectory samples against box.bt up to a synthetic 1e8-voxel grid",
                value=head["gqueries_per_s"], unit="Gqueries/s (1 M uniform samples, box.bt)", cases=cases,
                l2_gather_peak_gbs=l2_gather, hbm_peak_gbs=hbm,
                roofline=dict(bound="hbm", achieved=head["hbm_stream_gbs_25B"], peak=hbm, unit="GB/s", frac=head["frac_hbm"],
                              traffic=ncu_traffic("k_query_points"), kernel="k_query_points",
                              note="per query 24 B xyz in + 1 B flag out cross HBM; the 32 B map sector is an L2 hit for every map "
                                   "of the config (<= 12.5 MB bit-packed): frac = HBM-side bytes / measured copy peak; "
                                   "frac_l2_gather in `cases` = sector bytes / measured L2 random-sector gather rate; the 1 M-sample "
                                   "cases are launch-bound (one ~16 us kernel)"),
                cpu_baseline=dict(value=len(qh) / t_cpu / 1e9, unit="Gqueries/s", cores=1, kind="port",
                                  sample="the same 1 M uniform points through the oracle's occMap::isInflatedOccupied, one core"),
                not_measured="an int8 (1 B / voxel, 100 MB) HBM-bound map variant: the engine stores maps bit-packed only")


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of the kernel at the bench workload
    (profiles/ncu_traffic.json, written from the capture committed beside it); None when there is no capture."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[kernel]["bytes_per_step"]
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--mode", default=os.environ.get("TP_BENCH_MODE", "fast"), choices=["fast", "strict"])
    ap.add_argument("--cpu-sample", type=int, default=2048)
    ap.add_argument("--ref-sample", type=int, default=1024)
    ap.add_argument("--no-extras", action="store_true", help="skip cpu_baseline / sweeps / second mode")
    ap.add_argument("--workload", default="square4096", choices=["square4096", "octomap65536"],
                    help="square4096: BASELINE configs[1], weak scaling (default); octomap65536: configs[2], 65,536 solves on "
                         "maze.bt + tunnel.bt sharded over the ranks (strong scaling)")
    ap.add_argument("--octomap-total", type=int, default=65536)
    ap.add_argument("--seed-offset", type=int, default=0, help="workload seed = SEED + rank + offset (rank r of an N-GPU run uses offset r)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import trajectory_planner_b200 as tp

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    B = args.batch

    pmap = tp.OccMap.from_tpm(MAP_TPM)
    eng = tp.Engine(local)
    eng.set_map(pmap)
    p = tp.default_params()
    p.strict_order = 1 if args.mode == "strict" else 0
    if args.workload == "octomap65536":
        return run_octomap(args, tp, torch, dist, rank, world, local)
    if rank + args.seed_offset == 0 and B <= 4096 and os.path.exists(PAIRS_NPZ):
        offsets, ctrl = committed_workload(tp, pmap, p, B)   # the committed pairs: the reference arm solves a slice of these
    else:
        offsets, ctrl = make_workload(tp, pmap, eng.query_points, B, SEED + rank + args.seed_offset, p)
    total_pts = int(offsets[-1])
    Ns = np.diff(offsets)

    dev = torch.device("cuda", local)
    d_off = torch.from_numpy(offsets).to(dev)
    d_in = torch.from_numpy(ctrl).to(dev)
    d_out = torch.empty_like(d_in)
    d_res = torch.empty(B * tp.RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    tstream = torch.cuda.Stream(device=dev)   # the launching stream: events below are recorded on it
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    def step_device():
        eng.make_plan_batch_device(p, B, d_off.data_ptr(), d_in.data_ptr(), d_out.data_ptr(), d_res.data_ptr(), stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step_device()
    barrier()
    # ---- timed region: exactly K steps, L2 flushed between iterations, CUDA events on the stream
    sampler = ClockSampler(local)
    sampler.start()
    eng.profile_enable(True)
    eng.profile_get()
    launches0 = eng.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    for k in range(K):
        flush.zero_()
        ev[k][0].record()
        step_device()
        ev[k][1].record()
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    prof = eng.profile_get()
    eng.profile_enable(False)
    gpu_launches = eng.launch_count - launches0
    clocks = sampler.stop()
    total_ms = float(sum(step_ms))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * B * K / (total_ms * 1e-3)
    res = np.frombuffer(d_res.cpu().numpy().tobytes(), dtype=tp.RESULT_DTYPE)

    # ---- e2e: host (pinned) buffers through the public batched entry point, copies inside the timing
    h_in = torch.from_numpy(ctrl).pin_memory()
    h_off = torch.from_numpy(offsets).pin_memory()
    for _ in range(2):
        eng.make_plan_batch(p, h_off.numpy(), h_in.numpy())
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        out_h, res_h = eng.make_plan_batch(p, h_off.numpy(), h_in.numpy())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * B * K / e2e_s
    h2d = int(total_pts * 24 + (B + 1) * 4)
    d2h = int(total_pts * 24 + B * tp.RESULT_DTYPE.itemsize)

    # ---- N > 1: BASELINE.json configs[2] as well — the 65,536-solve maze.bt + tunnel.bt batch STRONG-scaled over the ranks
    # (contiguous shards from sharding.shard_bounds, no collective on the solve path), every rank takes part
    oct_line = None
    if world > 1 and not args.no_extras:
        try:
            oct_line = run_octomap(args, tp, torch, dist, rank, world, local, quiet=True, total=args.octomap_total, K=2, W=1)
        except Exception as ex:   # a secondary workload must not take the headline line down
            oct_line = dict(error=repr(ex))
        torch.cuda.set_stream(tstream)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: k_solve (one block per trajectory, the whole makePlan).  Its arithmetic is
    # the fused cost + L-BFGS iterate: FP64 pipe, peak measured live.  The size-class launches of one step overlap, so
    # the kernel's duration per step is the device time of the step itself (events above), not the sum of its launches.
    fp64_peak = eng.microbench_fp64()
    lb_ms = prof["ms"]["solve"]
    step_ms_mean = float(np.mean(step_ms))
    ach = prof["lbfgs_flops"] / K / (step_ms_mean * 1e-3) / 1e12
    kern_ms = {k: v / K for k, v in prof["ms"].items()}
    bytes_per_step = float(2 * 24 * total_pts + 64 * B)   # control points in + out, result records
    roofline = dict(bound="fp64", achieved=ach, peak=fp64_peak, unit="TFLOP/s", frac=ach / fp64_peak if fp64_peak else None,
                    traffic=ncu_traffic("k_solve"),
                    kernel="k_solve<team> (bsplineTraj::makePlan per thread block: segments, A*, guide points, fused "
                           "cost+gradient+L-BFGS in the lean team form, collision check, re-parameterisation)",
                    peak_source="measured live: dependent-free FP64 FMA micro-benchmark (tp_microbench_fp64); "
                                "MEASURED_PEAKS.json carries no FP64 figure",
                    launches_per_step=prof["launches"]["solve"] / K,
                    launch_ms_sum_per_step=lb_ms / K, step_ms=step_ms_mean,
                    flops_per_step=prof["lbfgs_flops"] / K,
                    flops_model="E(81N + 21G + 4n) + sum_k (8 b_k + 15) n per optimize() (SURVEY.md 8d), counted by the kernel",
                    lbfgs_iters_per_step=prof["lbfgs_iters"] / K, cost_evals_per_step=prof["lbfgs_evals"] / K,
                    hbm_algorithmic_gbs=bytes_per_step / (step_ms_mean * 1e-3) / 1e9,
                    note="latency-bound small-vector FP64 work (n <= 300 unknowns per problem): the fraction of the FMA "
                         "peak is low by construction; see profiles/ for issue-slot and stall breakdown",
                    kernel_ms_per_step=kern_ms)
    line = dict(metric="ViGO B-spline solves/sec", value=value, unit="solves/s", n_gpus=world, steps=K, warmup=W,
                ms_per_step=total_ms / K, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                data="synthetic",
                config=dict(workload=f"batch of {B:,} ViGO solves, random start/goal pairs on square_static_map.pcd "
                                     "(0.1 m voxels, 400x400x30 grid), per GPU", batch_per_gpu=B, map="square_static",
                            seed=SEED, reduction_order=args.mode, l2="flushed between timed iterations (256 MB write)",
                            control_points=dict(min=int(Ns.min()), mean=float(Ns.mean()), max=int(Ns.max())),
                            success_rate=float(np.mean(res["status"] == 1)),
                            lbfgs_iters_per_solve=float(res["lbfgs_iters"].mean()),
                            astar_expansions_per_solve=float(res["astar_expansions"].mean())),
                clocks=clocks, gpu_launches=int(gpu_launches),
                e2e=dict(value=e2e_value, unit="solves/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
                roofline=roofline)

    if world == 1 and not args.no_extras:
        # single-solve latency (batch of one, host buffers) — BASELINE.json "single-solve p50 ms": the FIRST 256 problems of
        # the batch, each solved alone (the CPU arm's p50 is over single solves of the same problems)
        for i in range(8):
            eng.make_plan_batch(p, np.array([0, Ns[i]], np.int32), ctrl[offsets[i]:offsets[i + 1]])
        lat = []
        for i in range(min(256, B)):
            o1 = np.array([0, Ns[i]], np.int32)
            c1 = ctrl[offsets[i]:offsets[i + 1]]
            t0 = time.perf_counter()
            eng.make_plan_batch(p, o1, c1)
            lat.append(1e3 * (time.perf_counter() - t0))
        line["single_solve_p50_ms"] = float(np.median(lat))
        line["single_solve"] = dict(p50_ms=float(np.median(lat)), p95_ms=float(np.percentile(lat, 95)), mean_ms=float(np.mean(lat)),
                                    problems=len(lat), note="one trajectory per call through the C ABI with host buffers; the path is "
                                    "latency-bound (a serial chain of ~600 L-BFGS iterations): one block of a 1.9 GHz GPU does not beat a "
                                    "CPU core on ONE solve, the engine's case is the batch")
        # the other reduction order, for the record
        p2 = tp.default_params()
        p2.strict_order = 0 if args.mode == "strict" else 1
        for _ in range(2):
            eng.make_plan_batch_device(p2, B, d_off.data_ptr(), d_in.data_ptr(), d_out.data_ptr(), d_res.data_ptr(), stream)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            eng.make_plan_batch_device(p2, B, d_off.data_ptr(), d_in.data_ptr(), d_out.data_ptr(), d_res.data_ptr(), stream)
        b.record()
        torch.cuda.synchronize()
        line["other_mode"] = dict(reduction_order="fast" if args.mode == "strict" else "strict",
                                  value=3 * B / (a.elapsed_time(b) * 1e-3), unit="solves/s")
        # map queries, 16 M random points: stream side vs HBM, sector side vs L2 gather
        nq = 16 << 20
        q = torch.empty((nq, 3), dtype=torch.float64, device=dev).uniform_(-12.0, 12.0)
        q[:, 2].uniform_(0.0, 2.8)
        hit = torch.empty(nq, dtype=torch.uint8, device=dev)
        for _ in range(3):
            eng.query_points_device(nq, q.data_ptr(), hit.data_ptr(), stream)
        torch.cuda.synchronize()
        flush.zero_()
        a.record()
        eng.query_points_device(nq, q.data_ptr(), hit.data_ptr(), stream)
        b.record()
        torch.cuda.synchronize()
        qms = a.elapsed_time(b)
        peaks = load_peaks()
        hbm = peaks.get("hbm_gbs", 6650.0)
        l2g = eng.microbench_gather(12_500_000)
        line["roofline_query"] = dict(bound="hbm", achieved=nq * 25.0 / (qms * 1e-3) / 1e9, peak=hbm, unit="GB/s",
                                      frac=nq * 25.0 / (qms * 1e-3) / 1e9 / hbm, traffic=ncu_traffic("k_query_points"),
                                      kernel="k_query_points", points=nq, ms=qms,
                                      bytes_per_query="HBM side: 24 B xyz in + 1 B flag out = 25 B; the 32 B map sector is an L2 hit "
                                                      "(640 KB map): reported separately",
                                      l2_sector_gbs=nq * 32.0 / (qms * 1e-3) / 1e9, l2_gather_peak_gbs=l2g,
                                      frac_l2_gather=nq * 32.0 / (qms * 1e-3) / 1e9 / l2g,
                                      algorithmic_gbs_57B=nq * 57.0 / (qms * 1e-3) / 1e9,
                                      peak_source="MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650")
        line["l2_gather_gbs"] = l2g
        del q, hit
        # front end: the same (start, goal) pairs -> control points on the device vs on the host cores
        rng = np.random.default_rng(SEED)
        S, G = random_pairs(eng.query_points, B, rng)
        eng.frontend_batch(p, S, G)
        t0 = time.perf_counter()
        off_d, ctrl_d, valid_d = eng.frontend_batch(p, S, G)
        fe_dev = time.perf_counter() - t0
        t0 = time.perf_counter()
        off_h, ctrl_h, valid_h = tp.frontend_batch(pmap, p, S, G)
        fe_host = time.perf_counter() - t0
        line["front_end"] = dict(pairs=B, device_ms=1e3 * fe_dev, host_1core_ms=1e3 * fe_host,
                                 identical_offsets=bool(np.array_equal(off_d, off_h)),
                                 max_abs_diff=float(np.max(np.abs(ctrl_d - ctrl_h))) if ctrl_d.shape == ctrl_h.shape else None)
        # CPU baseline on the host cores (bounded sample)
        threads = os.cpu_count() or 1
        v, info = cpu_baseline(tp, pmap, offsets, ctrl, args.cpu_sample, threads)
        line["cpu_baseline"] = dict(value=v, unit="solves/s", **info)
        # BASELINE.json configs[2..4] (secondary workloads), each with its own roofline and CPU baseline
        eng.close()
        extras = {}
        try:
            o = run_octomap(args, tp, torch, dist, 0, 1, local, quiet=True, total=65536, K=2, W=1)
            o["cpu_baseline"] = octomap_cpu_baseline(48, threads)
            extras["octomap65536"] = {k: o[k] for k in ("value", "unit", "ms_per_step", "scaling", "config", "e2e", "roofline", "cpu_baseline", "gpu_launches")}
        except Exception as ex:   # a secondary workload must not take the headline line down
            extras["octomap65536"] = dict(error=repr(ex))
        torch.cuda.set_stream(tstream)
        try:
            extras["minsnap16384"] = extra_minsnap(tp, local, fp64_peak)
        except Exception as ex:
            extras["minsnap16384"] = dict(error=repr(ex))
        try:
            extras["collision_sweep"] = extra_sweep(tp, torch, dev, local, flush, load_peaks().get("hbm_gbs", 6650.0))
        except Exception as ex:
            extras["collision_sweep"] = dict(error=repr(ex))
        line["extras"] = extras
    if oct_line is not None:
        keys = ("value", "unit", "n_gpus", "ms_per_step", "scaling", "config", "e2e", "roofline", "gpu_launches", "error")
        line["extras"] = {"octomap65536": {k: oct_line[k] for k in keys if k in oct_line}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
