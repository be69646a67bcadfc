#!/usr/bin/env python
"""bench.py — ViGO B-spline solves/sec on B200 (BASELINE.json metric), one process per GPU.

  python bench.py --gpus N --steps K --warmup W          (N>1: launched by torch.distributed.run)
  python bench.py --impl reference --gpus N ...          (reference arm: the CPU path on host cores)

A "step" is one pass of the hot path — bsplineTraj::makePlan for a whole batch — over one synthetic
batch: 4,096 random start/goal pairs on the square_static map (BASELINE.json configs[1]); with N GPUs
every rank solves its own 4,096-problem batch (weak scaling, no collective on the solve path).

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


def make_workload(tp, pmap, query, B, seed, params):
    rng = np.random.default_rng(seed)
    offs, chunks = [0], []
    while len(offs) - 1 < B:
        need = B - (len(offs) - 1)
        S, G = random_pairs(query, int(need * 1.15) + 16, rng)
        off, ctrl, valid = tp.frontend_batch(pmap, params, S, G)
        for b in range(len(S)):
            if valid[b] and off[b + 1] - off[b] >= 7 and len(offs) - 1 < B:
                chunks.append(ctrl[off[b]:off[b + 1]])
                offs.append(offs[-1] + len(chunks[-1]))
    return np.array(offs, np.int32), np.concatenate(chunks, 0)


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


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path on all host cores, same config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import trajectory_planner_b200 as tp
    pmap = tp.OccMap.from_tpm(MAP_TPM)
    p = tp.default_params()
    info = pmap.info()
    inflated = pmap.grid("inflated")

    def query_host(xyz):  # host-side occupancy lookup for workload generation only
        idx = np.floor((xyz - np.array(info["origin"])) / info["res"]).astype(int)
        inside = np.all((idx >= 0) & (idx < np.array(info["dims"])), axis=1)
        out = np.ones(len(xyz), np.uint8)
        ii = idx[inside]
        out[inside] = inflated[ii[:, 0], ii[:, 1], ii[:, 2]]
        return out

    threads = os.cpu_count() or 1
    # bounded sample per step so that the whole run ends within minutes
    sample = int(args.ref_sample)
    offsets, ctrl = make_workload(tp, pmap, query_host, sample, SEED, p)
    from oracle import oracle as O
    om = oracle_map(O, pmap)
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
    line = dict(metric="ViGO B-spline solves/sec", value=value, unit="solves/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * dt / args.steps, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f64", data="synthetic", impl="reference",
                config=dict(workload=f"batch of {sample} ViGO solves (bounded sample of the 4,096 batch), random "
                                     "start/goal pairs on square_static_map (0.1 m voxels), CPU path on host cores",
                            batch=sample, map="square_static", seed=SEED),
                cpu_baseline=dict(value=value, unit="solves/s", cores=threads, kind="port",
                                  sample=f"{sample} problems x {args.steps} steps, {threads} host threads; p50 "
                                         f"{np.median(ms):.3f} ms/solve"),
                e2e=dict(value=value, unit="solves/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                single_solve_p50_ms=float(np.median(ms)))
    print(json.dumps(line), flush=True)


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
                    kernel="k_solve<vector-free> (bsplineTraj::makePlan per thread block: segments, A*, guide points, "
                           "fused cost+gradient+L-BFGS, collision check, re-parameterisation)",
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
        # single-solve latency (batch of one, host buffers) — BASELINE.json "single-solve p50 ms"
        lat = []
        o1 = np.array([0, Ns[0]], np.int32)
        for i in range(60):
            t0 = time.perf_counter()
            eng.make_plan_batch(p, o1, ctrl[:Ns[0]])
            lat.append(1e3 * (time.perf_counter() - t0))
        line["single_solve_p50_ms"] = float(np.median(lat[10:]))
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
        # map-query sweep (BASELINE.json configs[4] in small): 16 M random points, HBM-bound stream
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
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = peaks.get("hbm_gbs", 6650.0)
        qach = nq * 57.0 / (qms * 1e-3) / 1e9
        line["roofline_query"] = dict(bound="hbm", achieved=qach, peak=hbm, unit="GB/s", frac=qach / hbm, traffic=None,
                                      kernel="k_query_points", points=nq, ms=qms,
                                      bytes_per_query="24 B xyz + 1 B flag + 32 B map sector = 57 B (SURVEY.md §8d)",
                                      peak_source="MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650")
        line["l2_gather_gbs"] = eng.microbench_gather(12_500_000)
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
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
