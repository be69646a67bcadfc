#!/usr/bin/env python
"""BASELINE.json configs[2] on one GPU: ViGO solves on maze.bt and tunnel.bt rasterised to the occMap contract (native
0.1 m resolution, inflation (4,4,2)); start/goal uniform over free inflated cells of the z = 1.0 slab, 2-20 m apart, half of
the batch per map (OM_B trajectories per map, default 4,096 = the per-GPU share of 65,536 over 8 GPUs).  The (start, goal)
pairs go through the DEVICE front end; the solve is timed device-resident with CUDA events.  Prints one JSON line."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import trajectory_planner_b200 as tp

B = int(os.environ.get("OM_B", "4096"))
out = {}
for name in ("maze", "tunnel"):
    m = tp.OccMap.from_tpm(os.path.join(ROOT, "data", "maps", name + ".tpm"))
    info = m.info()
    e = tp.Engine(0)
    e.set_map(m)
    p = tp.default_params()
    inf = m.grid("inflated")
    kz = int(np.floor((1.0 - info["origin"][2]) / info["res"]))
    free = np.argwhere(inf[:, :, kz] == 0)
    rng = np.random.default_rng(20261018)
    org, res = np.array(info["origin"]), info["res"]
    offs, chunks, fe_ms = [0], [], 0.0
    while len(offs) - 1 < B:
        n = int((B - (len(offs) - 1)) * 1.5) + 64
        a, b = free[rng.integers(len(free), size=n)], free[rng.integers(len(free), size=n)]
        S = org + (np.column_stack([a[:, 0], a[:, 1], np.full(n, kz)]) + 0.5) * res
        G = org + (np.column_stack([b[:, 0], b[:, 1], np.full(n, kz)]) + 0.5) * res
        S[:, 2] = G[:, 2] = 1.0
        d = np.linalg.norm(S - G, axis=1)
        S, G = S[(d >= 2) & (d <= 20)], G[(d >= 2) & (d <= 20)]
        t0 = time.perf_counter()
        off, ctrl, valid = e.frontend_batch(p, S, G)
        fe_ms += 1e3 * (time.perf_counter() - t0)
        for i in range(len(S)):
            if valid[i] and off[i + 1] - off[i] >= 7 and len(offs) - 1 < B:
                chunks.append(ctrl[off[i]:off[i + 1]])
                offs.append(offs[-1] + len(chunks[-1]))
    offsets, ctrl = np.array(offs, np.int32), np.concatenate(chunks, 0)
    dev = torch.device("cuda", 0)
    d_off, d_in = torch.from_numpy(offsets).to(dev), torch.from_numpy(ctrl).to(dev)
    d_out = torch.empty_like(d_in)
    d_res = torch.empty(B * tp.RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)
    for _ in range(3):
        e.make_plan_batch_device(p, B, d_off.data_ptr(), d_in.data_ptr(), d_out.data_ptr(), d_res.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a_.record()
    K = 3
    for _ in range(K):
        e.make_plan_batch_device(p, B, d_off.data_ptr(), d_in.data_ptr(), d_out.data_ptr(), d_res.data_ptr(), st.cuda_stream)
    b_.record()
    torch.cuda.synchronize()
    ms = a_.elapsed_time(b_) / K
    r = np.frombuffer(d_res.cpu().numpy().tobytes(), dtype=tp.RESULT_DTYPE)
    Ns = np.diff(offsets)
    out[name] = dict(batch=B, ms=ms, solves_per_s=B / (ms * 1e-3), success_rate=float(np.mean(r["status"] == 1)),
                     control_points=dict(min=int(Ns.min()), mean=float(Ns.mean()), max=int(Ns.max())),
                     lbfgs_iters_per_solve=float(r["lbfgs_iters"].mean()), astar_expansions_per_solve=float(r["astar_expansions"].mean()),
                     grid=info["dims"], front_end_device_ms=fe_ms)
    e.close()
tot = sum(v["batch"] for v in out.values()) / sum(v["ms"] * 1e-3 for v in out.values())
print(json.dumps(dict(metric="ViGO B-spline solves/sec", workload="maze.bt + tunnel.bt rasters, half the batch each (configs[2], one GPU's share)",
                      value=tot, unit="solves/s", per_map=out)))
