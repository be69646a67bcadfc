#!/usr/bin/env python
"""Collision-query sweep (BASELINE.json configs[4]): 1 M trajectory samples against box.bt and synthetic grids up
to 1e8 voxels; algorithmic GB/s (57 B / query: 24 B xyz + 1 B flag + one 32 B map sector) against the measured HBM
copy peak and the measured L2 random-sector gather rate.  Prints one JSON line."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import trajectory_planner_b200 as tp

dev = torch.device("cuda", 0)
eng = tp.Engine(0)
tstream = torch.cuda.Stream(device=dev)   # the launching stream (a NULL handle would select the engine's own stream)
torch.cuda.set_stream(tstream)
stream = tstream.cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
hbm = peaks.get("hbm_gbs", 6650.0)
out = dict(metric="collision queries", unit="GB/s (57 B/query)", hbm_peak=hbm, cases=[])


def timed(nq, q, hit, reps=5):
    for _ in range(3):
        eng.query_points_device(nq, q.data_ptr(), hit.data_ptr(), stream)
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.query_points_device(nq, q.data_ptr(), hit.data_ptr(), stream); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


def case(name, pmap, nq, coherent=False):
    eng.set_map(pmap)
    info = pmap.info()
    lo = np.array(info["origin"]); hi = lo + np.array(info["dims"]) * info["res"]
    g = torch.Generator(device=dev); g.manual_seed(1)
    if coherent:   # 2048 trajectories x 512 consecutive samples, 2.5 cm apart along random directions
        nt, ns = nq // 512, 512
        p0 = torch.rand((nt, 1, 3), generator=g, device=dev, dtype=torch.float64) * torch.tensor(hi - lo, device=dev) + torch.tensor(lo, device=dev)
        d = torch.randn((nt, 1, 3), generator=g, device=dev, dtype=torch.float64); d = d / d.norm(dim=2, keepdim=True)
        q = (p0 + d * (0.025 * torch.arange(ns, device=dev, dtype=torch.float64))[None, :, None]).reshape(-1, 3).contiguous()
    else:
        q = torch.rand((nq, 3), generator=g, device=dev, dtype=torch.float64) * torch.tensor(hi - lo, device=dev) + torch.tensor(lo, device=dev)
    hit = torch.empty(len(q), dtype=torch.uint8, device=dev)
    ms = timed(len(q), q, hit)
    gbs = len(q) * 57.0 / (ms * 1e-3) / 1e9
    out["cases"].append(dict(map=name, cells=int(np.prod(info["dims"])), packed_MB=info["packed_bytes"] / 1e6, queries=len(q),
                             pattern="trajectory-coherent" if coherent else "uniform", ms=ms, gqueries_per_s=len(q) / ms / 1e6,
                             gbs=gbs, frac_hbm=gbs / hbm, hit_rate=float(hit.float().mean().item())))


box = tp.OccMap.from_tpm(os.path.join(ROOT, "data", "maps", "box.tpm"))
case("box.bt", box, 1 << 20)
case("box.bt", box, 1 << 20, coherent=True)
case("box.bt", box, 16 << 20)
rng = np.random.default_rng(0)
for dims in ((100, 100, 100), (400, 250, 100), (1000, 1000, 100)):
    m = tp.OccMap(0.1, (0.0, 0.0, 0.0), dims, (0, 0, 0))
    nbox = int(0.1 * np.prod(dims) / (20 * 20 * 10))
    cells = []
    for _ in range(nbox):
        c = rng.integers(0, np.array(dims) - np.array([20, 20, 10]))
        ii, jj, kk = np.meshgrid(np.arange(20), np.arange(20), np.arange(10), indexing="ij")
        cells.append(np.stack([ii.ravel() + c[0], jj.ravel() + c[1], kk.ravel() + c[2]], 1))
    m.add_cells(np.concatenate(cells).astype(np.int32))
    case(f"synthetic {np.prod(dims):.0e} voxels, 10 % boxes", m, 16 << 20)
out["l2_gather_gbs_12p5MB"] = eng.microbench_gather(12_500_000)
print(json.dumps(out))
