# k_query_points timing for the current TP_QUERY_MINB: 16 M and 1 M uniform points on the square map and box.bt
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import trajectory_planner_b200 as tp, bench
dev = torch.device("cuda", 0)
eng = tp.Engine(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = torch.cuda.Stream(device=dev); torch.cuda.set_stream(ts); s = ts.cuda_stream
for name, path, lo, hi in (("square", bench.MAP_TPM, (-12, -12, 0), (12, 12, 2.8)), ("box", os.path.join(bench.ROOT, "data", "maps", "box.tpm"), None, None)):
    m = tp.OccMap.from_tpm(path); eng.set_map(m); info = m.info()
    if lo is None:
        lo = info["origin"]; hi = np.array(info["origin"]) + np.array(info["dims"]) * info["res"]
    for nq in (1 << 20, 16 << 20):
        q = torch.rand((nq, 3), dtype=torch.float64, device=dev) * torch.tensor(np.array(hi) - np.array(lo), device=dev) + torch.tensor(np.array(lo, float), device=dev)
        hit = torch.empty(nq, dtype=torch.uint8, device=dev)
        for _ in range(3): eng.query_points_device(nq, q.data_ptr(), hit.data_ptr(), s)
        torch.cuda.synchronize()
        ms = []
        for _ in range(7):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); eng.query_points_device(nq, q.data_ptr(), hit.data_ptr(), s); b.record(); torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        t = float(np.median(ms))
        print(f"minb {os.environ.get('TP_QUERY_MINB','4')} {name} {nq>>20} M: {1e3*t:.1f} us, {nq*25/t/1e6:.0f} GB/s HBM-side ({nq*25/t/1e6/6543.4:.3f} of copy peak), hits {float(hit.float().mean()):.4f}")
