# ncu target: k_query_points on 16 M uniform points against the square map (the bench's roofline_query case)
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import trajectory_planner_b200 as tp, bench
pmap = tp.OccMap.from_tpm(bench.MAP_TPM); eng = tp.Engine(0); eng.set_map(pmap)
dev = torch.device("cuda", 0)
nq = 16 << 20
q = torch.empty((nq, 3), dtype=torch.float64, device=dev).uniform_(-12.0, 12.0)
q[:, 2].uniform_(0.0, 2.8)
hit = torch.empty(nq, dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream(dev).cuda_stream
for _ in range(4):
    eng.query_points_device(nq, q.data_ptr(), hit.data_ptr(), s)
torch.cuda.synchronize()
print("ok", float(hit.float().mean().item()))
