#!/usr/bin/env python
"""Writes data/workloads/square4096_pairs.npz: the (start, goal) pairs of bench.py's N = 1 workload (BASELINE.json
configs[1]: 4,096 ViGO solves on square_static_map, seed 20261018) and the control-point count of each.  Both arms of
bench.py start from this file: the product arm turns the pairs into control points with its own front end, the
reference arm with the numpy restatement (oracle/frontend_np.py) — without loading the product library.  CPU only."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import trajectory_planner_b200 as tp

pmap = tp.OccMap.from_tpm(bench.MAP_TPM)
p = tp.default_params()
inflated = pmap.grid("inflated")
info = pmap.info()


def query(xyz):
    idx = np.floor((xyz - np.array(info["origin"])) / info["res"]).astype(int)
    inside = np.all((idx >= 0) & (idx < np.array(info["dims"])), axis=1)
    out = np.ones(len(xyz), np.uint8)
    ii = idx[inside]
    out[inside] = inflated[ii[:, 0], ii[:, 1], ii[:, 2]]
    return out


S, G, offsets, ctrl = bench.make_workload(tp, pmap, query, 4096, bench.SEED, p, want_pairs=True)
out = os.path.join(ROOT, "data", "workloads", "square4096_pairs.npz")
np.savez_compressed(out, starts=S, goals=G, n_ctrl=np.diff(offsets).astype(np.int32), seed=bench.SEED)
print(out, os.path.getsize(out), "bytes;", len(S), "pairs; control points", int(offsets[-1]))
