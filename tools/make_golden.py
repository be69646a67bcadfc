#!/usr/bin/env python
"""Generate tests/golden/vigo_golden.npz — golden input/output vectors for the ViGO solve path.

Run in the build container, where /root/reference exists:  python tools/make_golden.py
The outputs come from oracle/_ref/liborc_ref.so, i.e. the CPU restatement LINKED AGAINST THE
REFERENCE'S OWN solver/lbfgs.hpp (compiled from where it lies under /root/reference, use_ref_lbfgs=1):
the L-BFGS / More-Thuente iterate in these vectors is the reference's code, not this repo's port.
Everything else the reference needs (ROS, Eigen, map_manager) is absent from the container, so the cost
terms / A* / guide points in the vectors are the restatement's ("parity unpinned" there).

Two variants are stored: atan2 = libm (reference-faithful) and atan2 = the deterministic software
routine shared with the device (bit-comparable with the CUDA strict-order path).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import trajectory_planner_b200 as tp  # host-side map loader + front end only (no GPU needed)
from helpers import make_problems, oracle_map_from
from oracle import oracle as O

SEED = 20261018
B = 32


def main():
    L = O.lib(ref=True)
    assert L.L.orc_is_ref_build() == 1, "oracle/_ref was not built against the reference's lbfgs.hpp"
    pmap = tp.OccMap.from_tpm(os.path.join(ROOT, "data", "maps", "square_static.tpm"))
    om = oracle_map_from(O, pmap, ref=True)
    pr = make_problems(tp, pmap, om, B, SEED)
    off, ctrl = pr["offsets"], pr["ctrl"]
    out = dict(seed=SEED, offsets=off, ctrl=ctrl, starts=pr["starts"], goals=pr["goals"])
    rng = np.random.default_rng(11)
    pert = ctrl + rng.normal(0, 0.15, ctrl.shape)
    out["ctrl_perturbed"] = pert
    for soft in (0, 1):
        po = L.default_params()
        po.use_ref_lbfgs = 1
        po.soft_atan2 = soft
        tag = f"soft{soft}"
        g_cp, g_p, g_v, g_off = [], [], [], [0]
        f_all, grad_all, opt_ctrl, opt_stats, segs_all, nseg_all, hit_all = [], [], [], [], [], [], []
        for b in range(B):
            c = ctrl[off[b]:off[b + 1]]
            pl = O.Planner(om, po)
            pl.set_ctrl(c)
            hit_all.append(int(pl.has_collision()))
            s = pl.find_collision_seg()
            nseg_all.append(len(s))
            segs_all.append(np.asarray(s, np.int32).reshape(-1, 2))
            pl.init_guides()
            cp, gp, gv = pl.get_guides()
            g_cp.append(cp), g_p.append(gp), g_v.append(gv), g_off.append(g_off[-1] + len(cp))
            # cost + gradient at a perturbed point with the guides attached
            pc = pert[off[b]:off[b + 1]]
            pl2 = O.Planner(om, po)
            pl2.set_ctrl(pc)
            pl2.add_guides(cp, gp, gv)
            f, g, _ = pl2.cost(pc[3:-3].ravel())
            f_all.append(f), grad_all.append(g)
            # one optimize() from the seed control points
            o = pl.optimize()
            opt_ctrl.append(pl.get_ctrl())
            opt_stats.append([o["ret"], o["iters"], o["evals"], o["fx"]])
        out[f"{tag}_g_off"] = np.array(g_off, np.int32)
        out[f"{tag}_g_cp"] = np.concatenate(g_cp).astype(np.int32)
        out[f"{tag}_g_p"] = np.concatenate(g_p, 0)
        out[f"{tag}_g_v"] = np.concatenate(g_v, 0)
        out[f"{tag}_cost"] = np.array(f_all)
        out[f"{tag}_grad"] = np.concatenate(grad_all)
        out[f"{tag}_opt_ctrl"] = np.concatenate(opt_ctrl, 0)
        out[f"{tag}_opt_stats"] = np.array(opt_stats)
        out["has_collision"] = np.array(hit_all, np.uint8)
        out["nseg"] = np.array(nseg_all, np.int32)
        out["segs"] = np.concatenate(segs_all, 0) if segs_all else np.zeros((0, 2), np.int32)
        ok, plan_ctrl, st = O.make_plan_batch(om, po, off, ctrl, nthreads=4)
        out[f"{tag}_plan_ctrl"] = plan_ctrl
        out[f"{tag}_plan_stats"] = st
    # map-query vectors (bit-exact decisions), incl. voxel faces / outside / non-finite points
    n = 4096
    xyz = np.column_stack([rng.uniform(-21, 21, n), rng.uniform(-21, 21, n), rng.uniform(-0.5, 3.5, n)])
    xyz[:512] = np.round(xyz[:512], 1)
    xyz[512] = [1e300, 0, 1]
    xyz[513] = [np.nan, 0, 1]
    xyz[514] = [-20.0, -20.0, -0.1]
    xyz[515] = [20.0, 20.0, 2.9]
    b_ = xyz + rng.normal(0, 0.4, xyz.shape)
    out["q_xyz"], out["q_b"] = xyz, b_
    out["q_hit"], out["q_unknown"], out["q_line"] = om.query(xyz), om.query_unknown(xyz), om.query_lines(xyz, b_)
    # A* vectors
    starts, ends = [], []
    for b in range(B):
        c = ctrl[off[b]:off[b + 1]]
        for k in range(out["nseg"][b]):
            s0, s1 = out["segs"][int(out["nseg"][:b].sum()) + k]
            starts.append(c[s0]), ends.append(c[s1])
    starts += [[0.0, 0.0, 1.0], [-8.0, -8.0, 1.0], [0.0, 0.0, 2.5]]
    ends += [[30.0, 0.0, 1.0], [-7.0, -7.6, 1.0], [1.0, 0.5, 1.0]]
    pl = O.Planner(om, L.default_params())
    plen, pflat, pexp = [], [], []
    for s, e in zip(starts, ends):
        p_, ex = pl.astar(s, e)
        plen.append(-1 if p_ is None else len(p_))
        pexp.append(ex)
        if p_ is not None:
            pflat.append(p_)
    out["astar_starts"], out["astar_ends"] = np.array(starts), np.array(ends)
    out["astar_len"], out["astar_exp"] = np.array(plen, np.int32), np.array(pexp, np.int32)
    out["astar_paths"] = np.concatenate(pflat, 0)
    path = os.path.join(ROOT, "tests", "golden", "vigo_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", "plan success", out["soft0_plan_stats"]["success"].mean())


if __name__ == "__main__":
    main()
