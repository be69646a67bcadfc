"""Multi-GPU sharding of a ViGO batch: independent trajectories, contiguous ranges, no exchange
step on the solve path.  One process per GPU (torchrun); `torch.distributed` is
plumbing only — NCCL/gloo is touched solely by the optional final gather of the results."""
import numpy as np


def shard_bounds(B, world):
    """Contiguous ranges of ceil(B/world) problems per rank -> list of (begin, end)."""
    per = -(-int(B) // int(world))
    return [(min(r * per, B), min((r + 1) * per, B)) for r in range(world)]


def shard_batch(offsets, ctrl, rank, world):
    """The rank's slice of a ragged batch: (local_offsets[b1-b0+1] rebased to 0, local_ctrl, (b0, b1))."""
    offsets = np.asarray(offsets, dtype=np.int32)
    ctrl = np.asarray(ctrl, dtype=np.float64).reshape(-1, 3)
    b0, b1 = shard_bounds(len(offsets) - 1, world)[rank]
    loc = offsets[b0:b1 + 1] - offsets[b0]
    return loc.astype(np.int32), ctrl[offsets[b0]:offsets[b1]].copy(), (b0, b1)


def gather_batch(local_ctrl, local_results, offsets, group=None, dst=None):
    """Optional final gather (off the timed solve path): every rank contributes its shard's control
    points and result records; returns (ctrl[sum N,3], results[B]) on every rank (all_gather) or on
    `dst` only (gather; None elsewhere).  Works with the nccl and gloo backends: payloads travel as
    padded uint8 tensors on the backend's device."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    offsets = np.asarray(offsets, dtype=np.int64)
    B = len(offsets) - 1
    bounds = shard_bounds(B, world)
    rec = local_results.dtype.itemsize
    sizes = [int(offsets[b1] - offsets[b0]) * 24 + (b1 - b0) * rec for b0, b1 in bounds]
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    pay = np.concatenate([np.ascontiguousarray(local_ctrl, np.float64).reshape(-1).view(np.uint8),
                          np.ascontiguousarray(local_results).view(np.uint8).reshape(-1)])
    assert pay.size == sizes[rank], (pay.size, sizes[rank])
    buf = torch.zeros(max(sizes), dtype=torch.uint8, device=dev)
    buf[:pay.size] = torch.from_numpy(pay).to(dev)
    if dst is None:
        outs = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(outs, buf, group=group)
    else:
        outs = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
        dist.gather(buf, outs, dst=dst, group=group)
        if rank != dst:
            return None
    ctrl = np.zeros((int(offsets[-1]), 3))
    res = np.zeros(B, dtype=local_results.dtype)
    for r, (b0, b1) in enumerate(bounds):
        raw = outs[r].cpu().numpy()
        nc = int(offsets[b1] - offsets[b0]) * 24
        ctrl[offsets[b0]:offsets[b1]] = raw[:nc].view(np.float64).reshape(-1, 3)
        res[b0:b1] = raw[nc:nc + (b1 - b0) * rec].view(local_results.dtype)
    return ctrl, res
