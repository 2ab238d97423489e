"""Multi-GPU plumbing for the render path (new: the reference is single-process, single-GPU).

Every term of renderD is sharded by LANES, dealt to the ranks round-robin in
blocks of 32 (one warp; one pixel at spp = 32): rank r renders blocks r, r + world, r + 2 world, ... of
each term into a private full-frame buffer and the partial images are summed with one all-reduce.
Lane -> random stream is a function of the global lane index, so the sum is independent of the number
of ranks up to float summation order.  (Contiguous lane ranges, round 1, left the rank that owns the
empty top rows of the Cornell-box image with a quarter of the interior work of the others.)
"""
from __future__ import annotations


def shard_lanes(n: int, rank: int, world: int):
    """Global lane indices rendered by `rank` -- the deal psdr_b200's render_impl uses (csrc/capi.cpp set_shard,
    csrc/device_path.cuh global_lane): 32-lane blocks, round-robin."""
    import numpy as np
    if world < 1 or not (0 <= rank < world):
        raise ValueError("invalid shard")
    blocks = (n + 31) // 32
    mine = np.arange(rank, blocks, world, dtype=np.int64)
    lanes = (mine[:, None] * 32 + np.arange(32, dtype=np.int64)[None, :]).reshape(-1)
    return lanes[lanes < n]


def all_reduce_images(*tensors, group=None, dst=None):
    """Sum partial full-frame images over ranks in ONE collective (NCCL on GPUs, gloo in the CPU tests).  A single
    contiguous tensor (e.g. the [2, npix, 3] image + derivative-image buffer of renderD_fwd) is reduced in place;
    several tensors are flattened into one message first.  dst = rank: reduce to that rank only."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return tensors
    if len(tensors) == 1 and tensors[0].is_contiguous():
        flat = tensors[0].view(-1)
    else:
        flat = torch.cat([t.reshape(-1) for t in tensors])
    if dst is None:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(flat, dst=dst, op=dist.ReduceOp.SUM, group=group)
    if len(tensors) == 1 and tensors[0].is_contiguous():
        return tensors
    out, o = [], 0
    for t in tensors:
        out.append(flat[o:o + t.numel()].view_as(t))
        o += t.numel()
    return tuple(out)
