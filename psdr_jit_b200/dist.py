"""Multi-GPU plumbing for the render path (new: the reference is single-process, single-GPU).

Every term of renderD is sharded by LANE RANGE: rank r renders lanes
[cut(r), cut(r+1)) of each term into a private full-frame buffer and the partial images are summed
with one all-reduce.  Lane -> random stream is a function of the global lane index, so the sum is
independent of the number of ranks up to float summation order.
"""
from __future__ import annotations


def shard_range(n: int, rank: int, world: int):
    """Lane range of `rank` -- the same cut psdr_b200's render_impl uses (csrc/capi.cpp shard_of):
    cuts are rounded down to a multiple of 32 so that warps stay pixel-aligned."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("invalid shard")

    def cut(r):
        return n if r == world else (n * r // world) // 32 * 32
    return cut(rank), cut(rank + 1)


def all_reduce_images(*tensors, group=None):
    """Sum partial full-frame images over ranks in ONE collective (NCCL on GPUs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return tensors
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    out, o = [], 0
    for t in tensors:
        out.append(flat[o:o + t.numel()].view_as(t))
        o += t.numel()
    return tuple(out)
