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


def ordered_deal(span: int, n_warps: int, warp: int, block: int = 256):
    """Positions (in the bucketed sample order of csrc/edge_sort.cu) that warp `warp` of `n_warps` evaluates in the
    secondary-edge kernels -- the deal of csrc/device_path.cuh sec_edge_batches for ordered samples: every n_warps-th block of
    `block` consecutive positions, so that each warp walks the whole edge list and a block stays on one stretch of it."""
    import numpy as np
    per = (span + n_warps * block - 1) // (n_warps * block) * block
    t = np.arange(per, dtype=np.int64)
    j = ((t // block) * n_warps + warp) * block + t % block
    return j[j < span]


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


class PeerBuffers:
    """Fused render + reduction over NVLink (include/psdr_b200.h psdr_scene_set_output_multicast).

    Two symmetric float32 buffers of `numel` entries, mapped by every rank of `group` and bound to an NVLS multicast
    address (torch symmetric memory: cuMemCreate + cuMulticastBindMem underneath).  The term kernels of every rank
    accumulate into the multicast address with `multimem.red.add.f32`: the NVSwitch adds each contribution into the
    replica of EVERY GPU, so when the kernels of all ranks are done every rank holds the complete sum -- the all-reduce
    pass (and its extra read + write of the frame on every GPU) does not exist.

    Protocol of step k (b = k % 2), all stream-ordered, ONE device-side barrier per step:
      target()  -> (multicast address, local replica) of buffer b, which is all zeros on every rank
      ... the caller's kernels add into the multicast address ...
      finish()  -> barrier over the ranks (everyone's adds have landed), copy buffer b out, zero it for step k + 2.
    Buffer b is next written in step k + 2, i.e. after barrier k + 1, which every rank enters only after its zeroing
    of step k; step k + 1 writes the OTHER buffer, so a fast rank never adds into a replica that is still being read
    or cleared."""

    def __init__(self, numel: int, device, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        if not dist.is_initialized():
            raise RuntimeError("PeerBuffers needs an initialised process group")
        self.group = group if group is not None else dist.group.WORLD
        self.numel = int(numel)
        self.bufs, self.handles = [], []
        for _ in range(2):
            t = symm.empty(self.numel, dtype=torch.float32, device=device)
            h = symm.rendezvous(t, self.group)
            if not getattr(h, "multicast_ptr", 0):
                raise RuntimeError("this node has no NVLS multicast support (symmetric memory multicast_ptr = 0)")
            t.zero_()
            self.bufs.append(t)
            self.handles.append(h)
        self.k = 0
        torch.cuda.synchronize(device)
        self.handles[0].barrier(channel=0)
        torch.cuda.synchronize(device)

    def target(self):
        b = self.k % 2
        return int(self.handles[b].multicast_ptr), self.bufs[b]

    def finish(self, extract=None):
        """Barrier, then the summed buffer as a fresh tensor (`extract(buffer)` if given: the caller's own copy-out, e.g.
        the rgb part of float4 pixels); the slot is cleared for reuse."""
        b = self.k % 2
        self.handles[b].barrier(channel=0)
        out = extract(self.bufs[b]) if extract is not None else self.bufs[b].clone()
        self.bufs[b].zero_()
        self.k += 1
        return out


class SharedHostBuffer:
    """One page-locked host buffer mapped by every rank of a node: the device->host copy of a result that all ranks hold
    (the fused multicast reduction leaves the complete image on every GPU) is split N ways -- rank r copies slice r over
    ITS PCIe link -- instead of one rank moving the whole frame (6.3 MB per step for image + derivative image at 512^2:
    the largest fixed cost of the end-to-end step at 8 GPUs).

    The buffer is a file in /dev/shm mapped MAP_SHARED by all ranks and registered with CUDA (cudaHostRegister) in each;
    a second small mapping holds one step counter per rank.  gather(src, step): copy this rank's slice of the flat
    device tensor `src`, wait for the copy, publish `step`; wait_all(step) (typically on rank 0) returns once every
    rank's counter has reached `step` -- the whole buffer is then valid.  Plain host tensors work too (the CPU tests)."""

    def __init__(self, numel: int, rank: int, world: int, tag: str, group=None, dtype=None):
        import os
        import numpy as np
        import torch
        import torch.distributed as dist
        self.rank, self.world, self.numel = int(rank), int(world), int(numel)
        dtype = dtype or torch.float32
        base = "/dev/shm/psdr_b200_%s_%s" % (os.environ.get("MASTER_PORT", "0"), tag)
        self.paths = (base + ".data", base + ".flags")
        esize = torch.empty((), dtype=dtype).element_size()
        if self.rank == 0:
            for p, nbytes in zip(self.paths, (max(self.numel, 1) * esize, 8 * self.world)):
                with open(p, "wb") as f:
                    f.truncate(nbytes)
        if dist.is_available() and dist.is_initialized() and self.world > 1:
            dist.barrier(group=group)
        self.data = torch.from_file(self.paths[0], shared=True, size=max(self.numel, 1), dtype=dtype)[:self.numel]
        self.flags = np.memmap(self.paths[1], dtype=np.int64, mode="r+", shape=(self.world,))
        self.registered = False
        if torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self.data.data_ptr(), max(self.numel, 1) * esize, 0)
            self.registered = int(rc) == 0
        # slice r = [bounds[r], bounds[r + 1]): equal parts, multiples of 4 elements
        per = (self.numel + self.world - 1) // self.world
        per = (per + 3) // 4 * 4
        self.bounds = [min(r * per, self.numel) for r in range(self.world + 1)]
        if dist.is_available() and dist.is_initialized() and self.world > 1:
            dist.barrier(group=group)          # nobody unlinks or writes before everybody has mapped the files
        if self.rank == 0:
            for p in self.paths:               # the mappings stay valid; the names disappear with the job
                try:
                    os.unlink(p)
                except OSError:
                    pass

    def gather(self, src, step: int):
        """copy this rank's slice of the flat tensor `src` (device or host) into the shared buffer and publish `step`"""
        import torch
        lo, hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        if hi > lo:
            self.data[lo:hi].copy_(src.reshape(-1)[lo:hi], non_blocking=True)
            if src.is_cuda:
                torch.cuda.current_stream(src.device).synchronize()
        self.flags[self.rank] = int(step)

    def wait_all(self, step: int, timeout_s: float = 30.0):
        """spin until every rank has published `step` (or later); returns the full buffer"""
        import time
        t0 = time.perf_counter()
        while int(self.flags.min()) < int(step):
            if time.perf_counter() - t0 > timeout_s:
                raise RuntimeError("SharedHostBuffer.wait_all: ranks %s have not reached step %d" % ([r for r in range(self.world) if self.flags[r] < step], step))
        return self.data

    def close(self):
        import torch
        if self.registered:
            try:
                torch.cuda.cudart().cudaHostUnregister(self.data.data_ptr())
            except Exception:
                pass
            self.registered = False
