"""world_size-2 gloo test of the multi-GPU host logic: block-cyclic lane shards of every term,
accumulated into private full-frame images and summed with ONE all-reduce, reproduce the full image."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from psdr_jit_b200.dist import all_reduce_images, shard_lanes
    from tests.common import build_oracle, scenes
    w = h = 32
    spp = 4
    osc = build_oracle(scenes.cbox_meshes(), w, h, spp, 0, 0, move_mesh=0, axis_scale=(100.0, 0.0, 0.0))
    img, dimg, lanes = osc.render(2, seed=3, mode=1, terms=1, lane_out=True)
    mine = shard_lanes(w * h * spp, rank, world)
    part = np.zeros((w * h, 3), dtype=np.float64)
    np.add.at(part, mine // spp, lanes[mine].astype(np.float64) / spp)
    t = torch.from_numpy(part)
    t2 = torch.full((5,), float(rank + 1), dtype=torch.float64)
    (tot, tot2) = all_reduce_images(t, t2)
    if rank == 0:
        q.put((float(np.abs(tot.numpy() - img).max()), float(np.abs(img).max()), tot2.tolist(), (int(mine[0]), int(mine[32]), len(mine))))
    dist.destroy_process_group()


def test_lane_shards_allreduce_to_full_image(oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, mx, tot2, rng = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-5 * mx
    assert tot2 == [3.0] * 5
    assert rng == (0, 64, 2048)


def _worker_shared(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from psdr_jit_b200.dist import SharedHostBuffer
    n = 1001
    buf = SharedHostBuffer(n, rank, world, "t")
    ok = True
    for step in (1, 2, 3):
        src = torch.arange(n, dtype=torch.float32) * step          # every rank holds the complete result
        buf.gather(src, step)
        if rank == 0:
            full = buf.wait_all(step)
            ok = ok and bool(torch.equal(full, src))
        dist.barrier()                                             # the next step overwrites the buffer
    if rank == 0:
        q.put((ok, list(buf.bounds)))
    buf.close()
    dist.destroy_process_group()


def test_shared_host_buffer_gathers_slices_of_every_rank():
    """the N-way split of the device->host copy (psdr_jit_b200.dist.SharedHostBuffer): each rank writes its slice of a
    result all ranks hold, rank 0 sees the whole buffer once every rank has published the step"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_shared, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, bounds = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and bounds == [0, 504, 1001]


def test_shard_lanes_properties():
    from psdr_jit_b200.dist import shard_lanes
    for n in (0, 31, 32, 1000, 65536, 123457):
        for world in (1, 2, 3, 8):
            parts = [shard_lanes(n, r, world) for r in range(world)]
            allv = np.sort(np.concatenate(parts)) if n else np.zeros(0, np.int64)
            assert np.array_equal(allv, np.arange(n))                     # a partition of the lanes
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 32                           # balanced to one block
            for r, p in enumerate(parts):
                assert np.all((p // 32) % world == r)                      # block b belongs to rank b % world
    with pytest.raises(ValueError):
        shard_lanes(10, 2, 2)


def test_ordered_deal_properties():
    """the deal of bucket-ordered secondary-edge samples to warps (csrc/device_path.cuh sec_edge_batches, mirrored by
    dist.ordered_deal): a partition of the positions, whole 256-blocks, every warp spread over the whole order"""
    from psdr_jit_b200.dist import ordered_deal
    for span in (1, 255, 256, 40255, 32768, 1048576 + 77):
        for n_warps in (4, 1260, 148 * 32):
            parts = [ordered_deal(span, n_warps, w) for w in range(n_warps)]
            allv = np.sort(np.concatenate(parts))
            assert np.array_equal(allv, np.arange(span))
            for w in (0, n_warps // 2, n_warps - 1):
                p = parts[w]
                if len(p) == 0:
                    continue
                assert np.all((p // 256) % n_warps == w)                   # block b belongs to warp b % n_warps
                assert np.all(np.diff(p) > 0)                              # walked front to back
            sizes = np.array([len(p) for p in parts])
            assert sizes.max() - sizes.min() <= 256 or span < n_warps * 256
