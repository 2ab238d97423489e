"""2-GPU NCCL test of the sharded render path: block-cyclic lane shards + ONE all-reduce of the
[2, npix, 3] image buffer reproduce the single-GPU images, and the autograd path (forward all-reduce of the image, ONE
all-reduce of the flat device gradient table in backward) reproduces the single-GPU parameter gradient of a nonlinear
loss.  Skipped on boxes with fewer than two GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py`)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


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
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import psdr_jit_b200 as psdr
    from tests.common import build_product, scenes
    kw = dict(move_mesh=0, axis_scale=(100.0, 0.0, 0.0))
    integ = psdr.PathTracer(3)
    # ---- forward mode
    part = build_product(scenes.cbox_meshes(), 96, 96, 8, 8, 8, shard=(rank, world), **kw)
    integ.renderD_fwd(part, 0, seed=2)
    dist.all_reduce(integ.last_buffer)
    summed = integ.last_buffer.clone()
    out = {}
    if rank == 0:
        full = build_product(scenes.cbox_meshes(), 96, 96, 8, 8, 8, **kw)
        img, dimg = integ.renderD_fwd(full, 0, seed=2)
        out["fwd_img"] = float((summed[0] - img).norm() / img.norm())
        out["fwd_dimg"] = float((summed[1] - dimg).norm() / dimg.norm())
    # ---- reverse mode through autograd with a nonlinear loss
    def grad_of_loss(shard):
        sc = build_product(scenes.cbox_meshes(), 96, 96, 8, 8, 8, shard=shard)
        t = torch.eye(4, dtype=torch.float32, requires_grad=True)
        sc.param_map["Mesh[0]"].set_transform(t)
        r = torch.tensor([20.0, 20.0, 8.0], requires_grad=True)
        sc.param_map["Emitter[0]"].radiance = r
        sc.configure([0])
        img = integ.renderD(sc, 0, seed=2)
        target = torch.full_like(img, 0.3)
        loss = ((img - target) ** 2).sum()
        loss.backward()
        return float(loss), t.grad.clone(), r.grad.clone()
    loss_s, gt_s, gr_s = grad_of_loss((rank, world))
    if rank == 0:
        loss_f, gt_f, gr_f = grad_of_loss(None)
        out["loss"] = abs(loss_s - loss_f) / abs(loss_f)
        out["grad_transform"] = float((gt_s - gt_f).norm() / gt_f.norm())
        out["grad_radiance"] = float((gr_s - gr_f).norm() / gr_f.norm())
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_sharded_render_and_autograd():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    print(out)
    assert out["fwd_img"] < 1e-6 and out["fwd_dimg"] < 1e-4, out
    assert out["loss"] < 1e-5 and out["grad_transform"] < 2e-4 and out["grad_radiance"] < 2e-4, out


def _worker_peer(rank, world, port, q):
    """The fused path: kernels add through the NVLS multicast address (dist.PeerBuffers) -- no all-reduce anywhere."""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import psdr_jit_b200 as psdr
    from tests.common import build_product, scenes
    kw = dict(move_mesh=0, axis_scale=(100.0, 0.0, 0.0))
    integ = psdr.PathTracer(3)
    part = build_product(scenes.cbox_meshes(), 96, 96, 8, 8, 8, shard=(rank, world), **kw)
    out = {"multicast": bool(part.enable_peer_reduction())}
    if out["multicast"]:
        full = build_product(scenes.cbox_meshes(), 96, 96, 8, 8, 8, **kw)
        # several steps: the double-buffer protocol (zeroing, reuse) must hold beyond the first use of each buffer
        errs = []
        for step in range(5):
            img, dimg = integ.renderD_fwd(part, 0, seed=step)
            assert integ.last_reduced
            ref_img, ref_dimg = integ.renderD_fwd(full, 0, seed=step)
            errs.append((float((img - ref_img).norm() / ref_img.norm()), float((dimg - ref_dimg).norm() / ref_dimg.norm())))
        out["fwd_img"], out["fwd_dimg"] = max(e[0] for e in errs), max(e[1] for e in errs)
        # the device->host copy split over the ranks (dist.SharedHostBuffer): every rank holds the summed result, each
        # copies its slice into ONE page-locked buffer all ranks map, rank 0 reads the whole frame
        from psdr_jit_b200.dist import SharedHostBuffer
        shared = SharedHostBuffer(2 * 96 * 96 * 3, rank, world, "t2")
        bad = 0.0
        for step in range(1, 4):
            img, dimg = integ.renderD_fwd(part, 0, seed=10 + step)
            shared.gather(integ.last_buffer, step)
            if rank == 0:
                host = shared.wait_all(step).view(2, -1, 3)
                bad = max(bad, float((host[0] - img.cpu()).abs().max()), float((host[1] - dimg.cpu()).abs().max()))
        out["shared_host"] = bad
        out["shared_registered"] = bool(shared.registered)
        shared.close()
        imgc = integ.renderC(part, 0, seed=3)
        out["renderC"] = float((imgc - integ.renderC(full, 0, seed=3)).norm() / imgc.norm())

        def grad_of_loss(shard, peer):
            sc = build_product(scenes.cbox_meshes(), 96, 96, 8, 8, 8, shard=shard)
            if peer:
                assert sc.enable_peer_reduction()
            t = torch.eye(4, dtype=torch.float32, requires_grad=True)
            sc.param_map["Mesh[0]"].set_transform(t)
            r = torch.tensor([20.0, 20.0, 8.0], requires_grad=True)
            sc.param_map["Emitter[0]"].radiance = r
            sc.configure([0])
            res = []
            for step in range(3):
                if t.grad is not None:
                    t.grad = None
                    r.grad = None
                img = integ.renderD(sc, 0, seed=2 + step)
                loss = ((img - torch.full_like(img, 0.3)) ** 2).sum()
                loss.backward()
                res.append((float(loss), t.grad.clone(), r.grad.clone()))
            return res
        a = grad_of_loss((rank, world), True)
        b = grad_of_loss(None, False)
        out["loss"] = max(abs(x[0] - y[0]) / abs(y[0]) for x, y in zip(a, b))
        out["grad_transform"] = max(float((x[1] - y[1]).norm() / y[1].norm()) for x, y in zip(a, b))
        out["grad_radiance"] = max(float((x[2] - y[2]).norm() / y[2].norm()) for x, y in zip(a, b))
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_fused_multicast_reduction():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_peer, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    print(out)
    if not out["multicast"]:
        pytest.skip("no NVLS multicast support on this node")
    assert out["fwd_img"] < 1e-6 and out["fwd_dimg"] < 1e-4 and out["renderC"] < 1e-6, out
    assert out["shared_host"] == 0.0 and out["shared_registered"], out
    assert out["loss"] < 1e-5 and out["grad_transform"] < 2e-4 and out["grad_radiance"] < 2e-4, out
