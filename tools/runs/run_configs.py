#!/usr/bin/env python3
"""Times the other BASELINE.json configs (1, 3, 4, 5) on the GPU(s) visible to this process -- they are
parity-test cases, not bench lines, but their numbers go into DESIGN.md.  Under torchrun every term is sharded
by lane range (configs 1/3/4) or one sensor is rendered per rank (config 5).
    python tools/run_configs.py [--configs 1,3,4,5] [--reps 3] [--scale 1.0]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import psdr_jit_b200 as psdr  # noqa: E402
from psdr_jit_b200 import scenes  # noqa: E402


def base_scene(w, h, spp, sppe, sppse, bsdfs, meshes, cams=None, envmap=None, rank=0, world=1):
    sc = psdr.Scene()
    o = sc.opts
    o.width, o.height, o.spp, o.sppe, o.sppse, o.log_level = w, h, spp, sppe, sppse, 0
    for cam in (cams or [scenes.CBOX_CAMERA]):
        s = psdr.PerspectiveCamera(cam["fov"], cam["near"], cam["far"])
        s.to_world = cam["to_world"]
        sc.add_Sensor(s)
    for name, p in bsdfs:
        if len(p) == 3 and hasattr(p[0], "__len__"):
            sc.add_BSDF(psdr.MicrofacetBSDF(p[0], p[1], p[2]), name)
        else:
            sc.add_BSDF(psdr.DiffuseBSDF(p), name)
    if envmap is not None:
        e = psdr.EnvironmentMap(psdr.Bitmap3fD(envmap[1], envmap[2], envmap[0]))
        sc.add_EnvironmentMap(e)
    for m in meshes:
        mesh = psdr.Mesh()
        mesh.load_raw(m.v, m.f, m.uv, m.fuv)
        mesh.to_world = m.to_world
        sc.add_Mesh(mesh, m.bsdf, psdr.AreaLight(m.emitter) if m.emitter is not None else None)
    sc.set_shard(rank, world)
    return sc


def timed(fn, reps):
    fn(0)
    torch.cuda.synchronize()
    ts = []
    for it in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(it + 1)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,3,4,5")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--scale", type=float, default=1.0, help="scales spp of configs 3/4 (smoke runs)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tangent = np.zeros((4, 4), np.float32)
    tangent[0, 3] = 100.0
    out = []

    def reduce_(*ts):
        if world > 1:
            buf = torch.stack(ts)
            dist.all_reduce(buf)
            return buf
        return ts

    for cfg in [int(c) for c in args.configs.split(",")]:
        if cfg == 1:
            sc = base_scene(128, 128, 1, 0, 0, scenes.CBOX_BSDFS, scenes.cbox_meshes(), rank=rank, world=world)
            sc.configure(); sc.configure([0])
            integ = psdr.PathTracer(1)
            ms = timed(lambda it: reduce_(integ.renderC(sc, 0, seed=it)), args.reps)
            n = 128 * 128
            out.append({"config": 1, "what": "cbox 128^2 spp=1 depth=1 renderC", "ms": ms, "msamples_s": n / ms / 1e3})
        elif cfg == 3:
            spp = max(1, int(128 * args.scale))
            rng = np.random.default_rng(0)
            env = ((rng.random((512 * 1024, 3), dtype=np.float32) ** 4) * 4).astype(np.float32)
            mf = [("light", (0.0, 0.0, 0.0))] + [(n_, ((.2, .9, .9), (.01, .01, .01), 0.3)) for n_ in ("cat", "white", "green", "red")]
            sc = base_scene(1024, 1024, spp, 0, 0, mf, scenes.cbox_meshes(), envmap=(env, 1024, 512), rank=rank, world=world)
            sc.param_map["Mesh[0]"].set_transform(np.eye(4, dtype=np.float32), tangent=tangent)
            sc.configure(); sc.configure([0])
            integ = psdr.PathTracer(6)
            ms = timed(lambda it: reduce_(*integ.renderD_fwd(sc, 0, seed=it)), args.reps)
            cot = torch.ones((1024 * 1024, 3), device="cuda")
            ms_v = timed(lambda it: (integ.renderD_primal(sc, 0, seed=it), integ.render_vjp(sc, cot, 0, seed=it)), args.reps)
            n = 1024 * 1024 * spp
            out.append({"config": 3, "what": "cbox 1024^2 spp=%d depth=6 renderD Microfacet + envmap 1024x512" % spp, "ms_jvp": ms,
                        "msamples_s_jvp": n / ms / 1e3, "ms_vjp_step": ms_v, "msamples_s_vjp": n / ms_v / 1e3, "configure_ms": sc.last_configure_ms()})
        elif cfg == 4:
            spp = max(1, int(64 * args.scale))
            blob = scenes.icosphere(4, 90.0, (300.0, 200.0, 250.0))          # 5120 faces: bunny_low-sized stand-in (4968 faces)
            rng = np.random.default_rng(1)
            blob.v[:] = (blob.v - np.array([300.0, 200.0, 250.0], np.float32)) * (1.0 + 0.15 * rng.standard_normal((len(blob.v), 1)).astype(np.float32)) \
                + np.array([300.0, 200.0, 250.0], np.float32)
            sc = base_scene(512, 512, spp, 0, spp, scenes.CBOX_BSDFS, scenes.cbox_meshes() + [blob], rank=rank, world=world)
            sc.param_map["Mesh[8]"].set_transform(np.eye(4, dtype=np.float32), tangent=tangent)
            sc.configure(); sc.configure([0])
            integ = psdr.PathTracer(3)
            t0 = time.perf_counter()
            integ.preprocess_secondary_edges(sc, 0, [2000, 5, 5, 32], 1)
            torch.cuda.synchronize()
            prep = (time.perf_counter() - t0) * 1e3
            ms = timed(lambda it: reduce_(*integ.renderD_fwd(sc, 0, seed=it)), args.reps)
            n = 512 * 512 * 2 * spp
            out.append({"config": 4, "what": "cbox + 5120-face blob 512^2 spp=sppse=%d depth=3 guided renderD (BVH2)" % spp, "ms_jvp": ms,
                        "msamples_s_jvp": n / ms / 1e3, "guiding_prepass_ms": prep, "configure_ms": sc.last_configure_ms(),
                        "secondary_edges": sc.num_secondary_edges()})
        elif cfg == 5:
            cams = []
            for k in range(8):
                a = 2 * np.pi * k / 8
                eye = np.array([278 + 900 * np.sin(a) * 0.35, 273 + 60 * np.cos(2 * a), -800 + 120 * (1 - np.cos(a))], np.float32)
                cams.append(dict(fov=60.0, near=1e-6, far=1e7, to_world=scenes.translate(*eye)))
            sc = base_scene(512, 512, 32, 0, 0, scenes.CBOX_BSDFS, scenes.cbox_meshes(), cams=cams)
            sc.param_map["Mesh[0]"].set_transform(np.eye(4, dtype=np.float32), tangent=tangent)
            sc.configure(); sc.configure(list(range(8)))
            integ = psdr.PathTracer(3)
            mine = [k for k in range(8) if k % world == rank]

            def step(it):
                for k in mine:
                    integ.renderD_fwd(sc, k, seed=it)
            ms = timed(step, args.reps)
            n = 512 * 512 * 32 * 8
            out.append({"config": 5, "what": "8 sensors 512^2 spp=32 depth=3 renderD, sensor k on rank k %% %d" % world, "ms_jvp": ms,
                        "msamples_s_jvp": n / ms / 1e3})
    if world > 1:
        t = torch.tensor([o.get("ms_jvp", o.get("ms", 0.0)) for o in out], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        for o, v in zip(out, t.tolist()):
            o["ms_max_over_ranks"] = v
    if rank == 0:
        for o in out:
            o["n_gpus"] = world
            print(json.dumps(o), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
