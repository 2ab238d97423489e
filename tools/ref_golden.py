#!/usr/bin/env python3
"""Generate golden vectors by RUNNING the unmodified reference (baseline/_ref) on a GPU box.

Usage (through gpurun, from the repo root):
    python tools/ref_golden.py [--full]

Only the reference's public Python API is used (none of this repo's kernels).  Results go to
``gpurun_out/ref_golden/*.npz`` + ``gpurun_out/ref_golden/log.json``; the small ones are then
committed under ``tests/golden/`` (see tests/golden/README.md).  Scenes come from
``psdr_jit_b200/scenes.py`` (pure numpy) written out as triangle-only OBJ files so that the
reference and this repo consume an identical triangle list.
"""
import json
import os
import subprocess
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out", "ref_golden")
os.makedirs(OUT, exist_ok=True)
LOG = {"sections": {}}


def save_log():
    with open(os.path.join(OUT, "log.json"), "w") as fh:
        json.dump(LOG, fh, indent=1, default=str)


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=60).stdout.strip()
    except Exception as e:  # noqa
        return "ERR %r" % (e,)


LOG["nvidia_smi"] = sh("nvidia-smi --query-gpu=name,driver_version --format=csv,noheader")
LOG["optix_libs"] = sh("ldconfig -p | grep -i -E 'optix|libcuda' ; find / -xdev -name 'libnvoptix*' 2>/dev/null | head")
LOG["nproc"] = sh("nproc")
save_log()
print(LOG, flush=True)

import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("scenes", os.path.join(ROOT, "psdr_jit_b200", "scenes.py"))
scenes = importlib.util.module_from_spec(spec)
sys.modules["scenes"] = scenes
spec.loader.exec_module(scenes)

import drjit  # noqa: E402
import psdr_jit as psdr  # noqa: E402
from drjit.cuda import Float as FloatC, Matrix4f as Matrix4fC, UInt64 as U64  # noqa: E402
from drjit.cuda.ad import Float as FloatD, Matrix4f as Matrix4fD  # noqa: E402

OBJDIR = os.path.join(OUT, "obj")
os.makedirs(OBJDIR, exist_ok=True)


def mat(m):
    return [[float(m[i][j]) for j in range(4)] for i in range(4)]


def build(meshes, w, h, spp, sppe, sppse, cam=None, log_level=0):
    cam = cam or scenes.CBOX_CAMERA
    sc = psdr.Scene()
    o = sc.opts
    o.width, o.height, o.spp, o.sppe, o.sppse, o.log_level = w, h, spp, sppe, sppse, log_level
    sensor = psdr.PerspectiveCamera(cam["fov"], cam["near"], cam["far"])
    sensor.to_world = Matrix4fD(mat(cam["to_world"]))
    sc.add_Sensor(sensor)
    for name, refl in scenes.CBOX_BSDFS:
        sc.add_BSDF(psdr.DiffuseBSDF([float(x) for x in refl]), name)
    for i, m in enumerate(meshes):
        path = os.path.join(OBJDIR, "m%d_%s.obj" % (i, m.name))
        scenes.write_obj(m, path)
        em = psdr.AreaLight([float(x) for x in m.emitter]) if m.emitter is not None else None
        sc.add_Mesh(path, Matrix4fC(mat(m.to_world)), m.bsdf, em)
    return sc


def T(x, y, z):
    return [[1., 0., 0., x], [0., 1., 0., y], [0., 0., 1., z], [0., 0., 0., 1.]]


def render_d(sc, integ, seed, mesh_id, axis_scale):
    """renderD + forward-mode derivative w.r.t. scalar P moving mesh `mesh_id` by axis_scale*P."""
    P = FloatD(0.)
    drjit.enable_grad(P)
    ax = axis_scale
    sc.param_map["Mesh[%d]" % mesh_id].set_transform(Matrix4fD(T(P * ax[0], P * ax[1], P * ax[2])))
    sc.configure()
    sc.configure([0])
    drjit.sync_thread()
    t0 = time.perf_counter()
    img = integ.renderD(sc, 0, seed=seed)
    drjit.eval(img)
    drjit.sync_thread()
    t1 = time.perf_counter()
    drjit.set_grad(P, 1.0)
    drjit.forward_to(img)
    g = drjit.grad(img)
    drjit.eval(g)
    drjit.sync_thread()
    t2 = time.perf_counter()
    return np.asarray(img.numpy(), dtype=np.float32), np.asarray(g.numpy(), dtype=np.float32), t1 - t0, t2 - t1


def section(name):
    def deco(fn):
        t0 = time.time()
        try:
            LOG["sections"][name] = {"ok": True, "info": fn()}
        except Exception as e:  # noqa
            LOG["sections"][name] = {"ok": False, "err": repr(e), "tb": traceback.format_exc()}
            print("SECTION FAILED", name, repr(e), flush=True)
        LOG["sections"][name]["secs"] = time.time() - t0
        save_log()
        print("section", name, LOG["sections"][name].get("ok"), "%.1fs" % LOG["sections"][name]["secs"], flush=True)
        return fn
    return deco


@section("sampler")
def _sampler():
    out = {}
    for seed in (0, 7):
        s1, s2 = psdr.Sampler(), psdr.Sampler()
        s1.seed(drjit.arange(U64, 8) + seed)
        s2.seed(drjit.arange(U64, 8) + seed)
        d1 = np.stack([np.asarray(s1.next_1d().numpy()) for _ in range(8)])
        v = s2.next_2d()
        out["draws_seed%d" % seed] = d1.astype(np.float32)
        out["next2d_seed%d" % seed] = np.stack([np.asarray(v[0].numpy()), np.asarray(v[1].numpy())]).astype(np.float32)
    # large lane index (checks 64-bit TEA lanes)
    s3 = psdr.Sampler()
    s3.seed(drjit.arange(U64, 8) + 8388600)
    out["draws_big"] = np.stack([np.asarray(s3.next_1d().numpy()) for _ in range(4)]).astype(np.float32)
    np.savez(os.path.join(OUT, "sampler.npz"), **out)
    return {k: v.tolist() for k, v in out.items() if "seed0" in k}


@section("pmf")
def _pmf():
    rng = np.random.default_rng(1)
    pmf = (rng.random(37).astype(np.float32) * 3 + 0.01).astype(np.float32)
    d = psdr.DiscreteDistribution()
    d.init(FloatC(pmf))
    smp = rng.random(4096).astype(np.float32)
    idx, p = d.sample(FloatC(smp))
    np.savez(os.path.join(OUT, "pmf.npz"), pmf=pmf, samples=smp, idx=np.asarray(idx.numpy()),
             p=np.asarray(p.numpy()), sum=np.asarray(d.sum.numpy()), pmf_norm=np.asarray(d.pmf().numpy()))
    return {"sum": float(np.asarray(d.sum.numpy())[0])}


@section("edges")
def _edges():
    meshes = scenes.cbox_meshes() + [scenes.icosphere(2, 80.0, (185.0, 250.0, 169.0))]
    out = {}
    for i, m in enumerate(meshes):
        path = os.path.join(OBJDIR, "e%d_%s.obj" % (i, m.name))
        scenes.write_obj(m, path)
        mesh = psdr.Mesh()
        mesh.load(path, False)
        ei = mesh.edge_indices()
        out["mesh%d" % i] = np.stack([np.asarray(ei[k].numpy()) for k in range(4)]).astype(np.int32)
    np.savez(os.path.join(OUT, "edges.npz"), **out)
    return {k: list(v.shape) for k, v in out.items()}


@section("cfg1_renderC")
def _cfg1():
    sc = build(scenes.cbox_meshes(), 128, 128, 1, 0, 0)
    sc.configure()
    sc.configure([0])
    integ = psdr.PathTracer(1)
    img = integ.renderC(sc, 0, seed=0)
    drjit.eval(img)
    drjit.sync_thread()
    a = np.asarray(img.numpy(), dtype=np.float32)
    ts = []
    for it in range(5):
        drjit.sync_thread()
        t0 = time.perf_counter()
        im = integ.renderC(sc, 0, seed=it)
        drjit.eval(im)
        drjit.sync_thread()
        ts.append(time.perf_counter() - t0)
    # depth-3, spp 4 primal too
    sc2 = build(scenes.cbox_meshes(), 128, 128, 4, 0, 0)
    sc2.configure()
    sc2.configure([0])
    b = np.asarray(psdr.PathTracer(3).renderC(sc2, 0, seed=3).numpy(), dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, "cfg1_renderC.npz"), img=a, img_d3_spp4_seed3=b)
    return {"mean": float(a.mean()), "times": ts}


@section("aov")
def _aov():
    out = {}
    for scene_name, meshes in (("cbox", scenes.cbox_meshes()),
                               ("cboxsphere", scenes.cbox_meshes() + [scenes.icosphere(2, 80.0, (185.0, 250.0, 169.0))])):
        sc = build(meshes, 128, 128, 1, 0, 0)
        sc.configure()
        sc.configure([0])
        for fld in ("segmentation", "position", "depth", "geoNormal", "shNormal", "uv"):
            integ = psdr.FieldExtractionIntegrator(fld)
            im = integ.renderC(sc, 0, seed=0)
            out["%s_%s" % (scene_name, fld)] = np.asarray(im.numpy(), dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, "aov.npz"), **out)
    return {k: float(np.nanmean(v)) for k, v in out.items()}


def golden_renderD(tag, meshes, w, h, spps, depth, seed, mesh_id, axis_scale):
    out = {}
    info = {}
    for name, (a, b, c) in spps.items():
        sc = build(meshes, w, h, a, b, c)
        integ = psdr.PathTracer(depth)
        img, g, t_r, t_f = render_d(sc, integ, seed, mesh_id, axis_scale)
        out["img_" + name] = img
        out["grad_" + name] = g
        info[name] = {"img_mean": float(img.mean()), "grad_abs_mean": float(np.abs(g).mean()),
                      "nan_img": int(np.isnan(img).sum()), "nan_grad": int(np.isnan(g).sum()), "t": [t_r, t_f]}
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **out)
    return info


@section("renderD_small_light")
def _rd_small():
    return golden_renderD("renderD_128_s4_d3_light", scenes.cbox_meshes(), 128, 128,
                          {"all": (4, 4, 4), "interior": (4, 0, 0), "primary": (0, 4, 0), "secondary": (0, 0, 4)},
                          3, 0, 0, (100.0, 0.0, 0.0))


@section("renderD_small_box")
def _rd_box():
    return golden_renderD("renderD_128_s4_d2_smallbox", scenes.cbox_meshes(), 128, 128,
                          {"all": (4, 4, 4), "interior": (4, 0, 0), "primary": (0, 4, 0), "secondary": (0, 0, 4)},
                          2, 5, 1, (0.0, 30.0, 50.0))


@section("renderD_sphere")
def _rd_sphere():
    meshes = scenes.cbox_meshes() + [scenes.icosphere(2, 80.0, (185.0, 250.0, 169.0))]
    return golden_renderD("renderD_128_s4_d2_sphere", meshes, 128, 128,
                          {"all": (4, 4, 4), "interior": (4, 0, 0), "primary": (0, 4, 0), "secondary": (0, 0, 4)},
                          2, 1, 8, (40.0, 20.0, 0.0))


@section("renderD_128_s32")
def _rd_s32():
    return golden_renderD("renderD_128_s32_d3_light", scenes.cbox_meshes(), 128, 128,
                          {"all": (32, 32, 32)}, 3, 0, 0, (100.0, 0.0, 0.0))


@section("cfg2_full")
def _cfg2():
    sc = build(scenes.cbox_meshes(), 512, 512, 32, 32, 32)
    integ = psdr.PathTracer(3)
    rows = []
    img = g = None
    for it in range(5):
        im, gg, t_r, t_f = render_d(sc, integ, it, 0, (100.0, 0.0, 0.0))
        if it == 0:
            img, g = im, gg
        rows.append({"it": it, "t_render": t_r, "t_forward": t_f,
                     "msamples_s": 512 * 512 * 96 / 1e6 / (t_r + t_f)})
        print(rows[-1], flush=True)
    np.savez_compressed(os.path.join(OUT, "cfg2_512_s32_d3_light.npz"), img=img, grad=g)
    return rows


save_log()
print(json.dumps({k: (v.get("ok"), v.get("err")) for k, v in LOG["sections"].items()}, indent=1))
