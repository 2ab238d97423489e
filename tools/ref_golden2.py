#!/usr/bin/env python3
"""Second set of golden vectors from the RUNNING reference (baseline/_ref, GPU box only):
MicrofacetBSDF scenes (renderC + every term of renderD) and guided secondary-edge sampling
(PathTracer.preprocess_secondary_edges).  Output: gpurun_out/ref_golden2/*.npz + log.json; the small
files are committed under tests/golden/.  Only the reference's public Python API is used."""
import json
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
OUT = os.path.join(ROOT, "gpurun_out", "ref_golden2")
os.makedirs(OUT, exist_ok=True)
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("scenes", os.path.join(ROOT, "psdr_jit_b200", "scenes.py"))
scenes = importlib.util.module_from_spec(spec)
sys.modules["scenes"] = scenes
spec.loader.exec_module(scenes)

import drjit  # noqa: E402
import psdr_jit as psdr  # noqa: E402
from drjit.cuda import Matrix4f as Matrix4fC  # noqa: E402
from drjit.cuda.ad import Float as FloatD, Matrix4f as Matrix4fD  # noqa: E402

OBJDIR = os.path.join(OUT, "obj")
os.makedirs(OBJDIR, exist_ok=True)
LOG = {}


def mat(m):
    return [[float(m[i][j]) for j in range(4)] for i in range(4)]


def T(x, y, z):
    return [[1., 0., 0., x], [0., 1., 0., y], [0., 0., 1., z], [0., 0., 0., 1.]]


def is_mf(p):
    return len(p) == 3 and hasattr(p[0], "__len__")


def build(meshes, bsdfs, w, h, spp, sppe, sppse):
    cam = scenes.CBOX_CAMERA
    sc = psdr.Scene()
    o = sc.opts
    o.width, o.height, o.spp, o.sppe, o.sppse, o.log_level = w, h, spp, sppe, sppse, 0
    sensor = psdr.PerspectiveCamera(cam["fov"], cam["near"], cam["far"])
    sensor.to_world = Matrix4fD(mat(cam["to_world"]))
    sc.add_Sensor(sensor)
    for name, p in bsdfs:
        if is_mf(p):
            sc.add_BSDF(psdr.MicrofacetBSDF([float(x) for x in p[0]], [float(x) for x in p[1]], float(p[2])), name)
        else:
            sc.add_BSDF(psdr.DiffuseBSDF([float(x) for x in p]), name)
    for i, m in enumerate(meshes):
        path = os.path.join(OBJDIR, "m%d_%s.obj" % (i, m.name))
        scenes.write_obj(m, path)
        em = psdr.AreaLight([float(x) for x in m.emitter]) if m.emitter is not None else None
        sc.add_Mesh(path, Matrix4fC(mat(m.to_world)), m.bsdf, em)
    return sc


def render_d(sc, integ, seed, mesh_id, ax, prep=None):
    P = FloatD(0.)
    drjit.enable_grad(P)
    sc.param_map["Mesh[%d]" % mesh_id].set_transform(Matrix4fD(T(P * ax[0], P * ax[1], P * ax[2])))
    sc.configure()
    sc.configure([0])
    if prep is not None:
        with drjit.suspend_grad():
            integ.preprocess_secondary_edges(sc, 0, prep[0], prep[1], prep[2])
    img = integ.renderD(sc, 0, seed=seed)
    drjit.eval(img)
    drjit.set_grad(P, 1.0)
    drjit.forward_to(img)
    g = drjit.grad(img)
    drjit.eval(g)
    drjit.sync_thread()
    return np.asarray(img.numpy(), dtype=np.float32), np.asarray(g.numpy(), dtype=np.float32)


def section(name, fn):
    t0 = time.time()
    try:
        LOG[name] = {"ok": True, "info": fn()}
    except Exception as e:  # noqa
        LOG[name] = {"ok": False, "err": repr(e), "tb": traceback.format_exc()}
        print("SECTION FAILED", name, repr(e), flush=True)
    LOG[name]["secs"] = time.time() - t0
    json.dump(LOG, open(os.path.join(OUT, "log.json"), "w"), indent=1, default=str)
    print("section", name, LOG[name].get("ok"), "%.1fs" % LOG[name]["secs"], flush=True)


def sphere_meshes():
    return scenes.cbox_meshes() + [scenes.icosphere(2, 80.0, (185.0, 250.0, 169.0))]


def mf_renderC():
    sc = build(scenes.cbox_meshes(), scenes.CBOX_MF_BSDFS, 128, 128, 4, 0, 0)
    sc.configure()
    sc.configure([0])
    out = {}
    for depth, seed in ((1, 0), (3, 3)):
        img = psdr.PathTracer(depth).renderC(sc, 0, seed=seed)
        out["img_d%d_seed%d" % (depth, seed)] = np.asarray(img.numpy(), dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, "mf_renderC.npz"), **out)
    return {k: float(v.mean()) for k, v in out.items()}


def mf_renderD():
    out, info = {}, {}
    for name, spps in (("interior", (4, 0, 0)), ("primary", (0, 4, 0)), ("secondary", (0, 0, 4)), ("all", (4, 4, 4))):
        sc = build(scenes.cbox_meshes(), scenes.CBOX_MF_BSDFS, 128, 128, *spps)
        img, g = render_d(sc, psdr.PathTracer(2), 5, 1, (0.0, 30.0, 50.0))
        out["img_" + name], out["grad_" + name] = img, g
        info[name] = [float(img.mean()), float(np.abs(g).mean()), int(np.isnan(g).sum())]
    np.savez_compressed(os.path.join(OUT, "mf_renderD_128_s4_d2_smallbox.npz"), **out)
    return info


def guided():
    out, info = {}, {}
    for name, prep in (("unguided", None), ("guided_r1", ([200, 4, 4, 8], 1, 0)), ("guided_r2", ([64, 8, 8, 4], 2, 7))):
        sc = build(sphere_meshes(), scenes.CBOX_BSDFS, 128, 128, 0, 0, 8)
        img, g = render_d(sc, psdr.PathTracer(2), 1, 8, (40.0, 20.0, 0.0), prep)
        out["grad_" + name] = g
        info[name] = [float(np.abs(g).mean()), float(g.sum()), int(np.isnan(g).sum())]
    np.savez_compressed(os.path.join(OUT, "guided_sec_128_s8_sphere.npz"), **out)
    return info


section("mf_renderC", mf_renderC)
section("mf_renderD", mf_renderD)
section("guided", guided)
print(json.dumps({k: (v.get("ok"), v.get("err"), v.get("info")) for k, v in LOG.items()}, indent=1))
