#!/usr/bin/env python3
"""Is the reference's forward-mode secondary-edge derivative image DETERMINISTIC?  The same scene, seed and parameters rendered
three times in one process (and once more in a fresh scene object): if the images differ between runs, the per-channel losses
seen against our output (tests/golden/scaled_cfg2.npz: same 437 non-zero pixels, individual channels of individual pixels lower
by up to 40 %) are a race in the reference's own accumulation (Dr.Jit scatter_reduce under forward-mode AD), not an estimator
difference.  Output: gpurun_out/ref_probe5/determinism.npz + the printed summary."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", "ref_probe5")
os.makedirs(OUT, exist_ok=True)
ns = {"__file__": os.path.join(ROOT, "tools", "ref_golden2.py"), "__name__": "probe5"}
src = open(os.path.join(ROOT, "tools", "ref_golden2.py")).read().split('section("mf_renderC"')[0].replace('"ref_golden2"', '"ref_probe5"')
exec(compile(src, "ref_golden2_head", "exec"), ns)
psdr, scenes, drjit, T, mat = ns["psdr"], ns["scenes"], ns["drjit"], ns["T"], ns["mat"]
from drjit.cuda import Matrix4f as Matrix4fC  # noqa: E402
from drjit.cuda.ad import Float as FloatD, Matrix4f as Matrix4fD  # noqa: E402


def build(meshes, cam, spp, sppe, sppse, w=128, h=128):
    sc = psdr.Scene()
    o = sc.opts
    o.width, o.height, o.spp, o.sppe, o.sppse, o.log_level = w, h, spp, sppe, sppse, 0
    sensor = psdr.PerspectiveCamera(cam["fov"], cam["near"], cam["far"])
    sensor.to_world = Matrix4fD(mat(cam["to_world"]))
    sc.add_Sensor(sensor)
    for name, p in scenes.CBOX_BSDFS:
        sc.add_BSDF(psdr.DiffuseBSDF([float(x) for x in p]), name)
    for i, m in enumerate(meshes):
        path = os.path.join(ns["OBJDIR"], "m%d_%s.obj" % (i, m.name))
        scenes.write_obj(m, path)
        em = psdr.AreaLight([float(x) for x in m.emitter]) if m.emitter is not None else None
        sc.add_Mesh(path, Matrix4fC(mat(m.to_world)), m.bsdf, em)
    return sc


def grad_once(meshes, cam, ax, spps, sc=None):
    P = FloatD(0.); drjit.enable_grad(P)
    if sc is None:
        sc = build(meshes, cam, *spps)
    sc.param_map["Mesh[0]"].set_transform(Matrix4fD(T(P * ax, P * 0., P * 0.)))
    sc.configure(); sc.configure([0])
    img = psdr.PathTracer(3).renderD(sc, 0, seed=0)
    drjit.eval(img)
    drjit.set_grad(P, 1.0)
    drjit.forward_to(img)
    g = drjit.grad(img)
    drjit.eval(g)
    return np.asarray(g.numpy(), np.float32), sc


out = {}
for tag, (meshes, cam), ax in (("scaled", scenes.scaled_cbox(1.0 / 300.0), 100.0 / 300.0), ("full", (scenes.cbox_meshes(), scenes.CBOX_CAMERA), 100.0)):
    for term, spps in (("sec", (0, 0, 4)), ("int", (4, 0, 0)), ("pri", (0, 4, 0))):
        runs = []
        g, sc = grad_once(meshes, cam, ax, spps)
        runs.append(g)
        for _ in range(2):
            runs.append(grad_once(meshes, cam, ax, spps, sc)[0])       # same scene object
        runs.append(grad_once(meshes, cam, ax, spps)[0])               # fresh scene object
        a = np.stack(runs)
        spread = (a.max(axis=0) - a.min(axis=0))
        n_diff = int((spread.max(axis=1) > 0).sum())
        print("%s %s: pixels that differ between 4 identical runs: %d of %d non-zero; max spread %.4g (max |value| %.4g); sums per run %s" % (
            tag, term, n_diff, int((np.abs(a[0]).max(axis=1) > 0).sum()), float(spread.max()), float(np.abs(a).max()),
            [round(float(np.abs(r).sum()), 4) for r in runs]), flush=True)
        out["%s_%s" % (tag, term)] = a
np.savez_compressed(os.path.join(OUT, "determinism.npz"), **out)
