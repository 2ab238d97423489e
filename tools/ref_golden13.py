#!/usr/bin/env python3
"""Thirteenth set of golden vectors from the RUNNING reference: OrthographicCamera(near, far) (src/sensor/orthographic.cpp) on
the Cornell box shrunk by 300 (the orthographic view volume is 2 x 2 camera units), 128 x 128, spp 4, PathTracer(2): renderC,
and renderD's forward derivative image w.r.t. the small box's translation, interior and primary-edge term one at a time.
Output: gpurun_out/ref_golden13/ortho.npz"""
import os
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", "ref_golden13")
os.makedirs(OUT, exist_ok=True)
ns = {"__file__": os.path.join(ROOT, "tools", "ref_golden2.py"), "__name__": "golden13"}
src = open(os.path.join(ROOT, "tools", "ref_golden2.py")).read().split('section("mf_renderC"')[0].replace('"ref_golden2"', '"ref_golden13"')
exec(compile(src, "ref_golden2_head", "exec"), ns)
psdr, scenes, drjit, T, mat = ns["psdr"], ns["scenes"], ns["drjit"], ns["T"], ns["mat"]
from drjit.cuda import Matrix4f as Matrix4fC  # noqa: E402
from drjit.cuda.ad import Float as FloatD, Matrix4f as Matrix4fD  # noqa: E402

S = 1.0 / 300.0
MESHES, CAM = scenes.scaled_cbox(S)


def build(spp, sppe, sppse, w=128, h=128):
    sc = psdr.Scene()
    o = sc.opts
    o.width, o.height, o.spp, o.sppe, o.sppse, o.log_level = w, h, spp, sppe, sppse, 0
    sensor = psdr.OrthographicCamera(1e-3, 1e3)
    sensor.to_world = Matrix4fD(mat(CAM["to_world"]))
    sc.add_Sensor(sensor)
    for name, p in scenes.CBOX_BSDFS:
        sc.add_BSDF(psdr.DiffuseBSDF([float(x) for x in p]), name)
    for i, m in enumerate(MESHES):
        path = os.path.join(ns["OBJDIR"], "m%d_%s.obj" % (i, m.name))
        scenes.write_obj(m, path)
        em = psdr.AreaLight([float(x) for x in m.emitter]) if m.emitter is not None else None
        sc.add_Mesh(path, Matrix4fC(mat(m.to_world)), m.bsdf, em)
    return sc


out = {"scale": np.float32(S)}
integ = psdr.PathTracer(2)
try:
    sc = build(4, 0, 0)
    sc.configure(); sc.configure([0])
    out["imgC"] = np.asarray(integ.renderC(sc, 0, seed=0).numpy(), np.float32)
    print("renderC mean", float(out["imgC"].mean()), flush=True)
    for tag, (spp, sppe, sppse) in (("int", (4, 0, 0)), ("pri", (0, 4, 0))):
        P = FloatD(0.); drjit.enable_grad(P)
        sc = build(spp, sppe, sppse)
        sc.param_map["Mesh[1]"].set_transform(Matrix4fD(T(P * 0.1, P * 0.03, P * 0.)))
        sc.configure(); sc.configure([0])
        img = integ.renderD(sc, 0, seed=0)
        drjit.eval(img)
        drjit.set_grad(P, 1.0)
        drjit.forward_to(img)
        g = drjit.grad(img)
        drjit.eval(g)
        out["gradD_" + tag] = np.asarray(g.numpy(), np.float32)
        print(tag, "grad mean abs", float(np.abs(out["gradD_" + tag]).mean()), flush=True)
        np.savez_compressed(os.path.join(OUT, "ortho.npz"), **out)
except Exception:
    traceback.print_exc()
np.savez_compressed(os.path.join(OUT, "ortho.npz"), **out)
print("saved", sorted(out))
