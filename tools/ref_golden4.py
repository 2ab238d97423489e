#!/usr/bin/env python3
"""Fourth set of golden vectors from the RUNNING reference: textured DiffuseBSDF (Bitmap3fD with more than one
texel) on the meshes that carry UVs.  Output: gpurun_out/ref_golden4/tex_render.npz"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", "ref_golden4")
os.makedirs(OUT, exist_ok=True)
ns = {"__file__": os.path.join(ROOT, "tools", "ref_golden2.py"), "__name__": "golden4"}
src = open(os.path.join(ROOT, "tools", "ref_golden2.py")).read().split('section("mf_renderC"')[0].replace('"ref_golden2"', '"ref_golden4"')
exec(compile(src, "ref_golden2_head", "exec"), ns)
psdr, scenes, drjit, T, mat = ns["psdr"], ns["scenes"], ns["drjit"], ns["T"], ns["mat"]
from drjit.cuda import Matrix4f as Matrix4fC  # noqa: E402
from drjit.cuda.ad import Array3f as Vector3fD, Float as FloatD, Matrix4f as Matrix4fD  # noqa: E402


def textures(seed=4):
    rng = np.random.default_rng(seed)
    out = {}
    for name, (w, h) in (("white", (8, 6)), ("cat", (5, 7))):
        out[name] = (rng.random((h * w, 3), dtype=np.float32) * 0.8 + 0.1, w, h)
        rng.normal(size=(h * w, 3))          # keeps the stream aligned with tests/test_gpu_parity.py::_textures
    return out


def build(w, h, spp, sppe, sppse):
    cam = scenes.CBOX_CAMERA
    sc = psdr.Scene()
    o = sc.opts
    o.width, o.height, o.spp, o.sppe, o.sppse, o.log_level = w, h, spp, sppe, sppse, 0
    sensor = psdr.PerspectiveCamera(cam["fov"], cam["near"], cam["far"])
    sensor.to_world = Matrix4fD(mat(cam["to_world"]))
    sc.add_Sensor(sensor)
    tex = textures()
    for name, p in scenes.CBOX_BSDFS:
        if name in tex:
            d, tw, th = tex[name]
            sc.add_BSDF(psdr.DiffuseBSDF(psdr.Bitmap3fD(tw, th, Vector3fD(d[:, 0], d[:, 1], d[:, 2]))), name)
        else:
            sc.add_BSDF(psdr.DiffuseBSDF([float(x) for x in p]), name)
    for i, m in enumerate(scenes.cbox_meshes()):
        path = os.path.join(ns["OBJDIR"], "m%d_%s.obj" % (i, m.name))
        scenes.write_obj(m, path)
        em = psdr.AreaLight([float(x) for x in m.emitter]) if m.emitter is not None else None
        sc.add_Mesh(path, Matrix4fC(mat(m.to_world)), m.bsdf, em)
    return sc


out = {}
sc = build(128, 128, 4, 0, 0)
sc.configure(); sc.configure([0])
out["img_d3_seed3"] = np.asarray(psdr.PathTracer(3).renderC(sc, 0, seed=3).numpy(), np.float32)
# derivative with respect to a camera translation (moves the texture coordinate of the primary hit)
sc = build(128, 128, 4, 0, 0)
P = FloatD(0.)
drjit.enable_grad(P)
sc.param_map["Sensor[0]"].set_transform(Matrix4fD(T(P * 3., P * -2., P * 1.)))
sc.configure(); sc.configure([0])
img = psdr.PathTracer(2).renderD(sc, 0, seed=6)
drjit.eval(img)
drjit.set_grad(P, 1.0)
drjit.forward_to(img)
g = drjit.grad(img)
drjit.eval(g)
out["img_cam"], out["grad_cam"] = np.asarray(img.numpy(), np.float32), np.asarray(g.numpy(), np.float32)
np.savez_compressed(os.path.join(OUT, "tex_render.npz"), **out)
print({k: float(np.abs(v).mean()) for k, v in out.items()})
