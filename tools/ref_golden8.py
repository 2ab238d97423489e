#!/usr/bin/env python3
"""Eighth set of golden vectors from the RUNNING reference: RoughConductorBSDF (src/bsdf/roughconductor.cpp) on the two
Cornell-box blocks ("cat"), the parameters of tutorials/batch_render.ipynb (gold: eta, k) with alpha 0.15 (and 0.01 as in
the notebook), 128 x 128, spp 4, PathTracer(3): renderC, renderD's primal image, and forward-mode derivative images
w.r.t. (a) the luminaire translation (interior term), (b) alpha, (c) eta.x, (d) k.y.
Output: gpurun_out/ref_golden8/conductor.npz"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", "ref_golden8")
os.makedirs(OUT, exist_ok=True)
ns = {"__file__": os.path.join(ROOT, "tools", "ref_golden2.py"), "__name__": "golden8"}
src = open(os.path.join(ROOT, "tools", "ref_golden2.py")).read().split('section("mf_renderC"')[0].replace('"ref_golden2"', '"ref_golden8"')
exec(compile(src, "ref_golden2_head", "exec"), ns)
psdr, scenes, drjit, T, mat = ns["psdr"], ns["scenes"], ns["drjit"], ns["T"], ns["mat"]
from drjit.cuda import Matrix4f as Matrix4fC  # noqa: E402
from drjit.cuda.ad import Float as FloatD, Matrix4f as Matrix4fD, Array3f as Vector3fD  # noqa: E402

ETA, K = [0.155475, 0.116753, 0.138334], [4.83181, 3.12296, 2.1486]


def build(alpha, w=128, h=128, spp=4, P=None, wrt=None):
    cam = scenes.CBOX_CAMERA
    sc = psdr.Scene()
    o = sc.opts
    o.width, o.height, o.spp, o.sppe, o.sppse, o.log_level = w, h, spp, 0, 0, 0
    sensor = psdr.PerspectiveCamera(cam["fov"], cam["near"], cam["far"])
    sensor.to_world = Matrix4fD(mat(cam["to_world"]))
    sc.add_Sensor(sensor)
    for name, p in scenes.CBOX_BSDFS:
        if name == "cat":
            a_bm, e_bm, k_bm = psdr.Bitmap1fD(alpha), psdr.Bitmap3fD(ETA), psdr.Bitmap3fD(K)
            if wrt == "alpha":
                a_bm = psdr.Bitmap1fD(1, 1, FloatD(alpha) + P)
            elif wrt == "eta":
                e_bm = psdr.Bitmap3fD(1, 1, Vector3fD(FloatD(ETA[0]) + P, FloatD(ETA[1]), FloatD(ETA[2])))
            elif wrt == "k":
                k_bm = psdr.Bitmap3fD(1, 1, Vector3fD(FloatD(K[0]), FloatD(K[1]) + P, FloatD(K[2])))
            sc.add_BSDF(psdr.RoughConductorBSDF(a_bm, e_bm, k_bm), name)
        else:
            sc.add_BSDF(psdr.DiffuseBSDF([float(x) for x in p]), name)
    for i, m in enumerate(scenes.cbox_meshes()):
        path = os.path.join(ns["OBJDIR"], "m%d_%s.obj" % (i, m.name))
        scenes.write_obj(m, path)
        em = psdr.AreaLight([float(x) for x in m.emitter]) if m.emitter is not None else None
        sc.add_Mesh(path, Matrix4fC(mat(m.to_world)), m.bsdf, em)
    return sc


def fwd(sc, P, integ, seed=0):
    img = integ.renderD(sc, 0, seed=seed)
    drjit.eval(img)
    drjit.set_grad(P, 1.0)
    drjit.forward_to(img)
    g = drjit.grad(img)
    drjit.eval(g)
    return np.asarray(img.numpy(), np.float32), np.asarray(g.numpy(), np.float32)


out = {"eta": np.float32(ETA), "k": np.float32(K)}
integ = psdr.PathTracer(3)
for tag, alpha in (("a15", 0.15), ("a01", 0.01)):
    sc = build(alpha)
    sc.configure(); sc.configure([0])
    out["imgC_" + tag] = np.asarray(integ.renderC(sc, 0, seed=0).numpy(), np.float32)
    print(tag, "renderC mean", float(out["imgC_" + tag].mean()), flush=True)
# (a) geometry
P = FloatD(0.); drjit.enable_grad(P)
sc = build(0.15)
sc.param_map["Mesh[0]"].set_transform(Matrix4fD(T(P * 100., P * 0., P * 0.)))
sc.configure(); sc.configure([0])
out["imgD_geo"], out["gradD_geo"] = fwd(sc, P, integ)
# (b) alpha, (c) eta.x, (d) k.y  -- P enters the bitmap data of the cat BSDF
for tag in ("alpha", "eta", "k"):
    P = FloatD(0.); drjit.enable_grad(P)
    sc = build(0.15, P=P, wrt=tag)
    sc.configure(); sc.configure([0])
    out["imgD_" + tag], out["gradD_" + tag] = fwd(sc, P, integ)
    print(tag, float(np.abs(out["gradD_" + tag]).mean()), flush=True)
np.savez_compressed(os.path.join(OUT, "conductor.npz"), **out)
print("saved", {k: v.shape for k, v in out.items()})
