#!/usr/bin/env python3
"""Sixth set of golden vectors from the RUNNING reference: the Direct integrator (psdr.Direct(mis), src/integrator/direct.cpp)
in its three MIS modes on the Cornell box, 128 x 128, spp 4: renderC and the interior derivative image.
Output: gpurun_out/ref_golden6/direct.npz"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", "ref_golden6")
os.makedirs(OUT, exist_ok=True)
ns = {"__file__": os.path.join(ROOT, "tools", "ref_golden2.py"), "__name__": "golden6"}
src = open(os.path.join(ROOT, "tools", "ref_golden2.py")).read().split('section("mf_renderC"')[0].replace('"ref_golden2"', '"ref_golden6"')
exec(compile(src, "ref_golden2_head", "exec"), ns)
psdr, scenes, drjit, T, mat = ns["psdr"], ns["scenes"], ns["drjit"], ns["T"], ns["mat"]
from drjit.cuda import Matrix4f as Matrix4fC  # noqa: E402
from drjit.cuda.ad import Float as FloatD, Matrix4f as Matrix4fD  # noqa: E402


def build(w, h, spp, sppe, sppse):
    cam = scenes.CBOX_CAMERA
    sc = psdr.Scene()
    o = sc.opts
    o.width, o.height, o.spp, o.sppe, o.sppse, o.log_level = w, h, spp, sppe, sppse, 0
    sensor = psdr.PerspectiveCamera(cam["fov"], cam["near"], cam["far"])
    sensor.to_world = Matrix4fD(mat(cam["to_world"]))
    sc.add_Sensor(sensor)
    for name, p in scenes.CBOX_BSDFS:
        sc.add_BSDF(psdr.DiffuseBSDF([float(x) for x in p]), name)
    for i, m in enumerate(scenes.cbox_meshes()):
        path = os.path.join(ns["OBJDIR"], "m%d_%s.obj" % (i, m.name))
        scenes.write_obj(m, path)
        em = psdr.AreaLight([float(x) for x in m.emitter]) if m.emitter is not None else None
        sc.add_Mesh(path, Matrix4fC(mat(m.to_world)), m.bsdf, em)
    return sc


out = {}
for mis in (0, 1, 2):
    sc = build(128, 128, 4, 0, 0)
    sc.configure(); sc.configure([0])
    integ = psdr.Direct(mis)
    out["imgC_mis%d" % mis] = np.asarray(integ.renderC(sc, 0, seed=0).numpy(), np.float32)
    P = FloatD(0.)
    drjit.enable_grad(P)
    sc = build(128, 128, 4, 0, 0)
    sc.param_map["Mesh[0]"].set_transform(Matrix4fD(T(P * 100., P * 0., P * 0.)))
    sc.configure(); sc.configure([0])
    img = integ.renderD(sc, 0, seed=0)
    drjit.eval(img)
    drjit.set_grad(P, 1.0)
    drjit.forward_to(img)
    g = drjit.grad(img)
    drjit.eval(g)
    out["imgD_int_mis%d" % mis], out["gradD_int_mis%d" % mis] = np.asarray(img.numpy(), np.float32), np.asarray(g.numpy(), np.float32)
    print(mis, {k: float(np.abs(v).mean()) for k, v in out.items() if k.endswith("mis%d" % mis)}, flush=True)
np.savez_compressed(os.path.join(OUT, "direct.npz"), **out)
