#!/usr/bin/env python3
"""Ninth set of golden vectors from the RUNNING reference: FieldExtractionIntegrator(field).renderC / renderD with
forward-mode derivative images (src/integrator/field.cpp) for depth, position, shNormal, geoNormal, silhouette, uv on
(a) the Cornell box and (b) its two blocks + luminaire without the walls (a silhouette against the void), 128 x 128,
spp 4, sppe 4; the small box translates by (30 P, 10 P, 0).
Output: gpurun_out/ref_golden9/fields.npz"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", "ref_golden9")
os.makedirs(OUT, exist_ok=True)
ns = {"__file__": os.path.join(ROOT, "tools", "ref_golden2.py"), "__name__": "golden9"}
src = open(os.path.join(ROOT, "tools", "ref_golden2.py")).read().split('section("mf_renderC"')[0].replace('"ref_golden2"', '"ref_golden9"')
exec(compile(src, "ref_golden2_head", "exec"), ns)
psdr, scenes, drjit, T, mat = ns["psdr"], ns["scenes"], ns["drjit"], ns["T"], ns["mat"]
from drjit.cuda import Matrix4f as Matrix4fC  # noqa: E402
from drjit.cuda.ad import Float as FloatD, Matrix4f as Matrix4fD  # noqa: E402


def build(walls, spp, sppe, P=None):
    cam = scenes.CBOX_CAMERA
    sc = psdr.Scene()
    o = sc.opts
    o.width, o.height, o.spp, o.sppe, o.sppse, o.log_level = 128, 128, spp, sppe, 0, 0
    sensor = psdr.PerspectiveCamera(cam["fov"], cam["near"], cam["far"])
    sensor.to_world = Matrix4fD(mat(cam["to_world"]))
    sc.add_Sensor(sensor)
    for name, p in scenes.CBOX_BSDFS:
        sc.add_BSDF(psdr.DiffuseBSDF([float(x) for x in p]), name)
    ms = scenes.cbox_meshes() if walls else scenes.cbox_meshes()[:3]
    for i, m in enumerate(ms):
        path = os.path.join(ns["OBJDIR"], "m%d_%s.obj" % (i, m.name))
        scenes.write_obj(m, path)
        em = psdr.AreaLight([float(x) for x in m.emitter]) if m.emitter is not None else None
        sc.add_Mesh(path, Matrix4fC(mat(m.to_world)), m.bsdf, em)
    if P is not None:
        sc.param_map["Mesh[1]"].set_transform(Matrix4fD(T(P * 30., P * 10., P * 0.)))
    sc.configure(); sc.configure([0])
    return sc


out = {}
for tag, walls in (("box", True), ("open", False)):
    for field in ("depth", "position", "shNormal", "geoNormal", "silhouette", "uv"):
        integ = psdr.FieldExtractionIntegrator(field)
        sc = build(walls, 4, 0)
        out["%s_%s_C" % (tag, field)] = np.asarray(integ.renderC(sc, 0, seed=0).numpy(), np.float32)
        for terms, (spp, sppe) in (("int", (4, 0)), ("all", (4, 4))):
            P = FloatD(0.); drjit.enable_grad(P)
            sc = build(walls, spp, sppe, P)
            img = integ.renderD(sc, 0, seed=0)
            drjit.eval(img)
            drjit.set_grad(P, 1.0)
            out["%s_%s_D_%s" % (tag, field, terms)] = np.asarray(img.numpy(), np.float32)
            try:
                drjit.forward_to(img)
                g = drjit.grad(img)
                drjit.eval(g)
                out["%s_%s_G_%s" % (tag, field, terms)] = np.asarray(g.numpy(), np.float32)
            except TypeError:       # the image does not depend on P (e.g. the silhouette's interior part): zero
                out["%s_%s_G_%s" % (tag, field, terms)] = np.zeros_like(out["%s_%s_D_%s" % (tag, field, terms)])
        print(tag, field, float(np.abs(out["%s_%s_C" % (tag, field)]).mean()), float(np.abs(out["%s_%s_G_all" % (tag, field)]).mean()), flush=True)
np.savez_compressed(os.path.join(OUT, "fields.npz"), **out)
print("saved", len(out))
