#!/usr/bin/env python3
"""Seventh set of golden vectors from the RUNNING reference: secondary-edge GUIDING pinned on a scene where the
reference's epsilon-band tests cannot flip -- the Cornell box shrunk 100x (psdr_jit_b200/scenes.py scaled_cbox): the
pre-pass PathTracer.preprocess_secondary_edges keeps a sample only if two fp32 reconstructions of one point agree to
ShadowEpsilon = 1e-3; at full scale their typical distance is 3e-4 and a few flipped samples reshuffle the whole CDF,
at 1/100 scale it is ~1e-6.  The reference does not expose the mass vector, so the guided secondary-edge derivative
image is the pin: it only matches per pixel if the cell masses (hence the warped samples) match.
Output: gpurun_out/ref_golden7/guided_scaled.npz"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", "ref_golden7")
os.makedirs(OUT, exist_ok=True)
ns = {"__file__": os.path.join(ROOT, "tools", "ref_golden2.py"), "__name__": "golden7"}
src = open(os.path.join(ROOT, "tools", "ref_golden2.py")).read().split('section("mf_renderC"')[0].replace('"ref_golden2"', '"ref_golden7"')
exec(compile(src, "ref_golden2_head", "exec"), ns)
psdr, scenes, drjit, T, mat = ns["psdr"], ns["scenes"], ns["drjit"], ns["T"], ns["mat"]
from drjit.cuda import Matrix4f as Matrix4fC  # noqa: E402
from drjit.cuda.ad import Float as FloatD, Matrix4f as Matrix4fD  # noqa: E402

SCALE = 0.01
RES, SPPSE = 128, 16


def build():
    meshes, cam = scenes.scaled_cbox(SCALE)
    sc = psdr.Scene()
    o = sc.opts
    o.width, o.height, o.spp, o.sppe, o.sppse, o.log_level = RES, RES, 0, 0, SPPSE, 0
    sensor = psdr.PerspectiveCamera(cam["fov"], cam["near"], cam["far"])
    sensor.to_world = Matrix4fD(mat(cam["to_world"]))
    sc.add_Sensor(sensor)
    for name, p in scenes.CBOX_BSDFS:
        sc.add_BSDF(psdr.DiffuseBSDF([float(x) for x in p]), name)
    for i, m in enumerate(meshes):
        path = os.path.join(ns["OBJDIR"], "s%d_%s.obj" % (i, m.name))
        scenes.write_obj(m, path)
        em = psdr.AreaLight([float(x) for x in m.emitter]) if m.emitter is not None else None
        sc.add_Mesh(path, Matrix4fC(mat(m.to_world)), m.bsdf, em)
    return sc


def grad(sc, integ, seed):
    P = FloatD(0.)
    drjit.enable_grad(P)
    sc.param_map["Mesh[0]"].set_transform(Matrix4fD(T(P * 1., P * 0., P * 0.)))     # 100 * SCALE
    sc.configure(); sc.configure([0])
    return P


out = {}
for tag, reso in (("unguided", None), ("guided_40_4_4_8", [40, 4, 4, 8]), ("guided_200_3_3_16", [200, 3, 3, 16])):
    sc = build()
    integ = psdr.PathTracer(1)
    P = grad(sc, integ, 0)
    if reso is not None:
        integ.preprocess_secondary_edges(sc, 0, reso, 1)
    img = integ.renderD(sc, 0, seed=3)
    drjit.eval(img)
    drjit.set_grad(P, 1.0)
    drjit.forward_to(img)
    g = drjit.grad(img)
    drjit.eval(g)
    out["grad_" + tag] = np.asarray(g.numpy(), np.float32)
    print(tag, float(np.abs(out["grad_" + tag]).sum()), flush=True)
np.savez_compressed(os.path.join(OUT, "guided_scaled.npz"), **out)
