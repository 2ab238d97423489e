#!/usr/bin/env python3
"""Fifth set of golden vectors from the RUNNING reference: MicrofacetBSDF(Bitmap3fD, Bitmap3fD, Bitmap1fD) -- all three
bitmap slots textured -- with the bitmaps' uv transform (Bitmap.scale / .rotate / .translate, src/core/bitmap.cpp:64-72)
on the meshes that carry UVs.  The textures are tests/test_gpu_parity.py::_slot_textures.
Output: gpurun_out/ref_golden5/tex_slots.npz"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", "ref_golden5")
os.makedirs(OUT, exist_ok=True)
ns = {"__file__": os.path.join(ROOT, "tools", "ref_golden2.py"), "__name__": "golden5"}
src = open(os.path.join(ROOT, "tools", "ref_golden2.py")).read().split('section("mf_renderC"')[0].replace('"ref_golden2"', '"ref_golden5"')
exec(compile(src, "ref_golden2_head", "exec"), ns)
psdr, scenes, drjit, T, mat = ns["psdr"], ns["scenes"], ns["drjit"], ns["T"], ns["mat"]
from drjit.cuda import Matrix4f as Matrix4fC  # noqa: E402
from drjit.cuda.ad import Array2f as Vector2fD, Array3f as Vector3fD, Float as FloatD, Matrix4f as Matrix4fD  # noqa: E402


def slot_textures():
    """same stream as tests/test_gpu_parity.py::_slot_textures(with_tangent=False, with_xform=True)"""
    rng = np.random.default_rng(21)
    out = {}
    for name, dims in (("cat", ((7, 5), (4, 6), (5, 5))), ("white", ((6, 4), (3, 3), (8, 2)))):
        slots = {}
        for slot, (w, h) in enumerate(dims):
            ch = 1 if slot == 2 else 3
            lo, hi = ((0.05, 0.9), (0.02, 0.6), (0.15, 0.7))[slot]
            slots[slot] = dict(data=(rng.random((h * w, ch), dtype=np.float32) * (hi - lo) + lo), w=w, h=h,
                               xform=np.array([1.0 + 0.4 * slot, 0.3 - 0.25 * slot, 0.11 * (slot + 1), -0.07 * slot], np.float32))
        out[name] = slots
    return out


def bitmap(t, ch, P=None, dP=None):
    d = t["data"]
    if ch == 3:
        bm = psdr.Bitmap3fD(t["w"], t["h"], Vector3fD(d[:, 0], d[:, 1], d[:, 2]))
    else:
        bm = psdr.Bitmap1fD(t["w"], t["h"], FloatD(d[:, 0]))
    x = [float(v) for v in t["xform"]]
    if P is None:
        bm.scale, bm.rotate, bm.translate = FloatD(x[0]), FloatD(x[1]), Vector2fD(x[2], x[3])
    else:      # the uv transform moves with P
        bm.scale, bm.rotate = FloatD(x[0]) + P * dP[0], FloatD(x[1]) + P * dP[1]
        bm.translate = Vector2fD(FloatD(x[2]) + P * dP[2], FloatD(x[3]) + P * dP[3])
    return bm


def build(w, h, spp, sppe, sppse, P=None):
    cam = scenes.CBOX_CAMERA
    sc = psdr.Scene()
    o = sc.opts
    o.width, o.height, o.spp, o.sppe, o.sppse, o.log_level = w, h, spp, sppe, sppse, 0
    sensor = psdr.PerspectiveCamera(cam["fov"], cam["near"], cam["far"])
    sensor.to_world = Matrix4fD(mat(cam["to_world"]))
    sc.add_Sensor(sensor)
    tex = slot_textures()
    for name, p in scenes.CBOX_MF_BSDFS:
        if name in tex:
            t = tex[name]
            dP = [None, None, None]
            if P is not None:
                dP = [[0.2 * (k + 1), -0.3 * (k + 1), 0.05 * (k + 1), 0.1 * (k + 1)] for k in range(3)]
            sc.add_BSDF(psdr.MicrofacetBSDF(bitmap(t[1], 3, P, dP[1]), bitmap(t[0], 3, P, dP[0]), bitmap(t[2], 1, P, dP[2])), name)
        elif len(p) == 3 and hasattr(p[0], "__len__"):
            sc.add_BSDF(psdr.MicrofacetBSDF([float(x) for x in p[0]], [float(x) for x in p[1]], float(p[2])), name)
        else:
            sc.add_BSDF(psdr.DiffuseBSDF([float(x) for x in p]), name)
    for i, m in enumerate(scenes.cbox_meshes()):
        path = os.path.join(ns["OBJDIR"], "m%d_%s.obj" % (i, m.name))
        scenes.write_obj(m, path)
        em = psdr.AreaLight([float(x) for x in m.emitter]) if m.emitter is not None else None
        sc.add_Mesh(path, Matrix4fC(mat(m.to_world)), m.bsdf, em)
    return sc


def fwd(sc, P, depth, seed):
    sc.configure(); sc.configure([0])
    img = psdr.PathTracer(depth).renderD(sc, 0, seed=seed)
    drjit.eval(img)
    drjit.set_grad(P, 1.0)
    drjit.forward_to(img)
    g = drjit.grad(img)
    drjit.eval(g)
    return np.asarray(img.numpy(), np.float32), np.asarray(g.numpy(), np.float32)


out = {}
sc = build(128, 128, 4, 0, 0)
sc.configure(); sc.configure([0])
out["img_d3_seed5"] = np.asarray(psdr.PathTracer(3).renderC(sc, 0, seed=5).numpy(), np.float32)
# derivative with respect to the uv transforms of all six bitmaps (interior term)
P = FloatD(0.)
drjit.enable_grad(P)
sc = build(128, 128, 4, 0, 0, P=P)
out["img_uv"], out["grad_uv"] = fwd(sc, P, 2, 8)
# derivative with respect to a translation of the large box (moves uv at the primary hit), all three terms
P = FloatD(0.)
drjit.enable_grad(P)
sc = build(128, 128, 4, 4, 4)
sc.param_map["Mesh[2]"].set_transform(Matrix4fD(T(P * 10., P * 0., P * 20.)))
out["img_box"], out["grad_box"] = fwd(sc, P, 2, 8)
np.savez_compressed(os.path.join(OUT, "tex_slots.npz"), **out)
print({k: float(np.abs(v).mean()) for k, v in out.items()})
