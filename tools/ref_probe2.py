#!/usr/bin/env python3
"""Second probe of the unmodified reference (baseline/_ref) on a GPU box: characterises how the
forward-mode derivative image scales per term and per parameter kind (the first probe showed the
interior and secondary-edge tangents coming out exactly 2x the finite-difference-correct value
for a mesh translation, the primary-edge one 1x).  Output: gpurun_out/ref_probe2/*.npz + log.json
"""
import json
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
OUT = os.path.join(ROOT, "gpurun_out", "ref_probe2")
os.makedirs(OUT, exist_ok=True)
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("scenes", os.path.join(ROOT, "psdr_jit_b200", "scenes.py"))
scenes = importlib.util.module_from_spec(spec)
sys.modules["scenes"] = scenes
spec.loader.exec_module(scenes)

import drjit  # noqa: E402
import psdr_jit as psdr  # noqa: E402
from drjit.cuda import Matrix4f as Matrix4fC  # noqa: E402
from drjit.cuda.ad import Array3f as Vector3fD, Float as FloatD, Matrix4f as Matrix4fD  # noqa: E402

OBJDIR = os.path.join(OUT, "obj")
os.makedirs(OBJDIR, exist_ok=True)
LOG = {}


def mat(m):
    return [[float(m[i][j]) for j in range(4)] for i in range(4)]


def T(x, y, z):
    return [[1., 0., 0., x], [0., 1., 0., y], [0., 0., 1., z], [0., 0., 0., 1.]]


def build(w, h, spp, sppe, sppse):
    cam = scenes.CBOX_CAMERA
    sc = psdr.Scene()
    o = sc.opts
    o.width, o.height, o.spp, o.sppe, o.sppse, o.log_level = w, h, spp, sppe, sppse, 0
    sensor = psdr.PerspectiveCamera(cam["fov"], cam["near"], cam["far"])
    sensor.to_world = Matrix4fD(mat(cam["to_world"]))
    sc.add_Sensor(sensor)
    for name, refl in scenes.CBOX_BSDFS:
        sc.add_BSDF(psdr.DiffuseBSDF([float(x) for x in refl]), name)
    for i, m in enumerate(scenes.cbox_meshes()):
        path = os.path.join(OBJDIR, "m%d_%s.obj" % (i, m.name))
        scenes.write_obj(m, path)
        em = psdr.AreaLight([float(x) for x in m.emitter]) if m.emitter is not None else None
        sc.add_Mesh(path, Matrix4fC(mat(m.to_world)), m.bsdf, em)
    return sc


def finish(sc, integ, P, seed, configure_twice=True):
    if configure_twice:
        sc.configure()
    sc.configure([0])
    img = integ.renderD(sc, 0, seed=seed)
    drjit.eval(img)
    drjit.set_grad(P, 1.0)
    drjit.forward_to(img)
    g = drjit.grad(img)
    drjit.eval(g)
    drjit.sync_thread()
    return np.asarray(img.numpy(), dtype=np.float32), np.asarray(g.numpy(), dtype=np.float32)


def variant(name, spps, depth, seed, setup, configure_twice=True):
    try:
        out = {}
        for tname, (a, b, c) in spps.items():
            sc = build(64, 64, a, b, c)
            P = FloatD(0.)
            drjit.enable_grad(P)
            setup(sc, P)
            img, g = finish(sc, psdr.PathTracer(depth), P, seed, configure_twice)
            out["img_" + tname] = img
            out["grad_" + tname] = g
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        LOG[name] = {k: float(np.abs(v).mean()) for k, v in out.items()}
    except Exception as e:  # noqa
        LOG[name] = {"err": repr(e), "tb": traceback.format_exc()}
    print(name, LOG[name], flush=True)
    with open(os.path.join(OUT, "log.json"), "w") as fh:
        json.dump(LOG, fh, indent=1)


TERMS = {"interior": (4, 0, 0), "primary": (0, 4, 0), "secondary": (0, 0, 4), "all": (4, 4, 4)}
INT_ONLY = {"interior": (4, 0, 0)}


def s_radiance(sc, P):
    sc.param_map["Emitter[0]"].radiance = Vector3fD(20. * (1. + P), 20. * (1. + P), 8. * (1. + P))


def s_reflect(sc, P):
    b = sc.param_map["BSDF[id=white]"]
    b.reflectance.data = Vector3fD(0.95 * (1. + P), 0.95 * (1. + P), 0.95 * (1. + P))


def s_left(sc, P):
    sc.param_map["Mesh[0]"].set_transform(Matrix4fD(T(P * 100., P * 0., P * 0.)))


def s_raw(sc, P):
    sc.param_map["Mesh[0]"].to_world = Matrix4fD(T(P * 100., P * 0. - 0.5, P * 0.))


def s_verts(sc, P):
    m = sc.param_map["Mesh[0]"]
    v = m.vertex_positions
    m.vertex_positions = Vector3fD(v[0] + P * 100., v[1], v[2])


def s_camera(sc, P):
    sc.param_map["Sensor[0]"].set_transform(Matrix4fD(T(P * 30., P * 10., P * 0.)))


def s_box(sc, P):
    sc.param_map["Mesh[1]"].set_transform(Matrix4fD(T(P * 0., P * 30., P * 50.)))


variant("radiance", INT_ONLY, 2, 0, s_radiance)
variant("reflectance", INT_ONLY, 2, 0, s_reflect)
variant("left_twice", TERMS, 2, 0, s_left, True)
variant("left_once", TERMS, 2, 0, s_left, False)
variant("raw", INT_ONLY, 2, 0, s_raw)
variant("verts", INT_ONLY, 2, 0, s_verts)
variant("camera", TERMS, 2, 0, s_camera)
variant("box_once", TERMS, 2, 0, s_box, False)
print(json.dumps(LOG, indent=1)[:3000])
