#!/usr/bin/env python3
"""Third set of golden vectors from the RUNNING reference (GPU box): EnvironmentMap scenes (in-memory lat-long
Bitmap3fD, Microfacet materials, area light kept -> two emitters).  Output: gpurun_out/ref_golden3/*.npz."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", "ref_golden3")
os.makedirs(OUT, exist_ok=True)
ns = {"__file__": os.path.join(ROOT, "tools", "ref_golden2.py"), "__name__": "golden3"}
src = open(os.path.join(ROOT, "tools", "ref_golden2.py")).read().split('section("mf_renderC"')[0].replace('"ref_golden2"', '"ref_golden3"')
exec(compile(src, "ref_golden2_head", "exec"), ns)
build, psdr, scenes, drjit, section, T = ns["build"], ns["psdr"], ns["scenes"], ns["drjit"], ns["section"], ns["T"]
from drjit.cuda.ad import Array3f as Vector3fD, Float as FloatD, Matrix4f as Matrix4fD  # noqa: E402


def envmap_data(w=32, h=16, seed=0):
    rng = np.random.default_rng(seed)
    return (rng.random((h * w, 3), dtype=np.float32) ** 4 * 4).astype(np.float32)


def add_env(sc, rot=0.7, scale=1.5, w=32, h=16):
    d = envmap_data(w, h)
    env = psdr.EnvironmentMap()
    env.radiance = psdr.Bitmap3fD(w, h, Vector3fD(d[:, 0], d[:, 1], d[:, 2]))
    env.scale = FloatD(scale)
    c, s = float(np.cos(rot)), float(np.sin(rot))
    env.set_transform(Matrix4fD([[c, 0., s, 0.], [0., 1., 0., 0.], [-s, 0., c, 0.], [0., 0., 0., 1.]]))
    sc.add_EnvironmentMap(env)


def build_env(bsdfs, w, h, spp, sppe, sppse):
    # same order as tests/common.py: sensor, BSDFs, envmap, meshes -> Emitter[0] = envmap, Emitter[1] = area light
    cam = scenes.CBOX_CAMERA
    sc = psdr.Scene()
    o = sc.opts
    o.width, o.height, o.spp, o.sppe, o.sppse, o.log_level = w, h, spp, sppe, sppse, 0
    sensor = psdr.PerspectiveCamera(cam["fov"], cam["near"], cam["far"])
    sensor.to_world = Matrix4fD(ns["mat"](cam["to_world"]))
    sc.add_Sensor(sensor)
    for name, p in bsdfs:
        if ns["is_mf"](p):
            sc.add_BSDF(psdr.MicrofacetBSDF([float(x) for x in p[0]], [float(x) for x in p[1]], float(p[2])), name)
        else:
            sc.add_BSDF(psdr.DiffuseBSDF([float(x) for x in p]), name)
    add_env(sc)
    from drjit.cuda import Matrix4f as Matrix4fC
    for i, m in enumerate(scenes.cbox_meshes()):
        path = os.path.join(ns["OBJDIR"], "m%d_%s.obj" % (i, m.name))
        scenes.write_obj(m, path)
        em = psdr.AreaLight([float(x) for x in m.emitter]) if m.emitter is not None else None
        sc.add_Mesh(path, Matrix4fC(ns["mat"](m.to_world)), m.bsdf, em)
    return sc


def env_renderC():
    out = {}
    for tag, bsdfs in (("diffuse", scenes.CBOX_BSDFS), ("mf", scenes.CBOX_MF_BSDFS)):
        sc = build_env(bsdfs, 128, 128, 4, 0, 0)
        sc.configure()
        sc.configure([0])
        for depth, seed in ((1, 0), (3, 3)):
            img = psdr.PathTracer(depth).renderC(sc, 0, seed=seed)
            out["img_%s_d%d_seed%d" % (tag, depth, seed)] = np.asarray(img.numpy(), dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, "env_renderC.npz"), **out)
    return {k: float(v.mean()) for k, v in out.items()}


def env_renderD():
    out, info = {}, {}
    for name, spps in (("interior", (4, 0, 0)), ("all", (4, 4, 4))):
        sc = build_env(scenes.CBOX_MF_BSDFS, 128, 128, *spps)
        P = FloatD(0.)
        drjit.enable_grad(P)
        sc.param_map["Mesh[1]"].set_transform(Matrix4fD(T(P * 0., P * 30., P * 50.)))
        sc.configure()
        sc.configure([0])
        integ = psdr.PathTracer(3)
        img = integ.renderD(sc, 0, seed=5)
        drjit.eval(img)
        drjit.set_grad(P, 1.0)
        drjit.forward_to(img)
        g = drjit.grad(img)
        drjit.eval(g)
        drjit.sync_thread()
        out["img_" + name], out["grad_" + name] = np.asarray(img.numpy(), np.float32), np.asarray(g.numpy(), np.float32)
        info[name] = [float(out["img_" + name].mean()), float(np.abs(out["grad_" + name]).mean())]
    # derivative with respect to the envmap scale (interior only): the image is affine in it
    sc = build_env(scenes.CBOX_MF_BSDFS, 128, 128, 4, 0, 0)
    S = FloatD(1.5)
    drjit.enable_grad(S)
    sc.param_map["Emitter[0]"].scale = S
    sc.configure()
    sc.configure([0])
    img = psdr.PathTracer(3).renderD(sc, 0, seed=5)
    drjit.eval(img)
    drjit.set_grad(S, 1.0)
    drjit.forward_to(img)
    g = drjit.grad(img)
    drjit.eval(g)
    out["grad_scale"] = np.asarray(g.numpy(), np.float32)
    np.savez_compressed(os.path.join(OUT, "env_renderD_128_s4_d3_smallbox.npz"), **out)
    return info


section("env_renderC", env_renderC)
section("env_renderD", env_renderD)
import json  # noqa: E402
print(json.dumps({k: (v.get("ok"), v.get("err"), v.get("info")) for k, v in ns["LOG"].items()}, indent=1))
