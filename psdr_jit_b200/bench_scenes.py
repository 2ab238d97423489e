"""The five BASELINE.json workloads as plain data + builders for both bench arms (bench.py --config N).

Everything here is numpy; `build_ours` goes through this repo's psdr_jit-style surface, `build_reference` through the
UNMODIFIED reference's Python API (psdr_jit + Dr.Jit from baseline/_ref), so both arms consume the same triangles,
materials, cameras and environment map.  SURVEY.md 8(d) defines the workloads:

 1  cbox 128^2, spp 1, PathTracer(1), renderC
 2  cbox 512^2, spp = sppe = sppse = 32, PathTracer(3), renderD, DiffuseBSDF                      (the headline)
 3  cbox 1024^2, spp 128, PathTracer(6), renderD, one MicrofacetBSDF([.2,.9,.9],[.01,.01,.01],0.3) on every non-light
    mesh, area light kept + a 1024 x 512 lat-long environment map (rng(0).random**4 * 4)
 4  Bunny-in-Cornell: the five cbox walls + luminaire + tutorials/data/mesh/bunny_low.obj (4 968 faces, scaled 2.5x to
    (278, 102.5, 280)), 512^2, spp = sppse = 64 (sppe 0), secondary-edge guiding [2000, 5, 5, 32], PathTracer(3), renderD
 5  batch_render: cbox, 8 sensors on a ring in front of the box, 512^2, spp 32, PathTracer(3), renderD
The differentiated parameter is always the scalar P of the README: Mesh[k] translated by (100 P, 0, 0)."""
from __future__ import annotations

import os
from typing import Dict, List

import numpy as np

from . import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUNNY = os.path.join(ROOT, "baseline", "_ref", "data", "mesh", "bunny_low.obj")
BUNNY_TO_WORLD = np.array([[2.5, 0, 0, 278.0], [0, 2.5, 0, 102.5], [0, 0, 2.5, 280.0], [0, 0, 0, 1]], np.float32)
MF_CFG3 = ((0.2, 0.9, 0.9), (0.01, 0.01, 0.01), 0.3)


def ring_cameras(n: int = 8) -> List[dict]:
    cams = []
    for k in range(n):
        a = 2 * np.pi * k / n
        eye = np.array([278 + 900 * np.sin(a) * 0.35, 273 + 60 * np.cos(2 * a), -800 + 120 * (1 - np.cos(a))], np.float32)
        cams.append(dict(fov=60.0, near=1e-6, far=1e7, to_world=scenes.translate(*eye)))
    return cams


def workload(cfg: int, scale: float = 1.0) -> Dict:
    """scale < 1 shrinks spp (bounded samples for the CPU baseline and the reference arm)."""
    s = lambda n: max(1, int(round(n * scale)))      # noqa: E731
    base = dict(cfg=cfg, bsdfs=scenes.CBOX_BSDFS, meshes=scenes.cbox_meshes(), cams=[scenes.CBOX_CAMERA], envmap=None, moving=0,
                guiding=None, mode="renderD", sensors=[0])
    if cfg == 1:
        base.update(name="cbox 128x128 spp=1 PathTracer(1) renderC (BASELINE configs[0])", w=128, h=128, spp=1, sppe=0, sppse=0, depth=1, mode="renderC")
    elif cfg == 2:
        base.update(name="cbox 512x512 spp=32 sppe=32 sppse=32 PathTracer(3) renderD DiffuseBSDF (BASELINE configs[1])",
                    w=512, h=512, spp=s(32), sppe=s(32), sppse=s(32), depth=3)
    elif cfg == 3:
        rng = np.random.default_rng(0)
        env = ((rng.random((512 * 1024, 3), dtype=np.float32) ** 4) * 4).astype(np.float32)
        mf = [("light", (0.0, 0.0, 0.0))] + [(n_, MF_CFG3) for n_ in ("cat", "white", "green", "red")]
        base.update(name="cbox 1024x1024 spp=%d PathTracer(6) renderD MicrofacetBSDF + 1024x512 envmap (BASELINE configs[2])" % s(128),
                    w=1024, h=1024, spp=s(128), sppe=0, sppse=0, depth=6, bsdfs=mf, envmap=(env, 1024, 512))
    elif cfg == 4:
        ms = [m for m in scenes.cbox_meshes() if m.name not in ("smallbox", "largebox")]
        if os.path.exists(BUNNY):
            v, f, _, _ = scenes.load_obj(BUNNY)
            bunny = scenes.MeshData(name="bunny", v=v, f=f, to_world=BUNNY_TO_WORLD.copy(), bsdf="cat")
            what = "bunny_low.obj (%d faces)" % len(f)
        else:       # the data file did not travel: a stand-in of the same size (icosphere level 4, 5 120 faces)
            bunny = scenes.icosphere(4, 90.0, (278.0, 102.5, 280.0))
            what = "icosphere stand-in (5120 faces; baseline/_ref/data/mesh/bunny_low.obj missing)"
        ms.append(bunny)
        base.update(name="Bunny-in-Cornell (%s) 512x512 spp=%d sppse=%d guided [2000,5,5,32] PathTracer(3) renderD (BASELINE configs[3])" % (what, s(64), s(64)),
                    w=512, h=512, spp=s(64), sppe=0, sppse=s(64), depth=3, meshes=ms, moving=len(ms) - 1, guiding=[2000, 5, 5, 32])
    elif cfg == 5:
        base.update(name="batch_render: cbox, 8 sensors 512x512 spp=%d PathTracer(3) renderD, one sensor per GPU (BASELINE configs[4])" % s(32),
                    w=512, h=512, spp=s(32), sppe=0, sppse=0, depth=3, cams=ring_cameras(8), sensors=list(range(8)))
    else:
        raise ValueError("config must be 1..5")
    base["samples_per_call"] = base["w"] * base["h"] * (base["spp"] + base["sppe"] + base["sppse"])
    return base


def algorithmic_bytes(wl: Dict, term: int, lanes: int) -> int:
    """88 B per traced ray + 12 B per splat (24 B with a derivative image), one kernel launch."""
    d = wl["depth"]
    rays = {1: 1 + 2 * d, 2: 2 * (1 + 2 * d), 4: 3}[term]
    splat = (24 if wl["mode"] == "renderD" else 12) if term == 1 else 12
    return lanes * (88 * rays + splat)


def tangent_matrix() -> np.ndarray:
    t = np.zeros((4, 4), np.float32)
    t[0, 3] = 100.0
    return t


def build_ours(psdr, wl: Dict, rank: int = 0, world: int = 1):
    sc = psdr.Scene()
    o = sc.opts
    o.width, o.height, o.spp, o.sppe, o.sppse, o.log_level = wl["w"], wl["h"], wl["spp"], wl["sppe"], wl["sppse"], 0
    for cam in wl["cams"]:
        s = psdr.PerspectiveCamera(cam["fov"], cam["near"], cam["far"])
        s.to_world = cam["to_world"]
        sc.add_Sensor(s)
    for name, p in wl["bsdfs"]:
        if len(p) == 3 and hasattr(p[0], "__len__"):
            sc.add_BSDF(psdr.MicrofacetBSDF(p[0], p[1], p[2]), name)
        else:
            sc.add_BSDF(psdr.DiffuseBSDF(p), name)
    if wl["envmap"] is not None:
        env, ew, eh = wl["envmap"]
        sc.add_EnvironmentMap(psdr.EnvironmentMap(psdr.Bitmap3fD(ew, eh, env)))
    for m in wl["meshes"]:
        mesh = psdr.Mesh()
        mesh.load_raw(m.v, m.f, m.uv, m.fuv)
        mesh.to_world = m.to_world
        sc.add_Mesh(mesh, m.bsdf, psdr.AreaLight(m.emitter) if m.emitter is not None else None)
    if wl["mode"] == "renderD":
        sc.param_map["Mesh[%d]" % wl["moving"]].set_transform(np.eye(4, dtype=np.float32), tangent=tangent_matrix())
    sc.set_shard(rank, world)
    sc.set_accel(int(os.environ.get("PSDR_ACCEL", "-1")))
    sc.configure()
    sc.configure(wl["sensors"])
    return sc


def build_reference(ref, drjit, wl: Dict, objdir: str):
    """The same workload through the reference's API.  Returns (scene, set_param) where set_param() re-attaches the
    differentiable scalar P the way the README's loop does (returns P)."""
    from drjit.cuda import Matrix4f as Matrix4fC
    from drjit.cuda.ad import Array3f as Vector3fD, Float as FloatD, Matrix4f as Matrix4fD
    os.makedirs(objdir, exist_ok=True)
    mat = lambda m: [[float(m[i][j]) for j in range(4)] for i in range(4)]   # noqa: E731
    sc = ref.Scene()
    o = sc.opts
    o.width, o.height, o.spp, o.sppe, o.sppse, o.log_level = wl["w"], wl["h"], wl["spp"], wl["sppe"], wl["sppse"], 0
    for cam in wl["cams"]:
        s = ref.PerspectiveCamera(cam["fov"], cam["near"], cam["far"])
        s.to_world = Matrix4fD(mat(cam["to_world"]))
        sc.add_Sensor(s)
    for name, p in wl["bsdfs"]:
        if len(p) == 3 and hasattr(p[0], "__len__"):
            sc.add_BSDF(ref.MicrofacetBSDF([float(x) for x in p[0]], [float(x) for x in p[1]], float(p[2])), name)
        else:
            sc.add_BSDF(ref.DiffuseBSDF([float(x) for x in p]), name)
    if wl["envmap"] is not None:
        d, ew, eh = wl["envmap"]
        env = ref.EnvironmentMap()
        env.radiance = ref.Bitmap3fD(ew, eh, Vector3fD(d[:, 0], d[:, 1], d[:, 2]))
        env.scale = FloatD(1.0)
        sc.add_EnvironmentMap(env)
    for i, m in enumerate(wl["meshes"]):
        path = os.path.join(objdir, "cfg%d_m%d_%s.obj" % (wl["cfg"], i, m.name))
        scenes.write_obj(m, path)
        em = ref.AreaLight([float(x) for x in m.emitter]) if m.emitter is not None else None
        sc.add_Mesh(path, Matrix4fC(mat(m.to_world)), m.bsdf, em)

    def set_param():
        P = FloatD(0.)
        drjit.enable_grad(P)
        sc.param_map["Mesh[%d]" % wl["moving"]].set_transform(Matrix4fD([[1., 0., 0., P * 100.], [0., 1., 0., 0.], [0., 0., 1., 0.], [0., 0., 0., 1.]]))
        sc.configure()
        sc.configure(wl["sensors"])
        return P

    return sc, set_param
