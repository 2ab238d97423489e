#!/usr/bin/env python3
"""Tenth set of golden vectors from the RUNNING reference: the rest of its BSDF set on the Cornell box's tall block (the mesh
with UVs), 128 x 128, spp 4, PathTracer(3):
  * MicrofacetBSDFPerVertex (src/bsdf/microfacet_pv.cpp) -- renderC, renderD + forward derivative w.r.t. the luminaire
    translation;
  * NormalMapBSDF through Scene.add_normalmap_BSDF (src/scene/scene.cpp:128-145) -- the same three images -- and through
    Scene.add_BSDF (the reference then installs its defaults, scene.cpp:219-229) -- renderC;
  * RoughDielectricBSDF, which the reference creates from scene files only (src/scene/scene_loader.cpp:346-360) -- a scene
    file is written and loaded with Scene.load_file -- renderC.
Output: gpurun_out/ref_golden10/ext_bsdfs.npz"""
import copy
import os
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", "ref_golden10")
os.makedirs(OUT, exist_ok=True)
ns = {"__file__": os.path.join(ROOT, "tools", "ref_golden2.py"), "__name__": "golden10"}
src = open(os.path.join(ROOT, "tools", "ref_golden2.py")).read().split('section("mf_renderC"')[0].replace('"ref_golden2"', '"ref_golden10"')
exec(compile(src, "ref_golden2_head", "exec"), ns)
psdr, scenes, drjit, T, mat = ns["psdr"], ns["scenes"], ns["drjit"], ns["T"], ns["mat"]
from drjit.cuda import Matrix4f as Matrix4fC  # noqa: E402
from drjit.cuda.ad import Float as FloatD, Matrix4f as Matrix4fD, Array3f as Vector3fD, Array1f as Vector1fD  # noqa: E402

SPP = 4
rng = np.random.default_rng(5)
PV = (rng.uniform(0.02, 0.9, (8, 3)).astype(np.float32), rng.uniform(0.05, 0.8, (8, 3)).astype(np.float32),
      rng.uniform(0.15, 0.9, 8).astype(np.float32))
NM_NORMAL, NM_SPEC, NM_DIFF, NM_ROUGH = [0.42, 0.56, 0.93], [0.3, 0.6, 0.8], [0.4, 0.3, 0.2], 0.45


def box_meshes(bsdf="ext"):
    ms = copy.deepcopy(scenes.cbox_meshes())
    for m in ms:
        if m.name == "largebox":
            m.bsdf = bsdf
    return ms


def v3(a):
    a = np.asarray(a, np.float32)
    return Vector3fD(FloatD(a[:, 0].copy()), FloatD(a[:, 1].copy()), FloatD(a[:, 2].copy()))


def build(kind, w=128, h=128, spp=SPP, P=None, wrt=None):
    cam = scenes.CBOX_CAMERA
    sc = psdr.Scene()
    o = sc.opts
    o.width, o.height, o.spp, o.sppe, o.sppse, o.log_level = w, h, spp, 0, 0, 0
    sensor = psdr.PerspectiveCamera(cam["fov"], cam["near"], cam["far"])
    sensor.to_world = Matrix4fD(mat(cam["to_world"]))
    sc.add_Sensor(sensor)
    for name, p in scenes.CBOX_BSDFS:
        sc.add_BSDF(psdr.DiffuseBSDF([float(x) for x in p]), name)
    if kind == "pervertex":
        rough = FloatD(PV[2].copy())
        if wrt == "rough":
            rough = rough + P
        sc.add_BSDF(psdr.MicrofacetBSDFPerVertex(v3(PV[0]), v3(PV[1]), Vector1fD(rough)), "ext")
    elif kind == "normalmap":
        nm = psdr.NormalMapBSDF(NM_NORMAL)
        mf = psdr.MicrofacetBSDF(NM_SPEC, NM_DIFF, NM_ROUGH)
        if wrt == "rough":
            mf = psdr.MicrofacetBSDF(psdr.Bitmap3fD(NM_SPEC), psdr.Bitmap3fD(NM_DIFF), psdr.Bitmap1fD(1, 1, FloatD(NM_ROUGH) + P))
        elif wrt == "normal":
            nm.normal_map = psdr.Bitmap3fD(1, 1, Vector3fD(FloatD(NM_NORMAL[0]) + P, FloatD(NM_NORMAL[1]), FloatD(NM_NORMAL[2])))
        sc.add_normalmap_BSDF(nm, mf, "ext")
    elif kind == "normalmap_default":
        sc.add_BSDF(psdr.NormalMapBSDF(), "ext")
    for i, m in enumerate(box_meshes()):
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


def dielectric_scene_file():
    ms = box_meshes("glass")
    cam = scenes.CBOX_CAMERA
    tw = " ".join("%.9g" % x for x in np.asarray(cam["to_world"], np.float32).ravel())
    xml = ['<scene version="0.6.0">',
           '<sensor type="perspective"><float name="fov" value="%g"/><float name="near_clip" value="%g"/><float name="far_clip" value="%g"/>' % (cam["fov"], cam["near"], cam["far"]),
           '<transform name="to_world"><matrix value="%s"/></transform>' % tw,
           '<sampler type="independent"><integer name="sample_count" value="%d"/></sampler>' % SPP,
           '<film type="hdrfilm"><integer name="width" value="128"/><integer name="height" value="128"/></film></sensor>']
    for name, p in scenes.CBOX_BSDFS:
        xml.append('<bsdf type="diffuse" id="%s"><rgb name="reflectance" value="%g, %g, %g"/></bsdf>' % ((name,) + tuple(p)))
    xml.append('<bsdf type="roughdielectric" id="glass"><float name="alpha" value="0.2"/><float name="intIOR" value="1.5"/><float name="extIOR" value="1.0"/></bsdf>')
    for i, m in enumerate(ms):
        path = os.path.join(ns["OBJDIR"], "x%d_%s.obj" % (i, m.name))
        scenes.write_obj(m, path)
        em = '<emitter type="area"><rgb name="radiance" value="%g, %g, %g"/></emitter>' % tuple(m.emitter) if m.emitter is not None else ""
        mw = " ".join("%.9g" % x for x in np.asarray(m.to_world, np.float32).ravel())
        xml.append('<shape type="obj"><string name="filename" value="%s"/><transform name="to_world"><matrix value="%s"/></transform><ref id="%s"/>%s</shape>' % (path, mw, m.bsdf, em))
    xml.append("</scene>")
    f = os.path.join(OUT, "dielectric.xml")
    open(f, "w").write("\n".join(xml))
    return f


out = {"spp": np.int32(SPP), "pv_spec": PV[0], "pv_diff": PV[1], "pv_rough": PV[2], "nm_normal": np.float32(NM_NORMAL),
       "nm_spec": np.float32(NM_SPEC), "nm_diff": np.float32(NM_DIFF), "nm_rough": np.float32(NM_ROUGH)}
integ = psdr.PathTracer(3)


def save():
    np.savez_compressed(os.path.join(OUT, "ext_bsdfs.npz"), **out)


for kind in ("pervertex", "normalmap", "normalmap_default"):
    try:
        sc = build(kind)
        sc.configure(); sc.configure([0])
        out["imgC_" + kind] = np.asarray(integ.renderC(sc, 0, seed=0).numpy(), np.float32)
        print(kind, "renderC mean", float(out["imgC_" + kind].mean()), "finite", bool(np.isfinite(out["imgC_" + kind]).all()), flush=True)
        save()
        if kind == "normalmap_default":
            continue
        P = FloatD(0.); drjit.enable_grad(P)
        sc = build(kind)
        sc.param_map["Mesh[0]"].set_transform(Matrix4fD(T(P * 100., P * 0., P * 0.)))
        sc.configure(); sc.configure([0])
        out["imgD_" + kind], out["gradD_" + kind] = fwd(sc, P, integ)
        print(kind, "grad mean abs", float(np.abs(out["gradD_" + kind]).mean()), flush=True)
        save()
    except Exception:
        traceback.print_exc()
for kind, wrt in (("pervertex", "rough"), ("normalmap", "rough"), ("normalmap", "normal")):
    try:
        P = FloatD(0.); drjit.enable_grad(P)
        sc = build(kind, P=P, wrt=wrt)
        sc.configure(); sc.configure([0])
        _, out["gradD_%s_%s" % (kind, wrt)] = fwd(sc, P, integ)
        print(kind, wrt, "grad mean abs", float(np.abs(out["gradD_%s_%s" % (kind, wrt)]).mean()), flush=True)
        save()
    except Exception:
        traceback.print_exc()
try:
    f = dielectric_scene_file()
    sc = psdr.Scene()
    sc.load_file(f, False)
    sc.opts.log_level = 0
    sc.configure(); sc.configure([0])
    out["imgC_dielectric"] = np.asarray(integ.renderC(sc, 0, seed=0).numpy(), np.float32)
    print("dielectric renderC mean", float(out["imgC_dielectric"].mean()), flush=True)
    save()
    P = FloatD(0.); drjit.enable_grad(P)
    sc = psdr.Scene()
    sc.load_file(f, False)
    sc.opts.log_level = 0
    sc.param_map["Mesh[0]"].set_transform(Matrix4fD(T(P * 100., P * 0., P * 0.)))
    sc.configure(); sc.configure([0])
    out["imgD_dielectric"], out["gradD_dielectric"] = fwd(sc, P, integ)
    print("dielectric grad mean abs", float(np.abs(out["gradD_dielectric"]).mean()), flush=True)
    save()
except Exception:
    traceback.print_exc()
save()
print("saved", sorted(out))
