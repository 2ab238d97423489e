"""Shared helpers for the tests: scene construction for the oracle and metrics."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

from psdr_jit_b200 import scenes  # noqa: E402  (pure numpy module)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def sphere_meshes():
    return scenes.cbox_meshes() + [scenes.icosphere(2, 80.0, (185.0, 250.0, 169.0))]


def translation_tangent(axis_scale):
    d = np.zeros((4, 4), dtype=np.float32)
    d[0, 3], d[1, 3], d[2, 3] = axis_scale
    return d


def is_microfacet(params):
    return len(params) == 3 and hasattr(params[0], "__len__")


def is_conductor(params):
    """("name", {"conductor": (alpha, eta[3], k[3])}) or with a 4th entry specular_reflectance[3]"""
    return isinstance(params, dict) and "conductor" in params


def ext_kind(params):
    """extended material set: {"dielectric": (alpha, intIOR, extIOR)}, {"pervertex": (spec[n,3], diff[n,3], rough[n])},
    {"normalmap": {"normal": (x, y, z), "nested": <any non-normalmap spec>, "d_nested": tangent of the nested spec}}"""
    if isinstance(params, dict):
        for k in ("dielectric", "pervertex", "normalmap"):
            if k in params:
                return k
    return None


def test_envmap(w=32, h=16, seed=0):
    """synthetic lat-long environment map (BASELINE config 3 recipe: rng.random**4 * 4), [h*w, 3]"""
    rng = np.random.default_rng(seed)
    return (rng.random((h * w, 3), dtype=np.float32) ** 4 * 4).astype(np.float32)


test_envmap.__test__ = False


def tex_slots(t):
    """textures[name] -> {slot: dict(data, w, h, d_data, xform, d_xform)}.  The short form (data, w, h[, d_data]) is slot 0
    (reflectance / diffuseReflectance) without a uv transform; the long form is a dict keyed by slot (0, 1 specular,
    2 roughness) or slot name."""
    names = {"reflectance": 0, "diffuse": 0, "specular": 1, "roughness": 2}
    if isinstance(t, dict):
        return {names.get(k, k): v for k, v in t.items()}
    return {0: dict(data=t[0], w=t[1], h=t[2], d_data=t[3] if len(t) > 3 else None)}


def build_oracle(meshes, w, h, spp, sppe, sppse, move_mesh=None, axis_scale=None, active=(0,), cam=None, bsdfs=None, d_bsdf=None,
                 envmap=None, textures=None):
    """Scene = scenes.* meshes + CBOX bsdfs + camera; derivative parameter P translates mesh
    `move_mesh` by P*axis_scale through to_world_left (reference README.md:87-90)."""
    from oracle.psdr_oracle import OracleScene
    cam = cam or scenes.CBOX_CAMERA
    sc = OracleScene(w, h, spp, sppe, sppse)
    def add_oracle_bsdf(name, params, d):
        kind = ext_kind(params)
        if kind == "dielectric":        # d = (d_alpha,)
            c = params["dielectric"]
            sc.add_roughdielectric(name, c[0], c[1], c[2], d_alpha=0.0 if d is None else float(np.ravel(d)[0]))
        elif kind == "pervertex":       # d = [n, 7] table
            c = params["pervertex"]
            sc.add_microfacet_pervertex(name, c[0], c[1], c[2], d=d)
        elif kind == "normalmap":       # d = d_normal[3]; the nested BSDF gets a name of its own in the oracle
            c = params["normalmap"]
            add_oracle_bsdf(name + "/nested", c["nested"], c.get("d_nested"))
            sc.add_normalmap(name, name + "/nested", c.get("normal", (0.499999, 0.499999, 1.0)), d_normal=d)
        elif is_conductor(params):      # d = (d_alpha, d_eta[3], d_k[3], d_spec[3])
            c = params["conductor"]
            sc.add_roughconductor(name, c[0], c[1], c[2], c[3] if len(c) > 3 else (1.0, 1.0, 1.0), d=d)
        elif is_microfacet(params):
            sc.add_microfacet(name, params[0], params[1], params[2], d=d)
        else:
            sc.add_diffuse(name, params, d_refl=d)

    for name, params in (bsdfs or scenes.CBOX_BSDFS):
        add_oracle_bsdf(name, params, d_bsdf.get(name) if d_bsdf else None)
        if textures and name in textures:
            for slot, t in tex_slots(textures[name]).items():
                sc.set_bsdf_texture(name, t["data"], t["w"], t["h"], t.get("d_data"), slot=slot, xform=t.get("xform"), d_xform=t.get("d_xform"))
    if envmap is not None:      # dict(data, w, h, scale, to_world, d_data, d_scale, d_to_world_left); added before the meshes
        sc.add_envmap(envmap["data"], envmap["w"], envmap["h"], to_world=envmap.get("to_world"), scale=envmap.get("scale", 1.0),
                      d_data=envmap.get("d_data"), d_to_world_left=envmap.get("d_to_world_left"), d_scale=envmap.get("d_scale", 0.0))
    for i, m in enumerate(meshes):
        dtw = None
        if move_mesh is not None and i == move_mesh:
            dtw = {"left": translation_tangent(axis_scale)}
        sc.add_mesh(m.v, m.f, m.bsdf, uv=m.uv, fuv=m.fuv, to_world={"raw": m.to_world}, d_to_world=dtw, radiance=m.emitter)
    if cam.get("ortho"):                       # OrthographicCamera(near, far)
        sc.add_camera_orthographic(cam["near"], cam["far"], {"raw": cam["to_world"]})
    elif cam.get("intrinsics") is not None:    # PerspectiveCamera(fx, fy, cx, cy, near, far)
        sc.add_camera_intrinsic(*cam["intrinsics"], cam["near"], cam["far"], {"raw": cam["to_world"]})
    else:
        sc.add_camera(cam["fov"], cam["near"], cam["far"], {"raw": cam["to_world"]})
    sc.configure(active)
    return sc


def build_product(meshes, w, h, spp, sppe, sppse, move_mesh=None, axis_scale=None, active=(0,), cam=None, accel=-1,
                  shard=None, two_side=False, d_radiance=None, d_reflectance=None, d_cam_left=None, log_level=0, bsdfs=None, d_bsdf=None,
                  envmap=None, textures=None):
    """The same scene through the product's psdr_jit-style Python surface."""
    import psdr_jit_b200 as psdr
    cam = cam or scenes.CBOX_CAMERA
    sc = psdr.Scene()
    o = sc.opts
    o.width, o.height, o.spp, o.sppe, o.sppse, o.log_level = w, h, spp, sppe, sppse, log_level
    if cam.get("ortho"):
        sensor = psdr.OrthographicCamera(cam["near"], cam["far"])
    elif cam.get("intrinsics") is not None:
        sensor = psdr.PerspectiveCamera(*cam["intrinsics"], cam["near"], cam["far"])
    else:
        sensor = psdr.PerspectiveCamera(cam["fov"], cam["near"], cam["far"])
    sensor.to_world = cam["to_world"]
    sc.add_Sensor(sensor)
    def make_bsdf(params, d):
        kind = ext_kind(params)
        if kind == "dielectric":
            c = params["dielectric"]
            b = psdr.RoughDielectricBSDF(float(c[0]), float(c[1]), float(c[2]))
            if d is not None:
                b.d_alpha_u = np.float32(np.ravel(d)[0])
            return b
        if kind == "pervertex":
            c = params["pervertex"]
            b = psdr.MicrofacetBSDFPerVertex(c[0], c[1], c[2])
            if d is not None:
                t = np.float32(d).reshape(-1, 7)
                b.d_specularReflectance, b.d_diffuseReflectance, b.d_roughness = t[:, 0:3].copy(), t[:, 3:6].copy(), t[:, 6].copy()
            return b
        if is_conductor(params):
            c = params["conductor"]
            b = psdr.RoughConductorBSDF(psdr.Bitmap1fD(float(c[0])), psdr.Bitmap3fD(list(c[1])), psdr.Bitmap3fD(list(c[2])))
            if len(c) > 3:
                b.specular_reflectance = np.float32(c[3])
            if d is not None:
                b.d_alpha_u, b.d_eta, b.d_k, b.d_specular_reflectance = np.float32(d[0]), np.float32(d[1:4]), np.float32(d[4:7]), np.float32(d[7:10])
            return b
        if is_microfacet(params):
            b = psdr.MicrofacetBSDF(params[0], params[1], params[2])
            if d is not None:
                b.d_specularReflectance, b.d_diffuseReflectance, b.d_roughness = np.float32(d[0:3]), np.float32(d[3:6]), np.float32(d[6])
            return b
        b = psdr.DiffuseBSDF(params)
        if d is not None:
            b.d_reflectance = np.float32(d)
        return b

    for name, params in (bsdfs or scenes.CBOX_BSDFS):
        d = d_bsdf.get(name) if d_bsdf else None
        if ext_kind(params) == "normalmap":
            c = params["normalmap"]
            nm = psdr.NormalMapBSDF(c.get("normal", (0.499999, 0.499999, 1.0)))
            if d is not None:
                nm.d_normal_map = np.float32(d)
            if textures and name in textures:
                t = tex_slots(textures[name])[0]
                nm.normal_map = psdr.Bitmap3fD(t["w"], t["h"], t["data"])
                if t.get("d_data") is not None:
                    nm.normal_map.d_data = np.asarray(t["d_data"], dtype=np.float32)
            sc.add_normalmap_BSDF(nm, make_bsdf(c["nested"], c.get("d_nested")), name, twoSide=two_side)
            continue
        if ext_kind(params) in ("dielectric", "pervertex"):
            b = make_bsdf(params, d)
            if textures and name in textures and ext_kind(params) == "dielectric":
                t = tex_slots(textures[name])[2]
                b.alpha_u = psdr.Bitmap1fD(t["w"], t["h"], t["data"])
            sc.add_BSDF(b, name, twoSide=two_side)
            continue
        if is_conductor(params):
            c = params["conductor"]
            b = psdr.RoughConductorBSDF(psdr.Bitmap1fD(float(c[0])), psdr.Bitmap3fD(list(c[1])), psdr.Bitmap3fD(list(c[2])))
            if len(c) > 3:
                b.specular_reflectance = np.float32(c[3])
            if d is not None:
                b.d_alpha_u, b.d_eta, b.d_k, b.d_specular_reflectance = np.float32(d[0]), np.float32(d[1:4]), np.float32(d[4:7]), np.float32(d[7:10])
        elif is_microfacet(params):
            b = psdr.MicrofacetBSDF(params[0], params[1], params[2])
            if d is not None:
                b.d_specularReflectance, b.d_diffuseReflectance, b.d_roughness = np.float32(d[0:3]), np.float32(d[3:6]), np.float32(d[6])
        else:
            b = psdr.DiffuseBSDF(params)
            if d is not None:
                b.d_reflectance = np.float32(d)
        if textures and name in textures:
            for slot, t in tex_slots(textures[name]).items():
                bm = (psdr.Bitmap1fD if slot == 2 else psdr.Bitmap3fD)(t["w"], t["h"], t["data"])
                if t.get("d_data") is not None:
                    bm.d_data = np.asarray(t["d_data"], dtype=np.float32)
                if t.get("xform") is not None:
                    bm.scale, bm.rotate, bm.translate = np.float32(t["xform"][0]), np.float32(t["xform"][1]), np.float32(t["xform"][2:4])
                if t.get("d_xform") is not None:
                    bm.d_scale, bm.d_rotate, bm.d_translate = np.float32(t["d_xform"][0]), np.float32(t["d_xform"][1]), np.float32(t["d_xform"][2:4])
                if is_conductor(params):
                    setattr(b, (None, "specular_reflectance", "alpha_u")[slot], bm)
                elif not is_microfacet(params):
                    b.reflectance = bm
                else:
                    setattr(b, ("diffuseReflectance", "specularReflectance", "roughness")[slot], bm)
        sc.add_BSDF(b, name, twoSide=two_side)
    if envmap is not None:
        env = psdr.EnvironmentMap(psdr.Bitmap3fD(envmap["w"], envmap["h"], envmap["data"]))
        env.scale = np.float32(envmap.get("scale", 1.0))
        env.d_scale = np.float32(envmap.get("d_scale", 0.0))
        if envmap.get("to_world") is not None:
            env.to_world = np.asarray(envmap["to_world"], dtype=np.float32)
        if envmap.get("d_data") is not None:
            env.radiance.d_data = np.asarray(envmap["d_data"], dtype=np.float32)
        if envmap.get("d_to_world_left") is not None:
            env.set_transform(np.eye(4, dtype=np.float32), tangent=envmap["d_to_world_left"])
        sc.add_EnvironmentMap(env)
    for i, m in enumerate(meshes):
        mesh = psdr.Mesh()
        mesh.load_raw(m.v, m.f, m.uv, m.fuv)
        mesh.to_world = m.to_world
        sc.add_Mesh(mesh, m.bsdf, psdr.AreaLight(m.emitter) if m.emitter is not None else None)
    if move_mesh is not None:
        sc.param_map["Mesh[%d]" % move_mesh].set_transform(np.eye(4, dtype=np.float32), tangent=translation_tangent(axis_scale))
    if d_radiance is not None:
        sc.param_map["Emitter[%d]" % (1 if envmap is not None else 0)].d_radiance = np.asarray(d_radiance, dtype=np.float32)
    if d_reflectance is not None:
        name, d = d_reflectance
        sc.param_map["BSDF[id=%s]" % name].d_reflectance = np.asarray(d, dtype=np.float32)
    if d_cam_left is not None:
        sc.param_map["Sensor[0]"].set_transform(np.eye(4, dtype=np.float32), tangent=d_cam_left)
    sc.set_accel(accel)
    if shard is not None:
        sc.set_shard(*shard)
    sc.configure()
    sc.configure(list(active))
    return sc


def compare_stats(a, b, flip_rel=1e-3):
    """rel-L2, number of pixels whose max channel error exceeds flip_rel*max|b|, rel-L2 without them."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.abs(a - b).max(axis=1)
    bad = d > flip_rel * max(np.abs(b).max(), 1e-12)
    a2 = a.copy()
    a2[bad] = b[bad]
    return rel_l2(a, b), int(bad.sum()), rel_l2(a2, b)
