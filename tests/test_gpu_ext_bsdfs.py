"""The rest of the reference's BSDF set on the CUDA path: RoughDielectricBSDF
(src/bsdf/roughdielectric.cpp), MicrofacetBSDFPerVertex (src/bsdf/microfacet_pv.cpp) and NormalMapBSDF
(src/bsdf/normalmap.cpp) -- against the oracle (values and forward-mode tangents, all three terms), against the reference's
own output (tests/golden/ext_bsdfs.npz, tools/ref_golden10.py), through the scene-file loader; reverse mode of the
three against forward mode."""
import copy
import os

import numpy as np
import pytest

from tests.common import GOLDEN, build_oracle, build_product, compare_stats, rel_l2, scenes

pytestmark = pytest.mark.gpu
TOL = 1e-4


def box_meshes(bsdf="ext"):
    """Cornell box with the tall box (the mesh with UVs and 8 shared vertices) on a BSDF of its own"""
    ms = copy.deepcopy(scenes.cbox_meshes())
    for m in ms:
        if m.name == "largebox":
            m.bsdf = bsdf
    return ms


def with_ext(spec):
    return list(scenes.CBOX_BSDFS) + [("ext", spec)]


def pervertex_tables(n=8, seed=5):
    rng = np.random.default_rng(seed)
    return (rng.uniform(0.02, 0.9, (n, 3)).astype(np.float32), rng.uniform(0.05, 0.8, (n, 3)).astype(np.float32),
            rng.uniform(0.15, 0.9, n).astype(np.float32))


def normal_texture(w=8, h=6, seed=9):
    rng = np.random.default_rng(seed)
    n = rng.normal(size=(h * w, 3)).astype(np.float32) * 0.25 + np.float32([0, 0, 1])
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    return dict(data=(0.5 * n + 0.5).astype(np.float32), w=w, h=h)


SPECS = {
    "dielectric": {"dielectric": (0.2, 1.5, 1.0)},
    "dielectric_thin": {"dielectric": (0.05, 1.33, 1.0)},
    "pervertex": {"pervertex": pervertex_tables()},
    "normalmap_mf": {"normalmap": {"normal": (0.42, 0.56, 0.93), "nested": ([0.3, 0.6, 0.8], [0.4, 0.3, 0.2], 0.45)}},
    "normalmap_diffuse": {"normalmap": {"normal": (0.55, 0.47, 0.9), "nested": (0.7, 0.6, 0.5)}},
    "normalmap_dielectric": {"normalmap": {"normal": (0.45, 0.52, 0.95), "nested": {"dielectric": (0.3, 1.5, 1.0)}}},
}


@pytest.mark.parametrize("kind", sorted(SPECS))
def test_renderC_vs_oracle(kind):
    import psdr_jit_b200 as psdr
    bs = with_ext(SPECS[kind])
    ref = build_oracle(box_meshes(), 96, 96, 4, 0, 0, bsdfs=bs).render(4, seed=3, mode=0)
    got = psdr.PathTracer(4).renderC(build_product(box_meshes(), 96, 96, 4, 0, 0, bsdfs=bs), 0, seed=3).cpu().numpy()
    assert np.isfinite(got).all() and np.abs(ref).max() > 0
    assert rel_l2(got, ref) < TOL


def test_normalmap_texture_vs_oracle():
    import psdr_jit_b200 as psdr
    bs = with_ext(SPECS["normalmap_mf"])
    tex = {"ext": {0: normal_texture()}}
    ref = build_oracle(box_meshes(), 96, 96, 4, 0, 0, bsdfs=bs, textures=tex).render(3, seed=2, mode=0)
    got = psdr.PathTracer(3).renderC(build_product(box_meshes(), 96, 96, 4, 0, 0, bsdfs=bs, textures=tex), 0, seed=2).cpu().numpy()
    assert rel_l2(got, ref) < TOL


def material_tangent(kind):
    rng = np.random.default_rng(11)
    if kind.startswith("dielectric"):
        return dict(d_bsdf={"ext": np.float32([0.7])})
    if kind == "pervertex":
        return dict(d_bsdf={"ext": rng.normal(size=(8, 7)).astype(np.float32) * 0.3})
    return dict(d_bsdf={"ext": np.float32([0.3, -0.2, 0.1])})


@pytest.mark.parametrize("kind", ["dielectric", "pervertex", "normalmap_mf", "normalmap_dielectric"])
def test_renderD_all_terms_and_tangents_vs_oracle(kind):
    """image + forward derivative image: a moving luminaire AND a tangent on the BSDF's own parameters (alpha; the per-vertex
    tables; the normal map and the nested Microfacet), interior + primary-edge + secondary-edge terms"""
    import psdr_jit_b200 as psdr
    spec = copy.deepcopy(SPECS[kind])
    if kind == "normalmap_mf":
        spec["normalmap"]["d_nested"] = np.float32([0.2, -0.1, 0.3, 0.1, 0.2, -0.2, 0.4])
    bs = with_ext(spec)
    kw = dict(move_mesh=0, axis_scale=(40.0, 10.0, 0.0), bsdfs=bs, **material_tangent(kind))
    img_ref, dimg_ref = build_oracle(box_meshes(), 96, 96, 4, 4, 4, **kw).render(3, seed=6, mode=1, terms=7)
    sc = build_product(box_meshes(), 96, 96, 4, 4, 4, **kw)
    img, dimg = psdr.PathTracer(3).renderD_fwd(sc, 0, seed=6)
    assert rel_l2(img.cpu().numpy(), img_ref) < TOL
    assert np.abs(dimg_ref).max() > 0 and rel_l2(dimg.cpu().numpy(), dimg_ref) < TOL


@pytest.mark.parametrize("kind", ["pervertex", "normalmap_mf"])
def test_material_tangent_matches_finite_differences(kind):
    """the forward-mode image of a material parameter against central differences of the product's own image.  Only
    parameters that do not enter the (detached) sampling densities can be checked this way per seed -- the per-vertex
    diffuse colour, the diffuse colour of the BSDF under the normal map; for roughness / alpha / the normal itself the
    detached estimator and the difference quotient agree in expectation only (the oracle and the reference goldens pin
    those)."""
    import psdr_jit_b200 as psdr
    h = 1e-2

    def render(eps, tangent):
        spec = copy.deepcopy(SPECS[kind])
        if kind == "pervertex":
            s, df, r = spec["pervertex"]
            df = df.copy()
            df[:, 1] += np.float32(eps)
            spec["pervertex"] = (s, df, r)
            d = np.zeros((8, 7), np.float32)
            d[:, 4] = 1.0
            kw = dict(d_bsdf={"ext": d}) if tangent else {}
        else:
            sp, df, r = spec["normalmap"]["nested"]
            spec["normalmap"]["nested"] = (sp, [df[0], df[1] + eps, df[2]], r)
            if tangent:
                spec["normalmap"]["d_nested"] = np.float32([0, 0, 0, 0, 1, 0, 0])
            kw = {}
        sc = build_product(box_meshes(), 64, 64, 16, 0, 0, bsdfs=with_ext(spec), **kw)
        out = psdr.PathTracer(3).renderD_fwd(sc, 0, seed=4, terms=1)
        return out[1 if tangent else 0].cpu().numpy()

    fd = (render(h, False) - render(-h, False)) / (2 * h)
    ad = render(0.0, True)
    assert np.abs(ad).max() > 0
    assert rel_l2(ad, fd) < 2e-3, rel_l2(ad, fd)


def test_dielectric_transmits_and_conserves_energy():
    """a dielectric slab in front of the camera: light reaches the pixels behind it (transmission paths exist) and a
    furnace-like bound holds (no pixel brighter than the emitter)"""
    import psdr_jit_b200 as psdr
    sc = build_product(box_meshes(), 64, 64, 32, 0, 0, bsdfs=with_ext(SPECS["dielectric"]))
    img = psdr.PathTracer(5).renderC(sc, 0, seed=1).cpu().numpy()
    diffuse = build_product(box_meshes(), 64, 64, 32, 0, 0, bsdfs=with_ext((0.5, 0.5, 0.5)))
    base = psdr.PathTracer(5).renderC(diffuse, 0, seed=1).cpu().numpy()
    assert np.isfinite(img).all() and img.max() <= 20.0 + 1e-3
    assert rel_l2(img, base) > 1e-2          # the box looks different


def test_vs_reference_golden():
    """renderC and forward derivative images of the RUNNING reference (tools/ref_golden10.py)"""
    import psdr_jit_b200 as psdr
    path = os.path.join(GOLDEN, "ext_bsdfs.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/ext_bsdfs.npz not generated yet (tools/ref_golden10.py needs the reference on a GPU box)")
    g = np.load(path)
    integ = psdr.PathTracer(3)
    integ.reference_tangent_scaling = True
    spp = int(g["spp"])
    pv = (g["pv_spec"], g["pv_diff"], g["pv_rough"])
    cases = {"pervertex": {"pervertex": pv}, "normalmap": {"normalmap": {"normal": tuple(g["nm_normal"]), "nested": (list(g["nm_spec"]), list(g["nm_diff"]), float(g["nm_rough"]))}},
             "normalmap_default": {"normalmap": {"normal": (0.499999, 0.499999, 1.0), "nested": ([0.04] * 3, [0.5] * 3, 0.8)}}}
    for tag, spec in list(cases.items()) + [("dielectric", SPECS["dielectric"])]:
        if "imgC_" + tag not in g:
            continue
        sc = build_product(box_meshes(), 128, 128, spp, 0, 0, bsdfs=with_ext(spec))
        got = integ.renderC(sc, 0, seed=0).cpu().numpy()
        r, nbad, r_ex = compare_stats(got, g["imgC_" + tag])
        assert nbad <= 32 and r_ex < 1e-3, (tag, r, nbad, r_ex)
    cases["dielectric"] = SPECS["dielectric"]
    for tag in ("pervertex", "normalmap", "dielectric"):
        if "gradD_" + tag not in g:
            continue
        kw = dict(move_mesh=0, axis_scale=(100.0, 0.0, 0.0))
        sc = build_product(box_meshes(), 128, 128, spp, 0, 0, bsdfs=with_ext(cases[tag]), **kw)
        img, dimg = integ.renderD_fwd(sc, 0, seed=0)
        r, nbad, r_ex = compare_stats(img.cpu().numpy(), g["imgD_" + tag])
        assert nbad <= 128 and r_ex < 2e-3, (tag, r, nbad, r_ex)
        r, nbad, r_ex = compare_stats(dimg.cpu().numpy(), g["gradD_" + tag])
        assert nbad <= 256 and r_ex < 5e-3, (tag, r, nbad, r_ex)
    # material derivatives: per-vertex roughness (all vertices), roughness of the Microfacet under the normal map, normal.x
    d_pv = np.zeros((8, 7), np.float32)
    d_pv[:, 6] = 1.0
    nm_rough = copy.deepcopy(cases["normalmap"])
    nm_rough["normalmap"]["d_nested"] = np.float32([0, 0, 0, 0, 0, 0, 1])
    for tag, spec, d in (("pervertex_rough", cases["pervertex"], d_pv), ("normalmap_rough", nm_rough, None),
                         ("normalmap_normal", cases["normalmap"], np.float32([1, 0, 0]))):
        if "gradD_" + tag not in g:
            continue
        sc = build_product(box_meshes(), 128, 128, spp, 0, 0, bsdfs=with_ext(spec), d_bsdf=None if d is None else {"ext": d})
        _, dimg = integ.renderD_fwd(sc, 0, seed=0)
        r, nbad, r_ex = compare_stats(dimg.cpu().numpy(), g["gradD_" + tag])
        assert nbad <= 256 and r_ex < 5e-3, (tag, r, nbad, r_ex)


def test_scene_file_with_dielectric_and_normalmap(tmp_path):
    """Scene.load_file on a scene with <bsdf type="roughdielectric"> and <bsdf type="normalmap"> (scene_loader.cpp:346-424)
    gives the image of the same scene built through the Python surface"""
    import psdr_jit_b200 as psdr
    ms = box_meshes("glass")
    for m in ms:
        if m.name == "smallbox":
            m.bsdf = "bumpy"
        scenes.write_obj(m, str(tmp_path / (m.name + ".obj")))
    cam = scenes.CBOX_CAMERA
    tw = " ".join("%.9g" % x for x in np.asarray(cam["to_world"], np.float32).ravel())
    xml = ['<scene version="0.6.0">',
           '<sensor type="perspective"><float name="fov" value="%g"/><float name="near_clip" value="%g"/><float name="far_clip" value="%g"/>' % (cam["fov"], cam["near"], cam["far"]),
           '<transform name="to_world"><matrix value="%s"/></transform>' % tw,
           '<sampler type="independent"><integer name="sample_count" value="4"/></sampler>',
           '<film type="hdrfilm"><integer name="width" value="64"/><integer name="height" value="64"/></film></sensor>']
    for name, p in scenes.CBOX_BSDFS:
        xml.append('<bsdf type="diffuse" id="%s"><rgb name="reflectance" value="%g, %g, %g"/></bsdf>' % ((name,) + tuple(p)))
    xml.append('<bsdf type="roughdielectric" id="glass"><float name="alpha" value="0.2"/><float name="intIOR" value="1.5"/><float name="extIOR" value="1.0"/></bsdf>')
    xml.append('<bsdf type="normalmap" id="bumpy"><rgb name="normalmap" value="0.42, 0.56, 0.93"/><bsdf type="microfacet">'
               '<rgb name="specular_reflectance" value="0.3, 0.6, 0.8"/><rgb name="diffuse_reflectance" value="0.4, 0.3, 0.2"/><float name="roughness" value="0.45"/></bsdf></bsdf>')
    for m in ms:
        em = '<emitter type="area"><rgb name="radiance" value="%g, %g, %g"/></emitter>' % tuple(m.emitter) if m.emitter is not None else ""
        mw = " ".join("%.9g" % x for x in np.asarray(m.to_world, np.float32).ravel())
        xml.append('<shape type="obj"><string name="filename" value="%s.obj"/><transform name="to_world"><matrix value="%s"/></transform><ref id="%s"/>%s</shape>' % (m.name, mw, m.bsdf, em))
    xml.append("</scene>")
    f = tmp_path / "scene.xml"
    f.write_text("\n".join(xml))
    sc = psdr.Scene()
    sc.opts.log_level = 0
    sc.load_file(str(f), False)
    sc.configure()
    got = psdr.PathTracer(3).renderC(sc, 0, seed=5).cpu().numpy()
    bs = list(scenes.CBOX_BSDFS) + [("glass", SPECS["dielectric"]), ("bumpy", SPECS["normalmap_mf"])]
    ref = psdr.PathTracer(3).renderC(build_product(ms, 64, 64, 4, 0, 0, bsdfs=bs), 0, seed=5).cpu().numpy()
    assert np.abs(ref).max() > 0 and rel_l2(got, ref) < 1e-6


def test_add_BSDF_normalmap_installs_the_reference_defaults():
    """Scene.add_BSDF(NormalMapBSDF(...)) ignores the object's fields (scene.cpp:219-229)"""
    import psdr_jit_b200 as psdr
    sc = psdr.Scene()
    sc.add_BSDF(psdr.NormalMapBSDF([0.1, 0.2, 0.3]), "nm")
    b = sc.param_map["BSDF[id=nm]"]
    assert np.allclose(b.normal_map, [.499999, .499999, 1.0]) and isinstance(b.nested_bsdf, psdr.MicrofacetBSDF)
    assert np.allclose(b.nested_bsdf.roughness, 0.8) and np.allclose(b.nested_bsdf.specularReflectance, 0.04)


def test_vjp_per_parameter_normalmap():
    """NormalMap in reverse mode (the adjoint differentiates the forward code as a function of world-space quantities,
    adjoint.cuh normalmap_jet): one JVP per parameter against the matching entry of a single VJP -- the constant normal map,
    the nested Microfacet's three parameters, and texels of a bitmap normal map"""
    import torch
    import psdr_jit_b200 as psdr
    rng = np.random.default_rng(4)
    w = h = 48
    integ = psdr.PathTracer(3)
    cot = torch.as_tensor(rng.normal(size=(w * h, 3)).astype(np.float32), device="cuda")

    def check(dimg, grad, what):
        lhs = float((cot.double() * dimg.double()).sum())
        ref = float(torch.linalg.norm(cot.double()) * torch.linalg.norm(dimg.double()))
        assert abs(lhs - float(grad)) < 1e-3 * max(abs(lhs), 1e-2 * ref), (what, lhs, float(grad))

    bs = with_ext(SPECS["normalmap_mf"])
    sc = build_product(box_meshes(), w, h, 16, 0, 0, bsdfs=bs)
    integ.render_vjp(sc, cot, 0, seed=2, terms=1)
    g_nm = sc.grad_of("BSDF[id=ext]", "normal_map").ravel().copy()
    g_nested = np.concatenate([sc.grad_of("BSDF[id=ext]", "nested_bsdf.specularReflectance").ravel(), sc.grad_of("BSDF[id=ext]", "nested_bsdf.diffuseReflectance").ravel(),
                               sc.grad_of("BSDF[id=ext]", "nested_bsdf.roughness").ravel()])
    assert g_nm.size == 3 and g_nested.size == 7 and np.abs(g_nm).max() > 0 and np.abs(g_nested).max() > 0
    for k in range(3):
        d = np.zeros(3, np.float32)
        d[k] = 1.0
        sc2 = build_product(box_meshes(), w, h, 16, 0, 0, bsdfs=bs, d_bsdf={"ext": d})
        check(integ.renderD_fwd(sc2, 0, seed=2, terms=1)[1], g_nm[k], ("normal", k))
    for k in (0, 4, 6):
        spec = copy.deepcopy(SPECS["normalmap_mf"])
        d = np.zeros(7, np.float32)
        d[k] = 1.0
        spec["normalmap"]["d_nested"] = d
        sc2 = build_product(box_meshes(), w, h, 16, 0, 0, bsdfs=with_ext(spec))
        check(integ.renderD_fwd(sc2, 0, seed=2, terms=1)[1], g_nested[k], ("nested", k))
    # bitmap normal map: texel gradients
    tex = normal_texture()
    sc = build_product(box_meshes(), w, h, 16, 0, 0, bsdfs=bs, textures={"ext": {0: tex}})
    integ.render_vjp(sc, cot, 0, seed=2, terms=1)
    g_tex = sc.grad_of("BSDF[id=ext]", "normal_map").reshape(-1, 3)
    assert g_tex.shape == (tex["w"] * tex["h"], 3) and np.abs(g_tex).max() > 0
    hot = np.argsort(-np.abs(g_tex).sum(axis=1))[:3]
    for t in hot:
        d = np.zeros_like(tex["data"])
        d[t, 1] = 1.0
        sc2 = build_product(box_meshes(), w, h, 16, 0, 0, bsdfs=bs, textures={"ext": {0: dict(tex, d_data=d)}})
        check(integ.renderD_fwd(sc2, 0, seed=2, terms=1)[1], g_tex[t, 1], ("texel", int(t)))


def test_vjp_per_parameter_dielectric_and_pervertex():
    """one JVP per parameter against the matching entry of a single VJP: alpha of the dielectric, entries of the per-vertex
    tables (roughness, a diffuse and a specular channel at several vertices)"""
    import torch
    import psdr_jit_b200 as psdr
    rng = np.random.default_rng(4)
    w = h = 48
    integ = psdr.PathTracer(3)
    cot = torch.as_tensor(rng.normal(size=(w * h, 3)).astype(np.float32), device="cuda")

    def check(lhs_dimg, grad):
        lhs = float((cot.double() * lhs_dimg.double()).sum())
        ref = float(torch.linalg.norm(cot.double()) * torch.linalg.norm(lhs_dimg.double()))
        assert abs(lhs - float(grad)) < 5e-4 * max(abs(lhs), 1e-2 * ref), (lhs, float(grad))

    bs = with_ext(SPECS["dielectric"])
    sc = build_product(box_meshes(), w, h, 16, 0, 0, bsdfs=bs)
    integ.render_vjp(sc, cot, 0, seed=2, terms=1)
    g_alpha = sc.grad_of("BSDF[id=ext]", "alpha_u").ravel()
    assert g_alpha.size == 1 and abs(g_alpha[0]) > 0
    sc2 = build_product(box_meshes(), w, h, 16, 0, 0, bsdfs=bs, d_bsdf={"ext": np.float32([1.0])})
    check(integ.renderD_fwd(sc2, 0, seed=2, terms=1)[1], g_alpha[0])

    bs = with_ext(SPECS["pervertex"])
    sc = build_product(box_meshes(), w, h, 16, 0, 0, bsdfs=bs)
    integ.render_vjp(sc, cot, 0, seed=2, terms=1)
    g = np.concatenate([sc.grad_of("BSDF[id=ext]", "specularReflectance"), sc.grad_of("BSDF[id=ext]", "diffuseReflectance"),
                        sc.grad_of("BSDF[id=ext]", "roughness").reshape(-1, 1)], axis=1)
    assert g.shape == (8, 7) and np.abs(g).max() > 0
    for vtx, col in ((0, 6), (3, 6), (5, 4), (2, 0), (7, 2)):
        d = np.zeros((8, 7), np.float32)
        d[vtx, col] = 1.0
        sc2 = build_product(box_meshes(), w, h, 16, 0, 0, bsdfs=bs, d_bsdf={"ext": d})
        check(integ.renderD_fwd(sc2, 0, seed=2, terms=1)[1], g[vtx, col])


@pytest.mark.parametrize("kind", ["dielectric", "pervertex", "normalmap_mf"])
def test_vjp_is_transpose_with_geometry_and_edges(kind):
    """<cotangent, J t> == <J^T cotangent, t> with t = (translation of the luminaire and of the tall box, material tangent),
    all three terms: the box moves under the camera, so the per-vertex tables are also reached through the differentiable
    barycentrics of the primary hit"""
    import torch
    import psdr_jit_b200 as psdr
    rng = np.random.default_rng(8)
    w = h = 64
    mt = material_tangent(kind)["d_bsdf"]["ext"]
    sc = build_product(box_meshes(), w, h, 8, 8, 8, bsdfs=with_ext(SPECS[kind]), d_bsdf={"ext": mt})
    t = np.zeros((4, 4), np.float32)
    t[:3, 3] = rng.normal(size=3) * 30
    tang = {}
    for name in ("Mesh[0]", "Mesh[2]"):
        sc.param_map[name].d_to_world_left = t.copy()
        tang[(name, "to_world_left")] = t.copy()
    if kind == "dielectric":
        tang[("BSDF[id=ext]", "alpha_u")] = np.float32(mt).reshape(1)
    elif kind == "normalmap_mf":
        tang[("BSDF[id=ext]", "normal_map")] = np.float32(mt).reshape(3)
    else:
        m = np.float32(mt).reshape(8, 7)
        tang[("BSDF[id=ext]", "specularReflectance")] = m[:, 0:3]
        tang[("BSDF[id=ext]", "diffuseReflectance")] = m[:, 3:6]
        tang[("BSDF[id=ext]", "roughness")] = m[:, 6]
    sc.configure()
    sc.configure([0])
    integ = psdr.PathTracer(3)
    img, dimg = integ.renderD_fwd(sc, 0, seed=3)
    cot = torch.as_tensor(rng.normal(size=(w * h, 3)).astype(np.float32), device=img.device)
    lhs = float((cot.double() * dimg.double()).sum())
    integ.render_vjp(sc, cot, 0, seed=3)
    parts = {k: float((sc.grad_of(*k).reshape(np.shape(v)).astype(np.float64) * v.astype(np.float64)).sum()) for k, v in tang.items()}
    rhs = sum(parts.values())
    mag = max(abs(lhs), sum(abs(v) for v in parts.values()))
    assert mag > 0 and abs(lhs - rhs) < 5e-4 * mag, (lhs, rhs, parts)


def test_constant_microfacet_scene_uses_the_small_family():
    """bitmaps / conductor / dielectric / per-vertex / normal-map code lives in its own kernel family: a constant
    Microfacet + envmap scene (BASELINE config 3) must not select it, an extended scene must"""
    import psdr_jit_b200 as psdr
    mf = [(n, ([0.2, 0.9, 0.9], [0.01, 0.01, 0.01], 0.3)) for n, _ in scenes.CBOX_BSDFS]
    assert build_product(scenes.cbox_meshes(), 32, 32, 1, 0, 0, bsdfs=mf).kernel_family() == 2
    assert build_product(box_meshes(), 32, 32, 1, 0, 0, bsdfs=with_ext(SPECS["pervertex"])).kernel_family() == 10
