"""CPU tier (-m "not gpu"): pins the oracle (oracle/psdr_oracle.cpp) against golden vectors produced
by RUNNING the unmodified reference on a B200 (tools/ref_golden.py -> tests/golden/*.npz), and checks
the host logic + that the C-ABI library loads and exports every symbol include/psdr_b200.h declares.
No GPU compute here."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests.common import GOLDEN, ROOT, build_oracle, compare_stats, rel_l2, scenes, sphere_meshes


def test_sampler_bit_exact_vs_reference(oracle):
    """TEA-64 seeding + PCG32 (reference src/core/sampler.cpp:6-42) -- bit exact, incl. lanes near 2^23."""
    g = np.load(GOLDEN + "/sampler.npz")
    assert np.array_equal(oracle.sampler_draws(0, 8, 8), g["draws_seed0"])
    assert np.array_equal(oracle.sampler_draws(7, 8, 8), g["draws_seed7"])
    assert np.array_equal(oracle.sampler_draws(8388600, 8, 4), g["draws_big"])
    # next_2d: y receives the first draw (GCC right-to-left argument evaluation, sampler.h:19-21)
    assert np.array_equal(g["next2d_seed0"][1], g["draws_seed0"][0])
    assert np.array_equal(g["next2d_seed0"][0], g["draws_seed0"][1])


def test_pmf_sample_reuse_vs_reference(oracle):
    """DiscreteDistribution::init/sample (pmf.cpp:6-51): index bit-exact, sum bit-exact."""
    g = np.load(GOLDEN + "/pmf.npz")
    idx, p, s = oracle.pmf_sample(g["pmf"], g["samples"])
    assert np.array_equal(idx, g["idx"])
    assert np.float32(s) == g["sum"][0]
    np.testing.assert_allclose(p, g["p"], rtol=1e-6, atol=0)


def test_edge_tables_vs_reference(oracle):
    """Mesh.edge_indices() (mesh.cpp:244-305): std::map order, (v0, v1, f0, f1) bit-exact; 66 + 480 edges."""
    g = np.load(GOLDEN + "/edges.npz")
    osc = build_oracle(sphere_meshes(), 32, 32, 1, 1, 1)
    for i in range(9):
        assert np.array_equal(osc.mesh_edges(i), g["mesh%d" % i])
    assert osc.num_secondary_edges() == 6 * 5 + 2 * 18 + 480


def test_aov_vs_reference(oracle):
    g = np.load(GOLDEN + "/aov.npz")
    for name, meshes in (("cbox", scenes.cbox_meshes()), ("cboxsphere", sphere_meshes())):
        a = build_oracle(meshes, 128, 128, 1, 0, 0).aov()
        # FieldExtractionIntegrator returns exactly 2x the field in the reference binary
        assert np.array_equal(a[:, 0] * 2, g[name + "_segmentation"][:, 0])       # mesh ids bit-exact
        assert np.abs(a[:, 2:5] - g[name + "_position"] / 2).max() < 2e-3
        assert np.abs(a[:, 5] - g[name + "_depth"][:, 0] / 2).max() < 2e-3
        assert np.abs(a[:, 6:9] - g[name + "_geoNormal"] / 2).max() < 1e-5
        assert np.abs(a[:, 9:12] - g[name + "_shNormal"] / 2).max() < 1e-4


def test_renderC_vs_reference(oracle):
    g = np.load(GOLDEN + "/cfg1_renderC.npz")
    img = build_oracle(scenes.cbox_meshes(), 128, 128, 1, 0, 0).render(1, seed=0, mode=0)
    r, nbad, r_ex = compare_stats(img, g["img"])
    assert nbad <= 4 and r_ex < 1e-4, (r, nbad, r_ex)
    img = build_oracle(scenes.cbox_meshes(), 128, 128, 4, 0, 0).render(3, seed=3, mode=0)
    r, nbad, r_ex = compare_stats(img, g["img_d3_spp4_seed3"])
    assert nbad <= 40 and r_ex < 1e-3, (r, nbad, r_ex)


CASES = [
    ("renderD_128_s4_d3_light", "cbox", 3, 0, 0, (100.0, 0.0, 0.0)),
    ("renderD_128_s4_d2_smallbox", "cbox", 2, 5, 1, (0.0, 30.0, 50.0)),
    ("renderD_128_s4_d2_sphere", "sphere", 2, 1, 8, (40.0, 20.0, 0.0)),
]


@pytest.mark.parametrize("tag,scene,depth,seed,mesh,axis", CASES)
def test_renderD_terms_vs_reference(oracle, tag, scene, depth, seed, mesh, axis):
    """Each term of renderD + drjit.forward_to against the running reference.  The reference binary's
    interior and secondary-edge tangents are exactly 2x the finite-difference value."""
    g = np.load(GOLDEN + "/%s.npz" % tag)
    meshes = scenes.cbox_meshes() if scene == "cbox" else sphere_meshes()
    for term, spps, scale in (("interior", (4, 0, 0), 2.0), ("primary", (0, 4, 0), 1.0), ("secondary", (0, 0, 4), 2.0)):
        osc = build_oracle(meshes, 128, 128, *spps, move_mesh=mesh, axis_scale=axis)
        img, dimg = osc.render(depth, seed=seed, mode=1, terms=7)
        if spps[0]:
            r, nbad, r_ex = compare_stats(img, g["img_" + term])
            assert nbad <= 120 and r_ex < 1e-3, (term, r, nbad, r_ex)
        else:
            assert np.abs(img).max() == 0.0                     # boundary terms are zero-primal
        r, nbad, r_ex = compare_stats(dimg * scale, g["grad_" + term])
        assert nbad <= 0.02 * len(dimg) and r_ex < 1e-3, (term, r, nbad, r_ex)


def test_reference_tangent_is_twice_the_derivative():
    """Known answer from the reference itself: d img / d(radiance scale) must equal img; the reference
    returns 2*img (tests/golden/probe2_radiance.npz, tools/ref_probe2.py)."""
    g = np.load(GOLDEN + "/probe2_radiance.npz")
    assert rel_l2(g["grad_interior"], 2.0 * g["img_interior"]) < 1e-5


def test_oracle_finite_difference(oracle):
    """The oracle's forward tangent is the derivative: central differences on the interior term with a
    parameter that has no discontinuities in view (radiance), and linearity in the tangent."""
    from oracle.psdr_oracle import OracleScene

    def scene(rad, d_rad):
        sc = OracleScene(48, 48, 4, 0, 0)
        for name, refl in scenes.CBOX_BSDFS:
            sc.add_diffuse(name, refl)
        for m in scenes.cbox_meshes():
            sc.add_mesh(m.v, m.f, m.bsdf, uv=m.uv, fuv=m.fuv, to_world={"raw": m.to_world},
                        radiance=rad if m.emitter is not None else None, d_radiance=d_rad if m.emitter is not None else None)
        c = scenes.CBOX_CAMERA
        sc.add_camera(c["fov"], c["near"], c["far"], {"raw": c["to_world"]})
        sc.configure((0,))
        return sc
    img, dimg = scene((20.0, 20.0, 8.0), (1.0, 2.0, 3.0)).render(2, seed=1, mode=1, terms=1)
    h = 0.5
    # primal of the AD pass (renderD re-intersects the primary hit analytically, scene.cpp:772-801, so its
    # image differs from renderC's on a few grazing lanes -- compare like with like)
    ip = scene((20.0 + h, 20.0 + 2 * h, 8.0 + 3 * h), None).render(2, seed=1, mode=1, terms=1)[0]
    im = scene((20.0 - h, 20.0 - 2 * h, 8.0 - 3 * h), None).render(2, seed=1, mode=1, terms=1)[0]
    assert rel_l2(dimg, (ip - im) / (2 * h)) < 1e-4


def test_c_abi_library_exports_every_declared_symbol():
    """The product library loads and exports exactly what include/psdr_b200.h declares (no compute)."""
    from psdr_jit_b200 import _lib, build
    path = build.build_native()
    L = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "psdr_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(psdr_[a-z_0-9]+)\s*\(", header)))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(L, name), "missing export " + name
    assert sorted(_lib.EXPORTS) == declared
    L.psdr_version.restype = ctypes.c_int
    assert L.psdr_version() >= 100
    # host-only entry point: PCG32 streams must match the reference bit for bit
    g = np.load(GOLDEN + "/sampler.npz")
    out = np.zeros((8, 8), dtype=np.float32)
    L.psdr_sampler_draws.argtypes = [ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
    assert L.psdr_sampler_draws(7, 8, 8, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float))) == 0
    assert np.array_equal(out, g["draws_seed7"])


def test_library_staleness_is_decided_by_content_not_by_file_times():
    """The built library travels to other machines by copy (file times do not survive): build.is_stale() compares a
    hash of the sources with the stamp written next to the library."""
    from psdr_jit_b200 import build
    build.build_native()
    assert os.path.exists(build.STAMP) and not build.is_stale()
    src = os.path.join(build.CSRC, build.SOURCES[0])
    st = os.stat(src)
    try:
        os.utime(src, None)                     # newer than the library: still not stale
        assert not build.is_stale()
        with open(build.STAMP) as fh:
            good = fh.read()
        with open(build.STAMP, "w") as fh:
            fh.write("0" * 64 + "\n")
        assert build.is_stale()
        with open(build.STAMP, "w") as fh:
            fh.write(good)
        assert not build.is_stale()
    finally:
        os.utime(src, (st.st_atime, st.st_mtime))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under psdr_jit_b200/ may reference it."""
    pkg = os.path.join(ROOT, "psdr_jit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "psdr_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import psdr_jit_b200 as psdr
    sc = psdr.Scene()
    sc.add_Sensor(psdr.PerspectiveCamera(60, 1e-6, 1e7))
    sc.add_BSDF(psdr.DiffuseBSDF([0.5, 0.5, 0.5]), "a")
    m = scenes.cbox_meshes()[0]
    mesh = psdr.Mesh()
    mesh.load_raw(m.v, m.f)
    sc.add_Mesh(mesh, "a", None)
    with pytest.raises(RuntimeError, match="CUDA device"):
        sc.configure()


def test_python_surface_host_logic():
    """param_map naming (scene.cpp:29-47), add_* copy semantics, RenderOption ctor rules, errors."""
    import psdr_jit_b200 as psdr
    sc = psdr.Scene()
    assert (sc.opts.width, sc.opts.height, sc.opts.spp, sc.opts.sppe, sc.opts.sppse) == (128, 128, 1, 0, 0)
    o = psdr.RenderOption(64, 32, 8)
    assert (o.spp, o.sppe, o.sppse) == (8, 8, 8)
    o = psdr.RenderOption(64, 32, 8, 4)
    assert (o.sppe, o.sppse) == (4, 4)
    cam = psdr.PerspectiveCamera(60, 1e-6, 1e7)
    sc.add_Sensor(cam)
    b = psdr.DiffuseBSDF([0.1, 0.2, 0.3])
    sc.add_BSDF(b, "red")
    b.reflectance[:] = 9.0                                           # the scene owns a copy
    assert np.allclose(sc.param_map["BSDF[id=red]"].reflectance, [0.1, 0.2, 0.3])
    assert sc.param_map["BSDF[0]"] is sc.param_map["BSDF[id=red]"]
    with pytest.raises(RuntimeError, match="Duplicate BSDF id"):
        sc.add_BSDF(psdr.DiffuseBSDF(), "red")
    m = scenes.cbox_meshes()[0]
    mesh = psdr.Mesh()
    mesh.load_raw(m.v, m.f)
    with pytest.raises(RuntimeError, match="Unknown BSDF id"):
        sc.add_Mesh(mesh, "nope", None)
    sc.add_Mesh(mesh, "red", psdr.AreaLight([1.0, 2.0, 3.0]))
    assert sc.num_meshes == 1 and sc.get_num_emitters() == 1 and sc.num_sensors == 1
    assert set(sc.param_map) >= {"Mesh[0]", "Emitter[0]", "Sensor[0]", "BSDF[0]", "BSDF[id=red]"}
    sc.param_map["Mesh[0]"].set_transform(scenes.translate(1, 2, 3))
    sc.param_map["Mesh[0]"].append_transform(scenes.translate(1, 0, 0))
    assert np.allclose(sc._meshes[0].to_world_left[:3, 3], [2, 2, 3])
    with pytest.raises(RuntimeError, match="must be configured"):
        psdr.PathTracer(1).renderC(sc, 0, seed=0)
    with pytest.raises(RuntimeError, match="Missing meshes"):
        psdr.Scene().configure()


def test_microfacet_vs_reference(oracle):
    """MicrofacetBSDF (microfacet.cpp, ggx.cpp): renderC and every term of renderD against the running reference."""
    g = np.load(GOLDEN + "/mf_renderC.npz")
    for depth, seed in ((1, 0), (3, 3)):
        img = build_oracle(scenes.cbox_meshes(), 128, 128, 4, 0, 0, bsdfs=scenes.CBOX_MF_BSDFS).render(depth, seed=seed, mode=0)
        r, nbad, r_ex = compare_stats(img, g["img_d%d_seed%d" % (depth, seed)], flip_rel=2e-5)
        assert r < 5e-3 and nbad <= 40 and r_ex < 2e-5, (depth, r, nbad, r_ex)   # r carries the few flipped lanes
    g = np.load(GOLDEN + "/mf_renderD_128_s4_d2_smallbox.npz")
    for term, spps, scale in (("interior", (4, 0, 0), 2.0), ("primary", (0, 4, 0), 1.0), ("secondary", (0, 0, 4), 2.0)):
        osc = build_oracle(scenes.cbox_meshes(), 128, 128, *spps, move_mesh=1, axis_scale=(0.0, 30.0, 50.0), bsdfs=scenes.CBOX_MF_BSDFS)
        img, dimg = osc.render(2, seed=5, mode=1, terms=7)
        if spps[0]:
            r, nbad, r_ex = compare_stats(img, g["img_" + term])
            assert nbad <= 120 and r_ex < 1e-3, (term, r, nbad, r_ex)
        r, nbad, r_ex = compare_stats(dimg * scale, g["grad_" + term])
        assert nbad <= 0.02 * len(dimg) and r_ex < 1e-3, (term, r, nbad, r_ex)


def test_guided_secondary_edges_vs_reference(oracle):
    """HyperCubeDistribution3f semantics pinned on the cases that do not depend on the Monte-Carlo mass values:
    a 1-cell grid is the identity, a grid whose mass sits in a cell without moving edges yields exactly zero.
    Grids whose result depends on the mass values are chaotic in the reference itself (the pre-pass keeps a
    sample only if two fp32 reconstructions of the same point agree to 1e-3 while their typical distance is
    3e-4) -- for those only the order of magnitude is compared."""
    g = np.load(GOLDEN + "/guided_probe.npz")
    kw = dict(move_mesh=8, axis_scale=(40.0, 20.0, 0.0))

    def run(reso, seed):
        osc = build_oracle(sphere_meshes(), 128, 128, 0, 0, 8, **kw)
        osc.preprocess_secondary_edges(0, reso, 1, seed)
        return osc.render(2, seed=1, mode=1, terms=4)[1] * 2.0      # reference tangent scaling
    r, nbad, r_ex = compare_stats(run([1, 1, 1, 8], 0), g["g111"])
    assert nbad <= 100 and r_ex < 3e-4, (r, nbad, r_ex)
    assert np.abs(g["g311"]).max() == 0.0 and np.abs(run([3, 1, 1, 64], 3)).max() == 0.0
    for name, reso in (("g211", [2, 1, 1, 64]), ("g112", [1, 1, 2, 64]), ("g222", [2, 2, 2, 64])):
        a, b = np.abs(run(reso, 0)).sum(), np.abs(g[name]).sum()
        assert 0.6 < a / b < 1.6, (name, a, b)


def _golden_env():
    from tests.common import test_envmap
    c, s_ = np.cos(0.7), np.sin(0.7)
    return dict(data=test_envmap(32, 16), w=32, h=16, scale=1.5,
                to_world=np.array([[c, 0, s_, 0], [0, 1, 0, 0], [-s_, 0, c, 0], [0, 0, 0, 1]], np.float32))


def test_envmap_vs_reference(oracle):
    """EnvironmentMap (envmap.cpp, bitmap.cpp envmap mode, bounding mesh + emitter weights of scene.cpp:435-515):
    two emitters (envmap first, area light second), Diffuse and Microfacet materials, against the running reference."""
    g = np.load(GOLDEN + "/env_renderC.npz")
    for tag, bs in (("diffuse", None), ("mf", scenes.CBOX_MF_BSDFS)):
        for depth, seed in ((1, 0), (3, 3)):
            img = build_oracle(scenes.cbox_meshes(), 128, 128, 4, 0, 0, bsdfs=bs, envmap=_golden_env()).render(depth, seed=seed, mode=0)
            r, nbad, r_ex = compare_stats(img, g["img_%s_d%d_seed%d" % (tag, depth, seed)], flip_rel=2e-5)
            assert r < 1e-2 and nbad <= 60 and r_ex < 2e-5, (tag, depth, r, nbad, r_ex)
    g = np.load(GOLDEN + "/env_renderD_128_s4_d3_smallbox.npz")
    kw = dict(move_mesh=1, axis_scale=(0.0, 30.0, 50.0), bsdfs=scenes.CBOX_MF_BSDFS, envmap=_golden_env())
    img, d = build_oracle(scenes.cbox_meshes(), 128, 128, 4, 0, 0, **kw).render(3, seed=5, mode=1, terms=7)
    r, nbad, r_ex = compare_stats(img, g["img_interior"], flip_rel=2e-5)
    assert nbad <= 0.03 * len(img) and r_ex < 1e-4, (r, nbad, r_ex)
    r, nbad, r_ex = compare_stats(d * 2.0, g["grad_interior"], flip_rel=2e-5)      # reference tangent scaling
    assert nbad <= 0.05 * len(img) and r_ex < 2e-4, (r, nbad, r_ex)


def _golden_textures(seed=4):
    rng = np.random.default_rng(seed)
    out = {}
    for name, (w, h) in (("white", (8, 6)), ("cat", (5, 7))):
        out[name] = (rng.random((h * w, 3), dtype=np.float32) * 0.8 + 0.1, w, h, None)
        rng.normal(size=(h * w, 3))
    return out


def test_textured_reflectance_vs_reference(oracle):
    """Bitmap3fD::eval with more than one texel (bitmap.cpp:46-131) on the meshes with UVs, and the derivative
    through the texture coordinate of the primary hit (camera translation), against the running reference."""
    from oracle.psdr_oracle import OracleScene
    g = np.load(GOLDEN + "/tex_render.npz")
    img = build_oracle(scenes.cbox_meshes(), 128, 128, 4, 0, 0, textures=_golden_textures()).render(3, seed=3, mode=0)
    r, nbad, r_ex = compare_stats(img, g["img_d3_seed3"], flip_rel=2e-5)
    assert nbad <= 60 and r_ex < 1e-6, (r, nbad, r_ex)
    osc = OracleScene(128, 128, 4, 0, 0)
    tex = _golden_textures()
    for name, refl in scenes.CBOX_BSDFS:
        osc.add_diffuse(name, refl)
        if name in tex:
            osc.set_bsdf_texture(name, *tex[name])
    for m in scenes.cbox_meshes():
        osc.add_mesh(m.v, m.f, m.bsdf, uv=m.uv, fuv=m.fuv, to_world={"raw": m.to_world}, radiance=m.emitter)
    c = scenes.CBOX_CAMERA
    from tests.common import translation_tangent
    osc.add_camera(c["fov"], c["near"], c["far"], {"raw": c["to_world"]}, d_to_world={"left": translation_tangent((3.0, -2.0, 1.0))})
    osc.configure((0,))
    i, d = osc.render(2, seed=6, mode=1, terms=1)
    r, nbad, r_ex = compare_stats(d * 2.0, g["grad_cam"], flip_rel=2e-5)
    assert nbad <= 0.02 * len(d) and r_ex < 2e-4, (r, nbad, r_ex)


def test_oracle_texture_slots_vs_reference_golden(oracle):
    """Oracle vs the running reference for MicrofacetBSDF(Bitmap3fD, Bitmap3fD, Bitmap1fD) with the bitmaps' uv
    transforms (tests/golden/tex_slots.npz, tools/ref_golden5.py)."""
    from tests.common import compare_stats
    from tests.test_gpu_parity import _slot_textures
    g = np.load(os.path.join(GOLDEN, "tex_slots.npz"))
    got = build_oracle(scenes.cbox_meshes(), 128, 128, 4, 0, 0, bsdfs=scenes.CBOX_MF_BSDFS, textures=_slot_textures()).render(3, seed=5, mode=0)
    r, nbad, r_ex = compare_stats(got, g["img_d3_seed5"], flip_rel=1e-4)
    assert nbad < 60 and r_ex < 3e-5, (r, nbad, r_ex)                  # 25 pixels hold a flipped lane, the rest agrees to 1.2e-5
    texd = _slot_textures(with_tangent=True, texel_tangents=False)
    img, d = build_oracle(scenes.cbox_meshes(), 128, 128, 4, 0, 0, bsdfs=scenes.CBOX_MF_BSDFS, textures=texd).render(2, seed=8, mode=1, terms=1)
    r, nbad, r_ex = compare_stats(img, g["img_uv"], flip_rel=1e-4)
    assert nbad < 300 and r_ex < 1e-4, (r, nbad, r_ex)
    r, nbad, r_ex = compare_stats(2 * d, g["grad_uv"], flip_rel=1e-3)  # the reference's 2x interior scaling
    assert nbad < 250 and r_ex < 5e-3, (r, nbad, r_ex)


def test_oracle_direct_integrator_vs_reference_golden(oracle):
    """Direct(mis) (tests/golden/direct.npz): the reference binary returns 2x the radiance and 2x the interior derivative."""
    g = np.load(os.path.join(GOLDEN, "direct.npz"))
    osc = build_oracle(scenes.cbox_meshes(), 128, 128, 4, 0, 0, move_mesh=0, axis_scale=(100.0, 0.0, 0.0))
    for mis in (0, 1, 2):
        osc.set_mis(mis)
        r, nbad, r_ex = compare_stats(2.0 * osc.render(1, seed=0, mode=0), g["imgC_mis%d" % mis], flip_rel=2e-5)
        assert nbad < 250 and r_ex < 1e-5, (mis, r, nbad, r_ex)
        img, d = osc.render(1, seed=0, mode=1, terms=1)
        r, nbad, r_ex = compare_stats(2.0 * d, g["gradD_int_mis%d" % mis], flip_rel=1e-3)
        assert nbad < 250 and r_ex < 1e-4, (mis, r, nbad, r_ex)


def test_oracle_roughconductor_vs_reference_golden(oracle):
    """RoughConductorBSDF (tests/golden/conductor.npz, tools/ref_golden8.py: the RUNNING reference, gold eta / k of
    tutorials/batch_render.ipynb on the two blocks).  The specular lobe amplifies the reference's approximate rcp / sqrt:
    a few pixels flip a hit; the rest agrees to ~1e-4 (image) / ~2e-3 (derivative images, which the reference scales by 2)."""
    g = np.load(os.path.join(GOLDEN, "conductor.npz"))
    eta, k = g["eta"], g["k"]

    def bs(alpha):
        return [(n, {"conductor": (alpha, eta, k)}) if n == "cat" else (n, p) for n, p in scenes.CBOX_BSDFS]
    for tag, alpha in (("a15", 0.15), ("a01", 0.01)):
        img = build_oracle(scenes.cbox_meshes(), 128, 128, 4, 0, 0, bsdfs=bs(alpha)).render(3, seed=0, mode=0)
        r, nbad, r_ex = compare_stats(img, g["imgC_" + tag])
        assert nbad <= 16 and r_ex < 5e-4, (tag, r, nbad, r_ex)
    img, d = build_oracle(scenes.cbox_meshes(), 128, 128, 4, 0, 0, bsdfs=bs(0.15), move_mesh=0, axis_scale=(100.0, 0.0, 0.0)).render(3, seed=0, mode=1, terms=1)
    r, nbad, r_ex = compare_stats(img, g["imgD_geo"])
    assert nbad <= 128 and r_ex < 2e-3, (r, nbad, r_ex)
    r, nbad, r_ex = compare_stats(2.0 * d, g["gradD_geo"])
    assert nbad <= 256 and r_ex < 5e-3, (r, nbad, r_ex)
    for tag, j in (("alpha", 0), ("eta", 1), ("k", 5)):
        dd = np.zeros(10, np.float32)
        dd[j] = 1.0
        _, d = build_oracle(scenes.cbox_meshes(), 128, 128, 4, 0, 0, bsdfs=bs(0.15), d_bsdf={"cat": dd}).render(3, seed=0, mode=1, terms=1)
        r, nbad, r_ex = compare_stats(2.0 * d, g["gradD_" + tag])
        assert nbad <= 256 and r_ex < 5e-3, (tag, r, nbad, r_ex)


def test_oracle_ext_bsdfs_vs_reference_golden(oracle):
    """MicrofacetBSDFPerVertex, NormalMapBSDF (add_normalmap_BSDF and add_BSDF's defaults) and RoughDielectricBSDF (through
    a scene file) on the tall block: tests/golden/ext_bsdfs.npz, tools/ref_golden10.py -- the RUNNING reference.  As for the
    other glossy BSDFs a few pixels of 16 384 flip a hit or a lobe choice under the reference's approximate arithmetic
    (8-15 in renderC, up to ~160 in the derivative images, where single samples reach 1e2-1e4); the rest agrees to 2-4e-4
    (images) / 1-5e-3 (derivative images, which the reference scales by 2)."""
    import copy
    g = np.load(os.path.join(GOLDEN, "ext_bsdfs.npz"))
    spp = int(g["spp"])

    def meshes():
        ms = copy.deepcopy(scenes.cbox_meshes())
        for m in ms:
            if m.name == "largebox":
                m.bsdf = "ext"
        return ms

    def bs(spec):
        return list(scenes.CBOX_BSDFS) + [("ext", spec)]
    nm = {"normalmap": {"normal": tuple(g["nm_normal"]), "nested": (list(g["nm_spec"]), list(g["nm_diff"]), float(g["nm_rough"]))}}
    cases = {"pervertex": {"pervertex": (g["pv_spec"], g["pv_diff"], g["pv_rough"])}, "normalmap": nm,
             "normalmap_default": {"normalmap": {"normal": (0.499999, 0.499999, 1.0), "nested": ([0.04] * 3, [0.5] * 3, 0.8)}},
             "dielectric": {"dielectric": (0.2, 1.5, 1.0)}}
    for tag, spec in cases.items():
        img = build_oracle(meshes(), 128, 128, spp, 0, 0, bsdfs=bs(spec)).render(3, seed=0, mode=0)
        r, nbad, r_ex = compare_stats(img, g["imgC_" + tag])
        assert nbad <= 32 and r_ex < 6e-4, (tag, r, nbad, r_ex)
    for tag in ("normalmap", "dielectric"):
        img, d = build_oracle(meshes(), 128, 128, spp, 0, 0, bsdfs=bs(cases[tag]), move_mesh=0, axis_scale=(100.0, 0.0, 0.0)).render(3, seed=0, mode=1, terms=1)
        r, nbad, r_ex = compare_stats(img, g["imgD_" + tag])
        assert nbad <= 128 and r_ex < 1e-3, (tag, r, nbad, r_ex)
        r, nbad, r_ex = compare_stats(2.0 * d, g["gradD_" + tag])
        assert nbad <= 256 and r_ex < 5e-3, (tag, r, nbad, r_ex)
    d_pv = np.zeros((8, 7), np.float32)
    d_pv[:, 6] = 1.0
    for tag, spec, dd in (("pervertex_rough", cases["pervertex"], d_pv), ("normalmap_normal", nm, np.float32([1, 0, 0]))):
        _, d = build_oracle(meshes(), 128, 128, spp, 0, 0, bsdfs=bs(spec), d_bsdf={"ext": dd}).render(3, seed=0, mode=1, terms=1)
        r, nbad, r_ex = compare_stats(2.0 * d, g["gradD_" + tag])
        assert nbad <= 256 and r_ex < 7e-3, (tag, r, nbad, r_ex)


def test_oracle_pervertex_tangent_matches_finite_differences(oracle):
    """d/d(per-vertex diffuse colour): the oracle's forward-mode image against central differences (a parameter outside the
    detached sampling densities, so the two agree per seed)"""
    import copy
    rng = np.random.default_rng(5)
    sp, df, rg = rng.uniform(0.02, 0.9, (8, 3)).astype(np.float32), rng.uniform(0.05, 0.8, (8, 3)).astype(np.float32), rng.uniform(0.15, 0.9, 8).astype(np.float32)
    ms = copy.deepcopy(scenes.cbox_meshes())
    for m in ms:
        if m.name == "largebox":
            m.bsdf = "ext"

    def render(eps, tangent):
        d2 = df.copy()
        d2[:, 1] += np.float32(eps)
        d = np.zeros((8, 7), np.float32)
        d[:, 4] = 1.0
        sc = build_oracle(ms, 48, 48, 16, 0, 0, bsdfs=list(scenes.CBOX_BSDFS) + [("ext", {"pervertex": (sp, d2, rg)})], d_bsdf={"ext": d} if tangent else None)
        return sc.render(2, seed=4, mode=1, terms=1)[1 if tangent else 0]
    h = 1e-2
    fd = (render(h, False) - render(-h, False)) / (2 * h)
    ad = render(0.0, True)
    assert np.abs(ad).max() > 0 and rel_l2(ad, fd) < 1e-3


def test_oracle_field_integrator_vs_reference_golden(oracle):
    """FieldExtractionIntegrator (tests/golden/fields.npz, tools/ref_golden9.py: the RUNNING reference).  Like Direct, the
    reference binary returns exactly 2x the field (a fully covered pixel of "silhouette" reads 2) and 2x its interior
    derivative; with that factor the images agree to 1e-6.  Its primary-edge part is NOT reproduced: it comes out at about a
    sixth of the finite-difference-correct value (the covered area of the moving box changes by -56 pixels per unit P; the
    reference's edge term sums to -9.9); ours matches finite differences (below, and tests/test_gpu_fields.py)."""
    g = np.load(os.path.join(GOLDEN, "fields.npz"))
    for tag, meshes in (("box", scenes.cbox_meshes()), ("open", scenes.cbox_meshes()[:3])):
        osc = build_oracle(meshes, 128, 128, 4, 0, 0, move_mesh=1, axis_scale=(30.0, 10.0, 0.0))
        for f in ("depth", "position", "shNormal", "geoNormal", "silhouette", "uv"):
            img, d = osc.field_render_d(f, seed=0, terms=1)
            r, nbad, r_ex = compare_stats(2.0 * img, g["%s_%s_C" % (tag, f)], flip_rel=1e-4)
            assert nbad <= 4 and r_ex < 2e-6, (tag, f, r, nbad, r_ex)
            gi = g["%s_%s_G_int" % (tag, f)]
            if np.abs(gi).max() > 0:
                r, nbad, r_ex = compare_stats(2.0 * d, gi, flip_rel=1e-4)
                assert nbad <= 8 and r_ex < 1e-5, (tag, f, r, nbad, r_ex)
            else:
                assert np.abs(d).max() == 0
    # the edge part against finite differences of the covered area
    def area(P):
        ms = scenes.cbox_meshes()[:3]
        tw = ms[1].to_world.copy()
        tw[0, 3] += 30 * P
        tw[1, 3] += 10 * P
        ms[1].to_world = tw
        return float(build_oracle(ms, 128, 128, 64, 0, 0).field_render_d("silhouette", seed=1, terms=0)[0][:, 0].sum())
    fd = (area(0.05) - area(-0.05)) / 0.1
    osc = build_oracle(scenes.cbox_meshes()[:3], 128, 128, 4, 16, 0, move_mesh=1, axis_scale=(30.0, 10.0, 0.0))
    est = float(osc.field_render_d("silhouette", seed=1, terms=2)[1][:, 0].sum())
    assert abs(est - fd) < 0.03 * abs(fd), (est, fd)
    ref_edges = float((g["open_silhouette_G_all"] - g["open_silhouette_G_int"])[:, 0].sum())
    assert abs(ref_edges) < 0.3 * abs(fd)          # the reference's own edge term: not finite-difference consistent


def test_oracle_intrinsics_camera_matches_fov_camera(oracle):
    """PerspectiveCamera(fx, fy, cx, cy, near, far) (perspective.h:11-12, transform.h:63-71): with fx = fy = cot(fov/2)/2 and a
    centred principal point it is the fov camera of a square image; an off-centre principal point shifts the image"""
    import math
    f = 0.5 / math.tan(math.radians(30.0))
    a = build_oracle(scenes.cbox_meshes(), 48, 48, 2, 0, 0).render(2, seed=1, mode=0)
    b = build_oracle(scenes.cbox_meshes(), 48, 48, 2, 0, 0, cam=dict(scenes.CBOX_CAMERA, intrinsics=(f, f, 0.5, 0.5))).render(2, seed=1, mode=0)
    assert rel_l2(b, a) < 2e-3          # the two matrices round differently: a handful of grazing lanes flip
    c = build_oracle(scenes.cbox_meshes(), 48, 48, 2, 0, 0, cam=dict(scenes.CBOX_CAMERA, intrinsics=(f, f, 0.25, 0.5))).render(2, seed=1, mode=0)
    assert rel_l2(c, a) > 0.2
    g = os.path.join(GOLDEN, "intrinsics.npz")
    if os.path.exists(g):
        g = np.load(g)
        cam = dict(scenes.CBOX_CAMERA, intrinsics=tuple(float(x) for x in g["intrinsics"]))
        img = build_oracle(scenes.cbox_meshes(), 128, 128, 4, 0, 0, cam=cam).render(2, seed=0, mode=0)
        r, nbad, r_ex = compare_stats(img, g["imgC"])
        assert nbad <= 16 and r_ex < 3e-4, (r, nbad, r_ex)      # measured: 4 flipped pixels, 1.2e-4 on the rest


def test_oracle_collocated_vs_reference_golden(oracle):
    """CollocatedIntegrator (tests/golden/collocated.npz, tools/ref_golden12.py: the RUNNING reference).  As for Direct and
    FieldExtraction the binary returns exactly 2x what collocated.cpp computes, and 2x its interior derivative; its
    primary-edge image has the oracle's non-zero pixels but values that are not finite-difference consistent (DESIGN.md
    section 5), so only the support is compared."""
    g = np.load(os.path.join(GOLDEN, "collocated.npz"))
    osc = build_oracle(scenes.cbox_meshes(), 128, 128, 4, 0, 0)
    osc.set_collocated(float(g["intensity"]))
    r, nbad, r_ex = compare_stats(2.0 * osc.render(1, seed=0, mode=0), g["imgC"])
    assert nbad <= 4 and r_ex < 1e-5, (r, nbad, r_ex)
    osc = build_oracle(scenes.cbox_meshes(), 128, 128, 4, 0, 0, move_mesh=1, axis_scale=(30.0, 10.0, 0.0))
    osc.set_collocated(float(g["intensity"]))
    r, nbad, r_ex = compare_stats(2.0 * osc.render(1, seed=0, mode=1, terms=1)[1], g["gradD_int"])
    assert nbad <= 8 and r_ex < 1e-5, (r, nbad, r_ex)
    osc = build_oracle(scenes.cbox_meshes(), 128, 128, 0, 4, 0, move_mesh=1, axis_scale=(30.0, 10.0, 0.0))
    osc.set_collocated(float(g["intensity"]))
    d = osc.render(1, seed=0, mode=1, terms=2)[1]
    assert ((np.abs(d).max(axis=1) > 0) == (np.abs(g["gradD_pri"]).max(axis=1) > 0)).all()


def test_oracle_orthographic_camera_vs_reference_golden(oracle):
    """OrthographicCamera (tests/golden/ortho.npz, tools/ref_golden13.py: the RUNNING reference on the Cornell box shrunk by
    300): image 2.4e-7, primary-edge derivative 1.6e-7, interior derivative 3.9e-4 (which the reference scales by 2)"""
    g = np.load(os.path.join(GOLDEN, "ortho.npz"))
    ms, cam = scenes.scaled_cbox(1.0 / 300.0)
    cam = dict(cam, ortho=True, near=1e-3, far=1e3)
    r, nbad, r_ex = compare_stats(build_oracle(ms, 128, 128, 4, 0, 0, cam=cam).render(2, seed=0, mode=0), g["imgC"])
    assert nbad <= 4 and r_ex < 1e-5, (r, nbad, r_ex)
    for tag, (spp, sppe), term, f, tol in (("int", (4, 0), 1, 2.0, 2e-3), ("pri", (0, 4), 2, 1.0, 1e-5)):
        _, d = build_oracle(ms, 128, 128, spp, sppe, 0, cam=cam, move_mesh=1, axis_scale=(0.1, 0.03, 0.0)).render(2, seed=0, mode=1, terms=term)
        r, nbad, r_ex = compare_stats(f * d, g["gradD_" + tag])
        assert nbad <= 8 and r_ex < tol, (tag, r, nbad, r_ex)


def test_oracle_scaled_cfg2_vs_reference_golden(oracle):
    """BASELINE config 2's scene and derivative at a scene scale of 1/300 (tests/golden/scaled_cfg2.npz, tools/ref_golden14.py:
    the RUNNING reference).  With coordinates ~2 no decision sits on the reference's fixed epsilon bands any more: the image
    agrees to 3.7e-5 with no flipped pixel (north-star bar 1e-4), the primary-edge derivative to 7e-7, the interior
    derivative to 2e-6 outside ONE pixel.  The secondary-edge term is the subject of the next test."""
    g = np.load(os.path.join(GOLDEN, "scaled_cfg2.npz"))
    s, ax = float(g["scale"]), float(g["axis"])
    ms, cam = scenes.scaled_cbox(s)
    img = build_oracle(ms, 128, 128, 4, 0, 0, cam=cam).render(3, seed=0, mode=0)
    assert rel_l2(img, g["imgC"]) < 1e-4
    for tag, (spp, sppe), term, f, tol in (("int", (4, 0), 1, 2.0, 1e-5), ("pri", (0, 4), 2, 1.0, 1e-5)):
        im, d = build_oracle(ms, 128, 128, spp, sppe, 0, cam=cam, move_mesh=0, axis_scale=(ax, 0.0, 0.0)).render(3, seed=0, mode=1, terms=term)
        assert rel_l2(im, g["imgD_" + tag]) < 1e-4
        r, nbad, r_ex = compare_stats(f * d, g["gradD_" + tag])
        assert nbad <= 2 and r_ex < tol, (tag, r, nbad, r_ex)


def test_reference_secondary_edge_derivative_is_not_deterministic(oracle):
    """tests/golden/ref_determinism_sec.npz (tools/ref_probe5.py): the reference's forward-mode secondary-edge derivative image
    rendered FOUR times with identical scene, seed and parameters, at full scale and at scale 1/300.  The four images
    differ from each other in two thirds of their non-zero pixels (rel-L2 between two runs 0.13 - 0.18; interior and
    primary-edge terms are reproducible): individual (pixel, channel) sums come out lower by a different amount in every
    run -- lost updates in the reference's accumulation under forward-mode AD.  Ours is the complete sum: it exceeds every
    run's total, bounds 98 % of the entries in magnitude, and coincides with at least one of the four runs in three
    quarters of the entries.  This is the floor of the gradient-image parity; no implementation can
    agree with a moving target more closely than it agrees with itself."""
    D = np.load(os.path.join(GOLDEN, "ref_determinism_sec.npz"))
    for tag, (ms, cam), ax in (("scaled", scenes.scaled_cbox(1.0 / 300.0), 100.0 / 300.0), ("full", (scenes.cbox_meshes(), scenes.CBOX_CAMERA), 100.0)):
        a = D[tag].astype(np.float64)                      # [4 runs, npix, 3]
        d = 2.0 * build_oracle(ms, 128, 128, 0, 0, 4, cam=cam, move_mesh=0, axis_scale=(ax, 0.0, 0.0)).render(3, seed=0, mode=1, terms=4)[1]
        self_spread = min(rel_l2(a[i], a[j]) for i in range(4) for j in range(i))
        assert self_spread > 0.08, (tag, self_spread)      # the reference does not reproduce itself
        assert max(rel_l2(d, a[i]) for i in range(4)) < 2.0 * max(rel_l2(a[i], a[j]) for i in range(4) for j in range(i)), tag
        nz = np.abs(d) > 0
        tol = 2e-4 * np.abs(d).max()
        assert (np.abs(a - d[None]) < tol).any(axis=0)[nz].mean() > 0.7, tag          # equal to ours in some run
        assert (np.abs(d)[None] >= np.abs(a) - tol)[:, nz].mean() > 0.97, tag         # lost updates only lower a sum
        assert all(np.abs(d).sum() > np.abs(a[i]).sum() * 1.05 for i in range(4)), tag
