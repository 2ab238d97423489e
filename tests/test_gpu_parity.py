"""GPU parity: the CUDA path (through the psdr_jit-style surface -> C ABI -> sm_100a kernels)
against the CPU oracle on identical seeded inputs, and against golden vectors produced by the
unmodified reference.  Tolerance: rel-L2 < 1e-4 (BASELINE.json north_star) for radiance and the
forward-mode derivative image; hit ids bit-exact."""
import os

import numpy as np
import pytest

from tests.common import (GOLDEN, build_oracle, build_product, compare_stats, rel_l2, scenes, sphere_meshes, translation_tangent)

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _psdr():
    import psdr_jit_b200 as psdr
    return psdr


def test_extension_loaded_and_fails_without_fallback():
    psdr = _psdr()
    from psdr_jit_b200 import _lib
    L = _lib.load()
    assert L.psdr_version() >= 100
    import os
    assert os.path.exists(_lib.lib_path())


def test_aov_hit_ids_bit_exact(oracle):
    psdr = _psdr()
    for meshes in (scenes.cbox_meshes(), sphere_meshes()):
        ref = build_oracle(meshes, 128, 128, 1, 0, 0).aov()
        sc = build_product(meshes, 128, 128, 1, 0, 0)
        got = psdr.PathTracer(1).render_aov(sc, 0, seed=0).cpu().numpy()
        assert np.array_equal(got[:, 0], ref[:, 0])          # mesh ids
        assert np.array_equal(got[:, 1], ref[:, 1])          # triangle ids
        np.testing.assert_allclose(got[:, 2:12], ref[:, 2:12], rtol=0, atol=1e-4)


def test_aov_vs_reference_golden():
    psdr = _psdr()
    g = np.load(GOLDEN + "/aov.npz")
    for name, meshes in (("cbox", scenes.cbox_meshes()), ("cboxsphere", sphere_meshes())):
        sc = build_product(meshes, 128, 128, 1, 0, 0)
        got = psdr.PathTracer(1).render_aov(sc, 0, seed=0).cpu().numpy()
        # the reference's FieldExtractionIntegrator output is exactly 2x the field (vcall quirk)
        assert np.array_equal(got[:, 0] * 2, g[name + "_segmentation"][:, 0])
        assert np.abs(got[:, 2:5] - g[name + "_position"] / 2).max() < 2e-3
        assert np.abs(got[:, 6:9] - g[name + "_geoNormal"] / 2).max() < 1e-5
        assert np.abs(got[:, 9:12] - g[name + "_shNormal"] / 2).max() < 1e-4


@pytest.mark.parametrize("depth,spp,seed", [(1, 1, 0), (3, 4, 3), (0, 2, 1), (6, 2, 11)])
def test_renderC_vs_oracle(oracle, depth, spp, seed):
    psdr = _psdr()
    ref = build_oracle(scenes.cbox_meshes(), 128, 128, spp, 0, 0).render(depth, seed=seed, mode=0)
    sc = build_product(scenes.cbox_meshes(), 128, 128, spp, 0, 0)
    got = psdr.PathTracer(depth).renderC(sc, 0, seed=seed).cpu().numpy()
    assert rel_l2(got, ref) < TOL


def test_renderC_cfg1_vs_reference_golden():
    psdr = _psdr()
    g = np.load(GOLDEN + "/cfg1_renderC.npz")
    sc = build_product(scenes.cbox_meshes(), 128, 128, 1, 0, 0)
    got = psdr.PathTracer(1).renderC(sc, 0, seed=0).cpu().numpy()
    r, nbad, r_ex = compare_stats(got, g["img"])
    # grazing shadow rays flip between OptiX and our tracer on a handful of lanes
    assert nbad <= 4 and r_ex < 1e-4 and r < 5e-3


CASES = [
    ("light", "cbox", 3, 0, 0, (100.0, 0.0, 0.0)),
    ("smallbox", "cbox", 2, 5, 1, (0.0, 30.0, 50.0)),
    ("sphere", "sphere", 2, 1, 8, (40.0, 20.0, 0.0)),
]


@pytest.mark.parametrize("name,scene,depth,seed,mesh,axis", CASES)
@pytest.mark.parametrize("terms", [1, 2, 4, 7])
def test_renderD_terms_vs_oracle(oracle, name, scene, depth, seed, mesh, axis, terms):
    psdr = _psdr()
    meshes = scenes.cbox_meshes() if scene == "cbox" else sphere_meshes()
    spps = (4 if terms & 1 else 0, 4 if terms & 2 else 0, 4 if terms & 4 else 0)
    osc = build_oracle(meshes, 128, 128, *spps, move_mesh=mesh, axis_scale=axis)
    img_ref, dimg_ref = osc.render(depth, seed=seed, mode=1, terms=7)
    sc = build_product(meshes, 128, 128, *spps, move_mesh=mesh, axis_scale=axis)
    assert sc.num_primary_edges(0) == osc.num_primary_edges(0)
    assert sc.num_secondary_edges() == osc.num_secondary_edges()
    img, dimg = psdr.PathTracer(depth).renderD_fwd(sc, 0, seed=seed)
    img, dimg = img.cpu().numpy(), dimg.cpu().numpy()
    if terms & 1:
        assert rel_l2(img, img_ref) < TOL
    assert np.abs(dimg_ref).max() > 0
    assert rel_l2(dimg, dimg_ref) < TOL


@pytest.mark.parametrize("name,scene,depth,seed,mesh,axis", CASES)
def test_renderD_vs_reference_golden(name, scene, depth, seed, mesh, axis):
    """Per-term comparison with the running reference.  The reference binary's interior and
    secondary-edge tangents are exactly 2x the correct value (tests/golden/probe2_radiance.npz:
    d img / d radiance-scale = 2*img); reference_tangent_scaling reproduces that."""
    psdr = _psdr()
    tag = {"light": "renderD_128_s4_d3_light", "smallbox": "renderD_128_s4_d2_smallbox", "sphere": "renderD_128_s4_d2_sphere"}[name]
    g = np.load(GOLDEN + "/%s.npz" % tag)
    meshes = scenes.cbox_meshes() if scene == "cbox" else sphere_meshes()
    integ = psdr.PathTracer(depth)
    integ.reference_tangent_scaling = True
    for term, spps in (("interior", (4, 0, 0)), ("primary", (0, 4, 0)), ("secondary", (0, 0, 4)), ("all", (4, 4, 4))):
        sc = build_product(meshes, 128, 128, *spps, move_mesh=mesh, axis_scale=axis)
        img, dimg = integ.renderD_fwd(sc, 0, seed=seed)
        img, dimg = img.cpu().numpy(), dimg.cpu().numpy()
        if spps[0]:
            r, nbad, r_ex = compare_stats(img, g["img_" + term])
            assert nbad <= 120 and r_ex < 1e-3, (term, r, nbad, r_ex)
        r, nbad, r_ex = compare_stats(dimg, g["grad_" + term])
        assert nbad <= 0.06 * len(dimg) and r_ex < 5e-3, (term, r, nbad, r_ex)


def test_brute_force_and_bvh_agree(oracle):
    psdr = _psdr()
    out = []
    for accel in (0, 1):
        sc = build_product(sphere_meshes(), 96, 96, 2, 2, 2, move_mesh=8, axis_scale=(40.0, 20.0, 0.0), accel=accel)
        img, dimg = psdr.PathTracer(2).renderD_fwd(sc, 0, seed=4)
        aov = psdr.PathTracer(2).render_aov(sc, 0, seed=4)
        out.append((img.cpu().numpy(), dimg.cpu().numpy(), aov.cpu().numpy()))
    assert np.array_equal(out[0][2], out[1][2])
    assert rel_l2(out[0][0], out[1][0]) < 1e-6
    assert rel_l2(out[0][1], out[1][1]) < 1e-5


@pytest.mark.parametrize("extra", [1, 3])
def test_odd_triangle_count_in_the_paired_scan(oracle, extra):
    """The brute-force scan tests triangles two at a time (dscene.h bg_pair): an odd count is padded with a
    zero triangle that must never be hit, and the last real triangle must still be found."""
    psdr = _psdr()
    from psdr_jit_b200.scenes import MeshData
    v = np.array([[150, 330, 200], [400, 330, 200], [275, 330, 420], [150, 120, 420], [400, 120, 420]], np.float32)
    f = np.array([[0, 1, 2], [2, 3, 4], [0, 2, 3]][:extra], np.int32)
    fin = MeshData(name="fin", v=v, f=f, bsdf="white")
    meshes = scenes.cbox_meshes() + [fin]
    assert sum(len(m.f) for m in meshes) % 2 == 1
    osc = build_oracle(meshes, 96, 96, 2, 2, 2, move_mesh=8, axis_scale=(30.0, 10.0, 0.0))
    img_ref, dimg_ref = osc.render(3, seed=2, mode=1, terms=7)
    sc = build_product(meshes, 96, 96, 2, 2, 2, move_mesh=8, axis_scale=(30.0, 10.0, 0.0))
    integ = psdr.PathTracer(3)
    img, dimg = integ.renderD_fwd(sc, 0, seed=2)
    assert rel_l2(img.cpu().numpy(), img_ref) < TOL
    assert rel_l2(dimg.cpu().numpy(), dimg_ref) < TOL
    ref = build_oracle(meshes, 96, 96, 1, 0, 0).aov()
    got = psdr.PathTracer(1).render_aov(build_product(meshes, 96, 96, 1, 0, 0), 0, seed=0).cpu().numpy()
    assert np.array_equal(got[:, 0], ref[:, 0]) and np.array_equal(got[:, 1], ref[:, 1])     # mesh and triangle ids
    assert (got[:, 0] == 8).any()           # the fin (the last, unpaired triangles) is visible


def test_seed_continuation_matches_oracle_skip(oracle):
    """seed=-1 continues the sampler streams (reference integrator.cpp:23,60)."""
    psdr = _psdr()
    depth = 2
    sc = build_product(scenes.cbox_meshes(), 64, 64, 2, 2, 2, move_mesh=0, axis_scale=(100.0, 0.0, 0.0))
    integ = psdr.PathTracer(depth)
    integ.renderD_fwd(sc, 0, seed=9)
    img2, dimg2 = integ.renderD_fwd(sc, 0, seed=-1)
    osc = build_oracle(scenes.cbox_meshes(), 64, 64, 2, 2, 2, move_mesh=0, axis_scale=(100.0, 0.0, 0.0))
    ref_img, ref_dimg = osc.render(depth, seed=9, mode=1, terms=7, skip=(2 + 5 * depth, 1 + 10 * depth, 3))
    assert rel_l2(img2.cpu().numpy(), ref_img) < TOL
    assert rel_l2(dimg2.cpu().numpy(), ref_dimg) < TOL


def test_batch_pixels(oracle):
    psdr = _psdr()
    rng = np.random.default_rng(0)
    pix = rng.choice(96 * 96, size=500, replace=False).astype(np.int32)
    sc = build_product(scenes.cbox_meshes(), 96, 96, 4, 0, 0)
    got = psdr.PathTracer(2).renderC(sc, 0, seed=5, batch_pix=pix).cpu().numpy()
    ref = build_oracle(scenes.cbox_meshes(), 96, 96, 4, 0, 0).render(2, seed=5, mode=0, pix_id=pix)
    assert got.shape == (500, 3)
    assert rel_l2(got, ref) < TOL
    with pytest.raises(RuntimeError, match="seed must be set"):
        psdr.PathTracer(2).renderC(sc, 0, seed=-1, batch_pix=pix)


def test_hide_emitters_and_material_tangents(oracle):
    psdr = _psdr()
    from oracle.psdr_oracle import OracleScene
    w = h = 64
    osc = OracleScene(w, h, 4, 0, 0)
    for name, refl in scenes.CBOX_BSDFS:
        osc.add_diffuse(name, refl, d_refl=(0.3, 0.2, 0.1) if name == "white" else None)
    for m in scenes.cbox_meshes():
        osc.add_mesh(m.v, m.f, m.bsdf, uv=m.uv, fuv=m.fuv, to_world={"raw": m.to_world}, radiance=m.emitter,
                     d_radiance=(1.0, 2.0, 3.0) if m.emitter is not None else None)
    c = scenes.CBOX_CAMERA
    osc.add_camera(c["fov"], c["near"], c["far"], {"raw": c["to_world"]}, d_to_world={"left": translation_tangent((3.0, 1.0, 2.0))})
    osc.configure((0,))
    sc = build_product(scenes.cbox_meshes(), w, h, 4, 0, 0, d_radiance=(1.0, 2.0, 3.0), d_reflectance=("white", (0.3, 0.2, 0.1)),
                       d_cam_left=translation_tangent((3.0, 1.0, 2.0)))
    for hide in (False, True):
        ref_img, ref_dimg = osc.render(3, seed=2, mode=1, terms=1, hide_emitters=hide)
        integ = psdr.PathTracer(3)
        integ.hide_emitters = hide
        img, dimg = integ.renderD_fwd(sc, 0, seed=2, terms=1)
        assert rel_l2(img.cpu().numpy(), ref_img) < TOL
        assert rel_l2(dimg.cpu().numpy(), ref_dimg) < TOL


def test_host_buffer_api_matches_device_api():
    psdr = _psdr()
    sc = build_product(scenes.cbox_meshes(), 64, 64, 2, 2, 2, move_mesh=0, axis_scale=(100.0, 0.0, 0.0))
    integ = psdr.PathTracer(2)
    img, dimg = integ.renderD_fwd(sc, 0, seed=3)
    himg, hdimg = integ.renderD_host(sc, 0, seed=3)
    assert rel_l2(himg, img.cpu().numpy()) < 1e-6
    assert rel_l2(hdimg, dimg.cpu().numpy()) < 1e-5
    assert rel_l2(integ.renderC_host(sc, 0, seed=3), integ.renderC(sc, 0, seed=3).cpu().numpy()) < 1e-6


def test_lane_shards_sum_to_full_image():
    """Multi-GPU decomposition: every term sharded by lane range, partial full-frame images summed."""
    psdr = _psdr()
    kw = dict(move_mesh=0, axis_scale=(100.0, 0.0, 0.0))
    full = build_product(scenes.cbox_meshes(), 64, 64, 3, 2, 2, **kw)
    img, dimg = psdr.PathTracer(2).renderD_fwd(full, 0, seed=1)
    acc_i, acc_d = 0, 0
    for r in range(3):
        part = build_product(scenes.cbox_meshes(), 64, 64, 3, 2, 2, shard=(r, 3), **kw)
        a, b = psdr.PathTracer(2).renderD_fwd(part, 0, seed=1)
        acc_i, acc_d = acc_i + a, acc_d + b
    assert rel_l2(acc_i.cpu().numpy(), img.cpu().numpy()) < 1e-6
    assert rel_l2(acc_d.cpu().numpy(), dimg.cpu().numpy()) < 1e-5


def test_errors_mirror_reference_messages():
    psdr = _psdr()
    sc = psdr.Scene()
    with pytest.raises(RuntimeError, match="Missing meshes"):
        sc.configure()
    sc2 = build_product(scenes.cbox_meshes(), 32, 32, 1, 0, 0)
    with pytest.raises(RuntimeError, match="Invalid sensor id"):
        psdr.PathTracer(1).renderC(sc2, 3, seed=0)
    sc2.opts.spp = 2          # options changed -> must configure again
    sc2._native().psdr_scene_set_options(sc2._h, 32, 32, 2, 0, 0, 0)
    with pytest.raises(RuntimeError, match="must be configured"):
        psdr.PathTracer(1).renderC(sc2, 0, seed=0)


def test_full_size_properties_cfg2():
    """BASELINE config 2 (512x512, 32/32/32, depth 3): size-independent properties."""
    import torch
    psdr = _psdr()
    kw = dict(move_mesh=0, axis_scale=(100.0, 0.0, 0.0))
    sc = build_product(scenes.cbox_meshes(), 512, 512, 32, 32, 32, **kw)
    integ = psdr.PathTracer(3)
    img, dimg = integ.renderD_fwd(sc, 0, seed=0)
    img2, dimg2 = integ.renderD_fwd(sc, 0, seed=0)
    assert torch.isfinite(img).all() and torch.isfinite(dimg).all()
    # determinism up to atomic ordering
    assert rel_l2(img2.cpu().numpy(), img.cpu().numpy()) < 1e-6
    assert rel_l2(dimg2.cpu().numpy(), dimg.cpu().numpy()) < 1e-4
    # linearity: d img / d(radiance scale) == img  (interior term, tangent = radiance)
    sc_r = build_product(scenes.cbox_meshes(), 512, 512, 32, 0, 0, d_radiance=(20.0, 20.0, 8.0))
    a, da = integ.renderD_fwd(sc_r, 0, seed=0, terms=1)
    assert rel_l2(da.cpu().numpy(), a.cpu().numpy()) < 1e-5
    # against the running reference (golden, reference scaling), BASELINE config 2 at full size, seed 0.
    # What round 2's census (tools/ref_parity.py, profiles/r02c_parity_summary.json) established about the residual:
    #  * lanes whose discrete decisions differ (OptiX vs our closest hit) are sparse, zero-mean noise -- except on the tall
    #    box's side face (triangles 22 / 23), where renderD's primary hit o + t d lands on one or the other side of the face
    #    depending on how the reciprocal in ray_intersect_triangle is rounded and grazing next-event rays re-intersect the
    #    face: +1.5 % mean radiance with IEEE division, -0.4 % with Dr.Jit's rcp.approx (Scene.reference_arithmetic);
    #  * the reference's one-call derivative image contradicts the sum of its own three terms in the BLUE channel of the
    #    ~30 pixels on the luminaire's silhouette (rel-L2 0.25; red / green agree bit for bit), so the gradient is anchored
    #    on the reference's terms rendered one at a time (cfg2_512_grad_terms.npz).
    g = np.load(GOLDEN + "/cfg2_512_s32_d3_light.npz")
    gt = np.load(GOLDEN + "/cfg2_512_grad_terms.npz")
    face = gt["face22"]
    integ.reference_tangent_scaling = True
    img3, dimg3 = integ.renderD_fwd(sc, 0, seed=0)
    img3, dimg3 = img3.cpu().numpy(), dimg3.cpu().numpy()
    r, nbad, r_ex = compare_stats(img3, g["img"], flip_rel=2e-5)
    bias = float((img3[face].astype(np.float64) - g["img"][face]).sum() / g["img"][face].astype(np.float64).sum())
    print("cfg2 image vs reference: rel-L2 %.3e, %d/%d pixels with a flipped lane, rel-L2 of the rest %.3e, face 22/23 bias %+.4f"
          % (r, nbad, len(img3), r_ex, bias))
    assert r < 3e-3 and nbad < 0.035 * len(img3) and r_ex < 3e-5, (r, nbad, r_ex)     # measured: 2.29e-3, 8241 pixels, 1.3e-5
    assert 0.005 < bias < 0.03, bias                                                  # measured: +1.66 %
    assert rel_l2(dimg3, g["grad"]) > 0.1                                             # the reference's one-call blue channel (0.23)
    assert rel_l2(g["grad"], gt["grad_terms"]) > 0.1                                  # ... contradicts its own terms (0.25)
    rg, nbadg, rg_ex = compare_stats(dimg3, gt["grad_terms"], flip_rel=1e-3)
    print("cfg2 derivative image vs the reference's terms: rel-L2 %.3e, %d pixels off by > 1e-3 of the maximum, rel-L2 of the rest %.3e"
          % (rg, nbadg, rg_ex))
    assert rg < 1e-3 and nbadg <= 8, (rg, nbadg)                                      # measured: 5.7e-4, 2 pixels
    # Dr.Jit's approximate reciprocal at the analytic primary hit moves the face bias to the other side of zero
    sc.reference_arithmetic = True
    sc.configure([0])
    img4 = integ.renderD_fwd(sc, 0, seed=0)[0].cpu().numpy()
    bias4 = float((img4[face].astype(np.float64) - g["img"][face]).sum() / g["img"][face].astype(np.float64).sum())
    print("   with Scene.reference_arithmetic: rel-L2 %.3e, face 22/23 bias %+.4f" % (rel_l2(img4, g["img"]), bias4))
    assert abs(bias4) < 0.01 and rel_l2(img4, g["img"]) < r, bias4                   # measured: -0.38 %, 2.06e-3


def test_unbiased_against_reference_seed_mean():
    """Is the residual against the reference a bias or zero-mean noise of flipped lanes?  64 consecutive renderD calls
    (seed 0, then continuing streams) at 128 x 128, spp = sppe = sppse = 32, depth 3, on both sides
    (tools/ref_parity.py -> tests/golden/seedmean_128.npz); means compared."""
    psdr = _psdr()
    g = np.load(GOLDEN + "/seedmean_128.npz")
    n, face = int(g["calls"]), g["face22"]
    integ = psdr.PathTracer(3)
    integ.reference_tangent_scaling = True
    res = {}
    for tag, ra in (("ieee", False), ("refarith", True)):
        sc = build_product(scenes.cbox_meshes(), 128, 128, 32, 32, 32, move_mesh=0, axis_scale=(100.0, 0.0, 0.0))
        sc.reference_arithmetic = ra
        sc.configure([0])
        acc_i = acc_g = first_i = first_g = None
        for k in range(n):
            img, dimg = integ.renderD_fwd(sc, 0, seed=0 if k == 0 else -1)
            if k == 0:
                first_i, first_g = img.cpu().numpy(), dimg.cpu().numpy()
                acc_i, acc_g = img.double(), dimg.double()
            else:
                acc_i += img
                acc_g += dimg
        res[tag] = (first_i, first_g, (acc_i / n).cpu().numpy(), (acc_g / n).cpu().numpy())
    f_i, f_g, m_i, m_g = res["ieee"]
    one = rel_l2(f_i[~face], g["first_img"][~face])
    mean = rel_l2(m_i[~face], g["mean_img"][~face])
    bias = float((m_i[face].astype(np.float64) - g["mean_img"][face]).sum() / g["mean_img"][face].astype(np.float64).sum())
    print("image off face 22/23: rel-L2 1 call %.2e -> %d-call mean %.2e; on the face: mean bias %+.4f" % (one, n, mean, bias))
    assert mean < 1.5e-4 and mean < one / 2.5, (one, mean)            # measured 3.9e-4 -> 9.9e-5: noise, falls with the call count
    assert 0.005 < bias < 0.025, bias                                 # measured +1.46 %: the rounding artifact described above
    m_ra = res["refarith"][2]
    bias_ra = float((m_ra[face].astype(np.float64) - g["mean_img"][face]).sum() / g["mean_img"][face].astype(np.float64).sum())
    print("   with Scene.reference_arithmetic: whole image rel-L2 %.2e, face bias %+.4f" % (rel_l2(m_ra, g["mean_img"]), bias_ra))
    assert rel_l2(m_ra, g["mean_img"]) < 4e-4 and abs(bias_ra) < 0.008           # measured 2.4e-4, -0.40 %
    # derivative image: against the reference's three terms rendered one at a time (its one-call result contradicts them)
    assert rel_l2(g["mean_grad"], g["mean_grad_terms"]) > 0.03                   # measured 0.070
    r1, rm = rel_l2(f_g, g["first_grad_terms"]), rel_l2(m_g, g["mean_grad_terms"])
    print("derivative image vs the reference's terms: rel-L2 1 call %.2e -> %d-call mean %.2e" % (r1, n, rm))
    assert rm < 6e-4 and r1 < 1e-3, (r1, rm)                                     # measured 6.9e-4 -> 4.3e-4


# ---- MicrofacetBSDF (reference src/bsdf/microfacet.cpp, src/bsdf/ggx.cpp) --------------------------------
MF_TANGENT = {"cat": np.array([0.1, -0.2, 0.05, 0.02, 0.01, -0.01, 0.2], np.float32),
              "white": np.array([0.01, 0.02, 0.03, -0.1, 0.1, 0.05, -0.3], np.float32)}


@pytest.mark.parametrize("depth,spp,seed", [(1, 2, 0), (3, 4, 3), (6, 1, 2)])
def test_microfacet_renderC_vs_oracle(oracle, depth, spp, seed):
    psdr = _psdr()
    ref = build_oracle(scenes.cbox_meshes(), 128, 128, spp, 0, 0, bsdfs=scenes.CBOX_MF_BSDFS).render(depth, seed=seed, mode=0)
    sc = build_product(scenes.cbox_meshes(), 128, 128, spp, 0, 0, bsdfs=scenes.CBOX_MF_BSDFS)
    got = psdr.PathTracer(depth).renderC(sc, 0, seed=seed).cpu().numpy()
    assert np.isfinite(got).all() and rel_l2(got, ref) < TOL


@pytest.mark.parametrize("terms", [1, 2, 4, 7])
def test_microfacet_renderD_vs_oracle(oracle, terms):
    """geometry tangent (small box) + material tangents (specular, diffuse, roughness of two BSDFs)"""
    psdr = _psdr()
    spps = (4 if terms & 1 else 0, 4 if terms & 2 else 0, 4 if terms & 4 else 0)
    kw = dict(move_mesh=1, axis_scale=(0.0, 30.0, 50.0), bsdfs=scenes.CBOX_MF_BSDFS, d_bsdf=MF_TANGENT)
    img_ref, dimg_ref = build_oracle(scenes.cbox_meshes(), 128, 128, *spps, **kw).render(2, seed=5, mode=1, terms=7)
    sc = build_product(scenes.cbox_meshes(), 128, 128, *spps, **kw)
    img, dimg = psdr.PathTracer(2).renderD_fwd(sc, 0, seed=5)
    if terms & 1:
        assert rel_l2(img.cpu().numpy(), img_ref) < TOL
    assert np.abs(dimg_ref).max() > 0
    assert rel_l2(dimg.cpu().numpy(), dimg_ref) < TOL


def test_microfacet_two_sided_vs_oracle(oracle):
    from oracle.psdr_oracle import OracleScene
    psdr = _psdr()
    osc = OracleScene(64, 64, 4, 0, 0)
    for name, p in scenes.CBOX_MF_BSDFS:
        if len(p) == 3 and hasattr(p[0], "__len__"):
            osc.add_microfacet(name, p[0], p[1], p[2], two_side=True)
        else:
            osc.add_diffuse(name, p, two_side=True)
    for m in scenes.cbox_meshes():
        osc.add_mesh(m.v, m.f, m.bsdf, uv=m.uv, fuv=m.fuv, to_world={"raw": m.to_world}, radiance=m.emitter)
    c = scenes.CBOX_CAMERA
    osc.add_camera(c["fov"], c["near"], c["far"], {"raw": c["to_world"]})
    osc.configure((0,))
    ref = osc.render(3, seed=1, mode=0)
    sc = build_product(scenes.cbox_meshes(), 64, 64, 4, 0, 0, bsdfs=scenes.CBOX_MF_BSDFS, two_side=True)
    got = psdr.PathTracer(3).renderC(sc, 0, seed=1).cpu().numpy()
    assert rel_l2(got, ref) < TOL


# ---- guided secondary-edge sampling (reference src/integrator/path.cpp:130-168, src/core/cube_distrb.cpp) ----
@pytest.mark.parametrize("reso,nrounds,pseed", [([200, 4, 4, 8], 1, 0), ([64, 8, 8, 4], 2, 7)])
def test_guided_secondary_edges_vs_oracle(oracle, reso, nrounds, pseed):
    psdr = _psdr()
    kw = dict(move_mesh=8, axis_scale=(40.0, 20.0, 0.0))
    osc = build_oracle(sphere_meshes(), 128, 128, 0, 0, 8, **kw)
    mass_ref = osc.preprocess_secondary_edges(0, reso, nrounds, pseed)
    _, dimg_ref = osc.render(2, seed=1, mode=1, terms=4)
    sc = build_product(sphere_meshes(), 128, 128, 0, 0, 8, **kw)
    integ = psdr.PathTracer(2)
    plain = integ.renderD_fwd(sc, 0, seed=1)[1].cpu().numpy()
    integ.preprocess_secondary_edges(sc, 0, reso, nrounds, pseed)
    mass = integ.guiding_mass(sc, 0)
    assert mass.shape == mass_ref.shape and mass_ref.sum() > 0
    np.testing.assert_allclose(mass, mass_ref, rtol=2e-5, atol=1e-9)
    dimg = integ.renderD_fwd(sc, 0, seed=1)[1].cpu().numpy()
    assert rel_l2(dimg, dimg_ref) < TOL
    assert rel_l2(dimg, plain) > 1e-2                      # guiding changed the samples
    # the grid survives configure() and belongs to this integrator only
    sc.configure([0])
    assert rel_l2(integ.renderD_fwd(sc, 0, seed=1)[1].cpu().numpy(), dimg_ref) < TOL
    assert rel_l2(psdr.PathTracer(2).renderD_fwd(sc, 0, seed=1)[1].cpu().numpy(), plain) < 1e-6
    with pytest.raises(RuntimeError, match="nrounds > 0"):
        integ.preprocess_secondary_edges(sc, 0, reso, 0)


def test_microfacet_vs_reference_golden():
    psdr = _psdr()
    g = np.load(GOLDEN + "/mf_renderC.npz")
    sc = build_product(scenes.cbox_meshes(), 128, 128, 4, 0, 0, bsdfs=scenes.CBOX_MF_BSDFS)
    for depth, seed in ((1, 0), (3, 3)):
        got = psdr.PathTracer(depth).renderC(sc, 0, seed=seed).cpu().numpy()
        r, nbad, r_ex = compare_stats(got, g["img_d%d_seed%d" % (depth, seed)], flip_rel=2e-5)
        assert r < 5e-3 and nbad <= 40 and r_ex < 2e-5, (depth, r, nbad, r_ex)   # r carries the few flipped lanes
    g = np.load(GOLDEN + "/mf_renderD_128_s4_d2_smallbox.npz")
    integ = psdr.PathTracer(2)
    integ.reference_tangent_scaling = True
    sc = build_product(scenes.cbox_meshes(), 128, 128, 4, 4, 4, move_mesh=1, axis_scale=(0.0, 30.0, 50.0), bsdfs=scenes.CBOX_MF_BSDFS)
    img, dimg = integ.renderD_fwd(sc, 0, seed=5)
    r, nbad, r_ex = compare_stats(img.cpu().numpy(), g["img_all"])
    assert nbad <= 120 and r_ex < 1e-3, (r, nbad, r_ex)
    r, nbad, r_ex = compare_stats(dimg.cpu().numpy(), g["grad_all"])
    assert nbad <= 0.06 * len(dimg) and r_ex < 5e-3, (r, nbad, r_ex)


# ---- EnvironmentMap (reference src/emitter/envmap.cpp, src/core/bitmap.cpp envmap mode, scene.cpp:435-515) -----
def _env(rot=None, **kw):
    from tests.common import test_envmap
    e = dict(data=test_envmap(32, 16), w=32, h=16, scale=1.5)
    if rot is not None:
        c, s_ = np.cos(rot), np.sin(rot)
        e["to_world"] = np.array([[c, 0, s_, 0], [0, 1, 0, 0], [-s_, 0, c, 0], [0, 0, 0, 1]], np.float32)
    e.update(kw)
    return e


@pytest.mark.parametrize("depth,spp,seed,bs", [(1, 2, 0, "diffuse"), (3, 4, 3, "mf"), (6, 1, 2, "mf")])
def test_envmap_renderC_vs_oracle(oracle, depth, spp, seed, bs):
    psdr = _psdr()
    bsdfs = scenes.CBOX_MF_BSDFS if bs == "mf" else None
    env = _env(rot=0.7)
    ref = build_oracle(scenes.cbox_meshes(), 128, 128, spp, 0, 0, bsdfs=bsdfs, envmap=env).render(depth, seed=seed, mode=0)
    sc = build_product(scenes.cbox_meshes(), 128, 128, spp, 0, 0, bsdfs=bsdfs, envmap=env)
    assert sc.get_num_emitters() == 2
    got = psdr.PathTracer(depth).renderC(sc, 0, seed=seed).cpu().numpy()
    assert np.isfinite(got).all() and rel_l2(got, ref) < TOL


@pytest.mark.parametrize("terms", [1, 7])
def test_envmap_renderD_vs_oracle(oracle, terms):
    """tangents: small box translation, envmap texels, envmap scale, envmap rotation, camera"""
    psdr = _psdr()
    rng = np.random.default_rng(3)
    d_left = np.zeros((4, 4), np.float32)
    d_left[0, 2], d_left[2, 0] = 1.0, -1.0                      # d/dangle of a rotation about y at angle 0
    env = _env(rot=0.7, d_data=rng.normal(size=(16 * 32, 3)).astype(np.float32), d_scale=0.5, d_to_world_left=d_left)
    spps = (4 if terms & 1 else 0, 4 if terms & 2 else 0, 4 if terms & 4 else 0)
    kw = dict(move_mesh=1, axis_scale=(0.0, 30.0, 50.0), bsdfs=scenes.CBOX_MF_BSDFS, envmap=env)
    img_ref, dimg_ref = build_oracle(scenes.cbox_meshes(), 128, 128, *spps, **kw).render(3, seed=5, mode=1, terms=7)
    sc = build_product(scenes.cbox_meshes(), 128, 128, *spps, **kw)
    img, dimg = psdr.PathTracer(3).renderD_fwd(sc, 0, seed=5)
    assert rel_l2(img.cpu().numpy(), img_ref) < TOL
    assert np.abs(dimg_ref).max() > 0 and rel_l2(dimg.cpu().numpy(), dimg_ref) < TOL


def test_envmap_vs_reference_golden():
    psdr = _psdr()
    c, s_ = np.cos(0.7), np.sin(0.7)
    from tests.common import test_envmap
    env = dict(data=test_envmap(32, 16), w=32, h=16, scale=1.5,
               to_world=np.array([[c, 0, s_, 0], [0, 1, 0, 0], [-s_, 0, c, 0], [0, 0, 0, 1]], np.float32))
    g = np.load(GOLDEN + "/env_renderC.npz")
    for tag, bs in (("diffuse", None), ("mf", scenes.CBOX_MF_BSDFS)):
        sc = build_product(scenes.cbox_meshes(), 128, 128, 4, 0, 0, bsdfs=bs, envmap=env)
        for depth, seed in ((1, 0), (3, 3)):
            got = psdr.PathTracer(depth).renderC(sc, 0, seed=seed).cpu().numpy()
            r, nbad, r_ex = compare_stats(got, g["img_%s_d%d_seed%d" % (tag, depth, seed)], flip_rel=2e-5)
            assert r < 1e-2 and nbad <= 60 and r_ex < 2e-5, (tag, depth, r, nbad, r_ex)
    g = np.load(GOLDEN + "/env_renderD_128_s4_d3_smallbox.npz")
    integ = psdr.PathTracer(3)
    integ.reference_tangent_scaling = True
    sc = build_product(scenes.cbox_meshes(), 128, 128, 4, 4, 4, move_mesh=1, axis_scale=(0.0, 30.0, 50.0), bsdfs=scenes.CBOX_MF_BSDFS, envmap=env)
    img, dimg = integ.renderD_fwd(sc, 0, seed=5)
    r, nbad, r_ex = compare_stats(dimg.cpu().numpy(), g["grad_all"], flip_rel=2e-5)
    assert nbad <= 0.08 * len(dimg) and r_ex < 1e-3, (r, nbad, r_ex)


# ---- textured reflectance: Bitmap3fD::eval (reference src/core/bitmap.cpp:46-131) --------------------------
def _textures(with_tangent=False, seed=4):
    rng = np.random.default_rng(seed)
    out = {}
    for name, (w, h) in (("white", (8, 6)), ("cat", (5, 7))):
        data = rng.random((h * w, 3), dtype=np.float32) * 0.8 + 0.1
        out[name] = (data, w, h, rng.normal(size=(h * w, 3)).astype(np.float32) * 0.2 if with_tangent else None)
    return out


@pytest.mark.parametrize("bs", ["diffuse", "mf"])
def test_textured_bsdf_renderC_vs_oracle(oracle, bs):
    psdr = _psdr()
    bsdfs = scenes.CBOX_MF_BSDFS if bs == "mf" else None
    ref = build_oracle(scenes.cbox_meshes(), 128, 128, 4, 0, 0, bsdfs=bsdfs, textures=_textures()).render(3, seed=3, mode=0)
    plain = build_oracle(scenes.cbox_meshes(), 128, 128, 4, 0, 0, bsdfs=bsdfs).render(3, seed=3, mode=0)
    sc = build_product(scenes.cbox_meshes(), 128, 128, 4, 0, 0, bsdfs=bsdfs, textures=_textures())
    got = psdr.PathTracer(3).renderC(sc, 0, seed=3).cpu().numpy()
    assert rel_l2(got, ref) < TOL
    assert rel_l2(ref, plain) > 1e-2           # the textures are visible


@pytest.mark.parametrize("terms", [1, 7])
def test_textured_bsdf_renderD_vs_oracle(oracle, terms):
    """tangents: texels, camera translation (moves uv at the primary hit), large-box translation"""
    psdr = _psdr()
    spps = (4 if terms & 1 else 0, 4 if terms & 2 else 0, 4 if terms & 4 else 0)
    from oracle.psdr_oracle import OracleScene
    tex = _textures(with_tangent=True)
    cam_t = translation_tangent((3.0, -2.0, 1.0))
    osc = OracleScene(128, 128, *spps)
    for name, refl in scenes.CBOX_BSDFS:
        osc.add_diffuse(name, refl)
        if name in tex:
            osc.set_bsdf_texture(name, *tex[name])
    for i, m in enumerate(scenes.cbox_meshes()):
        osc.add_mesh(m.v, m.f, m.bsdf, uv=m.uv, fuv=m.fuv, to_world={"raw": m.to_world}, radiance=m.emitter,
                     d_to_world={"left": translation_tangent((10.0, 0.0, 20.0))} if i == 2 else None)
    c = scenes.CBOX_CAMERA
    osc.add_camera(c["fov"], c["near"], c["far"], {"raw": c["to_world"]}, d_to_world={"left": cam_t})
    osc.configure((0,))
    img_ref, dimg_ref = osc.render(2, seed=6, mode=1, terms=7)
    sc = build_product(scenes.cbox_meshes(), 128, 128, *spps, move_mesh=2, axis_scale=(10.0, 0.0, 20.0), textures=tex, d_cam_left=cam_t)
    img, dimg = psdr.PathTracer(2).renderD_fwd(sc, 0, seed=6)
    assert rel_l2(img.cpu().numpy(), img_ref) < TOL
    assert np.abs(dimg_ref).max() > 0 and rel_l2(dimg.cpu().numpy(), dimg_ref) < TOL


# ---- all three bitmap slots of a MicrofacetBSDF + the bitmaps' uv transform (reference src/core/bitmap.cpp:64-72,
#      include/psdr/bsdf/microfacet.h:17,33-35) ------------------------------------------------------------------
def _slot_textures(with_tangent=False, with_xform=True, texel_tangents=True):
    """texel data from ONE stream (tools/ref_golden5.py draws the same); tangents from a second one"""
    rng, trng = np.random.default_rng(21), np.random.default_rng(22)
    out = {}
    for name, dims in (("cat", ((7, 5), (4, 6), (5, 5))), ("white", ((6, 4), (3, 3), (8, 2)))):
        slots = {}
        for slot, (w, h) in enumerate(dims):
            ch = 1 if slot == 2 else 3
            lo, hi = ((0.05, 0.9), (0.02, 0.6), (0.15, 0.7))[slot]
            t = dict(data=(rng.random((h * w, ch), dtype=np.float32) * (hi - lo) + lo), w=w, h=h)
            if with_tangent and texel_tangents:
                t["d_data"] = trng.normal(size=(h * w, ch)).astype(np.float32) * 0.1
            if with_xform:
                t["xform"] = np.array([1.0 + 0.4 * slot, 0.3 - 0.25 * slot, 0.11 * (slot + 1), -0.07 * slot], np.float32)
                if with_tangent:
                    t["d_xform"] = np.array([0.2, -0.3, 0.05, 0.1], np.float32) * (slot + 1)
            slots[slot] = t
        out[name] = slots
    return out


def test_texture_slots_and_uv_transform_renderC_vs_oracle(oracle):
    psdr = _psdr()
    tex = _slot_textures()
    ref = build_oracle(scenes.cbox_meshes(), 128, 128, 4, 0, 0, bsdfs=scenes.CBOX_MF_BSDFS, textures=tex).render(3, seed=5, mode=0)
    no_xf = build_oracle(scenes.cbox_meshes(), 128, 128, 4, 0, 0, bsdfs=scenes.CBOX_MF_BSDFS, textures=_slot_textures(with_xform=False)).render(3, seed=5, mode=0)
    sc = build_product(scenes.cbox_meshes(), 128, 128, 4, 0, 0, bsdfs=scenes.CBOX_MF_BSDFS, textures=tex)
    got = psdr.PathTracer(3).renderC(sc, 0, seed=5).cpu().numpy()
    assert rel_l2(got, ref) < TOL
    assert rel_l2(ref, no_xf) > 1e-3            # the uv transforms are visible


@pytest.mark.parametrize("terms", [1, 7])
def test_texture_slots_and_uv_transform_renderD_vs_oracle(oracle, terms):
    """tangents: texels of all three slots, the uv transforms (scale, rotate, translate), large-box translation"""
    psdr = _psdr()
    spps = (4 if terms & 1 else 0, 4 if terms & 2 else 0, 4 if terms & 4 else 0)
    tex = _slot_textures(with_tangent=True)
    kw = dict(move_mesh=2, axis_scale=(10.0, 0.0, 20.0), bsdfs=scenes.CBOX_MF_BSDFS, textures=tex)
    img_ref, dimg_ref = build_oracle(scenes.cbox_meshes(), 128, 128, *spps, **kw).render(2, seed=8, mode=1, terms=7)
    sc = build_product(scenes.cbox_meshes(), 128, 128, *spps, **kw)
    img, dimg = psdr.PathTracer(2).renderD_fwd(sc, 0, seed=8)
    assert rel_l2(img.cpu().numpy(), img_ref) < TOL
    assert np.abs(dimg_ref).max() > 0 and rel_l2(dimg.cpu().numpy(), dimg_ref) < TOL


def test_texture_slots_vs_reference_golden():
    """MicrofacetBSDF(Bitmap3fD, Bitmap3fD, Bitmap1fD) with uv transforms against the running reference
    (tools/ref_golden5.py -> tests/golden/tex_slots.npz): primal image, derivative w.r.t. the six uv transforms,
    derivative w.r.t. a translation of the textured large box."""
    psdr = _psdr()
    g = np.load(GOLDEN + "/tex_slots.npz")
    tex = _slot_textures()
    sc = build_product(scenes.cbox_meshes(), 128, 128, 4, 0, 0, bsdfs=scenes.CBOX_MF_BSDFS, textures=tex)
    got = psdr.PathTracer(3).renderC(sc, 0, seed=5).cpu().numpy()
    r, nbad, r_ex = compare_stats(got, g["img_d3_seed5"], flip_rel=1e-4)
    print("texture slots renderC vs reference: rel-L2 %.3e, %d pixels with a flipped lane, rest %.3e" % (r, nbad, r_ex))
    assert nbad < 0.03 * len(got) and r_ex < 2e-5, (r, nbad, r_ex)
    texd = _slot_textures(with_tangent=True, texel_tangents=False)
    integ = psdr.PathTracer(2)
    integ.reference_tangent_scaling = True
    sc = build_product(scenes.cbox_meshes(), 128, 128, 4, 0, 0, bsdfs=scenes.CBOX_MF_BSDFS, textures=texd)
    img, dimg = integ.renderD_fwd(sc, 0, seed=8)
    r, nbad, r_ex = compare_stats(dimg.cpu().numpy(), g["grad_uv"], flip_rel=1e-3)
    print("d/d(uv transforms) vs reference: rel-L2 %.3e, %d pixels off, rest %.3e" % (r, nbad, r_ex))
    assert nbad < 0.03 * len(got) and r_ex < 5e-3 and r < 6e-2, (r, nbad, r_ex)        # CPU oracle vs this golden: 2.6e-2, 105 px, 1.7e-3
    sc = build_product(scenes.cbox_meshes(), 128, 128, 4, 4, 4, bsdfs=scenes.CBOX_MF_BSDFS, textures=tex, move_mesh=2, axis_scale=(10.0, 0.0, 20.0))
    img, dimg = integ.renderD_fwd(sc, 0, seed=8)
    r, nbad, r_ex = compare_stats(dimg.cpu().numpy(), g["grad_box"], flip_rel=1e-3)
    print("d/d(box translation) vs reference: rel-L2 %.3e, %d pixels off, rest %.3e" % (r, nbad, r_ex))
    assert nbad < 0.03 * len(got) and r_ex < 1.5e-2 and r < 0.1, (r, nbad, r_ex)        # CPU oracle vs this golden: 5.2e-2, 129 px, 5.7e-3


# ---- Direct integrator (reference src/integrator/direct.cpp) and field extraction (src/integrator/field.cpp) ----------
@pytest.mark.parametrize("mis", [0, 1, 2])
def test_direct_integrator_vs_oracle(oracle, mis):
    psdr = _psdr()
    kw = dict(move_mesh=0, axis_scale=(100.0, 0.0, 0.0))
    osc = build_oracle(scenes.cbox_meshes(), 96, 96, 4, 4, 4, **kw)
    osc.set_mis(mis)
    ref_c = osc.render(1, seed=4, mode=0)
    ref_i, ref_d = osc.render(1, seed=4, mode=1, terms=7)
    sc = build_product(scenes.cbox_meshes(), 96, 96, 4, 4, 4, **kw)
    integ = psdr.Direct(mis)
    assert rel_l2(integ.renderC(sc, 0, seed=4).cpu().numpy(), ref_c) < TOL
    img, dimg = integ.renderD_fwd(sc, 0, seed=4)
    assert rel_l2(img.cpu().numpy(), ref_i) < TOL and rel_l2(dimg.cpu().numpy(), ref_d) < TOL
    # continuation of the streams consumes the mode's own number of draws
    img2, _ = integ.renderD_fwd(sc, 0, seed=-1)
    draws = {0: 2, 1: 3, 2: 5}[mis]
    ref2, _ = osc.render(1, seed=4, mode=1, terms=7, skip=[2 + draws, 1 + 2 * draws, 3])
    assert rel_l2(img2.cpu().numpy(), ref2) < TOL
    if mis == 2:      # Direct(2) is PathTracer(1)
        assert rel_l2(psdr.PathTracer(1).renderC(sc, 0, seed=4).cpu().numpy(), ref_c) < TOL


def test_direct_integrator_vs_reference_golden():
    """psdr.Direct(mis) against the running reference (tools/ref_golden6.py -> tests/golden/direct.npz).  The reference
    binary's Direct integrator returns exactly TWICE the radiance PathTracer(1) returns for the same samples (a directly
    visible (20, 20, 8) luminaire reads (40, 40, 16)); the factor is applied to the golden here, the product returns the
    radiance.  Its interior derivative image carries the same 2x as PathTracer's (reference_tangent_scaling)."""
    psdr = _psdr()
    g = np.load(GOLDEN + "/direct.npz")
    kw = dict(move_mesh=0, axis_scale=(100.0, 0.0, 0.0))
    assert np.allclose(g["imgC_mis2"].max(axis=0), [40.0, 40.0, 16.0])
    for mis in (0, 1, 2):
        sc = build_product(scenes.cbox_meshes(), 128, 128, 4, 0, 0, **kw)
        integ = psdr.Direct(mis)
        integ.reference_tangent_scaling = True
        r, nbad, r_ex = compare_stats(2.0 * integ.renderC(sc, 0, seed=0).cpu().numpy(), g["imgC_mis%d" % mis], flip_rel=2e-5)
        print("Direct(%d) renderC vs reference: rel-L2 %.3e, %d pixels with a flipped lane, rest %.3e" % (mis, r, nbad, r_ex))
        assert nbad < 0.02 * 128 * 128 and r_ex < 2e-5, (mis, r, nbad, r_ex)               # CPU oracle vs this golden: <= 108 pixels, 3e-6
        img, dimg = integ.renderD_fwd(sc, 0, seed=0, terms=1)
        r, nbad, r_ex = compare_stats(dimg.cpu().numpy(), g["gradD_int_mis%d" % mis], flip_rel=1e-3)
        print("Direct(%d) interior derivative vs reference: rel-L2 %.3e, %d pixels off, rest %.3e" % (mis, r, nbad, r_ex))
        assert nbad < 0.02 * 128 * 128 and r_ex < 2e-4, (mis, r, nbad, r_ex)               # CPU oracle: 0 / 5e-7 (mis 1), 5e-5 (mis 0, 2)


def test_field_extraction_integrator(oracle):
    psdr = _psdr()
    sc = build_product(scenes.cbox_meshes(), 64, 64, 4, 0, 0)
    aov = build_oracle(scenes.cbox_meshes(), 64, 64, 4, 0, 0).aov(0, seed=0).reshape(64 * 64, 4, 14)
    valid = (aov[:, :, 0] > 0)[..., None]
    for field, cols in (("position", slice(2, 5)), ("geoNormal", slice(6, 9)), ("shNormal", slice(9, 12))):
        got = psdr.FieldExtractionIntegrator(field).renderC(sc, 0, seed=0).cpu().numpy()
        assert rel_l2(got, (aov[:, :, cols] * valid).sum(1) / 4.0) < 1e-6, field
    depth = psdr.FieldExtractionIntegrator("depth").renderC(sc, 0, seed=0).cpu().numpy()
    assert rel_l2(depth[:, 0], (aov[:, :, 5] * valid[..., 0]).sum(1) / 4.0) < 1e-6
    sil = psdr.FieldExtractionIntegrator("silhouette 1").renderC(sc, 0, seed=0).cpu().numpy()       # the small box only
    assert rel_l2(sil[:, 0], (aov[:, :, 0] == 2.0).sum(1) / 4.0) < 1e-6 and 0 < sil.sum() < 3 * 64 * 64
    with pytest.raises(RuntimeError, match="Unsupported field"):
        psdr.FieldExtractionIntegrator("albedo")


def test_bvh_gpu_refit_matches_fresh_build(oracle):
    """Scenes above 64 triangles: the BVH topology is built once on the host, every later configure() refits boxes and
    leaf blocks on the GPU (device_upload.cu bvh_refit_kernel; the reference rebuilds its OptiX GAS per configure,
    scene_optix.cpp:254-333).  After a mesh moved, hit ids and images must equal those of a freshly built tree and of
    the brute-force oracle."""
    psdr = _psdr()
    from psdr_jit_b200 import _lib
    L = _lib.load()
    meshes = sphere_meshes()                                   # 36 + 320 triangles
    sc = build_product(meshes, 96, 96, 2, 0, 0)
    assert L.psdr_scene_query(sc._h, _lib.Q_USES_BVH, 0) == 1
    b0, r0 = L.psdr_scene_query(sc._h, _lib.Q_BVH_BUILDS, 0), L.psdr_scene_query(sc._h, _lib.Q_BVH_REFITS, 0)
    move = scenes.translate(60.0, -35.0, 40.0)
    sc.param_map["Mesh[8]"].set_transform(move)                # the sphere
    sc.param_map["Mesh[2]"].set_transform(scenes.translate(-20.0, 0.0, 15.0))
    sc.configure([0])
    assert L.psdr_scene_query(sc._h, _lib.Q_BVH_BUILDS, 0) == b0 and L.psdr_scene_query(sc._h, _lib.Q_BVH_REFITS, 0) == r0 + 1
    aov_refit = psdr.PathTracer(1).render_aov(sc, 0, seed=0).cpu().numpy()
    img_refit = psdr.PathTracer(3).renderC(sc, 0, seed=2).cpu().numpy()
    sc.set_accel(1)                                            # forces a fresh topology for the moved geometry
    sc.configure([0])
    assert L.psdr_scene_query(sc._h, _lib.Q_BVH_BUILDS, 0) == b0 + 1
    aov_fresh = psdr.PathTracer(1).render_aov(sc, 0, seed=0).cpu().numpy()
    img_fresh = psdr.PathTracer(3).renderC(sc, 0, seed=2).cpu().numpy()
    assert np.array_equal(aov_refit[:, :2], aov_fresh[:, :2])              # mesh and triangle ids, bit exact
    assert rel_l2(img_refit, img_fresh) < 1e-6
    import copy
    moved = copy.deepcopy(meshes)
    moved[8].to_world = move @ moved[8].to_world
    moved[2].to_world = scenes.translate(-20.0, 0.0, 15.0) @ moved[2].to_world
    osc = build_oracle(moved, 96, 96, 2, 0, 0)
    assert np.array_equal(aov_refit[:, 1], osc.aov(0, seed=0)[:, 1])       # triangle ids vs the brute-force oracle
    assert rel_l2(img_refit, osc.render(3, seed=2, mode=0)) < TOL


INTRINSIC_CAM = dict(scenes.CBOX_CAMERA, intrinsics=(0.9, 1.1, 0.45, 0.55))


def test_intrinsics_camera_vs_oracle():
    """PerspectiveCamera(fx, fy, cx, cy, near, far) (reference perspective.h:11-12, perspective.cpp:15-20): image and forward
    derivative image, all three terms, off-centre principal point and fx != fy"""
    import psdr_jit_b200 as psdr
    kw = dict(move_mesh=1, axis_scale=(30.0, 10.0, 0.0), cam=INTRINSIC_CAM)
    img_ref, dimg_ref = build_oracle(scenes.cbox_meshes(), 96, 96, 4, 4, 4, **kw).render(2, seed=5, mode=1, terms=7)
    sc = build_product(scenes.cbox_meshes(), 96, 96, 4, 4, 4, **kw)
    img, dimg = psdr.PathTracer(2).renderD_fwd(sc, 0, seed=5)
    assert rel_l2(img.cpu().numpy(), img_ref) < 1e-6
    assert np.abs(dimg_ref).max() > 0 and rel_l2(dimg.cpu().numpy(), dimg_ref) < 1e-4


def test_intrinsics_camera_vs_reference_golden():
    """the same camera against the RUNNING reference (tests/golden/intrinsics.npz, tools/ref_golden11.py), term by term"""
    import psdr_jit_b200 as psdr
    path = os.path.join(GOLDEN, "intrinsics.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/intrinsics.npz not generated yet")
    g = np.load(path)
    cam = dict(scenes.CBOX_CAMERA, intrinsics=tuple(float(x) for x in g["intrinsics"]))
    integ = psdr.PathTracer(2)
    integ.reference_tangent_scaling = True
    got = integ.renderC(build_product(scenes.cbox_meshes(), 128, 128, 4, 0, 0, cam=cam), 0, seed=0).cpu().numpy()
    r, nbad, r_ex = compare_stats(got, g["imgC"])
    assert nbad <= 16 and r_ex < 3e-4, (r, nbad, r_ex)      # measured: 4 flipped pixels, 1.2e-4 on the rest
    for tag, (spp, sppe, sppse), term in (("int", (4, 0, 0), 1), ("pri", (0, 4, 0), 2), ("sec", (0, 0, 4), 4)):
        sc = build_product(scenes.cbox_meshes(), 128, 128, spp, sppe, sppse, move_mesh=1, axis_scale=(30.0, 10.0, 0.0), cam=cam)
        dimg = integ.renderD_fwd(sc, 0, seed=0, terms=term)[1].cpu().numpy()
        r, nbad, r_ex = compare_stats(dimg, g["gradD_" + tag])
        assert nbad <= 256 and r_ex < 5e-3, (tag, r, nbad, r_ex)


def test_scaled_cfg2_vs_reference_golden_and_reference_nondeterminism():
    """The CUDA path on BASELINE config 2's scene at scale 1/300 against the running reference (tests/golden/scaled_cfg2.npz):
    image 3.7e-5 (no flipped pixel; north-star bar 1e-4), primary-edge derivative 7e-7, interior derivative 2e-6 outside one
    pixel.  The secondary-edge derivative of the reference is NOT reproducible between identical runs
    (tests/golden/ref_determinism_sec.npz, tools/ref_probe5.py: two runs of the reference differ by rel-L2 0.13 - 0.18, lost
    updates in its accumulation); ours is the complete sum -- above every run's total, equal to some run in most entries."""
    import psdr_jit_b200 as psdr
    g = np.load(os.path.join(GOLDEN, "scaled_cfg2.npz"))
    s, ax = float(g["scale"]), float(g["axis"])
    ms, cam = scenes.scaled_cbox(s)
    integ = psdr.PathTracer(3)
    integ.reference_tangent_scaling = True
    img = integ.renderC(build_product(ms, 128, 128, 4, 0, 0, cam=cam), 0, seed=0).cpu().numpy()
    assert rel_l2(img, g["imgC"]) < 1e-4
    for tag, (spp, sppe), term, tol in (("int", (4, 0), 1, 1e-5), ("pri", (0, 4), 2, 1e-5)):
        sc = build_product(ms, 128, 128, spp, sppe, 0, cam=cam, move_mesh=0, axis_scale=(ax, 0.0, 0.0))
        im, d = integ.renderD_fwd(sc, 0, seed=0, terms=term)
        assert rel_l2(im.cpu().numpy(), g["imgD_" + tag]) < 1e-4
        r, nbad, r_ex = compare_stats(d.cpu().numpy(), g["gradD_" + tag])
        assert nbad <= 2 and r_ex < tol, (tag, r, nbad, r_ex)
    D = np.load(os.path.join(GOLDEN, "ref_determinism_sec.npz"))
    for tag, (ms2, cam2), ax2 in (("scaled", (ms, cam), ax), ("full", (scenes.cbox_meshes(), scenes.CBOX_CAMERA), 100.0)):
        a = D[tag].astype(np.float64)
        sc = build_product(ms2, 128, 128, 0, 0, 4, cam=cam2, move_mesh=0, axis_scale=(ax2, 0.0, 0.0))
        d = integ.renderD_fwd(sc, 0, seed=0, terms=4)[1].cpu().numpy().astype(np.float64)
        assert min(rel_l2(a[i], a[j]) for i in range(4) for j in range(i)) > 0.08
        nz = np.abs(d) > 0
        tol = 2e-4 * np.abs(d).max()
        assert (np.abs(a - d[None]) < tol).any(axis=0)[nz].mean() > 0.7, tag
        assert (np.abs(d)[None] >= np.abs(a) - tol)[:, nz].mean() > 0.97, tag
        assert all(np.abs(d).sum() > np.abs(a[i]).sum() * 1.05 for i in range(4)), tag
