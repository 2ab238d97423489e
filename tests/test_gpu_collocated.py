"""CollocatedIntegrator (reference src/psdr.cpp:427-429, src/integrator/collocated.cpp): a point light at the camera.  CUDA vs
the oracle (image, forward derivative image incl. the primary-edge term and the intensity tangent), vs the running reference
(tests/golden/collocated.npz, tools/ref_golden12.py), and reverse mode against forward mode."""
import os

import numpy as np
import pytest

from tests.common import GOLDEN, build_oracle, build_product, compare_stats, rel_l2, scenes

pytestmark = pytest.mark.gpu
MF = [(n, ([0.2, 0.6, 0.8], [0.5, 0.4, 0.3], 0.4)) if n == "cat" else (n, p) for n, p in scenes.CBOX_BSDFS]


@pytest.mark.parametrize("bsdfs", [None, MF])
def test_renderC_and_renderD_vs_oracle(bsdfs):
    import psdr_jit_b200 as psdr
    kw = dict(move_mesh=1, axis_scale=(30.0, 10.0, 0.0), bsdfs=bsdfs)
    osc = build_oracle(scenes.cbox_meshes(), 96, 96, 4, 4, 4, **kw)
    osc.set_collocated(1e6, 2e5)
    img_ref, dimg_ref = osc.render(1, seed=5, mode=1, terms=7)
    imgc_ref = osc.render(1, seed=5, mode=0)
    sc = build_product(scenes.cbox_meshes(), 96, 96, 4, 4, 4, **kw)
    integ = psdr.CollocatedIntegrator(1e6)
    integ.d_m_intensity = np.float32(2e5)
    img, dimg = integ.renderD_fwd(sc, 0, seed=5)
    assert rel_l2(img.cpu().numpy(), img_ref) < 1e-5
    assert np.abs(dimg_ref).max() > 0 and rel_l2(dimg.cpu().numpy(), dimg_ref) < 1e-4
    assert rel_l2(integ.renderC(sc, 0, seed=5).cpu().numpy(), imgc_ref) < 1e-5


def test_sampler_streams_continue_without_path_draws():
    """seed = -1 continuation: Li draws nothing, so two calls consume 2 (jitter) and 1 (edge) numbers per lane each"""
    import psdr_jit_b200 as psdr
    sc = build_product(scenes.cbox_meshes(), 48, 48, 2, 2, 0, move_mesh=1, axis_scale=(30.0, 10.0, 0.0))
    integ = psdr.CollocatedIntegrator(5e5)
    a0 = integ.renderD_fwd(sc, 0, seed=3)
    a1 = integ.renderD_fwd(sc, 0, seed=-1)
    osc = build_oracle(scenes.cbox_meshes(), 48, 48, 2, 2, 0, move_mesh=1, axis_scale=(30.0, 10.0, 0.0))
    osc.set_collocated(5e5)
    b0 = osc.render(1, seed=3, mode=1, terms=3)
    b1 = osc.render(1, seed=3, mode=1, terms=3, skip=(2, 1, 0))       # the same streams, 2 / 1 draws further on
    for a, b in ((a0, b0), (a1, b1)):
        assert rel_l2(a[0].cpu().numpy(), b[0]) < 1e-5 and rel_l2(a[1].cpu().numpy(), b[1]) < 1e-4
    assert rel_l2(a1[0].cpu().numpy(), b0[0]) > 1e-3      # the second call really used other samples


def test_vs_reference_golden():
    """The running reference (tests/golden/collocated.npz).  Like its Direct and FieldExtraction integrators, the binary
    returns exactly 2x the value collocated.cpp computes (and 2x its interior derivative): with that factor image and
    interior derivative image agree to 1e-6.  Its primary-edge image has the same 91 non-zero pixels as ours but values a
    factor 0.2 - 17 apart and channel-dependent sums on a grey box in front of grey walls -- as for FieldExtraction
     its edge term is not consistent with finite differences; ours is (next test)."""
    import psdr_jit_b200 as psdr
    path = os.path.join(GOLDEN, "collocated.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/collocated.npz not generated yet (tools/ref_golden12.py)")
    g = np.load(path)
    integ = psdr.CollocatedIntegrator(float(g["intensity"]))
    sc = build_product(scenes.cbox_meshes(), 128, 128, 4, 0, 0)
    r, nbad, r_ex = compare_stats(2.0 * integ.renderC(sc, 0, seed=0).cpu().numpy(), g["imgC"])
    assert nbad <= 4 and r_ex < 1e-5, (r, nbad, r_ex)
    sc = build_product(scenes.cbox_meshes(), 128, 128, 4, 0, 0, move_mesh=1, axis_scale=(30.0, 10.0, 0.0))
    dimg = integ.renderD_fwd(sc, 0, seed=0, terms=1)[1].cpu().numpy()
    r, nbad, r_ex = compare_stats(2.0 * dimg, g["gradD_int"])
    assert nbad <= 8 and r_ex < 1e-5, (r, nbad, r_ex)
    sc = build_product(scenes.cbox_meshes(), 128, 128, 0, 4, 0, move_mesh=1, axis_scale=(30.0, 10.0, 0.0))
    dimg = integ.renderD_fwd(sc, 0, seed=0, terms=2)[1].cpu().numpy()
    assert ((np.abs(dimg).max(axis=1) > 0) == (np.abs(g["gradD_pri"]).max(axis=1) > 0)).all()      # the same edge samples land in the same pixels


def test_derivative_matches_finite_differences():
    """interior + primary-edge derivative, summed over the image, against the central difference of the summed image of
    the moved scene (the box slides along the floor).  The two terms nearly cancel (interior -7, edges +6 per unit of
    motion at 128^2), so the comparison is made on the scale of the edge term; the difference quotient of a Monte-Carlo
    coverage image carries a few per cent of that scale in noise itself (three seeds are averaged)."""
    import psdr_jit_b200 as psdr
    integ = psdr.CollocatedIntegrator(1e6)
    w, spp, h = 192, 64, 0.2

    def total(P, seed):
        ms = scenes.cbox_meshes()
        tw = ms[1].to_world.copy()
        tw[0, 3] += 30 * P
        tw[2, 3] += 10 * P
        ms[1].to_world = tw
        return float(integ.renderC(build_product(ms, w, w, spp, 0, 0), 0, seed=seed).sum())
    fd, est, interior = [], [], []
    for seed in (1, 2, 3):
        fd.append((total(h, seed) - total(-h, seed)) / (2 * h))
        sc = build_product(scenes.cbox_meshes(), w, w, spp, spp, 0, move_mesh=1, axis_scale=(30.0, 0.0, 10.0))
        est.append(float(integ.renderD_fwd(sc, 0, seed=seed)[1].sum()))
        interior.append(float(integ.renderD_fwd(sc, 0, seed=seed, terms=1)[1].sum()))
    fd, est, interior = np.mean(fd), np.mean(est), np.mean(interior)
    edge = est - interior
    assert abs(edge) > 0.5 * abs(interior)                 # the edge term matters here
    assert abs(est - fd) < 0.08 * abs(edge), (est, interior, fd)


@pytest.mark.parametrize("bsdfs", [None, MF])
def test_vjp_is_transpose(bsdfs):
    """<cot, J t> == <J^T cot, t> for t = (the small box's and the camera's translation, a reflectance, the intensity)"""
    import torch
    import psdr_jit_b200 as psdr
    rng = np.random.default_rng(3)
    w = h = 64
    sc = build_product(scenes.cbox_meshes(), w, h, 8, 8, 0, bsdfs=bsdfs)
    t = np.zeros((4, 4), np.float32)
    t[:3, 3] = rng.normal(size=3) * 30
    tc = np.zeros((4, 4), np.float32)
    tc[:3, 3] = rng.normal(size=3) * 10
    sc.param_map["Mesh[1]"].d_to_world_left = t.copy()
    sc.param_map["Sensor[0]"].d_to_world_left = tc.copy()
    tang = {("Mesh[1]", "to_world_left"): t, ("Sensor[0]", "to_world_left"): tc}
    if bsdfs is None:
        sc.param_map["BSDF[id=white]"].d_reflectance = np.float32([0.3, -0.2, 0.5])
        tang[("BSDF[id=white]", "reflectance")] = np.float32([0.3, -0.2, 0.5])
    else:
        sc.param_map["BSDF[id=cat]"].d_roughness = np.float32(0.7)
        tang[("BSDF[id=cat]", "roughness")] = np.float32([0.7])
    sc.configure()
    sc.configure([0])
    integ = psdr.CollocatedIntegrator(8e5)
    integ.d_m_intensity = np.float32(1e5)
    img, dimg = integ.renderD_fwd(sc, 0, seed=3)
    cot = torch.as_tensor(rng.normal(size=(w * h, 3)).astype(np.float32), device=img.device)
    lhs = float((cot.double() * dimg.double()).sum())
    integ.render_vjp(sc, cot, 0, seed=3)
    parts = {k: float((sc.grad_of(*k).reshape(np.shape(v)).astype(np.float64) * v.astype(np.float64)).sum()) for k, v in tang.items()}
    parts["intensity"] = integ.grad_intensity(sc) * 1e5
    rhs = sum(parts.values())
    mag = max(abs(lhs), sum(abs(v) for v in parts.values()))
    assert mag > 0 and abs(lhs - rhs) < 5e-4 * mag, (lhs, rhs, parts)


def test_autograd_intensity_leaf():
    import torch
    import psdr_jit_b200 as psdr
    sc = build_product(scenes.cbox_meshes(), 32, 32, 4, 0, 0)
    inten = torch.tensor(5e5, requires_grad=True)
    integ = psdr.CollocatedIntegrator(inten)
    img = integ.renderD(sc, 0, seed=1)
    img.sum().backward()
    assert inten.grad is not None and abs(float(inten.grad) - float(img.sum()) / 5e5) < 1e-4 * abs(float(inten.grad))


def test_bsdf_field_vs_oracle():
    """FieldExtractionIntegrator("bsdf") (reference src/integrator/field.cpp:72-92): BSDF(wi, wi) at the primary hit"""
    import psdr_jit_b200 as psdr
    kw = dict(move_mesh=1, axis_scale=(30.0, 10.0, 0.0), bsdfs=MF)
    osc = build_oracle(scenes.cbox_meshes(), 96, 96, 4, 4, 0, **kw)
    osc.set_collocated(1.0, 0.0, bsdf_field=True)
    img_ref, dimg_ref = osc.render(1, seed=5, mode=1, terms=3)
    sc = build_product(scenes.cbox_meshes(), 96, 96, 4, 4, 0, **kw)
    integ = psdr.FieldExtractionIntegrator("bsdf")
    img = integ.renderD(sc, 0, seed=5)
    assert rel_l2(img.cpu().numpy(), img_ref) < 1e-5
    assert np.abs(dimg_ref).max() > 0 and rel_l2(integ.grad_image.cpu().numpy(), dimg_ref) < 1e-4
    assert float(img.max()) < 2.0          # a BSDF value, not a radiance
