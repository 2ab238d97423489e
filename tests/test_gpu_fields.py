"""FieldExtractionIntegrator.renderD in forward mode: CUDA vs the oracle, and the oracle-independent checks."""
import numpy as np
import pytest

from tests.common import build_oracle, build_product, rel_l2, scenes, sphere_meshes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("field", ["depth", "position", "geoNormal", "shNormal", "uv", "silhouette", "segmentation"])
@pytest.mark.parametrize("scene,accel", [("open", 0), ("sphere", 1)])
def test_field_renderD_vs_oracle(field, scene, accel):
    import psdr_jit_b200 as psdr
    meshes = scenes.cbox_meshes()[:3] if scene == "open" else sphere_meshes()
    kw = dict(move_mesh=1, axis_scale=(30.0, 10.0, 0.0))
    osc = build_oracle(meshes, 64, 64, 4, 4, 0, **kw)
    ref_img, ref_d = osc.field_render_d(field, seed=3)
    sc = build_product(meshes, 64, 64, 4, 4, 0, accel=accel, **kw)
    integ = psdr.FieldExtractionIntegrator(field)
    img = integ.renderD(sc, 0, seed=3)
    dimg = integ.grad_image
    assert rel_l2(img.cpu().numpy(), ref_img) < 1e-5
    if np.abs(ref_d).max() > 0:
        assert rel_l2(dimg.cpu().numpy(), ref_d) < 1e-4
    else:
        assert float(dimg.abs().max()) == 0.0
    if field in ("depth", "position", "silhouette") and scene == "open":
        assert np.abs(ref_d).max() > 0          # the moving box has a silhouette against the void
    # one object only
    if field == "depth":
        ref_img1, ref_d1 = osc.field_render_d(field, seed=3, obj=1)
        integ1 = psdr.FieldExtractionIntegrator("depth 1")
        img1 = integ1.renderD(sc, 0, seed=3)
        assert rel_l2(img1.cpu().numpy(), ref_img1) < 1e-5 and rel_l2(integ1.grad_image.cpu().numpy(), ref_d1) < 1e-4


def test_field_vs_reference_golden_and_finite_differences():
    """tests/golden/fields.npz (the RUNNING reference, tools/ref_golden9.py): field images and interior derivative images
    are exactly 2x ours (the reference's Direct integrator shows the same factor); the primary-edge part is checked against
    finite differences of the covered area instead (the reference's own edge term is ~1/6 of it, tests/test_cpu_oracle.py)."""
    import os
    import psdr_jit_b200 as psdr
    from tests.common import GOLDEN, compare_stats
    g = np.load(os.path.join(GOLDEN, "fields.npz"))
    kw = dict(move_mesh=1, axis_scale=(30.0, 10.0, 0.0))
    for tag, meshes in (("box", scenes.cbox_meshes()), ("open", scenes.cbox_meshes()[:3])):
        sc = build_product(meshes, 128, 128, 4, 0, 0, **kw)
        for f in ("depth", "position", "shNormal", "geoNormal", "silhouette", "uv"):
            integ = psdr.FieldExtractionIntegrator(f)
            img = integ.renderD(sc, 0, seed=0)
            r, nbad, r_ex = compare_stats(2.0 * img.cpu().numpy(), g["%s_%s_C" % (tag, f)], flip_rel=1e-4)
            assert nbad <= 4 and r_ex < 2e-6, (tag, f, r, nbad, r_ex)
            gi = g["%s_%s_G_int" % (tag, f)]
            if np.abs(gi).max() > 0:
                r, nbad, r_ex = compare_stats(2.0 * integ.grad_image.cpu().numpy(), gi, flip_rel=1e-4)
                assert nbad <= 8 and r_ex < 1e-5, (tag, f, r, nbad, r_ex)

    def area(P):
        ms = scenes.cbox_meshes()[:3]
        tw = ms[1].to_world.copy()
        tw[0, 3] += 30 * P
        tw[1, 3] += 10 * P
        ms[1].to_world = tw
        return float(psdr.FieldExtractionIntegrator("silhouette").renderC(build_product(ms, 256, 256, 64, 0, 0), 0, seed=1)[:, 0].sum())
    fd = (area(0.05) - area(-0.05)) / 0.1
    sc = build_product(scenes.cbox_meshes()[:3], 256, 256, 1, 64, 0, **kw)
    integ = psdr.FieldExtractionIntegrator("silhouette")
    integ.renderD(sc, 0, seed=1)
    est = float(integ.grad_image[:, 0].sum())
    assert abs(est - fd) < 0.04 * abs(fd), (est, fd)      # (the finite difference of a 64-spp coverage image carries ~2 % noise itself)


@pytest.mark.parametrize("field", ["silhouette", "depth", "position", "shNormal", "geoNormal", "uv", "silhouette 1"])
def test_field_vjp_is_transpose_of_forward_mode(field):
    """reverse mode of FieldExtractionIntegrator (psdr_render_field_vjp): <cot, J t> == <J^T cot, t> for t = translations /
    a small rotation of two meshes, a vertex field and the camera; interior + primary-edge parts"""
    import torch
    import psdr_jit_b200 as psdr
    rng = np.random.default_rng(11)
    w = h = 64
    sc = build_product(scenes.cbox_meshes(), w, h, 8, 8, 0)
    tang = {}
    t = np.zeros((4, 4), np.float32)
    t[:3, 3] = rng.normal(size=3) * 30
    t[:3, :3] = rng.normal(size=(3, 3)) * 0.05
    for name in ("Mesh[1]", "Mesh[2]"):
        sc.param_map[name].d_to_world_left = t.copy()
        tang[(name, "to_world_left")] = t.copy()
    m = sc.param_map["Mesh[2]"]
    tv = (rng.normal(size=m.vertex_positions.shape) * 3).astype(np.float32)
    m.d_vertex_positions = tv
    tang[("Mesh[2]", "vertex_positions")] = tv
    tc = np.zeros((4, 4), np.float32)
    tc[:3, 3] = rng.normal(size=3) * 5
    sc.param_map["Sensor[0]"].d_to_world_left = tc
    tang[("Sensor[0]", "to_world_left")] = tc
    sc.configure()
    sc.configure([0])
    integ = psdr.FieldExtractionIntegrator(field)
    img, dimg = integ.renderD_fwd(sc, 0, seed=3)
    cot = torch.as_tensor(rng.normal(size=(w * h, 3)).astype(np.float32), device=img.device)
    lhs = float((cot.double() * dimg.double()).sum())
    integ.render_vjp(sc, cot, 0, seed=3)
    parts = {k: float((sc.grad_of(*k).reshape(np.shape(v)).astype(np.float64) * v.astype(np.float64)).sum()) for k, v in tang.items()}
    rhs = sum(parts.values())
    mag = max(abs(lhs), sum(abs(v) for v in parts.values()))
    assert mag > 0 and abs(lhs - rhs) < 5e-4 * mag, (field, lhs, rhs, parts)


def test_silhouette_loss_gradient_through_autograd():
    """a mask loss in an optimisation loop: d/dP ||silhouette(P) - silhouette(P*)||^2 for a box translated by P along x has
    the sign that moves P towards P*, on both sides of it, and vanishes at P*"""
    import torch
    import psdr_jit_b200 as psdr
    integ = psdr.FieldExtractionIntegrator("silhouette 1")

    def scene_at(p):
        sc = build_product(scenes.cbox_meshes(), 96, 96, 16, 16, 0)
        T = torch.eye(4)
        T[0, 3] = p
        return sc, T

    sc, T = scene_at(0.0)
    sc.param_map["Mesh[1]"].set_transform(T.numpy())
    sc.configure([0])
    target = integ.renderC(sc, 0, seed=1).clone()
    grads = {}
    for p0 in (-25.0, 0.0, 25.0):
        sc, _ = scene_at(p0)
        p = torch.tensor(p0, requires_grad=True)
        M = torch.eye(4)
        M[0, 3] = p                      # (in place on a fresh tensor: M joins p's graph)
        sc.param_map["Mesh[1]"].set_transform(M)
        sc.configure([0])
        img = integ.renderD(sc, 0, seed=1)
        loss = ((img - target) ** 2).sum()
        loss.backward()
        grads[p0] = float(p.grad)
    assert grads[-25.0] < 0 < grads[25.0], grads            # descent moves P back to 0 from either side
    assert abs(grads[0.0]) < 0.05 * min(abs(grads[-25.0]), abs(grads[25.0])), grads
