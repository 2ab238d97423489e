"""Every term kernel exists in two CTA shapes (include/psdr_b200.h psdr_set_cta_policy): 128-thread CTAs for small
launches, and large CTAs with block barriers.  The small test scenes would only ever launch the first, so the parity
tests of this file force each shape in turn: both must reproduce the oracle, and each other."""
import numpy as np
import pytest

from tests.common import build_oracle, build_product, rel_l2, scenes, sphere_meshes

pytestmark = pytest.mark.gpu
TOL = 1e-4      # as tests/test_gpu_parity.py


@pytest.fixture
def policy():
    import psdr_jit_b200 as psdr
    yield psdr.set_cta_policy
    psdr.set_cta_policy(0)


@pytest.mark.parametrize("accel", [0, 1])
def test_forward_images_of_both_shapes_match_oracle(policy, accel):
    import psdr_jit_b200 as psdr
    kw = dict(move_mesh=0, axis_scale=(100.0, 0.0, 0.0))
    meshes = sphere_meshes() if accel else scenes.cbox_meshes()
    w, h, spp = 40, 36, 3          # lanes not a multiple of any CTA size: dead lanes ride through the block barriers
    orc = build_oracle(meshes, w, h, spp, spp, spp, **kw)
    ref_img, ref_dimg = orc.render(3, seed=5, mode=1, terms=7)
    ref_c = build_oracle(meshes, w, h, spp, 0, 0).render(3, seed=5, mode=0)
    sc = build_product(meshes, w, h, spp, spp, spp, accel=accel, **kw)
    integ = psdr.PathTracer(3)
    out = {}
    for p in (1, 2):
        policy(p)
        img, dimg = integ.renderD_fwd(sc, 0, seed=5)
        out[p] = (img.cpu().numpy(), dimg.cpu().numpy())
        assert rel_l2(out[p][0], ref_img) < TOL and rel_l2(out[p][1], ref_dimg) < TOL, p
        assert rel_l2(integ.renderC(sc, 0, seed=5).cpu().numpy(), ref_c) < TOL, p
    assert np.array_equal(out[1][0], out[2][0])                    # primal image: one deterministic writer per pixel
    assert rel_l2(out[1][1], out[2][1]) < 1e-5                     # derivative image: float atomics, order differs


def test_gradient_tables_of_both_shapes_match(policy):
    import torch
    import psdr_jit_b200 as psdr
    kw = dict(move_mesh=0, axis_scale=(100.0, 0.0, 0.0))
    sc = build_product(scenes.cbox_meshes(), 40, 36, 3, 3, 3, **kw)
    rng = np.random.default_rng(2)
    cot = torch.as_tensor(rng.normal(size=(40 * 36, 3)).astype(np.float32), device="cuda")
    integ = psdr.PathTracer(3)
    tabs = {}
    for p in (1, 2):
        policy(p)
        tabs[p] = integ.render_vjp_table(sc, cot, 0, seed=7).cpu().numpy().astype(np.float64)
    scale = np.abs(tabs[1]).max()
    assert scale > 0 and np.abs(tabs[1] - tabs[2]).max() < 2e-4 * scale
    # and the large shape is the transpose of forward mode (the small shape is covered by test_gpu_adjoint.py)
    policy(2)
    _, dimg = integ.renderD_fwd(sc, 0, seed=7)
    lhs = float((dimg.double() * cot.double()).sum())
    integ.render_vjp(sc, cot, 0, seed=7)
    g = sc.grad_of("Mesh[0]", "to_world_left").astype(np.float64)
    rhs = float(g[0, 3] * 100.0)
    assert abs(lhs) > 0 and abs(lhs - rhs) < 5e-4 * abs(lhs), (lhs, rhs)
