"""Primary-edge launches of 32768 lanes or more walk their samples bucketed by position along the edge list
(include/psdr_b200.h psdr_set_edge_sort, csrc/edge_sort.cu).  Which thread evaluates a sample must not change anything but
the order of the float atomics: every bucket count reproduces the unsorted launch and the oracle, for fresh and continued
sampler streams, both CTA shapes, lane shards and the adjoint kernel."""
import numpy as np
import pytest

from tests.common import build_oracle, build_product, rel_l2, scenes

pytestmark = pytest.mark.gpu
KW = dict(move_mesh=0, axis_scale=(100.0, 0.0, 0.0))
W, H, SPPE = 97, 83, 5          # 40 255 lanes: above the threshold, not a multiple of 32 / any CTA size / the sort's chunk


@pytest.fixture
def knobs():
    import psdr_jit_b200 as psdr
    yield psdr
    psdr.set_edge_sort(512)
    psdr.set_cta_policy(0)


def test_primary_edge_image_does_not_depend_on_the_lane_order(knobs):
    psdr = knobs
    orc = build_oracle(scenes.cbox_meshes(), W, H, 1, SPPE, 0, **KW)
    _, ref = orc.render(3, seed=5, mode=1, terms=2)
    integ = psdr.PathTracer(3)
    out = {}
    for bins in (0, 2, 512, 2048):
        for policy in (1, 2):
            psdr.set_edge_sort(bins)
            psdr.set_cta_policy(policy)
            sc = build_product(scenes.cbox_meshes(), W, H, 1, SPPE, 0, **KW)
            first = integ.renderD_fwd(sc, 0, seed=5, terms=2)[1].cpu().numpy()
            second = integ.renderD_fwd(sc, 0, seed=-1, terms=2)[1].cpu().numpy()      # continued streams: rp.skip != 0
            out[(bins, policy)] = (first, second)
            assert rel_l2(first, ref) < 1e-4, (bins, policy)
    base = out[(0, 1)]
    assert np.abs(base[0]).max() > 0 and rel_l2(base[0], base[1]) > 0.1      # the continuation is a different sample set
    for k, v in out.items():
        assert rel_l2(v[0], base[0]) < 1e-5 and rel_l2(v[1], base[1]) < 1e-5, k


def test_sorted_lane_shards_add_up(knobs):
    psdr = knobs
    integ = psdr.PathTracer(3)
    psdr.set_edge_sort(0)
    whole = integ.renderD_fwd(build_product(scenes.cbox_meshes(), 128, 128, 1, 8, 0, **KW), 0, seed=9, terms=2)[1].cpu().numpy()
    psdr.set_edge_sort(256)
    parts = 0
    for r in range(3):
        sc = build_product(scenes.cbox_meshes(), 128, 128, 1, 8, 0, **KW)
        sc.set_shard(r, 3)
        parts = parts + integ.renderD_fwd(sc, 0, seed=9, terms=2)[1].cpu().numpy().astype(np.float64)
    assert rel_l2(parts, whole) < 1e-5


def test_primary_edge_gradient_table_does_not_depend_on_the_lane_order(knobs):
    import torch
    psdr = knobs
    rng = np.random.default_rng(4)
    cot = torch.as_tensor(rng.normal(size=(W * H, 3)).astype(np.float32), device="cuda")
    integ = psdr.PathTracer(3)
    tabs = {}
    for bins in (0, 512):
        psdr.set_edge_sort(bins)
        sc = build_product(scenes.cbox_meshes(), W, H, 1, SPPE, 0, **KW)
        tabs[bins] = integ.render_vjp_table(sc, cot, 0, seed=7, terms=2).cpu().numpy().astype(np.float64)
    scale = np.abs(tabs[0]).max()
    assert scale > 0 and np.abs(tabs[0] - tabs[512]).max() < 2e-4 * scale


def test_bucket_count_is_validated(knobs):
    psdr = knobs
    for bad in (-1, 1, 2049):
        with pytest.raises(RuntimeError):
            psdr.set_edge_sort(bad)


def test_secondary_edge_image_does_not_depend_on_the_lane_order(knobs):
    """the secondary-edge kernels are ordered by the sample dimension that selects the edge; ordered samples are dealt to the
    warps in blocks of 256 (csrc/device_path.cuh sec_edge_batches)"""
    psdr = knobs
    orc = build_oracle(scenes.cbox_meshes(), W, H, 1, 0, SPPE, **KW)
    _, ref = orc.render(3, seed=5, mode=1, terms=4)
    integ = psdr.PathTracer(3)
    out = {}
    for bins in (0, 2, 512, 2048):
        for policy in (1, 2):
            psdr.set_edge_sort(bins)
            psdr.set_cta_policy(policy)
            sc = build_product(scenes.cbox_meshes(), W, H, 1, 0, SPPE, **KW)
            first = integ.renderD_fwd(sc, 0, seed=5, terms=4)[1].cpu().numpy()
            second = integ.renderD_fwd(sc, 0, seed=-1, terms=4)[1].cpu().numpy()
            out[(bins, policy)] = (first, second)
            assert rel_l2(first, ref) < 1e-4, (bins, policy)
    base = out[(0, 1)]
    assert np.abs(base[0]).max() > 0 and rel_l2(base[0], base[1]) > 0.1
    for k, v in out.items():
        assert rel_l2(v[0], base[0]) < 1e-5 and rel_l2(v[1], base[1]) < 1e-5, k


def test_guided_secondary_edges_do_not_depend_on_the_lane_order(knobs):
    """with a guiding grid the key is the sample AFTER HyperCubeDistribution<3>::sample_reuse has warped it"""
    from tests.common import sphere_meshes
    psdr = knobs
    kw = dict(move_mesh=8, axis_scale=(40.0, 20.0, 0.0))
    out = {}
    for bins in (0, 512):
        psdr.set_edge_sort(bins)
        sc = build_product(sphere_meshes(), 128, 128, 0, 0, 8, **kw)
        integ = psdr.PathTracer(2)
        integ.preprocess_secondary_edges(sc, 0, [8, 4, 4, 2], 2, 3)
        out[bins] = integ.renderD_fwd(sc, 0, seed=1)[1].cpu().numpy()
    assert np.abs(out[0]).max() > 0 and rel_l2(out[512], out[0]) < 1e-5


def test_secondary_edge_gradient_table_and_shards_do_not_depend_on_the_lane_order(knobs):
    import torch
    psdr = knobs
    rng = np.random.default_rng(6)
    cot = torch.as_tensor(rng.normal(size=(W * H, 3)).astype(np.float32), device="cuda")
    integ = psdr.PathTracer(3)
    tabs = {}
    for bins in (0, 512):
        psdr.set_edge_sort(bins)
        sc = build_product(scenes.cbox_meshes(), W, H, 1, 0, SPPE, **KW)
        tabs[bins] = integ.render_vjp_table(sc, cot, 0, seed=7, terms=4).cpu().numpy().astype(np.float64)
    scale = np.abs(tabs[0]).max()
    assert scale > 0 and np.abs(tabs[0] - tabs[512]).max() < 2e-4 * scale
    psdr.set_edge_sort(0)
    whole = integ.renderD_fwd(build_product(scenes.cbox_meshes(), 128, 128, 1, 0, 8, **KW), 0, seed=9, terms=4)[1].cpu().numpy()
    psdr.set_edge_sort(256)
    parts = 0
    for r in range(3):
        sc = build_product(scenes.cbox_meshes(), 128, 128, 1, 0, 8, **KW)
        sc.set_shard(r, 3)
        parts = parts + integ.renderD_fwd(sc, 0, seed=9, terms=4)[1].cpu().numpy().astype(np.float64)
    assert rel_l2(parts, whole) < 1e-5
