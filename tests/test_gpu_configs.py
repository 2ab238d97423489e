"""BASELINE configs 3, 4, 5 at their FULL sizes through size-independent properties (the oracle cannot run them in
seconds): finite output, determinism, independence of the CTA shape and of the number of lane shards (the sum of the
shards' partial images is the image), linearity in the emitter radiance, per-sensor independence of the batch.
Their kernels' parity against the oracle / reference at small sizes is in test_gpu_parity.py."""
import numpy as np
import pytest

from tests.common import rel_l2

pytestmark = pytest.mark.gpu


def _render(psdr, wl, integ, sc, sensor=0, seed=0):
    if wl["mode"] == "renderC":
        return integ.renderC(sc, sensor, seed=seed), None
    return integ.renderD_fwd(sc, sensor, seed=seed)


@pytest.mark.parametrize("cfg", [3, 4])
def test_full_size_properties(cfg):
    import torch
    import psdr_jit_b200 as psdr
    from psdr_jit_b200 import bench_scenes
    wl = bench_scenes.workload(cfg)
    sc = bench_scenes.build_ours(psdr, wl)
    integ = psdr.PathTracer(wl["depth"])
    if wl["guiding"]:
        integ.preprocess_secondary_edges(sc, 0, wl["guiding"], 1)
    img, dimg = _render(psdr, wl, integ, sc)
    assert torch.isfinite(img).all() and torch.isfinite(dimg).all()
    assert float(img.abs().max()) > 0 and float(dimg.abs().max()) > 0
    # determinism up to the order of float atomics (a pixel's spp lanes span several warps above spp 32)
    img2, dimg2 = _render(psdr, wl, integ, sc)
    assert rel_l2(img2.cpu().numpy(), img.cpu().numpy()) < 1e-6
    assert rel_l2(dimg2.cpu().numpy(), dimg.cpu().numpy()) < 1e-4
    # the CTA shape does not change a lane's value
    try:
        psdr.set_cta_policy(1)
        img3, dimg3 = _render(psdr, wl, integ, sc)
    finally:
        psdr.set_cta_policy(0)
    assert rel_l2(img3.cpu().numpy(), img.cpu().numpy()) < 1e-6
    assert rel_l2(dimg3.cpu().numpy(), dimg.cpu().numpy()) < 1e-4
    # lane shards: the partial images of 3 ranks add up to the image (what the multi-GPU reduction computes)
    acc = torch.zeros_like(integ.last_buffer)
    for r in range(3):
        part = bench_scenes.build_ours(psdr, wl, r, 3)
        if wl["guiding"]:
            integ.preprocess_secondary_edges(part, 0, wl["guiding"], 1)
        _render(psdr, wl, integ, part)
        acc += integ.last_buffer
    assert rel_l2(acc[0].cpu().numpy(), img.cpu().numpy()) < 1e-6
    assert rel_l2(acc[1].cpu().numpy(), dimg.cpu().numpy()) < 1e-4


def test_cfg3_linear_in_environment_scale():
    """cfg 3 without its area light: the image is linear in EnvironmentMap.scale, so d img / d scale * scale == img."""
    import psdr_jit_b200 as psdr
    from psdr_jit_b200 import bench_scenes
    wl = bench_scenes.workload(3, scale=1.0 / 16)
    wl["meshes"] = [m for m in wl["meshes"] if m.emitter is None]
    wl["bsdfs"] = [b for b in wl["bsdfs"] if b[0] != "light"]
    wl["moving"] = 0
    sc = bench_scenes.build_ours(psdr, wl)
    sc.param_map["Mesh[0]"].set_transform(np.eye(4, dtype=np.float32), tangent=np.zeros((4, 4), np.float32))
    env = sc.param_map["Emitter[0]"]
    env.scale, env.d_scale = np.float32(1.7), np.float32(1.0)
    sc.configure()
    sc.configure([0])
    img, dimg = psdr.PathTracer(wl["depth"]).renderD_fwd(sc, 0, seed=1, terms=1)
    assert float(img.abs().max()) > 0
    assert rel_l2((dimg * 1.7).cpu().numpy(), img.cpu().numpy()) < 1e-5


def test_cfg5_sensors_are_independent():
    """batch_render (cfg 5): rendering sensor k of the 8-sensor scene equals rendering a scene that holds only sensor k."""
    import torch
    import psdr_jit_b200 as psdr
    from psdr_jit_b200 import bench_scenes
    wl = bench_scenes.workload(5)
    sc = bench_scenes.build_ours(psdr, wl)
    integ = psdr.PathTracer(wl["depth"])
    for k in (0, 3, 7):
        img, dimg = integ.renderD_fwd(sc, k, seed=5)
        assert torch.isfinite(img).all() and torch.isfinite(dimg).all() and float(img.abs().max()) > 0
        one = dict(wl)
        one["cams"], one["sensors"] = [wl["cams"][k]], [0]
        sc1 = bench_scenes.build_ours(psdr, one)
        img1, dimg1 = integ.renderD_fwd(sc1, 0, seed=5)
        assert torch.equal(img, img1)
        assert rel_l2(dimg.cpu().numpy(), dimg1.cpu().numpy()) < 1e-6
