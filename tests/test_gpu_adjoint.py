"""GPU tests of the reverse pass (psdr_render_vjp): the adjoint kernels + host chain must be the exact
transpose of the forward-mode pass (which is pinned against the oracle and the reference):
<w, J v> == <J^T w, v> for random tangents v over every parameter kind and random cotangent images w."""
import numpy as np
import pytest

from tests.common import build_product, scenes, sphere_meshes

pytestmark = pytest.mark.gpu


def _scene_with_tangents(psdr, meshes, w, h, spps, rng, what):
    mf = "microfacet" in what
    envmap = None
    if "envmap" in what:
        from tests.common import test_envmap
        envmap = dict(data=test_envmap(32, 16), w=32, h=16, scale=1.5,
                      to_world=np.array([[0.8, 0, 0.6, 0], [0, 1, 0, 0], [-0.6, 0, 0.8, 0], [0, 0, 0, 1]], np.float32))
    textures = None
    if "textures" in what:
        trng = np.random.default_rng(9)
        textures = {n: (trng.random((hh * ww, 3), dtype=np.float32) * 0.8 + 0.1, ww, hh, trng.normal(size=(hh * ww, 3)).astype(np.float32) * 0.2)
                    for n, (ww, hh) in (("white", (8, 6)), ("cat", (5, 7)))}
    slot_tex = None
    if "slots" in what:        # all three bitmap slots of the Microfacet BSDFs textured, with uv transforms (no transform tangents)
        from tests.test_gpu_parity import _slot_textures
        slot_tex = _slot_textures(with_tangent=True)
        if "xform" not in what:      # "slots xform": the tangents of the bitmaps' uv transforms stay in
            for n in slot_tex:
                for k in slot_tex[n]:
                    slot_tex[n][k].pop("d_xform", None)
        textures, mf = slot_tex, True
    sc = build_product(meshes, w, h, *spps, bsdfs=scenes.CBOX_MF_BSDFS if mf else None, envmap=envmap, textures=textures)
    tang = {}
    if slot_tex is not None:
        for n in slot_tex:
            for k, field in ((0, "diffuseReflectance.data"), (1, "specularReflectance.data"), (2, "roughness.data")):
                tang[("BSDF[id=%s]" % n, field)] = slot_tex[n][k]["d_data"]
                dx = slot_tex[n][k].get("d_xform")
                if dx is not None:      # (scale, rotate, translate.x, translate.y) of the slot's bitmap
                    base = field[:-len(".data")]
                    tang[("BSDF[id=%s]" % n, base + ".scale")] = np.float32(dx[0:1])
                    tang[("BSDF[id=%s]" % n, base + ".rotate")] = np.float32(dx[1:2])
                    tang[("BSDF[id=%s]" % n, base + ".translate")] = np.float32(dx[2:4])
        what = [x for x in what if x not in ("slots", "xform")]
        textures = None
    if textures is not None:
        for n in textures:
            tang[("BSDF[id=%s]" % n, "diffuseReflectance.data" if mf else "reflectance.data")] = textures[n][3]
        what = [x for x in what if x != "textures"]
    if envmap is not None:
        e = sc.param_map["Emitter[0]"]
        e.radiance.d_data = rng.normal(size=(16 * 32, 3)).astype(np.float32)
        e.d_scale = np.float32(0.7)
        t = np.zeros((4, 4), np.float32)
        t[:3, :3] = rng.normal(size=(3, 3)) * 0.3
        e.d_to_world_left = t
        tang[("Emitter[0]", "radiance.data")] = e.radiance.d_data
        tang[("Emitter[0]", "scale")] = np.reshape(e.d_scale, (1,))
        tang[("Emitter[0]", "to_world_left")] = t
        what = [x for x in what if x != "envmap"]
    if mf:
        for name in (("BSDF[id=red]", "BSDF[id=green]") if slot_tex is not None else ("BSDF[id=cat]", "BSDF[id=white]", "BSDF[id=red]")):
            b = sc.param_map[name]
            b.d_specularReflectance = (rng.normal(size=3) * 0.1).astype(np.float32)
            b.d_diffuseReflectance = (rng.normal(size=3) * 0.1).astype(np.float32)
            b.d_roughness = np.float32(rng.normal() * 0.1)
            tang[(name, "specularReflectance")] = b.d_specularReflectance
            if not hasattr(b.diffuseReflectance, "resolution"):
                tang[(name, "diffuseReflectance")] = b.d_diffuseReflectance
            tang[(name, "roughness")] = np.reshape(b.d_roughness, (1,))
        what = [x for x in what if x != "materials"]
    if "mesh_left" in what:
        t = np.zeros((4, 4), np.float32)
        t[:3, 3] = rng.normal(size=3) * 30
        t[:3, :3] = rng.normal(size=(3, 3)) * 0.05
        for name in ("Mesh[0]", "Mesh[1]"):
            sc.param_map[name].d_to_world_left = t.copy()
            tang[(name, "to_world_left")] = t.copy()
    if "mesh_raw" in what:
        t = np.zeros((4, 4), np.float32)
        t[:3, 3] = rng.normal(size=3) * 10
        sc.param_map["Mesh[2]"].d_to_world = t
        tang[("Mesh[2]", "to_world")] = t
    if "vertices" in what:
        for name in ("Mesh[2]", "Mesh[0]", "Mesh[%d]" % (len(meshes) - 1)):
            m = sc.param_map[name]
            t = (rng.normal(size=m.vertex_positions.shape) * 3).astype(np.float32)
            m.d_vertex_positions = t
            tang[(name, "vertex_positions")] = t
    if "materials" in what:
        sc.param_map["BSDF[id=white]"].d_reflectance = np.array([0.3, -0.2, 0.1], np.float32)
        tang[("BSDF[id=white]", "reflectance")] = sc.param_map["BSDF[id=white]"].d_reflectance
        sc.param_map["BSDF[id=red]"].d_reflectance = np.array([0.1, 0.2, -0.3], np.float32)
        tang[("BSDF[id=red]", "reflectance")] = sc.param_map["BSDF[id=red]"].d_reflectance
        sc.param_map["Emitter[0]"].d_radiance = np.array([1.0, -2.0, 3.0], np.float32)
        tang[("Emitter[0]", "radiance")] = sc.param_map["Emitter[0]"].d_radiance
    if "camera" in what:
        t = np.zeros((4, 4), np.float32)
        t[:3, 3] = rng.normal(size=3) * 5
        sc.param_map["Sensor[0]"].d_to_world_left = t
        tang[("Sensor[0]", "to_world_left")] = t
    sc.configure()
    sc.configure([0])
    return sc, tang


CASES = [
    ("materials", 1, 3, "cbox"), ("mesh_left", 1, 3, "cbox"), ("vertices", 1, 2, "cbox"), ("camera", 1, 3, "cbox"),
    ("mesh_left", 2, 2, "cbox"), ("vertices", 2, 2, "sphere"), ("camera", 2, 2, "cbox"),
    ("mesh_left", 4, 2, "cbox"), ("vertices", 4, 2, "sphere"), ("camera", 4, 2, "cbox"), ("mesh_raw", 4, 2, "cbox"),
    ("mesh_left vertices materials camera mesh_raw", 7, 3, "sphere"),
    ("microfacet", 1, 3, "cbox"), ("microfacet mesh_left vertices camera", 1, 2, "cbox"),
    ("microfacet mesh_left vertices camera", 7, 2, "sphere"),
    ("textures", 1, 2, "cbox"), ("textures camera mesh_left", 7, 2, "cbox"), ("textures microfacet camera", 1, 3, "cbox"),
    ("slots", 1, 2, "cbox"), ("slots camera mesh_left", 7, 3, "cbox"), ("slots xform", 1, 2, "cbox"), ("slots xform camera mesh_left", 7, 2, "cbox"),
    ("envmap", 1, 2, "cbox"), ("envmap microfacet", 1, 3, "cbox"), ("envmap microfacet mesh_left vertices camera", 7, 3, "cbox"),
]


@pytest.mark.parametrize("what,terms,depth,scene", CASES)
def test_vjp_is_transpose_of_jvp(what, terms, depth, scene):
    import torch
    import psdr_jit_b200 as psdr
    import zlib
    rng = np.random.default_rng(zlib.crc32(("%s/%d" % (what, terms)).encode()))
    meshes = scenes.cbox_meshes() if scene == "cbox" else sphere_meshes()
    w = h = 64
    spps = (8 if terms & 1 else 0, 8 if terms & 2 else 0, 8 if terms & 4 else 0)
    sc, tang = _scene_with_tangents(psdr, meshes, w, h, spps, rng, what.split())
    integ = psdr.PathTracer(depth)
    img, dimg = integ.renderD_fwd(sc, 0, seed=3, terms=terms)
    cot = torch.as_tensor(rng.normal(size=(w * h, 3)).astype(np.float32), device=img.device)
    lhs = float((cot.double() * dimg.double()).sum())
    integ.render_vjp(sc, cot, 0, seed=3, terms=terms)
    rhs, parts = 0.0, {}
    for (name, field), t in tang.items():
        g = sc.grad_of(name, field).reshape(np.shape(t))
        parts[(name, field)] = float((g.astype(np.float64) * t.astype(np.float64)).sum())
        rhs += parts[(name, field)]
    scale = float(torch.linalg.norm(cot.double()) * torch.linalg.norm(dimg.double()))
    mag = max(abs(lhs), sum(abs(v) for v in parts.values()))      # the per-parameter terms may cancel
    assert scale > 0 and mag > 1e-6 * scale, (lhs, scale)
    tol = 5e-4 if ("microfacet" in what or "slots" in what) else 2e-4        # GGX roughness derivatives are spiky: more fp32 cancellation
    assert abs(lhs - rhs) < tol * max(mag, 1e-3 * scale), (lhs, rhs, parts)


def test_vjp_per_parameter_microfacet():
    """one JVP per material parameter against the matching entry of a single VJP"""
    import torch
    import psdr_jit_b200 as psdr
    rng = np.random.default_rng(11)
    w = h = 48
    sc = build_product(scenes.cbox_meshes(), w, h, 16, 0, 0, bsdfs=scenes.CBOX_MF_BSDFS)
    integ = psdr.PathTracer(3)
    cot = torch.as_tensor(rng.normal(size=(w * h, 3)).astype(np.float32), device="cuda")
    integ.render_vjp(sc, cot, 0, seed=2, terms=1)
    grads = {(n, f): sc.grad_of(n, f).ravel().copy() for n in ("BSDF[id=cat]", "BSDF[id=white]")
             for f in ("specularReflectance", "diffuseReflectance", "roughness")}
    for (n, f), g in grads.items():
        for k in range(len(g)):
            sc2 = build_product(scenes.cbox_meshes(), w, h, 16, 0, 0, bsdfs=scenes.CBOX_MF_BSDFS)
            b = sc2.param_map[n]
            if f == "roughness":
                b.d_roughness = np.float32(1.0)
            else:
                t = np.zeros(3, np.float32)
                t[k] = 1.0
                setattr(b, "d_" + f, t)
            sc2.configure()
            sc2.configure([0])
            dimg = integ.renderD_fwd(sc2, 0, seed=2, terms=1)[1]
            lhs = float((cot.double() * dimg.double()).sum())
            ref = float(torch.linalg.norm(cot.double()) * torch.linalg.norm(dimg.double()))
            assert abs(lhs - float(g[k])) < 3e-4 * max(abs(lhs), 1e-2 * ref), (n, f, k, lhs, float(g[k]))


def test_autograd_backward_matches_forward_mode():
    """renderD returns an image with an autograd node; d loss / dP through it equals <dloss/dimg, dimg/dP>."""
    import torch
    import psdr_jit_b200 as psdr
    meshes = scenes.cbox_meshes()
    w = h = 64
    sc = build_product(meshes, w, h, 8, 8, 8, move_mesh=0, axis_scale=(100.0, 0.0, 0.0))
    integ = psdr.PathTracer(2)
    img_f, dimg = integ.renderD_fwd(sc, 0, seed=5)
    target = torch.full_like(img_f, 0.25)
    dl_dimg = 2.0 * (img_f - target) / img_f.numel()
    expect = float((dl_dimg.double() * dimg.double()).sum())

    sc2 = build_product(meshes, w, h, 8, 8, 8)
    P = torch.zeros((), dtype=torch.float32, requires_grad=True)
    M = torch.eye(4).clone()
    M[0, 3] = P * 100.0
    sc2.param_map["Mesh[0]"].set_transform(M)
    R = torch.tensor([20.0, 20.0, 8.0], requires_grad=True)
    sc2.param_map["Emitter[0]"].radiance = R
    sc2.configure()
    sc2.configure([0])
    img = integ.renderD(sc2, 0, seed=5)
    assert img.requires_grad
    loss = ((img - target) ** 2).mean()
    loss.backward()
    assert abs(float(P.grad) - expect) < 2e-4 * abs(expect), (float(P.grad), expect)
    # d loss / d radiance: the image is linear in the radiance (interior term only, detached elsewhere)
    sc3 = build_product(meshes, w, h, 8, 0, 0, d_radiance=(1.0, 0.0, 0.0))
    a, da = integ.renderD_fwd(sc3, 0, seed=5, terms=1)
    assert abs(float(R.grad[0]) - float((dl_dimg.double() * da.double()).sum())) < 2e-4 * abs(float(R.grad[0]))
    # the primal of the autograd path is the primal of renderD
    assert float((img.detach() - img_f).abs().max()) < 1e-6


def test_vjp_replays_continued_sampler_streams():
    """seed = -1: backward replays the draws its forward call consumed and leaves the streams advanced."""
    import torch
    import psdr_jit_b200 as psdr
    sc = build_product(scenes.cbox_meshes(), 48, 48, 4, 4, 4)
    R = torch.tensor([20.0, 20.0, 8.0], requires_grad=True)
    sc.param_map["Emitter[0]"].radiance = R
    sc.configure()
    sc.configure([0])
    integ = psdr.PathTracer(2)
    integ.renderD_primal(sc, 0, seed=11)
    img = integ.renderD(sc, 0, seed=-1)
    st_after = list(sc._sampler_state())
    img.sum().backward()
    assert list(sc._sampler_state()) == st_after
    # image is linear in the radiance: <grad, R> == sum(img)
    assert abs(float((R.grad * R.detach()).sum()) - float(img.detach().sum())) < 1e-4 * float(img.detach().sum())


def test_vjp_of_guided_secondary_edges():
    import torch
    import psdr_jit_b200 as psdr
    rng = np.random.default_rng(5)
    sc, tang = _scene_with_tangents(psdr, sphere_meshes(), 64, 64, (0, 0, 8), rng, ["mesh_left", "vertices", "camera"])
    integ = psdr.PathTracer(2)
    integ.preprocess_secondary_edges(sc, 0, [100, 4, 4, 8], 1, 0)
    img, dimg = integ.renderD_fwd(sc, 0, seed=3, terms=4)
    cot = torch.as_tensor(rng.normal(size=(64 * 64, 3)).astype(np.float32), device=img.device)
    lhs = float((cot.double() * dimg.double()).sum())
    integ.render_vjp(sc, cot, 0, seed=3, terms=4)
    rhs = sum(float((sc.grad_of(n, f).astype(np.float64) * t.astype(np.float64)).sum()) for (n, f), t in tang.items())
    assert abs(lhs) > 0 and abs(lhs - rhs) < 2e-4 * abs(lhs), (lhs, rhs)


def test_shard_vjp_tables_sum_to_full_vjp():
    """Multi-GPU reverse mode: the gradient tables of the lane shards add up to the table of the whole
    frame (what the single NCCL all-reduce computes), and back-propagating the summed table gives the same parameter
    gradients as the unsharded pass."""
    import torch
    import psdr_jit_b200 as psdr
    kw = dict(move_mesh=0, axis_scale=(100.0, 0.0, 0.0))
    rng = np.random.default_rng(5)
    cot = torch.as_tensor(rng.normal(size=(48 * 48, 3)).astype(np.float32), device="cuda")
    integ = psdr.PathTracer(3)
    full = build_product(scenes.cbox_meshes(), 48, 48, 3, 2, 2, **kw)
    t_full = integ.render_vjp_table(full, cot, 0, seed=4)
    integ.render_vjp(full, cot, 0, seed=4)
    g_full = full.grad_of("Mesh[0]", "to_world_left").copy()
    acc = torch.zeros_like(t_full)
    parts = []
    for r in range(3):
        part = build_product(scenes.cbox_meshes(), 48, 48, 3, 2, 2, shard=(r, 3), **kw)
        parts.append(part)
        acc += integ.render_vjp_table(part, cot, 0, seed=4)
    scale = float(t_full.abs().max())
    assert float((acc - t_full).abs().max()) < 2e-4 * scale
    integ.backprop_table(parts[0], acc, 0)
    g_sum = parts[0].grad_of("Mesh[0]", "to_world_left")
    assert np.abs(g_sum - g_full).max() < 2e-4 * max(np.abs(g_full).max(), 1e-12)


@pytest.mark.parametrize("mis", [0, 1])
def test_direct_integrator_vjp_is_transpose_of_jvp(mis):
    import torch
    import psdr_jit_b200 as psdr
    rng = np.random.default_rng(31 + mis)
    sc, tang = _scene_with_tangents(psdr, scenes.cbox_meshes(), 64, 64, (8, 8, 8), rng, "mesh_left vertices materials camera".split())
    integ = psdr.Direct(mis)
    img, dimg = integ.renderD_fwd(sc, 0, seed=3)
    cot = torch.as_tensor(rng.normal(size=(64 * 64, 3)).astype(np.float32), device=img.device)
    lhs = float((cot.double() * dimg.double()).sum())
    integ.render_vjp(sc, cot, 0, seed=3)
    parts = [float((sc.grad_of(n, f).reshape(np.shape(t)).astype(np.float64) * t.astype(np.float64)).sum()) for (n, f), t in tang.items()]
    mag = max(abs(lhs), sum(abs(v) for v in parts))
    assert abs(lhs - sum(parts)) < 2e-4 * mag, (lhs, sum(parts))
