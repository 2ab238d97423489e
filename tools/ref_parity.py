#!/usr/bin/env python3
"""Parity of this repo's CUDA path against the RUNNING reference (baseline/_ref, Dr.Jit + OptiX), on a GPU box.

    python tools/ref_parity.py all            # = ref, ours (separate processes), then analyze
    python tools/ref_parity.py ref|ours|analyze

Two experiments on the BASELINE config-2 scene (Cornell box, PathTracer(3), P = x-translation of the luminaire):

A. seed mean (is the residual against the reference a BIAS or discrete, zero-mean noise?)
   renderD with spp = sppe = sppse = 32, first call seed = 0, then N-1 calls with seed = -1 (the sampler streams
   continue, reference README.md:96); image and forward derivative image.  Both sides render the same N sample sets.
   If the two implementations differ only by lanes whose discrete decisions flip (closest-hit ties, the
   t > dist - ShadowEpsilon test), ours_k - ref_k is sparse zero-mean noise and the rel-L2 of the N-call MEANS falls
   like 1/sqrt(N); a bias would leave a floor.

B. flip census (which lanes differ, and why?)
   spp = 1 (lane = pixel), interior term only, renderC and renderD, S calls with continuing streams.  A lane is
   "flipped" when its value differs from the reference's by more than 1e-3 relative.  The CPU oracle re-traces every
   lane with decision margins (oracle/psdr_oracle.cpp LaneDiag): flipped lanes are classified by the margin class that
   is critical for them and compared with the margin distribution of all lanes.

   Every experiment renders the reference four ways -- all three terms in one renderD call, and each term alone (the
   pattern of tutorials/Forward_AD_envmap.ipynb cells 6/10/12): round 2 found that the reference's one-call derivative
   image DISAGREES WITH THE SUM OF ITS OWN TERMS in the blue channel at the luminaire's silhouette pixels (red and
   green agree bit for bit), so the sum of the separately rendered terms is the consistent gradient anchor.

Outputs: gpurun_out/parity/summary.json, seedmean_128.npz (committed as tests/golden/seedmean_128.npz),
cfg2_512_grad_terms.npz, census_flips.npz.
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", "parity")
TMP = os.environ.get("PARITY_TMP", "/tmp/parity")
DEPTH = 3
SEEDMEAN = [(128, 64), (512, 16)]          # (resolution, number of calls)
CENSUS_RES, CENSUS_CALLS = 512, 4
AXIS = (100.0, 0.0, 0.0)
COUNTS = (1, 2, 4, 8, 16, 32, 64)
# all three terms in one renderD call, and each term alone (the pattern of tutorials/Forward_AD_envmap.ipynb cells 6/10/12)
VARIANTS = {"all": (32, 32, 32), "int": (32, 0, 0), "pri": (0, 32, 0), "sec": (0, 0, 32)}
if os.environ.get("PARITY_SMALL"):          # dry run of the plumbing
    SEEDMEAN, CENSUS_RES, CENSUS_CALLS = [(32, 4)], 32, 2
if os.environ.get("PARITY_SKIP_CENSUS"):
    CENSUS_CALLS = 0


def T(x, y, z):
    return [[1., 0., 0., x], [0., 1., 0., y], [0., 0., 1., z], [0., 0., 0., 1.]]


# ------------------------------------------------------------------------------------------------ reference side
def side_ref():
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("scenes", os.path.join(ROOT, "psdr_jit_b200", "scenes.py"))
    scenes = importlib.util.module_from_spec(spec)
    sys.modules["scenes"] = scenes
    spec.loader.exec_module(scenes)
    import drjit
    import psdr_jit as psdr
    from drjit.cuda import Matrix4f as Matrix4fC
    from drjit.cuda.ad import Float as FloatD, Matrix4f as Matrix4fD
    objdir = os.path.join(TMP, "obj")
    os.makedirs(objdir, exist_ok=True)
    mat = lambda m: [[float(m[i][j]) for j in range(4)] for i in range(4)]   # noqa: E731

    def build(w, h, spp, sppe, sppse):
        sc = psdr.Scene()
        o = sc.opts
        o.width, o.height, o.spp, o.sppe, o.sppse, o.log_level = w, h, spp, sppe, sppse, 0
        cam = scenes.CBOX_CAMERA
        sensor = psdr.PerspectiveCamera(cam["fov"], cam["near"], cam["far"])
        sensor.to_world = Matrix4fD(mat(cam["to_world"]))
        sc.add_Sensor(sensor)
        for name, refl in scenes.CBOX_BSDFS:
            sc.add_BSDF(psdr.DiffuseBSDF([float(x) for x in refl]), name)
        for i, m in enumerate(scenes.cbox_meshes()):
            path = os.path.join(objdir, "m%d_%s.obj" % (i, m.name))
            scenes.write_obj(m, path)
            em = psdr.AreaLight([float(x) for x in m.emitter]) if m.emitter is not None else None
            sc.add_Mesh(path, Matrix4fC(mat(m.to_world)), m.bsdf, em)
        return sc

    integ = psdr.PathTracer(DEPTH)

    def render_d(sc, seed, want_grad=True):
        P = FloatD(0.)
        drjit.enable_grad(P)
        sc.param_map["Mesh[0]"].set_transform(Matrix4fD(T(P * AXIS[0], P * AXIS[1], P * AXIS[2])))
        sc.configure()
        sc.configure([0])
        img = integ.renderD(sc, 0) if seed == -1 else integ.renderD(sc, 0, seed=seed)
        drjit.eval(img)
        g = None
        if want_grad:
            drjit.set_grad(P, 1.0)
            drjit.forward_to(img)
            g = drjit.grad(img)
            drjit.eval(g)
        drjit.sync_thread()
        return np.asarray(img.numpy(), np.float32), (None if g is None else np.asarray(g.numpy(), np.float32))

    times = {}
    for res, n in SEEDMEAN:
        for vname, (a, b, c) in VARIANTS.items():
            sc = build(res, res, a, b, c)
            acc_i = np.zeros((res * res, 3), np.float64)
            acc_g = np.zeros((res * res, 3), np.float64)
            t0 = time.time()
            for k in range(n):
                img, g = render_d(sc, 0 if k == 0 else -1)
                acc_i += img
                acc_g += g
                if k == 0:
                    np.save(os.path.join(TMP, "ref_sm%d_%s_first_img.npy" % (res, vname)), img)
                    np.save(os.path.join(TMP, "ref_sm%d_%s_first_grad.npy" % (res, vname)), g)
                if (k + 1) in COUNTS:
                    np.save(os.path.join(TMP, "ref_sm%d_%s_mean%d_img.npy" % (res, vname, k + 1)), (acc_i / (k + 1)).astype(np.float32))
                    np.save(os.path.join(TMP, "ref_sm%d_%s_mean%d_grad.npy" % (res, vname, k + 1)), (acc_g / (k + 1)).astype(np.float32))
            times["seedmean_%d_%s" % (res, vname)] = time.time() - t0
            print("ref seedmean", res, vname, n, "calls", times["seedmean_%d_%s" % (res, vname)], "s", flush=True)
    # census: spp 1, interior term only
    res = CENSUS_RES
    sc = build(res, res, 1, 0, 0)
    sc.configure()
    sc.configure([0])
    t0 = time.time()
    for k in range(CENSUS_CALLS):
        img = integ.renderC(sc, 0) if k else integ.renderC(sc, 0, seed=0)
        drjit.eval(img)
        np.save(os.path.join(TMP, "ref_census_C%d.npy" % k), np.asarray(img.numpy(), np.float32))
    sc = build(res, res, 1, 0, 0)
    for k in range(CENSUS_CALLS):
        img, _ = render_d(sc, 0 if k == 0 else -1, want_grad=False)
        np.save(os.path.join(TMP, "ref_census_D%d.npy" % k), img)
    times["census"] = time.time() - t0
    json.dump(times, open(os.path.join(TMP, "ref_times.json"), "w"))


# ------------------------------------------------------------------------------------------------ this repo's side
def side_ours():
    sys.path.insert(0, ROOT)
    import torch
    import psdr_jit_b200 as psdr
    from tests.common import build_product, scenes
    integ = psdr.PathTracer(DEPTH)
    integ.reference_tangent_scaling = True
    # ours: IEEE arithmetic (the default, bit-comparable with the CPU oracle); oursra: Scene.reference_arithmetic
    # (approximate rcp at the analytic primary hit only); oursfa: the whole library compiled with Dr.Jit-like approximate
    # division / sqrt (PSDR_REFERENCE_ARITHMETIC=1, psdr_jit_b200/build.py variant "refarith")
    full_approx = os.environ.get("PSDR_REFERENCE_ARITHMETIC") == "1"
    for tag, ref_arith in ((("oursfa", True),) if full_approx else (("ours", False), ("oursra", True))):
        for res, n in SEEDMEAN:
            sc = build_product(scenes.cbox_meshes(), res, res, 32, 32, 32, move_mesh=0, axis_scale=AXIS)
            sc.reference_arithmetic = ref_arith
            sc.configure([0])
            acc_i = torch.zeros((res * res, 3), dtype=torch.float64, device="cuda")
            acc_g = torch.zeros_like(acc_i)
            for k in range(n):
                sc.configure([0])
                img, g = integ.renderD_fwd(sc, 0, seed=0 if k == 0 else -1)
                acc_i += img
                acc_g += g
                if (k + 1) in COUNTS:
                    np.save(os.path.join(TMP, "%s_sm%d_mean%d_img.npy" % (tag, res, k + 1)), (acc_i / (k + 1)).float().cpu().numpy())
                    np.save(os.path.join(TMP, "%s_sm%d_mean%d_grad.npy" % (tag, res, k + 1)), (acc_g / (k + 1)).float().cpu().numpy())
    if full_approx:
        return
    # primary-hit triangle ids (which pixels look at the tall box's side face, triangles 22 / 23)
    for res, n in SEEDMEAN:
        sc = build_product(scenes.cbox_meshes(), res, res, 16, 0, 0, move_mesh=0, axis_scale=AXIS)
        tri = psdr.PathTracer(1).render_aov(sc, 0, seed=0).cpu().numpy()[:, 1].astype(np.int32).reshape(res * res, 16)
        np.save(os.path.join(TMP, "ours_sm%d_face22.npy" % res), np.isin(tri, [22, 23]).any(axis=1))
    res = CENSUS_RES
    sc = build_product(scenes.cbox_meshes(), res, res, 1, 0, 0, move_mesh=0, axis_scale=AXIS)
    for k in range(CENSUS_CALLS):
        np.save(os.path.join(TMP, "ours_census_C%d.npy" % k), integ.renderC(sc, 0, seed=0 if k == 0 else -1).cpu().numpy())
    sc = build_product(scenes.cbox_meshes(), res, res, 1, 0, 0, move_mesh=0, axis_scale=AXIS)
    for k in range(CENSUS_CALLS):
        np.save(os.path.join(TMP, "ours_census_D%d.npy" % k), integ.renderD_primal(sc, 0, seed=0 if k == 0 else -1).cpu().numpy())


# ------------------------------------------------------------------------------------------------ analysis
def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def analyze():
    sys.path.insert(0, ROOT)
    os.makedirs(OUT, exist_ok=True)
    S = {"scene": "cbox, PathTracer(3), P = translate(100 P, 0, 0) of Mesh[0]; reference_tangent_scaling on our side",
         "seedmean": {}, "census": {}}
    ld = lambda n: np.load(os.path.join(TMP, n))   # noqa: E731
    for res, n in SEEDMEAN:
        rows = []
        face = ld("ours_sm%d_face22.npy" % res)
        for k in COUNTS:
            if k > n:
                break
            R = {v: (ld("ref_sm%d_%s_mean%d_img.npy" % (res, v, k)), ld("ref_sm%d_%s_mean%d_grad.npy" % (res, v, k))) for v in VARIANTS}
            ri, rg = R["all"]
            rg_terms = R["int"][1] + R["pri"][1] + R["sec"][1]         # the reference's terms rendered one at a time, summed
            row = {"calls": k, "reference_all_vs_its_own_terms_rel_l2_grad": rel_l2(rg, rg_terms),
                   "reference_all_vs_its_own_interior_rel_l2_img": rel_l2(ri, R["int"][0])}
            for tag in ("ours", "oursra", "oursfa"):
                if not os.path.exists(os.path.join(TMP, "%s_sm%d_mean%d_img.npy" % (tag, res, k))):
                    continue
                oi, og = ld("%s_sm%d_mean%d_img.npy" % (tag, res, k)), ld("%s_sm%d_mean%d_grad.npy" % (tag, res, k))
                dg = np.abs(og - rg_terms).max(axis=1)
                row[tag] = {"rel_l2_img": rel_l2(oi, ri), "rel_l2_img_off_face22": rel_l2(oi[~face], ri[~face]),
                            "face22_mean_bias": float((oi[face].astype(np.float64) - ri[face]).sum() / ri[face].astype(np.float64).sum()),
                            "rel_l2_grad_vs_all": rel_l2(og, rg), "rel_l2_grad_vs_terms": rel_l2(og, rg_terms),
                            "pixels_grad_vs_terms_gt_1e-3": int((dg > 1e-3 * np.abs(rg_terms).max()).sum())}
            rows.append(row)
            print("seedmean", res, json.dumps(row), flush=True)
        S["seedmean"][str(res)] = rows
        S["seedmean"]["face22_pixels_%d" % res] = int(face.sum())
        if res == 128:
            np.savez_compressed(os.path.join(OUT, "seedmean_128.npz"),
                                first_img=ld("ref_sm128_all_first_img.npy"), first_grad=ld("ref_sm128_all_first_grad.npy"),
                                first_grad_terms=ld("ref_sm128_int_first_grad.npy") + ld("ref_sm128_pri_first_grad.npy") + ld("ref_sm128_sec_first_grad.npy"),
                                mean_img=ld("ref_sm128_all_mean%d_img.npy" % n), mean_grad=ld("ref_sm128_all_mean%d_grad.npy" % n),
                                mean_grad_terms=ld("ref_sm128_int_mean%d_grad.npy" % n) + ld("ref_sm128_pri_mean%d_grad.npy" % n) + ld("ref_sm128_sec_mean%d_grad.npy" % n),
                                face22=face, calls=np.int32(n), depth=np.int32(DEPTH), spp=np.int32(32), axis=np.float32(AXIS))
        if res == 512:      # the consistent full-size gradient anchor: the reference's three terms of call 1, summed
            g = ld("ref_sm512_int_first_grad.npy") + ld("ref_sm512_pri_first_grad.npy") + ld("ref_sm512_sec_first_grad.npy")
            np.savez_compressed(os.path.join(OUT, "cfg2_512_grad_terms.npz"), grad_terms=g.astype(np.float16 if False else np.float32), face22=face)
    # ---- census
    from tests.common import build_oracle, scenes
    res = CENSUS_RES
    osc = build_oracle(scenes.cbox_meshes(), res, res, 1, 0, 0, move_mesh=0, axis_scale=AXIS)
    names = ("shadow", "border", "self", "tie")
    flips_out = {}
    for mode, tag, draws in ((0, "C", 2 + 5 * DEPTH), (1, "D", 2 + 5 * DEPTH)):
        if CENSUS_CALLS == 0:
            break
        tot = {"lanes": 0, "flipped": 0, "ours_vs_oracle_flipped": 0}
        diag_all, flip_all, ids = [], [], []
        for k in range(CENSUS_CALLS):
            ref, ours = ld("ref_census_%s%d.npy" % (tag, k)), ld("ours_census_%s%d.npy" % (tag, k))
            oimg, lanes, diag = osc.render_diag(DEPTH, seed=0, mode=mode, skip=[k * draws, 0, 0])
            scale = np.maximum(np.abs(ref).max(axis=1), 1e-2)
            flipped = np.abs(ours - ref).max(axis=1) > 1e-3 * scale
            tot["lanes"] += len(ref)
            tot["flipped"] += int(flipped.sum())
            tot["ours_vs_oracle_flipped"] += int((np.abs(ours - oimg).max(axis=1) > 1e-3 * scale).sum())
            diag_all.append(diag)
            flip_all.append(flipped)
            ids.append(np.nonzero(flipped)[0] + k * res * res)
            tot.setdefault("rel_l2_per_call", []).append(rel_l2(ours, ref))
            tot.setdefault("rel_l2_per_call_without_flipped", []).append(rel_l2(np.where(flipped[:, None], ref, ours), ref))
            tot.setdefault("mean_signed_diff_over_mean_ref", []).append(float((ours.astype(np.float64) - ref).sum() / ref.astype(np.float64).sum()))
        diag, flipped = np.concatenate(diag_all), np.concatenate(flip_all)
        cls = {}
        # thresholds: the 2 % quantile of each margin over ALL lanes; a flipped lane is "explained" by a class when its
        # margin is below that threshold (for the shadow class: closer to the decision boundary than 10 % of ShadowEpsilon)
        thr = {}
        for j, nm in enumerate(names):
            v = diag[:, j]
            fin = v < 1e29
            thr[nm] = float(np.quantile(v[fin], 0.02)) if fin.any() else 0.0
        explained = np.zeros(len(diag), bool)
        for j, nm in enumerate(names):
            below = diag[:, j] <= thr[nm]
            cls[nm] = {"threshold_2pct_quantile": thr[nm], "all_lanes_below": float(below.mean()),
                       "flipped_lanes_below": float(below[flipped].mean()) if flipped.any() else None,
                       "median_all": float(np.median(diag[:, j][diag[:, j] < 1e29])) if (diag[:, j] < 1e29).any() else None,
                       "median_flipped": float(np.median(diag[flipped, j][diag[flipped, j] < 1e29])) if (diag[flipped, j] < 1e29).any() else None}
            explained |= below
        tot["classes"] = cls
        tot["flipped_explained_by_any_class"] = float(explained[flipped].mean()) if flipped.any() else None
        tot["all_lanes_in_any_class"] = float(explained.mean())
        S["census"]["render" + tag] = tot
        flips_out["ids_" + tag] = np.concatenate(ids).astype(np.int64)
        flips_out["diag_" + tag] = diag[flipped]
        print("census", tag, json.dumps(tot)[:1500], flush=True)
    if flips_out:
        np.savez_compressed(os.path.join(OUT, "census_flips.npz"), **flips_out)
    try:
        S["ref_times_s"] = json.load(open(os.path.join(TMP, "ref_times.json")))
    except Exception:
        pass
    json.dump(S, open(os.path.join(OUT, "summary.json"), "w"), indent=1)


if __name__ == "__main__":
    os.makedirs(TMP, exist_ok=True)
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what == "all":
        for side in ("ref", "ours"):
            t0 = time.time()
            rc = subprocess.call([sys.executable, os.path.abspath(__file__), side])
            print("side", side, "rc", rc, "%.1f s" % (time.time() - t0), flush=True)
            if rc:
                sys.exit(rc)
        env = dict(os.environ, PSDR_REFERENCE_ARITHMETIC="1")
        t0 = time.time()
        rc = subprocess.call([sys.executable, os.path.abspath(__file__), "ours"], env=env)
        print("side oursfa rc", rc, "%.1f s" % (time.time() - t0), flush=True)
        analyze()
    elif what == "ref":
        side_ref()
    elif what == "ours":
        side_ours()
    else:
        analyze()
