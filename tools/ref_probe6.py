#!/usr/bin/env python3
"""A DETERMINISTIC anchor for the secondary-edge term: the reference in REVERSE mode.

Its forward-mode secondary-edge derivative image is not reproducible (tools/ref_probe5.py: lost updates in the scatter under
forward-mode AD).  Backward mode turns that scatter into a gather: d/dP of a linear functional <w, image> is a sum over samples
with no write conflict.  This script evaluates, in the reference, d<w_k, renderD(sec term)>/dP by drjit.backward for 32
functionals w_k (16 bands of rows, 16 bands of columns; P = x-translation of the luminaire), twice each (is it reproducible?),
unguided and with PathTracer.preprocess_secondary_edges -- and the same numbers from this repo's forward derivative image,
<w_k, dimg>.  With the guided runs this also pins the guiding masses: the estimator divides by them.

    python tools/ref_probe6.py all     # = ref, ours (separate processes), then the summary
Output: gpurun_out/ref_probe6/summary.json (committed as profiles/r04r_sec_reverse_anchor.json)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", "ref_probe6")
TMP = "/tmp/ref_probe6"
RES, SPPSE, DEPTH, NB = 128, 16, 1, 16
CASES = {"full": (1.0, None), "scaled": (1.0 / 300.0, None), "full_guided": (1.0, [40, 4, 4, 8]), "scaled_guided": (1.0 / 300.0, [40, 4, 4, 8])}


def weights():
    """[32, npix]: 16 row bands, 16 column bands"""
    ys, xs = np.arange(RES * RES) // RES, np.arange(RES * RES) % RES
    w = [(ys // (RES // NB) == k).astype(np.float32) for k in range(NB)] + [(xs // (RES // NB) == k).astype(np.float32) for k in range(NB)]
    return np.stack(w)


def side_ref():
    ns = {"__file__": os.path.join(ROOT, "tools", "ref_golden2.py"), "__name__": "probe6"}
    src = open(os.path.join(ROOT, "tools", "ref_golden2.py")).read().split('section("mf_renderC"')[0].replace('"ref_golden2"', '"ref_probe6_obj"')
    exec(compile(src, "ref_golden2_head", "exec"), ns)
    psdr, scenes, drjit, T, mat = ns["psdr"], ns["scenes"], ns["drjit"], ns["T"], ns["mat"]
    from drjit.cuda import Matrix4f as Matrix4fC
    from drjit.cuda.ad import Float as FloatD, Matrix4f as Matrix4fD
    W = weights()
    out = {}
    for tag, (scale, reso) in CASES.items():
        meshes, cam = scenes.scaled_cbox(scale) if scale != 1.0 else (scenes.cbox_meshes(), scenes.CBOX_CAMERA)
        ax = 100.0 * scale
        sc = psdr.Scene()
        o = sc.opts
        o.width, o.height, o.spp, o.sppe, o.sppse, o.log_level = RES, RES, 0, 0, SPPSE, 0
        sensor = psdr.PerspectiveCamera(cam["fov"], cam["near"], cam["far"])
        sensor.to_world = Matrix4fD(mat(cam["to_world"]))
        sc.add_Sensor(sensor)
        for name, p in scenes.CBOX_BSDFS:
            sc.add_BSDF(psdr.DiffuseBSDF([float(x) for x in p]), name)
        for i, m in enumerate(meshes):
            path = os.path.join(ns["OBJDIR"], "%s%d_%s.obj" % (tag[0], i, m.name))
            scenes.write_obj(m, path)
            em = psdr.AreaLight([float(x) for x in m.emitter]) if m.emitter is not None else None
            sc.add_Mesh(path, Matrix4fC(mat(m.to_world)), m.bsdf, em)
        integ = psdr.PathTracer(DEPTH)
        res = np.zeros((2, len(W)), np.float64)
        for rep in range(2):
            for k in range(len(W)):
                P = FloatD(0.)
                drjit.enable_grad(P)
                sc.param_map["Mesh[0]"].set_transform(Matrix4fD(T(P * ax, P * 0., P * 0.)))
                sc.configure(); sc.configure([0])
                if reso is not None:
                    with drjit.suspend_grad():
                        integ.preprocess_secondary_edges(sc, 0, reso, 1)
                img = integ.renderD(sc, 0, seed=3)
                w = FloatD(W[k].copy())
                loss = drjit.sum(img[0] * w + img[1] * w + img[2] * w)
                drjit.backward(loss)
                res[rep, k] = float(np.asarray(drjit.grad(P).numpy()).ravel()[0])
        out[tag] = res
        print("ref", tag, "rep spread", float(np.abs(res[0] - res[1]).max()), "max |value|", float(np.abs(res).max()), flush=True)
    os.makedirs(TMP, exist_ok=True)
    np.savez(os.path.join(TMP, "ref.npz"), **out)


def side_ours():
    sys.path.insert(0, ROOT)
    import psdr_jit_b200 as psdr
    from tests.common import build_product, scenes
    W = weights().astype(np.float64)
    out = {}
    for tag, (scale, reso) in CASES.items():
        meshes, cam = scenes.scaled_cbox(scale) if scale != 1.0 else (scenes.cbox_meshes(), scenes.CBOX_CAMERA)
        sc = build_product(meshes, RES, RES, 0, 0, SPPSE, cam=cam, move_mesh=0, axis_scale=(100.0 * scale, 0.0, 0.0))
        integ = psdr.PathTracer(DEPTH)
        if reso is not None:
            integ.preprocess_secondary_edges(sc, 0, reso, 1)
        dimg = integ.renderD_fwd(sc, 0, seed=3, terms=4)[1].cpu().numpy().astype(np.float64)
        out[tag] = W @ dimg.sum(axis=1)
    os.makedirs(TMP, exist_ok=True)
    np.savez(os.path.join(TMP, "ours.npz"), **out)


def analyze():
    r, o = np.load(os.path.join(TMP, "ref.npz")), np.load(os.path.join(TMP, "ours.npz"))
    s = {"what": "d<w_k, secondary-edge derivative image>/dP for 16 row bands + 16 column bands; reference: drjit.backward (two runs), ours: <w_k, forward derivative image>",
         "resolution": RES, "sppse": SPPSE}
    for tag in CASES:
        a, b = r[tag], o[tag]
        ratio = float(np.dot(a[0], b) / max(np.dot(b, b), 1e-30))          # least-squares scale ref = ratio * ours
        s[tag] = {"ref_run_to_run_max_abs": float(np.abs(a[0] - a[1]).max()), "ref_max_abs": float(np.abs(a).max()),
                  "ref_over_ours_scale": ratio,
                  "rel_l2_ours_vs_ref": float(np.linalg.norm(ratio * b - a[0]) / max(np.linalg.norm(a[0]), 1e-30)),
                  "rel_l2_ours_vs_ref_scale_2": float(np.linalg.norm(2.0 * b - a[0]) / max(np.linalg.norm(a[0]), 1e-30)),
                  "ref": [float(x) for x in a[0]], "ours": [float(x) for x in b]}
    os.makedirs(OUT, exist_ok=True)
    json.dump(s, open(os.path.join(OUT, "summary.json"), "w"), indent=1)
    for tag in CASES:
        print(tag, {k: v for k, v in s[tag].items() if k not in ("ref", "ours")})


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what == "all":
        for side in ("ref", "ours"):
            rc = subprocess.call([sys.executable, os.path.abspath(__file__), side])
            if rc != 0:
                print("side", side, "failed", rc)
        analyze()
    else:
        {"ref": side_ref, "ours": side_ours, "analyze": analyze}[what]()
