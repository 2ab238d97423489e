#!/usr/bin/env python3
"""BASELINE config 2 at FULL SIZE (512 x 512, spp = sppe = sppse = 32, PathTracer(3), the luminaire translated along x) on the
scene scaled by 1/300 -- where no decision sits on the reference's fixed 1e-3 epsilon bands -- this repo's CUDA path against
the RUNNING reference, term by term, plus the reference against itself (two identical runs per term).

    python tools/ref_parity_scaled.py all        # = ref, ours (separate processes), then the summary
Output: gpurun_out/parity_scaled/summary.json (committed as profiles/r04n_parity_scaled.json)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", "parity_scaled")
TMP = "/tmp/parity_scaled"
S, RES, SPP, DEPTH = 1.0 / 300.0, int(os.environ.get("PARITY_RES", "512")), int(os.environ.get("PARITY_SPP", "32")), 3
AX = 100.0 * S
TERMS = {"int": (SPP, 0, 0), "pri": (0, SPP, 0), "sec": (0, 0, SPP)}


def side_ref():
    ns = {"__file__": os.path.join(ROOT, "tools", "ref_golden2.py"), "__name__": "parity_scaled"}
    src = open(os.path.join(ROOT, "tools", "ref_golden2.py")).read().split('section("mf_renderC"')[0].replace('"ref_golden2"', '"parity_scaled_obj"')
    exec(compile(src, "ref_golden2_head", "exec"), ns)
    psdr, scenes, drjit, T, mat = ns["psdr"], ns["scenes"], ns["drjit"], ns["T"], ns["mat"]
    from drjit.cuda import Matrix4f as Matrix4fC
    from drjit.cuda.ad import Float as FloatD, Matrix4f as Matrix4fD
    meshes, cam = scenes.scaled_cbox(S)

    def build(spp, sppe, sppse):
        sc = psdr.Scene()
        o = sc.opts
        o.width, o.height, o.spp, o.sppe, o.sppse, o.log_level = RES, RES, spp, sppe, sppse, 0
        sensor = psdr.PerspectiveCamera(cam["fov"], cam["near"], cam["far"])
        sensor.to_world = Matrix4fD(mat(cam["to_world"]))
        sc.add_Sensor(sensor)
        for name, p in scenes.CBOX_BSDFS:
            sc.add_BSDF(psdr.DiffuseBSDF([float(x) for x in p]), name)
        for i, m in enumerate(meshes):
            path = os.path.join(ns["OBJDIR"], "m%d_%s.obj" % (i, m.name))
            scenes.write_obj(m, path)
            em = psdr.AreaLight([float(x) for x in m.emitter]) if m.emitter is not None else None
            sc.add_Mesh(path, Matrix4fC(mat(m.to_world)), m.bsdf, em)
        return sc

    integ = psdr.PathTracer(DEPTH)
    out = {}
    sc = build(SPP, 0, 0)
    sc.configure(); sc.configure([0])
    out["imgC"] = np.asarray(integ.renderC(sc, 0, seed=0).numpy(), np.float32)
    for tag, spps in list(TERMS.items()) + [("all", (SPP, SPP, SPP))]:
        for rep in range(2):
            P = FloatD(0.); drjit.enable_grad(P)
            sc = build(*spps)
            sc.param_map["Mesh[0]"].set_transform(Matrix4fD(T(P * AX, P * 0., P * 0.)))
            sc.configure(); sc.configure([0])
            img = integ.renderD(sc, 0, seed=0)
            drjit.eval(img)
            drjit.set_grad(P, 1.0)
            drjit.forward_to(img)
            g = drjit.grad(img)
            drjit.eval(g)
            if rep == 0:
                out["imgD_" + tag] = np.asarray(img.numpy(), np.float32)
            out["gradD_%s_%d" % (tag, rep)] = np.asarray(g.numpy(), np.float32)
            print("ref", tag, rep, float(np.abs(out["gradD_%s_%d" % (tag, rep)]).mean()), flush=True)
    os.makedirs(TMP, exist_ok=True)
    np.savez(os.path.join(TMP, "ref.npz"), **out)


def side_ours():
    sys.path.insert(0, ROOT)
    import psdr_jit_b200 as psdr
    from tests.common import build_product, scenes
    meshes, cam = scenes.scaled_cbox(S)
    integ = psdr.PathTracer(DEPTH)
    integ.reference_tangent_scaling = True
    out = {"imgC": integ.renderC(build_product(meshes, RES, RES, SPP, 0, 0, cam=cam), 0, seed=0).cpu().numpy()}
    for tag, spps in list(TERMS.items()) + [("all", (SPP, SPP, SPP))]:
        sc = build_product(meshes, RES, RES, *spps, cam=cam, move_mesh=0, axis_scale=(AX, 0.0, 0.0))
        img, d = integ.renderD_fwd(sc, 0, seed=0)
        out["imgD_" + tag], out["gradD_" + tag] = img.cpu().numpy(), d.cpu().numpy()
    os.makedirs(TMP, exist_ok=True)
    np.savez(os.path.join(TMP, "ours.npz"), **out)


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def stats(a, b, flip_rel=1e-3):
    """rel-L2, pixels whose max channel error exceeds flip_rel * max|b|, rel-L2 without them (tests/common.py compare_stats)"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    bad = np.abs(a - b).max(axis=1) > flip_rel * max(np.abs(b).max(), 1e-12)
    a2 = a.copy()
    a2[bad] = b[bad]
    return {"rel_l2": rel_l2(a, b), "pixels_off": int(bad.sum()), "rel_l2_without_them": rel_l2(a2, b)}


def analyze():
    r, o = np.load(os.path.join(TMP, "ref.npz")), np.load(os.path.join(TMP, "ours.npz"))
    s = {"workload": "BASELINE configs[1] scaled by 1/300: cbox %dx%d spp=sppe=sppse=%d PathTracer(%d), d/d(x-translation of the luminaire)" % (RES, RES, SPP, DEPTH),
         "image_renderC_rel_l2": rel_l2(o["imgC"], r["imgC"]), "image_renderD_rel_l2": rel_l2(o["imgD_all"], r["imgD_all"])}
    for tag in ("int", "pri", "sec", "all"):
        s["grad_%s_ours_vs_ref_rel_l2" % tag] = rel_l2(o["gradD_" + tag], r["gradD_%s_0" % tag])
        s["grad_%s_ref_vs_ref_rel_l2" % tag] = rel_l2(r["gradD_%s_1" % tag], r["gradD_%s_0" % tag])
        s["grad_%s_energy" % tag] = float(np.linalg.norm(r["gradD_%s_0" % tag].astype(np.float64)))
    ours_ip = o["gradD_int"].astype(np.float64) + o["gradD_pri"]
    ref_ip = r["gradD_int_0"].astype(np.float64) + r["gradD_pri_0"]
    s["grad_interior_plus_primary_rel_l2"] = rel_l2(ours_ip, ref_ip)
    ref_sum = ref_ip + r["gradD_sec_0"]
    s["grad_sum_of_terms_ours_vs_ref_rel_l2"] = rel_l2(ours_ip + o["gradD_sec"], ref_sum)
    s["grad_ref_one_call_vs_ref_sum_of_terms_rel_l2"] = rel_l2(r["gradD_all_0"], ref_sum)
    s["grad_sum_of_terms_if_sec_were_exact_rel_l2"] = rel_l2(ours_ip + r["gradD_sec_0"], ref_sum)
    s["pixels"] = int(RES * RES)
    s["image_renderC"] = stats(o["imgC"], r["imgC"])
    for tag in ("int", "pri", "sec"):
        s["grad_%s_ours_vs_ref" % tag] = stats(o["gradD_" + tag], r["gradD_%s_0" % tag])
        s["grad_%s_ref_vs_ref" % tag] = stats(r["gradD_%s_1" % tag], r["gradD_%s_0" % tag])
    s["grad_sum_of_terms_ours_vs_ref"] = stats(ours_ip + o["gradD_sec"], ref_sum)
    os.makedirs(OUT, exist_ok=True)
    json.dump(s, open(os.path.join(OUT, "summary.json"), "w"), indent=1)
    print(json.dumps(s, indent=1))


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
