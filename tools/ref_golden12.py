#!/usr/bin/env python3
"""Twelfth set of golden vectors from the RUNNING reference: CollocatedIntegrator(1e6) (src/integrator/collocated.cpp) on
the Cornell box, 128 x 128, spp 4: renderC, and renderD's forward derivative image w.r.t. the small box's translation, the
interior and the primary-edge term one at a time.  Output: gpurun_out/ref_golden12/collocated.npz"""
import os
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", "ref_golden12")
os.makedirs(OUT, exist_ok=True)
ns = {"__file__": os.path.join(ROOT, "tools", "ref_golden2.py"), "__name__": "golden12"}
src = open(os.path.join(ROOT, "tools", "ref_golden2.py")).read().split('section("mf_renderC"')[0].replace('"ref_golden2"', '"ref_golden12"')
exec(compile(src, "ref_golden2_head", "exec"), ns)
psdr, scenes, drjit, T, mat, build = ns["psdr"], ns["scenes"], ns["drjit"], ns["T"], ns["mat"], ns["build"]
from drjit.cuda.ad import Float as FloatD, Matrix4f as Matrix4fD  # noqa: E402

INTENSITY = 1e6
out = {"intensity": np.float32(INTENSITY)}
integ = psdr.CollocatedIntegrator(FloatD(INTENSITY))
try:
    sc = build(scenes.cbox_meshes(), scenes.CBOX_BSDFS, 128, 128, 4, 0, 0)
    sc.configure(); sc.configure([0])
    out["imgC"] = np.asarray(integ.renderC(sc, 0, seed=0).numpy(), np.float32)
    print("renderC mean", float(out["imgC"].mean()), flush=True)
    for tag, (spp, sppe) in (("int", (4, 0)), ("pri", (0, 4))):
        P = FloatD(0.); drjit.enable_grad(P)
        sc = build(scenes.cbox_meshes(), scenes.CBOX_BSDFS, 128, 128, spp, sppe, 0)
        sc.param_map["Mesh[1]"].set_transform(Matrix4fD(T(P * 30., P * 10., P * 0.)))
        sc.configure(); sc.configure([0])
        img = integ.renderD(sc, 0, seed=0)
        drjit.eval(img)
        drjit.set_grad(P, 1.0)
        drjit.forward_to(img)
        g = drjit.grad(img)
        drjit.eval(g)
        out["gradD_" + tag] = np.asarray(g.numpy(), np.float32)
        print(tag, "grad mean abs", float(np.abs(out["gradD_" + tag]).mean()), flush=True)
        np.savez_compressed(os.path.join(OUT, "collocated.npz"), **out)
except Exception:
    traceback.print_exc()
np.savez_compressed(os.path.join(OUT, "collocated.npz"), **out)
print("saved", sorted(out))
