#!/usr/bin/env python3
"""Prints CUDA-vs-oracle and CUDA-vs-reference-golden statistics (no assertions); dev aid."""
import sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.common import *
import psdr_jit_b200 as psdr
import torch

def show(tag, a, b):
    r, nbad, rex = compare_stats(a, b)
    print("  %-34s rel_l2 %.3e  off %5d/%d  rel_l2(ex) %.3e" % (tag, r, nbad, len(a), rex), flush=True)

print("== aov")
for nm, meshes in (("cbox", scenes.cbox_meshes()), ("sphere", sphere_meshes())):
    ref = build_oracle(meshes, 128, 128, 1, 0, 0).aov()
    for accel in (0, 1):
        sc = build_product(meshes, 128, 128, 1, 0, 0, accel=accel)
        got = psdr.PathTracer(1).render_aov(sc, 0, seed=0).cpu().numpy()
        print(nm, "accel", accel, "mesh id mismatches", (got[:, 0] != ref[:, 0]).sum(), "tri", (got[:, 1] != ref[:, 1]).sum(), "pos maxdiff", np.abs(got[:, 2:5] - ref[:, 2:5]).max())
print("== renderC vs oracle")
for depth, spp, seed in ((1, 1, 0), (3, 4, 3), (6, 2, 11)):
    ref = build_oracle(scenes.cbox_meshes(), 128, 128, spp, 0, 0).render(depth, seed=seed, mode=0)
    sc = build_product(scenes.cbox_meshes(), 128, 128, spp, 0, 0)
    got = psdr.PathTracer(depth).renderC(sc, 0, seed=seed).cpu().numpy()
    show("depth %d spp %d" % (depth, spp), got, ref)
print("== renderD vs oracle")
CASES = [("light", scenes.cbox_meshes(), 3, 0, 0, (100.0, 0.0, 0.0)), ("smallbox", scenes.cbox_meshes(), 2, 5, 1, (0.0, 30.0, 50.0)),
         ("sphere", sphere_meshes(), 2, 1, 8, (40.0, 20.0, 0.0))]
for name, meshes, depth, seed, mesh, axis in CASES:
    for terms in (1, 2, 4, 7):
        spps = (4 if terms & 1 else 0, 4 if terms & 2 else 0, 4 if terms & 4 else 0)
        osc = build_oracle(meshes, 128, 128, *spps, move_mesh=mesh, axis_scale=axis)
        img_ref, dimg_ref = osc.render(depth, seed=seed, mode=1, terms=7)
        sc = build_product(meshes, 128, 128, *spps, move_mesh=mesh, axis_scale=axis)
        img, dimg = psdr.PathTracer(depth).renderD_fwd(sc, 0, seed=seed)
        if terms & 1: show("%s terms %d img" % (name, terms), img.cpu().numpy(), img_ref)
        show("%s terms %d dimg" % (name, terms), dimg.cpu().numpy(), dimg_ref)
print("== cfg2 timing + golden")
kw = dict(move_mesh=0, axis_scale=(100.0, 0.0, 0.0))
sc = build_product(scenes.cbox_meshes(), 512, 512, 32, 32, 32, **kw)
integ = psdr.PathTracer(3)
for term in (1, 2, 4, 7):
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        img, dimg = integ.renderD_fwd(sc, 0, seed=it, terms=term)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    n = 512 * 512 * 32 * bin(term).count("1")
    print("  terms %d: %.2f ms  (%.1f Msamples/s)" % (term, dt * 1e3, n / dt / 1e6), flush=True)
t0 = time.perf_counter(); c = integ.renderC(sc, 0, seed=0); torch.cuda.synchronize(); print("  renderC %.2f ms" % ((time.perf_counter() - t0) * 1e3))
g = np.load(GOLDEN + "/cfg2_512_s32_d3_light.npz")
integ.reference_tangent_scaling = True
img, dimg = integ.renderD_fwd(sc, 0, seed=0)
show("cfg2 img vs reference", img.cpu().numpy(), g["img"])
show("cfg2 grad vs reference", dimg.cpu().numpy(), g["grad"])
print("configure ms", sc.last_configure_ms())
