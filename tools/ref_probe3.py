#!/usr/bin/env python3
"""Probe of the reference's guided secondary-edge sampling with tiny grids (isolates the semantics of
HyperCubeDistribution3f per dimension).  Output: gpurun_out/ref_probe3/guided_probe.npz"""
import os, sys, runpy
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", "ref_probe3")
os.makedirs(OUT, exist_ok=True)
sys.argv = ["x"]
ns = {"__file__": os.path.join(ROOT, "tools", "ref_golden2.py"), "__name__": "probe"}
src = open(os.path.join(ROOT, "tools", "ref_golden2.py")).read().split('section("mf_renderC"')[0].replace('"ref_golden2"', '"ref_probe3"')
exec(compile(src, "ref_golden2_head", "exec"), ns)
build, render_d, psdr, scenes = ns["build"], ns["render_d"], ns["psdr"], ns["scenes"]
sphere = ns["sphere_meshes"]
out = {}
for name, prep in (("g111", ([1, 1, 1, 8], 1, 0)), ("g211", ([2, 1, 1, 64], 1, 0)), ("g121", ([1, 2, 1, 64], 1, 0)), ("g112", ([1, 1, 2, 64], 1, 0)),
                   ("g311", ([3, 1, 1, 64], 1, 3)), ("g222", ([2, 2, 2, 64], 1, 0)), ("g811", ([8, 1, 1, 64], 1, 0))):
    sc = build(sphere(), scenes.CBOX_BSDFS, 128, 128, 0, 0, 8)
    img, g = render_d(sc, psdr.PathTracer(2), 1, 8, (40.0, 20.0, 0.0), prep)
    out[name] = g
    print(name, float(np.abs(g).sum()), float(g.sum()), flush=True)
np.savez_compressed(os.path.join(OUT, "guided_probe.npz"), **out)
