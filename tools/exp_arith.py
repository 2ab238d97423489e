#!/usr/bin/env python3
"""cfg-2 image / derivative image of the CURRENT libpsdr_b200.so against the reference goldens: rel-L2, pixels off, and the
bias on the pixels whose primary hit is the tall box's side face (triangles 22/23).  Used to A/B arithmetic variants."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import psdr_jit_b200 as psdr
from tests.common import GOLDEN, build_product, compare_stats, rel_l2, scenes

tag = sys.argv[1] if len(sys.argv) > 1 else "variant"
kw = dict(move_mesh=0, axis_scale=(100.0, 0.0, 0.0))
integ = psdr.PathTracer(3)
integ.reference_tangent_scaling = True
for res, gname in ((512, "cfg2_512_s32_d3_light.npz"), (128, "renderD_128_s32_d3_light.npz")):
    g = np.load(os.path.join(GOLDEN, gname))
    sc = build_product(scenes.cbox_meshes(), res, res, 32, 32, 32, **kw)
    img, dimg = integ.renderD_fwd(sc, 0, seed=0)
    img, dimg = img.cpu().numpy(), dimg.cpu().numpy()
    sc1 = build_product(scenes.cbox_meshes(), res, res, 1, 0, 0, **kw)
    tri = psdr.PathTracer(1).render_aov(sc1, 0, seed=0).cpu().numpy()[:, 1].astype(int)
    m = np.isin(tri, [22, 23])
    r, nbad, r_ex = compare_stats(img, g["img"], flip_rel=2e-5)
    rg, nbadg, rg_ex = compare_stats(dimg, g["grad"], flip_rel=2e-5)
    print("[%s] %d^2 image: rel-L2 %.3e, %d pixels off, rel-L2 of the rest %.3e | on tri 22/23 (%d px): mean(ours-ref)/mean(ref) %.4f, rel-L2 complement %.3e"
          % (tag, res, r, nbad, r_ex, int(m.sum()), float((img[m] - g["img"][m]).sum() / g["img"][m].sum()), rel_l2(img[~m], g["img"][~m])))
    print("[%s] %d^2 derivative image: rel-L2 %.3e, %d pixels off, rel-L2 of the rest %.3e" % (tag, res, rg, nbadg, rg_ex), flush=True)
