"""Interior adjoint: kernel time vs lanes, with and without the L2 flush, at both CTA shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import psdr_jit_b200 as psdr
from psdr_jit_b200 import _lib
from tests.common import build_product, scenes
L = _lib.load()
kw = dict(move_mesh=0, axis_scale=(100.0, 0.0, 0.0))
cot = torch.ones((512 * 512, 3), device="cuda")
integ = psdr.PathTracer(3)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for policy in (0, 1):
    psdr.set_cta_policy(policy)
    for do_flush in (True, False):
        row = []
        for spp in (1, 2, 4, 8, 16, 32):
            sc = build_product(scenes.cbox_meshes(), 512, 512, spp, 0, 0, **kw)
            _lib.check(L.psdr_scene_enable_timing(sc._h, 1))
            res = []
            for it in range(6):
                if do_flush:
                    flush.fill_(it)
                integ.render_vjp_table(sc, cot, 0, seed=it, terms=1)
                torch.cuda.synchronize()
                res.append(L.psdr_scene_kernel_ms(sc._h, 1))
            row.append(round(float(np.mean(res[2:])), 3))
        print("policy", policy, "flush", do_flush, "interior adjoint ms at spp 1,2,4,8,16,32:", row, flush=True)
