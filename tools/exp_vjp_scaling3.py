"""Is the slow first ~2 ms of the interior adjoint a ramp from idle?  Time it (a) after an idle gap, (b) right behind other GPU work."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import psdr_jit_b200 as psdr
from psdr_jit_b200 import _lib
from tests.common import build_product, scenes
L = _lib.load()
kw = dict(move_mesh=0, axis_scale=(100.0, 0.0, 0.0))
cot = torch.ones((512 * 512, 3), device="cuda")
integ = psdr.PathTracer(3)
big = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
for policy in (0, 1):
    psdr.set_cta_policy(policy)
    for spp in (4, 8, 32):
        sc = build_product(scenes.cbox_meshes(), 512, 512, spp, 0, 0, **kw)
        _lib.check(L.psdr_scene_enable_timing(sc._h, 1))
        out = {}
        for mode in ("idle 20 ms before", "behind 3 ms of memsets", "behind a forward render", "twice back to back (2nd)"):
            res = []
            for it in range(6):
                torch.cuda.synchronize()
                if mode.startswith("idle"):
                    time.sleep(0.02)
                elif mode.startswith("behind 3"):
                    for _ in range(40): big.fill_(1.0)
                elif mode.startswith("behind a"):
                    integ.renderD_fwd(sc, 0, seed=it)
                else:
                    integ.render_vjp_table(sc, cot, 0, seed=it, terms=1)
                integ.render_vjp_table(sc, cot, 0, seed=it, terms=1)
                torch.cuda.synchronize()
                res.append(L.psdr_scene_kernel_ms(sc._h, 1))
            out[mode] = round(float(np.mean(res[2:])), 3)
        print("policy", policy, "spp", spp, out, flush=True)
