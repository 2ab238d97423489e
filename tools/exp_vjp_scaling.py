"""Where does the interior adjoint lose time when the frame is sharded?  One GPU: kernel time vs spp and vs shard."""
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
for spp, shard in ((32, None), (16, None), (8, None), (32, (0, 2)), (32, (1, 2)), (32, (0, 4)), (32, (0, 8))):
    sc = build_product(scenes.cbox_meshes(), 512, 512, spp, spp, spp, shard=shard, **kw)
    _lib.check(L.psdr_scene_enable_timing(sc._h, 1))
    res = {1: [], 2: [], 4: []}
    fres = {1: [], 2: [], 4: []}
    for it in range(6):
        flush.fill_(it)
        integ.render_vjp_table(sc, cot, 0, seed=it)
        torch.cuda.synchronize()
        for t in res: res[t].append(L.psdr_scene_kernel_ms(sc._h, t))
        flush.fill_(it)
        integ.renderD_fwd(sc, 0, seed=it)
        torch.cuda.synchronize()
        for t in fres: fres[t].append(L.psdr_scene_kernel_ms(sc._h, t))
    print("spp", spp, "shard", shard, "adjoint", {t: round(float(np.mean(v[2:])), 3) for t, v in res.items()},
          "forward", {t: round(float(np.mean(v[2:])), 3) for t, v in fres.items()}, flush=True)
