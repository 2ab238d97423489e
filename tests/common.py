"""Shared helpers for the tests: scene construction for the oracle and metrics."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

from psdr_jit_b200 import scenes  # noqa: E402  (pure numpy module)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def sphere_meshes():
    return scenes.cbox_meshes() + [scenes.icosphere(2, 80.0, (185.0, 250.0, 169.0))]


def translation_tangent(axis_scale):
    d = np.zeros((4, 4), dtype=np.float32)
    d[0, 3], d[1, 3], d[2, 3] = axis_scale
    return d


def build_oracle(meshes, w, h, spp, sppe, sppse, move_mesh=None, axis_scale=None, active=(0,), cam=None):
    """Scene = scenes.* meshes + CBOX bsdfs + camera; derivative parameter P translates mesh
    `move_mesh` by P*axis_scale through to_world_left (reference README.md:87-90)."""
    from oracle.psdr_oracle import OracleScene
    cam = cam or scenes.CBOX_CAMERA
    sc = OracleScene(w, h, spp, sppe, sppse)
    for name, refl in scenes.CBOX_BSDFS:
        sc.add_diffuse(name, refl)
    for i, m in enumerate(meshes):
        dtw = None
        if move_mesh is not None and i == move_mesh:
            dtw = {"left": translation_tangent(axis_scale)}
        sc.add_mesh(m.v, m.f, m.bsdf, uv=m.uv, fuv=m.fuv, to_world={"raw": m.to_world}, d_to_world=dtw, radiance=m.emitter)
    sc.add_camera(cam["fov"], cam["near"], cam["far"], {"raw": cam["to_world"]})
    sc.configure(active)
    return sc
