"""Synthetic benchmark scenes (in-memory geometry; no files needed at run time).

The Cornell box below is the scene BASELINE.json's configs are quoted on: the
eight meshes the reference README builds (reference README.md:72-85 lists them;
the geometry is the standard Cornell-box measurement set).  Quads are already
split into two triangles ``(a,b,c),(a,c,d)`` so that every consumer -- this
package, the CPU oracle and the reference (through ``Mesh.load_raw`` or OBJ files
written by :func:`write_obj`) -- sees the *same* triangle list, face order and
therefore the same edge tables and light-face pmf order.

Everything is plain numpy; nothing here touches the GPU.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

F32 = np.float32


def translate(x: float, y: float, z: float) -> np.ndarray:
    m = np.eye(4, dtype=F32)
    m[0, 3], m[1, 3], m[2, 3] = x, y, z
    return m


def _quad(a: int, b: int, c: int, d: int):
    return [(a, b, c), (a, c, d)]


@dataclass
class MeshData:
    name: str
    v: np.ndarray                      # [nv,3] float32 object-space vertices
    f: np.ndarray                      # [nf,3] int32
    uv: Optional[np.ndarray] = None    # [nuv,2] float32
    fuv: Optional[np.ndarray] = None   # [nf,3] int32
    to_world: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=F32))
    bsdf: str = ""
    emitter: Optional[Sequence[float]] = None   # area-light radiance or None


def _m(name, v, f, bsdf, uv=None, fuv=None, to_world=None, emitter=None) -> MeshData:
    return MeshData(
        name=name,
        v=np.asarray(v, dtype=F32).reshape(-1, 3),
        f=np.asarray(f, dtype=np.int32).reshape(-1, 3),
        uv=None if uv is None else np.asarray(uv, dtype=F32).reshape(-1, 2),
        fuv=None if fuv is None else np.asarray(fuv, dtype=np.int32).reshape(-1, 3),
        to_world=np.eye(4, dtype=F32) if to_world is None else np.asarray(to_world, dtype=F32),
        bsdf=bsdf,
        emitter=emitter,
    )


def cbox_meshes() -> List[MeshData]:
    """The 8 Cornell-box meshes (36 triangles), in the README's add_Mesh order."""
    luminaire = _m(
        "luminaire",
        [(343, 540.79999, 227), (343, 540.79999, 332), (213, 540.79999, 332), (213, 540.79999, 227)],
        _quad(0, 1, 2, 3), "light", to_world=translate(0.0, -0.5, 0.0), emitter=(20.0, 20.0, 8.0))
    smallbox = _m(
        "smallbox",
        [(130, 165, 65), (82, 165, 225), (240, 165, 272), (290, 165, 114),
         (290, 0, 114), (240, 0, 272), (130, 0, 65), (82, 0, 225)],
        [(0, 1, 2), (0, 2, 3), (4, 3, 2), (4, 2, 5), (6, 0, 3), (6, 3, 4),
         (7, 1, 0), (7, 0, 6), (5, 2, 1), (5, 1, 7), (4, 5, 7), (4, 7, 6)], "cat")
    largebox = _m(
        "largebox",
        [(423.0, 329.999969, 247.000061), (265.0, 329.999939, 296.000061),
         (314.0, 329.999939, 456.000061), (472.0, 329.999939, 406.000061),
         (423.0, -0.000040, 247.0), (472.0, -0.000066, 406.0),
         (314.0, -0.000074, 456.0), (265.0, -0.000048, 296.0)],
        [(0, 1, 2), (0, 2, 3), (4, 0, 3), (4, 3, 5), (5, 3, 2), (5, 2, 6),
         (6, 2, 1), (6, 1, 7), (7, 1, 0), (7, 0, 4), (5, 6, 7), (5, 7, 4)], "cat",
        uv=[(0.994838, 0.753967), (0.663615, 0.753540), (0.663030, 0.500000), (0.994838, 0.501892),
            (0.335054, 0.500000), (0.335054, 0.000000), (0.668172, 0.000000), (0.668172, 0.500000),
            (1.000000, 0.000000), (1.000000, 0.500000), (0.335054, 0.000000), (0.335054, 0.500000),
            (0.000000, 0.500000), (0.000000, 0.000000), (0.000000, 1.000000), (0.331223, 0.500000),
            (0.331223, 1.000000), (0.663030, 0.752075), (0.331223, 0.753998), (0.331808, 0.500458),
            (0.663030, 0.500000)],
        fuv=[(0, 1, 2), (0, 2, 3), (4, 5, 6), (4, 6, 7), (7, 6, 8), (7, 8, 9),
             (10, 11, 12), (10, 12, 13), (14, 12, 15), (14, 15, 16), (17, 18, 19), (17, 19, 20)])
    floor = _m(
        "floor",
        [(552.79999, 0, 0), (0, 0, 0), (0, 0, 559.20001), (549.59998, 0, 559.20001)],
        _quad(0, 1, 2, 3), "white")
    ceiling = _m(
        "ceiling",
        [(556, 548.79999, 0), (556, 548.79999, 559.20001), (0, 548.79999, 559.20001), (0, 548.79999, 0)],
        _quad(0, 1, 2, 3), "white")
    back = _m(
        "back",
        [(549.599976, -0.000091, 559.200012), (0.0, -0.000091, 559.200012),
         (0.0, 548.799927, 559.200073), (556.0, 548.799927, 559.200073)],
        _quad(0, 1, 2, 3), "white",
        uv=[(0.987061, 0.011536), (0.987061, 1.0), (0.0, 1.0), (0.0, 0.0)],
        fuv=_quad(0, 1, 2, 3))
    greenwall = _m(
        "greenwall",
        [(0, 0, 559.20001), (0, 0, 0), (0, 548.79999, 0), (0, 548.79999, 559.20001)],
        _quad(0, 1, 2, 3), "green")
    redwall = _m(
        "redwall",
        [(552.79999, 0, 0), (549.59998, 0, 559.20001), (556, 548.79999, 559.20001), (556, 548.79999, 0)],
        _quad(0, 1, 2, 3), "red")
    return [luminaire, smallbox, largebox, floor, ceiling, back, greenwall, redwall]


# (name, reflectance) in the README's add_BSDF order; "cat" is DiffuseBSDF() = 0.5 grey
CBOX_BSDFS = [
    ("light", (0.0, 0.0, 0.0)),
    ("cat", (0.5, 0.5, 0.5)),
    ("white", (0.95, 0.95, 0.95)),
    ("green", (0.20, 0.90, 0.20)),
    ("red", (0.90, 0.20, 0.20)),
]

# The same scene with MicrofacetBSDF(specular, diffuse, roughness) materials (BASELINE config 3 uses
# ([.2,.9,.9], [.01,.01,.01], 0.3) on every non-light mesh; the walls get a rougher, more diffuse one so that
# both lobes matter).  Entries with 3 fields are Microfacet, with 1 field Diffuse.
CBOX_MF_BSDFS = [
    ("light", (0.0, 0.0, 0.0)),
    ("cat", ((0.2, 0.9, 0.9), (0.01, 0.01, 0.01), 0.3)),
    ("white", ((0.04, 0.04, 0.04), (0.7, 0.7, 0.7), 0.5)),
    ("green", ((0.1, 0.3, 0.1), (0.10, 0.60, 0.10), 0.4)),
    ("red", ((0.3, 0.1, 0.1), (0.60, 0.10, 0.10), 0.6)),
]

CBOX_CAMERA = dict(fov=60.0, near=1e-6, far=1e7, to_world=translate(278.0, 273.0, -800.0))


def scaled_cbox(scale: float):
    """The Cornell box shrunk by `scale` (vertices, the luminaire's offset and the camera position).  At scale 0.01 the
    coordinates are ~5 and fp32 reconstruction errors (~1e-6) sit three orders below the reference's fixed epsilons
    (RayEpsilon = ShadowEpsilon = 1e-3), so epsilon-band decisions that flip between implementations at full scale
    (DESIGN.md "parity") are excluded by construction.  Returns (meshes, camera)."""
    ms = cbox_meshes()
    for m in ms:
        m.v = (m.v.astype(np.float64) * scale).astype(F32)
        m.to_world = m.to_world.copy()
        m.to_world[:3, 3] *= F32(scale)
    cam = dict(CBOX_CAMERA)
    cam["to_world"] = translate(278.0 * scale, 273.0 * scale, -800.0 * scale)
    return ms, cam


def icosphere(subdiv: int, radius: float, center) -> MeshData:
    """A closed smooth-shaded sphere (used for curved-silhouette test scenes)."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t),
         (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4),
         (11, 10, 2), (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8),
         (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = [np.asarray(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdiv):
        cache = {}
        nf = []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[key] = len(v) - 1
            return cache[key]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    vv = np.asarray(v) * radius + np.asarray(center, dtype=np.float64)
    return _m("sphere", vv.astype(F32), f, "cat")


def load_obj(path: str):
    """Minimal OBJ reader: v / vt / f, polygons fan-triangulated (a,b,c),(a,c,d),..."""
    v, vt, f, ft = [], [], [], []
    with open(path) as fh:
        for line in fh:
            tok = line.split()
            if not tok or tok[0].startswith("#"):
                continue
            if tok[0] == "v":
                v.append([float(x) for x in tok[1:4]])
            elif tok[0] == "vt":
                vt.append([float(x) for x in tok[1:3]])
            elif tok[0] == "f":
                vi, ti = [], []
                for c in tok[1:]:
                    parts = c.split("/")
                    a = int(parts[0])
                    vi.append(a - 1 if a > 0 else len(v) + a)
                    if len(parts) > 1 and parts[1]:
                        b = int(parts[1])
                        ti.append(b - 1 if b > 0 else len(vt) + b)
                for k in range(1, len(vi) - 1):
                    f.append([vi[0], vi[k], vi[k + 1]])
                    if len(ti) == len(vi):
                        ft.append([ti[0], ti[k], ti[k + 1]])
    has_uv = len(vt) > 0 and len(ft) == len(f)
    return (np.asarray(v, np.float32), np.asarray(f, np.int32), np.asarray(vt, np.float32) if has_uv else None,
            np.asarray(ft, np.int32) if has_uv else None)


def write_obj(mesh: MeshData, path: str) -> None:
    """Write a triangle-only OBJ (float32 round-trip exact) for consumers that load files."""
    with open(path, "w") as fh:
        for p in mesh.v:
            fh.write("v %.9g %.9g %.9g\n" % (float(p[0]), float(p[1]), float(p[2])))
        if mesh.uv is not None:
            for t in mesh.uv:
                fh.write("vt %.9g %.9g\n" % (float(t[0]), float(t[1])))
        for i, tri in enumerate(mesh.f):
            if mesh.uv is not None:
                tu = mesh.fuv[i]
                fh.write("f %d/%d %d/%d %d/%d\n" % (tri[0] + 1, tu[0] + 1, tri[1] + 1, tu[1] + 1, tri[2] + 1, tu[2] + 1))
            else:
                fh.write("f %d %d %d\n" % (tri[0] + 1, tri[1] + 1, tri[2] + 1))


def write_cbox_objs(dirname: str) -> List[str]:
    os.makedirs(dirname, exist_ok=True)
    paths = []
    for m in cbox_meshes():
        p = os.path.join(dirname, "cbox_%s.obj" % m.name)
        write_obj(m, p)
        paths.append(p)
    return paths
