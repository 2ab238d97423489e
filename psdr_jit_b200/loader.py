"""Scene I/O either side of the render path: the Mitsuba-style XML scene description
(reference src/scene/scene_loader.cpp), OpenEXR bitmaps (src/core/bitmap_loader.cpp, tinyexr there) and OBJ dumps
(src/shape/mesh.cpp:469-554).  Host-side plumbing only: everything ends up in the same Scene / Mesh / BSDF objects the
hand-built scenes use.  EXR decoding goes through OpenCV when it is importable (the image has cv2 with OpenEXR)."""
from __future__ import annotations

import os
import xml.etree.ElementTree as ET

import numpy as np


def load_exr(path: str):
    """-> (width, height, data[h*w, 3] float32 RGB), rows top to bottom as stored (BitmapLoader::load_openexr_rgba)."""
    if not os.path.exists(path):
        raise RuntimeError("Failed to load EXR from: " + path)
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    try:
        import cv2
    except Exception as e:      # pragma: no cover
        raise RuntimeError("reading OpenEXR files needs OpenCV (cv2) built with OpenEXR: %s" % e)
    im = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if im is None:
        raise RuntimeError("Failed to load EXR from: " + path + " (is OPENCV_IO_ENABLE_OPENEXR=1 set before cv2 is imported?)")
    im = np.asarray(im, dtype=np.float32)
    if im.ndim == 2:
        im = np.repeat(im[:, :, None], 3, axis=2)
    rgb = im[:, :, 2::-1] if im.shape[2] >= 3 else np.repeat(im[:, :, :1], 3, axis=2)      # BGR(A) -> RGB
    h, w = rgb.shape[:2]
    return w, h, np.ascontiguousarray(rgb.reshape(h * w, 3))


def parse_vector(s: str, n: int, allow_empty: bool = False):
    """scene_loader.cpp:25-54: numbers separated by commas / blanks; a short vector repeats its last entry if allowed"""
    vals = [float(t) for t in s.replace(",", " ").split()]
    if len(vals) > n:
        raise RuntimeError("tot < length")
    if len(vals) < n:
        if not allow_empty:
            raise RuntimeError("Vector too short: [" + s + "]")
        vals += [vals[-1] if vals else 0.0] * (n - len(vals))
    return np.asarray(vals, dtype=np.float32)


def _translate(v):
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = v
    return m


def _scale(v):
    return np.diag(np.asarray([v[0], v[1], v[2], 1.0], dtype=np.float32))


def _rotate(axis, angle_deg):
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    t = np.deg2rad(angle_deg)
    c, s = np.cos(t), np.sin(t)
    x, y, z = a
    r = np.array([[c + x * x * (1 - c), x * y * (1 - c) - z * s, x * z * (1 - c) + y * s, 0],
                  [y * x * (1 - c) + z * s, c + y * y * (1 - c), y * z * (1 - c) - x * s, 0],
                  [z * x * (1 - c) - y * s, z * y * (1 - c) + x * s, c + z * z * (1 - c), 0], [0, 0, 0, 1]])
    return r.astype(np.float32)


def _look_at(origin, target, up):
    """include/psdr/core/transform.h:85-103: columns (left, new_up, dir, origin)"""
    o, t, u = (np.asarray(x, dtype=np.float64) for x in (origin, target, up))
    d = (t - o) / np.linalg.norm(t - o)
    left = np.cross(u, d)
    left /= np.linalg.norm(left)
    nu = np.cross(d, left)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = left, nu, d, o
    return m.astype(np.float32)


def load_transform(node) -> np.ndarray:
    """scene_loader.cpp:83-129: children applied in document order, each multiplied from the LEFT"""
    result = np.eye(4, dtype=np.float32)
    if node is None:
        return result
    name = node.get("name", "")
    if name not in ("to_world", "toWorld"):
        raise RuntimeError("Invalid transformation name: " + name)
    for c in node:
        if c.tag == "translate":
            m = _translate([float(c.get(k, 0.0)) for k in "xyz"])
        elif c.tag == "rotate":
            m = _rotate([float(c.get(k, 0.0)) for k in "xyz"], float(c.get("angle", 0.0)))
        elif c.tag == "scale":
            m = _scale([float(c.get(k, 1.0)) for k in "xyz"])
        elif c.tag in ("look_at", "lookAt", "lookat"):
            m = _look_at(parse_vector(c.get("origin"), 3), parse_vector(c.get("target"), 3), parse_vector(c.get("up"), 3))
        elif c.tag == "matrix":
            m = parse_vector(c.get("value"), 16).reshape(4, 4)       # row-major text (the reference transposes a column-major load)
        else:
            raise RuntimeError("Unsupported transformation: " + c.tag)
        result = m @ result
    return result.astype(np.float32)


def _child_by_name(node, names, allow_empty=False):
    for c in node:
        if c.get("name") in names:
            return c
    if not allow_empty:
        raise RuntimeError("Missing child node: " + sorted(names)[0])
    return None


def load_scene_xml(scene, root, base_dir: str = "."):
    """SceneLoader::load_scene (scene_loader.cpp:203-236): sensors, BSDFs, environment emitter, shapes -- in that order."""
    import psdr_jit_b200 as psdr
    if root.tag != "scene":
        root = root.find("scene")
        if root is None:
            raise RuntimeError("XML parsing failed")
    path = lambda p: p if os.path.isabs(p) else os.path.join(base_dir, p)      # noqa: E731

    def texture(node, channels):
        if node.tag == "texture":
            if node.get("type") != "bitmap":
                raise RuntimeError("Unsupported texture type: " + str(node.get("type")))
            fn = node.find("string")
            if fn is None or fn.get("name") != "filename" or not fn.get("value"):
                raise RuntimeError("Failed to retrieve bitmap filename")
            return (psdr.Bitmap1fD if channels == 1 else psdr.Bitmap3fD)(path(fn.get("value")))
        if channels == 1:
            return np.float32(float(node.get("value")))
        if node.tag == "float":
            return np.full(3, float(node.get("value")), np.float32)
        if node.tag in ("rgb", "spectrum"):
            return parse_vector(node.get("value"), 3, True)
        raise RuntimeError("Unsupported RGB type: " + node.tag)

    first = len(scene._sensors) == 0
    for node in root.findall("sensor"):
        film, sampler = node.find("film"), node.find("sampler")
        if first:
            if film is None:
                raise RuntimeError("Missing film node")
            if sampler is None:
                raise RuntimeError("Missing sampler node")
            o = scene.opts
            o.width = int(_child_by_name(film, {"width"}).get("value"))
            o.height = int(_child_by_name(film, {"height"}).get("value"))
            o.spp = int(sampler.find("integer").get("value"))
            o.sppe = o.sppse = 0
            first = False
        elif film is not None or sampler is not None:
            raise RuntimeError("Duplicate film node" if film is not None else "Duplicate sampler node")
        if node.get("type") != "perspective":
            raise RuntimeError("Unsupported sensor: " + str(node.get("type")))
        fov = float(_child_by_name(node, {"fov"}).get("value"))
        axis = _child_by_name(node, {"fov_axis", "fovAxis"}, True)
        if axis is not None and axis.get("value") != "x":
            raise RuntimeError("Unsupported fov-axis: " + str(axis.get("value")))
        near = _child_by_name(node, {"near_clip", "nearClip"}, True)
        far = _child_by_name(node, {"far_clip", "farClip"}, True)
        s = psdr.PerspectiveCamera(fov, float(near.get("value")) if near is not None else 0.1, float(far.get("value")) if far is not None else 1e4)
        s.to_world = load_transform(node.find("transform"))
        scene.add_Sensor(s)
    for node in root.findall("bsdf"):
        bid = node.get("id")
        if not bid:
            raise RuntimeError("BSDF must have an id")
        t = node.get("type")

        def leaf(n, t, what="Unsupported BSDF: "):
            if t == "diffuse":
                return psdr.DiffuseBSDF(texture(_child_by_name(n, {"reflectance"}), 3))
            if t == "microfacet":
                return psdr.MicrofacetBSDF(texture(_child_by_name(n, {"specular_reflectance", "specularReflectance"}), 3),
                                           texture(_child_by_name(n, {"diffuse_reflectance", "diffuseReflectance"}), 3),
                                           texture(_child_by_name(n, {"roughness"}), 1))
            if t == "roughconductor":      # scene_loader.cpp:334-345: alpha (-> alpha_u = alpha_v), eta, k
                return psdr.RoughConductorBSDF(texture(_child_by_name(n, {"alpha"}), 1), texture(_child_by_name(n, {"eta"}), 3),
                                               texture(_child_by_name(n, {"k"}), 3))
            if t == "roughdielectric":     # scene_loader.cpp:346-360: alpha, intIOR, extIOR
                return psdr.RoughDielectricBSDF(texture(_child_by_name(n, {"alpha"}), 1), float(_child_by_name(n, {"intIOR"}).get("value")),
                                                float(_child_by_name(n, {"extIOR"}).get("value")))
            raise RuntimeError(what + str(t))

        if t == "normalmap":               # scene_loader.cpp:372-424: <normalmap> texture + one nested <bsdf>
            inner = node.find("bsdf")
            if inner is None:
                raise RuntimeError("Unsupported normal map nested BSDF: ")
            nm = psdr.NormalMapBSDF(texture(_child_by_name(node, {"normalmap"}), 3))
            scene.add_normalmap_BSDF(nm, leaf(inner, inner.get("type"), "Unsupported normal map nested BSDF: "), bid)
            continue
        b = leaf(node, t)
        scene.add_BSDF(b, bid)
    for node in root.findall("emitter"):
        if node.get("type") != "envmap":
            raise RuntimeError("Unsupported emitter: " + str(node.get("type")))
        fn = node.find("string")
        if fn is None or fn.get("name") != "filename" or not fn.get("value"):
            raise RuntimeError("Failed to retrieve bitmap filename")
        sc_node = _child_by_name(node, {"scale"}, True)
        scene.add_EnvironmentMap(path(fn.get("value")), load_transform(node.find("transform")), float(sc_node.get("value")) if sc_node is not None else 1.0)
    for node in root.findall("shape"):
        if node.get("type") != "obj":
            raise RuntimeError("Unsupported shape: " + str(node.get("type")))
        fn = node.find("string")
        if fn is None or fn.get("name") != "filename":
            raise RuntimeError('strcmp(name_node.attribute("name").value(), "filename") == 0')
        ref = node.find("ref")
        if ref is None:
            raise RuntimeError("Missing BSDF reference")
        if node.find("bsdf") is not None:
            raise RuntimeError("BSDFs declared under shapes are not supported.")
        mesh = psdr.Mesh()
        mesh.load(path(fn.get("value")))
        mesh.id = node.get("id") or ""
        fnn = _child_by_name(node, {"face_normals", "faceNormals"}, True)
        mesh.use_face_normal = fnn is not None and fnn.get("value") == "true"
        mesh.to_world = load_transform(node.find("transform"))
        em = node.find("emitter")
        light = None
        if em is not None:
            if em.get("type") != "area":
                raise RuntimeError("Unsupported emitter: " + str(em.get("type")))
            light = psdr.AreaLight(texture(_child_by_name(em, {"radiance"}), 3))
        scene.add_Mesh(mesh, ref.get("id"), light)


def dump_obj(mesh, fname: str, normals=None):
    """Mesh::dump (mesh.cpp:469-554): v (+ vn per vertex unless face normals are used), vt, f v/vt/vn"""
    v, f = np.asarray(mesh.vertex_positions, np.float32), np.asarray(mesh.face_indices, np.int32)
    with open(fname, "wt") as out:
        for i in range(len(v)):
            out.write("v %.6e %.6e %.6e\n" % tuple(float(x) for x in v[i]))
            if normals is not None:
                out.write("vn %.6e %.6e %.6e\n" % tuple(float(x) for x in normals[i]))
        if mesh.vertex_uv is not None:
            for t in np.asarray(mesh.vertex_uv, np.float32):
                out.write("vt %.6e %.6e\n" % (float(t[0]), float(t[1])))
            fu = np.asarray(mesh.face_uv_indices, np.int32)
            for i in range(len(f)):
                a = [int(x) + 1 for x in f[i]]
                u = [int(x) + 1 for x in fu[i]]
                if normals is None:
                    out.write("f %d/%d %d/%d %d/%d\n" % (a[0], u[0], a[1], u[1], a[2], u[2]))
                else:
                    out.write("f %d/%d/%d %d/%d/%d %d/%d/%d\n" % (a[0], u[0], a[0], a[1], u[1], a[1], a[2], u[2], a[2]))
        else:
            for i in range(len(f)):
                a = [int(x) + 1 for x in f[i]]
                if normals is None:
                    out.write("f %d %d %d\n" % tuple(a))
                else:
                    out.write("f %d//%d %d//%d %d//%d\n" % (a[0], a[0], a[1], a[1], a[2], a[2]))


def vertex_normals(v, f):
    """area-weighted vertex normals as Mesh::process_mesh (mesh.cpp:23-62): normalise(sum of unnormalised face normals)"""
    v, f = np.asarray(v, np.float64), np.asarray(f, np.int64)
    fn = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    n = np.zeros_like(v)
    for k in range(3):
        np.add.at(n, f[:, k], fn)
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    return (n / np.where(ln > 0, ln, 1.0)).astype(np.float32)
