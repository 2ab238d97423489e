"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the CPU oracle (oracle/psdr_oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (psdr_jit_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f = C.POINTER(C.c_float)
_i = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libpsdr_oracle.so")
    srcs = [os.path.join(_HERE, n) for n in ("psdr_oracle.cpp", "orc_math.h")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int] * 5
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_error.restype = C.c_char_p
        L.orc_error.argtypes = [C.c_void_p]
        L.orc_set_li_order.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_mis.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_collocated.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int]
        L.orc_add_diffuse.argtypes = [C.c_void_p, _f, _f, C.c_int]
        L.orc_add_microfacet.argtypes = [C.c_void_p, _f, _f, C.c_float, _f, C.c_int]
        L.orc_aov_d.argtypes = [C.c_void_p, C.c_int, C.c_int, _f, _f]
        L.orc_field_edges.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, _f]
        L.orc_add_roughconductor.argtypes = [C.c_void_p, C.c_float, _f, _f, _f, _f, C.c_int]
        L.orc_add_roughdielectric.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int]
        L.orc_add_microfacet_pervertex.argtypes = [C.c_void_p, _f, _f, C.c_int, C.c_int]
        L.orc_add_normalmap.argtypes = [C.c_void_p, _f, _f, C.c_int, C.c_int]
        L.orc_add_envmap.argtypes = [C.c_void_p, _f, _f, C.c_int, C.c_int, _f, _f, C.c_float, C.c_float]
        L.orc_set_bsdf_texture.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _f, _f]
        L.orc_set_bsdf_texture_slot.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, _f, _f, _f, _f]
        L.orc_add_mesh.argtypes = [C.c_void_p, _f, _f, C.c_int, _i, C.c_int, _f, C.c_int, _i, _f, _f, C.c_int, _f, _f, C.c_int, C.c_int]
        L.orc_add_camera.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, _f, _f]
        L.orc_add_camera_intrinsic.argtypes = [C.c_void_p] + [C.c_float] * 6 + [_f, _f]
        L.orc_add_camera_orthographic.argtypes = [C.c_void_p, C.c_float, C.c_float, _f, _f]
        L.orc_configure.argtypes = [C.c_void_p, _i, C.c_int]
        L.orc_num_primary_edges.argtypes = [C.c_void_p, C.c_int]
        L.orc_num_secondary_edges.argtypes = [C.c_void_p]
        L.orc_num_mesh_edges.argtypes = [C.c_void_p, C.c_int]
        L.orc_mesh_edges.argtypes = [C.c_void_p, C.c_int, _i]
        L.orc_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i, _i, C.c_int, _f, _f, _f]
        L.orc_render_diag.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, _i, _f, _f, _f]
        L.orc_preprocess_secondary_edges.argtypes = [C.c_void_p, C.c_int, _i, C.c_int, C.c_int, _f]
        L.orc_sampler_draws.argtypes = [C.c_int64, C.c_int, C.c_int, _f]
        L.orc_pmf_sample.argtypes = [_f, C.c_int, _f, C.c_int, _i, _f, _f]
        L.orc_aov.argtypes = [C.c_void_p, C.c_int, C.c_int, _f]
        L.orc_num_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _fp(a):
    return None if a is None else a.ctypes.data_as(_f)


def _ip(a):
    return None if a is None else a.ctypes.data_as(_i)


def _f32(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    return a if shape is None else a.reshape(shape)


def _mats(m):
    """3x(4x4): (left, raw, right); accepts a single 4x4 (= raw) or a dict."""
    out = np.tile(np.eye(4, dtype=np.float32), (3, 1, 1))
    if m is None:
        return out
    if isinstance(m, dict):
        for k, name in enumerate(("left", "raw", "right")):
            if name in m and m[name] is not None:
                out[k] = np.asarray(m[name], dtype=np.float32)
        return out
    out[1] = np.asarray(m, dtype=np.float32)
    return out


def _dmats(m):
    out = np.zeros((3, 4, 4), dtype=np.float32)
    if m is None:
        return None
    if isinstance(m, dict):
        for k, name in enumerate(("left", "raw", "right")):
            if name in m and m[name] is not None:
                out[k] = np.asarray(m[name], dtype=np.float32)
        return out
    out[1] = np.asarray(m, dtype=np.float32)
    return out


class OracleScene:
    def __init__(self, width, height, spp, sppe=0, sppse=0):
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_create(width, height, spp, sppe, sppse))
        self.width, self.height, self.spp, self.sppe, self.sppse = width, height, spp, sppe, sppse
        self._keep = []
        self.bsdf_ids = {}

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    def add_diffuse(self, name, refl, d_refl=None, two_side=False):
        r, dr = _f32(refl), _f32(d_refl)
        idx = self.L.orc_add_diffuse(self.h, _fp(r), _fp(dr), int(two_side))
        self.bsdf_ids[name] = idx
        return idx

    def add_microfacet(self, name, spec, diff, rough, d=None, two_side=False):
        """d = (d_spec[3], d_diff[3], d_rough) flattened to 7 floats, or None"""
        sp, df, dd = _f32(spec), _f32(diff), _f32(d)
        idx = self.L.orc_add_microfacet(self.h, _fp(sp), _fp(df), float(rough), _fp(dd), int(two_side))
        self.bsdf_ids[name] = idx
        return idx

    def add_roughconductor(self, name, alpha, eta, k, spec=(1.0, 1.0, 1.0), d=None, two_side=False):
        """RoughConductorBSDF(alpha, eta, k) (reference src/bsdf/roughconductor.cpp); d = (d_alpha, d_eta[3], d_k[3],
        d_spec[3]) flattened to 10 floats, or None"""
        idx = self.L.orc_add_roughconductor(self.h, float(alpha), _fp(_f32(eta)), _fp(_f32(k)), _fp(_f32(spec)), _fp(_f32(d)), int(two_side))
        self.bsdf_ids[name] = idx
        return idx

    def add_roughdielectric(self, name, alpha, int_ior, ext_ior, d_alpha=0.0, two_side=False):
        """RoughDielectricBSDF(alpha, intIOR, extIOR) (reference src/bsdf/roughdielectric.cpp)"""
        idx = self.L.orc_add_roughdielectric(self.h, float(alpha), float(d_alpha), float(int_ior), float(ext_ior), int(two_side))
        self.bsdf_ids[name] = idx
        return idx

    def add_microfacet_pervertex(self, name, spec, diff, rough, d=None, two_side=False):
        """MicrofacetBSDFPerVertex(spec[n,3], diff[n,3], rough[n]) (reference src/bsdf/microfacet_pv.cpp);
        d = tangents as an [n,7] table (specular rgb, diffuse rgb, roughness) or None"""
        sp, df, rg = _f32(spec).reshape(-1, 3), _f32(diff).reshape(-1, 3), _f32(rough).reshape(-1, 1)
        pv = np.ascontiguousarray(np.concatenate([sp, df, rg], axis=1), np.float32)
        dd = None if d is None else np.ascontiguousarray(_f32(d).reshape(-1, 7))
        idx = self.L.orc_add_microfacet_pervertex(self.h, _fp(pv), _fp(dd), int(pv.shape[0]), int(two_side))
        self.bsdf_ids[name] = idx
        return idx

    def add_normalmap(self, name, nested_name, normal=(0.499999, 0.499999, 1.0), d_normal=None, two_side=False):
        """NormalMapBSDF around the BSDF added before under `nested_name` (reference src/bsdf/normalmap.cpp,
        src/scene/scene.cpp:128-145); a normal-map bitmap goes through set_bsdf_texture(name, ..., slot=0)"""
        idx = self.L.orc_add_normalmap(self.h, _fp(_f32(normal)), _fp(_f32(d_normal)), int(self.bsdf_ids[nested_name]), int(two_side))
        self.bsdf_ids[name] = idx
        return idx

    def set_bsdf_texture(self, name, data, w, h, d_data=None, slot=0, xform=None, d_xform=None):
        """slot 0 reflectance / diffuseReflectance, 1 specularReflectance, 2 roughness (1 channel);
        xform = (scale, rotate, translate.x, translate.y) of the bitmap's uv transform, d_xform its tangent"""
        dat, dd = _f32(data).reshape(-1), (None if d_data is None else _f32(d_data).reshape(-1))
        xf, dxf = (None if xform is None else _f32(xform, (4,))), (None if d_xform is None else _f32(d_xform, (4,)))
        return self.L.orc_set_bsdf_texture_slot(self.h, self.bsdf_ids[name], int(slot), int(w), int(h), _fp(dat), _fp(dd), _fp(xf), _fp(dxf))

    def add_envmap(self, data, w, h, to_world=None, scale=1.0, d_data=None, d_to_world_left=None, d_scale=0.0):
        """data: [h*w, 3]; to_world = raw 4x4 (left = identity); d_to_world_left = tangent of the left factor"""
        dat, dd = _f32(data).reshape(-1), (None if d_data is None else _f32(d_data).reshape(-1))
        tw = np.tile(np.eye(4, dtype=np.float32), (2, 1, 1))
        if to_world is not None:
            tw[1] = np.asarray(to_world, dtype=np.float32)
        dtw = None
        if d_to_world_left is not None:
            dtw = np.zeros((2, 4, 4), dtype=np.float32)
            dtw[0] = np.asarray(d_to_world_left, dtype=np.float32)
        return self.L.orc_add_envmap(self.h, _fp(dat), _fp(dd), int(w), int(h), _fp(tw), _fp(dtw), float(scale), float(d_scale))

    def add_mesh(self, v, f, bsdf, uv=None, fuv=None, to_world=None, d_to_world=None, dv=None, radiance=None,
                 d_radiance=None, use_face_normals=False, enable_edges=True):
        v = _f32(v, (-1, 3))
        f = np.ascontiguousarray(np.asarray(f, dtype=np.int32).reshape(-1, 3))
        dv = _f32(dv, (-1, 3))
        uv_ = _f32(uv, (-1, 2)) if uv is not None else None
        fuv_ = np.ascontiguousarray(np.asarray(fuv, dtype=np.int32).reshape(-1, 3)) if fuv is not None else None
        tw, dtw = _mats(to_world), _dmats(d_to_world)
        rad, drad = _f32(radiance), _f32(d_radiance)
        b = self.bsdf_ids[bsdf] if isinstance(bsdf, str) else int(bsdf)
        return self.L.orc_add_mesh(self.h, _fp(v), _fp(dv), len(v), _ip(f), len(f), _fp(uv_), 0 if uv_ is None else len(uv_),
                                   _ip(fuv_), _fp(tw), _fp(dtw), b, _fp(rad), _fp(drad), int(use_face_normals), int(enable_edges))

    def add_camera(self, fov, near, far, to_world, d_to_world=None):
        tw, dtw = _mats(to_world), _dmats(d_to_world)
        return self.L.orc_add_camera(self.h, fov, near, far, _fp(tw), _fp(dtw))

    def add_camera_intrinsic(self, fx, fy, cx, cy, near, far, to_world, d_to_world=None):
        """PerspectiveCamera(fx, fy, cx, cy, near, far) (reference include/psdr/sensor/perspective.h:11-12)"""
        tw, dtw = _mats(to_world), _dmats(d_to_world)
        return self.L.orc_add_camera_intrinsic(self.h, fx, fy, cx, cy, near, far, _fp(tw), _fp(dtw))

    def add_camera_orthographic(self, near, far, to_world, d_to_world=None):
        """OrthographicCamera(near, far) (reference src/sensor/orthographic.cpp)"""
        tw, dtw = _mats(to_world), _dmats(d_to_world)
        return self.L.orc_add_camera_orthographic(self.h, near, far, _fp(tw), _fp(dtw))

    def configure(self, active=(0,)):
        a = np.asarray(list(active), dtype=np.int32)
        rc = self.L.orc_configure(self.h, _ip(a), len(a))
        if rc:
            raise RuntimeError(self.L.orc_error(self.h).decode())

    def set_collocated(self, intensity, d_intensity=0.0, bsdf_field=False):
        """CollocatedIntegrator(intensity) (reference src/integrator/collocated.cpp); render with any depth.
        bsdf_field: FieldExtractionIntegrator("bsdf") (src/integrator/field.cpp:72-92), the BSDF term alone"""
        self.L.orc_set_collocated(self.h, float(intensity), float(d_intensity), int(bool(bsdf_field)))

    def set_mis(self, mis=2):
        """2: PathTracer / Direct(2); 0 / 1: Direct(0) / Direct(1) -- render with depth 1"""
        self.L.orc_set_mis(self.h, int(mis))

    def set_li_order(self, p_first=True):
        self.L.orc_set_li_order(self.h, int(p_first))

    def num_primary_edges(self, sensor=0):
        return self.L.orc_num_primary_edges(self.h, sensor)

    def num_secondary_edges(self):
        return self.L.orc_num_secondary_edges(self.h)

    def mesh_edges(self, mesh):
        n = self.L.orc_num_mesh_edges(self.h, mesh)
        out = np.zeros((4, n), dtype=np.int32)
        self.L.orc_mesh_edges(self.h, mesh, _ip(out))
        return out

    def render(self, depth, seed=0, mode=1, terms=7, sensor=0, hide_emitters=False, skip=None, pix_id=None, lane_out=False):
        if pix_id is not None:
            pix = np.ascontiguousarray(np.asarray(pix_id, dtype=np.int32))
            npix = len(pix)
        else:
            pix, npix = None, self.width * self.height
        img = np.zeros((npix, 3), dtype=np.float32)
        dimg = np.zeros((npix, 3), dtype=np.float32) if mode == 1 else None
        lanes = np.zeros((npix * self.spp, 3), dtype=np.float32) if lane_out else None
        sk = None if skip is None else np.asarray(skip, dtype=np.int32)
        rc = self.L.orc_render(self.h, sensor, depth, seed, mode, terms, int(hide_emitters), _ip(sk), _ip(pix),
                               0 if pix is None else npix, _fp(img), _fp(dimg), _fp(lanes))
        if rc:
            raise RuntimeError(self.L.orc_error(self.h).decode())
        if lane_out:
            return img, dimg, lanes
        return (img, dimg) if mode == 1 else img

    def render_diag(self, depth, seed=0, mode=0, sensor=0, skip=None):
        """interior term + per-lane decision margins [N0, 4] = (shadow, border, self, tie); see LaneDiag in psdr_oracle.cpp"""
        n = self.width * self.height
        img = np.zeros((n, 3), dtype=np.float32)
        lanes = np.zeros((n * self.spp, 3), dtype=np.float32)
        diag = np.zeros((n * self.spp, 4), dtype=np.float32)
        sk = None if skip is None else np.asarray(skip, dtype=np.int32)
        rc = self.L.orc_render_diag(self.h, sensor, depth, seed, mode, _ip(sk), _fp(img), _fp(lanes), _fp(diag))
        if rc:
            raise RuntimeError(self.L.orc_error(self.h).decode())
        return img, lanes, diag

    def preprocess_secondary_edges(self, sensor, reso, nrounds=1, seed=0):
        r = np.asarray(reso, dtype=np.int32)
        mass = np.zeros(int(r[0]) * int(r[1]) * int(r[2]), dtype=np.float32)
        rc = self.L.orc_preprocess_secondary_edges(self.h, sensor, _ip(r), nrounds, seed, _fp(mass))
        if rc:
            raise RuntimeError(self.L.orc_error(self.h).decode())
        return mass

    def aov(self, sensor=0, seed=0):
        out = np.zeros((self.width * self.height * self.spp, 14), dtype=np.float32)
        self.L.orc_aov(self.h, sensor, seed, _fp(out))
        return out

    FIELDS = ("segmentation", "silhouette", "position", "depth", "geoNormal", "shNormal", "uv")

    def field_render_d(self, field, sensor=0, seed=0, obj=-1, terms=3):
        """FieldExtractionIntegrator(field).renderD in forward mode (reference src/integrator/field.cpp through
        Integrator::renderD): (field image, derivative image); terms bit 0 interior, bit 1 primary edges."""
        n, spp = self.width * self.height, max(self.spp, 1)
        a = np.zeros((n * spp, 14), dtype=np.float32)
        da = np.zeros_like(a)
        self.L.orc_aov_d(self.h, sensor, seed, _fp(a), _fp(da))
        a, da = a.reshape(n, spp, 14), da.reshape(n, spp, 14)
        valid = a[:, :, 0] > 0
        if obj >= 0:
            valid &= a[:, :, 0] == float(obj + 1)

        def sel(x):
            return {"segmentation": np.repeat(x[:, :, 0:1], 3, axis=2) - 1.0, "silhouette": np.ones_like(x[:, :, 2:5]), "position": x[:, :, 2:5],
                    "depth": np.repeat(x[:, :, 5:6], 3, axis=2), "geoNormal": x[:, :, 6:9], "shNormal": x[:, :, 9:12],
                    "uv": np.concatenate((x[:, :, 12:14], np.zeros_like(x[:, :, 0:1])), axis=2)}[field]
        m = valid[:, :, None]
        img = (sel(a) * m).sum(axis=1) / float(spp)
        dimg = np.zeros_like(img)
        if (terms & 1) and field not in ("segmentation", "silhouette"):
            dimg += (sel(da) * m).sum(axis=1) / float(spp)
        if (terms & 2) and self.sppe > 0:
            e = np.zeros((n, 3), dtype=np.float32)
            self.L.orc_field_edges(self.h, sensor, seed, self.FIELDS.index(field), obj, _fp(e))
            dimg += e
        return img.astype(np.float32), dimg.astype(np.float32)


def sampler_draws(seed, n, ndraws):
    out = np.zeros((ndraws, n), dtype=np.float32)
    lib().orc_sampler_draws(seed, n, ndraws, _fp(out))
    return out


def pmf_sample(pmf, samples):
    pmf, samples = _f32(pmf), _f32(samples)
    idx = np.zeros(len(samples), dtype=np.int32)
    p = np.zeros(len(samples), dtype=np.float32)
    s = np.zeros(1, dtype=np.float32)
    lib().orc_pmf_sample(_fp(pmf), len(pmf), _fp(samples), len(samples), _ip(idx), _fp(p), _fp(s))
    return idx, p, float(s[0])


def num_threads():
    return lib().orc_num_threads()
