"""psdr_jit_b200 -- B200-native drop-in for the integrator path of psdr_jit.

The Python surface mirrors the reference module (reference src/psdr.cpp:100-441):
``Scene``, ``RenderOption``, ``PerspectiveCamera``, ``DiffuseBSDF``, ``AreaLight``, ``Mesh``,
``PathTracer(max_depth).renderC / renderD``, ``Scene.param_map``.  Arrays are numpy / torch instead
of ``drjit.cuda(.ad)`` types; results are torch CUDA tensors.  All numerics run in the C++/CUDA
library behind the C ABI of ``include/psdr_b200.h`` -- this file is host-side plumbing only.

Differences a reference user has to know (DESIGN.md has the rationale):
  * forward-mode derivatives: set tangents on the parameter objects (``d_*`` fields or the
    ``tangent=`` argument of ``set_transform``) and call ``renderD``; the derivative image is
    returned by ``renderD_fwd`` / kept in ``integrator.grad_image`` (the reference needs
    ``drjit.set_grad(P, 1); drjit.forward_to(img); drjit.grad(img)``).
  * ``PathTracer.reference_tangent_scaling = True`` reproduces the reference binary's scaling of the
    interior and secondary-edge tangents (2x); the default is the finite-difference-correct value.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib, scenes  # noqa: F401
from . import loader as _loader

__all__ = ["Scene", "RenderOption", "PerspectiveCamera", "OrthographicCamera", "DiffuseBSDF", "MicrofacetBSDF", "RoughConductorBSDF", "RoughDielectricBSDF", "MicrofacetBSDFPerVertex", "NormalMapBSDF", "AreaLight", "EnvironmentMap", "Bitmap3fD", "Bitmap1fD", "Mesh", "PathTracer", "CollocatedIntegrator", "Direct", "DirectIntegrator", "FieldExtractionIntegrator", "Sampler",
           "Integrator", "Object", "scenes", "kernel_launch_count"]


def _f32(a, shape=None) -> np.ndarray:
    if hasattr(a, "__psdr_value__"):   # Dr.Jit stand-in types (psdr_jit_b200.compat)
        a = a.__psdr_value__()
    elif hasattr(a, "detach"):   # torch tensor
        a = a.detach().cpu().numpy()
    if shape is not None and np.ndim(a) == 0:      # DiffuseBSDF(0.5): a scalar fills the vector
        a = np.full(shape, a, dtype=np.float32)
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    return a if shape is None else a.reshape(shape)


def _mat4(m) -> np.ndarray:
    if m is None:
        return np.eye(4, dtype=np.float32)
    m = _f32(m)
    if m.shape != (4, 4):
        raise RuntimeError("expected a 4x4 matrix")
    return m.copy()


def _fp(a):
    return None if a is None else a.ctypes.data_as(_lib.P_F)


def _ip(a):
    return None if a is None else a.ctypes.data_as(_lib.P_I)


def kernel_launch_count() -> int:
    return int(_lib.load().psdr_kernel_launch_count())


def set_cta_policy(policy: int = 0):
    """CTA shape of the term kernels: 0 = by launch size, 1 = 128-thread CTAs, 2 = large CTAs with block barriers
    (include/psdr_b200.h psdr_set_cta_policy).  Results do not depend on it."""
    _lib.check(_lib.load().psdr_set_cta_policy(int(policy)))


def set_edge_sort(bins: int = 512):
    """Lane order of the primary- and secondary-edge kernels: launches of 32768 lanes or more bucket their samples by position
    along the edge list (``bins`` buckets, 2..2048) so that a warp's rays start next to each other; 0 = lane order
    (include/psdr_b200.h psdr_set_edge_sort).  Results do not depend on it."""
    _lib.check(_lib.load().psdr_set_edge_sort(int(bins)))


class Object:
    """reference include/psdr/object.h"""
    id = ""

    def type_name(self) -> str:
        return type(self).__name__

    def __repr__(self) -> str:
        return "%s[id=%s]" % (type(self).__name__, self.id)


class RenderOption:
    """reference include/psdr/types.h:217-228"""

    def __init__(self, *args):
        """RenderOption() | (w, h, spp) | (w, h, spp, sppe) | (w, h, spp, sppe, sppse); the 3-argument form sets
        sppe = sppse = spp, the default constructor leaves both boundary terms off (types.h:218-221)."""
        if len(args) not in (0, 3, 4, 5):
            raise TypeError("RenderOption(): expected 0, 3, 4 or 5 arguments")
        self.width, self.height, self.spp, self.sppe, self.sppse = 128, 128, 1, 0, 0
        if args:
            self.width, self.height, self.spp = int(args[0]), int(args[1]), int(args[2])
            self.sppe = int(args[3]) if len(args) > 3 else self.spp
            self.sppse = int(args[4]) if len(args) > 4 else self.sppe
        self.log_level = 1

    def __repr__(self):
        return "[width: %d, height: %d, spp: %d, sppe: %d, sppse: %d, log_level: %d]" % (
            self.width, self.height, self.spp, self.sppe, self.sppse, self.log_level)


class _Transformable(Object):
    """to_world = to_world_left * to_world * to_world_right (reference mesh.h:25-41, sensor.h:34-48)."""

    def __init__(self):
        self.to_world = np.eye(4, dtype=np.float32)
        self.to_world_left = np.eye(4, dtype=np.float32)
        self.to_world_right = np.eye(4, dtype=np.float32)
        self.d_to_world = np.zeros((4, 4), dtype=np.float32)
        self.d_to_world_left = np.zeros((4, 4), dtype=np.float32)
        self.d_to_world_right = np.zeros((4, 4), dtype=np.float32)

    def set_transform(self, mat, set_left: bool = True, tangent=None):
        # a matrix whose entries depend on AD leaves (Dr.Jit stand-in): remember d(matrix)/d(leaf) for forward_to
        if hasattr(mat, "__psdr_tangents__"):
            self.__dict__.setdefault("_leaf_tangents", {})["to_world_left" if set_left else "to_world_right"] = mat.__psdr_tangents__()
        t = np.zeros((4, 4), dtype=np.float32) if tangent is None else _f32(tangent, (4, 4)).copy()
        m = mat if hasattr(mat, "requires_grad") else _mat4(mat)     # torch tensors stay live for autograd
        if set_left:
            self.to_world_left, self.d_to_world_left = m, t
        else:
            self.to_world_right, self.d_to_world_right = m, t

    def append_transform(self, mat, append_left: bool = True):
        m = _mat4(mat)
        if append_left:
            self.d_to_world_left = m @ self.d_to_world_left
            self.to_world_left = m @ self.to_world_left
        else:
            self.d_to_world_right = self.d_to_world_right @ m
            self.to_world_right = self.to_world_right @ m

    def _copy_transform_from(self, other: "_Transformable"):
        for n in ("to_world", "to_world_left", "to_world_right", "d_to_world", "d_to_world_left", "d_to_world_right"):
            setattr(self, n, _f32(getattr(other, n), (4, 4)).copy())


class PerspectiveCamera(_Transformable):
    """reference src/psdr.cpp:365-375, src/sensor/perspective.cpp"""

    def __init__(self, *args):
        """PerspectiveCamera(fov_x, near, far) | PerspectiveCamera(fx, fy, cx, cy, near, far): the second form takes pinhole
        intrinsics in units of the image size (include/psdr/sensor/perspective.h:10-12, src/sensor/perspective.cpp:15-20)."""
        super().__init__()
        if len(args) == 3:
            self.intrinsics = None
            self.fov, self.near, self.far = (float(a) for a in args)
        elif len(args) == 6:
            self.intrinsics = tuple(float(a) for a in args[:4])
            self.fov, self.near, self.far = 0.0, float(args[4]), float(args[5])
        else:
            raise TypeError("PerspectiveCamera(): expected (fov_x, near, far) or (fx, fy, cx, cy, near, far)")

    def _clone(self):
        c = PerspectiveCamera(self.fov, self.near, self.far) if self.intrinsics is None else PerspectiveCamera(*self.intrinsics, self.near, self.far)
        c._copy_transform_from(self)
        return c


class OrthographicCamera(PerspectiveCamera):
    """reference src/psdr.cpp:375-383, src/sensor/orthographic.cpp: ``OrthographicCamera(near, far)``; rays leave the near
    plane along the camera's +z, the view volume is 2 x 2/aspect camera units."""

    def __init__(self, near: float, far: float):
        _Transformable.__init__(self)
        self.intrinsics = None
        self.fov, self.near, self.far = 0.0, float(near), float(far)

    def _clone(self):
        c = OrthographicCamera(self.near, self.far)
        c._copy_transform_from(self)
        return c


class BSDF(Object):
    twoSide = False

    def anisotropic(self) -> bool:
        return False


class DiffuseBSDF(BSDF):
    """reference src/psdr.cpp:279-284, src/bsdf/diffuse.cpp (1x1 reflectance bitmap)."""

    def __init__(self, reflectance=None):
        if isinstance(reflectance, str):                # DiffuseBSDF(file_name): src/psdr.cpp:281
            reflectance = Bitmap3fD(reflectance)
        if isinstance(reflectance, Bitmap3fD):          # textured reflectance (src/psdr.cpp:282)
            self.reflectance = reflectance
        else:
            self.reflectance = np.full(3, 0.5, dtype=np.float32) if reflectance is None else _f32(reflectance, (3,)).copy()
        self.d_reflectance = np.zeros(3, dtype=np.float32)

    def _clone(self):
        r = self.reflectance
        if isinstance(r, Bitmap3fD):
            r = r._clone()
        b = DiffuseBSDF(r)
        b.d_reflectance = _f32(self.d_reflectance).copy()
        b.twoSide = self.twoSide
        return b


class MicrofacetBSDF(BSDF):
    """reference src/psdr.cpp:298-304, src/bsdf/microfacet.cpp.  Argument order as in the reference:
    (specular, diffuse, roughness) (include/psdr/bsdf/microfacet.h:12); each slot is a constant (1x1 bitmap) or a
    texture: Bitmap3fD, Bitmap3fD, Bitmap1fD (microfacet.h:17)."""

    def __init__(self, specular=None, diffuse=None, roughness=None):
        if isinstance(specular, Bitmap3fD):
            self.specularReflectance = specular
        else:
            self.specularReflectance = np.full(3, 0.04, dtype=np.float32) if specular is None else _f32(specular, (3,)).copy()
        if isinstance(diffuse, Bitmap3fD):
            self.diffuseReflectance = diffuse
        else:
            self.diffuseReflectance = np.full(3, 0.5, dtype=np.float32) if diffuse is None else _f32(diffuse, (3,)).copy()
        if isinstance(roughness, Bitmap1fD):
            self.roughness = roughness
        else:
            self.roughness = np.float32(0.8) if roughness is None else (roughness if hasattr(roughness, "requires_grad") else np.float32(roughness))
        self.d_specularReflectance = np.zeros(3, dtype=np.float32)
        self.d_diffuseReflectance = np.zeros(3, dtype=np.float32)
        self.d_roughness = np.float32(0.0)

    def _clone(self):
        cl = lambda x: x._clone() if isinstance(x, _Bitmap) else x      # noqa: E731
        b = MicrofacetBSDF(cl(self.specularReflectance), cl(self.diffuseReflectance), cl(self.roughness))
        b.d_specularReflectance = _f32(self.d_specularReflectance, (3,)).copy()
        b.d_diffuseReflectance = _f32(self.d_diffuseReflectance).copy()
        b.d_roughness = np.float32(self.d_roughness)
        b.twoSide = self.twoSide
        return b


class RoughConductorBSDF(BSDF):
    """reference src/psdr.cpp:286-293, include/psdr/bsdf/roughconductor.h, src/bsdf/roughconductor.cpp: GGX conductor with
    complex index of refraction eta + i k.  ``RoughConductorBSDF(alpha, eta, k)`` with Bitmap1fD / Bitmap3fD arguments as
    in tutorials/batch_render.ipynb (plain numbers / triples are accepted too).  ``alpha_u`` (= ``alpha_v``: the isotropic
    form) and ``specular_reflectance`` may be textures; ``eta`` and ``k`` are per-channel constants."""

    def __init__(self, alpha=None, eta=None, k=None, specular_reflectance=None):
        self.alpha_u = alpha if isinstance(alpha, Bitmap1fD) else (np.float32(0.1) if alpha is None else
                                                                   (alpha if hasattr(alpha, "requires_grad") else np.float32(alpha)))
        self.eta = self._const3(eta, 0.0)
        self.k = self._const3(k, 1.0)
        sr = specular_reflectance
        self.specular_reflectance = sr if isinstance(sr, Bitmap3fD) and sr._textured() else self._const3(sr, 1.0)
        self.d_alpha_u = np.float32(0.0)
        self.d_eta, self.d_k, self.d_specular_reflectance = (np.zeros(3, dtype=np.float32) for _ in range(3))

    @staticmethod
    def _const3(v, default):
        if v is None:
            return np.full(3, default, dtype=np.float32)
        if isinstance(v, _Bitmap):
            if v._textured():
                raise RuntimeError("RoughConductorBSDF: eta and k are constants (1x1 bitmaps) on this path")
            v = v.data
        if hasattr(v, "requires_grad"):
            return v
        return _f32(v, (3,)).copy()

    @property
    def alpha_v(self):
        return self.alpha_u

    @alpha_v.setter
    def alpha_v(self, v):
        self.alpha_u = v

    def _clone(self):
        cl = lambda x: x._clone() if isinstance(x, _Bitmap) else x      # noqa: E731
        b = RoughConductorBSDF(cl(self.alpha_u), cl(self.eta), cl(self.k), cl(self.specular_reflectance))
        b.d_alpha_u = np.float32(self.d_alpha_u)
        b.d_eta, b.d_k = _f32(self.d_eta, (3,)).copy(), _f32(self.d_k, (3,)).copy()
        b.d_specular_reflectance = _f32(self.d_specular_reflectance, (3,)).copy()
        b.twoSide = self.twoSide
        return b


class RoughDielectricBSDF(BSDF):
    """reference include/psdr/bsdf/roughdielectric.h, src/bsdf/roughdielectric.cpp: GGX reflection + refraction through an
    interface with eta = intIOR / extIOR.  The reference binds the class without a constructor (src/psdr.cpp:295) and
    creates it from scene files only (src/scene/scene_loader.cpp:346-360); ``RoughDielectricBSDF(alpha, intIOR, extIOR)``
    mirrors the C++ constructors (roughdielectric.h:20-29) so that a scene can be built without XML as well.
    ``alpha_u`` (= ``alpha_v``) may be a Bitmap1fD; the indices of refraction are fixed at add_BSDF time."""

    def __init__(self, alpha=None, intIOR: float = 1.5, extIOR: float = 1.0):
        self.alpha_u = alpha if isinstance(alpha, Bitmap1fD) else (np.float32(0.1) if alpha is None else np.float32(alpha))
        self.intIOR, self.extIOR = float(intIOR), float(extIOR)
        self.d_alpha_u = np.float32(0.0)

    @property
    def alpha_v(self):
        return self.alpha_u

    @alpha_v.setter
    def alpha_v(self, v):
        self.alpha_u = v

    def _clone(self):
        b = RoughDielectricBSDF(self.alpha_u._clone() if isinstance(self.alpha_u, _Bitmap) else self.alpha_u, self.intIOR, self.extIOR)
        b.d_alpha_u = np.float32(self.d_alpha_u)
        b.twoSide = self.twoSide
        return b


class MicrofacetBSDFPerVertex(BSDF):
    """reference src/psdr.cpp:306-310, src/bsdf/microfacet_pv.cpp: Microfacet parameters given per VERTEX of the mesh the
    BSDF is attached to -- ``MicrofacetBSDFPerVertex(specular[n,3], diffuse[n,3], roughness[n])`` -- and interpolated
    with the hit's barycentrics.  Forward-mode tangents: d_specularReflectance, d_diffuseReflectance, d_roughness."""

    def __init__(self, specular, diffuse, roughness):
        self.specularReflectance = _f32(specular).reshape(-1, 3).copy()
        self.diffuseReflectance = _f32(diffuse).reshape(-1, 3).copy()
        self.roughness = _f32(roughness).reshape(-1).copy()
        n = len(self.roughness)
        if len(self.specularReflectance) != n or len(self.diffuseReflectance) != n:
            raise RuntimeError("MicrofacetBSDFPerVertex: the three arrays must cover the same vertices")
        self.d_specularReflectance = np.zeros((n, 3), np.float32)
        self.d_diffuseReflectance = np.zeros((n, 3), np.float32)
        self.d_roughness = np.zeros(n, np.float32)

    def _table(self, tangent: bool):
        p = "d_" if tangent else ""
        return np.ascontiguousarray(np.concatenate([_f32(getattr(self, p + "specularReflectance")).reshape(-1, 3),
                                                    _f32(getattr(self, p + "diffuseReflectance")).reshape(-1, 3),
                                                    _f32(getattr(self, p + "roughness")).reshape(-1, 1)], axis=1), np.float32)

    def _clone(self):
        b = MicrofacetBSDFPerVertex(self.specularReflectance, self.diffuseReflectance, self.roughness)
        b.d_specularReflectance = _f32(self.d_specularReflectance).reshape(-1, 3).copy()
        b.d_diffuseReflectance = _f32(self.d_diffuseReflectance).reshape(-1, 3).copy()
        b.d_roughness = _f32(self.d_roughness).reshape(-1).copy()
        b.twoSide = self.twoSide
        return b


class NormalMapBSDF(BSDF):
    """reference src/psdr.cpp:273-277, include/psdr/bsdf/normalmap.h, src/bsdf/normalmap.cpp: a nested BSDF evaluated on
    the facet the normal map selects (+ one tangent facet).  ``NormalMapBSDF()`` / ``NormalMapBSDF([x, y, z])`` /
    ``NormalMapBSDF(Bitmap3fD)``; fields ``normal_map`` and ``nested_bsdf`` as in the reference.  It enters a scene through
    ``Scene.add_normalmap_BSDF(normalmap, microfacet, name)`` (src/scene/scene.cpp:128-145) or ``Scene.add_BSDF``, which --
    as the reference does (scene.cpp:219-229) -- ignores the object's fields and installs the constant map
    (.499999, .499999, 1) around a default MicrofacetBSDF."""

    def __init__(self, normal_map=None):
        if isinstance(normal_map, Bitmap3fD):
            self.normal_map = normal_map
        else:
            self.normal_map = np.zeros(3, np.float32) if normal_map is None else _f32(normal_map, (3,)).copy()
        self.d_normal_map = np.zeros(3, dtype=np.float32)
        self.nested_bsdf = None

    def _clone(self):
        nm = self.normal_map
        b = NormalMapBSDF(nm._clone() if isinstance(nm, _Bitmap) else nm)
        b.d_normal_map = _f32(self.d_normal_map, (3,)).copy()
        b.nested_bsdf = None if self.nested_bsdf is None else self.nested_bsdf._clone()
        b.twoSide = self.twoSide
        return b


class Emitter(Object):
    pass


class _Bitmap:
    """reference src/psdr.cpp:195-219 (Bitmap1fD / Bitmap3fD): w x h bitmap, data = [h*w, channels] row-major
    (pixel = y*w + x), plus the uv transform applied by eval (src/core/bitmap.cpp:64-72): ``scale``, ``rotate``
    (radians, about the centre), ``translate`` (2 floats); ``d_*`` = forward-mode tangents."""
    channels = 3

    def __init__(self, width=1, height: int = 1, data=None):
        c = self.channels
        if isinstance(width, str):                     # Bitmap(file_name): src/core/bitmap.cpp:16-19
            w, h, rgb = _loader.load_exr(width)
            width, height, data = w, h, (rgb[:, :1] if c == 1 else rgb)
        elif data is None and not np.isscalar(width) and np.ndim(width) == 1 and len(width) == c:   # Bitmap3fD([r, g, b])
            width, data = 1, _f32(width, (1, c))
        if np.isscalar(width) and data is None and not isinstance(width, (int, np.integer)):   # Bitmap(value)
            width, data = 1, np.full((1, c), float(width), np.float32)
        self.resolution = (int(width), int(height))
        if data is None:
            data = np.zeros((self.resolution[0] * self.resolution[1], c), dtype=np.float32)
        self.data = data if hasattr(data, "requires_grad") else _f32(data).reshape(-1, c).copy()
        self.d_data = None
        n = int(np.prod(self.data.shape))
        if n != c * self.resolution[0] * self.resolution[1]:
            raise RuntimeError("Bitmap: invalid data size!")
        self.scale, self.rotate, self.translate = np.float32(1.0), np.float32(0.0), np.zeros(2, np.float32)
        self.d_scale, self.d_rotate, self.d_translate = np.float32(0.0), np.float32(0.0), np.zeros(2, np.float32)

    def _clone(self):
        c = type(self)(self.resolution[0], self.resolution[1], self.data)
        c.d_data = None if self.d_data is None else _f32(self.d_data).reshape(-1, self.channels).copy()
        for n in ("scale", "rotate", "d_scale", "d_rotate"):
            setattr(c, n, np.float32(getattr(self, n)))
        c.translate, c.d_translate = _f32(self.translate, (2,)).copy(), _f32(self.d_translate, (2,)).copy()
        return c

    def load_openexr(self, file_name: str):
        w, h, rgb = _loader.load_exr(file_name)
        self.resolution, self.data, self.d_data = (w, h), (rgb[:, :1].copy() if self.channels == 1 else rgb), None

    def _uv(self, tangent: bool):
        if tangent:
            return np.array([self.d_scale, self.d_rotate, self.d_translate[0], self.d_translate[1]], np.float32)
        return np.array([self.scale, self.rotate, self.translate[0], self.translate[1]], np.float32)

    def _textured(self) -> bool:
        return self.resolution[0] * self.resolution[1] > 1


class Bitmap3fD(_Bitmap):
    channels = 3


class Bitmap1fD(_Bitmap):
    channels = 1


class EnvironmentMap(Emitter):
    """reference src/psdr.cpp:349-355, src/emitter/envmap.cpp.  ``radiance`` is a lat-long Bitmap3fD; EXR loading
    (EnvironmentMap(path)) is I/O outside this path -- build the bitmap from an array instead."""

    def __init__(self, radiance: Optional[Bitmap3fD] = None):
        if isinstance(radiance, str):                   # EnvironmentMap(file_name): include/psdr/emitter/envmap.h:14-16
            radiance = Bitmap3fD(radiance)
        self.radiance = radiance if radiance is not None else Bitmap3fD(2, 2, np.ones((4, 3), np.float32))
        self.scale = np.float32(1.0)
        self.d_scale = np.float32(0.0)
        self.to_world = np.eye(4, dtype=np.float32)           # read-only in the reference (to_world_raw)
        self.to_world_left = np.eye(4, dtype=np.float32)
        self.d_to_world_left = np.zeros((4, 4), dtype=np.float32)

    def set_transform(self, mat, tangent=None):
        self.to_world_left = mat if hasattr(mat, "requires_grad") else _mat4(mat)
        self.d_to_world_left = np.zeros((4, 4), np.float32) if tangent is None else _f32(tangent, (4, 4)).copy()

    def _clone(self):
        e = EnvironmentMap(self.radiance._clone())
        e.scale, e.d_scale = self.scale, self.d_scale
        e.to_world = _mat4(self.to_world)
        e.to_world_left = self.to_world_left if hasattr(self.to_world_left, "requires_grad") else _mat4(self.to_world_left)
        e.d_to_world_left = _f32(self.d_to_world_left, (4, 4)).copy()
        return e


class AreaLight(Emitter):
    """reference src/psdr.cpp:344-347, src/emitter/area.cpp"""

    def __init__(self, radiance, mesh=None):
        self.radiance = _f32(radiance, (3,)).copy()
        self.d_radiance = np.zeros(3, dtype=np.float32)


_load_obj = scenes.load_obj


class Mesh(_Transformable):
    """reference src/psdr.cpp:314-339, src/shape/mesh.cpp"""

    def __init__(self):
        super().__init__()
        self.vertex_positions = np.zeros((0, 3), dtype=np.float32)
        self.d_vertex_positions = None
        self.face_indices = np.zeros((0, 3), dtype=np.int32)
        self.vertex_uv = None
        self.face_uv_indices = None
        self.use_face_normal = False
        self.enable_edges = True
        self.bsdf = None
        self._scene = None
        self._index = -1

    num_vertices = property(lambda self: len(self.vertex_positions))
    num_faces = property(lambda self: len(self.face_indices))

    def load_raw(self, v, f, uv=None, f_uv=None, verbose: bool = False):
        self.vertex_positions = _f32(v).reshape(-1, 3).copy()
        self.face_indices = np.ascontiguousarray(np.asarray(f, dtype=np.int32).reshape(-1, 3))
        if uv is not None and len(uv) > 0:
            self.vertex_uv = _f32(uv).reshape(-1, 2).copy()
            self.face_uv_indices = np.ascontiguousarray(np.asarray(f_uv, dtype=np.int32).reshape(-1, 3))
        else:
            self.vertex_uv, self.face_uv_indices = None, None
        self.d_vertex_positions = None

    def load(self, filename: str, verbose: bool = False):
        v, f, uv, fuv = _load_obj(filename)
        if len(f) == 0:
            raise RuntimeError("Failed to load OBJ from: " + filename)
        self.load_raw(v, f, uv, fuv, verbose)

    def dump(self, fname: str, raw: bool = False):
        """reference Mesh::dump (src/shape/mesh.cpp:469-554): OBJ with per-vertex normals unless face normals are used.
        raw = False writes the object-space vertices, raw = True the world-space ones (the reference's flag naming)."""
        v = self.vertex_positions
        if raw:
            tw = _f32(self.to_world_left, (4, 4)) @ _f32(self.to_world, (4, 4)) @ _f32(self.to_world_right, (4, 4))
            v = (np.c_[_f32(v).reshape(-1, 3), np.ones(len(v), np.float32)] @ tw.T)[:, :3].astype(np.float32)
        m = self._clone()
        m.vertex_positions = _f32(v).reshape(-1, 3)
        _loader.dump_obj(m, fname, None if self.use_face_normal else _loader.vertex_normals(m.vertex_positions, m.face_indices))

    def edge_indices(self):
        if self._scene is None or self._scene._h is None:
            raise RuntimeError("edge_indices() needs a mesh that was added to a configured scene")
        L = _lib.load()
        n = L.psdr_scene_query(self._scene._h, _lib.Q_NUM_MESH_EDGES, self._index)
        out = np.zeros((4, n), dtype=np.int32)
        _lib.check(L.psdr_scene_mesh_edges(self._scene._h, self._index, _ip(out)))
        return out

    def _clone(self):
        m = Mesh()
        m._copy_transform_from(self)
        m.vertex_positions = self.vertex_positions.copy()
        m.d_vertex_positions = None if self.d_vertex_positions is None else _f32(self.d_vertex_positions).reshape(-1, 3).copy()
        m.face_indices = self.face_indices.copy()
        m.vertex_uv = None if self.vertex_uv is None else self.vertex_uv.copy()
        m.face_uv_indices = None if self.face_uv_indices is None else self.face_uv_indices.copy()
        m.use_face_normal, m.enable_edges = self.use_face_normal, self.enable_edges
        m.id = self.id
        return m


class Sampler:
    """reference src/psdr.cpp:181-185 -- host-side view of the PCG32 streams the kernels use."""

    def __init__(self):
        self._seed, self._n, self._k = None, 0, 0

    def seed(self, seed_value):
        sv = np.asarray(seed_value, dtype=np.int64).ravel()
        if len(sv) == 0 or not np.array_equal(sv - sv[0], np.arange(len(sv))):
            raise NotImplementedError("only seeds of the form arange(n) + s are supported")
        self._seed, self._n, self._k = int(sv[0]), len(sv), 0

    def _draw(self, n):
        if self._seed is None:
            raise RuntimeError("Sampler::seed() must be invoked before using this sampler!")
        out = np.zeros((self._k + n, self._n), dtype=np.float32)
        _lib.check(_lib.load().psdr_sampler_draws(self._seed, self._n, self._k + n, _fp(out)))
        self._k += n
        return out[-n:]

    def next_1d(self):
        return self._draw(1)[0]

    def next_2d(self):
        d = self._draw(2)
        return np.stack([d[1], d[0]])   # y receives the first draw (sampler.h:19-21, GCC argument order)


class Scene(Object):
    """reference src/psdr.cpp:393-414, src/scene/scene.cpp"""

    def __init__(self, device: Optional[int] = None):
        self.opts = RenderOption()
        self.seed = 0
        self.param_map: Dict[str, Object] = {}
        self._sensors: List[PerspectiveCamera] = []
        self._bsdfs: List[BSDF] = []
        self._meshes: List[Mesh] = []
        self._emitters: List[AreaLight] = []
        self._mesh_emitter: List[int] = []
        self._h = None
        self._pushed = [0, 0, 0]   # bsdfs, creation events (meshes / envmap), sensors already in the native scene
        self._events = []          # ("mesh", i) / ("env", emitter index) in add_* order = native emitter order
        self._env = None
        self._device = device
        self._shard = (0, 1)
        self._peer_group = None    # process group of the fused NVLink reduction (enable_peer_reduction)
        self._peer_bufs = {}
        self._accel = -1
        self.reference_arithmetic = False     # Dr.Jit's approximate rcp in renderD's analytic primary hit (psdr_b200.h)

    def __del__(self):
        try:
            if self._h is not None:
                _lib.load().psdr_scene_destroy(self._h)
        except Exception:
            pass

    num_sensors = property(lambda self: len(self._sensors))
    num_meshes = property(lambda self: len(self._meshes))

    def get_num_emitters(self) -> int:
        return len(self._emitters)

    def _register(self, kind: str, arr: Sequence[Object]):
        for i, o in enumerate(arr):
            self.param_map["%s[%d]" % (kind, i)] = o
            if o.id:
                self.param_map["%s[id=%s]" % (kind, o.id)] = o

    def add_Sensor(self, sensor: PerspectiveCamera):
        if not isinstance(sensor, PerspectiveCamera):
            raise RuntimeError("Unknown sensor type!")
        self._sensors.append(sensor._clone())
        self._register("Sensor", self._sensors)

    def add_BSDF(self, bsdf: BSDF, name: str, twoSide: bool = False):
        if not isinstance(bsdf, (DiffuseBSDF, MicrofacetBSDF, RoughConductorBSDF, RoughDielectricBSDF, MicrofacetBSDFPerVertex, NormalMapBSDF)):
            raise RuntimeError("Unknown BSDF type!")
        if ("BSDF[id=%s]" % name) in self.param_map:
            raise RuntimeError("Duplicate BSDF id: " + name)
        if isinstance(bsdf, NormalMapBSDF):
            # scene.cpp:219-229: a fresh NormalMap((.499999, .499999, 1)) around Microfacet(), whatever was passed in
            b = NormalMapBSDF([.499999, .499999, 1.0])
            b.nested_bsdf = MicrofacetBSDF()
        else:
            b = bsdf._clone()
        b.id = name
        b.twoSide = bool(twoSide)
        self._bsdfs.append(b)
        self._register("BSDF", self._bsdfs)

    def add_normalmap_BSDF(self, bsdf1: "NormalMapBSDF", bsdf2: BSDF, name: str, twoSide: bool = False):
        """reference src/psdr.cpp:402, src/scene/scene.cpp:128-145: a NormalMap with bsdf1's normal map around a copy of the
        MicrofacetBSDF bsdf2.  (The scene-file loader also nests Diffuse, RoughConductor and RoughDielectric BSDFs,
        scene_loader.cpp:372-424; they are accepted here too.)"""
        if not isinstance(bsdf1, NormalMapBSDF) or not isinstance(bsdf2, (MicrofacetBSDF, DiffuseBSDF, RoughConductorBSDF, RoughDielectricBSDF)):
            raise RuntimeError("Unknown BSDF type!")
        if ("BSDF[id=%s]" % name) in self.param_map:
            raise RuntimeError("Duplicate BSDF id: " + name)
        b = bsdf1._clone()
        b.nested_bsdf = bsdf2._clone()
        b.nested_bsdf.twoSide = False
        b.id = name
        b.twoSide = bool(twoSide)
        self._bsdfs.append(b)
        self._register("BSDF", self._bsdfs)

    def add_EnvironmentMap(self, envmap, to_world=None, scale: float = 1.0):
        """add_EnvironmentMap(EnvironmentMap) (reference src/scene/scene.cpp:97-105); the (path, to_world, scale)
        overload needs an EXR reader and is not part of this path."""
        if self._env is not None:
            raise RuntimeError("A scene is only allowed to have one envmap!")
        if isinstance(envmap, str):                     # add_EnvironmentMap(fname, to_world, scale): scene.cpp:85-95
            envmap = EnvironmentMap(envmap)
            envmap.to_world = _mat4(to_world)
            envmap.scale = np.float32(scale)
        if not isinstance(envmap, EnvironmentMap):
            raise RuntimeError("Unknown emitter type!")
        e = envmap._clone()
        self._env = e
        self._events.append(("env", len(self._emitters)))
        self._emitters.append(e)
        self._register("Emitter", self._emitters)

    def add_Mesh(self, mesh_or_path, *args):
        """add_Mesh(path, to_world, bsdf_id, emitter) or add_Mesh(mesh, bsdf_id, emitter=None)"""
        if isinstance(mesh_or_path, Mesh):
            bsdf_id = args[0]
            emitter = args[1] if len(args) > 1 else None
            mesh = mesh_or_path._clone()
        else:
            to_world, bsdf_id = args[0], args[1]
            emitter = args[2] if len(args) > 2 else None
            mesh = Mesh()
            mesh.load(mesh_or_path)
            mesh.to_world = _mat4(to_world)
        if ("BSDF[id=%s]" % bsdf_id) not in self.param_map:
            raise RuntimeError("Unknown BSDF id: " + str(bsdf_id))
        mesh.bsdf = bsdf_id
        mesh._scene, mesh._index = self, len(self._meshes)
        self._events.append(("mesh", len(self._meshes)))
        if emitter is not None:
            if not isinstance(emitter, AreaLight):
                raise RuntimeError("Unknown emitter type!")
            e = AreaLight(emitter.radiance)
            e.d_radiance = emitter.d_radiance.copy()
            self._mesh_emitter.append(len(self._emitters))
            self._emitters.append(e)
            self._register("Emitter", self._emitters)
        else:
            self._mesh_emitter.append(-1)
        self._meshes.append(mesh)
        self._register("Mesh", self._meshes)

    def load_file(self, file_name: str, auto_configure: bool = True):
        """reference Scene::load_file (src/psdr.cpp:407, src/scene/scene_loader.cpp): Mitsuba-style XML scene"""
        import xml.etree.ElementTree as ET
        try:
            root = ET.parse(file_name).getroot()
        except Exception:
            raise RuntimeError("XML parsing failed")
        _loader.load_scene_xml(self, root, os.path.dirname(os.path.abspath(file_name)))
        if auto_configure:
            self.configure()

    def load_string(self, scene_xml: str, auto_configure: bool = True):
        import xml.etree.ElementTree as ET
        try:
            root = ET.fromstring(scene_xml)
        except Exception:
            raise RuntimeError("XML parsing failed")
        _loader.load_scene_xml(self, root, os.getcwd())
        if auto_configure:
            self.configure()

    # -- multi-GPU / acceleration knobs (new; the reference is single-GPU OptiX)
    def set_shard(self, rank: int, world: int):
        self._shard = (int(rank), int(world))
        if self._h is not None:
            _lib.check(_lib.load().psdr_scene_set_shard(self._h, rank, world))

    def enable_peer_reduction(self, group=None, strict: bool = False) -> bool:
        """Fuse the multi-GPU image / gradient reduction INTO the term kernels: on a sharded scene (set_shard) the kernels
        of every rank accumulate through an NVLS multicast address (multimem.red over NVLink / NVSwitch; dist.PeerBuffers),
        so renderC / renderD_fwd / renderD_primal / render_vjp return the complete result on every rank with no
        all-reduce.  Returns False (and leaves the NCCL path in place) when the node has no multicast support."""
        import torch.distributed as dist
        from . import dist as _dist
        if self._shard[1] <= 1 or not dist.is_initialized():
            if strict:
                raise RuntimeError("enable_peer_reduction needs a sharded scene and an initialised process group")
            return False
        import torch
        dev = torch.device("cuda", self._device_index())
        probe, err = None, None
        try:
            probe = _dist.PeerBuffers(32, dev, group)
        except Exception as e:      # no symmetric memory / no multicast on this node
            err = e
        # the ranks must take the same path: one of them falling back to NCCL while the others wait in the device
        # barrier would hang the job
        ok = torch.tensor([1 if probe is not None else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            if strict:
                raise RuntimeError("no NVLS multicast path on every rank: %s" % (err,))
            return False
        self._peer_group = group if group is not None else dist.group.WORLD
        self._peer_bufs = {("probe", 32): probe}
        return True

    def _peer(self, key: str, numel: int):
        """PeerBuffers for outputs of `numel` floats, or None when the fused reduction is off."""
        if self._peer_group is None or self._shard[1] <= 1:
            return None
        pb = self._peer_bufs.get((key, numel))
        if pb is None:
            import torch
            from . import dist as _dist
            pb = self._peer_bufs[(key, numel)] = _dist.PeerBuffers(numel, torch.device("cuda", self._device_index()), self._peer_group)
        return pb

    def set_accel(self, mode: int):
        self._accel = int(mode)
        if self._h is not None:
            _lib.check(_lib.load().psdr_scene_set_accel(self._h, self._accel))

    def _device_index(self) -> int:
        if self._device is None:
            try:
                import torch
                self._device = torch.cuda.current_device() if torch.cuda.is_available() else 0
            except Exception:
                self._device = 0
        return int(self._device)

    def _native(self):
        L = _lib.load()
        if self._h is None:
            dev = self._device_index()
            h = L.psdr_scene_create(int(dev))
            if not h:
                raise RuntimeError(L.psdr_last_error().decode())
            self._h = C.c_void_p(h)
            _lib.check(L.psdr_scene_set_shard(self._h, *self._shard))
            _lib.check(L.psdr_scene_set_accel(self._h, self._accel))
        return L

    def configure(self, active_sensor: Sequence[int] = ()):
        """Scene.configure (reference src/scene/scene.cpp:311-601): pushes every parameter (and
        tangent) to the native scene, which rebuilds triangle records, distributions, edges and BVH."""
        if not self._meshes:
            raise RuntimeError("Missing meshes!")
        if not self._sensors:
            raise RuntimeError("Missing sensor!")
        L = self._native()
        o = self.opts
        _lib.check(L.psdr_scene_set_options(self._h, o.width, o.height, o.spp, o.sppe, o.sppse, o.log_level))
        _lib.check(L.psdr_scene_set_seed(self._h, int(self.seed)))
        _lib.check(L.psdr_scene_set_reference_arithmetic(self._h, int(bool(self.reference_arithmetic))))
        def add_native_bsdf(b, name):
            if isinstance(b, MicrofacetBSDF):
                d0 = np.full(3, 0.5, np.float32) if isinstance(b.diffuseReflectance, _Bitmap) else _f32(b.diffuseReflectance)
                s0 = np.full(3, 0.04, np.float32) if isinstance(b.specularReflectance, _Bitmap) else _f32(b.specularReflectance)
                r0 = 0.8 if isinstance(b.roughness, _Bitmap) else float(_f32(b.roughness).ravel()[0])
                rc = L.psdr_scene_add_bsdf_microfacet(self._h, name, _fp(s0), _fp(d0), r0, int(b.twoSide))
            elif isinstance(b, RoughConductorBSDF):
                a0 = 0.1 if isinstance(b.alpha_u, _Bitmap) else float(_f32(b.alpha_u).ravel()[0])
                s0 = np.ones(3, np.float32) if isinstance(b.specular_reflectance, _Bitmap) else _f32(b.specular_reflectance, (3,))
                rc = L.psdr_scene_add_bsdf_roughconductor(self._h, name, a0, _fp(_f32(b.eta, (3,))), _fp(_f32(b.k, (3,))), _fp(s0), int(b.twoSide))
            elif isinstance(b, RoughDielectricBSDF):
                a0 = 0.1 if isinstance(b.alpha_u, _Bitmap) else float(_f32(b.alpha_u).ravel()[0])
                rc = L.psdr_scene_add_bsdf_roughdielectric(self._h, name, a0, b.intIOR, b.extIOR, int(b.twoSide))
            elif isinstance(b, MicrofacetBSDFPerVertex):
                rc = L.psdr_scene_add_bsdf_microfacet_pervertex(self._h, name, _fp(_f32(b.specularReflectance).reshape(-1)), _fp(_f32(b.diffuseReflectance).reshape(-1)),
                                                                _fp(_f32(b.roughness).reshape(-1)), len(_f32(b.roughness).reshape(-1)), int(b.twoSide))
            elif isinstance(b, NormalMapBSDF):
                if b.nested_bsdf is None:
                    raise RuntimeError("NormalMapBSDF: nested_bsdf is not set")
                _lib.check(L.psdr_scene_begin_nested_bsdf(self._h))
                b._nested_handle = add_native_bsdf(b.nested_bsdf, b"")
                n0 = np.array([.5, .5, 1.], np.float32) if isinstance(b.normal_map, _Bitmap) else _f32(b.normal_map, (3,))
                rc = L.psdr_scene_add_bsdf_normalmap(self._h, name, _fp(n0), b._nested_handle, int(b.twoSide))
            else:
                r0 = np.full(3, 0.5, np.float32) if isinstance(b.reflectance, Bitmap3fD) else _f32(b.reflectance)
                rc = L.psdr_scene_add_bsdf_diffuse(self._h, name, _fp(r0), int(b.twoSide))
            if rc < 0:
                raise RuntimeError(L.psdr_last_error().decode())
            return rc

        for b in self._bsdfs[self._pushed[0]:]:
            add_native_bsdf(b, b.id.encode())
        self._pushed[0] = len(self._bsdfs)
        for kind, i in self._events[self._pushed[1]:]:
            if kind == "env":
                e = self._env
                w, h = e.radiance.resolution
                if L.psdr_scene_add_envmap(self._h, _fp(_f32(e.radiance.data).reshape(-1)), w, h, _fp(_mat4(e.to_world)), float(_f32(e.scale).ravel()[0])) < 0:
                    raise RuntimeError(L.psdr_last_error().decode())
                continue
            m = self._meshes[i]
            v, f = _f32(m.vertex_positions).reshape(-1, 3), np.ascontiguousarray(m.face_indices, dtype=np.int32)
            uv = None if m.vertex_uv is None else _f32(m.vertex_uv).reshape(-1, 2)
            fuv = None if m.face_uv_indices is None else np.ascontiguousarray(m.face_uv_indices, dtype=np.int32)
            e = self._mesh_emitter[i]
            rad = None if e < 0 else _f32(self._emitters[e].radiance)
            if L.psdr_scene_add_mesh(self._h, _fp(v), len(v), _ip(f), len(f), _fp(uv), 0 if uv is None else len(uv), _ip(fuv),
                                     _fp(_mat4(m.to_world)), m.bsdf.encode(), _fp(rad), int(m.use_face_normal), int(m.enable_edges)) < 0:
                raise RuntimeError(L.psdr_last_error().decode())
        self._pushed[1] = len(self._events)
        for s in self._sensors[self._pushed[2]:]:
            if isinstance(s, OrthographicCamera):
                rc = L.psdr_scene_add_orthographic(self._h, s.near, s.far, _fp(_mat4(s.to_world)))
            else:
                rc = (L.psdr_scene_add_perspective(self._h, s.fov, s.near, s.far, _fp(_mat4(s.to_world))) if s.intrinsics is None else
                      L.psdr_scene_add_perspective_intrinsic(self._h, *s.intrinsics, s.near, s.far, _fp(_mat4(s.to_world))))
            if rc < 0:
                raise RuntimeError(L.psdr_last_error().decode())
        self._pushed[2] = len(self._sensors)

        sent = self.__dict__.setdefault("_sent", {})      # what the native scene already holds: (kind, index, is_tangent) -> key

        def key_of(a):
            """cheap identity of a parameter value: bytes for small arrays, (object, in-place version) for torch tensors,
            a checksum for mid-sized numpy arrays; None = always push"""
            if hasattr(a, "_version") and hasattr(a, "data_ptr"):
                return ("t", id(a), a._version, tuple(a.shape))
            if a.size <= 64:
                return a.tobytes()
            if a.nbytes <= (1 << 20):
                import zlib
                return ("c", a.shape, zlib.crc32(a))
            return None

        def push1(kind, idx, value, is_tangent, fn):
            raw = value
            value = _f32(value).ravel()
            k = key_of(raw if hasattr(raw, "_version") else value)
            if k is not None and sent.get((kind, idx, is_tangent)) == k:
                return                                   # unchanged since the last configure: nothing to send
            _lib.check(fn(self._h, kind, idx, _fp(value), value.size))
            sent[(kind, idx, is_tangent)] = k

        def push(kind, idx, value, tangent):
            push1(kind, idx, value, False, L.psdr_scene_set_param)
            n = int(np.prod(np.shape(_f32(value)))) if tangent is None else 0
            push1(kind, idx, np.zeros(n, np.float32) if tangent is None else tangent, True, L.psdr_scene_set_tangent)

        for i, m in enumerate(self._meshes):
            push(_lib.MESH_VERTICES, i, m.vertex_positions, m.d_vertex_positions)
            push(_lib.MESH_TO_WORLD_LEFT, i, m.to_world_left, m.d_to_world_left)
            push(_lib.MESH_TO_WORLD_RAW, i, m.to_world, m.d_to_world)
            push(_lib.MESH_TO_WORLD_RIGHT, i, m.to_world_right, m.d_to_world_right)
        for i, s in enumerate(self._sensors):
            push(_lib.SENSOR_TO_WORLD_LEFT, i, s.to_world_left, s.d_to_world_left)
            push(_lib.SENSOR_TO_WORLD_RAW, i, s.to_world, s.d_to_world)
            push(_lib.SENSOR_TO_WORLD_RIGHT, i, s.to_world_right, s.d_to_world_right)
        def push_slot(i, slot, kind, r, d_const):
            """one bitmap slot of a BSDF: texture (texels + uv transform) or constant"""
            res = tuple(r.resolution) if isinstance(r, _Bitmap) and r._textured() else (1, 1)
            if sent.get(("slot", i, slot)) != res:       # the native side re-allocates the slot: forget what it held
                sent[("slot", i, slot)] = res
                sent.pop((kind, i, False), None)
                sent.pop((kind, i, True), None)
            if isinstance(r, _Bitmap) and r._textured():
                _lib.check(L.psdr_scene_set_bsdf_texture_slot(self._h, i, slot, r.resolution[0], r.resolution[1]))
                push(kind, i, r.data, r.d_data)
                push(_lib.BSDF_REFLECTANCE_UV + slot, i, r._uv(False), r._uv(True))
            else:
                _lib.check(L.psdr_scene_set_bsdf_texture_slot(self._h, i, slot, 1, 1))
                push(kind, i, r.data if isinstance(r, _Bitmap) else np.reshape(_f32(r), (-1,)), np.reshape(_f32(d_const), (-1,)))

        def push_bsdf(i, b):
            if isinstance(b, MicrofacetBSDF):
                push_slot(i, _lib.TEX_REFLECTANCE, _lib.BSDF_REFLECTANCE, b.diffuseReflectance, b.d_diffuseReflectance)
                push_slot(i, _lib.TEX_SPECULAR, _lib.BSDF_SPECULAR, b.specularReflectance, b.d_specularReflectance)
                push_slot(i, _lib.TEX_ROUGHNESS, _lib.BSDF_ROUGHNESS, b.roughness, b.d_roughness)
            elif isinstance(b, RoughConductorBSDF):
                push_slot(i, _lib.TEX_SPECULAR, _lib.BSDF_SPECULAR, b.specular_reflectance, b.d_specular_reflectance)
                push_slot(i, _lib.TEX_ROUGHNESS, _lib.BSDF_ROUGHNESS, b.alpha_u, b.d_alpha_u)
                push(_lib.BSDF_ETA, i, np.reshape(_f32(b.eta), (-1,)), np.reshape(_f32(b.d_eta), (-1,)))
                push(_lib.BSDF_K, i, np.reshape(_f32(b.k), (-1,)), np.reshape(_f32(b.d_k), (-1,)))
            elif isinstance(b, RoughDielectricBSDF):
                push_slot(i, _lib.TEX_ROUGHNESS, _lib.BSDF_ROUGHNESS, b.alpha_u, b.d_alpha_u)
            elif isinstance(b, MicrofacetBSDFPerVertex):
                push(_lib.BSDF_PERVERTEX, i, b._table(False), b._table(True))
            elif isinstance(b, NormalMapBSDF):
                push_slot(i, _lib.TEX_REFLECTANCE, _lib.BSDF_REFLECTANCE, b.normal_map, b.d_normal_map)
                push_bsdf(b._nested_handle, b.nested_bsdf)
            else:
                push_slot(i, _lib.TEX_REFLECTANCE, _lib.BSDF_REFLECTANCE, b.reflectance, b.d_reflectance)

        for i, b in enumerate(self._bsdfs):
            push_bsdf(i, b)
        for i, e in enumerate(self._emitters):
            if isinstance(e, EnvironmentMap):
                push(_lib.ENVMAP_RADIANCE, i, e.radiance.data, e.radiance.d_data)
                push(_lib.ENVMAP_SCALE, i, np.reshape(_f32(e.scale), (1,)), np.reshape(_f32(e.d_scale), (1,)))
                push(_lib.ENVMAP_TO_WORLD_LEFT, i, e.to_world_left, e.d_to_world_left)
            else:
                push(_lib.EMITTER_RADIANCE, i, e.radiance, e.d_radiance)
        act = np.asarray(list(active_sensor), dtype=np.int32)
        self._last_active = [int(a) for a in act]
        _lib.check(L.psdr_scene_configure(self._h, _ip(act), len(act)))
        if o.log_level > 0 and o.sppe > 0:
            print("(%s) primary edges initialized." % ", ".join(str(self.num_primary_edges(i)) for i in range(len(self._sensors))))
        if o.log_level > 0 and o.sppse > 0:
            print("%d secondary edges initialized." % self.num_secondary_edges())

    # -- reverse mode -------------------------------------------------------------------------
    _GRAD_FIELDS = {"Mesh": (("vertex_positions", _lib.MESH_VERTICES), ("to_world_left", _lib.MESH_TO_WORLD_LEFT),
                             ("to_world", _lib.MESH_TO_WORLD_RAW), ("to_world_right", _lib.MESH_TO_WORLD_RIGHT)),
                    "Sensor": (("to_world_left", _lib.SENSOR_TO_WORLD_LEFT), ("to_world", _lib.SENSOR_TO_WORLD_RAW),
                               ("to_world_right", _lib.SENSOR_TO_WORLD_RIGHT)),
                    "BSDF": (("reflectance", _lib.BSDF_REFLECTANCE), ("diffuseReflectance", _lib.BSDF_REFLECTANCE),
                             ("reflectance.data", _lib.BSDF_REFLECTANCE), ("diffuseReflectance.data", _lib.BSDF_REFLECTANCE),
                             ("specularReflectance", _lib.BSDF_SPECULAR), ("roughness", _lib.BSDF_ROUGHNESS),
                             ("specularReflectance.data", _lib.BSDF_SPECULAR), ("roughness.data", _lib.BSDF_ROUGHNESS),
                             ("alpha_u", _lib.BSDF_ROUGHNESS), ("alpha_u.data", _lib.BSDF_ROUGHNESS), ("eta", _lib.BSDF_ETA), ("k", _lib.BSDF_K),
                             ("specular_reflectance", _lib.BSDF_SPECULAR), ("specular_reflectance.data", _lib.BSDF_SPECULAR),
                             ("normal_map", _lib.BSDF_REFLECTANCE), ("normal_map.data", _lib.BSDF_REFLECTANCE)),
                    "Emitter": (("radiance", _lib.EMITTER_RADIANCE),),
                    "EnvironmentMap": (("radiance.data", _lib.ENVMAP_RADIANCE), ("scale", _lib.ENVMAP_SCALE),
                                       ("to_world_left", _lib.ENVMAP_TO_WORLD_LEFT))}

    @staticmethod
    def _field(o, dotted):
        for part in dotted.split("."):
            o = getattr(o, part, None)
            if o is None:
                return None
        if isinstance(o, _Bitmap):        # a bitmap-valued field named without ".data": its texels are the parameter
            o = o.data
        return o

    def _fields_of(self, kind_name, o):
        return self._GRAD_FIELDS["EnvironmentMap" if isinstance(o, EnvironmentMap) else kind_name]

    def _objects(self):
        return (("Mesh", self._meshes), ("Sensor", self._sensors), ("BSDF", self._bsdfs), ("Emitter", self._emitters))

    def _grad_leaves(self):
        """[(tensor, kind, index)] for every parameter field that is a torch tensor requiring grad."""
        out, seen = [], set()
        for kind_name, objs in self._objects():
            for i, o in enumerate(objs):
                for field, kind in self._fields_of(kind_name, o):
                    t = self._field(o, field)
                    if hasattr(t, "requires_grad") and t.requires_grad and id(t) not in seen:    # ("x" and "x.data" name the same texels)
                        seen.add(id(t))
                        out.append((t, kind, i))
                if isinstance(o, NormalMapBSDF) and o.nested_bsdf is not None and hasattr(o, "_nested_handle"):
                    for field, kind in self._GRAD_FIELDS["BSDF"]:      # the wrapped BSDF: addressed by its nested handle
                        t = self._field(o.nested_bsdf, field)
                        if hasattr(t, "requires_grad") and t.requires_grad and id(t) not in seen:
                            seen.add(id(t))
                            out.append((t, kind, o._nested_handle))
        return out

    def grad_of(self, name: str, field: str) -> np.ndarray:
        """Gradient of the last adjoint pass for ``scene.param_map[name].<field>`` (drjit.grad in the reference)."""
        obj = self.param_map[name]
        if "." in field and field.rsplit(".", 1)[1] in ("scale", "rotate", "translate") and isinstance(obj, BSDF):
            # the uv transform of a bitmap-valued slot (reference include/psdr/core/bitmap.h:36-38)
            attr, what = field.rsplit(".", 1)
            slots = {"reflectance": 0, "diffuseReflectance": 0, "normal_map": 0, "specularReflectance": 1, "specular_reflectance": 1,
                     "roughness": 2, "alpha_u": 2}
            bm = getattr(obj, attr, None)
            if attr in slots and isinstance(bm, _Bitmap) and bm._textured():
                g = self._read_grad(_lib.BSDF_REFLECTANCE_UV + slots[attr], self._bsdfs.index(obj), (4,))
                return {"scale": g[0:1], "rotate": g[1:2], "translate": g[2:4]}[what].copy()
            raise RuntimeError("no bitmap-valued field %s on %s" % (attr, name))
        for kind_name, objs in self._objects():
            for i, o in enumerate(objs):
                if o is obj and isinstance(o, MicrofacetBSDFPerVertex):
                    cols = {"specularReflectance": slice(0, 3), "diffuseReflectance": slice(3, 6), "roughness": 6}
                    if field not in cols:
                        break
                    return self._read_grad(_lib.BSDF_PERVERTEX, i, (len(o.roughness), 7))[:, cols[field]].copy()
                if o is obj and isinstance(o, NormalMapBSDF) and field.startswith("nested_bsdf."):
                    sub = field[len("nested_bsdf."):]
                    for f, kind in self._GRAD_FIELDS["BSDF"]:
                        if f == sub and self._field(o.nested_bsdf, f) is not None:
                            return self._read_grad(kind, o._nested_handle, np.asarray(_f32(self._field(o.nested_bsdf, f))).shape)
                    break
                if o is obj:
                    for f, kind in self._fields_of(kind_name, o):
                        if f == field:
                            return self._read_grad(kind, i, np.asarray(_f32(self._field(o, f))).shape)
        raise RuntimeError("no differentiable field %s on %s" % (field, name))

    def _read_grad(self, kind: int, index: int, shape) -> np.ndarray:
        out = np.zeros(int(np.prod(shape)), dtype=np.float32)
        _lib.check(_lib.load().psdr_scene_get_grad(self._h, kind, index, _fp(out), out.size))
        return out.reshape(shape)

    def _sampler_state(self):
        st = (C.c_longlong * 6)()
        _lib.check(_lib.load().psdr_scene_get_sampler_state(self._h, st))
        return st

    def _set_sampler_state(self, st):
        _lib.check(_lib.load().psdr_scene_set_sampler_state(self._h, st))

    def is_ready(self) -> bool:
        return self._h is not None and _lib.load().psdr_scene_query(self._h, _lib.Q_IS_CONFIGURED, 0) == 1

    def num_primary_edges(self, sensor: int = 0) -> int:
        return _lib.load().psdr_scene_query(self._h, _lib.Q_NUM_PRIMARY_EDGES, sensor)

    def num_secondary_edges(self) -> int:
        return _lib.load().psdr_scene_query(self._h, _lib.Q_NUM_SECONDARY_EDGES, 0)

    def kernel_family(self) -> int:
        """which kernel instantiation the configured scene runs (PSDR_Q_KERNEL_FAMILY): bit 0 BVH2, bit 1 Microfacet /
        envmap code, bit 3 the extended material set"""
        return _lib.load().psdr_scene_query(self._h, _lib.Q_KERNEL_FAMILY, 0)

    def last_configure_ms(self) -> float:
        return float(_lib.load().psdr_scene_last_configure_ms(self._h))


class Integrator(Object):
    """reference src/psdr.cpp:419-421, src/integrator/integrator.cpp"""
    max_depth = 1
    hide_emitters = False
    reference_tangent_scaling = False
    grad_image = None

    _kind, _mis = 0, 2      # PSDR_INTEGRATOR_PATH; Direct overrides

    def _check(self, scene: Scene):
        if scene._h is None or not scene.is_ready():
            raise RuntimeError("Input scene must be configured!")
        _lib.check(_lib.load().psdr_scene_set_integrator(scene._h, self._kind, self._mis))

    @staticmethod
    def _torch():
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("psdr_jit_b200 needs a CUDA device (no CPU fallback)")
        return torch

    @staticmethod
    def _dev_stream(torch, scene):
        """Outputs live on the SCENE's device and launches go to torch's current stream of that device (the native
        side does cudaSetDevice(scene device); a stream of another device would be an invalid handle)."""
        dev = torch.device("cuda", scene._device_index())
        return dev, torch.cuda.current_stream(dev).cuda_stream

    def _pix(self, torch, batch_pix, device):
        if batch_pix is None or (np.isscalar(batch_pix) and int(batch_pix) == -1):
            return None
        if hasattr(batch_pix, "to"):
            return batch_pix.to(device=device, dtype=torch.int32).contiguous()
        return torch.as_tensor(np.asarray(batch_pix, dtype=np.int32), device=device)

    last_reduced = False     # the last render already holds the sum over the ranks (fused NVLink reduction)

    @staticmethod
    def _multicast(scene, pb, launch, mode=1, extract=None):
        """Run `launch(base address)` with the scene's outputs switched to the multicast address of `pb`, then barrier
        and fetch the summed buffer (dist.PeerBuffers).  mode 2: float4 pixels (include/psdr_b200.h)."""
        L = _lib.load()
        base, _ = pb.target()
        _lib.check(L.psdr_scene_set_output_multicast(scene._h, mode))
        try:
            launch(base)
        finally:
            _lib.check(L.psdr_scene_set_output_multicast(scene._h, 0))
        return pb.finish(extract)

    def renderC(self, scene: Scene, sensor_id: int = 0, seed: int = -1, batch_pix=-1):
        """Primal image, float32[H*W, 3] on the GPU (reference Integrator::renderC)."""
        self._check(scene)
        torch = self._torch()
        dev, st = self._dev_stream(torch, scene)
        pix = self._pix(torch, batch_pix, dev)
        n = scene.opts.width * scene.opts.height if pix is None else pix.numel()
        L = _lib.load()
        args = (scene._h, sensor_id, self.max_depth, int(seed), int(self.hide_emitters),
                None if pix is None else pix.data_ptr(), 0 if pix is None else pix.numel())
        pb = scene._peer("img", 4 * n)
        self.last_reduced = pb is not None
        if pb is not None:
            return self._multicast(scene, pb, lambda base: _lib.check(L.psdr_render_c(*args, base, st)), 2, lambda b: b.view(n, 4)[:, :3].contiguous())
        img = torch.empty((n, 3), dtype=torch.float32, device=dev)
        _lib.check(L.psdr_render_c(*args, img.data_ptr(), st))
        return img

    def renderD_fwd(self, scene: Scene, sensor_id: int = 0, seed: int = -1, batch_pix=-1, terms: int = _lib.TERM_ALL):
        """(image, forward-mode derivative image) for the tangents configured on the scene."""
        self._check(scene)
        torch = self._torch()
        dev, st = self._dev_stream(torch, scene)
        pix = self._pix(torch, batch_pix, dev)
        n = scene.opts.width * scene.opts.height if pix is None else pix.numel()
        L = _lib.load()
        args = (scene._h, sensor_id, self.max_depth, int(seed), int(self.hide_emitters), int(terms), int(self.reference_tangent_scaling),
                None if pix is None else pix.data_ptr(), 0 if pix is None else pix.numel())
        pb = scene._peer("img2", 8 * n)
        self.last_reduced = pb is not None
        if pb is not None:      # fused reduction: the kernels of all ranks add straight into every rank's replica (float4 pixels)
            buf = self._multicast(scene, pb, lambda base: _lib.check(L.psdr_render_d(*args, base, base + 16 * n, st)), 2,
                                  lambda b: b.view(2, n, 4)[:, :, :3].contiguous())
        else:
            # ONE [2, n, 3] buffer: the NCCL path sums image and derivative image with a single all-reduce of it
            buf = torch.empty((2, n, 3), dtype=torch.float32, device=dev)
            _lib.check(L.psdr_render_d(*args, buf[0].data_ptr(), buf[1].data_ptr(), st))
        self.last_buffer = buf
        return buf[0], buf[1]

    def renderD(self, scene: Scene, sensor_id: int = 0, seed: int = -1, batch_pix=-1):
        """reference Integrator::renderD.  If any parameter reachable through ``scene.param_map`` is a torch
        tensor with ``requires_grad`` the returned image carries an autograd node whose backward runs the
        adjoint kernels (``psdr_render_vjp``) -- ``loss(img).backward()`` then fills ``param.grad`` like
        ``drjit.backward`` does in the reference.  Otherwise: primal image, and the forward-mode derivative
        image for the configured tangents is kept in ``self.grad_image``."""
        leaves = scene._grad_leaves() + self._own_leaves()
        if leaves:
            return _render_d_autograd(self, scene, sensor_id, int(seed), batch_pix, leaves)
        img, self.grad_image = self.renderD_fwd(scene, sensor_id, seed, batch_pix)
        return img

    def _own_leaves(self):
        """[(tensor, kind, index)] for differentiable parameters of the integrator itself (CollocatedIntegrator.m_intensity)"""
        return []

    def renderD_primal(self, scene: Scene, sensor_id: int = 0, seed: int = -1, batch_pix=-1):
        """Image of renderD without any derivative (psdr_render_d with dimg = NULL)."""
        self._check(scene)
        torch = self._torch()
        dev, st = self._dev_stream(torch, scene)
        pix = self._pix(torch, batch_pix, dev)
        n = scene.opts.width * scene.opts.height if pix is None else pix.numel()
        L = _lib.load()
        args = (scene._h, sensor_id, self.max_depth, int(seed), int(self.hide_emitters), _lib.TERM_ALL, 0,
                None if pix is None else pix.data_ptr(), 0 if pix is None else pix.numel())
        pb = scene._peer("img", 4 * n)
        self.last_reduced = pb is not None
        if pb is not None:
            return self._multicast(scene, pb, lambda base: _lib.check(L.psdr_render_d(*args, base, None, st)), 2, lambda b: b.view(n, 4)[:, :3].contiguous())
        img = torch.empty((n, 3), dtype=torch.float32, device=dev)
        _lib.check(L.psdr_render_d(*args, img.data_ptr(), None, st))
        return img

    def render_vjp(self, scene: Scene, d_img, sensor_id: int = 0, seed: int = -1, batch_pix=-1, terms: int = _lib.TERM_ALL, group=None):
        """Adjoint pass: accumulates d<d_img, img>/d(parameter) for every parameter; read them with
        ``scene.grad_of(name, field)``.  Replays the sample streams of a forward call with the same seed.
        With a sharded scene (``set_shard(rank, world > 1)``) and an initialised process group the flat device
        gradient table is summed over the ranks with ONE all-reduce (NCCL) before it is copied to the host, so every
        rank ends up with the full gradients (``d_img`` must be the same, complete cotangent image on every rank)."""
        self._check(scene)
        torch = self._torch()
        dev, st = self._dev_stream(torch, scene)
        pix = self._pix(torch, batch_pix, dev)
        d_img = torch.as_tensor(d_img, dtype=torch.float32, device=dev).contiguous()
        n = scene.opts.width * scene.opts.height if pix is None else pix.numel()
        if d_img.numel() != 3 * n:
            raise RuntimeError("cotangent image must have %d x 3 entries" % n)
        L = _lib.load()
        args = (scene._h, sensor_id, self.max_depth, int(seed), int(self.hide_emitters), int(terms), int(self.reference_tangent_scaling),
                None if pix is None else pix.data_ptr(), 0 if pix is None else pix.numel(), d_img.data_ptr())
        if scene._shard[1] > 1 and torch.distributed.is_available() and torch.distributed.is_initialized():
            nt = L.psdr_grad_table_size(scene._h, sensor_id)
            if nt < 0:
                raise RuntimeError(L.psdr_last_error().decode())
            pb = scene._peer("table", nt) if L.psdr_scene_query(scene._h, _lib.Q_GRAD_TABLE_MULTICAST, sensor_id) == 1 else None
            if pb is not None:      # the adjoint kernels of all ranks add into every rank's table (multimem.red)
                table = self._multicast(scene, pb, lambda base: _lib.check(L.psdr_render_vjp_device(*args, base, nt, st)))
            else:
                table = torch.empty(nt, dtype=torch.float32, device=dev)
                _lib.check(L.psdr_render_vjp_device(*args, table.data_ptr(), nt, st))
                torch.distributed.all_reduce(table, group=group)          # the single gradient all-reduce
            _lib.check(L.psdr_scene_backprop_table(scene._h, sensor_id, table.data_ptr(), nt, st))
        else:
            _lib.check(L.psdr_render_vjp(*args, st))

    def render_vjp_table(self, scene: Scene, d_img, sensor_id: int = 0, seed: int = -1, terms: int = _lib.TERM_ALL):
        """The adjoint kernels only: this rank's flat device gradient table (psdr_render_vjp_device), a torch tensor.
        Tables of lane shards add up to the table of the whole frame; ``backprop_table`` finishes the pass."""
        self._check(scene)
        torch = self._torch()
        dev, st = self._dev_stream(torch, scene)
        d_img = torch.as_tensor(d_img, dtype=torch.float32, device=dev).contiguous()
        L = _lib.load()
        nt = L.psdr_grad_table_size(scene._h, sensor_id)
        if nt < 0:
            raise RuntimeError(L.psdr_last_error().decode())
        table = torch.empty(nt, dtype=torch.float32, device=dev)
        _lib.check(L.psdr_render_vjp_device(scene._h, sensor_id, self.max_depth, int(seed), int(self.hide_emitters), int(terms),
                                            int(self.reference_tangent_scaling), None, 0, d_img.data_ptr(), table.data_ptr(), nt, st))
        return table

    def backprop_table(self, scene: Scene, table, sensor_id: int = 0):
        """Host reverse chain of configure() for a (summed) gradient table; afterwards ``scene.grad_of`` reads gradients."""
        torch = self._torch()
        _, st = self._dev_stream(torch, scene)
        _lib.check(_lib.load().psdr_scene_backprop_table(scene._h, sensor_id, table.data_ptr(), table.numel(), st))

    # host-buffer entry points (numpy in/out; used for the end-to-end measurement)
    def renderC_host(self, scene: Scene, sensor_id: int = 0, seed: int = -1, out: Optional[np.ndarray] = None):
        self._check(scene)
        n = scene.opts.width * scene.opts.height
        img = np.empty((n, 3), dtype=np.float32) if out is None else out
        _lib.check(_lib.load().psdr_render_c_host(scene._h, sensor_id, self.max_depth, int(seed), int(self.hide_emitters), None, 0, _fp(img)))
        return img

    def renderD_host(self, scene: Scene, sensor_id: int = 0, seed: int = -1, terms: int = _lib.TERM_ALL, out=None, dout=None):
        self._check(scene)
        n = scene.opts.width * scene.opts.height
        img = np.empty((n, 3), dtype=np.float32) if out is None else out
        dimg = np.empty((n, 3), dtype=np.float32) if dout is None else dout
        _lib.check(_lib.load().psdr_render_d_host(scene._h, sensor_id, self.max_depth, int(seed), int(self.hide_emitters), int(terms),
                                                  int(self.reference_tangent_scaling), None, 0, _fp(img), _fp(dimg)))
        return img, dimg

    def render_aov(self, scene: Scene, sensor_id: int = 0, seed: int = 0):
        self._check(scene)
        torch = self._torch()
        dev, st = self._dev_stream(torch, scene)
        n = scene.opts.width * scene.opts.height * max(scene.opts.spp, 1)
        out = torch.empty((n, 14), dtype=torch.float32, device=dev)
        _lib.check(_lib.load().psdr_render_aov(scene._h, sensor_id, int(seed), out.data_ptr(), st))
        return out


class PathTracer(Integrator):
    """reference src/psdr.cpp:431-434, src/integrator/path.cpp"""

    def __init__(self, max_depth: int = 1):
        if max_depth < 0:
            raise RuntimeError("max_depth >= 0")
        self.max_depth = int(max_depth)
        self.hide_emitters = False
        self._guided = set()          # (id(scene), sensor) pairs this integrator has a guiding grid for

    def preprocess_secondary_edges(self, scene: Scene, sensor_id: int, reso, nrounds: int = 1, seed: int = 0):
        """reference PathTracer::preprocess_secondary_edges (src/integrator/path.cpp:130-168)."""
        if nrounds <= 0:
            raise RuntimeError("nrounds > 0")
        if scene._h is None or not scene.is_ready():
            raise RuntimeError("Scene needs to be configured!")
        torch = self._torch()
        r = np.asarray(list(reso), dtype=np.int32)
        if r.shape != (4,):
            raise RuntimeError("reso = [rx, ry, rz, samples per cell]")
        _lib.check(_lib.load().psdr_preprocess_secondary_edges(scene._h, sensor_id, _ip(r), int(nrounds), int(seed),
                                                               self._dev_stream(torch, scene)[1]))
        self._guided.add((id(scene), sensor_id))

    def guiding_mass(self, scene: Scene, sensor_id: int = 0):
        L = _lib.load()
        n = L.psdr_scene_query(scene._h, _lib.Q_GUIDING_CELLS, sensor_id)
        out = np.zeros(max(n, 0), dtype=np.float32)
        _lib.check(L.psdr_scene_guiding_mass(scene._h, sensor_id, _fp(out), out.size))
        return out

    def _check(self, scene: Scene):
        super()._check(scene)
        for i in range(scene.num_sensors):      # the grid belongs to the integrator in the reference
            _lib.check(_lib.load().psdr_scene_set_guiding(scene._h, i, int((id(scene), i) in getattr(self, "_guided", ()))))


class Direct(PathTracer):
    """reference src/psdr.cpp:436-439, src/integrator/direct.cpp: direct illumination, one bounce.  mis = 2: emitter and
    BSDF sampling combined with the power heuristic, 0: emitter sampling only, 1: BSDF sampling only."""

    def __init__(self, mis: int = 2):
        if not 0 <= int(mis) <= 2:
            raise RuntimeError("mis >= 0 && mis <= 2")
        super().__init__(1)
        self._kind, self._mis = 1, int(mis)


DirectIntegrator = Direct


class CollocatedIntegrator(Integrator):
    """reference src/psdr.cpp:427-429, src/integrator/collocated.cpp: a point light of ``m_intensity`` at the camera --
    Li = BSDF(wi, wo = wi) * intensity / t^2 at the primary hit (flash photography); no emitters are used, Li draws no
    random numbers, and there is no secondary-edge term.  renderC / renderD / forward mode (``d_m_intensity`` is the
    tangent of the intensity) / reverse mode (``grad_intensity(scene)``, or a torch tensor with requires_grad as
    ``m_intensity``) as for PathTracer."""

    def __init__(self, intensity):
        self.m_intensity = intensity if hasattr(intensity, "requires_grad") else np.float32(intensity)
        self.d_m_intensity = np.float32(0.0)
        self.max_depth = 1

    def _check(self, scene: Scene):
        if scene._h is None or not scene.is_ready():
            raise RuntimeError("Input scene must be configured!")
        _lib.check(_lib.load().psdr_scene_set_integrator_collocated(scene._h, float(_f32(self.m_intensity).ravel()[0]), float(_f32(self.d_m_intensity).ravel()[0]), 0))

    def _own_leaves(self):
        t = self.m_intensity
        return [(t, _lib.INTEGRATOR_INTENSITY, 0)] if hasattr(t, "requires_grad") and t.requires_grad else []

    def grad_intensity(self, scene: Scene) -> float:
        """d<d_img, img>/d(m_intensity) of the last ``render_vjp``"""
        return float(scene._read_grad(_lib.INTEGRATOR_INTENSITY, 0, (1,))[0])

    def renderD_fwd(self, scene: Scene, sensor_id: int = 0, seed: int = -1, batch_pix=-1, terms: int = _lib.TERM_ALL):
        return super().renderD_fwd(scene, sensor_id, seed, batch_pix, terms & ~_lib.TERM_SECONDARY_EDGES)

    def render_vjp(self, scene: Scene, d_img, sensor_id: int = 0, seed: int = -1, batch_pix=-1, terms: int = _lib.TERM_ALL, group=None):
        return super().render_vjp(scene, d_img, sensor_id, seed, batch_pix, terms & ~_lib.TERM_SECONDARY_EDGES, group)


class FieldExtractionIntegrator(Integrator):
    """reference src/psdr.cpp:423-425, src/integrator/field.cpp: the value of an intersection field at the primary hit,
    averaged over the pixel's samples.  Fields: silhouette, position, depth, geoNormal, shNormal, uv, segmentation
    (mesh index; the reference reports the id parsed from the mesh's name), optionally "<field> <mesh index>" to keep
    one object only.  renderD: forward mode only -- interior part (the field at the analytically re-intersected primary
    hit, differentiated) + primary-edge part (jump of the field across the pixel-space edges x their normal velocity);
    the derivative image for the tangents configured on the scene is kept in ``grad_image``."""
    FIELDS = ("segmentation", "silhouette", "position", "depth", "geoNormal", "shNormal", "uv")

    def __init__(self, field: str):
        tok = field.split()
        if not tok or tok[0] not in self.FIELDS + ("bsdf",):
            raise RuntimeError("Unsupported field: " + (tok[0] if tok else ""))
        if tok[0] == "bsdf" and len(tok) > 1:
            raise NotImplementedError("FieldExtractionIntegrator('bsdf <object>'): the object filter is not implemented for this field")
        self.field, self.object = tok[0], (tok[1] if len(tok) > 1 else "")
        self.max_depth = 0

    def _check(self, scene: Scene):
        if self.field != "bsdf":
            return super()._check(scene)
        # the "bsdf" field (field.cpp:72-92) is BSDF(wi, wi) at the primary hit: the CollocatedIntegrator's kernels without
        # intensity / t^2
        if scene._h is None or not scene.is_ready():
            raise RuntimeError("Input scene must be configured!")
        _lib.check(_lib.load().psdr_scene_set_integrator_collocated(scene._h, 1.0, 0.0, 1))

    def renderC(self, scene: Scene, sensor_id: int = 0, seed: int = -1, batch_pix=-1):
        if self.field == "bsdf":
            return Integrator.renderC(self, scene, sensor_id, seed, batch_pix)
        torch = self._torch()
        spp = max(scene.opts.spp, 1)
        a = self.render_aov(scene, sensor_id, 0 if seed < 0 else seed).view(-1, spp, 14)
        valid = a[:, :, 0] > 0
        if self.object:
            valid = valid & (a[:, :, 0] == float(int(self.object) + 1))
        f = {"segmentation": a[:, :, 0:1].expand(-1, -1, 3) - 1.0, "silhouette": torch.ones_like(a[:, :, 2:5]), "position": a[:, :, 2:5],
             "depth": a[:, :, 5:6].expand(-1, -1, 3), "geoNormal": a[:, :, 6:9], "shNormal": a[:, :, 9:12],
             "uv": torch.cat((a[:, :, 12:14], torch.zeros_like(a[:, :, 0:1])), dim=2)}[self.field]
        return (f * valid.unsqueeze(-1)).sum(dim=1) / float(spp)

    def _select(self, torch, a):
        """field columns of a [npix, spp, 14] tap (or tangent-tap) array"""
        return {"segmentation": a[:, :, 0:1].expand(-1, -1, 3) - 1.0, "silhouette": torch.ones_like(a[:, :, 2:5]), "position": a[:, :, 2:5],
                "depth": a[:, :, 5:6].expand(-1, -1, 3), "geoNormal": a[:, :, 6:9], "shNormal": a[:, :, 9:12],
                "uv": torch.cat((a[:, :, 12:14], torch.zeros_like(a[:, :, 0:1])), dim=2)}[self.field]

    def renderD_fwd(self, scene: Scene, sensor_id: int = 0, seed: int = -1, batch_pix=-1, terms: int = _lib.TERM_ALL):
        """(field image, forward-mode derivative image) -- Integrator::renderD with Li = the field (field.cpp:47-121)."""
        if self.field == "bsdf":
            return Integrator.renderD_fwd(self, scene, sensor_id, seed, batch_pix, terms & ~_lib.TERM_SECONDARY_EDGES)
        self._check(scene)
        torch = self._torch()
        dev, st = self._dev_stream(torch, scene)
        L = _lib.load()
        spp = max(scene.opts.spp, 1)
        npix = scene.opts.width * scene.opts.height
        sd = 0 if seed < 0 else int(seed)
        a = torch.empty((npix * spp, 14), dtype=torch.float32, device=dev)
        da = torch.empty_like(a)
        _lib.check(L.psdr_render_aov_d(scene._h, sensor_id, sd, a.data_ptr(), da.data_ptr(), st))
        a, da = a.view(-1, spp, 14), da.view(-1, spp, 14)
        valid = a[:, :, 0] > 0
        obj = int(self.object) if self.object else -1
        if obj >= 0:
            valid = valid & (a[:, :, 0] == float(obj + 1))
        m = valid.unsqueeze(-1)
        img = (self._select(torch, a) * m).sum(dim=1) / float(spp)
        dimg = torch.zeros_like(img)
        if (terms & _lib.TERM_INTERIOR) and self.field not in ("segmentation", "silhouette"):
            dimg = dimg + (self._select(torch, da) * m).sum(dim=1) / float(spp)
        if (terms & _lib.TERM_PRIMARY_EDGES) and scene.opts.sppe > 0:
            e = torch.empty((npix, 3), dtype=torch.float32, device=dev)
            _lib.check(L.psdr_render_field_edges(scene._h, sensor_id, sd, self.FIELDS.index(self.field), obj, e.data_ptr(), st))
            dimg = dimg + e
        return img, dimg

    def renderD(self, scene: Scene, sensor_id: int = 0, seed: int = -1, batch_pix=-1):
        """The field image; with torch parameters that require grad (``scene.param_map``) it carries an autograd node whose
        backward runs the adjoint (``render_vjp``) -- e.g. a silhouette / mask loss in an optimisation loop.  Otherwise the
        forward-mode derivative image for the configured tangents is kept in ``grad_image``."""
        leaves = scene._grad_leaves() if self.field != "bsdf" else []
        if leaves:
            return _render_d_autograd(self, scene, sensor_id, int(seed), batch_pix, leaves)
        img, self.grad_image = self.renderD_fwd(scene, sensor_id, seed, batch_pix)
        return img

    last_reduced = True          # (autograd node: nothing to all-reduce, the field integrator runs unsharded)

    def renderD_primal(self, scene: Scene, sensor_id: int = 0, seed: int = -1, batch_pix=-1):
        return self.renderC(scene, sensor_id, seed, batch_pix)

    def render_vjp(self, scene: Scene, d_img, sensor_id: int = 0, seed: int = -1, batch_pix=-1, terms: int = _lib.TERM_ALL, group=None):
        """Adjoint of the field image: accumulates d<d_img, field image>/d(parameter) for every scene parameter (read with
        ``scene.grad_of``): the interior part through the analytically re-intersected primary hit (position, depth, normals,
        uv) and the primary-edge part (every field; the only part of a silhouette)."""
        if self.field == "bsdf":
            raise RuntimeError("FieldExtractionIntegrator('bsdf') has no reverse mode: use the forward-mode derivative image")
        self._check(scene)
        torch = self._torch()
        dev, st = self._dev_stream(torch, scene)
        n = scene.opts.width * scene.opts.height
        d_img = torch.as_tensor(d_img, dtype=torch.float32, device=dev).contiguous()
        if d_img.numel() != 3 * n:
            raise RuntimeError("cotangent image must have %d x 3 entries" % n)
        obj = int(self.object) if self.object else -1
        _lib.check(_lib.load().psdr_render_field_vjp(scene._h, sensor_id, 0 if seed < 0 else int(seed), self.FIELDS.index(self.field), obj, int(terms),
                                                      0, d_img.data_ptr(), st))      # (the forward image carries no reference scaling either)


def _render_d_autograd(integ: Integrator, scene: Scene, sensor_id: int, seed: int, batch_pix, leaves):
    import torch

    class _RenderD(torch.autograd.Function):
        """forward = primal kernels of renderD; backward = adjoint kernels + host chain (psdr_render_vjp)."""

        @staticmethod
        def forward(ctx, *tensors):
            ctx.state0 = scene._sampler_state()
            img = integ.renderD_primal(scene, sensor_id, seed, batch_pix)
            if not integ.last_reduced and scene._shard[1] > 1 and torch.distributed.is_available() and torch.distributed.is_initialized():
                # every rank rendered its lane shard into a full-frame buffer: a loss must see the SUM (a nonlinear
                # loss of a partial image has the wrong cotangent, different on every rank)
                torch.distributed.all_reduce(img)
            return img

        @staticmethod
        def backward(ctx, d_img):
            now = scene._sampler_state()            # the streams as they are when backward starts (other renders may
            if seed == -1:                          # have consumed draws since this node's forward)
                scene._set_sampler_state(ctx.state0)     # replay the streams the forward call consumed
            integ.render_vjp(scene, d_img, sensor_id, seed, batch_pix)     # sharded: one all-reduce of the gradient table
            scene._set_sampler_state(now)
            grads = []
            for t, kind, index in leaves:
                g = scene._read_grad(kind, index, tuple(t.shape))
                grads.append(torch.from_numpy(g).to(device=t.device, dtype=t.dtype))
            return tuple(grads)

    return _RenderD.apply(*[t for t, _, _ in leaves])
