"""`import psdr_jit as psdr` -- the reference's module name (reference src/psdr.cpp:100) served by the B200-native
implementation in psdr_jit_b200.  Same classes and call shapes as the reference's README.md:45-107 and tutorials;
arrays come back as light-weight objects with ``.numpy()`` / ``.torch()``.  If Dr.Jit is not installed, a small stand-in
(psdr_jit_b200.compat) provides the handful of Dr.Jit names that code uses around the renderer (FloatD, Matrix4fD,
enable_grad, set_grad, forward_to, grad ...)."""
from __future__ import annotations

import numpy as np

import psdr_jit_b200 as _b
from psdr_jit_b200 import compat as _compat
from psdr_jit_b200 import (AreaLight, Bitmap1fD, Bitmap3fD, DiffuseBSDF, EnvironmentMap, Mesh, MicrofacetBSDF, Object,  # noqa: F401
                           RoughConductorBSDF, RoughDielectricBSDF, MicrofacetBSDFPerVertex, NormalMapBSDF,
                           OrthographicCamera, PerspectiveCamera, RenderOption, Sampler, Scene)

STAND_IN_DRJIT = _compat.install()
__version__ = "0.2.1+b200"


class ArrayXf:
    """What renderC / renderD / grad return: a device image with the accessors reference code uses."""

    def __init__(self, tensor):
        self._t = tensor
        self._grad = None
        self._fwd = None

    def numpy(self):
        return self._t.detach().cpu().numpy()

    def torch(self):
        return self._t

    def __array__(self, dtype=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    def _forward(self, seeds):
        if self._fwd is None:
            raise RuntimeError("forward_to: this image was not produced by renderD")
        self._grad = ArrayXf(self._fwd(seeds))


def _leaf_params(scene):
    """[(object, field, {leaf: tangent})] for every parameter that was assigned a value depending on an AD leaf"""
    out = []
    for objs in (scene._meshes, scene._sensors, scene._bsdfs, scene._emitters):
        for o in objs:
            for field, tans in getattr(o, "_leaf_tangents", {}).items():
                if tans:
                    out.append((o, field, tans))
    return out


class _IntegratorMixin:
    def renderC(self, scene, sensor_id=0, seed=-1, batch_pix=-1):
        return ArrayXf(super().renderC(scene, sensor_id, seed, batch_pix))

    def renderD(self, scene, sensor_id=0, seed=-1, batch_pix=-1):
        params = _leaf_params(scene)
        if not params:
            r = super().renderD(scene, sensor_id, seed, batch_pix)
            return ArrayXf(r)
        state0 = scene._sampler_state()
        img = ArrayXf(self.renderD_primal(scene, sensor_id, seed, batch_pix))
        integ = self

        def fwd(seeds):
            # tangent of every parameter = sum over seeded leaves; one forward-mode pass replaying the same streams
            saved = []
            for o, field, tans in params:
                t = sum(np.float32(s) * tans[leaf] for leaf, s in seeds.items() if leaf in tans)
                saved.append((o, field, getattr(o, "d_" + field)))
                setattr(o, "d_" + field, np.zeros_like(getattr(o, "d_" + field)) + t)
            after = scene._sampler_state()
            scene.configure(scene._last_active)
            if seed == -1:
                scene._set_sampler_state(state0)
            _, dimg = integ.renderD_fwd(scene, sensor_id, seed, batch_pix)
            scene._set_sampler_state(after)
            for o, field, old in saved:
                setattr(o, "d_" + field, old)
            scene.configure(scene._last_active)
            return dimg

        img._fwd = fwd
        return img


class PathTracer(_IntegratorMixin, _b.PathTracer):
    pass


class Direct(_IntegratorMixin, _b.Direct):
    pass


class CollocatedIntegrator(_IntegratorMixin, _b.CollocatedIntegrator):
    """psdr.CollocatedIntegrator(intensity) (reference src/psdr.cpp:427-429); a Dr.Jit (stand-in) float is accepted"""

    def __init__(self, intensity):
        _b.CollocatedIntegrator.__init__(self, float(np.asarray(intensity.numpy() if hasattr(intensity, "numpy") else intensity).ravel()[0]))


class FieldExtractionIntegrator(_b.FieldExtractionIntegrator):
    def renderC(self, scene, sensor_id=0, seed=-1, batch_pix=-1):
        return ArrayXf(super().renderC(scene, sensor_id, seed, batch_pix))
