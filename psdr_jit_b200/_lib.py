"""ctypes binding of libpsdr_b200.so (C ABI in include/psdr_b200.h).

There is no CPU fallback: if the shared library is missing it is built with nvcc; if that fails, or
if no CUDA device is present when a scene is created, the error propagates.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None

P_F = C.POINTER(C.c_float)
P_I = C.POINTER(C.c_int)

# parameter kinds / queries / terms (include/psdr_b200.h)
MESH_VERTICES, MESH_TO_WORLD_LEFT, MESH_TO_WORLD_RAW, MESH_TO_WORLD_RIGHT = 0, 1, 2, 3
SENSOR_TO_WORLD_LEFT, SENSOR_TO_WORLD_RAW, SENSOR_TO_WORLD_RIGHT = 4, 5, 6
BSDF_REFLECTANCE, EMITTER_RADIANCE, BSDF_SPECULAR, BSDF_ROUGHNESS = 7, 8, 9, 10
ENVMAP_RADIANCE, ENVMAP_SCALE, ENVMAP_TO_WORLD_LEFT = 11, 12, 13
BSDF_REFLECTANCE_UV, BSDF_SPECULAR_UV, BSDF_ROUGHNESS_UV = 14, 15, 16
BSDF_ETA, BSDF_K, BSDF_PERVERTEX, INTEGRATOR_INTENSITY = 17, 18, 19, 20
NESTED_BSDF_BASE = 1 << 20
TEX_REFLECTANCE, TEX_SPECULAR, TEX_ROUGHNESS = 0, 1, 2
Q_NUM_MESHES, Q_NUM_SENSORS, Q_NUM_EMITTERS, Q_NUM_TRIANGLES, Q_NUM_PRIMARY_EDGES, Q_NUM_SECONDARY_EDGES = 0, 1, 2, 3, 4, 5
Q_NUM_MESH_EDGES, Q_NUM_MESH_VERTICES, Q_NUM_MESH_FACES, Q_IS_CONFIGURED, Q_USES_BVH, Q_UPLOAD_BYTES, Q_GUIDING_CELLS = 6, 7, 8, 9, 10, 11, 12
Q_BVH_BUILDS, Q_BVH_REFITS, Q_GRAD_TABLE_MULTICAST, Q_KERNEL_FAMILY = 13, 14, 15, 16
TERM_INTERIOR, TERM_PRIMARY_EDGES, TERM_SECONDARY_EDGES, TERM_ALL = 1, 2, 4, 7

EXPORTS = [
    "psdr_last_error", "psdr_version", "psdr_kernel_launch_count", "psdr_scene_create", "psdr_scene_destroy",
    "psdr_scene_set_options", "psdr_scene_set_seed", "psdr_scene_set_shard", "psdr_scene_set_accel", "psdr_scene_set_reference_arithmetic", "psdr_scene_set_integrator", "psdr_scene_set_integrator_collocated", "psdr_scene_set_output_multicast", "psdr_set_cta_policy", "psdr_set_edge_sort",
    "psdr_scene_add_bsdf_diffuse", "psdr_scene_add_bsdf_microfacet", "psdr_scene_add_bsdf_roughconductor", "psdr_scene_add_bsdf_roughdielectric", "psdr_scene_add_bsdf_microfacet_pervertex", "psdr_scene_begin_nested_bsdf", "psdr_scene_add_bsdf_normalmap", "psdr_scene_add_envmap", "psdr_scene_set_bsdf_texture", "psdr_scene_set_bsdf_texture_slot", "psdr_scene_add_mesh", "psdr_scene_add_perspective", "psdr_scene_add_perspective_intrinsic", "psdr_scene_add_orthographic", "psdr_scene_set_param",
    "psdr_scene_set_tangent", "psdr_scene_clear_tangents", "psdr_scene_configure", "psdr_scene_last_configure_ms",
    "psdr_scene_query", "psdr_scene_mesh_edges", "psdr_render_c", "psdr_render_d", "psdr_render_c_host",
    "psdr_render_d_host", "psdr_render_aov", "psdr_render_aov_d", "psdr_render_field_edges", "psdr_render_field_vjp", "psdr_sampler_draws", "psdr_scene_enable_timing", "psdr_scene_kernel_ms",
    "psdr_preprocess_secondary_edges", "psdr_scene_set_guiding", "psdr_scene_guiding_mass", "psdr_render_vjp", "psdr_grad_table_size", "psdr_render_vjp_device", "psdr_scene_backprop_table", "psdr_scene_get_grad", "psdr_scene_get_sampler_state", "psdr_scene_set_sampler_state",
]


def lib_path() -> str:
    return _build.LIB


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    # raises if nvcc / sources are missing; PSDR_REFERENCE_ARITHMETIC=1 selects the approximate-division variant (build.py)
    # PSDR_B200_LIB: load exactly this library (A/B runs of kernel variants built beforehand, tools/build_variant.py)
    path = os.environ.get("PSDR_B200_LIB") or _build.build_native(variant="refarith" if os.environ.get("PSDR_REFERENCE_ARITHMETIC") == "1" else "")
    L = C.CDLL(path)
    vp, ll, i, f = C.c_void_p, C.c_longlong, C.c_int, C.c_float
    L.psdr_last_error.restype = C.c_char_p
    L.psdr_kernel_launch_count.restype = ll
    L.psdr_scene_create.restype = vp
    L.psdr_scene_create.argtypes = [i]
    L.psdr_scene_destroy.argtypes = [vp]
    L.psdr_scene_set_options.argtypes = [vp, i, i, i, i, i, i]
    L.psdr_scene_set_seed.argtypes = [vp, ll]
    L.psdr_scene_set_shard.argtypes = [vp, i, i]
    L.psdr_scene_set_accel.argtypes = [vp, i]
    L.psdr_scene_set_reference_arithmetic.argtypes = [vp, i]
    L.psdr_scene_set_integrator.argtypes = [vp, i, i]
    L.psdr_scene_set_integrator_collocated.argtypes = [vp, f, f, i]
    L.psdr_scene_set_output_multicast.argtypes = [vp, i]
    L.psdr_set_cta_policy.argtypes = [i]
    L.psdr_set_edge_sort.argtypes = [i]
    L.psdr_scene_add_bsdf_diffuse.argtypes = [vp, C.c_char_p, P_F, i]
    L.psdr_scene_add_bsdf_microfacet.argtypes = [vp, C.c_char_p, P_F, P_F, f, i]
    L.psdr_scene_add_bsdf_roughconductor.argtypes = [vp, C.c_char_p, f, P_F, P_F, P_F, i]
    L.psdr_scene_add_bsdf_roughdielectric.argtypes = [vp, C.c_char_p, f, f, f, i]
    L.psdr_scene_add_bsdf_microfacet_pervertex.argtypes = [vp, C.c_char_p, P_F, P_F, P_F, i, i]
    L.psdr_scene_begin_nested_bsdf.argtypes = [vp]
    L.psdr_scene_add_bsdf_normalmap.argtypes = [vp, C.c_char_p, P_F, i, i]
    L.psdr_scene_add_envmap.argtypes = [vp, P_F, i, i, P_F, f]
    L.psdr_scene_set_bsdf_texture.argtypes = [vp, i, i, i]
    L.psdr_scene_set_bsdf_texture_slot.argtypes = [vp, i, i, i, i]
    L.psdr_scene_add_mesh.argtypes = [vp, P_F, i, P_I, i, P_F, i, P_I, P_F, C.c_char_p, P_F, i, i]
    L.psdr_scene_add_perspective.argtypes = [vp, f, f, f, P_F]
    L.psdr_scene_add_perspective_intrinsic.argtypes = [vp, f, f, f, f, f, f, P_F]
    L.psdr_scene_add_orthographic.argtypes = [vp, f, f, P_F]
    L.psdr_scene_set_param.argtypes = [vp, i, i, P_F, i]
    L.psdr_scene_set_tangent.argtypes = [vp, i, i, P_F, i]
    L.psdr_scene_clear_tangents.argtypes = [vp]
    L.psdr_scene_configure.argtypes = [vp, P_I, i]
    L.psdr_scene_last_configure_ms.restype = C.c_double
    L.psdr_scene_last_configure_ms.argtypes = [vp]
    L.psdr_scene_query.argtypes = [vp, i, i]
    L.psdr_scene_enable_timing.argtypes = [vp, i]
    L.psdr_scene_kernel_ms.restype = C.c_double
    L.psdr_scene_kernel_ms.argtypes = [vp, i]
    L.psdr_scene_mesh_edges.argtypes = [vp, i, P_I]
    L.psdr_render_c.argtypes = [vp, i, i, ll, i, vp, i, vp, vp]
    L.psdr_render_d.argtypes = [vp, i, i, ll, i, i, i, vp, i, vp, vp, vp]
    L.psdr_render_c_host.argtypes = [vp, i, i, ll, i, P_I, i, P_F]
    L.psdr_render_d_host.argtypes = [vp, i, i, ll, i, i, i, P_I, i, P_F, P_F]
    L.psdr_render_aov.argtypes = [vp, i, ll, vp, vp]
    L.psdr_render_aov_d.argtypes = [vp, i, ll, vp, vp, vp]
    L.psdr_render_field_edges.argtypes = [vp, i, ll, i, i, vp, vp]
    L.psdr_render_field_vjp.argtypes = [vp, i, ll, i, i, i, i, vp, vp]
    L.psdr_sampler_draws.argtypes = [ll, i, i, P_F]
    L.psdr_preprocess_secondary_edges.argtypes = [vp, i, P_I, i, ll, vp]
    L.psdr_scene_set_guiding.argtypes = [vp, i, i]
    L.psdr_scene_guiding_mass.argtypes = [vp, i, P_F, i]
    L.psdr_render_vjp.argtypes = [vp, i, i, ll, i, i, i, vp, i, vp, vp]
    L.psdr_grad_table_size.argtypes = [vp, i]
    L.psdr_render_vjp_device.argtypes = [vp, i, i, ll, i, i, i, vp, i, vp, vp, i, vp]
    L.psdr_scene_backprop_table.argtypes = [vp, i, vp, i, vp]
    L.psdr_scene_get_grad.argtypes = [vp, i, i, P_F, i]
    L.psdr_scene_get_sampler_state.argtypes = [vp, C.POINTER(ll)]
    L.psdr_scene_set_sampler_state.argtypes = [vp, C.POINTER(ll)]
    _LIB = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(load().psdr_last_error().decode())
