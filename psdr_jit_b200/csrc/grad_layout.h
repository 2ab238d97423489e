// Layout of the device gradient table the adjoint kernels scatter into (host + device view).
#pragma once

namespace psdr {

// per triangle 24 floats: 0 p0 | 3 e1 | 6 e2 | 9 area | 10 n0 | 13 n1 | 16 n2 | 19 fn | 22,23 pad
constexpr int kGradTri = 24;
constexpr int kGradCam = 40;
constexpr int kGradBsdf = 16;
constexpr int kGradEnvHead = 16;
struct GradLayout {
    float *base;      // global table (device)
    int off_bsdf;     // kGradBsdf floats per BSDF: 0..2 d reflectance (diffuse) | 3 d roughness / alpha | 4..6 d specular | 8..10 d eta | 12..14 d k
    int off_emit;     // 4 floats per emitter (d radiance rgb)
    int off_cam;      // 40 floats: 0..15 d to_world | 16..31 d world_to_sample | 32..34 d pos | 35..37 d dir
    int off_pe;       // 4 floats per primary edge of the rendered sensor (d p0.xy, d p1.xy)
    int off_se;       // 6 floats per secondary edge (d p0, d e1)
    // (texel gradients of textured BSDFs follow the environment map block; their offsets are in DBsdf::tex_goff)
    int off_env;      // environment map: 0 d scale | 1..9 d from_world (3x3) | 16.. d texels (3*w*h); == total if none
    int total;
};

// The adjoint kernels keep a shared-memory copy of their section of the table when it fits (48 KB, no opt-in).
constexpr int kGradSharedMaxFloats = 12 * 1024;
// Multicast tables (psdr_scene_set_output_multicast) are written only by the flush of that shared copy (adjoint.cuh
// grad_acc_end): every section must fit it and nothing may live outside it (environment-map / BSDF texel gradients do).
inline bool grad_table_multicast_ok(const GradLayout &gl) { return gl.total == gl.off_env && gl.off_env <= kGradSharedMaxFloats; }

}  // namespace psdr
