// Host-callable launchers (kernels.h): pick the kernel family a scene needs -- bit 0 BVH2 traversal, bit 1
// Microfacet / EnvironmentMap code, bit 3 the extended material set (bitmaps, conductor, dielectric, per-vertex, normal
// map; implies bit 1) -- and forward to its translation unit (kern_cfg*.cu, vjp_cfg*.cu).
#include <cuda_runtime.h>

#include "kernels.h"
#include "launch_decl.h"

namespace psdr {

int scene_cfg(const DScene &sc) { return (sc.use_bvh ? 1 : 0) | (sc.full_features ? 2 : 0) | (sc.ext_features ? 10 : 0); }

#define PSDR_DISPATCH(CALL)                   \
    switch (scene_cfg(sc)) {                  \
        case 0: return fwd0::CALL;            \
        case 1: return fwd1::CALL;            \
        case 2: return fwd2::CALL;            \
        case 3: return fwd3::CALL;            \
        case 10: return fwd10::CALL;          \
        default: return fwd11::CALL;          \
    }

cudaError_t launch_interior(const DScene &sc, const DCamera &cam, const RenderParams &rp, bool ad, float *img, float *dimg, cudaStream_t st) {
    if (rp.lane_end - rp.lane_begin <= 0) return cudaSuccess;
    PSDR_DISPATCH(interior(sc, cam, rp, ad, img, dimg, st))
}
cudaError_t launch_primary_edges(const DScene &sc, const DCamera &cam, const RenderParams &rp, float *dimg, cudaStream_t st) {
    if (rp.lane_end - rp.lane_begin <= 0 || cam.n_edges <= 0) return cudaSuccess;
    PSDR_DISPATCH(primary(sc, cam, rp, dimg, st))
}
cudaError_t launch_secondary_edges(const DScene &sc, const DCamera &cam, const RenderParams &rp, float *dimg, cudaStream_t st) {
    if (rp.lane_end - rp.lane_begin <= 0 || sc.n_sec_edges <= 0) return cudaSuccess;
    PSDR_DISPATCH(secondary(sc, cam, rp, dimg, st))
}
cudaError_t launch_guiding(const DScene &sc, const DCamera &cam, const int reso[4], int nrounds, long long seed, float *mass, cudaStream_t st) {
    if ((long long) reso[0] * reso[1] * reso[2] <= 0) return cudaSuccess;
    PSDR_DISPATCH(guiding(sc, cam, reso, nrounds, seed, mass, st))
}
cudaError_t launch_aov(const DScene &sc, const DCamera &cam, const RenderParams &rp, float *out, cudaStream_t st) {
    if (rp.lane_end - rp.lane_begin <= 0) return cudaSuccess;
    PSDR_DISPATCH(aov(sc, cam, rp, out, st))
}
cudaError_t launch_aov_d(const DScene &sc, const DCamera &cam, const RenderParams &rp, float *out, float *dout, cudaStream_t st) {
    if (rp.lane_end - rp.lane_begin <= 0) return cudaSuccess;
    PSDR_DISPATCH(aov_d(sc, cam, rp, out, dout, st))
}
cudaError_t launch_field_edges(const DScene &sc, const DCamera &cam, const RenderParams &rp, int field, int object, float *dimg, cudaStream_t st) {
    if (rp.lane_end - rp.lane_begin <= 0 || cam.n_edges <= 0) return cudaSuccess;
    PSDR_DISPATCH(field_edges(sc, cam, rp, field, object, dimg, st))
}
#undef PSDR_DISPATCH

#define PSDR_DISPATCH_V(CALL)                 \
    switch (scene_cfg(sc)) {                  \
        case 0: return vjp0::CALL;            \
        case 1: return vjp1::CALL;            \
        case 2: return vjp2::CALL;            \
        case 3: return vjp3::CALL;            \
        case 10: return vjp10::CALL;          \
        default: return vjp11::CALL;          \
    }
cudaError_t launch_interior_vjp(const DScene &sc, const DCamera &cam, const RenderParams &rp, const GradLayout &gl, const float *d_img, cudaStream_t st) {
    if (rp.lane_end - rp.lane_begin <= 0) return cudaSuccess;
    if (rp.max_depth > 8) return cudaErrorInvalidValue;
    PSDR_DISPATCH_V(interior(sc, cam, rp, gl, d_img, st))
}
cudaError_t launch_primary_edges_vjp(const DScene &sc, const DCamera &cam, const RenderParams &rp, const GradLayout &gl, const float *d_img, cudaStream_t st) {
    if (rp.lane_end - rp.lane_begin <= 0 || cam.n_edges <= 0) return cudaSuccess;
    PSDR_DISPATCH_V(primary(sc, cam, rp, gl, d_img, st))
}
cudaError_t launch_secondary_edges_vjp(const DScene &sc, const DCamera &cam, const RenderParams &rp, const GradLayout &gl, const float *d_img, cudaStream_t st) {
    if (rp.lane_end - rp.lane_begin <= 0 || sc.n_sec_edges <= 0) return cudaSuccess;
    PSDR_DISPATCH_V(secondary(sc, cam, rp, gl, d_img, st))
}
#undef PSDR_DISPATCH_V

}  // namespace psdr
