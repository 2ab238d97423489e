// Adjoint kernels of configuration 10 (bit 0: BVH2 traversal, bit 1: Microfacet + EnvironmentMap code, bit 3: extended material set).
#include "kernels_vjp_impl.cuh"
#include "launch_decl.h"

namespace psdr {
namespace vjp10 {
cudaError_t primary(const DScene &sc, const DCamera &cam, const RenderParams &rp, const GradLayout &gl, const float *d_img, cudaStream_t st) {
    return AdjointLaunch<10>::primary(sc, cam, rp, gl, d_img, st);
}
cudaError_t secondary(const DScene &sc, const DCamera &cam, const RenderParams &rp, const GradLayout &gl, const float *d_img, cudaStream_t st) {
    return AdjointLaunch<10>::secondary(sc, cam, rp, gl, d_img, st);
}
}  // namespace vjp10
}  // namespace psdr
