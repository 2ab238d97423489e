// (part a: the interior kernels)  Forward kernels (interior / primary-edge / secondary-edge / guiding / AOV) of configuration 11
// (bit 0: BVH2 traversal, bit 1: Microfacet + EnvironmentMap code, bit 3: extended material set).  See kernels_impl.cuh.
#include "kernels_impl.cuh"
#include "launch_decl.h"

namespace psdr {
namespace fwd11 {
cudaError_t interior(const DScene &sc, const DCamera &cam, const RenderParams &rp, bool ad, float *img, float *dimg, cudaStream_t st) {
    return ForwardLaunch<11>::interior(sc, cam, rp, ad, img, dimg, st);
}
}  // namespace fwd11
}  // namespace psdr
