// (part a: the interior kernels)  Forward kernels (interior / primary-edge / secondary-edge / guiding / AOV) of configuration 10
// (bit 0: BVH2 traversal, bit 1: Microfacet + EnvironmentMap code, bit 3: extended material set).  See kernels_impl.cuh.
#include "kernels_impl.cuh"
#include "launch_decl.h"

namespace psdr {
namespace fwd10 {
cudaError_t interior(const DScene &sc, const DCamera &cam, const RenderParams &rp, bool ad, float *img, float *dimg, cudaStream_t st) {
    return ForwardLaunch<10>::interior(sc, cam, rp, ad, img, dimg, st);
}
}  // namespace fwd10
}  // namespace psdr
