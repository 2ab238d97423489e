// Forward kernels (interior / primary-edge / secondary-edge / guiding / AOV) of configuration 10
// (bit 0: BVH2 traversal, bit 1: Microfacet + EnvironmentMap code, bit 3: extended material set).  See kernels_impl.cuh.
#include "kernels_impl.cuh"
#include "launch_decl.h"

namespace psdr {
namespace fwd10 {
cudaError_t primary(const DScene &sc, const DCamera &cam, const RenderParams &rp, float *dimg, cudaStream_t st) { return ForwardLaunch<10>::primary(sc, cam, rp, dimg, st); }
cudaError_t secondary(const DScene &sc, const DCamera &cam, const RenderParams &rp, float *dimg, cudaStream_t st) { return ForwardLaunch<10>::secondary(sc, cam, rp, dimg, st); }
cudaError_t guiding(const DScene &sc, const DCamera &cam, const int reso[4], int nrounds, long long seed, float *mass, cudaStream_t st) {
    return ForwardLaunch<10>::guiding(sc, cam, reso, nrounds, seed, mass, st);
}
cudaError_t aov(const DScene &sc, const DCamera &cam, const RenderParams &rp, float *out, cudaStream_t st) { return ForwardLaunch<10>::aov(sc, cam, rp, out, st); }
cudaError_t aov_d(const DScene &sc, const DCamera &cam, const RenderParams &rp, float *out, float *dout, cudaStream_t st) { return ForwardLaunch<10>::aov_d(sc, cam, rp, out, dout, st); }
cudaError_t field_edges(const DScene &sc, const DCamera &cam, const RenderParams &rp, int field, int object, float *dimg, cudaStream_t st) {
    return ForwardLaunch<10>::field_edges(sc, cam, rp, field, object, dimg, st);
}
}  // namespace fwd10
}  // namespace psdr
