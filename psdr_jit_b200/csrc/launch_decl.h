// Per-configuration launch entry points (defined in kern_cfg<N>.cu / vjp_cfg<N>.cu).
#pragma once
#include <cuda_runtime.h>

#include "dscene.h"
#include "grad_layout.h"

namespace psdr {
#define PSDR_DECL_FWD(NS)                                                                                                                      \
    namespace NS {                                                                                                                             \
    cudaError_t interior(const DScene &sc, const DCamera &cam, const RenderParams &rp, bool ad, float *img, float *dimg, cudaStream_t st);      \
    cudaError_t primary(const DScene &sc, const DCamera &cam, const RenderParams &rp, float *dimg, cudaStream_t st);                           \
    cudaError_t secondary(const DScene &sc, const DCamera &cam, const RenderParams &rp, float *dimg, cudaStream_t st);                         \
    cudaError_t guiding(const DScene &sc, const DCamera &cam, const int reso[4], int nrounds, long long seed, float *mass, cudaStream_t st);   \
    cudaError_t aov(const DScene &sc, const DCamera &cam, const RenderParams &rp, float *out, cudaStream_t st);                                \
    cudaError_t aov_d(const DScene &sc, const DCamera &cam, const RenderParams &rp, float *out, float *dout, cudaStream_t st);                 \
    cudaError_t field_edges(const DScene &sc, const DCamera &cam, const RenderParams &rp, int field, int object, float *dimg, cudaStream_t st); \
    }
PSDR_DECL_FWD(fwd0) PSDR_DECL_FWD(fwd1) PSDR_DECL_FWD(fwd2) PSDR_DECL_FWD(fwd3) PSDR_DECL_FWD(fwd10) PSDR_DECL_FWD(fwd11)
#undef PSDR_DECL_FWD
#define PSDR_DECL_VJP(NS)                                                                                                                              \
    namespace NS {                                                                                                                                     \
    cudaError_t interior(const DScene &sc, const DCamera &cam, const RenderParams &rp, const GradLayout &gl, const float *d_img, cudaStream_t st);     \
    cudaError_t primary(const DScene &sc, const DCamera &cam, const RenderParams &rp, const GradLayout &gl, const float *d_img, cudaStream_t st);      \
    cudaError_t secondary(const DScene &sc, const DCamera &cam, const RenderParams &rp, const GradLayout &gl, const float *d_img, cudaStream_t st);    \
    }
PSDR_DECL_VJP(vjp0) PSDR_DECL_VJP(vjp1) PSDR_DECL_VJP(vjp2) PSDR_DECL_VJP(vjp3) PSDR_DECL_VJP(vjp10) PSDR_DECL_VJP(vjp11)
#undef PSDR_DECL_VJP
}  // namespace psdr
