// Host-callable launchers of the sm_100a kernels (kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include "dscene.h"
#include "grad_layout.h"

namespace psdr {
int scene_cfg(const DScene &sc);
cudaError_t launch_interior(const DScene &sc, const DCamera &cam, const RenderParams &rp, bool ad, float *img, float *dimg, cudaStream_t st);
cudaError_t launch_primary_edges(const DScene &sc, const DCamera &cam, const RenderParams &rp, float *dimg, cudaStream_t st);
cudaError_t launch_secondary_edges(const DScene &sc, const DCamera &cam, const RenderParams &rp, float *dimg, cudaStream_t st);
cudaError_t launch_guiding(const DScene &sc, const DCamera &cam, const int reso[4], int nrounds, long long seed, float *mass, cudaStream_t st);
cudaError_t launch_aov(const DScene &sc, const DCamera &cam, const RenderParams &rp, float *out, cudaStream_t st);
cudaError_t launch_aov_d(const DScene &sc, const DCamera &cam, const RenderParams &rp, float *out, float *dout, cudaStream_t st);
cudaError_t launch_field_edges(const DScene &sc, const DCamera &cam, const RenderParams &rp, int field, int object, float *dimg, cudaStream_t st);

// edge_sort.cu: perm = the local lanes of a primary-edge (secondary-edge) launch bucketed by the sample dimension that
// selects the edge and the point on it (= position along the edge list)
constexpr int kEdgeSortMaxBins = 2048;
cudaError_t launch_edge_sort(const RenderParams &rp, const DCamera &cam, bool secondary, int bins, unsigned short *key, int *work, int *perm, cudaStream_t st);

// reverse mode (kernels_vjp.cu); GradLayout = adjoint.cuh
cudaError_t launch_interior_vjp(const DScene &sc, const DCamera &cam, const RenderParams &rp, const GradLayout &gl, const float *d_img, cudaStream_t st);
cudaError_t launch_primary_edges_vjp(const DScene &sc, const DCamera &cam, const RenderParams &rp, const GradLayout &gl, const float *d_img, cudaStream_t st);
cudaError_t launch_secondary_edges_vjp(const DScene &sc, const DCamera &cam, const RenderParams &rp, const GradLayout &gl, const float *d_img, cudaStream_t st);
}  // namespace psdr
