// Lane ordering for the primary-edge kernels (kernels_impl.cuh primary_edge_kernel, kernels_vjp_impl.cuh
// primary_edge_vjp_kernel).  Sample i of the term (reference Integrator::render_primary_edges,
// src/integrator/integrator.cpp:144-189) is a pure function of (seed, i): its first draw s1 picks the pixel-space edge AND
// the position on it (PerspectiveCamera::sample_primary_edge reuses the sample, src/sensor/perspective.cpp:117-141), and the
// edge cdf is monotone in s1.  In lane order the 32 rays of a warp therefore start at 32 unrelated places of the image
// (18 of 32 lanes active in the closest-hit scans, profiles/r04b).  Here the lanes are bucketed by the top bits of s1 --
// one counting-sort pass: histogram, exclusive scan, scatter -- and the kernels walk the buckets: a warp's rays start on
// the same stretch of the same edge, hit the same two surfaces and take the same branches.  Which thread evaluates a sample
// does not change its value; only the (already unordered) order of the atomic adds into the image changes.
// The secondary-edge kernels (PathTracer::render_secondary_edges, src/integrator/path.cpp:268-302) get the same treatment: the
// first dimension of a sample -- after the guiding distribution has warped it, if there is one -- selects the edge and the
// point on it (Scene::sample_boundary_segment_direct, src/scene/scene.cpp), so bucketing by it puts the three traces of a
// warp's candidates next to each other; device_path.cuh sec_edge_batches deals the ordered samples to the warps in blocks.
#include <cuda_runtime.h>

#include "device_path.cuh"
#include "kernels.h"
#include "pmath.h"

namespace psdr {
namespace {

constexpr int kSortBlock = 1024, kSortPerThread = 8, kSortChunk = kSortBlock * kSortPerThread;

// pass 1: bucket of every local lane (dead lanes of the last 32-block go to the last bucket) + histogram
// kSecondary: the key is the first dimension of the secondary-edge sample (device_path.cuh sec_edge_draw), else the first draw
template <bool kSecondary>
__global__ void __launch_bounds__(kSortBlock) edge_bucket_kernel(const __grid_constant__ RenderParams rp, const __grid_constant__ DCamera cam, int nb,
                                                                  unsigned short *__restrict__ key, int *__restrict__ hist) {
    extern __shared__ int s_hist[];
    for (int t = threadIdx.x; t < nb; t += kSortBlock) s_hist[t] = 0;
    __syncthreads();
    const long long span = rp.lane_end - rp.lane_begin, stride = (long long) gridDim.x * kSortBlock;
    for (long long j = (long long) blockIdx.x * kSortBlock + threadIdx.x; j < span; j += stride) {
        const long long i = global_lane(rp, j);
        int b = nb - 1;
        if (i < rp.n_lanes) {
            Pcg32 rng;
            rng.seed((unsigned long long) (i + rp.seed), (unsigned long long) i);
            if (rp.skip) rng.advance(rp.skip);
            float x = rng.next_1d();
            if (kSecondary) {
                const float d2 = rng.next_1d(), d3 = rng.next_1d();
                V3f sample3(d3, d2, x);
                if (cam.guided) guide_sample_reuse(cam, sample3);
                x = sample3.x;
            }
            b = max(min((int) (x * (float) nb), nb - 1), 0);
        }
        key[j] = (unsigned short) b;
        atomicAdd(&s_hist[b], 1);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nb; t += kSortBlock)
        if (s_hist[t]) atomicAdd(&hist[t], s_hist[t]);
}

// pass 2 (one CTA): cursor[b] = first slot of bucket b; the histogram is cleared for the next call
__global__ void __launch_bounds__(kSortBlock) edge_scan_kernel(int nb, int *__restrict__ hist, int *__restrict__ cursor) {
    __shared__ int s[2 * kSortBlock];
    const int t = threadIdx.x;
    const int a = 2 * t < nb ? hist[2 * t] : 0, b = 2 * t + 1 < nb ? hist[2 * t + 1] : 0;
    int *cur = s, *nxt = s + kSortBlock;
    cur[t] = a + b;
    __syncthreads();
    for (int d = 1; d < kSortBlock; d <<= 1) {       // inclusive scan of the pair sums
        nxt[t] = cur[t] + (t >= d ? cur[t - d] : 0);
        __syncthreads();
        int *tmp = cur; cur = nxt; nxt = tmp;
    }
    const int before = cur[t] - (a + b);
    if (2 * t < nb) { cursor[2 * t] = before; hist[2 * t] = 0; }
    if (2 * t + 1 < nb) { cursor[2 * t + 1] = before + a; hist[2 * t + 1] = 0; }
}

// pass 3: every CTA takes chunks of kSortChunk lanes; ranks inside (chunk, bucket) from a shared-memory counter, one global
// reservation per non-empty (chunk, bucket)
__global__ void __launch_bounds__(kSortBlock) edge_scatter_kernel(long long span, int nb, const unsigned short *__restrict__ key, int *__restrict__ cursor,
                                                                   int *__restrict__ perm) {
    extern __shared__ int s_cnt[];
    const long long nchunks = (span + kSortChunk - 1) / kSortChunk;
    for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
        for (int t = threadIdx.x; t < nb; t += kSortBlock) s_cnt[t] = 0;
        __syncthreads();
        int bk[kSortPerThread], rk[kSortPerThread];
#pragma unroll
        for (int m = 0; m < kSortPerThread; ++m) {
            const long long j = c * kSortChunk + m * kSortBlock + threadIdx.x;
            bk[m] = -1;
            if (j < span) {
                bk[m] = key[j];
                rk[m] = atomicAdd(&s_cnt[bk[m]], 1);
            }
        }
        __syncthreads();
        for (int t = threadIdx.x; t < nb; t += kSortBlock) {
            const int n = s_cnt[t];
            if (n) s_cnt[t] = atomicAdd(&cursor[t], n);
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < kSortPerThread; ++m) {
            const long long j = c * kSortChunk + m * kSortBlock + threadIdx.x;
            if (bk[m] >= 0) perm[s_cnt[bk[m]] + rk[m]] = (int) j;
        }
        __syncthreads();
    }
}

}  // namespace

// key: span uint16, work: 2 * bins ints (zero before the first call; left zeroed), perm: span ints
cudaError_t launch_edge_sort(const RenderParams &rp, const DCamera &cam, bool secondary, int bins, unsigned short *key, int *work, int *perm, cudaStream_t st) {
    const long long span = rp.lane_end - rp.lane_begin;
    if (span <= 0 || bins < 2 || bins > kEdgeSortMaxBins || span > 2147483647LL) return cudaErrorInvalidValue;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = sizeof(int) * bins;
    const long long need = (span + kSortChunk - 1) / kSortChunk;
    const int grid = (int) (need < 2LL * sms ? need : 2LL * sms);
    if (secondary) edge_bucket_kernel<true><<<grid, kSortBlock, smem, st>>>(rp, cam, bins, key, work);
    else edge_bucket_kernel<false><<<grid, kSortBlock, smem, st>>>(rp, cam, bins, key, work);
    edge_scan_kernel<<<1, kSortBlock, 0, st>>>(bins, work, work + bins);
    edge_scatter_kernel<<<grid, kSortBlock, smem, st>>>(span, bins, key, work + bins, perm);
    return cudaGetLastError();
}

}  // namespace psdr
