// sm_100a adjoint kernels: renderD's reverse pass (what drjit.backward(loss(img)) triggers in the
// reference: AD-graph traversal kernels scattering into triangle-record / texel / radiance gradient
// arrays).  One fused kernel per term, same lane -> random-stream mapping as the
// forward kernels (kernels.cu), so a backward call with the forward call's seed replays the same paths.
#include <cuda_runtime.h>

#include "adjoint.cuh"
#include "kernels.h"

namespace psdr {

constexpr int kBlockV = 128;

template <bool kBvh, int kD, bool kSmem>
__global__ void __launch_bounds__(kBlockV) interior_vjp_kernel(const __grid_constant__ DScene sc, const __grid_constant__ DCamera cam,
                                                                const __grid_constant__ RenderParams rp, const __grid_constant__ GradLayout gl,
                                                                const float *__restrict__ d_img) {
    extern __shared__ float smem[];
    const GradAcc acc = grad_acc_begin(gl, smem, 0, gl.off_pe, kSmem);
    const long long stride = (long long) gridDim.x * kBlockV;
    const float inv_spp = (sc.spp > 1 ? 1.f / (float) sc.spp : 1.f) * rp.tangent_scale;
    for (long long i = rp.lane_begin + (long long) blockIdx.x * kBlockV + threadIdx.x; i < rp.lane_end; i += stride) {
        const int idx = (int) (sc.spp > 1 ? i / sc.spp : i);
        const int pix = rp.pix_id ? __ldg(rp.pix_id + idx) : idx;
        const unsigned long long seed_value = rp.pix_id ? (unsigned long long) ((long long) pix + rp.seed) : (unsigned long long) (i + rp.seed);
        Pcg32 rng;
        rng.seed(seed_value, (unsigned long long) i);
        if (rp.skip) rng.advance(rp.skip);
        const float jy = rng.next_1d(), jx = rng.next_1d();
        const float sx = ((float) (pix % sc.width) + jx) / (float) sc.width;
        const float sy = ((float) (pix / sc.width) + jy) / (float) sc.height;
        const V3f dc = normalize(xform_pos(cam.sample_to_camera, V3f(sx, sy, 0.f)));
        const V3f o = xform_pos(cam.to_world, V3f(0.f, 0.f, 0.f)), d = xform_dir(cam.to_world, dc);
        PathRecord<kD> R;
        R.reset();
        const V3f v = Li<float, kBvh, true, PathRecord<kD>>(sc, rng, o, d, true, rp.max_depth, rp.hide_emitters != 0, R);
        // cotangent of this lane's value; channels the forward pass scrubbed (non-finite) carry none
        V3f g(__ldg(d_img + 3 * idx) * inv_spp, __ldg(d_img + 3 * idx + 1) * inv_spp, __ldg(d_img + 3 * idx + 2) * inv_spp);
        if (!isfinite(v.x)) g.x = 0.f;
        if (!isfinite(v.y)) g.y = 0.f;
        if (!isfinite(v.z)) g.z = 0.f;
        if (g.x == 0.f && g.y == 0.f && g.z == 0.f) continue;
        path_adjoint<kD>(sc, gl, acc, R, o, d, dc, g, rp.hide_emitters != 0);
    }
    grad_acc_end(acc);
}

template <bool kBvh, bool kSmem>
__global__ void __launch_bounds__(kBlockV) primary_edge_vjp_kernel(const __grid_constant__ DScene sc, const __grid_constant__ DCamera cam,
                                                                    const __grid_constant__ RenderParams rp, const __grid_constant__ GradLayout gl,
                                                                    const float *__restrict__ d_img) {
    extern __shared__ float smem[];
    const GradAcc acc = grad_acc_begin(gl, smem, gl.off_pe, gl.off_se, kSmem);
    const long long stride = (long long) gridDim.x * kBlockV;
    const float inv_sppe = sc.sppe > 1 ? 1.f / (float) sc.sppe : 1.f;
    for (long long i = rp.lane_begin + (long long) blockIdx.x * kBlockV + threadIdx.x; i < rp.lane_end; i += stride) {
        Pcg32 rng;
        rng.seed((unsigned long long) (i + rp.seed), (unsigned long long) i);
        if (rp.skip) rng.advance(rp.skip);
        float s1 = rng.next_1d(), prob;
        const int ei = sample_reuse(cam.pe_pmf, cam.pe_cmf, cam.n_edges, cam.edge_sum, s1, prob);
        const float4 a = __ldg(cam.pe_a + ei), bq = __ldg(cam.pe_b + ei);
        const float pdf = prob / bq.z;
        const float w0 = 1.0f - s1;
        const float px = fmaf(a.x, w0, a.z * s1), py = fmaf(a.y, w0, a.w * s1);
        const float x_dot_n = fmaf(py, bq.y, px * bq.x);
        const int ix = (int) floorf(px * (float) sc.width), iy = (int) floorf(py * (float) sc.height);
        const bool valid = ix >= 0 && ix < sc.width && iy >= 0 && iy < sc.height;
        V3f Lside[2];
#pragma unroll 1
        for (int side = 0; side < 2; ++side) {
            const float sg = side == 0 ? kEdgeEpsilon : -kEdgeEpsilon;
            V3f ro, rd;
            sample_primary_ray<float>(cam, V2f(px + sg * bq.x, py + sg * bq.y), ro, rd);
            Lside[side] = Li<float, kBvh>(sc, rng, ro, rd, valid, rp.max_depth, rp.hide_emitters != 0);
        }
        if (!valid) continue;
        const int pix = iy * sc.width + ix;
        const float dl[3] = {(Lside[1].x - Lside[0].x) / pdf, (Lside[1].y - Lside[0].y) / pdf, (Lside[1].z - Lside[0].z) / pdf};
        float gsum = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float primal = x_dot_n * dl[c];
            if (!isfinite(primal)) continue;
            gsum += __ldg(d_img + 3 * pix + c) * dl[c];
        }
        gsum *= inv_sppe;
        if (gsum == 0.f || !isfinite(gsum)) continue;
        // x_dot_n = <lerp(p0, p1, s), n>
        const int b = gl.off_pe + 4 * ei;
        acc.add(b, gsum * w0 * bq.x);
        acc.add(b + 1, gsum * w0 * bq.y);
        acc.add(b + 2, gsum * s1 * bq.x);
        acc.add(b + 3, gsum * s1 * bq.y);
    }
    grad_acc_end(acc);
}

template <bool kBvh, bool kSmem>
__global__ void __launch_bounds__(kBlockV) secondary_edge_vjp_kernel(const __grid_constant__ DScene sc, const __grid_constant__ DCamera cam,
                                                                      const __grid_constant__ RenderParams rp, const __grid_constant__ GradLayout gl,
                                                                      const float *__restrict__ d_img) {
    extern __shared__ float smem[];
    SecEdgeAdjoint adj;
    adj.acc = grad_acc_begin(gl, smem, 0, gl.total, kSmem);
    adj.gl = gl;
    adj.d_img = d_img;
    adj.scale = rp.tangent_scale * (sc.sppse > 1 ? 1.f / (float) sc.sppse : 1.f);
    const long long stride = (long long) gridDim.x * kBlockV;
    for (long long i = rp.lane_begin + (long long) blockIdx.x * kBlockV + threadIdx.x; i < rp.lane_end; i += stride) {
        Pcg32 rng;
        rng.seed((unsigned long long) (i + rp.seed), (unsigned long long) i);
        if (rp.skip) rng.advance(rp.skip);
        const float d1 = rng.next_1d(), d2 = rng.next_1d(), d3 = rng.next_1d();
        V3f sample3(d3, d2, d1);
        SecEdgeAdjoint a2 = adj;
        if (cam.guided) {
            const float pdf0 = guide_sample_reuse(cam, sample3);
            if (pdf0 > kEpsilon) a2.scale = adj.scale / pdf0;
        }
        V3f value0, tangent;
        eval_secondary_edge<kBvh, SecEdgeAdjoint>(sc, cam, sample3, value0, tangent, a2);
    }
    grad_acc_end(adj.acc);
}

// ---- launchers -------------------------------------------------------------------------------
static int g_sms = 0;
static int vjp_grid(long long lanes, int blocks_per_sm) {
    if (g_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_sms <= 0) g_sms = 148;
    }
    const long long need = (lanes + kBlockV - 1) / kBlockV, cap = (long long) g_sms * blocks_per_sm;
    return (int) (need < cap ? (need > 0 ? need : 1) : cap);
}
constexpr int kSmemGradMaxFloats = 12 * 1024;   // 48 KB: no opt-in needed

template <class K> static cudaError_t launch_k(K kern, int grid, size_t smem, cudaStream_t st, const DScene &sc, const DCamera &cam,
                                               const RenderParams &rp, const GradLayout &gl, const float *d_img) {
    kern<<<grid, kBlockV, smem, st>>>(sc, cam, rp, gl, d_img);
    return cudaGetLastError();
}

cudaError_t launch_interior_vjp(const DScene &sc, const DCamera &cam, const RenderParams &rp, const GradLayout &gl, const float *d_img,
                                cudaStream_t st) {
    const long long lanes = rp.lane_end - rp.lane_begin;
    if (lanes <= 0) return cudaSuccess;
    const int n = gl.off_pe;
    const bool sm = n <= kSmemGradMaxFloats;
    const size_t bytes = sm ? sizeof(float) * n : 0;
    const int grid = vjp_grid(lanes, 4);
    const bool deep = rp.max_depth > 4;
    if (rp.max_depth > 8) return cudaErrorInvalidValue;
#define PSDR_LAUNCH_I(BVH, D, SM) return launch_k(interior_vjp_kernel<BVH, D, SM>, grid, bytes, st, sc, cam, rp, gl, d_img)
    if (sc.use_bvh) {
        if (deep) { if (sm) PSDR_LAUNCH_I(true, 8, true); else PSDR_LAUNCH_I(true, 8, false); }
        else { if (sm) PSDR_LAUNCH_I(true, 4, true); else PSDR_LAUNCH_I(true, 4, false); }
    } else {
        if (deep) { if (sm) PSDR_LAUNCH_I(false, 8, true); else PSDR_LAUNCH_I(false, 8, false); }
        else { if (sm) PSDR_LAUNCH_I(false, 4, true); else PSDR_LAUNCH_I(false, 4, false); }
    }
#undef PSDR_LAUNCH_I
}

cudaError_t launch_primary_edges_vjp(const DScene &sc, const DCamera &cam, const RenderParams &rp, const GradLayout &gl, const float *d_img,
                                     cudaStream_t st) {
    const long long lanes = rp.lane_end - rp.lane_begin;
    if (lanes <= 0 || cam.n_edges <= 0) return cudaSuccess;
    const int n = gl.off_se - gl.off_pe;
    const bool sm = n <= kSmemGradMaxFloats;
    const size_t bytes = sm ? sizeof(float) * n : 0;
    const int grid = vjp_grid(lanes, 8);
    if (sc.use_bvh) return sm ? launch_k(primary_edge_vjp_kernel<true, true>, grid, bytes, st, sc, cam, rp, gl, d_img)
                              : launch_k(primary_edge_vjp_kernel<true, false>, grid, bytes, st, sc, cam, rp, gl, d_img);
    return sm ? launch_k(primary_edge_vjp_kernel<false, true>, grid, bytes, st, sc, cam, rp, gl, d_img)
              : launch_k(primary_edge_vjp_kernel<false, false>, grid, bytes, st, sc, cam, rp, gl, d_img);
}

cudaError_t launch_secondary_edges_vjp(const DScene &sc, const DCamera &cam, const RenderParams &rp, const GradLayout &gl, const float *d_img,
                                       cudaStream_t st) {
    const long long lanes = rp.lane_end - rp.lane_begin;
    if (lanes <= 0 || sc.n_sec_edges <= 0) return cudaSuccess;
    const int n = gl.total;
    const bool sm = n <= kSmemGradMaxFloats;
    const size_t bytes = sm ? sizeof(float) * n : 0;
    const int grid = vjp_grid(lanes, 8);
    if (sc.use_bvh) return sm ? launch_k(secondary_edge_vjp_kernel<true, true>, grid, bytes, st, sc, cam, rp, gl, d_img)
                              : launch_k(secondary_edge_vjp_kernel<true, false>, grid, bytes, st, sc, cam, rp, gl, d_img);
    return sm ? launch_k(secondary_edge_vjp_kernel<false, true>, grid, bytes, st, sc, cam, rp, gl, d_img)
              : launch_k(secondary_edge_vjp_kernel<false, false>, grid, bytes, st, sc, cam, rp, gl, d_img);
}

}  // namespace psdr
