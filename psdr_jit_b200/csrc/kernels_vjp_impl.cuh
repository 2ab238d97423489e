// sm_100a adjoint kernels: renderD's reverse pass (what drjit.backward(loss(img)) triggers in the
// reference: AD-graph traversal kernels scattering into triangle-record / texel / radiance gradient
// arrays).  One fused kernel per term, same lane -> random-stream mapping as the
// forward kernels (kernels.cu), so a backward call with the forward call's seed replays the same paths.
#pragma once
#include <cuda_runtime.h>

#include "adjoint.cuh"
#include "kernels.h"

namespace psdr {

constexpr int kBlockV = 128;
// CTA shapes + block barriers of the adjoint kernels, for the same reason as the forward kernels (kernels_impl.cuh):
// keep the warps of an SM inside the same stretch of code.  The interior adjoint has two phases per path -- primal replay,
// reverse sweep -- with a barrier in front of each (6.76 -> 5.28 ms at 2 x 256 threads); primary-edge adjoint 6.76 ->
// 6.01 ms (1 x 1024); secondary-edge adjoint 1.68 -> 1.35 ms (2 x 384).  profiles/r02l, r02n.
#ifndef PSDR_LB_IVJP
#define PSDR_LB_IVJP 2          // resident CTAs per SM the interior adjoint is compiled for
#endif
#ifndef PSDR_BLOCK_IVJP
#define PSDR_BLOCK_IVJP 256
#endif
#ifndef PSDR_VJP_PHASE_SYNC
#define PSDR_VJP_PHASE_SYNC 1   // 1: block barrier before the primal replay and before the reverse sweep; 2: + every replay step; 3: + every sweep step
#endif
#ifndef PSDR_BLOCK_PVJP
#define PSDR_BLOCK_PVJP 1024
#endif
#ifndef PSDR_LB_PVJP
#define PSDR_LB_PVJP 1
#endif
#ifndef PSDR_SYNC_PVJP
#define PSDR_SYNC_PVJP 1
#endif
#ifndef PSDR_BLOCK_SVJP
#define PSDR_BLOCK_SVJP 384
#endif
#ifndef PSDR_LB_SVJP
#define PSDR_LB_SVJP 2
#endif
#ifndef PSDR_SYNC_SVJP
#define PSDR_SYNC_SVJP 1
#endif
constexpr int kBlockIV = PSDR_BLOCK_IVJP, kBlockPV = PSDR_BLOCK_PVJP, kBlockSV = PSDR_BLOCK_SVJP;

// kBig: the large-CTA shape with block barriers (above); else 128 threads, for launches too small to fill it
// kColloc: CollocatedIntegrator (device_path.cuh Li_collocated): the "path" is the primary hit alone; 128-thread shape only
template <int kCfg, int kD, bool kBig, bool kColloc = false>
__global__ void __launch_bounds__(kBig ? kBlockIV : kBlockV, kBig ? PSDR_LB_IVJP : 5) interior_vjp_kernel(const __grid_constant__ DScene sc, const __grid_constant__ DCamera cam,
                                                                const __grid_constant__ RenderParams rp, const __grid_constant__ GradLayout gl,
                                                                const float *__restrict__ d_img) {
    constexpr int kBlockIV = kBig ? psdr::kBlockIV : kBlockV, kSync = kBig ? PSDR_VJP_PHASE_SYNC : 0;
    extern __shared__ float smem[];
    brute_init<kCfg>(sc, kBlockIV);
    const GradAcc acc = grad_acc_begin(gl, smem, 0, gl.off_pe, rp.smem_grad != 0, true, rp.out_multicast);
    const long long stride = (long long) gridDim.x * kBlockIV;
    const float inv_spp = (sc.spp > 1 ? 1.f / (float) sc.spp : 1.f) * rp.tangent_scale;
    // every lane of a warp runs the same number of iterations (the span is padded to 32) and the warp re-converges
    // at the top of each one: without the barrier lanes that finish a path early run ahead into the next lane's
    // closest-hit scans and the warp stays split (profiles/r01d: 5 of 32 lanes active in the adjoint kernel)
    const long long span = rp.lane_end - rp.lane_begin, span_pad = (span + kBlockIV - 1) / kBlockIV * kBlockIV;   // CTA-uniform trip count
    // large CTAs: chunks handed out dynamically (device_path.cuh ChunkSched; its barriers are the block barrier in front
    // of the replay phase); 128-thread CTAs: the rotated static schedule (rotated_lane)
    __shared__ long long s_chunk;
    const bool dynamic = kSync != 0 && rp.sched != nullptr;
    ChunkSched sched{rp.sched};
    const long long iters = (span_pad + stride - 1) / stride;
    for (long long k = 0; dynamic || k < iters; ++k) {
        long long j;
        if (dynamic) {
            j = sched.next(&s_chunk) * kBlockIV + threadIdx.x;
            if (j >= span_pad) break;
        } else {
            j = rotated_lane(k, kBlockIV);     // (CTA-uniform: span_pad is a multiple of the CTA size)
            if (j >= span_pad) continue;
        }
        // the whole body is warp-uniform control flow: lanes past the end of the span ride along inactive (the
        // span is padded to 32), so the barriers here and inside path_adjoint are full-mask barriers at the top
        // level -- the only form that really re-converges the warp (see path_adjoint)
        if (kSync && !dynamic) __syncthreads();
        else if (!kSync) __syncwarp();
        const long long gi = global_lane(rp, j);
        const bool live = j < span && gi < rp.n_lanes;
        const long long i = live ? gi : 0;
        const int idx = (int) (sc.spp > 1 ? i / sc.spp : i);
        const int pix = rp.pix_id ? __ldg(rp.pix_id + idx) : idx;
        const unsigned long long seed_value = rp.pix_id ? (unsigned long long) ((long long) pix + rp.seed) : (unsigned long long) (i + rp.seed);
        Pcg32 rng;
        rng.seed(seed_value, (unsigned long long) i);
        if (rp.skip) rng.advance(rp.skip);
        const float jy = rng.next_1d(), jx = rng.next_1d();
        const float sx = ((float) (pix % sc.width) + jx) / (float) sc.width;
        const float sy = ((float) (pix / sc.width) + jy) / (float) sc.height;
        V3f oc, dc;
        camera_ray_local(cam, V2f(sx, sy), oc, dc);
        const V3f o = xform_pos(cam.to_world, oc), d = xform_dir(cam.to_world, dc);
        PathRecord<kD> R;
        R.reset();
        V3f v(0.f, 0.f, 0.f);
        if (kColloc) {
            const Its<float> its = ray_intersect<float, kCfg, true>(sc, o, d, live, false);
            if (its.valid) {
                R.vertex(0, its.tri, its.bu, its.bv);
                v = bsdf_eval<float, kCfg>(sc, its, its.wi, true) / sqr(its.t) * sc.colloc_intensity;
            }
        } else v = Li<float, kCfg, true, PathRecord<kD>>(sc, rng, o, d, live, rp.max_depth, rp.hide_emitters != 0, R, 0xffffffffu, rp.mis, kSync >= 2);
        // cotangent of this lane's value; channels the forward pass scrubbed (non-finite) carry none
        V3f g(__ldg(d_img + 3 * idx) * inv_spp, __ldg(d_img + 3 * idx + 1) * inv_spp, __ldg(d_img + 3 * idx + 2) * inv_spp);
        if (!isfinite(v.x)) g.x = 0.f;
        if (!isfinite(v.y)) g.y = 0.f;
        if (!isfinite(v.z)) g.z = 0.f;
        const bool has_cotangent = !(g.x == 0.f && g.y == 0.f && g.z == 0.f);
        if (kSync) __syncthreads();
        else __syncwarp();
        path_adjoint<kD, kCfg, kColloc>(sc, gl, acc, R, o, d, dc, g, rp.hide_emitters != 0, live && has_cotangent, oc);
    }
    if (dynamic) sched.finish();
    grad_acc_end(acc);
}

// kField: Li = the field of FieldExtractionIntegrator at the primary hit (rp.field, rp.field_object); 128-thread shape only
template <int kCfg, bool kBig, bool kColloc = false, bool kField = false>
__global__ void __launch_bounds__(kBig ? kBlockPV : kBlockV, kBig ? PSDR_LB_PVJP : 8) primary_edge_vjp_kernel(const __grid_constant__ DScene sc, const __grid_constant__ DCamera cam,
                                                                    const __grid_constant__ RenderParams rp, const __grid_constant__ GradLayout gl,
                                                                    const float *__restrict__ d_img) {
    constexpr int kBlockPV = kBig ? psdr::kBlockPV : kBlockV, kSync = kBig ? PSDR_SYNC_PVJP : 0;
    extern __shared__ float smem[];
    brute_init<kCfg>(sc, kBlockPV);
    const GradAcc acc = grad_acc_begin(gl, smem, gl.off_pe, gl.off_se, rp.smem_grad != 0, false, rp.out_multicast);
    const long long stride = (long long) gridDim.x * kBlockPV;
    const float inv_sppe = sc.sppe > 1 ? 1.f / (float) sc.sppe : 1.f;
    // every lane of a warp runs the same number of iterations (the span is padded to 32) and the warp re-converges
    // at the top of each one: without the barrier lanes that finish a path early run ahead into the next lane's
    // closest-hit scans and the warp stays split (profiles/r01d: 5 of 32 lanes active in the adjoint kernel)
    const long long span = rp.lane_end - rp.lane_begin, span_pad = (span + kBlockPV - 1) / kBlockPV * kBlockPV;
    // large CTAs: dynamic chunk hand-out, as in the forward kernel (kernels_impl.cuh primary_edge_kernel)
    __shared__ long long s_chunk;
    const bool dynamic = kSync != 0 && rp.sched != nullptr;
    ChunkSched sched{rp.sched};
    for (long long k = 0;; ++k) {
        long long j;
        if (dynamic) {
            j = sched.next(&s_chunk) * kBlockPV + threadIdx.x;
            if (j >= span_pad) break;
        } else {
            j = (long long) blockIdx.x * kBlockPV + threadIdx.x + k * stride;
            if (j >= span_pad) break;
            if (kSync) __syncthreads();
            else __syncwarp();
        }
        const long long i = global_lane(rp, rp.perm && j < span ? (long long) __ldg(rp.perm + j) : j);
        const bool live = j < span && i < rp.n_lanes;
        const unsigned live_mask = __ballot_sync(0xffffffffu, live);
        if (!kSync && !live) continue;
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
        const bool valid = live && ix >= 0 && ix < sc.width && iy >= 0 && iy < sc.height;
        V3f Lside[2];
#pragma unroll 1
        for (int side = 0; side < 2; ++side) {
            if (kSync) __syncthreads();
            else __syncwarp(live_mask);
            const float sg = side == 0 ? kEdgeEpsilon : -kEdgeEpsilon;
            V3f ro, rd;
            sample_primary_ray<float>(cam, V2f(px + sg * bq.x, py + sg * bq.y), ro, rd);
            if (kField) Lside[side] = field_value(ray_intersect<float, kCfg>(sc, ro, rd, valid, false), rp.field, rp.field_object);
            else if (kColloc) Lside[side] = Li_collocated<float, kCfg, false>(sc, ro, rd, valid);
            else Lside[side] = Li<float, kCfg>(sc, rng, ro, rd, valid, rp.max_depth, rp.hide_emitters != 0, rp.mis);
        }
        float gsum = 0.f;
        if (valid) {
            const int pix = iy * sc.width + ix;
            const float inv_pdf = 1.f / pdf;
            const float dl[3] = {(Lside[1].x - Lside[0].x) * inv_pdf, (Lside[1].y - Lside[0].y) * inv_pdf, (Lside[1].z - Lside[0].z) * inv_pdf};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float primal = x_dot_n * dl[c];
                if (!isfinite(primal)) continue;
                gsum += __ldg(d_img + 3 * pix + c) * dl[c];
            }
            gsum *= inv_sppe;
        }
        const bool ok = valid && gsum != 0.f && isfinite(gsum);
        // x_dot_n = <lerp(p0, p1, s), n>: four table entries per edge.  With the lanes ordered along the edge list (rp.perm) a
        // warp usually holds ONE edge, and 32 lanes adding into the same four shared-memory words is a 32-way conflict of
        // compare-and-swap loops: such warps add once, after a shuffle reduction.
        __syncwarp(live_mask);
        bool merged = false;
        if (kSync || live_mask == 0xffffffffu) {        // all 32 lanes are here
            const unsigned okm = __ballot_sync(0xffffffffu, ok);
            const int e0 = __shfl_sync(0xffffffffu, ei, okm ? __ffs(okm) - 1 : 0);
            if (okm != 0u && __popc(okm) > 2 && __all_sync(0xffffffffu, !ok || ei == e0)) {
                float ga = ok ? gsum * w0 : 0.f, gb = ok ? gsum * s1 : 0.f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    ga += __shfl_xor_sync(0xffffffffu, ga, o);
                    gb += __shfl_xor_sync(0xffffffffu, gb, o);
                }
                if ((threadIdx.x & 31) == 0) {
                    const float4 be = __ldg(cam.pe_b + e0);
                    const int b = gl.off_pe + 4 * e0;
                    acc.add(b, ga * be.x);
                    acc.add(b + 1, ga * be.y);
                    acc.add(b + 2, gb * be.x);
                    acc.add(b + 3, gb * be.y);
                }
                merged = true;
            }
        }
        if (merged || !ok) continue;
        const int b = gl.off_pe + 4 * ei;
        acc.add(b, gsum * w0 * bq.x);
        acc.add(b + 1, gsum * w0 * bq.y);
        acc.add(b + 2, gsum * s1 * bq.x);
        acc.add(b + 3, gsum * s1 * bq.y);
    }
    if (dynamic) sched.finish();
    grad_acc_end(acc);
}

// ---- FieldExtractionIntegrator, interior part in reverse mode: the adjoint of the analytically re-intersected primary hit
// (aov_d_kernel's taps: p = o + t d, t, face normal, shading normal at the differentiable (u, v), uv) for the cotangent of the
// field image.  Same sample positions as aov_kernel (seed + lane, no stream continuation).
template <int kCfg>
__global__ void __launch_bounds__(kBlockV, 5) field_vjp_kernel(const __grid_constant__ DScene sc, const __grid_constant__ DCamera cam,
                                                               const __grid_constant__ RenderParams rp, const __grid_constant__ GradLayout gl,
                                                               const float *__restrict__ d_img) {
    extern __shared__ float smem[];
    brute_init<kCfg>(sc, kBlockV);
    const GradAcc acc = grad_acc_begin(gl, smem, 0, gl.off_pe, rp.smem_grad != 0, true, 0);
    const long long stride = (long long) gridDim.x * kBlockV;
    const float inv_spp = (sc.spp > 1 ? 1.f / (float) sc.spp : 1.f) * rp.tangent_scale;
    const long long span = rp.lane_end - rp.lane_begin, span_pad = (span + 31) / 32 * 32;
    const int field = rp.field;
    for (long long j = (long long) blockIdx.x * kBlockV + threadIdx.x; j < span_pad; j += stride) {
        __syncwarp();
        const long long i = global_lane(rp, j);
        const bool live = j < span && i < rp.n_lanes;
        const long long il = live ? i : 0;
        const int idx = (int) (sc.spp > 1 ? il / sc.spp : il);
        Pcg32 rng;
        rng.seed((unsigned long long) (il + rp.seed), (unsigned long long) il);
        const float jy = rng.next_1d(), jx = rng.next_1d();
        const float sx = ((float) (idx % sc.width) + jx) / (float) sc.width, sy = ((float) (idx / sc.width) + jy) / (float) sc.height;
        V3f oc, dc;
        camera_ray_local(cam, V2f(sx, sy), oc, dc);
        const V3f o = xform_pos(cam.to_world, oc), d = xform_dir(cam.to_world, dc);
        const Its<float> its = ray_intersect<float, kCfg, true>(sc, o, d, live, false);
        const V3f g(__ldg(d_img + 3 * idx) * inv_spp, __ldg(d_img + 3 * idx + 1) * inv_spp, __ldg(d_img + 3 * idx + 2) * inv_spp);
        const bool on = live && its.valid && (rp.field_object < 0 || its.mesh == rp.field_object) && field >= 2 &&
                        !(g.x == 0.f && g.y == 0.f && g.z == 0.f);
        // lanes without work ride along with zeros: the camera-ray scatter below is warp-uniform
        V3f o_bar(0.f, 0.f, 0.f), d_bar(0.f, 0.f, 0.f);
        if (on) {
            const TriRec<float> T = load_tri<float>(sc, its.tri);
            float u, v, t;
            ray_intersect_triangle<float>(T.p0, T.e1, T.e2, o, d, u, v, t);
            float u_bar = 0.f, v_bar = 0.f, t_bar = 0.f;
            const int b = kGradTri * its.tri;
            if (field == 2) {                                   // position: p = o + t d
                o_bar = g;
                d_bar = g * t;
                t_bar = dot(d, g);
            } else if (field == 3) t_bar = g.x + g.y + g.z;     // depth, replicated to the three channels
            else if (field == 4) acc.add3(b + 19, g);           // geoNormal: the face normal record
            else if (field == 5) {                              // shNormal: normalised interpolated normal at (u, v)
                const VtxGeo geo = vertex_geo<kCfg>(sc, its.tri, u, v);
                V3f m_bar;
                scatter_shading_normal(acc, geo, g, m_bar);
                if (!geo.face_normals) {
                    const ShadeRec<float> N = load_shade<float>(sc, its.tri);
                    u_bar = dot(N.n1 - N.n0, m_bar);
                    v_bar = dot(N.n2 - N.n0, m_bar);
                }
            } else if (field == 6 && (sc.meshes[its.mesh].flags & 2)) {      // uv = uv0 + u duv0 + v duv1
                const float2 t0 = __ldg(sc.uv + 3 * its.tri), t1 = __ldg(sc.uv + 3 * its.tri + 1), t2 = __ldg(sc.uv + 3 * its.tri + 2);
                u_bar = g.x * (t1.x - t0.x) + g.y * (t1.y - t0.y);
                v_bar = g.x * (t2.x - t0.x) + g.y * (t2.y - t0.y);
            }
            const V3f r_bar = isect_adj(T.e1, T.e2, d, u_bar, v_bar, t_bar);
            scatter_isect_tri(acc, its.tri, u, v, r_bar);
            o_bar = o_bar + r_bar;
            d_bar = d_bar + r_bar * t;
        }
        __syncwarp();
        scatter_camera_ray(acc, gl, dc, o_bar, d_bar, oc);
    }
    grad_acc_end(acc);
}

template <int kCfg, bool kBig>
__global__ void __launch_bounds__(kBig ? kBlockSV : kBlockV, kBig ? PSDR_LB_SVJP : 6) secondary_edge_vjp_kernel(const __grid_constant__ DScene sc, const __grid_constant__ DCamera cam,
                                                                      const __grid_constant__ RenderParams rp, const __grid_constant__ GradLayout gl,
                                                                      const float *__restrict__ d_img) {
    constexpr int kBlockSV = kBig ? psdr::kBlockSV : kBlockV, kSync = kBig ? PSDR_SYNC_SVJP : 0;
    extern __shared__ float smem[];
    brute_init<kCfg>(sc, kBlockSV);
    SecEdgeAdjoint adj;
    adj.acc = grad_acc_begin(gl, smem, 0, gl.off_env, rp.smem_grad != 0, false, rp.out_multicast);
    adj.gl = gl;
    adj.d_img = d_img;
    adj.scale = rp.tangent_scale * (sc.sppse > 1 ? 1.f / (float) sc.sppse : 1.f);
    // batches of stage-0 survivors, as in the forward kernel (kernels_impl.cuh sec_edge_batches); this kernel needs the
    // warp-uniform scan loop to keep its lanes together across the three traces of stage 1 (device_path.cuh trace())
    constexpr int kC = kCfg | kCfgUniformScan;
    sec_edge_batches<kC>(sc, cam, rp, kBlockSV, [&](const SecSample &smp) {
        SecEdgeAdjoint a2 = adj;
        if (cam.guided && smp.pdf0 > kEpsilon) a2.scale = adj.scale / smp.pdf0;
        V3f value0, tangent;
        sec_edge_stage1<kC, SecEdgeAdjoint>(sc, cam, smp.cand, value0, tangent, a2);
    }, kSync != 0);
    grad_acc_end(adj.acc);
}

// ---- per-configuration launchers (one translation unit per configuration: vjp_cfg*.cu) -----------------
template <class K> inline int vjp_grid(K kernel, size_t smem, long long lanes, int kBlockV = psdr::kBlockV) {   // resident CTAs only (see kernels_impl.cuh)
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlockV, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    const long long need = (lanes + kBlockV - 1) / kBlockV, cap = (long long) sms * per_sm;
    return (int) (need < cap ? (need > 0 ? need : 1) : cap);
}
constexpr int kSmemGradMaxFloats = kGradSharedMaxFloats;   // 48 KB: no opt-in needed (grad_layout.h)

extern int g_cta_policy;        // kernels_impl.cuh use_big_cta
inline bool vjp_big_cta(long long lanes, int block) {
    if (g_cta_policy == 1) return false;
    if (g_cta_policy == 2) return true;
    return lanes >= 4LL * 148 * block;
}

template <int kCfg> struct AdjointLaunch {
    template <int kD> static void interior_as(const DScene &sc, const DCamera &cam, const RenderParams &rq, const GradLayout &gl, const float *d_img, size_t bytes, cudaStream_t st) {
        const long long n = rq.lane_end - rq.lane_begin;
        if (vjp_big_cta(n, 2 * kBlockIV)) interior_vjp_kernel<kCfg, kD, true><<<vjp_grid(interior_vjp_kernel<kCfg, kD, true>, bytes, n, kBlockIV), kBlockIV, bytes, st>>>(sc, cam, rq, gl, d_img);
        else interior_vjp_kernel<kCfg, kD, false><<<vjp_grid(interior_vjp_kernel<kCfg, kD, false>, bytes, n), kBlockV, bytes, st>>>(sc, cam, rq, gl, d_img);
    }
    static cudaError_t interior(const DScene &sc, const DCamera &cam, const RenderParams &rp, const GradLayout &gl, const float *d_img, cudaStream_t st) {
        RenderParams rq = rp;
        rq.smem_grad = gl.off_pe <= kSmemGradMaxFloats ? 1 : 0;
        const size_t bytes = rq.smem_grad ? sizeof(float) * gl.off_pe : 0;
        if (rp.use_field) {     // FieldExtractionIntegrator: the adjoint of the primary hit alone
            const long long n = rq.lane_end - rq.lane_begin;
            field_vjp_kernel<kCfg><<<vjp_grid(field_vjp_kernel<kCfg>, bytes, n), kBlockV, bytes, st>>>(sc, cam, rq, gl, d_img);
            return cudaGetLastError();
        }
        if (rp.mis == 3) {      // CollocatedIntegrator
            const long long n = rq.lane_end - rq.lane_begin;
            interior_vjp_kernel<kCfg, 1, false, true><<<vjp_grid(interior_vjp_kernel<kCfg, 1, false, true>, bytes, n), kBlockV, bytes, st>>>(sc, cam, rq, gl, d_img);
            return cudaGetLastError();
        }
        if (rp.max_depth > 4) interior_as<8>(sc, cam, rq, gl, d_img, bytes, st);
        else interior_as<4>(sc, cam, rq, gl, d_img, bytes, st);
        return cudaGetLastError();
    }
    static cudaError_t primary(const DScene &sc, const DCamera &cam, const RenderParams &rp, const GradLayout &gl, const float *d_img, cudaStream_t st) {
        RenderParams rq = rp;
        const int n = gl.off_se - gl.off_pe;
        rq.smem_grad = n <= kSmemGradMaxFloats ? 1 : 0;
        const size_t bytes = rq.smem_grad ? sizeof(float) * n : 0;
        const long long lanes = rp.lane_end - rp.lane_begin;
        if (rp.use_field) {     // FieldExtractionIntegrator
            primary_edge_vjp_kernel<kCfg, false, false, true><<<vjp_grid(primary_edge_vjp_kernel<kCfg, false, false, true>, bytes, lanes), kBlockV, bytes, st>>>(sc, cam, rq, gl, d_img);
            return cudaGetLastError();
        }
        if (rp.mis == 3) {      // CollocatedIntegrator
            primary_edge_vjp_kernel<kCfg, false, true><<<vjp_grid(primary_edge_vjp_kernel<kCfg, false, true>, bytes, lanes), kBlockV, bytes, st>>>(sc, cam, rq, gl, d_img);
            return cudaGetLastError();
        }
        if (vjp_big_cta(lanes, kBlockPV)) primary_edge_vjp_kernel<kCfg, true><<<vjp_grid(primary_edge_vjp_kernel<kCfg, true>, bytes, lanes, kBlockPV), kBlockPV, bytes, st>>>(sc, cam, rq, gl, d_img);
        else primary_edge_vjp_kernel<kCfg, false><<<vjp_grid(primary_edge_vjp_kernel<kCfg, false>, bytes, lanes), kBlockV, bytes, st>>>(sc, cam, rq, gl, d_img);
        return cudaGetLastError();
    }
    static cudaError_t secondary(const DScene &sc, const DCamera &cam, const RenderParams &rp, const GradLayout &gl, const float *d_img, cudaStream_t st) {
        RenderParams rq = rp;
        const int n = gl.off_env;            // everything but the envmap texels
        rq.smem_grad = n <= kSmemGradMaxFloats ? 1 : 0;
        const size_t bytes = rq.smem_grad ? sizeof(float) * n : 0;
        const long long lanes = rp.lane_end - rp.lane_begin;
        if (vjp_big_cta(lanes, 2 * kBlockSV)) secondary_edge_vjp_kernel<kCfg, true><<<vjp_grid(secondary_edge_vjp_kernel<kCfg, true>, bytes, lanes, kBlockSV), kBlockSV, bytes, st>>>(sc, cam, rq, gl, d_img);
        else secondary_edge_vjp_kernel<kCfg, false><<<vjp_grid(secondary_edge_vjp_kernel<kCfg, false>, bytes, lanes), kBlockV, bytes, st>>>(sc, cam, rq, gl, d_img);
        return cudaGetLastError();
    }
};

}  // namespace psdr
