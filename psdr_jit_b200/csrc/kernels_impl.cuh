// sm_100a kernels of the PathTracer hot path (reference src/integrator/integrator.cpp:104-198,
// src/integrator/path.cpp:35-294).  One fused kernel per term: a lane generates its camera ray /
// edge sample from a counter-based PCG32 stream, walks the whole path in registers (closest-hit
// queries against the L1/L2-resident triangle tables) and splats into the image.  Nothing but the
// final image (and derivative image) goes to HBM.
#pragma once
#include <cuda_runtime.h>

#include "device_path.cuh"
#include "kernels.h"

// ---- CTA shape of the three term kernels (tuned on cfg 2, B200; profiles/r02m, r02n) ---------------------------------
// These kernels are 7-10 K SASS instructions (110-160 KB) against a 32 KB L1.5 instruction cache, and their warps walk
// different paths: with 6-8 independent 128-thread CTAs per SM, 39 % of the stall samples of the interior and
// secondary-edge kernels (16 % of the primary-edge kernel) were "no instruction" (profiles/r02a).  ONE large CTA per SM
// with a block barrier where every path (edge side, batch) starts keeps all warps of the SM inside the same stretch of
// code: interior 2.26 -> 1.76 ms, primary edges 6.52 -> 5.92 ms, secondary edges 1.26 -> 1.07 ms.  A barrier at every
// path STEP (mode 2) costs more in idle warps than it saves (+10 %).
// PSDR_LB_* = resident CTAs per SM each kernel is compiled for (register cap = 65536 / (block * n)).
#ifndef PSDR_LB_INTERIOR
#define PSDR_LB_INTERIOR 1
#endif
#ifndef PSDR_LB_INTERIOR_DUAL
#define PSDR_LB_INTERIOR_DUAL 1
#endif
#ifndef PSDR_LB_PRIMARY
#define PSDR_LB_PRIMARY 1
#endif
#ifndef PSDR_LB_SECONDARY
#define PSDR_LB_SECONDARY 1
#endif
// CTA-wide phase barriers: 0 = none; 1 = one block barrier per path (and per edge side / batch); 2 = additionally one
// per path step.
#ifndef PSDR_SYNC_I
#define PSDR_SYNC_I 1
#endif
#ifndef PSDR_SYNC_P
#define PSDR_SYNC_P 1
#endif
#ifndef PSDR_SYNC_S
#define PSDR_SYNC_S 1
#endif
#ifndef PSDR_BLOCK_I
#define PSDR_BLOCK_I 640
#endif
#ifndef PSDR_BLOCK_I_FULL
#define PSDR_BLOCK_I_FULL 896   // interior kernel of the full-feature family (Microfacet / envmap / textures): cfg 3 214.6 -> 209.2 ms (profiles/r03ab)
#endif
#ifndef PSDR_BLOCK_P
#define PSDR_BLOCK_P 896
#endif
#ifndef PSDR_BLOCK_S
#define PSDR_BLOCK_S 1024
#endif

namespace psdr {

constexpr int kBlock = 128;
constexpr int kBlockI = PSDR_BLOCK_I, kBlockP = PSDR_BLOCK_P, kBlockS = PSDR_BLOCK_S;
template <int kCfg> struct InteriorBlock { static constexpr int value = (kCfg & kCfgFull) ? PSDR_BLOCK_I_FULL : PSDR_BLOCK_I; };

// Reduce values over runs of consecutive lanes that share a pixel (lanes are pixel-major, so a
// warp holds at most a few contiguous runs; with spp a multiple of 32 it is one run) and issue one
// atomicAdd per run and channel.  Reference: scatter_reduce(Add) in integrator.cpp:128.
__device__ __forceinline__ void splat_runs(float *img, int pix, float r, float g, float b, bool valid, int mc) {
    const unsigned lane = threadIdx.x & 31u;
    const int key = valid ? pix : -1;
    if (!valid) { r = g = b = 0.f; }
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int k2 = __shfl_down_sync(0xffffffffu, key, off);
        const float r2 = __shfl_down_sync(0xffffffffu, r, off), g2 = __shfl_down_sync(0xffffffffu, g, off),
                    b2 = __shfl_down_sync(0xffffffffu, b, off);
        if (lane + off < 32 && k2 == key) { r += r2; g += g2; b += b2; }
    }
    const int kprev = __shfl_up_sync(0xffffffffu, key, 1);
    if (valid && (lane == 0 || kprev != key)) {
        if (mc == 2) out_add_rgb(img, pix, r, g, b, mc);
        else {      // (zero channels are added too: the reference's scatter_reduce writes every lane)
            out_add(img + 3 * pix, r, mc);
            out_add(img + 3 * pix + 1, g, mc);
            out_add(img + 3 * pix + 2, b, mc);
        }
    }
}

__device__ __forceinline__ float scrub(float x) { return isfinite(x) ? x : 0.f; }

// ---- interior term: Integrator::__render / __render_batch ------------------------------------
// kAD = use the formulas of the reference's renderD instantiation (the primary hit re-intersected
// analytically, scene.cpp:772-801) -- <float, kBvh, true> is the primal image of renderD.
// kBig: the one-CTA-per-SM shape with block barriers (above); else the 128-thread shape for launches too small to fill it.
// kColloc: Li = CollocatedIntegrator (device_path.cuh Li_collocated) instead of the path tracer; 128-thread shape only.
template <class S, int kCfg, bool kAD, bool kBig, bool kColloc = false>
__global__ void __launch_bounds__(kBig ? InteriorBlock<kCfg>::value : kBlock, kBig ? (IsDual<S>::value ? PSDR_LB_INTERIOR_DUAL : PSDR_LB_INTERIOR) : 6)
    interior_kernel(const __grid_constant__ DScene sc, const __grid_constant__ DCamera cam, const __grid_constant__ RenderParams rp, float *__restrict__ img,
                    float *__restrict__ dimg) {
    constexpr int kBlockI = kBig ? InteriorBlock<kCfg>::value : kBlock, kSync = kBig ? PSDR_SYNC_I : 0;
    brute_init<kCfg>(sc, kBlockI);
    const long long stride = (long long) gridDim.x * kBlockI;
    const long long span = rp.lane_end - rp.lane_begin;
    const long long span_pad = (span + kBlockI - 1) / kBlockI * kBlockI;   // keep warps converged for the shuffles (and the trip count CTA-uniform)
    const float inv_spp = sc.spp > 1 ? 1.f / (float) sc.spp : 1.f;
    // large CTAs: chunks handed out dynamically (device_path.cuh ChunkSched; its barriers are the per-path block barrier);
    // 128-thread CTAs: plain grid-stride loop
    __shared__ long long s_chunk;
    const bool dynamic = kSync != 0 && rp.sched != nullptr;
    ChunkSched sched{rp.sched};
    for (long long k = 0;; ++k) {
        long long j;
        if (dynamic) {
            j = sched.next(&s_chunk) * kBlockI + threadIdx.x;
            if (j >= span_pad) break;
        } else {
            j = (long long) blockIdx.x * kBlockI + threadIdx.x + k * stride;
            if (j >= span_pad) break;
            if (kSync) __syncthreads();
        }
        const long long gi = global_lane(rp, j);
        const bool live = j < span && gi < rp.n_lanes;
        const long long i = live ? gi : 0;
        int idx = 0;
        V3<S> v(S(0.f));
        if (kSync >= 2 || live) {
            idx = (int) (sc.spp > 1 ? i / sc.spp : i);
            const int pix = rp.pix_id ? __ldg(rp.pix_id + idx) : idx;
            const unsigned long long seed_value = rp.pix_id ? (unsigned long long) ((long long) pix + rp.seed) : (unsigned long long) (i + rp.seed);
            Pcg32 rng;
            rng.seed(seed_value, (unsigned long long) i);
            if (rp.skip) rng.advance(rp.skip);
            const float jy = rng.next_1d(), jx = rng.next_1d();
            const float sx = ((float) (pix % sc.width) + jx) / (float) sc.width;
            const float sy = ((float) (pix / sc.width) + jy) / (float) sc.height;
            V3<S> o, d;
            sample_primary_ray<S>(cam, V2f(sx, sy), o, d);
            NoRecord rec;
            if (kColloc) v = Li_collocated<S, kCfg, kAD>(sc, o, d, live);
            else v = Li<S, kCfg, kAD, NoRecord>(sc, rng, o, d, live, rp.max_depth, rp.hide_emitters != 0, rec, kSync >= 2 ? 0xffffffffu : 0u, rp.mis, kSync >= 2);
        }
        float r = val(v.x), g = val(v.y), b = val(v.z), dr = tang(v.x), dg = tang(v.y), db = tang(v.z);
        // masked(value, ~isfinite(value)) = 0 zeroes value and tangent (integrator.cpp:126)
        if (!isfinite(r)) { r = 0.f; dr = 0.f; }
        if (!isfinite(g)) { g = 0.f; dg = 0.f; }
        if (!isfinite(b)) { b = 0.f; db = 0.f; }
        splat_runs(img, idx, r * inv_spp, g * inv_spp, b * inv_spp, live, rp.out_multicast);
        if (IsDual<S>::value) {
            const float ts = rp.tangent_scale * inv_spp;
            splat_runs(dimg, idx, scrub(dr) * ts, scrub(dg) * ts, scrub(db) * ts, live, rp.out_multicast);
        }
    }
    if (dynamic) sched.finish();
}

// ---- primary (pixel) edges: PerspectiveCamera::sample_primary_edge + Integrator::render_primary_edges
template <int kCfg, bool kBig, bool kColloc = false>
__global__ void __launch_bounds__(kBig ? kBlockP : kBlock, kBig ? PSDR_LB_PRIMARY : 8) primary_edge_kernel(const __grid_constant__ DScene sc, const __grid_constant__ DCamera cam,
                                                               const __grid_constant__ RenderParams rp, float *__restrict__ dimg) {
    constexpr int kBlockP = kBig ? psdr::kBlockP : kBlock, kSync = kBig ? PSDR_SYNC_P : 0;
    brute_init<kCfg>(sc, kBlockP);
    const long long stride = (long long) gridDim.x * kBlockP;
    const float inv_sppe = sc.sppe > 1 ? 1.f / (float) sc.sppe : 1.f;
    // every lane of a warp runs the same number of iterations (the span is padded to 32) and the warp re-converges
    // at the top of each one: without the barrier lanes that finish a path early run ahead into the next lane's
    // closest-hit scans and the warp stays split (profiles/r01d: 5 of 32 lanes active in the adjoint kernel)
    const long long span = rp.lane_end - rp.lane_begin, span_pad = (span + kBlockP - 1) / kBlockP * kBlockP;
    // large CTAs: chunks (one CTA-load of lanes) handed out dynamically, as in the interior kernel (device_path.cuh ChunkSched;
    // its barriers are the per-path block barrier).  With the lanes ordered along the edge list (rp.perm) a chunk lies inside
    // one bucket and chunks differ in cost -- off-screen stretches are free, a stretch in front of the light is not -- so a
    // static slice per CTA ends with the CTAs that drew the expensive stretches.
    __shared__ long long s_chunk;
    const bool dynamic = kSync != 0 && rp.sched != nullptr;
    ChunkSched sched{rp.sched};
    for (long long k = 0;; ++k) {
        long long j;
        if (dynamic) {
            j = sched.next(&s_chunk) * kBlockP + threadIdx.x;
            if (j >= span_pad) break;
        } else {
            j = (long long) blockIdx.x * kBlockP + threadIdx.x + k * stride;
            if (j >= span_pad) break;
            if (kSync) __syncthreads();
            else __syncwarp();
        }
        const long long i = global_lane(rp, rp.perm && j < span ? (long long) __ldg(rp.perm + j) : j);
        const bool live = j < span && i < rp.n_lanes;
        const unsigned live_mask = __ballot_sync(0xffffffffu, live);
        if (!kSync && !live) continue;        // (with block barriers dead lanes ride along inactive)
        Pcg32 rng;
        rng.seed((unsigned long long) (i + rp.seed), (unsigned long long) i);
        if (rp.skip) rng.advance(rp.skip);
        float s1 = rng.next_1d(), prob;
        const int ei = sample_reuse(cam.pe_pmf, cam.pe_cmf, cam.n_edges, cam.edge_sum, s1, prob);
        const float4 a = __ldg(cam.pe_a + ei), da = __ldg(cam.pe_da + ei), bq = __ldg(cam.pe_b + ei);
        const float pdf = prob / bq.z;
        const float w0 = 1.0f - s1;
        const Dual px = fmadd(Dual(a.x, da.x), Dual(w0), Dual(a.z, da.z) * s1), py = fmadd(Dual(a.y, da.y), Dual(w0), Dual(a.w, da.w) * s1);
        const Dual x_dot_n = dot(V2d(px, py), V2d(Dual(bq.x), Dual(bq.y)));
        const int ix = (int) floorf(px.v * (float) sc.width), iy = (int) floorf(py.v * (float) sc.height);
        const bool valid = live && ix >= 0 && ix < sc.width && iy >= 0 && iy < sc.height;
        // Li(ray_n) - Li(ray_p): the reference binary evaluates Li(ray_p) first (verified on the
        // running reference, tests/golden/renderD_*: primary-only images).  One rolled loop over the two
        // sides keeps a single copy of Li in the kernel.
        V3f Lside[2];
#pragma unroll 1
        for (int side = 0; side < 2; ++side) {
            if (kSync) __syncthreads();
            else __syncwarp(live_mask);
            const float sg = side == 0 ? kEdgeEpsilon : -kEdgeEpsilon;
            V3f ro, rd;
            sample_primary_ray<float>(cam, V2f(px.v + sg * bq.x, py.v + sg * bq.y), ro, rd);
            if (kColloc) Lside[side] = Li_collocated<float, kCfg, false>(sc, ro, rd, valid);
            else Lside[side] = Li<float, kCfg>(sc, rng, ro, rd, valid, rp.max_depth, rp.hide_emitters != 0, rp.mis, kSync >= 2);
        }
        const V3f Lp = Lside[0], Ln = Lside[1];
        if (!valid) continue;
        const int pix = iy * sc.width + ix;
        const float inv_pdf = 1.f / pdf;
        const float dl[3] = {(Ln.x - Lp.x) * inv_pdf, (Ln.y - Lp.y) * inv_pdf, (Ln.z - Lp.z) * inv_pdf};
        float t3[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float primal = x_dot_n.v * dl[c];
            const float t = x_dot_n.d * dl[c] * inv_sppe;
            t3[c] = (isfinite(primal) && isfinite(t)) ? t : 0.f;
        }
        if (t3[0] != 0.f || t3[1] != 0.f || t3[2] != 0.f) out_add_rgb(dimg, pix, t3[0], t3[1], t3[2], rp.out_multicast);
    }
    if (dynamic) sched.finish();
}

// ---- secondary (shadow) edges: PathTracer::render_secondary_edges -----------------------------
// Sample i of the term is a pure function of (seed, i), so it does not matter which lane evaluates it.  Each warp
// owns a contiguous slice of the samples and works in batches: lanes without a candidate pull the next untested
// samples of the slice through stage 0 (cheap, ~30 % pass) until the warp is (nearly) full, then all candidates run
// stage 1 -- the three closest-hit scans -- together.  One sample per lane per iteration ran those scans with
// 12, 10 and 2.5 of 32 lanes (profiles/r01i).
template <int kCfg, bool kBig>
__global__ void __launch_bounds__(kBig ? kBlockS : kBlock, kBig ? PSDR_LB_SECONDARY : 8) secondary_edge_kernel(const __grid_constant__ DScene sc, const __grid_constant__ DCamera cam,
                                                                 const __grid_constant__ RenderParams rp, float *__restrict__ dimg) {
    constexpr int kBlockS = kBig ? psdr::kBlockS : kBlock, kSync = kBig ? PSDR_SYNC_S : 0;
    brute_init<kCfg>(sc, kBlockS);
    const float scale = rp.tangent_scale * (sc.sppse > 1 ? 1.f / (float) sc.sppse : 1.f);
    sec_edge_batches<kCfg>(sc, cam, rp, kBlockS, [&](const SecSample &smp) {
        V3f value0, tangent;
        const int pix = sec_edge_stage1<kCfg, NoSecAdjoint>(sc, cam, smp.cand, value0, tangent, NoSecAdjoint());
        if (pix < 0) return;
        float t[3] = {tangent.x, tangent.y, tangent.z};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (smp.pdf0 > kEpsilon) t[c] = t[c] / smp.pdf0;                 // masked(value, pdf0 > Epsilon) /= pdf0
            t[c] = isfinite(t[c]) ? t[c] * scale : 0.f;
        }
        if (t[0] != 0.f || t[1] != 0.f || t[2] != 0.f) out_add_rgb(dimg, pix, t[0], t[1], t[2], rp.out_multicast);
    }, kSync != 0);
}

// ---- guiding pre-pass: PathTracer::preprocess_secondary_edges (reference src/integrator/path.cpp:130-168)
// One thread per grid cell walks that cell's reso[3] x nrounds samples in lane order (a fixed summation
// order; the reference's scatter_reduce order is unspecified) and writes mass[cell].
template <int kCfg>
__global__ void __launch_bounds__(kBlock) guiding_kernel(const __grid_constant__ DScene sc, const __grid_constant__ DCamera cam, int r0, int r1, int r2,
                                                          int r3, int nrounds, long long seed, float *__restrict__ mass) {
    brute_init<kCfg>(sc, kBlock);
    const int ncells = r0 * r1 * r2;
    for (int cell = blockIdx.x * kBlock + threadIdx.x; cell < ncells; cell += gridDim.x * kBlock) {
        const int c0 = cell / (r1 * r2), rem = cell - c0 * (r1 * r2), c1 = rem / r2, c2 = rem - c1 * r2;
        float total = 0.f;
        for (int j = 0; j < nrounds; ++j)
            for (int k = 0; k < r3; ++k) {
                const long long i = (long long) cell * r3 + k;
                Pcg32 rng;
                rng.seed((unsigned long long) (i + seed), (unsigned long long) i);
                if (j) rng.advance(3ull * (unsigned long long) j);
                const float d1 = rng.next_1d(), d2 = rng.next_1d(), d3 = rng.next_1d();
                const V3f s3(((float) c0 + d3) * (1.f / (float) r0), ((float) c1 + d2) * (1.f / (float) r1), ((float) c2 + d1) * (1.f / (float) r2));
                V3f value0, tangent;
                eval_secondary_edge<kCfg>(sc, cam, s3, value0, tangent);
                const float v[3] = {value0.x, value0.y, value0.z};
                float m = 0.f;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float x = isfinite(v[c]) ? v[c] : 0.f;
                    if (r3 > 1) x /= (float) r3;
                    m = c == 0 ? x : fmaxf(m, x);
                }
                total += m;
            }
        mass[cell] = total;
    }
}

// ---- AOV tap: what the reference's FieldExtractionIntegrator exposes (src/integrator/field.cpp:47-121)
template <int kCfg>
__global__ void __launch_bounds__(kBlock) aov_kernel(const __grid_constant__ DScene sc, const __grid_constant__ DCamera cam,
                                                      const __grid_constant__ RenderParams rp, float *__restrict__ out) {
    brute_init<kCfg>(sc, kBlock);
    const long long stride = (long long) gridDim.x * kBlock;
    // every lane of a warp runs the same number of iterations (the span is padded to 32) and the warp re-converges
    // at the top of each one: without the barrier lanes that finish a path early run ahead into the next lane's
    // closest-hit scans and the warp stays split (profiles/r01d: 5 of 32 lanes active in the adjoint kernel)
    const long long span = rp.lane_end - rp.lane_begin, span_pad = (span + 31) / 32 * 32;
    for (long long j = (long long) blockIdx.x * kBlock + threadIdx.x; j < span_pad; j += stride) {
        __syncwarp();
        const long long i = global_lane(rp, j);
        if (!(j < span && i < rp.n_lanes)) continue;
        const int idx = (int) (sc.spp > 1 ? i / sc.spp : i);
        Pcg32 rng;
        rng.seed((unsigned long long) (i + rp.seed), (unsigned long long) i);
        const float jy = rng.next_1d(), jx = rng.next_1d();
        const float sx = ((float) (idx % sc.width) + jx) / (float) sc.width, sy = ((float) (idx / sc.width) + jy) / (float) sc.height;
        V3f o, d;
        sample_primary_ray<float>(cam, V2f(sx, sy), o, d);
        const Its<float> its = ray_intersect<float, kCfg>(sc, o, d, true, false);
        float *r = out + 14 * i;
        for (int k = 0; k < 14; ++k) r[k] = 0.f;
        if (!its.valid) { r[1] = -1.f; continue; }
        r[0] = (float) (its.mesh + 1); r[1] = (float) its.tri;
        r[2] = its.p.x; r[3] = its.p.y; r[4] = its.p.z; r[5] = its.t;
        r[6] = its.n.x; r[7] = its.n.y; r[8] = its.n.z;
        r[9] = its.sh_n.x; r[10] = its.sh_n.y; r[11] = its.sh_n.z;
        r[12] = its.uv.x; r[13] = its.uv.y;
    }
}

// ---- FieldExtractionIntegrator::renderD, forward mode (reference src/integrator/field.cpp:47-121 through
// Integrator::renderD, src/integrator/integrator.cpp:51-100,179-198) ----------------------------------------------
// Interior part: the taps of aov_kernel from the AD instantiation of ray_intersect (primary hit re-intersected
// analytically, scene.cpp:772-801) with their forward tangents: 14 values + 14 tangents per lane.
template <int kCfg>
__global__ void __launch_bounds__(kBlock) aov_d_kernel(const __grid_constant__ DScene sc, const __grid_constant__ DCamera cam,
                                                        const __grid_constant__ RenderParams rp, float *__restrict__ out, float *__restrict__ dout) {
    brute_init<kCfg>(sc, kBlock);
    const long long stride = (long long) gridDim.x * kBlock;
    const long long span = rp.lane_end - rp.lane_begin, span_pad = (span + 31) / 32 * 32;
    for (long long j = (long long) blockIdx.x * kBlock + threadIdx.x; j < span_pad; j += stride) {
        __syncwarp();
        const long long i = global_lane(rp, j);
        if (!(j < span && i < rp.n_lanes)) continue;
        const int idx = (int) (sc.spp > 1 ? i / sc.spp : i);
        Pcg32 rng;
        rng.seed((unsigned long long) (i + rp.seed), (unsigned long long) i);
        const float jy = rng.next_1d(), jx = rng.next_1d();
        const float sx = ((float) (idx % sc.width) + jx) / (float) sc.width, sy = ((float) (idx / sc.width) + jy) / (float) sc.height;
        V3d o, d;
        sample_primary_ray<Dual>(cam, V2f(sx, sy), o, d);
        const Its<Dual> its = ray_intersect<Dual, kCfg, true>(sc, o, d, true, false);
        float *r = out + 14 * i, *t = dout + 14 * i;
        for (int k = 0; k < 14; ++k) { r[k] = 0.f; t[k] = 0.f; }
        if (!its.valid) { r[1] = -1.f; continue; }
        r[0] = (float) (its.mesh + 1); r[1] = (float) its.tri;
        const Dual f[12] = {its.p.x, its.p.y, its.p.z, its.t, its.n.x, its.n.y, its.n.z, its.sh_n.x, its.sh_n.y, its.sh_n.z, its.uv.x, its.uv.y};
        for (int k = 0; k < 12; ++k) { r[2 + k] = f[k].v; t[2 + k] = isfinite(f[k].d) ? f[k].d : 0.f; }
    }
}

// Primary-edge part (Integrator::render_primary_edges with Li = the field at the primary hit): the jump of the field across
// the sampled pixel-space edge times the edge's normal velocity.
template <int kCfg>
__global__ void __launch_bounds__(kBlock) field_edge_kernel(const __grid_constant__ DScene sc, const __grid_constant__ DCamera cam,
                                                             const __grid_constant__ RenderParams rp, int field, int object, float *__restrict__ dimg) {
    brute_init<kCfg>(sc, kBlock);
    const long long stride = (long long) gridDim.x * kBlock;
    const float inv_sppe = sc.sppe > 1 ? 1.f / (float) sc.sppe : 1.f;
    const long long span = rp.lane_end - rp.lane_begin, span_pad = (span + 31) / 32 * 32;
    for (long long j = (long long) blockIdx.x * kBlock + threadIdx.x; j < span_pad; j += stride) {
        __syncwarp();
        const long long i = global_lane(rp, j);
        if (!(j < span && i < rp.n_lanes)) continue;
        Pcg32 rng;
        rng.seed((unsigned long long) (i + rp.seed), (unsigned long long) i);
        float s1 = rng.next_1d(), prob;
        const int ei = sample_reuse(cam.pe_pmf, cam.pe_cmf, cam.n_edges, cam.edge_sum, s1, prob);
        const float4 a = __ldg(cam.pe_a + ei), da = __ldg(cam.pe_da + ei), bq = __ldg(cam.pe_b + ei);
        const float pdf = prob / bq.z;
        const float w0 = 1.0f - s1;
        const Dual px = fmadd(Dual(a.x, da.x), Dual(w0), Dual(a.z, da.z) * s1), py = fmadd(Dual(a.y, da.y), Dual(w0), Dual(a.w, da.w) * s1);
        const Dual x_dot_n = dot(V2d(px, py), V2d(Dual(bq.x), Dual(bq.y)));
        const int ix = (int) floorf(px.v * (float) sc.width), iy = (int) floorf(py.v * (float) sc.height);
        if (!(ix >= 0 && ix < sc.width && iy >= 0 && iy < sc.height)) continue;
        V3f side[2];
#pragma unroll 1
        for (int k = 0; k < 2; ++k) {
            const float sg = k == 0 ? kEdgeEpsilon : -kEdgeEpsilon;
            V3f ro, rd;
            sample_primary_ray<float>(cam, V2f(px.v + sg * bq.x, py.v + sg * bq.y), ro, rd);
            side[k] = field_value(ray_intersect<float, kCfg>(sc, ro, rd, true, false), field, object);
        }
        const float inv_pdf = 1.f / pdf;
        const float dl[3] = {(side[1].x - side[0].x) * inv_pdf, (side[1].y - side[0].y) * inv_pdf, (side[1].z - side[0].z) * inv_pdf};
        float t3[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float primal = x_dot_n.v * dl[c], t = x_dot_n.d * dl[c] * inv_sppe;
            t3[c] = (isfinite(primal) && isfinite(t)) ? t : 0.f;
        }
        if (t3[0] != 0.f || t3[1] != 0.f || t3[2] != 0.f) out_add_rgb(dimg, iy * sc.width + ix, t3[0], t3[1], t3[2], 0);
    }
}

// ---- per-configuration launchers: one explicit instantiation of ForwardLaunch<kCfg> per translation unit
// (kern_cfg*.cu) so that the four kernel families compile in parallel ----------------------------------------
// Persistent grid = exactly the number of CTAs that are resident at once (occupancy API x SM count): a larger
// grid runs in more than one wave and the last, partially filled wave costs as much as a full one.
template <class K> inline int persistent_grid(K kernel, int block, size_t smem, long long lanes) {
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    const long long need = (lanes + block - 1) / block, cap = (long long) sms * per_sm;
    return (int) (need < cap ? (need > 0 ? need : 1) : cap);
}

// Launches with fewer lanes than ~4 per thread of the one-CTA-per-SM grid run the 128-thread shape (more CTAs, all SMs busy).
// g_cta_policy (psdr_set_cta_policy): 0 = by size, 1 = always the 128-thread shape, 2 = always the large-CTA shape.
extern int g_cta_policy;
inline bool use_big_cta(long long lanes, int block) {
    if (g_cta_policy == 1) return false;
    if (g_cta_policy == 2) return true;
    return lanes >= 4LL * 148 * block;
}

template <int kCfg> struct ForwardLaunch {
    template <class S, bool kAD> static void interior_as(const DScene &sc, const DCamera &cam, const RenderParams &rp, float *img, float *dimg, cudaStream_t st) {
        const long long n = rp.lane_end - rp.lane_begin;
        constexpr int kBlockI = InteriorBlock<kCfg>::value;
        if (rp.mis == 3) {      // CollocatedIntegrator
            interior_kernel<S, kCfg, kAD, false, true><<<persistent_grid(interior_kernel<S, kCfg, kAD, false, true>, kBlock, 0, n), kBlock, 0, st>>>(sc, cam, rp, img, dimg);
            return;
        }
        if (use_big_cta(n, kBlockI)) interior_kernel<S, kCfg, kAD, true><<<persistent_grid(interior_kernel<S, kCfg, kAD, true>, kBlockI, 0, n), kBlockI, 0, st>>>(sc, cam, rp, img, dimg);
        else interior_kernel<S, kCfg, kAD, false><<<persistent_grid(interior_kernel<S, kCfg, kAD, false>, kBlock, 0, n), kBlock, 0, st>>>(sc, cam, rp, img, dimg);
    }
    static cudaError_t interior(const DScene &sc, const DCamera &cam, const RenderParams &rp, bool ad, float *img, float *dimg, cudaStream_t st) {
        if (ad && dimg) interior_as<Dual, true>(sc, cam, rp, img, dimg, st);
        else if (ad) interior_as<float, true>(sc, cam, rp, img, dimg, st);     // primal image of renderD
        else interior_as<float, false>(sc, cam, rp, img, dimg, st);
        return cudaGetLastError();
    }
    static cudaError_t primary(const DScene &sc, const DCamera &cam, const RenderParams &rp, float *dimg, cudaStream_t st) {
        const long long n = rp.lane_end - rp.lane_begin;
        if (rp.mis == 3) {      // CollocatedIntegrator
            primary_edge_kernel<kCfg, false, true><<<persistent_grid(primary_edge_kernel<kCfg, false, true>, kBlock, 0, n), kBlock, 0, st>>>(sc, cam, rp, dimg);
            return cudaGetLastError();
        }
        if (use_big_cta(n, kBlockP)) primary_edge_kernel<kCfg, true><<<persistent_grid(primary_edge_kernel<kCfg, true>, kBlockP, 0, n), kBlockP, 0, st>>>(sc, cam, rp, dimg);
        else primary_edge_kernel<kCfg, false><<<persistent_grid(primary_edge_kernel<kCfg, false>, kBlock, 0, n), kBlock, 0, st>>>(sc, cam, rp, dimg);
        return cudaGetLastError();
    }
    static cudaError_t secondary(const DScene &sc, const DCamera &cam, const RenderParams &rp, float *dimg, cudaStream_t st) {
        const long long n = rp.lane_end - rp.lane_begin;
        if (use_big_cta(n, kBlockS)) secondary_edge_kernel<kCfg, true><<<persistent_grid(secondary_edge_kernel<kCfg, true>, kBlockS, 0, n), kBlockS, 0, st>>>(sc, cam, rp, dimg);
        else secondary_edge_kernel<kCfg, false><<<persistent_grid(secondary_edge_kernel<kCfg, false>, kBlock, 0, n), kBlock, 0, st>>>(sc, cam, rp, dimg);
        return cudaGetLastError();
    }
    static cudaError_t guiding(const DScene &sc, const DCamera &cam, const int reso[4], int nrounds, long long seed, float *mass, cudaStream_t st) {
        const long long cells = (long long) reso[0] * reso[1] * reso[2];
        guiding_kernel<kCfg><<<persistent_grid(guiding_kernel<kCfg>, kBlock, 0, cells), kBlock, 0, st>>>(sc, cam, reso[0], reso[1], reso[2], reso[3], nrounds, seed, mass);
        return cudaGetLastError();
    }
    static cudaError_t aov(const DScene &sc, const DCamera &cam, const RenderParams &rp, float *out, cudaStream_t st) {
        aov_kernel<kCfg><<<persistent_grid(aov_kernel<kCfg>, kBlock, 0, rp.lane_end - rp.lane_begin), kBlock, 0, st>>>(sc, cam, rp, out);
        return cudaGetLastError();
    }
    static cudaError_t aov_d(const DScene &sc, const DCamera &cam, const RenderParams &rp, float *out, float *dout, cudaStream_t st) {
        aov_d_kernel<kCfg><<<persistent_grid(aov_d_kernel<kCfg>, kBlock, 0, rp.lane_end - rp.lane_begin), kBlock, 0, st>>>(sc, cam, rp, out, dout);
        return cudaGetLastError();
    }
    static cudaError_t field_edges(const DScene &sc, const DCamera &cam, const RenderParams &rp, int field, int object, float *dimg, cudaStream_t st) {
        field_edge_kernel<kCfg><<<persistent_grid(field_edge_kernel<kCfg>, kBlock, 0, rp.lane_end - rp.lane_begin), kBlock, 0, st>>>(sc, cam, rp, field, object, dimg);
        return cudaGetLastError();
    }
};

}  // namespace psdr
