// Device-side path-space integrand: intersection records, Diffuse BSDF, area-light sampling,
// the PathTracer radiance estimator (interior term) and the secondary-edge boundary estimator.
// Everything is templated on the scalar S: float = primal (renderC and the detached Li calls of the
// edge terms), Dual = value + forward tangent (renderD interior term; what the reference gets from
// Dr.Jit's AD with detach() at the same places).  One thread = one lane of the reference wavefront.
#pragma once
#include "dscene.h"
#include "pmath.h"
#include "texture.h"

namespace psdr {

// kernel configuration bits (template parameter kCfg): which code a kernel instantiation contains
constexpr int kCfgBvh = 1;    // BVH2 traversal instead of the parameter-space triangle scan
constexpr int kCfgFull = 2;   // MicrofacetBSDF + EnvironmentMap code paths (otherwise Diffuse + AreaLight only)
constexpr int kCfgUniformScan = 4;   // brute-force scan: warp-uniform trip count in the per-lane candidate loop (see trace())
// Extended material set (implies kCfgFull): bitmap-valued BSDF slots, RoughConductor, RoughDielectric, MicrofacetPerVertex,
// NormalMap.  A family of its own because code size is a first-order term for the full-feature kernels (32 KB instruction
// cache): with the textures and the conductor compiled out, cfg 3 (constant Microfacet + envmap) runs 12 % faster
// (profiles/r03g_cfg3_texcond.log), so scenes that use none of this do not carry it.
constexpr int kCfgExt = 8;

struct Hit {
    int tri;
    float u, v, t;
};

// local lane index of this rank -> global lane index of the term (RenderParams: block-cyclic deal of 32-lane blocks)
// Output accumulation.  mc = 0: red.global into this GPU's buffer.  mc = 1: `p` is an NVLS multicast address of a buffer
// that every rank of the node maps -- ONE multimem.red travels to the NVSwitch, which adds the value into the replica of
// every GPU: the term kernels of the N ranks then build the complete image on all ranks with no reduction pass.
__device__ __forceinline__ void out_add(float *p, float v, int mc) {
    if (mc) asm volatile("multimem.red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
    else atomicAdd(p, v);
}
// One pixel (rgb) of an image.  mc = 0 / 1: float32[npix][3], three adds (zero channels skipped).  mc = 2: the multicast
// image is float32[npix][4] and the pixel travels as ONE 16-byte multimem.red.v4 -- every multimem.red is replicated by
// the switch to all N GPUs, so each GPU receives the reds of ALL ranks and the packet count, not the byte count, is what
// the links carry: three 4-byte reds per splat slowed the secondary-edge kernel by 1.5x at 8 GPUs (profiles/r02v).
__device__ __forceinline__ void out_add_rgb(float *img, long long pix, float r, float g, float b, int mc) {
    if (mc == 2) {
        asm volatile("multimem.red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(img + 4 * pix), "f"(r), "f"(g), "f"(b), "f"(0.f)
                     : "memory");
        return;
    }
    float *p = img + 3 * pix;
    if (r != 0.f) out_add(p, r, mc);
    if (g != 0.f) out_add(p + 1, g, mc);
    if (b != 0.f) out_add(p + 2, b, mc);
}
// Persistent-loop schedule of the interior kernels.  Iteration k covers local lanes [k G B, (k + 1) G B) (G CTAs of B
// threads); CTA c takes chunk (c + k) mod G of it.  With the plain grid-stride loop (chunk c every time) a CTA revisits
// the same image COLUMNS whenever G B / spp (x the number of ranks) shares a large factor with the image width: the
// interior adjoint's 2368 pixels per iteration at 512 columns give 8 column sets, at 8 ranks every iteration of a CTA
// lands on the same 64 columns, and the kernel ends with the CTAs that own the expensive columns -- the sharded
// interior adjoint took 1.5x its share at 4 ranks and 1.65x at 8 (profiles/r02q).  Shifting by one chunk per iteration
// walks every CTA across the row.
__device__ __forceinline__ long long rotated_lane(long long k, int block) {
    const unsigned G = gridDim.x;
    unsigned chunk = blockIdx.x + (unsigned) (k % G);
    if (chunk >= G) chunk -= G;
    return (k * (long long) G + chunk) * block + threadIdx.x;
}
// Dynamic chunk hand-out for the large-CTA interior kernels: a chunk is one CTA-load of local lanes; thread 0 takes the
// next chunk index from a global counter and the block barrier that starts every path broadcasts it.  The static
// schedules end with the CTAs whose pixels were the expensive ones (the image is not uniform, and a CTA is 1/148 of the
// machine): at 131 k pixels per rank the interior kernel took 1.27x its share.  Here the tail is bounded by one chunk.
// The last CTA out re-arms the two counters, so consecutive launches need no memset.
struct ChunkSched {
    int *ctr;
    __device__ __forceinline__ long long next(long long *slot) {      // every thread of the CTA calls it together
        __syncthreads();        // everyone is done with the previous chunk (and has read *slot)
        if (threadIdx.x == 0) *slot = (long long) atomicAdd(ctr, 1);
        __syncthreads();
        return *slot;
    }
    __device__ __forceinline__ void finish() {
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(ctr + 1, 1) == (int) gridDim.x - 1) {     // every CTA has taken its terminating chunk
                ctr[0] = 0;
                ctr[1] = 0;
                __threadfence();
            }
        }
    }
};
__device__ __forceinline__ long long global_lane(const RenderParams &rp, long long j) {
    return rp.shard_world <= 1 ? rp.lane_begin + j : (((j >> 5) * rp.shard_world + rp.shard_rank) << 5) + (j & 31);
}

template <class S> struct TriRec {
    V3<S> p0, e1, e2;
    S area;
    int mesh;
};
template <class S> struct ShadeRec {
    V3<S> n0, n1, n2, fn;
};

__device__ __forceinline__ V3f f3(float4 a) { return V3f(a.x, a.y, a.z); }

template <class S> __device__ __forceinline__ TriRec<S> load_tri(const DScene &sc, int i);
template <> __device__ __forceinline__ TriRec<float> load_tri<float>(const DScene &sc, int i) {
    const float4 a = __ldg(sc.geo + 3 * i), b = __ldg(sc.geo + 3 * i + 1), c = __ldg(sc.geo + 3 * i + 2);
    TriRec<float> t;
    t.p0 = V3f(a.x, a.y, a.z);
    t.e1 = V3f(a.w, b.x, b.y);
    t.e2 = V3f(b.z, b.w, c.x);
    t.area = c.y;
    t.mesh = __float_as_int(c.z);
    return t;
}
template <> __device__ __forceinline__ TriRec<Dual> load_tri<Dual>(const DScene &sc, int i) {
    const float4 a = __ldg(sc.geo + 3 * i), b = __ldg(sc.geo + 3 * i + 1), c = __ldg(sc.geo + 3 * i + 2);
    const float4 da = __ldg(sc.dgeo + 3 * i), db = __ldg(sc.dgeo + 3 * i + 1), dc = __ldg(sc.dgeo + 3 * i + 2);
    TriRec<Dual> t;
    t.p0 = V3d(Dual(a.x, da.x), Dual(a.y, da.y), Dual(a.z, da.z));
    t.e1 = V3d(Dual(a.w, da.w), Dual(b.x, db.x), Dual(b.y, db.y));
    t.e2 = V3d(Dual(b.z, db.z), Dual(b.w, db.w), Dual(c.x, dc.x));
    t.area = Dual(c.y, dc.y);
    t.mesh = __float_as_int(c.z);
    return t;
}
template <class S> __device__ __forceinline__ ShadeRec<S> load_shade(const DScene &sc, int i);
template <> __device__ __forceinline__ ShadeRec<float> load_shade<float>(const DScene &sc, int i) {
    const float4 a = __ldg(sc.shade + 3 * i), b = __ldg(sc.shade + 3 * i + 1), c = __ldg(sc.shade + 3 * i + 2);
    ShadeRec<float> s;
    s.n0 = V3f(a.x, a.y, a.z);
    s.n1 = V3f(a.w, b.x, b.y);
    s.n2 = V3f(b.z, b.w, c.x);
    s.fn = V3f(c.y, c.z, c.w);
    return s;
}
template <> __device__ __forceinline__ ShadeRec<Dual> load_shade<Dual>(const DScene &sc, int i) {
    const float4 a = __ldg(sc.shade + 3 * i), b = __ldg(sc.shade + 3 * i + 1), c = __ldg(sc.shade + 3 * i + 2);
    const float4 da = __ldg(sc.dshade + 3 * i), db = __ldg(sc.dshade + 3 * i + 1), dc = __ldg(sc.dshade + 3 * i + 2);
    ShadeRec<Dual> s;
    s.n0 = V3d(Dual(a.x, da.x), Dual(a.y, da.y), Dual(a.z, da.z));
    s.n1 = V3d(Dual(a.w, da.w), Dual(b.x, db.x), Dual(b.y, db.y));
    s.n2 = V3d(Dual(b.z, db.z), Dual(b.w, db.w), Dual(c.x, dc.x));
    s.fn = V3d(Dual(c.y, dc.y), Dual(c.z, dc.z), Dual(c.w, dc.w));
    return s;
}

// ---- closest hit in (RayEpsilon, 1e8): replaces OptiX (reference src/scene/scene_optix.cpp:343-410)
// Moeller-Trumbore numerators with a fixed operation order (cross = fused multiply-subtract, dot = fma chain).
// Everything a candidate has to pass -- inside test on the sign-normalised numerators, t in (eps, tmax), closer
// than the best so far -- is decided on the NUMERATORS by cross-multiplication (t1 < t2  <=>  tn1 |det2| < tn2 |det1|),
// so the scan is branch-free: the best candidate is carried as (tn, det, un, vn, id) with predicated moves and
// the one IEEE division per ray happens after the scan.  With a division per pierced triangle inside the loop the
// 1-3 lanes of a warp that pierce the current triangle ran ~20 instructions alone while 30 lanes waited
// (profiles/r01e: 17 of 32 lanes active in kernels whose paths were all alive).  Ties resolve to the lowest id.
struct HitCand {
    float ts, adet;        // sign-normalised t numerator and |det| of the best candidate (t = ts / adet)
    int tri;
};
__device__ __forceinline__ void hit_init(HitCand &b) {
    b.ts = kTraceTMax;     // "closer than the best" therefore also means t < kTraceTMax
    b.adet = 1.f;
    b.tri = -1;
}
// the four numerators of one triangle (scalar form: BVH leaves, and the winner of a scan)
struct TriNum {
    float tn, det, un, vn;
};
__device__ __forceinline__ TriNum tri_numerators(V3f p0, V3f e1, V3f e2, V3f o, V3f d) {
    TriNum r;
    const V3f h = cross_fms(d, e2);
    r.det = dot(e1, h);
    const V3f s = o - p0;
    r.un = dot(s, h);
    const V3f q = cross_fms(s, e1);
    r.vn = dot(d, q);
    r.tn = dot(e2, q);
    return r;
}
// acceptance of one candidate on its numerators
template <bool kOrdered>   // kOrdered: candidates arrive in ascending id (strict "closer" keeps the lowest id on ties)
__device__ __forceinline__ void hit_consider(float tn, float det, float un, float vn, int id, HitCand &b) {
    const float adet = fabsf(det);
    const bool neg = det < 0.f;
    const float us = neg ? -un : un, vs = neg ? -vn : vn, ts = neg ? -tn : tn;
    const float lhs = ts * b.adet, rhs = b.ts * adet;
    const bool closer = kOrdered ? (lhs < rhs) : (lhs < rhs || (lhs == rhs && id < b.tri));
    const bool ok = us >= 0.f && vs >= 0.f && us + vs <= adet && adet > 0.f && ts > kRayEpsilon * adet && closer;
    b.ts = ok ? ts : b.ts;
    b.adet = ok ? adet : b.adet;
    b.tri = ok ? id : b.tri;
}
// the same acceptance for the ordered brute-force scan, on operands the packed code has already normalised
__device__ __forceinline__ void hit_consider_ordered(float us, float vs, float sum, float ts, float adet, float eps_adet, int id, HitCand &b) {
    const float lhs = ts * b.adet, rhs = b.ts * adet;
    const bool ok = us >= 0.f && vs >= 0.f && sum <= adet && ts > eps_adet && lhs < rhs;
    b.ts = ok ? ts : b.ts;
    b.adet = ok ? adet : b.adet;
    b.tri = ok ? id : b.tri;
}
template <bool kOrdered>
__device__ __forceinline__ void tri_test(V3f p0, V3f e1, V3f e2, int id, V3f o, V3f d, HitCand &b) {
    const TriNum n = tri_numerators(p0, e1, e2, o, d);
    hit_consider<kOrdered>(n.tn, n.det, n.un, n.vn, id, b);
}
// the winner's numerators are recomputed (same operations, same bits) and divided once
__device__ __forceinline__ Hit hit_finish(const HitCand &b, V3f p0, V3f e1, V3f e2, V3f o, V3f d) {
    Hit h;
    h.tri = b.tri;
    h.u = h.v = 0.f;
    h.t = kTraceTMax;
    if (b.tri >= 0) {
        const TriNum n = tri_numerators(p0, e1, e2, o, d);
        const float f = 1.f / n.det;
        h.t = f * n.tn;
        h.u = f * n.un;
        h.v = f * n.vn;
    }
    return h;
}

// ---- packed fp32x2 arithmetic (sm_100 FFMA2/FMUL2/FADD2): one instruction, two IEEE round-to-nearest results --
// the brute-force scan tests triangles (2j, 2j+1) in the two halves, so its 27 multiply/add/fma per triangle
// cost 27 issue slots per PAIR; the results are bit-identical to the scalar forms above.
struct F2 {
    unsigned long long v;
};
__device__ __forceinline__ F2 f2_dup(float x) {
    F2 r;
    asm("mov.b64 %0, {%1, %1};" : "=l"(r.v) : "f"(x));
    return r;
}
__device__ __forceinline__ F2 f2_pack(float lo, float hi) {
    F2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_split(F2 a, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ F2 f2_mul(F2 a, F2 b) {
    F2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ F2 f2_add(F2 a, F2 b) {
    F2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ F2 f2_sub(F2 a, F2 b) {
    F2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ F2 f2_fma(F2 a, F2 b, F2 c) {
    F2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
struct V3p {
    F2 x, y, z;
};
__device__ __forceinline__ F2 dot(V3p a, V3p b) { return f2_fma(a.z, b.z, f2_fma(a.y, b.y, f2_mul(a.x, b.x))); }
// cross_fms(a, b) with the subtrahend's sign carried by an operand: fma(a.y, b.z, -(a.z * b.y)) = fma(a.y, b.z, na.z * b.y)
__device__ __forceinline__ V3p cross_fms_na(V3p a, V3p na, V3p b) {
    V3p r;
    r.x = f2_fma(a.y, b.z, f2_mul(na.z, b.y));
    r.y = f2_fma(a.z, b.x, f2_mul(na.x, b.z));
    r.z = f2_fma(a.x, b.y, f2_mul(na.y, b.x));
    return r;
}
__device__ __forceinline__ V3p cross_fms_nb(V3p a, V3p b, V3p nb) {
    V3p r;
    r.x = f2_fma(a.y, b.z, f2_mul(a.z, nb.y));
    r.y = f2_fma(a.z, b.x, f2_mul(a.x, nb.z));
    r.z = f2_fma(a.x, b.y, f2_mul(a.y, nb.x));
    return r;
}

// ---- brute-force mode: per-CTA shared-memory copy of the triangle pairs ------------------------------------------
// The per-lane stage of the scan reads pair j with a per-lane j, which the constant bank serialises; shared memory
// serves it as five 128-bit loads (pair stride 80 B).  One static table per kernel; brute_init() fills it.
#ifndef PSDR_BRUTE_CULL
#define PSDR_BRUTE_CULL 1     // 0: every pair is tested for every ray (A/B switch for the box cull)
#endif
__device__ __forceinline__ ulonglong2 *brute_table() {
    __shared__ ulonglong2 s_pairs[kMaxBruteTris / 2 * kBruteSmemStride];
    return s_pairs;
}
// called once by every thread of the CTA at kernel entry (all kernels that trace in brute-force mode)
template <int kCfg> __device__ __forceinline__ void brute_init(const DScene &sc, int block) {
    if (kCfg & kCfgBvh) return;
    unsigned long long *t = reinterpret_cast<unsigned long long *>(brute_table());
    const int n_pairs = (sc.n_tris + 1) >> 1;
    for (int k = threadIdx.x; k < n_pairs * kBrutePairWords; k += block) {
        const int j = k / kBrutePairWords, c = k - j * kBrutePairWords;
        t[2 * kBruteSmemStride * j + c] = sc.bg_pair[k];
    }
    __syncthreads();
}

// Closest hit, brute-force mode, in two stages.
// Stage 1 (warp-uniform, operands from the constant bank): slab test of the ray against the padded box of every
//   triangle PAIR, two boxes per packed instruction -> a per-lane bit mask of the pairs the ray can hit.  In the
//   Cornell box a ray passes 1.7 - 2.8 of the 18 boxes on average (the warp's slowest lane ~4).
// Stage 2 (per lane): each lane walks ITS mask in ascending pair order and runs the packed Moeller-Trumbore test on
//   the two triangles of the pair; the loop runs as long as the lane with the most candidates.  Ascending order and
//   the strict "closer" keep the lowest triangle id on ties, as a full ascending scan would.
// The box test is part of the definition of the closest hit (the CPU checker used by the tests applies the identical
// test), so hit ids agree bit for bit whether or not a pad was generous enough.
#ifndef PSDR_BVH_LOOP
#define PSDR_BVH_LOOP 0     // 1: one unit of work per traversal iteration + distance-tagged stack (see trace(), BVH branch);
                            // 2: the shipped loop + distance-tagged stack only.  Both bit-identical and both measured SLOWER on
                            // cfg 4: 21.6 / 21.2 vs 20.55 ms (profiles/r04q_cfg4_bvh_loop_variants.log), kept as switches
#endif
#ifndef PSDR_TRACE_NOINLINE
#define PSDR_TRACE_NOINLINE 0   // 1: one out-of-line copy of the closest-hit query per kernel (instruction-cache experiments)
#endif
#if PSDR_TRACE_NOINLINE
template <int kCfg> __device__ __noinline__ Hit trace(const DScene &sc, V3f o, V3f d) {
#else
template <int kCfg> __device__ __forceinline__ Hit trace(const DScene &sc, V3f o, V3f d) {
#endif
    HitCand best;
    hit_init(best);
    Hit miss;
    miss.tri = -1;
    miss.u = miss.v = 0.f;
    miss.t = kTraceTMax;
    constexpr bool kBvh = (kCfg & kCfgBvh) != 0;
    constexpr bool kUniform = (kCfg & kCfgUniformScan) != 0;
    const bool nan_ray = isnan(o.x) || isnan(o.y) || isnan(o.z) || isnan(d.x) || isnan(d.y) || isnan(d.z);
    if ((kBvh || !kUniform) && nan_ray) return miss;
    if (!kBvh) {
        // kUniform: the lanes that trace together (converged on entry); nothing below returns before the warp-level reduction
        const unsigned lanes = kUniform ? __activemask() : 0u;
        const int n_pairs = (sc.n_tris + 1) >> 1;
        V3p O;
        O.x = f2_dup(o.x); O.y = f2_dup(o.y); O.z = f2_dup(o.z);
        unsigned mask = n_pairs >= 32 ? 0xffffffffu : ((1u << n_pairs) - 1u);
#if PSDR_BRUTE_CULL
        {
            const float ix = 1.f / d.x, iy = 1.f / d.y, iz = 1.f / d.z;
            const F2 IX = f2_dup(ix), IY = f2_dup(iy), IZ = f2_dup(iz);
            const F2 AX = f2_dup(fabsf(ix)), AY = f2_dup(fabsf(iy)), AZ = f2_dup(fabsf(iz));
            const F2 NIX = f2_dup(-ix), NIY = f2_dup(-iy), NIZ = f2_dup(-iz);   // (negation and |.| become operand modifiers)
            unsigned pass = 0u;
            const int n_box_pairs = (n_pairs + 1) >> 1;
#pragma unroll 2
            for (int k = 0; k < n_box_pairs; ++k) {
                const unsigned long long *w = sc.bg_box + kBruteBoxWords * k;
                F2 cx, cy, cz, hx, hy, hz;
                cx.v = w[0]; cy.v = w[1]; cz.v = w[2]; hx.v = w[3]; hy.v = w[4]; hz.v = w[5];
                // far = (c - o) * inv + h * |inv|, near = (c - o) * inv - h * |inv|, each ONE fused multiply-add (ptxas
                // contracts a packed mul.rn + add.rn pair anyway, so the fusion is spelled out and the oracle uses fmaf);
                // near is formed negated -- fma(a, -b, c) = -fma(a, b, -c) exactly -- so that h * |inv| is shared
                const F2 dx = f2_sub(cx, O.x), dy = f2_sub(cy, O.y), dz = f2_sub(cz, O.z);
                const F2 thx = f2_mul(hx, AX), thy = f2_mul(hy, AY), thz = f2_mul(hz, AZ);
                const F2 fx = f2_fma(dx, IX, thx), fy = f2_fma(dy, IY, thy), fz = f2_fma(dz, IZ, thz);
                const F2 mx = f2_fma(dx, NIX, thx), my = f2_fma(dy, NIY, thy), mz = f2_fma(dz, NIZ, thz);   // -near
                float mx0, mx1, my0, my1, mz0, mz1, fx0, fx1, fy0, fy1, fz0, fz1;
                f2_split(mx, mx0, mx1); f2_split(my, my0, my1); f2_split(mz, mz0, mz1);
                f2_split(fx, fx0, fx1); f2_split(fy, fy0, fy1); f2_split(fz, fz0, fz1);
                // fmaxf / fminf drop the NaNs of 0 * inf (an axis the ray does not move along constrains nothing);
                // t_near = max(near.xyz, RayEpsilon) = -min(-near.xyz, -RayEpsilon)
                const float tn0 = -fminf(fminf(fminf(mx0, my0), mz0), -kRayEpsilon), tf0 = fminf(fminf(fx0, fy0), fz0);
                const float tn1 = -fminf(fminf(fminf(mx1, my1), mz1), -kRayEpsilon), tf1 = fminf(fminf(fx1, fy1), fz1);
                pass |= (tn0 <= tf0 ? 1u : 0u) << (2 * k);
                pass |= (tn1 <= tf1 ? 2u : 0u) << (2 * k);
            }
            mask &= pass;
        }
#endif
        if (kUniform && nan_ray) mask = 0u;
        if (!kUniform && mask == 0u) return miss;
        V3p D, ND;
        D.x = f2_dup(d.x); D.y = f2_dup(d.y); D.z = f2_dup(d.z);
        ND.x = f2_dup(-d.x); ND.y = f2_dup(-d.y); ND.z = f2_dup(-d.z);
        const ulonglong2 *tab = brute_table();
        const F2 EPSV = f2_dup(kRayEpsilon);
        // kUniform: every lane runs the loop as often as the lane with the most candidates -- the trip count is the same
        // register value in all lanes, so the loop branch never diverges and the lanes stay converged through it.  A plain
        // `while (mask)` leaves the lanes split by trip count for the rest of the caller wherever ptxas places no
        // reconvergence point behind the loop: the secondary-edge adjoint ran its second and third trace with 5.7 of 27
        // lanes (profiles/r02e; 3.6 -> 1.7 ms with the uniform loop).  The forward kernels and the other adjoints do
        // reconverge behind the plain loop and lose 8 % to the extra vote, so the choice is per kernel (kCfgUniformScan).
        auto test_pair = [&](int j) {
                const ulonglong2 *w = tab + kBruteSmemStride * j;
                const ulonglong2 w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
                V3p P0, E1, E2, S, NS;
                P0.x.v = w0.x; P0.y.v = w0.y; P0.z.v = w1.x;
                E1.x.v = w1.y; E1.y.v = w2.x; E1.z.v = w2.y;
                E2.x.v = w3.x; E2.y.v = w3.y; E2.z.v = w4.x;
                const V3p h = cross_fms_na(D, ND, E2);
                const F2 det = dot(E1, h);
                S.x = f2_sub(O.x, P0.x); S.y = f2_sub(O.y, P0.y); S.z = f2_sub(O.z, P0.z);
                NS.x = f2_sub(P0.x, O.x); NS.y = f2_sub(P0.y, O.y); NS.z = f2_sub(P0.z, O.z);   // = -S exactly
                const F2 un = dot(S, h);
                // cross_fms(S, E1): fma(S.y, E1.z, -(S.z * E1.y)) with the sign carried by -S (exact)
                V3p q;
                q.x = f2_fma(S.y, E1.z, f2_mul(NS.z, E1.y));
                q.y = f2_fma(S.z, E1.x, f2_mul(NS.x, E1.z));
                q.z = f2_fma(S.x, E1.y, f2_mul(NS.y, E1.x));
                const F2 vn = dot(D, q), tn = dot(E2, q);
                // sign normalisation by an exact packed multiply with copysign(1, det); a zero determinant cannot pass:
                // its rhs below is 0 and lhs >= 0, so "adet > 0" is implied by the strict "closer" of the ordered scan
                float det0, det1;
                f2_split(det, det0, det1);
                const F2 SG = f2_pack(__int_as_float((__float_as_int(det0) & 0x80000000) | 0x3f800000),
                                      __int_as_float((__float_as_int(det1) & 0x80000000) | 0x3f800000));
                const F2 US = f2_mul(un, SG), VS = f2_mul(vn, SG), TS = f2_mul(tn, SG), AD = f2_mul(det, SG);
                const F2 SUM = f2_add(US, VS), EPS = f2_mul(AD, EPSV);
                float us0, us1, vs0, vs1, ts0, ts1, ad0, ad1, sum0, sum1, eps0, eps1;
                f2_split(US, us0, us1); f2_split(VS, vs0, vs1); f2_split(TS, ts0, ts1);
                f2_split(AD, ad0, ad1); f2_split(SUM, sum0, sum1); f2_split(EPS, eps0, eps1);
                hit_consider_ordered(us0, vs0, sum0, ts0, ad0, eps0, 2 * j, best);
                hit_consider_ordered(us1, vs1, sum1, ts1, ad1, eps1, 2 * j + 1, best);
        };
        if (kUniform) {
            const int n_it = __reduce_max_sync(lanes, __popc(mask));
#pragma unroll 1
            for (int it = 0; it < n_it; ++it) {
                if (mask == 0u) continue;
                const int j = __ffs(mask) - 1;
                mask &= mask - 1u;
                test_pair(j);
            }
        } else {
            while (mask) {
                const int j = __ffs(mask) - 1;
                mask &= mask - 1u;
                test_pair(j);
            }
        }
        if (best.tri < 0) return miss;
        const float *f = reinterpret_cast<const float *>(tab + kBruteSmemStride * (best.tri >> 1)) + (best.tri & 1);
        return hit_finish(best, V3f(f[0], f[2], f[4]), V3f(f[6], f[8], f[10]), V3f(f[12], f[14], f[16]), o, d);
    } else {
        // BVH2 with both child boxes in the parent record and near-child-first descent: the nearer subtree usually yields a
        // hit that prunes the farther one.  Leaves (<= 4 triangles, stored contiguously in traversal order) are tested as
        // soon as their box is hit.  The slab test is conservative (padded boxes, slack on the comparison), candidates
        // pass the same numerator test as the brute-force scan and ties go to the lowest id whatever the visiting order.
        const float ix = 1.f / d.x, iy = 1.f / d.y, iz = 1.f / d.z;
#if PSDR_BVH_LOOP == 1
        // ONE unit of work per loop iteration -- a triangle test of the pending leaf, or a node visit -- so that lanes in
        // a leaf and lanes descending stay in the same loop instead of one group waiting for the other's inner loop
        // (profiles/r02w: 9 - 11 of 32 lanes active in the kernels of cfg 4 with the leaf loop inline); entries popped from the
        // stack carry their entry distance and are skipped when a closer hit has been found since they were pushed.
        int stack[32];
        float stack_t[32];
        int sp = 0;
        int node = 0;
        float best_t = kTraceTMax;          // conservative bound for the box pruning only
        int lf0 = 0, ln0 = 0, lf1 = 0, ln1 = 0;          // pending leaf slots: [lf0, lf0 + ln0) first, then [lf1, lf1 + ln1)
        float lt1 = 0.f;                                 // entry distance of the second pending leaf
        while (true) {
            if (ln0 > 0) {
                const float4 a = __ldg(sc.leaf_tri + 3 * lf0), b = __ldg(sc.leaf_tri + 3 * lf0 + 1), c2 = __ldg(sc.leaf_tri + 3 * lf0 + 2);
                tri_test<false>(V3f(a.x, a.y, a.z), V3f(b.x, b.y, b.z), V3f(c2.x, c2.y, c2.z), __float_as_int(a.w), o, d, best);
                ++lf0;
                if (--ln0 == 0) {
                    best_t = best.ts / best.adet * 1.000001f;
                    lf0 = lf1; ln0 = (ln1 > 0 && lt1 <= best_t) ? ln1 : 0; ln1 = 0;      // the nearer leaf may have pruned the other
                }
                continue;
            }
            if (node < 0) {
                if (sp == 0) break;
                --sp;
                if (stack_t[sp] > best_t) continue;          // pruned by a hit found after the push
                node = stack[sp];
            }
            const float4 *nd = reinterpret_cast<const float4 *>(sc.nodes2 + node);
            const float4 n0 = __ldg(nd), n1 = __ldg(nd + 1), n2 = __ldg(nd + 2);
            const int4 lk = __ldg(reinterpret_cast<const int4 *>(nd + 3));
            float tn[2], tf[2];
            {
                const float x0 = (n0.x - o.x) * ix, x1 = (n0.w - o.x) * ix, y0 = (n0.y - o.y) * iy, y1 = (n1.x - o.y) * iy;
                const float z0 = (n0.z - o.z) * iz, z1 = (n1.y - o.z) * iz;
                tn[0] = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));
                tf[0] = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), best_t));
            }
            {
                const float x0 = (n1.z - o.x) * ix, x1 = (n2.y - o.x) * ix, y0 = (n1.w - o.y) * iy, y1 = (n2.z - o.y) * iy;
                const float z0 = (n2.x - o.z) * iz, z1 = (n2.w - o.z) * iz;
                tn[1] = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));
                tf[1] = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), best_t));
            }
            bool h0 = tn[0] <= tf[0] * 1.0000005f + 1e-6f, h1 = tn[1] <= tf[1] * 1.0000005f + 1e-6f;
            int c0 = lk.x, c1 = lk.y;
            if (h0 && h1 && tn[1] < tn[0]) {      // nearer child first
                const int t = c0; c0 = c1; c1 = t;
                const float q = tn[0]; tn[0] = tn[1]; tn[1] = q;
            } else if (!h0) { c0 = c1; tn[0] = tn[1]; h0 = h1; h1 = false; }
            node = -1;
            if (h0) {
                if (c0 < 0) { const int code = ~c0; lf0 = code >> 3; ln0 = code & 7; }
                else node = c0;
            }
            if (h1) {
                if (c1 < 0) {
                    const int code = ~c1;
                    if (ln0 > 0) { lf1 = code >> 3; ln1 = code & 7; lt1 = tn[1]; }
                    else { lf0 = code >> 3; ln0 = code & 7; }
                } else if (node < 0) node = c1;
                else if (sp < 31) { stack[sp] = c1; stack_t[sp] = tn[1]; ++sp; }
            }
        }
#else
        int stack[32];
#if PSDR_BVH_LOOP == 2
        float stack_t[32];      // entry distance of a pushed node: skipped when a closer hit has been found since (variant 2)
#endif
        int sp = 0;
        int node = 0;
        float best_t = kTraceTMax;          // conservative bound for the box pruning only
        auto leaf = [&](int c) {
            const int code = ~c, first = code >> 3, cnt = code & 7;
            for (int k = 0; k < cnt; ++k) {
                const float4 a = __ldg(sc.leaf_tri + 3 * (first + k)), b = __ldg(sc.leaf_tri + 3 * (first + k) + 1), c2 = __ldg(sc.leaf_tri + 3 * (first + k) + 2);
                tri_test<false>(V3f(a.x, a.y, a.z), V3f(b.x, b.y, b.z), V3f(c2.x, c2.y, c2.z), __float_as_int(a.w), o, d, best);
            }
            best_t = best.ts / best.adet * 1.000001f;
        };
        while (true) {
            const float4 *nd = reinterpret_cast<const float4 *>(sc.nodes2 + node);
            const float4 n0 = __ldg(nd), n1 = __ldg(nd + 1), n2 = __ldg(nd + 2);
            const int4 lk = __ldg(reinterpret_cast<const int4 *>(nd + 3));
            // slab tests; fminf/fmaxf drop the NaNs of 0*inf
            float tn[2], tf[2];
            {
                const float x0 = (n0.x - o.x) * ix, x1 = (n0.w - o.x) * ix, y0 = (n0.y - o.y) * iy, y1 = (n1.x - o.y) * iy;
                const float z0 = (n0.z - o.z) * iz, z1 = (n1.y - o.z) * iz;
                tn[0] = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));
                tf[0] = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), best_t));
            }
            {
                const float x0 = (n1.z - o.x) * ix, x1 = (n2.y - o.x) * ix, y0 = (n1.w - o.y) * iy, y1 = (n2.z - o.y) * iy;
                const float z0 = (n2.x - o.z) * iz, z1 = (n2.w - o.z) * iz;
                tn[1] = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));
                tf[1] = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), best_t));
            }
            bool h0 = tn[0] <= tf[0] * 1.0000005f + 1e-6f, h1 = tn[1] <= tf[1] * 1.0000005f + 1e-6f;
            int c0 = lk.x, c1 = lk.y;
            if (h0 && h1 && tn[1] < tn[0]) {      // nearer child first
                const int t = c0; c0 = c1; c1 = t;
                const float q = tn[0]; tn[0] = tn[1]; tn[1] = q;
            } else if (!h0) { c0 = c1; tn[0] = tn[1]; h0 = h1; h1 = false; }
            // c0 = the (nearer) hit child if h0, c1 = the other hit child if h1
            int next = -1;
            if (h0) {
                if (c0 < 0) leaf(c0);
                else next = c0;
            }
            if (h1 && tn[1] <= best_t) {          // the nearer leaf may have pruned it
                if (c1 < 0) leaf(c1);
                else if (next < 0) next = c1;
                else if (sp < 31) {
#if PSDR_BVH_LOOP == 2
                    stack_t[sp] = tn[1];
#endif
                    stack[sp++] = c1;
                }
            }
            if (next < 0) {
#if PSDR_BVH_LOOP == 2
                while (sp > 0 && stack_t[sp - 1] > best_t) --sp;
#endif
                if (sp == 0) break;
                next = stack[--sp];
            }
            node = next;
        }
#endif
        if (best.tri < 0) return miss;
        const float4 a = __ldg(sc.geo + 3 * best.tri), b = __ldg(sc.geo + 3 * best.tri + 1);
        const float c = __ldg(&sc.geo[3 * best.tri + 2].x);
        return hit_finish(best, V3f(a.x, a.y, a.z), V3f(a.w, b.x, b.y), V3f(b.z, b.w, c), o, d);
    }
}

// reference include/psdr/core/frame.h:9-28 (Duff et al.)
template <class S> __device__ __forceinline__ void coordinate_system(V3<S> n, V3<S> &s, V3<S> &t) {
    const bool neg = signbit_(val(n.z));
    const float sign = neg ? -1.f : 1.f;
    const S a = -rcp_(sign + n.z);
    const S b = n.x * n.y * a;
    const S sx = sqr(n.x) * a;
    s = V3<S>((neg ? -sx : sx) + 1.f, neg ? -b : b, neg ? n.x : -n.x);
    t = V3<S>(b, sign + sqr(n.y) * a, -n.y);
}

template <class S> struct Its {   // reference include/psdr/core/intersection.h:23-60
    bool valid;
    int mesh, tri;
    V3<S> p, n, wi, sh_s, sh_t, sh_n;
    S t, J;
    float bu, bv;   // detached barycentrics of the hit (p = p0 + bu e1 + bv e2)
    V2<S> uv;       // texture coordinate (differentiable at the primary hit of renderD, scene.cpp:755-788)
    V2<S> bc;       // its.bc: the barycentrics as the reference keeps them (tangent-carrying at the analytic primary hit)
    V3<S> dp_du;    // world-space position derivative along u (zero without UVs); bc and dp_du are read by the extended
                    // material set only (MicrofacetPerVertex, NormalMap) and vanish from the other kernel families
    __device__ __forceinline__ V3<S> to_local(V3<S> v) const { return V3<S>(dot(v, sh_s), dot(v, sh_t), dot(v, sh_n)); }
    __device__ __forceinline__ V3<S> to_world(V3<S> v) const { return sh_s * v.x + sh_t * v.y + sh_n * v.z; }
};

// reference include/psdr/utils.h:82-93
// rcp() as Dr.Jit emits it for CUDA floats: rcp.approx.ftz.f32 (ext/drjit/ext/drjit-core/src/cuda_eval.cpp:638-640),
// within 1 ulp of 1/x but not correctly rounded.  Used only on request (DScene::ref_rcp): the primary hit of renderD is
// reconstructed as o + t d, and on faces lit at grazing angles the share of next-event rays that re-intersect the face
// itself depends on which side of the surface that rounding leaves the point (DESIGN.md "parity").
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_sel(float a, bool approx) { return approx ? rcp_approx(a) : 1.f / a; }
__device__ __forceinline__ Dual rcp_sel(Dual a, bool approx) {
    if (!approx) return rcp_(a);
    const float r = rcp_approx(a.v);
    return Dual(r, -a.d * r * r);
}
template <class S>
__device__ __forceinline__ void ray_intersect_triangle(V3<S> p0, V3<S> e1, V3<S> e2, V3<S> o, V3<S> d, S &u, S &v, S &t, bool approx_rcp = false) {
    const V3<S> h = cross(d, e2);
    const S a = dot(e1, h);
    const S f = rcp_sel(a, approx_rcp);
    const V3<S> s = o - p0;
    u = f * dot(s, h);
    const V3<S> q = cross(s, e1);
    v = f * dot(d, q);
    t = f * dot(e2, q);
}

template <class S> struct IsDual { static constexpr bool value = false; };
template <> struct IsDual<Dual> { static constexpr bool value = true; };

// Scene::ray_intersect<ad, path_space> (reference src/scene/scene.cpp:612-806).  The material-form
// ("path-space") variant pins the hit to the triangle by detached barycentrics; the solid-angle
// variant (S = Dual, path_space = false: primary hits) re-intersects analytically.
// kAD selects the formulas of the reference's ad=true instantiation; it defaults to "S carries a
// tangent", the adjoint's primal replay uses <float, kBvh, true> to walk the same path as renderD.
template <class S, int kCfg, bool kAD = IsDual<S>::value>
__device__ __forceinline__ Its<S> ray_intersect(const DScene &sc, V3<S> o, V3<S> d, bool active, bool path_space, int *out_tri = nullptr) {
    constexpr bool ad = kAD;
    Its<S> its;
    its.valid = false;
    its.mesh = -1;
    its.tri = -1;
    its.t = S(0.f);
    its.J = S(1.f);
    its.bu = its.bv = 0.f;
    if (out_tri) *out_tri = -1;
    if (!active) return its;
    const Hit h = trace<kCfg>(sc, val(o), val(d));
    if (h.tri < 0) return its;
    if (out_tri) *out_tri = h.tri;
    const TriRec<S> T = load_tri<S>(sc, h.tri);
    const ShadeRec<S> N = load_shade<S>(sc, h.tri);
    its.valid = true;
    its.tri = h.tri;
    its.mesh = T.mesh;
    its.n = N.fn;
    const DMesh mesh = sc.meshes[T.mesh];
    V3<S> sh_n, dir;
    S bary_u(0.f), bary_v(0.f);
    if (!ad || path_space) {
        const V2f uv(h.u, h.v);
        bary_u = S(h.u);
        bary_v = S(h.v);
        sh_n = normalize(bilinear(N.n0, N.n1 - N.n0, N.n2 - N.n0, uv));
        its.p = bilinear(T.p0, T.e1, T.e2, uv);
        dir = its.p - o;
        its.t = norm(dir);
        dir = dir / its.t;
        if (ad) its.J = T.area / detach(T.area);
        its.bu = h.u;
        its.bv = h.v;
    } else {
        S u, v, t;
        ray_intersect_triangle(T.p0, T.e1, T.e2, o, d, u, v, t, sc.ref_rcp != 0);
        const V2<S> uv(u, v);
        bary_u = u;
        bary_v = v;
        sh_n = normalize(bilinear(N.n0, N.n1 - N.n0, N.n2 - N.n0, uv));
        its.p = V3<S>(fmadd(d.x, t, o.x), fmadd(d.y, t, o.y), fmadd(d.z, t, o.z));
        its.t = t;
        dir = d;
        its.bu = val(u);
        its.bv = val(v);
    }
    if (mesh.flags & 1) sh_n = its.n;
    its.sh_n = sh_n;
    coordinate_system(sh_n, its.sh_s, its.sh_t);
    its.uv = V2<S>(S(0.f), S(0.f));
    its.bc = V2<S>(bary_u, bary_v);
    its.dp_du = V3<S>(S(0.f));
    if (mesh.flags & 2) {   // uv-derived tangent frame (scene.cpp:716-728, 757-765)
        const float2 t0 = __ldg(sc.uv + 3 * h.tri), t1 = __ldg(sc.uv + 3 * h.tri + 1), t2 = __ldg(sc.uv + 3 * h.tri + 2);
        const float du0x = t1.x - t0.x, du0y = t1.y - t0.y, du1x = t2.x - t0.x, du1y = t2.y - t0.y;
        its.uv = V2<S>(fmadd(S(du0x), bary_u, fmadd(S(du1x), bary_v, S(t0.x))), fmadd(S(du0y), bary_u, fmadd(S(du1y), bary_v, S(t0.y))));
        const float det = du0x * du1y - du0y * du1x;
        if (det != 0.f) {
            const float inv_det = 1.f / det;
            const V3<S> dp_du = (T.e1 * S(du1y) - T.e2 * S(du0y)) * S(inv_det);
            its.dp_du = dp_du;
            its.sh_s = normalize(dp_du - sh_n * dot(sh_n, dp_du));
            its.sh_t = cross(sh_n, its.sh_s);
        }
    }
    its.wi = its.to_local(-dir);
    return its;
}

// ---- Diffuse BSDF (reference src/bsdf/diffuse.cpp:23-108) ------------------------------------
template <class S> __device__ __forceinline__ V3<S> bsdf_reflectance_const(const DBsdf &b);
template <> __device__ __forceinline__ V3f bsdf_reflectance_const<float>(const DBsdf &b) { return V3f(b.refl[0], b.refl[1], b.refl[2]); }
template <> __device__ __forceinline__ V3d bsdf_reflectance_const<Dual>(const DBsdf &b) {
    return V3d(Dual(b.refl[0], b.d_refl[0]), Dual(b.refl[1], b.d_refl[1]), Dual(b.refl[2], b.d_refl[2]));
}
// Bitmap::eval(its.uv): 1x1 -> the constant, else transformed bilinear texture lookup (reference src/core/bitmap.cpp:46-131);
// textures belong to the "full" kernel family
template <class S, int kCfg> __device__ __forceinline__ V3<S> bsdf_reflectance(const DBsdf &b, V2<S> uv) {
    if ((kCfg & kCfgExt) && b.tex[0].w > 0) return tex_eval_uv<S>(b.tex[0], IsDual<S>::value, uv);
    return bsdf_reflectance_const<S>(b);
}

template <class S> __device__ __forceinline__ V3<S> bsdf_specular_const(const DBsdf &b);
template <> __device__ __forceinline__ V3f bsdf_specular_const<float>(const DBsdf &b) { return V3f(b.spec[0], b.spec[1], b.spec[2]); }
template <> __device__ __forceinline__ V3d bsdf_specular_const<Dual>(const DBsdf &b) {
    return V3d(Dual(b.spec[0], b.d_spec[0]), Dual(b.spec[1], b.d_spec[1]), Dual(b.spec[2], b.d_spec[2]));
}
template <class S, int kCfg> __device__ __forceinline__ V3<S> bsdf_specular(const DBsdf &b, V2<S> uv) {     // Microfacet: full family only
    if ((kCfg & kCfgExt) && b.tex[1].w > 0) return tex_eval_uv<S>(b.tex[1], IsDual<S>::value, uv);
    return bsdf_specular_const<S>(b);
}
template <class S> __device__ __forceinline__ S bsdf_roughness_const(const DBsdf &b);
template <> __device__ __forceinline__ float bsdf_roughness_const<float>(const DBsdf &b) { return b.rough; }
template <> __device__ __forceinline__ Dual bsdf_roughness_const<Dual>(const DBsdf &b) { return Dual(b.rough, b.d_rough); }
template <class S, int kCfg> __device__ __forceinline__ S bsdf_roughness(const DBsdf &b, V2<S> uv) {
    if ((kCfg & kCfgExt) && b.tex[2].w > 0) return tex_eval_uv<S>(b.tex[2], IsDual<S>::value, uv).x;
    return bsdf_roughness_const<S>(b);
}

// ---- GGX (reference src/bsdf/ggx.cpp:13-109), isotropic: alpha_u = alpha_v = alpha ---------------
template <class S> __device__ __forceinline__ S ggx_eval(S alpha, V3<S> m) {
    const S alpha_uv = alpha * alpha;
    const S t = sqr(m.x / alpha) + sqr(m.y / alpha) + sqr(m.z);
    const S result = rcp_(S(kPi) * alpha_uv * sqr(t));
    return (val(result) * val(m.z) > 1e-20f) ? result : S(0.f);
}
template <class S> __device__ __forceinline__ S ggx_smith_g1(S alpha, V3<S> v, V3<S> m) {
    const S xy_alpha_2 = sqr(alpha * v.x) + sqr(alpha * v.y);
    const S tan_theta_alpha_2 = xy_alpha_2 / sqr(v.z);
    S result = S(2.f) / (S(1.f) + sqrt_(S(1.f) + tan_theta_alpha_2));
    if (val(xy_alpha_2) == 0.f) result = S(1.f);
    if (val(dot(v, m)) * val(v.z) <= 0.f) result = S(0.f);
    return result;
}

// Microfacet::__eval (reference src/bsdf/microfacet.cpp:22-68): Lambert + GGX specular with the
// Schlick-Gaussian Fresnel fit; wi, wo in the shading frame
// The full-feature kernels are 250 KB of code with everything inlined (each BSDF evaluation twice per path step, in
// dual numbers) against a 32 KB instruction cache; PSDR_FULL_OUTLINE = 1 keeps ONE copy of the Microfacet / conductor /
// environment-map routines per kernel (calls instead of inlining).
#ifndef PSDR_FULL_OUTLINE
#define PSDR_FULL_OUTLINE 0
#endif
#if PSDR_FULL_OUTLINE
#define PSDR_FULL_FN __device__ __noinline__
#else
#define PSDR_FULL_FN __device__ __forceinline__
#endif
template <class S, int kCfg> PSDR_FULL_FN V3<S> microfacet_eval(const DBsdf &b, V3<S> wi, V3<S> wo, V2<S> uv) {
    if (b.two_side) {
        if (signbit_(val(wi.z))) wo.z = -wo.z;
        wi.z = abs_(wi.z);
    }
    const S cos_nv = wi.z, cos_nl = wo.z;
    if (!(val(cos_nv) > 0.f && val(cos_nl) > 0.f)) return V3<S>(S(0.f));
    const V3<S> diffuse = bsdf_reflectance<S, kCfg>(b, uv) * S(kInvPi);
    const V3<S> H = normalize(wi + wo);
    const S cos_vh = dot(H, wi);
    const V3<S> F0 = bsdf_specular<S, kCfg>(b, uv);
    const S alpha = sqr(bsdf_roughness<S, kCfg>(b, uv));
    const S ggx = ggx_eval<S>(alpha, H);
    const S coeff = cos_vh * (S(-5.55473f) * cos_vh - S(6.8316f));
    const S e = exp2_(coeff);
    const V3<S> fresnel = F0 + (V3<S>(S(1.f)) - F0) * e;
    const S smithG = ggx_smith_g1<S>(alpha, wi, H) * ggx_smith_g1<S>(alpha, wo, H);
    const V3<S> numerator = fresnel * (ggx * smithG);
    const S denominator = S(4.f) * cos_nl * cos_nv;
    const V3<S> specular = numerator / (denominator + S(1e-6f));
    return (diffuse + specular) * cos_nl;
}

// Fresnel reflectance of a conductor with complex index eta + i k, unpolarised (reference include/psdr/utils.h:167-183)
template <class S> __device__ __forceinline__ S fresnel_conductor(S eta, S k, S cos_theta_i) {
    const S c2 = sqr(cos_theta_i), s2 = S(1.f) - c2, s4 = sqr(s2);
    const S temp_1 = sqr(eta) - sqr(k) - s2;
    const S a_2_pb_2 = safe_sqrt(sqr(temp_1) + S(4.f) * sqr(k * eta));
    const S a = safe_sqrt(S(.5f) * (a_2_pb_2 + temp_1));
    const S term_1 = a_2_pb_2 + c2, term_2 = S(2.f) * cos_theta_i * a;
    const S r_s = (term_1 - term_2) / (term_1 + term_2);
    const S term_3 = a_2_pb_2 * c2 + s4, term_4 = term_2 * s2;
    const S r_p = r_s * (term_3 - term_4) / (term_3 + term_4);
    return S(.5f) * (r_s + r_p);
}
template <class S> __device__ __forceinline__ V3<S> bsdf_eta(const DBsdf &b);
template <> __device__ __forceinline__ V3f bsdf_eta<float>(const DBsdf &b) { return V3f(b.eta[0], b.eta[1], b.eta[2]); }
template <> __device__ __forceinline__ V3d bsdf_eta<Dual>(const DBsdf &b) { return V3d(Dual(b.eta[0], b.d_eta[0]), Dual(b.eta[1], b.d_eta[1]), Dual(b.eta[2], b.d_eta[2])); }
template <class S> __device__ __forceinline__ V3<S> bsdf_k(const DBsdf &b);
template <> __device__ __forceinline__ V3f bsdf_k<float>(const DBsdf &b) { return V3f(b.kk[0], b.kk[1], b.kk[2]); }
template <> __device__ __forceinline__ V3d bsdf_k<Dual>(const DBsdf &b) { return V3d(Dual(b.kk[0], b.d_kk[0]), Dual(b.kk[1], b.d_kk[1]), Dual(b.kk[2], b.d_kk[2])); }

// RoughConductor::__eval (reference src/bsdf/roughconductor.cpp:38-66), isotropic (alpha_u = alpha_v):
// F(eta, k, <wi, H>) D(H) G(wi, wo, H) / (4 cos_theta_i) * specular_reflectance
template <class S, int kCfg> PSDR_FULL_FN V3<S> conductor_eval(const DBsdf &b, V3<S> wi, V3<S> wo, V2<S> uv) {
    if (b.two_side) {
        if (signbit_(val(wi.z))) wo.z = -wo.z;
        wi.z = abs_(wi.z);
    }
    if (!(val(wi.z) > 0.f && val(wo.z) > 0.f)) return V3<S>(S(0.f));
    const S alpha = bsdf_roughness<S, kCfg>(b, uv);
    const V3<S> H = normalize(wo + wi);
    const S D = ggx_eval<S>(alpha, H);
    if (val(D) == 0.f) return V3<S>(S(0.f));
    const S G = ggx_smith_g1<S>(alpha, wi, H) * ggx_smith_g1<S>(alpha, wo, H);
    const S result = D * G / (S(4.f) * wi.z);
    const S cos_ih = dot(wi, H);
    const V3<S> eta = bsdf_eta<S>(b), k = bsdf_k<S>(b);
    const V3<S> F(fresnel_conductor<S>(eta.x, k.x, cos_ih), fresnel_conductor<S>(eta.y, k.y, cos_ih), fresnel_conductor<S>(eta.z, k.z, cos_ih));
    return F * result * bsdf_specular<S, kCfg>(b, uv);
}

// ---- extended material set (kCfgExt) --------------------------------------------------------------
// fresnel_dielectric (reference include/psdr/utils.h:185-215): unpolarised Fresnel reflectance of a dielectric interface
// with relative index eta, plus the cosine of the transmitted direction and the two relative indices
template <class S> struct FresnelDielectric {
    S r, cos_theta_t, eta_it, eta_ti;
};
template <class S> __device__ __forceinline__ FresnelDielectric<S> fresnel_dielectric(S eta, S cos_theta_i) {
    FresnelDielectric<S> f;
    const bool outside = val(cos_theta_i) >= 0.f;
    const S rcp_eta = rcp_(eta);
    f.eta_it = outside ? eta : rcp_eta;
    f.eta_ti = outside ? rcp_eta : eta;
    const S sin_i_2 = fmadd(-cos_theta_i, cos_theta_i, S(1.f));
    const S cos_theta_t_sqr = fmadd(-sin_i_2, f.eta_ti * f.eta_ti, S(1.f));
    const S ci_abs = abs_(cos_theta_i), ct_abs = safe_sqrt(cos_theta_t_sqr);
    const bool index_matched = val(eta) == 1.f, special_case = index_matched || val(ci_abs) == 0.f;
    const S a_s = fmadd(-f.eta_it, ct_abs, ci_abs) / fmadd(f.eta_it, ct_abs, ci_abs);
    const S a_p = fmadd(-f.eta_it, ci_abs, ct_abs) / fmadd(f.eta_it, ci_abs, ct_abs);
    f.r = S(.5f) * (sqr(a_s) + sqr(a_p));
    if (special_case) f.r = S(index_matched ? 0.f : 1.f);
    f.cos_theta_t = signbit_(val(cos_theta_i)) ? ct_abs : -ct_abs;      // mulsign_neg(cos_theta_t_abs, cos_theta_i)
    return f;
}
template <class S> __device__ __forceinline__ S bsdf_scalar(float v, float d);
template <> __device__ __forceinline__ float bsdf_scalar<float>(float v, float) { return v; }
template <> __device__ __forceinline__ Dual bsdf_scalar<Dual>(float v, float d) { return Dual(v, d); }

// RoughDielectric::__eval (reference src/bsdf/roughdielectric.cpp:37-122): GGX reflection + refraction through an
// interface with eta = intIOR / extIOR (DBsdf::eta[0]) and its separately rounded inverse extIOR / intIOR (eta[1]);
// alpha_u = alpha_v = the roughness slot
template <class S, int kCfg> PSDR_FULL_FN V3<S> dielectric_eval(const DBsdf &b, V3<S> wi, V3<S> wo, V2<S> uv) {
    if (b.two_side) {
        if (signbit_(val(wi.z))) wo.z = -wo.z;
        wi.z = abs_(wi.z);
    }
    const S cos_theta_i = wi.z, cos_theta_o = wo.z;
    if (val(cos_theta_i) == 0.f) return V3<S>(S(0.f));
    const bool reflect = val(cos_theta_i) * val(cos_theta_o) > 0.f;
    const S m_eta = bsdf_scalar<S>(b.eta[0], b.d_eta[0]), m_inv_eta = bsdf_scalar<S>(b.eta[1], b.d_eta[1]);
    const bool front = val(cos_theta_i) > 0.f;
    const S eta = front ? m_eta : m_inv_eta, inv_eta = front ? m_inv_eta : m_eta;
    V3<S> m = normalize(wi + wo * (reflect ? S(1.f) : eta));
    if (signbit_(val(m.z))) m = -m;                                     // mulsign(m, cos_theta(m))
    const S alpha = bsdf_roughness<S, kCfg>(b, uv);
    const S D = ggx_eval<S>(alpha, m);
    const S F = fresnel_dielectric<S>(m_eta, dot(wi, m)).r;
    const S G = ggx_smith_g1<S>(alpha, wi, m) * ggx_smith_g1<S>(alpha, wo, m);
    if (reflect) return V3<S>(F * D * G / (S(4.f) * abs_(cos_theta_i)));
    const S scale = sqr(inv_eta);
    const S wi_m = dot(wi, m), wo_m = dot(wo, m);
    const S value = abs_((scale * (S(1.f) - F) * D * G * eta * eta * wi_m * wo_m) / (cos_theta_i * sqr(wi_m + eta * wo_m)));
    return V3<S>(value);
}

// MicrofacetPerVertex::__interpolate (reference src/bsdf/microfacet_pv.cpp:146-160): channel c of the per-vertex table
// (DBsdf::pv, 7 floats per vertex: specular rgb, diffuse rgb, roughness) at the hit's barycentrics; the vertex indices are
// the mesh-local face indices of the triangle (DScene::face_idx)
template <class S> __device__ __forceinline__ S pv_load(const DBsdf &b, int vtx, int c);
template <> __device__ __forceinline__ float pv_load<float>(const DBsdf &b, int vtx, int c) { return __ldg(b.pv + 7 * vtx + c); }
template <> __device__ __forceinline__ Dual pv_load<Dual>(const DBsdf &b, int vtx, int c) {
    return Dual(__ldg(b.pv + 7 * vtx + c), b.d_pv ? __ldg(b.d_pv + 7 * vtx + c) : 0.f);
}
template <class S> __device__ __forceinline__ S pv_interp(const DScene &sc, const DBsdf &b, const Its<S> &its, int c) {
    const int i0 = __ldg(sc.face_idx + 3 * its.tri), i1 = __ldg(sc.face_idx + 3 * its.tri + 1), i2 = __ldg(sc.face_idx + 3 * its.tri + 2);
    // a table shorter than the mesh's vertex list reads as zeros (Dr.Jit gathers out of range are undefined; we define them)
    const S v0 = i0 < b.pv_n ? pv_load<S>(b, i0, c) : S(0.f), v1 = i1 < b.pv_n ? pv_load<S>(b, i1, c) : S(0.f),
            v2 = i2 < b.pv_n ? pv_load<S>(b, i2, c) : S(0.f);
    return fmadd(v1 - v0, its.bc.x, fmadd(v2 - v0, its.bc.y, v0));
}
// MicrofacetPerVertex::__eval (reference src/bsdf/microfacet_pv.cpp:20-68): Lambert + GGX with the Schlick-Smith k form
// (NOT GGXDistribution: its own NDF / geometry expressions)
template <class S, int kCfg> PSDR_FULL_FN V3<S> microfacet_pv_eval(const DScene &sc, const DBsdf &b, const Its<S> &its, V3<S> wi, V3<S> wo) {
    if (b.two_side) {
        if (signbit_(val(wi.z))) wo.z = -wo.z;
        wi.z = abs_(wi.z);
    }
    const S cos_nv = wi.z, cos_nl = wo.z;
    if (!(val(cos_nv) > 0.f && val(cos_nl) > 0.f)) return V3<S>(S(0.f));
    const V3<S> F0(pv_interp<S>(sc, b, its, 0), pv_interp<S>(sc, b, its, 1), pv_interp<S>(sc, b, its, 2));
    const V3<S> diffuse = V3<S>(pv_interp<S>(sc, b, its, 3), pv_interp<S>(sc, b, its, 4), pv_interp<S>(sc, b, its, 5)) * S(kInvPi);
    const S roughness = pv_interp<S>(sc, b, its, 6);
    const V3<S> H = normalize(wi + wo);
    const S cos_nh = H.z, cos_vh = dot(H, wi);
    const S alpha = sqr(roughness);
    const S k = sqr(roughness + S(1.f)) / S(8.f);
    const S tmp = alpha / (cos_nh * cos_nh * (sqr(alpha) - S(1.f)) + S(1.f));
    const S ggx = tmp * tmp * S(kInvPi);
    const S coeff = cos_vh * (S(-5.55473f) * cos_vh - S(6.8316f));
    const V3<S> fresnel = F0 + (V3<S>(S(1.f)) - F0) * exp2_(coeff);
    const S smithG1 = cos_nv / (cos_nv * (S(1.f) - k) + k);
    const S smithG2 = cos_nl / (cos_nl * (S(1.f) - k) + k);
    const S smithG = smithG1 * smithG2;
    const V3<S> numerator = fresnel * (ggx * smithG);
    const S denominator = S(4.f) * cos_nl * cos_nv;
    const V3<S> specular = numerator / (denominator + S(1e-6f));
    return (diffuse + specular) * cos_nl;
}

// BSDF::eval of a record that is not a NormalMap, with the incident direction given explicitly (NormalMap evaluates its
// nested BSDF with perturbed directions); `its` supplies uv, barycentrics and the triangle
template <class S, int kCfg> __device__ __forceinline__ V3<S> bsdf_eval_leaf(const DScene &sc, const DBsdf &b, const Its<S> &its, V3<S> wi, V3<S> wo) {
    if ((kCfg & kCfgFull) && b.type == 1) return microfacet_eval<S, kCfg>(b, wi, wo, its.uv);
    if ((kCfg & kCfgExt) && b.type == 2) return conductor_eval<S, kCfg>(b, wi, wo, its.uv);
    if ((kCfg & kCfgExt) && b.type == 3) return dielectric_eval<S, kCfg>(b, wi, wo, its.uv);
    if ((kCfg & kCfgExt) && b.type == 4) return microfacet_pv_eval<S, kCfg>(sc, b, its, wi, wo);
    S wiz = wi.z;
    if (b.two_side) {
        if (signbit_(val(wiz))) wo.z = -wo.z;
        wiz = abs_(wiz);
    }
    if (!(val(wiz) > 0.f && val(wo.z) > 0.f)) return V3<S>(S(0.f));
    return bsdf_reflectance<S, kCfg>(b, its.uv) * S(kInvPi) * wo.z;
}

// ---- NormalMap (reference src/bsdf/normalmap.cpp): microfacet-based normal mapping, one tangent facet -----------------
// The perturbed normal wp comes from the normal map (2 rgb - 1, normalised) in the SHADING frame; the reference then builds
// the perturbed frame from wp and its.dp_du, a WORLD-space vector, as written (normalmap.cpp:61) -- reproduced, not fixed.
template <class S> struct NmFrame {
    V3<S> s, t, n;      // Frame(n, s) (include/psdr/core/frame.h:42-45)
    __device__ __forceinline__ V3<S> to_local(V3<S> v) const { return V3<S>(dot(v, s), dot(v, t), dot(v, n)); }
    __device__ __forceinline__ V3<S> to_world(V3<S> v) const { return s * v.x + t * v.y + n * v.z; }
};
template <class S> __device__ __forceinline__ S nm_pdot(V3<S> a, V3<S> b) {      // maximum(0, dot): a NaN dot gives 0, as max.f32 does
    const S d = dot(a, b);
    return val(d) > 0.f ? d : S(0.f);
}
template <class S> __device__ __forceinline__ S nm_sin_theta(V3<S> v) { return safe_sqrt(fmadd(v.x, v.x, sqr(v.y))); }
template <class S> __device__ __forceinline__ V3<S> nm_wt(V3<S> wp) { return normalize(V3<S>(-wp.x, -wp.y, S(0.f))); }
template <class S> __device__ __forceinline__ S nm_G1(V3<S> wp, V3<S> w) {
    const S cw = val(w.z) > 0.f ? w.z : S(0.f), cp = val(wp.z) > 0.f ? wp.z : S(0.f);
    const S g = cw * cp / (nm_pdot<S>(w, wp) + nm_pdot<S>(w, nm_wt<S>(wp)) * nm_sin_theta<S>(wp));
    return val(g) < 1.f ? g : S(1.f);      // minimum(1, g): a NaN g gives 1
}
template <class S> __device__ __forceinline__ S nm_lambda_p(V3<S> wp, V3<S> wi) {
    const S i_dot_p = nm_pdot<S>(wp, wi);
    return i_dot_p / (i_dot_p + nm_pdot<S>(nm_wt<S>(wp), wi) * nm_sin_theta<S>(wp));
}
template <class S, int kCfg> __device__ __forceinline__ void nm_setup(const DBsdf &b, const Its<S> &its, V3<S> &wp, NmFrame<S> &fr) {
    const V3<S> c = bsdf_reflectance<S, kCfg>(b, its.uv);                // the normal map lives in slot 0
    wp = normalize(V3<S>(fmadd(c.x, S(2.f), S(-1.f)), fmadd(c.y, S(2.f), S(-1.f)), fmadd(c.z, S(2.f), S(-1.f))));
    const S d = dot(wp, its.dp_du);
    const V3<S> s0 = normalize(V3<S>(fmadd(-wp.x, d, its.dp_du.x), fmadd(-wp.y, d, its.dp_du.y), fmadd(-wp.z, d, its.dp_du.z)));
    fr.n = wp;
    fr.t = normalize(cross(wp, s0));
    fr.s = normalize(cross(fr.t, wp));
}
template <class S> __device__ __forceinline__ V3<S> nm_reflect(V3<S> w, V3<S> wt) {      // normalize(w - 2 <w, wt> wt)
    const S k = S(2.f) * dot(w, wt);
    return normalize(w - wt * k);
}
// NormalMap::__eval (normalmap.cpp:42-86): i -> p -> o and i -> t -> p -> o
// (nb = the nested record, passed explicitly: the adjoint evaluates this with tangent-free local copies of both records)
template <class S, int kCfg> PSDR_FULL_FN V3<S> normalmap_eval_nb(const DScene &sc, const DBsdf &b, const DBsdf &nb, const Its<S> &its, V3<S> wi, V3<S> wo) {
    if (b.two_side) {
        if (signbit_(val(wi.z))) wo.z = -wo.z;
        wi.z = abs_(wi.z);
    }
    if (!(val(wi.z) > 0.f && val(wo.z) > 0.f)) return V3<S>(S(0.f));
    V3<S> wp;
    NmFrame<S> fr;
    nm_setup<S, kCfg>(b, its, wp, fr);
    const V3<S> p_wo = fr.to_local(wo);
    const S shadowing = nm_G1<S>(wp, wo);
    const S lambda_p = nm_lambda_p<S>(wp, wi);
    const V3<S> wt = nm_wt<S>(wp);
    V3<S> value = bsdf_eval_leaf<S, kCfg>(sc, nb, its, fr.to_local(wi), p_wo) * lambda_p * shadowing;
    if (val(dot(wi, wt)) > 0.f) {
        const V3<S> wi_r = nm_reflect<S>(wi, wt);
        value = value + bsdf_eval_leaf<S, kCfg>(sc, nb, its, fr.to_local(wi_r), p_wo) * ((S(1.f) - lambda_p) * shadowing);
    }
    return value;
}

template <class S, int kCfg> PSDR_FULL_FN V3<S> normalmap_eval(const DScene &sc, const DBsdf &b, const Its<S> &its, V3<S> wi, V3<S> wo) {
    return normalmap_eval_nb<S, kCfg>(sc, b, sc.bsdfs[b.nested], its, wi, wo);
}

template <class S, int kCfg> __device__ __forceinline__ V3<S> bsdf_eval(const DScene &sc, const Its<S> &its, V3<S> wo, bool active) {
    if (!active || !its.valid) return V3<S>(S(0.f));
    const int bi = sc.meshes[its.mesh].bsdf;
    if (bi < 0) return V3<S>(S(0.f));
    const DBsdf &b = sc.bsdfs[bi];
    if ((kCfg & kCfgExt) && b.type == 5) return normalmap_eval<S, kCfg>(sc, b, its, its.wi, wo);
    return bsdf_eval_leaf<S, kCfg>(sc, b, its, its.wi, wo);
}

// alpha of the GGX lobe that __pdf / __sample of a record use (detached): Microfacet and MicrofacetPerVertex square the
// roughness (microfacet.cpp:92,123, microfacet_pv.cpp:93,135), RoughConductor and RoughDielectric store alpha itself
template <class S, int kCfg> __device__ __forceinline__ float bsdf_alpha(const DScene &sc, const DBsdf &b, const Its<S> &its) {
    if ((kCfg & kCfgExt) && b.type == 4) {
        Its<float> f;
        f.tri = its.tri;
        f.bc = V2f(val(its.bc.x), val(its.bc.y));
        return sqr(pv_interp<float>(sc, b, f, 6));
    }
    const float r = bsdf_roughness<float, kCfg>(b, val(its.uv));
    return ((kCfg & kCfgExt) && (b.type == 2 || b.type == 3)) ? r : sqr(r);
}

// Microfacet::__pdf (reference src/bsdf/microfacet.cpp:108-133), detached; RoughConductor::__pdf
// (src/bsdf/roughconductor.cpp:70-95) and MicrofacetPerVertex::__pdf (microfacet_pv.cpp:120-143) are the same expression
static PSDR_FULL_FN float ggx_reflect_pdf(float alpha, bool two_side, V3f wi, V3f wo) {
    if (two_side) {
        if (signbit_(wi.z)) wo.z = -wo.z;
        wi.z = fabsf(wi.z);
    }
    const V3f m = normalize(wo + wi);
    if (!(wi.z > 0.f && wo.z > 0.f && dot(wi, m) > 0.f && dot(wo, m) > 0.f)) return 0.f;
    return ggx_eval<float>(alpha, m) * ggx_smith_g1<float>(alpha, wi, m) / (4.f * wi.z);
}
// RoughDielectric::__pdf (reference src/bsdf/roughdielectric.cpp:125-176)
static PSDR_FULL_FN float dielectric_pdf(const DBsdf &b, float alpha, V3f wi, V3f wo) {
    if (b.two_side) {
        if (signbit_(wi.z)) wo.z = -wo.z;
        wi.z = fabsf(wi.z);
    }
    const float cos_theta_i = wi.z, cos_theta_o = wo.z;
    if (cos_theta_i == 0.f) return 0.f;
    const bool reflect = cos_theta_i * cos_theta_o > 0.f;
    const float eta = cos_theta_i > 0.f ? b.eta[0] : b.eta[1];
    V3f m = normalize(wi + wo * (reflect ? 1.f : eta));
    if (signbit_(m.z)) m = -m;
    const float wi_m = dot(wi, m), wo_m = dot(wo, m);
    if (!(wi_m * wi.z > 0.f && wo_m * wo.z > 0.f)) return 0.f;
    const float dwh_dwo = reflect ? 1.f / (4.f * wo_m) : (eta * eta * wo_m) / sqr(wi_m + eta * wo_m);
    const V3f pwi = signbit_(wi.z) ? -wi : wi;
    float prob = ggx_eval<float>(alpha, m) * ggx_smith_g1<float>(alpha, pwi, m) / pwi.z;
    const float F = fresnel_dielectric<float>(b.eta[0], wi_m).r;
    prob *= reflect ? F : 1.f - F;
    return prob * fabsf(dwh_dwo);
}
template <class S, int kCfg> __device__ __forceinline__ float bsdf_pdf_leaf(const DScene &sc, const DBsdf &b, const Its<S> &its, V3f wi, V3f wo) {
    if ((kCfg & kCfgExt) && b.type == 3) return dielectric_pdf(b, bsdf_alpha<S, kCfg>(sc, b, its), wi, wo);
    if ((kCfg & kCfgFull) && b.type != 0) return ggx_reflect_pdf(bsdf_alpha<S, kCfg>(sc, b, its), b.two_side != 0, wi, wo);
    float wiz = wi.z, woz = wo.z;
    if (b.two_side) {
        if (signbit_(wiz)) woz = -woz;
        wiz = fabsf(wiz);
    }
    if (!(wiz > 0.f && woz > 0.f)) return 0.f;
    return kInvPi * woz;
}
// NormalMap::__pdf (normalmap.cpp:109-144), detached
template <class S, int kCfg> PSDR_FULL_FN float normalmap_pdf(const DScene &sc, const DBsdf &b, const Its<S> &its, V3f wi, V3f wo) {
    if (b.two_side) {
        if (signbit_(wi.z)) wo.z = -wo.z;
        wi.z = fabsf(wi.z);
    }
    if (!(wi.z > 0.f && wo.z > 0.f)) return 0.f;
    const DBsdf &nb = sc.bsdfs[b.nested];
    Its<float> f;
    f.tri = its.tri;
    f.uv = val(its.uv);
    f.bc = V2f(val(its.bc.x), val(its.bc.y));
    f.dp_du = val(its.dp_du);
    V3f wp;
    NmFrame<float> fr;
    nm_setup<float, kCfg>(b, f, wp, fr);
    const V3f p_wo = fr.to_local(wo);
    const float probability_wp = nm_lambda_p<float>(wp, wi);
    const V3f wi_r = nm_reflect<float>(wi, nm_wt<float>(wp));
    return probability_wp * bsdf_pdf_leaf<float, kCfg>(sc, nb, f, fr.to_local(wi), p_wo) +
           (1.f - probability_wp) * bsdf_pdf_leaf<float, kCfg>(sc, nb, f, fr.to_local(wi_r), p_wo);
}

template <class S, int kCfg> __device__ __forceinline__ float bsdf_pdf(const DScene &sc, const Its<S> &its, V3<S> wo, bool active) {
    if (!active || !its.valid) return 0.f;
    const int bi = sc.meshes[its.mesh].bsdf;
    if (bi < 0) return 0.f;
    const DBsdf &b = sc.bsdfs[bi];
    if ((kCfg & kCfgExt) && b.type == 5) return normalmap_pdf<S, kCfg>(sc, b, its, val(its.wi), val(wo));
    return bsdf_pdf_leaf<S, kCfg>(sc, b, its, val(its.wi), val(wo));
}

struct BsdfSample {
    V3f wo;
    float pdf;
    bool valid;
};

// concentric disk map (reference include/psdr/core/warp.h:15-63) with polynomial sincos
__device__ __forceinline__ V2f square_to_uniform_disk_concentric(V2f s) {
    const float x = fmaf(2.f, s.x, -1.f), y = fmaf(2.f, s.y, -1.f);
    const bool is_zero = (x == 0.f && y == 0.f), q13 = fabsf(x) < fabsf(y);
    const float r = q13 ? y : x, rp = q13 ? x : y;
    float sn, cs;
    sincos_quarter(.25f * kPi * rp / r, sn, cs);
    if (is_zero) { sn = 0.f; cs = 1.f; }
    if (q13 && !is_zero) { const float tmp = sn; sn = cs; cs = tmp; }
    return V2f(r * cs, r * sn);
}

// GGXDistribution::sample_visible_11 (reference src/bsdf/ggx.cpp:99-109)
__device__ __forceinline__ V2f ggx_sample_visible_11(float cos_theta_i, V2f sample) {
    V2f p = square_to_uniform_disk_concentric(sample);
    const float s = .5f * (1.f + cos_theta_i);
    const float a = safe_sqrt(1.f - sqr(p.x));
    p.y = fmaf(p.y, s, fmaf(-a, s, a));                       // lerp(a, p.y, s)
    const float x = p.x, y = p.y, z = safe_sqrt(1.f - fmaf(p.y, p.y, p.x * p.x));
    const float sin_theta_i = safe_sqrt(1.f - sqr(cos_theta_i));
    const float norm = 1.f / fmaf(sin_theta_i, y, cos_theta_i * z);
    return V2f(fmaf(cos_theta_i, y, -(sin_theta_i * z)) * norm, x * norm);
}

// GGXDistribution::sample (reference src/bsdf/ggx.cpp:36-79): visible-normal sample m and its density
__device__ __forceinline__ void ggx_sample_m(float alpha, V3f wi, V3f sample, V3f &m, float &m_pdf) {
    const V3f wi_p = normalize(V3f(alpha * wi.x, alpha * wi.y, wi.z));
    const float sin_theta_2 = fmaf(wi_p.x, wi_p.x, sqr(wi_p.y)), inv_sin_theta = 1.f / sqrtf(sin_theta_2);
    const bool pole = fabsf(sin_theta_2) <= 4.f * kEpsilon;
    const float sin_phi = pole ? 0.f : fminf(fmaxf(wi_p.y * inv_sin_theta, -1.f), 1.f);
    const float cos_phi = pole ? 1.f : fminf(fmaxf(wi_p.x * inv_sin_theta, -1.f), 1.f);
    V2f slope = ggx_sample_visible_11(wi_p.z, V2f(sample.x, sample.y));
    slope = V2f(fmaf(cos_phi, slope.x, -(sin_phi * slope.y)) * alpha, fmaf(sin_phi, slope.x, cos_phi * slope.y) * alpha);
    m = normalize(V3f(-slope.x, -slope.y, 1.f));
    m_pdf = ggx_smith_g1<float>(alpha, wi, m) * fabsf(dot(wi, m)) * ggx_eval<float>(alpha, m) / fabsf(wi.z);
}
// Microfacet::__sample (reference src/bsdf/microfacet.cpp:80-98); RoughConductor::__sample (roughconductor.cpp:99-122) and
// MicrofacetPerVertex::__sample (microfacet_pv.cpp:82-106) are the same with their own alpha
static PSDR_FULL_FN BsdfSample ggx_reflect_sample(float alpha, bool two_side, V3f wi, V3f sample, bool active) {
    BsdfSample bs;
    if (two_side) wi.z = fabsf(wi.z);
    V3f m;
    float m_pdf;
    ggx_sample_m(alpha, wi, sample, m, m_pdf);
    const float k = 2.f * dot(wi, m);
    bs.wo = V3f(fmaf(m.x, k, -wi.x), fmaf(m.y, k, -wi.y), fmaf(m.z, k, -wi.z));
    bs.pdf = m_pdf / (4.f * dot(bs.wo, m));
    bs.valid = active && (wi.z > 0.f) && (bs.pdf != 0.f) && (bs.wo.z > 0.f);
    return bs;
}
// RoughDielectric::__sample (reference src/bsdf/roughdielectric.cpp:179-236): sample.z chooses reflection / refraction
static PSDR_FULL_FN BsdfSample dielectric_sample(const DBsdf &b, float alpha, V3f wi, V3f sample, bool active) {
    BsdfSample bs;
    if (b.two_side) wi.z = fabsf(wi.z);
    const float cos_theta_i = wi.z;
    active = active && cos_theta_i != 0.f;
    V3f m;
    ggx_sample_m(alpha, signbit_(cos_theta_i) ? -wi : wi, sample, m, bs.pdf);
    active = active && bs.pdf != 0.f;
    const float wi_m = dot(wi, m);
    const FresnelDielectric<float> f = fresnel_dielectric<float>(b.eta[0], wi_m);
    const bool selected_r = sample.z <= f.r && active, selected_t = !selected_r && active;
    bs.pdf *= selected_r ? f.r : 1.f - f.r;
    const float bs_eta = selected_r ? 1.f : f.eta_it;
    bs.wo = V3f(0.f, 0.f, 0.f);
    float dwh_dwo = 0.f;
    if (selected_r) {
        const float k = 2.f * wi_m;
        bs.wo = V3f(fmaf(m.x, k, -wi.x), fmaf(m.y, k, -wi.y), fmaf(m.z, k, -wi.z));
    }
    // the reflection Jacobian is evaluated for every lane (with wo = 0 where nothing was selected: rcp(0) = inf)
    dwh_dwo = 1.f / (4.f * dot(bs.wo, m));
    if (selected_t) {
        const float k = fmaf(wi_m, f.eta_ti, f.cos_theta_t);
        bs.wo = V3f(fmaf(m.x, k, -(wi.x * f.eta_ti)), fmaf(m.y, k, -(wi.y * f.eta_ti)), fmaf(m.z, k, -(wi.z * f.eta_ti)));
        const float wo_m = dot(bs.wo, m);
        dwh_dwo = (sqr(bs_eta) * wo_m) / sqr(wi_m + bs_eta * wo_m);
    }
    bs.pdf *= fabsf(dwh_dwo) * ggx_smith_g1<float>(alpha, bs.wo, m);
    bs.valid = active && (selected_t || selected_r);
    return bs;
}
template <class S, int kCfg> __device__ __forceinline__ BsdfSample bsdf_sample_leaf(const DScene &sc, const DBsdf &b, const Its<S> &its, V3f wi, V3f sample, bool active) {
    if ((kCfg & kCfgExt) && b.type == 3) return dielectric_sample(b, bsdf_alpha<S, kCfg>(sc, b, its), wi, sample, active);
    if ((kCfg & kCfgFull) && b.type != 0) return ggx_reflect_sample(bsdf_alpha<S, kCfg>(sc, b, its), b.two_side != 0, wi, sample, active);
    BsdfSample bs;
    float wiz = wi.z;
    if (b.two_side) wiz = fabsf(wiz);
    const V2f p = square_to_uniform_disk_concentric(V2f(sample.y, sample.z));
    const float z = safe_sqrt(1.f - fmaf(p.y, p.y, p.x * p.x));
    bs.wo = V3f(p.x, p.y, z);
    bs.pdf = kInvPi * z;
    bs.valid = active && (wiz > 0.f);
    return bs;
}
// NormalMap::__sample (normalmap.cpp:147-187): sample.z picks the facet the nested BSDF is sampled on
template <class S, int kCfg> PSDR_FULL_FN BsdfSample normalmap_sample(const DScene &sc, const DBsdf &b, const Its<S> &its, V3f wi, V3f sample, bool active) {
    if (b.two_side) wi.z = fabsf(wi.z);
    const DBsdf &nb = sc.bsdfs[b.nested];
    Its<float> f;
    f.tri = its.tri;
    f.uv = val(its.uv);
    f.bc = V2f(val(its.bc.x), val(its.bc.y));
    f.dp_du = val(its.dp_du);
    V3f wp;
    NmFrame<float> fr;
    nm_setup<float, kCfg>(b, f, wp, fr);
    const V3f p_wi = fr.to_local(wi);
    const float probability_wp = nm_lambda_p<float>(wp, wi);
    const bool itpo = sample.z >= probability_wp;
    // i -> t -> p -> o: the reflected incident direction (the reference leaves its.wi.z of this copy as passed in)
    const V3f r_wi = fr.to_local(nm_reflect<float>(wi, nm_wt<float>(wp)));
    BsdfSample bs = bsdf_sample_leaf<float, kCfg>(sc, nb, f, itpo ? r_wi : p_wi, sample, active);
    const float pdf1 = bsdf_pdf_leaf<float, kCfg>(sc, nb, f, p_wi, bs.wo), pdf2 = bsdf_pdf_leaf<float, kCfg>(sc, nb, f, r_wi, bs.wo);
    bs.pdf = probability_wp * pdf1 + (1.f - probability_wp) * pdf2;
    bs.wo = fr.to_world(bs.wo);
    return bs;
}

template <class S, int kCfg> __device__ __forceinline__ BsdfSample bsdf_sample(const DScene &sc, const Its<S> &its, V3f sample, bool active) {
    BsdfSample bs;
    bs.wo = V3f(0.f, 0.f, 0.f);
    bs.pdf = 0.f;
    bs.valid = false;
    if (!its.valid) return bs;
    const int bi = sc.meshes[its.mesh].bsdf;
    if (bi < 0) return bs;
    const DBsdf &b = sc.bsdfs[bi];
    if ((kCfg & kCfgExt) && b.type == 5) return normalmap_sample<S, kCfg>(sc, b, its, val(its.wi), sample, active);
    return bsdf_sample_leaf<S, kCfg>(sc, b, its, val(its.wi), sample, active);
}

// ---- emitters (reference src/emitter/area.cpp, src/shape/mesh.cpp:413-466) --------------------
template <class S> __device__ __forceinline__ V3<S> emitter_radiance(const DEmitter &e);
template <> __device__ __forceinline__ V3f emitter_radiance<float>(const DEmitter &e) { return V3f(e.radiance[0], e.radiance[1], e.radiance[2]); }
template <> __device__ __forceinline__ V3d emitter_radiance<Dual>(const DEmitter &e) {
    return V3d(Dual(e.radiance[0], e.d_radiance[0]), Dual(e.radiance[1], e.d_radiance[1]), Dual(e.radiance[2], e.d_radiance[2]));
}
template <class S> __device__ __forceinline__ bool is_emitter(const DScene &sc, const Its<S> &its) {
    return its.valid && sc.meshes[its.mesh].emitter >= 0;
}
// EnvironmentMap::eval_direction (reference src/emitter/envmap.cpp:56-73)
template <class S> __device__ __forceinline__ V3<S> env_eval_direction(const DEnv &e, V3<S> wi, EnvTexelTaps *taps = nullptr) {
    const V3<S> v = mul3x3<S>(e.from_world, e.d_from_world, wi);
    const V2<S> uv = envmap_dir_to_uv<S>(v);
    const V3<S> r = bitmap_eval_envmap<S>(e.data, IsDual<S>::value ? e.ddata : nullptr, e.w, e.h, uv, taps);
    return r * Lift<S>::s(e.scale, e.d_scale);
}
template <class S, int kCfg> __device__ __forceinline__ V3<S> Le(const DScene &sc, const Its<S> &its, bool active) {
    if (!its.valid) return V3<S>(S(0.f));
    const int e = sc.meshes[its.mesh].emitter;
    if (e < 0) return V3<S>(S(0.f));
    if ((kCfg & kCfgFull) && sc.emitters[e].type == 1) {   // EnvironmentMap::eval: radiance arriving along -wi, no cosine test
        if (!active) return V3<S>(S(0.f));
        return env_eval_direction<S>(sc.env, -its.to_world(its.wi));
    }
    if (!(active && val(its.wi.z) > 0.f)) return V3<S>(S(0.f));
    return emitter_radiance<S>(sc.emitters[e]);
}

// DiscreteDistribution::sample_reuse (reference src/core/pmf.cpp:31-51): binary search of the
// unnormalised fp32 CDF, sample re-stretched to [0,1]
__device__ __forceinline__ int sample_reuse(const float *pmf, const float *cmf, int size, float sum, float &s, float &prob) {
    if (size == 1) { prob = 1.f; return 0; }
    s *= sum;
    int start = 0, end = size - 1;
    const int iterations = 32 - __clz(end - start);   // log2i(end-start)+1, end-start >= 1
    for (int i = 0; i < iterations; ++i) {
        const int middle = (start + end) >> 1;
        if (__ldg(cmf + middle) < s) start = min(middle + 1, end);
        else end = middle;
    }
    const int idx = start;
    if (idx > 0) s -= __ldg(cmf + idx - 1);
    const float p = __ldg(pmf + idx);
    if (p > 0.f) s /= p;
    s = fminf(fmaxf(s, 0.f), 1.f);
    prob = p / sum;
    return idx;
}

// The same search narrowed by a bucket table: lut[b] = first index whose CDF value is >= b * sum / n (n buckets, n + 1
// entries, built in double precision by device_upload.cu build_cdf_lut), so the answer for a scaled sample in bucket b
// lies in [lut[b - 1], lut[b + 2]] (one bucket of slack on either side absorbs the fp32 rounding of the bucket index)
// and the bisection -- the same lower bound, hence the same index as sample_reuse -- takes 1-2 dependent loads instead
// of log2(size): the 50 000-cell guiding grid and the 7 000-edge distribution of cfg 4 cost 16 and 13 (profiles/r02w:
// a quarter of the secondary-edge kernel's stall samples sat on the search's load).
__device__ __forceinline__ int sample_reuse_lut(const float *pmf, const float *cmf, int size, float sum, const int *lut, int n, float &s, float &prob) {
    if (size == 1) { prob = 1.f; return 0; }
    if (n <= 0) return sample_reuse(pmf, cmf, size, sum, s, prob);
    s *= sum;
    int b = (int) (s * ((float) n / sum));
    b = min(max(b, 0), n - 1);
    int start = __ldg(lut + max(b - 1, 0)), end = __ldg(lut + min(b + 2, n));
    while (start < end) {
        const int middle = (start + end) >> 1;
        if (__ldg(cmf + middle) < s) start = middle + 1;
        else end = middle;
    }
    const int idx = start;
    if (idx > 0) s -= __ldg(cmf + idx - 1);
    const float p = __ldg(pmf + idx);
    if (p > 0.f) s /= p;
    s = fminf(fmaxf(s, 0.f), 1.f);
    prob = p / sum;
    return idx;
}

template <class S> struct PosSample {
    V3<S> p, n;
    S J;
    float pdf;
    int tri;     // global id of the sampled emitter triangle
    V2f st;      // barycentrics of the sample on it
};

// EnvironmentMap::sample_direction + __sample_position (reference src/emitter/envmap.cpp:86-129) and
// ray_intersect_scene_aabb (include/psdr/utils.h:144-164): a direction drawn from the cell distribution,
// turned into the point where it leaves the scene bounding box.  Everything is detached.
static PSDR_FULL_FN void env_sample_position(const DEnv &e, V3f ref_p, V2f sample2, V3f &p, V3f &n, float &pdf_out) {
    const int ncells = e.cw * e.ch;
    float prob;
    const int idx = sample_reuse_lut(e.cell_pmf, e.cell_cmf, ncells, e.cell_sum, e.cell_lut, e.cell_lut_n, sample2.y, prob);   // HyperCube<2>: last dimension
    const int cx = idx / e.ch, cy = idx - cx * e.ch;
    const float u = (sample2.x + (float) cx) * (1.f / (float) e.cw), v = (sample2.y + (float) cy) * (1.f / (float) e.ch);
    float pdf = prob * (float) ncells;
    float st, ct, sp, cp;
    sincos_full(v * kPi, st, ct);
    sincos_full(u * (2.f * kPi), sp, cp);
    V3f d(sp * st, ct, -(cp * st));                                   // sphdir(theta, phi) -> (y, z, -x)
    const float inv_sin_theta = 1.f / sqrtf(fmaxf(sqr(d.x) + sqr(d.z), sqr(kEpsilon)));
    if (pdf > kEpsilon) pdf *= inv_sin_theta * (.5f / sqr(kPi));
    d = mul3x3<float>(e.to_world, nullptr, d);
    const float dd[3] = {d.x, d.y, d.z}, oo[3] = {ref_p.x, ref_p.y, ref_p.z};
    float t = 0.f;
    int axis = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float t1 = (e.lower[i] - oo[i]) / dd[i], t2 = (e.upper[i] - oo[i]) / dd[i];
        const float tm = fmaxf(t1, t2);
        if (i == 0 || tm < t) { t = tm; axis = i; }
    }
    float nn[3] = {0.f, 0.f, 0.f};
    nn[axis] = signbit_(dd[axis]) ? 1.f : -1.f;                       // -sign(d[axis])
    n = V3f(nn[0], nn[1], nn[2]);
    const float G = dot(n, -d) * (1.f / sqr(t));
    p = V3f(fmaf(d.x, t, ref_p.x), fmaf(d.y, t, ref_p.y), fmaf(d.z, t, ref_p.z));
    pdf_out = pdf * G;
}

// EnvironmentMap::__sample_position_pdf (reference src/emitter/envmap.cpp:142-162) + HyperCubeDistribution::pdf
template <class S> __device__ __forceinline__ float env_position_pdf(const DEnv &e, V3f ref_p, const Its<S> &its) {
    V3f d = val(its.p) - ref_p;
    const float dist2 = squared_norm(d);
    d = d / safe_sqrt(dist2);
    const float G = fabsf(dot(d, val(its.n))) / dist2;
    d = mul3x3<float>(e.from_world, nullptr, d);
    const float factor = G * (1.f / sqrtf(fmaxf(sqr(d.x) + sqr(d.z), sqr(kEpsilon)))) * (.5f / sqr(kPi));
    const V2f uv = envmap_dir_to_uv<float>(d);
    const int ix = (int) floorf(uv.x * (float) e.cw), iy = (int) floorf(uv.y * (float) e.ch);
    if (!(ix >= 0 && ix < e.cw && iy >= 0 && iy < e.ch)) return 0.f;
    const int idx = ix * e.ch + iy;
    return (__ldg(e.cell_pmf + idx) / e.cell_sum) * (float) (e.cw * e.ch) * factor;
}

// Scene::sample_emitter_position (reference src/scene/scene.cpp:987-1013) -> Mesh::sample_position / EnvironmentMap
template <class S, int kCfg> __device__ __forceinline__ PosSample<S> sample_emitter_position(const DScene &sc, V3f ref_p, V2f sample2) {
    PosSample<S> ps;
    int ei = 0;
    float emitter_pdf = 1.f;
    if (sc.n_emitters != 1) ei = sample_reuse(sc.emitter_pmf, sc.emitter_cmf, sc.n_emitters, sc.emitter_sum, sample2.y, emitter_pdf);
    const DEmitter &em = sc.emitters[ei];
    if ((kCfg & kCfgFull) && em.type == 1) {
        V3f p, n;
        float pdf;
        env_sample_position(sc.env, ref_p, sample2, p, n, pdf);
        ps.p = lift3<S>(p);
        ps.n = lift3<S>(n);
        ps.J = S(1.f);
        ps.pdf = pdf;
        if (sc.n_emitters != 1) ps.pdf *= emitter_pdf;
        ps.tri = -1;
        ps.st = V2f(0.f, 0.f);
        return ps;
    }
    float dummy;
    const int fi = sample_reuse(sc.face_pmf + em.distrb_offset, sc.face_cmf + em.distrb_offset, em.nfaces, em.face_sum, sample2.x, dummy);
    const float t = safe_sqrt(1.f - sample2.x);
    const V2f st(1.f - t, t * sample2.y);
    const TriRec<S> T = load_tri<S>(sc, em.face_offset + fi);
    ps.tri = em.face_offset + fi;
    ps.st = st;
    ps.J = S(1.f);
    if (IsDual<S>::value) ps.J = T.area / detach(T.area);
    ps.p = bilinear(T.p0, T.e1, T.e2, st);
    const float4 c = __ldg(sc.shade + 3 * (em.face_offset + fi) + 2);
    if (IsDual<S>::value) {
        const float4 dc = __ldg(sc.dshade + 3 * (em.face_offset + fi) + 2);
        ps.n = Lift<S>::v3(V3f(c.y, c.z, c.w), V3f(dc.y, dc.z, dc.w));
    } else {
        ps.n = Lift<S>::v3(V3f(c.y, c.z, c.w), V3f(0.f, 0.f, 0.f));
    }
    ps.pdf = em.inv_total_area;
    if (sc.n_emitters != 1) ps.pdf *= emitter_pdf;
    return ps;
}

template <class S, int kCfg> __device__ __forceinline__ float emitter_position_pdf(const DScene &sc, V3f ref_p, const Its<S> &its, bool active) {
    if (!its.valid || !active) return 0.f;
    const int e = sc.meshes[its.mesh].emitter;
    if (e < 0) return 0.f;
    if ((kCfg & kCfgFull) && sc.emitters[e].type == 1) return env_position_pdf<S>(sc.env, ref_p, its);
    return sc.emitters[e].sampling_weight * sc.emitters[e].inv_total_area;
}

__device__ __forceinline__ float mis_weight(float a, float b) {
    const float w1 = a * a, w2 = b * b;
    return w1 / (w1 + w2);
}

// ---- camera (reference src/sensor/perspective.cpp:160-197) ------------------------------------
__device__ __forceinline__ V3f xform_pos(const float *M, V3f p) {
    float t[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) t[i] = fmaf(M[4 * i + 2], p.z, fmaf(M[4 * i + 1], p.y, M[4 * i] * p.x)) + M[4 * i + 3];
    return V3f(t[0], t[1], t[2]) / t[3];
}
__device__ __forceinline__ V3f xform_dir(const float *M, V3f p) {
    return V3f(fmaf(M[2], p.z, fmaf(M[1], p.y, M[0] * p.x)), fmaf(M[6], p.z, fmaf(M[5], p.y, M[4] * p.x)),
               fmaf(M[10], p.z, fmaf(M[9], p.y, M[8] * p.x)));
}
__device__ __forceinline__ V3d xform_pos_d(const float *M, const float *dM, V3d p) {
    Dual t[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const Dual a(M[4 * i], dM[4 * i]), b(M[4 * i + 1], dM[4 * i + 1]), c(M[4 * i + 2], dM[4 * i + 2]), w(M[4 * i + 3], dM[4 * i + 3]);
        t[i] = fmadd(c, p.z, fmadd(b, p.y, a * p.x)) + w;
    }
    return V3d(t[0], t[1], t[2]) / t[3];
}
__device__ __forceinline__ V3d xform_dir_d(const float *M, const float *dM, V3d p) {
    Dual t[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const Dual a(M[4 * i], dM[4 * i]), b(M[4 * i + 1], dM[4 * i + 1]), c(M[4 * i + 2], dM[4 * i + 2]);
        t[i] = fmadd(c, p.z, fmadd(b, p.y, a * p.x));
    }
    return V3d(t[0], t[1], t[2]);
}

template <class S> __device__ __forceinline__ void sample_primary_ray(const DCamera &cam, V2f s, V3<S> &o, V3<S> &d);
// the ray in CAMERA space: perspective -- origin 0, direction through the sample on the near plane; orthographic
// (reference src/sensor/orthographic.cpp:109-131) -- origin = the sample on the near plane, direction +z
__device__ __forceinline__ void camera_ray_local(const DCamera &cam, V2f s, V3f &oc, V3f &dc) {
    const V3f q = xform_pos(cam.sample_to_camera, V3f(s.x, s.y, 0.f));
    if (cam.ortho) { oc = q; dc = V3f(0.f, 0.f, 1.f); }
    else { oc = V3f(0.f, 0.f, 0.f); dc = normalize(q); }
}
template <> __device__ __forceinline__ void sample_primary_ray<float>(const DCamera &cam, V2f s, V3f &o, V3f &d) {
    V3f oc, dc;
    camera_ray_local(cam, s, oc, dc);
    o = xform_pos(cam.to_world, oc);
    d = xform_dir(cam.to_world, dc);
}
template <> __device__ __forceinline__ void sample_primary_ray<Dual>(const DCamera &cam, V2f s, V3d &o, V3d &d) {
    V3f oc, dc;
    camera_ray_local(cam, s, oc, dc);                                               // detached in camera space
    o = xform_pos_d(cam.to_world, cam.d_to_world, lift3<Dual>(oc));
    d = xform_dir_d(cam.to_world, cam.d_to_world, lift3<Dual>(dc));
}

// ---- PathTracer::__Li (reference src/integrator/path.cpp:35-127) ------------------------------
// NEE + BSDF sampling with the power heuristic, fixed max_depth, no Russian roulette.  Every lane
// draws 5 numbers per bounce whether it is alive or not, so draw k of a lane is a closed-form
// function of (seed, lane, k).
// The reference unrolls "primary hit; for each bounce {NEE; BSDF ray}" into one megakernel; here the
// same arithmetic is rolled into ONE loop whose iteration is {main ray (camera ray or the BSDF ray of
// the previous bounce) -> vertex; NEE shadow ray}, so the kernel holds exactly two copies of the
// closest-hit scan and stays inside the instruction cache (profiles/r01a: the unrolled form stalled
// 49 % of the cycles on instruction fetch).
// Rec = recorder policy: the adjoint's primal replay passes a PathRecord (adjoint.cuh) that keeps what the
// reverse sweep needs (vertices, light samples, detached pdfs/MIS weights, throughputs); NoRecord
// compiles to nothing.
struct NoRecord {
    __device__ __forceinline__ void vertex(int, int, float, float) {}
    __device__ __forceinline__ void throughput(int, V3f) {}
    __device__ __forceinline__ void bounce(int, bool, float, float) {}
    __device__ __forceinline__ void nee(int, bool, int, V2f, V3f, int, float, float) {}
};

// The loop body is exposed as a state machine (li_begin / li_step) so that the persistent kernels can let every
// lane walk its OWN sequence of paths: a lane whose path ends starts its next one in the same warp iteration
// instead of idling until the slowest lane of the warp is done ("lane regeneration").
template <class S> struct LiState {
    V3<S> throughput, result, ray_o, ray_d;
    Its<S> its;
    BsdfSample bs;
    int depth;
    bool active;
};

template <class S> __device__ __forceinline__ void li_begin(LiState<S> &st, V3<S> ro, V3<S> rd, bool active) {
    st.throughput = V3<S>(S(1.f));
    st.result = V3<S>(S(0.f));
    st.its.valid = false;
    st.bs.wo = V3f(0.f, 0.f, 1.f);
    st.bs.pdf = 1.f;
    st.bs.valid = true;
    st.ray_o = ro;
    st.ray_d = rd;
    st.depth = -1;
    st.active = active;
}

// one iteration {main ray -> vertex, NEE shadow ray, BSDF sample}; returns true when the path is finished
template <class S, int kCfg, bool kAD, class Rec>
__device__ __forceinline__ bool li_step(const DScene &sc, Pcg32 &rng, LiState<S> &st, int max_depth, bool hide_emitters, Rec &R, int mis = 2) {
    constexpr bool ad = kAD;
    const int depth = st.depth;
    Its<S> &its = st.its;
    // ---- main ray: primary hit (solid-angle form under AD) or the BSDF-sampled ray (path-space form)
    const Its<S> its1 = ray_intersect<S, kCfg, kAD>(sc, st.ray_o, st.ray_d, st.active, ad && depth >= 0);
    if (its1.valid) R.vertex(depth + 1, its1.tri, its1.bu, its1.bv);
    if (depth < 0) {
        st.active = st.active && its1.valid;
        if (!hide_emitters) st.result = Le<S, kCfg>(sc, its1, st.active);
    } else {
        st.active = st.active && st.bs.valid && its1.valid;
        if (st.active) {
            V3<S> bsdf_val;
            float pdf0;
            if (ad) {
                V3<S> wo = its1.p - its.p;
                wo = wo / its1.t;
                const S cos_val = dot(its1.n, -wo);
                const S G_val = abs_(cos_val) / sqr(its1.t);
                pdf0 = st.bs.pdf * val(G_val);
                if (val(its1.t) < kEpsilon) bsdf_val = V3<S>(S(0.f));
                else bsdf_val = bsdf_eval<S, kCfg>(sc, its, its.to_local(wo), st.active) * (G_val * its1.J / S(pdf0));
            } else {
                const S cos_val = dot(its1.n, -st.ray_d);
                const S G_val = abs_(cos_val) / sqr(its1.t);
                pdf0 = st.bs.pdf * val(G_val);
                if (val(its1.t) < kEpsilon) bsdf_val = V3<S>(S(0.f));
                else bsdf_val = bsdf_eval<S, kCfg>(sc, its, lift3<S>(st.bs.wo), st.active) / S(st.bs.pdf);
            }
            // DirectIntegrator(mis = 1): BSDF sampling only, weight 1 (reference src/integrator/direct.cpp:121-124)
            const float weight2 = mis == 1 ? 1.f : mis_weight(pdf0, emitter_position_pdf<S, kCfg>(sc, val(its.p), its1, st.active));
            R.bounce(depth, val(its1.t) >= kEpsilon, pdf0, weight2);
            st.throughput = st.throughput * bsdf_val;
            st.result = st.result + Le<S, kCfg>(sc, its1, st.active) * st.throughput * S(weight2);
        }
    }
    const int bounces_left = max_depth - depth - 1;
    // draws per bounce: next_2d (emitter sampling) + next_nd<3> (BSDF sampling); the Direct integrator's one-strategy
    // modes draw only what they use (mis = 0: emitter sampling only, mis = 1: BSDF sampling only; direct.cpp:47,84-91)
    const unsigned long long per_bounce = mis == 0 ? 2ull : mis == 1 ? 3ull : 5ull;
    if (!st.active) {   // dead lanes only burn their draws
        if (bounces_left > 0) rng.advance(per_bounce * (unsigned long long) bounces_left);
        return true;
    }
    if (bounces_left <= 0) return true;
    its = its1;
    R.throughput(depth + 1, val(st.throughput));
    float s_y = 0.f, s_x = 0.f, s3_z = 0.f, s3_y = 0.f, s3_x = 0.f;
    if (mis != 1) { s_y = rng.next_1d(); s_x = rng.next_1d(); }                        // next_2d: y first
    if (mis != 0) { s3_z = rng.next_1d(); s3_y = rng.next_1d(); s3_x = rng.next_1d(); }   // next_nd<3> = (d3,d2,d1)
    if (mis != 1) {   // ---- emitter sampling
        const PosSample<S> ps = sample_emitter_position<S, kCfg>(sc, val(its.p), V2f(s_x, s_y));
        bool active_direct = !is_emitter(sc, its);
        V3<S> wod = ps.p - its.p;
        const S dist_sqr = squared_norm(wod);
        const S dist = safe_sqrt(dist_sqr);
        wod = wod / dist;
        const Its<S> its2 = ray_intersect<S, kCfg, kAD>(sc, its.p, wod, active_direct, ad);
        active_direct = active_direct && its2.valid;
        active_direct = active_direct && (val(its2.t) > val(dist) - kShadowEpsilon) && is_emitter(sc, its2);
        if (active_direct) {
            const S cos_val = dot(its2.n, -wod);
            const S G_val = abs_(cos_val) / dist_sqr;
            const V3<S> emitter_val = Le<S, kCfg>(sc, its2, true);
            const V3<S> wo_local = its.to_local(wod);
            V3<S> bsdf_val2 = bsdf_eval<S, kCfg>(sc, its, wo_local, active_direct);
            bsdf_val2 = bsdf_val2 * (G_val * ps.J / S(ps.pdf));
            const float pdf1 = bsdf_pdf<S, kCfg>(sc, its, wo_local, active_direct) * val(G_val);
            if (pdf1 != 0.f) {
                const float weight1 = mis == 0 ? 1.f : mis_weight(ps.pdf, pdf1);
                R.nee(depth + 1, ps.tri < 0 || val(its2.wi.z) > 0.f, ps.tri, ps.st, val(ps.p), its2.tri, ps.pdf, weight1);
                st.result = st.result + st.throughput * emitter_val * bsdf_val2 * S(weight1);
            }
        }
    }
    if (mis == 0) return true;      // emitter sampling only: no BSDF-sampled ray
    // ---- BSDF sampling: the ray is traced by the next step
    st.bs = bsdf_sample<S, kCfg>(sc, its, V3f(s3_x, s3_y, s3_z), true);
    st.ray_o = its.p;
    st.ray_d = its.to_world(lift3<S>(st.bs.wo));
    st.depth = depth + 1;
    return false;
}

// `lanes` != 0: the lanes of the warp that call Li together (converged); the step loop is then warp-uniform -- every
// lane stays in it until the longest path of the warp is finished and the vote re-converges the warp once per step.
// The adjoint's replay uses it (its backward sweep wants the warp whole: 14.4 -> 13.0 ms); the forward kernels run
// the plain loop (`lanes` = 0) and re-converge at the caller's barrier, which measured 2-7 % faster for them.
template <class S, int kCfg, bool kAD, class Rec>
__device__ __forceinline__ V3<S> Li(const DScene &sc, Pcg32 &rng, V3<S> ro, V3<S> rd, bool active, int max_depth, bool hide_emitters, Rec &R,
                                    unsigned lanes = 0u, int mis = 2, bool cta = false) {
    // `cta`: every thread of the CTA calls Li together and the step loop is CTA-uniform (one block barrier per step), so
    // that all warps of the CTA execute the same stretch of code at the same time (instruction-cache locality)
    LiState<S> st;
    li_begin<S>(st, ro, rd, active);
    if (lanes != 0u) {
        bool done = false;
#pragma unroll 1
        while (true) {
            if (!done) done = li_step<S, kCfg, kAD, Rec>(sc, rng, st, max_depth, hide_emitters, R, mis);
            if (cta ? __syncthreads_and(done) != 0 : __all_sync(lanes, done)) break;
        }
    } else {
#pragma unroll 1
        while (!li_step<S, kCfg, kAD, Rec>(sc, rng, st, max_depth, hide_emitters, R, mis)) {}
    }
    return st.result;
}

template <class S, int kCfg>
__device__ __forceinline__ V3<S> Li(const DScene &sc, Pcg32 &rng, V3<S> ro, V3<S> rd, bool active, int max_depth, bool hide_emitters, int mis = 2,
                                    bool cta = false) {
    NoRecord rec;
    return Li<S, kCfg, IsDual<S>::value, NoRecord>(sc, rng, ro, rd, active, max_depth, hide_emitters, rec, cta ? 0xffffffffu : 0u, mis, cta);
}

// field ids of psdr_render_field_edges
__device__ __forceinline__ V3f field_value(const Its<float> &its, int field, int object) {
    if (!its.valid || (object >= 0 && its.mesh != object)) return V3f(0.f, 0.f, 0.f);
    switch (field) {
        case 0: return V3f((float) its.mesh, (float) its.mesh, (float) its.mesh);      // segmentation
        case 1: return V3f(1.f, 1.f, 1.f);                                                // silhouette
        case 2: return its.p;                                                             // position
        case 3: return V3f(its.t, its.t, its.t);                                          // depth
        case 4: return its.n;                                                             // geoNormal
        case 5: return its.sh_n;                                                          // shNormal
        default: return V3f(its.uv.x, its.uv.y, 0.f);                                     // uv
    }
}

// CollocatedIntegrator::__Li (reference src/integrator/collocated.cpp:21-53): a point light at the camera -- the BSDF towards
// the viewer for light arriving from the viewer, times intensity / t^2; no emitters, no sampling, no random numbers.
// Kept out of li_step (its own small kernel instantiations, kernels_impl.cuh kColloc): the path tracer's kernels do not
// carry a second inlined copy of the BSDF code for it.
template <class S, int kCfg, bool kAD>
__device__ __forceinline__ V3<S> Li_collocated(const DScene &sc, V3<S> ro, V3<S> rd, bool active) {
    const Its<S> its = ray_intersect<S, kCfg, kAD>(sc, ro, rd, active, false);
    if (!(active && its.valid)) return V3<S>(S(0.f));
    const V3<S> f = bsdf_eval<S, kCfg>(sc, its, its.wi, true);
    if (sc.colloc_field) return f;      // FieldExtractionIntegrator("bsdf") (reference src/integrator/field.cpp:72-92)
    return f / sqr(its.t) * bsdf_scalar<S>(sc.colloc_intensity, sc.d_colloc_intensity);
}

// ---- secondary (shadow) edges: reference src/scene/scene.cpp:1027-1068, src/integrator/path.cpp:172-270
struct SensorDirect {
    V2f q;
    int pixel;
    float sensor_val;
    bool valid;
};
__device__ __forceinline__ SensorDirect sample_direct(const DScene &sc, const DCamera &cam, V3f p) {
    SensorDirect r;
    const V3f q = xform_pos(cam.world_to_sample, p);
    r.q = V2f(q.x, q.y);
    const int ix = (int) floorf(q.x * (float) sc.width), iy = (int) floorf(q.y * (float) sc.height);
    r.valid = ix >= 0 && ix < sc.width && iy >= 0 && iy < sc.height;
    r.pixel = r.valid ? iy * sc.width + ix : -1;
    V3f dir = p - V3f(cam.pos[0], cam.pos[1], cam.pos[2]);
    const float dist2 = squared_norm(dir);
    dir = dir / safe_sqrt(dist2);
    const float cosTheta = dot(V3f(cam.dir[0], cam.dir[1], cam.dir[2]), dir);
    const float ic = 1.f / cosTheta;
    r.sensor_val = (1.f / dist2) * (ic * ic * ic) * cam.inv_area;
    return r;
}

__device__ __forceinline__ int sign_eps(float x, float eps) { return x > eps ? 1 : (x < -eps ? -1 : 0); }
__device__ __forceinline__ float sign1(float x) { return signbit_(x) ? -1.f : 1.f; }

// returns the pixel (-1: no contribution); value0 = primal boundary value (guiding pre-pass),
// tangent = d/dP of the zero-primal estimator.
// Adj = reverse-mode policy of the secondary-edge estimator: NoSecAdjoint (forward mode) or
// SecEdgeAdjoint (adjoint.cuh), whose tail() scatters the gradients instead of forming the tangent.
struct NoSecAdjoint {
    static constexpr bool enabled = false;
    template <int kCfg>
    __device__ __forceinline__ void tail(const DScene &, const DCamera &, int, V3f, V3f, int, float, V3d, int, const Its<Dual> &, V3f, V2f) const {}
};

// The estimator runs in two stages so that a kernel can batch the survivors of the first one: stage 0 (edge and
// light-point sample, the orientation test: no ray) passes ~30 % of the samples, stage 1 (three closest-hit scans)
// is where the time goes.  A candidate carries what stage 1 cannot recompute from (edge, position on it).
struct SecCand {
    int ei;              // secondary edge
    float sample1;       // position on it (re-stretched first sample)
    V3f p2, bn;          // sampled light point and its normal
    float bss_pdf;
};
struct SecEdgeGeo {      // the sampled boundary point with its tangent, and the edge frame
    V3d bp0;
    V3f edge, edge2;
};
__device__ __forceinline__ SecEdgeGeo sec_edge_geo(const DScene &sc, int ei, float sample1, V3f &n0, V3f &n1, bool &is_boundary, float &e1_norm) {
    const float4 q0 = __ldg(sc.sec_edges + 6 * ei), q1 = __ldg(sc.sec_edges + 6 * ei + 1), q2 = __ldg(sc.sec_edges + 6 * ei + 2),
                 q3 = __ldg(sc.sec_edges + 6 * ei + 3), q4 = __ldg(sc.sec_edges + 6 * ei + 4), q5 = __ldg(sc.sec_edges + 6 * ei + 5);
    const V3d ep0(Dual(q0.x, q1.z), Dual(q0.y, q1.w), Dual(q0.z, q2.x));
    const V3d ee1(Dual(q0.w, q2.y), Dual(q1.x, q2.z), Dual(q1.y, q2.w));
    n0 = V3f(q3.x, q3.y, q3.z);
    n1 = V3f(q3.w, q4.x, q4.y);
    const V3f ep2(q4.z, q4.w, q5.x);
    is_boundary = q5.y != 0.f;
    SecEdgeGeo g;
    g.bp0 = V3d(fmadd(ee1.x, sample1, ep0.x), fmadd(ee1.y, sample1, ep0.y), fmadd(ee1.z, sample1, ep0.z));
    const V3f e1v = val(ee1);
    g.edge = normalize(e1v);
    g.edge2 = ep2 - val(ep0);
    e1_norm = norm(e1v);
    return g;
}

// -- sample_boundary_segment_direct
template <int kCfg>
__device__ __forceinline__ bool sec_edge_stage0(const DScene &sc, V3f sample3, SecCand &c) {
    float sample1 = sample3.x, pdf0;
    const int ei = sample_reuse_lut(sc.sec_pmf, sc.sec_cmf, sc.n_sec_edges, sc.sec_sum, sc.sec_lut, sc.sec_lut_n, sample1, pdf0);
    V3f n0, n1;
    bool is_boundary;
    float e1_norm;
    const SecEdgeGeo g = sec_edge_geo(sc, ei, sample1, n0, n1, is_boundary, e1_norm);
    const V3f _p0 = val(g.bp0);
    pdf0 /= e1_norm;
    const PosSample<float> ps2 = sample_emitter_position<float, kCfg>(sc, _p0, V2f(sample3.y, sample3.z));
    const V3f _p2 = ps2.p, bn = ps2.n;
    V3f e = _p2 - _p0;
    const float distSqr = squared_norm(e);
    e = e / safe_sqrt(distSqr);
    const float cosTheta = dot(bn, -e);
    const int sgn0 = sign_eps(dot(n0, e), kEdgeEpsilon), sgn1 = sign_eps(dot(n1, e), kEdgeEpsilon);
    const bool valid = (cosTheta > kEpsilon) && ((is_boundary && sgn0 != 0) || (!is_boundary && sgn0 * sgn1 < 0));
    c.ei = ei;
    c.sample1 = sample1;
    c.p2 = _p2;
    c.bn = bn;
    c.bss_pdf = pdf0 * ps2.pdf * (distSqr / cosTheta);
    return valid;
}

// -- eval_secondary_edge
template <int kCfg, class Adj>
__device__ __forceinline__ int sec_edge_stage1(const DScene &sc, const DCamera &cam, const SecCand &c, V3f &value0_out, V3f &tangent_out,
                                               const Adj &adj) {
    value0_out = V3f(0.f, 0.f, 0.f);
    tangent_out = V3f(0.f, 0.f, 0.f);
    const int ei = c.ei;
    const float sample1 = c.sample1;
    V3f n0_, n1_;
    bool is_boundary_;
    float e1_norm_;
    const SecEdgeGeo g = sec_edge_geo(sc, ei, sample1, n0_, n1_, is_boundary_, e1_norm_);
    const V3d bp0 = g.bp0;
    const V3f edge = g.edge, edge2 = g.edge2, _p0 = val(g.bp0), _p2 = c.p2, bn = c.bn;
    const float bss_pdf = c.bss_pdf;
    bool valid = true;
    const V3f _dir = normalize(_p2 - _p0);
    int light_tri = -1;
    const Its<float> _its2 = ray_intersect<float, kCfg>(sc, _p0, _dir, valid, false, &light_tri);
    valid = valid && is_emitter(sc, _its2) && _its2.valid && norm(_its2.p - _p2) < kShadowEpsilon;
    if (!valid) return -1;
    const Its<float> _its1 = ray_intersect<float, kCfg>(sc, _p0, -_dir, valid, false);
    valid = valid && _its1.valid;
    if (!valid) return -1;
    const V3f _p1 = _its1.p;
    const SensorDirect sds = sample_direct(sc, cam, _p1);
    valid = valid && sds.valid;
    if (!valid) return -1;
    V3d co, cd;
    sample_primary_ray<Dual>(cam, sds.q, co, cd);
    const Its<Dual> its1 = ray_intersect<Dual, kCfg>(sc, co, cd, valid, false);
    valid = valid && its1.valid && norm(val(its1.p) - _p1) < kShadowEpsilon;
    valid = valid && its1.valid && sc.meshes[its1.mesh].bsdf >= 0;
    if (!valid) return -1;
    const float dist = norm(_p2 - _p1), cos2 = fabsf(dot(bn, -_dir));
    const V3f ec = cross(edge, _dir);
    const float sinphi = norm(ec);
    const V3f proj = normalize(cross(ec, bn));
    const float sinphi2 = norm(cross(_dir, proj));
    const float base_v = (_its1.t / dist) * (sinphi / sinphi2) * cos2;
    valid = valid && (sinphi > kEpsilon) && (sinphi2 > kEpsilon);
    if (!valid) return -1;
    const V3f d0 = -val(cd);
    const V3f d0_local = _its1.to_local(d0);
    V3f bsdf_val = bsdf_eval<float, kCfg>(sc, _its1, d0_local, valid);
    const float correction = fabsf((_its1.wi.z * dot(d0, _its1.n)) / (d0_local.z * dot(_dir, _its1.n)));
    bsdf_val = bsdf_val * correction;
    V3f value0 = bsdf_val * Le<float, kCfg>(sc, _its2, valid) * (base_v * sds.sensor_val / bss_pdf);
    value0_out = value0;
    const V3f n = normalize(cross(bn, proj));
    value0 = value0 * (sign1(dot(ec, edge2)) * sign1(dot(ec, n)));
    if (Adj::enabled) {
        adj.template tail<kCfg>(sc, cam, sds.pixel, value0, n, ei, sample1, bp0, light_tri, its1, val(cd), sds.q);
        return sds.pixel;
    }
    const TriRec<Dual> T = load_tri<Dual>(sc, light_tri);
    const V3d sdir = normalize(bp0 - its1.p);
    Dual u, v, t;
    ray_intersect_triangle<Dual>(T.p0, T.e1, T.e2, its1.p, sdir, u, v, t);
    const V3d u2 = bilinear(detach(T.p0), detach(T.e1), detach(T.e2), V2d(u, v));
    const Dual dn = dot(lift3<Dual>(n), u2);
    tangent_out = V3f(value0.x * dn.d, value0.y * dn.d, value0.z * dn.d);
    return sds.pixel;
}
// both stages back to back (guiding pre-pass: one thread per grid cell)
template <int kCfg, class Adj>
__device__ __forceinline__ int eval_secondary_edge(const DScene &sc, const DCamera &cam, V3f sample3, V3f &value0_out, V3f &tangent_out,
                                                   const Adj &adj) {
    value0_out = V3f(0.f, 0.f, 0.f);
    tangent_out = V3f(0.f, 0.f, 0.f);
    SecCand c;
    if (!sec_edge_stage0<kCfg>(sc, sample3, c)) return -1;
    return sec_edge_stage1<kCfg, Adj>(sc, cam, c, value0_out, tangent_out, adj);
}

// HyperCubeDistribution<3>::sample_reuse (reference src/core/cube_distrb.cpp:41-48): the cell is picked with the
// LAST sample dimension; the sample becomes (cell + sample) * unit; returns pdf = pmf * num_cells
__device__ __forceinline__ float guide_sample_reuse(const DCamera &cam, V3f &s) {
    const int ncells = cam.greso[0] * cam.greso[1] * cam.greso[2];
    float prob;
    const int idx = sample_reuse_lut(cam.guide_pmf, cam.guide_cmf, ncells, cam.guide_sum, cam.guide_lut, cam.guide_lut_n, s.z, prob);
    const int r12 = cam.greso[1] * cam.greso[2];
    const int c0 = idx / r12, rem = idx - c0 * r12, c1 = rem / cam.greso[2], c2 = rem - c1 * cam.greso[2];
    s.x = (s.x + (float) c0) * (1.f / (float) cam.greso[0]);
    s.y = (s.y + (float) c1) * (1.f / (float) cam.greso[1]);
    s.z = (s.z + (float) c2) * (1.f / (float) cam.greso[2]);
    return prob * (float) ncells;
}

template <int kCfg>
__device__ __forceinline__ int eval_secondary_edge(const DScene &sc, const DCamera &cam, V3f sample3, V3f &value0_out, V3f &tangent_out) {
    return eval_secondary_edge<kCfg, NoSecAdjoint>(sc, cam, sample3, value0_out, tangent_out, NoSecAdjoint());
}

// ---- batches of stage-0 survivors (used by the forward and the adjoint secondary-edge kernels; see kernels_impl.cuh)
#ifndef PSDR_SEC_FILL_ROUNDS
#define PSDR_SEC_FILL_ROUNDS 4
#endif
constexpr int kSecFillRounds = PSDR_SEC_FILL_ROUNDS;   // cap on the fill loop (tools/gpu_lb_sweep.sh)
struct SecSample {      // what the fill loop hands to stage 1
    SecCand cand;
    float pdf0;         // guiding pdf (1 without guiding)
};
template <int kCfg>
__device__ __forceinline__ bool sec_edge_draw(const DScene &sc, const DCamera &cam, const RenderParams &rp, long long j, SecSample &out) {
    const long long i = global_lane(rp, rp.perm ? (long long) __ldg(rp.perm + j) : j);
    if (i >= rp.n_lanes) return false;
    Pcg32 rng;
    rng.seed((unsigned long long) (i + rp.seed), (unsigned long long) i);
    if (rp.skip) rng.advance(rp.skip);
    const float d1 = rng.next_1d(), d2 = rng.next_1d(), d3 = rng.next_1d();
    V3f sample3(d3, d2, d1);
    out.pdf0 = 1.f;
    if (cam.guided) out.pdf0 = guide_sample_reuse(cam, sample3);     // path.cpp:279-281
    return sec_edge_stage0<kCfg>(sc, sample3, out.cand);
}
// warp-uniform batch loop; `body(sample)` is called by the lanes that hold a candidate
template <int kCfg, class F>
__device__ __forceinline__ void sec_edge_batches(const DScene &sc, const DCamera &cam, const RenderParams &rp, int block, F body, bool cta = false) {
    const long long span = rp.lane_end - rp.lane_begin;
    const unsigned lane = threadIdx.x & 31u;
    const long long n_warps = (long long) gridDim.x * (block / 32), warp = (long long) blockIdx.x * (block / 32) + (threadIdx.x >> 5);
    // Lane order (rp.perm == nullptr): warp w owns the contiguous slice [w per, (w + 1) per) -- all slices are statistically
    // alike.  Samples ordered along the edge list (rp.perm, edge_sort.cu): consecutive samples sit on the same stretch of
    // the same edge, whole stretches fail stage 0 (edges facing away from the light) and others pass it three times as
    // often as the average, so a warp takes every n_warps-th block of kSecBlock ordered samples instead -- coherent inside
    // a block, and every warp walks the whole edge list.
    constexpr long long kSecBlock = 256;
    const bool ordered = rp.perm != nullptr;
    const long long per = ordered ? (span + n_warps * kSecBlock - 1) / (n_warps * kSecBlock) * kSecBlock
                                  : ((span + n_warps - 1) / n_warps + 31) / 32 * 32;      // samples of this warp
    long long next = ordered ? 0 : warp * per;
    const long long end = ordered ? per : (next + per < span ? next + per : span);
    bool finished = false;      // (cta mode) this warp has drained its slice and only keeps the block barriers company
    while (true) {
        if (cta) {
            if (!__syncthreads_or(!finished)) break;
        } else __syncwarp();
        bool have = false;
        SecSample smp;
        for (int round = 0; round < kSecFillRounds; ++round) {
            const unsigned need = __ballot_sync(0xffffffffu, !have);
            if (need == 0u || next >= end) break;
            if (!have) {
                const long long t = next + __popc(need & ((1u << lane) - 1u));
                const long long j = ordered ? ((t / kSecBlock) * n_warps + warp) * kSecBlock + (t % kSecBlock) : t;
                if (t < end && j < span) have = sec_edge_draw<kCfg>(sc, cam, rp, j, smp);
            }
            next += __popc(need);
        }
        if (cta) {
            if (!__any_sync(0xffffffffu, have) && next >= end) finished = true;
            __syncthreads();        // fill phase | stage-1 phase
        } else if (!__any_sync(0xffffffffu, have)) {
            if (next >= end) break;
            continue;
        }
        if (have) body(smp);
    }
}

}  // namespace psdr
