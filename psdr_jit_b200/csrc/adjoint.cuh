// Reverse-mode (adjoint / VJP) of the path-space estimators: "radiative backprop with path replay".
//
// The reference obtains d(loss)/d(parameters) from Dr.Jit's AD graph, which materialises every
// intermediate of Li as an N-lane array in HBM.  Here a lane
//   1. replays its path with the primal code (Li<float, kBvh, /*kAD=*/true>) and records only what cannot be
//      recomputed cheaply: the vertices (triangle id + detached barycentrics), the light samples and the
//      detached pdfs / MIS weights / throughputs -- a few dozen words in registers or local memory;
//   2. sweeps the bounces backwards: with the suffix radiance known, every bounce is a product of a few
//      differentiable geometric factors whose adjoints are written out by hand below;
//   3. scatters the adjoints of the triangle records / materials / camera with atomics -- into a
//      shared-memory copy of the gradient table when it fits, flushed once per block, else into HBM.
//
// What carries derivatives is exactly SURVEY.md A.9: triangle records of every hit (p, face normal, shading
// normal, area Jacobian), light-sample position/area, BSDF and emitter parameters, the camera matrix.
// Sampling decisions, pdfs and MIS weights are detached.
#pragma once
#include "device_path.cuh"
#include "grad_layout.h"

#ifndef PSDR_AGG_PARTIAL
#define PSDR_AGG_PARTIAL 1      // 0: lanes aggregate their adds only when the whole warp arrives together
#endif

namespace psdr {

// Accumulator: either the block's shared-memory copy of [lo, hi) of the table or the global table.
// The atomic itself lives in ONE non-inlined function per translation unit: inlined, the ~300 scatter sites of
// the adjoint made up half of the kernel's instructions (generic-address atomics expand to ~45 SASS
// instructions each) and the kernel stalled on instruction fetch (profiles/r01e).
// Shared-memory float adds are compare-and-swap loops on sm_100 (ATOMS.CAST.SPIN): a k-way address conflict inside
// the warp costs k rounds, and the 32 lanes of a warp are the 32 samples of ONE pixel -- camera, emitter, BSDF and
// primary-triangle entries conflict 32 ways.  So the lanes that arrive together first check (MATCH.ALL) whether
// they all target the same entry; if so their values are summed with shuffles and one lane issues one add.
// Mixed targets fall back to one predicated add per lane.  No data-dependent early return: lanes that leave a
// non-inlined function in separate groups are not re-merged before the end of the caller's enclosing region.
__device__ __forceinline__ void grad_red(unsigned m, bool aggregate, unsigned smem_addr, float *g, int lo, int hi, int i, float v) {
    float val = (v != 0.f && isfinite(v)) ? v : 0.f;
    int same = 0;
    if (aggregate && (PSDR_AGG_PARTIAL || m == 0xffffffffu)) __match_all_sync(m, i, &same);      // (kernel-wide constant: the edge kernels' lanes sit on different edges)
    bool mine = true;
    if (same && (m & (m - 1u)) != 0u) {            // uniform over m: every lane targets entry i
        if (m == 0xffffffffu) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
        } else {
            float sum = 0.f;
            for (unsigned r = m; r; r &= r - 1u) sum += __shfl_sync(m, val, __ffs((int) r) - 1);
            val = sum;
        }
        mine = (threadIdx.x & 31) == __ffs((int) m) - 1;
    }
    const bool ok = mine && val != 0.f && isfinite(val);
    const bool in = smem_addr != 0u && i >= lo && i < hi;
    const int to_shared = ok && in, to_global = ok && !in;
    asm volatile(
        "{\n\t.reg .pred ps, pg;\n\t"
        "setp.ne.s32 ps, %0, 0;\n\t"
        "setp.ne.s32 pg, %1, 0;\n\t"
        "@ps red.shared.add.f32 [%2], %4;\n\t"
        "@pg red.global.add.f32 [%3], %4;\n\t}"
        ::"r"(to_shared), "r"(to_global), "r"(smem_addr + 4u * (unsigned) (i - lo)), "l"(g + i), "f"(val)
        : "memory");
}
// three consecutive entries (a vector): one MATCH on the base index, the three sums interleaved
static __device__ __noinline__ void grad_add3_impl(bool aggregate, unsigned smem_addr, float *g, int lo, int hi, int idx, float x, float y, float z) {
    const unsigned m = __activemask();
    float v0 = (x != 0.f && isfinite(x)) ? x : 0.f, v1 = (y != 0.f && isfinite(y)) ? y : 0.f, v2 = (z != 0.f && isfinite(z)) ? z : 0.f;
    bool mine = true;
    int same = 0;
    if (aggregate && (PSDR_AGG_PARTIAL || m == 0xffffffffu)) __match_all_sync(m, idx, &same);
    if (same && (m & (m - 1u)) != 0u) {            // uniform over m: every lane targets the same three entries
        if (m == 0xffffffffu) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                v0 += __shfl_xor_sync(0xffffffffu, v0, d);
                v1 += __shfl_xor_sync(0xffffffffu, v1, d);
                v2 += __shfl_xor_sync(0xffffffffu, v2, d);
            }
        } else {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
            for (unsigned r = m; r; r &= r - 1u) {
                const int src = __ffs((int) r) - 1;
                s0 += __shfl_sync(m, v0, src);
                s1 += __shfl_sync(m, v1, src);
                s2 += __shfl_sync(m, v2, src);
            }
            v0 = s0; v1 = s1; v2 = s2;
        }
        mine = (threadIdx.x & 31) == __ffs((int) m) - 1;
    }
    // a vector never straddles the shared window (its bounds are record-aligned section offsets of grad_layout.h)
    const bool in = smem_addr != 0u && idx >= lo && idx + 2 < hi;
    const int s0 = mine && in && v0 != 0.f && isfinite(v0), s1 = mine && in && v1 != 0.f && isfinite(v1), s2 = mine && in && v2 != 0.f && isfinite(v2);
    const int g0 = mine && !in && v0 != 0.f && isfinite(v0), g1 = mine && !in && v1 != 0.f && isfinite(v1), g2 = mine && !in && v2 != 0.f && isfinite(v2);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.s32 p, %0, 0;\n\t@p red.shared.add.f32 [%6], %8;\n\t"
        "setp.ne.s32 p, %1, 0;\n\t@p red.shared.add.f32 [%6+4], %9;\n\t"
        "setp.ne.s32 p, %2, 0;\n\t@p red.shared.add.f32 [%6+8], %10;\n\t"
        "setp.ne.s32 p, %3, 0;\n\t@p red.global.add.f32 [%7], %8;\n\t"
        "setp.ne.s32 p, %4, 0;\n\t@p red.global.add.f32 [%7+4], %9;\n\t"
        "setp.ne.s32 p, %5, 0;\n\t@p red.global.add.f32 [%7+8], %10;\n\t}"
        ::"r"(s0), "r"(s1), "r"(s2), "r"(g0), "r"(g1), "r"(g2), "r"(smem_addr + 4u * (unsigned) (idx - lo)), "l"(g + idx), "f"(v0), "f"(v1), "f"(v2)
        : "memory");
    __syncwarp(m);
}
static __device__ __noinline__ void grad_add1_impl(bool aggregate, unsigned smem_addr, float *g, int lo, int hi, int idx, float v) {
    const unsigned m = __activemask();
    grad_red(m, aggregate, smem_addr, g, lo, hi, idx, v);
    __syncwarp(m);
}
struct GradAcc {
    float *g;         // global table
    float *s;         // shared copy (nullptr = none)
    unsigned s_addr;  // its shared-window address (0 = none)
    int lo, hi;
    bool aggregate;   // sum across the lanes that target the same entry before adding (interior kernel: one pixel per warp)
    int mc;           // 1 = g is an NVLS multicast address (RenderParams::out_multicast): the shared copy is flushed with
                      // multimem.red (out_add); only tables that fit the shared window run in this mode (capi.cpp vjp_launch)
    __device__ __forceinline__ void add(int idx, float v) const { grad_add1_impl(aggregate, s_addr, g, lo, hi, idx, v); }
    __device__ __forceinline__ void add3(int idx, V3f v) const { grad_add3_impl(aggregate, s_addr, g, lo, hi, idx, v.x, v.y, v.z); }
};

__device__ __forceinline__ GradAcc grad_acc_begin(const GradLayout &gl, float *smem, int lo, int hi, bool use_smem, bool aggregate, int mc = 0) {
    GradAcc a;
    a.aggregate = aggregate;
    a.mc = mc;
    a.g = gl.base;
    a.s = use_smem ? smem : nullptr;
    a.s_addr = use_smem ? (unsigned) __cvta_generic_to_shared(smem) : 0u;
    a.lo = lo;
    a.hi = hi;
    if (use_smem) {
        for (int i = threadIdx.x; i < hi - lo; i += blockDim.x) smem[i] = 0.f;
        __syncthreads();
    }
    return a;
}
__device__ __forceinline__ void grad_acc_end(const GradAcc &a) {
    if (!a.s) return;
    __syncthreads();
    for (int i = threadIdx.x; i < a.hi - a.lo; i += blockDim.x) {
        const float v = a.s[i];
        if (v != 0.f) out_add(a.g + a.lo + i, v, a.mc);
    }
}

// ---- what the primal replay records -------------------------------------------------------------
template <int kD> struct PathRecord {
    int nv, nsh;                        // vertices found; vertices that were shaded (NEE + BSDF sample)
    int vtri[kD + 1];
    float vu[kD + 1], vv[kD + 1];
    V3f T[kD];                          // throughput in front of bounce k
    unsigned char nee_ok[kD], bnc_ok[kD], bnc_nz[kD];
    int ltri[kD], htri[kD];
    float la[kD], lb[kD], lc[kD], lpdf[kD], w1[kD], pdf0[kD], w2[kD];   // (la, lb) = barycentrics, or (la, lb, lc) = envmap position if ltri < 0
    __device__ __forceinline__ void reset() {
        nv = nsh = 0;
#pragma unroll
        for (int k = 0; k < kD; ++k) nee_ok[k] = bnc_ok[k] = bnc_nz[k] = 0;
    }
    __device__ __forceinline__ void vertex(int k, int tri, float u, float v) {
        if (k <= kD) { vtri[k] = tri; vu[k] = u; vv[k] = v; nv = k + 1; }
    }
    __device__ __forceinline__ void throughput(int k, V3f t) {
        if (k < kD) { T[k] = t; nsh = k + 1; }
    }
    __device__ __forceinline__ void bounce(int k, bool nonzero, float p0, float w) {
        if (k < kD) { bnc_ok[k] = 1; bnc_nz[k] = nonzero ? 1 : 0; pdf0[k] = p0; w2[k] = w; }
    }
    __device__ __forceinline__ void nee(int k, bool ok, int lt, V2f st, V3f p, int ht, float pdf, float w) {
        if (k < kD) {
            nee_ok[k] = ok ? 1 : 0; ltri[k] = lt; htri[k] = ht; lpdf[k] = pdf; w1[k] = w;
            if (lt >= 0) { la[k] = st.x; lb[k] = st.y; lc[k] = 0.f; } else { la[k] = p.x; lb[k] = p.y; lc[k] = p.z; }
        }
    }
};

// ---- isotropic BSDF as a function of the three cosines (what the adjoint differentiates) ---------
// ci = cos(wi, sh_n), co = cos(wo, sh_n), cio = dot(wi, wo).  Diffuse: reference src/bsdf/diffuse.cpp:23-55;
// Microfacet: src/bsdf/microfacet.cpp:22-68 + src/bsdf/ggx.cpp rewritten in frame-invariant quantities
// (|wi + wo|^2 = 2 + 2 cio, H.z = (ci + co)/|wi + wo|, <wi,H> = <wo,H> = (1 + cio)/|wi + wo|).
template <class S> struct BsdfP {
    int type, two_side;
    V3<S> diff, spec;
    S rough;
    V3<S> eta, k;       // RoughConductor
};
struct BsdfVals {       // the three parameter slots of a BSDF evaluated at a vertex (constants or texture lookups)
    V3f refl, spec;
    float rough;
};
template <class S> __device__ __forceinline__ BsdfP<S> bsdf_params(const DBsdf &b, const BsdfVals &v) {
    BsdfP<S> p;
    p.type = b.type;
    p.two_side = b.two_side;
    p.diff = V3<S>(S(v.refl.x), S(v.refl.y), S(v.refl.z));
    p.spec = V3<S>(S(v.spec.x), S(v.spec.y), S(v.spec.z));
    p.rough = S(v.rough);
    p.eta = V3<S>(S(b.eta[0]), S(b.eta[1]), S(b.eta[2]));
    p.k = V3<S>(S(b.kk[0]), S(b.kk[1]), S(b.kk[2]));
    return p;
}
template <class S> __device__ __forceinline__ S iso_smith_g1(S alpha, S vz, S vdoth) {
    const S s2 = S(1.f) - sqr(vz);
    const S xy_alpha_2 = sqr(alpha) * (val(s2) > 0.f ? s2 : S(0.f));
    S result = S(2.f) / (S(1.f) + sqrt_(S(1.f) + xy_alpha_2 / sqr(vz)));
    if (val(xy_alpha_2) == 0.f) result = S(1.f);
    if (val(vdoth) * val(vz) <= 0.f) result = S(0.f);
    return result;
}
// specular lobe without the Fresnel colour: D * G / (4 ci co + 1e-6), and the Fresnel blend weight e
template <class S> __device__ __forceinline__ void iso_specular(S rough, S ci, S co, S cio, S &dg, S &e) {
    const S alpha = sqr(rough);
    const S L = sqrt_(S(2.f) + S(2.f) * cio);
    const S hz = (ci + co) / L, vh = (S(1.f) + cio) / L;
    const S s2 = S(1.f) - sqr(hz);
    const S t = (val(s2) > 0.f ? s2 : S(0.f)) / sqr(alpha) + sqr(hz);
    S ggx = rcp_(S(kPi) * sqr(alpha) * sqr(t));
    if (!(val(ggx) * val(hz) > 1e-20f)) ggx = S(0.f);
    e = exp2_(vh * (S(-5.55473f) * vh - S(6.8316f)));
    dg = ggx * iso_smith_g1<S>(alpha, ci, vh) * iso_smith_g1<S>(alpha, co, vh) / (S(4.f) * co * ci + S(1e-6f));
}
// RoughConductor (src/bsdf/roughconductor.cpp:38-66) in the same quantities: D G / (4 ci), and <wi, H>
template <class S> __device__ __forceinline__ void iso_conductor(S alpha, S ci, S co, S cio, S &res, S &vh) {
    const S L = sqrt_(S(2.f) + S(2.f) * cio);
    const S hz = (ci + co) / L;
    vh = (S(1.f) + cio) / L;
    const S s2 = S(1.f) - sqr(hz);
    const S t = (val(s2) > 0.f ? s2 : S(0.f)) / sqr(alpha) + sqr(hz);
    S ggx = rcp_(S(kPi) * sqr(alpha) * sqr(t));
    if (!(val(ggx) * val(hz) > 1e-20f)) ggx = S(0.f);
    res = ggx * iso_smith_g1<S>(alpha, ci, vh) * iso_smith_g1<S>(alpha, co, vh) / (S(4.f) * ci);
}
// RoughDielectric (src/bsdf/roughdielectric.cpp:37-122) in the same quantities; with e = 1 (reflection) or eta (refraction):
// |wi + e wo|^2 = 1 + e^2 + 2 e cio, m.z = (ci + e co)/|.|, <wi, m> = (1 + e cio)/|.|, <wo, m> = (cio + e)/|.|
template <class S> __device__ __forceinline__ S iso_dielectric(S alpha, S m_eta, S m_inv_eta, S ci, S co, S cio) {
    if (val(ci) == 0.f) return S(0.f);
    const bool reflect = val(ci) * val(co) > 0.f, front = val(ci) > 0.f;
    const S eta = front ? m_eta : m_inv_eta, inv_eta = front ? m_inv_eta : m_eta;
    const S e = reflect ? S(1.f) : eta;
    const S L = sqrt_(S(1.f) + sqr(e) + S(2.f) * e * cio);
    S mz = (ci + e * co) / L, wim = (S(1.f) + e * cio) / L, wom = (cio + e) / L;
    if (signbit_(val(mz))) { mz = -mz; wim = -wim; wom = -wom; }
    const S s2 = S(1.f) - sqr(mz);
    const S t = (val(s2) > 0.f ? s2 : S(0.f)) / sqr(alpha) + sqr(mz);
    S D = rcp_(S(kPi) * sqr(alpha) * sqr(t));
    if (!(val(D) * val(mz) > 1e-20f)) D = S(0.f);
    const S F = fresnel_dielectric<S>(m_eta, wim).r;
    const S G = iso_smith_g1<S>(alpha, ci, wim) * iso_smith_g1<S>(alpha, co, wom);
    if (reflect) return F * D * G / (S(4.f) * abs_(ci));
    return abs_((sqr(inv_eta) * (S(1.f) - F) * D * G * eta * eta * wim * wom) / (ci * sqr(wim + eta * wom)));
}
// MicrofacetPerVertex (src/bsdf/microfacet_pv.cpp:20-68): specular lobe without the Fresnel colour,
// ggx smithG / (4 ci co + 1e-6) in its own NDF / Schlick-Smith form, and the Fresnel blend weight e
template <class S> __device__ __forceinline__ void iso_specular_pv(S rough, S ci, S co, S cio, S &dg, S &e) {
    const S L = sqrt_(S(2.f) + S(2.f) * cio);
    const S hz = (ci + co) / L, vh = (S(1.f) + cio) / L;
    const S alpha = sqr(rough), k = sqr(rough + S(1.f)) / S(8.f);
    const S tmp = alpha / (hz * hz * (sqr(alpha) - S(1.f)) + S(1.f));
    const S ggx = tmp * tmp * S(kInvPi);
    e = exp2_(vh * (S(-5.55473f) * vh - S(6.8316f)));
    const S smithG = (ci / (ci * (S(1.f) - k) + k)) * (co / (co * (S(1.f) - k) + k));
    dg = ggx * smithG / (S(4.f) * co * ci + S(1e-6f));
}
template <class S, int kCfg> __device__ __forceinline__ V3<S> bsdf_iso(const BsdfP<S> &b, S ci, S co, S cio) {
    if (b.two_side) {
        if (signbit_(val(ci))) co = -co;
        ci = abs_(ci);
    }
    if ((kCfg & kCfgExt) && b.type == 3) return V3<S>(iso_dielectric<S>(b.rough, b.eta.x, b.eta.y, ci, co, cio));
    if (!(val(ci) > 0.f && val(co) > 0.f)) return V3<S>(S(0.f));
    if ((kCfg & kCfgExt) && b.type == 4) {
        S dg, e;
        iso_specular_pv<S>(b.rough, ci, co, cio, dg, e);
        const V3<S> fresnel = b.spec + (V3<S>(S(1.f)) - b.spec) * e;
        return (b.diff * S(kInvPi) + fresnel * dg) * co;
    }
    if ((kCfg & kCfgFull) && b.type == 1) {
        S dg, e;
        iso_specular<S>(b.rough, ci, co, cio, dg, e);
        const V3<S> fresnel = b.spec + (V3<S>(S(1.f)) - b.spec) * e;
        return (b.diff * S(kInvPi) + fresnel * dg) * co;
    }
    if ((kCfg & kCfgExt) && b.type == 2) {
        S res, vh;
        iso_conductor<S>(b.rough, ci, co, cio, res, vh);
        if (val(res) == 0.f) return V3<S>(S(0.f));
        const V3<S> F(fresnel_conductor<S>(b.eta.x, b.k.x, vh), fresnel_conductor<S>(b.eta.y, b.k.y, vh), fresnel_conductor<S>(b.eta.z, b.k.z, vh));
        return F * res * b.spec;
    }
    return b.diff * S(kInvPi) * co;
}

struct BsdfJet {        // value and partials of sum_c W_c f_c
    V3f f;
    float d_ci, d_co, d_cio;
};
template <int kCfg> __device__ __forceinline__ BsdfJet bsdf_jet(const DBsdf &b, const BsdfVals &bv, float ci, float co, float cio, V3f W) {
    BsdfJet j;
    const BsdfP<Dual> p = bsdf_params<Dual>(b, bv);
    const V3d a = bsdf_iso<Dual, kCfg>(p, Dual(ci, 1.f), Dual(co), Dual(cio));
    const V3d c = bsdf_iso<Dual, kCfg>(p, Dual(ci), Dual(co, 1.f), Dual(cio));
    j.f = val(a);
    j.d_ci = W.x * a.x.d + W.y * a.y.d + W.z * a.z.d;
    j.d_co = W.x * c.x.d + W.y * c.y.d + W.z * c.z.d;
    j.d_cio = 0.f;
    if ((kCfg & kCfgFull) && b.type != 0) {
        const V3d e = bsdf_iso<Dual, kCfg>(p, Dual(ci), Dual(co), Dual(cio, 1.f));
        j.d_cio = W.x * e.x.d + W.y * e.y.d + W.z * e.z.d;
    }
    return j;
}
// texel gradients of one texture slot: value_bar (rgb, or x for 1 channel) scattered through the four bilinear taps, and
// the lookup's uv derivative added to uv_bar
__device__ __forceinline__ void tex_slot_grad(const GradAcc &acc, const GradLayout &gl, const DTex &t, V2f uv, V3f value_bar, V2f &uv_bar) {
    EnvTexelTaps tp;
    tex_eval_uv<float>(t, false, uv, &tp);
    const int idx[4] = {tp.i00, tp.i10, tp.i01, tp.i11};
    const float bw[4] = {tp.w0y * tp.w0x, tp.w0y * tp.w1x, tp.w1y * tp.w0x, tp.w1y * tp.w1x};
    const int tbase = gl.total + t.goff;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (t.ch == 1) acc.add(tbase + idx[k], value_bar.x * bw[k]);
        else acc.add3(tbase + 3 * idx[k], value_bar * bw[k]);
    }
    // derivatives of the lookup w.r.t. the texture coordinate and w.r.t. the bitmap's uv transform (scale, rotation,
    // translation: bitmap.cpp:64-72): six dual evaluations on a copy of the slot WITHOUT the forward-mode tangents a user
    // may have set on the transform.  The four transform gradients follow the slot's texels in the table.
    DTex q = t;
    q.d_cr = q.d_sr = q.d_scale = q.d_tx = q.d_ty = 0.f;
    q.ddata = nullptr;
    const V3d ru = tex_eval_uv<Dual>(q, false, V2d(Dual(uv.x, 1.f), Dual(uv.y)));
    const V3d rv = tex_eval_uv<Dual>(q, false, V2d(Dual(uv.x), Dual(uv.y, 1.f)));
    uv_bar.x += value_bar.x * ru.x.d + value_bar.y * ru.y.d + value_bar.z * ru.z.d;
    uv_bar.y += value_bar.x * rv.x.d + value_bar.y * rv.y.d + value_bar.z * rv.z.d;
    const int ubase = tbase + t.ch * t.w * t.h;
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        DTex qk = q;
        if (k == 0) qk.d_scale = 1.f;
        else if (k == 1) { qk.d_cr = -q.sr; qk.d_sr = q.cr; }      // d cos(rot), d sin(rot)
        else if (k == 2) qk.d_tx = 1.f;
        else qk.d_ty = 1.f;
        const V3d r = tex_eval_uv<Dual>(qk, false, V2d(Dual(uv.x), Dual(uv.y)));
        acc.add(ubase + k, value_bar.x * r.x.d + value_bar.y * r.y.d + value_bar.z * r.z.d);
    }
}
// d(sum_c W_c f_c * scale)/d(params): reflectance (Diffuse / Microfacet diffuse), Microfacet specular + roughness;
// constants go to the BSDF block of the table, textured slots to their texel blocks.
// uv_bar: d(contribution)/d(texture coordinate) through textured slots (used at the primary hit only)
// (tri, bu, bv2): the vertex's triangle and barycentrics -- MicrofacetPerVertex scatters through them into the per-vertex
// gradient block, and bc_bar receives d(contribution)/d(barycentrics) (used at the primary hit only, like uv_bar)
template <int kCfg> __device__ __forceinline__ void bsdf_param_grad(const GradAcc &acc, const GradLayout &gl, const DScene &sc, int bi, const DBsdf &b, const BsdfVals &bv,
                                                float ci, float co, float cio, V3f W, float scale, V2f uv, V2f &uv_bar, int tri, float bu, float bv2, V2f &bc_bar) {
    if (b.two_side) {
        if (signbit_(ci)) co = -co;
        ci = fabsf(ci);
    }
    const int base = gl.off_bsdf + kGradBsdf * bi;
    if ((kCfg & kCfgExt) && b.type == 3) {
        // f_c = iso_dielectric(alpha): one value for the three channels; eta is fixed at creation
        const Dual f = iso_dielectric<Dual>(Dual(bv.rough, 1.f), Dual(b.eta[0]), Dual(b.eta[1]), Dual(ci), Dual(co), Dual(cio));
        const float g_alpha = (W.x + W.y + W.z) * f.d * scale;
        if (b.tex[2].w > 0) tex_slot_grad(acc, gl, b.tex[2], uv, V3f(g_alpha, 0.f, 0.f), uv_bar);
        else acc.add(base + 3, g_alpha);
        return;
    }
    if (!(ci > 0.f && co > 0.f)) return;
    if ((kCfg & kCfgExt) && b.type == 4) {
        // f_c = (diff_c / pi + (F0_c + (1 - F0_c) e) dg(rough)) co with the three parameters interpolated from the corners
        Dual dg, e;
        iso_specular_pv<Dual>(Dual(bv.rough, 1.f), Dual(ci), Dual(co), Dual(cio), dg, e);
        const float kd = kInvPi * co * scale, ks = (1.f - e.v) * dg.v * co * scale;
        const float F0[3] = {bv.spec.x, bv.spec.y, bv.spec.z}, Wc[3] = {W.x, W.y, W.z};
        float vbar[7];
        float gr = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            vbar[c] = Wc[c] * ks;
            vbar[3 + c] = Wc[c] * kd;
            gr += Wc[c] * (F0[c] + (1.f - F0[c]) * e.v) * dg.d;
        }
        vbar[6] = gr * co * scale;
        const int i0 = __ldg(sc.face_idx + 3 * tri), i1 = __ldg(sc.face_idx + 3 * tri + 1), i2 = __ldg(sc.face_idx + 3 * tri + 2);
        const int pbase = gl.total + b.pv_goff;
        const float w0 = 1.f - bu - bv2;
        float dbu = 0.f, dbv = 0.f;
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            if (vbar[c] == 0.f) continue;
            const float a0 = i0 < b.pv_n ? __ldg(b.pv + 7 * i0 + c) : 0.f, a1 = i1 < b.pv_n ? __ldg(b.pv + 7 * i1 + c) : 0.f,
                        a2 = i2 < b.pv_n ? __ldg(b.pv + 7 * i2 + c) : 0.f;
            if (i0 < b.pv_n) acc.add(pbase + 7 * i0 + c, vbar[c] * w0);
            if (i1 < b.pv_n) acc.add(pbase + 7 * i1 + c, vbar[c] * bu);
            if (i2 < b.pv_n) acc.add(pbase + 7 * i2 + c, vbar[c] * bv2);
            dbu += vbar[c] * (a1 - a0);
            dbv += vbar[c] * (a2 - a0);
        }
        bc_bar.x += dbu;
        bc_bar.y += dbv;
        return;
    }
    if ((kCfg & kCfgExt) && b.type == 2) {
        // f_c = F(eta_c, k_c, vh) res(alpha) spec_c
        Dual res, vh;
        iso_conductor<Dual>(Dual(bv.rough, 1.f), Dual(ci), Dual(co), Dual(cio), res, vh);
        if (res.v == 0.f) return;
        const float sp[3] = {bv.spec.x, bv.spec.y, bv.spec.z}, Wc[3] = {W.x, W.y, W.z};
        float F[3], g_alpha = 0.f;
        V3f g_eta, g_k;
        float *ge = &g_eta.x, *gk = &g_k.x;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const Dual fe = fresnel_conductor<Dual>(Dual(b.eta[c], 1.f), Dual(b.kk[c]), Dual(vh.v));
            const Dual fk = fresnel_conductor<Dual>(Dual(b.eta[c]), Dual(b.kk[c], 1.f), Dual(vh.v));
            F[c] = fe.v;
            ge[c] = Wc[c] * fe.d * res.v * sp[c] * scale;
            gk[c] = Wc[c] * fk.d * res.v * sp[c] * scale;
            g_alpha += Wc[c] * F[c] * sp[c];
        }
        acc.add3(base + 8, g_eta);
        acc.add3(base + 12, g_k);
        const V3f g_spec(W.x * F[0] * res.v * scale, W.y * F[1] * res.v * scale, W.z * F[2] * res.v * scale);
        if (b.tex[1].w > 0) tex_slot_grad(acc, gl, b.tex[1], uv, g_spec, uv_bar);
        else acc.add3(base + 4, g_spec);
        if (b.tex[2].w > 0) tex_slot_grad(acc, gl, b.tex[2], uv, V3f(g_alpha * res.d * scale, 0.f, 0.f), uv_bar);
        else acc.add(base + 3, g_alpha * res.d * scale);
        return;
    }
    const float k = kInvPi * co * scale;
    if ((kCfg & kCfgExt) && b.tex[0].w > 0) tex_slot_grad(acc, gl, b.tex[0], uv, W * k, uv_bar);
    else acc.add3(base, V3f(W.x * k, W.y * k, W.z * k));
    if ((kCfg & kCfgFull) && b.type == 1) {
        Dual dg, e;
        iso_specular<Dual>(Dual(bv.rough, 1.f), Dual(ci), Dual(co), Dual(cio), dg, e);
        // f_spec,c = (F0_c + (1 - F0_c) e) dg co
        const float ks = (1.f - e.v) * dg.v * co * scale;
        if ((kCfg & kCfgExt) && b.tex[1].w > 0) tex_slot_grad(acc, gl, b.tex[1], uv, W * ks, uv_bar);
        else acc.add3(base + 4, V3f(W.x * ks, W.y * ks, W.z * ks));
        float gr = 0.f;
        const float F0[3] = {bv.spec.x, bv.spec.y, bv.spec.z}, Wc[3] = {W.x, W.y, W.z};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const Dual fr = Dual(F0[c]) + Dual(1.f - F0[c]) * e;
            gr += Wc[c] * (fr * dg).d;
        }
        if ((kCfg & kCfgExt) && b.tex[2].w > 0) tex_slot_grad(acc, gl, b.tex[2], uv, V3f(gr * co * scale, 0.f, 0.f), uv_bar);
        else acc.add(base + 3, gr * co * scale);
    }
}

// ---- geometry of a recorded vertex ---------------------------------------------------------------
struct VtxGeo {
    V3f p, fn, shn, m;      // position, face normal, shading normal, un-normalised interpolated normal
    float minv;             // 1 / |m|
    float area, u, v;
    int tri, mesh, bsdf, emitter;
    bool face_normals;
    V2f uv, duv0, duv1;     // texture coordinate and its edge differences (uv = uv0 + u duv0 + v duv1)
    BsdfVals bv;            // reflectance / diffuseReflectance, specularReflectance, roughness of the BSDF at uv
};
#ifndef PSDR_VJP_GEO_NOINLINE
#define PSDR_VJP_GEO_NOINLINE 1     // out-of-line vertex_geo: interior adjoint 8.92 -> 8.63 ms (profiles/r02g vjp sweep)
#endif
#if PSDR_VJP_GEO_NOINLINE
template <int kCfg> static __device__ __noinline__ VtxGeo vertex_geo(const DScene &sc, int tri, float u, float v) {
#else
template <int kCfg> __device__ __forceinline__ VtxGeo vertex_geo(const DScene &sc, int tri, float u, float v) {
#endif
    VtxGeo g;
    const TriRec<float> T = load_tri<float>(sc, tri);
    const ShadeRec<float> N = load_shade<float>(sc, tri);
    const DMesh mesh = sc.meshes[T.mesh];
    g.tri = tri;
    g.mesh = T.mesh;
    g.bsdf = mesh.bsdf;
    g.emitter = mesh.emitter;
    g.face_normals = (mesh.flags & 1) != 0;
    g.u = u;
    g.v = v;
    g.p = bilinear(T.p0, T.e1, T.e2, V2f(u, v));
    g.fn = N.fn;
    g.area = T.area;
    g.m = bilinear(N.n0, N.n1 - N.n0, N.n2 - N.n0, V2f(u, v));
    g.minv = 1.f / norm(g.m);
    g.shn = g.face_normals ? g.fn : g.m * g.minv;
    g.uv = g.duv0 = g.duv1 = V2f(0.f, 0.f);
    if (mesh.flags & 2) {
        const float2 t0 = __ldg(sc.uv + 3 * tri), t1 = __ldg(sc.uv + 3 * tri + 1), t2 = __ldg(sc.uv + 3 * tri + 2);
        g.duv0 = V2f(t1.x - t0.x, t1.y - t0.y);
        g.duv1 = V2f(t2.x - t0.x, t2.y - t0.y);
        g.uv = V2f(fmaf(g.duv0.x, u, fmaf(g.duv1.x, v, t0.x)), fmaf(g.duv0.y, u, fmaf(g.duv1.y, v, t0.y)));
    }
    g.bv.refl = g.bv.spec = V3f(0.f, 0.f, 0.f);
    g.bv.rough = 0.f;
    if (g.bsdf >= 0) {
        const DBsdf &b = sc.bsdfs[g.bsdf];
        constexpr bool kTex = (kCfg & kCfgExt) != 0;      // bitmap-valued slots exist in the extended family only
        g.bv.refl = kTex && b.tex[0].w > 0 ? tex_eval_uv<float>(b.tex[0], false, g.uv) : V3f(b.refl[0], b.refl[1], b.refl[2]);
        g.bv.spec = kTex && b.tex[1].w > 0 ? tex_eval_uv<float>(b.tex[1], false, g.uv) : V3f(b.spec[0], b.spec[1], b.spec[2]);
        g.bv.rough = kTex && b.tex[2].w > 0 ? tex_eval_uv<float>(b.tex[2], false, g.uv).x : b.rough;
        if (kTex && b.type == 4) {      // MicrofacetPerVertex: the three parameters interpolated from the triangle's corners
            Its<float> f;
            f.tri = tri;
            f.bc = V2f(u, v);
            g.bv.spec = V3f(pv_interp<float>(sc, b, f, 0), pv_interp<float>(sc, b, f, 1), pv_interp<float>(sc, b, f, 2));
            g.bv.refl = V3f(pv_interp<float>(sc, b, f, 3), pv_interp<float>(sc, b, f, 4), pv_interp<float>(sc, b, f, 5));
            g.bv.rough = pv_interp<float>(sc, b, f, 6);
        }
    }
    return g;
}

// the three parameter slots of a record at a texture coordinate (constants or bitmap lookups)
template <int kCfg> __device__ __forceinline__ BsdfVals bsdf_vals_at(const DBsdf &b, V2f uv) {
    constexpr bool kTex = (kCfg & kCfgExt) != 0;
    BsdfVals v;
    v.refl = kTex && b.tex[0].w > 0 ? tex_eval_uv<float>(b.tex[0], false, uv) : V3f(b.refl[0], b.refl[1], b.refl[2]);
    v.spec = kTex && b.tex[1].w > 0 ? tex_eval_uv<float>(b.tex[1], false, uv) : V3f(b.spec[0], b.spec[1], b.spec[2]);
    v.rough = kTex && b.tex[2].w > 0 ? tex_eval_uv<float>(b.tex[2], false, uv).x : b.rough;
    return v;
}

// ---- NormalMap in reverse mode ---------------------------------------------------------------------------------------
// Its value is not a function of the three cosines: the perturbed frame is built from the normal map (shading frame) and
// dp_du (world frame).  The adjoint therefore differentiates the forward code itself, written as a function of
// world-space quantities -- wi, wo, the shading normal, dp_du and the normal-map value -- by forward mode: fifteen dual
// evaluations per event in ONE rolled loop (one copy of the code).  Both records are local copies without the
// forward-mode tangents a user may have set on the scene (psdr_scene_set_tangent), which would leak into the duals.
__device__ __forceinline__ void strip_tangents(DBsdf &b) {
#pragma unroll
    for (int c = 0; c < 3; ++c) b.d_refl[c] = b.d_spec[c] = b.d_eta[c] = b.d_kk[c] = 0.f;
    b.d_rough = 0.f;
    b.d_pv = nullptr;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        b.tex[k].ddata = nullptr;
        b.tex[k].d_cr = b.tex[k].d_sr = b.tex[k].d_scale = b.tex[k].d_tx = b.tex[k].d_ty = 0.f;
    }
}
template <class S> struct NmVertex {      // what Scene::ray_intersect leaves in the intersection record, as arguments
    int tri;
    V2f bc, uv;
    bool valid_dp;
};
template <class S, int kCfg>
__device__ __forceinline__ Its<S> nm_its(const NmVertex<S> &v, V3<S> shn, V3<S> dp_du) {
    Its<S> its;
    its.valid = true;
    its.tri = v.tri;
    its.mesh = -1;
    its.bc = V2<S>(S(v.bc.x), S(v.bc.y));
    its.uv = V2<S>(S(v.uv.x), S(v.uv.y));
    its.dp_du = dp_du;
    its.sh_n = shn;
    coordinate_system(shn, its.sh_s, its.sh_t);
    if (v.valid_dp) {      // scene.cpp:764-765
        its.sh_s = normalize(dp_du - shn * dot(shn, dp_du));
        its.sh_t = cross(shn, its.sh_s);
    }
    return its;
}
struct NmJet {
    V3f f;            // BSDF value
    float g[15];      // d(sum_c W_c f_c) / d(wi, wo, shn, dp_du, normal-map value)
    bool ok;          // all of it finite
};
template <int kCfg>
static __device__ __noinline__ NmJet normalmap_jet(const DScene &sc, DBsdf &bl, const DBsdf &nbl, const NmVertex<Dual> &v, V3f shn, V3f dp_du, V3f wi, V3f wo, V3f W) {
    NmJet j;
    j.ok = true;
    j.f = V3f(0.f, 0.f, 0.f);
#pragma unroll 1
    for (int a = 0; a < 15; ++a) {
        const int grp = a / 3, c = a - 3 * grp;
        auto seed = [&](V3f x, int gsel) {
            return V3d(Dual(x.x, (grp == gsel && c == 0) ? 1.f : 0.f), Dual(x.y, (grp == gsel && c == 1) ? 1.f : 0.f), Dual(x.z, (grp == gsel && c == 2) ? 1.f : 0.f));
        };
#pragma unroll
        for (int k = 0; k < 3; ++k) bl.d_refl[k] = (grp == 4 && c == k) ? 1.f : 0.f;
        const Its<Dual> its = nm_its<Dual, kCfg>(v, seed(shn, 2), seed(dp_du, 3));
        const V3d r = normalmap_eval_nb<Dual, kCfg>(sc, bl, nbl, its, its.to_local(seed(wi, 0)), its.to_local(seed(wo, 1)));
        j.g[a] = W.x * r.x.d + W.y * r.y.d + W.z * r.z.d;
        j.f = val(r);
        j.ok = j.ok && isfinite(j.g[a]) && isfinite(r.x.v) && isfinite(r.y.v) && isfinite(r.z.v);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) bl.d_refl[k] = 0.f;
    return j;
}

struct VtxAdj {
    V3f p, shn, fn;
    float area;
    V2f uv;
    V2f bc;       // d/d(barycentrics) through per-vertex BSDF tables (consumed at the primary hit, where (u, v) are differentiable)
};

// adjoint of v -> (w = v / |v|, t = |v|):  v_bar = (w_bar - w <w, w_bar>)/t + t_bar w
__device__ __forceinline__ V3f unit_adj(V3f w, float t, V3f w_bar, float t_bar) {
    const float k = dot(w, w_bar);
    return (w_bar - w * k) * (1.f / t) + w * t_bar;
}

// scatter the adjoint of a vertex pinned to its triangle: p = p0 + u e1 + v e2, sh_n = normalize(n0 + u (n1-n0) + v (n2-n0))
__device__ __forceinline__ void scatter_shading_normal(const GradAcc &acc, const VtxGeo &g, V3f shn_bar, V3f &m_bar_out) {
    m_bar_out = V3f(0.f, 0.f, 0.f);
    const int b = kGradTri * g.tri;
    if (g.face_normals) {
        acc.add3(b + 19, shn_bar);
        return;
    }
    const V3f m_bar = (shn_bar - g.shn * dot(g.shn, shn_bar)) * g.minv;
    acc.add3(b + 10, m_bar * (1.f - g.u - g.v));
    acc.add3(b + 13, m_bar * g.u);
    acc.add3(b + 16, m_bar * g.v);
    m_bar_out = m_bar;
}
__device__ __forceinline__ void scatter_pinned_vertex(const GradAcc &acc, const VtxGeo &g, const VtxAdj &a) {
    const int b = kGradTri * g.tri;
    acc.add3(b, a.p);
    acc.add3(b + 3, a.p * g.u);
    acc.add3(b + 6, a.p * g.v);
    acc.add3(b + 19, a.fn);
    acc.add(b + 9, a.area);
    V3f unused;
    scatter_shading_normal(acc, g, a.shn, unused);
}

// adjoint of (u, v, t) = ray_intersect_triangle(p0, e1, e2, o, d)  [o + t d = p0 + u e1 + v e2]:
// with A = [e1 e2 -d], A dx = do - dp0 - u de1 - v de2 + t dd  =>  r_bar = A^{-T} x_bar.
__device__ __forceinline__ V3f isect_adj(V3f e1, V3f e2, V3f d, float u_bar, float v_bar, float t_bar) {
    const V3f nd = -d;
    const V3f c0 = cross(e2, nd), c1 = cross(nd, e1), c2 = cross(e1, e2);
    const float det = dot(e1, c0);
    return (c0 * u_bar + c1 * v_bar + c2 * t_bar) * (1.f / det);
}
// scatter r_bar of isect_adj into the triangle's geometry; returns nothing for o/d (the caller owns them)
__device__ __forceinline__ void scatter_isect_tri(const GradAcc &acc, int tri, float u, float v, V3f r_bar) {
    const int b = kGradTri * tri;
    acc.add3(b, -r_bar);
    acc.add3(b + 3, r_bar * (-u));
    acc.add3(b + 6, r_bar * (-v));
}
// camera ray o = to_world * (0,0,0,1), d = to_world[:3,:3] * d_cam
// (oc = camera-space origin: zero for the perspective camera, the near-plane point of the orthographic one)
__device__ __forceinline__ void scatter_camera_ray(const GradAcc &acc, const GradLayout &gl, V3f dc, V3f o_bar, V3f d_bar, V3f oc = V3f(0.f, 0.f, 0.f)) {
    const int b = gl.off_cam;
    const float ob[3] = {o_bar.x, o_bar.y, o_bar.z}, db[3] = {d_bar.x, d_bar.y, d_bar.z}, c[3] = {dc.x, dc.y, dc.z}, q[3] = {oc.x, oc.y, oc.z};
    float f[12];      // rows 0..2 of d to_world: (d_bar_i * dc + o_bar_i * oc, o_bar_i); twelve consecutive entries = four vector adds
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) f[4 * i + j] = fmaf(ob[i], q[j], db[i] * c[j]);
        f[4 * i + 3] = ob[i];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) acc.add3(b + 3 * k, V3f(f[3 * k], f[3 * k + 1], f[3 * k + 2]));
}

// Environment-map radiance along `dir` weighted by Wc (contribution = sum_c Wc_c Le_c(dir)): scatters the
// gradients of the texels, the scale and the from_world rotation, returns d contribution / d dir and Le.
static __device__ __noinline__ V3f env_le_adjoint(const GradAcc &acc, const GradLayout &gl, const DEnv &env, V3f dir, V3f Wc, V3f &Le_out) {
    EnvTexelTaps tp;
    const V3f v = mul3x3<float>(env.from_world, nullptr, dir);
    const V3f tex = bitmap_eval_envmap<float>(env.data, nullptr, env.w, env.h, envmap_dir_to_uv<float>(v), &tp);
    Le_out = tex * env.scale;
    const float wsum = Wc.x + Wc.y + Wc.z;
    if (wsum == 0.f && Wc.x == 0.f && Wc.y == 0.f) return V3f(0.f, 0.f, 0.f);
    const int base = gl.off_env;
    acc.add(base, Wc.x * tex.x + Wc.y * tex.y + Wc.z * tex.z);                 // d scale
    const int idx[4] = {tp.i00, tp.i10, tp.i01, tp.i11};
    const float bw[4] = {tp.w0y * tp.w0x, tp.w0y * tp.w1x, tp.w1y * tp.w0x, tp.w1y * tp.w1x};
#pragma unroll
    for (int k = 0; k < 4; ++k) acc.add3(base + kGradEnvHead + 3 * idx[k], Wc * (env.scale * bw[k]));
    // direction: three dual evaluations in the map's local frame (parameter tangents switched off)
    V3f v_bar;
    {
        float gb[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const V3d vd(Dual(v.x, a == 0 ? 1.f : 0.f), Dual(v.y, a == 1 ? 1.f : 0.f), Dual(v.z, a == 2 ? 1.f : 0.f));
            const V3d r = bitmap_eval_envmap<Dual>(env.data, nullptr, env.w, env.h, envmap_dir_to_uv<Dual>(vd));
            gb[a] = (Wc.x * r.x.d + Wc.y * r.y.d + Wc.z * r.z.d) * env.scale;
        }
        v_bar = V3f(gb[0], gb[1], gb[2]);
    }
    const float vb[3] = {v_bar.x, v_bar.y, v_bar.z}, dd[3] = {dir.x, dir.y, dir.z};
#pragma unroll
    for (int i = 0; i < 3; ++i) acc.add3(base + 1 + 3 * i, V3f(vb[i] * dd[0], vb[i] * dd[1], vb[i] * dd[2]));      // v = F dir
    const float *F = env.from_world;
    return V3f(F[0] * vb[0] + F[3] * vb[1] + F[6] * vb[2], F[1] * vb[0] + F[4] * vb[1] + F[7] * vb[2], F[2] * vb[0] + F[5] * vb[1] + F[8] * vb[2]);
}

// One scattering event x -> y seen from x: C = sum_c W_c f_c(ci, co, cio) * |cos_y| / t^2 * J_y * scale
// (scale = detached 1/pdf * MIS weight).  Accumulates the adjoints of x (p, sh_n), of the previous point
// (through wi) and returns those of y.
struct EventAdj {
    V3f py, ny;       // d/d p_y, d/d n_y (face normal at y)
    float area_y;     // d/d area_y
    V3f wi_bar;       // adjoint of the unit direction wi at x (caller maps it to the previous point / camera)
    V3f f;            // BSDF value (rgb)
    float geo;        // |cos_y| / t^2 * scale  (J = 1 in the primal)
};
// wo_extra(f * geo) lets the caller add d(contribution)/d(wo) coming from a direction-dependent emitter (envmap).
template <int kCfg, class F>
__device__ __forceinline__ EventAdj event_adjoint(const GradAcc &acc, const GradLayout &gl, const DScene &sc, const VtxGeo &x, V3f wi,
                                                  V3f py, V3f ny, float area_y, V3f W, float scale, VtxAdj &xa, F wo_extra) {
    EventAdj r;
    r.py = r.ny = r.wi_bar = V3f(0.f, 0.f, 0.f);
    r.area_y = 0.f;
    r.f = V3f(0.f, 0.f, 0.f);
    r.geo = 0.f;
    if (x.bsdf < 0) return r;
    const DBsdf &b = sc.bsdfs[x.bsdf];
    const V3f vec = py - x.p;
    const float t = norm(vec);
    const V3f wo = vec / t;
    const float cy = -dot(ny, wo);
    const float G = fabsf(cy) / (t * t);
    // adjoints of wi, wo, sh_n (d(sum_c W_c f_c)/d(.) times phi_bar = G scale): through the three cosines, or (NormalMap)
    // from the world-space jet
    V3f f, wi_bar_v, wo_bar_v, shn_bar_v;
    if ((kCfg & kCfgExt) && b.type == 5) {
        DBsdf bl = b, nbl = sc.bsdfs[b.nested];
        strip_tangents(bl);
        strip_tangents(nbl);
        const V3f nm_value = bsdf_vals_at<kCfg>(b, x.uv).refl;       // the map at this vertex: constant or bitmap lookup
        bl.tex[0].w = bl.tex[0].h = 0;                                // the duals differentiate w.r.t. the VALUE; the texel
        bl.refl[0] = nm_value.x; bl.refl[1] = nm_value.y; bl.refl[2] = nm_value.z;     // scatter follows below
        const TriRec<float> T = load_tri<float>(sc, x.tri);
        const float det = x.duv0.x * x.duv1.y - x.duv0.y * x.duv1.x;
        NmVertex<Dual> nv;
        nv.tri = x.tri;
        nv.bc = V2f(x.u, x.v);
        nv.uv = x.uv;
        nv.valid_dp = det != 0.f;
        const float inv_det = nv.valid_dp ? 1.f / det : 0.f;
        const V3f dp_du = nv.valid_dp ? (T.e1 * x.duv1.y - T.e2 * x.duv0.y) * inv_det : V3f(0.f, 0.f, 0.f);
        const NmJet j = normalmap_jet<kCfg>(sc, bl, nbl, nv, x.shn, dp_du, wi, wo, W);
        if (!j.ok) return r;                                          // a NaN value is scrubbed by the forward pass: no gradient
        f = j.f;
        const float ps = G * scale;                                   // = phi_bar below
        wi_bar_v = V3f(j.g[0], j.g[1], j.g[2]) * ps;
        wo_bar_v = V3f(j.g[3], j.g[4], j.g[5]) * ps;
        shn_bar_v = V3f(j.g[6], j.g[7], j.g[8]) * ps;
        // dp_du = (e1 duv1.y - e2 duv0.y) / det
        const V3f dpdu_bar = V3f(j.g[9], j.g[10], j.g[11]) * ps;
        acc.add3(kGradTri * x.tri + 3, dpdu_bar * (x.duv1.y * inv_det));
        acc.add3(kGradTri * x.tri + 6, dpdu_bar * (-x.duv0.y * inv_det));
        // the normal map itself
        const V3f nm_bar = V3f(j.g[12], j.g[13], j.g[14]) * ps;
        if (b.tex[0].w > 0) tex_slot_grad(acc, gl, b.tex[0], x.uv, nm_bar, xa.uv);
        else acc.add3(gl.off_bsdf + kGradBsdf * x.bsdf, nm_bar);
        // parameters of the nested BSDF: f = N(pwi, pwo) lp sh + [<wi, wt> > 0] N(rwi, pwo) (1 - lp) sh
        {
            const Its<float> itf = nm_its<float, kCfg>(NmVertex<float>{x.tri, V2f(x.u, x.v), x.uv, nv.valid_dp}, x.shn, dp_du);
            V3f wil = itf.to_local(wi), wol = itf.to_local(wo);
            if (b.two_side) {
                if (signbit_(wil.z)) wol.z = -wol.z;
                wil.z = fabsf(wil.z);
            }
            if (wil.z > 0.f && wol.z > 0.f) {
                V3f wp;
                NmFrame<float> fr;
                nm_setup<float, kCfg>(bl, itf, wp, fr);
                const V3f p_wo = fr.to_local(wol), p_wi = fr.to_local(wil), wt = nm_wt<float>(wp);
                const float sh = nm_G1<float>(wp, wol), lp = nm_lambda_p<float>(wp, wil);
                const BsdfVals nbv = bsdf_vals_at<kCfg>(nbl, x.uv);
                V2f bc_unused(0.f, 0.f);
                bsdf_param_grad<kCfg>(acc, gl, sc, b.nested, nbl, nbv, p_wi.z, p_wo.z, dot(p_wi, p_wo), W, ps * lp * sh, x.uv, xa.uv, x.tri, x.u, x.v, bc_unused);
                if (dot(wil, wt) > 0.f) {
                    const V3f r_wi = fr.to_local(nm_reflect<float>(wil, wt));
                    bsdf_param_grad<kCfg>(acc, gl, sc, b.nested, nbl, nbv, r_wi.z, p_wo.z, dot(r_wi, p_wo), W, ps * (1.f - lp) * sh, x.uv, xa.uv, x.tri, x.u, x.v, bc_unused);
                }
            }
        }
    } else {
        const float ci = dot(wi, x.shn), co = dot(wo, x.shn), cio = dot(wi, wo);
        const BsdfJet j = bsdf_jet<kCfg>(b, x.bv, ci, co, cio, W);
        f = j.f;
        bsdf_param_grad<kCfg>(acc, gl, sc, x.bsdf, b, x.bv, ci, co, cio, W, G * scale, x.uv, xa.uv, x.tri, x.u, x.v, xa.bc);
        const float phi_bar0 = G * scale;
        const float ci_bar = phi_bar0 * j.d_ci, co_bar = phi_bar0 * j.d_co, cio_bar = phi_bar0 * j.d_cio;
        wo_bar_v = x.shn * co_bar + wi * cio_bar;
        wi_bar_v = wo * cio_bar + x.shn * ci_bar;
        shn_bar_v = wo * co_bar + wi * ci_bar;
    }
    r.f = f;
    r.geo = G * scale;
    const float phi = W.x * f.x + W.y * f.y + W.z * f.z;          // sum_c W_c f_c
    // C = phi * G * J * scale
    const float G_bar = phi * scale, J_bar = phi * G * scale;
    // adjoints are often exactly 0 (W = 0): multiply by reciprocals, a zero numerator sends div.rn.f32 down its slow path
    const float inv_t = 1.f / t, inv_t2 = 1.f / (t * t);
    const float cy_bar = G_bar * (cy < 0.f ? -1.f : 1.f) * inv_t2;
    const float t_bar = G_bar * (-2.f * fabsf(cy) * (inv_t2 * inv_t));
    V3f wo_bar = ny * (-cy_bar) + wo_bar_v;
    wo_bar = wo_bar + wo_extra(wo, f * (G * scale));
    r.ny = wo * (-cy_bar);
    r.wi_bar = wi_bar_v;
    xa.shn = xa.shn + shn_bar_v;
    const V3f vec_bar = unit_adj(wo, t, wo_bar, t_bar);
    r.py = vec_bar;
    xa.p = xa.p - vec_bar;
    r.area_y = area_y > 0.f ? J_bar * (1.f / area_y) : 0.f;
    return r;
}

// ---- reverse sweep of one interior path --------------------------------------------------------------
// g = d(loss)/d(lane value) (rgb, already divided by spp).  o/d/dc = camera ray (world origin, world
// direction, camera-space direction).  The loop runs k = nsh-1 .. 0 with a rolling window of three vertices
// (previous, x = vertex k, y = vertex k+1) held in registers; a vertex's adjoint is complete -- and scattered
// -- one iteration after it was x.
// kColloc: CollocatedIntegrator -- the lane's value is ONE event at the primary hit, written as an event towards the camera
// position with "normal" wo there: sum_c g_c f_c(wi, wo = wi) |cos_y| / t^2 * intensity with |cos_y| = 1.
template <int kD, int kCfg, bool kColloc = false>
__device__ __forceinline__ void path_adjoint(const DScene &sc, const GradLayout &gl, const GradAcc &acc, const PathRecord<kD> &R, V3f o, V3f d,
                                             V3f dc, V3f g, bool hide_emitters, bool enabled, V3f oc = V3f(0.f, 0.f, 0.f)) {
    // Called by ALL 32 lanes of the warp from warp-uniform control flow (`enabled` = this lane has a path and a
    // cotangent).  The sweep loop below therefore sits at the top level, runs the warp's maximum trip count and
    // re-converges with a full-mask barrier every iteration; everything lane-specific hangs off plain `if`s.
    // (A barrier over a lane subset inside divergent code becomes a collective wait in SASS that lets the
    // groups through one by one without merging them -- profiles/r01g: 3 of 32 lanes active in the sweep.)
    constexpr bool kFull = (kCfg & kCfgFull) != 0;
    const bool has_v0 = enabled && R.nv > 0;
    const bool sweep = has_v0 && (kColloc || R.nsh > 0);
    // Emitter-radiance gradients are summed per lane in registers and scattered ONCE, from the warp-uniform epilogue:
    // every event of every lane of the warp adds to the same few entries (one emitter in most scenes), and inside the
    // divergent event code that was a same-target add with a partial lane mask -- the slow path of grad_add3_impl
    // (a serial loop of variable-lane shuffles: 9 % of the kernel's instructions, profiles/r02e).
    int em_idx = -1;
    V3f em_acc(0.f, 0.f, 0.f);
    auto add_emitter = [&](int e, V3f v) {
        if (em_idx < 0) em_idx = e;
        if (e == em_idx) em_acc = em_acc + V3f(isfinite(v.x) ? v.x : 0.f, isfinite(v.y) ? v.y : 0.f, isfinite(v.z) ? v.z : 0.f);
        else acc.add3(gl.off_emit + 4 * e, v);           // a second emitter on the same path: rare, scattered directly
    };
    // vertex 0: solid-angle form -- (u, v, t) are functions of the triangle and the camera ray
    TriRec<float> T0;
    float u0 = 0.f, v0 = 0.f, t0 = 0.f;
    VtxGeo v0geo;
    V3f o_bar(0.f, 0.f, 0.f), d_bar(0.f, 0.f, 0.f);
    if (has_v0) {
        T0 = load_tri<float>(sc, R.vtri[0]);
        ray_intersect_triangle<float>(T0.p0, T0.e1, T0.e2, o, d, u0, v0, t0);
        v0geo = vertex_geo<kCfg>(sc, R.vtri[0], u0, v0);
        v0geo.p = V3f(fmaf(d.x, t0, o.x), fmaf(d.y, t0, o.y), fmaf(d.z, t0, o.z));
        // Le at the primary hit
        if (!kColloc && !hide_emitters && v0geo.emitter >= 0) {
            if (kFull && sc.emitters[v0geo.emitter].type == 1) {
                V3f le;
                d_bar = d_bar + env_le_adjoint(acc, gl, sc.env, d, g, le);
                if (R.nsh <= 0) scatter_camera_ray(acc, gl, dc, o_bar, d_bar, oc);
            } else if (dot(-d, v0geo.shn) > 0.f) add_emitter(v0geo.emitter, g);
        }
    }

    const VtxAdj zero_adj = {V3f(0.f, 0.f, 0.f), V3f(0.f, 0.f, 0.f), V3f(0.f, 0.f, 0.f), 0.f, V2f(0.f, 0.f), V2f(0.f, 0.f)};
    const int ktop = R.nsh - 1;
    auto geo_of = [&](int k) { return k == 0 ? v0geo : vertex_geo<kCfg>(sc, R.vtri[k], R.vu[k], R.vv[k]); };
    VtxGeo y = v0geo, x = v0geo;
    if (sweep && !kColloc) {
        if (ktop + 1 < R.nv) y = geo_of(ktop + 1);      // only read when the bounce exists
        x = geo_of(ktop);
    }
    VtxAdj ya = zero_adj, xa = zero_adj, pa = zero_adj;
    if (kColloc && sweep) {
        const V3f wi = -d;
        auto none = [](V3f, V3f) { return V3f(0.f, 0.f, 0.f); };
        const EventAdj ev = event_adjoint<kCfg>(acc, gl, sc, v0geo, wi, o, wi, 0.f, g, sc.colloc_intensity, ya, none);
        // wo = (o - p)/|o - p| and wi = -d: ev.py is d/d(o), ev.wi_bar d/d(wi); ya.p (= -ev.py) follows p = o + t d in the epilogue
        o_bar = o_bar + ev.py;
        d_bar = d_bar - ev.wi_bar;
        // d/d(intensity): the lane's value over the intensity
        const V3f fb = ev.f * ev.geo;
        if (sc.colloc_intensity != 0.f) acc.add(gl.off_cam + 38, (g.x * fb.x + g.y * fb.y + g.z * fb.z) / sc.colloc_intensity);
    }
    V3f Lnext(0.f, 0.f, 0.f);     // R_{k+1}: radiance gathered after vertex k+1 (without E_{k+1})
    // every lane executes iteration kk together, whatever its own k is
    const int iters = __reduce_max_sync(0xffffffffu, (sweep && !kColloc) ? R.nsh : 0);
#pragma unroll 1
    for (int kk = 0;; ++kk) {
#if defined(PSDR_VJP_PHASE_SYNC) && PSDR_VJP_PHASE_SYNC >= 3
        if (!__syncthreads_or(kk < iters)) break;      // CTA-uniform trip count (called by every thread of the CTA)
#else
        if (kk >= iters) break;
        __syncwarp();
#endif
        const int k = ktop - kk;
        if (sweep && k >= 0) {
        VtxGeo prev = k > 0 ? geo_of(k - 1) : v0geo;
        const V3f A = g * R.T[k];
        float tprev = 1.f;
        V3f wi;
        if (k == 0) wi = -d;
        else {
            const V3f vp = prev.p - x.p;
            tprev = norm(vp);
            wi = vp / tprev;
        }
        V3f wi_bar(0.f, 0.f, 0.f);
        V3f Rk(0.f, 0.f, 0.f);
        // The two events of vertex k -- e = 0: BSDF bounce to vertex k+1, e = 1: emitter sampling -- run through ONE
        // copy of event_adjoint (a rolled loop): three inlined copies made the kernel 3x the instruction cache.
#pragma unroll 1
        for (int e = 0; e < 2; ++e) {
            // mode 0: nothing; 1: bounce; 2: area-light sample; 3: environment-map sample
            int mode = 0;
            V3f py, ny, W, envW(0.f, 0.f, 0.f), Ltot(0.f, 0.f, 0.f), Le(0.f, 0.f, 0.f);
            float area_y = 0.f, scale = 0.f;
            bool y_emits = false, y_env = false;
            int emi = -1;
            if (e == 0) {
                if (R.bnc_ok[k] && k + 1 < R.nv) {
                    const V3f wo = normalize(y.p - x.p);
                    V3f E(0.f, 0.f, 0.f);
                    y_env = kFull && y.emitter >= 0 && sc.emitters[y.emitter].type == 1;
                    y_emits = y.emitter >= 0 && (y_env || dot(-wo, y.shn) > 0.f);
                    if (y_env) {
                        const V2f uv = envmap_dir_to_uv<float>(mul3x3<float>(sc.env.from_world, nullptr, wo));
                        E = bitmap_eval_envmap<float>(sc.env.data, nullptr, sc.env.w, sc.env.h, uv) * (sc.env.scale * R.w2[k]);
                    } else if (y_emits) {
                        const DEmitter &em = sc.emitters[y.emitter];
                        E = V3f(em.radiance[0], em.radiance[1], em.radiance[2]) * R.w2[k];
                    }
                    Ltot = E + Lnext;
                    if (R.bnc_nz[k]) {
                        mode = 1;
                        py = y.p; ny = y.fn; area_y = y.area;
                        W = A * Ltot;
                        scale = 1.f / R.pdf0[k];
                        // a direction-dependent emitter at y adds d(E)/d(wo) and the texel / scale gradients
                        if (y_env) envW = A * R.w2[k];
                    }
                }
            } else if (R.nee_ok[k]) {
                const float4 c = __ldg(sc.shade + 3 * R.htri[k] + 2);
                ny = V3f(c.y, c.z, c.w);
                scale = R.w1[k] / R.lpdf[k];
                if (kFull && R.ltri[k] < 0) {
                    // environment-map sample: the position on the bounding box is detached (envmap.cpp:95-101), J = 1;
                    // derivatives flow through the direction (x.p) into the radiance lookup and the geometric term
                    py = V3f(R.la[k], R.lb[k], R.lc[k]);
                    const V3f wod = normalize(py - x.p);
                    const V2f uv = envmap_dir_to_uv<float>(mul3x3<float>(sc.env.from_world, nullptr, wod));
                    Le = bitmap_eval_envmap<float>(sc.env.data, nullptr, sc.env.w, sc.env.h, uv) * sc.env.scale;
                    mode = 3;
                    envW = A;
                    W = A * Le;
                } else {
                    const TriRec<float> TL = load_tri<float>(sc, R.ltri[k]);
                    py = bilinear(TL.p0, TL.e1, TL.e2, V2f(R.la[k], R.lb[k]));
                    area_y = TL.area;
                    emi = sc.meshes[__float_as_int(__ldg(&sc.geo[3 * R.htri[k] + 2].z))].emitter;
                    if (emi >= 0) {
                        const DEmitter &em = sc.emitters[emi];
                        Le = V3f(em.radiance[0], em.radiance[1], em.radiance[2]);
                        mode = 2;
                        W = A * Le;
                    }
                }
            }
            if (mode != 0) {
                const bool with_env = kFull && (mode == 3 || (mode == 1 && y_env));
                auto extra = [&](V3f wo_, V3f fgeo) {
                    if (!with_env) return V3f(0.f, 0.f, 0.f);
                    V3f le;
                    return env_le_adjoint(acc, gl, sc.env, wo_, envW * fgeo, le);
                };
                const EventAdj ev = event_adjoint<kCfg>(acc, gl, sc, x, wi, py, ny, area_y, W, scale, xa, extra);
                wi_bar = wi_bar + ev.wi_bar;
                const V3f fb = ev.f * ev.geo;
                if (mode == 1) {
                    ya.p = ya.p + ev.py;
                    ya.fn = ya.fn + ev.ny;
                    ya.area += ev.area_y;
                    if (y_emits && !y_env) add_emitter(y.emitter, A * fb * R.w2[k]);
                    Rk = Rk + fb * Ltot;
                } else {
                    if (mode == 2) {
                        const int lb = kGradTri * R.ltri[k];
                        acc.add3(lb, ev.py);
                        acc.add3(lb + 3, ev.py * R.la[k]);
                        acc.add3(lb + 6, ev.py * R.lb[k]);
                        acc.add(lb + 9, ev.area_y);
                        acc.add3(kGradTri * R.htri[k] + 19, ev.ny);
                        add_emitter(emi, A * fb);
                    }
                    Rk = Rk + Le * fb;
                }
            }
        }
        // ---- wi at vertex k: through the previous vertex, or the camera ray direction
        if (k == 0) d_bar = d_bar - wi_bar;
        else {
            const V3f vp_bar = unit_adj(wi, tprev, wi_bar, 0.f);
            pa.p = pa.p + vp_bar;
            xa.p = xa.p - vp_bar;
        }
        Lnext = Rk;
        // vertex k+1 has received everything: scatter it, then slide the window down
        if (k + 1 < R.nv) scatter_pinned_vertex(acc, y, ya);
        y = x; ya = xa;
        x = prev; xa = pa;
        pa = zero_adj;
        }
    }
    // ---- epilogue, executed by ALL 32 lanes (lanes without a path carry zeros): vertex 0 (now in y / ya: p = o + t d,
    // sh_n from the differentiable (u, v)), the camera ray and the emitter accumulator.  The 32 lanes are the samples of
    // one pixel, so these adds mostly share their target: with the whole warp present they take the 5-step butterfly
    // of grad_add3_impl instead of its partial-mask loop.
    __syncwarp();
    const unsigned sw = __ballot_sync(0xffffffffu, sweep);
    const unsigned ew = __ballot_sync(0xffffffffu, em_idx >= 0);
    if (sw != 0u) {
        const int leader = __ffs((int) sw) - 1;
        const int tri_l = __shfl_sync(0xffffffffu, y.tri, leader);
        const bool fnrm_l = __shfl_sync(0xffffffffu, (int) y.face_normals, leader) != 0;
        const int tri = sweep ? y.tri : tri_l;                  // idle lanes adopt the leader's target and add zeros
        const bool face_normals = sweep ? y.face_normals : fnrm_l;
        const VtxAdj a0 = sweep ? ya : zero_adj;
        const int b = kGradTri * tri;
        acc.add3(b + 19, a0.fn);
        acc.add(b + 9, a0.area);
        V3f m_bar(0.f, 0.f, 0.f);
        if (face_normals) acc.add3(b + 19, a0.shn);             // scatter_shading_normal, with the branch on a per-lane flag that
        else {                                                  // is uniform whenever the lanes share the triangle
            const float gu = sweep ? y.u : 0.f, gv = sweep ? y.v : 0.f, minv = sweep ? y.minv : 0.f;
            const V3f shn = sweep ? y.shn : V3f(0.f, 0.f, 0.f);
            m_bar = (a0.shn - shn * dot(shn, a0.shn)) * minv;
            acc.add3(b + 10, m_bar * (1.f - gu - gv));
            acc.add3(b + 13, m_bar * gu);
            acc.add3(b + 16, m_bar * gv);
        }
        V3f r_bar(0.f, 0.f, 0.f);
        if (sweep) {
            const ShadeRec<float> N = load_shade<float>(sc, y.tri);
            // uv = uv0 + u duv0 + v duv1 is differentiable at the primary hit (scene.cpp:785-788)
            const float u_bar = dot(N.n1 - N.n0, m_bar) + dot(y.duv0, a0.uv) + a0.bc.x, v_bar = dot(N.n2 - N.n0, m_bar) + dot(y.duv1, a0.uv) + a0.bc.y;
            o_bar = o_bar + a0.p;
            d_bar = d_bar + a0.p * t0;
            const float t_bar = dot(d, a0.p);
            r_bar = isect_adj(T0.e1, T0.e2, d, u_bar, v_bar, t_bar);
            o_bar = o_bar + r_bar;
            d_bar = d_bar + r_bar * t0;
        }
        acc.add3(b, -r_bar);                                    // scatter_isect_tri
        acc.add3(b + 3, r_bar * (-(sweep ? u0 : 0.f)));
        acc.add3(b + 6, r_bar * (-(sweep ? v0 : 0.f)));
        scatter_camera_ray(acc, gl, dc, sweep ? o_bar : V3f(0.f, 0.f, 0.f), sweep ? d_bar : V3f(0.f, 0.f, 0.f), oc);
    }
    if (ew != 0u) {
        const int leader = __ffs((int) ew) - 1;
        const int e_l = __shfl_sync(0xffffffffu, em_idx, leader);
        acc.add3(gl.off_emit + 4 * (em_idx >= 0 ? em_idx : e_l), em_acc);
    }
}

// ---- secondary-edge estimator, reverse mode (forward code: device_path.cuh eval_secondary_edge) ---------
// tangent = value0 * d<n, u2>, u2 = point on the (detached) emitter triangle at the differentiable
// barycentrics of the ray its1.p -> bp0 (reference src/integrator/path.cpp:252-265).
struct SecEdgeAdjoint {
    static constexpr bool enabled = true;
    const float *d_img;     // cotangent image
    float scale;            // tangent_scale / sppse
    GradLayout gl;
    GradAcc acc;
    template <int kCfg>
    __device__ __forceinline__ void tail(const DScene &sc, const DCamera &cam, int pixel, V3f value0, V3f n, int edge, float s, V3d bp0d,
                                         int light_tri, const Its<Dual> &its1, V3f cd, V2f q) const {
        const SecEdgeAdjoint *adj = this;
    const float *gp = adj->d_img + 3 * pixel;
    float w = 0.f;
    {
        const float gv[3] = {__ldg(gp), __ldg(gp + 1), __ldg(gp + 2)}, v0[3] = {value0.x, value0.y, value0.z};
#pragma unroll
        for (int c = 0; c < 3; ++c)
            if (isfinite(v0[c])) w += gv[c] * v0[c];
    }
    w *= adj->scale;
    if (w == 0.f || !isfinite(w)) return;
    const GradAcc &acc = adj->acc;
    const TriRec<float> TL = load_tri<float>(sc, light_tri);
    const V3f p1 = val(its1.p), bp0 = val(bp0d);
    const V3f vec = bp0 - p1;
    const float len = norm(vec);
    const V3f sdir = vec / len;
    float u, v, t;
    ray_intersect_triangle<float>(TL.p0, TL.e1, TL.e2, p1, sdir, u, v, t);
    const float u_bar = w * dot(n, TL.e1), v_bar = w * dot(n, TL.e2);
    const V3f r_bar = isect_adj(TL.e1, TL.e2, sdir, u_bar, v_bar, 0.f);
    scatter_isect_tri(acc, light_tri, u, v, r_bar);
    V3f p1_bar = r_bar;                      // origin of the ray
    const V3f sdir_bar = r_bar * t;
    const V3f vec_bar = unit_adj(sdir, len, sdir_bar, 0.f);
    p1_bar = p1_bar - vec_bar;
    const int eb = adj->gl.off_se + 6 * edge;   // bp0 = edge.p0 + s * edge.e1
    acc.add3(eb, vec_bar);
    acc.add3(eb + 3, vec_bar * s);
    // its1.p = o + t1 d: camera ray re-intersected with the triangle under p1 (solid-angle form)
    const TriRec<float> TP = load_tri<float>(sc, its1.tri);
    const float t1 = val(its1.t);
    V3f o_bar = p1_bar, d_bar = p1_bar * t1;
    const float t1_bar = dot(cd, p1_bar);
    const V3f r1 = isect_adj(TP.e1, TP.e2, cd, 0.f, 0.f, t1_bar);
    scatter_isect_tri(acc, its1.tri, its1.bu, its1.bv, r1);
    o_bar = o_bar + r1;
    d_bar = d_bar + r1 * t1;
    V3f oc, dc;
    camera_ray_local(cam, q, oc, dc);
    scatter_camera_ray(acc, adj->gl, dc, o_bar, d_bar, oc);
    }
};

}  // namespace psdr
