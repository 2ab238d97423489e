// psdr_jit_b200 -- fp32 + forward-mode dual-number maths shared by the host scene code and the
// sm_100a kernels.  Arithmetic is written with an explicit operation order (explicit fmaf where a
// fused multiply-add is wanted) and the library is compiled with contraction disabled
// (nvcc -fmad=false, host -ffp-contract=off), so results do not depend on compiler fusion choices
// and can be checked lane-by-lane against the CPU oracle.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define PSDR_HD __host__ __device__ __forceinline__
#else
#define PSDR_HD inline
#endif

namespace psdr {

constexpr float kEpsilon = 1e-5f;        // reference include/psdr/constants.h:12
constexpr float kRayEpsilon = 1e-3f;     // :13
constexpr float kShadowEpsilon = 1e-3f;  // :14
constexpr float kEdgeEpsilon = 1e-5f;    // :15
constexpr float kPi = 3.14159265358979323846f;
constexpr float kInvPi = 0.31830988618379067154f;
constexpr float kTraceTMax = 100000000.f;  // reference src/scene/scene_optix.cpp:376
constexpr float kFltEps = 1.1920929e-07f;

struct Dual {
    float v, d;
    PSDR_HD Dual() : v(0.f), d(0.f) {}
    PSDR_HD Dual(float v_) : v(v_), d(0.f) {}
    PSDR_HD Dual(float v_, float d_) : v(v_), d(d_) {}
};

PSDR_HD float val(float x) { return x; }
PSDR_HD float val(const Dual &x) { return x.v; }
PSDR_HD float tang(float) { return 0.f; }
PSDR_HD float tang(const Dual &x) { return x.d; }
PSDR_HD float detach(float x) { return x; }
PSDR_HD Dual detach(const Dual &x) { return Dual(x.v, 0.f); }

PSDR_HD Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
PSDR_HD Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
PSDR_HD Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
PSDR_HD Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, a.d * b.v + a.v * b.d); }
// The tangent is multiplied by the reciprocal of the denominator instead of divided: tangents are exactly 0 for
// everything the derivative parameter does not move, and div.rn.f32 with a zero numerator takes its out-of-line
// slow path (FCHK) every time; 1 / (a normal number) never does.  The oracle spells it identically.
PSDR_HD Dual operator/(Dual a, Dual b) {
    float q = a.v / b.v;
    return Dual(q, (a.d - q * b.d) * (1.f / b.v));
}
PSDR_HD Dual operator+(Dual a, float b) { return Dual(a.v + b, a.d); }
PSDR_HD Dual operator+(float a, Dual b) { return Dual(a + b.v, b.d); }
PSDR_HD Dual operator-(Dual a, float b) { return Dual(a.v - b, a.d); }
PSDR_HD Dual operator-(float a, Dual b) { return Dual(a - b.v, -b.d); }
PSDR_HD Dual operator*(Dual a, float b) { return Dual(a.v * b, a.d * b); }
PSDR_HD Dual operator*(float a, Dual b) { return Dual(a * b.v, a * b.d); }
PSDR_HD Dual operator/(Dual a, float b) { return Dual(a.v / b, a.d * (1.f / b)); }
PSDR_HD Dual operator/(float a, Dual b) { return Dual(a) / b; }
PSDR_HD Dual &operator+=(Dual &a, Dual b) { a = a + b; return a; }
PSDR_HD Dual &operator*=(Dual &a, Dual b) { a = a * b; return a; }

PSDR_HD float sqrt_(float x) { return sqrtf(x); }
PSDR_HD Dual sqrt_(Dual x) {
    float s = sqrtf(x.v);
    return Dual(s, x.d * (1.f / (2.f * s)));
}
// drjit safe_sqrt: sqrt(max(a,0)), derivative taken at max(a, eps)
PSDR_HD float safe_sqrt(float x) { return sqrtf(fmaxf(x, 0.f)); }
PSDR_HD Dual safe_sqrt(Dual x) {
    float s = sqrtf(fmaxf(x.v, 0.f));
    float sg = sqrtf(fmaxf(x.v, kFltEps));
    return Dual(s, x.d * (1.f / (2.f * sg)));
}
PSDR_HD bool signbit_(float x) {      // std::signbit (the sign bit itself: -0.f counts as negative)
#if defined(__CUDA_ARCH__)
    return __float_as_int(x) < 0;
#else
    return std::signbit(x);
#endif
}
PSDR_HD float abs_(float x) { return fabsf(x); }
PSDR_HD Dual abs_(Dual x) { return Dual(fabsf(x.v), signbit_(x.v) ? -x.d : x.d); }
PSDR_HD float rcp_(float x) { return 1.f / x; }
PSDR_HD Dual rcp_(Dual x) { return 1.f / x; }
PSDR_HD float exp2_(float x) { return exp2f(x); }
PSDR_HD Dual exp2_(Dual x) {
    const float e = exp2f(x.v);
    return Dual(e, 0.69314718055994530942f * e * x.d);
}
PSDR_HD float sqr(float x) { return x * x; }
PSDR_HD Dual sqr(Dual x) { return x * x; }
PSDR_HD float fmadd(float a, float b, float c) { return fmaf(a, b, c); }
PSDR_HD Dual fmadd(Dual a, Dual b, Dual c) { return Dual(fmaf(a.v, b.v, c.v), a.d * b.v + a.v * b.d + c.d); }
PSDR_HD Dual fmadd(Dual a, float b, Dual c) { return Dual(fmaf(a.v, b, c.v), a.d * b + c.d); }
PSDR_HD Dual fmadd(float a, Dual b, Dual c) { return Dual(fmaf(a, b.v, c.v), a * b.d + c.d); }

// sin/cos on [-pi/4, pi/4]: fixed single-precision minimax polynomials (no range reduction is
// needed by the concentric disk map, the only trigonometry on the hot path).
PSDR_HD void sincos_quarter(float x, float &sn, float &cs) {
    float z = x * x;
    float ps = fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f);
    sn = fmaf(ps * z, x, x);
    float pc = fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f);
    cs = fmaf(pc * z, z, fmaf(-0.5f, z, 1.f));
}


// ---- elementary functions with a fixed operation order, shared bit for bit by host and device code
// (library sincosf/atan2f/acosf differ in the last ulp between glibc and libdevice, and the environment
// map turns them into discrete cell / texel choices).  Polynomials: Cephes single precision.
PSDR_HD void sincos_full(float x, float &sn, float &cs) {   // any |x| < ~1e4
    const float kf = rintf(x * 0.63661977236758134308f);    // x / (pi/2)
    const int k = (int) kf;
    float r = fmaf(-kf, 1.5703125f, x);                      // Cody-Waite, pi/2 split in three
    r = fmaf(-kf, 4.837512969970703125e-4f, r);
    r = fmaf(-kf, 7.54978995489188216e-8f, r);
    float s, c;
    sincos_quarter(r, s, c);
    switch (k & 3) {
        case 0: sn = s; cs = c; break;
        case 1: sn = c; cs = -s; break;
        case 2: sn = -s; cs = -c; break;
        default: sn = -c; cs = s; break;
    }
}
PSDR_HD float atan_pos(float x) {   // x >= 0
    float y = 0.f;
    if (x > 2.414213562373095f) { y = 1.5707963267948966f; x = -1.f / x; }
    else if (x > 0.4142135623730950f) { y = 0.7853981633974483f; x = (x - 1.f) / (x + 1.f); }
    const float z = x * x;
    const float p = fmaf(fmaf(fmaf(8.05374449538e-2f, z, -1.38776856032e-1f), z, 1.99777106478e-1f), z, -3.33329491539e-1f);
    return y + fmaf(p * z, x, x);
}
PSDR_HD float atan2_(float y, float x) {
    if (x == 0.f && y == 0.f) return 0.f;
    const float ax = fabsf(x), ay = fabsf(y);
    float a = ax >= ay ? atan_pos(ay / ax) : 1.5707963267948966f - atan_pos(ax / ay);
    if (x < 0.f) a = 3.14159265358979323846f - a;
    return y < 0.f ? -a : a;
}
PSDR_HD float asin_small(float x) {   // |x| <= 0.5
    const float z = x * x;
    const float p = fmaf(fmaf(fmaf(fmaf(4.2163199048e-2f, z, 2.4181311049e-2f), z, 4.5470025998e-2f), z, 7.4953002686e-2f), z, 1.6666752422e-1f);
    return fmaf(p * z, x, x);
}
PSDR_HD float safe_acos_(float x) {   // drjit safe_acos: clamps to [-1, 1]
    x = fminf(fmaxf(x, -1.f), 1.f);
    if (x < -0.5f) return 3.14159265358979323846f - 2.f * asin_small(sqrtf(0.5f * (1.f + x)));
    if (x > 0.5f) return 2.f * asin_small(sqrtf(0.5f * (1.f - x)));
    return 1.5707963267948966f - asin_small(x);
}
PSDR_HD Dual atan2_(Dual y, Dual x) {
    const float r2 = x.v * x.v + y.v * y.v;
    return Dual(atan2_(y.v, x.v), r2 > 0.f ? (x.v * y.d - y.v * x.d) / r2 : 0.f);
}
PSDR_HD Dual safe_acos_(Dual x) {
    const float s = 1.f - x.v * x.v;
    return Dual(safe_acos_(x.v), s > 0.f ? -x.d / sqrtf(s) : 0.f);
}
PSDR_HD float floor_(float x) { return floorf(x); }
PSDR_HD Dual floor_(Dual x) { return Dual(floorf(x.v), 0.f); }

template <class S> struct V2 {
    S x, y;
    PSDR_HD V2() : x(S(0.f)), y(S(0.f)) {}
    PSDR_HD V2(S x_, S y_) : x(x_), y(y_) {}
};
template <class S> struct V3 {
    S x, y, z;
    PSDR_HD V3() : x(S(0.f)), y(S(0.f)), z(S(0.f)) {}
    PSDR_HD V3(S x_, S y_, S z_) : x(x_), y(y_), z(z_) {}
    PSDR_HD explicit V3(S s) : x(s), y(s), z(s) {}
};
using V3f = V3<float>;
using V3d = V3<Dual>;
using V2f = V2<float>;
using V2d = V2<Dual>;

template <class S> PSDR_HD V3<S> operator+(V3<S> a, V3<S> b) { return V3<S>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class S> PSDR_HD V3<S> operator-(V3<S> a, V3<S> b) { return V3<S>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class S> PSDR_HD V3<S> operator-(V3<S> a) { return V3<S>(-a.x, -a.y, -a.z); }
template <class S> PSDR_HD V3<S> operator*(V3<S> a, V3<S> b) { return V3<S>(a.x * b.x, a.y * b.y, a.z * b.z); }
template <class S> PSDR_HD V3<S> operator*(V3<S> a, S b) { return V3<S>(a.x * b, a.y * b, a.z * b); }
// vector / scalar = vector * (1 / scalar): one IEEE division instead of three (Dr.Jit's array / scalar does the same);
// the oracle (oracle/orc_math.h) spells it identically
template <class S> PSDR_HD V3<S> operator/(V3<S> a, S b) {
    const S r = S(1.f) / b;
    return V3<S>(a.x * r, a.y * r, a.z * r);
}
PSDR_HD V3d operator*(V3d a, float b) { return V3d(a.x * b, a.y * b, a.z * b); }
PSDR_HD V3d operator/(V3d a, float b) {
    const float r = 1.f / b;
    return V3d(a.x * r, a.y * r, a.z * r);
}
template <class S> PSDR_HD V2<S> operator+(V2<S> a, V2<S> b) { return V2<S>(a.x + b.x, a.y + b.y); }
template <class S> PSDR_HD V2<S> operator-(V2<S> a, V2<S> b) { return V2<S>(a.x - b.x, a.y - b.y); }

template <class S> PSDR_HD S dot(V3<S> a, V3<S> b) { return fmadd(a.z, b.z, fmadd(a.y, b.y, a.x * b.x)); }
template <class S> PSDR_HD S dot(V2<S> a, V2<S> b) { return fmadd(a.y, b.y, a.x * b.x); }
template <class S> PSDR_HD S squared_norm(V3<S> a) { return dot(a, a); }
template <class S> PSDR_HD S norm(V3<S> a) { return sqrt_(dot(a, a)); }
template <class S> PSDR_HD S norm(V2<S> a) { return sqrt_(dot(a, a)); }
template <class S> PSDR_HD V3<S> normalize(V3<S> a) { return a / norm(a); }
template <class S> PSDR_HD V3<S> cross(V3<S> a, V3<S> b) {
    return V3<S>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// cross product with one rounding less per component (fused multiply-subtract, what Dr.Jit's cross()
// emits: ext/drjit/include/drjit/array_router.h fmsub) -- used by the triangle tests
PSDR_HD V3f cross_fms(V3f a, V3f b) {
    return V3f(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
PSDR_HD V3f detach(V3f a) { return a; }
PSDR_HD V3d detach(V3d a) { return V3d(detach(a.x), detach(a.y), detach(a.z)); }
PSDR_HD V3f val(V3f a) { return a; }
PSDR_HD V3f val(V3d a) { return V3f(a.x.v, a.y.v, a.z.v); }
PSDR_HD V3f tang(V3f) { return V3f(0.f, 0.f, 0.f); }
PSDR_HD V3f tang(V3d a) { return V3f(a.x.d, a.y.d, a.z.d); }
PSDR_HD V2f val(V2f a) { return a; }
PSDR_HD V2f val(V2d a) { return V2f(a.x.v, a.y.v); }

template <class S> struct Lift;
template <> struct Lift<float> {
    static PSDR_HD float s(float v, float) { return v; }
    static PSDR_HD V3f v3(V3f v, V3f) { return v; }
};
template <> struct Lift<Dual> {
    static PSDR_HD Dual s(float v, float d) { return Dual(v, d); }
    static PSDR_HD V3d v3(V3f v, V3f d) { return V3d(Dual(v.x, d.x), Dual(v.y, d.y), Dual(v.z, d.z)); }
};
template <class S> PSDR_HD V3<S> lift3(V3f a) { return V3<S>(S(a.x), S(a.y), S(a.z)); }
template <class S> PSDR_HD V2<S> lift2(V2f a) { return V2<S>(S(a.x), S(a.y)); }

// reference include/psdr/utils.h:64-72
template <class S, class W> PSDR_HD V3<S> bilinear(V3<S> p0, V3<S> e1, V3<S> e2, V2<W> st) {
    return V3<S>(fmadd(e1.x, st.x, fmadd(e2.x, st.y, p0.x)), fmadd(e1.y, st.x, fmadd(e2.y, st.y, p0.y)),
                 fmadd(e1.z, st.x, fmadd(e2.z, st.y, p0.z)));
}
template <class S, class W> PSDR_HD V2<S> bilinear2(V2<S> p0, V2<S> e1, V2<S> e2, V2<W> st) {
    return V2<S>(fmadd(e1.x, st.x, fmadd(e2.x, st.y, p0.x)), fmadd(e1.y, st.x, fmadd(e2.y, st.y, p0.y)));
}

template <class S> struct M4 {
    S m[4][4];
    static PSDR_HD M4 identity() {
        M4 r;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) r.m[i][j] = S(i == j ? 1.f : 0.f);
        return r;
    }
};
template <class S> PSDR_HD M4<S> operator*(const M4<S> &a, const M4<S> &b) {
    M4<S> r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            S acc = a.m[i][0] * b.m[0][j];
            for (int k = 1; k < 4; ++k) acc = fmadd(a.m[i][k], b.m[k][j], acc);
            r.m[i][j] = acc;
        }
    return r;
}
// reference include/psdr/core/transform.h:117-128
template <class S> PSDR_HD V3<S> transform_pos(const M4<S> &M, V3<S> p) {
    S t[4];
    for (int i = 0; i < 4; ++i) t[i] = fmadd(M.m[i][2], p.z, fmadd(M.m[i][1], p.y, M.m[i][0] * p.x)) + M.m[i][3];
    return V3<S>(t[0], t[1], t[2]) / t[3];      // one reciprocal (vector / scalar)
}
template <class S> PSDR_HD V3<S> transform_dir(const M4<S> &M, V3<S> p) {
    S t[3];
    for (int i = 0; i < 3; ++i) t[i] = fmadd(M.m[i][2], p.z, fmadd(M.m[i][1], p.y, M.m[i][0] * p.x));
    return V3<S>(t[0], t[1], t[2]);
}

// ---- PCG32 + 64-bit-lane TEA seeding (reference src/core/sampler.cpp:6-30,
//      ext/drjit/include/drjit/random.h:55-75,130-132) ----------------------------------------
PSDR_HD uint64_t sample_tea_64(uint64_t v0, uint64_t v1) {
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cull) ^ (v1 + (uint64_t) sum) ^ ((v1 >> 5) + 0xc8013ea4ull);
        v1 += ((v0 << 4) + 0xad90777dull) ^ (v0 + (uint64_t) sum) ^ ((v0 >> 5) + 0x7e95761eull);
    }
    return v0 + (v1 << 32);
}

constexpr uint64_t kPcgMult = 0x5851f42d4c957f2dull;

struct Pcg32 {
    uint64_t state, inc;
    PSDR_HD uint32_t next_u32() {
        uint64_t old = state;
        state = old * kPcgMult + inc;
        uint32_t xs = (uint32_t) (((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t) (old >> 59u);
        return (xs >> rot) | (xs << ((0u - rot) & 31u));
    }
    PSDR_HD float next_1d() {
        uint32_t b = (next_u32() >> 9) | 0x3f800000u;
#if defined(__CUDA_ARCH__)
        return __uint_as_float(b) - 1.f;
#else
        union { uint32_t u; float f; } c;
        c.u = b;
        return c.f - 1.f;
#endif
    }
    // Sampler::seed(seed_value) for lane idx of the seeded array
    PSDR_HD void seed(uint64_t seed_value, uint64_t idx) {
        seed_value += 0x853c49e6748fea9bull;
        uint64_t initstate = sample_tea_64(seed_value, idx), initseq = sample_tea_64(idx, seed_value);
        state = 0;
        inc = (initseq << 1) | 1u;
        next_u32();
        state += initstate;
        next_u32();
    }
    // jump ahead by `delta` draws (O(log delta)); used to continue a stream when seed == -1
    PSDR_HD void advance(uint64_t delta) {
        uint64_t cur_mult = kPcgMult, cur_plus = inc, acc_mult = 1u, acc_plus = 0u;
        while (delta > 0) {
            if (delta & 1) {
                acc_mult *= cur_mult;
                acc_plus = acc_plus * cur_mult + cur_plus;
            }
            cur_plus = (cur_mult + 1) * cur_plus;
            cur_mult *= cur_mult;
            delta >>= 1;
        }
        state = acc_mult * state + acc_plus;
    }
};

}  // namespace psdr
