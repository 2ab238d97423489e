// TEST INFRASTRUCTURE ONLY -- CPU oracle for the psdr-jit PathTracer hot path.
// Nothing under oracle/ may be imported, linked or executed by the product path
// (psdr_jit_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, and only as the checker.
//
// Scalar + dual-number (forward-mode) maths.  `Dual` restates what Dr.Jit's
// DiffArray gives the reference in forward mode: value + one tangent; detach() drops
// the tangent (reference: drjit::detach everywhere in src/).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace orc {

constexpr float kEpsilon = 1e-5f;        // reference include/psdr/constants.h:12
constexpr float kRayEpsilon = 1e-3f;     // :13
constexpr float kShadowEpsilon = 1e-3f;  // :14
constexpr float kEdgeEpsilon = 1e-5f;    // :15
constexpr float kPi = 3.14159265358979323846f;
constexpr float kInvPi = 0.31830988618379067154f;

struct Dual {
    float v = 0.f, d = 0.f;
    Dual() = default;
    Dual(float v_) : v(v_), d(0.f) {}
    Dual(float v_, float d_) : v(v_), d(d_) {}
};

inline float val(float x) { return x; }
inline float val(const Dual &x) { return x.v; }
inline float tan_(float) { return 0.f; }
inline float tan_(const Dual &x) { return x.d; }
inline float detach(float x) { return x; }
inline Dual detach(const Dual &x) { return Dual(x.v, 0.f); }

inline Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
inline Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
inline Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
inline Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, a.d * b.v + a.v * b.d); }
// tangent * (1 / denominator): same spelling as psdr_jit_b200/csrc/pmath.h (see the note there)
inline Dual operator/(Dual a, Dual b) {
    float q = a.v / b.v;
    return Dual(q, (a.d - q * b.d) * (1.f / b.v));
}
inline Dual &operator+=(Dual &a, Dual b) { a = a + b; return a; }
inline Dual &operator-=(Dual &a, Dual b) { a = a - b; return a; }
inline Dual &operator*=(Dual &a, Dual b) { a = a * b; return a; }
inline Dual &operator/=(Dual &a, Dual b) { a = a / b; return a; }
inline Dual operator+(Dual a, float b) { return Dual(a.v + b, a.d); }
inline Dual operator+(float a, Dual b) { return Dual(a + b.v, b.d); }
inline Dual operator-(Dual a, float b) { return Dual(a.v - b, a.d); }
inline Dual operator-(float a, Dual b) { return Dual(a - b.v, -b.d); }
inline Dual operator*(Dual a, float b) { return Dual(a.v * b, a.d * b); }
inline Dual operator*(float a, Dual b) { return Dual(a * b.v, a * b.d); }
inline Dual operator/(Dual a, float b) { return Dual(a.v / b, a.d * (1.f / b)); }
inline Dual operator/(float a, Dual b) { return Dual(a) / b; }

inline float sqrt_(float x) { return std::sqrt(x); }
inline Dual sqrt_(Dual x) {
    float s = std::sqrt(x.v);
    return Dual(s, x.d * (1.f / (2.f * s)));
}
// drjit safe_sqrt (ext/drjit/include/drjit/array_router.h:1987): sqrt(max(a,0)); the
// derivative is taken at max(a, eps).
inline float safe_sqrt(float x) { return std::sqrt(std::fmax(x, 0.f)); }
inline Dual safe_sqrt(Dual x) {
    float s = std::sqrt(std::fmax(x.v, 0.f));
    float sg = std::sqrt(std::fmax(x.v, std::numeric_limits<float>::epsilon()));
    return Dual(s, x.d * (1.f / (2.f * sg)));
}
inline float abs_(float x) { return std::fabs(x); }
inline Dual abs_(Dual x) { return Dual(std::fabs(x.v), std::signbit(x.v) ? -x.d : x.d); }
inline float rcp_(float x) { return 1.f / x; }
inline Dual rcp_(Dual x) { return 1.f / x; }
inline float exp2_(float x) { return std::exp2(x); }
inline Dual exp2_(Dual x) {
    float e = std::exp2(x.v);
    return Dual(e, 0.69314718055994530942f * e * x.d);
}
inline float sqr(float x) { return x * x; }
inline Dual sqr(Dual x) { return x * x; }
// fmadd(a,b,c) = a*b+c with a single rounding on the value (drjit emits fma.rn.ftz)
inline float fmadd(float a, float b, float c) { return std::fmaf(a, b, c); }
inline Dual fmadd(Dual a, Dual b, Dual c) { return Dual(std::fmaf(a.v, b.v, c.v), a.d * b.v + a.v * b.d + c.d); }
inline Dual fmadd(Dual a, float b, Dual c) { return Dual(std::fmaf(a.v, b, c.v), a.d * b + c.d); }
inline Dual fmadd(float a, Dual b, Dual c) { return Dual(std::fmaf(a, b.v, c.v), a * b.d + c.d); }
inline float powf_(float x, float e) { return std::pow(x, e); }

// sin/cos on [-pi/4, pi/4] (Cephes single-precision minimax polynomials, fixed fma order)
inline void sincos_quarter(float x, float &sn, float &cs) {
    float z = x * x;
    float ps = std::fmaf(std::fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f);
    sn = std::fmaf(ps * z, x, x);
    float pc = std::fmaf(std::fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f);
    cs = std::fmaf(pc * z, z, std::fmaf(-0.5f, z, 1.f));
}

template <class S> inline S select(bool m, S a, S b) { return m ? a : b; }


// ---- elementary functions with a fixed operation order (the CUDA side evaluates the same polynomials,
// so cell / texel choices of the environment map agree lane by lane)
// (library sincosf/atan2f/acosf differ in the last ulp between glibc and libdevice, and the environment
// map turns them into discrete cell / texel choices).  Polynomials: Cephes single precision.
inline void sincos_full(float x, float &sn, float &cs) {   // any |x| < ~1e4
    const float kf = std::rint(x * 0.63661977236758134308f);    // x / (pi/2)
    const int k = (int) kf;
    float r = std::fmaf(-kf, 1.5703125f, x);                      // Cody-Waite, pi/2 split in three
    r = std::fmaf(-kf, 4.837512969970703125e-4f, r);
    r = std::fmaf(-kf, 7.54978995489188216e-8f, r);
    float s, c;
    sincos_quarter(r, s, c);
    switch (k & 3) {
        case 0: sn = s; cs = c; break;
        case 1: sn = c; cs = -s; break;
        case 2: sn = -s; cs = -c; break;
        default: sn = -c; cs = s; break;
    }
}
inline float atan_pos(float x) {   // x >= 0
    float y = 0.f;
    if (x > 2.414213562373095f) { y = 1.5707963267948966f; x = -1.f / x; }
    else if (x > 0.4142135623730950f) { y = 0.7853981633974483f; x = (x - 1.f) / (x + 1.f); }
    const float z = x * x;
    const float p = std::fmaf(std::fmaf(std::fmaf(8.05374449538e-2f, z, -1.38776856032e-1f), z, 1.99777106478e-1f), z, -3.33329491539e-1f);
    return y + std::fmaf(p * z, x, x);
}
inline float atan2_(float y, float x) {
    if (x == 0.f && y == 0.f) return 0.f;
    const float ax = std::fabs(x), ay = std::fabs(y);
    float a = ax >= ay ? atan_pos(ay / ax) : 1.5707963267948966f - atan_pos(ax / ay);
    if (x < 0.f) a = 3.14159265358979323846f - a;
    return y < 0.f ? -a : a;
}
inline float asin_small(float x) {   // |x| <= 0.5
    const float z = x * x;
    const float p = std::fmaf(std::fmaf(std::fmaf(std::fmaf(4.2163199048e-2f, z, 2.4181311049e-2f), z, 4.5470025998e-2f), z, 7.4953002686e-2f), z, 1.6666752422e-1f);
    return std::fmaf(p * z, x, x);
}
inline float safe_acos_(float x) {   // drjit safe_acos: clamps to [-1, 1]
    x = std::fmin(std::fmax(x, -1.f), 1.f);
    if (x < -0.5f) return 3.14159265358979323846f - 2.f * asin_small(std::sqrt(0.5f * (1.f + x)));
    if (x > 0.5f) return 2.f * asin_small(std::sqrt(0.5f * (1.f - x)));
    return 1.5707963267948966f - asin_small(x);
}
inline Dual atan2_(Dual y, Dual x) {
    const float r2 = x.v * x.v + y.v * y.v;
    return Dual(atan2_(y.v, x.v), r2 > 0.f ? (x.v * y.d - y.v * x.d) / r2 : 0.f);
}
inline Dual safe_acos_(Dual x) {
    const float s = 1.f - x.v * x.v;
    return Dual(safe_acos_(x.v), s > 0.f ? -x.d / std::sqrt(s) : 0.f);
}
inline float floor_(float x) { return std::floor(x); }
inline Dual floor_(Dual x) { return Dual(std::floor(x.v), 0.f); }

template <class S> struct V2 {
    S x{}, y{};
    V2() = default;
    V2(S x_, S y_) : x(x_), y(y_) {}
};
template <class S> struct V3 {
    S x{}, y{}, z{};
    V3() = default;
    V3(S x_, S y_, S z_) : x(x_), y(y_), z(z_) {}
    explicit V3(S s) : x(s), y(s), z(s) {}
    S &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    const S &operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
using V3f = V3<float>;
using V3d = V3<Dual>;
using V2f = V2<float>;
using V2d = V2<Dual>;

template <class S> inline V3<S> operator+(V3<S> a, V3<S> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class S> inline V3<S> operator-(V3<S> a, V3<S> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class S> inline V3<S> operator-(V3<S> a) { return {-a.x, -a.y, -a.z}; }
template <class S> inline V3<S> operator*(V3<S> a, V3<S> b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
template <class S> inline V3<S> operator*(V3<S> a, S b) { return {a.x * b, a.y * b, a.z * b}; }
template <class S> inline V3<S> operator*(S b, V3<S> a) { return {a.x * b, a.y * b, a.z * b}; }
// vector / scalar = vector * (1 / scalar) (Dr.Jit's array / scalar); same spelling as psdr_jit_b200/csrc/pmath.h
template <class S> inline V3<S> operator/(V3<S> a, S b) {
    const S r = S(1.f) / b;
    return {a.x * r, a.y * r, a.z * r};
}
inline V3d operator*(V3d a, float b) { return {a.x * b, a.y * b, a.z * b}; }
inline V3d operator/(V3d a, float b) {
    const float r = 1.f / b;
    return {a.x * r, a.y * r, a.z * r};
}
template <class S> inline V3<S> &operator+=(V3<S> &a, V3<S> b) { a = a + b; return a; }
template <class S> inline V3<S> &operator*=(V3<S> &a, V3<S> b) { a = a * b; return a; }
template <class S> inline V2<S> operator+(V2<S> a, V2<S> b) { return {a.x + b.x, a.y + b.y}; }
template <class S> inline V2<S> operator-(V2<S> a, V2<S> b) { return {a.x - b.x, a.y - b.y}; }
template <class S> inline V2<S> operator*(V2<S> a, S b) { return {a.x * b, a.y * b}; }

// drjit dot(): x*y fused left to right: fmadd(a.z,b.z, fmadd(a.y,b.y, a.x*b.x))
template <class S> inline S dot(V3<S> a, V3<S> b) { return fmadd(a.z, b.z, fmadd(a.y, b.y, a.x * b.x)); }
template <class S> inline S dot(V2<S> a, V2<S> b) { return fmadd(a.y, b.y, a.x * b.x); }
template <class S> inline S squared_norm(V3<S> a) { return dot(a, a); }
template <class S> inline S norm(V3<S> a) { return sqrt_(dot(a, a)); }
template <class S> inline S norm(V2<S> a) { return sqrt_(dot(a, a)); }
template <class S> inline V3<S> normalize(V3<S> a) { return a / norm(a); }
// fused multiply-subtract cross product (Dr.Jit's cross() emits fmsub); used by the triangle tests
inline V3<float> cross_fms(V3<float> a, V3<float> b) {
    return V3<float>(std::fmaf(a.y, b.z, -(a.z * b.y)), std::fmaf(a.z, b.x, -(a.x * b.z)), std::fmaf(a.x, b.y, -(a.y * b.x)));
}
template <class S> inline V3<S> cross(V3<S> a, V3<S> b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline V3f detach(V3f a) { return a; }
inline V3d detach(V3d a) { return {detach(a.x), detach(a.y), detach(a.z)}; }
inline V3f val(V3f a) { return a; }
inline V3f val(V3d a) { return {a.x.v, a.y.v, a.z.v}; }
inline V3f tan_(V3d a) { return {a.x.d, a.y.d, a.z.d}; }
inline V3f tan_(V3f) { return {0.f, 0.f, 0.f}; }
inline V2f val(V2f a) { return a; }
inline V2f val(V2d a) { return {a.x.v, a.y.v}; }
template <class S> inline V3<S> lift(V3f a) { return {S(a.x), S(a.y), S(a.z)}; }
template <class S> inline V2<S> lift(V2f a) { return {S(a.x), S(a.y)}; }

// reference include/psdr/utils.h:64-72
template <class S, class W> inline V3<S> bilinear(V3<S> p0, V3<S> e1, V3<S> e2, V2<W> st) {
    return {fmadd(e1.x, st.x, fmadd(e2.x, st.y, p0.x)), fmadd(e1.y, st.x, fmadd(e2.y, st.y, p0.y)),
            fmadd(e1.z, st.x, fmadd(e2.z, st.y, p0.z))};
}
template <class S, class W> inline V2<S> bilinear2(V2<S> p0, V2<S> e1, V2<S> e2, V2<W> st) {
    return {fmadd(e1.x, st.x, fmadd(e2.x, st.y, p0.x)), fmadd(e1.y, st.x, fmadd(e2.y, st.y, p0.y))};
}

template <class S> struct M4 {
    S m[4][4];
    static M4 identity() {
        M4 r;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) r.m[i][j] = S(i == j ? 1.f : 0.f);
        return r;
    }
};
template <class S> inline M4<S> operator*(const M4<S> &a, const M4<S> &b) {
    M4<S> r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            S acc = a.m[i][0] * b.m[0][j];
            for (int k = 1; k < 4; ++k) acc = fmadd(a.m[i][k], b.m[k][j], acc);
            r.m[i][j] = acc;
        }
    return r;
}
template <class S> inline M4<S> lift(const M4<float> &a) {
    M4<S> r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r.m[i][j] = S(a.m[i][j]);
    return r;
}
inline M4<float> val(const M4<Dual> &a) {
    M4<float> r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r.m[i][j] = a.m[i][j].v;
    return r;
}
inline M4<float> val(const M4<float> &a) { return a; }

// General 4x4 inverse (Gauss-Jordan with partial pivoting on values; works on duals).
template <class S> inline M4<S> inverse(const M4<S> &a) {
    S w[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            w[i][j] = a.m[i][j];
            w[i][j + 4] = S(i == j ? 1.f : 0.f);
        }
    for (int c = 0; c < 4; ++c) {
        int p = c;
        for (int r = c + 1; r < 4; ++r)
            if (std::fabs(val(w[r][c])) > std::fabs(val(w[p][c]))) p = r;
        if (p != c)
            for (int j = 0; j < 8; ++j) std::swap(w[p][j], w[c][j]);
        S inv = S(1.f) / w[c][c];
        for (int j = 0; j < 8; ++j) w[c][j] = w[c][j] * inv;
        for (int r = 0; r < 4; ++r)
            if (r != c) {
                S f = w[r][c];
                if (val(f) == 0.f && tan_(f) == 0.f) continue;
                for (int j = 0; j < 8; ++j) w[r][j] = w[r][j] - f * w[c][j];
            }
    }
    M4<S> r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r.m[i][j] = w[i][j + 4];
    return r;
}

// reference include/psdr/core/transform.h:117-128
template <class S> inline V3<S> transform_pos(const M4<S> &M, V3<S> p) {
    S t[4];
    for (int i = 0; i < 4; ++i) t[i] = fmadd(M.m[i][2], p.z, fmadd(M.m[i][1], p.y, M.m[i][0] * p.x)) + M.m[i][3];
    return V3<S>{t[0], t[1], t[2]} / t[3];      // one reciprocal (vector / scalar)
}
template <class S> inline V3<S> transform_dir(const M4<S> &M, V3<S> p) {
    S t[3];
    for (int i = 0; i < 3; ++i) t[i] = fmadd(M.m[i][2], p.z, fmadd(M.m[i][1], p.y, M.m[i][0] * p.x));
    return {t[0], t[1], t[2]};
}

// ---- PCG32 + TEA-64 seeding (reference src/core/sampler.cpp:6-30,
//      ext/drjit/include/drjit/random.h:55-75,130-132) -------------------------------
inline uint64_t sample_tea_64(uint64_t v0, uint64_t v1, int rounds = 4) {
    uint32_t sum = 0;  // UIntC: 32-bit wrap-around
    for (int i = 0; i < rounds; ++i) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cull) ^ (v1 + (uint64_t) sum) ^ ((v1 >> 5) + 0xc8013ea4ull);
        v1 += ((v0 << 4) + 0xad90777dull) ^ (v0 + (uint64_t) sum) ^ ((v0 >> 5) + 0x7e95761eull);
    }
    return v0 + (v1 << 32);
}

struct Pcg32 {
    uint64_t state = 0, inc = 0;
    uint32_t next_u32() {
        uint64_t old = state;
        state = old * 0x5851f42d4c957f2dull + inc;
        uint32_t xs = (uint32_t) (((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t) (old >> 59u);
        return (xs >> rot) | (xs << ((-(int32_t) rot) & 31));
    }
    void seed(uint64_t initstate, uint64_t initseq) {
        state = 0;
        inc = (initseq << 1) | 1u;
        next_u32();
        state += initstate;
        next_u32();
    }
    float next_1d() {
        uint32_t b = (next_u32() >> 9) | 0x3f800000u;
        float f;
        std::memcpy(&f, &b, 4);
        return f - 1.f;
    }
};

// Sampler::seed(seed_value) for lane `idx` of the seeded array (sampler.cpp:19-30)
inline Pcg32 make_sampler(uint64_t seed_value, uint64_t idx) {
    seed_value += 0x853c49e6748fea9bull;  // m_base_seed = PCG32_DEFAULT_STATE (sampler.h:38)
    Pcg32 r;
    r.seed(sample_tea_64(seed_value, idx), sample_tea_64(idx, seed_value));
    return r;
}

}  // namespace orc
