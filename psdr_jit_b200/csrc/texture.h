// Bitmap<3>::eval in environment-map mode and the lat-long direction <-> uv maps, shared by the host
// (cell masses in Scene::configure) and the kernels.  Reference: src/core/bitmap.cpp:46-131 (rotation 0,
// scale 1, translation 0, flip_v = false, envmap_mode = true -- what EnvironmentMap uses,
// src/emitter/envmap.cpp:27,70-72), src/emitter/envmap.cpp:56-73.
#pragma once
#include "pmath.h"

namespace psdr {

template <class S> struct TexelLoad;
template <> struct TexelLoad<float> {
    static PSDR_HD V3f get(const float *data, const float *, int i) { return V3f(data[3 * i], data[3 * i + 1], data[3 * i + 2]); }
};
template <> struct TexelLoad<Dual> {
    static PSDR_HD V3d get(const float *data, const float *ddata, int i) {
        if (!ddata) return V3d(Dual(data[3 * i]), Dual(data[3 * i + 1]), Dual(data[3 * i + 2]));
        return V3d(Dual(data[3 * i], ddata[3 * i]), Dual(data[3 * i + 1], ddata[3 * i + 1]), Dual(data[3 * i + 2], ddata[3 * i + 2]));
    }
};

struct EnvTexelTaps {      // the four taps and weights of one lookup (the adjoint scatters through them)
    int i00, i10, i01, i11;
    float w0x, w1x, w0y, w1y;
};

// uv in [0,1)^2 (already wrapped by the caller as the reference does), w x h bitmap
template <class S> PSDR_HD V3<S> bitmap_eval_envmap(const float *data, const float *ddata, int w, int h, V2<S> uv, EnvTexelTaps *taps = nullptr) {
    uv = V2<S>((uv.x - 0.5f) + 0.5f, (uv.y - 0.5f) + 0.5f);              // rotation by 0 about the centre
    uv.x = uv.x - (float) (0.5 / (double) w);
    uv = V2<S>(uv.x - floor_(uv.x), uv.y - floor_(uv.y));
    uv.x = uv.x * (float) w;
    uv.y = uv.y * (float) (h - 1);
    int px = (int) floorf(val(uv.x)), py = (int) floorf(val(uv.y));
    const S w1x = uv.x - (float) px, w1y = uv.y - (float) py;
    const S w0x = 1.0f - w1x, w0y = 1.0f - w1y;
    const int yw = (py < h - 2 ? py : h - 2) * w;
    const int xp1 = (px + 1) % w;
    const int last = w * h - 1;
    int i00 = yw + px, i10 = yw + xp1, i01 = yw + px + w, i11 = yw + xp1 + w;
    i00 = i00 < last ? i00 : last; i10 = i10 < last ? i10 : last; i01 = i01 < last ? i01 : last; i11 = i11 < last ? i11 : last;
    if (taps) {
        taps->i00 = i00; taps->i10 = i10; taps->i01 = i01; taps->i11 = i11;
        taps->w0x = val(w0x); taps->w1x = val(w1x); taps->w0y = val(w0y); taps->w1y = val(w1y);
    }
    const V3<S> v00 = TexelLoad<S>::get(data, ddata, i00), v10 = TexelLoad<S>::get(data, ddata, i10),
                v01 = TexelLoad<S>::get(data, ddata, i01), v11 = TexelLoad<S>::get(data, ddata, i11);
    const V3<S> v0(fmadd(w0x, v00.x, w1x * v10.x), fmadd(w0x, v00.y, w1x * v10.y), fmadd(w0x, v00.z, w1x * v10.z));
    const V3<S> v1(fmadd(w0x, v01.x, w1x * v11.x), fmadd(w0x, v01.y, w1x * v11.y), fmadd(w0x, v01.z, w1x * v11.z));
    return V3<S>(fmadd(w0y, v0.x, w1y * v1.x), fmadd(w0y, v0.y, w1y * v1.y), fmadd(w0y, v0.z, w1y * v1.z));
}

// A surface texture slot of a BSDF (reference Bitmap<1|3>, src/core/bitmap.cpp): w x h texels of `ch` channels
// (interleaved, pixel = y*w + x) and the uv transform rotate / scale / translate (bitmap.cpp:64-72) with forward tangents.
// w * h == 0: the slot holds a constant (1x1 bitmap) and the BSDF record's scalar fields are used instead.
struct DTex {
    int w, h;
    const float *data, *ddata;
    float cr, sr, d_cr, d_sr;        // cos / sin of m_rot and their tangents
    float scale, d_scale;
    float tx, ty, d_tx, d_ty;
    int goff;                        // offset of the texel gradients, relative to the END of the adjoint's gradient table
    int ch;                          // 1 or 3
};

template <class S> struct Lift2 {
    static PSDR_HD S s(float v, float) { return S(v); }
};
template <> struct Lift2<Dual> {
    static PSDR_HD Dual s(float v, float d) { return Dual(v, d); }
};

// Bitmap<channels>::eval for surface textures (flip_v = true, envmap_mode = false): reference src/core/bitmap.cpp:60-131.
// uv is rotated about the centre, flipped, scaled about the centre, translated, wrapped; bilinear over (w-1) x (h-1) cells.
// Returns rgb for 3-channel data, (x, 0, 0) for 1-channel data.
template <class S> PSDR_HD V3<S> tex_eval_uv(const DTex &t, bool with_texel_tangents, V2<S> uv, EnvTexelTaps *taps = nullptr) {
    const S cr = Lift2<S>::s(t.cr, t.d_cr), sr = Lift2<S>::s(t.sr, t.d_sr), sc = Lift2<S>::s(t.scale, t.d_scale);
    const S ux = uv.x - 0.5f, uy = uv.y - 0.5f;
    S rx = ux * cr + uy * sr, ry = -ux * sr + uy * cr;
    rx = rx + 0.5f;
    ry = -(ry + 0.5f);                                                    // flip v
    rx = rx * sc;
    ry = ry * sc;
    const S off = sc * 0.5f + (-0.5f);                                    // -.5 + scale / 2
    rx = rx - off;
    ry = ry + off;
    rx = rx + Lift2<S>::s(t.tx, t.d_tx);
    ry = ry + Lift2<S>::s(t.ty, t.d_ty);
    uv = V2<S>(rx - floor_(rx), ry - floor_(ry));
    const int w = t.w, h = t.h;
    uv.x = uv.x * (float) (w - 1);
    uv.y = uv.y * (float) (h - 1);
    int px = (int) floorf(val(uv.x)), py = (int) floorf(val(uv.y));
    const S w1x = uv.x - (float) px, w1y = uv.y - (float) py;
    const S w0x = 1.0f - w1x, w0y = 1.0f - w1y;
    px = px < w - 2 ? px : w - 2;
    py = py < h - 2 ? py : h - 2;
    const int i00 = py * w + px, i10 = i00 + 1, i01 = i00 + w, i11 = i01 + 1;
    if (taps) {
        taps->i00 = i00; taps->i10 = i10; taps->i01 = i01; taps->i11 = i11;
        taps->w0x = val(w0x); taps->w1x = val(w1x); taps->w0y = val(w0y); taps->w1y = val(w1y);
    }
    const float *dd = with_texel_tangents ? t.ddata : nullptr;
    if (t.ch == 1) {
        auto get = [&](int i) { return (dd ? Lift2<S>::s(t.data[i], dd[i]) : S(t.data[i])); };
        const S v00 = get(i00), v10 = get(i10), v01 = get(i01), v11 = get(i11);
        const S v0 = fmadd(w0x, v00, w1x * v10), v1 = fmadd(w0x, v01, w1x * v11);
        return V3<S>(fmadd(w0y, v0, w1y * v1), S(0.f), S(0.f));
    }
    const V3<S> v00 = TexelLoad<S>::get(t.data, dd, i00), v10 = TexelLoad<S>::get(t.data, dd, i10),
                v01 = TexelLoad<S>::get(t.data, dd, i01), v11 = TexelLoad<S>::get(t.data, dd, i11);
    const V3<S> v0(fmadd(w0x, v00.x, w1x * v10.x), fmadd(w0x, v00.y, w1x * v10.y), fmadd(w0x, v00.z, w1x * v10.z));
    const V3<S> v1(fmadd(w0x, v01.x, w1x * v11.x), fmadd(w0x, v01.y, w1x * v11.y), fmadd(w0x, v01.z, w1x * v11.z));
    return V3<S>(fmadd(w0y, v0.x, w1y * v1.x), fmadd(w0y, v0.y, w1y * v1.y), fmadd(w0y, v0.z, w1y * v1.z));
}

// lat-long uv of a direction in the map's local frame (envmap.cpp:66-67)
template <class S> PSDR_HD V2<S> envmap_dir_to_uv(V3<S> v) {
    V2<S> uv(atan2_(v.x, -v.z) * 0.15915494309189533577f, safe_acos_(v.y) * kInvPi);
    return V2<S>(uv.x - floor_(uv.x), uv.y - floor_(uv.y));
}

template <class S> PSDR_HD V3<S> mul3x3(const float *M, const float *dM, V3<S> p);
template <> PSDR_HD V3f mul3x3<float>(const float *M, const float *, V3f p) {
    return V3f(fmaf(M[2], p.z, fmaf(M[1], p.y, M[0] * p.x)), fmaf(M[5], p.z, fmaf(M[4], p.y, M[3] * p.x)),
               fmaf(M[8], p.z, fmaf(M[7], p.y, M[6] * p.x)));
}
template <> PSDR_HD V3d mul3x3<Dual>(const float *M, const float *dM, V3d p) {
    Dual t[3];
    for (int i = 0; i < 3; ++i) {
        const Dual a(M[3 * i], dM[3 * i]), b(M[3 * i + 1], dM[3 * i + 1]), c(M[3 * i + 2], dM[3 * i + 2]);
        t[i] = fmadd(c, p.z, fmadd(b, p.y, a * p.x));
    }
    return V3d(t[0], t[1], t[2]);
}

}  // namespace psdr
