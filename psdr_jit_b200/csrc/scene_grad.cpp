// Reverse mode of Scene::configure: maps the gradient table the adjoint kernels produce (per-triangle
// records, primary/secondary edge records, camera matrices, material parameters) back to the parameters
// a user owns through Scene.param_map: object-space vertices, the three to_world factors of every mesh
// and sensor, reflectances, radiances.  The reference gets this from Dr.Jit's AD graph of
// Mesh::configure / PerspectiveCamera::configure (src/shape/mesh.cpp:23-62,317-382,
// src/sensor/perspective.cpp:10-152); here it is the hand-written transpose of scene.cpp, in double.
#include <cmath>
#include <cstring>
#include <stdexcept>

#include "scene.h"

namespace psdr {

namespace {

struct D3 {
    double x = 0, y = 0, z = 0;
};
inline D3 operator+(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline D3 operator-(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline D3 operator*(D3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline D3 cross(D3 a, D3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline D3 d3(const V3d &v) { return {v.x.v, v.y.v, v.z.v}; }
inline D3 d3(const float *p) { return {p[0], p[1], p[2]}; }

struct Mat {
    double m[4][4];
};
Mat mat_of(const M4<Dual> &a) {
    Mat r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r.m[i][j] = a.m[i][j].v;
    return r;
}
Mat mat_of(const M4<float> &a) {
    Mat r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r.m[i][j] = a.m[i][j];
    return r;
}
Mat zero() {
    Mat r;
    std::memset(&r, 0, sizeof(r));
    return r;
}
Mat mul(const Mat &a, const Mat &b) {
    Mat r = zero();
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            for (int k = 0; k < 4; ++k) r.m[i][j] += a.m[i][k] * b.m[k][j];
    return r;
}
Mat transpose(const Mat &a) {
    Mat r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r.m[i][j] = a.m[j][i];
    return r;
}
Mat inverse(const Mat &a) {
    double w[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            w[i][j] = a.m[i][j];
            w[i][4 + j] = i == j ? 1.0 : 0.0;
        }
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        for (int r = c + 1; r < 4; ++r)
            if (std::fabs(w[r][c]) > std::fabs(w[piv][c])) piv = r;
        if (w[piv][c] == 0.0) throw std::runtime_error("singular transform in backprop");
        for (int j = 0; j < 8; ++j) std::swap(w[c][j], w[piv][j]);
        const double inv = 1.0 / w[c][c];
        for (int j = 0; j < 8; ++j) w[c][j] *= inv;
        for (int r = 0; r < 4; ++r)
            if (r != c) {
                const double f = w[r][c];
                for (int j = 0; j < 8; ++j) w[r][j] -= f * w[c][j];
            }
    }
    Mat r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r.m[i][j] = w[i][4 + j];
    return r;
}

// M = L * R * Rt: gradients of the three factors from the gradient of the product
void split_product(const M4<Dual> f[3], const Mat &gM, double out[3][16]) {
    const Mat L = mat_of(f[0]), R = mat_of(f[1]), Rt = mat_of(f[2]);
    const Mat gL = mul(gM, transpose(mul(R, Rt)));
    const Mat gR = mul(mul(transpose(L), gM), transpose(Rt));
    const Mat gRt = mul(transpose(mul(L, R)), gM);
    const Mat *g[3] = {&gL, &gR, &gRt};
    for (int k = 0; k < 3; ++k)
        for (int i = 0; i < 16; ++i) out[k][i] += g[k]->m[i / 4][i % 4];
}

// q = (M [p;1]).xyz / w  -- accumulates d/dp and d/dM from d/dq
void transform_pos_adj(const Mat &M, D3 p, D3 gq, D3 &gp, Mat &gM) {
    const double h[4] = {p.x, p.y, p.z, 1.0};
    double t[4];
    for (int i = 0; i < 4; ++i) t[i] = M.m[i][0] * h[0] + M.m[i][1] * h[1] + M.m[i][2] * h[2] + M.m[i][3];
    const double iw = 1.0 / t[3];
    const double q[3] = {t[0] * iw, t[1] * iw, t[2] * iw};
    const double gt[4] = {gq.x * iw, gq.y * iw, gq.z * iw, -(gq.x * q[0] + gq.y * q[1] + gq.z * q[2]) * iw};
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) gM.m[i][j] += gt[i] * h[j];
    gp.x += M.m[0][0] * gt[0] + M.m[1][0] * gt[1] + M.m[2][0] * gt[2] + M.m[3][0] * gt[3];
    gp.y += M.m[0][1] * gt[0] + M.m[1][1] * gt[1] + M.m[2][1] * gt[2] + M.m[3][1] * gt[3];
    gp.z += M.m[0][2] * gt[0] + M.m[1][2] * gt[1] + M.m[2][2] * gt[2] + M.m[3][2] * gt[3];
}

}  // namespace

// gradient block of a textured slot: its texels, then 4 floats for the bitmap's uv transform (scale, rotation, translation)
static int tex_block_floats(const HBsdf &b, int slot) {
    const HBsdf::Tex &t = b.tex[slot];
    return t.w > 0 ? HBsdf::tex_channels(slot) * t.w * t.h + 4 : 0;
}

int Scene::pervertex_grad_offset(int bsdf) const {
    int back = 0;
    for (int i = num_bsdf_records() - 1; i >= bsdf; --i) back += (int) bsdf_record(i).pv.size();
    return -back;
}

int Scene::texture_grad_offset(int bsdf, int slot) const {
    // tail of the table: [.. secondary edges | envmap block | textures of BSDF 0 (slots 0, 1, 2), BSDF 1, ... | per-vertex
    // tables of BSDF 0, 1, ...]; the sensor-dependent primary-edge block sits before, so the offset is taken relative to
    // the END of the table
    int back = -pervertex_grad_offset(0);
    for (int i = num_bsdf_records() - 1; i >= bsdf; --i)
        for (int k = 2; k >= (i == bsdf ? slot : 0); --k) back += tex_block_floats(bsdf_record(i), k);
    return -back;   // negative: relative to GradLayout::total
}

GradLayout Scene::grad_layout(int sensor) const {
    GradLayout gl{};
    int ntris = 0;
    for (const HMesh &m : meshes) ntris += (int) m.tris.size();
    gl.base = nullptr;
    gl.off_bsdf = kGradTri * ntris;
    gl.off_emit = gl.off_bsdf + kGradBsdf * num_bsdf_records();
    gl.off_cam = gl.off_emit + 4 * (int) emitters.size();
    gl.off_pe = gl.off_cam + kGradCam;
    const int npe = (sensor >= 0 && sensor < (int) cameras.size()) ? (int) cameras[sensor].edges.size() : 0;
    gl.off_se = gl.off_pe + 4 * npe;
    gl.off_env = gl.off_se + 6 * (int) sec_edges.size();
    gl.total = gl.off_env + (env.present ? kGradEnvHead + 3 * env.w * env.h : 0);
    for (int i = 0; i < num_bsdf_records(); ++i)
        for (int k = 0; k < 3; ++k) gl.total += tex_block_floats(bsdf_record(i), k);      // same order as texture_grad_offset()
    for (int i = 0; i < num_bsdf_records(); ++i) gl.total += (int) bsdf_record(i).pv.size();
    return gl;
}

void Scene::backprop(const float *table, const GradLayout &gl, int sensor) {
    grads = ParamGrads{};
    grads.meshes.resize(meshes.size());
    grads.cameras.resize(cameras.size());
    grads.bsdf_refl.assign(3 * (size_t) num_bsdf_records(), 0.0);
    grads.emitter_rad.assign(3 * emitters.size(), 0.0);
    for (auto &c : grads.cameras) std::memset(c.to_world, 0, sizeof(c.to_world));

    // world-space vertex gradients per mesh
    std::vector<std::vector<D3>> gvw(meshes.size());
    for (size_t mi = 0; mi < meshes.size(); ++mi) gvw[mi].assign(meshes[mi].v_world.size(), D3{});

    // ---- primary edges of the rendered sensor + camera matrices
    Mat gW = zero(), gM = zero();   // d world_to_sample, d to_world (full product)
    const HCamera &cam = cameras[sensor];
    for (int i = 0; i < 16; ++i) {
        gM.m[i / 4][i % 4] += table[gl.off_cam + i];
        gW.m[i / 4][i % 4] += table[gl.off_cam + 16 + i];
    }
    const Mat W = mat_of(cam.world_to_sample);
    for (size_t e = 0; e < cam.edges.size(); ++e) {
        const HPrimEdge &pe = cam.edges[e];
        const float *g = table + gl.off_pe + 4 * e;
        transform_pos_adj(W, d3(meshes[pe.mesh].v_world[pe.v0]), D3{g[0], g[1], 0.0}, gvw[pe.mesh][pe.v0], gW);
        transform_pos_adj(W, d3(meshes[pe.mesh].v_world[pe.v1]), D3{g[2], g[3], 0.0}, gvw[pe.mesh][pe.v1], gW);
    }
    {   // world_to_sample = C * inverse(M)  =>  dM = -M^{-T} (C^T dW) M^{-T}
        const Mat C = mat_of(cam.camera_to_sample), Minv = inverse(mat_of(cam.to_world_full));
        const Mat gInv = mul(transpose(C), gW);
        const Mat MinvT = transpose(Minv);
        const Mat t = mul(mul(MinvT, gInv), MinvT);
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) gM.m[i][j] -= t.m[i][j];
        // pos = M[:, 3] (w = 1), dir = M[:3, :3] * (0, 0, 1)
        for (int i = 0; i < 3; ++i) {
            gM.m[i][3] += table[gl.off_cam + 32 + i];
            gM.m[i][2] += table[gl.off_cam + 35 + i];
        }
        split_product(cam.to_world, gM, grads.cameras[sensor].to_world);
    }

    // ---- secondary edges: p0 = V[v0], e1 = V[v1] - V[v0]
    for (size_t e = 0; e < sec_edges.size(); ++e) {
        const HSecEdge &s = sec_edges[e];
        const float *g = table + gl.off_se + 6 * e;
        const D3 gp0 = d3(g), ge1 = d3(g + 3);
        gvw[s.mesh][s.v0] = gvw[s.mesh][s.v0] + gp0 - ge1;
        gvw[s.mesh][s.v1] = gvw[s.mesh][s.v1] + ge1;
    }

    // ---- triangle records (make_triangle_records in scene.cpp)
    for (size_t mi = 0; mi < meshes.size(); ++mi) {
        const HMesh &m = meshes[mi];
        const size_t nf = m.tris.size(), nv = m.v_world.size();
        std::vector<D3> gvn(nv), S(nv), gcross(nf);
        for (size_t i = 0; i < nf; ++i) {
            const float *g = table + (size_t) kGradTri * (m.face_offset + i);
            gvn[m.f[3 * i]] = gvn[m.f[3 * i]] + d3(g + 10);
            gvn[m.f[3 * i + 1]] = gvn[m.f[3 * i + 1]] + d3(g + 13);
            gvn[m.f[3 * i + 2]] = gvn[m.f[3 * i + 2]] + d3(g + 16);
            const D3 e1 = d3(m.tris[i].e1), e2 = d3(m.tris[i].e2);
            const D3 c = cross(e1, e2);
            for (int k = 0; k < 3; ++k) S[m.f[3 * i + k]] = S[m.f[3 * i + k]] + c;
        }
        // vn = normalize(S / w) = S / |S|: the weight w drops out of the derivative
        std::vector<D3> gS(nv);
        for (size_t v = 0; v < nv; ++v) {
            const double len = std::sqrt(dot(S[v], S[v]));
            if (!(len > 0.0)) continue;
            const D3 n = S[v] * (1.0 / len);
            gS[v] = (gvn[v] - n * dot(n, gvn[v])) * (1.0 / len);
        }
        for (size_t i = 0; i < nf; ++i) {
            const float *g = table + (size_t) kGradTri * (m.face_offset + i);
            const int a = m.f[3 * i], b = m.f[3 * i + 1], c = m.f[3 * i + 2];
            const D3 e1 = d3(m.tris[i].e1), e2 = d3(m.tris[i].e2);
            const D3 cr = cross(e1, e2);
            const double len = std::sqrt(dot(cr, cr));
            D3 gc = gS[a] + gS[b] + gS[c];
            if (len > 0.0) {
                const D3 fn = cr * (1.0 / len), gfn = d3(g + 19);
                gc = gc + (gfn - fn * dot(fn, gfn)) * (1.0 / len) + fn * (0.5 * (double) g[9]);   // fn = c/|c|, area = |c|/2
            }
            const D3 ge1 = d3(g + 3) + cross(e2, gc), ge2 = d3(g + 6) + cross(gc, e1), gp0 = d3(g);
            gvw[mi][a] = gvw[mi][a] + gp0 - ge1 - ge2;
            gvw[mi][b] = gvw[mi][b] + ge1;
            gvw[mi][c] = gvw[mi][c] + ge2;
        }
        // ---- world vertices = transform_pos(L * R * Rt, raw vertices)
        const Mat tw = mul(mul(mat_of(m.to_world[0]), mat_of(m.to_world[1])), mat_of(m.to_world[2]));
        Mat gtw = zero();
        ParamGrads::MeshG &out = grads.meshes[mi];
        out.v.assign(3 * nv, 0.0);
        std::memset(out.to_world, 0, sizeof(out.to_world));
        for (size_t v = 0; v < nv; ++v) {
            D3 gp;
            transform_pos_adj(tw, d3(m.v_raw[v]), gvw[mi][v], gp, gtw);
            out.v[3 * v] = gp.x; out.v[3 * v + 1] = gp.y; out.v[3 * v + 2] = gp.z;
        }
        split_product(m.to_world, gtw, out.to_world);
    }
    for (int k = 0; k < 3; ++k) {
        grads.bsdf_tex[k].assign((size_t) num_bsdf_records(), std::vector<float>());
        grads.bsdf_tex_uv[k].assign((size_t) num_bsdf_records(), std::vector<float>());
        for (size_t i = 0; i < (size_t) num_bsdf_records(); ++i)
            if (bsdf_record((int) i).tex[k].w > 0) {
                const float *g = table + gl.total + texture_grad_offset((int) i, k);
                const size_t nt = (size_t) HBsdf::tex_channels(k) * bsdf_record((int) i).tex[k].w * bsdf_record((int) i).tex[k].h;
                grads.bsdf_tex[k][i].assign(g, g + nt);
                grads.bsdf_tex_uv[k][i].assign(g + nt, g + nt + 4);
            }
    }
    grads.bsdf_pv.assign((size_t) num_bsdf_records(), std::vector<float>());
    for (size_t i = 0; i < (size_t) num_bsdf_records(); ++i)
        if (!bsdf_record((int) i).pv.empty()) {
            const float *g = table + gl.total + pervertex_grad_offset((int) i);
            grads.bsdf_pv[i].assign(g, g + bsdf_record((int) i).pv.size());
        }
    grads.bsdf_spec.assign(3 * (size_t) num_bsdf_records(), 0.0);
    grads.bsdf_rough.assign((size_t) num_bsdf_records(), 0.0);
    grads.bsdf_eta.assign(3 * (size_t) num_bsdf_records(), 0.0);
    grads.bsdf_k.assign(3 * (size_t) num_bsdf_records(), 0.0);
    for (size_t i = 0; i < (size_t) num_bsdf_records(); ++i) {
        for (int c = 0; c < 3; ++c) {
            grads.bsdf_refl[3 * i + c] = table[gl.off_bsdf + kGradBsdf * i + c];
            grads.bsdf_spec[3 * i + c] = table[gl.off_bsdf + kGradBsdf * i + 4 + c];
            grads.bsdf_eta[3 * i + c] = table[gl.off_bsdf + kGradBsdf * i + 8 + c];
            grads.bsdf_k[3 * i + c] = table[gl.off_bsdf + kGradBsdf * i + 12 + c];
        }
        grads.bsdf_rough[i] = table[gl.off_bsdf + kGradBsdf * i + 3];
    }
    for (size_t i = 0; i < emitters.size(); ++i)
        for (int c = 0; c < 3; ++c) grads.emitter_rad[3 * i + c] = table[gl.off_emit + 4 * i + c];
    grads.colloc_intensity = table[gl.off_cam + 38];
    grads.env_radiance.clear();
    grads.env_scale = 0.0;
    std::memset(grads.env_to_world_left, 0, sizeof(grads.env_to_world_left));
    if (env.present && gl.total > gl.off_env) {
        const float *g = table + gl.off_env;
        grads.env_scale = g[0];
        grads.env_radiance.assign(g + kGradEnvHead, g + kGradEnvHead + (size_t) 3 * env.w * env.h);
        // from_world = inverse(to_world_full), to_world_full = left * raw
        Mat gF = zero();
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) gF.m[i][j] = g[1 + 3 * i + j];
        const Mat F = mat_of(env.from_world), FT = transpose(F);
        const Mat t = mul(mul(FT, gF), FT);
        Mat gM = zero();
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) gM.m[i][j] = -t.m[i][j];
        const Mat gL = mul(gM, transpose(mat_of(env.to_world[1])));
        for (int i = 0; i < 16; ++i) grads.env_to_world_left[i] = gL.m[i / 4][i % 4];
    }
    grads.valid = true;
}

}  // namespace psdr
