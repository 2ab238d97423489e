// TEST INFRASTRUCTURE ONLY -- CPU oracle for the psdr-jit PathTracer hot path.
//
// A scalar C++ restatement of the reference algorithm (andyyankai/psdr-jit @ 50cc4f6), one
// lane at a time, with forward-mode dual numbers standing in for Dr.Jit's AD.  It is pinned
// against golden vectors produced by RUNNING the unmodified reference on a B200
// (tools/ref_golden.py -> tests/golden/*.npz); tests/test_oracle_golden.py checks that.
// Ray casting is brute force over all triangles (the reference uses OptiX).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library.  The product (psdr_jit_b200/) never does.
//
// Each function cites the reference file:line it follows.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "orc_math.h"

#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

namespace {

// ------------------------------------------------------------------------------------------
// Discrete distribution (reference src/core/pmf.cpp:6-51, include/psdr/core/pmf.h:12-38)
// ------------------------------------------------------------------------------------------
// drjit::sum on the GPU = block tree reduction (ext/drjit/ext/drjit-core/resources/reduce.cuh:
// 12-60): 1024 threads, thread t starts with data[t] + data[t+1024], then halving strides.
static float drjit_sum(const std::vector<float> &x) {
    size_t n = x.size();
    std::vector<float> partial;
    for (size_t base = 0; base < n || partial.empty(); base += 2048) {
        float s[1024];
        for (int t = 0; t < 1024; ++t) {
            float v = 0.f;
            size_t i = base + t;
            if (i < n) {
                v = v + x[i];
                if (i + 1024 < n) v = v + x[i + 1024];
            }
            s[t] = v;
        }
        for (int stride = 512; stride >= 1; stride >>= 1)
            for (int t = 0; t < stride; ++t) s[t] = s[t] + s[t + stride];
        partial.push_back(s[0]);
        if (n == 0) break;
    }
    if (partial.size() == 1) return partial[0];
    return drjit_sum(partial);
}

struct Distrib {
    int size = 0;
    float sum = 0.f;
    std::vector<float> pmf, cmf;
    void init(const std::vector<float> &p) {
        size = (int) p.size();
        pmf = p;
        sum = drjit_sum(p);
        cmf.resize(p.size());
        double acc = 0.0;  // pmf.h:19-25: double accumulation, rounded to float per entry
        for (size_t i = 0; i < p.size(); ++i) {
            acc += (double) p[i];
            cmf[i] = (float) acc;
        }
    }
    // drjit binary_search(0, size-1, cmf[i] < s)  (ext/drjit/include/drjit/util.h:172-217)
    int search(float s) const {
        int start = 0, end = size - 1;
        int iterations = 0;
        if (start < end) {
            unsigned r = (unsigned) (end - start);
            int lg = 0;
            while (r >>= 1) ++lg;
            iterations = lg + 1;
        }
        for (int i = 0; i < iterations; ++i) {
            int middle = (start + end) >> 1;
            bool cond = cmf[middle] < s;
            if (cond) start = std::min(middle + 1, end);
            else end = middle;
        }
        return start;
    }
    // pmf.cpp:18-28
    std::pair<int, float> sample(float s) const {
        if (size == 1) return {0, 1.f};
        s *= sum;
        int idx = search(s);
        return {idx, pmf[idx] / sum};
    }
    // pmf.cpp:31-51 -- mutates the sample
    std::pair<int, float> sample_reuse(float &s) const {
        if (size == 1) return {0, 1.f};
        s *= sum;
        int idx = search(s);
        if (idx > 0) s -= cmf[idx - 1];
        float p = pmf[idx];
        if (p > 0.f) s /= p;
        s = std::min(std::max(s, 0.f), 1.f);
        return {idx, p / sum};
    }
};

// ------------------------------------------------------------------------------------------
// Scene description
// ------------------------------------------------------------------------------------------
template <class S> struct Tri {  // reference include/psdr/types.h:162-175
    V3<S> p0, e1, e2, n0, n1, n2, fn;
    S area;
};

struct Edge {  // reference src/shape/mesh.cpp:244-305 (m_edge_indices rows)
    int v0, v1, f0, f1, v2;
};

struct Bsdf {
    int type = 0;  // 0 Diffuse, 1 Microfacet, 2 RoughConductor, 3 RoughDielectric, 4 MicrofacetPerVertex, 5 NormalMap
    V3d reflectance;  // Diffuse reflectance / Microfacet diffuseReflectance / NormalMap: constant normal map
    V3d specular;     // Microfacet specularReflectance / RoughConductor specular_reflectance
    Dual roughness;   // Microfacet roughness / RoughConductor alpha (alpha_u = alpha_v)
    V3d eta, k;       // RoughConductor eta + i k; RoughDielectric: eta.x = intIOR / extIOR, eta.y = extIOR / intIOR
    bool two_side = false;
    int nested = -1;                 // NormalMap: index (into Scene::bsdfs) of the BSDF it perturbs
    std::vector<float> pv, d_pv;     // MicrofacetPerVertex: 7 floats per vertex (specular rgb, diffuse rgb, roughness) + tangents
    // bitmap slots with more than one texel (channels interleaved, pixel = y*w + x): 0 reflectance / diffuseReflectance
    // (Bitmap3fD), 1 specularReflectance (Bitmap3fD), 2 roughness (Bitmap1fD); each with the bitmap's uv transform
    // (reference include/psdr/core/bitmap.h:36-38)
    struct Tex {
        int w = 0, h = 0, ch = 3;
        std::vector<float> data, ddata;
        Dual scale = Dual(1.f), rot = Dual(0.f), tx = Dual(0.f), ty = Dual(0.f);
    };
    Tex tex[3];
};

struct MeshRec {
    std::vector<V3d> v_raw;
    std::vector<int> f;  // 3*nf
    std::vector<V2f> uv;
    std::vector<int> fuv;
    bool has_uv = false;
    M4<Dual> to_world[3];  // left, raw, right
    int bsdf = -1;
    int emitter = -1;
    bool use_face_normals = false, enable_edges = true;
    // configured
    std::vector<V3d> v_world;
    std::vector<Tri<Dual>> tris;
    std::vector<Edge> edges;
    Distrib face_distrb;
    float total_area = 0.f, inv_total_area = 0.f;
    int face_offset = 0;
};

struct EmitterRec {
    int type = 0;  // 0 AreaLight, 1 EnvironmentMap
    V3d radiance;
    int mesh = -1;
    float sampling_weight = 0.f;
};

// EnvironmentMap (reference src/emitter/envmap.cpp, include/psdr/emitter/envmap.h)
struct Envmap {
    bool present = false, has_bounds = false;
    int emitter = -1, mesh = -1;
    int w = 0, h = 0;
    std::vector<float> data, ddata;  // rgb interleaved; forward tangents (may be empty)
    Dual scale = Dual(1.f);
    M4<Dual> to_world[2];            // left, raw
    M4<Dual> to_world_full, from_world;
    V3f lower, upper;
    int cw = 0, ch = 0;
    Distrib cell;
};

struct PrimEdge {  // reference include/psdr/edge/edge.h:26-40
    V2d p0, p1;
    V2f normal;
    float length;
};

struct SecEdge {  // edge.h:49-66
    V3d p0, e1;
    V3f n0, n1, p2;
    bool is_boundary;
};

struct Camera {
    float fov, near_, far_;
    bool ortho = false;           // OrthographicCamera(near, far) (src/sensor/orthographic.cpp)
    bool use_intrinsic = false;   // PerspectiveCamera(fx, fy, cx, cy, near, far) (include/psdr/sensor/perspective.h:11-12)
    float fx = 0.f, fy = 0.f, cx = 0.f, cy = 0.f;
    M4<Dual> to_world[3];
    // configured (reference src/sensor/perspective.cpp:10-46)
    M4<Dual> to_world_full, world_to_sample, sample_to_camera;
    V3d pos, dir;
    float inv_area;
    std::vector<PrimEdge> edges;
    Distrib edge_distrb;
    bool enable_edges = false;
    // secondary-edge guiding: HyperCubeDistribution<3> (reference src/core/cube_distrb.cpp:9-64)
    bool guided = false;
    int greso[3] = {0, 0, 0};
    Distrib guide;
};

struct Scene {
    int width = 128, height = 128, spp = 1, sppe = 0, sppse = 0;
    std::vector<Bsdf> bsdfs;
    std::vector<MeshRec> meshes;
    std::vector<EmitterRec> emitters;
    std::vector<Camera> cameras;
    Envmap env;
    // configured
    std::vector<Tri<Dual>> tris;
    std::vector<int> tri_mesh;
    std::vector<V2f> tri_uv;  // 3 per triangle
    std::vector<int> tri_fidx;  // 3 per triangle: mesh-local vertex indices (m_triangle_info.face_indices, mesh.cpp:28)
    // per triangle PAIR (2j, 2j+1): padded bounding box, centre / half extent (scenes of <= 64 triangles; see trace())
    std::vector<float> cull_c, cull_h;
    std::vector<SecEdge> sec_edges;
    Distrib sec_edge_distrb, emitter_distrb;
    bool configured = false;
    std::string error;
    // options
    bool li_p_first = true;  // evaluation order of Li(ray_n) - Li(ray_p), integrator.cpp:185-186
    int mis = 2;             // 2: PathTracer / Direct(2); 0 / 1: Direct(0) / Direct(1) (reference src/integrator/direct.cpp);
                             // 3: CollocatedIntegrator (reference src/integrator/collocated.cpp)
    Dual colloc_intensity = Dual(0.f);
    bool colloc_field = false;   // FieldExtractionIntegrator("bsdf") (field.cpp:72-92): the BSDF term alone
};

// The product's closest-hit query (psdr_jit_b200/csrc/device_path.cuh trace<brute>, the replacement of OptiX) tests a
// triangle pair only if the ray passes the pair's padded bounding box; that slab test is part of the DEFINITION of the
// closest hit in scenes of <= 64 triangles, so it is restated here operation by operation (boxes: device_upload.cu).
static void build_cull_boxes(Scene &sc) {
    sc.cull_c.clear();
    sc.cull_h.clear();
    const int ntris = (int) sc.tris.size();
    if (ntris > 64) return;
    const int npairs = (ntris + 1) / 2;
    std::vector<float> blo((size_t) 3 * npairs, 1e30f), bhi((size_t) 3 * npairs, -1e30f);
    float slo[3] = {1e30f, 1e30f, 1e30f}, shi[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = 0; i < ntris; ++i) {
        const V3f p0 = val(sc.tris[i].p0), e1 = val(sc.tris[i].e1), e2 = val(sc.tris[i].e2);
        const float P[3] = {p0.x, p0.y, p0.z}, A[3] = {e1.x, e1.y, e1.z}, B[3] = {e2.x, e2.y, e2.z};
        for (int k = 0; k < 3; ++k) {
            const float v[3] = {P[k], P[k] + A[k], P[k] + B[k]};
            for (float x : v) {
                blo[3 * (i >> 1) + k] = std::fmin(blo[3 * (i >> 1) + k], x);
                bhi[3 * (i >> 1) + k] = std::fmax(bhi[3 * (i >> 1) + k], x);
                slo[k] = std::fmin(slo[k], x);
                shi[k] = std::fmax(shi[k], x);
            }
        }
    }
    float ext = 0.f;
    for (int k = 0; k < 3; ++k) ext = std::fmax(ext, shi[k] - slo[k]);
    const float pad_fat = 2.5e-4f * ext, pad_flat = 2.5e-7f * ext, flat_below = 1e-5f * ext;
    sc.cull_c.resize((size_t) 3 * npairs);
    sc.cull_h.resize((size_t) 3 * npairs);
    for (int j = 0; j < npairs; ++j)
        for (int k = 0; k < 3; ++k) {
            const float lo = blo[3 * j + k], hi = bhi[3 * j + k];
            const float half = 0.5f * (hi - lo);
            sc.cull_c[3 * j + k] = 0.5f * (lo + hi);
            sc.cull_h[3 * j + k] = half + (half < flat_below ? pad_flat : pad_fat);
        }
}
static bool cull_box_hit(const Scene &sc, int pair, V3f o, V3f d) {
    const float O[3] = {o.x, o.y, o.z}, D[3] = {d.x, d.y, d.z};
    float mn[3], fr[3];
    for (int k = 0; k < 3; ++k) {
        const float inv = 1.f / D[k];
        const float dx = sc.cull_c[3 * pair + k] - O[k];
        const float th = sc.cull_h[3 * pair + k] * std::fabs(inv);
        fr[k] = std::fmaf(dx, inv, th);
        mn[k] = std::fmaf(dx, -inv, th);       // = -(near)
    }
    const float tn = -std::fmin(std::fmin(std::fmin(mn[0], mn[1]), mn[2]), -kRayEpsilon);   // fmin / fmax drop NaNs
    const float tf = std::fmin(std::fmin(fr[0], fr[1]), fr[2]);
    return tn <= tf;
}

// reference src/shape/mesh.cpp:23-62
static void process_mesh(const std::vector<V3d> &vp, const std::vector<int> &f, std::vector<Tri<Dual>> &out) {
    size_t nf = f.size() / 3, nv = vp.size();
    out.resize(nf);
    std::vector<V3d> vn(nv);
    std::vector<Dual> vw(nv);
    std::vector<V3d> fnorm(nf);
    std::vector<Dual> farea(nf);
    for (size_t i = 0; i < nf; ++i) {
        Tri<Dual> &t = out[i];
        t.p0 = vp[f[3 * i]];
        t.e1 = vp[f[3 * i + 1]] - t.p0;
        t.e2 = vp[f[3 * i + 2]] - t.p0;
        fnorm[i] = cross(t.e1, t.e2);
        farea[i] = norm(fnorm[i]);
    }
    // scatter_reduce order on the GPU is unspecified; accumulate in face order
    for (int k = 0; k < 3; ++k)
        for (size_t i = 0; i < nf; ++i) {
            int vi = f[3 * i + k];
            vn[vi] = vn[vi] + fnorm[i];
            vw[vi] = vw[vi] + farea[i];
        }
    for (size_t i = 0; i < nv; ++i) vn[i] = normalize(vn[i] / vw[i]);
    for (size_t i = 0; i < nf; ++i) {
        Tri<Dual> &t = out[i];
        t.n0 = vn[f[3 * i]];
        t.n1 = vn[f[3 * i + 1]];
        t.n2 = vn[f[3 * i + 2]];
        t.fn = fnorm[i] / farea[i];
        t.area = farea[i] * 0.5f;
    }
}

// reference src/shape/mesh.cpp:244-305 (std::map keyed by (min,max) vertex id)
static void build_edges(MeshRec &m) {
    m.edges.clear();
    if (!m.enable_edges) return;
    std::map<std::pair<int, int>, std::vector<int>> edge_map;
    size_t nf = m.f.size() / 3;
    for (size_t fi = 0; fi < nf; ++fi)
        for (int i = 0; i < 3; ++i) {
            int i1 = m.f[3 * fi + i], i2 = m.f[3 * fi + (i + 1) % 3], i3 = m.f[3 * fi + (i + 2) % 3];
            auto key = i1 < i2 ? std::make_pair(i1, i2) : std::make_pair(i2, i1);
            if (edge_map.find(key) == edge_map.end()) edge_map[key].push_back(i3);
            edge_map[key].push_back((int) fi);
        }
    for (auto &it : edge_map) {
        Edge e;
        e.v0 = it.first.first;
        e.v1 = it.first.second;
        e.f0 = it.second[1];
        e.f1 = it.second.size() >= 3 ? it.second[2] : -1;
        e.v2 = it.second[0];
        m.edges.push_back(e);
    }
}

static inline float rgb2luminance(V3f c) { return c.x * .2126f + c.y * .7152f + c.z * .0722f; }

// reference include/psdr/core/transform.h:48-61
static M4<float> perspective(float fov, float near_, float far_) {
    float recip = 1.f / (far_ - near_);
    float tn = std::tan(fov * .5f * (kPi / 180.f)), cot = 1.f / tn;
    M4<float> t = M4<float>::identity();
    t.m[0][0] = cot;
    t.m[1][1] = cot;
    t.m[2][2] = far_ * recip;
    t.m[3][3] = 0.f;
    t.m[2][3] = -near_ * far_ * recip;
    t.m[3][2] = 1.f;
    return t;
}

// reference include/psdr/core/transform.h:63-71
static M4<float> perspective_intrinsic(float fx, float fy, float cx, float cy, float near_, float far_) {
    float recip = 1.f / (far_ - near_);
    M4<float> t = M4<float>::identity();
    t.m[2][2] = far_ * recip;
    t.m[3][3] = 0.f;
    t.m[2][3] = -near_ * far_ * recip;
    t.m[3][2] = 1.f;
    M4<float> tr = M4<float>::identity(), sc_ = M4<float>::identity();
    tr.m[0][3] = 1.f - 2.f * cx;
    tr.m[1][3] = 1.f - 2.f * cy;
    sc_.m[0][0] = 2.f * fx;
    sc_.m[1][1] = 2.f * fy;
    return (tr * sc_) * t;
}

// reference src/shape/mesh.cpp:317-382
static void configure_mesh(MeshRec &m) {
    M4<Dual> tw = (m.to_world[0] * m.to_world[1]) * m.to_world[2];
    m.v_world.resize(m.v_raw.size());
    for (size_t i = 0; i < m.v_raw.size(); ++i) m.v_world[i] = transform_pos(tw, m.v_raw[i]);
    process_mesh(m.v_world, m.f, m.tris);
    std::vector<float> areas(m.tris.size());
    for (size_t i = 0; i < m.tris.size(); ++i) areas[i] = m.tris[i].area.v;
    m.total_area = drjit_sum(areas);
    m.inv_total_area = 1.f / m.total_area;
    m.face_distrb.init(areas);
}

// reference src/sensor/perspective.cpp:10-152
static bool configure_camera(Scene &sc, Camera &cam, bool build_primary_edges) {
    float aspect = (float) sc.width / (float) sc.height;
    M4<float> sc_ = M4<float>::identity(), tr = M4<float>::identity();
    sc_.m[0][0] = -0.5f;
    sc_.m[1][1] = -0.5f * aspect;
    tr.m[0][3] = -1.f;
    tr.m[1][3] = -1.f / aspect;
    if (cam.use_intrinsic) {   // perspective.cpp:15-20
        sc_.m[1][1] = -0.5f;
        tr.m[1][3] = -1.f;
    }
    M4<float> proj;
    if (cam.ortho) {   // transform::orthographic (transform.h:73-76), orthographic.cpp:17-20
        M4<float> os = M4<float>::identity(), ot = M4<float>::identity();
        os.m[2][2] = 1.f / (cam.far_ - cam.near_);
        ot.m[2][3] = -cam.near_;
        proj = os * ot;
    } else proj = cam.use_intrinsic ? perspective_intrinsic(cam.fx, cam.fy, cam.cx, cam.cy, cam.near_, cam.far_) : perspective(cam.fov, cam.near_, cam.far_);
    M4<float> c2s = (sc_ * tr) * proj;
    M4<Dual> camera_to_sample = lift<Dual>(c2s);
    cam.sample_to_camera = lift<Dual>(inverse(c2s));
    cam.to_world_full = (cam.to_world[0] * cam.to_world[1]) * cam.to_world[2];
    cam.world_to_sample = camera_to_sample * inverse(cam.to_world_full);
    cam.pos = transform_pos(cam.to_world_full, V3d(Dual(0.f), Dual(0.f), Dual(0.f)));
    cam.dir = transform_dir(cam.to_world_full, V3d(Dual(0.f), Dual(0.f), Dual(1.f)));
    M4<float> s2c = val(cam.sample_to_camera);
    V3f v00 = transform_pos(s2c, V3f(0.f, 0.f, 0.f)), v10 = transform_pos(s2c, V3f(1.f, 0.f, 0.f)),
        v11 = transform_pos(s2c, V3f(1.f, 1.f, 0.f)), vc = transform_pos(s2c, V3f(.5f, .5f, 0.f));
    cam.inv_area = (1.f / (norm(v00 - v10) * norm(v11 - v10))) * squared_norm(vc);

    cam.edges.clear();
    cam.enable_edges = false;
    if (sc.sppe > 0 && build_primary_edges) {
        for (auto &m : sc.meshes) {
            if (!m.enable_edges) continue;
            size_t before = cam.edges.size();
            for (auto &e : m.edges) {
                bool valid = e.f1 >= 0;
                V3f camp = val(cam.pos);
                V3f e0 = normalize(camp - val(m.tris[e.f0].p0));
                V3f e1 = normalize(camp - (valid ? val(m.tris[e.f1].p0) : V3f(0.f, 0.f, 0.f)));
                V3f n0 = val(m.tris[e.f0].fn);
                V3f n1 = valid ? val(m.tris[e.f1].fn) : V3f(0.f, 0.f, 0.f);
                bool uv_mask = false;
                if (m.has_uv) {  // perspective.cpp:73-94
                    int a[3] = {m.fuv[3 * e.f0], m.fuv[3 * e.f0 + 1], m.fuv[3 * e.f0 + 2]};
                    int b[3] = {0, 0, 0};
                    if (valid) { b[0] = m.fuv[3 * e.f1]; b[1] = m.fuv[3 * e.f1 + 1]; b[2] = m.fuv[3 * e.f1 + 2]; }
                    int cut = 0;  // number of uv indices of face 0 that face 1 also uses
                    for (int k = 0; k < 3; ++k)
                        if (a[k] == b[0] || a[k] == b[1] || a[k] == b[2]) cut++;
                    uv_mask = (cut != 2);
                }
                bool keep;
                if (m.use_face_normals) {
                    bool skip = valid && ((dot(e0, n0) < kEpsilon && dot(e1, n1) < kEpsilon) || (dot(n0, n1) > 1.f - kEpsilon));
                    keep = !skip || (m.has_uv && uv_mask);
                } else {
                    bool active = !valid;
                    active |= (dot(e0, n0) > kEpsilon) != (dot(e1, n1) > kEpsilon);
                    keep = active || (m.has_uv && uv_mask);
                }
                if (!keep) continue;
                PrimEdge pe;
                V3d q0 = transform_pos(cam.world_to_sample, m.v_world[e.v0]);
                V3d q1 = transform_pos(cam.world_to_sample, m.v_world[e.v1]);
                pe.p0 = V2d(q0.x, q0.y);
                pe.p1 = V2d(q1.x, q1.y);
                V2f ev(q1.x.v - q0.x.v, q1.y.v - q0.y.v);
                float len = norm(ev);
                ev.x /= len;
                ev.y /= len;
                pe.normal = V2f(-ev.y, ev.x);
                pe.length = len;
                cam.edges.push_back(pe);
            }
            if (cam.edges.size() == before) {
                sc.error = "PSDR_ASSERT(slices(info) > 0): a mesh produced no primary edges";
                return false;
            }
        }
        if (!cam.edges.empty()) {
            std::vector<float> lens(cam.edges.size());
            for (size_t i = 0; i < lens.size(); ++i) lens[i] = cam.edges[i].length;
            cam.edge_distrb.init(lens);
            cam.enable_edges = true;
        }
    }
    return true;
}

// reference src/scene/scene.cpp:311-601
// ---- EnvironmentMap ---------------------------------------------------------------------------
// Bitmap<3>::eval, envmap mode, no uv transform (reference src/core/bitmap.cpp:46-131)
template <class S> static V3<S> texel(const Envmap &e, int i);
template <> V3<float> texel<float>(const Envmap &e, int i) { return V3f(e.data[3 * i], e.data[3 * i + 1], e.data[3 * i + 2]); }
template <> V3<Dual> texel<Dual>(const Envmap &e, int i) {
    if (e.ddata.empty()) return lift<Dual>(V3f(e.data[3 * i], e.data[3 * i + 1], e.data[3 * i + 2]));
    return V3d(Dual(e.data[3 * i], e.ddata[3 * i]), Dual(e.data[3 * i + 1], e.ddata[3 * i + 1]), Dual(e.data[3 * i + 2], e.ddata[3 * i + 2]));
}
template <class S> static V3<S> envmap_bitmap_eval(const Envmap &e, V2<S> uv) {
    const int w = e.w, h = e.h;
    uv = V2<S>((uv.x - 0.5f) + 0.5f, (uv.y - 0.5f) + 0.5f);
    uv.x = uv.x - (float) (0.5 / (double) w);
    uv = V2<S>(uv.x - floor_(uv.x), uv.y - floor_(uv.y));
    uv.x = uv.x * (float) w;
    uv.y = uv.y * (float) (h - 1);
    int px = (int) std::floor(val(uv.x)), py = (int) std::floor(val(uv.y));
    S w1x = uv.x - (float) px, w1y = uv.y - (float) py;
    S w0x = 1.0f - w1x, w0y = 1.0f - w1y;
    int yw = std::min(py, h - 2) * w, xp1 = (px + 1) % w, last = w * h - 1;
    int i00 = std::min(yw + px, last), i10 = std::min(yw + xp1, last), i01 = std::min(yw + px + w, last), i11 = std::min(yw + xp1 + w, last);
    V3<S> v00 = texel<S>(e, i00), v10 = texel<S>(e, i10), v01 = texel<S>(e, i01), v11 = texel<S>(e, i11);
    V3<S> v0(fmadd(w0x, v00.x, w1x * v10.x), fmadd(w0x, v00.y, w1x * v10.y), fmadd(w0x, v00.z, w1x * v10.z));
    V3<S> v1(fmadd(w0x, v01.x, w1x * v11.x), fmadd(w0x, v01.y, w1x * v11.y), fmadd(w0x, v01.z, w1x * v11.z));
    return V3<S>(fmadd(w0y, v0.x, w1y * v1.x), fmadd(w0y, v0.y, w1y * v1.y), fmadd(w0y, v0.z, w1y * v1.z));
}
template <class S> static V2<S> dir_to_uv(V3<S> v) {  // envmap.cpp:66-67
    V2<S> uv(atan2_(v.x, -v.z) * 0.15915494309189533577f, safe_acos_(v.y) * kInvPi);
    return V2<S>(uv.x - floor_(uv.x), uv.y - floor_(uv.y));
}
template <class S> static M4<S> mat_as(const M4<Dual> &m);
template <> M4<Dual> mat_as<Dual>(const M4<Dual> &m) { return m; }
template <> M4<float> mat_as<float>(const M4<Dual> &m) { return val(m); }
template <class S> static S scale_of(const Envmap &e);
template <> Dual scale_of<Dual>(const Envmap &e) { return e.scale; }
template <> float scale_of<float>(const Envmap &e) { return e.scale.v; }
// EnvironmentMap::eval_direction (envmap.cpp:56-73)
template <class S> static V3<S> env_eval_direction(const Envmap &e, V3<S> wi) {
    V3<S> v = transform_dir(mat_as<S>(e.from_world), wi);
    V3<S> r = envmap_bitmap_eval<S>(e, dir_to_uv<S>(v));
    return r * scale_of<S>(e);
}

static void configure_mesh(MeshRec &m);
// scene.cpp:355-368,435-485 (bounding box fixed at the first configure) + envmap.cpp:17-41 (cell masses)
static bool configure_envmap(Scene &sc) {
    Envmap &env = sc.env;
    if (!env.present) return true;
    if (env.w < 2 || env.h < 2) { sc.error = "src/emitter/envmap.cpp (21): width > 1 && height > 1"; return false; }
    env.to_world_full = env.to_world[0] * env.to_world[1];
    env.from_world = inverse(env.to_world_full);
    if (!env.has_bounds) {
        V3f lo(3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f), hi(1.175494351e-38f, 1.175494351e-38f, 1.175494351e-38f);
        auto grow = [&](V3f p) {
            lo = V3f(std::fmin(lo.x, p.x), std::fmin(lo.y, p.y), std::fmin(lo.z, p.z));
            hi = V3f(std::fmax(hi.x, p.x), std::fmax(hi.y, p.y), std::fmax(hi.z, p.z));
        };
        for (auto &m : sc.meshes) for (auto &p : m.v_world) grow(val(p));
        for (auto &c : sc.cameras) grow(val(c.pos));
        float margin = std::fmin(std::fmin((hi.x - lo.x) * 0.05f, (hi.y - lo.y) * 0.05f), (hi.z - lo.z) * 0.05f);
        env.lower = V3f(lo.x - margin, lo.y - margin, lo.z - margin);
        env.upper = V3f(hi.x + margin, hi.y + margin, hi.z + margin);
        static const int face_data[3][12] = {{0, 0, 1, 1, 2, 2, 0, 0, 0, 0, 4, 4}, {1, 3, 5, 7, 3, 7, 5, 4, 2, 6, 7, 6}, {3, 2, 7, 3, 7, 6, 1, 5, 6, 4, 5, 7}};
        MeshRec b;
        for (int i = 0; i < 8; ++i) {
            float c[3];
            for (int j = 0; j < 3; ++j) c[j] = (i & (1 << j)) ? env.upper[j] : env.lower[j];
            b.v_raw.push_back(V3d(Dual(c[0]), Dual(c[1]), Dual(c[2])));
        }
        for (int k = 0; k < 12; ++k) for (int r = 0; r < 3; ++r) b.f.push_back(face_data[r][k]);
        for (auto &M : b.to_world) M = M4<Dual>::identity();
        b.use_face_normals = true;
        b.enable_edges = false;
        b.bsdf = -1;
        b.emitter = env.emitter;
        configure_mesh(b);
        b.face_offset = 0;
        for (auto &m : sc.meshes) b.face_offset += (int) m.tris.size();
        env.mesh = (int) sc.meshes.size();
        sc.emitters[env.emitter].mesh = env.mesh;
        sc.meshes.push_back(b);
        env.has_bounds = true;
    }
    env.cw = (env.w - 1) << 1;
    env.ch = (env.h - 1) << 1;
    size_t ncells = (size_t) env.cw * env.ch;
    std::vector<float> mass(ncells);
    float ux = 1.f / (float) env.cw, uy = 1.f / (float) env.ch, dtheta = kPi / (float) env.ch;
    for (size_t idx = 0; idx < ncells; ++idx) {
        int cx = (int) (idx / env.ch), cy = (int) (idx % env.ch);
        V3f v = envmap_bitmap_eval<float>(env, V2f(((float) cx + .5f) * ux, ((float) cy + .5f) * uy));
        float sn, cs;
        sincos_full(((float) cy + .5f) * dtheta, sn, cs);
        mass[idx] = rgb2luminance(v) * sn;
    }
    env.cell.init(mass);
    return true;
}

static bool configure_scene(Scene &sc, const int *active, int nactive) {
    sc.error.clear();
    if (sc.meshes.empty()) { sc.error = "Missing meshes!"; return false; }
    if (sc.cameras.empty()) { sc.error = "Missing sensor!"; return false; }
    int off = 0;
    for (auto &m : sc.meshes) {
        build_edges(m);
        configure_mesh(m);
        m.face_offset = off;
        off += (int) m.tris.size();
    }
    // sensors: primary edges only for the sensors named in `active` (scene.cpp:381-416)
    for (size_t i = 0; i < sc.cameras.size(); ++i) {
        bool act = false;
        for (int k = 0; k < nactive; ++k) act |= (active[k] == (int) i);
        if (!configure_camera(sc, sc.cameras[i], act)) return false;
    }
    if (!configure_envmap(sc)) return false;
    // emitters (scene.cpp:489-515, src/emitter/area.cpp:9-14); the envmap's weight = sum of the others
    if (!sc.emitters.empty()) {
        std::vector<float> w;
        double total = 0.0;
        for (auto &e : sc.emitters) {
            e.sampling_weight = e.type == 1 ? 0.f : sc.meshes[e.mesh].total_area * rgb2luminance(val(e.radiance));
            total += e.sampling_weight;
        }
        for (auto &e : sc.emitters) if (e.type == 1) e.sampling_weight = (float) total;
        for (auto &e : sc.emitters) w.push_back(e.sampling_weight);
        sc.emitter_distrb.init(w);
        float inv_total = 1.f / sc.emitter_distrb.sum;
        for (auto &e : sc.emitters) e.sampling_weight *= inv_total;
    }
    // global triangle arrays (scene.cpp:529-542)
    sc.tris.clear();
    sc.tri_mesh.clear();
    sc.tri_uv.clear();
    sc.tri_fidx.clear();
    for (size_t mi = 0; mi < sc.meshes.size(); ++mi) {
        auto &m = sc.meshes[mi];
        for (size_t i = 0; i < m.tris.size(); ++i) {
            sc.tris.push_back(m.tris[i]);
            sc.tri_mesh.push_back((int) mi);
            for (int k = 0; k < 3; ++k) sc.tri_uv.push_back(m.has_uv ? m.uv[m.fuv[3 * i + k]] : V2f(0.f, 0.f));
            for (int k = 0; k < 3; ++k) sc.tri_fidx.push_back(3 * i + k < m.f.size() ? m.f[3 * i + k] : 0);
        }
    }
    build_cull_boxes(sc);
    // secondary edges (scene.cpp:547-571, mesh.cpp:353-369)
    sc.sec_edges.clear();
    if (sc.sppse > 0) {
        for (auto &m : sc.meshes) {
            if (!m.enable_edges) continue;
            for (auto &e : m.edges) {
                SecEdge s;
                s.is_boundary = e.f1 < 0;
                s.p0 = m.v_world[e.v0];
                s.e1 = m.v_world[e.v1] - s.p0;
                s.n0 = val(m.tris[e.f0].fn);
                s.n1 = s.is_boundary ? V3f(0.f, 0.f, 0.f) : val(m.tris[e.f1].fn);
                s.p2 = val(m.v_world[e.v2]);
                sc.sec_edges.push_back(s);
            }
        }
        std::vector<float> lens(sc.sec_edges.size());
        for (size_t i = 0; i < lens.size(); ++i) lens[i] = norm(val(sc.sec_edges[i].e1));
        if (lens.empty()) { sc.error = "DiscreteDistribution: empty distribution!"; return false; }
        sc.sec_edge_distrb.init(lens);
    }
    sc.configured = true;
    return true;
}

// ------------------------------------------------------------------------------------------
// Ray casting: brute force closest hit in (RayEpsilon, 1e8); reference
// src/scene/scene_optix.cpp:343-410 delegates this to OptiX.
// ------------------------------------------------------------------------------------------
struct Hit {
    int tri = -1;
    float u = 0.f, v = 0.f, t = 0.f;
};

// Decision margins of one path (tools/ref_parity.py: which lanes can flip against OptiX and why).  Every closest-hit
// query and every shadow test of the path lowers the minimum of its class:
//   [0] shadow: |t - (dist - ShadowEpsilon)| of an emitter hit by a next-event ray (path.cpp:58), world units
//   [1] border: |min(u, v, 1-u-v)| of any triangle whose plane is crossed no farther than the winner (the ray passes
//       that close, in barycentric units, to an edge where the closest triangle changes)
//   [2] self:   |t - RayEpsilon| / RayEpsilon of a triangle pierced near the tmin threshold (self-intersection)
//   [3] tie:    (t2 - t1) / t1 between the winner and the runner-up
struct LaneDiag {
    float m[4] = {1e30f, 1e30f, 1e30f, 1e30f};
};
static thread_local LaneDiag *g_diag = nullptr;

static void trace_margins(const Scene &sc, V3f o, V3f d, int best_tri, float best_t) {
    LaneDiag &D = *g_diag;
    for (size_t i = 0; i < sc.tris.size(); ++i) {
        const Tri<Dual> &T = sc.tris[i];
        V3f e1 = val(T.e1), e2 = val(T.e2), p0 = val(T.p0);
        V3f h = cross_fms(d, e2);
        float det = dot(e1, h);
        if (det == 0.f) continue;
        V3f s = o - p0;
        float f = 1.f / det;
        float u = f * dot(s, h);
        V3f q = cross_fms(s, e1);
        float v = f * dot(d, q), t = f * dot(e2, q);
        float m = std::fmin(std::fmin(u, v), 1.f - u - v);
        bool no_farther = best_tri < 0 ? (t > 0.5f * kRayEpsilon) : (t > 0.5f * kRayEpsilon && t <= best_t * 1.001f);
        if (no_farther) D.m[1] = std::fmin(D.m[1], std::fabs(m));
        if (m > -1e-3f && t < 4.f * kRayEpsilon && t > -2.f * kRayEpsilon) D.m[2] = std::fmin(D.m[2], std::fabs(t - kRayEpsilon) / kRayEpsilon);
        if (best_tri >= 0 && (int) i != best_tri && m >= 0.f && t > kRayEpsilon) D.m[3] = std::fmin(D.m[3], (t - best_t) / best_t);
    }
}

// fp32 Moeller-Trumbore numerators with a FIXED operation order (cross = fused multiply-subtract, dot = fma
// chain).  A candidate is accepted on the numerators alone: inside test on the sign-normalised numerators,
// t in (RayEpsilon, 1e8) and "closer than the best so far" by cross-multiplication (t1 < t2 <=> tn1 |det2| <
// tn2 |det1|); the IEEE divide happens once, for the winner.  OptiX's own arithmetic is closed source, so this is
// the repo's definition of the closest hit; the CUDA kernels use the same order so that hit ids and (u,v,t) agree
// bit for bit with this oracle.  Ascending scan, strict comparison: ties keep the lowest triangle id.
static Hit trace(const Scene &sc, V3f o, V3f d) {
    Hit best;
    if (std::isnan(o.x) || std::isnan(o.y) || std::isnan(o.z) || std::isnan(d.x) || std::isnan(d.y) || std::isnan(d.z))
        return best;
    float b_ts = 1e8f, b_adet = 1.f, b_tn = 0.f, b_det = 0.f, b_un = 0.f, b_vn = 0.f;
    int b_tri = -1;
    const bool cull = !sc.cull_c.empty();
    bool pair_ok = true;
    for (size_t i = 0; i < sc.tris.size(); ++i) {
        if (cull && (i & 1) == 0) pair_ok = cull_box_hit(sc, (int) (i >> 1), o, d);
        if (!pair_ok) continue;
        const Tri<Dual> &T = sc.tris[i];
        V3f e1 = val(T.e1), e2 = val(T.e2), p0 = val(T.p0);
        V3f h = cross_fms(d, e2);
        float det = dot(e1, h);
        V3f s = o - p0;
        float un = dot(s, h);
        V3f q = cross_fms(s, e1);
        float vn = dot(d, q);
        float tn = dot(e2, q);
        float adet = std::fabs(det);
        bool neg = det < 0.f;
        float us = neg ? -un : un, vs = neg ? -vn : vn, ts = neg ? -tn : tn;
        bool ok = us >= 0.f && vs >= 0.f && us + vs <= adet && adet > 0.f && ts > kRayEpsilon * adet &&
                  ts * b_adet < b_ts * adet;   // b_ts starts at 1e8 (b_adet = 1): closer also means t < 1e8
        if (!ok) continue;
        b_ts = ts; b_adet = adet; b_tn = tn; b_det = det; b_un = un; b_vn = vn; b_tri = (int) i;
    }
    if (b_tri < 0) {
        if (g_diag) trace_margins(sc, o, d, -1, 0.f);
        return best;
    }
    float f = 1.f / b_det;
    best.tri = b_tri;
    best.t = f * b_tn;
    best.u = f * b_un;
    best.v = f * b_vn;
    if (g_diag) trace_margins(sc, o, d, best.tri, best.t);
    return best;
}

// reference include/psdr/core/frame.h:9-28 (Duff et al. ONB)
template <class S> static void coordinate_system(V3<S> n, V3<S> &s, V3<S> &t) {
    float sign = std::copysign(1.f, val(n.z));
    S a = -rcp_(sign + n.z);
    S b = n.x * n.y * a;
    auto mulsign = [](S x, float z) { return std::signbit(z) ? -x : x; };
    s = V3<S>(mulsign(sqr(n.x) * a, val(n.z)) + 1.f, mulsign(b, val(n.z)), mulsign(-n.x, val(n.z)));
    t = V3<S>(b, sign + sqr(n.y) * a, -n.y);
}

template <class S> struct Its {  // reference include/psdr/core/intersection.h:23-60
    bool valid = false;
    int mesh = -1, tri = -1;
    V3<S> p, n, wi, sh_s, sh_t, sh_n;
    S t = S(0.f), J = S(1.f);
    V2<S> uv;
    float bu = 0.f, bv = 0.f;
    V2<S> bc;        // its.bc (scene.cpp:769,799)
    V3<S> dp_du;     // its.dp_du (scene.cpp:760-763,789-792): zero without UVs
    V3<S> to_local(V3<S> v) const { return {dot(v, sh_s), dot(v, sh_t), dot(v, sh_n)}; }
    V3<S> to_world(V3<S> v) const { return sh_s * v.x + sh_t * v.y + sh_n * v.z; }
};

template <class S> struct TriS {
    V3<S> p0, e1, e2, n0, n1, n2, fn;
    S area;
};
template <class S> static TriS<S> get_tri(const Scene &sc, int i);
template <> TriS<Dual> get_tri<Dual>(const Scene &sc, int i) {
    const Tri<Dual> &t = sc.tris[i];
    return {t.p0, t.e1, t.e2, t.n0, t.n1, t.n2, t.fn, t.area};
}
template <> TriS<float> get_tri<float>(const Scene &sc, int i) {
    const Tri<Dual> &t = sc.tris[i];
    return {val(t.p0), val(t.e1), val(t.e2), val(t.n0), val(t.n1), val(t.n2), val(t.fn), t.area.v};
}

// reference include/psdr/utils.h:82-93
template <class S> static void ray_intersect_triangle(V3<S> p0, V3<S> e1, V3<S> e2, V3<S> o, V3<S> d, S &u, S &v, S &t) {
    V3<S> h = cross(d, e2);
    S a = dot(e1, h);
    S f = rcp_(a);
    V3<S> s = o - p0;
    u = f * dot(s, h);
    V3<S> q = cross(s, e1);
    v = f * dot(d, q);
    t = f * dot(e2, q);
}

// reference src/scene/scene.cpp:612-806.  path_space=false with S=Dual is the solid-angle AD
// formulation (analytic re-intersection); everything else is the material-form one.
template <class S> static Its<S> ray_intersect(const Scene &sc, V3<S> o, V3<S> d, bool active, bool path_space,
                                               int *out_tri = nullptr) {
    Its<S> its;
    if (out_tri) *out_tri = -1;
    if (!active) return its;
    Hit h = trace(sc, val(o), val(d));
    if (h.tri < 0) return its;
    if (out_tri) *out_tri = h.tri;
    constexpr bool ad = std::is_same<S, Dual>::value;
    TriS<S> T = get_tri<S>(sc, h.tri);
    its.valid = true;
    its.tri = h.tri;
    its.mesh = sc.tri_mesh[h.tri];
    its.n = T.fn;
    const MeshRec &mesh = sc.meshes[its.mesh];
    V2<S> uv0 = lift<S>(sc.tri_uv[3 * h.tri]), uv1 = lift<S>(sc.tri_uv[3 * h.tri + 1]), uv2 = lift<S>(sc.tri_uv[3 * h.tri + 2]);
    V2<S> duv0 = uv1 - uv0, duv1 = uv2 - uv0;
    S det = duv0.x * duv1.y - duv0.y * duv1.x;  // fmsub in the reference
    bool valid_dp = val(det) != 0.f;
    S inv_det = rcp_(det);
    V3<S> sh_n, dir;
    if (!ad || path_space) {
        V2f uv(h.u, h.v);
        sh_n = normalize(bilinear(T.n0, T.n1 - T.n0, T.n2 - T.n0, uv));
        if (mesh.use_face_normals) sh_n = its.n;
        its.p = bilinear(T.p0, T.e1, T.e2, uv);
        dir = its.p - o;
        its.t = norm(dir);
        dir = dir / its.t;
        its.uv = bilinear2(uv0, duv0, duv1, uv);
        its.bu = h.u;
        its.bv = h.v;
        its.bc = V2<S>(S(h.u), S(h.v));
        if (ad) its.J = T.area / detach(T.area);
    } else {
        S u, v, t;
        ray_intersect_triangle(T.p0, T.e1, T.e2, o, d, u, v, t);
        V2<S> uv(u, v);
        sh_n = normalize(bilinear(T.n0, T.n1 - T.n0, T.n2 - T.n0, uv));
        if (mesh.use_face_normals) sh_n = its.n;
        its.p = V3<S>(fmadd(d.x, t, o.x), fmadd(d.y, t, o.y), fmadd(d.z, t, o.z));
        its.t = t;
        its.uv = bilinear2(uv0, duv0, duv1, uv);
        its.bu = val(u);
        its.bv = val(v);
        its.bc = uv;
        dir = d;
    }
    its.sh_n = sh_n;
    coordinate_system(sh_n, its.sh_s, its.sh_t);
    its.dp_du = V3<S>(S(0.f));
    if (valid_dp) {
        V3<S> dp_du = (T.e1 * duv1.y - T.e2 * duv0.y) * inv_det;
        its.dp_du = dp_du;
        its.sh_s = normalize(dp_du - sh_n * dot(sh_n, dp_du));
        its.sh_t = cross(sh_n, its.sh_s);
    }
    its.wi = its.to_local(-dir);
    return its;
}

// ---- BSDF (reference src/bsdf/diffuse.cpp:23-108) ----------------------------------------
template <class S> static V3<S> refl_const(const Bsdf &b);
template <> V3<Dual> refl_const<Dual>(const Bsdf &b) { return b.reflectance; }
template <> V3<float> refl_const<float>(const Bsdf &b) { return val(b.reflectance); }
template <class S> static S lift_d(Dual x);
template <> float lift_d<float>(Dual x) { return x.v; }
template <> Dual lift_d<Dual>(Dual x) { return x; }
template <class S> static V3<S> tex_texel(const Bsdf::Tex &t, int i) {
    V3<S> r(S(0.f));
    for (int c = 0; c < t.ch; ++c) {
        Dual x(t.data[t.ch * i + c], t.ddata.empty() ? 0.f : t.ddata[t.ch * i + c]);
        (c == 0 ? r.x : c == 1 ? r.y : r.z) = lift_d<S>(x);
    }
    return r;
}
// Bitmap<channels>::eval(uv) (reference src/core/bitmap.cpp:46-131): rotate about the centre, flip_v = true, scale about
// the centre, translate, wrap, bilinear.  cos / sin of the rotation are evaluated once in fp32 (std::cos / std::sin)
template <class S> static V3<S> tex_eval(const Bsdf::Tex &t, V2<S> uv) {
    const int w = t.w, h = t.h;
    const float crv = std::cos(t.rot.v), srv = std::sin(t.rot.v);
    const S cr = lift_d<S>(Dual(crv, -srv * t.rot.d)), sr = lift_d<S>(Dual(srv, crv * t.rot.d)), sc = lift_d<S>(t.scale);
    const S ux = uv.x - 0.5f, uy = uv.y - 0.5f;
    S rx = ux * cr + uy * sr, ry = -ux * sr + uy * cr;
    rx = rx + 0.5f;
    ry = -(ry + 0.5f);
    rx = rx * sc;
    ry = ry * sc;
    const S off = sc * 0.5f + (-0.5f);
    rx = rx - off;
    ry = ry + off;
    rx = rx + lift_d<S>(t.tx);
    ry = ry + lift_d<S>(t.ty);
    uv = V2<S>(rx - floor_(rx), ry - floor_(ry));
    uv.x = uv.x * (float) (w - 1);
    uv.y = uv.y * (float) (h - 1);
    int px = (int) std::floor(val(uv.x)), py = (int) std::floor(val(uv.y));
    S w1x = uv.x - (float) px, w1y = uv.y - (float) py;
    S w0x = 1.0f - w1x, w0y = 1.0f - w1y;
    px = std::min(px, w - 2);
    py = std::min(py, h - 2);
    int i00 = py * w + px;
    V3<S> v00 = tex_texel<S>(t, i00), v10 = tex_texel<S>(t, i00 + 1), v01 = tex_texel<S>(t, i00 + w), v11 = tex_texel<S>(t, i00 + w + 1);
    V3<S> v0(fmadd(w0x, v00.x, w1x * v10.x), fmadd(w0x, v00.y, w1x * v10.y), fmadd(w0x, v00.z, w1x * v10.z));
    V3<S> v1(fmadd(w0x, v01.x, w1x * v11.x), fmadd(w0x, v01.y, w1x * v11.y), fmadd(w0x, v01.z, w1x * v11.z));
    return V3<S>(fmadd(w0y, v0.x, w1y * v1.x), fmadd(w0y, v0.y, w1y * v1.y), fmadd(w0y, v0.z, w1y * v1.z));
}
template <class S> static V3<S> refl_of(const Bsdf &b, V2<S> uv) {
    if (b.tex[0].w <= 0) return refl_const<S>(b);
    return tex_eval<S>(b.tex[0], uv);
}

template <class S> static V3<S> spec_const(const Bsdf &b);
template <> V3<Dual> spec_const<Dual>(const Bsdf &b) { return b.specular; }
template <> V3<float> spec_const<float>(const Bsdf &b) { return val(b.specular); }
template <class S> static V3<S> spec_of(const Bsdf &b, V2<S> uv) { return b.tex[1].w > 0 ? tex_eval<S>(b.tex[1], uv) : spec_const<S>(b); }
template <class S> static S rough_of(const Bsdf &b, V2<S> uv) { return b.tex[2].w > 0 ? tex_eval<S>(b.tex[2], uv).x : lift_d<S>(b.roughness); }

// GGXDistribution::eval (reference src/bsdf/ggx.cpp:13-33), alpha_u = alpha_v
template <class S> static S ggx_eval(S alpha, V3<S> m) {
    S alpha_uv = alpha * alpha;
    S t = sqr(m.x / alpha) + sqr(m.y / alpha) + sqr(m.z);
    S result = rcp_(S(kPi) * alpha_uv * sqr(t));
    return (val(result) * val(m.z) > 1e-20f) ? result : S(0.f);
}
// GGXDistribution::smith_g1 (ggx.cpp:82-96)
template <class S> static S ggx_smith_g1(S alpha, V3<S> v, V3<S> m) {
    S xy_alpha_2 = sqr(alpha * v.x) + sqr(alpha * v.y);
    S tan_theta_alpha_2 = xy_alpha_2 / sqr(v.z);
    S result = S(2.f) / (S(1.f) + sqrt_(S(1.f) + tan_theta_alpha_2));
    if (val(xy_alpha_2) == 0.f) result = S(1.f);
    if (val(dot(v, m)) * val(v.z) <= 0.f) result = S(0.f);
    return result;
}
// Microfacet::__eval (reference src/bsdf/microfacet.cpp:22-68)
template <class S> static V3<S> microfacet_eval(const Bsdf &b, V3<S> wi, V3<S> wo, V2<S> uv) {
    if (b.two_side) {
        if (std::signbit(val(wi.z))) wo.z = -wo.z;
        wi.z = abs_(wi.z);
    }
    S cos_nv = wi.z, cos_nl = wo.z;
    if (!(val(cos_nv) > 0.f && val(cos_nl) > 0.f)) return V3<S>(S(0.f));
    V3<S> diffuse = refl_of<S>(b, uv) * S(kInvPi);
    V3<S> H = normalize(wi + wo);
    S cos_vh = dot(H, wi);
    V3<S> F0 = spec_of<S>(b, uv);
    S alpha = sqr(rough_of<S>(b, uv));
    S ggx = ggx_eval<S>(alpha, H);
    S coeff = cos_vh * (S(-5.55473f) * cos_vh - S(6.8316f));
    S e = exp2_(coeff);
    V3<S> fresnel = F0 + (V3<S>(S(1.f)) - F0) * e;
    S smithG = ggx_smith_g1<S>(alpha, wi, H) * ggx_smith_g1<S>(alpha, wo, H);
    V3<S> numerator = fresnel * (ggx * smithG);
    S denominator = S(4.f) * cos_nl * cos_nv;
    V3<S> specular = numerator / (denominator + S(1e-6f));
    return (diffuse + specular) * cos_nl;
}
// fresnel<ad>(eta_r, eta_i, cos_theta_i) for conductors (reference include/psdr/utils.h:167-183), one channel
template <class S> static S fresnel_conductor(S eta_r, S eta_i, S cos_theta_i) {
    S cos_theta_i_2 = sqr(cos_theta_i), sin_theta_i_2 = S(1.f) - cos_theta_i_2, sin_theta_i_4 = sqr(sin_theta_i_2);
    S temp_1 = sqr(eta_r) - sqr(eta_i) - sin_theta_i_2;
    S a_2_pb_2 = safe_sqrt(sqr(temp_1) + S(4.f) * sqr(eta_i * eta_r));
    S a = safe_sqrt(S(.5f) * (a_2_pb_2 + temp_1));
    S term_1 = a_2_pb_2 + cos_theta_i_2, term_2 = S(2.f) * cos_theta_i * a;
    S r_s = (term_1 - term_2) / (term_1 + term_2);
    S term_3 = a_2_pb_2 * cos_theta_i_2 + sin_theta_i_4, term_4 = term_2 * sin_theta_i_2;
    S r_p = r_s * (term_3 - term_4) / (term_3 + term_4);
    return S(.5f) * (r_s + r_p);
}
template <class S> static V3<S> lift_v3(const V3d &v);
template <> V3<Dual> lift_v3<Dual>(const V3d &v) { return v; }
template <> V3<float> lift_v3<float>(const V3d &v) { return val(v); }
// RoughConductor::__eval (reference src/bsdf/roughconductor.cpp:38-66), isotropic
template <class S> static V3<S> conductor_eval(const Bsdf &b, V3<S> wi, V3<S> wo, V2<S> uv) {
    if (b.two_side) {
        if (std::signbit(val(wi.z))) wo.z = -wo.z;
        wi.z = abs_(wi.z);
    }
    if (!(val(wi.z) > 0.f && val(wo.z) > 0.f)) return V3<S>(S(0.f));
    S alpha = rough_of<S>(b, uv);
    V3<S> H = normalize(wo + wi);
    S D = ggx_eval<S>(alpha, H);
    if (val(D) == 0.f) return V3<S>(S(0.f));
    S G = ggx_smith_g1<S>(alpha, wi, H) * ggx_smith_g1<S>(alpha, wo, H);
    S result = D * G / (S(4.f) * wi.z);
    S c = dot(wi, H);
    V3<S> eta = lift_v3<S>(b.eta), k = lift_v3<S>(b.k);
    V3<S> F(fresnel_conductor<S>(eta.x, k.x, c), fresnel_conductor<S>(eta.y, k.y, c), fresnel_conductor<S>(eta.z, k.z, c));
    return F * result * spec_of<S>(b, uv);
}
// fresnel_dielectric (reference include/psdr/utils.h:185-215)
template <class S> struct FresnelDielectric {
    S r, cos_theta_t, eta_it, eta_ti;
};
template <class S> static FresnelDielectric<S> fresnel_dielectric(S eta, S cos_theta_i) {
    FresnelDielectric<S> f;
    bool outside_mask = val(cos_theta_i) >= 0.f;
    S rcp_eta = rcp_(eta);
    f.eta_it = outside_mask ? eta : rcp_eta;
    f.eta_ti = outside_mask ? rcp_eta : eta;
    S cos_theta_t_sqr = fmadd(-fmadd(-cos_theta_i, cos_theta_i, S(1.f)), f.eta_ti * f.eta_ti, S(1.f));
    S cos_theta_i_abs = abs_(cos_theta_i), cos_theta_t_abs = safe_sqrt(cos_theta_t_sqr);
    bool index_matched = val(eta) == 1.f, special_case = index_matched || val(cos_theta_i_abs) == 0.f;
    S a_s = fmadd(-f.eta_it, cos_theta_t_abs, cos_theta_i_abs) / fmadd(f.eta_it, cos_theta_t_abs, cos_theta_i_abs);
    S a_p = fmadd(-f.eta_it, cos_theta_i_abs, cos_theta_t_abs) / fmadd(f.eta_it, cos_theta_i_abs, cos_theta_t_abs);
    f.r = S(.5f) * (sqr(a_s) + sqr(a_p));
    if (special_case) f.r = S(index_matched ? 0.f : 1.f);
    f.cos_theta_t = std::signbit(val(cos_theta_i)) ? cos_theta_t_abs : -cos_theta_t_abs;   // mulsign_neg
    return f;
}
// RoughDielectric::__eval (reference src/bsdf/roughdielectric.cpp:37-122), alpha_u = alpha_v
template <class S> static V3<S> dielectric_eval(const Bsdf &b, V3<S> wi, V3<S> wo, V2<S> uv) {
    if (b.two_side) {
        if (std::signbit(val(wi.z))) wo.z = -wo.z;
        wi.z = abs_(wi.z);
    }
    S cos_theta_i = wi.z, cos_theta_o = wo.z;
    if (val(cos_theta_i) == 0.f) return V3<S>(S(0.f));
    bool reflect = val(cos_theta_i) * val(cos_theta_o) > 0.f;
    S m_eta = lift_d<S>(b.eta.x), m_inv_eta = lift_d<S>(b.eta.y);
    S eta = val(cos_theta_i) > 0.f ? m_eta : m_inv_eta, inv_eta = val(cos_theta_i) > 0.f ? m_inv_eta : m_eta;
    V3<S> m = normalize(wi + wo * (reflect ? S(1.f) : eta));
    if (std::signbit(val(m.z))) m = -m;
    S alpha = rough_of<S>(b, uv);
    S D = ggx_eval<S>(alpha, m);
    S F = fresnel_dielectric<S>(m_eta, dot(wi, m)).r;
    S G = ggx_smith_g1<S>(alpha, wi, m) * ggx_smith_g1<S>(alpha, wo, m);
    if (reflect) return V3<S>(F * D * G / (S(4.f) * abs_(cos_theta_i)));
    S scale = sqr(inv_eta);
    S wi_m = dot(wi, m), wo_m = dot(wo, m);
    S value = abs_((scale * (S(1.f) - F) * D * G * eta * eta * wi_m * wo_m) / (cos_theta_i * sqr(wi_m + eta * wo_m)));
    return V3<S>(value);
}
// MicrofacetPerVertex::__interpolate (reference src/bsdf/microfacet_pv.cpp:146-160)
template <class S> static S pv_interp(const Scene &sc, const Bsdf &b, const Its<S> &its, int c) {
    auto at = [&](int vtx) -> S {
        if (7 * vtx + c >= (int) b.pv.size()) return S(0.f);
        return lift_d<S>(Dual(b.pv[7 * vtx + c], b.d_pv.empty() ? 0.f : b.d_pv[7 * vtx + c]));
    };
    S v0 = at(sc.tri_fidx[3 * its.tri]), v1 = at(sc.tri_fidx[3 * its.tri + 1]), v2 = at(sc.tri_fidx[3 * its.tri + 2]);
    return fmadd(v1 - v0, its.bc.x, fmadd(v2 - v0, its.bc.y, v0));
}
// MicrofacetPerVertex::__eval (reference src/bsdf/microfacet_pv.cpp:20-68)
template <class S> static V3<S> microfacet_pv_eval(const Scene &sc, const Bsdf &b, const Its<S> &its, V3<S> wi, V3<S> wo) {
    if (b.two_side) {
        if (std::signbit(val(wi.z))) wo.z = -wo.z;
        wi.z = abs_(wi.z);
    }
    S cos_nv = wi.z, cos_nl = wo.z;
    if (!(val(cos_nv) > 0.f && val(cos_nl) > 0.f)) return V3<S>(S(0.f));
    V3<S> F0(pv_interp<S>(sc, b, its, 0), pv_interp<S>(sc, b, its, 1), pv_interp<S>(sc, b, its, 2));
    V3<S> diffuse = V3<S>(pv_interp<S>(sc, b, its, 3), pv_interp<S>(sc, b, its, 4), pv_interp<S>(sc, b, its, 5)) * S(kInvPi);
    S roughness = pv_interp<S>(sc, b, its, 6);
    V3<S> H = normalize(wi + wo);
    S cos_nh = H.z, cos_vh = dot(H, wi);
    S alpha = sqr(roughness);
    S k = sqr(roughness + S(1.f)) / S(8.f);
    S tmp = alpha / (cos_nh * cos_nh * (sqr(alpha) - S(1.f)) + S(1.f));
    S ggx = tmp * tmp * S(kInvPi);
    S coeff = cos_vh * (S(-5.55473f) * cos_vh - S(6.8316f));
    V3<S> fresnel = F0 + (V3<S>(S(1.f)) - F0) * exp2_(coeff);
    S smithG1 = cos_nv / (cos_nv * (S(1.f) - k) + k);
    S smithG2 = cos_nl / (cos_nl * (S(1.f) - k) + k);
    S smithG = smithG1 * smithG2;
    V3<S> numerator = fresnel * (ggx * smithG);
    S denominator = S(4.f) * cos_nl * cos_nv;
    V3<S> specular = numerator / (denominator + S(1e-6f));
    return (diffuse + specular) * cos_nl;
}
// BSDF::eval of a record that is not a NormalMap, incident direction given explicitly
template <class S> static V3<S> bsdf_eval_leaf(const Scene &sc, const Bsdf &b, const Its<S> &its, V3<S> wi, V3<S> wo) {
    if (b.type == 1) return microfacet_eval<S>(b, wi, wo, its.uv);
    if (b.type == 2) return conductor_eval<S>(b, wi, wo, its.uv);
    if (b.type == 3) return dielectric_eval<S>(b, wi, wo, its.uv);
    if (b.type == 4) return microfacet_pv_eval<S>(sc, b, its, wi, wo);
    S wiz = wi.z;
    if (b.two_side) {
        if (std::signbit(val(wiz))) wo.z = -wo.z;
        wiz = abs_(wiz);
    }
    if (!(val(wiz) > 0.f && val(wo.z) > 0.f)) return V3<S>(S(0.f));
    V3<S> r = refl_of<S>(b, its.uv);
    return r * S(kInvPi) * wo.z;
}
// ---- NormalMap (reference src/bsdf/normalmap.cpp:17-40 helpers) ----
template <class S> struct NmFrame {   // Frame(n, s), include/psdr/core/frame.h:42-45
    V3<S> s, t, n;
    V3<S> to_local(V3<S> v) const { return {dot(v, s), dot(v, t), dot(v, n)}; }
    V3<S> to_world(V3<S> v) const { return s * v.x + t * v.y + n * v.z; }
};
template <class S> static S nm_pdot(V3<S> a, V3<S> b) {
    S d = dot(a, b);
    return val(d) > 0.f ? d : S(0.f);
}
template <class S> static S nm_sin_theta(V3<S> v) { return safe_sqrt(fmadd(v.x, v.x, sqr(v.y))); }
template <class S> static V3<S> nm_wt(V3<S> wp) { return normalize(V3<S>(-wp.x, -wp.y, S(0.f))); }
template <class S> static S nm_G1(V3<S> wp, V3<S> w) {
    S cw = val(w.z) > 0.f ? w.z : S(0.f), cp = val(wp.z) > 0.f ? wp.z : S(0.f);
    S g = cw * cp / (nm_pdot<S>(w, wp) + nm_pdot<S>(w, nm_wt<S>(wp)) * nm_sin_theta<S>(wp));
    return val(g) < 1.f ? g : S(1.f);
}
template <class S> static S nm_lambda_p(V3<S> wp, V3<S> wi) {
    S i_dot_p = nm_pdot<S>(wp, wi);
    return i_dot_p / (i_dot_p + nm_pdot<S>(nm_wt<S>(wp), wi) * nm_sin_theta<S>(wp));
}
template <class S> static void nm_setup(const Bsdf &b, const Its<S> &its, V3<S> &wp, NmFrame<S> &fr) {
    V3<S> c = refl_of<S>(b, its.uv);
    wp = normalize(V3<S>(fmadd(c.x, S(2.f), S(-1.f)), fmadd(c.y, S(2.f), S(-1.f)), fmadd(c.z, S(2.f), S(-1.f))));
    S d = dot(wp, its.dp_du);   // a shading-frame vector against a world-space one, as the reference writes it (normalmap.cpp:61)
    V3<S> s0 = normalize(V3<S>(fmadd(-wp.x, d, its.dp_du.x), fmadd(-wp.y, d, its.dp_du.y), fmadd(-wp.z, d, its.dp_du.z)));
    fr.n = wp;
    fr.t = normalize(cross(wp, s0));
    fr.s = normalize(cross(fr.t, wp));
}
template <class S> static V3<S> nm_reflect(V3<S> w, V3<S> wt) {
    S k = S(2.f) * dot(w, wt);
    return normalize(w - wt * k);
}
// NormalMap::__eval (normalmap.cpp:42-86)
template <class S> static V3<S> normalmap_eval(const Scene &sc, const Bsdf &b, const Its<S> &its, V3<S> wi, V3<S> wo) {
    if (b.two_side) {
        if (std::signbit(val(wi.z))) wo.z = -wo.z;
        wi.z = abs_(wi.z);
    }
    if (!(val(wi.z) > 0.f && val(wo.z) > 0.f)) return V3<S>(S(0.f));
    const Bsdf &nb = sc.bsdfs[b.nested];
    V3<S> wp;
    NmFrame<S> fr;
    nm_setup<S>(b, its, wp, fr);
    V3<S> p_wo = fr.to_local(wo);
    S shadowing = nm_G1<S>(wp, wo);
    S lambda_p = nm_lambda_p<S>(wp, wi);
    V3<S> wt = nm_wt<S>(wp);
    V3<S> value = bsdf_eval_leaf<S>(sc, nb, its, fr.to_local(wi), p_wo) * lambda_p * shadowing;
    if (val(dot(wi, wt)) > 0.f) {
        V3<S> wi_r = nm_reflect<S>(wi, wt);
        value = value + bsdf_eval_leaf<S>(sc, nb, its, fr.to_local(wi_r), p_wo) * ((S(1.f) - lambda_p) * shadowing);
    }
    return value;
}

template <class S> static V3<S> bsdf_eval(const Scene &sc, const Its<S> &its, V3<S> wo, bool active) {
    if (!active || !its.valid) return V3<S>(S(0.f));
    if (sc.meshes[its.mesh].bsdf < 0) return V3<S>(S(0.f));   // bsdf == nullptr (envmap bounding mesh): a Dr.Jit vcall on null yields 0
    const Bsdf &b = sc.bsdfs[sc.meshes[its.mesh].bsdf];
    if (b.type == 5) return normalmap_eval<S>(sc, b, its, its.wi, wo);
    return bsdf_eval_leaf<S>(sc, b, its, its.wi, wo);
}

template <class S> static Its<float> its_detached(const Its<S> &its) {
    Its<float> f;
    f.valid = its.valid; f.mesh = its.mesh; f.tri = its.tri;
    f.uv = V2f(val(its.uv.x), val(its.uv.y));
    f.bc = V2f(val(its.bc.x), val(its.bc.y));
    f.dp_du = val(its.dp_du);
    return f;
}
// alpha of the GGX lobe used by __pdf / __sample (microfacet.cpp:92,123; microfacet_pv.cpp:93,135; roughconductor.cpp:83,107;
// roughdielectric.cpp:158,191)
static float bsdf_alpha(const Scene &sc, const Bsdf &b, const Its<float> &its) {
    if (b.type == 4) return sqr(pv_interp<float>(sc, b, its, 6));
    float r = rough_of<float>(b, its.uv);
    return (b.type == 2 || b.type == 3) ? r : sqr(r);
}
// Microfacet::__pdf (microfacet.cpp:108-133), detached; RoughConductor::__pdf (roughconductor.cpp:70-95) and
// MicrofacetPerVertex::__pdf (microfacet_pv.cpp:120-143) are the same expression with their own alpha
static float ggx_reflect_pdf(float alpha, bool two_side, V3f wi, V3f wo) {
    if (two_side) {
        if (std::signbit(wi.z)) wo.z = -wo.z;
        wi.z = std::fabs(wi.z);
    }
    V3f m = normalize(wo + wi);
    if (!(wi.z > 0.f && wo.z > 0.f && dot(wi, m) > 0.f && dot(wo, m) > 0.f)) return 0.f;
    return ggx_eval<float>(alpha, m) * ggx_smith_g1<float>(alpha, wi, m) / (4.f * wi.z);
}
// RoughDielectric::__pdf (roughdielectric.cpp:125-176)
static float dielectric_pdf(const Bsdf &b, float alpha, V3f wi, V3f wo) {
    if (b.two_side) {
        if (std::signbit(wi.z)) wo.z = -wo.z;
        wi.z = std::fabs(wi.z);
    }
    float cos_theta_i = wi.z, cos_theta_o = wo.z;
    if (cos_theta_i == 0.f) return 0.f;
    bool reflect = cos_theta_i * cos_theta_o > 0.f;
    float eta = cos_theta_i > 0.f ? b.eta.x.v : b.eta.y.v;
    V3f m = normalize(wi + wo * (reflect ? 1.f : eta));
    if (std::signbit(m.z)) m = -m;
    float wi_m = dot(wi, m), wo_m = dot(wo, m);
    if (!(wi_m * wi.z > 0.f && wo_m * wo.z > 0.f)) return 0.f;
    float dwh_dwo = reflect ? 1.f / (4.f * wo_m) : (eta * eta * wo_m) / sqr(wi_m + eta * wo_m);
    V3f pwi = std::signbit(wi.z) ? -wi : wi;
    float prob = ggx_eval<float>(alpha, m) * ggx_smith_g1<float>(alpha, pwi, m) / pwi.z;
    float F = fresnel_dielectric<float>(b.eta.x.v, wi_m).r;
    prob *= reflect ? F : 1.f - F;
    return prob * std::fabs(dwh_dwo);
}
static float bsdf_pdf_leaf(const Scene &sc, const Bsdf &b, const Its<float> &its, V3f wi, V3f wo) {
    if (b.type == 3) return dielectric_pdf(b, bsdf_alpha(sc, b, its), wi, wo);
    if (b.type != 0) return ggx_reflect_pdf(bsdf_alpha(sc, b, its), b.two_side, wi, wo);
    float wiz = wi.z, woz = wo.z;
    if (b.two_side) {
        if (std::signbit(wiz)) woz = -woz;
        wiz = std::fabs(wiz);
    }
    if (!(wiz > 0.f && woz > 0.f)) return 0.f;
    return kInvPi * woz;
}
// NormalMap::__pdf (normalmap.cpp:109-144)
static float normalmap_pdf(const Scene &sc, const Bsdf &b, const Its<float> &its, V3f wi, V3f wo) {
    if (b.two_side) {
        if (std::signbit(wi.z)) wo.z = -wo.z;
        wi.z = std::fabs(wi.z);
    }
    if (!(wi.z > 0.f && wo.z > 0.f)) return 0.f;
    const Bsdf &nb = sc.bsdfs[b.nested];
    V3f wp;
    NmFrame<float> fr;
    nm_setup<float>(b, its, wp, fr);
    V3f p_wo = fr.to_local(wo);
    float probability_wp = nm_lambda_p<float>(wp, wi);
    V3f wi_r = nm_reflect<float>(wi, nm_wt<float>(wp));
    return probability_wp * bsdf_pdf_leaf(sc, nb, its, fr.to_local(wi), p_wo) +
           (1.f - probability_wp) * bsdf_pdf_leaf(sc, nb, its, fr.to_local(wi_r), p_wo);
}

template <class S> static float bsdf_pdf(const Scene &sc, const Its<S> &its, V3<S> wo, bool active) {
    if (!active || !its.valid) return 0.f;
    if (sc.meshes[its.mesh].bsdf < 0) return 0.f;
    const Bsdf &b = sc.bsdfs[sc.meshes[its.mesh].bsdf];
    const Its<float> f = its_detached(its);
    if (b.type == 5) return normalmap_pdf(sc, b, f, val(its.wi), val(wo));
    return bsdf_pdf_leaf(sc, b, f, val(its.wi), val(wo));
}

struct BsdfSample {
    V3f wo;
    float pdf = 0.f, eta = 1.f;
    bool valid = false;
};

// reference include/psdr/core/warp.h:15-63
static V2f square_to_uniform_disk_concentric(V2f s) {
    float x = std::fmaf(2.f, s.x, -1.f), y = std::fmaf(2.f, s.y, -1.f);
    bool is_zero = (x == 0.f && y == 0.f), q13 = std::fabs(x) < std::fabs(y);
    float r = q13 ? y : x, rp = q13 ? x : y;
    // sincos(phi): |base phi| <= pi/4, evaluated with fixed polynomials (orc_math.h) so that the
    // CUDA path can reproduce it bit for bit; quadrants 1/3 use sin(pi/2-x)=cos x.
    float sn, cs;
    sincos_quarter(.25f * kPi * rp / r, sn, cs);
    if (is_zero) { sn = 0.f; cs = 1.f; }
    if (q13 && !is_zero) std::swap(sn, cs);
    return V2f(r * cs, r * sn);
}

// GGXDistribution::sample_visible_11 (reference src/bsdf/ggx.cpp:99-109)
static V2f ggx_sample_visible_11(float cos_theta_i, V2f sample) {
    V2f p = square_to_uniform_disk_concentric(sample);
    float s = .5f * (1.f + cos_theta_i);
    float a = safe_sqrt(1.f - sqr(p.x));
    p.y = std::fmaf(p.y, s, std::fmaf(-a, s, a));  // drjit lerp(a, b, t) = fmadd(b, t, fnmadd(a, t, a))
    float x = p.x, y = p.y, z = safe_sqrt(1.f - std::fmaf(p.y, p.y, p.x * p.x));
    float sin_theta_i = safe_sqrt(1.f - sqr(cos_theta_i));
    float norm = 1.f / std::fmaf(sin_theta_i, y, cos_theta_i * z);
    return V2f(std::fmaf(cos_theta_i, y, -(sin_theta_i * z)) * norm, x * norm);
}
// GGXDistribution::sample (ggx.cpp:36-79); sin/cos phi: frame.h:104-122
static void ggx_sample_m(float alpha, V3f wi, V3f sample, V3f &m, float &m_pdf) {
    V3f wi_p = normalize(V3f(alpha * wi.x, alpha * wi.y, wi.z));
    float sin_theta_2 = std::fmaf(wi_p.x, wi_p.x, sqr(wi_p.y)), inv_sin_theta = 1.f / std::sqrt(sin_theta_2);
    bool pole = std::fabs(sin_theta_2) <= 4.f * kEpsilon;
    float sin_phi = pole ? 0.f : std::fmin(std::fmax(wi_p.y * inv_sin_theta, -1.f), 1.f);
    float cos_phi = pole ? 1.f : std::fmin(std::fmax(wi_p.x * inv_sin_theta, -1.f), 1.f);
    V2f slope = ggx_sample_visible_11(wi_p.z, V2f(sample.x, sample.y));
    slope = V2f(std::fmaf(cos_phi, slope.x, -(sin_phi * slope.y)) * alpha, std::fmaf(sin_phi, slope.x, cos_phi * slope.y) * alpha);
    m = normalize(V3f(-slope.x, -slope.y, 1.f));
    m_pdf = ggx_smith_g1<float>(alpha, wi, m) * std::fabs(dot(wi, m)) * ggx_eval<float>(alpha, m) / std::fabs(wi.z);
}
// Microfacet::__sample (microfacet.cpp:80-98); RoughConductor::__sample (roughconductor.cpp:99-122) and
// MicrofacetPerVertex::__sample (microfacet_pv.cpp:82-106) are the same with their own alpha
static BsdfSample ggx_reflect_sample(float alpha, bool two_side, V3f wi, V3f sample, bool active) {
    BsdfSample bs;
    if (two_side) wi.z = std::fabs(wi.z);
    V3f m;
    float m_pdf;
    ggx_sample_m(alpha, wi, sample, m, m_pdf);
    float k = 2.f * dot(wi, m);
    bs.wo = V3f(std::fmaf(m.x, k, -wi.x), std::fmaf(m.y, k, -wi.y), std::fmaf(m.z, k, -wi.z));
    bs.pdf = m_pdf / (4.f * dot(bs.wo, m));
    bs.valid = active && (wi.z > 0.f) && (bs.pdf != 0.f) && (bs.wo.z > 0.f);
    return bs;
}
// RoughDielectric::__sample (roughdielectric.cpp:179-236)
static BsdfSample dielectric_sample(const Bsdf &b, float alpha, V3f wi, V3f sample, bool active) {
    BsdfSample bs;
    if (b.two_side) wi.z = std::fabs(wi.z);
    float cos_theta_i = wi.z;
    active = active && cos_theta_i != 0.f;
    V3f m;
    ggx_sample_m(alpha, std::signbit(cos_theta_i) ? -wi : wi, sample, m, bs.pdf);
    active = active && bs.pdf != 0.f;
    float wi_m = dot(wi, m);
    FresnelDielectric<float> f = fresnel_dielectric<float>(b.eta.x.v, wi_m);
    bool selected_r = sample.z <= f.r && active, selected_t = !selected_r && active;
    bs.pdf *= selected_r ? f.r : 1.f - f.r;
    bs.eta = selected_r ? 1.f : f.eta_it;
    bs.wo = V3f(0.f, 0.f, 0.f);
    if (selected_r) {
        float k = 2.f * wi_m;
        bs.wo = V3f(std::fmaf(m.x, k, -wi.x), std::fmaf(m.y, k, -wi.y), std::fmaf(m.z, k, -wi.z));
    }
    float dwh_dwo = 1.f / (4.f * dot(bs.wo, m));   // assigned for every lane in the reference (:218), overwritten where refracting
    if (selected_t) {
        float k = std::fmaf(wi_m, f.eta_ti, f.cos_theta_t);
        bs.wo = V3f(std::fmaf(m.x, k, -(wi.x * f.eta_ti)), std::fmaf(m.y, k, -(wi.y * f.eta_ti)), std::fmaf(m.z, k, -(wi.z * f.eta_ti)));
        float wo_m = dot(bs.wo, m);
        dwh_dwo = (sqr(bs.eta) * wo_m) / sqr(wi_m + bs.eta * wo_m);
    }
    bs.pdf *= std::fabs(dwh_dwo) * ggx_smith_g1<float>(alpha, bs.wo, m);
    bs.valid = active && (selected_t || selected_r);
    return bs;
}
static BsdfSample bsdf_sample_leaf(const Scene &sc, const Bsdf &b, const Its<float> &its, V3f wi, V3f sample, bool active) {
    if (b.type == 3) return dielectric_sample(b, bsdf_alpha(sc, b, its), wi, sample, active);
    if (b.type != 0) return ggx_reflect_sample(bsdf_alpha(sc, b, its), b.two_side, wi, sample, active);
    BsdfSample bs;
    float wiz = wi.z;
    if (b.two_side) wiz = std::fabs(wiz);
    V2f p = square_to_uniform_disk_concentric(V2f(sample.y, sample.z));  // tail<2>(sample)
    float z = safe_sqrt(1.f - (std::fmaf(p.y, p.y, p.x * p.x)));
    bs.wo = V3f(p.x, p.y, z);
    bs.pdf = kInvPi * z;
    bs.valid = active && (wiz > 0.f);
    return bs;
}
// NormalMap::__sample (normalmap.cpp:147-187)
static BsdfSample normalmap_sample(const Scene &sc, const Bsdf &b, const Its<float> &its, V3f wi, V3f sample, bool active) {
    if (b.two_side) wi.z = std::fabs(wi.z);
    const Bsdf &nb = sc.bsdfs[b.nested];
    V3f wp;
    NmFrame<float> fr;
    nm_setup<float>(b, its, wp, fr);
    V3f p_wi = fr.to_local(wi);
    float probability_wp = nm_lambda_p<float>(wp, wi);
    bool itpo = sample.z >= probability_wp;
    V3f r_wi = fr.to_local(nm_reflect<float>(wi, nm_wt<float>(wp)));
    BsdfSample bs = bsdf_sample_leaf(sc, nb, its, itpo ? r_wi : p_wi, sample, active);
    float pdf1 = bsdf_pdf_leaf(sc, nb, its, p_wi, bs.wo), pdf2 = bsdf_pdf_leaf(sc, nb, its, r_wi, bs.wo);
    bs.pdf = probability_wp * pdf1 + (1.f - probability_wp) * pdf2;
    bs.wo = fr.to_world(bs.wo);
    return bs;
}

template <class S> static BsdfSample bsdf_sample(const Scene &sc, const Its<S> &its, V3f sample, bool active) {
    BsdfSample bs;
    if (!its.valid) return bs;
    if (sc.meshes[its.mesh].bsdf < 0) return bs;
    const Bsdf &b = sc.bsdfs[sc.meshes[its.mesh].bsdf];
    const Its<float> f = its_detached(its);
    if (b.type == 5) return normalmap_sample(sc, b, f, val(its.wi), sample, active);
    return bsdf_sample_leaf(sc, b, f, val(its.wi), sample, active);
}

// ---- emitters (reference src/emitter/area.cpp, src/shape/mesh.cpp:413-466) ---------------
template <class S> static V3<S> radiance_of(const EmitterRec &e);
template <> V3<Dual> radiance_of<Dual>(const EmitterRec &e) { return e.radiance; }
template <> V3<float> radiance_of<float>(const EmitterRec &e) { return val(e.radiance); }

template <class S> static bool is_emitter(const Scene &sc, const Its<S> &its) {
    return its.valid && sc.meshes[its.mesh].emitter >= 0;
}
template <class S> static V3<S> Le(const Scene &sc, const Its<S> &its, bool active) {
    if (!its.valid || sc.meshes[its.mesh].emitter < 0) return V3<S>(S(0.f));
    if (sc.emitters[sc.meshes[its.mesh].emitter].type == 1) {  // EnvironmentMap::eval (envmap.cpp:44-53)
        if (!active) return V3<S>(S(0.f));
        return env_eval_direction<S>(sc.env, -its.to_world(its.wi));
    }
    if (!(active && val(its.wi.z) > 0.f)) return V3<S>(S(0.f));
    return radiance_of<S>(sc.emitters[sc.meshes[its.mesh].emitter]);
}

template <class S> struct PosSample {
    V3<S> p, n;
    S J = S(1.f);
    float pdf = 0.f;
    bool valid = false;
};

// reference src/scene/scene.cpp:987-1013 + mesh.cpp:413-454
// EnvironmentMap::sample_direction + __sample_position (envmap.cpp:86-129), ray_intersect_scene_aabb (utils.h:144-164)
static void env_sample_position(const Envmap &e, V3f ref_p, V2f sample2, V3f &p, V3f &n, float &pdf_out) {
    int ncells = e.cw * e.ch;
    auto r = e.cell.sample_reuse(sample2.y);
    int idx = r.first, cx = idx / e.ch, cy = idx - cx * e.ch;
    float u = (sample2.x + (float) cx) * (1.f / (float) e.cw), v = (sample2.y + (float) cy) * (1.f / (float) e.ch);
    float pdf = r.second * (float) ncells;
    float st, ct, sp, cp;
    sincos_full(v * kPi, st, ct);
    sincos_full(u * (2.f * kPi), sp, cp);
    V3f d(sp * st, ct, -(cp * st));
    float inv_sin_theta = 1.f / std::sqrt(std::fmax(sqr(d.x) + sqr(d.z), sqr(kEpsilon)));
    if (pdf > kEpsilon) pdf *= inv_sin_theta * (.5f / sqr(kPi));
    d = transform_dir(val(e.to_world_full), d);
    float t = 0.f;
    int axis = 0;
    for (int i = 0; i < 3; ++i) {
        float t1 = (e.lower[i] - ref_p[i]) / d[i], t2 = (e.upper[i] - ref_p[i]) / d[i];
        float tm = std::fmax(t1, t2);
        if (i == 0 || tm < t) { t = tm; axis = i; }
    }
    float nn[3] = {0.f, 0.f, 0.f};
    nn[axis] = std::signbit(d[axis]) ? 1.f : -1.f;
    n = V3f(nn[0], nn[1], nn[2]);
    float G = dot(n, -d) * (1.f / sqr(t));
    p = V3f(std::fmaf(d.x, t, ref_p.x), std::fmaf(d.y, t, ref_p.y), std::fmaf(d.z, t, ref_p.z));
    pdf_out = pdf * G;
}
// EnvironmentMap::__sample_position_pdf (envmap.cpp:142-162) + HyperCubeDistribution::pdf (cube_distrb.cpp:51-63)
template <class S> static float env_position_pdf(const Envmap &e, V3f ref_p, const Its<S> &its) {
    V3f d = val(its.p) - ref_p;
    float dist2 = squared_norm(d);
    d = d / safe_sqrt(dist2);
    float G = std::fabs(dot(d, val(its.n))) / dist2;
    d = transform_dir(val(e.from_world), d);
    float factor = G * (1.f / std::sqrt(std::fmax(sqr(d.x) + sqr(d.z), sqr(kEpsilon)))) * (.5f / sqr(kPi));
    V2f uv = dir_to_uv<float>(d);
    int ix = (int) std::floor(uv.x * (float) e.cw), iy = (int) std::floor(uv.y * (float) e.ch);
    if (!(ix >= 0 && ix < e.cw && iy >= 0 && iy < e.ch)) return 0.f;
    return (e.cell.pmf[ix * e.ch + iy] / e.cell.sum) * (float) (e.cw * e.ch) * factor;
}

template <class S> static PosSample<S> sample_emitter_position(const Scene &sc, V3f ref_p, V2f sample2) {
    PosSample<S> ps;
    int ei = 0;
    float emitter_pdf = 1.f;
    if (sc.emitters.size() != 1) {
        auto r = sc.emitter_distrb.sample_reuse(sample2.y);
        ei = r.first;
        emitter_pdf = r.second;
    }
    if (sc.emitters[ei].type == 1) {
        V3f p, n;
        float pdf;
        env_sample_position(sc.env, ref_p, sample2, p, n, pdf);
        ps.p = lift<S>(p);
        ps.n = lift<S>(n);
        ps.pdf = pdf;
        if (sc.emitters.size() != 1) ps.pdf *= emitter_pdf;
        ps.valid = true;
        return ps;
    }
    const MeshRec &m = sc.meshes[sc.emitters[ei].mesh];
    int fi = m.face_distrb.sample_reuse(sample2.x).first;
    float t = safe_sqrt(1.f - sample2.x);
    V2f st(1.f - t, t * sample2.y);
    TriS<S> T = get_tri<S>(sc, m.face_offset + fi);
    constexpr bool ad = std::is_same<S, Dual>::value;
    if (ad) ps.J = T.area / detach(T.area);
    ps.p = bilinear(T.p0, T.e1, T.e2, st);
    ps.n = T.fn;
    ps.pdf = m.inv_total_area;
    if (sc.emitters.size() != 1) ps.pdf *= emitter_pdf;
    ps.valid = true;
    return ps;
}

// reference src/scene/scene.cpp:1016-1024, area.cpp:48-59, mesh.cpp:457-466
template <class S> static float emitter_position_pdf(const Scene &sc, V3f ref_p, const Its<S> &its, bool active) {
    if (!its.valid || !active) return 0.f;
    int e = sc.meshes[its.mesh].emitter;
    if (e < 0) return 0.f;
    if (sc.emitters[e].type == 1) return env_position_pdf<S>(sc.env, ref_p, its);
    return sc.emitters[e].sampling_weight * sc.meshes[its.mesh].inv_total_area;
}

static inline float mis_weight(float a, float b) {
    float w1 = a * a, w2 = b * b;
    return w1 / (w1 + w2);
}

// ---- camera ------------------------------------------------------------------------------
// reference src/sensor/perspective.cpp:160-178
template <class S> static void sample_primary_ray(const Camera &cam, V2f s, V3<S> &o, V3<S> &d);
template <> void sample_primary_ray<float>(const Camera &cam, V2f s, V3f &o, V3f &d) {
    M4<float> s2c = val(cam.sample_to_camera), tw = val(cam.to_world_full);
    if (cam.ortho) {   // OrthographicCamera::sample_primary_ray (orthographic.cpp:109-119)
        o = transform_pos(tw, transform_pos(s2c, V3f(s.x, s.y, 0.f)));
        d = transform_dir(tw, V3f(0.f, 0.f, 1.f));
        return;
    }
    V3f dc = normalize(transform_pos(s2c, V3f(s.x, s.y, 0.f)));
    o = transform_pos(tw, V3f(0.f, 0.f, 0.f));
    d = transform_dir(tw, dc);
}
template <> void sample_primary_ray<Dual>(const Camera &cam, V2f s, V3d &o, V3d &d) {
    M4<float> s2c = val(cam.sample_to_camera);
    if (cam.ortho) {   // orthographic.cpp:122-131
        o = transform_pos(cam.to_world_full, lift<Dual>(transform_pos(s2c, V3f(s.x, s.y, 0.f))));
        d = transform_dir(cam.to_world_full, V3d(Dual(0.f), Dual(0.f), Dual(1.f)));
        return;
    }
    V3f dc = normalize(transform_pos(s2c, V3f(s.x, s.y, 0.f)));
    o = transform_pos(cam.to_world_full, V3d(Dual(0.f), Dual(0.f), Dual(0.f)));
    d = transform_dir(cam.to_world_full, lift<Dual>(dc));
}

struct SensorDirect {
    V2f q;
    int pixel = -1;
    float sensor_val = 0.f;
    bool valid = false;
};
// reference src/sensor/perspective.cpp:181-197
static SensorDirect sample_direct(const Scene &sc, const Camera &cam, V3f p) {
    SensorDirect r;
    V3f q = transform_pos(val(cam.world_to_sample), p);
    r.q = V2f(q.x, q.y);
    int ix = (int) std::floor(q.x * (float) sc.width), iy = (int) std::floor(q.y * (float) sc.height);
    r.valid = ix >= 0 && ix < sc.width && iy >= 0 && iy < sc.height;
    r.pixel = r.valid ? iy * sc.width + ix : -1;
    V3f dir = p - val(cam.pos);
    float dist2 = squared_norm(dir);
    dir = dir / safe_sqrt(dist2);
    float cosTheta = dot(val(cam.dir), dir);
    float ic = 1.f / cosTheta;
    r.sensor_val = (1.f / dist2) * (ic * ic * ic) * cam.inv_area;  // pow(rcp(cos),3)
    return r;
}

// ------------------------------------------------------------------------------------------
// PathTracer::__Li  (reference src/integrator/path.cpp:35-127)
// ------------------------------------------------------------------------------------------
template <class S>
static V3<S> Li(const Scene &sc, Pcg32 &rng, V3<S> ro, V3<S> rd, bool active, int max_depth, bool hide_emitters) {
    constexpr bool ad = std::is_same<S, Dual>::value;
    // primary hit: ray_intersect<ad> (path_space = false)
    Its<S> its = ray_intersect<S>(sc, ro, rd, active, false);
    active = active && its.valid;
    if (sc.mis == 3) {   // CollocatedIntegrator::__Li (collocated.cpp:21-53): evalD(its, its.wi) / sqr(its.t) * m_intensity
        if (!active) return V3<S>(S(0.f));
        if (sc.colloc_field) return bsdf_eval<S>(sc, its, its.wi, true);
        return bsdf_eval<S>(sc, its, its.wi, true) / sqr(its.t) * lift_d<S>(sc.colloc_intensity);
    }
    V3<S> throughput(S(1.f));
    V3<S> result = hide_emitters ? V3<S>(S(0.f)) : Le(sc, its, active);
    for (int depth = 0; depth < max_depth; ++depth) {
        // all lanes draw, masked or not (GCC evaluates next_2d's arguments right-to-left:
        // y gets the first draw; sampler.h:19-21)
        // DirectIntegrator's one-strategy modes draw only what they use (direct.cpp:47,84-91)
        const int mis = sc.mis;
        float s_y = 0.f, s_x = 0.f, s3_z = 0.f, s3_y = 0.f, s3_x = 0.f;
        if (mis != 1) { s_y = rng.next_1d(); s_x = rng.next_1d(); }
        if (mis != 0) { s3_z = rng.next_1d(); s3_y = rng.next_1d(); s3_x = rng.next_1d(); }  // next_nd<3> = (d3,d2,d1)
        if (!active) continue;
        if (mis != 1) {   // ---- emitter sampling
            PosSample<S> ps = sample_emitter_position<S>(sc, val(its.p), V2f(s_x, s_y));
            bool active_direct = active && ps.valid && !is_emitter(sc, its);
            V3<S> wod = ps.p - its.p;
            S dist_sqr = squared_norm(wod);
            S dist = safe_sqrt(dist_sqr);
            wod = wod / dist;
            Its<S> its1 = ray_intersect<S>(sc, its.p, wod, active_direct, ad);
            active_direct = active_direct && its1.valid;
            if (g_diag && active_direct && is_emitter(sc, its1))
                g_diag->m[0] = std::fmin(g_diag->m[0], std::fabs(val(its1.t) - (val(dist) - kShadowEpsilon)));
            active_direct = active_direct && (val(its1.t) > val(dist) - kShadowEpsilon) && is_emitter(sc, its1);
            S cos_val = dot(its1.n, -wod);
            S G_val = abs_(cos_val) / dist_sqr;
            V3<S> emitter_val = Le(sc, its1, active);
            V3<S> wo_local = its.to_local(wod);
            V3<S> bsdf_val2 = bsdf_eval(sc, its, wo_local, active_direct);
            bsdf_val2 = bsdf_val2 * (G_val * ps.J / S(ps.pdf));
            float pdf1 = bsdf_pdf(sc, its, wo_local, active_direct) * val(G_val);
            active_direct = active_direct && (pdf1 != 0.f);
            float weight1 = mis == 0 ? 1.f : mis_weight(ps.pdf, pdf1);
            if (active_direct) result += throughput * emitter_val * bsdf_val2 * S(weight1);
        }
        if (mis != 0) {   // ---- BSDF sampling
            BsdfSample bs = bsdf_sample(sc, its, V3f(s3_x, s3_y, s3_z), active);
            V3<S> wdir = its.to_world(lift<S>(bs.wo));
            Its<S> its1 = ray_intersect<S>(sc, its.p, wdir, active, ad);
            active = active && bs.valid;
            active = active && its1.valid;
            V3<S> bsdf_val;
            float pdf0;
            if (ad) {
                V3<S> wo = its1.p - its.p;
                wo = wo / its1.t;
                S cos_val = dot(its1.n, -wo);
                S G_val = abs_(cos_val) / sqr(its1.t);
                S J = its1.valid ? its1.J : S(1.f);
                if (!its1.valid) G_val = S(1.f);
                pdf0 = bs.pdf * val(G_val);
                if (val(its1.t) < kEpsilon) bsdf_val = V3<S>(S(0.f));
                else bsdf_val = bsdf_eval(sc, its, its.to_local(wo), active) * (G_val * J / S(pdf0));
            } else {
                S cos_val = dot(its1.n, -wdir);
                S G_val = abs_(cos_val) / sqr(its1.t);
                pdf0 = bs.pdf * val(G_val);
                if (val(its1.t) < kEpsilon) bsdf_val = V3<S>(S(0.f));
                else bsdf_val = bsdf_eval(sc, its, lift<S>(bs.wo), active) / S(bs.pdf);
            }
            float weight2 = mis == 1 ? 1.f : mis_weight(pdf0, emitter_position_pdf(sc, val(its.p), its1, active));
            throughput *= bsdf_val;
            if (active) result += Le(sc, its1, active) * throughput * S(weight2);
            its = its1;
        }
    }
    return result;
}

static inline bool finite3(float x) { return std::isfinite(x); }

struct RenderArgs {
    int sensor = 0, max_depth = 1, seed = 0;
    bool hide_emitters = false;
    int skip[3] = {0, 0, 0};  // draws already consumed per lane (seed = -1 continuation)
};

static void splat(float *img, int pix, int c, float v) {
#pragma omp atomic
    img[3 * pix + c] += v;
}

// reference src/integrator/integrator.cpp:104-136 (renderC: ad=false; renderD: ad=true)
template <class S>
static void render_interior(const Scene &sc, const RenderArgs &ra, float *img, float *dimg, const int *pix_id, int npix_sel,
                            float *lane_out, float *diag_out = nullptr) {
    const Camera &cam = sc.cameras[ra.sensor];
    int64_t npix = pix_id ? npix_sel : (int64_t) sc.width * sc.height;
    int64_t N = npix * sc.spp;
    float inv_spp = sc.spp > 1 ? 1.f / (float) sc.spp : 1.f;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < N; ++i) {
        LaneDiag diag;
        g_diag = diag_out ? &diag : nullptr;
        int64_t idx = sc.spp > 1 ? i / sc.spp : i;
        int pix = pix_id ? pix_id[idx] : (int) idx;
        uint64_t seed_value = pix_id ? (uint64_t) ((int64_t) pix + ra.seed) : (uint64_t) (i + ra.seed);
        Pcg32 rng = make_sampler(seed_value, (uint64_t) i);
        for (int k = 0; k < ra.skip[0]; ++k) rng.next_1d();
        float jy = rng.next_1d(), jx = rng.next_1d();
        float sx = ((float) (pix % sc.width) + jx) / (float) sc.width;
        float sy = ((float) (pix / sc.width) + jy) / (float) sc.height;
        V3<S> o, d;
        sample_primary_ray<S>(cam, V2f(sx, sy), o, d);
        V3<S> v = Li<S>(sc, rng, o, d, true, ra.max_depth, ra.hide_emitters);
        for (int c = 0; c < 3; ++c) {
            float x = val(v[c]), dx = tan_(v[c]);
            if (!std::isfinite(x)) { x = 0.f; dx = 0.f; }
            if (lane_out) lane_out[3 * i + c] = x;
            splat(img, (int) idx, c, x * inv_spp);   // sum then divide in the reference; see tests' tolerance
            if (dimg) splat(dimg, (int) idx, c, dx * inv_spp);
        }
        if (diag_out) for (int k = 0; k < 4; ++k) diag_out[4 * i + k] = diag.m[k];
        g_diag = nullptr;
    }
}

// reference src/sensor/perspective.cpp:200-226 + src/integrator/integrator.cpp:179-198
static void render_primary_edges(const Scene &sc, const RenderArgs &ra, float *dimg) {
    const Camera &cam = sc.cameras[ra.sensor];
    if (!cam.enable_edges || sc.sppe <= 0) return;
    int64_t N = (int64_t) sc.width * sc.height * sc.sppe;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < N; ++i) {
        Pcg32 rng = make_sampler((uint64_t) (i + ra.seed), (uint64_t) i);
        for (int k = 0; k < ra.skip[1]; ++k) rng.next_1d();
        float s1 = rng.next_1d();
        auto r = cam.edge_distrb.sample_reuse(s1);
        const PrimEdge &e = cam.edges[r.first];
        float pdf = r.second / e.length;
        V2d p_(fmadd(e.p0.x, Dual(1.0f - s1), e.p1.x * s1), fmadd(e.p0.y, Dual(1.0f - s1), e.p1.y * s1));
        V2f p = val(p_);
        Dual x_dot_n = dot(p_, lift<Dual>(e.normal));
        int ix = (int) std::floor(p.x * (float) sc.width), iy = (int) std::floor(p.y * (float) sc.height);
        bool valid = ix >= 0 && ix < sc.width && iy >= 0 && iy < sc.height;
        int idx = valid ? iy * sc.width + ix : -1;
        V3f op, dp, on, dn;
        sample_primary_ray<float>(cam, V2f(p.x + kEdgeEpsilon * e.normal.x, p.y + kEdgeEpsilon * e.normal.y), op, dp);
        sample_primary_ray<float>(cam, V2f(p.x - kEdgeEpsilon * e.normal.x, p.y - kEdgeEpsilon * e.normal.y), on, dn);
        V3f Lp, Ln;
        if (sc.li_p_first) {
            Lp = Li<float>(sc, rng, op, dp, valid, ra.max_depth, ra.hide_emitters);
            Ln = Li<float>(sc, rng, on, dn, valid, ra.max_depth, ra.hide_emitters);
        } else {
            Ln = Li<float>(sc, rng, on, dn, valid, ra.max_depth, ra.hide_emitters);
            Lp = Li<float>(sc, rng, op, dp, valid, ra.max_depth, ra.hide_emitters);
        }
        if (!valid) continue;
        const float inv_pdf = 1.f / pdf;      // (Ln - Lp) / pdf as a Spectrum / scalar: one reciprocal
        for (int c = 0; c < 3; ++c) {
            float dl = (Ln[c] - Lp[c]) * inv_pdf;
            Dual value = x_dot_n * Dual(dl);
            if (!std::isfinite(value.v)) continue;   // masked(value, ~isfinite(value)) = 0
            float t = value.d;
            if (sc.sppe > 1) t /= (float) sc.sppe;
            splat(dimg, idx, c, t);
        }
    }
}

// reference src/scene/scene.cpp:1027-1068
struct BoundarySeg {
    V3d p0;
    V3f edge, edge2, p2, n;
    float pdf = 0.f;
    bool valid = false;
};
static int sign_eps(float x, float eps) { return x > eps ? 1 : (x < -eps ? -1 : 0); }

static BoundarySeg sample_boundary_segment_direct(const Scene &sc, V3f sample3) {
    BoundarySeg r;
    float sample1 = sample3.x;
    auto er = sc.sec_edge_distrb.sample_reuse(sample1);
    const SecEdge &info = sc.sec_edges[er.first];
    float pdf0 = er.second;
    r.p0 = V3d(fmadd(info.e1.x, sample1, info.p0.x), fmadd(info.e1.y, sample1, info.p0.y), fmadd(info.e1.z, sample1, info.p0.z));
    V3f e1 = val(info.e1);
    r.edge = normalize(e1);
    r.edge2 = info.p2 - val(info.p0);
    V3f p0 = val(r.p0);
    pdf0 /= norm(e1);
    PosSample<float> ps2 = sample_emitter_position<float>(sc, p0, V2f(sample3.y, sample3.z));
    r.p2 = ps2.p;
    r.n = ps2.n;
    V3f e = r.p2 - p0;
    float distSqr = squared_norm(e);
    e = e / safe_sqrt(distSqr);
    float cosTheta = dot(r.n, -e);
    int sgn0 = sign_eps(dot(info.n0, e), kEdgeEpsilon), sgn1 = sign_eps(dot(info.n1, e), kEdgeEpsilon);
    r.valid = (cosTheta > kEpsilon) && ((info.is_boundary && sgn0 != 0) || (!info.is_boundary && sgn0 * sgn1 < 0));
    r.pdf = r.valid ? pdf0 * ps2.pdf * (distSqr / cosTheta) : 0.f;
    return r;
}

static inline float sign1(float x) { return std::signbit(x) ? -1.f : 1.f; }  // drjit::sign

// reference src/integrator/path.cpp:172-270; returns pixel (-1 = invalid), value0 (primal
// boundary value, what the guiding pre-pass accumulates) and the tangent contribution.
static int eval_secondary_edge(const Scene &sc, const Camera &cam, V3f sample3, V3f &value0_out, V3f &tangent_out) {
    value0_out = V3f(0.f, 0.f, 0.f);
    tangent_out = V3f(0.f, 0.f, 0.f);
    BoundarySeg bss = sample_boundary_segment_direct(sc, sample3);
    bool valid = bss.valid;
    V3f _p0 = val(bss.p0), _p2 = bss.p2;
    V3f _dir = normalize(_p2 - _p0);
    int light_tri = -1;
    Its<float> _its2 = ray_intersect<float>(sc, _p0, _dir, valid, false, &light_tri);
    valid = valid && is_emitter(sc, _its2) && _its2.valid && norm(_its2.p - _p2) < kShadowEpsilon;
    Its<float> _its1 = ray_intersect<float>(sc, _p0, -_dir, valid, false);
    valid = valid && _its1.valid;
    V3f _p1 = _its1.p;
    SensorDirect sds = sample_direct(sc, cam, _p1);
    valid = valid && sds.valid;
    V3d co, cd;
    sample_primary_ray<Dual>(cam, sds.q, co, cd);
    Its<Dual> its1 = ray_intersect<Dual>(sc, co, cd, valid, false);
    valid = valid && its1.valid && norm(val(its1.p) - _p1) < kShadowEpsilon;
    valid = valid && its1.valid && sc.meshes[its1.mesh].bsdf >= 0;
    if (!valid) return -1;

    float dist = norm(_p2 - _p1), cos2 = std::fabs(dot(bss.n, -_dir));
    V3f e = cross(bss.edge, _dir);
    float sinphi = norm(e);
    V3f proj = normalize(cross(e, bss.n));
    float sinphi2 = norm(cross(_dir, proj));
    float base_v = (_its1.t / dist) * (sinphi / sinphi2) * cos2;
    valid = valid && (sinphi > kEpsilon) && (sinphi2 > kEpsilon);
    if (!valid) return -1;

    V3f d0 = -val(cd);
    V3f d0_local = _its1.to_local(d0);
    V3f bsdf_val = bsdf_eval<float>(sc, _its1, d0_local, valid);
    float correction = std::fabs((_its1.wi.z * dot(d0, _its1.n)) / (d0_local.z * dot(_dir, _its1.n)));
    bsdf_val = bsdf_val * correction;
    V3f value0 = bsdf_val * Le(sc, _its2, valid) * (base_v * sds.sensor_val / bss.pdf);
    value0_out = value0;

    V3f n = normalize(cross(bss.n, proj));
    value0 = value0 * (sign1(dot(e, bss.edge2)) * sign1(dot(e, n)));
    const Tri<Dual> &T = sc.tris[light_tri];
    V3d sdir = normalize(bss.p0 - its1.p);
    Dual u, v, t;
    ray_intersect_triangle<Dual>(T.p0, T.e1, T.e2, its1.p, sdir, u, v, t);
    V3d u2 = bilinear(detach(T.p0), detach(T.e1), detach(T.e2), V2d(u, v));
    Dual dn = dot(lift<Dual>(n), u2);
    tangent_out = V3f(value0.x * dn.d, value0.y * dn.d, value0.z * dn.d);
    if (getenv("ORC_DEBUG") && !(std::isfinite(tangent_out.x)))
        fprintf(stderr, "sec-edge NaN: value0 %g %g %g dn %g %g base_v %g sensor %g pdf %g corr %g bsdf %g sinphi %g %g u %g %g v %g %g t %g %g\n", value0.x, value0.y, value0.z,
                dn.v, dn.d, base_v, sds.sensor_val, bss.pdf, correction, bsdf_val.x, sinphi, sinphi2, u.v, u.d, v.v, v.d, t.v, t.d);
    return sds.pixel;
}

// reference src/integrator/path.cpp:274-294 (no guiding distribution: pdf0 = 1)
// HyperCubeDistribution<3>::sample_reuse (reference src/core/cube_distrb.cpp:41-48): the cell is chosen with
// the LAST sample dimension, the sample becomes (cell + sample) * unit, pdf = pmf * num_cells
static float guide_sample_reuse(const Camera &cam, V3f &s) {
    auto r = cam.guide.sample_reuse(s.z);
    int idx = r.first;
    int c0 = idx / (cam.greso[1] * cam.greso[2]);
    int rem = idx - c0 * (cam.greso[1] * cam.greso[2]);
    int c1 = rem / cam.greso[2], c2 = rem - c1 * cam.greso[2];
    s.x = (s.x + (float) c0) * (1.f / (float) cam.greso[0]);
    s.y = (s.y + (float) c1) * (1.f / (float) cam.greso[1]);
    s.z = (s.z + (float) c2) * (1.f / (float) cam.greso[2]);
    return r.second * (float) (cam.greso[0] * cam.greso[1] * cam.greso[2]);
}

static void render_secondary_edges(const Scene &sc, const RenderArgs &ra, float *dimg) {
    if (sc.sppse <= 0) return;
    const Camera &cam = sc.cameras[ra.sensor];
    int64_t N = (int64_t) sc.width * sc.height * sc.sppse;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < N; ++i) {
        Pcg32 rng = make_sampler((uint64_t) (i + ra.seed), (uint64_t) i);
        for (int k = 0; k < ra.skip[2]; ++k) rng.next_1d();
        float d1 = rng.next_1d(), d2 = rng.next_1d(), d3 = rng.next_1d();
        V3f sample3(d3, d2, d1);
        float pdf0 = 1.f;
        if (cam.guided) pdf0 = guide_sample_reuse(cam, sample3);   // path.cpp:279-281
        V3f value0, tangent;
        int pix = eval_secondary_edge(sc, cam, sample3, value0, tangent);
        if (pix < 0) continue;
        for (int c = 0; c < 3; ++c) {
            float t = tangent[c];
            if (pdf0 > kEpsilon) t /= pdf0;                          // masked(value, pdf0 > Epsilon) /= pdf0
            // deviation: the reference leaves non-finite tangents in (its scrub is commented out,
            // path.cpp:284); a shadow ray parallel to the emitter triangle gives 0*inf here.
            if (!std::isfinite(t)) continue;
            if (sc.sppse > 1) t /= (float) sc.sppse;
            splat(dimg, pix, c, t);
        }
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C API (ctypes)
// ------------------------------------------------------------------------------------------
static M4<Dual> load_m4(const float *v, const float *d) {
    M4<Dual> m = M4<Dual>::identity();
    if (v)
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) m.m[i][j] = Dual(v[4 * i + j], d ? d[4 * i + j] : 0.f);
    return m;
}

extern "C" {

void *orc_create(int width, int height, int spp, int sppe, int sppse) {
    Scene *s = new Scene();
    s->width = width; s->height = height; s->spp = spp; s->sppe = sppe; s->sppse = sppse;
    return s;
}
void orc_destroy(void *h) { delete (Scene *) h; }
const char *orc_error(void *h) { return ((Scene *) h)->error.c_str(); }
void orc_set_li_order(void *h, int p_first) { ((Scene *) h)->li_p_first = p_first != 0; }
void orc_set_mis(void *h, int mis) { ((Scene *) h)->mis = mis; }
void orc_set_collocated(void *h, float intensity, float d_intensity, int bsdf_field) {
    ((Scene *) h)->mis = 3;
    ((Scene *) h)->colloc_intensity = Dual(intensity, d_intensity);
    ((Scene *) h)->colloc_field = bsdf_field != 0;
}

int orc_add_diffuse(void *h, const float *refl, const float *d_refl, int two_side) {
    Scene *s = (Scene *) h;
    Bsdf b;
    b.reflectance = V3d(Dual(refl[0], d_refl ? d_refl[0] : 0.f), Dual(refl[1], d_refl ? d_refl[1] : 0.f),
                        Dual(refl[2], d_refl ? d_refl[2] : 0.f));
    b.two_side = two_side != 0;
    s->bsdfs.push_back(b);
    return (int) s->bsdfs.size() - 1;
}

// Bitmap texture for slot `slot` of BSDF `bsdf` (0 reflectance / diffuseReflectance, 1 specularReflectance, 2 roughness):
// data [h*w*channels], optional tangents; xform = (scale, rotate, tx, ty), d_xform its tangent (NULL: identity / zero)
int orc_set_bsdf_texture_slot(void *h, int bsdf, int slot, int w, int hh, const float *data, const float *ddata, const float *xform,
                              const float *d_xform) {
    Scene *s = (Scene *) h;
    Bsdf::Tex &t = s->bsdfs[bsdf].tex[slot];
    t.ch = slot == 2 ? 1 : 3;
    t.w = w; t.h = hh;
    t.data.assign(data, data + (size_t) t.ch * w * hh);
    if (ddata) t.ddata.assign(ddata, ddata + (size_t) t.ch * w * hh); else t.ddata.clear();
    const float id[4] = {1.f, 0.f, 0.f, 0.f}, z[4] = {0.f, 0.f, 0.f, 0.f};
    const float *x = xform ? xform : id, *dx = d_xform ? d_xform : z;
    t.scale = Dual(x[0], dx[0]); t.rot = Dual(x[1], dx[1]); t.tx = Dual(x[2], dx[2]); t.ty = Dual(x[3], dx[3]);
    return 0;
}
int orc_set_bsdf_texture(void *h, int bsdf, int w, int hh, const float *data, const float *ddata) {
    return orc_set_bsdf_texture_slot(h, bsdf, 0, w, hh, data, ddata, nullptr, nullptr);
}

// MicrofacetBSDF(specular, diffuse, roughness); d = [d_spec(3), d_diff(3), d_rough(1)] or NULL
int orc_add_microfacet(void *h, const float *spec, const float *diff, float rough, const float *d, int two_side) {
    Scene *s = (Scene *) h;
    Bsdf b;
    b.type = 1;
    b.specular = V3d(Dual(spec[0], d ? d[0] : 0.f), Dual(spec[1], d ? d[1] : 0.f), Dual(spec[2], d ? d[2] : 0.f));
    b.reflectance = V3d(Dual(diff[0], d ? d[3] : 0.f), Dual(diff[1], d ? d[4] : 0.f), Dual(diff[2], d ? d[5] : 0.f));
    b.roughness = Dual(rough, d ? d[6] : 0.f);
    b.two_side = two_side != 0;
    s->bsdfs.push_back(b);
    return (int) s->bsdfs.size() - 1;
}

// RoughConductorBSDF(alpha, eta, k[, specular_reflectance]); d = [d_alpha, d_eta(3), d_k(3), d_spec(3)] or NULL
int orc_add_roughconductor(void *h, float alpha, const float *eta, const float *k, const float *spec, const float *d, int two_side) {
    Scene *s = (Scene *) h;
    Bsdf b;
    b.type = 2;
    b.roughness = Dual(alpha, d ? d[0] : 0.f);
    b.eta = V3d(Dual(eta[0], d ? d[1] : 0.f), Dual(eta[1], d ? d[2] : 0.f), Dual(eta[2], d ? d[3] : 0.f));
    b.k = V3d(Dual(k[0], d ? d[4] : 0.f), Dual(k[1], d ? d[5] : 0.f), Dual(k[2], d ? d[6] : 0.f));
    b.specular = V3d(Dual(spec[0], d ? d[7] : 0.f), Dual(spec[1], d ? d[8] : 0.f), Dual(spec[2], d ? d[9] : 0.f));
    b.reflectance = V3d(Dual(0.f), Dual(0.f), Dual(0.f));
    b.two_side = two_side != 0;
    s->bsdfs.push_back(b);
    return (int) s->bsdfs.size() - 1;
}

// RoughDielectricBSDF(alpha, intIOR, extIOR) (roughdielectric.h:20-25: m_eta and m_inv_eta are two separate quotients)
int orc_add_roughdielectric(void *h, float alpha, float d_alpha, float int_ior, float ext_ior, int two_side) {
    Scene *s = (Scene *) h;
    Bsdf b;
    b.type = 3;
    b.roughness = Dual(alpha, d_alpha);
    b.eta = V3d(Dual(int_ior / ext_ior), Dual(ext_ior / int_ior), Dual(0.f));
    b.two_side = two_side != 0;
    s->bsdfs.push_back(b);
    return (int) s->bsdfs.size() - 1;
}
// MicrofacetBSDFPerVertex: pv / d_pv = n*7 floats (specular rgb, diffuse rgb, roughness per vertex); d_pv may be NULL
int orc_add_microfacet_pervertex(void *h, const float *pv, const float *d_pv, int n, int two_side) {
    Scene *s = (Scene *) h;
    Bsdf b;
    b.type = 4;
    b.pv.assign(pv, pv + (size_t) 7 * n);
    if (d_pv) b.d_pv.assign(d_pv, d_pv + (size_t) 7 * n);
    b.two_side = two_side != 0;
    s->bsdfs.push_back(b);
    return (int) s->bsdfs.size() - 1;
}
// NormalMapBSDF around BSDF `nested` (an index returned by an earlier orc_add_*; tests add the nested record LAST so that
// the numbering of the visible BSDFs matches the product's); normal / d_normal: the constant normal map (a bitmap goes
// through orc_set_bsdf_texture_slot, slot 0)
int orc_add_normalmap(void *h, const float *normal, const float *d_normal, int nested, int two_side) {
    Scene *s = (Scene *) h;
    Bsdf b;
    b.type = 5;
    b.reflectance = V3d(Dual(normal[0], d_normal ? d_normal[0] : 0.f), Dual(normal[1], d_normal ? d_normal[1] : 0.f),
                        Dual(normal[2], d_normal ? d_normal[2] : 0.f));
    b.nested = nested;
    b.two_side = two_side != 0;
    s->bsdfs.push_back(b);
    return (int) s->bsdfs.size() - 1;
}
int orc_set_normalmap_nested(void *h, int bsdf, int nested) {
    Scene *s = (Scene *) h;
    if (bsdf < 0 || bsdf >= (int) s->bsdfs.size() || nested < 0 || nested >= (int) s->bsdfs.size()) return -1;
    s->bsdfs[bsdf].nested = nested;
    return 0;
}

// EnvironmentMap: radiance [h*w*3], optional tangents; to_world = (left, raw) 2x16 floats or NULL; returns emitter index
int orc_add_envmap(void *h, const float *data, const float *ddata, int w, int hh, const float *to_world, const float *d_to_world,
                   float scale, float d_scale) {
    Scene *s = (Scene *) h;
    Envmap &e = s->env;
    e.present = true;
    e.w = w; e.h = hh;
    e.data.assign(data, data + (size_t) 3 * w * hh);
    if (ddata) e.ddata.assign(ddata, ddata + (size_t) 3 * w * hh);
    e.scale = Dual(scale, d_scale);
    for (int k = 0; k < 2; ++k) e.to_world[k] = load_m4(to_world ? to_world + 16 * k : nullptr, d_to_world ? d_to_world + 16 * k : nullptr);
    EmitterRec em;
    em.type = 1;
    e.emitter = (int) s->emitters.size();
    s->emitters.push_back(em);
    return e.emitter;
}

// to_world / d_to_world: 3 consecutive row-major 4x4 (left, raw, right); NULL = identity / zero
int orc_add_mesh(void *h, const float *v, const float *dv, int nv, const int *f, int nf, const float *uv, int nuv, const int *fuv,
                 const float *to_world, const float *d_to_world, int bsdf, const float *radiance, const float *d_radiance,
                 int use_face_normals, int enable_edges) {
    Scene *s = (Scene *) h;
    MeshRec m;
    m.v_raw.resize(nv);
    for (int i = 0; i < nv; ++i)
        m.v_raw[i] = V3d(Dual(v[3 * i], dv ? dv[3 * i] : 0.f), Dual(v[3 * i + 1], dv ? dv[3 * i + 1] : 0.f),
                         Dual(v[3 * i + 2], dv ? dv[3 * i + 2] : 0.f));
    m.f.assign(f, f + 3 * nf);
    m.has_uv = nuv > 0;
    if (m.has_uv) {
        m.uv.resize(nuv);
        for (int i = 0; i < nuv; ++i) m.uv[i] = V2f(uv[2 * i], uv[2 * i + 1]);
        m.fuv.assign(fuv, fuv + 3 * nf);
    }
    for (int k = 0; k < 3; ++k) m.to_world[k] = load_m4(to_world ? to_world + 16 * k : nullptr, d_to_world ? d_to_world + 16 * k : nullptr);
    m.bsdf = bsdf;
    m.use_face_normals = use_face_normals != 0;
    m.enable_edges = enable_edges != 0;
    if (radiance) {
        EmitterRec e;
        e.radiance = V3d(Dual(radiance[0], d_radiance ? d_radiance[0] : 0.f), Dual(radiance[1], d_radiance ? d_radiance[1] : 0.f),
                         Dual(radiance[2], d_radiance ? d_radiance[2] : 0.f));
        e.mesh = (int) s->meshes.size();
        m.emitter = (int) s->emitters.size();
        s->emitters.push_back(e);
    }
    s->meshes.push_back(m);
    s->configured = false;
    return (int) s->meshes.size() - 1;
}

int orc_add_camera(void *h, float fov, float near_, float far_, const float *to_world, const float *d_to_world) {
    Scene *s = (Scene *) h;
    Camera c;
    c.fov = fov; c.near_ = near_; c.far_ = far_;
    for (int k = 0; k < 3; ++k) c.to_world[k] = load_m4(to_world ? to_world + 16 * k : nullptr, d_to_world ? d_to_world + 16 * k : nullptr);
    s->cameras.push_back(c);
    s->configured = false;
    return (int) s->cameras.size() - 1;
}

int orc_add_camera_intrinsic(void *h, float fx, float fy, float cx, float cy, float near_, float far_, const float *to_world, const float *d_to_world) {
    int i = orc_add_camera(h, 0.f, near_, far_, to_world, d_to_world);
    Camera &c = ((Scene *) h)->cameras[i];
    c.use_intrinsic = true;
    c.fx = fx; c.fy = fy; c.cx = cx; c.cy = cy;
    return i;
}

int orc_add_camera_orthographic(void *h, float near_, float far_, const float *to_world, const float *d_to_world) {
    int i = orc_add_camera(h, 0.f, near_, far_, to_world, d_to_world);
    ((Scene *) h)->cameras[i].ortho = true;
    return i;
}

int orc_configure(void *h, const int *active, int nactive) { return configure_scene(*(Scene *) h, active, nactive) ? 0 : 1; }

int orc_num_primary_edges(void *h, int sensor) { return (int) ((Scene *) h)->cameras[sensor].edges.size(); }
int orc_num_secondary_edges(void *h) { return (int) ((Scene *) h)->sec_edges.size(); }
int orc_num_mesh_edges(void *h, int mesh) { return (int) ((Scene *) h)->meshes[mesh].edges.size(); }
// out: [4][ne] rows v0, v1, f0, f1 (what the reference's Mesh.edge_indices() returns)
void orc_mesh_edges(void *h, int mesh, int *out) {
    auto &E = ((Scene *) h)->meshes[mesh].edges;
    size_t n = E.size();
    for (size_t i = 0; i < n; ++i) { out[i] = E[i].v0; out[n + i] = E[i].v1; out[2 * n + i] = E[i].f0; out[3 * n + i] = E[i].f1; }
}

// mode 0 = renderC, 1 = renderD.  terms: bit0 interior, bit1 primary edges, bit2 secondary edges.
// img/dimg: [npix*3] (zeroed here).  pix_id may be NULL.  lane_out may be NULL ([N0*3] primal lane values).
int orc_render(void *h, int sensor, int max_depth, int seed, int mode, int terms, int hide_emitters, const int *skip,
               const int *pix_id, int npix_sel, float *img, float *dimg, float *lane_out) {
    Scene &sc = *(Scene *) h;
    if (!sc.configured) { sc.error = "Input scene must be configured!"; return 1; }
    if (sensor < 0 || sensor >= (int) sc.cameras.size()) { sc.error = "Invalid sensor id!"; return 1; }
    if (pix_id && seed < 0) { sc.error = "While using batch rendering, seed must be set!"; return 1; }
    RenderArgs ra;
    ra.sensor = sensor; ra.max_depth = max_depth; ra.seed = seed; ra.hide_emitters = hide_emitters != 0;
    if (skip) for (int k = 0; k < 3; ++k) ra.skip[k] = skip[k];
    int64_t npix = pix_id ? npix_sel : (int64_t) sc.width * sc.height;
    std::fill(img, img + 3 * npix, 0.f);
    if (dimg) std::fill(dimg, dimg + 3 * npix, 0.f);
    if (sc.spp > 0 && (terms & 1)) {
        if (mode == 0) render_interior<float>(sc, ra, img, nullptr, pix_id, npix_sel, lane_out);
        else render_interior<Dual>(sc, ra, img, dimg, pix_id, npix_sel, lane_out);
    }
    if (mode == 1 && dimg) {
        if ((terms & 2) && sc.sppe > 0) render_primary_edges(sc, ra, dimg);
        if ((terms & 4) && sc.sppse > 0 && sc.mis != 3) render_secondary_edges(sc, ra, dimg);   // (Integrator::render_secondary_edges is empty)
    }
    return 0;
}

// interior term only, with the decision margins of every lane: diag_out [N0][4] (see LaneDiag)
int orc_render_diag(void *h, int sensor, int max_depth, int seed, int mode, const int *skip, float *img, float *lane_out, float *diag_out) {
    Scene &sc = *(Scene *) h;
    if (!sc.configured) { sc.error = "Input scene must be configured!"; return 1; }
    RenderArgs ra;
    ra.sensor = sensor; ra.max_depth = max_depth; ra.seed = seed;
    if (skip) for (int k = 0; k < 3; ++k) ra.skip[k] = skip[k];
    int64_t npix = (int64_t) sc.width * sc.height;
    std::fill(img, img + 3 * npix, 0.f);
    std::vector<float> dimg((size_t) 3 * npix, 0.f);
    if (mode == 0) render_interior<float>(sc, ra, img, nullptr, nullptr, 0, lane_out, diag_out);
    else render_interior<Dual>(sc, ra, img, dimg.data(), nullptr, 0, lane_out, diag_out);
    return 0;
}

// PathTracer::preprocess_secondary_edges (reference src/integrator/path.cpp:130-168)
int orc_preprocess_secondary_edges(void *h, int sensor, const int *reso, int nrounds, int seed, float *mass_out) {
    Scene &sc = *(Scene *) h;
    if (!sc.configured) { sc.error = "Scene needs to be configured!"; return 1; }
    Camera &cam = sc.cameras[sensor];
    const int ncells = reso[0] * reso[1] * reso[2];
    const int64_t N = (int64_t) ncells * reso[3];
    std::vector<double> acc(ncells, 0.0);
    std::vector<float> mass(ncells, 0.f);
    cam.guided = false;
    (void) N;
    (void) acc;
    // one cell at a time, its samples in lane order: a fixed summation order (the reference's
    // scatter_reduce order is unspecified)
#pragma omp parallel for schedule(dynamic, 64)
    for (int cell = 0; cell < ncells; ++cell) {
        int c0 = cell / (reso[1] * reso[2]);
        int rem = cell - c0 * (reso[1] * reso[2]);
        int c1 = rem / reso[2], c2 = rem - c1 * reso[2];
        float total = 0.f;
        for (int j = 0; j < nrounds; ++j)
            for (int k2 = 0; k2 < reso[3]; ++k2) {
                int64_t i = (int64_t) cell * reso[3] + k2;
                Pcg32 rng = make_sampler((uint64_t) (i + seed), (uint64_t) i);
                for (int k = 0; k < 3 * j; ++k) rng.next_1d();
                float d1 = rng.next_1d(), d2 = rng.next_1d(), d3 = rng.next_1d();
                V3f s3(((float) c0 + d3) * (1.f / (float) reso[0]), ((float) c1 + d2) * (1.f / (float) reso[1]),
                       ((float) c2 + d1) * (1.f / (float) reso[2]));
                V3f value0, tangent;
                eval_secondary_edge(sc, cam, s3, value0, tangent);
                float m = 0.f;
                for (int c = 0; c < 3; ++c) {
                    float v = value0[c];
                    if (!std::isfinite(v)) v = 0.f;
                    if (reso[3] > 1) v /= (float) reso[3];
                    m = c == 0 ? v : std::fmax(m, v);
                }
                total += m;
            }
        mass[cell] = total;
    }
    if (nrounds > 1) for (float &m : mass) m /= (float) nrounds;
    cam.greso[0] = reso[0]; cam.greso[1] = reso[1]; cam.greso[2] = reso[2];
    cam.guide.init(mass);
    cam.guided = true;
    if (mass_out) std::copy(mass.begin(), mass.end(), mass_out);
    return 0;
}

// Sampler tap: out[ndraws][n] = successive next_1d() of a sampler seeded with arange(n)+seed
void orc_sampler_draws(int64_t seed, int n, int ndraws, float *out) {
    for (int i = 0; i < n; ++i) {
        Pcg32 r = make_sampler((uint64_t) (i + seed), (uint64_t) i);
        for (int k = 0; k < ndraws; ++k) out[(size_t) k * n + i] = r.next_1d();
    }
}

// DiscreteDistribution tap (pmf.cpp:18-28)
void orc_pmf_sample(const float *pmf, int n, const float *samples, int m, int *idx, float *p, float *sum) {
    Distrib d;
    d.init(std::vector<float>(pmf, pmf + n));
    *sum = d.sum;
    for (int i = 0; i < m; ++i) {
        auto r = d.sample(samples[i]);
        idx[i] = r.first;
        p[i] = r.second;
    }
}

// AOV tap = what FieldExtractionIntegrator returns at spp=1 (reference src/integrator/field.cpp:47-121):
// per lane: mesh id+1, tri id, position, depth t, geometric normal, shading normal, uv.  out: [N][14]
void orc_aov(void *h, int sensor, int seed, float *out) {
    Scene &sc = *(Scene *) h;
    const Camera &cam = sc.cameras[sensor];
    int64_t N = (int64_t) sc.width * sc.height * sc.spp;
#pragma omp parallel for
    for (int64_t i = 0; i < N; ++i) {
        int64_t idx = sc.spp > 1 ? i / sc.spp : i;
        Pcg32 rng = make_sampler((uint64_t) (i + seed), (uint64_t) i);
        float jy = rng.next_1d(), jx = rng.next_1d();
        float sx = ((float) (idx % sc.width) + jx) / (float) sc.width, sy = ((float) (idx / sc.width) + jy) / (float) sc.height;
        V3f o, d;
        sample_primary_ray<float>(cam, V2f(sx, sy), o, d);
        Its<float> its = ray_intersect<float>(sc, o, d, true, false);
        float *r = out + 14 * i;
        for (int k = 0; k < 14; ++k) r[k] = 0.f;
        if (!its.valid) { r[1] = -1.f; continue; }
        r[0] = (float) (its.mesh + 1); r[1] = (float) its.tri;
        r[2] = its.p.x; r[3] = its.p.y; r[4] = its.p.z; r[5] = its.t;
        r[6] = its.n.x; r[7] = its.n.y; r[8] = its.n.z;
        r[9] = its.sh_n.x; r[10] = its.sh_n.y; r[11] = its.sh_n.z; r[12] = its.uv.x; r[13] = its.uv.y;
    }
}

// FieldExtractionIntegrator::renderD in forward mode (reference src/integrator/field.cpp:47-121 through Integrator::renderD,
// src/integrator/integrator.cpp:51-100): interior part = the taps of orc_aov from ray_intersect<Dual> with tangents
void orc_aov_d(void *h, int sensor, int seed, float *out, float *dout) {
    Scene &sc = *(Scene *) h;
    const Camera &cam = sc.cameras[sensor];
    int64_t N = (int64_t) sc.width * sc.height * sc.spp;
#pragma omp parallel for
    for (int64_t i = 0; i < N; ++i) {
        int64_t idx = sc.spp > 1 ? i / sc.spp : i;
        Pcg32 rng = make_sampler((uint64_t) (i + seed), (uint64_t) i);
        float jy = rng.next_1d(), jx = rng.next_1d();
        float sx = ((float) (idx % sc.width) + jx) / (float) sc.width, sy = ((float) (idx / sc.width) + jy) / (float) sc.height;
        V3d o, d;
        sample_primary_ray<Dual>(cam, V2f(sx, sy), o, d);
        Its<Dual> its = ray_intersect<Dual>(sc, o, d, true, false);
        float *r = out + 14 * i, *t = dout + 14 * i;
        for (int k = 0; k < 14; ++k) { r[k] = 0.f; t[k] = 0.f; }
        if (!its.valid) { r[1] = -1.f; continue; }
        r[0] = (float) (its.mesh + 1); r[1] = (float) its.tri;
        const Dual f[12] = {its.p.x, its.p.y, its.p.z, its.t, its.n.x, its.n.y, its.n.z, its.sh_n.x, its.sh_n.y, its.sh_n.z, its.uv.x, its.uv.y};
        for (int k = 0; k < 12; ++k) { r[2 + k] = f[k].v; t[2 + k] = std::isfinite(f[k].d) ? f[k].d : 0.f; }
    }
}

static V3f field_value(const Its<float> &its, int field, int object) {
    if (!its.valid || (object >= 0 && its.mesh != object)) return V3f(0.f, 0.f, 0.f);
    switch (field) {
        case 0: return V3f((float) its.mesh, (float) its.mesh, (float) its.mesh);
        case 1: return V3f(1.f, 1.f, 1.f);
        case 2: return its.p;
        case 3: return V3f(its.t, its.t, its.t);
        case 4: return its.n;
        case 5: return its.sh_n;
        default: return V3f(its.uv.x, its.uv.y, 0.f);
    }
}
// primary-edge part (Integrator::render_primary_edges, integrator.cpp:179-198, with Li = the field): dimg[W*H*3], zeroed here
void orc_field_edges(void *h, int sensor, int seed, int field, int object, float *dimg) {
    Scene &sc = *(Scene *) h;
    const Camera &cam = sc.cameras[sensor];
    for (int64_t k = 0; k < (int64_t) sc.width * sc.height * 3; ++k) dimg[k] = 0.f;
    if (!cam.enable_edges || sc.sppe <= 0) return;
    int64_t N = (int64_t) sc.width * sc.height * sc.sppe;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < N; ++i) {
        Pcg32 rng = make_sampler((uint64_t) (i + seed), (uint64_t) i);
        float s1 = rng.next_1d();
        auto r = cam.edge_distrb.sample_reuse(s1);
        const PrimEdge &e = cam.edges[r.first];
        float pdf = r.second / e.length;
        V2d p_(fmadd(e.p0.x, Dual(1.0f - s1), e.p1.x * s1), fmadd(e.p0.y, Dual(1.0f - s1), e.p1.y * s1));
        V2f p = val(p_);
        Dual x_dot_n = dot(p_, lift<Dual>(e.normal));
        int ix = (int) std::floor(p.x * (float) sc.width), iy = (int) std::floor(p.y * (float) sc.height);
        if (!(ix >= 0 && ix < sc.width && iy >= 0 && iy < sc.height)) continue;
        V3f op, dp, on, dn;
        sample_primary_ray<float>(cam, V2f(p.x + kEdgeEpsilon * e.normal.x, p.y + kEdgeEpsilon * e.normal.y), op, dp);
        sample_primary_ray<float>(cam, V2f(p.x - kEdgeEpsilon * e.normal.x, p.y - kEdgeEpsilon * e.normal.y), on, dn);
        V3f Lp = field_value(ray_intersect<float>(sc, op, dp, true, false), field, object);
        V3f Ln = field_value(ray_intersect<float>(sc, on, dn, true, false), field, object);
        const float inv_pdf = 1.f / pdf;
        for (int c = 0; c < 3; ++c) {
            float dl = (Ln[c] - Lp[c]) * inv_pdf;
            Dual value = x_dot_n * Dual(dl);
            if (!std::isfinite(value.v)) continue;
            float t = value.d;
            if (sc.sppe > 1) t /= (float) sc.sppe;
            splat(dimg, iy * sc.width + ix, c, t);
        }
    }
}

int orc_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
