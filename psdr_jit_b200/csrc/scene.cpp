// Host-side Scene::configure for the B200 path tracer.  Produces every table the kernels read:
// triangle records (+ forward-mode tangents), light/edge distributions, camera matrices, the
// primary/secondary edge lists and the BVH2.  Follows reference src/scene/scene.cpp:311-601,
// src/shape/mesh.cpp:23-62,244-382, src/sensor/perspective.cpp:10-152, src/emitter/area.cpp:9-14.
#include <thread>

#include "scene.h"
#include "texture.h"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <stdexcept>

namespace psdr {

// ---------------------------------------------------------------------------------------------
// fp32 sum in the order of Dr.Jit's GPU block reduction (the reference's DiscreteDistribution
// normalises with drjit::sum, ext/drjit/ext/drjit-core/resources/reduce.cuh:12-60): 1024 lanes,
// lane t owns x[t] + x[t+1024] of each 2048-chunk, then strides 512,256,...,1.
// ---------------------------------------------------------------------------------------------
static float block_tree_sum(const std::vector<float> &x) {
    const size_t n = x.size();
    if (n == 0) return 0.f;
    std::vector<float> partials;
    std::vector<float> lane(1024);
    for (size_t base = 0; base < n; base += 2048) {
        for (size_t t = 0; t < 1024; ++t) {
            float acc = 0.f;
            if (base + t < n) {
                acc = acc + x[base + t];
                if (base + t + 1024 < n) acc = acc + x[base + t + 1024];
            }
            lane[t] = acc;
        }
        for (size_t stride = 512; stride > 0; stride /= 2)
            for (size_t t = 0; t < stride; ++t) lane[t] = lane[t] + lane[t + stride];
        partials.push_back(lane[0]);
    }
    return partials.size() == 1 ? partials[0] : block_tree_sum(partials);
}

void Distrib::init(const std::vector<float> &p) {
    if (p.empty()) throw std::runtime_error("DiscreteDistribution: empty distribution!");
    size = (int) p.size();
    pmf = p;
    sum = block_tree_sum(p);
    cmf.resize(p.size());
    double running = 0.0;   // CDF accumulated in double, rounded per entry (pmf.h:19-25)
    for (size_t i = 0; i < p.size(); ++i) {
        if (p[i] < 0.f) throw std::runtime_error("DiscreteDistribution: entries must be non-negative!");
        running += (double) p[i];
        cmf[i] = (float) running;
    }
}

Scene::Scene() {}

int Scene::find_bsdf(const std::string &id) const {
    for (size_t i = 0; i < bsdfs.size(); ++i)
        if (bsdfs[i].id == id) return (int) i;
    return -1;
}

// --- triangle records: p0/e1/e2, area-weighted vertex normals, unit face normal, area ----------
static void make_triangle_records(const std::vector<V3d> &vp, const std::vector<int> &f, std::vector<HTri> &out) {
    const size_t nf = f.size() / 3, nv = vp.size();
    out.resize(nf);
    std::vector<V3d> vn(nv), fcross(nf);
    std::vector<Dual> vw(nv), flen(nf);
    for (size_t i = 0; i < nf; ++i) {
        HTri &t = out[i];
        t.p0 = vp[f[3 * i]];
        t.e1 = vp[f[3 * i + 1]] - t.p0;
        t.e2 = vp[f[3 * i + 2]] - t.p0;
        fcross[i] = cross(t.e1, t.e2);
        flen[i] = norm(fcross[i]);
    }
    for (int corner = 0; corner < 3; ++corner)
        for (size_t i = 0; i < nf; ++i) {
            const int vi = f[3 * i + corner];
            vn[vi] = vn[vi] + fcross[i];
            vw[vi] = vw[vi] + flen[i];
        }
    for (size_t i = 0; i < nv; ++i) vn[i] = normalize(vn[i] / vw[i]);
    for (size_t i = 0; i < nf; ++i) {
        HTri &t = out[i];
        t.n0 = vn[f[3 * i]];
        t.n1 = vn[f[3 * i + 1]];
        t.n2 = vn[f[3 * i + 2]];
        t.fn = fcross[i] / flen[i];
        t.area = flen[i] * 0.5f;
    }
}

// --- unique undirected edges, ordered by (min vertex, max vertex); each keeps the first two
//     faces that use it and the vertex opposite to it in the first face -------------------------
static void make_edge_list(HMesh &m) {
    m.edges.clear();
    m.edges_dirty = false;
    if (!m.enable_edges) return;
    struct Use { int lo, hi, face, opposite; };
    const size_t nf = m.f.size() / 3;
    std::vector<Use> uses;
    uses.reserve(3 * nf);
    for (size_t fi = 0; fi < nf; ++fi)
        for (int c = 0; c < 3; ++c) {
            const int a = m.f[3 * fi + c], b = m.f[3 * fi + (c + 1) % 3], o = m.f[3 * fi + (c + 2) % 3];
            uses.push_back({std::min(a, b), std::max(a, b), (int) fi, o});
        }
    std::stable_sort(uses.begin(), uses.end(), [](const Use &x, const Use &y) { return x.lo != y.lo ? x.lo < y.lo : x.hi < y.hi; });
    for (size_t i = 0; i < uses.size();) {
        size_t j = i;
        while (j < uses.size() && uses[j].lo == uses[i].lo && uses[j].hi == uses[i].hi) ++j;
        HEdge e;
        e.v0 = uses[i].lo;
        e.v1 = uses[i].hi;
        e.f0 = uses[i].face;
        e.f1 = (j - i >= 2) ? uses[i + 1].face : -1;
        e.v2 = uses[i].opposite;
        m.edges.push_back(e);
        i = j;
    }
}

static float luminance(V3f c) { return c.x * .2126f + c.y * .7152f + c.z * .0722f; }

static M4<float> lower(const M4<Dual> &a) {
    M4<float> r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r.m[i][j] = a.m[i][j].v;
    return r;
}
static M4<Dual> raise(const M4<float> &a) {
    M4<Dual> r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r.m[i][j] = Dual(a.m[i][j]);
    return r;
}

// 4x4 inverse by Gauss-Jordan elimination with partial pivoting on the primal values.
template <class S> static M4<S> invert(const M4<S> &a) {
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
                S fct = w[r][c];
                if (val(fct) == 0.f && tang(fct) == 0.f) continue;
                for (int j = 0; j < 8; ++j) w[r][j] = w[r][j] - fct * w[c][j];
            }
    }
    M4<S> r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r.m[i][j] = w[i][j + 4];
    return r;
}

// reference include/psdr/core/transform.h:48-61
static M4<float> perspective_matrix(float fov, float near_, float far_) {
    const float recip = 1.f / (far_ - near_);
    const float tn = std::tan(fov * .5f * (kPi / 180.f)), cot = 1.f / tn;
    M4<float> t = M4<float>::identity();
    t.m[0][0] = cot;
    t.m[1][1] = cot;
    t.m[2][2] = far_ * recip;
    t.m[3][3] = 0.f;
    t.m[2][3] = -near_ * far_ * recip;
    t.m[3][2] = 1.f;
    return t;
}

// reference include/psdr/core/transform.h:63-71: translate(1 - 2cx, 1 - 2cy, 0) * scale(2fx, 2fy, 1) * trafo
static M4<float> perspective_intrinsic_matrix(float fx, float fy, float cx, float cy, float near_, float far_) {
    const float recip = 1.f / (far_ - near_);
    M4<float> t = M4<float>::identity();
    t.m[2][2] = far_ * recip;
    t.m[3][3] = 0.f;
    t.m[2][3] = -near_ * far_ * recip;
    t.m[3][2] = 1.f;
    M4<float> trn = M4<float>::identity(), scl = M4<float>::identity();
    trn.m[0][3] = 1.f - 2.f * cx;
    trn.m[1][3] = 1.f - 2.f * cy;
    scl.m[0][0] = 2.f * fx;
    scl.m[1][1] = 2.f * fy;
    return (trn * scl) * t;
}

static void configure_mesh(HMesh &m) {
    if (m.edges_dirty) make_edge_list(m);
    const M4<Dual> tw = (m.to_world[0] * m.to_world[1]) * m.to_world[2];
    m.v_world.resize(m.v_raw.size());
    for (size_t i = 0; i < m.v_raw.size(); ++i) m.v_world[i] = transform_pos(tw, m.v_raw[i]);
    make_triangle_records(m.v_world, m.f, m.tris);
    std::vector<float> areas(m.tris.size());
    for (size_t i = 0; i < areas.size(); ++i) areas[i] = m.tris[i].area.v;
    m.total_area = block_tree_sum(areas);
    m.inv_total_area = 1.f / m.total_area;
    m.face_distrb.init(areas);
}

static void configure_camera(const Scene &sc, HCamera &cam, bool with_primary_edges) {
    const float aspect = (float) sc.width / (float) sc.height;
    M4<float> scl = M4<float>::identity(), trn = M4<float>::identity();
    scl.m[0][0] = -0.5f;
    scl.m[1][1] = -0.5f * aspect;
    trn.m[0][3] = -1.f;
    trn.m[1][3] = -1.f / aspect;
    if (cam.use_intrinsic) {      // perspective.cpp:15-20: no aspect ratio in this branch
        scl.m[1][1] = -0.5f;
        trn.m[1][3] = -1.f;
    }
    M4<float> proj;
    if (cam.ortho) {          // transform::orthographic (transform.h:73-76): scale(1, 1, 1 / (far - near)) * translate(0, 0, -near)
        M4<float> os = M4<float>::identity(), ot = M4<float>::identity();
        os.m[2][2] = 1.f / (cam.far_ - cam.near_);
        ot.m[2][3] = -cam.near_;
        proj = os * ot;
    } else proj = cam.use_intrinsic ? perspective_intrinsic_matrix(cam.fx, cam.fy, cam.cx, cam.cy, cam.near_, cam.far_)
                                    : perspective_matrix(cam.fov, cam.near_, cam.far_);
    const M4<float> c2s = (scl * trn) * proj;
    cam.sample_to_camera = invert(c2s);
    cam.camera_to_sample = c2s;
    cam.to_world_full = (cam.to_world[0] * cam.to_world[1]) * cam.to_world[2];
    {   // "Sensor transformation should not involve scaling!" (sensor.cpp:12-13)
        const M4<float> t = lower(cam.to_world_full);
        const float det3 = t.m[0][0] * (t.m[1][1] * t.m[2][2] - t.m[1][2] * t.m[2][1]) - t.m[0][1] * (t.m[1][0] * t.m[2][2] - t.m[1][2] * t.m[2][0]) +
                           t.m[0][2] * (t.m[1][0] * t.m[2][1] - t.m[1][1] * t.m[2][0]);
        if (!(std::fabs(det3 - 1.f) < kEpsilon)) throw std::runtime_error("Sensor transformation should not involve scaling!");
    }
    cam.world_to_sample = raise(c2s) * invert(cam.to_world_full);
    cam.pos = transform_pos(cam.to_world_full, V3d(Dual(0.f), Dual(0.f), Dual(0.f)));
    cam.dir = transform_dir(cam.to_world_full, V3d(Dual(0.f), Dual(0.f), Dual(1.f)));
    const M4<float> &s2c = cam.sample_to_camera;
    const V3f v00 = transform_pos(s2c, V3f(0.f, 0.f, 0.f)), v10 = transform_pos(s2c, V3f(1.f, 0.f, 0.f)),
              v11 = transform_pos(s2c, V3f(1.f, 1.f, 0.f)), vc = transform_pos(s2c, V3f(.5f, .5f, 0.f));
    cam.inv_area = (1.f / (norm(v00 - v10) * norm(v11 - v10))) * squared_norm(vc);

    cam.edges.clear();
    if (sc.sppe <= 0 || !with_primary_edges) return;
    const V3f camp = val(cam.pos);
    for (const HMesh &m : sc.meshes) {
        if (!m.enable_edges) continue;
        const size_t before = cam.edges.size();
        for (const HEdge &e : m.edges) {
            const bool two_faces = e.f1 >= 0;
            const V3f zero(0.f, 0.f, 0.f);
            const V3f to_cam0 = normalize(camp - val(m.tris[e.f0].p0));
            const V3f to_cam1 = normalize(camp - (two_faces ? val(m.tris[e.f1].p0) : zero));
            const V3f n0 = val(m.tris[e.f0].fn), n1 = two_faces ? val(m.tris[e.f1].fn) : zero;
            bool uv_seam = false;
            if (m.has_uv) {   // an edge whose faces do not share exactly two uv indices is a seam
                int shared = 0;
                for (int k = 0; k < 3; ++k) {
                    const int a = m.fuv[3 * e.f0 + k];
                    bool hit = false;
                    for (int l = 0; l < 3; ++l) hit |= (a == (two_faces ? m.fuv[3 * e.f1 + l] : 0));
                    shared += hit ? 1 : 0;
                }
                uv_seam = shared != 2;
            }
            bool keep;
            if (m.use_face_normals) {
                const bool skip = two_faces && ((dot(to_cam0, n0) < kEpsilon && dot(to_cam1, n1) < kEpsilon) || dot(n0, n1) > 1.f - kEpsilon);
                keep = !skip || uv_seam;
            } else {
                const bool silhouette = (dot(to_cam0, n0) > kEpsilon) != (dot(to_cam1, n1) > kEpsilon);
                keep = !two_faces || silhouette || uv_seam;
            }
            if (!keep) continue;
            HPrimEdge pe;
            const V3d q0 = transform_pos(cam.world_to_sample, m.v_world[e.v0]);
            const V3d q1 = transform_pos(cam.world_to_sample, m.v_world[e.v1]);
            pe.p0 = V2d(q0.x, q0.y);
            pe.p1 = V2d(q1.x, q1.y);
            pe.mesh = (int) (&m - &sc.meshes[0]);
            pe.v0 = e.v0;
            pe.v1 = e.v1;
            V2f dir2(q1.x.v - q0.x.v, q1.y.v - q0.y.v);
            const float len = norm(dir2);
            dir2.x /= len;
            dir2.y /= len;
            pe.normal = V2f(-dir2.y, dir2.x);
            pe.length = len;
            cam.edges.push_back(pe);
        }
        if (cam.edges.size() == before)   // PSDR_ASSERT(slices(info) > 0), perspective.cpp:113
            throw std::runtime_error("src/sensor/perspective.cpp (113): slices(info) > 0");
    }
    if (!cam.edges.empty()) {
        std::vector<float> lens(cam.edges.size());
        for (size_t i = 0; i < lens.size(); ++i) lens[i] = cam.edges[i].length;
        cam.edge_distrb.init(lens);
    }
}

// ---------------------------------------------------------------------------------------------
// BVH2, binned SAH (16 bins), built on the host from the world-space triangles.  Boxes are padded
// so that the fp32 slab test can never cull a triangle the fp32 triangle test would accept.
// ---------------------------------------------------------------------------------------------
namespace {
struct Box {
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; ++a) { lo[a] = 3.4e38f; hi[a] = -3.4e38f; } }
    void grow(const float *p) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    void grow(const Box &b) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    float half_area() const {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return dx * dy + dy * dz + dz * dx;
    }
};
struct Builder {
    std::vector<Box> tri_box;
    std::vector<float> centroid;   // 3 per triangle
    std::vector<int> &order;
    std::vector<DBvhNode> &nodes;
    int leaf_size;
    float pad;
    Builder(std::vector<int> &o, std::vector<DBvhNode> &n, int ls) : order(o), nodes(n), leaf_size(ls), pad(0.f) {}

    void store_box(DBvhNode &nd, const Box &b) {
        for (int a = 0; a < 3; ++a) {
            nd.lo[a] = b.lo[a] - pad;
            nd.hi[a] = b.hi[a] + pad;
        }
    }
    int build(int first, int count) {
        const int me = (int) nodes.size();
        nodes.push_back(DBvhNode{});
        Box bounds, cb;
        bounds.reset();
        cb.reset();
        for (int i = first; i < first + count; ++i) {
            bounds.grow(tri_box[order[i]]);
            cb.grow(&centroid[3 * order[i]]);
        }
        store_box(nodes[me], bounds);
        auto make_leaf = [&]() {
            std::sort(order.begin() + first, order.begin() + first + count);
            nodes[me].a = first;
            nodes[me].b = -count;
            return me;
        };
        if (count <= leaf_size) return make_leaf();
        constexpr int kBins = 16;
        float best_cost = 3.4e38f;
        int best_axis = -1, best_split = -1;
        for (int axis = 0; axis < 3; ++axis) {
            const float ext = cb.hi[axis] - cb.lo[axis];
            if (!(ext > 0.f)) continue;
            Box bb[kBins];
            int bc[kBins] = {0};
            for (auto &b : bb) b.reset();
            const float scale = (float) kBins / ext;
            for (int i = first; i < first + count; ++i) {
                int k = std::min(kBins - 1, (int) ((centroid[3 * order[i] + axis] - cb.lo[axis]) * scale));
                bc[k]++;
                bb[k].grow(tri_box[order[i]]);
            }
            float right_area[kBins];
            int right_cnt[kBins];
            Box acc;
            acc.reset();
            int cnt = 0;
            for (int k = kBins - 1; k > 0; --k) {
                acc.grow(bb[k]);
                cnt += bc[k];
                right_area[k] = acc.half_area();
                right_cnt[k] = cnt;
            }
            acc.reset();
            cnt = 0;
            for (int k = 0; k < kBins - 1; ++k) {
                acc.grow(bb[k]);
                cnt += bc[k];
                if (cnt == 0 || right_cnt[k + 1] == 0) continue;
                const float cost = acc.half_area() * (float) cnt + right_area[k + 1] * (float) right_cnt[k + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_split = k; }
            }
        }
        int mid;
        if (best_axis < 0) {
            if (count <= 2 * leaf_size) return make_leaf();
            mid = first + count / 2;   // all centroids coincide
        } else {
            const float ext = cb.hi[best_axis] - cb.lo[best_axis];
            const float scale = (float) kBins / ext;
            auto it = std::partition(order.begin() + first, order.begin() + first + count, [&](int t) {
                int k = std::min(kBins - 1, (int) ((centroid[3 * t + best_axis] - cb.lo[best_axis]) * scale));
                return k <= best_split;
            });
            mid = (int) (it - order.begin());
            if (mid == first || mid == first + count) mid = first + count / 2;
        }
        const int l = build(first, mid - first);
        const int r = build(mid, first + count - mid);
        nodes[me].a = l;
        nodes[me].b = r;
        return me;
    }
};
}  // namespace

void build_bvh(const std::vector<HTri> &tris, std::vector<DBvhNode> &nodes, std::vector<int> &order, int leaf_size) {
    const int n = (int) tris.size();
    nodes.clear();
    order.resize(n);
    Builder b(order, nodes, leaf_size);
    b.tri_box.resize(n);
    b.centroid.resize(3 * (size_t) n);
    Box all;
    all.reset();
    for (int i = 0; i < n; ++i) {
        order[i] = i;
        const V3f p0 = val(tris[i].p0), e1 = val(tris[i].e1), e2 = val(tris[i].e2);
        const float v[3][3] = {{p0.x, p0.y, p0.z}, {p0.x + e1.x, p0.y + e1.y, p0.z + e1.z}, {p0.x + e2.x, p0.y + e2.y, p0.z + e2.z}};
        b.tri_box[i].reset();
        for (int k = 0; k < 3; ++k) b.tri_box[i].grow(v[k]);
        for (int a = 0; a < 3; ++a) b.centroid[3 * i + a] = 0.5f * (b.tri_box[i].lo[a] + b.tri_box[i].hi[a]);
        all.grow(b.tri_box[i]);
    }
    const float dx = all.hi[0] - all.lo[0], dy = all.hi[1] - all.lo[1], dz = all.hi[2] - all.lo[2];
    b.pad = 2e-5f * std::sqrt(dx * dx + dy * dy + dz * dz) + 1e-6f;
    nodes.reserve(2 * (size_t) n);
    if (n > 0) b.build(0, n);
}

// ---------------------------------------------------------------------------------------------
// Scene::configure
// ---------------------------------------------------------------------------------------------
void upload_scene(Scene &sc);   // device_upload.cu

// EnvironmentMap part of Scene::configure (reference src/scene/scene.cpp:355-368,384-416,435-485 and
// src/emitter/envmap.cpp:17-41).  Runs after the meshes and sensors are configured.
static void configure_envmap(Scene &sc, bool plain_configure) {
    HEnvmap &env = sc.env;
    if (!env.present) return;
    if (env.w < 2 || env.h < 2) throw std::runtime_error("src/emitter/envmap.cpp (21): width > 1 && height > 1");
    env.to_world_full = env.to_world[0] * env.to_world[1];
    env.from_world = invert(env.to_world_full);
    if (!env.has_bounds) {
        // scene AABB over mesh vertices and camera positions; the reference initialises the upper corner with
        // numeric_limits<float>::min() (the smallest POSITIVE float), kept
        V3f lo(3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f), hi(1.175494351e-38f, 1.175494351e-38f, 1.175494351e-38f);
        auto grow = [&](V3f p) {
            lo = V3f(std::fmin(lo.x, p.x), std::fmin(lo.y, p.y), std::fmin(lo.z, p.z));
            hi = V3f(std::fmax(hi.x, p.x), std::fmax(hi.y, p.y), std::fmax(hi.z, p.z));
        };
        for (const HMesh &m : sc.meshes)
            for (const V3d &p : m.v_world) grow(val(p));
        (void) plain_configure;
        for (const HCamera &c : sc.cameras) grow(val(c.pos));
        const float margin = std::fmin(std::fmin((hi.x - lo.x) * 0.05f, (hi.y - lo.y) * 0.05f), (hi.z - lo.z) * 0.05f);
        env.lower = V3f(lo.x - margin, lo.y - margin, lo.z - margin);
        env.upper = V3f(hi.x + margin, hi.y + margin, hi.z + margin);
        // bounding mesh: 8 corners, 12 triangles, face normals, no edges, no BSDF, emitter = the envmap
        static const int face_data[3][12] = {{0, 0, 1, 1, 2, 2, 0, 0, 0, 0, 4, 4}, {1, 3, 5, 7, 3, 7, 5, 4, 2, 6, 7, 6}, {3, 2, 7, 3, 7, 6, 1, 5, 6, 4, 5, 7}};
        HMesh b;
        const float lo3[3] = {env.lower.x, env.lower.y, env.lower.z}, hi3[3] = {env.upper.x, env.upper.y, env.upper.z};
        for (int i = 0; i < 8; ++i) {
            float c[3];
            for (int j = 0; j < 3; ++j) c[j] = (i & (1 << j)) ? hi3[j] : lo3[j];
            b.v_raw.push_back(V3d(Dual(c[0]), Dual(c[1]), Dual(c[2])));
        }
        for (int k = 0; k < 12; ++k)
            for (int r = 0; r < 3; ++r) b.f.push_back(face_data[r][k]);
        for (auto &M : b.to_world) M = M4<Dual>::identity();
        b.use_face_normals = true;
        b.enable_edges = false;
        b.bsdf = -1;
        b.emitter = env.emitter;
        b.is_bound_mesh = true;
        configure_mesh(b);
        b.face_offset = 0;
        for (const HMesh &m : sc.meshes) b.face_offset += (int) m.tris.size();
        env.mesh = (int) sc.meshes.size();
        sc.emitters[env.emitter].mesh = env.mesh;
        sc.meshes.push_back(std::move(b));
        env.has_bounds = true;
    }
    // cell distribution: luminance * sin(theta) at the centres of 2(w-1) x 2(h-1) cells, x-major.  Depends on the texels
    // only: rebuilt when they changed (2 M cells for a 1024 x 512 map: 107 ms single-threaded in round 1), the cell loop
    // split over the host cores (cells are independent, so the result does not depend on the split)
    if (env.cell_built == env.data_version && env.cw == ((env.w - 1) << 1) && env.ch == ((env.h - 1) << 1)) return;
    env.cw = (env.w - 1) << 1;
    env.ch = (env.h - 1) << 1;
    const size_t ncells = (size_t) env.cw * env.ch;
    std::vector<float> mass(ncells);
    const float ux = 1.f / (float) env.cw, uy = 1.f / (float) env.ch, dtheta = kPi / (float) env.ch;
    auto fill = [&](size_t lo, size_t hi) {
        for (size_t idx = lo; idx < hi; ++idx) {
            const int cx = (int) (idx / env.ch), cy = (int) (idx % env.ch);
            const V2f uv(((float) cx + .5f) * ux, ((float) cy + .5f) * uy);
            const V3f v = bitmap_eval_envmap<float>(env.data.data(), nullptr, env.w, env.h, uv);
            float sn, cs;
            sincos_full(((float) cy + .5f) * dtheta, sn, cs);
            mass[idx] = luminance(v) * sn;
        }
    };
    const unsigned hw = std::thread::hardware_concurrency();
    const size_t nthreads = ncells < (1u << 16) ? 1 : std::min<size_t>(hw ? hw : 4, 32);
    if (nthreads <= 1) fill(0, ncells);
    else {
        std::vector<std::thread> pool;
        for (size_t t = 0; t < nthreads; ++t) pool.emplace_back(fill, ncells * t / nthreads, ncells * (t + 1) / nthreads);
        for (auto &th : pool) th.join();
    }
    env.cell.init(mass);
    env.cell_built = env.data_version;
}

void Scene::configure(const int *active, int nactive) {
    const auto t0 = std::chrono::high_resolution_clock::now();
    configured = false;
    // samplers: (re)seed when the lane count changed (scene.cpp:330-344)
    const long long npix = (long long) width * height;
    const int per[3] = {spp, sppe, sppse};
    for (int k = 0; k < 3; ++k)
        if (per[k] > 0 && samplers[k].sample_count != npix * per[k]) {
            samplers[k].ready = true;
            samplers[k].sample_count = npix * per[k];
            samplers[k].seed = seed;
            samplers[k].consumed = 0;
        }
    if (meshes.empty()) throw std::runtime_error("Missing meshes!");
    if (cameras.empty()) throw std::runtime_error("Missing sensor!");
    int off = 0;
    for (HMesh &m : meshes) {
        if (m.bsdf >= 0 && m.bsdf >= (int) bsdfs.size()) throw std::runtime_error("Unknown BSDF id");
        configure_mesh(m);
        m.face_offset = off;
        off += (int) m.tris.size();
    }
    // with an explicit sensor list only those sensors get primary edges; a plain configure()
    // leaves every sensor without (scene.cpp:381-416)
    for (size_t i = 0; i < cameras.size(); ++i) {
        bool act = false;
        for (int k = 0; k < nactive; ++k) {
            if (active[k] < 0 || active[k] >= (int) cameras.size()) throw std::runtime_error("Invalid sensor id!");
            act |= (active[k] == (int) i);
        }
        configure_camera(*this, cameras[i], act);
    }
    configure_envmap(*this, nactive == 0);
    if (!emitters.empty()) {
        std::vector<float> w;
        double total_weight = 0.0;          // scene.cpp:489-503: the envmap's weight is the sum of the others
        for (HEmitter &e : emitters) {
            e.raw_weight = e.type == 1 ? 0.f : meshes[e.mesh].total_area * luminance(val(e.radiance));
            total_weight += e.raw_weight;
        }
        for (HEmitter &e : emitters)
            if (e.type == 1) e.raw_weight = (float) total_weight;
        for (HEmitter &e : emitters) w.push_back(e.raw_weight);
        emitter_distrb.init(w);
        const float inv_total = 1.f / emitter_distrb.sum;
        for (HEmitter &e : emitters) e.sampling_weight = e.raw_weight * inv_total;
    }
    sec_edges.clear();
    if (sppse > 0) {
        for (const HMesh &m : meshes) {
            if (!m.enable_edges) continue;
            for (const HEdge &e : m.edges) {
                HSecEdge s;
                s.mesh = (int) (&m - &meshes[0]);
                s.v0 = e.v0;
                s.v1 = e.v1;
                s.is_boundary = e.f1 < 0;
                s.p0 = m.v_world[e.v0];
                s.e1 = m.v_world[e.v1] - s.p0;
                s.n0 = val(m.tris[e.f0].fn);
                s.n1 = s.is_boundary ? V3f(0.f, 0.f, 0.f) : val(m.tris[e.f1].fn);
                s.p2 = val(m.v_world[e.v2]);
                sec_edges.push_back(s);
            }
        }
        std::vector<float> lens(sec_edges.size());
        for (size_t i = 0; i < lens.size(); ++i) lens[i] = norm(val(sec_edges[i].e1));
        sec_edge_distrb.init(lens);
    }
    upload_scene(*this);
    configured = true;
    last_configure_ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
}

void Scene::refresh_tables() {
    if (!configured) throw std::runtime_error("Scene needs to be configured!");
    upload_scene(*this);
}

}  // namespace psdr
