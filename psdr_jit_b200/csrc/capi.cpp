// C ABI of libpsdr_b200.so (declared in include/psdr_b200.h).  Translates handles + plain buffers
// into the host Scene (scene.h) and the kernel launchers (kernels.h); converts every C++ exception
// into an error code + psdr_last_error().
#include <cuda_runtime.h>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/psdr_b200.h"
#include "kernels.h"
#include "scene.h"

using namespace psdr;

struct psdr_scene {
    Scene sc;
    // staging for the *_host entry points
    float *d_img = nullptr, *d_dimg = nullptr;
    int *d_pix = nullptr;
    size_t img_cap = 0, pix_cap = 0;
    cudaStream_t stream = nullptr;
    bool next_bsdf_nested = false;        // psdr_scene_begin_nested_bsdf: the next add_bsdf_* call creates a nested record
    // reverse mode: device gradient table + pinned host copy
    float *d_grad = nullptr, *h_grad = nullptr;
    size_t grad_cap = 0;
    // optional per-kernel timing (psdr_scene_enable_timing)
    bool timing = false;
    cudaEvent_t ev[3][2] = {};
    bool ev_used[3] = {false, false, false};
    // the three term kernels of one call run on three streams (TermStreams below)
    int *d_sched = nullptr;               // chunk hand-out counters of the large-CTA kernels (ChunkSched), two ints each: interior forward, interior adjoint, primary edges forward, primary edges adjoint
    // primary- and secondary-edge lane ordering (edge_sort.cu): one set of buffers per kernel that can be in flight
    struct EdgeSort { unsigned short *key = nullptr; int *perm = nullptr, *work = nullptr; size_t cap = 0; } edge_sort[4];   // primary fwd / adjoint, secondary fwd / adjoint
    float *early_img_host = nullptr;      // psdr_render_d_host: copy the primal image out as soon as the interior kernel is done
    cudaStream_t side[2] = {nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
    ~psdr_scene() {
        if (d_sched) cudaFree(d_sched);
        for (auto &e : edge_sort) {
            if (e.key) cudaFree(e.key);
            if (e.perm) cudaFree(e.perm);
            if (e.work) cudaFree(e.work);
        }
        for (auto &q : side) if (q) cudaStreamDestroy(q);
        if (ev_fork) cudaEventDestroy(ev_fork);
        for (auto &e : ev_join) if (e) cudaEventDestroy(e);
        for (auto &p : ev)
            for (auto &e : p)
                if (e) cudaEventDestroy(e);
        if (d_img) cudaFree(d_img);
        if (d_dimg) cudaFree(d_dimg);
        if (d_pix) cudaFree(d_pix);
        if (d_grad) cudaFree(d_grad);
        if (h_grad) cudaFreeHost(h_grad);
        if (stream) cudaStreamDestroy(stream);
    }
};

static thread_local std::string g_error;
static std::atomic<long long> g_launches{0};
static int g_edge_sort_bins = 512;          // psdr_set_edge_sort
namespace psdr { int g_cta_policy = 0; }      // psdr_set_cta_policy; read by the launchers (kernels_impl.cuh use_big_cta)

static int fail(const std::string &msg) {
    g_error = msg;
    return 1;
}
#define PSDR_TRY try {
#define PSDR_CATCH                                                  \
    }                                                               \
    catch (const std::exception &e) { return fail(e.what()); }      \
    catch (...) { return fail("unknown error"); }

static void cuda_ok(cudaError_t e, const char *what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e));
}

static M4<Dual> mat_from(const float *v, const M4<Dual> *keep_tangent = nullptr) {
    M4<Dual> m = M4<Dual>::identity();
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            m.m[i][j].v = v ? v[4 * i + j] : (i == j ? 1.f : 0.f);
            m.m[i][j].d = keep_tangent ? keep_tangent->m[i][j].d : 0.f;
        }
    return m;
}

extern "C" {

const char *psdr_last_error(void) { return g_error.c_str(); }
int psdr_version(void) { return 100; }
long long psdr_kernel_launch_count(void) { return g_launches.load(); }

psdr_scene *psdr_scene_create(int device) {
    try {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count <= 0) {
            g_error = std::string("psdr_b200 needs a CUDA device (no CPU fallback): ") + cudaGetErrorString(e);
            return nullptr;
        }
        if (device < 0 || device >= count) {
            g_error = "invalid CUDA device ordinal";
            return nullptr;
        }
        psdr_scene *s = new psdr_scene();
        s->sc.device = device;
        return s;
    } catch (const std::exception &e) {
        g_error = e.what();
        return nullptr;
    }
}

void psdr_scene_destroy(psdr_scene *s) { delete s; }

int psdr_scene_set_options(psdr_scene *s, int width, int height, int spp, int sppe, int sppse, int log_level) {
    if (!s) return fail("null scene");
    if (width <= 0 || height <= 0 || spp < 0 || sppe < 0 || sppse < 0) return fail("invalid render options");
    Scene &sc = s->sc;
    if (sc.width != width || sc.height != height || sc.spp != spp || sc.sppe != sppe || sc.sppse != sppse) sc.configured = false;
    sc.width = width; sc.height = height; sc.spp = spp; sc.sppe = sppe; sc.sppse = sppse; sc.log_level = log_level;
    return 0;
}

int psdr_scene_set_seed(psdr_scene *s, long long seed) {
    if (!s) return fail("null scene");
    s->sc.seed = seed;
    return 0;
}

int psdr_scene_set_shard(psdr_scene *s, int rank, int world) {
    if (!s) return fail("null scene");
    if (world < 1 || rank < 0 || rank >= world) return fail("invalid shard");
    s->sc.rank = rank;
    s->sc.world = world;
    return 0;
}

int psdr_scene_set_accel(psdr_scene *s, int mode) {
    if (!s) return fail("null scene");
    s->sc.force_bvh = mode;
    s->sc.bvh_rebuild = true;      // also: rebuild the BVH topology (not only refit it) at the next configure
    s->sc.configured = false;
    return 0;
}

int psdr_scene_set_integrator(psdr_scene *s, int kind, int mis) {
    if (!s) return fail("null scene");
    if (kind != PSDR_INTEGRATOR_PATH && kind != PSDR_INTEGRATOR_DIRECT) return fail("unknown integrator kind");
    if (kind == PSDR_INTEGRATOR_DIRECT && (mis < 0 || mis > 2)) return fail("mis >= 0 && mis <= 2");
    s->sc.integrator_mis = kind == PSDR_INTEGRATOR_DIRECT ? mis : 2;
    return 0;
}

int psdr_scene_set_integrator_collocated(psdr_scene *s, float intensity, float d_intensity, int bsdf_field) {
    if (!s) return fail("null scene");
    s->sc.integrator_mis = 3;
    s->sc.colloc_intensity = Dual(intensity, d_intensity);
    s->sc.colloc_field = bsdf_field != 0;
    return 0;
}

int psdr_set_cta_policy(int policy) {
    if (policy < 0 || policy > 2) return fail("policy >= 0 && policy <= 2");
    psdr::g_cta_policy = policy;
    return 0;
}

int psdr_set_edge_sort(int bins) {
    if (bins < 0 || bins == 1 || bins > psdr::kEdgeSortMaxBins) return fail("bins == 0 || (bins >= 2 && bins <= 2048)");
    g_edge_sort_bins = bins;
    return 0;
}

int psdr_scene_set_output_multicast(psdr_scene *s, int on) {
    if (!s) return fail("null scene");
    if (on < 0 || on > 2) return fail("mode >= 0 && mode <= 2");
    s->sc.out_multicast = on;
    return 0;
}

int psdr_scene_set_reference_arithmetic(psdr_scene *s, int on) {
    if (!s) return fail("null scene");
    s->sc.ref_rcp = on != 0;
    s->sc.configured = false;
    return 0;
}

// A BSDF record joins the numbered list ("BSDF[i]"), or -- right after psdr_scene_begin_nested_bsdf -- the un-numbered list of
// BSDFs that NormalMap records wrap; the handle of a nested record is PSDR_NESTED_BSDF_BASE + its position.
static bool bsdf_id_taken(psdr_scene *s, const char *id) {
    if (s->next_bsdf_nested) return false;
    if (s->sc.find_bsdf(id) >= 0) { fail(std::string("Duplicate BSDF id: ") + id); return true; }
    return false;
}
static int push_bsdf(psdr_scene *s, const HBsdf &b) {
    Scene &sc = s->sc;
    sc.configured = false;
    if (s->next_bsdf_nested) {
        s->next_bsdf_nested = false;
        sc.nested_bsdfs.push_back(b);
        return PSDR_NESTED_BSDF_BASE + (int) sc.nested_bsdfs.size() - 1;
    }
    sc.bsdfs.push_back(b);
    return (int) sc.bsdfs.size() - 1;
}
static HBsdf *bsdf_at(Scene &sc, int index) {
    if (index >= PSDR_NESTED_BSDF_BASE) {
        const int k = index - PSDR_NESTED_BSDF_BASE;
        return k < (int) sc.nested_bsdfs.size() ? &sc.nested_bsdfs[k] : nullptr;
    }
    return (index >= 0 && index < (int) sc.bsdfs.size()) ? &sc.bsdfs[index] : nullptr;
}

int psdr_scene_begin_nested_bsdf(psdr_scene *s) {
    if (!s) return fail("null scene");
    s->next_bsdf_nested = true;
    return 0;
}

int psdr_scene_add_bsdf_diffuse(psdr_scene *s, const char *id, const float reflectance[3], int two_side) {
    if (!s || !id || !reflectance) { fail("null argument"); return -1; }
    if (bsdf_id_taken(s, id)) return -1;
    HBsdf b;
    b.id = id;
    b.type = 0;
    b.reflectance = V3d(Dual(reflectance[0]), Dual(reflectance[1]), Dual(reflectance[2]));
    b.two_side = two_side != 0;
    return push_bsdf(s, b);
}

int psdr_scene_add_bsdf_microfacet(psdr_scene *s, const char *id, const float specular[3], const float diffuse[3], float roughness,
                                   int two_side) {
    if (!s || !id || !specular || !diffuse) { fail("null argument"); return -1; }
    if (bsdf_id_taken(s, id)) return -1;
    HBsdf b;
    b.id = id;
    b.type = 1;
    b.reflectance = V3d(Dual(diffuse[0]), Dual(diffuse[1]), Dual(diffuse[2]));
    b.specular = V3d(Dual(specular[0]), Dual(specular[1]), Dual(specular[2]));
    b.roughness = Dual(roughness);
    b.two_side = two_side != 0;
    return push_bsdf(s, b);
}

int psdr_scene_add_bsdf_roughconductor(psdr_scene *s, const char *id, float alpha, const float eta[3], const float k[3], const float specular[3],
                                       int two_side) {
    if (!s || !id || !eta || !k || !specular) { fail("null argument"); return -1; }
    if (bsdf_id_taken(s, id)) return -1;
    HBsdf b;
    b.id = id;
    b.type = 2;
    b.reflectance = V3d(Dual(0.f), Dual(0.f), Dual(0.f));
    b.specular = V3d(Dual(specular[0]), Dual(specular[1]), Dual(specular[2]));
    b.roughness = Dual(alpha);
    b.eta = V3d(Dual(eta[0]), Dual(eta[1]), Dual(eta[2]));
    b.k = V3d(Dual(k[0]), Dual(k[1]), Dual(k[2]));
    b.two_side = two_side != 0;
    return push_bsdf(s, b);
}

int psdr_scene_add_bsdf_roughdielectric(psdr_scene *s, const char *id, float alpha, float int_ior, float ext_ior, int two_side) {
    if (!s || !id) { fail("null argument"); return -1; }
    if (bsdf_id_taken(s, id)) return -1;
    HBsdf b;
    b.id = id;
    b.type = 3;
    b.roughness = Dual(alpha);
    // m_eta(intIOR / extIOR), m_inv_eta(extIOR / intIOR): two separately rounded quotients (roughdielectric.h:20-25)
    b.eta = V3d(Dual(int_ior / ext_ior), Dual(ext_ior / int_ior), Dual(0.f));
    b.two_side = two_side != 0;
    return push_bsdf(s, b);
}

int psdr_scene_add_bsdf_microfacet_pervertex(psdr_scene *s, const char *id, const float *specular, const float *diffuse, const float *roughness,
                                             int n_vertices, int two_side) {
    if (!s || !id || !specular || !diffuse || !roughness || n_vertices <= 0) { fail("null argument"); return -1; }
    if (bsdf_id_taken(s, id)) return -1;
    HBsdf b;
    b.id = id;
    b.type = 4;
    b.pv.resize((size_t) 7 * n_vertices);
    for (int i = 0; i < n_vertices; ++i) {
        for (int c = 0; c < 3; ++c) { b.pv[7 * i + c] = specular[3 * i + c]; b.pv[7 * i + 3 + c] = diffuse[3 * i + c]; }
        b.pv[7 * i + 6] = roughness[i];
    }
    b.two_side = two_side != 0;
    return push_bsdf(s, b);
}

int psdr_scene_add_bsdf_normalmap(psdr_scene *s, const char *id, const float normal[3], int nested, int two_side) {
    if (!s || !id || !normal) { fail("null argument"); return -1; }
    if (s->next_bsdf_nested) { s->next_bsdf_nested = false; fail("a NormalMap cannot be nested"); return -1; }
    Scene &sc = s->sc;
    if (nested < PSDR_NESTED_BSDF_BASE || nested - PSDR_NESTED_BSDF_BASE >= (int) sc.nested_bsdfs.size()) { fail("NormalMap: invalid nested BSDF handle"); return -1; }
    if (bsdf_id_taken(s, id)) return -1;
    HBsdf b;
    b.id = id;
    b.type = 5;
    b.reflectance = V3d(Dual(normal[0]), Dual(normal[1]), Dual(normal[2]));
    b.nested = nested - PSDR_NESTED_BSDF_BASE;
    b.two_side = two_side != 0;
    return push_bsdf(s, b);
}

int psdr_scene_set_bsdf_texture_slot(psdr_scene *s, int index, int slot, int w, int h) {
    if (!s) return fail("null scene");
    Scene &sc = s->sc;
    HBsdf *bp = bsdf_at(sc, index);
    if (!bp) return fail("invalid BSDF index");
    if (slot < 0 || slot > 2) return fail("invalid texture slot");
    const int type = bp->type;
    if (slot > 0 && type == 0) return fail("specular / roughness textures need a MicrofacetBSDF or a RoughConductorBSDF");
    if (slot == 0 && type == 2) return fail("a RoughConductorBSDF has no diffuse reflectance");
    if (type == 3 && slot != 2) return fail("a RoughDielectricBSDF has an alpha bitmap only");
    if (type == 4) return fail("a MicrofacetBSDFPerVertex has no bitmaps");
    if (type == 5 && slot != 0) return fail("a NormalMapBSDF has one bitmap: the normal map");
    if (w < 1 || h < 1 || (w * h > 1 && (w < 2 || h < 2))) return fail("Bitmap: invalid resolution!");
    HBsdf::Tex &t = bp->tex[slot];
    if (w * h == 1) { t.w = t.h = 0; t.data.clear(); t.ddata.clear(); }
    else if (t.w != w || t.h != h) {
        t.w = w; t.h = h;
        t.data.assign((size_t) HBsdf::tex_channels(slot) * w * h, 0.5f);
        t.ddata.clear();
    }
    sc.configured = false;
    return 0;
}
int psdr_scene_set_bsdf_texture(psdr_scene *s, int index, int w, int h) { return psdr_scene_set_bsdf_texture_slot(s, index, PSDR_TEX_REFLECTANCE, w, h); }

int psdr_scene_add_mesh(psdr_scene *s, const float *v, int nv, const int *f, int nf, const float *uv, int nuv, const int *fuv,
                        const float *to_world, const char *bsdf_id, const float *radiance, int use_face_normals, int enable_edges) {
    if (!s || !v || !f || !bsdf_id || nv <= 0 || nf <= 0) { fail("invalid mesh arguments"); return -1; }
    Scene &sc = s->sc;
    const int bi = sc.find_bsdf(bsdf_id);
    if (bi < 0) { fail(std::string("Unknown BSDF id: ") + bsdf_id); return -1; }
    for (int i = 0; i < 3 * nf; ++i)
        if (f[i] < 0 || f[i] >= nv) { fail("face index out of range"); return -1; }
    HMesh m;
    m.v_raw.resize(nv);
    for (int i = 0; i < nv; ++i) m.v_raw[i] = V3d(Dual(v[3 * i]), Dual(v[3 * i + 1]), Dual(v[3 * i + 2]));
    m.f.assign(f, f + 3 * nf);
    m.has_uv = uv != nullptr && nuv > 0 && fuv != nullptr;
    if (m.has_uv) {
        m.uv.resize(nuv);
        for (int i = 0; i < nuv; ++i) m.uv[i] = V2f(uv[2 * i], uv[2 * i + 1]);
        m.fuv.assign(fuv, fuv + 3 * nf);
        for (int i = 0; i < 3 * nf; ++i)
            if (fuv[i] < 0 || fuv[i] >= nuv) { fail("uv index out of range"); return -1; }
    }
    m.to_world[0] = M4<Dual>::identity();
    m.to_world[1] = mat_from(to_world);
    m.to_world[2] = M4<Dual>::identity();
    m.bsdf = bi;
    m.use_face_normals = use_face_normals != 0;
    m.enable_edges = enable_edges != 0;
    if (radiance) {
        HEmitter e;
        e.radiance = V3d(Dual(radiance[0]), Dual(radiance[1]), Dual(radiance[2]));
        e.mesh = (int) sc.meshes.size();
        m.emitter = (int) sc.emitters.size();
        sc.emitters.push_back(e);
    }
    sc.meshes.push_back(std::move(m));
    sc.configured = false;
    return (int) sc.meshes.size() - 1;
}

int psdr_scene_add_envmap(psdr_scene *s, const float *radiance, int w, int h, const float *to_world, float scale) {
    if (!s || !radiance) { fail("null argument"); return -1; }
    Scene &sc = s->sc;
    if (sc.env.present) { fail("A scene is only allowed to have one envmap!"); return -1; }
    if (w < 2 || h < 2) { fail("Bitmap: invalid resolution!"); return -1; }
    HEnvmap &e = sc.env;
    e.present = true;
    e.w = w; e.h = h;
    e.data.assign(radiance, radiance + (size_t) 3 * w * h);
    e.ddata.clear();
    e.scale = Dual(scale);
    e.to_world[0] = M4<Dual>::identity();
    e.to_world[1] = mat_from(to_world);
    HEmitter em;
    em.type = 1;
    em.mesh = -1;
    e.emitter = (int) sc.emitters.size();
    sc.emitters.push_back(em);
    sc.configured = false;
    return e.emitter;
}

int psdr_scene_add_perspective(psdr_scene *s, float fov_x, float near_clip, float far_clip, const float *to_world) {
    if (!s) { fail("null scene"); return -1; }
    HCamera c;
    c.fov = fov_x;
    c.near_ = near_clip;
    c.far_ = far_clip;
    c.to_world[0] = M4<Dual>::identity();
    c.to_world[1] = mat_from(to_world);
    c.to_world[2] = M4<Dual>::identity();
    s->sc.cameras.push_back(c);
    s->sc.configured = false;
    return (int) s->sc.cameras.size() - 1;
}

int psdr_scene_add_perspective_intrinsic(psdr_scene *s, float fx, float fy, float cx, float cy, float near_clip, float far_clip, const float *to_world) {
    const int i = psdr_scene_add_perspective(s, 0.f, near_clip, far_clip, to_world);
    if (i < 0) return i;
    HCamera &c = s->sc.cameras[i];
    c.use_intrinsic = true;
    c.fx = fx; c.fy = fy; c.cx = cx; c.cy = cy;
    return i;
}

int psdr_scene_add_orthographic(psdr_scene *s, float near_clip, float far_clip, const float *to_world) {
    const int i = psdr_scene_add_perspective(s, 0.f, near_clip, far_clip, to_world);
    if (i < 0) return i;
    s->sc.cameras[i].ortho = true;
    return i;
}

static int set_param_impl(psdr_scene *s, int kind, int index, const float *data, int n, bool tangent) {
    if (!s || !data) return fail("null argument");
    Scene &sc = s->sc;
    auto put = [&](Dual &x, float val_) { if (tangent) x.d = val_; else x.v = val_; };
    // texel data (or its tangent) of a textured slot; returns -1 if the slot holds a constant
    auto put_texels = [&](HBsdf::Tex &t, int channels) -> int {
        if (t.w <= 0) return -1;
        if (n != channels * t.w * t.h) return fail("texture size mismatch");
        if (tangent) {
            bool any = false;
            for (int i = 0; i < n; ++i) any |= data[i] != 0.f;
            if (any) t.ddata.assign(data, data + n); else t.ddata.clear();
        } else t.data.assign(data, data + n);
        return 0;
    };
    switch (kind) {
        case PSDR_MESH_VERTICES: {
            if (index < 0 || index >= (int) sc.meshes.size()) return fail("invalid mesh index");
            HMesh &m = sc.meshes[index];
            if (n != 3 * (int) m.v_raw.size()) return fail("vertex buffer size mismatch");
            for (size_t i = 0; i < m.v_raw.size(); ++i) { put(m.v_raw[i].x, data[3 * i]); put(m.v_raw[i].y, data[3 * i + 1]); put(m.v_raw[i].z, data[3 * i + 2]); }
            break;
        }
        case PSDR_MESH_TO_WORLD_LEFT: case PSDR_MESH_TO_WORLD_RAW: case PSDR_MESH_TO_WORLD_RIGHT: {
            if (index < 0 || index >= (int) sc.meshes.size()) return fail("invalid mesh index");
            if (n != 16) return fail("a transform is 16 floats");
            M4<Dual> &M = sc.meshes[index].to_world[kind - PSDR_MESH_TO_WORLD_LEFT];
            for (int i = 0; i < 16; ++i) put(M.m[i / 4][i % 4], data[i]);
            break;
        }
        case PSDR_SENSOR_TO_WORLD_LEFT: case PSDR_SENSOR_TO_WORLD_RAW: case PSDR_SENSOR_TO_WORLD_RIGHT: {
            if (index < 0 || index >= (int) sc.cameras.size()) return fail("Invalid sensor id!");
            if (n != 16) return fail("a transform is 16 floats");
            M4<Dual> &M = sc.cameras[index].to_world[kind - PSDR_SENSOR_TO_WORLD_LEFT];
            for (int i = 0; i < 16; ++i) put(M.m[i / 4][i % 4], data[i]);
            break;
        }
        case PSDR_BSDF_REFLECTANCE: {
            HBsdf *bp = bsdf_at(sc, index);
            if (!bp) return fail("invalid BSDF index");
            const int rc = put_texels(bp->tex[0], 3);
            if (rc > 0) return rc;
            if (rc == 0) break;
            if (n != 3) return fail("reflectance is 3 floats");
            put(bp->reflectance.x, data[0]); put(bp->reflectance.y, data[1]); put(bp->reflectance.z, data[2]);
            break;
        }
        case PSDR_BSDF_SPECULAR: {
            HBsdf *bp = bsdf_at(sc, index);
            if (!bp) return fail("invalid BSDF index");
            const int rc = put_texels(bp->tex[1], 3);
            if (rc > 0) return rc;
            if (rc == 0) break;
            if (n != 3) return fail("specular reflectance is 3 floats");
            put(bp->specular.x, data[0]); put(bp->specular.y, data[1]); put(bp->specular.z, data[2]);
            break;
        }
        case PSDR_BSDF_ROUGHNESS: {
            HBsdf *bp = bsdf_at(sc, index);
            if (!bp) return fail("invalid BSDF index");
            const int rc = put_texels(bp->tex[2], 1);
            if (rc > 0) return rc;
            if (rc == 0) break;
            if (n != 1) return fail("roughness is 1 float");
            put(bp->roughness, data[0]);
            break;
        }
        case PSDR_BSDF_ETA: case PSDR_BSDF_K: {
            HBsdf *bp = bsdf_at(sc, index);
            if (!bp) return fail("invalid BSDF index");
            if (bp->type != 2) return fail("eta / k belong to a RoughConductorBSDF");
            if (n != 3) return fail("eta / k are 3 floats");
            V3d &v = kind == PSDR_BSDF_ETA ? bp->eta : bp->k;
            put(v.x, data[0]); put(v.y, data[1]); put(v.z, data[2]);
            break;
        }
        case PSDR_BSDF_PERVERTEX: {
            HBsdf *bp = bsdf_at(sc, index);
            if (!bp) return fail("invalid BSDF index");
            if (bp->type != 4) return fail("per-vertex tables belong to a MicrofacetBSDFPerVertex");
            if (n != (int) bp->pv.size()) return fail("per-vertex table size mismatch (7 floats per vertex: specular rgb, diffuse rgb, roughness)");
            if (tangent) {
                bool any = false;
                for (int i = 0; i < n; ++i) any |= data[i] != 0.f;
                if (any) bp->d_pv.assign(data, data + n); else bp->d_pv.clear();
            } else bp->pv.assign(data, data + n);
            break;
        }
        case PSDR_BSDF_REFLECTANCE_UV: case PSDR_BSDF_SPECULAR_UV: case PSDR_BSDF_ROUGHNESS_UV: {
            HBsdf *bp = bsdf_at(sc, index);
            if (!bp) return fail("invalid BSDF index");
            if (n != 4) return fail("a uv transform is 4 floats (scale, rotate, translate.x, translate.y)");
            HBsdf::Tex &t = bp->tex[kind - PSDR_BSDF_REFLECTANCE_UV];
            put(t.scale, data[0]); put(t.rot, data[1]); put(t.tx, data[2]); put(t.ty, data[3]);
            break;
        }
        case PSDR_ENVMAP_RADIANCE: {
            if (!sc.env.present) return fail("the scene has no environment map");
            if (n != 3 * sc.env.w * sc.env.h) return fail("envmap radiance size mismatch");
            if (tangent) {
                bool any = false;
                for (int i = 0; i < n; ++i) any |= data[i] != 0.f;
                if (any || !sc.env.ddata.empty()) sc.env.ddata_version++;
                if (any) sc.env.ddata.assign(data, data + n); else sc.env.ddata.clear();
            } else if (sc.env.data.size() != (size_t) n || std::memcmp(sc.env.data.data(), data, sizeof(float) * n) != 0) {
                sc.env.data.assign(data, data + n);
                sc.env.data_version++;       // the cell table and the device copy follow (scene.cpp configure_envmap, device_upload.cu)
            }
            break;
        }
        case PSDR_ENVMAP_SCALE: {
            if (!sc.env.present) return fail("the scene has no environment map");
            if (n != 1) return fail("scale is 1 float");
            put(sc.env.scale, data[0]);
            break;
        }
        case PSDR_ENVMAP_TO_WORLD_LEFT: {
            if (!sc.env.present) return fail("the scene has no environment map");
            if (n != 16) return fail("a transform is 16 floats");
            for (int i = 0; i < 16; ++i) put(sc.env.to_world[0].m[i / 4][i % 4], data[i]);
            break;
        }
        case PSDR_EMITTER_RADIANCE: {
            if (index < 0 || index >= (int) sc.emitters.size()) return fail("invalid emitter index");
            if (sc.emitters[index].type != 0) return fail("not an AreaLight");
            if (n != 3) return fail("radiance is 3 floats");
            put(sc.emitters[index].radiance.x, data[0]); put(sc.emitters[index].radiance.y, data[1]); put(sc.emitters[index].radiance.z, data[2]);
            break;
        }
        default: return fail("unknown parameter kind");
    }
    sc.configured = false;   // as in the reference, configure() must follow a parameter change
    return 0;
}

int psdr_scene_set_param(psdr_scene *s, int kind, int index, const float *value, int n) { return set_param_impl(s, kind, index, value, n, false); }
int psdr_scene_set_tangent(psdr_scene *s, int kind, int index, const float *tangent, int n) { return set_param_impl(s, kind, index, tangent, n, true); }

int psdr_scene_clear_tangents(psdr_scene *s) {
    if (!s) return fail("null scene");
    Scene &sc = s->sc;
    for (HMesh &m : sc.meshes) {
        for (V3d &p : m.v_raw) p = detach(p);
        for (auto &M : m.to_world)
            for (int i = 0; i < 16; ++i) M.m[i / 4][i % 4].d = 0.f;
    }
    for (HCamera &c : sc.cameras)
        for (auto &M : c.to_world)
            for (int i = 0; i < 16; ++i) M.m[i / 4][i % 4].d = 0.f;
    for (std::vector<HBsdf> *list : {&sc.bsdfs, &sc.nested_bsdfs})
        for (HBsdf &b : *list) {
            for (HBsdf::Tex &t : b.tex) { t.ddata.clear(); t.scale = detach(t.scale); t.rot = detach(t.rot); t.tx = detach(t.tx); t.ty = detach(t.ty); }
            b.reflectance = detach(b.reflectance); b.specular = detach(b.specular); b.roughness = detach(b.roughness);
            b.eta = detach(b.eta); b.k = detach(b.k);
            b.d_pv.clear();
        }
    for (HEmitter &e : sc.emitters) e.radiance = detach(e.radiance);
    if (!sc.env.ddata.empty()) sc.env.ddata_version++;
    sc.env.ddata.clear();
    sc.env.scale = detach(sc.env.scale);
    for (int i = 0; i < 16; ++i) sc.env.to_world[0].m[i / 4][i % 4].d = 0.f;
    sc.configured = false;
    return 0;
}

int psdr_scene_configure(psdr_scene *s, const int *active_sensors, int n_active) {
    if (!s) return fail("null scene");
    PSDR_TRY
    s->sc.configure(active_sensors, n_active);
    return 0;
    PSDR_CATCH
}

double psdr_scene_last_configure_ms(psdr_scene *s) { return s ? s->sc.last_configure_ms : 0.0; }

int psdr_scene_enable_timing(psdr_scene *s, int on) {
    if (!s) return fail("null scene");
    PSDR_TRY
    s->timing = on != 0;
    if (s->timing && !s->ev[0][0]) {
        cuda_ok(cudaSetDevice(s->sc.device), "cudaSetDevice");
        for (auto &p : s->ev)
            for (auto &e : p) cuda_ok(cudaEventCreate(&e), "cudaEventCreate");
    }
    return 0;
    PSDR_CATCH
}

double psdr_scene_kernel_ms(psdr_scene *s, int term) {
    if (!s || !s->timing) return -1.0;
    const int k = term == PSDR_TERM_INTERIOR ? 0 : term == PSDR_TERM_PRIMARY_EDGES ? 1 : term == PSDR_TERM_SECONDARY_EDGES ? 2 : -1;
    if (k < 0 || !s->ev_used[k]) return -1.0;
    float ms = 0.f;
    if (cudaEventSynchronize(s->ev[k][1]) != cudaSuccess || cudaEventElapsedTime(&ms, s->ev[k][0], s->ev[k][1]) != cudaSuccess) return -1.0;
    return (double) ms;
}

int psdr_scene_query(psdr_scene *s, int what, int index) {
    if (!s) { fail("null scene"); return -1; }
    const Scene &sc = s->sc;
    switch (what) {
        case PSDR_Q_NUM_MESHES: return (int) sc.meshes.size();
        case PSDR_Q_NUM_SENSORS: return (int) sc.cameras.size();
        case PSDR_Q_NUM_EMITTERS: return (int) sc.emitters.size();
        case PSDR_Q_NUM_TRIANGLES: { int n = 0; for (auto &m : sc.meshes) n += (int) m.f.size() / 3; return n; }
        case PSDR_Q_NUM_PRIMARY_EDGES: return (index >= 0 && index < (int) sc.cameras.size()) ? (int) sc.cameras[index].edges.size() : -1;
        case PSDR_Q_NUM_SECONDARY_EDGES: return (int) sc.sec_edges.size();
        case PSDR_Q_NUM_MESH_EDGES: return (index >= 0 && index < (int) sc.meshes.size()) ? (int) sc.meshes[index].edges.size() : -1;
        case PSDR_Q_NUM_MESH_VERTICES: return (index >= 0 && index < (int) sc.meshes.size()) ? (int) sc.meshes[index].v_raw.size() : -1;
        case PSDR_Q_NUM_MESH_FACES: return (index >= 0 && index < (int) sc.meshes.size()) ? (int) sc.meshes[index].f.size() / 3 : -1;
        case PSDR_Q_IS_CONFIGURED: return sc.configured ? 1 : 0;
        case PSDR_Q_USES_BVH: return sc.dscene.use_bvh;
        case PSDR_Q_UPLOAD_BYTES: return (int) sc.upload_bytes;
        case PSDR_Q_BVH_BUILDS: return sc.bvh_builds;
        case PSDR_Q_BVH_REFITS: return sc.bvh_refits;
        case PSDR_Q_KERNEL_FAMILY: return sc.configured ? scene_cfg(sc.dscene) : -1;
        case PSDR_Q_GRAD_TABLE_MULTICAST: {
            if (index < 0 || index >= (int) sc.cameras.size()) { fail("sensor index out of range"); return -1; }
            const GradLayout gl = sc.grad_layout(index);
            return grad_table_multicast_ok(gl) ? 1 : 0;
        }
        case PSDR_Q_GUIDING_CELLS:
            return (index >= 0 && index < (int) sc.cameras.size() && sc.cameras[index].guide_ready) ? (int) sc.cameras[index].guide.pmf.size() : 0;
        default: fail("unknown query"); return -1;
    }
}

int psdr_scene_mesh_edges(psdr_scene *s, int mesh, int *out) {
    if (!s || !out) return fail("null argument");
    if (mesh < 0 || mesh >= (int) s->sc.meshes.size()) return fail("invalid mesh index");
    const auto &E = s->sc.meshes[mesh].edges;
    const size_t n = E.size();
    for (size_t i = 0; i < n; ++i) { out[i] = E[i].v0; out[n + i] = E[i].v1; out[2 * n + i] = E[i].f0; out[3 * n + i] = E[i].f1; }
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// render
// ---------------------------------------------------------------------------------------------
namespace {

// Block-cyclic deal of 32-lane blocks (dscene.h RenderParams): fills the sharding fields for a term of n lanes
void set_shard(RenderParams &rp, long long n, int rank, int world) {
    const long long blocks = (n + 31) / 32;
    const long long owned = blocks > rank ? (blocks - rank + world - 1) / world : 0;
    rp.lane_begin = 0;
    rp.lane_end = world == 1 ? n : owned * 32;
    rp.n_lanes = n;
    rp.shard_rank = rank;
    rp.shard_world = world;
}

// Integrator::renderC / renderD front matter (integrator.cpp:12-31, 51-73): argument checks and
// sampler (re)seeding; returns the per-sampler (seed, skip) pair of this call.
void begin_render(Scene &sc, int sensor, long long seed, const int *pix_id, bool ad, int terms, int max_depth, RenderParams rp[3]) {
    const int mis = sc.integrator_mis;
    // draws per bounce (device_path.cuh li_step); the CollocatedIntegrator's Li draws nothing
    const unsigned long long per_bounce = mis == 3 ? 0ull : mis == 0 ? 2ull : mis == 1 ? 3ull : 5ull;
    for (int k = 0; k < 3; ++k) rp[k].mis = mis;
    sc.dscene.colloc_intensity = sc.colloc_intensity.v;
    sc.dscene.d_colloc_intensity = sc.colloc_intensity.d;
    sc.dscene.colloc_field = sc.colloc_field ? 1 : 0;
    if (pix_id && seed == -1) throw std::runtime_error("While using batch rendering, seed must be set!");
    if (!sc.configured) throw std::runtime_error("Input scene must be configured!");
    if (sensor < 0 || sensor >= (int) sc.cameras.size()) throw std::runtime_error("Invalid sensor id!");
    if (max_depth < 0) throw std::runtime_error("max_depth >= 0");
    const int per[3] = {sc.spp, ad ? sc.sppe : 0, ad ? sc.sppse : 0};
    const unsigned long long draws[3] = {2ull + per_bounce * max_depth, 1ull + 2ull * per_bounce * max_depth, 3ull};
    for (int k = 0; k < 3; ++k) {
        if (per[k] <= 0) continue;
        SamplerState &st = sc.samplers[k];
        if (seed != -1) { st.seed = seed; st.consumed = 0; st.ready = true; }
        if (!st.ready) throw std::runtime_error("Sampler::seed() must be invoked before using this sampler!");
        rp[k].seed = st.seed;
        rp[k].skip = st.consumed;
        if (terms & (1 << k)) st.consumed += draws[k];
    }
}

// PSDR_EDGE_DYNAMIC=0: static slices in the large-CTA primary-edge kernels (A/B switch; default: dynamic chunk hand-out)
bool edge_dynamic() {
    static const bool on = [] { const char *e = getenv("PSDR_EDGE_DYNAMIC"); return !(e && e[0] == '0'); }();
    return on;
}

// PSDR_SEC_EDGE_SORT=0: the secondary-edge kernels keep the lane order whatever psdr_set_edge_sort says (A/B switch)
bool sec_edge_sort() {
    static const bool on = [] { const char *e = getenv("PSDR_SEC_EDGE_SORT"); return !(e && e[0] == '0'); }();
    return on;
}

int *sched_counters(psdr_scene *s, int which) {
    if (!s->d_sched) {
        cuda_ok(cudaMalloc(&s->d_sched, sizeof(int) * 8), "cudaMalloc(sched)");
        cuda_ok(cudaMemset(s->d_sched, 0, sizeof(int) * 8), "cudaMemset(sched)");
    }
    return s->d_sched + 2 * which;
}

// Primary-edge launches of at least kEdgeSortMinLanes local lanes walk their lanes in the order of edge_sort.cu (three small
// kernels on the term's stream, counted as launches); smaller ones keep the lane order.
constexpr long long kEdgeSortMinLanes = 32768;
void order_edge_lanes(psdr_scene *s, int which, RenderParams &rp, const DCamera &cam, cudaStream_t q) {
    const long long span = rp.lane_end - rp.lane_begin;
    if (g_edge_sort_bins < 2 || span < kEdgeSortMinLanes || span > 2147483647LL) return;
    auto &e = s->edge_sort[which];
    if ((size_t) span > e.cap) {
        if (e.key) cudaFree(e.key);
        if (e.perm) cudaFree(e.perm);
        e.key = nullptr; e.perm = nullptr; e.cap = 0;
        cuda_ok(cudaMalloc(&e.key, sizeof(unsigned short) * span), "cudaMalloc(edge sort keys)");
        cuda_ok(cudaMalloc(&e.perm, sizeof(int) * span), "cudaMalloc(edge sort order)");
        e.cap = (size_t) span;
    }
    if (!e.work) {
        cuda_ok(cudaMalloc(&e.work, sizeof(int) * 2 * kEdgeSortMaxBins), "cudaMalloc(edge sort counters)");
        cuda_ok(cudaMemsetAsync(e.work, 0, sizeof(int) * 2 * kEdgeSortMaxBins, q), "memset(edge sort counters)");
    }
    cuda_ok(launch_edge_sort(rp, cam, which >= 2, g_edge_sort_bins, e.key, e.work, e.perm, q), "edge sort kernels");
    g_launches += 3;
    rp.perm = e.perm;
}

void tick(psdr_scene *s, int k, int which, cudaStream_t st) {
    if (!s->timing) return;
    cuda_ok(cudaEventRecord(s->ev[k][which], st), "cudaEventRecord");
    if (which) s->ev_used[k] = true;
}

// The term kernels of one call are independent (they only ADD into the outputs), each is a persistent grid that fills
// the GPU, and each ends with a tail in which the SMs whose CTA finished early sit idle.  Launched on three streams
// the next kernel's CTAs move into those SMs as they free up, so only the LAST kernel's tail is exposed: the caller's
// stream forks after the output memsets and joins behind the last kernel.  Order: interior, secondary edges, primary
// edges -- the primary-edge kernel's lanes are statistically identical, its tail is the shortest.  With per-kernel
// timing enabled (psdr_scene_enable_timing) everything stays on the caller's stream, back to back.
struct TermStreams {
    psdr_scene *s;
    cudaStream_t main;
    bool overlap;
    int used = 0;
    TermStreams(psdr_scene *s_, cudaStream_t st, int n_kernels) : s(s_), main(st), overlap(!s_->timing && n_kernels > 1) {
        if (!overlap) return;
        for (int i = 0; i < 2; ++i) {
            if (!s->side[i]) cuda_ok(cudaStreamCreateWithFlags(&s->side[i], cudaStreamNonBlocking), "cudaStreamCreate");
            if (!s->ev_join[i]) cuda_ok(cudaEventCreateWithFlags(&s->ev_join[i], cudaEventDisableTiming), "cudaEventCreate");
        }
        if (!s->ev_fork) cuda_ok(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming), "cudaEventCreate");
        cuda_ok(cudaEventRecord(s->ev_fork, main), "cudaEventRecord(fork)");
    }
    // stream of the next kernel: the first runs on the caller's stream, the others on the side streams
    cudaStream_t next() {
        if (!overlap || used == 0) { ++used; return main; }
        cudaStream_t q = s->side[used - 1];
        ++used;
        cuda_ok(cudaStreamWaitEvent(q, s->ev_fork, 0), "cudaStreamWaitEvent(fork)");
        return q;
    }
    void join() {
        if (!overlap) return;
        for (int i = 0; i + 1 < used; ++i) {
            cuda_ok(cudaEventRecord(s->ev_join[i], s->side[i]), "cudaEventRecord(join)");
            cuda_ok(cudaStreamWaitEvent(main, s->ev_join[i], 0), "cudaStreamWaitEvent(join)");
        }
    }
};

int render_impl(psdr_scene *s, int sensor, int max_depth, long long seed, int hide_emitters, bool ad, int terms, int reference_scaling,
                const int *pix_id, int npix_sel, float *img, float *dimg, cudaStream_t st) {
    Scene &sc = s->sc;
    cuda_ok(cudaSetDevice(sc.device), "cudaSetDevice");
    RenderParams rp[3];
    for (auto &r : rp) {
        r = RenderParams{};
        r.max_depth = max_depth;
        r.hide_emitters = hide_emitters;
        r.pix_id = nullptr;
        r.tangent_scale = 1.f;
    }
    begin_render(sc, sensor, seed, pix_id, ad, terms, max_depth, rp);
    s->ev_used[0] = s->ev_used[1] = s->ev_used[2] = false;
    const long long npix_full = (long long) sc.width * sc.height;
    const long long npix = pix_id ? npix_sel : npix_full;
    if (pix_id && npix_sel <= 0) throw std::runtime_error("empty pixel batch");
    if (npix * (long long) std::max(sc.spp, 1) > 2147483647LL) throw std::runtime_error("num_samples <= std::numeric_limits<int>::max()");
    if (!img) throw std::runtime_error("null image buffer");
    const bool primal_only = ad && !dimg;   // renderD's image without the forward-mode derivative image
    const DCamera &cam = sc.dcameras[sensor];
    // multicast outputs: the caller zeroed every rank's replica (and synchronised the ranks) before this call
    for (auto &r : rp) r.out_multicast = sc.out_multicast;
    if (!sc.out_multicast) {
        cuda_ok(cudaMemsetAsync(img, 0, sizeof(float) * 3 * npix, st), "memset(img)");
        if (ad && dimg) cuda_ok(cudaMemsetAsync(dimg, 0, sizeof(float) * 3 * npix, st), "memset(dimg)");
    }
    const bool do_int = sc.spp > 0 && (terms & PSDR_TERM_INTERIOR);
    const bool do_pri = ad && !primal_only && sc.sppe > 0 && (terms & PSDR_TERM_PRIMARY_EDGES) && cam.n_edges > 0;
    // (the CollocatedIntegrator has no secondary-edge term: Integrator::render_secondary_edges is empty, integrator.h:22)
    const bool do_sec = ad && !primal_only && sc.sppse > 0 && (terms & PSDR_TERM_SECONDARY_EDGES) && sc.dscene.n_sec_edges > 0 && sc.integrator_mis != 3;
    if ((do_pri || do_sec) && pix_id) throw std::runtime_error("batch rendering supports the interior term only");
    TermStreams ts(s, st, (int) do_int + (int) do_pri + (int) do_sec);
    if (do_int) {
        set_shard(rp[0], npix * sc.spp, sc.rank, sc.world);
        rp[0].pix_id = pix_id; rp[0].npix = (int) npix;
        rp[0].tangent_scale = reference_scaling ? 2.f : 1.f;
        rp[0].sched = sched_counters(s, 0);
        cudaStream_t q = ts.next();
        tick(s, 0, 0, q);
        cuda_ok(launch_interior(sc.dscene, cam, rp[0], ad, img, dimg, q), "interior kernel");
        tick(s, 0, 1, q);
        g_launches++;
        // the primal image is final here (the edge kernels only add to the derivative image): its D2H overlaps them
        if (s->early_img_host) cuda_ok(cudaMemcpyAsync(s->early_img_host, img, sizeof(float) * 3 * npix, cudaMemcpyDeviceToHost, q), "D2H(img)");
    }
    if (do_sec) {
        set_shard(rp[2], npix_full * sc.sppse, sc.rank, sc.world);
        rp[2].tangent_scale = reference_scaling ? 2.f : 1.f;
        cudaStream_t q = ts.next();
        tick(s, 2, 0, q);
        if (sec_edge_sort()) order_edge_lanes(s, 2, rp[2], cam, q);
        cuda_ok(launch_secondary_edges(sc.dscene, cam, rp[2], dimg, q), "secondary-edge kernel");
        tick(s, 2, 1, q);
        g_launches++;
    }
    if (do_pri) {
        set_shard(rp[1], npix_full * sc.sppe, sc.rank, sc.world);
        cudaStream_t q = ts.next();
        tick(s, 1, 0, q);
        order_edge_lanes(s, 0, rp[1], cam, q);
        if (edge_dynamic()) rp[1].sched = sched_counters(s, 2);
        cuda_ok(launch_primary_edges(sc.dscene, cam, rp[1], dimg, q), "primary-edge kernel");
        tick(s, 1, 1, q);
        g_launches++;
    }
    ts.join();
    return 0;
}

void ensure_staging(psdr_scene *s, size_t npix, size_t npix_ids) {
    cuda_ok(cudaSetDevice(s->sc.device), "cudaSetDevice");
    if (!s->stream) cuda_ok(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking), "cudaStreamCreate");
    if (npix > s->img_cap) {
        if (s->d_img) cudaFree(s->d_img);
        if (s->d_dimg) cudaFree(s->d_dimg);
        cuda_ok(cudaMalloc(&s->d_img, sizeof(float) * 3 * npix), "cudaMalloc(img)");
        cuda_ok(cudaMalloc(&s->d_dimg, sizeof(float) * 3 * npix), "cudaMalloc(dimg)");
        s->img_cap = npix;
    }
    if (npix_ids > s->pix_cap) {
        if (s->d_pix) cudaFree(s->d_pix);
        cuda_ok(cudaMalloc(&s->d_pix, sizeof(int) * npix_ids), "cudaMalloc(pix)");
        s->pix_cap = npix_ids;
    }
}

}  // namespace

extern "C" {

int psdr_render_c(psdr_scene *s, int sensor, int max_depth, long long seed, int hide_emitters, const int *pix_id, int npix, float *img,
                  void *cuda_stream) {
    if (!s) return fail("null scene");
    PSDR_TRY
    return render_impl(s, sensor, max_depth, seed, hide_emitters, false, PSDR_TERM_INTERIOR, 0, pix_id, npix, img, nullptr, (cudaStream_t) cuda_stream);
    PSDR_CATCH
}

int psdr_render_d(psdr_scene *s, int sensor, int max_depth, long long seed, int hide_emitters, int terms, int reference_scaling,
                  const int *pix_id, int npix, float *img, float *dimg, void *cuda_stream) {
    if (!s) return fail("null scene");
    PSDR_TRY
    return render_impl(s, sensor, max_depth, seed, hide_emitters, true, terms, reference_scaling, pix_id, npix, img, dimg, (cudaStream_t) cuda_stream);
    PSDR_CATCH
}

// adjoint kernels of all requested terms into a device gradient table (zeroed here); asynchronous on `st`
static void vjp_launch(psdr_scene *s, int sensor, int max_depth, long long seed, int hide_emitters, int terms, int reference_scaling,
                       const int *pix_id, int npix_sel, const float *d_img, float *table, cudaStream_t st) {
    Scene &sc = s->sc;
    cuda_ok(cudaSetDevice(sc.device), "cudaSetDevice");
    if (!d_img) throw std::runtime_error("null cotangent image");
    if (max_depth > 8) throw std::runtime_error("the adjoint supports max_depth <= 8");
    if (sc.integrator_mis == 3 && sc.colloc_field) throw std::runtime_error("FieldExtractionIntegrator has no reverse mode: use the forward-mode derivative image");
    RenderParams rp[3];
    for (auto &r : rp) {
        r = RenderParams{};
        r.max_depth = max_depth;
        r.hide_emitters = hide_emitters;
        r.tangent_scale = 1.f;
    }
    SamplerState saved[3] = {sc.samplers[0], sc.samplers[1], sc.samplers[2]};
    begin_render(sc, sensor, seed, pix_id, true, terms, max_depth, rp);
    for (int k = 0; k < 3; ++k) sc.samplers[k] = saved[k];    // a replay does not consume the streams
    const long long npix_full = (long long) sc.width * sc.height;
    const long long npix = pix_id ? npix_sel : npix_full;
    if (pix_id && npix_sel <= 0) throw std::runtime_error("empty pixel batch");
    const DCamera &cam = sc.dcameras[sensor];
    GradLayout gl = sc.grad_layout(sensor);
    gl.base = table;
    for (auto &r : rp) r.out_multicast = sc.out_multicast ? 1 : 0;
    if (sc.out_multicast && !grad_table_multicast_ok(gl)) throw std::runtime_error("gradient table too large for a multicast target (PSDR_Q_GRAD_TABLE_MULTICAST)");
    if (!sc.out_multicast) cuda_ok(cudaMemsetAsync(table, 0, sizeof(float) * gl.total, st), "memset(grad table)");
    s->ev_used[0] = s->ev_used[1] = s->ev_used[2] = false;
    const bool do_int = sc.spp > 0 && (terms & PSDR_TERM_INTERIOR);
    const bool do_pri = sc.sppe > 0 && (terms & PSDR_TERM_PRIMARY_EDGES) && cam.n_edges > 0;
    const bool do_sec = sc.sppse > 0 && (terms & PSDR_TERM_SECONDARY_EDGES) && sc.dscene.n_sec_edges > 0 && sc.integrator_mis != 3;
    if ((do_pri || do_sec) && pix_id) throw std::runtime_error("batch rendering supports the interior term only");
    TermStreams ts(s, st, (int) do_int + (int) do_pri + (int) do_sec);      // see render_impl
    if (do_int) {
        set_shard(rp[0], npix * sc.spp, sc.rank, sc.world);
        rp[0].pix_id = pix_id; rp[0].npix = (int) npix;
        rp[0].tangent_scale = reference_scaling ? 2.f : 1.f;
        rp[0].sched = sched_counters(s, 1);
        cudaStream_t q = ts.next();
        tick(s, 0, 0, q);
        cuda_ok(launch_interior_vjp(sc.dscene, cam, rp[0], gl, d_img, q), "interior adjoint kernel");
        tick(s, 0, 1, q);
        g_launches++;
    }
    if (do_sec) {
        set_shard(rp[2], npix_full * sc.sppse, sc.rank, sc.world);
        rp[2].tangent_scale = reference_scaling ? 2.f : 1.f;
        cudaStream_t q = ts.next();
        tick(s, 2, 0, q);
        if (sec_edge_sort()) order_edge_lanes(s, 3, rp[2], cam, q);
        cuda_ok(launch_secondary_edges_vjp(sc.dscene, cam, rp[2], gl, d_img, q), "secondary-edge adjoint kernel");
        tick(s, 2, 1, q);
        g_launches++;
    }
    if (do_pri) {
        set_shard(rp[1], npix_full * sc.sppe, sc.rank, sc.world);
        cudaStream_t q = ts.next();
        tick(s, 1, 0, q);
        order_edge_lanes(s, 1, rp[1], cam, q);
        if (edge_dynamic()) rp[1].sched = sched_counters(s, 3);
        cuda_ok(launch_primary_edges_vjp(sc.dscene, cam, rp[1], gl, d_img, q), "primary-edge adjoint kernel");
        tick(s, 1, 1, q);
        g_launches++;
    }
    ts.join();
}

static void ensure_grad_table(psdr_scene *s, size_t total) {
    if (total <= s->grad_cap) return;
    if (s->d_grad) cudaFree(s->d_grad);
    if (s->h_grad) cudaFreeHost(s->h_grad);
    s->grad_cap = total * 2;
    cuda_ok(cudaMalloc(&s->d_grad, sizeof(float) * s->grad_cap), "cudaMalloc(grad table)");
    cuda_ok(cudaMallocHost(&s->h_grad, sizeof(float) * s->grad_cap), "cudaMallocHost(grad table)");
}

// D2H of a device gradient table + the host reverse chain of configure()
static void vjp_finish(psdr_scene *s, int sensor, const float *table, cudaStream_t st) {
    Scene &sc = s->sc;
    GradLayout gl = sc.grad_layout(sensor);
    ensure_grad_table(s, (size_t) gl.total);
    cuda_ok(cudaMemcpyAsync(s->h_grad, table, sizeof(float) * gl.total, cudaMemcpyDeviceToHost, st), "D2H(grad table)");
    cuda_ok(cudaStreamSynchronize(st), "stream sync");
    sc.backprop(s->h_grad, gl, sensor);
}

int psdr_render_vjp(psdr_scene *s, int sensor, int max_depth, long long seed, int hide_emitters, int terms, int reference_scaling,
                    const int *pix_id, int npix_sel, const float *d_img, void *cuda_stream) {
    if (!s) return fail("null scene");
    PSDR_TRY
    Scene &sc = s->sc;
    if (!sc.configured) throw std::runtime_error("Input scene must be configured!");
    if (sensor < 0 || sensor >= (int) sc.cameras.size()) throw std::runtime_error("Invalid sensor id!");
    cuda_ok(cudaSetDevice(sc.device), "cudaSetDevice");
    ensure_grad_table(s, (size_t) sc.grad_layout(sensor).total);
    vjp_launch(s, sensor, max_depth, seed, hide_emitters, terms, reference_scaling, pix_id, npix_sel, d_img, s->d_grad, (cudaStream_t) cuda_stream);
    vjp_finish(s, sensor, s->d_grad, (cudaStream_t) cuda_stream);
    return 0;
    PSDR_CATCH
}

int psdr_grad_table_size(psdr_scene *s, int sensor) {
    if (!s) { fail("null scene"); return -1; }
    Scene &sc = s->sc;
    if (!sc.configured) { fail("Input scene must be configured!"); return -1; }
    if (sensor < 0 || sensor >= (int) sc.cameras.size()) { fail("Invalid sensor id!"); return -1; }
    return sc.grad_layout(sensor).total;
}

int psdr_render_vjp_device(psdr_scene *s, int sensor, int max_depth, long long seed, int hide_emitters, int terms, int reference_scaling,
                           const int *pix_id, int npix_sel, const float *d_img, float *grad_table, int n_table, void *cuda_stream) {
    if (!s) return fail("null scene");
    PSDR_TRY
    Scene &sc = s->sc;
    if (!sc.configured) throw std::runtime_error("Input scene must be configured!");
    if (sensor < 0 || sensor >= (int) sc.cameras.size()) throw std::runtime_error("Invalid sensor id!");
    if (!grad_table || n_table != sc.grad_layout(sensor).total) throw std::runtime_error("gradient table size mismatch (psdr_grad_table_size)");
    vjp_launch(s, sensor, max_depth, seed, hide_emitters, terms, reference_scaling, pix_id, npix_sel, d_img, grad_table, (cudaStream_t) cuda_stream);
    return 0;
    PSDR_CATCH
}

int psdr_scene_backprop_table(psdr_scene *s, int sensor, const float *grad_table, int n_table, void *cuda_stream) {
    if (!s) return fail("null scene");
    PSDR_TRY
    Scene &sc = s->sc;
    if (!sc.configured) throw std::runtime_error("Input scene must be configured!");
    if (sensor < 0 || sensor >= (int) sc.cameras.size()) throw std::runtime_error("Invalid sensor id!");
    if (!grad_table || n_table != sc.grad_layout(sensor).total) throw std::runtime_error("gradient table size mismatch (psdr_grad_table_size)");
    cuda_ok(cudaSetDevice(sc.device), "cudaSetDevice");
    vjp_finish(s, sensor, grad_table, (cudaStream_t) cuda_stream);
    return 0;
    PSDR_CATCH
}

int psdr_scene_get_grad(psdr_scene *s, int kind, int index, float *out, int n) {
    if (s && index >= PSDR_NESTED_BSDF_BASE && kind != PSDR_MESH_VERTICES) {
        // a nested BSDF (psdr_scene_begin_nested_bsdf): its gradients follow those of the numbered records
        const int k = index - PSDR_NESTED_BSDF_BASE;
        if (k >= (int) s->sc.nested_bsdfs.size()) return fail("invalid BSDF index");
        index = (int) s->sc.bsdfs.size() + k;
    }
    if (!s || !out) return fail("null argument");
    const Scene &sc = s->sc;
    const ParamGrads &g = sc.grads;
    if (!g.valid) return fail("no gradients: call psdr_render_vjp first");
    auto copy = [&](const double *src, int count) {
        if (n != count) return fail("gradient buffer size mismatch");
        for (int i = 0; i < count; ++i) out[i] = (float) src[i];
        return 0;
    };
    auto copy_tex = [&](const std::vector<float> &src) {
        if (n != (int) src.size()) return fail("gradient buffer size mismatch");
        std::memcpy(out, src.data(), sizeof(float) * n);
        return 0;
    };
    switch (kind) {
        case PSDR_MESH_VERTICES:
            if (index < 0 || index >= (int) g.meshes.size()) return fail("invalid mesh index");
            return copy(g.meshes[index].v.data(), (int) g.meshes[index].v.size());
        case PSDR_MESH_TO_WORLD_LEFT: case PSDR_MESH_TO_WORLD_RAW: case PSDR_MESH_TO_WORLD_RIGHT:
            if (index < 0 || index >= (int) g.meshes.size()) return fail("invalid mesh index");
            return copy(g.meshes[index].to_world[kind - PSDR_MESH_TO_WORLD_LEFT], 16);
        case PSDR_SENSOR_TO_WORLD_LEFT: case PSDR_SENSOR_TO_WORLD_RAW: case PSDR_SENSOR_TO_WORLD_RIGHT:
            if (index < 0 || index >= (int) g.cameras.size()) return fail("Invalid sensor id!");
            return copy(g.cameras[index].to_world[kind - PSDR_SENSOR_TO_WORLD_LEFT], 16);
        case PSDR_BSDF_REFLECTANCE:
            if (index < 0 || 3 * index + 3 > (int) g.bsdf_refl.size()) return fail("invalid BSDF index");
            if (index < (int) g.bsdf_tex[0].size() && !g.bsdf_tex[0][index].empty()) return copy_tex(g.bsdf_tex[0][index]);
            return copy(g.bsdf_refl.data() + 3 * index, 3);
        case PSDR_ENVMAP_RADIANCE:
            if (n != (int) g.env_radiance.size()) return fail("gradient buffer size mismatch");
            std::memcpy(out, g.env_radiance.data(), sizeof(float) * n);
            return 0;
        case PSDR_ENVMAP_SCALE: return copy(&g.env_scale, 1);
        case PSDR_ENVMAP_TO_WORLD_LEFT: return copy(g.env_to_world_left, 16);
        case PSDR_BSDF_SPECULAR:
            if (index < 0 || 3 * index + 3 > (int) g.bsdf_spec.size()) return fail("invalid BSDF index");
            if (index < (int) g.bsdf_tex[1].size() && !g.bsdf_tex[1][index].empty()) return copy_tex(g.bsdf_tex[1][index]);
            return copy(g.bsdf_spec.data() + 3 * index, 3);
        case PSDR_BSDF_ROUGHNESS:
            if (index < 0 || index >= (int) g.bsdf_rough.size()) return fail("invalid BSDF index");
            if (index < (int) g.bsdf_tex[2].size() && !g.bsdf_tex[2][index].empty()) return copy_tex(g.bsdf_tex[2][index]);
            return copy(g.bsdf_rough.data() + index, 1);
        case PSDR_BSDF_ETA:
            if (index < 0 || 3 * index + 3 > (int) g.bsdf_eta.size()) return fail("invalid BSDF index");
            return copy(g.bsdf_eta.data() + 3 * index, 3);
        case PSDR_BSDF_K:
            if (index < 0 || 3 * index + 3 > (int) g.bsdf_k.size()) return fail("invalid BSDF index");
            return copy(g.bsdf_k.data() + 3 * index, 3);
        case PSDR_BSDF_PERVERTEX:
            if (index < 0 || index >= (int) g.bsdf_pv.size() || g.bsdf_pv[index].empty()) return fail("not a MicrofacetBSDFPerVertex");
            return copy_tex(g.bsdf_pv[index]);
        case PSDR_INTEGRATOR_INTENSITY: return copy(&g.colloc_intensity, 1);
        case PSDR_BSDF_REFLECTANCE_UV: case PSDR_BSDF_SPECULAR_UV: case PSDR_BSDF_ROUGHNESS_UV: {
            const int k = kind - PSDR_BSDF_REFLECTANCE_UV;
            if (index < 0 || index >= (int) g.bsdf_tex_uv[k].size() || g.bsdf_tex_uv[k][index].empty()) return fail("the slot holds a constant: no uv transform");
            return copy_tex(g.bsdf_tex_uv[k][index]);
        }
        case PSDR_EMITTER_RADIANCE:
            if (index < 0 || 3 * index + 3 > (int) g.emitter_rad.size()) return fail("invalid emitter index");
            return copy(g.emitter_rad.data() + 3 * index, 3);
        default: return fail("unknown parameter kind");
    }
}

int psdr_scene_get_sampler_state(psdr_scene *s, long long state[6]) {
    if (!s || !state) return fail("null argument");
    for (int k = 0; k < 3; ++k) {
        state[2 * k] = s->sc.samplers[k].ready ? s->sc.samplers[k].seed : -1;
        state[2 * k + 1] = (long long) s->sc.samplers[k].consumed;
    }
    return 0;
}
int psdr_scene_set_sampler_state(psdr_scene *s, const long long state[6]) {
    if (!s || !state) return fail("null argument");
    for (int k = 0; k < 3; ++k) {
        s->sc.samplers[k].ready = state[2 * k] >= 0;
        s->sc.samplers[k].seed = state[2 * k] >= 0 ? state[2 * k] : 0;
        s->sc.samplers[k].consumed = (unsigned long long) state[2 * k + 1];
    }
    return 0;
}

int psdr_preprocess_secondary_edges(psdr_scene *s, int sensor, const int reso[4], int nrounds, long long seed, void *cuda_stream) {
    if (!s || !reso) return fail("null argument");
    PSDR_TRY
    Scene &sc = s->sc;
    cudaStream_t st = (cudaStream_t) cuda_stream;
    if (nrounds <= 0) throw std::runtime_error("nrounds > 0");
    if (!sc.configured) throw std::runtime_error("Scene needs to be configured!");
    if (sensor < 0 || sensor >= (int) sc.cameras.size()) throw std::runtime_error("Invalid sensor id!");
    if (reso[0] <= 0 || reso[1] <= 0 || reso[2] <= 0 || reso[3] <= 0) throw std::runtime_error("invalid guiding resolution");
    const long long cells = (long long) reso[0] * reso[1] * reso[2];
    if (cells * reso[3] > 2147483647LL) throw std::runtime_error("num_samples <= std::numeric_limits<int>::max()");
    if (sc.dscene.n_sec_edges <= 0) throw std::runtime_error("no secondary edges (sppse must be > 0 at configure)");
    cuda_ok(cudaSetDevice(sc.device), "cudaSetDevice");
    float *d_mass = nullptr;
    cuda_ok(cudaMalloc(&d_mass, sizeof(float) * cells), "cudaMalloc(guiding mass)");
    std::vector<float> mass((size_t) cells);
    DCamera cam = sc.dcameras[sensor];
    cam.guided = 0;                                  // the pre-pass samples the unit cube directly
    cudaError_t e = launch_guiding(sc.dscene, cam, reso, nrounds, seed, d_mass, st);
    g_launches++;
    if (e == cudaSuccess) e = cudaMemcpyAsync(mass.data(), d_mass, sizeof(float) * cells, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_mass);
    cuda_ok(e, "guiding pre-pass");
    if (nrounds > 1) for (float &m : mass) m /= (float) nrounds;
    HCamera &hc = sc.cameras[sensor];
    hc.guide.init(mass);
    for (int k = 0; k < 3; ++k) hc.greso[k] = reso[k];
    hc.guide_ready = true;
    hc.guide_enabled = true;
    sc.refresh_tables();                             // re-upload so that the kernels see the grid
    return 0;
    PSDR_CATCH
}

int psdr_scene_set_guiding(psdr_scene *s, int sensor, int enabled) {
    if (!s) return fail("null scene");
    Scene &sc = s->sc;
    if (sensor < 0 || sensor >= (int) sc.cameras.size()) return fail("Invalid sensor id!");
    HCamera &hc = sc.cameras[sensor];
    const bool on = enabled != 0 && hc.guide_ready;
    if (on != hc.guide_enabled) {
        hc.guide_enabled = on;
        if (sensor < (int) sc.dcameras.size()) sc.dcameras[sensor].guided = on ? 1 : 0;
    }
    return 0;
}

int psdr_scene_guiding_mass(psdr_scene *s, int sensor, float *out, int n) {
    if (!s || !out) return fail("null argument");
    const Scene &sc = s->sc;
    if (sensor < 0 || sensor >= (int) sc.cameras.size()) return fail("Invalid sensor id!");
    const HCamera &hc = sc.cameras[sensor];
    if (!hc.guide_ready || n != (int) hc.guide.pmf.size()) return fail("no guiding grid of that size");
    std::memcpy(out, hc.guide.pmf.data(), sizeof(float) * n);
    return 0;
}

int psdr_render_c_host(psdr_scene *s, int sensor, int max_depth, long long seed, int hide_emitters, const int *pix_id_host, int npix,
                       float *img_host) {
    if (!s) return fail("null scene");
    PSDR_TRY
    const size_t n = pix_id_host ? (size_t) npix : (size_t) s->sc.width * s->sc.height;
    ensure_staging(s, n, pix_id_host ? n : 0);
    if (pix_id_host) cuda_ok(cudaMemcpyAsync(s->d_pix, pix_id_host, sizeof(int) * n, cudaMemcpyHostToDevice, s->stream), "H2D(pix)");
    render_impl(s, sensor, max_depth, seed, hide_emitters, false, PSDR_TERM_INTERIOR, 0, pix_id_host ? s->d_pix : nullptr, npix, s->d_img, nullptr, s->stream);
    cuda_ok(cudaMemcpyAsync(img_host, s->d_img, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, s->stream), "D2H(img)");
    cuda_ok(cudaStreamSynchronize(s->stream), "stream sync");
    return 0;
    PSDR_CATCH
}

int psdr_render_d_host(psdr_scene *s, int sensor, int max_depth, long long seed, int hide_emitters, int terms, int reference_scaling,
                       const int *pix_id_host, int npix, float *img_host, float *dimg_host) {
    if (!s) return fail("null scene");
    PSDR_TRY
    const size_t n = pix_id_host ? (size_t) npix : (size_t) s->sc.width * s->sc.height;
    ensure_staging(s, n, pix_id_host ? n : 0);
    if (pix_id_host) cuda_ok(cudaMemcpyAsync(s->d_pix, pix_id_host, sizeof(int) * n, cudaMemcpyHostToDevice, s->stream), "H2D(pix)");
    const bool early = s->sc.spp > 0 && (terms & PSDR_TERM_INTERIOR);
    s->early_img_host = early ? img_host : nullptr;
    try {
        render_impl(s, sensor, max_depth, seed, hide_emitters, true, terms, reference_scaling, pix_id_host ? s->d_pix : nullptr, npix, s->d_img, s->d_dimg, s->stream);
    } catch (...) {
        s->early_img_host = nullptr;
        throw;
    }
    s->early_img_host = nullptr;
    if (!early) cuda_ok(cudaMemcpyAsync(img_host, s->d_img, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, s->stream), "D2H(img)");
    cuda_ok(cudaMemcpyAsync(dimg_host, s->d_dimg, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, s->stream), "D2H(dimg)");
    cuda_ok(cudaStreamSynchronize(s->stream), "stream sync");
    return 0;
    PSDR_CATCH
}

int psdr_render_aov(psdr_scene *s, int sensor, long long seed, float *out, void *cuda_stream) {
    if (!s) return fail("null scene");
    PSDR_TRY
    Scene &sc = s->sc;
    if (!sc.configured) throw std::runtime_error("Input scene must be configured!");
    if (sensor < 0 || sensor >= (int) sc.cameras.size()) throw std::runtime_error("Invalid sensor id!");
    cuda_ok(cudaSetDevice(sc.device), "cudaSetDevice");
    RenderParams rp{};
    rp.seed = seed < 0 ? 0 : seed;
    set_shard(rp, (long long) sc.width * sc.height * std::max(sc.spp, 1), 0, 1);
    cuda_ok(launch_aov(sc.dscene, sc.dcameras[sensor], rp, out, (cudaStream_t) cuda_stream), "aov kernel");
    g_launches++;
    return 0;
    PSDR_CATCH
}

int psdr_render_aov_d(psdr_scene *s, int sensor, long long seed, float *out, float *dout, void *cuda_stream) {
    if (!s) return fail("null scene");
    PSDR_TRY
    Scene &sc = s->sc;
    if (!sc.configured) throw std::runtime_error("Input scene must be configured!");
    if (sensor < 0 || sensor >= (int) sc.cameras.size()) throw std::runtime_error("Invalid sensor id!");
    if (!out || !dout) throw std::runtime_error("null output buffer");
    cuda_ok(cudaSetDevice(sc.device), "cudaSetDevice");
    RenderParams rp{};
    rp.seed = seed < 0 ? 0 : seed;
    set_shard(rp, (long long) sc.width * sc.height * std::max(sc.spp, 1), 0, 1);
    cuda_ok(launch_aov_d(sc.dscene, sc.dcameras[sensor], rp, out, dout, (cudaStream_t) cuda_stream), "aov (forward mode) kernel");
    g_launches++;
    return 0;
    PSDR_CATCH
}

int psdr_render_field_edges(psdr_scene *s, int sensor, long long seed, int field, int object, float *dimg, void *cuda_stream) {
    if (!s) return fail("null scene");
    PSDR_TRY
    Scene &sc = s->sc;
    if (!sc.configured) throw std::runtime_error("Input scene must be configured!");
    if (sensor < 0 || sensor >= (int) sc.cameras.size()) throw std::runtime_error("Invalid sensor id!");
    if (field < 0 || field > 6) throw std::runtime_error("Unsupported field");
    if (!dimg) throw std::runtime_error("null output buffer");
    cuda_ok(cudaSetDevice(sc.device), "cudaSetDevice");
    cudaStream_t st = (cudaStream_t) cuda_stream;
    const long long npix = (long long) sc.width * sc.height;
    cuda_ok(cudaMemsetAsync(dimg, 0, sizeof(float) * 3 * npix, st), "memset(dimg)");
    if (sc.sppe <= 0 || sc.dcameras[sensor].n_edges <= 0) return 0;
    RenderParams rp{};
    rp.seed = seed < 0 ? 0 : seed;
    set_shard(rp, npix * sc.sppe, 0, 1);
    cuda_ok(launch_field_edges(sc.dscene, sc.dcameras[sensor], rp, field, object, dimg, st), "field edge kernel");
    g_launches++;
    return 0;
    PSDR_CATCH
}

int psdr_render_field_vjp(psdr_scene *s, int sensor, long long seed, int field, int object, int terms, int reference_scaling,
                          const float *d_img, void *cuda_stream) {
    if (!s) return fail("null scene");
    PSDR_TRY
    Scene &sc = s->sc;
    if (!sc.configured) throw std::runtime_error("Input scene must be configured!");
    if (sensor < 0 || sensor >= (int) sc.cameras.size()) throw std::runtime_error("Invalid sensor id!");
    if (field < 0 || field > 6) throw std::runtime_error("Unsupported field");
    if (!d_img) throw std::runtime_error("null cotangent image");
    if (sc.world > 1) throw std::runtime_error("FieldExtractionIntegrator: reverse mode runs unsharded");
    cuda_ok(cudaSetDevice(sc.device), "cudaSetDevice");
    cudaStream_t st = (cudaStream_t) cuda_stream;
    GradLayout gl = sc.grad_layout(sensor);
    ensure_grad_table(s, (size_t) gl.total);
    gl.base = s->d_grad;
    cuda_ok(cudaMemsetAsync(s->d_grad, 0, sizeof(float) * gl.total, st), "memset(grad table)");
    const long long npix = (long long) sc.width * sc.height;
    const DCamera &cam = sc.dcameras[sensor];
    RenderParams rp{};
    rp.seed = seed < 0 ? 0 : seed;
    rp.use_field = 1;
    rp.field = field;
    rp.field_object = object;
    // the reference's field images and their interior derivatives come out 2x; its edge term is 1x
    if ((terms & PSDR_TERM_INTERIOR) && sc.spp > 0 && field >= 2) {
        rp.tangent_scale = reference_scaling ? 2.f : 1.f;
        set_shard(rp, npix * sc.spp, 0, 1);
        cuda_ok(launch_interior_vjp(sc.dscene, cam, rp, gl, d_img, st), "field adjoint kernel");
        g_launches++;
    }
    if ((terms & PSDR_TERM_PRIMARY_EDGES) && sc.sppe > 0 && cam.n_edges > 0) {
        rp.tangent_scale = 1.f;
        set_shard(rp, npix * sc.sppe, 0, 1);
        cuda_ok(launch_primary_edges_vjp(sc.dscene, cam, rp, gl, d_img, st), "field edge adjoint kernel");
        g_launches++;
    }
    vjp_finish(s, sensor, s->d_grad, st);
    return 0;
    PSDR_CATCH
}

int psdr_sampler_draws(long long seed, int n, int ndraws, float *out) {
    if (!out || n < 0 || ndraws < 0) return fail("invalid arguments");
    for (int i = 0; i < n; ++i) {
        Pcg32 r;
        r.seed((unsigned long long) (i + seed), (unsigned long long) i);
        for (int k = 0; k < ndraws; ++k) out[(size_t) k * n + i] = r.next_1d();
    }
    return 0;
}

}  // extern "C"
