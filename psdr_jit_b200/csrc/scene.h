// Host-side scene graph of the B200 path tracer: owns the objects a user adds through the C ABI
// (mirroring psdr_jit.Scene: reference src/scene/scene.cpp), evaluates `configure()` in fp32 +
// forward-mode duals on the host, builds the BVH2 and uploads float4-packed tables (dscene.h).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "dscene.h"
#include "grad_layout.h"
#include "pmath.h"

namespace psdr {

struct Distrib {   // reference src/core/pmf.cpp:6-15, include/psdr/core/pmf.h:12-38
    int size = 0;
    float sum = 0.f;
    std::vector<float> pmf, cmf;
    void init(const std::vector<float> &p);
};

struct HTri {
    V3d p0, e1, e2, n0, n1, n2, fn;
    Dual area;
};

struct HEdge {
    int v0, v1, f0, f1, v2;
};

struct HBsdf {
    std::string id;
    int type = 0;             // 0 Diffuse, 1 Microfacet, 2 RoughConductor, 3 RoughDielectric, 4 MicrofacetPerVertex, 5 NormalMap
    V3d reflectance;          // Diffuse reflectance / Microfacet diffuseReflectance / NormalMap: the (constant) normal map
    V3d specular;             // Microfacet specularReflectance / RoughConductor specular_reflectance
    Dual roughness;           // Microfacet roughness / RoughConductor alpha
    V3d eta, k;               // RoughConductor eta, k; RoughDielectric: eta.x = intIOR / extIOR, eta.y = extIOR / intIOR
    bool two_side = false;
    int nested = -1;          // NormalMap: index into Scene::nested_bsdfs of the BSDF it perturbs
    std::vector<float> pv, d_pv;   // MicrofacetPerVertex: 7 floats per vertex (specular rgb, diffuse rgb, roughness) + tangents
    // texture slots (Bitmap with more than one texel; channels interleaved, pixel = y*w + x):
    // 0 reflectance / diffuseReflectance (3 channels), 1 specularReflectance (3), 2 roughness (1);
    // each with the bitmap's uv transform (reference include/psdr/core/bitmap.h:36-38: m_scale, m_rot, m_trans)
    struct Tex {
        int w = 0, h = 0;
        std::vector<float> data, ddata;
        Dual scale = Dual(1.f), rot = Dual(0.f), tx = Dual(0.f), ty = Dual(0.f);
    };
    Tex tex[3];
    static int tex_channels(int slot) { return slot == 2 ? 1 : 3; }
};

struct HMesh {
    std::vector<V3d> v_raw;
    std::vector<int> f, fuv;
    std::vector<V2f> uv;
    bool has_uv = false;
    M4<Dual> to_world[3];   // left, raw, right
    int bsdf = -1, emitter = -1;
    bool use_face_normals = false, enable_edges = true;
    bool edges_dirty = true;
    bool is_bound_mesh = false;   // the envmap's bounding box (appended by configure)
    // configured
    std::vector<V3d> v_world;
    std::vector<HTri> tris;
    std::vector<HEdge> edges;
    Distrib face_distrb;
    float total_area = 0.f, inv_total_area = 0.f;
    int face_offset = 0;
};

struct HEmitter {
    int type = 0;             // 0 AreaLight, 1 EnvironmentMap
    V3d radiance;
    int mesh = -1;
    float sampling_weight = 0.f, raw_weight = 0.f;
};

// EnvironmentMap (reference src/emitter/envmap.cpp, include/psdr/emitter/envmap.h): lat-long radiance bitmap,
// scale, to_world = left * raw; `cell` = HyperCubeDistribution2f over 2(w-1) x 2(h-1) cells (x-major)
struct HEnvmap {
    bool present = false;
    int emitter = -1, mesh = -1;      // index into emitters / the bounding mesh (appended by configure)
    int w = 0, h = 0;
    std::vector<float> data, ddata;   // rgb interleaved, row-major (pixel = y*w + x); forward tangents
    Dual scale = Dual(1.f);
    M4<Dual> to_world[2];             // left, raw
    // configured
    M4<Dual> to_world_full, from_world;
    bool has_bounds = false;          // the bounding box is fixed by the first configure() (scene.cpp:435)
    V3f lower, upper;
    int cw = 0, ch = 0;
    Distrib cell;
    // versions: data_version changes with every write of the radiance texels (psdr_scene_set_param), ddata_version with
    // their tangents; the cell table is rebuilt and the big device tables re-uploaded only when they differ from the
    // *_built / *_uploaded copies (an optimisation loop that only moves geometry pays neither)
    unsigned data_version = 1, ddata_version = 1, cell_built = 0;
};

struct HPrimEdge {
    V2d p0, p1;
    V2f normal;
    float length;
    int mesh, v0, v1;   // where the endpoints come from (reverse mode)
};

struct HCamera {
    float fov = 60.f, near_ = 1e-6f, far_ = 1e7f;
    bool ortho = false;                  // OrthographicCamera(near, far) (include/psdr/sensor/orthographic.h, src/sensor/orthographic.cpp)
    bool use_intrinsic = false;          // PerspectiveCamera(fx, fy, cx, cy, near, far) (perspective.h:11-12)
    float fx = 0.f, fy = 0.f, cx = 0.f, cy = 0.f;
    M4<Dual> to_world[3];
    M4<Dual> to_world_full, world_to_sample;
    M4<float> sample_to_camera, camera_to_sample;
    V3d pos, dir;
    float inv_area = 0.f;
    std::vector<HPrimEdge> edges;
    Distrib edge_distrb;
    // secondary-edge guiding (PathTracer::preprocess_secondary_edges); survives configure()
    bool guide_ready = false, guide_enabled = false;
    int greso[3] = {0, 0, 0};
    Distrib guide;
};

struct HSecEdge {
    V3d p0, e1;
    V3f n0, n1, p2;
    bool is_boundary;
    int mesh, v0, v1;
};

// Gradients of one scalar loss with respect to every parameter (filled by Scene::backprop).
struct ParamGrads {
    struct MeshG { std::vector<double> v; double to_world[3][16]; };
    struct CamG { double to_world[3][16]; };
    std::vector<MeshG> meshes;
    std::vector<CamG> cameras;
    std::vector<double> bsdf_refl, emitter_rad, bsdf_spec, bsdf_eta, bsdf_k;   // 3 per object
    std::vector<double> bsdf_rough;                          // 1 per BSDF
    std::vector<float> env_radiance;                         // 3*w*h
    std::vector<std::vector<float>> bsdf_tex[3];             // per BSDF and texture slot: channels*w*h (empty: not textured)
    std::vector<std::vector<float>> bsdf_tex_uv[3];          // per BSDF and textured slot: d/d(scale, rotation, translate.x, translate.y)
    std::vector<std::vector<float>> bsdf_pv;                 // per BSDF: 7 floats per vertex (MicrofacetPerVertex; empty otherwise)
    double env_scale = 0.0, env_to_world_left[16] = {};
    double colloc_intensity = 0.0;                           // CollocatedIntegrator::m_intensity
    bool valid = false;
};

struct DeviceBuffers;  // owns every cudaMalloc'ed table

struct SamplerState {   // reference Sampler (src/core/sampler.cpp): stateless on the device --
    bool ready = false; // a stream is (seed, lane) + number of draws consumed so far
    long long sample_count = 0;
    long long seed = 0;
    unsigned long long consumed = 0;
};

struct Scene {
    int width = 128, height = 128, spp = 1, sppe = 0, sppse = 0, log_level = 1;  // RenderOption defaults (types.h:217-228)
    long long seed = 0;
    std::vector<HBsdf> bsdfs;
    std::vector<HBsdf> nested_bsdfs;   // the BSDFs NormalMap records wrap: not numbered, not in param_map (scene.cpp:128-145)
    std::vector<HMesh> meshes;
    std::vector<HEmitter> emitters;
    std::vector<HCamera> cameras;
    std::vector<HSecEdge> sec_edges;
    HEnvmap env;
    Distrib sec_edge_distrb, emitter_distrb;
    SamplerState samplers[3];
    bool configured = false;
    int device = 0;
    int rank = 0, world = 1;   // lane-range sharding across GPUs
    int force_bvh = -1;        // -1 auto, 0 brute force, 1 bvh
    bool bvh_rebuild = false;  // force a topology rebuild at the next configure (psdr_scene_set_accel)
    int bvh_builds = 0, bvh_refits = 0;   // host topology builds / GPU refits so far (psdr_scene_query)
    int integrator_mis = 2;    // 2 PathTracer / Direct(2); 0, 1: Direct(0), Direct(1); 3: CollocatedIntegrator (psdr_scene_set_integrator)
    Dual colloc_intensity = Dual(0.f);   // CollocatedIntegrator::m_intensity
    bool colloc_field = false;           // FieldExtractionIntegrator("bsdf"): the BSDF term alone
    int out_multicast = 0;        // outputs are NVLS multicast addresses: 1 = same layout, 2 = float4 pixels (psdr_scene_set_output_multicast)
    bool ref_rcp = false;      // reference arithmetic for the analytic primary hit (psdr_scene_set_reference_arithmetic)
    DeviceBuffers *dev = nullptr;
    DScene dscene{};
    std::vector<DCamera> dcameras;
    double last_configure_ms = 0.0;
    size_t upload_bytes = 0;
    ParamGrads grads;

    // reverse mode: table layout for `sensor` (base = nullptr) and the host chain
    // table gradients -> world vertices -> raw vertices / to_world / camera matrices (scene_grad.cpp)
    GradLayout grad_layout(int sensor) const;
    // reverse mode numbers the BSDF records as the device does: the numbered ones, then the nested ones
    int num_bsdf_records() const { return (int) (bsdfs.size() + nested_bsdfs.size()); }
    const HBsdf &bsdf_record(int i) const { return i < (int) bsdfs.size() ? bsdfs[i] : nested_bsdfs[i - (int) bsdfs.size()]; }
    int texture_grad_offset(int bsdf, int slot) const;   // relative to GradLayout::total (negative)
    int pervertex_grad_offset(int bsdf) const;           // MicrofacetPerVertex tables: the last blocks of the table (negative)
    void backprop(const float *table, const GradLayout &gl, int sensor);

    Scene();
    ~Scene();
    int find_bsdf(const std::string &id) const;
    void configure(const int *active, int nactive);   // throws std::runtime_error
    void refresh_tables();                             // re-pack + upload the device tables of a configured scene
};

// BVH2 over world-space triangles (binned SAH); nodes/order are what DScene points to.
void build_bvh(const std::vector<HTri> &tris, std::vector<DBvhNode> &nodes, std::vector<int> &order, int leaf_size);

}  // namespace psdr
