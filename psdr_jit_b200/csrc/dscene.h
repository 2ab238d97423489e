// Device-side scene tables read by the sm_100a kernels (all read-only during a render call).
// Layout: float4-packed structure-of-arrays so that one triangle's geometry is three 16-byte
// loads; everything a Cornell-box-class scene needs is a few KB and stays L1/L2 resident,
// bigger meshes stream through L2 (126 MB) -- HBM is touched once per call.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "texture.h"

namespace psdr {

// triangle i: geo[3i+0] = (p0.x,p0.y,p0.z,e1.x) geo[3i+1] = (e1.y,e1.z,e2.x,e2.y)
//             geo[3i+2] = (e2.z, face_area, int_as_float(mesh_id), 0)
//             shade[3i+0] = (n0.xyz, n1.x) shade[3i+1] = (n1.yz, n2.xy) shade[3i+2] = (n2.z, fn.xyz)
// dgeo/dshade: the forward-mode tangents in the same layout (dgeo[3i+2].z/.w unused).
// uv[3i+k] = texture coordinate of corner k (zeros for meshes without UVs)
constexpr int kMaxBruteTris = 64;
constexpr int kBrutePairWords = 9;    // 64-bit words per triangle pair in DScene::bg_pair
constexpr int kBruteSmemStride = 5;   // 16-byte units per pair in the kernels' shared-memory copy (9 words + 1 pad:
                                      // an odd unit stride keeps the 8 lanes of a 128-bit load phase on distinct banks)
constexpr int kBruteBoxWords = 6;     // 64-bit words per PAIR of cull boxes in DScene::bg_box

struct DBvhNode {        // 32 B
    float lo[3];
    int a;               // inner: left child;  leaf: first index into tri_order
    float hi[3];
    int b;               // inner: right child; leaf: -(count)
};

// Traversal node: BOTH children's boxes and links in one 64-byte record (four 16-byte loads), so a visit decides both
// children without touching them; leaves are not nodes -- a leaf child is a run of slots in DScene::leaf_tri.
//   box0 = (f[0..2], f[3..5]), box1 = (f[6..8], f[9..11]);  c0 / c1 >= 0: inner node index, < 0: leaf ~((first_slot << 3) | count)
//   parent / pslot: where this node's own box lives (refit walks up), parent < 0 for the root
struct DBvhNode2 {
    float f[12];
    int c0, c1, parent, pslot;
};

struct DMesh {
    int bsdf;            // index into bsdf tables, -1 = none
    int emitter;         // index into emitter table, -1 = none
    int flags;           // bit0 use_face_normals, bit1 has_uv
    int face_offset;
};

struct DEmitter {
    int type;                // 0 AreaLight, 1 EnvironmentMap (fields below unused, see DEnv)
    int pad_;
    float radiance[3];
    float d_radiance[3];
    int mesh;
    float sampling_weight;   // normalised: weight * rcp(sum of weights)
    int face_offset;         // global id of the mesh's first triangle
    int nfaces;
    float inv_total_area;
    int distrb_offset;       // into face_pmf / face_cmf
    float face_sum;          // drjit-style fp32 sum of the face areas
    float emitter_pmf;       // unnormalised selection weight (area*luminance)
    float pad0, pad1;
};

// EnvironmentMap tables (reference src/emitter/envmap.cpp)
struct DEnv {
    int present, emitter;
    int w, h;                          // radiance bitmap (lat-long)
    const float *data, *ddata;         // rgb interleaved [h][w][3] + forward tangents
    float scale, d_scale;
    float to_world[9], d_to_world[9];  // 3x3 direction transforms, row-major
    float from_world[9], d_from_world[9];
    float lower[3], upper[3];          // scene bounding box the "positions" live on
    int cw, ch;                        // cell grid 2(w-1) x 2(h-1)
    float cell_sum;
    const float *cell_pmf, *cell_cmf;
    const int *cell_lut;               // bucket table of the cell CDF (sample_reuse_lut; cell_lut_n buckets, 0 = none):
    int cell_lut_n;                    // 21 dependent loads of the bisection over 2 M cells become ~4
};

struct DBsdf {
    float refl[3];           // Diffuse: reflectance; Microfacet: diffuseReflectance
    float d_refl[3];
    int type;                // 0 Diffuse, 1 Microfacet, 2 RoughConductor, 3 RoughDielectric, 4 MicrofacetPerVertex, 5 NormalMap
    int two_side;
    float spec[3];           // Microfacet: specularReflectance (F0); RoughConductor: specular_reflectance
    float d_spec[3];
    float rough, d_rough;    // Microfacet: roughness (alpha = roughness^2); RoughConductor: alpha (alpha_u = alpha_v)
    float eta[3], d_eta[3];  // RoughConductor: complex index of refraction eta + i k per channel;
                             // RoughDielectric: eta[0] = intIOR / extIOR, eta[1] = extIOR / intIOR (rounded separately, as the reference does)
    float kk[3], d_kk[3];
    // texture slots (texture.h DTex): 0 reflectance / diffuseReflectance (Bitmap3fD), 1 specularReflectance (Bitmap3fD),
    // 2 roughness (Bitmap1fD); w * h == 0: the constant above
    // NormalMap: slot 0 (refl / tex[0]) is the normal map, `nested` the record of the BSDF it perturbs (nested records
    // follow the n_bsdfs user-visible ones in DScene::bsdfs)
    DTex tex[3];
    int nested;
    // MicrofacetPerVertex: pv_n vertices x 7 floats (specular rgb, diffuse rgb, roughness) and their forward tangents
    int pv_n;
    const float *pv, *d_pv;
    int pv_goff;             // offset of the table's gradients, relative to the END of the adjoint's gradient table
};

struct DCamera {
    float sample_to_camera[16];
    float to_world[16], d_to_world[16];
    float world_to_sample[16], d_world_to_sample[16];
    float pos[3], d_pos[3], dir[3], d_dir[3];
    float inv_area;
    int n_edges;             // primary edges (0 = none / not an active sensor)
    float edge_sum;
    int ortho;               // 1: OrthographicCamera -- rays leave the near plane along +z of the camera (orthographic.cpp:109-131)
    // primary edges: pe_a = (p0.x,p0.y,p1.x,p1.y), pe_da = tangents, pe_b = (nx,ny,len,0)
    const float4 *pe_a, *pe_da, *pe_b;
    const float *pe_pmf, *pe_cmf;
    // secondary-edge guiding grid (HyperCubeDistribution<3>, reference src/core/cube_distrb.cpp:9-64)
    int guided;              // 0 = sample3 is used as drawn
    int greso[3];
    float guide_sum;
    const float *guide_pmf, *guide_cmf;
    const int *guide_lut;    // bucket table of the grid's CDF (sample_reuse_lut, device_path.cuh); guide_lut_n buckets (0 = none)
    int guide_lut_n;
};

struct DScene {
    int width, height, spp, sppe, sppse;
    int n_tris, n_meshes, n_emitters, n_bsdfs, n_sec_edges, n_nodes;
    int use_bvh;             // 0: brute force over all triangles (tiny scenes)
    int ref_rcp;             // 1: the analytic primary hit of renderD uses Dr.Jit's approximate rcp (device_path.cuh rcp_approx)
    int full_features;       // 1: some BSDF is a Microfacet or an EnvironmentMap exists (selects the kernel variant)
    int ext_features;        // 1: bitmap-valued BSDF slots or a BSDF of type >= 2 (kCfgExt kernel family; implies full_features)
    float colloc_intensity, d_colloc_intensity;   // CollocatedIntegrator::m_intensity and its forward tangent (RenderParams::mis == 3)
    int colloc_field;        // 1: the "bsdf" field of FieldExtractionIntegrator -- BSDF(wi, wi) alone, no intensity / t^2 (field.cpp:72-92)
    const float4 *geo, *shade, *dgeo, *dshade;
    const float2 *uv;
    const int *face_idx;     // 3 mesh-local vertex indices per triangle (MicrofacetPerVertex gathers through them); nullptr if unused
    const DMesh *meshes;
    const DEmitter *emitters;
    const DBsdf *bsdfs;
    const float *face_pmf, *face_cmf;          // concatenated per-emitter face distributions
    const float *emitter_pmf, *emitter_cmf;    // emitter selection (size n_emitters)
    float emitter_sum;
    // secondary edges: 6 float4 per edge:
    //  [0]=(p0.xyz,e1.x) [1]=(e1.yz, dp0.xy) [2]=(dp0.z, de1.xyz) [3]=(n0.xyz, n1.x) [4]=(n1.yz, p2.xy) [5]=(p2.z, boundary,0,0)
    const float4 *sec_edges;
    const float *sec_pmf, *sec_cmf;
    float sec_sum;
    const int *sec_lut;      // bucket table of the secondary-edge CDF; sec_lut_n buckets (0 = none)
    int sec_lut_n;
    const DBvhNode2 *nodes2;                   // BVH mode: wide-fetch BVH2 (device_upload.cu builds / refits it)
    const float4 *leaf_tri;                    // 3 float4 per leaf slot: (p0.xyz, int_as_float(triangle id)), (e1.xyz, 0), (e2.xyz, 0)
    DEnv env;
    // brute-force mode (n_tris <= kMaxBruteTris): the triangle geometry again, by value -- it travels in
    // the kernel parameters, two triangles at a time: entry [9 j + c] holds component c of triangles (2j, 2j+1)
    // as the two halves of one 64-bit word, c = p0.xyz, e1.xyz, e2.xyz, the operand layout of the packed fp32x2
    // closest-hit test (device_path.cuh); every CTA copies it to shared memory once (brute_init).
    // An odd count is padded with an all-zero triangle (det = 0: never accepted).
    unsigned long long bg_pair[kMaxBruteTris / 2 * kBrutePairWords];
    // one padded bounding box per triangle pair, two boxes per 64-bit word (boxes 2k, 2k+1 in the low / high
    // half): entry [6 k + c], c = centre.xyz, half-extent.xyz.  A ray tests a pair only if it passes that pair's
    // slab test (trace<brute>, stage 1); read warp-uniformly through the constant bank.
    unsigned long long bg_box[kMaxBruteTris / 4 * kBruteBoxWords];
};

struct RenderParams {
    int max_depth;
    int hide_emitters;
    int mis;                 // 2: PathTracer / Direct(2) (both strategies, power heuristic); 0 / 1: Direct(0) / Direct(1);
                             // 3: CollocatedIntegrator (Li = BSDF(wi, wi) intensity / t^2 at the primary hit, no sampling)
    long long seed;          // >= 0
    unsigned long long skip; // draws already consumed per lane of this sampler (seed = -1 continuation)
    long long lane_begin, lane_end;   // LOCAL lane index range of this call: [0, 32 * owned blocks)
    // multi-GPU sharding: the lanes of a term are dealt to the ranks in blocks of 32 (one warp; one pixel at spp = 32),
    // round-robin -- local lane j is global lane ((j / 32) * shard_world + shard_rank) * 32 + j % 32; lanes >= n_lanes
    // (the tail of the last block) are dead.  Contiguous ranges gave the rank that owns the empty top rows of the
    // image a quarter of the interior work of the others.
    long long n_lanes;
    int shard_rank, shard_world;
    const int *pix_id;       // batch mode: pixel list (device), else nullptr
    int npix;                // number of output pixels (W*H or len(pix_id))
    int smem_grad;           // adjoint kernels: 1 = accumulate into a shared-memory copy of the gradient table
    float tangent_scale;     // 1, or 2 to reproduce the reference's forward-mode scaling
    int *sched;              // large-CTA interior kernels: {next chunk, CTAs done} of the dynamic chunk hand-out (device, zero between launches)
    int use_field;           // 1: the adjoint launches evaluate FieldExtractionIntegrator (field, field_object) instead of an Li
    int field, field_object; // FieldExtractionIntegrator in reverse mode (kernels_vjp_impl.cuh kField): field id (device_path.cuh
                             // field_value) and the mesh to keep (-1 = all)
    int out_multicast;       // 1, 2 = the output pointers are NVLS multicast addresses: every add is a multimem.red that lands
                             // in the replica of EVERY rank; 2: images are float32[npix][4] (psdr_scene_set_output_multicast)
    const int *perm;         // primary-edge kernels: thread j evaluates local lane perm[j] (edge_sort.cu: the lanes bucketed by
                             // their position along the edge list, so that a warp's rays start next to each other); nullptr = j
};

}  // namespace psdr
