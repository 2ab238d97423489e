/* psdr_b200.h -- C ABI of the B200-native path-space differentiable path tracer.
 *
 * Drop-in boundary for the integrator path of andyyankai/psdr-jit.  The reference has no FFI seam
 * of its own: its boundary is the pybind11 module surface in src/psdr.cpp, so every entry point
 * below names the reference binding it stands in for.  Plain pointers and sizes only; no C++ or
 * torch types; no exception crosses this boundary (functions return 0 on success, non-zero on
 * error, psdr_last_error() gives the message -- the reference throws psdr_jit::Exception,
 * include/misc/Exception.h:7-40, which pybind11 turns into RuntimeError).
 *
 * Threading: re-entrant per scene handle, not thread-safe on one handle.  Render calls are
 * asynchronous on the given CUDA stream (pass NULL for the default stream) and never synchronise;
 * the *_host variants copy through pinned staging and return when the host buffers are filled.
 */
#ifndef PSDR_B200_H
#define PSDR_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct psdr_scene psdr_scene;

/* Parameter kinds for psdr_scene_set_param / psdr_scene_set_tangent.  They are the fields the
 * reference exposes through Scene.param_map objects (src/psdr.cpp:314-339,344-347,279-284,357-370). */
enum {
    PSDR_MESH_VERTICES = 0,        /* Mesh.vertex_positions (object space), n = 3*num_vertices, xyz interleaved */
    PSDR_MESH_TO_WORLD_LEFT = 1,   /* Mesh.to_world_left  = Mesh.set_transform(mat, set_left=True), 16 floats row-major */
    PSDR_MESH_TO_WORLD_RAW = 2,    /* Mesh.to_world */
    PSDR_MESH_TO_WORLD_RIGHT = 3,  /* Mesh.to_world_right */
    PSDR_SENSOR_TO_WORLD_LEFT = 4, /* Sensor.to_world_left = Sensor.set_transform */
    PSDR_SENSOR_TO_WORLD_RAW = 5,  /* Sensor.to_world */
    PSDR_SENSOR_TO_WORLD_RIGHT = 6,
    PSDR_BSDF_REFLECTANCE = 7,     /* DiffuseBSDF.reflectance / MicrofacetBSDF.diffuseReflectance (1x1 bitmap), 3 floats */
    PSDR_EMITTER_RADIANCE = 8,     /* AreaLight.radiance, 3 floats */
    PSDR_BSDF_SPECULAR = 9,        /* MicrofacetBSDF.specularReflectance (1x1 bitmap), 3 floats */
    PSDR_BSDF_ROUGHNESS = 10,      /* MicrofacetBSDF.roughness (1x1 bitmap), 1 float */
    PSDR_ENVMAP_RADIANCE = 11,     /* EnvironmentMap.radiance (lat-long Bitmap3fD), 3*w*h floats, rgb interleaved, index ignored */
    PSDR_ENVMAP_SCALE = 12,        /* EnvironmentMap.scale, 1 float */
    PSDR_ENVMAP_TO_WORLD_LEFT = 13,/* EnvironmentMap.set_transform(mat), 16 floats */
    /* Bitmap.scale / .rotate / .translate (src/psdr.cpp:204-206,217-219; the uv transform of src/core/bitmap.cpp:64-72)
     * of a BSDF's texture slots: 4 floats (scale, rotate, translate.x, translate.y); forward-mode tangents supported,
     * no reverse-mode gradient */
    PSDR_BSDF_REFLECTANCE_UV = 14,
    PSDR_BSDF_SPECULAR_UV = 15,
    PSDR_BSDF_ROUGHNESS_UV = 16,
    PSDR_BSDF_ETA = 17,            /* RoughConductorBSDF.eta (1x1 Bitmap3fD), 3 floats */
    PSDR_BSDF_K = 18,              /* RoughConductorBSDF.k (1x1 Bitmap3fD), 3 floats */
    PSDR_BSDF_PERVERTEX = 19,      /* MicrofacetBSDFPerVertex tables, 7 floats per vertex: specularReflectance rgb,
                                      diffuseReflectance rgb, roughness (src/psdr.cpp:306-310) */
    PSDR_INTEGRATOR_INTENSITY = 20 /* psdr_scene_get_grad only: d/d CollocatedIntegrator.m_intensity of the last adjoint pass */
};

/* Texture slots of a BSDF (psdr_scene_set_bsdf_texture_slot). */
enum { PSDR_TEX_REFLECTANCE = 0, PSDR_TEX_SPECULAR = 1, PSDR_TEX_ROUGHNESS = 2 };

/* What psdr_scene_query() can return. */
enum {
    PSDR_Q_NUM_MESHES = 0,          /* Scene.num_meshes */
    PSDR_Q_NUM_SENSORS = 1,         /* Scene.num_sensors */
    PSDR_Q_NUM_EMITTERS = 2,        /* Scene.get_num_emitters() */
    PSDR_Q_NUM_TRIANGLES = 3,
    PSDR_Q_NUM_PRIMARY_EDGES = 4,   /* index = sensor; "(n) primary edges initialized" (scene.cpp:425-431) */
    PSDR_Q_NUM_SECONDARY_EDGES = 5, /* "n secondary edges initialized" (scene.cpp:563-567) */
    PSDR_Q_NUM_MESH_EDGES = 6,      /* index = mesh; columns of Mesh.edge_indices() */
    PSDR_Q_NUM_MESH_VERTICES = 7,
    PSDR_Q_NUM_MESH_FACES = 8,
    PSDR_Q_IS_CONFIGURED = 9,
    PSDR_Q_USES_BVH = 10,
    PSDR_Q_UPLOAD_BYTES = 11,       /* bytes of device tables the last configure() copied host->device */
    PSDR_Q_GUIDING_CELLS = 12,      /* index = sensor; cells of its secondary-edge guiding grid (0 = none) */
    PSDR_Q_BVH_BUILDS = 13,         /* host BVH topology builds so far (first configure, or after the set of meshes changed) */
    PSDR_Q_BVH_REFITS = 14,         /* GPU BVH refits so far (every configure of a scene above 64 triangles) */
    PSDR_Q_GRAD_TABLE_MULTICAST = 15, /* index = sensor; 1 = psdr_render_vjp_device may target a multicast table
                                      * (psdr_scene_set_output_multicast): the table fits the kernels' shared-memory copy */
    PSDR_Q_KERNEL_FAMILY = 16        /* which kernel instantiation the configured scene runs: bit 0 BVH2 traversal, bit 1
                                      * Microfacet / EnvironmentMap code, bit 3 extended materials (bitmaps, RoughConductor,
                                      * RoughDielectric, MicrofacetPerVertex, NormalMap); -1 before configure() */
};

/* Terms of renderD (bit mask). */
enum { PSDR_TERM_INTERIOR = 1, PSDR_TERM_PRIMARY_EDGES = 2, PSDR_TERM_SECONDARY_EDGES = 4, PSDR_TERM_ALL = 7 };

const char *psdr_last_error(void);
int psdr_version(void);
/* Number of this library's kernel launches so far in the calling process (bench.py's gpu_launches). */
long long psdr_kernel_launch_count(void);

/* Scene() -- src/psdr.cpp:393, src/scene/scene.cpp:49-57.  `device` = CUDA ordinal. */
psdr_scene *psdr_scene_create(int device);
/* Scene::~Scene -- src/scene/scene.cpp:60-71 */
void psdr_scene_destroy(psdr_scene *s);

/* Scene.opts (RenderOption, src/psdr.cpp:125-144, include/psdr/types.h:217-228) */
int psdr_scene_set_options(psdr_scene *s, int width, int height, int spp, int sppe, int sppse, int log_level);
/* Scene.seed -- src/psdr.cpp:411 (used by configure() when it seeds the samplers) */
int psdr_scene_set_seed(psdr_scene *s, long long seed);
/* New (the reference is single-GPU): the lanes of every term are dealt to the ranks round-robin in blocks of
 * 32; this process renders blocks rank, rank + world, ... into a full-frame buffer and the caller sums the
 * buffers over ranks (one NCCL all-reduce).  In reverse mode the caller sums the gradient TABLES
 * (psdr_render_vjp_device) over ranks before psdr_scene_backprop_table. */
int psdr_scene_set_shard(psdr_scene *s, int rank, int world);
/* Which Li the following render calls evaluate.  PSDR_INTEGRATOR_PATH: PathTracer(max_depth) (src/integrator/path.cpp:35-127).
 * PSDR_INTEGRATOR_DIRECT: Direct(mis) (src/psdr.cpp:436-439, src/integrator/direct.cpp:35-131): one bounce (call the render
 * entry points with max_depth = 1); mis = 2 both strategies with the power heuristic (= PathTracer(1)), mis = 0 emitter
 * sampling only, mis = 1 BSDF sampling only.  The sample streams consume only the draws the mode uses. */
enum { PSDR_INTEGRATOR_PATH = 0, PSDR_INTEGRATOR_DIRECT = 1 };
int psdr_scene_set_integrator(psdr_scene *s, int kind, int mis);
/* CollocatedIntegrator(intensity) -- src/psdr.cpp:427-429, src/integrator/collocated.cpp:21-53: a point light at the camera;
 * Li = BSDF(wi, wo = wi) * intensity / t^2 at the primary hit, no emitters, no random numbers in Li, no secondary-edge
 * term (Integrator::render_secondary_edges is empty, include/psdr/integrator/integrator.h:22).  d_intensity: forward-mode
 * tangent of m_intensity.  The following render calls (max_depth ignored) evaluate it until psdr_scene_set_integrator.
 * bsdf_field = 1: the "bsdf" field of FieldExtractionIntegrator (src/integrator/field.cpp:72-92) -- the BSDF term alone,
 * without intensity / t^2 (forward mode only). */
int psdr_scene_set_integrator_collocated(psdr_scene *s, float intensity, float d_intensity, int bsdf_field);
/* Multi-GPU output fusion (new: SURVEY.md 8e).  on = 1 or 2: the img / dimg pointers of psdr_render_c / psdr_render_d and the
 * grad_table pointer of psdr_render_vjp_device are NVLS MULTICAST addresses of a buffer that every rank of the node has
 * mapped (cuMulticast* / torch symmetric memory).  The term kernels then accumulate with multimem.red: the NVSwitch adds
 * every contribution into the replica of every GPU, so the complete image (or gradient table) is present on all ranks
 * when the kernels of all ranks have finished -- no all-reduce pass.  The calls do NOT zero the buffer in this mode: the
 * caller zeroes every replica and synchronises the ranks before, and synchronises them again before reading
 * (psdr_jit_b200/dist.py PeerBuffers does both with one device-side barrier per step).
 * on = 2: additionally the multicast images are float32[npix][4] (rgb + one unused float, 16-byte aligned) and a pixel
 * travels as ONE multimem.red.v4 -- a third of the packets, which is what the NVLink fabric is limited by when every
 * GPU receives the adds of all ranks.  The gradient table keeps its layout in both modes. */
int psdr_scene_set_output_multicast(psdr_scene *s, int on);
/* CTA shape of the term kernels (process-wide; new).  Every term kernel exists in two shapes: large CTAs (one or two
 * per SM) with block barriers that keep the warps of an SM in the same stretch of code, and 128-thread CTAs for launches
 * too small to fill the large shape.  0 = choose by launch size (default), 1 = always 128 threads, 2 = always large.
 * Results do not depend on the shape (a lane's value is a function of its global index and the seed). */
int psdr_set_cta_policy(int policy);
/* Lane order of the primary- and secondary-edge kernels (process-wide; new).  Sample i of Integrator::render_primary_edges
 * (src/integrator/integrator.cpp:144-189) is a function of (seed, i) alone; its first draw selects the edge and the point on
 * it (src/sensor/perspective.cpp:117-141).  bins >= 2 (default 512, at most 2048): launches of 32768 lanes or more bucket
 * their lanes by that draw first (one counting-sort pass on the GPU) so that the rays of a warp start on the same stretch
 * of the same edge; 0 = lane order.  PathTracer::render_secondary_edges (src/integrator/path.cpp:268-302) likewise, by the
 * sample dimension that Scene::sample_boundary_segment_direct uses for the edge (after the guiding distribution's warp).
 * Values do not depend on it, only the order of the atomic adds into the image. */
int psdr_set_edge_sort(int bins);
/* New.  on != 0: the analytic re-intersection of the primary hit in renderD (src/scene/scene.cpp:772-801,
 * include/psdr/utils.h:82-93) takes its reciprocal with rcp.approx.ftz.f32, the instruction Dr.Jit emits for rcp()
 * (drjit-core cuda_eval.cpp:638-640), instead of the correctly rounded 1/x.  Everything else stays IEEE.  Reproduces the
 * reference's self-shadowing statistics on faces lit at grazing angles (DESIGN.md "parity"); off by default. */
int psdr_scene_set_reference_arithmetic(psdr_scene *s, int on);
/* -1 = automatic (BVH2 above 64 triangles), 0 = brute force, 1 = BVH2.  Replaces the OptiX GAS build of
 * Scene_OptiX::configure (src/scene/scene_optix.cpp:254-333): the BVH topology is built once (host, binned SAH) and every
 * later configure() refits boxes and leaf blocks on the GPU; calling this function forces a fresh topology build. */
int psdr_scene_set_accel(psdr_scene *s, int mode);

/* Scene.add_BSDF(DiffuseBSDF([r,g,b]), name, twoSide) -- src/psdr.cpp:401, src/scene/scene.cpp:148-247.
 * Returns the BSDF index (>= 0) or -1. */
int psdr_scene_add_bsdf_diffuse(psdr_scene *s, const char *id, const float reflectance[3], int two_side);

/* Scene.add_BSDF(MicrofacetBSDF([spec], [diff], roughness), name, twoSide) -- src/psdr.cpp:298-304,
 * include/psdr/bsdf/microfacet.h:12 (argument order: specular, diffuse, roughness), src/bsdf/microfacet.cpp. */
int psdr_scene_add_bsdf_microfacet(psdr_scene *s, const char *id, const float specular[3], const float diffuse[3], float roughness,
                                   int two_side);

/* Scene.add_BSDF(RoughConductorBSDF(alpha, eta, k), name, twoSide) -- src/psdr.cpp:286-293,
 * include/psdr/bsdf/roughconductor.h:10-17 (isotropic form: alpha_u = alpha_v = alpha; specular_reflectance defaults to 1),
 * src/bsdf/roughconductor.cpp:38-122, conductor Fresnel include/psdr/utils.h:167-183.  eta, k and specular_reflectance
 * are per-channel constants (1x1 Bitmap3fD) -- parameters PSDR_BSDF_ETA / PSDR_BSDF_K / PSDR_BSDF_SPECULAR; alpha is
 * PSDR_BSDF_ROUGHNESS.  The anisotropic constructors (alpha_u != alpha_v) are not supported. */
int psdr_scene_add_bsdf_roughconductor(psdr_scene *s, const char *id, float alpha, const float eta[3], const float k[3], const float specular[3],
                                       int two_side);

/* RoughDielectricBSDF(alpha, intIOR, extIOR) -- include/psdr/bsdf/roughdielectric.h:10-29, src/bsdf/roughdielectric.cpp:37-236
 * (GGX reflection + refraction, Fresnel include/psdr/utils.h:185-215); the reference creates it from scene files only
 * (src/scene/scene_loader.cpp:346-360; src/psdr.cpp:295 binds the class without a constructor).  alpha is PSDR_BSDF_ROUGHNESS
 * (constant or Bitmap1fD through slot PSDR_TEX_ROUGHNESS); the indices of refraction are fixed at creation. */
int psdr_scene_add_bsdf_roughdielectric(psdr_scene *s, const char *id, float alpha, float int_ior, float ext_ior, int two_side);

/* Scene.add_BSDF(MicrofacetBSDFPerVertex(specular, diffuse, roughness), name, twoSide) -- src/psdr.cpp:306-310,
 * src/bsdf/microfacet_pv.cpp: the three parameters are arrays over the VERTICES of the mesh the BSDF is attached to
 * (specular / diffuse: n_vertices*3 floats, roughness: n_vertices floats), interpolated with the hit's barycentrics
 * (microfacet_pv.cpp:146-160).  Afterwards PSDR_BSDF_PERVERTEX takes / returns all three as one 7*n_vertices table. */
int psdr_scene_add_bsdf_microfacet_pervertex(psdr_scene *s, const char *id, const float *specular, const float *diffuse, const float *roughness,
                                             int n_vertices, int two_side);

/* Scene.add_normalmap_BSDF(NormalMapBSDF, MicrofacetBSDF, name, twoSide) -- src/psdr.cpp:273-277,402, src/scene/scene.cpp:128-145,
 * src/bsdf/normalmap.cpp; scene files may nest a Diffuse, RoughConductor, RoughDielectric or Microfacet BSDF
 * (src/scene/scene_loader.cpp:372-424).  The wrapped BSDF is created first: psdr_scene_begin_nested_bsdf() makes the NEXT
 * psdr_scene_add_bsdf_* call create an un-numbered record and return its handle (>= PSDR_NESTED_BSDF_BASE), which is passed
 * here as `nested` and accepted as `index` by psdr_scene_set_param / set_tangent / set_bsdf_texture_slot.  normal: the
 * constant normal map (Scene.add_BSDF(NormalMapBSDF) uses (.499999, .499999, 1), scene.cpp:221); a bitmap goes through slot
 * PSDR_TEX_REFLECTANCE + PSDR_BSDF_REFLECTANCE of the NormalMap's own index. */
#define PSDR_NESTED_BSDF_BASE (1 << 20)
int psdr_scene_begin_nested_bsdf(psdr_scene *s);
int psdr_scene_add_bsdf_normalmap(psdr_scene *s, const char *id, const float normal[3], int nested, int two_side);

/* DiffuseBSDF.reflectance / MicrofacetBSDF.diffuseReflectance = Bitmap3fD(w, h, data) with more than one texel
 * (src/psdr.cpp:209-219, src/core/bitmap.cpp:46-131): declares the texture resolution of BSDF `index`; afterwards
 * PSDR_BSDF_REFLECTANCE takes / returns 3*w*h floats (rgb interleaved, pixel = y*w + x).  w = h = 1 switches back
 * to the constant. */
int psdr_scene_set_bsdf_texture(psdr_scene *s, int index, int w, int h);
/* The same for any of the three bitmaps of a BSDF: slot PSDR_TEX_REFLECTANCE (Bitmap3fD reflectance / diffuseReflectance),
 * PSDR_TEX_SPECULAR (Bitmap3fD specularReflectance) or PSDR_TEX_ROUGHNESS (Bitmap1fD roughness) -- MicrofacetBSDF(Bitmap3fD,
 * Bitmap3fD, Bitmap1fD), src/psdr.cpp:301, include/psdr/bsdf/microfacet.h:17,33-35.  Afterwards PSDR_BSDF_SPECULAR takes /
 * returns 3*w*h floats and PSDR_BSDF_ROUGHNESS w*h floats. */
int psdr_scene_set_bsdf_texture_slot(psdr_scene *s, int index, int slot, int w, int h);

/* Scene.add_Mesh(mesh, bsdf_id, emitter) with mesh = Mesh.load_raw(v, f, uv, f_uv) -- src/psdr.cpp:399-400,
 * src/scene/scene.cpp:249-309, src/shape/mesh.cpp:74-162.  v: nv*3 floats (object space), f: nf*3 ints,
 * uv: nuv*2 floats or NULL, fuv: nf*3 ints or NULL, to_world: 16 floats row-major or NULL (identity),
 * radiance: 3 floats (makes the mesh an AreaLight, src/emitter/area.cpp) or NULL.
 * Returns the mesh index (>= 0) or -1 ("Unknown BSDF id: ..."). */
int psdr_scene_add_mesh(psdr_scene *s, const float *v, int nv, const int *f, int nf, const float *uv, int nuv, const int *fuv,
                        const float *to_world, const char *bsdf_id, const float *radiance, int use_face_normals, int enable_edges);

/* Scene.add_EnvironmentMap(envmap) with envmap.radiance = Bitmap3fD(w, h, data), envmap.scale, envmap.to_world
 * -- src/psdr.cpp:349-355,397-398, src/scene/scene.cpp:85-105, src/emitter/envmap.cpp.  radiance: h*w*3 floats
 * (row-major, pixel = y*w + x).  configure() appends the 12-triangle bounding mesh (scene.cpp:435-485).
 * Returns the emitter index or -1 ("A scene is only allowed to have one envmap!"). */
int psdr_scene_add_envmap(psdr_scene *s, const float *radiance, int w, int h, const float *to_world, float scale);

/* Scene.add_Sensor(PerspectiveCamera(fov, near, far)) with sensor.to_world -- src/psdr.cpp:365-375,396 */
int psdr_scene_add_perspective(psdr_scene *s, float fov_x, float near_clip, float far_clip, const float *to_world);
/* Scene.add_Sensor(PerspectiveCamera(fx, fy, cx, cy, near, far)) -- src/psdr.cpp:366, include/psdr/sensor/perspective.h:11-12,
 * src/sensor/perspective.cpp:15-20, include/psdr/core/transform.h:63-71: pinhole intrinsics in units of the image size
 * (focal lengths fx, fy and principal point cx, cy as fractions of width / height). */
int psdr_scene_add_perspective_intrinsic(psdr_scene *s, float fx, float fy, float cx, float cy, float near_clip, float far_clip, const float *to_world);
/* Scene.add_Sensor(OrthographicCamera(near, far)) -- src/psdr.cpp:375-383, src/sensor/orthographic.cpp: rays leave the
 * sample's point on the near plane along the camera's +z; the view volume is 2 x 2/aspect camera units (the sensor transform
 * may not scale, sensor.cpp:12-13).  Everything else -- sample_direct, the primary-edge list -- is the perspective camera's
 * code in the reference (orthographic.cpp:46-104,134-175) and here. */
int psdr_scene_add_orthographic(psdr_scene *s, float near_clip, float far_clip, const float *to_world);

/* Writes through Scene.param_map[...] (README.md:87-90): new primal value of a parameter ... */
int psdr_scene_set_param(psdr_scene *s, int kind, int index, const float *value, int n);
/* ... and its forward-mode tangent d(param)/dP (what drjit.set_grad + forward_to propagate in the
 * reference, README.md:102-104).  Tangents persist until cleared. */
int psdr_scene_set_tangent(psdr_scene *s, int kind, int index, const float *tangent, int n);
int psdr_scene_clear_tangents(psdr_scene *s);

/* Scene.configure(active_sensor=[...]) -- src/psdr.cpp:409, src/scene/scene.cpp:311-601 */
int psdr_scene_configure(psdr_scene *s, const int *active_sensors, int n_active);
double psdr_scene_last_configure_ms(psdr_scene *s);

/* New (measurement): with timing on, every render call brackets each of its kernels with CUDA events
 * on the launch stream (the reference only prints std::chrono wall time, integrator.cpp:14,40-45).
 * psdr_scene_kernel_ms waits for the events of the LAST render call and returns the device time of
 * the kernel of `term` (one of PSDR_TERM_INTERIOR / _PRIMARY_EDGES / _SECONDARY_EDGES), or -1. */
int psdr_scene_enable_timing(psdr_scene *s, int on);
double psdr_scene_kernel_ms(psdr_scene *s, int term);

int psdr_scene_query(psdr_scene *s, int what, int index);
/* Mesh.edge_indices() -- src/psdr.cpp:338: out = 4 rows (v0, v1, face0, face1) of n_edges ints */
int psdr_scene_mesh_edges(psdr_scene *s, int mesh, int *out);

/* PathTracer(max_depth).renderC(scene, sensor_id, seed, batch_pix) -- src/psdr.cpp:419-420,431-434,
 * src/integrator/integrator.cpp:12-48.  seed = -1 continues the sampler streams.  pix_id (device, may be
 * NULL) = batch_pix pixel list of length npix; img (device) = float32[npix or W*H][3]. */
int psdr_render_c(psdr_scene *s, int sensor, int max_depth, long long seed, int hide_emitters, const int *pix_id, int npix,
                  float *img, void *cuda_stream);

/* PathTracer.renderD(...) followed by drjit.forward_to(img) -- src/integrator/integrator.cpp:51-100,179-198,
 * src/integrator/path.cpp:172-294, README.md:96-104: primal image and the forward-mode derivative image
 * for the tangents set with psdr_scene_set_tangent.  terms = PSDR_TERM_* mask.  reference_scaling != 0
 * reproduces the reference binary's output, whose interior and secondary-edge tangents come out
 * exactly 2x the finite-difference-correct value ("derivative scaling").  dimg = NULL: only
 * the primal image of renderD (the zero-primal edge kernels are skipped, their sampler streams still
 * advance) -- the forward half of an autograd step whose backward half is psdr_render_vjp. */
int psdr_render_d(psdr_scene *s, int sensor, int max_depth, long long seed, int hide_emitters, int terms, int reference_scaling,
                  const int *pix_id, int npix, float *img, float *dimg, void *cuda_stream);

/* Reverse mode of renderD -- what drjit.backward(loss(img)) computes in the reference (README.md:96-104,
 * the AD graph recorded by src/integrator/integrator.cpp:51-100 and src/integrator/path.cpp): given the
 * cotangent image d_img = d(loss)/d(img) (device, float32[npix][3]) the adjoint kernels replay the paths of
 * the forward call (same sensor / max_depth / seed / terms) and accumulate d(loss)/d(parameter) for EVERY
 * parameter reachable through Scene.param_map.  Synchronises the stream (the gradients are returned on
 * the host).  The sampler streams are left exactly as they were before the call.  reference_scaling as in
 * psdr_render_d. */
int psdr_render_vjp(psdr_scene *s, int sensor, int max_depth, long long seed, int hide_emitters, int terms, int reference_scaling,
                    const int *pix_id, int npix, const float *d_img, void *cuda_stream);
/* The same in two halves, for multi-GPU runs ("one ncclAllReduce over the flat parameter-gradient
 * buffer"): psdr_render_vjp_device launches the adjoint kernels into a CALLER-owned device table of
 * psdr_grad_table_size(s, sensor) floats (zeroed by the call; asynchronous, no synchronisation); the caller
 * all-reduces the table over ranks on the same stream; psdr_scene_backprop_table copies it to the host and runs
 * the host reverse chain of configure() (drjit.backward through Scene::configure in the reference). */
int psdr_grad_table_size(psdr_scene *s, int sensor);
int psdr_render_vjp_device(psdr_scene *s, int sensor, int max_depth, long long seed, int hide_emitters, int terms, int reference_scaling,
                           const int *pix_id, int npix, const float *d_img, float *grad_table, int n_table, void *cuda_stream);
int psdr_scene_backprop_table(psdr_scene *s, int sensor, const float *grad_table, int n_table, void *cuda_stream);
/* Gradient of parameter (kind, index) from the last psdr_render_vjp, same shapes as psdr_scene_set_param;
 * host buffer.  drjit.grad(param) in the reference. */
int psdr_scene_get_grad(psdr_scene *s, int kind, int index, float *out, int n);
/* Sampler streams (reference Sampler state, src/core/sampler.cpp): state[2k] = seed, state[2k+1] = draws
 * consumed per lane of sampler k (0 interior, 1 primary edges, 2 secondary edges); seed < 0 = not seeded. */
int psdr_scene_get_sampler_state(psdr_scene *s, long long state[6]);
int psdr_scene_set_sampler_state(psdr_scene *s, const long long state[6]);

/* PathTracer.preprocess_secondary_edges(scene, sensor_id, [rx, ry, rz, n], nrounds, seed) -- src/psdr.cpp:431-434,
 * src/integrator/path.cpp:130-168: builds the guiding grid (rx*ry*rz cells, n samples per cell and round) of the
 * secondary-edge sampler for `sensor` and enables it.  The grid belongs to the scene handle and survives
 * configure(); psdr_scene_set_guiding switches its use on/off (the reference keeps it in the integrator object).
 * Synchronises the stream. */
int psdr_preprocess_secondary_edges(psdr_scene *s, int sensor, const int reso[4], int nrounds, long long seed, void *cuda_stream);
int psdr_scene_set_guiding(psdr_scene *s, int sensor, int enabled);
/* Mass per cell of the last pre-pass (host buffer, rx*ry*rz floats). */
int psdr_scene_guiding_mass(psdr_scene *s, int sensor, float *out, int n);

/* Same calls with HOST output buffers (pageable or pinned): device work + device->host copies,
 * returns after the buffers are filled. */
int psdr_render_c_host(psdr_scene *s, int sensor, int max_depth, long long seed, int hide_emitters, const int *pix_id_host, int npix,
                       float *img_host);
int psdr_render_d_host(psdr_scene *s, int sensor, int max_depth, long long seed, int hide_emitters, int terms, int reference_scaling,
                       const int *pix_id_host, int npix, float *img_host, float *dimg_host);

/* FieldExtractionIntegrator taps (src/integrator/field.cpp:47-121) at the scene's spp: per lane 14 floats
 * (mesh id + 1, triangle id, position xyz, distance, geometric normal xyz, shading normal xyz, uv). */
int psdr_render_aov(psdr_scene *s, int sensor, long long seed, float *out, void *cuda_stream);
/* FieldExtractionIntegrator.renderD followed by drjit.forward_to (src/integrator/field.cpp:47-121 through
 * Integrator::renderD, src/integrator/integrator.cpp:51-100,179-198), in two parts.
 * psdr_render_aov_d: the interior part -- the same 14 taps from the AD instantiation of the primary hit (re-intersected
 * analytically, src/scene/scene.cpp:772-801) in `out`, and their forward-mode tangents for the tangents set with
 * psdr_scene_set_tangent in `dout` (both float32[lanes][14]).
 * psdr_render_field_edges: the primary-edge part at the scene's sppe -- dimg (device, float32[W*H][3], zeroed by the call)
 * receives jump-of-the-field x normal velocity of the sampled pixel-space edges.  field: 0 segmentation (mesh index),
 * 1 silhouette, 2 position, 3 depth, 4 geoNormal, 5 shNormal, 6 uv; object >= 0 keeps that mesh only ("<field> <id>"). */
int psdr_render_aov_d(psdr_scene *s, int sensor, long long seed, float *out, float *dout, void *cuda_stream);
int psdr_render_field_edges(psdr_scene *s, int sensor, long long seed, int field, int object, float *dimg, void *cuda_stream);
/* Reverse mode of FieldExtractionIntegrator::renderD: gradients of <d_img, field image> with respect to every parameter
 * (read with psdr_scene_get_grad) -- the interior part (the adjoint of the analytically re-intersected primary hit: position,
 * depth, normals, uv; zero for silhouette / segmentation) and the primary-edge part (the jump of the field across the
 * pixel-space edges).  d_img: width*height*3 floats on the device. */
int psdr_render_field_vjp(psdr_scene *s, int sensor, long long seed, int field, int object, int terms, int reference_scaling,
                          const float *d_img, void *cuda_stream);

/* Sampler.seed / next_1d (src/psdr.cpp:181-185): out[ndraws][n], host memory, computed on the host. */
int psdr_sampler_draws(long long seed, int n, int ndraws, float *out);

#ifdef __cplusplus
}
#endif
#endif /* PSDR_B200_H */
