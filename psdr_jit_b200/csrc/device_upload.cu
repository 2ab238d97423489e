// Packs the configured host scene into the float4 tables of dscene.h and uploads them.
// All tables of one scene live in ONE device allocation (a few KB for a Cornell box, ~1 MB for a
// 5k-triangle mesh), refreshed by a single cudaMemcpyAsync per configure() from pinned staging.
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "scene.h"

namespace psdr {

struct DeviceBuffers {
    void *dev = nullptr;
    void *host = nullptr;     // pinned staging
    size_t capacity = 0;
    // the environment map's big tables (texels, texel tangents, cell pmf / cdf: 22 MB for a 1024 x 512 map) live in their
    // own allocation and are re-uploaded only when their versions change (HEnvmap::data_version / ddata_version)
    void *big = nullptr;
    size_t big_capacity = 0, big_off[5] = {0, 0, 0, 0, 0};
    std::vector<int> env_lut;           // bucket table of the envmap cell CDF (rebuilt with the cell table)
    int env_lut_n = 0;
    unsigned big_data_version = 0, big_ddata_version = 0;
    // BVH (scenes above 64 triangles): the TOPOLOGY (node links, leaf slots -> triangle ids) is built on the host (binned
    // SAH) when the set of meshes / face counts changes and stays on the device; every configure() after that only
    // refits the boxes and refreshes the leaf triangle blocks ON THE GPU from the freshly uploaded triangle table
    // (bvh_refit_kernel below) -- the reference rebuilds its OptiX GAS on every configure (scene_optix.cpp:254-333).
    void *bvh = nullptr;
    size_t bvh_capacity = 0;
    std::vector<int> bvh_key;          // face count per mesh the topology was built for
    int n_nodes2 = 0, n_slots = 0, n_leaves = 0;
    float bvh_pad = 0.f;
    size_t off_nodes2 = 0, off_leaf_tri = 0, off_slot_tri = 0, off_leaf_list = 0, off_ready = 0;
    ~DeviceBuffers() {
        if (dev) cudaFree(dev);
        if (host) cudaFreeHost(host);
        if (big) cudaFree(big);
        if (bvh) cudaFree(bvh);
    }
};

// ---- GPU refit of the BVH ----------------------------------------------------------------------------------------
// One thread per leaf: copy the leaf's triangles from the triangle table into its contiguous slots, bound them, write
// the box into the parent's child slot and walk up; the second thread to arrive at a node (atomic counter) unions the
// two child boxes and continues, so every box is final before it is read (Karras-style bottom-up pass).
struct BvhLeaf {
    int node, slot, first, count;     // parent node, which child of it (0 / 1), first leaf slot, triangles
};
__global__ void bvh_refit_kernel(DBvhNode2 *nodes, float4 *leaf_tri, const int *slot_tri, const BvhLeaf *leaves, int n_leaves,
                                 const float4 *geo, int *ready, float pad) {
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_leaves) return;
    const BvhLeaf L = leaves[li];
    float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int k = 0; k < L.count; ++k) {
        const int id = slot_tri[L.first + k];
        const float4 a = geo[3 * id], b = geo[3 * id + 1], c = geo[3 * id + 2];
        const float p0[3] = {a.x, a.y, a.z}, e1[3] = {a.w, b.x, b.y}, e2[3] = {b.z, b.w, c.x};
        leaf_tri[3 * (L.first + k)] = make_float4(a.x, a.y, a.z, __int_as_float(id));
        leaf_tri[3 * (L.first + k) + 1] = make_float4(a.w, b.x, b.y, 0.f);
        leaf_tri[3 * (L.first + k) + 2] = make_float4(b.z, b.w, c.x, 0.f);
        for (int ax = 0; ax < 3; ++ax) {
            const float v1 = p0[ax] + e1[ax], v2 = p0[ax] + e2[ax];
            lo[ax] = fminf(lo[ax], fminf(p0[ax], fminf(v1, v2)));
            hi[ax] = fmaxf(hi[ax], fmaxf(p0[ax], fmaxf(v1, v2)));
        }
    }
    for (int ax = 0; ax < 3; ++ax) { lo[ax] -= pad; hi[ax] += pad; }
    int node = L.node, slot = L.slot;
    while (node >= 0) {
        float *f = nodes[node].f + 6 * slot;
        for (int ax = 0; ax < 3; ++ax) { f[ax] = lo[ax]; f[3 + ax] = hi[ax]; }
        __threadfence();
        if (atomicAdd(ready + node, 1) == 0) return;      // the sibling subtree is not finished: its thread continues
        const float *g = nodes[node].f + 6 * (1 - slot);  // written before the sibling's fence + atomic
        for (int ax = 0; ax < 3; ++ax) {
            lo[ax] = fminf(lo[ax], ((volatile const float *) g)[ax]);
            hi[ax] = fmaxf(hi[ax], ((volatile const float *) g)[3 + ax]);
        }
        slot = nodes[node].pslot;
        node = nodes[node].parent;
    }
}

static void check(cudaError_t e, const char *what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e));
}

Scene::~Scene() { delete dev; }

namespace {
struct Packer {
    std::vector<unsigned char> bytes;
    template <class T> size_t add(const std::vector<T> &v) {
        size_t off = (bytes.size() + 255) / 256 * 256;
        bytes.resize(off + std::max<size_t>(v.size() * sizeof(T), 16));
        if (!v.empty()) std::memcpy(bytes.data() + off, v.data(), v.size() * sizeof(T));
        return off;
    }
};
// Bucket table of an unnormalised fp32 CDF for sample_reuse_lut (device_path.cuh): n = the power of two >= size (at most
// 2^20) buckets; entry b = first index i with cmf[i] >= b * sum / n (clipped to size - 1), entry n = size - 1.  Appends
// n + 1 ints to `out`, returns n (0 for distributions too small to bother).
inline int build_cdf_lut(const std::vector<float> &cmf, float sum, std::vector<int> &out) {
    const int size = (int) cmf.size();
    if (size < 64 || !(sum > 0.f)) return 0;
    int n = 64;
    while (n < size && n < (1 << 20)) n <<= 1;
    int i = 0;
    for (int b = 0; b < n; ++b) {
        const double v = (double) b * (double) sum / (double) n;
        while (i < size - 1 && (double) cmf[i] < v) ++i;
        out.push_back(i);
    }
    out.push_back(size - 1);
    return n;
}
inline float as_float(int i) {
    float f;
    std::memcpy(&f, &i, 4);
    return f;
}
}  // namespace

void upload_scene(Scene &sc) {
    check(cudaSetDevice(sc.device), "cudaSetDevice");
    std::vector<float4> geo, shade, dgeo, dshade, sec, pe_a, pe_da, pe_b;
    std::vector<float2> uv;
    std::vector<int> face_idx;
    std::vector<DMesh> dmeshes;
    std::vector<DEmitter> demit;
    std::vector<DBsdf> dbsdf;
    std::vector<float> face_pmf, face_cmf, em_pmf, em_cmf, sec_pmf, sec_cmf, pe_pmf, pe_cmf;
    std::vector<HTri> all_tris;

    for (size_t mi = 0; mi < sc.meshes.size(); ++mi) {
        const HMesh &m = sc.meshes[mi];
        DMesh dm;
        dm.bsdf = m.bsdf;
        dm.emitter = m.emitter;
        dm.flags = (m.use_face_normals ? 1 : 0) | (m.has_uv ? 2 : 0);
        dm.face_offset = m.face_offset;
        dmeshes.push_back(dm);
        for (size_t i = 0; i < m.tris.size(); ++i) {
            const HTri &t = m.tris[i];
            all_tris.push_back(t);
            geo.push_back(make_float4(t.p0.x.v, t.p0.y.v, t.p0.z.v, t.e1.x.v));
            geo.push_back(make_float4(t.e1.y.v, t.e1.z.v, t.e2.x.v, t.e2.y.v));
            geo.push_back(make_float4(t.e2.z.v, t.area.v, as_float((int) mi), 0.f));
            dgeo.push_back(make_float4(t.p0.x.d, t.p0.y.d, t.p0.z.d, t.e1.x.d));
            dgeo.push_back(make_float4(t.e1.y.d, t.e1.z.d, t.e2.x.d, t.e2.y.d));
            dgeo.push_back(make_float4(t.e2.z.d, t.area.d, 0.f, 0.f));
            shade.push_back(make_float4(t.n0.x.v, t.n0.y.v, t.n0.z.v, t.n1.x.v));
            shade.push_back(make_float4(t.n1.y.v, t.n1.z.v, t.n2.x.v, t.n2.y.v));
            shade.push_back(make_float4(t.n2.z.v, t.fn.x.v, t.fn.y.v, t.fn.z.v));
            dshade.push_back(make_float4(t.n0.x.d, t.n0.y.d, t.n0.z.d, t.n1.x.d));
            dshade.push_back(make_float4(t.n1.y.d, t.n1.z.d, t.n2.x.d, t.n2.y.d));
            dshade.push_back(make_float4(t.n2.z.d, t.fn.x.d, t.fn.y.d, t.fn.z.d));
            for (int k = 0; k < 3; ++k) {
                const V2f c = m.has_uv ? m.uv[m.fuv[3 * i + k]] : V2f(0.f, 0.f);
                uv.push_back(make_float2(c.x, c.y));
                face_idx.push_back(3 * i + k < m.f.size() ? m.f[3 * i + k] : 0);   // the envmap's bounding mesh has no index list of its own
            }
        }
    }
    // numbered BSDFs first, then the records NormalMaps wrap (DBsdf::nested indexes the combined table)
    std::vector<const HBsdf *> all_bsdfs;
    for (const HBsdf &b : sc.bsdfs) all_bsdfs.push_back(&b);
    for (const HBsdf &b : sc.nested_bsdfs) all_bsdfs.push_back(&b);
    bool any_pv = false;
    for (const HBsdf *bp : all_bsdfs) {
        const HBsdf &b = *bp;
        DBsdf d{};
        d.refl[0] = b.reflectance.x.v; d.refl[1] = b.reflectance.y.v; d.refl[2] = b.reflectance.z.v;
        d.d_refl[0] = b.reflectance.x.d; d.d_refl[1] = b.reflectance.y.d; d.d_refl[2] = b.reflectance.z.d;
        d.type = b.type;
        d.two_side = b.two_side ? 1 : 0;
        d.spec[0] = b.specular.x.v; d.spec[1] = b.specular.y.v; d.spec[2] = b.specular.z.v;
        d.d_spec[0] = b.specular.x.d; d.d_spec[1] = b.specular.y.d; d.d_spec[2] = b.specular.z.d;
        d.rough = b.roughness.v;
        d.d_rough = b.roughness.d;
        d.eta[0] = b.eta.x.v; d.eta[1] = b.eta.y.v; d.eta[2] = b.eta.z.v;
        d.d_eta[0] = b.eta.x.d; d.d_eta[1] = b.eta.y.d; d.d_eta[2] = b.eta.z.d;
        d.kk[0] = b.k.x.v; d.kk[1] = b.k.y.v; d.kk[2] = b.k.z.v;
        d.d_kk[0] = b.k.x.d; d.d_kk[1] = b.k.y.d; d.d_kk[2] = b.k.z.d;
        d.nested = b.nested >= 0 ? (int) sc.bsdfs.size() + b.nested : -1;
        d.pv_n = (int) b.pv.size() / 7;
        any_pv |= b.type == 4;
        dbsdf.push_back(d);
    }
    for (const HEmitter &e : sc.emitters) {
        if (e.type == 1) {   // EnvironmentMap: everything lives in DEnv
            DEmitter d{};
            d.type = 1;
            d.mesh = e.mesh;
            d.sampling_weight = e.sampling_weight;
            d.emitter_pmf = e.raw_weight;
            demit.push_back(d);
            continue;
        }
        const HMesh &m = sc.meshes[e.mesh];
        DEmitter d{};
        d.radiance[0] = e.radiance.x.v; d.radiance[1] = e.radiance.y.v; d.radiance[2] = e.radiance.z.v;
        d.d_radiance[0] = e.radiance.x.d; d.d_radiance[1] = e.radiance.y.d; d.d_radiance[2] = e.radiance.z.d;
        d.mesh = e.mesh;
        d.sampling_weight = e.sampling_weight;
        d.face_offset = m.face_offset;
        d.nfaces = (int) m.tris.size();
        d.inv_total_area = m.inv_total_area;
        d.distrb_offset = (int) face_pmf.size();
        d.face_sum = m.face_distrb.sum;
        d.emitter_pmf = e.raw_weight;
        face_pmf.insert(face_pmf.end(), m.face_distrb.pmf.begin(), m.face_distrb.pmf.end());
        face_cmf.insert(face_cmf.end(), m.face_distrb.cmf.begin(), m.face_distrb.cmf.end());
        demit.push_back(d);
    }
    if (!sc.emitters.empty()) { em_pmf = sc.emitter_distrb.pmf; em_cmf = sc.emitter_distrb.cmf; }
    for (const HSecEdge &s : sc.sec_edges) {
        sec.push_back(make_float4(s.p0.x.v, s.p0.y.v, s.p0.z.v, s.e1.x.v));
        sec.push_back(make_float4(s.e1.y.v, s.e1.z.v, s.p0.x.d, s.p0.y.d));
        sec.push_back(make_float4(s.p0.z.d, s.e1.x.d, s.e1.y.d, s.e1.z.d));
        sec.push_back(make_float4(s.n0.x, s.n0.y, s.n0.z, s.n1.x));
        sec.push_back(make_float4(s.n1.y, s.n1.z, s.p2.x, s.p2.y));
        sec.push_back(make_float4(s.p2.z, s.is_boundary ? 1.f : 0.f, 0.f, 0.f));
    }
    if (!sc.sec_edges.empty()) { sec_pmf = sc.sec_edge_distrb.pmf; sec_cmf = sc.sec_edge_distrb.cmf; }

    std::vector<size_t> cam_edge_first(sc.cameras.size());
    for (size_t ci = 0; ci < sc.cameras.size(); ++ci) {
        const HCamera &c = sc.cameras[ci];
        cam_edge_first[ci] = pe_a.size();
        for (const HPrimEdge &e : c.edges) {
            pe_a.push_back(make_float4(e.p0.x.v, e.p0.y.v, e.p1.x.v, e.p1.y.v));
            pe_da.push_back(make_float4(e.p0.x.d, e.p0.y.d, e.p1.x.d, e.p1.y.d));
            pe_b.push_back(make_float4(e.normal.x, e.normal.y, e.length, 0.f));
        }
        if (!c.edges.empty()) {
            pe_pmf.insert(pe_pmf.end(), c.edge_distrb.pmf.begin(), c.edge_distrb.pmf.end());
            pe_cmf.insert(pe_cmf.end(), c.edge_distrb.cmf.begin(), c.edge_distrb.cmf.end());
        }
    }

    std::vector<float> g_pmf, g_cmf;
    std::vector<int> g_lut, s_lut;
    std::vector<size_t> cam_guide_first(sc.cameras.size(), 0), cam_lut_first(sc.cameras.size(), 0);
    std::vector<int> cam_lut_n(sc.cameras.size(), 0);
    for (size_t ci = 0; ci < sc.cameras.size(); ++ci) {
        const HCamera &c = sc.cameras[ci];
        cam_guide_first[ci] = g_pmf.size();
        cam_lut_first[ci] = g_lut.size();
        if (c.guide_ready) {
            g_pmf.insert(g_pmf.end(), c.guide.pmf.begin(), c.guide.pmf.end());
            g_cmf.insert(g_cmf.end(), c.guide.cmf.begin(), c.guide.cmf.end());
            cam_lut_n[ci] = build_cdf_lut(c.guide.cmf, c.guide.sum, g_lut);
        }
    }
    const int sec_lut_n = sc.sec_edges.empty() ? 0 : build_cdf_lut(sc.sec_edge_distrb.cmf, sc.sec_edge_distrb.sum, s_lut);

    const int ntris = (int) all_tris.size();
    const bool use_bvh = ntris > kMaxBruteTris || (sc.force_bvh < 0 ? false : sc.force_bvh != 0);

    Packer pk;
    const size_t o_geo = pk.add(geo), o_shade = pk.add(shade), o_dgeo = pk.add(dgeo), o_dshade = pk.add(dshade), o_uv = pk.add(uv),
                 o_mesh = pk.add(dmeshes), o_emit = pk.add(demit), o_bsdf = pk.add(dbsdf), o_fp = pk.add(face_pmf), o_fc = pk.add(face_cmf),
                 o_ep = pk.add(em_pmf), o_ec = pk.add(em_cmf), o_sec = pk.add(sec), o_sp = pk.add(sec_pmf), o_scm = pk.add(sec_cmf),
                 o_pa = pk.add(pe_a), o_pda = pk.add(pe_da), o_pb = pk.add(pe_b), o_pp = pk.add(pe_pmf), o_pc = pk.add(pe_cmf),
                 o_gp = pk.add(g_pmf), o_gc = pk.add(g_cmf), o_gl = pk.add(g_lut), o_sl = pk.add(s_lut),
                 o_small_end = 0;
    (void) o_small_end;
    if (!any_pv) face_idx.clear();
    const size_t o_fi = pk.add(face_idx);
    std::vector<size_t> o_tex(3 * all_bsdfs.size(), 0), o_dtex(3 * all_bsdfs.size(), 0), o_pv(all_bsdfs.size(), 0), o_dpv(all_bsdfs.size(), 0);
    for (size_t i = 0; i < all_bsdfs.size(); ++i) {
        for (int k = 0; k < 3; ++k) if (all_bsdfs[i]->tex[k].w > 0) {
            o_tex[3 * i + k] = pk.add(all_bsdfs[i]->tex[k].data);
            o_dtex[3 * i + k] = pk.add(all_bsdfs[i]->tex[k].ddata);
        }
        if (!all_bsdfs[i]->pv.empty()) {
            o_pv[i] = pk.add(all_bsdfs[i]->pv);
            o_dpv[i] = pk.add(all_bsdfs[i]->d_pv);
        }
    }

    if (!sc.dev) sc.dev = new DeviceBuffers();
    DeviceBuffers &db = *sc.dev;
    if (pk.bytes.size() > db.capacity) {
        if (db.dev) cudaFree(db.dev);
        if (db.host) cudaFreeHost(db.host);
        db.capacity = pk.bytes.size() * 2;
        check(cudaMalloc(&db.dev, db.capacity), "cudaMalloc(scene tables)");
        check(cudaMallocHost(&db.host, db.capacity), "cudaMallocHost(scene staging)");
    }
    // the previous tables may still be read by kernels in flight on other streams
    check(cudaDeviceSynchronize(), "cudaDeviceSynchronize(before table refresh)");
    size_t big_uploaded = 0;
    if (sc.env.present) {
        const HEnvmap &e = sc.env;
        const std::vector<float> *tabs[4] = {&e.data, &e.ddata, &e.cell.pmf, &e.cell.cmf};
        size_t need = 0, off[5];
        for (int k = 0; k < 4; ++k) { off[k] = need; need += (std::max<size_t>(tabs[k]->size() * sizeof(float), 16) + 255) / 256 * 256; }
        // bucket table of the cell CDF (sample_reuse_lut): at most 2^20 + 1 ints
        int lut_cap = 64;
        while (lut_cap < e.cell.size && lut_cap < (1 << 20)) lut_cap <<= 1;
        off[4] = need;
        need += ((size_t) (lut_cap + 1) * sizeof(int) + 255) / 256 * 256;
        const bool realloc = need > db.big_capacity;
        if (realloc) {
            if (db.big) cudaFree(db.big);
            db.big_capacity = need;
            check(cudaMalloc(&db.big, db.big_capacity), "cudaMalloc(envmap tables)");
        }
        const bool data_changed = realloc || db.big_data_version != e.data_version || std::memcmp(off, db.big_off, sizeof(off)) != 0;
        const bool ddata_changed = data_changed || db.big_ddata_version != e.ddata_version;
        for (int k = 0; k < 4; ++k) {
            if (!(k == 1 ? ddata_changed : data_changed) || tabs[k]->empty()) continue;
            check(cudaMemcpy((unsigned char *) db.big + off[k], tabs[k]->data(), tabs[k]->size() * sizeof(float), cudaMemcpyHostToDevice), "cudaMemcpy(envmap tables)");
            big_uploaded += tabs[k]->size() * sizeof(float);
        }
        if (data_changed) {
            db.env_lut.clear();
            db.env_lut_n = build_cdf_lut(e.cell.cmf, e.cell.sum, db.env_lut);
            if (db.env_lut_n > 0) {
                check(cudaMemcpy((unsigned char *) db.big + off[4], db.env_lut.data(), db.env_lut.size() * sizeof(int), cudaMemcpyHostToDevice), "cudaMemcpy(envmap cell table)");
                big_uploaded += db.env_lut.size() * sizeof(int);
            }
        }
        std::memcpy(db.big_off, off, sizeof(off));
        db.big_data_version = e.data_version;
        db.big_ddata_version = e.ddata_version;
    }
    // texture pointers of the BSDF records can only be filled in once the allocation is known
    for (size_t i = 0; i < all_bsdfs.size(); ++i) {
        DBsdf *rec = reinterpret_cast<DBsdf *>(pk.bytes.data() + o_bsdf) + i;
        if (!all_bsdfs[i]->pv.empty()) {
            rec->pv = (const float *) ((const unsigned char *) db.dev + o_pv[i]);
            rec->d_pv = all_bsdfs[i]->d_pv.empty() ? nullptr : (const float *) ((const unsigned char *) db.dev + o_dpv[i]);
            rec->pv_goff = sc.pervertex_grad_offset((int) i);
        }
        for (int k = 0; k < 3; ++k) {
            const HBsdf::Tex &t = all_bsdfs[i]->tex[k];
            DTex &dt = rec->tex[k];
            dt = DTex{};
            dt.ch = HBsdf::tex_channels(k);
            dt.goff = sc.texture_grad_offset((int) i, k);   // relative to the end of the gradient table (nested records included)
            // cos / sin of the rotation on the host (the reference evaluates them per lookup on the device)
            const float cr = std::cos(t.rot.v), sr = std::sin(t.rot.v);
            dt.cr = cr; dt.sr = sr; dt.d_cr = -sr * t.rot.d; dt.d_sr = cr * t.rot.d;
            dt.scale = t.scale.v; dt.d_scale = t.scale.d;
            dt.tx = t.tx.v; dt.d_tx = t.tx.d; dt.ty = t.ty.v; dt.d_ty = t.ty.d;
            if (t.w > 0) {
                dt.w = t.w;
                dt.h = t.h;
                dt.data = (const float *) ((const unsigned char *) db.dev + o_tex[3 * i + k]);
                dt.ddata = t.ddata.empty() ? nullptr : (const float *) ((const unsigned char *) db.dev + o_dtex[3 * i + k]);
            }
        }
    }
    std::memcpy(db.host, pk.bytes.data(), pk.bytes.size());
    check(cudaMemcpy(db.dev, db.host, pk.bytes.size(), cudaMemcpyHostToDevice), "cudaMemcpy(scene tables)");
    sc.upload_bytes = pk.bytes.size() + big_uploaded;
    if (use_bvh) {
        std::vector<int> key;
        for (const HMesh &m : sc.meshes) key.push_back((int) m.tris.size());
        if (key != db.bvh_key || sc.bvh_rebuild) {
            // ---- topology: host binned-SAH build, converted to the wide-fetch layout
            std::vector<DBvhNode> nodes;
            std::vector<int> order;
            build_bvh(all_tris, nodes, order, 4);
            if (nodes.empty() || nodes[0].b < 0) throw std::runtime_error("BVH mode needs more than one leaf");
            std::vector<int> inner_index(nodes.size(), -1);
            int n_inner = 0;
            for (size_t i = 0; i < nodes.size(); ++i)
                if (nodes[i].b >= 0) inner_index[i] = n_inner++;
            std::vector<DBvhNode2> n2((size_t) n_inner);
            std::vector<BvhLeaf> leaves;
            for (size_t i = 0; i < nodes.size(); ++i) {
                if (nodes[i].b < 0) continue;
                DBvhNode2 &nd = n2[inner_index[i]];
                if (i == 0) { nd.parent = -1; nd.pslot = 0; }
                const int ch[2] = {nodes[i].a, nodes[i].b};
                for (int k = 0; k < 2; ++k) {
                    const DBvhNode &c = nodes[ch[k]];
                    for (int ax = 0; ax < 3; ++ax) { nd.f[6 * k + ax] = c.lo[ax]; nd.f[6 * k + 3 + ax] = c.hi[ax]; }
                    int link;
                    if (c.b < 0) {      // leaf: first slot = position in `order`, count = -b
                        if (-c.b > 7) throw std::runtime_error("BVH leaf too large");
                        link = ~((c.a << 3) | (-c.b));
                        leaves.push_back(BvhLeaf{inner_index[i], k, c.a, -c.b});
                    } else {
                        link = inner_index[ch[k]];
                        n2[link].parent = inner_index[i];
                        n2[link].pslot = k;
                    }
                    (k == 0 ? nd.c0 : nd.c1) = link;
                }
            }
            const float dx = nodes[0].hi[0] - nodes[0].lo[0], dy = nodes[0].hi[1] - nodes[0].lo[1], dz = nodes[0].hi[2] - nodes[0].lo[2];
            db.bvh_pad = 2e-5f * std::sqrt(dx * dx + dy * dy + dz * dz) + 1e-6f;     // as build_bvh pads (boxes above include it twice at most)
            db.n_nodes2 = n_inner;
            db.n_slots = (int) order.size();
            db.n_leaves = (int) leaves.size();
            auto align = [](size_t x) { return (x + 255) / 256 * 256; };
            db.off_nodes2 = 0;
            db.off_leaf_tri = align(sizeof(DBvhNode2) * n2.size());
            db.off_slot_tri = db.off_leaf_tri + align(sizeof(float4) * 3 * order.size());
            db.off_leaf_list = db.off_slot_tri + align(sizeof(int) * order.size());
            db.off_ready = db.off_leaf_list + align(sizeof(BvhLeaf) * leaves.size());
            const size_t need = db.off_ready + align(sizeof(int) * n2.size());
            if (need > db.bvh_capacity) {
                if (db.bvh) cudaFree(db.bvh);
                db.bvh_capacity = need;
                check(cudaMalloc(&db.bvh, need), "cudaMalloc(BVH)");
            }
            unsigned char *bv = (unsigned char *) db.bvh;
            check(cudaMemcpy(bv + db.off_nodes2, n2.data(), sizeof(DBvhNode2) * n2.size(), cudaMemcpyHostToDevice), "cudaMemcpy(BVH nodes)");
            check(cudaMemcpy(bv + db.off_slot_tri, order.data(), sizeof(int) * order.size(), cudaMemcpyHostToDevice), "cudaMemcpy(BVH slots)");
            check(cudaMemcpy(bv + db.off_leaf_list, leaves.data(), sizeof(BvhLeaf) * leaves.size(), cudaMemcpyHostToDevice), "cudaMemcpy(BVH leaves)");
            db.bvh_key = key;
            sc.bvh_rebuild = false;
            sc.upload_bytes += sizeof(DBvhNode2) * n2.size() + sizeof(int) * order.size() + sizeof(BvhLeaf) * leaves.size();
            sc.bvh_builds++;
        }
        // ---- boxes + leaf triangle blocks: always refit on the GPU from the triangle table just uploaded
        unsigned char *bv = (unsigned char *) db.bvh;
        check(cudaMemsetAsync(bv + db.off_ready, 0, sizeof(int) * db.n_nodes2, 0), "memset(BVH counters)");
        bvh_refit_kernel<<<(db.n_leaves + 127) / 128, 128>>>((DBvhNode2 *) (bv + db.off_nodes2), (float4 *) (bv + db.off_leaf_tri),
                                                             (const int *) (bv + db.off_slot_tri), (const BvhLeaf *) (bv + db.off_leaf_list), db.n_leaves,
                                                             (const float4 *) ((const unsigned char *) db.dev + o_geo), (int *) (bv + db.off_ready), db.bvh_pad);
        check(cudaGetLastError(), "bvh_refit_kernel");
        check(cudaDeviceSynchronize(), "bvh_refit_kernel");
        sc.bvh_refits++;
    }
    const unsigned char *base = (const unsigned char *) db.dev;

    DScene &d = sc.dscene;
    d = DScene{};
    d.width = sc.width; d.height = sc.height; d.spp = sc.spp; d.sppe = sc.sppe; d.sppse = sc.sppse;
    d.n_tris = ntris;
    d.n_meshes = (int) sc.meshes.size();
    d.n_emitters = (int) sc.emitters.size();
    d.n_bsdfs = (int) sc.bsdfs.size();
    d.n_sec_edges = (int) sc.sec_edges.size();
    d.n_nodes = use_bvh ? db.n_nodes2 : 0;
    d.use_bvh = use_bvh ? 1 : 0;
    d.ref_rcp = sc.ref_rcp ? 1 : 0;
    d.full_features = sc.env.present ? 1 : 0;
    for (const HBsdf *b : all_bsdfs) {
        d.ext_features |= (b->type >= 2 || b->tex[0].w > 0 || b->tex[1].w > 0 || b->tex[2].w > 0) ? 1 : 0;
        d.full_features |= b->type != 0 ? 1 : 0;
    }
    d.full_features |= d.ext_features;
    d.geo = (const float4 *) (base + o_geo);
    d.shade = (const float4 *) (base + o_shade);
    d.dgeo = (const float4 *) (base + o_dgeo);
    d.dshade = (const float4 *) (base + o_dshade);
    d.uv = (const float2 *) (base + o_uv);
    d.face_idx = face_idx.empty() ? nullptr : (const int *) (base + o_fi);
    d.meshes = (const DMesh *) (base + o_mesh);
    d.emitters = (const DEmitter *) (base + o_emit);
    d.bsdfs = (const DBsdf *) (base + o_bsdf);
    d.face_pmf = (const float *) (base + o_fp);
    d.face_cmf = (const float *) (base + o_fc);
    d.emitter_pmf = (const float *) (base + o_ep);
    d.emitter_cmf = (const float *) (base + o_ec);
    d.emitter_sum = sc.emitters.empty() ? 0.f : sc.emitter_distrb.sum;
    d.sec_edges = (const float4 *) (base + o_sec);
    d.sec_pmf = (const float *) (base + o_sp);
    d.sec_cmf = (const float *) (base + o_scm);
    d.sec_sum = sc.sec_edges.empty() ? 0.f : sc.sec_edge_distrb.sum;
    d.sec_lut = (const int *) (base + o_sl);
    d.sec_lut_n = sec_lut_n;
    d.nodes2 = use_bvh ? (const DBvhNode2 *) ((const unsigned char *) db.bvh + db.off_nodes2) : nullptr;
    d.leaf_tri = use_bvh ? (const float4 *) ((const unsigned char *) db.bvh + db.off_leaf_tri) : nullptr;
    if (!use_bvh) {
        // (even, odd) triangle of a pair -> (low, high) half of each 64-bit word
        float *w = reinterpret_cast<float *>(d.bg_pair);
        std::memset(d.bg_pair, 0, sizeof(d.bg_pair));
        for (int i = 0; i < ntris; ++i) {
            const float4 a = geo[3 * i], b = geo[3 * i + 1];
            const float c = geo[3 * i + 2].x;
            const float comp[kBrutePairWords] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c};
            for (int k = 0; k < kBrutePairWords; ++k) w[2 * (kBrutePairWords * (i >> 1) + k) + (i & 1)] = comp[k];
        }
        // Cull boxes, one per pair.  Part of the closest-hit DEFINITION (the oracle builds the same boxes with the
        // same fp32 operations): vertices are p0, p0+e1, p0+e2 of the stored records; the box is padded on every
        // axis so that any hit Moeller-Trumbore accepts lies inside it -- by 2.5e-4 of the scene extent on axes where
        // the pair has thickness (fp32 noise of the barycentric test moves accepted hits sideways by far less), by
        // 2.5e-7 of it on an axis where the pair is flat (the hit lies on the pair's plane by construction; the small
        // pad keeps rays that leave a flat quad from testing that quad again unless they graze it).
        const int npairs = (ntris + 1) / 2;
        float slo[3] = {1e30f, 1e30f, 1e30f}, shi[3] = {-1e30f, -1e30f, -1e30f};
        std::vector<float> blo((size_t) 3 * npairs, 1e30f), bhi((size_t) 3 * npairs, -1e30f);
        for (int i = 0; i < ntris; ++i) {
            const float4 a = geo[3 * i], b = geo[3 * i + 1];
            const float c = geo[3 * i + 2].x;
            const float p0[3] = {a.x, a.y, a.z}, e1[3] = {a.w, b.x, b.y}, e2[3] = {b.z, b.w, c};
            for (int k = 0; k < 3; ++k) {
                const float v[3] = {p0[k], p0[k] + e1[k], p0[k] + e2[k]};
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
        float *bw = reinterpret_cast<float *>(d.bg_box);
        std::memset(d.bg_box, 0, sizeof(d.bg_box));
        for (int j = 0; j < npairs; ++j)
            for (int k = 0; k < 3; ++k) {
                const float lo = blo[3 * j + k], hi = bhi[3 * j + k];
                const float half = 0.5f * (hi - lo);
                bw[2 * (kBruteBoxWords * (j >> 1) + k) + (j & 1)] = 0.5f * (lo + hi);
                bw[2 * (kBruteBoxWords * (j >> 1) + 3 + k) + (j & 1)] = half + (half < flat_below ? pad_flat : pad_fat);
            }
    }

    d.env = DEnv{};
    if (sc.env.present) {
        const HEnvmap &e = sc.env;
        DEnv &de = d.env;
        de.present = 1;
        de.emitter = e.emitter;
        de.w = e.w; de.h = e.h;
        const unsigned char *big = (const unsigned char *) db.big;
        de.data = (const float *) (big + db.big_off[0]);
        de.ddata = e.ddata.empty() ? nullptr : (const float *) (big + db.big_off[1]);
        de.scale = e.scale.v; de.d_scale = e.scale.d;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                de.to_world[3 * i + j] = e.to_world_full.m[i][j].v; de.d_to_world[3 * i + j] = e.to_world_full.m[i][j].d;
                de.from_world[3 * i + j] = e.from_world.m[i][j].v; de.d_from_world[3 * i + j] = e.from_world.m[i][j].d;
            }
        de.lower[0] = e.lower.x; de.lower[1] = e.lower.y; de.lower[2] = e.lower.z;
        de.upper[0] = e.upper.x; de.upper[1] = e.upper.y; de.upper[2] = e.upper.z;
        de.cw = e.cw; de.ch = e.ch;
        de.cell_sum = e.cell.sum;
        de.cell_pmf = (const float *) (big + db.big_off[2]);
        de.cell_cmf = (const float *) (big + db.big_off[3]);
        de.cell_lut = (const int *) (big + db.big_off[4]);
        de.cell_lut_n = db.env_lut_n;
    }

    sc.dcameras.assign(sc.cameras.size(), DCamera{});
    size_t pmf_off = 0;
    for (size_t ci = 0; ci < sc.cameras.size(); ++ci) {
        const HCamera &c = sc.cameras[ci];
        DCamera &dc = sc.dcameras[ci];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                dc.sample_to_camera[4 * i + j] = c.sample_to_camera.m[i][j];
                dc.to_world[4 * i + j] = c.to_world_full.m[i][j].v;
                dc.d_to_world[4 * i + j] = c.to_world_full.m[i][j].d;
                dc.world_to_sample[4 * i + j] = c.world_to_sample.m[i][j].v;
                dc.d_world_to_sample[4 * i + j] = c.world_to_sample.m[i][j].d;
            }
        dc.pos[0] = c.pos.x.v; dc.pos[1] = c.pos.y.v; dc.pos[2] = c.pos.z.v;
        dc.d_pos[0] = c.pos.x.d; dc.d_pos[1] = c.pos.y.d; dc.d_pos[2] = c.pos.z.d;
        dc.dir[0] = c.dir.x.v; dc.dir[1] = c.dir.y.v; dc.dir[2] = c.dir.z.v;
        dc.d_dir[0] = c.dir.x.d; dc.d_dir[1] = c.dir.y.d; dc.d_dir[2] = c.dir.z.d;
        dc.inv_area = c.inv_area;
        dc.ortho = c.ortho ? 1 : 0;
        dc.n_edges = (int) c.edges.size();
        dc.edge_sum = c.edges.empty() ? 0.f : c.edge_distrb.sum;
        dc.pe_a = (const float4 *) (base + o_pa) + cam_edge_first[ci];
        dc.pe_da = (const float4 *) (base + o_pda) + cam_edge_first[ci];
        dc.pe_b = (const float4 *) (base + o_pb) + cam_edge_first[ci];
        dc.pe_pmf = (const float *) (base + o_pp) + pmf_off;
        dc.pe_cmf = (const float *) (base + o_pc) + pmf_off;
        pmf_off += c.edges.size();
        dc.guided = (c.guide_ready && c.guide_enabled) ? 1 : 0;
        for (int k = 0; k < 3; ++k) dc.greso[k] = c.greso[k];
        dc.guide_sum = c.guide_ready ? c.guide.sum : 0.f;
        dc.guide_pmf = (const float *) (base + o_gp) + cam_guide_first[ci];
        dc.guide_cmf = (const float *) (base + o_gc) + cam_guide_first[ci];
        dc.guide_lut = (const int *) (base + o_gl) + cam_lut_first[ci];
        dc.guide_lut_n = cam_lut_n[ci];
    }
}

}  // namespace psdr
