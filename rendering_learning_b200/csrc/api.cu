// C ABI of librl_b200.so (include/rl_b200.h): context, scene upload (flatten -> HBM -> LBVH build),
// ray batches, renders.  No CPU fallback anywhere: without an sm_100 device rl_create fails.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "device.cuh"
#include "kernels.h"
#include "lbvh.h"
#include "obj_ingest.h"
#include "scene.h"

namespace rl {
bool invert_affine_4x4(const double* m16, double* inv12, std::string* err);
cudaError_t measure_peaks(cudaStream_t s, int sm_count, double* fp32_tflops, double* l2_gbs, double* hbm_gbs);
}

using namespace rl;

namespace {
thread_local std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
        if (e == cudaSuccess) cap = bytes ? bytes : 16;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};
}  // namespace

struct rl_ctx {
    int device = 0;
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    size_t hbm_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string error;
    bool instrumented = false;
    bool has_scene = false;
    // scene buffers
    DevBuf prims, tri_verts, tri_shade, tri_plane, xforms, spheres, quads, sphere_node, quad_node, ord_node, materials, textures,
        images, lights, nodes;
    std::vector<DevBuf> image_texels;
    DevBuf big_refs, csg, media, medium_refs, perlin_vec, perlin_perm, bvh_aabb, bvh_ref, bvh_node_id, bounds, keys, sorted_prim, keys_tmp, idx_tmp, left, right, parent,
        node_aabb, lbvh_counters;
    DevScene ds{};
    rl_scene_info info{};
    float upload_ms = 0.0f;
    int upload_launches = 0;
    // work buffers
    DevBuf counters, queue, jobs, prefix, frame, frame8, partial, rays, hits;
    DevBuf wf_state, wf_rays, wf_ctr;  // the global-wavefront variant's path state (ow.variant = 7)
    ObjMesh mesh;                  // the mesh rl_obj_parse left on this device (RL_RTC_MESH / RL_OW_MESH nodes instance it)
    rl_obj_info mesh_info{};
    bool has_mesh = false;
    std::vector<rl_ctx*> group;    // rl_create_multi: [this, peer 1, ...]; empty for a single-GPU ctx
    cudaEvent_t ev_go = nullptr, ev_done = nullptr;  // group renders: leader's "queue is reset", member's "kernel finished"
    OwTuning tune;                 // scheduling parameters of the OW kernel (rl_set_option)
    std::vector<int> node_ref;     // scene node id -> leaf ref (-1: not a leaf), for rl_trace_batch_ex's self nodes
    DevBuf self_refs;
    std::vector<DevBuf> job_tables;  // job tables of launches still in flight (> JOBS_INLINE jobs); freed at rl_synchronize
    // cross-GPU queue (CUDA IPC): the owner allocates it, peers map it
    DevBuf shared_queue_own, shared_partial_own;
    float* shared_partial = nullptr;
    uint64_t shared_partial_bytes = 0;  // bytes per slot behind shared_partial (owner: allocated; peer: told at import)
    int shared_partial_slots = 0;
    bool shared_partial_imported = false;
    unsigned long long* shared_queue = nullptr;
    bool shared_queue_imported = false;
};

#define CK(ctx, call)                                                                           \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            (ctx)->error = std::string(#call) + ": " + cudaGetErrorString(e_);                  \
            return RL_E_CUDA;                                                                   \
        }                                                                                       \
    } while (0)

static int fail(rl_ctx* c, int code, const std::string& msg) {
    c->error = msg;
    return code;
}

template <class T>
static cudaError_t upload(DevBuf& b, const std::vector<T>& v, cudaStream_t s) {
    cudaError_t e = b.reserve(v.size() * sizeof(T));
    if (e != cudaSuccess || v.empty()) return e;
    return cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
}

// what a flattened scene will occupy / contain (shared by rl_scene_upload and the host-only rl_scene_check)
static rl_scene_info scene_info_of(const FlatScene& fs) {
    rl_scene_info si{};
    const int n = (int)fs.bvh_ref.size();
    const int nn = n >= 2 ? n - 1 : (n == 1 ? 1 : 0);
    size_t image_bytes = 0;
    for (const auto& im : fs.images) image_bytes += im.texels.size() * sizeof(float4);
    si.flavor = fs.flavor;
    si.n_prims = fs.flavor == RL_FLAVOR_OW ? (int)fs.big_refs.size() : (int)fs.prims.size();
    si.n_bvh_prims = n;
    si.n_bvh_nodes = n >= 2 ? n - 1 : 0;
    si.n_materials = (int)fs.materials.size();
    si.n_textures = (int)fs.textures.size();
    si.n_lights = (int)fs.lights.size();
    si.has_transparency = fs.has_transparency;
    si.device_bytes = (int64_t)(fs.prims.size() * sizeof(RtcPrim) + fs.tri_verts.size() * (sizeof(TriVerts) + sizeof(TriShade)) +
                                fs.spheres.size() * sizeof(OwSphere) + fs.quads.size() * sizeof(OwQuad) +
                                (size_t)nn * sizeof(BvhNode) + image_bytes + fs.materials.size() * sizeof(DevMaterial));
    return si;
}

extern "C" {

int rl_abi_version(void) { return RL_B200_ABI_VERSION; }

int rl_scene_check(const rl_scene_desc* scene, rl_scene_info* out, char* err, int32_t err_cap) {
    FlatScene fs;
    std::string msg;
    int rc = flatten_scene(scene, &fs, &msg);
    if (err && err_cap > 0) {
        size_t k = msg.size() < (size_t)(err_cap - 1) ? msg.size() : (size_t)(err_cap - 1);
        memcpy(err, msg.data(), k);
        err[k] = 0;
    }
    if (rc == RL_OK && out) *out = scene_info_of(fs);
    return rc;
}

const char* rl_last_error(const rl_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int rl_create(int device_id, rl_ctx** out) {
    if (!out) {
        g_create_error = "rl_create: out is NULL";
        return RL_E_INVALID;
    }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        g_create_error = std::string("no CUDA device visible (") + (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0") +
                         "); this library has no CPU fallback";
        return RL_E_NO_DEVICE;
    }
    if (device_id < 0 || device_id >= n) {
        g_create_error = "device id out of range";
        return RL_E_INVALID;
    }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device_id);
    if (e != cudaSuccess) {
        g_create_error = cudaGetErrorString(e);
        return RL_E_CUDA;
    }
    if (prop.major != 10) {
        char buf[160];
        snprintf(buf, sizeof(buf), "device %d is sm_%d%d; librl_b200 carries sm_100a code only (no fallback)", device_id,
                 prop.major, prop.minor);
        g_create_error = buf;
        return RL_E_NO_DEVICE;
    }
    rl_ctx* c = new rl_ctx();
    c->device = device_id;
    c->sm_count = prop.multiProcessorCount;
    c->cc_major = prop.major;
    c->cc_minor = prop.minor;
    c->hbm_bytes = prop.totalGlobalMem;
    if ((e = cudaSetDevice(device_id)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&c->ev0)) != cudaSuccess || (e = cudaEventCreate(&c->ev1)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&c->ev_go, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming)) != cudaSuccess ||
        (e = c->counters.reserve(sizeof(Counters))) != cudaSuccess || (e = c->queue.reserve(2 * sizeof(unsigned long long))) != cudaSuccess) {
        g_create_error = cudaGetErrorString(e);
        delete c;
        return RL_E_CUDA;
    }
    cudaMemset(c->counters.p, 0, sizeof(Counters));
    *out = c;
    return RL_OK;
}

void rl_destroy(rl_ctx* c) {
    if (!c) return;
    for (size_t g = 1; g < c->group.size(); g++) {  // a multi-GPU ctx owns its peers
        c->group[g]->group.clear();
        rl_destroy(c->group[g]);
    }
    c->group.clear();
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->shared_queue_imported && c->shared_queue) cudaIpcCloseMemHandle(c->shared_queue);
    if (c->shared_partial_imported && c->shared_partial) cudaIpcCloseMemHandle(c->shared_partial);
    c->shared_partial_own.release();
    c->shared_queue_own.release();
    DevBuf* all[] = {&c->prims, &c->tri_verts, &c->tri_shade, &c->tri_plane, &c->xforms, &c->spheres, &c->quads, &c->sphere_node,
                     &c->quad_node, &c->ord_node, &c->materials, &c->textures, &c->images, &c->lights, &c->nodes, &c->big_refs, &c->csg, &c->media, &c->medium_refs, &c->perlin_vec, &c->perlin_perm, &c->bvh_aabb,
                     &c->bvh_ref, &c->bvh_node_id, &c->bounds, &c->keys, &c->sorted_prim, &c->keys_tmp, &c->idx_tmp,
                     &c->left, &c->right, &c->parent, &c->node_aabb, &c->lbvh_counters, &c->counters, &c->queue,
                     &c->jobs, &c->prefix, &c->frame, &c->frame8, &c->partial, &c->rays, &c->hits, &c->self_refs,
                     &c->wf_state, &c->wf_rays, &c->wf_ctr};
    for (DevBuf* b : all) b->release();
    for (DevBuf& b : c->image_texels) b.release();
    for (DevBuf& b : c->job_tables) b.release();
    obj_mesh_free(&c->mesh);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev_go) cudaEventDestroy(c->ev_go);
    if (c->ev_done) cudaEventDestroy(c->ev_done);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int rl_create_multi(const int32_t* device_ids, int32_t n, rl_ctx** out) {
    if (!out || !device_ids || n < 1 || n > 64) {
        g_create_error = "rl_create_multi: bad arguments";
        return RL_E_INVALID;
    }
    *out = nullptr;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < i; j++)
            if (device_ids[i] == device_ids[j]) {
                g_create_error = "rl_create_multi: device ids must be distinct";
                return RL_E_INVALID;
            }
    std::vector<rl_ctx*> g;
    auto undo = [&]() { for (rl_ctx* m : g) rl_destroy(m); };
    for (int i = 0; i < n; i++) {
        rl_ctx* m = nullptr;
        int rc = rl_create(device_ids[i], &m);
        if (rc != RL_OK) { undo(); return rc; }
        g.push_back(m);
    }
    // every GPU's kernels pop GPU 0's counter and store into GPU 0's partial-sum buffer: peers need access to GPU 0
    for (int i = 1; i < n; i++) {
        int can = 0;
        cudaError_t e = cudaDeviceCanAccessPeer(&can, device_ids[i], device_ids[0]);
        if (e == cudaSuccess && can) {
            cudaSetDevice(device_ids[i]);
            e = cudaDeviceEnablePeerAccess(device_ids[0], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
        }
        if (e != cudaSuccess || !can) {
            char buf[160];
            snprintf(buf, sizeof(buf), "GPU %d cannot access GPU %d's memory (%s)", device_ids[i], device_ids[0],
                     e != cudaSuccess ? cudaGetErrorString(e) : "no peer path");
            g_create_error = buf;
            undo();
            return RL_E_CUDA;
        }
    }
    cudaSetDevice(device_ids[0]);
    if (n > 1) g[0]->group = g;
    *out = g[0];
    return RL_OK;
}

int rl_device_count(const rl_ctx* c) { return !c ? 0 : (c->group.empty() ? 1 : (int)c->group.size()); }

int rl_device_info(rl_ctx* c, int* sm_count, int* cc_major, int* cc_minor, int64_t* hbm_bytes) {
    if (!c) return RL_E_INVALID;
    if (sm_count) *sm_count = c->sm_count;
    if (cc_major) *cc_major = c->cc_major;
    if (cc_minor) *cc_minor = c->cc_minor;
    if (hbm_bytes) *hbm_bytes = (int64_t)c->hbm_bytes;
    return RL_OK;
}

int rl_synchronize(rl_ctx* c) {
    if (!c) return RL_E_INVALID;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize(c->stream));
    CK(c, cudaDeviceSynchronize());  // asynchronous launches may have gone to a caller-provided stream
    for (DevBuf& b : c->job_tables) b.release();  // nothing is in flight any more
    c->job_tables.clear();
    // The overflow counter is STICKY: no entry point zeroes it except the one that reports it (here and read_counters),
    // so an overflow of an asynchronous launch cannot be wiped by a later call before anyone has seen it.
    Counters h;
    CK(c, cudaMemcpy(&h, c->counters.p, sizeof(h), cudaMemcpyDeviceToHost));
    if (h.overflow) {
        CK(c, cudaMemset(&c->counters.as<Counters>()->overflow, 0, sizeof(h.overflow)));
        c->error = "a traversal stack / work list overflowed on the device (or a render kernel's watchdog fired)";
        return RL_E_OVERFLOW;
    }
    return RL_OK;
}

static const struct { const char* name; int OwTuning::*field; int lo, hi; } k_options[] = {
    {"ow.variant", &OwTuning::variant, 5, 7},        {"ow.slots", &OwTuning::slots, 0, 1 << 22},
    {"ow.minb", &OwTuning::minb, 0, 4},              {"ow.ctas_per_sm", &OwTuning::ctas_per_sm, 0, 8},
    {"ow.svc_lo", &OwTuning::svc_lo, 0, 32},         {"ow.exit_min", &OwTuning::exit_min, 0, 32},
    {"ow.leaf_min", &OwTuning::leaf_min, 0, 32},     {"ow.svc_min", &OwTuning::svc_min, 0, 32},
};

int rl_set_option(rl_ctx* c, const char* name, int32_t value) {
    if (!c || !name) return RL_E_INVALID;
    for (const auto& o : k_options) {
        if (strcmp(o.name, name) != 0) continue;
        if (value < o.lo || value > o.hi || (o.field == &OwTuning::minb && value != 0 && value != 3 && value != 4))
            return fail(c, RL_E_INVALID, std::string("option value out of range: ") + name);
        c->tune.*(o.field) = value;
        return RL_OK;
    }
    return fail(c, RL_E_INVALID, std::string("unknown option: ") + name);
}

int rl_measure_peaks(rl_ctx* c, double* fp32_tflops, double* l2_gbs, double* hbm_gbs) {
    if (!c) return RL_E_INVALID;
    CK(c, cudaSetDevice(c->device));
    CK(c, measure_peaks(c->stream, c->sm_count, fp32_tflops, l2_gbs, hbm_gbs));
    return RL_OK;
}

int rl_set_instrumented(rl_ctx* c, int enabled) {
    if (!c) return RL_E_INVALID;
    c->instrumented = enabled != 0;
    return RL_OK;
}

static int upload_flat(rl_ctx* c, const FlatScene& fs, int n_scene_nodes);

int rl_scene_upload(rl_ctx* c, const rl_scene_desc* scene) {
    if (!c) return RL_E_INVALID;
    c->has_scene = false;
    FlatScene fs;
    std::string err;
    MeshMeta mm{};
    if (c->has_mesh) {
        mm.n_triangles = c->mesh_info.n_triangles;
        for (int k = 0; k < 6; k++) mm.bounds[k] = c->mesh_info.bounds[k];
    }
    int rc = flatten_scene(scene, &fs, &err, c->has_mesh ? &mm : nullptr);
    if (rc != RL_OK) return fail(c, rc, err);
    rc = upload_flat(c, fs, scene->n_nodes);
    // a multi-GPU ctx replicates the scene (and builds the LBVH) on every GPU: <= a few MB, bit-identical builds
    for (size_t g = 1; rc == RL_OK && g < c->group.size(); g++) {
        rc = upload_flat(c->group[g], fs, scene->n_nodes);
        if (rc != RL_OK) c->error = c->group[g]->error;
    }
    return rc;
}

static int upload_flat(rl_ctx* c, const FlatScene& fs, int n_scene_nodes) {
    c->has_scene = false;
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    CK(c, cudaEventRecord(c->ev0, s));
    CK(c, upload(c->prims, fs.prims, s));
    CK(c, upload(c->tri_verts, fs.tri_verts, s));
    CK(c, upload(c->tri_shade, fs.tri_shade, s));
    CK(c, upload(c->xforms, fs.xforms, s));
    CK(c, upload(c->spheres, fs.spheres, s));
    CK(c, upload(c->quads, fs.quads, s));
    CK(c, upload(c->sphere_node, fs.sphere_node, s));
    CK(c, upload(c->quad_node, fs.quad_node, s));
    CK(c, upload(c->ord_node, fs.ord_node, s));
    CK(c, upload(c->materials, fs.materials, s));
    CK(c, upload(c->textures, fs.textures, s));
    CK(c, upload(c->lights, fs.lights, s));
    // images
    // image buffers are reused across uploads (cudaFree synchronises the device and was measured at 100+ ms on a
    // context that also holds the 800 MB partial-sum buffer of a 4K render)
    while (c->image_texels.size() > fs.images.size()) {
        c->image_texels.back().release();
        c->image_texels.pop_back();
    }
    c->image_texels.resize(fs.images.size());
    std::vector<DevImage> dimg(fs.images.size());
    size_t image_bytes = 0;
    for (size_t i = 0; i < fs.images.size(); i++) {
        CK(c, upload(c->image_texels[i], fs.images[i].texels, s));
        dimg[i].texels = c->image_texels[i].as<float4>();
        dimg[i].width = fs.images[i].w;
        dimg[i].height = fs.images[i].h;
        image_bytes += fs.images[i].texels.size() * sizeof(float4);
    }
    CK(c, upload(c->images, dimg, s));
    // LBVH
    int n = (int)fs.bvh_ref.size();
    int nn = n >= 2 ? n - 1 : (n == 1 ? 1 : 0);
    CK(c, upload(c->big_refs, fs.big_refs, s));
    CK(c, upload(c->csg, fs.csg, s));
    CK(c, upload(c->media, fs.media, s));
    CK(c, upload(c->medium_refs, fs.medium_refs, s));
    CK(c, upload(c->perlin_vec, fs.perlin_vec, s));
    CK(c, upload(c->perlin_perm, fs.perlin_perm, s));
    CK(c, upload(c->bvh_aabb, fs.bvh_aabb, s));
    CK(c, upload(c->bvh_ref, fs.bvh_ref, s));
    CK(c, upload(c->bvh_node_id, fs.bvh_node_id, s));
    c->upload_launches = 0;
    // uses of the device-resident mesh: transform bake, f32 packing and LBVH input boxes of its triangles, on the device
    for (const FlatScene::MeshUse& u : fs.meshes) {
        if (!c->has_mesh || u.bvh_first < 0) return fail(c, RL_E_INVALID, "the scene instances a mesh this GPU does not hold");
        MeshInstance mi{};
        memcpy(mi.fwd, u.fwd, sizeof(mi.fwd));
        memcpy(mi.inv, u.inv, sizeof(mi.inv));
        mi.flavor = fs.flavor;
        mi.material = u.material;
        mi.node = u.node;
        mi.report_node = (fs.flavor == RL_FLAVOR_RTC && u.node >= 0 && u.node < (int)fs.ord_node.size()) ? fs.ord_node[u.node] : u.node;
        mi.tri_first = u.tri_first;
        mi.bvh_first = u.bvh_first;
        mi.xf = u.xf;
        CK(c, launch_mesh_instance(c->mesh, mi, c->tri_verts.as<TriVerts>(), c->tri_shade.as<TriShade>(), c->bvh_aabb.as<float>(),
                                   c->bvh_ref.as<int>(), c->bvh_node_id.as<int>(), s));
        c->upload_launches++;
    }
    if (fs.flavor == RL_FLAVOR_OW && !fs.tri_verts.empty()) {  // the plane form of every triangle, host-flattened or instanced above
        CK(c, c->tri_plane.reserve(fs.tri_verts.size() * sizeof(OwTriPlane)));
        CK(c, launch_tri_planes(c->tri_verts.as<TriVerts>(), (int)fs.tri_verts.size(), c->tri_plane.as<OwTriPlane>(), s));
        c->upload_launches++;
    }
    if (n > 0) {
        CK(c, c->bounds.reserve(6 * sizeof(float)));
        CK(c, c->keys.reserve(n * sizeof(uint64_t)));
        CK(c, c->keys_tmp.reserve(n * sizeof(uint64_t)));
        CK(c, c->sorted_prim.reserve(n * sizeof(int)));
        CK(c, c->idx_tmp.reserve(n * sizeof(int)));
        CK(c, c->left.reserve((size_t)nn * sizeof(int)));
        CK(c, c->right.reserve((size_t)nn * sizeof(int)));
        CK(c, c->parent.reserve((size_t)(2 * n) * sizeof(int)));
        CK(c, c->node_aabb.reserve((size_t)nn * 6 * sizeof(float)));
        CK(c, c->lbvh_counters.reserve((size_t)nn * sizeof(int)));
        CK(c, c->nodes.reserve((size_t)nn * sizeof(BvhNode)));
        LbvhBuffers lb;
        lb.prim_aabb = c->bvh_aabb.as<float>();
        lb.prim_ref = c->bvh_ref.as<int>();
        lb.bounds = c->bounds.as<float>();
        lb.keys = c->keys.as<uint64_t>();
        lb.sorted_prim = c->sorted_prim.as<int>();
        lb.keys_tmp = c->keys_tmp.as<uint64_t>();
        lb.idx_tmp = c->idx_tmp.as<int>();
        lb.left = c->left.as<int>();
        lb.right = c->right.as<int>();
        lb.parent = c->parent.as<int>();
        lb.node_aabb = c->node_aabb.as<float>();
        lb.counters = c->lbvh_counters.as<int>();
        lb.nodes = c->nodes.as<BvhNode>();
        CK(c, lbvh_build(lb, n, s, &c->upload_launches));
    }
    CK(c, cudaEventRecord(c->ev1, s));
    CK(c, cudaStreamSynchronize(s));  // the host vectors die with this scope
    CK(c, cudaEventElapsedTime(&c->upload_ms, c->ev0, c->ev1));

    DevScene& d = c->ds;
    d = DevScene{};
    d.flavor = fs.flavor;
    d.n_prims = (int)fs.prims.size();
    d.n_tris = (int)fs.tri_verts.size();
    d.n_spheres = (int)fs.spheres.size();
    d.n_quads = (int)fs.quads.size();
    d.n_bvh_prims = n;
    d.n_bvh_nodes = nn;
    d.n_big = (int)fs.big_refs.size();
    d.n_csg = (int)fs.csg.size();
    d.n_media = (int)fs.media.size();
    d.n_perlins = (int)fs.perlin_vec.size() / 256;
    d.n_materials = (int)fs.materials.size();
    d.n_textures = (int)fs.textures.size();
    d.n_lights = (int)fs.lights.size();
    d.n_images = (int)fs.images.size();
    d.n_xforms = (int)fs.xforms.size();
    d.has_transparency = fs.has_transparency;
    d.max_reflection_depth = fs.max_reflection_depth;
    for (int k = 0; k < 3; k++) d.void_color[k] = fs.void_color[k];
    d.prims = c->prims.as<RtcPrim>();
    d.tri_verts = c->tri_verts.as<TriVerts>();
    d.tri_shade = c->tri_shade.as<TriShade>();
    d.tri_plane = c->tri_plane.as<OwTriPlane>();
    d.xforms = c->xforms.as<Xform>();
    d.spheres = c->spheres.as<OwSphere>();
    d.quads = c->quads.as<OwQuad>();
    d.sphere_node = c->sphere_node.as<int>();
    d.quad_node = c->quad_node.as<int>();
    d.ord_node = c->ord_node.as<int>();
    d.n_ord = (int)fs.ord_node.size();
    d.materials = c->materials.as<DevMaterial>();
    d.textures = c->textures.as<DevTexture>();
    d.images = c->images.as<DevImage>();
    d.lights = c->lights.as<DevLight>();
    d.nodes = c->nodes.as<BvhNode>();
    d.big_refs = c->big_refs.as<int>();
    d.csg = c->csg.as<int4>();
    d.media = c->media.as<OwMedium>();
    d.medium_refs = c->medium_refs.as<int>();
    d.perlin_vec = c->perlin_vec.as<float4>();
    d.perlin_perm = c->perlin_perm.as<int>();

    // node id -> leaf ref (rl_trace_batch_ex takes the node a ray starts on)
    c->node_ref.assign((size_t)(n_scene_nodes > 0 ? n_scene_nodes : 0), -1);
    auto note = [&](int node, int ref) { if (node >= 0 && node < (int)c->node_ref.size()) c->node_ref[node] = ref; };
    for (size_t i = 0; i < fs.sphere_node.size(); i++) note(fs.sphere_node[i], make_ref(REF_SPHERE, (int)i));
    for (size_t i = 0; i < fs.quad_node.size(); i++) note(fs.quad_node[i], make_ref(REF_QUAD, (int)i));
    for (size_t i = 0; fs.flavor == RL_FLAVOR_OW && i < fs.tri_verts.size(); i++) {  // (RTC triangles carry DFS ordinals there)
        int node;
        memcpy(&node, &fs.tri_verts[i].p1.w, sizeof(int));
        note(node, make_ref(REF_TRI, (int)i));
    }

    c->info = scene_info_of(fs);
    c->has_scene = true;
    return RL_OK;
}

int rl_obj_parse(rl_ctx* c, const char* text, uint64_t len, int32_t flavor, rl_obj_info* info) {
    if (!c || (len > 0 && !text)) return RL_E_INVALID;
    if (flavor != RL_FLAVOR_RTC && flavor != RL_FLAVOR_OW) return fail(c, RL_E_INVALID, "unknown flavor");
    const size_t members = c->group.empty() ? 1 : c->group.size();
    for (size_t g = 0; g < members; g++) {  // a multi-GPU ctx parses on every GPU: each one instances its own copy
        rl_ctx* m = c->group.empty() ? c : c->group[g];
        CK(c, cudaSetDevice(m->device));
        obj_mesh_free(&m->mesh);
        m->has_mesh = false;
        std::string err;
        int rc = obj_parse_device(text, len, flavor, m->stream, &m->mesh, &m->mesh_info, &err);
        if (rc != RL_OK) return fail(c, rc, err);
        m->has_mesh = true;
    }
    CK(c, cudaSetDevice(c->device));
    if (info) *info = c->mesh_info;
    return RL_OK;
}

int rl_obj_download(rl_ctx* c, double* tri_p, double* tri_n, double* tri_uv, uint8_t* flags) {
    if (!c) return RL_E_INVALID;
    if (!c->has_mesh) return fail(c, RL_E_INVALID, "no mesh parsed on this ctx");
    const size_t n = (size_t)c->mesh.n_triangles;
    if (n == 0) return RL_OK;
    CK(c, cudaSetDevice(c->device));
    if (tri_p) CK(c, cudaMemcpyAsync(tri_p, c->mesh.tri_p, 9 * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (tri_n) CK(c, cudaMemcpyAsync(tri_n, c->mesh.tri_n, 9 * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (tri_uv) CK(c, cudaMemcpyAsync(tri_uv, c->mesh.tri_uv, 6 * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (flags) CK(c, cudaMemcpyAsync(flags, c->mesh.tri_flags, n, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return RL_OK;
}

int rl_scene_info_get(rl_ctx* c, rl_scene_info* out) {
    if (!c || !out) return RL_E_INVALID;
    if (!c->has_scene) return fail(c, RL_E_NO_SCENE, "no scene uploaded");
    *out = c->info;
    return RL_OK;
}

int rl_lbvh_download(rl_ctx* c, rl_lbvh_host* out) {
    if (!c || !out) return RL_E_INVALID;
    if (!c->has_scene) return fail(c, RL_E_NO_SCENE, "no scene uploaded");
    int n = c->info.n_bvh_prims, m = c->info.n_bvh_nodes;
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    auto dl = [&](void* dst, const DevBuf& src, size_t bytes) -> cudaError_t {
        if (!dst || bytes == 0) return cudaSuccess;
        return cudaMemcpyAsync(dst, src.p, bytes, cudaMemcpyDeviceToHost, s);
    };
    if (n > 0) {
        CK(c, dl(out->prim_aabb, c->bvh_aabb, (size_t)n * 6 * sizeof(float)));
        CK(c, dl(out->prim_node, c->bvh_node_id, (size_t)n * sizeof(int)));
        CK(c, dl(out->morton, c->keys, (size_t)n * sizeof(uint64_t)));
        CK(c, dl(out->sorted_prim, c->sorted_prim, (size_t)n * sizeof(int)));
        float b[6];
        CK(c, cudaMemcpyAsync(b, c->bounds.p, sizeof(b), cudaMemcpyDeviceToHost, s));
        CK(c, cudaStreamSynchronize(s));
        for (int k = 0; k < 3; k++) {
            out->scene_lo[k] = b[k];
            out->scene_hi[k] = b[3 + k];
        }
    }
    if (m > 0) {
        CK(c, dl(out->left, c->left, (size_t)m * sizeof(int)));
        CK(c, dl(out->right, c->right, (size_t)m * sizeof(int)));
        CK(c, dl(out->parent, c->parent, (size_t)(m + n) * sizeof(int)));
        CK(c, dl(out->node_aabb, c->node_aabb, (size_t)m * 6 * sizeof(float)));
    }
    CK(c, cudaStreamSynchronize(s));
    return RL_OK;
}

}  // extern "C"

// ---- job tables ------------------------------------------------------------------------------------------
static int make_job_table(rl_ctx* c, const rl_job* jobs, int n_jobs, int width, int height, int n_chunks, bool ow,
                          cudaStream_t s, JobTable* jt) {
    if (n_jobs < 0 || (n_jobs > 0 && !jobs)) return fail(c, RL_E_INVALID, "bad job list");
    std::vector<long long> prefix(n_jobs + 1, 0);
    for (int i = 0; i < n_jobs; i++) {
        const rl_job& j = jobs[i];
        if (j.x0 < 0 || j.y0 < 0 || j.x1 > width || j.y1 > height || j.x0 >= j.x1 || j.y0 >= j.y1)
            return fail(c, RL_E_INVALID, "job rectangle outside the image");
        long long items = padded_pixels(j.x1 - j.x0, j.y1 - j.y0);
        if (ow) {
            if (j.chunk_begin < 0 || j.chunk_end > n_chunks || j.chunk_begin >= j.chunk_end)
                return fail(c, RL_E_INVALID, "job chunk range outside [0, n_chunks)");
            items *= (j.chunk_end - j.chunk_begin);
        }
        prefix[i + 1] = prefix[i] + items;
    }
    jt->n_jobs = n_jobs;
    jt->n_items = prefix[n_jobs];
    jt->jobs = nullptr;
    jt->prefix = nullptr;
    if (n_jobs <= JOBS_INLINE) {
        // small tables ride in the kernel parameters: nothing to upload, safe for back-to-back async launches
        jt->inline_jobs = 1;
        for (int i = 0; i < n_jobs; i++) jt->ijobs[i] = jobs[i];
        for (int i = 0; i <= n_jobs; i++) jt->iprefix[i] = prefix[i];
        return RL_OK;
    }
    // Larger tables get their OWN device buffer per launch (freed at rl_synchronize): a launch on another caller stream may
    // still be reading the previous table, so the buffer is never reused while anything can be in flight.
    jt->inline_jobs = 0;
    c->job_tables.emplace_back();
    DevBuf& buf = c->job_tables.back();
    const size_t jb = (sizeof(rl_job) * (size_t)n_jobs + 15) / 16 * 16;
    CK(c, buf.reserve(jb + sizeof(long long) * (size_t)(n_jobs + 1)));
    CK(c, cudaMemcpyAsync(buf.p, jobs, sizeof(rl_job) * (size_t)n_jobs, cudaMemcpyHostToDevice, s));
    CK(c, cudaMemcpyAsync((char*)buf.p + jb, prefix.data(), sizeof(long long) * (size_t)(n_jobs + 1), cudaMemcpyHostToDevice, s));
    CK(c, cudaStreamSynchronize(s));  // `prefix` is a local; `jobs` is the caller's
    jt->jobs = reinterpret_cast<const rl_job*>(buf.p);
    jt->prefix = reinterpret_cast<const long long*>((char*)buf.p + jb);
    return RL_OK;
}

// zero the statistics counters of a synchronous call, leaving the sticky overflow counter alone
static cudaError_t reset_stats(rl_ctx* c, cudaStream_t s) {
    return cudaMemsetAsync(c->counters.p, 0, offsetof(Counters, overflow), s);
}

static int read_counters(rl_ctx* c, cudaStream_t s, rl_stats* st) {
    Counters h;
    CK(c, cudaMemcpyAsync(&h, c->counters.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    CK(c, cudaStreamSynchronize(s));
    if (st) {
        st->rays = h.rays;
        st->node_visits = h.node_visits;
        st->prim_tests = h.prim_tests;
        st->tri_tests = h.tri_tests;
        st->shades = h.shades;
        st->overflow = h.overflow;
    }
    if (h.overflow) {
        CK(c, cudaMemsetAsync(&c->counters.as<Counters>()->overflow, 0, sizeof(h.overflow), s));  // reported: clear
        return fail(c, RL_E_OVERFLOW, "a traversal stack / work list overflowed on the device");
    }
    return RL_OK;
}

extern "C" {

int rl_trace_batch_ex(rl_ctx* c, const rl_ray* rays, const int32_t* self_nodes, uint64_t n, rl_hit* out) {
    if (!c || (n > 0 && (!rays || !out))) return RL_E_INVALID;
    if (!c->has_scene) return fail(c, RL_E_NO_SCENE, "no scene uploaded");
    if (n == 0) return RL_OK;
    if (self_nodes && c->ds.flavor != RL_FLAVOR_OW) return fail(c, RL_E_INVALID, "self nodes apply to OW scenes only");
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    CK(c, c->rays.reserve(n * sizeof(rl_ray)));
    CK(c, c->hits.reserve(n * sizeof(rl_hit)));
    CK(c, cudaMemcpyAsync(c->rays.p, rays, n * sizeof(rl_ray), cudaMemcpyHostToDevice, s));
    std::vector<int> refs;  // must outlive the asynchronous copy: read_counters below synchronises
    const int* d_self = nullptr;
    if (self_nodes) {
        refs.resize(n);
        for (uint64_t i = 0; i < n; i++) {
            const int nd = self_nodes[i];
            refs[i] = (nd >= 0 && nd < (int)c->node_ref.size()) ? c->node_ref[nd] : -1;
        }
        CK(c, c->self_refs.reserve(n * sizeof(int)));
        CK(c, cudaMemcpyAsync(c->self_refs.p, refs.data(), n * sizeof(int), cudaMemcpyHostToDevice, s));
        d_self = c->self_refs.as<int>();
    }
    CK(c, reset_stats(c, s));
    if (c->ds.flavor == RL_FLAVOR_RTC)
        CK(c, launch_rtc_trace(c->ds, c->rays.as<rl_ray>(), n, c->hits.as<rl_hit>(), c->counters.as<Counters>(), c->instrumented, s));
    else
        CK(c, launch_ow_trace(c->ds, c->rays.as<rl_ray>(), d_self, n, c->hits.as<rl_hit>(), c->queue.as<unsigned long long>(),
                              c->counters.as<Counters>(), c->instrumented, c->sm_count, s, c->tune));
    CK(c, cudaMemcpyAsync(out, c->hits.p, n * sizeof(rl_hit), cudaMemcpyDeviceToHost, s));
    return read_counters(c, s, nullptr);
}

int rl_trace_batch(rl_ctx* c, const rl_ray* rays, uint64_t n, rl_hit* out) { return rl_trace_batch_ex(c, rays, nullptr, n, out); }

int rl_render_rtc_device(rl_ctx* c, const rl_rtc_camera* cam, uint32_t aa, const rl_job* jobs, int32_t n_jobs,
                         void* d_out_rgb, void* stream, rl_stats* stats) {
    if (!c || !cam || !d_out_rgb) return RL_E_INVALID;
    if (!c->has_scene || c->ds.flavor != RL_FLAVOR_RTC) return fail(c, RL_E_NO_SCENE, "no RTC scene uploaded");
    if (aa < 1) return fail(c, RL_E_INVALID, "anti_aliasing_samples must be >= 1");
    if (cam->hsize < 1 || cam->vsize < 1) return fail(c, RL_E_INVALID, "empty image");
    double inv[12];
    std::string err;
    if (!invert_affine_4x4(cam->transform, inv, &err)) return fail(c, RL_E_INVALID, err);
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    JobTable jt;
    int rc = make_job_table(c, jobs, n_jobs, cam->hsize, cam->vsize, 1, false, s, &jt);
    if (rc != RL_OK) return rc;
    if (!stats) {
        // asynchronous mode: launch and return; overflow stays sticky in the counters until rl_synchronize
        CK(c, launch_rtc_render(c->ds, cam, inv, aa, jt, (float*)d_out_rgb, c->counters.as<Counters>(), false, s));
        return RL_OK;
    }
    CK(c, reset_stats(c, s));
    CK(c, cudaEventRecord(c->ev0, s));
    CK(c, launch_rtc_render(c->ds, cam, inv, aa, jt, (float*)d_out_rgb, c->counters.as<Counters>(), c->instrumented, s));
    CK(c, cudaEventRecord(c->ev1, s));
    rl_stats st{};
    rc = read_counters(c, s, &st);
    float ms = 0.0f;
    CK(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    st.kernel_ms = ms;
    st.upload_ms = c->upload_ms;
    st.kernel_launches = jt.n_items > 0 ? 1 : 0;
    long long px = 0;
    for (int i = 0; i < n_jobs; i++) px += (long long)(jobs[i].x1 - jobs[i].x0) * (jobs[i].y1 - jobs[i].y0);
    st.samples = (uint64_t)px * aa * aa;
    *stats = st;
    return rc;
}

// RTC on a multi-GPU ctx: the frame is cut into G x JOBS_INLINE interleaved row bands; GPU g renders bands g, g + G, ...
// with ONE launch and copies them (f32, or 8-bit after the device-side encoder) straight into the caller's buffer, so
// there is no gather at all.  Pixels are independent and every GPU holds the same scene: bit-identical to one GPU.
static int group_render_rtc(rl_ctx* c, const rl_rtc_camera* cam, uint32_t aa, float* out_rgb, uint8_t* out_rgb8, rl_stats* stats) {
    const int W = cam->hsize, H = cam->vsize, G = (int)c->group.size();
    int rows = (H + G * JOBS_INLINE - 1) / (G * JOBS_INLINE);
    rows = (rows + 3) / 4 * 4;  // keep the 8x4 micro-tiles whole
    const size_t row_px = (size_t)W * 3;
    rl_stats total{};
    std::vector<std::vector<rl_job>> bands(G);
    for (int y0 = 0, k = 0; y0 < H; y0 += rows, k++) bands[k % G].push_back(rl_job{0, y0, W, y0 + rows < H ? y0 + rows : H, 0, 1});
    for (int g = 0; g < G; g++) {
        rl_ctx* m = c->group[g];
        if (bands[g].empty()) continue;
        CK(c, cudaSetDevice(m->device));
        CK(c, m->frame.reserve(row_px * H * sizeof(float)));
        if (out_rgb8) CK(c, m->frame8.reserve(row_px * H));
        int rc = rl_render_rtc_device(m, cam, aa, bands[g].data(), (int)bands[g].size(), m->frame.p, nullptr, nullptr);  // async
        if (rc != RL_OK) return fail(c, rc, m->error);
        for (const rl_job& j : bands[g]) {
            const size_t off = row_px * j.y0, n = row_px * (size_t)(j.y1 - j.y0);
            if (out_rgb8) {
                CK(c, launch_encode_rtc_u8(m->frame.as<float>() + off, m->frame8.as<uint8_t>() + off, n, m->stream));
                CK(c, cudaMemcpyAsync(out_rgb8 + off, m->frame8.as<uint8_t>() + off, n, cudaMemcpyDeviceToHost, m->stream));
                total.kernel_launches += 1;
            } else {
                CK(c, cudaMemcpyAsync(out_rgb + off, m->frame.as<float>() + off, n * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
            }
        }
        total.kernel_launches += 1;
    }
    for (rl_ctx* m : c->group) {
        int rc = rl_synchronize(m);
        if (rc != RL_OK) return fail(c, rc, m->error);
    }
    CK(c, cudaSetDevice(c->device));
    total.samples = (uint64_t)W * H * aa * aa;
    total.upload_ms = c->upload_ms;
    if (stats) *stats = total;
    return RL_OK;
}

int rl_render_rtc(rl_ctx* c, const rl_rtc_camera* cam, uint32_t aa, float* out_rgb, rl_stats* stats) {
    if (!c || !cam || !out_rgb) return RL_E_INVALID;
    if (cam->hsize < 1 || cam->vsize < 1) return fail(c, RL_E_INVALID, "empty image");
    if (c->group.size() > 1) return group_render_rtc(c, cam, aa, out_rgb, nullptr, stats);
    size_t bytes = (size_t)cam->hsize * cam->vsize * 3 * sizeof(float);
    CK(c, cudaSetDevice(c->device));
    CK(c, c->frame.reserve(bytes));
    rl_job job{0, 0, cam->hsize, cam->vsize, 0, 1};
    rl_stats st{};
    int rc = rl_render_rtc_device(c, cam, aa, &job, 1, c->frame.p, nullptr, &st);
    if (stats) *stats = st;
    if (rc != RL_OK) return rc;
    CK(c, cudaMemcpyAsync(out_rgb, c->frame.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return RL_OK;
}

int rl_render_rtc_u8(rl_ctx* c, const rl_rtc_camera* cam, uint32_t aa, uint8_t* out_rgb8, rl_stats* stats) {
    if (!c || !cam || !out_rgb8) return RL_E_INVALID;
    if (cam->hsize < 1 || cam->vsize < 1) return fail(c, RL_E_INVALID, "empty image");
    if (c->group.size() > 1) return group_render_rtc(c, cam, aa, nullptr, out_rgb8, stats);
    size_t n = (size_t)cam->hsize * cam->vsize * 3;
    CK(c, cudaSetDevice(c->device));
    CK(c, c->frame.reserve(n * sizeof(float)));
    CK(c, c->frame8.reserve(n));
    rl_job job{0, 0, cam->hsize, cam->vsize, 0, 1};
    rl_stats st{};
    int rc = rl_render_rtc_device(c, cam, aa, &job, 1, c->frame.p, nullptr, &st);
    if (rc != RL_OK) return rc;
    CK(c, launch_encode_rtc_u8(c->frame.as<float>(), c->frame8.as<uint8_t>(), n, c->stream));
    CK(c, cudaMemcpyAsync(out_rgb8, c->frame8.p, n, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    st.kernel_launches += 1;
    if (stats) *stats = st;
    return RL_OK;
}

int rl_ow_image_height(const rl_ow_camera* cam) { return cam ? ow_image_height(cam) : 0; }
int rl_ow_num_chunks(const rl_ow_camera* cam) { return cam ? ow_num_chunks(cam->samples_per_pixel) : 0; }

// Camera::new's preconditions (OW/src/camera.rs:72-118), shared by every OW render entry point
static int check_ow_camera(rl_ctx* c, const rl_ow_camera* cam) {
    if (cam->image_width < 1 || cam->image_width > 65535 || cam->samples_per_pixel < 1 || !(cam->aspect_ratio > 0.0))
        return fail(c, RL_E_INVALID, "bad camera parameters");
    if (ow_image_height(cam) > 65535) return fail(c, RL_E_INVALID, "image taller than 65535 rows");
    if (cam->max_depth < 0) return fail(c, RL_E_INVALID, "max_depth must be >= 0");
    double dx = cam->lookfrom[0] - cam->lookat[0], dy = cam->lookfrom[1] - cam->lookat[1], dz = cam->lookfrom[2] - cam->lookat[2];
    if (dx * dx + dy * dy + dz * dz <= 1e-16) return fail(c, RL_E_INVALID, "cannot normalize vector with magnitude 0");
    // vup parallel to the view direction: normalize(vup x w) of a zero vector (camera.rs:89)
    const double* v = cam->vup;
    double cx = v[1] * dz - v[2] * dy, cy = v[2] * dx - v[0] * dz, cz = v[0] * dy - v[1] * dx;
    if (cx * cx + cy * cy + cz * cz <= 1e-24 * (dx * dx + dy * dy + dz * dz))
        return fail(c, RL_E_INVALID, "cannot normalize vector with magnitude 0");
    return RL_OK;
}

int rl_render_ow_device(rl_ctx* c, const rl_ow_camera* cam, uint32_t first_sample, const rl_job* jobs, int32_t n_jobs,
                        void* d_partial, void* stream, rl_stats* stats) {
    if (!c || !cam || !d_partial) return RL_E_INVALID;
    if (!c->has_scene || c->ds.flavor != RL_FLAVOR_OW) return fail(c, RL_E_NO_SCENE, "no OW scene uploaded");
    if (int rcc = check_ow_camera(c, cam)) return rcc;
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    int H = ow_image_height(cam), nc = ow_num_chunks(cam->samples_per_pixel);
    JobTable jt;
    int rc = make_job_table(c, jobs, n_jobs, cam->image_width, H, nc, true, s, &jt);
    if (rc != RL_OK) return rc;
    if (!stats) {
        // asynchronous mode (multi-GPU tile loop): launch and return; errors surface at rl_synchronize
        CK(c, launch_ow_render(c->ds, cam, first_sample, jt, (float*)d_partial, c->queue.as<unsigned long long>(),
                               c->counters.as<Counters>(), false, c->sm_count, s, false, c->tune));
        return RL_OK;
    }
    CK(c, reset_stats(c, s));
    int wf_launches = 0;
    if (c->tune.variant == 7) {  // the global-wavefront A/B variant: path state in global memory
        int S = c->tune.slots >= 1024 ? c->tune.slots : (1 << 19);
        S = (S + 255) / 256 * 256;
        CK(c, c->wf_state.reserve(sizeof(float) * (size_t)OW_WF_WORDS * S));
        CK(c, c->wf_rays.reserve(sizeof(int) * (size_t)S));
        CK(c, c->wf_ctr.reserve(256));
        WavefrontBuffers wb{c->wf_state.as<float>(), c->wf_rays.as<int>(), c->wf_ctr.p, S};
        CK(c, cudaMemsetAsync(c->queue.p, 0, 2 * sizeof(unsigned long long), s));
        CK(c, cudaEventRecord(c->ev0, s));
        CK(c, launch_ow_wavefront(c->ds, cam, first_sample, jt, (float*)d_partial, c->queue.as<unsigned long long>(),
                                  c->counters.as<Counters>(), wb, c->sm_count, s, c->tune, &wf_launches));
    } else {
        CK(c, cudaEventRecord(c->ev0, s));
        CK(c, launch_ow_render(c->ds, cam, first_sample, jt, (float*)d_partial, c->queue.as<unsigned long long>(),
                               c->counters.as<Counters>(), c->instrumented, c->sm_count, s, false, c->tune));
    }
    CK(c, cudaEventRecord(c->ev1, s));
    rl_stats st{};
    rc = read_counters(c, s, &st);
    float ms = 0.0f;
    CK(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    st.kernel_ms = ms;
    st.upload_ms = c->upload_ms;
    st.kernel_launches = wf_launches ? wf_launches : (jt.n_items > 0 ? 1 : 0);
    uint64_t samples = 0;
    for (int i = 0; i < n_jobs; i++) {
        uint64_t px = (uint64_t)(jobs[i].x1 - jobs[i].x0) * (uint64_t)(jobs[i].y1 - jobs[i].y0);
        for (int ck = jobs[i].chunk_begin; ck < jobs[i].chunk_end; ck++) {
            int s0, s1;
            ow_chunk_range(cam->samples_per_pixel, nc, ck, &s0, &s1);
            samples += px * (uint64_t)(s1 - s0);
        }
    }
    st.samples = samples;
    *stats = st;
    return rc;
}

// ---- cross-GPU queue, one process per GPU (CUDA IPC) ---------------------------------------------------------------
// control block: RL_QUEUE_SLOTS x {work counter, completion counter}, 16 bytes per slot
static unsigned long long* queue_slot(rl_ctx* c, int slot) { return c->shared_queue + 2 * slot; }
static bool bad_slot(int slot) { return slot < 0 || slot >= RL_QUEUE_SLOTS; }

int rl_queue_export(rl_ctx* c, void* handle64) {
    if (!c || !handle64) return RL_E_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CK(c, cudaSetDevice(c->device));
    if (c->shared_queue_imported && c->shared_queue) cudaIpcCloseMemHandle(c->shared_queue);
    if (!c->shared_queue_own.p) {
        CK(c, c->shared_queue_own.reserve(256));
        CK(c, cudaMemset(c->shared_queue_own.p, 0, 256));
    }
    cudaIpcMemHandle_t h;
    CK(c, cudaIpcGetMemHandle(&h, c->shared_queue_own.p));
    memcpy(handle64, &h, sizeof(h));
    c->shared_queue = c->shared_queue_own.as<unsigned long long>();
    c->shared_queue_imported = false;
    return RL_OK;
}

int rl_queue_import(rl_ctx* c, const void* handle64) {
    if (!c || !handle64) return RL_E_INVALID;
    CK(c, cudaSetDevice(c->device));
    if (c->shared_queue_imported && c->shared_queue) cudaIpcCloseMemHandle(c->shared_queue);  // the previous mapping
    c->shared_queue = nullptr;
    c->shared_queue_imported = false;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* p = nullptr;
    CK(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->shared_queue = (unsigned long long*)p;
    c->shared_queue_imported = true;
    return RL_OK;
}

int rl_partial_export(rl_ctx* c, uint64_t bytes_per_slot, int32_t n_slots, void* handle64) {
    if (!c || !handle64 || bytes_per_slot == 0 || n_slots < 1 || n_slots > RL_QUEUE_SLOTS) return RL_E_INVALID;
    CK(c, cudaSetDevice(c->device));
    if (c->shared_partial_imported && c->shared_partial) cudaIpcCloseMemHandle(c->shared_partial);
    bytes_per_slot = (bytes_per_slot + 255) / 256 * 256;
    CK(c, c->shared_partial_own.reserve(bytes_per_slot * (uint64_t)n_slots));
    cudaIpcMemHandle_t h;
    CK(c, cudaIpcGetMemHandle(&h, c->shared_partial_own.p));
    memcpy(handle64, &h, sizeof(h));
    c->shared_partial = c->shared_partial_own.as<float>();
    c->shared_partial_bytes = bytes_per_slot;
    c->shared_partial_slots = n_slots;
    c->shared_partial_imported = false;
    return RL_OK;
}

int rl_partial_import(rl_ctx* c, const void* handle64, uint64_t bytes_per_slot, int32_t n_slots) {
    if (!c || !handle64 || bytes_per_slot == 0 || n_slots < 1 || n_slots > RL_QUEUE_SLOTS) return RL_E_INVALID;
    CK(c, cudaSetDevice(c->device));
    if (c->shared_partial_imported && c->shared_partial) cudaIpcCloseMemHandle(c->shared_partial);
    c->shared_partial = nullptr;
    c->shared_partial_bytes = 0;
    c->shared_partial_slots = 0;
    c->shared_partial_imported = false;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* p = nullptr;
    CK(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->shared_partial = (float*)p;
    c->shared_partial_bytes = (bytes_per_slot + 255) / 256 * 256;  // the exporter's geometry: every launch is checked against it
    c->shared_partial_slots = n_slots;
    c->shared_partial_imported = true;
    return RL_OK;
}

int rl_queue_reset(rl_ctx* c, void* stream, int32_t slot) {
    if (!c || bad_slot(slot)) return RL_E_INVALID;
    if (!c->shared_queue || c->shared_queue_imported) return fail(c, RL_E_INVALID, "only the exporting ctx resets the shared queue");
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    CK(c, cudaMemsetAsync(queue_slot(c, slot), 0, 2 * sizeof(unsigned long long), s));  // work counter + completion counter
    return RL_OK;
}

int rl_queue_completed(rl_ctx* c, void* stream, int32_t slot, uint64_t* items) {
    if (!c || !items || bad_slot(slot)) return RL_E_INVALID;
    if (!c->shared_queue || c->shared_queue_imported) return fail(c, RL_E_INVALID, "only the exporting ctx reads the completion counter");
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    unsigned long long v = 0;
    CK(c, cudaMemcpyAsync(&v, queue_slot(c, slot) + 1, sizeof(v), cudaMemcpyDeviceToHost, s));
    CK(c, cudaStreamSynchronize(s));
    *items = v;
    return RL_OK;
}

int64_t rl_ow_job_items(const rl_ow_camera* cam, const rl_job* jobs, int32_t n_jobs) {
    if (!cam || n_jobs < 0 || (n_jobs > 0 && !jobs)) return -1;
    int64_t items = 0;
    for (int i = 0; i < n_jobs; i++)
        items += padded_pixels(jobs[i].x1 - jobs[i].x0, jobs[i].y1 - jobs[i].y0) * (int64_t)(jobs[i].chunk_end - jobs[i].chunk_begin);
    return items;
}

// the partial-sum buffer of `slot` behind the shared mapping, or an error when this camera does not fit into it
static int shared_partial_of(rl_ctx* c, const rl_ow_camera* cam, int slot, float** out) {
    if (!c->shared_partial) return fail(c, RL_E_INVALID, "no partial buffer: pass one or call rl_partial_export / rl_partial_import");
    if (slot >= c->shared_partial_slots) return fail(c, RL_E_INVALID, "partial-sum slot was not exported");
    const uint64_t need = (uint64_t)cam->image_width * ow_image_height(cam) * ow_num_chunks(cam->samples_per_pixel) * 16;
    if (need > c->shared_partial_bytes)
        return fail(c, RL_E_INVALID, "the shared partial-sum buffer is smaller than width x height x chunks x 16 bytes");
    *out = reinterpret_cast<float*>(reinterpret_cast<char*>(c->shared_partial) + (size_t)slot * c->shared_partial_bytes);
    return RL_OK;
}

int rl_render_ow_shared(rl_ctx* c, const rl_ow_camera* cam, uint32_t first_sample, const rl_job* jobs, int32_t n_jobs,
                        void* d_partial, void* stream, int32_t slot) {
    if (!c || !cam || bad_slot(slot)) return RL_E_INVALID;
    if (!c->has_scene || c->ds.flavor != RL_FLAVOR_OW) return fail(c, RL_E_NO_SCENE, "no OW scene uploaded");
    if (!c->shared_queue) return fail(c, RL_E_INVALID, "no shared queue: call rl_queue_export / rl_queue_import first");
    if (int rcc = check_ow_camera(c, cam)) return rcc;
    if (!d_partial) {  // fused gather: store straight into the owner's buffer — which must be large enough for THIS camera
        float* sp = nullptr;
        if (int rcc = shared_partial_of(c, cam, slot, &sp)) return rcc;
        d_partial = sp;
    }
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    int H = ow_image_height(cam), nc = ow_num_chunks(cam->samples_per_pixel);
    JobTable jt;
    int rc = make_job_table(c, jobs, n_jobs, cam->image_width, H, nc, true, s, &jt);
    if (rc != RL_OK) return rc;
    CK(c, launch_ow_render(c->ds, cam, first_sample, jt, (float*)d_partial, queue_slot(c, slot), c->counters.as<Counters>(),
                           false, c->sm_count, s, true, c->tune));
    return RL_OK;
}

int rl_ow_reduce_device(rl_ctx* c, const rl_ow_camera* cam, const void* d_partial, void* d_out, void* stream) {
    if (!c || !cam || !d_partial || !d_out) return RL_E_INVALID;
    if (int rcc = check_ow_camera(c, cam)) return rcc;
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    CK(c, launch_ow_reduce(cam, (const float*)d_partial, (float*)d_out, s));
    return RL_OK;
}

int rl_ow_reduce_shared(rl_ctx* c, const rl_ow_camera* cam, int32_t slot, void* d_out, void* stream) {
    if (!c || !cam || !d_out || bad_slot(slot)) return RL_E_INVALID;
    if (int rcc = check_ow_camera(c, cam)) return rcc;
    float* sp = nullptr;
    if (int rcc = shared_partial_of(c, cam, slot, &sp)) return rcc;
    return rl_ow_reduce_device(c, cam, sp, d_out, stream);
}

// ---- one process, several GPUs (rl_create_multi) ---------------------------------------------------------------------
// OW: ONE persistent launch per GPU, all popping the leader's work counter and storing into the leader's partial-sum
// buffer over NVLink peer memory; CUDA events order "counter reset -> every launch" and "every kernel -> fold".
static int group_render_ow_partials(rl_ctx* c, const rl_ow_camera* cam, uint32_t first_sample, int H, int nc, rl_stats* st) {
    rl_job job{0, 0, cam->image_width, H, 0, nc};
    const int64_t items = rl_ow_job_items(cam, &job, 1);
    CK(c, cudaSetDevice(c->device));
    CK(c, reset_stats(c, c->stream));
    CK(c, cudaMemsetAsync(c->queue.p, 0, 2 * sizeof(unsigned long long), c->stream));
    CK(c, cudaEventRecord(c->ev0, c->stream));
    CK(c, cudaEventRecord(c->ev_go, c->stream));
    for (rl_ctx* g : c->group) {
        CK(c, cudaSetDevice(g->device));
        CK(c, cudaStreamWaitEvent(g->stream, c->ev_go, 0));
        JobTable jt;
        int rc = make_job_table(g, &job, 1, cam->image_width, H, nc, true, g->stream, &jt);
        if (rc != RL_OK) return rc;
        CK(c, launch_ow_render(g->ds, cam, first_sample, jt, c->partial.as<float>(), c->queue.as<unsigned long long>(),
                               g->counters.as<Counters>(), false, g->sm_count, g->stream, true, c->tune));
        CK(c, cudaEventRecord(g->ev_done, g->stream));
    }
    CK(c, cudaSetDevice(c->device));
    for (rl_ctx* g : c->group) CK(c, cudaStreamWaitEvent(c->stream, g->ev_done, 0));
    CK(c, cudaEventRecord(c->ev1, c->stream));
    unsigned long long done = 0;
    CK(c, cudaMemcpyAsync(&done, c->queue.as<unsigned long long>() + 1, sizeof(done), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    float ms = 0.0f;
    CK(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    st->kernel_ms = ms;
    st->upload_ms = c->upload_ms;
    st->kernel_launches = (int)c->group.size();
    st->samples = (uint64_t)cam->image_width * H * (uint64_t)cam->samples_per_pixel;
    for (rl_ctx* g : c->group) {  // overflow on ANY GPU fails the render
        Counters h;
        CK(c, cudaSetDevice(g->device));
        CK(c, cudaMemcpy(&h, g->counters.p, sizeof(h), cudaMemcpyDeviceToHost));
        if (h.overflow) {
            cudaMemset(&g->counters.as<Counters>()->overflow, 0, sizeof(h.overflow));
            st->overflow += h.overflow;
        }
    }
    CK(c, cudaSetDevice(c->device));
    if (st->overflow) return fail(c, RL_E_OVERFLOW, "a traversal stack / work list overflowed on the device");
    if ((int64_t)done != items) return fail(c, RL_E_CUDA, "multi-GPU render incomplete: a GPU stored fewer items than it popped");
    return RL_OK;
}

// renders into c->frame (sums); shared by rl_render_ow and rl_render_ow_u8
static int render_ow_to_frame(rl_ctx* c, const rl_ow_camera* cam, uint32_t first_sample, rl_stats* st, size_t* frame_bytes) {
    if (int rcc = check_ow_camera(c, cam)) return rcc;
    if (!c->has_scene || c->ds.flavor != RL_FLAVOR_OW) return fail(c, RL_E_NO_SCENE, "no OW scene uploaded");
    int H = ow_image_height(cam), nc = ow_num_chunks(cam->samples_per_pixel);
    size_t frame = (size_t)cam->image_width * H * 3 * sizeof(float);
    CK(c, cudaSetDevice(c->device));
    CK(c, c->partial.reserve(frame / 3 * 4 * nc));  // [n_chunks][H][W] float4
    CK(c, c->frame.reserve(frame));
    int rc;
    if (c->group.size() > 1) {
        rc = group_render_ow_partials(c, cam, first_sample, H, nc, st);
    } else {
        rl_job job{0, 0, cam->image_width, H, 0, nc};
        rc = rl_render_ow_device(c, cam, first_sample, &job, 1, c->partial.p, nullptr, st);
    }
    if (rc != RL_OK) return rc;
    CK(c, cudaEventRecord(c->ev0, c->stream));
    rc = rl_ow_reduce_device(c, cam, c->partial.p, c->frame.p, nullptr);
    if (rc != RL_OK) return rc;
    CK(c, cudaEventRecord(c->ev1, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    float ms = 0.0f;
    CK(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    st->kernel_ms += ms;
    st->kernel_launches += 1;
    *frame_bytes = frame;
    return RL_OK;
}

int rl_render_ow(rl_ctx* c, const rl_ow_camera* cam, uint32_t first_sample, float* out_rgb_sum, rl_stats* stats) {
    if (!c || !cam || !out_rgb_sum) return RL_E_INVALID;
    rl_stats st{};
    size_t frame = 0;
    int rc = render_ow_to_frame(c, cam, first_sample, &st, &frame);
    if (rc != RL_OK) return rc;
    CK(c, cudaMemcpyAsync(out_rgb_sum, c->frame.p, frame, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    if (stats) *stats = st;
    return RL_OK;
}

int rl_render_ow_u8(rl_ctx* c, const rl_ow_camera* cam, uint32_t first_sample, uint8_t* out_rgb8, rl_stats* stats) {
    if (!c || !cam || !out_rgb8) return RL_E_INVALID;
    rl_stats st{};
    size_t frame = 0;
    int rc = render_ow_to_frame(c, cam, first_sample, &st, &frame);
    if (rc != RL_OK) return rc;
    size_t n = frame / sizeof(float);
    CK(c, c->frame8.reserve(n));
    // Canvas.samples of a fresh render = samples_per_pixel; a resumed render is merged on the host (sums), not here
    CK(c, launch_encode_ow_u8(c->frame.as<float>(), c->frame8.as<uint8_t>(), n, cam->samples_per_pixel, c->stream));
    CK(c, cudaMemcpyAsync(out_rgb8, c->frame8.p, n, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    st.kernel_launches += 1;
    if (stats) *stats = st;
    return RL_OK;
}

}  // extern "C"
