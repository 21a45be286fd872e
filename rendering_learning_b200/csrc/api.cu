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
    DevBuf prims, tri_verts, tri_shade, xforms, spheres, quads, sphere_node, quad_node, materials, textures,
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
    // cross-GPU queue (CUDA IPC): the owner allocates it, peers map it
    DevBuf shared_queue_own, shared_partial_own;
    float* shared_partial = nullptr;
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
        (e = c->counters.reserve(sizeof(Counters))) != cudaSuccess || (e = c->queue.reserve(sizeof(unsigned long long))) != cudaSuccess) {
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
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->shared_queue_imported && c->shared_queue) cudaIpcCloseMemHandle(c->shared_queue);
    if (c->shared_partial_imported && c->shared_partial) cudaIpcCloseMemHandle(c->shared_partial);
    c->shared_partial_own.release();
    c->shared_queue_own.release();
    DevBuf* all[] = {&c->prims, &c->tri_verts, &c->tri_shade, &c->xforms, &c->spheres, &c->quads, &c->sphere_node,
                     &c->quad_node, &c->materials, &c->textures, &c->images, &c->lights, &c->nodes, &c->big_refs, &c->csg, &c->media, &c->medium_refs, &c->perlin_vec, &c->perlin_perm, &c->bvh_aabb,
                     &c->bvh_ref, &c->bvh_node_id, &c->bounds, &c->keys, &c->sorted_prim, &c->keys_tmp, &c->idx_tmp,
                     &c->left, &c->right, &c->parent, &c->node_aabb, &c->lbvh_counters, &c->counters, &c->queue,
                     &c->jobs, &c->prefix, &c->frame, &c->frame8, &c->partial, &c->rays, &c->hits};
    for (DevBuf* b : all) b->release();
    for (DevBuf& b : c->image_texels) b.release();
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

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
    Counters h;
    CK(c, cudaMemcpy(&h, c->counters.p, sizeof(h), cudaMemcpyDeviceToHost));
    if (h.overflow) {
        cudaMemset(c->counters.p, 0, sizeof(Counters));
        c->error = "a traversal stack / work list overflowed on the device";
        return RL_E_OVERFLOW;
    }
    return RL_OK;
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

int rl_scene_upload(rl_ctx* c, const rl_scene_desc* scene) {
    if (!c) return RL_E_INVALID;
    c->has_scene = false;
    FlatScene fs;
    std::string err;
    int rc = flatten_scene(scene, &fs, &err);
    if (rc != RL_OK) return fail(c, rc, err);
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
    d.xforms = c->xforms.as<Xform>();
    d.spheres = c->spheres.as<OwSphere>();
    d.quads = c->quads.as<OwQuad>();
    d.sphere_node = c->sphere_node.as<int>();
    d.quad_node = c->quad_node.as<int>();
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

    c->info = scene_info_of(fs);
    c->has_scene = true;
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
    jt->inline_jobs = 0;
    CK(c, c->jobs.reserve(sizeof(rl_job) * (size_t)n_jobs));
    CK(c, c->prefix.reserve(sizeof(long long) * (size_t)(n_jobs + 1)));
    CK(c, cudaStreamSynchronize(s));  // a previous launch on this stream may still read the old table
    CK(c, cudaMemcpyAsync(c->jobs.p, jobs, sizeof(rl_job) * (size_t)n_jobs, cudaMemcpyHostToDevice, s));
    CK(c, cudaMemcpyAsync(c->prefix.p, prefix.data(), sizeof(long long) * (size_t)(n_jobs + 1), cudaMemcpyHostToDevice, s));
    CK(c, cudaStreamSynchronize(s));  // `prefix` is a local
    jt->jobs = c->jobs.as<rl_job>();
    jt->prefix = c->prefix.as<long long>();
    return RL_OK;
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
    if (h.overflow) return fail(c, RL_E_OVERFLOW, "a traversal stack / work list overflowed on the device");
    return RL_OK;
}

extern "C" {

int rl_trace_batch(rl_ctx* c, const rl_ray* rays, uint64_t n, rl_hit* out) {
    if (!c || (n > 0 && (!rays || !out))) return RL_E_INVALID;
    if (!c->has_scene) return fail(c, RL_E_NO_SCENE, "no scene uploaded");
    if (n == 0) return RL_OK;
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    CK(c, c->rays.reserve(n * sizeof(rl_ray)));
    CK(c, c->hits.reserve(n * sizeof(rl_hit)));
    CK(c, cudaMemcpyAsync(c->rays.p, rays, n * sizeof(rl_ray), cudaMemcpyHostToDevice, s));
    CK(c, cudaMemsetAsync(c->counters.p, 0, sizeof(Counters), s));
    if (c->ds.flavor == RL_FLAVOR_RTC)
        CK(c, launch_rtc_trace(c->ds, c->rays.as<rl_ray>(), n, c->hits.as<rl_hit>(), c->counters.as<Counters>(), c->instrumented, s));
    else
        CK(c, launch_ow_trace(c->ds, c->rays.as<rl_ray>(), n, c->hits.as<rl_hit>(), c->counters.as<Counters>(), c->instrumented, s));
    CK(c, cudaMemcpyAsync(out, c->hits.p, n * sizeof(rl_hit), cudaMemcpyDeviceToHost, s));
    return read_counters(c, s, nullptr);
}

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
    CK(c, cudaMemsetAsync(c->counters.p, 0, sizeof(Counters), s));
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

int rl_render_rtc(rl_ctx* c, const rl_rtc_camera* cam, uint32_t aa, float* out_rgb, rl_stats* stats) {
    if (!c || !cam || !out_rgb) return RL_E_INVALID;
    if (cam->hsize < 1 || cam->vsize < 1) return fail(c, RL_E_INVALID, "empty image");
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

int rl_render_ow_device(rl_ctx* c, const rl_ow_camera* cam, uint32_t first_sample, const rl_job* jobs, int32_t n_jobs,
                        void* d_partial, void* stream, rl_stats* stats) {
    if (!c || !cam || !d_partial) return RL_E_INVALID;
    if (!c->has_scene || c->ds.flavor != RL_FLAVOR_OW) return fail(c, RL_E_NO_SCENE, "no OW scene uploaded");
    if (cam->image_width < 1 || cam->samples_per_pixel < 1 || !(cam->aspect_ratio > 0.0))
        return fail(c, RL_E_INVALID, "bad camera parameters");
    if (cam->max_depth < 0) return fail(c, RL_E_INVALID, "max_depth must be >= 0");
    double dx = cam->lookfrom[0] - cam->lookat[0], dy = cam->lookfrom[1] - cam->lookat[1], dz = cam->lookfrom[2] - cam->lookat[2];
    if (dx * dx + dy * dy + dz * dz <= 1e-16) return fail(c, RL_E_INVALID, "cannot normalize vector with magnitude 0");
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    int H = ow_image_height(cam), nc = ow_num_chunks(cam->samples_per_pixel);
    JobTable jt;
    int rc = make_job_table(c, jobs, n_jobs, cam->image_width, H, nc, true, s, &jt);
    if (rc != RL_OK) return rc;
    if (!stats) {
        // asynchronous mode (multi-GPU tile loop): launch and return; errors surface at rl_synchronize
        CK(c, launch_ow_render(c->ds, cam, first_sample, jt, (float*)d_partial, c->queue.as<unsigned long long>(),
                               c->counters.as<Counters>(), false, c->sm_count, s));
        return RL_OK;
    }
    CK(c, cudaMemsetAsync(c->counters.p, 0, sizeof(Counters), s));
    CK(c, cudaEventRecord(c->ev0, s));
    CK(c, launch_ow_render(c->ds, cam, first_sample, jt, (float*)d_partial, c->queue.as<unsigned long long>(),
                           c->counters.as<Counters>(), c->instrumented, c->sm_count, s));
    CK(c, cudaEventRecord(c->ev1, s));
    rl_stats st{};
    rc = read_counters(c, s, &st);
    float ms = 0.0f;
    CK(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    st.kernel_ms = ms;
    st.upload_ms = c->upload_ms;
    st.kernel_launches = jt.n_items > 0 ? 1 : 0;
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

int rl_queue_export(rl_ctx* c, void* handle64) {
    if (!c || !handle64) return RL_E_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CK(c, cudaSetDevice(c->device));
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
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* p = nullptr;
    CK(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->shared_queue = (unsigned long long*)p;
    c->shared_queue_imported = true;
    return RL_OK;
}

int rl_partial_export(rl_ctx* c, uint64_t bytes, void* handle64) {
    if (!c || !handle64 || bytes == 0) return RL_E_INVALID;
    CK(c, cudaSetDevice(c->device));
    CK(c, c->shared_partial_own.reserve(bytes));
    cudaIpcMemHandle_t h;
    CK(c, cudaIpcGetMemHandle(&h, c->shared_partial_own.p));
    memcpy(handle64, &h, sizeof(h));
    c->shared_partial = c->shared_partial_own.as<float>();
    c->shared_partial_imported = false;
    return RL_OK;
}

int rl_partial_import(rl_ctx* c, const void* handle64) {
    if (!c || !handle64) return RL_E_INVALID;
    CK(c, cudaSetDevice(c->device));
    if (c->shared_partial_imported && c->shared_partial) cudaIpcCloseMemHandle(c->shared_partial);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* p = nullptr;
    CK(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->shared_partial = (float*)p;
    c->shared_partial_imported = true;
    return RL_OK;
}

int rl_queue_reset(rl_ctx* c, void* stream) {
    if (!c) return RL_E_INVALID;
    if (!c->shared_queue || c->shared_queue_imported) return fail(c, RL_E_INVALID, "only the exporting ctx resets the shared queue");
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    CK(c, cudaMemsetAsync(c->shared_queue, 0, sizeof(unsigned long long), s));
    return RL_OK;
}

int rl_render_ow_shared(rl_ctx* c, const rl_ow_camera* cam, uint32_t first_sample, const rl_job* jobs, int32_t n_jobs,
                        void* d_partial, void* stream) {
    if (!c || !cam) return RL_E_INVALID;
    if (!d_partial) d_partial = c->shared_partial;  // fused gather: store straight into the owner's buffer
    if (!d_partial) return fail(c, RL_E_INVALID, "no partial buffer: pass one or call rl_partial_export / rl_partial_import");
    if (!c->has_scene || c->ds.flavor != RL_FLAVOR_OW) return fail(c, RL_E_NO_SCENE, "no OW scene uploaded");
    if (!c->shared_queue) return fail(c, RL_E_INVALID, "no shared queue: call rl_queue_export / rl_queue_import first");
    if (cam->image_width < 1 || cam->samples_per_pixel < 1 || !(cam->aspect_ratio > 0.0) || cam->max_depth < 0)
        return fail(c, RL_E_INVALID, "bad camera parameters");
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    int H = ow_image_height(cam), nc = ow_num_chunks(cam->samples_per_pixel);
    JobTable jt;
    int rc = make_job_table(c, jobs, n_jobs, cam->image_width, H, nc, true, s, &jt);
    if (rc != RL_OK) return rc;
    CK(c, launch_ow_render(c->ds, cam, first_sample, jt, (float*)d_partial, c->shared_queue, c->counters.as<Counters>(),
                           false, c->sm_count, s, true));
    return RL_OK;
}

int rl_ow_reduce_device(rl_ctx* c, const rl_ow_camera* cam, const void* d_partial, void* d_out, void* stream) {
    if (c && !d_partial) d_partial = c->shared_partial;
    if (!c || !cam || !d_partial || !d_out) return RL_E_INVALID;
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    CK(c, launch_ow_reduce(cam, (const float*)d_partial, (float*)d_out, s));
    return RL_OK;
}

// renders into c->frame (sums); shared by rl_render_ow and rl_render_ow_u8
static int render_ow_to_frame(rl_ctx* c, const rl_ow_camera* cam, uint32_t first_sample, rl_stats* st, size_t* frame_bytes) {
    if (cam->image_width < 1 || cam->samples_per_pixel < 1 || !(cam->aspect_ratio > 0.0))
        return fail(c, RL_E_INVALID, "bad camera parameters");
    int H = ow_image_height(cam), nc = ow_num_chunks(cam->samples_per_pixel);
    size_t frame = (size_t)cam->image_width * H * 3 * sizeof(float);
    CK(c, cudaSetDevice(c->device));
    CK(c, c->partial.reserve(frame / 3 * 4 * nc));  // [n_chunks][H][W] float4
    CK(c, c->frame.reserve(frame));
    rl_job job{0, 0, cam->image_width, H, 0, nc};
    int rc = rl_render_ow_device(c, cam, first_sample, &job, 1, c->partial.p, nullptr, st);
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
