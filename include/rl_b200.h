/*
 * rl_b200.h — C ABI of the B200-native per-pixel ray loop for marcantony/rendering-learning.
 *
 * This is the drop-in boundary.  The reference has no FFI layer; its operator boundary is the
 * Rust method pair
 *     ray-tracer-challenge/src/scene/camera.rs:93   Camera::render(&self, &World, &RenderOpts) -> Canvas
 *     ray-tracing-one-weekend/src/camera.rs:122     Camera::render<M,H>(&self, world: H) -> Canvas
 *     ray-tracing-one-weekend/src/camera.rs:136     Camera::render_from_checkpoint(world, &Canvas) -> Canvas
 * Every entry point below is what a Rust `extern "C"` block in a `rl-b200-sys` crate would bind
 * (see INTEGRATION.md for the binding and the `lower()` trait methods that emit these structs).
 *
 * Conventions
 *   - every call returns RL_OK (0) or a negative RL_E_* code; nothing throws or aborts across the ABI;
 *   - all pointers are caller-owned host memory that only has to stay valid for the duration of the call
 *     (except the `*_device` entry points, which take CUDA device pointers of the ctx's device);
 *   - structs are POD, little-endian, natural alignment; matrices are row-major;
 *   - a ctx is used by one caller thread at a time; a ctx from rl_create drives one GPU, a ctx from rl_create_multi
 *     drives several GPUs of one node from that one thread;
 *   - there is NO CPU fallback: rl_create fails with RL_E_NO_DEVICE when no sm_100 GPU is visible.
 *
 * The scene description is a direct, loss-free image of the reference's own scene types (an object
 * tree), NOT the flattened device layout.  The library's flattener (csrc/flatten.cpp) composes and
 * pre-inverts the transforms in f64 and lowers the tree to structure-of-arrays f32 buffers in HBM.
 */
#ifndef RL_B200_H
#define RL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RL_B200_ABI_VERSION 3

/* ---- status codes -------------------------------------------------------------------------- */
enum {
    RL_OK = 0,
    RL_E_INVALID = -1,     /* bad argument / malformed scene description                        */
    RL_E_NO_DEVICE = -2,   /* no sm_100 CUDA device (there is no CPU fallback)                  */
    RL_E_CUDA = -3,        /* a CUDA runtime call failed; see rl_last_error                      */
    RL_E_UNSUPPORTED = -4, /* scene uses a reference type the device path does not lower yet    */
    RL_E_NO_SCENE = -5,    /* render/trace called before rl_scene_upload                        */
    RL_E_OVERFLOW = -6     /* a fixed-capacity device structure overflowed (reported, not UB)   */
};

/* ---- scene description ----------------------------------------------------------------------- */
enum { RL_FLAVOR_RTC = 1, RL_FLAVOR_OW = 2 };

/* node kinds.  RTC_* mirror ray-tracer-challenge/src/scene/object/{sphere,plane,cube,cylinder,cone,
 * triangle,transformed,group,bounded,csg}.rs; OW_* mirror ray-tracing-one-weekend/src/hittable/
 * {sphere,flat/quad,flat/triangle,transform,translate}.rs, src/bvh.rs and the `[H]` slice impl
 * (src/hittable/mod.rs:86-111). */
enum {
    RL_RTC_SPHERE = 1,
    RL_RTC_PLANE = 2,
    RL_RTC_CUBE = 3,
    RL_RTC_CYLINDER = 4,   /* params: [minimum, maximum] (-inf/+inf when None); flags bit0 = closed */
    RL_RTC_CONE = 5,       /* same params as cylinder                                               */
    RL_RTC_TRIANGLE = 6,   /* params: p1[3] p2[3] p3[3] n1[3] n2[3] n3[3]; flags bit0 = smooth      */
    RL_RTC_TRANSFORMED = 7,/* params: m[16] forward 4x4; child_begin = child node id                */
    RL_RTC_GROUP = 8,      /* children[child_begin .. child_end) are node ids                        */
    RL_RTC_BOUNDED = 9,    /* child_begin = child node id                                            */
    RL_RTC_CSG = 10,       /* child_begin = left node id, child_end = right node id, flags = op      */
    RL_RTC_MESH = 11,      /* WavefrontObj::to_object() of the mesh rl_obj_parse left on the ctx: Bounded<Group<Triangle>>,
                            * every triangle with `material` (wavefront_obj.rs:164-166 uses Material::default())   */

    RL_OW_SPHERE = 32,     /* params: c1[3] c2[3] radius; flags bit0 = moving                        */
    RL_OW_QUAD = 33,       /* params: q[3] u[3] v[3]                                                 */
    RL_OW_TRIANGLE = 34,   /* params: p1[3] p2[3] p3[3] uv[6] n[9]; flags bit0 = has uv, bit1 = has normals */
    RL_OW_TRANSFORM = 35,  /* params: M[9] Minv[9] (3x3 row-major); child_begin = child node id       */
    RL_OW_TRANSLATE = 36,  /* params: offset[3]; child_begin = child node id                          */
    RL_OW_BVH = 37,        /* children[child_begin .. child_end) — Bvh::new(vec)                      */
    RL_OW_LIST = 38,       /* children[child_begin .. child_end) — slice / array of hittables         */
    RL_OW_CONSTANT_MEDIUM = 39,/* hittable/constant_medium.rs:8-23: params: density; material = the phase
                              * function; child_begin = boundary node id                              */
    RL_OW_MESH = 40        /* WavefrontObj::to_object(material) of the mesh rl_obj_parse left on the ctx: a Bvh of
                            * Triangle::from_model(points, texture_coords, normals, material)          */
};

enum { RL_CSG_UNION = 0, RL_CSG_INTERSECTION = 1, RL_CSG_DIFFERENCE = 2 };

/* A node's id is its index in rl_scene_desc.nodes[].  Ids may be assigned in ANY order: where the reference orders objects
 * (equal-t ties, "later object wins", RTC/src/scene/intersect.rs:159-168; the n1 / n2 and shadow walks) the library follows
 * the order a depth-first walk from roots[] meets the leaves — the reference's World / Group order — not the ids; rl_hit.node
 * reports the caller's id.  A subtree may be referenced from several parents (each reference is its own instance). */
typedef struct rl_node {
    int32_t kind;
    int32_t material;    /* leaf shapes: index into materials[]; -1 otherwise */
    int32_t child_begin;
    int32_t child_end;
    int32_t flags;
    int32_t param;       /* offset (in doubles) into params[]; -1 when the kind has none */
} rl_node;

enum {
    RL_MAT_RTC_PHONG = 1,        /* RTC/src/scene/material.rs:22-31 */
    RL_MAT_OW_LAMBERTIAN = 16,   /* OW/src/material.rs:69-97   (texture) */
    RL_MAT_OW_METAL = 17,        /* OW/src/material.rs:99-127  (color = albedo, fuzz) */
    RL_MAT_OW_DIELECTRIC = 18,   /* OW/src/material.rs:134-170 (refractive_index) */
    RL_MAT_OW_DIFFUSE_LIGHT = 19,/* OW/src/material.rs:178-195 (texture) */
    RL_MAT_OW_ISOTROPIC = 20     /* OW/src/material.rs:197-221 (texture) */
};

typedef struct rl_material {
    int32_t kind;
    int32_t texture;  /* RTC: pattern index or -1 for Surface::Color(color); OW: texture index */
    double color[3];
    double ambient, diffuse, specular, shininess, reflectivity, transparency, refractive_index;
    double fuzz;
} rl_material;

enum {
    RL_TEX_RTC_STRIPE = 1,   /* RTC/src/scene/pattern/stripe.rs:21-27    */
    RL_TEX_RTC_CHECKER3D = 2,/* RTC/src/scene/pattern/checker3d.rs:19-25 */
    RL_TEX_RTC_GRADIENT = 3, /* RTC/src/scene/pattern/gradient.rs:20-25  */
    RL_TEX_RTC_RING = 4,     /* RTC/src/scene/pattern/ring.rs:20-28      */
    RL_TEX_OW_SOLID = 16,    /* OW/src/texture.rs:15-23 (a = albedo)     */
    RL_TEX_OW_CHECKER = 17,  /* OW/src/texture.rs:25-55 (tex_a = even, tex_b = odd, scale) */
    RL_TEX_OW_IMAGE = 18,    /* OW/src/texture.rs:58-82 (image index)    */
    RL_TEX_OW_NOISE = 19     /* OW/src/texture.rs:84-94 (image = index into perlins[], scale) */
};

typedef struct rl_texture {
    int32_t kind;
    int32_t tex_a, tex_b; /* OW checker: even / odd texture indices */
    int32_t image;        /* OW image: index into images[]; OW noise: index into perlins[] */
    double a[3], b[3];
    double scale;         /* OW checker scale (the reference stores 1/scale); OW noise scale */
    double transform[16]; /* RTC pattern: forward 4x4 */
} rl_texture;

typedef struct rl_image {
    int32_t width, height;
    const float* rgb; /* width*height*3, row-major, top row first, LINEAR colour (image::Rgb32FImage) */
} rl_image;

typedef struct rl_perlin {  /* OW/src/perlin.rs:9-14 — the tables Perlin::new(rng) drew */
    double randvec[256][3];
    int32_t perm_x[256], perm_y[256], perm_z[256];
} rl_perlin;

typedef struct rl_light {   /* RTC/src/scene/light.rs:3-7 */
    double position[3];
    double intensity[3];
} rl_light;

typedef struct rl_scene_desc {
    int32_t abi_version; /* RL_B200_ABI_VERSION */
    int32_t flavor;      /* RL_FLAVOR_RTC | RL_FLAVOR_OW */
    const rl_node* nodes;          int32_t n_nodes;
    const int32_t* children;       int32_t n_children;
    const double* params;          int64_t n_params;
    const int32_t* roots;          int32_t n_roots;  /* RTC: World.objects in order; OW: exactly one root */
    const rl_material* materials;  int32_t n_materials;
    const rl_texture* textures;    int32_t n_textures;
    const rl_image* images;        int32_t n_images;
    const rl_light* lights;        int32_t n_lights; /* RTC World.lights */
    int32_t max_reflection_depth;  /* RTC World.max_reflection_depth (world.rs:26-31) */
    double void_color[3];          /* RTC World.void_color */
    const rl_perlin* perlins;      int32_t n_perlins; /* OW Noise textures (ABI version 2) */
} rl_scene_desc;

/* ---- cameras --------------------------------------------------------------------------------- */
typedef struct rl_rtc_camera {  /* RTC/src/scene/camera.rs:11-19 — Camera::new(hsize, vsize, fov, transform) */
    int32_t hsize, vsize;
    double fov;
    double transform[16];       /* forward view transform; the library inverts it */
} rl_rtc_camera;

typedef struct rl_ow_camera {   /* OW/src/camera.rs:24-39 — CameraParams */
    double aspect_ratio;
    int32_t image_width;
    int32_t samples_per_pixel;
    int32_t max_depth;
    int32_t _pad;
    double vfov;
    double lookfrom[3], lookat[3], vup[3];
    double defocus_angle;
    double focus_dist;
    double background[3];
    uint64_t seed;
} rl_ow_camera;

/* ---- ray batches (parity harness) -------------------------------------------------------------- */
typedef struct rl_ray {
    float origin[3];
    float direction[3];
    float time;    /* OW ray time; ignored for RTC */
    float _pad;
} rl_ray;

typedef struct rl_hit {
    int32_t node;  /* id (index into rl_scene_desc.nodes) of the leaf that was hit, -1 on miss */
    float t;
    float u, v;    /* triangle barycentrics / quad alpha,beta / 0 */
} rl_hit;

/* ---- statistics -------------------------------------------------------------------------------- */
typedef struct rl_stats {
    uint64_t rays;          /* every ray cast: camera + secondary + shadow             */
    uint64_t node_visits;   /* BVH2 node visits (two child slabs each)                  */
    uint64_t prim_tests;    /* analytic primitive tests (sphere/plane/cube/cyl/cone/quad) */
    uint64_t tri_tests;     /* ray-triangle tests                                       */
    uint64_t shades;        /* RTC: shaded hits; OW: scatter evaluations                */
    uint64_t samples;       /* W*H*spp rendered by this call                            */
    uint64_t overflow;      /* traversal-stack / work-list overflows (must be 0)        */
    float kernel_ms;        /* device time of the render kernels (CUDA events)          */
    float upload_ms;        /* device time of the last scene upload incl. LBVH build    */
    int32_t kernel_launches;/* kernels launched by this call                            */
    int32_t _pad;
} rl_stats;

/* ---- LBVH download (bit-exact host rebuild check) ---------------------------------------------- */
typedef struct rl_scene_info {
    int32_t flavor;
    int32_t n_prims;       /* analytic (non-BVH) primitives, always tested   */
    int32_t n_bvh_prims;   /* primitives under the LBVH (leaves)             */
    int32_t n_bvh_nodes;   /* internal nodes = n_bvh_prims - 1 (0 when < 2)  */
    int32_t n_materials, n_textures, n_lights;
    int32_t has_transparency;
    int64_t device_bytes;  /* HBM held by the scene                          */
} rl_scene_info;

typedef struct rl_lbvh_host {   /* caller allocates every array; sizes from rl_scene_info */
    float* prim_aabb;      /* [n_bvh_prims][6]  lo.xyz hi.xyz — the LBVH builder's input       */
    int32_t* prim_node;    /* [n_bvh_prims]     source node id of each BVH primitive           */
    uint64_t* morton;      /* [n_bvh_prims]     sorted 63-bit Morton keys                      */
    int32_t* sorted_prim;  /* [n_bvh_prims]     primitive index at each sorted position        */
    int32_t* left;         /* [n_bvh_nodes]     child ids: >=0 internal, ~leaf_pos when leaf   */
    int32_t* right;        /* [n_bvh_nodes]                                                     */
    int32_t* parent;       /* [n_bvh_nodes + n_bvh_prims] internal parents then leaf parents   */
    float* node_aabb;      /* [n_bvh_nodes][6]                                                  */
    float scene_lo[3], scene_hi[3]; /* centroid bounds used for Morton quantisation             */
} rl_lbvh_host;

/* ---- work partitioning (multi-GPU tile queue) --------------------------------------------------- */
typedef struct rl_job {   /* a rectangle of pixels x a range of samples */
    int32_t x0, y0, x1, y1;       /* pixel rectangle [x0,x1) x [y0,y1)            */
    int32_t chunk_begin, chunk_end; /* OW: sample-chunk range; RTC: ignored         */
} rl_job;

typedef struct rl_ctx rl_ctx;

/* ---- lifecycle --------------------------------------------------------------------------------- */
int rl_create(int device_id, rl_ctx** out);
/* ONE context over n GPUs of one node (n >= 1, distinct device ids), for a caller that is a single process — the way
 * a Rust `Camera::render` (OW/src/camera.rs:122-124, RTC/src/scene/camera.rs:93) would use a whole HGX box with one
 * call.  rl_scene_upload replicates the scene on every GPU; rl_render_ow / rl_render_ow_u8 run ONE persistent launch per
 * GPU whose warps pop (pixel x sample-chunk) items from one counter in GPU 0's HBM (system-scope atomics over NVLink
 * peer memory) and store finished items straight into GPU 0's partial-sum buffer, which GPU 0 folds; rl_render_rtc /
 * rl_render_rtc_u8 give each GPU a band of rows that it copies to the caller's buffer itself.  The image is bit-identical
 * to a 1-GPU render.  Every other entry point acts on GPU 0.  Needs peer access between GPU 0 and the others. */
int rl_create_multi(const int32_t* device_ids, int32_t n, rl_ctx** out);
int rl_device_count(const rl_ctx* ctx); /* GPUs behind the ctx: 1, or rl_create_multi's n */
void rl_destroy(rl_ctx* ctx);
const char* rl_last_error(const rl_ctx* ctx); /* ctx may be NULL: error of the last failed rl_create */
int rl_abi_version(void);
int rl_device_info(rl_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, int64_t* hbm_bytes);
/* block until everything queued on the ctx's own stream has finished (the `*_device` entry points that
 * are given stream = 0 run there) */
int rl_synchronize(rl_ctx* ctx);

/* Roofline denominators of this path, measured on the ctx's device: non-tensor FP32 FMA throughput
 * (TFLOP/s), L2-resident read bandwidth and HBM read bandwidth (GB/s).  Takes a few milliseconds. */
int rl_measure_peaks(rl_ctx* ctx, double* fp32_tflops, double* l2_gbs, double* hbm_gbs);

/* ---- scene -------------------------------------------------------------------------------------- */
/* Replaces walking `World.objects` (RTC/src/scene/world.rs:46-55) / `world.hit` (OW/src/camera.rs:247):
 * flattens the tree, uploads SoA buffers and builds the LBVH on the device. */
int rl_scene_upload(rl_ctx* ctx, const rl_scene_desc* scene);
int rl_scene_info_get(rl_ctx* ctx, rl_scene_info* out);
/* The HOST half of rl_scene_upload on its own (no device needed): validates and flattens the description exactly like
 * the upload does and reports what it would produce; `err` (optional, err_cap bytes) receives the message on failure.
 * Lets a caller reject a scene the device path cannot lower (RL_E_UNSUPPORTED) before committing to it. */
int rl_scene_check(const rl_scene_desc* scene, rl_scene_info* out, char* err, int32_t err_cap);
int rl_lbvh_download(rl_ctx* ctx, rl_lbvh_host* out);

/* ---- OBJ ingest on the device ------------------------------------------------ */
/* Replaces WavefrontObj::parse (RTC/src/io/wavefront_obj.rs:22-76, OW/src/io/wavefront_obj.rs:32-104) and to_object for
 * meshes large enough that a sequential host parser is the bottleneck: the OBJ text is copied to the GPU once and parsed
 * there (lines, records, exact decimal -> f64 conversion, fan triangulation, group replacement order), and a scene node
 * of kind RL_RTC_MESH / RL_OW_MESH then instances the parsed triangles — transform bake, f32 packing, LBVH input boxes —
 * on the device as part of rl_scene_upload, so the triangles never visit the host.  `flavor` selects the record set
 * (OW reads `vt`, RTC ignores it) and the face-index rules of the respective parser.
 * A ctx holds ONE parsed mesh at a time (a new rl_obj_parse replaces it).  Errors: RL_E_INVALID for a face index that
 * is out of bounds (the reference panics), RL_E_UNSUPPORTED for a number with more than 19 significant digits or a
 * decimal exponent beyond +-60 (the only ones the device parser does not convert exactly: it fails rather than round
 * differently from `str::parse::<f64>`). */
typedef struct rl_obj_info {
    int32_t n_vertices, n_normals, n_texcoords; /* `v` / `vn` / `vt` records accepted                 */
    int32_t n_triangles;                        /* after fan triangulation and group replacement      */
    int32_t n_groups;                           /* `g` records + the default group                    */
    int32_t ignored;                            /* WavefrontObj.ignored: lines that yielded nothing   */
    int32_t kernel_launches;
    int32_t _pad;
    double bounds[6];                           /* lo.xyz hi.xyz of the triangles' points (object space) */
} rl_obj_info;
int rl_obj_parse(rl_ctx* ctx, const char* text, uint64_t len, int32_t flavor, rl_obj_info* info);
/* the parsed triangles, for parity checks against the host parser: tri_p [n][3][3], tri_n [n][3][3], tri_uv [n][3][2]
 * (f64), flags [n] (bit0 = has normals, bit1 = has texture coordinates); any pointer may be NULL */
int rl_obj_download(rl_ctx* ctx, double* tri_p, double* tri_n, double* tri_uv, uint8_t* flags);

/* ---- ray batches -------------------------------------------------------------------------------- */
/* RTC: closest hit per `intersect::hit` (RTC/src/scene/intersect.rs:159-168) over `World::intersect`.
 * OW : `world.hit(r, [t_min, inf))` (OW/src/camera.rs:242-247).                                       */
int rl_trace_batch(rl_ctx* ctx, const rl_ray* rays, uint64_t n, rl_hit* out);
/* OW only.  The same, for rays that START on a surface (the scattered rays of camera.rs:248-255): self_nodes[i] is the
 * leaf node ray i starts on, or -1.  The reference relies on f64 and t_min = 1e-10 (camera.rs:242) not to re-hit that
 * surface at t ~ 0; the f32 device path instead never re-hits the planar primitive a ray starts on and takes only the FAR
 * root of its own sphere.  OW rays, with or without self nodes, run through the render kernel itself (a TRACE instantiation
 * of k_ow_render5: work items are ray indices, the same big list at ray start, node steps, leaf rounds and service
 * thresholds): this is the production code path, not a second one.
 * t_min is the renderer's: 1e-5 * max|origin_k| + 1e-6 in units of the NORMALISED direction. */
int rl_trace_batch_ex(rl_ctx* ctx, const rl_ray* rays, const int32_t* self_nodes, uint64_t n, rl_hit* out);

/* ---- renders ------------------------------------------------------------------------------------ */
/* RTC Camera::render (RTC/src/scene/camera.rs:93-124): out_rgb = W*H*3 f32, row-major (width*y + x),
 * mean over anti_aliasing_samples^2 rays per pixel. */
int rl_render_rtc(rl_ctx* ctx, const rl_rtc_camera* cam, uint32_t anti_aliasing_samples,
                  float* out_rgb, rl_stats* stats);

/* OW Camera::_render (OW/src/camera.rs:145-199): out_rgb_sum = W*H*3 f32 SUMS over
 * samples [first_sample, first_sample + samples_per_pixel) so Canvas::merge / checkpoints keep working.
 * image height = max(1, trunc(image_width / aspect_ratio)) (camera.rs:75). */
int rl_render_ow(rl_ctx* ctx, const rl_ow_camera* cam, uint32_t first_sample,
                 float* out_rgb_sum, rl_stats* stats);
/* The same renders with the reference's 8-bit output encoders run on the device, so the D2H copy is
 * 3 bytes per pixel instead of 12; out_rgb8 = W*H*3 bytes, row-major.
 *   RTC: Canvas::ppm's `translate` (RTC/src/draw/canvas.rs:53-56): round(c * 255) half away from zero, clamped to [0,255].
 *   OW : Canvas::pixel_data + Color::write_ppm (OW/src/camera.rs:293-295, color.rs:22-57, 130-136):
 *        c = sum * (1 / samples); linear_to_srgb; floor(c * 255.999) clamped to [0,255].
 * Both are evaluated in f64 from the f32 framebuffer, i.e. exactly what the reference's encoder makes of the values
 * rl_render_rtc / rl_render_ow return. */
int rl_render_rtc_u8(rl_ctx* ctx, const rl_rtc_camera* cam, uint32_t anti_aliasing_samples, uint8_t* out_rgb8,
                     rl_stats* stats);
int rl_render_ow_u8(rl_ctx* ctx, const rl_ow_camera* cam, uint32_t first_sample, uint8_t* out_rgb8, rl_stats* stats);
int rl_ow_image_height(const rl_ow_camera* cam);
/* number of sample chunks the OW renderer splits samples_per_pixel into (deterministic reduction) */
int rl_ow_num_chunks(const rl_ow_camera* cam);

/* Device-resident variants used for multi-GPU sharding: render only `jobs`, accumulate into device
 * buffers that the caller owns (e.g. torch tensors) so the framebuffer can be gathered with NCCL.
 *   RTC: d_out_rgb [H][W][3] f32, pixels outside the jobs are left untouched.
 *   OW : d_partial [n_chunks][H][W][4] f32 per-chunk partial sums, rgb + one pad float so that a finished item is
 *        ONE 16-byte store (untouched outside the jobs);
 *        rl_ow_reduce_device folds the chunks in order into d_out_rgb_sum [H][W][3].
 * `stream` is a cudaStream_t; 0 = the ctx's own non-blocking stream (pass cudaStreamLegacy, 0x1, for the
 * legacy default stream).  With stats == NULL the call only LAUNCHES (asynchronous); errors and overflows
 * then surface at rl_synchronize. */
int rl_render_rtc_device(rl_ctx* ctx, const rl_rtc_camera* cam, uint32_t anti_aliasing_samples,
                         const rl_job* jobs, int32_t n_jobs, void* d_out_rgb, void* stream,
                         rl_stats* stats);
int rl_render_ow_device(rl_ctx* ctx, const rl_ow_camera* cam, uint32_t first_sample,
                        const rl_job* jobs, int32_t n_jobs, void* d_partial, void* stream,
                        rl_stats* stats);
int rl_ow_reduce_device(rl_ctx* ctx, const rl_ow_camera* cam, const void* d_partial,
                        void* d_out_rgb_sum, void* stream);

/* Cross-GPU dynamic tile queue for ONE PROCESS PER GPU (torchrun-style launches — a single process uses
 * rl_create_multi instead).  The exporting ctx owns a control block in its HBM with RL_QUEUE_SLOTS independent
 * {work counter, completion counter} pairs; peers map it with CUDA IPC and their persistent render kernels pop
 * (pixel x sample-chunk) items from the work counter with system-scope atomics over NVLink, a whole batch per atomic,
 * the next batch prefetched while the current one renders — no host in the loop.
 *   owner : rl_queue_export -> 64-byte handle (ship it to the peers); rl_queue_reset(slot) before a render on that slot
 *   peers : rl_queue_import
 *   all   : rl_render_ow_shared(slot) — asynchronous; every rank passes the SAME job list.
 * Two slots let consecutive renders alternate: the owner resets slot (i + 1) % 2 while render i drains slot i % 2, so a
 * rank launches render i + 1 without waiting for anybody — a late rank simply pops fewer items — and the only rendezvous
 * left is the one at the end of a render, after which the owner folds.
 * Fused gather: the owner also exports n_slots partial-sum buffers of bytes_per_slot each; with d_partial == NULL
 * rl_render_ow_shared stores every finished item straight into the OWNER's buffer of that slot over NVLink (each
 * [chunk][y][x] entry is written exactly once, by whichever GPU popped it), and the launch is refused when the camera
 * needs more than bytes_per_slot.  After the end-of-render rendezvous the owner checks rl_queue_completed == the number
 * of items (a rank that died leaves it short) and folds with rl_ow_reduce_shared. */
#define RL_QUEUE_SLOTS 2
int rl_queue_export(rl_ctx* ctx, void* handle64);
int rl_queue_import(rl_ctx* ctx, const void* handle64);
int rl_queue_reset(rl_ctx* ctx, void* stream, int32_t slot);
int rl_queue_completed(rl_ctx* ctx, void* stream, int32_t slot, uint64_t* items);
int rl_partial_export(rl_ctx* ctx, uint64_t bytes_per_slot, int32_t n_slots, void* handle64);
int rl_partial_import(rl_ctx* ctx, const void* handle64, uint64_t bytes_per_slot, int32_t n_slots);
int rl_render_ow_shared(rl_ctx* ctx, const rl_ow_camera* cam, uint32_t first_sample, const rl_job* jobs,
                        int32_t n_jobs, void* d_partial, void* stream, int32_t slot);
int rl_ow_reduce_shared(rl_ctx* ctx, const rl_ow_camera* cam, int32_t slot, void* d_out_rgb_sum, void* stream);
/* work items (pixel x sample-chunk, in padded 8x4 micro-tiles) a job list holds for this camera: what
 * rl_queue_completed must reach */
int64_t rl_ow_job_items(const rl_ow_camera* cam, const rl_job* jobs, int32_t n_jobs);

/* Scheduling parameters of the OW render kernel ("ow.variant", "ow.slots", "ow.minb", "ow.ctas_per_sm", "ow.svc_lo",
 * "ow.exit_min", "ow.leaf_min", "ow.svc_min"; 0 = the measured default).  They change WHEN work is done, never what is
 * computed: the image is bit-identical for every setting.  Unknown names / out-of-range values: RL_E_INVALID. */
int rl_set_option(rl_ctx* ctx, const char* name, int32_t value);

/* enable/disable the instrumented (counting) kernel variants; counters cost atomics, so timing runs
 * keep them off and a separate instrumented pass with the same seed fills rl_stats. */
int rl_set_instrumented(rl_ctx* ctx, int enabled);

#ifdef __cplusplus
}
#endif
#endif /* RL_B200_H */
