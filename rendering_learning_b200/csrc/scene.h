// Flat (device) scene layout — what the flattener lowers the reference's object tree into.
// Everything the kernels touch is 16-byte vectors so node / triangle / primitive fetches are single
// LDG.128 / LDS.128 instructions.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/rl_b200.h"

namespace rl {

// primitive kinds on the device
enum : int {
    PK_RTC_SPHERE = 1, PK_RTC_PLANE = 2, PK_RTC_CUBE = 3, PK_RTC_CYLINDER = 4, PK_RTC_CONE = 5,
    PK_TRIANGLE = 8, PK_OW_SPHERE = 9, PK_OW_QUAD = 10
};

// BVH leaf reference: (type << 28) | index into the per-type array
constexpr int REF_TRI = 0, REF_SPHERE = 1, REF_QUAD = 2, REF_MEDIUM = 3;
__host__ __device__ inline int make_ref(int type, int index) { return (type << 28) | index; }
__host__ __device__ inline int ref_type(int ref) { return (ref >> 28) & 7; }
__host__ __device__ inline int ref_index(int ref) { return ref & 0x0FFFFFFF; }

// OW spheres at least this large evaluate their quadratic in f64: seen from near its surface (the r = 1000 ground of the
// cover scene) r^2 - |perp|^2 cancels ~7 digits, more than f32 has.
constexpr float OW_BIG_RADIUS = 64.0f;
// Primitives that are large against the rest of the scene stay OUT of the LBVH (flatten.cpp select_big_prims): their box
// would cover everything, so every ray tests them anyway.  They go on a short "big" list that is tested once per ray
// before the traversal, which then starts with a finite tmax.
constexpr int OW_MAX_BIG = 8;  // more than this and the rest go through the LBVH like any other sphere

// RTC analytic primitive (unit shape + composed, pre-inverted transform). 176 B.
struct RtcPrim {
    float4 inv[3];   // world -> object affine rows (inv_total = inv_leaf * ... * inv_root)
    float4 fwd[3];   // object -> world affine rows (maps the surface-snapped local hit point back)
    float4 pat[3];   // object -> pattern space rows (pattern.inv); identity if no pattern
    float ymin, ymax;  // cylinder / cone truncation (+-inf when None)
    int kind, flags;   // flags bit0 = closed
    int material, node;
    int csg_first, csg_count;  // DevScene.csg[csg_first .. +csg_count): the CSG nodes above this leaf, post-order (0 = none)
};

// triangle: traversal data (48 B) and shading data (64 B) in separate arrays
struct TriVerts {
    float4 p0;  // xyz, w = material (int bits)
    float4 p1;  // xyz, w = source node id (int bits)
    float4 p2;  // xyz, w = flags (bit0 smooth normals, bit1 has uv) | (xform index + 1) << 8
};
// OW triangle in the reference's own form (flat/plane.rs:23-80 + flat/triangle.rs:60-67): plane (unit normal, d) and the two
// vectors that turn a point of the plane into (alpha, beta) with one dot product each — the layout of OwQuad's first 48 bytes.
// Derived ON THE DEVICE from the f32 TriVerts (ow_kernels.cu k_tri_planes, f64 arithmetic), so host-flattened and
// device-ingested triangles share one code path and the same bits.
struct OwTriPlane {
    float4 n;  // unit normal, w = d = n . p0
    float4 a;  // alpha = a.xyz . p - a.w   (weight of p1)
    float4 b;  // beta  = b.xyz . p - b.w   (weight of p2)
};
struct TriShade {
    float4 s0;  // n0.xyz, uv0.x      (flat: n0 = the face normal, world space)
    float4 s1;  // n1.xyz, uv0.y
    float4 s2;  // n2.xyz, uv1.x
    float4 s3;  // uv1.y, uv2.x, uv2.y, 0
};
struct Xform {  // world -> pattern space for triangles whose material carries a pattern
    float4 r[3];
};

struct OwSphere {  // 32 B
    float4 c;   // centre at time 0, w = radius
    float4 dc;  // centre(1) - centre(0), w = material (int bits)
};
struct OwQuad {  // 64 B; the intersection test reads the first 48
    float4 n;  // unit normal, w = d = n . q
    float4 a;  // alpha = a.xyz . p - a.w, with a.xyz = v x w and a.w = a.xyz . q   (w = n_raw / (n_raw . n_raw))
    float4 b;  // beta  = b.xyz . p - b.w, with b.xyz = w x u and b.w = b.xyz . q
    int4 m;    // x = material
};
struct OwMedium {  // 16 B — hittable/constant_medium.rs:8-12
    int ref_begin, ref_count;  // boundary primitives: medium_refs[ref_begin .. +ref_count) (never in the LBVH themselves)
    float neg_inv_density;
    int material;              // the phase function (Isotropic)
};

struct DevMaterial {  // 48 B
    float4 color;  // rgb (RTC surface colour / OW metal albedo), w = texture index (int bits, -1 none)
    float4 a;      // RTC: ambient, diffuse, specular, shininess | OW: fuzz, refractive_index, 0, 0
    float4 b;      // RTC: reflectivity, transparency, refractive_index, 0 | w = kind (int bits)
};
struct DevTexture {  // 48 B
    float4 a;  // rgb, w = kind (int bits)
    float4 b;  // rgb, w = OW checker 1/scale
    int4 idx;  // OW: tex_a, tex_b, image, 0
};
struct DevImage {
    const float4* texels;  // rgba f32, row-major, top row first
    int width, height;
};
struct DevLight {
    float4 pos, intensity;
};

// BVH2 traversal node: both children's boxes live in the parent, as CENTRE + HALF EXTENT, the two children's values of
// each coordinate ADJACENT (an aligned register pair after the LDG.128). 64 B = 4 x LDG.128.
// The slab test is then centre * (1/d) - o/d -+ half * |1/d| per axis: FMAs only, no per-axis min / max, and each FMA is
// ONE packed FFMA2 (fma.rn.f32x2, new on sm_100) that serves both children (device.cuh bvh2_step).  lbvh.cu k_pack
// rounds the half extent UP so that [centre - half, centre + half] contains the f32 box the hierarchy was refitted with;
// an empty child slot has half = -1 (entry > exit for every ray).
struct BvhNode {
    float4 a;  // c0.centre.x c1.centre.x c0.centre.y c1.centre.y
    float4 b;  // c0.centre.z c1.centre.z c0.half.x   c1.half.x
    float4 c;  // c0.half.y   c1.half.y   c0.half.z   c1.half.z
    int4 d;    // child0, child1 (>=0 internal node, <0 = ~leaf ref), 0, 0
};

// everything a kernel needs, passed by value
struct DevScene {
    int flavor;
    int n_prims;      // RTC analytic prims (brute force)
    int n_tris, n_spheres, n_quads;
    int n_bvh_prims, n_bvh_nodes;  // traversal nodes (>= 1 when n_bvh_prims >= 1)
    int n_big;                     // OW: leaf refs tested brute force before the traversal
    int n_csg;                     // RTC: CSG nodes
    int n_media, n_perlins;        // OW: constant media, Perlin tables
    int n_materials, n_textures, n_lights, n_images, n_xforms;
    int has_transparency;
    int max_reflection_depth;
    float void_color[3];
    const RtcPrim* prims;
    const TriVerts* tri_verts;
    const TriShade* tri_shade;
    const OwTriPlane* tri_plane;  // OW only, [n_tris]
    const Xform* xforms;
    const OwSphere* spheres;
    const OwQuad* quads;
    const int* sphere_node;  // source node ids
    const int* quad_node;
    const DevMaterial* materials;
    const DevTexture* textures;
    const DevImage* images;
    const DevLight* lights;
    const BvhNode* nodes;
    const int* big_refs;
    const OwMedium* media;
    const int* medium_refs;
    const float4* perlin_vec;  // [n_perlins][256] gradient vectors
    const int* perlin_perm;    // [n_perlins][3][256] perm_x, perm_y, perm_z
    const int4* csg;  // (operation, lo, mid, hi): left child = prims [lo, mid), right child = prims [mid, hi)
    const int* ord_node;  // RTC: DFS leaf ordinal (RtcPrim.node, TriVerts.p1.w) -> the caller's node id
    int n_ord;
};

// host-side result of flattening
struct FlatScene {
    int flavor = 0;
    std::vector<RtcPrim> prims;
    std::vector<TriVerts> tri_verts;
    std::vector<TriShade> tri_shade;
    std::vector<Xform> xforms;
    std::vector<OwSphere> spheres;
    std::vector<OwQuad> quads;
    std::vector<int> sphere_node, quad_node;
    std::vector<DevMaterial> materials;
    std::vector<DevTexture> textures;
    std::vector<DevLight> lights;
    struct Img { int w, h; std::vector<float4> texels; };
    std::vector<Img> images;
    // LBVH input: one AABB + leaf ref + source node per bounded primitive
    std::vector<float> bvh_aabb;   // [n][6]
    std::vector<int> bvh_ref;      // [n]
    std::vector<int> bvh_node_id;  // [n]
    std::vector<int> ord_node;     // RTC: DFS leaf ordinal -> the caller's node id (flatten.cpp leaf_ordinal)
    std::vector<int> big_refs;     // OW leaf refs kept out of the LBVH (OW_BIG_RADIUS)
    std::vector<int4> csg;         // RTC CSG nodes, post-order
    std::vector<OwMedium> media;
    std::vector<int> medium_refs;
    std::vector<float4> perlin_vec;
    std::vector<int> perlin_perm;
    int has_transparency = 0;
    int max_reflection_depth = 5;
    float void_color[3] = {0, 0, 0};
    // uses of the ctx's device-resident mesh (RL_RTC_MESH / RL_OW_MESH): the flattener only RESERVES their triangle slots
    // (zero-filled here) and records the composed transform; obj_ingest.cu's k_mesh_instance fills them on the device
    struct MeshUse {
        double fwd[3][4], inv[3][4];
        int material, node, tri_first, bvh_first, xf;
    };
    std::vector<MeshUse> meshes;
};

// what the flattener may know about the ctx's parsed mesh: its size and object-space extent, never its triangles
struct MeshMeta {
    int n_triangles;
    double bounds[6];
};

// returns RL_OK or an RL_E_* code and fills `err`; `mesh` = the ctx's parsed mesh (nullptr: RL_*_MESH nodes are an error)
int flatten_scene(const rl_scene_desc* d, FlatScene* out, std::string* err, const MeshMeta* mesh = nullptr);

}  // namespace rl
