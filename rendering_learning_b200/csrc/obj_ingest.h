// OBJ ingest on the device (obj_ingest.cu): parsed per-triangle arrays in HBM and their instancing into a flat scene.
#pragma once
#include <string>

#include <cuda_runtime.h>

#include "scene.h"

namespace rl {

// what rl_obj_parse leaves on the device: the triangles in the order WavefrontObj::to_object() would hand them on
struct ObjMesh {
    int n_triangles = 0;
    double* tri_p = nullptr;            // [n][3][3] points
    double* tri_n = nullptr;            // [n][3][3] normals (zeros when the triangle has none)
    double* tri_uv = nullptr;           // [n][3][2] texture coordinates (zeros when none)
    unsigned char* tri_flags = nullptr; // [n] bit0 = has normals (smooth), bit1 = has texture coordinates
};

// one use of the ctx's mesh in a scene: composed transform of the node (f64), where its triangles go in the flat arrays
struct MeshInstance {
    double fwd[3][4];  // object -> world
    double inv[3][4];  // world -> object (normals: inv^T)
    int flavor, material, node;  // node: what the triangles carry (RTC: the DFS leaf ordinal, OW: the caller's node id)
    int report_node;             // the caller's node id (rl_scene_download's prim_node)
    int tri_first;     // first slot in tri_verts / tri_shade
    int bvh_first;     // first slot in the LBVH input arrays
    int xf;            // RTC: index + 1 of the world -> pattern transform of the material's pattern (0 = none)
};

int obj_parse_device(const char* text, uint64_t len, int flavor, cudaStream_t s, ObjMesh* out, rl_obj_info* info, std::string* err);
void obj_mesh_free(ObjMesh* m);
cudaError_t launch_mesh_instance(const ObjMesh& m, const MeshInstance& inst, TriVerts* tv, TriShade* ts, float* aabb, int* refs,
                                 int* node_ids, cudaStream_t s);

}  // namespace rl
