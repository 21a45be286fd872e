#pragma once
#include <cuda_runtime.h>

#include "scene.h"

namespace rl {

// device buffers of one LBVH (all sized for n primitives / n-1 internal nodes)
struct LbvhBuffers {
    float* prim_aabb = nullptr;     // [n][6]   input
    int* prim_ref = nullptr;        // [n]      input: leaf refs (type<<28 | index)
    float* bounds = nullptr;        // [6]      centroid bounds
    uint64_t* keys = nullptr;       // [n]      sorted Morton keys
    int* sorted_prim = nullptr;     // [n]      primitive at each sorted position
    uint64_t* keys_tmp = nullptr;   // [n]
    int* idx_tmp = nullptr;         // [n]
    int* left = nullptr;            // [n-1]
    int* right = nullptr;           // [n-1]
    int* parent = nullptr;          // [2n-1]   internal parents, then leaf parents
    float* node_aabb = nullptr;     // [n-1][6]
    int* counters = nullptr;        // [n-1]
    BvhNode* nodes = nullptr;       // [max(n-1,1)] packed traversal nodes
};

cudaError_t lbvh_build(const LbvhBuffers& b, int n, cudaStream_t stream, int* launches);

}  // namespace rl
