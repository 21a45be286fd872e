// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// Serial host rebuild of the device LBVH (rendering_learning_b200/csrc/lbvh.cu): same f32 arithmetic
// (compiled with -ffp-contract=off), a stable sort by 63-bit Morton key, Karras' 2012 hierarchy with
// position tie-breaks, recursive box union.  tests/test_gpu_lbvh.py demands bit equality of every array.
// The reference has no LBVH (its BVH is the median split of OW/src/bvh.rs:22-61); this file checks the
// north_star's "bit-exact against a host rebuild" requirement, not a reference algorithm.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

namespace {

uint64_t expand21(uint32_t v) {
    uint64_t x = v & 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
inline float centroid(float lo, float hi) { return (lo + hi) * 0.5f; }
inline int clz64(uint64_t x) { return x ? __builtin_clzll(x) : 64; }
inline int clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }

struct Builder {
    const uint64_t* keys;
    int n;
    int delta(int i, int j) const {
        if (j < 0 || j >= n) return -1;
        uint64_t a = keys[i], b = keys[j];
        if (a == b) return 64 + clz32((uint32_t)(i ^ j));
        return clz64(a ^ b);
    }
};

void box_of(int child, const float* aabb, const int* sorted_prim, const int* left, const int* right, float* node_aabb,
            float out[6]) {
    if (child < 0) {
        const float* b = aabb + 6 * sorted_prim[~child];
        for (int k = 0; k < 6; k++) out[k] = b[k];
        return;
    }
    float l[6], r[6];
    box_of(left[child], aabb, sorted_prim, left, right, node_aabb, l);
    box_of(right[child], aabb, sorted_prim, left, right, node_aabb, r);
    for (int k = 0; k < 3; k++) {
        out[k] = std::fmin(std::fmin(FLT_MAX, l[k]), r[k]);
        out[3 + k] = std::fmax(std::fmax(-FLT_MAX, l[3 + k]), r[3 + k]);
    }
    for (int k = 0; k < 6; k++) node_aabb[6 * child + k] = out[k];
}

}  // namespace

extern "C" int orc_lbvh_build(const float* aabb, int n, uint64_t* morton, int32_t* sorted_prim, int32_t* left,
                              int32_t* right, int32_t* parent, float* node_aabb, float* bounds) {
    if (n <= 0) return 0;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) {
            float c = centroid(aabb[6 * i + k], aabb[6 * i + 3 + k]);
            lo[k] = std::fmin(lo[k], c);
            hi[k] = std::fmax(hi[k], c);
        }
    for (int k = 0; k < 3; k++) {
        bounds[k] = lo[k];
        bounds[3 + k] = hi[k];
    }
    std::vector<uint64_t> key(n);
    for (int i = 0; i < n; i++) {
        uint32_t q[3];
        // one scale for all axes: cubic Morton cells (see k_morton)
        const float ext = std::fmax(std::fmax(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
        for (int k = 0; k < 3; k++) {
            float c = centroid(aabb[6 * i + k], aabb[6 * i + 3 + k]);
            float t = ext > 0.0f ? (c - lo[k]) / ext : 0.0f;
            float s = std::fmin(std::fmax(t * 2097152.0f, 0.0f), 2097151.0f);
            q[k] = (uint32_t)s;
        }
        key[i] = (expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]);
    }
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b]; });
    for (int i = 0; i < n; i++) {
        sorted_prim[i] = order[i];
        morton[i] = key[order[i]];
    }
    if (n < 2) return 0;
    Builder B{morton, n};
    for (int i = 0; i < n - 1; i++) {
        int d = (B.delta(i, i + 1) - B.delta(i, i - 1)) >= 0 ? 1 : -1;
        int dmin = B.delta(i, i - d);
        int lmax = 2;
        while (B.delta(i, i + lmax * d) > dmin) lmax <<= 1;
        int l = 0;
        for (int t = lmax >> 1; t >= 1; t >>= 1)
            if (B.delta(i, i + (l + t) * d) > dmin) l += t;
        int j = i + l * d;
        int dnode = B.delta(i, j);
        int s = 0, t = l;
        do {
            t = (t + 1) >> 1;
            if (B.delta(i, i + (s + t) * d) > dnode) s += t;
        } while (t > 1);
        int gamma = i + s * d + std::min(d, 0);
        int a = std::min(i, j), b = std::max(i, j);
        int lc = (a == gamma) ? ~gamma : gamma;
        int rc = (b == gamma + 1) ? ~(gamma + 1) : gamma + 1;
        left[i] = lc;
        right[i] = rc;
        if (lc >= 0) parent[lc] = i; else parent[(n - 1) + gamma] = i;
        if (rc >= 0) parent[rc] = i; else parent[(n - 1) + gamma + 1] = i;
    }
    parent[0] = -1;
    float root[6];
    box_of(0, aabb, sorted_prim, left, right, node_aabb, root);
    return 0;
}
