// GPU LBVH build: Morton codes -> radix sort -> Karras hierarchy -> bottom-up refit -> packed BVH2 nodes.
//
// Replaces the reference's recursive median-split `Bvh::new` (OW/src/bvh.rs:22-61) and RTC's
// `Bounded<Group<Triangle>>` brute-force mesh (RTC/src/io/wavefront_obj.rs:72-75).  Closest-hit results
// do not depend on the hierarchy, so any correct BVH gives the reference's hits.
//
// Every step is bit-reproducible so the host rebuild in oracle/oracle_lbvh.cpp matches exactly:
//   * centroid / quantisation arithmetic uses explicit round-to-nearest intrinsics (no FMA contraction),
//   * the sort is a STABLE least-significant-digit radix sort, so equal keys keep ascending primitive order,
//   * Karras' delta() breaks key ties with the sorted position,
//   * node boxes are pure min / max unions (order independent).
//
// Scenes on this path have <= ~6 k primitives, i.e. the build is latency bound, not
// bandwidth bound: the whole sort runs in ONE CTA (8 passes x 8 bits, no inter-kernel round trips).
#include <cfloat>
#include <cstdint>

#include "lbvh.h"

namespace rl {

namespace {

constexpr int SORT_THREADS = 1024;
constexpr int SORT_WARPS = SORT_THREADS / 32;

__device__ __forceinline__ uint64_t expand21(uint32_t v) {
    // spread the low 21 bits of v so that there are two zero bits between each
    uint64_t x = v & 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__device__ __forceinline__ float centroid(float lo, float hi) { return __fmul_rn(__fadd_rn(lo, hi), 0.5f); }

// 1. centroid bounds (single CTA; n is small)
__global__ void k_centroid_bounds(const float* __restrict__ aabb, int n, float* __restrict__ bounds /*6*/) {
    __shared__ float s_lo[3][32], s_hi[3][32];
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        for (int k = 0; k < 3; k++) {
            float c = centroid(aabb[6 * i + k], aabb[6 * i + 3 + k]);
            lo[k] = fminf(lo[k], c);
            hi[k] = fmaxf(hi[k], c);
        }
    }
    for (int k = 0; k < 3; k++) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
        if ((threadIdx.x & 31) == 0) {
            s_lo[k][threadIdx.x >> 5] = lo[k];
            s_hi[k][threadIdx.x >> 5] = hi[k];
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        int nw = blockDim.x >> 5;
        for (int k = 0; k < 3; k++) {
            float l = threadIdx.x < nw ? s_lo[k][threadIdx.x] : FLT_MAX;
            float h = threadIdx.x < nw ? s_hi[k][threadIdx.x] : -FLT_MAX;
            for (int o = 16; o > 0; o >>= 1) {
                l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
                h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o));
            }
            if (threadIdx.x == 0) {
                bounds[k] = l;
                bounds[3 + k] = h;
            }
        }
    }
}

// 2. 63-bit Morton keys (21 bits per axis) of the box centroids
__global__ void k_morton(const float* __restrict__ aabb, int n, const float* __restrict__ bounds,
                         uint64_t* __restrict__ keys, int* __restrict__ idx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t q[3];
    // ONE scale for all three axes (the largest centroid extent): the Morton grid cells are cubes.  With a scale per
    // axis a flat scene (the cover scene: 22 x 0.8 x 22) spends every third bit on a direction in which nothing is
    // separated; the cubic grid lowers the LBVH's SAH cost from 13.97 to 10.46 there (full-sweep SAH: 9.44) and from
    // 18.39 to 17.93 on the Cornell box + spot.
    const float ext = fmaxf(fmaxf(__fsub_rn(bounds[3], bounds[0]), __fsub_rn(bounds[4], bounds[1])),
                            __fsub_rn(bounds[5], bounds[2]));
    for (int k = 0; k < 3; k++) {
        float c = centroid(aabb[6 * i + k], aabb[6 * i + 3 + k]);
        float t = ext > 0.0f ? __fdiv_rn(__fsub_rn(c, bounds[k]), ext) : 0.0f;
        float s = fminf(fmaxf(__fmul_rn(t, 2097152.0f), 0.0f), 2097151.0f);
        q[k] = __float2uint_rz(s);
    }
    keys[i] = (expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]);
    idx[i] = i;
}

// 3. stable LSD radix sort, 8 passes x 8 bits, one CTA.  (keys_a, idx_a) -> ... -> result back in (keys_a, idx_a)
__global__ void __launch_bounds__(SORT_THREADS)
k_radix_sort(uint64_t* keys_a, int* idx_a, uint64_t* keys_b, int* idx_b, int n) {
    __shared__ int s_base[256];                 // running global offset of each digit
    __shared__ int s_warp[SORT_WARPS][256];     // per-warp digit counts of the current chunk
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t* kin = keys_a;
    int* iin = idx_a;
    uint64_t* kout = keys_b;
    int* iout = idx_b;
    for (int pass = 0; pass < 8; pass++) {
        const int shift = pass * 8;
        // histogram
        for (int d = tid; d < 256; d += SORT_THREADS) s_base[d] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += SORT_THREADS) atomicAdd(&s_base[(int)((kin[i] >> shift) & 255)], 1);
        __syncthreads();
        // exclusive scan of 256 bins by warp 0 (8 bins per lane)
        if (warp == 0) {
            int v[8], sum = 0;
            for (int k = 0; k < 8; k++) { v[k] = s_base[lane * 8 + k]; sum += v[k]; }
            int incl = sum;
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            int run = incl - sum;
            for (int k = 0; k < 8; k++) { s_base[lane * 8 + k] = run; run += v[k]; }
        }
        __syncthreads();
        // stable scatter, chunk by chunk in input order
        for (int c0 = 0; c0 < n; c0 += SORT_THREADS) {
            for (int d = tid; d < SORT_WARPS * 256; d += SORT_THREADS) (&s_warp[0][0])[d] = 0;
            __syncthreads();
            int i = c0 + tid;
            bool valid = i < n;
            uint64_t key = valid ? kin[i] : 0;
            int id = valid ? iin[i] : 0;
            int digit = (int)((key >> shift) & 255);
            unsigned act = __ballot_sync(0xffffffffu, valid);
            int rank = 0;
            if (valid) {
                unsigned same = __match_any_sync(act, digit);
                rank = __popc(same & ((1u << lane) - 1u));
                if (rank == 0) s_warp[warp][digit] = __popc(same);
            }
            __syncthreads();
            if (valid) {
                int off = s_base[digit] + rank;
                for (int w = 0; w < warp; w++) off += s_warp[w][digit];
                kout[off] = key;
                iout[off] = id;
            }
            __syncthreads();
            for (int d = tid; d < 256; d += SORT_THREADS) {
                int s = 0;
                for (int w = 0; w < SORT_WARPS; w++) s += s_warp[w][d];
                s_base[d] += s;
            }
            __syncthreads();
        }
        uint64_t* tk = kin; kin = kout; kout = tk;
        int* ti = iin; iin = iout; iout = ti;
        __threadfence_block();
        __syncthreads();
    }
    // 8 passes (even) -> the sorted data is back in (keys_a, idx_a)
}

// 3b. the same stable LSD radix sort across many CTAs, for primitive counts where ONE CTA is the bottleneck (1 M
// primitives: 28 of the 32 ms of the build).  Per 8-bit pass: per-tile digit histograms -> one exclusive scan over the
// [digit][tile] table -> each CTA scatters its own tile with exactly the chunk loop of k_radix_sort, starting from its
// scanned offsets.  Stable, so the result (and everything built on it) is identical to the one-CTA sort and to the host
// rebuild.
constexpr int SORT_TILE = 8 * SORT_THREADS;   // keys per CTA
constexpr int SORT_MULTI_MIN = 16384;         // below this the single launch wins (24 launches of ~4 us each here)

__global__ void __launch_bounds__(SORT_THREADS)
k_sort_hist(const uint64_t* __restrict__ kin, int n, int shift, int nb, int* __restrict__ hist /*[256][nb]*/) {
    __shared__ int s_h[256];
    for (int d = threadIdx.x; d < 256; d += SORT_THREADS) s_h[d] = 0;
    __syncthreads();
    const int begin = blockIdx.x * SORT_TILE, end = min(n, begin + SORT_TILE);
    for (int i = begin + threadIdx.x; i < end; i += SORT_THREADS) atomicAdd(&s_h[(int)((kin[i] >> shift) & 255)], 1);
    __syncthreads();
    for (int d = threadIdx.x; d < 256; d += SORT_THREADS) hist[d * nb + blockIdx.x] = s_h[d];
}

// exclusive scan of m ints in place, one CTA, 1024 per round with a running carry
__global__ void __launch_bounds__(SORT_THREADS) k_sort_scan(int* __restrict__ a, int m) {
    __shared__ int s_w[SORT_WARPS];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int c0 = 0; c0 < m; c0 += SORT_THREADS) {
        int i = c0 + tid;
        int v = i < m ? a[i] : 0;
        int incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = s_w[lane], wi = w;
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            s_w[lane] = wi - w;  // exclusive prefix of the warp totals
        }
        __syncthreads();
        const int carry = s_carry;
        if (i < m) a[i] = carry + s_w[warp] + incl - v;
        __syncthreads();
        if (tid == SORT_THREADS - 1) s_carry = carry + s_w[warp] + incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SORT_THREADS)
k_sort_scatter(const uint64_t* __restrict__ kin, const int* __restrict__ iin, uint64_t* __restrict__ kout,
               int* __restrict__ iout, int n, int shift, int nb, const int* __restrict__ offs /*[256][nb], scanned*/) {
    __shared__ int s_base[256];
    __shared__ int s_warp[SORT_WARPS][256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int d = tid; d < 256; d += SORT_THREADS) s_base[d] = offs[d * nb + blockIdx.x];
    __syncthreads();
    const int begin = blockIdx.x * SORT_TILE, end = min(n, begin + SORT_TILE);
    for (int c0 = begin; c0 < end; c0 += SORT_THREADS) {
        for (int d = tid; d < SORT_WARPS * 256; d += SORT_THREADS) (&s_warp[0][0])[d] = 0;
        __syncthreads();
        int i = c0 + tid;
        bool valid = i < end;
        uint64_t key = valid ? kin[i] : 0;
        int id = valid ? iin[i] : 0;
        int digit = (int)((key >> shift) & 255);
        unsigned act = __ballot_sync(0xffffffffu, valid);
        int rank = 0;
        if (valid) {
            unsigned same = __match_any_sync(act, digit);
            rank = __popc(same & ((1u << lane) - 1u));
            if (rank == 0) s_warp[warp][digit] = __popc(same);
        }
        __syncthreads();
        if (valid) {
            int off = s_base[digit] + rank;
            for (int w = 0; w < warp; w++) off += s_warp[w][digit];
            kout[off] = key;
            iout[off] = id;
        }
        __syncthreads();
        for (int d = tid; d < 256; d += SORT_THREADS) {
            int t = 0;
            for (int w = 0; w < SORT_WARPS; w++) t += s_warp[w][d];
            s_base[d] += t;
        }
        __syncthreads();
    }
}

// Karras 2012: length of the common prefix of sorted keys i and j (ties broken by position)
__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

// 4. hierarchy: one thread per internal node
__global__ void k_hierarchy(const uint64_t* __restrict__ keys, int n, int* __restrict__ left,
                            int* __restrict__ right, int* __restrict__ parent) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int lc = (lo == gamma) ? ~gamma : gamma;            // ~x encodes leaf (sorted position x)
    int rc = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    left[i] = lc;
    right[i] = rc;
    if (lc >= 0) parent[lc] = i; else parent[(n - 1) + gamma] = i;
    if (rc >= 0) parent[rc] = i; else parent[(n - 1) + gamma + 1] = i;
    if (i == 0) parent[0] = -1;
}

// 5. bottom-up refit: the second thread to reach a node merges its children
__global__ void k_refit(const float* __restrict__ aabb, const int* __restrict__ sorted_prim, int n,
                        const int* __restrict__ left, const int* __restrict__ right,
                        const int* __restrict__ parent, float* __restrict__ node_aabb,
                        int* __restrict__ counters) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int node = parent[(n - 1) + i];
    while (node >= 0) {
        int prev = atomicAdd(&counters[node], 1);
        if (prev == 0) return;  // first arrival: the sibling subtree is not finished yet
        __threadfence();
        float box[6] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
        int ch[2] = {left[node], right[node]};
        for (int c = 0; c < 2; c++) {
            const volatile float* src = ch[c] < 0 ? (const volatile float*)(aabb + 6 * sorted_prim[~ch[c]])
                                                  : (const volatile float*)(node_aabb + 6 * ch[c]);
            for (int k = 0; k < 3; k++) {
                box[k] = fminf(box[k], src[k]);
                box[3 + k] = fmaxf(box[3 + k], src[3 + k]);
            }
        }
        for (int k = 0; k < 6; k++) node_aabb[6 * node + k] = box[k];
        __threadfence();
        node = parent[node];
    }
}

// 6. pack traversal nodes (children's boxes inside the parent, as centre + half extent: scene.h BvhNode)
// centre = fl(0.5 (lo + hi)); half = max(fl(hi - centre), fl(centre - lo)) scaled up by 1 + 2^-21, which covers the
// rounding of the two subtractions, so [centre - half, centre + half] contains [lo, hi] in exact arithmetic.
__device__ __forceinline__ void centre_half(const float* box, float* c, float* h) {
    for (int k = 0; k < 3; k++) {
        const float lo = box[k], hi = box[3 + k];
        const float m = __fmul_rn(0.5f, __fadd_rn(lo, hi));
        const float e = fmaxf(__fsub_rn(hi, m), __fsub_rn(m, lo));
        c[k] = m;
        h[k] = __fmul_rn(e, 1.0000005f);
    }
}
__global__ void k_pack(const float* __restrict__ aabb, const int* __restrict__ sorted_prim,
                       const int* __restrict__ prim_ref, int n, const int* __restrict__ left,
                       const int* __restrict__ right, const float* __restrict__ node_aabb,
                       BvhNode* __restrict__ nodes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n == 1) {
        if (i == 0) {
            float c[3], h[3];
            centre_half(aabb, c, h);
            BvhNode nd;
            nd.a = make_float4(c[0], 0.0f, c[1], 0.0f);
            nd.b = make_float4(c[2], 0.0f, h[0], -1.0f);  // empty second slot (half = -1): entry > exit for every ray
            nd.c = make_float4(h[1], -1.0f, h[2], -1.0f);
            nd.d = make_int4(~prim_ref[0], ~prim_ref[0], 0, 0);
            nodes[0] = nd;
        }
        return;
    }
    if (i >= n - 1) return;
    int ch[2] = {left[i], right[i]};
    float cc[2][3], hh[2][3];
    int id[2];
    for (int c = 0; c < 2; c++) {
        const float* src;
        if (ch[c] < 0) {
            int p = sorted_prim[~ch[c]];
            src = aabb + 6 * p;
            id[c] = ~prim_ref[p];
        } else {
            src = node_aabb + 6 * ch[c];
            id[c] = ch[c];
        }
        centre_half(src, cc[c], hh[c]);
    }
    BvhNode nd;
    nd.a = make_float4(cc[0][0], cc[1][0], cc[0][1], cc[1][1]);
    nd.b = make_float4(cc[0][2], cc[1][2], hh[0][0], hh[1][0]);
    nd.c = make_float4(hh[0][1], hh[1][1], hh[0][2], hh[1][2]);
    nd.d = make_int4(id[0], id[1], 0, 0);
    nodes[i] = nd;
}

}  // namespace

#define LB_CHECK(x)                          \
    do {                                     \
        cudaError_t e_ = (x);                \
        if (e_ != cudaSuccess) return e_;    \
    } while (0)

cudaError_t lbvh_build(const LbvhBuffers& b, int n, cudaStream_t stream, int* launches) {
    if (n <= 0) return cudaSuccess;
    int nl = 0;
    k_centroid_bounds<<<1, 1024, 0, stream>>>(b.prim_aabb, n, b.bounds);
    nl++;
    int blocks = (n + 255) / 256;
    k_morton<<<blocks, 256, 0, stream>>>(b.prim_aabb, n, b.bounds, b.keys, b.sorted_prim);
    nl++;
    if (n < SORT_MULTI_MIN) {
        k_radix_sort<<<1, SORT_THREADS, 0, stream>>>(b.keys, b.sorted_prim, b.keys_tmp, b.idx_tmp, n);
        nl++;
    } else {
        const int nb = (n + SORT_TILE - 1) / SORT_TILE;  // 256 * nb ints of scratch fit the (n - 1)-int refit counters
        uint64_t *kin = b.keys, *kout = b.keys_tmp;
        int *iin = b.sorted_prim, *iout = b.idx_tmp;
        for (int pass = 0; pass < 8; pass++) {
            k_sort_hist<<<nb, SORT_THREADS, 0, stream>>>(kin, n, pass * 8, nb, b.counters);
            k_sort_scan<<<1, SORT_THREADS, 0, stream>>>(b.counters, 256 * nb);
            k_sort_scatter<<<nb, SORT_THREADS, 0, stream>>>(kin, iin, kout, iout, n, pass * 8, nb, b.counters);
            nl += 3;
            uint64_t* tk = kin; kin = kout; kout = tk;
            int* ti = iin; iin = iout; iout = ti;
        }  // 8 passes (even): the sorted data is back in (keys, sorted_prim)
    }
    if (n >= 2) {
        k_hierarchy<<<(n - 1 + 255) / 256, 256, 0, stream>>>(b.keys, n, b.left, b.right, b.parent);
        nl++;
        LB_CHECK(cudaMemsetAsync(b.counters, 0, sizeof(int) * (n - 1), stream));
        k_refit<<<blocks, 256, 0, stream>>>(b.prim_aabb, b.sorted_prim, n, b.left, b.right, b.parent,
                                            b.node_aabb, b.counters);
        nl++;
    }
    k_pack<<<(n >= 2 ? (n - 1 + 255) / 256 : 1), 256, 0, stream>>>(b.prim_aabb, b.sorted_prim, b.prim_ref, n,
                                                                    b.left, b.right, b.node_aabb, b.nodes);
    nl++;
    if (launches) *launches += nl;
    return cudaGetLastError();
}

}  // namespace rl
