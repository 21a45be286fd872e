// Device-side building blocks shared by the RTC and OW kernels: f32 vector math, scene staging into
// shared memory, analytic primitive roots, watertight ray-triangle, stack-based BVH2 traversal.
#pragma once
#include <cfloat>
#include <cstdint>

#include "scene.h"

namespace rl {

#define RL_INF __int_as_float(0x7f800000)

// ---- vector math ----------------------------------------------------------------------------------
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 f3(float4 v) { return make_float3(v.x, v.y, v.z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float dot(float3 a, float3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ float3 cross(float3 a, float3 b) {
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float3 fma3(float3 a, float s, float3 b) {  // a*s + b
    return f3(fmaf(a.x, s, b.x), fmaf(a.y, s, b.y), fmaf(a.z, s, b.z));
}
__device__ __forceinline__ float3 normalize(float3 a) {
    float inv = rsqrtf(dot(a, a));
    return a * inv;
}
// exact-ish normalisation (sqrt + div), used where the reference's result feeds thresholds
__device__ __forceinline__ float3 normalize_precise(float3 a) {
    float m = sqrtf(dot(a, a));
    return f3(a.x / m, a.y / m, a.z / m);
}
__device__ __forceinline__ float comp(float3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
__device__ __forceinline__ float max_abs(float3 v) { return fmaxf(fabsf(v.x), fmaxf(fabsf(v.y), fabsf(v.z))); }

__device__ __forceinline__ float3 xf_point(const float4 r[3], float3 p) {
    return f3(fmaf(r[0].x, p.x, fmaf(r[0].y, p.y, fmaf(r[0].z, p.z, r[0].w))),
              fmaf(r[1].x, p.x, fmaf(r[1].y, p.y, fmaf(r[1].z, p.z, r[1].w))),
              fmaf(r[2].x, p.x, fmaf(r[2].y, p.y, fmaf(r[2].z, p.z, r[2].w))));
}
__device__ __forceinline__ float3 xf_vec(const float4 r[3], float3 v) {
    return f3(fmaf(r[0].x, v.x, fmaf(r[0].y, v.y, r[0].z * v.z)),
              fmaf(r[1].x, v.x, fmaf(r[1].y, v.y, r[1].z * v.z)),
              fmaf(r[2].x, v.x, fmaf(r[2].y, v.y, r[2].z * v.z)));
}
// n_world = inv^T * n_local
__device__ __forceinline__ float3 xf_normal(const float4 r[3], float3 n) {
    return f3(fmaf(r[0].x, n.x, fmaf(r[1].x, n.y, r[2].x * n.z)),
              fmaf(r[0].y, n.x, fmaf(r[1].y, n.y, r[2].y * n.z)),
              fmaf(r[0].z, n.x, fmaf(r[1].z, n.y, r[2].z * n.z)));
}

// ---- counters (instrumented builds only) ----------------------------------------------------------------
struct Counters {
    unsigned long long rays, node_visits, prim_tests, tri_tests, shades, overflow;
};
template <bool COUNT>
struct LocalCount {
    unsigned rays = 0, nodes = 0, prims = 0, tris = 0, shades = 0, overflow = 0;
    __device__ __forceinline__ void flush(Counters* c) {
        if (COUNT) {
            // warp-aggregate, one atomic per warp per counter
            unsigned v[6] = {rays, nodes, prims, tris, shades, overflow};
            unsigned long long* dst = &c->rays;
            for (int k = 0; k < 6; k++) {
                unsigned s = v[k];
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if ((threadIdx.x & 31) == 0 && s) atomicAdd(dst + k, (unsigned long long)s);
            }
        } else {
            // overflow is always reported
            unsigned s = overflow;
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if ((threadIdx.x & 31) == 0 && s) atomicAdd(&c->overflow, (unsigned long long)s);
        }
    }
};

// ---- rays ------------------------------------------------------------------------------------------
struct Ray {
    float3 o, d;
};

// per-ray constants of the watertight triangle test (Woop, Benthin, Wald 2013) and the slab test
struct RayPre {
    float3 o, d, inv_d;
    int kx, ky, kz;
    float Sx, Sy, Sz;
};
__device__ __forceinline__ RayPre make_pre(float3 o, float3 d) {
    RayPre r;
    r.o = o;
    r.d = d;
    r.inv_d = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int kz = (ax > ay) ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
    int kx = kz == 2 ? 0 : kz + 1;
    int ky = kx == 2 ? 0 : kx + 1;
    float dz = comp(d, kz);
    if (dz < 0.0f) { int t = kx; kx = ky; ky = t; }
    r.kx = kx; r.ky = ky; r.kz = kz;
    r.Sx = comp(d, kx) / dz;
    r.Sy = comp(d, ky) / dz;
    r.Sz = 1.0f / dz;
    return r;
}

// watertight ray-triangle. returns true and t, and barycentrics (b1 = weight of p1, b2 = weight of p2)
__device__ __forceinline__ bool tri_hit(const RayPre& r, float3 p0, float3 p1, float3 p2, float* t, float* b1,
                                        float* b2) {
    float3 A = p0 - r.o, B = p1 - r.o, C = p2 - r.o;
    float Akz = comp(A, r.kz), Bkz = comp(B, r.kz), Ckz = comp(C, r.kz);
    float Ax = fmaf(-r.Sx, Akz, comp(A, r.kx)), Ay = fmaf(-r.Sy, Akz, comp(A, r.ky));
    float Bx = fmaf(-r.Sx, Bkz, comp(B, r.kx)), By = fmaf(-r.Sy, Bkz, comp(B, r.ky));
    float Cx = fmaf(-r.Sx, Ckz, comp(C, r.kx)), Cy = fmaf(-r.Sy, Ckz, comp(C, r.ky));
    float U = __fmul_rn(Cx, By) - __fmul_rn(Cy, Bx);
    float V = __fmul_rn(Ax, Cy) - __fmul_rn(Ay, Cx);
    float W = __fmul_rn(Bx, Ay) - __fmul_rn(By, Ax);
    if (U == 0.0f || V == 0.0f || W == 0.0f) {  // exact edge: redo the edge functions in double
        double CxBy = (double)Cx * (double)By, CyBx = (double)Cy * (double)Bx;
        U = (float)(CxBy - CyBx);
        double AxCy = (double)Ax * (double)Cy, AyCx = (double)Ay * (double)Cx;
        V = (float)(AxCy - AyCx);
        double BxAy = (double)Bx * (double)Ay, ByAx = (double)By * (double)Ax;
        W = (float)(BxAy - ByAx);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    float det = U + V + W;
    if (det == 0.0f) return false;
    float Az = r.Sz * Akz, Bz = r.Sz * Bkz, Cz = r.Sz * Ckz;
    float T = fmaf(U, Az, fmaf(V, Bz, W * Cz));
    float rdet = 1.0f / det;
    *t = T * rdet;
    *b1 = V * rdet;
    *b2 = W * rdet;
    return true;
}

// ---- BVH2 traversal ----------------------------------------------------------------------------------
constexpr int BVH_STACK = 64;
__device__ __forceinline__ float slack(float f) { return fmaf(fabsf(f), 5e-7f, f); }

// Visits every leaf whose box overlaps [tmin, tmax] along the ray.  `leaf(ref, tmax)` returns the new
// tmax (closest-hit queries shrink it; enumeration queries return it unchanged; any-hit queries return
// a negative value to stop).  NodeLoad abstracts global vs shared memory residency of the nodes.
template <bool COUNT, class LeafFn>
__device__ __forceinline__ void bvh_traverse(const BvhNode* __restrict__ nodes, int n_bvh_prims, const RayPre& r,
                                             float tmin, float tmax, LocalCount<COUNT>& lc, LeafFn leaf) {
    if (n_bvh_prims <= 0) return;
    int stack_node[BVH_STACK];
    float stack_t[BVH_STACK];
    int sp = 0;
    int node = 0;
    while (true) {
        const float4* np = reinterpret_cast<const float4*>(nodes + node);
        float4 a = np[0], b = np[1], c = np[2];
        int4 d = *reinterpret_cast<const int4*>(np + 3);
        if (COUNT) lc.nodes++;
        // slabs of both children
        float t0x = (a.x - r.o.x) * r.inv_d.x, t1x = (a.w - r.o.x) * r.inv_d.x;
        float t0y = (a.y - r.o.y) * r.inv_d.y, t1y = (b.x - r.o.y) * r.inv_d.y;
        float t0z = (a.z - r.o.z) * r.inv_d.z, t1z = (b.y - r.o.z) * r.inv_d.z;
        float n0 = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), tmin));
        float f0 = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), tmax));
        float u0x = (b.z - r.o.x) * r.inv_d.x, u1x = (c.y - r.o.x) * r.inv_d.x;
        float u0y = (b.w - r.o.y) * r.inv_d.y, u1y = (c.z - r.o.y) * r.inv_d.y;
        float u0z = (c.x - r.o.z) * r.inv_d.z, u1z = (c.w - r.o.z) * r.inv_d.z;
        float n1 = fmaxf(fmaxf(fminf(u0x, u1x), fminf(u0y, u1y)), fmaxf(fminf(u0z, u1z), tmin));
        float f1 = fminf(fminf(fmaxf(u0x, u1x), fmaxf(u0y, u1y)), fminf(fmaxf(u0z, u1z), tmax));
        // a few ulp of slack keeps the f32 slab test conservative (Ize 2013)
        bool h0 = n0 <= slack(f0), h1 = n1 <= slack(f1);
        if (h0 && d.x < 0) {
            tmax = leaf(~d.x, tmax);
            if (tmax == -RL_INF) return;  // any-hit queries stop here
            h0 = false;
            h1 = h1 && n1 <= slack(tmax);
        }
        if (h1 && d.y < 0) {
            tmax = leaf(~d.y, tmax);
            if (tmax == -RL_INF) return;
            h1 = false;
        }
        if (h0 && h1) {
            int nearc = d.x, farc = d.y;
            float tf = n1;
            if (n1 < n0) { nearc = d.y; farc = d.x; tf = n0; }
            if (sp < BVH_STACK) {
                stack_node[sp] = farc;
                stack_t[sp] = tf;
                sp++;
            } else {
                lc.overflow++;
            }
            node = nearc;
            continue;
        }
        if (h0) { node = d.x; continue; }
        if (h1) { node = d.y; continue; }
        // pop
        bool found = false;
        while (sp > 0) {
            sp--;
            if (stack_t[sp] <= slack(tmax)) {
                node = stack_node[sp];
                found = true;
                break;
            }
        }
        if (!found) return;
    }
}

// "while-while" traversal (Aila & Laine 2009): the inner loop only descends inner nodes; a lane that reaches a
// leaf parks until every lane of the warp has one (or is done), then the leaves are tested together.  ncu on the
// if-if loop above showed primitive tests running with 2-4 of 32 lanes and the pop loop with < 3
// (profiles/r01_ncu_k_ow_render_v1.json); here the slab tests use the FMA form (lo * inv - o * inv) and the pop
// is a single predicated stack read — a popped node whose entry distance is now beyond tmax fails its own slab
// test, so no cull loop is needed.
constexpr int TRAV_END = (int)0x80000000;
template <bool COUNT, class LeafFn>
__device__ __forceinline__ void bvh_traverse_ww(const BvhNode* __restrict__ nodes, int n_bvh_prims, const RayPre& r,
                                                float tmin, float tmax, LocalCount<COUNT>& lc, LeafFn leaf) {
    if (n_bvh_prims <= 0) return;
    int stack_node[BVH_STACK];
    int sp = 0;
    int node = 0;
    const float3 oi = f3(r.o.x * r.inv_d.x, r.o.y * r.inv_d.y, r.o.z * r.inv_d.z);
    while (node != TRAV_END) {
        while (node >= 0) {
            const float4* np = reinterpret_cast<const float4*>(nodes + node);
            float4 a = np[0], b = np[1], c = np[2];
            int4 d = *reinterpret_cast<const int4*>(np + 3);
            if (COUNT) lc.nodes++;
            float t0x = fmaf(a.x, r.inv_d.x, -oi.x), t1x = fmaf(a.w, r.inv_d.x, -oi.x);
            float t0y = fmaf(a.y, r.inv_d.y, -oi.y), t1y = fmaf(b.x, r.inv_d.y, -oi.y);
            float t0z = fmaf(a.z, r.inv_d.z, -oi.z), t1z = fmaf(b.y, r.inv_d.z, -oi.z);
            float n0 = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), tmin));
            float f0 = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), tmax));
            float u0x = fmaf(b.z, r.inv_d.x, -oi.x), u1x = fmaf(c.y, r.inv_d.x, -oi.x);
            float u0y = fmaf(b.w, r.inv_d.y, -oi.y), u1y = fmaf(c.z, r.inv_d.y, -oi.y);
            float u0z = fmaf(c.x, r.inv_d.z, -oi.z), u1z = fmaf(c.w, r.inv_d.z, -oi.z);
            float n1 = fmaxf(fmaxf(fminf(u0x, u1x), fminf(u0y, u1y)), fmaxf(fminf(u0z, u1z), tmin));
            float f1 = fminf(fminf(fmaxf(u0x, u1x), fmaxf(u0y, u1y)), fminf(fmaxf(u0z, u1z), tmax));
            bool h0 = n0 <= slack(f0), h1 = n1 <= slack(f1);
            int nearc = d.x, farc = d.y;
            if (n1 < n0) { nearc = d.y; farc = d.x; }
            if (h0 && h1) {
                if (sp < BVH_STACK) stack_node[sp++] = farc; else lc.overflow++;
                node = nearc;
            } else if (h0 || h1) {
                node = h0 ? d.x : d.y;
            } else {
                node = sp > 0 ? stack_node[--sp] : TRAV_END;
            }
        }
        if (node != TRAV_END) {
            tmax = leaf(~node, tmax);
            if (tmax == -RL_INF) return;
            node = sp > 0 ? stack_node[--sp] : TRAV_END;
        }
    }
}

// ---- work distribution ----------------------------------------------------------------------------------
// jobs = pixel rectangles (x sample-chunk ranges for OW); item enumeration walks 8x4 pixel micro-tiles so
// that the 32 lanes of a warp start on neighbouring pixels.
constexpr int JOBS_INLINE = 8;
struct JobTable {
    const rl_job* jobs;      // device copy (used when n_jobs > JOBS_INLINE)
    const long long* prefix; // [n_jobs + 1] exclusive prefix of item counts
    int n_jobs;
    int inline_jobs;         // 1: the table travels in the kernel parameters (no device buffer, async-safe)
    long long n_items;
    rl_job ijobs[JOBS_INLINE];
    long long iprefix[JOBS_INLINE + 1];
};
__device__ __forceinline__ long long jt_prefix(const JobTable& jt, int j) {
    return jt.inline_jobs ? jt.iprefix[j] : jt.prefix[j];
}
__device__ __forceinline__ rl_job jt_job(const JobTable& jt, int j) { return jt.inline_jobs ? jt.ijobs[j] : jt.jobs[j]; }

__device__ __forceinline__ int find_job(const JobTable& jt, long long item) {
    int lo = 0, hi = jt.n_jobs - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (jt_prefix(jt, mid) <= item) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// pixel index within a w x h rectangle, enumerated in 8x4 micro-tiles (row-major tiles, row-major inside)
__device__ __forceinline__ void tile_pixel(int w, int h, long long p, int* x, int* y) {
    int tiles_x = (w + 7) >> 3;
    long long band_px = (long long)tiles_x * 32;           // padded pixels per band
    long long band = p / band_px;
    int in_band = (int)(p - band * band_px);
    int tile = in_band >> 5, lane = in_band & 31;
    *x = tile * 8 + (lane & 7);
    *y = (int)band * 4 + (lane >> 3);
}
__host__ __device__ __forceinline__ long long padded_pixels(int w, int h) {
    return (long long)((w + 7) >> 3) * 32 * (long long)((h + 3) >> 2);
}

}  // namespace rl
