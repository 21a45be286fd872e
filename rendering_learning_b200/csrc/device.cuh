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
    // IEEE sqrt and ONE IEEE division, then three multiplications (1.5 ulp per component; three divisions were 3.6 % of the
    // mirror scene's warp instructions)
    const float inv = 1.0f / sqrtf(dot(a, a));
    return f3(a.x * inv, a.y * inv, a.z * inv);
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

// ---- bounds-asserting debug build (-DRL_DEBUG: librl_b200_debug.so, tools/sanitize_small.py) -------------------------------
// compute-sanitizer is closed on this pool, so the debug library checks every index the kernels form against the scene's
// own counts; a failed check bumps the overflow counter (the render then fails with RL_E_OVERFLOW) instead of reading or
// writing out of bounds.  In the release build RL_CHECK compiles to nothing.
#ifdef RL_DEBUG
#define RL_CHECK(cond, lc)        \
    do {                          \
        if (!(cond)) (lc).overflow += 1u << 20; \
    } while (0)
#define RL_CHECK_OR(cond, lc, stmt) \
    do {                            \
        if (!(cond)) {              \
            (lc).overflow += 1u << 20; \
            stmt;                   \
        }                           \
    } while (0)
#else
#define RL_CHECK(cond, lc) \
    do {                   \
    } while (0)
#define RL_CHECK_OR(cond, lc, stmt) \
    do {                            \
    } while (0)
#endif

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
// Reciprocal direction for the slab tests, kept FINITE: with an exactly zero component 1 / d is +-inf and the FMA form
// of the slab test (centre * inv - o * inv) turns into inf - inf = NaN for every box that straddles the origin's
// coordinate; fminf / fmaxf drop the NaN and the node is falsely missed (axis-aligned cameras, reflections off
// axis-aligned quads).  |d_k| < 1e-20 is replaced by +-1e-20: the slab then spans |t| ~ 1e20 x distance, which is
// "the whole ray" for any coordinate a scene holds, and every product stays far below FLT_MAX.
__device__ __forceinline__ float safe_rcp(float x) {
    return 1.0f / (fabsf(x) < 1e-20f ? copysignf(1e-20f, x) : x);
}
__device__ __forceinline__ float3 safe_inv(float3 d) { return f3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z)); }

__device__ __forceinline__ RayPre make_pre(float3 o, float3 d) {
    RayPre r;
    r.o = o;
    r.d = d;
    r.inv_d = safe_inv(d);
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
        RL_CHECK_OR(node >= 0 && node < (n_bvh_prims > 1 ? n_bvh_prims - 1 : 1), lc, return);
        const float4* np = reinterpret_cast<const float4*>(nodes + node);
        float4 a = np[0], b = np[1], c = np[2];
        int4 d = *reinterpret_cast<const int4*>(np + 3);
        if (COUNT) lc.nodes++;
        // slabs of both children (scene.h BvhNode: centre + half extent, the children's values paired per coordinate;
        // entry / exit = centre -+ half * |1 / d|)
        const float aix = fabsf(r.inv_d.x), aiy = fabsf(r.inv_d.y), aiz = fabsf(r.inv_d.z);
        float c0x = (a.x - r.o.x) * r.inv_d.x, c0y = (a.z - r.o.y) * r.inv_d.y, c0z = (b.x - r.o.z) * r.inv_d.z;
        float n0 = fmaxf(fmaxf(fmaf(-b.z, aix, c0x), fmaf(-c.x, aiy, c0y)), fmaxf(fmaf(-c.z, aiz, c0z), tmin));
        float f0 = fminf(fminf(fmaf(b.z, aix, c0x), fmaf(c.x, aiy, c0y)), fminf(fmaf(c.z, aiz, c0z), tmax));
        float c1x = (a.y - r.o.x) * r.inv_d.x, c1y = (a.w - r.o.y) * r.inv_d.y, c1z = (b.y - r.o.z) * r.inv_d.z;
        float n1 = fmaxf(fmaxf(fmaf(-b.w, aix, c1x), fmaf(-c.y, aiy, c1y)), fmaxf(fmaf(-c.w, aiz, c1z), tmin));
        float f1 = fminf(fminf(fmaf(b.w, aix, c1x), fmaf(c.y, aiy, c1y)), fminf(fmaf(c.w, aiz, c1z), tmax));
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

// ---- resumable traversal (OW render / trace kernel) --------------------------------------------------------------
constexpr int TRAV_END = (int)0x80000000;

// Per-lane traversal stack.  The first STACK_SM entries live in shared memory as [entry][thread] (bank = thread: every
// push / pop of a warp is ONE conflict-free wavefront whatever the lanes' depths are); deeper entries spill to a local
// array that only the rare deep path touches.  The stack pointer `top` is a plain register holding the ADDRESS of the
// next free shared-memory entry (so push / pop are one predicated STS / LDS and one add, no index arithmetic) while the
// depth is <= STACK_SM; `deep` counts the entries beyond that.
// History (ncu, cover scene): round 1's all-local stack cost 803 M LDL / STL requests per launch and the pop was the top
// stall line; a struct {sp, spill[]} version had ptxas put sp itself into local memory; three divergent paths for push /
// single / pop were 21 % of ALL warp instructions at ~5 of 32 lanes (profiles/r02_ncu_v5_c4_struct_stack.json).
constexpr int STACK_SM = 8;
struct TravStack {
    unsigned top;   // shared-space ADDRESS of the next free entry of this thread's column; beyond `base + STACK_SM rows` it
                    // keeps counting in the same units and the entries live in the spill array
    unsigned base;  // shared-space address of the column (entry e at base + e * THREADS * 4)
};
struct StackSpill {
    int loc[BVH_STACK - STACK_SM];
};
template <int THREADS>
__device__ __forceinline__ void stack_init(TravStack& st, const int* column) {
    st.base = (unsigned)__cvta_generic_to_shared(column);
    asm volatile("" : "+r"(st.base));  // opaque: otherwise ptxas re-derives it from %tid in every node step (6 instructions)
    st.top = st.base;
}
__device__ __forceinline__ void stack_reset(TravStack& st) { st.top = st.base; }
// predicated shared-memory accesses, pinned with inline PTX (left to itself ptxas re-derived the column address from
// %tid in every node step and branched around the store)
__device__ __forceinline__ void sts_if(unsigned addr, int v, bool p) {
    asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; @q st.shared.b32 [%0], %1; }" ::"r"(addr), "r"(v), "r"((int)p) : "memory");
}
__device__ __forceinline__ int lds_if(unsigned addr, bool p, int otherwise) {
    int r;
    asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; mov.b32 %0, %3; @q ld.shared.b32 %0, [%1]; }"
                 : "=r"(r) : "r"(addr), "r"((int)p), "r"(otherwise) : "memory");
    return r;
}

// approximate reciprocal (one MUFU.RCP): the slab test is conservative by construction (slack()), and one ulp on 1 / d
// scales entry and exit of an axis together
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float3 safe_inv_fast(float3 d) {
    return f3(fast_rcp(fabsf(d.x) < 1e-20f ? copysignf(1e-20f, d.x) : d.x),
              fast_rcp(fabsf(d.y) < 1e-20f ? copysignf(1e-20f, d.y) : d.y),
              fast_rcp(fabsf(d.z) < 1e-20f ? copysignf(1e-20f, d.z) : d.z));
}

// pop (after a leaf test, or from the rare deep path of a node step); TRAV_END when the stack is empty
template <int THREADS>
__device__ __forceinline__ int stack_pop(TravStack& st, const StackSpill& spill) {
    constexpr unsigned ROW = THREADS * 4u, LIMIT = STACK_SM * ROW;
    if (st.top == st.base) return TRAV_END;
    st.top -= ROW;
    const unsigned off = st.top - st.base;
    if (off >= LIMIT) return spill.loc[(off - LIMIT) / ROW];
    return lds_if(st.top, true, TRAV_END);
}
template <int THREADS>
__device__ __forceinline__ bool stack_push_deep(TravStack& st, StackSpill& spill, int v) {  // top is at or beyond the column's end
    constexpr unsigned ROW = THREADS * 4u, LIMIT = STACK_SM * ROW;
    const unsigned e = (st.top - st.base - LIMIT) / ROW;
    if (e >= (unsigned)(BVH_STACK - STACK_SM)) return false;
    spill.loc[e] = v;
    st.top += ROW;
    return true;
}

// packed dual FP32 FMA (sm_100 FFMA2): (a.x, a.y) * s + (c.x, c.y) and (a.x, a.y) * s + t with scalar s, t
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra, rb, rc, rd;
    ra = *reinterpret_cast<unsigned long long*>(&a);
    rb = *reinterpret_cast<unsigned long long*>(&b);
    rc = *reinterpret_cast<unsigned long long*>(&c);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 ffma2(float2 a, float s, float2 c) { return ffma2(a, make_float2(s, s), c); }
__device__ __forceinline__ float2 ffma2(float2 a, float s, float t) { return ffma2(a, make_float2(s, s), make_float2(t, t)); }

// ONE node step of a lane standing at inner node `node`: slab-test both children, step into the nearer hit child and
// push the other one, or pop.
//  * Boxes are stored as centre + half extent (scene.h), so entry / exit per axis are fma(-+half, |1/d|, fma(centre, 1/d,
//    -o/d)): three FMAs per axis and child and NO per-axis min / max — round 1's lo / hi form spent 20 FMNMX per step on
//    the ALU pipe (61 % busy, the limiter) against 12 FFMA.  The two children's values sit in aligned register pairs,
//    so each of those FMAs is ONE FFMA2 for both children: 10 FFMA2 + 8 FMNMX(3) per step, where round 1 had 14 + 20.
//  * The push / pop tail has no divergent paths in the common case: selects, one predicated shared-memory store or load
//    and one predicated pointer add.  Only a lane whose stack is STACK_SM deep takes the generic path.
//  * A popped node is not culled against the current hit: its children fail their own slab tests.
// One 32-byte load per lane (sm_100 LDG.E.256): a 64-byte node is 2 load instructions instead of 4.  Every lane of a
// divergent load touches its own 128-byte line and the L1 spends one wavefront per line PER INSTRUCTION, so halving the
// instructions halves the LSU wavefronts — the pipe ncu shows at 84 % on the cover scene (profiles/r02_ncu_ow_c4_500.json).
// p must be 32-byte aligned and the data read-only for the kernel's lifetime (ld.global.nc).
#ifndef RL_LDG256
#define RL_LDG256 1
#endif
__device__ __forceinline__ void ldg256(const void* p, float4& lo, float4& hi) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
                 : "l"(p));
}

template <bool COUNT, int THREADS>
__device__ __forceinline__ void bvh2_step(const BvhNode* __restrict__ nodes, int& node, TravStack& st, StackSpill& spill,
                                          const float3 inv_d, const float3 oi, const float tmin, const float tmax,
                                          LocalCount<COUNT>& lc, const int n_nodes = 0x7fffffff) {
    RL_CHECK_OR(node >= 0 && node < n_nodes, lc, { node = TRAV_END; return; });
#if RL_LDG256
    float4 a, b, c, dd;
    ldg256(nodes + node, a, b);
    ldg256(reinterpret_cast<const char*>(nodes + node) + 32, c, dd);
    const int2 d = make_int2(__float_as_int(dd.x), __float_as_int(dd.y));
#else
    const float4* np = reinterpret_cast<const float4*>(nodes + node);
    const float4 a = np[0], b = np[1], c = np[2];
    const int2 d = *reinterpret_cast<const int2*>(np + 3);
#endif
    if (COUNT) lc.nodes++;
    // both children per instruction: (c0, c1) pairs x scalar 1/d (FFMA2 broadcasts a scalar operand)
    const float2 tcx = ffma2(make_float2(a.x, a.y), inv_d.x, -oi.x);
    const float2 tcy = ffma2(make_float2(a.z, a.w), inv_d.y, -oi.y);
    const float2 tcz = ffma2(make_float2(b.x, b.y), inv_d.z, -oi.z);
    const float2 hx = make_float2(b.z, b.w), hy = make_float2(c.x, c.y), hz = make_float2(c.z, c.w);
    const float aix = fabsf(inv_d.x), aiy = fabsf(inv_d.y), aiz = fabsf(inv_d.z);
    const float2 nx = ffma2(hx, -aix, tcx), ny = ffma2(hy, -aiy, tcy), nz = ffma2(hz, -aiz, tcz);
    const float2 fx = ffma2(hx, aix, tcx), fy = ffma2(hy, aiy, tcy), fz = ffma2(hz, aiz, tcz);
    const float n0 = fmaxf(fmaxf(nx.x, ny.x), fmaxf(nz.x, tmin)), n1 = fmaxf(fmaxf(nx.y, ny.y), fmaxf(nz.y, tmin));
    const float2 f = make_float2(fminf(fminf(fx.x, fy.x), fminf(fz.x, tmax)), fminf(fminf(fx.y, fy.y), fminf(fz.y, tmax)));
    const float2 fs = ffma2(make_float2(fabsf(f.x), fabsf(f.y)), 5e-7f, f);  // slack() of both
    const bool h0 = n0 <= fs.x, h1 = n1 <= fs.y;
    const bool both = h0 && h1, any = h0 || h1;
    const bool second_first = h1 && (!h0 || n1 < n0);
    const int nearc = second_first ? d.y : d.x;
    const int farc = second_first ? d.x : d.y;
    constexpr unsigned ROW = THREADS * 4u, LIMIT = STACK_SM * ROW;
    if (st.top - st.base >= LIMIT) {  // rare: the column is full, the spill rows are involved
        if (both) {
            if (!stack_push_deep<THREADS>(st, spill, farc)) lc.overflow++;
            node = nearc;
        } else {
            node = any ? nearc : stack_pop<THREADS>(st, spill);
        }
        return;
    }
    sts_if(st.top, farc, both);
    st.top += both ? ROW : 0u;
    const bool popping = !any && st.top != st.base;
    st.top -= popping ? ROW : 0u;
    const int popped = lds_if(st.top, popping, TRAV_END);
    node = any ? nearc : popped;
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
