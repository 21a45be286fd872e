// RTC (ray-tracer-challenge) render path: deterministic Whitted tracer in f32.
//
// Replaces, per pixel, the recursion
//   Camera::render -> rays_for_pixel -> World::color_at -> intersect / hit / prepare_computations ->
//   shade_hit -> shadow_attenuation / lighting / reflected_color / refracted_color
// (RTC/src/scene/camera.rs:63-124, world.rs:46-159, intersect.rs:47-168, material.rs:54-90).
//
// The recursion `colour = surface + r*reflect + t*refract` is linear, so it is unrolled into an explicit
// per-thread stack of (ray, scalar weight, remaining depth) entries (<= max_reflection_depth + 1 live).
// Sorted intersection lists are never materialised:
//   * closest hit   = min over (t, leaf order) with t >= 0, later leaf wins ties (intersect.rs:159-168);
//   * n1 / n2       = streaming container analysis: an object is "open" at the hit iff it has an odd number
//                     of crossings before the hit, and the open objects are ordered by their last crossing
//                     (equivalent to the Vec push / remove walk of intersect.rs:72-99);
//   * shadows       = any-hit when no material is transparent, else the first-occurrence product of
//                     world.rs:116-123 evaluated in two streaming passes.
#include "device.cuh"
#include "kernels.h"

namespace rl {
namespace {

constexpr float RTC_EPS = 1e-8f;       // plane / cylinder / cone epsilons of the reference
constexpr float RTC_BIAS = 1e-5f;      // POINT_OFFSET_BIAS (intersect.rs:8)
// Pending (ray, weight, depth) entries of the unrolled recursion.  An entry is popped before its children are pushed, so the
// stack only grows where a hit pushes BOTH a reflected and a refracted ray: a mirror-only chain of any depth needs one
// entry.  66 entries = 64 such levels pending at once (a tree the reference itself could not finish); beyond that the
// render reports RL_E_OVERFLOW instead of a wrong image.  max_reflection_depth itself is unbounded, as in world.rs:29.
constexpr int RTC_STACK = 66;

struct RtcCam {
    int hsize, vsize;
    float pixel_size, half_width, half_height;
    float4 inv[3];  // camera -> world
    int aa;
};

// ---- analytic primitives -------------------------------------------------------------------------------
// roots in the reference's push order; returns the count (<= 4)
// tags: 2 bits per root — 0 wall / body, 1 lower cap, 2 upper cap (f32 cannot re-derive the cap from the
// hit point with the reference's 1e-8 window, so the root remembers what it is)
// Root divisions: x * rcp.approx(y), 2 ulp, same inf / NaN behaviour as IEEE for y = 0 (the cube's axis-parallel rays).  The
// IEEE sequence was ~15 % of the mirror scene's warp instructions (profiles/r02_lines_rtc_c2.txt: plane, cube and sphere roots);
// RL_RTC_IEEE_DIV=1 restores it (tools/build_alt.py A/B).
#ifndef RL_RTC_IEEE_DIV
#define RL_RTC_IEEE_DIV 0
#endif
__device__ __forceinline__ float fdiv(float x, float y) {
#if RL_RTC_IEEE_DIV
    return x / y;
#else
    return __fdividef(x, y);
#endif
}
#ifndef RL_RTC_NOINLINE  // experiment switch (tools/build_alt.py): 1 = prim_roots out of line, 2 = + closest / shadow / n1-n2
#define RL_RTC_NOINLINE 0
#endif
#if RL_RTC_NOINLINE >= 1
#define RL_ROOTS_FN __device__ __noinline__
#else
#define RL_ROOTS_FN __device__ __forceinline__
#endif
#if RL_RTC_NOINLINE >= 2
#define RL_TRACER_FN __device__ __noinline__
#else
#define RL_TRACER_FN __device__
#endif
RL_ROOTS_FN int prim_roots(const RtcPrim& p, float3 o, float3 d, float ts[4], unsigned* tags) {
    int n = 0;
    *tags = 0u;
    switch (p.kind) {
        case PK_RTC_SPHERE: {  // sphere.rs:35-59 (numerically robust quarter-discriminant form)
            float a = dot(d, d);
            float hb = dot(d, o);
            float tc = fdiv(-hb, a);
            float3 perp = fma3(d, tc, o);
            float disc = a * (1.0f - dot(perp, perp));
            if (disc >= 0.0f) {
                float q = fdiv(sqrtf(disc), a);
                ts[0] = tc - q;
                ts[1] = tc + q;
                n = 2;
            }
            break;
        }
        case PK_RTC_PLANE: {  // plane.rs:26-39
            if (!(fabsf(d.y) < RTC_EPS)) {
                ts[0] = fdiv(-o.y, d.y);
                n = 1;
            }
            break;
        }
        case PK_RTC_CUBE: {  // cube.rs:38-79
            float ax = fdiv(-1.0f - o.x, d.x), bx = fdiv(1.0f - o.x, d.x);
            float ay = fdiv(-1.0f - o.y, d.y), by = fdiv(1.0f - o.y, d.y);
            float az = fdiv(-1.0f - o.z, d.z), bz = fdiv(1.0f - o.z, d.z);
            float tmin = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
            float tmax = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
            if (!(tmin > tmax)) {
                ts[0] = tmin;
                ts[1] = tmax;
                n = 2;
            }
            break;
        }
        case PK_RTC_CYLINDER: {  // cylinder.rs:94-139, caps 30-64
            float a = d.x * d.x + d.z * d.z;
            if (!(fabsf(a) < RTC_EPS)) {
                float hb = o.x * d.x + o.z * d.z;
                float tc = fdiv(-hb, a);
                float px = fmaf(d.x, tc, o.x), pz = fmaf(d.z, tc, o.z);
                float disc = a * (1.0f - (px * px + pz * pz));
                if (disc >= 0.0f) {
                    float q = fdiv(sqrtf(disc), a);
                    float t0 = tc - q, t1 = tc + q;
                    float y0 = fmaf(t0, d.y, o.y);
                    if (y0 > p.ymin && y0 < p.ymax) ts[n++] = t0;
                    float y1 = fmaf(t1, d.y, o.y);
                    if (y1 > p.ymin && y1 < p.ymax) ts[n++] = t1;
                }
            }
            if ((p.flags & 1) && !(fabsf(d.y) < RTC_EPS)) {
                if (p.ymin > -RL_INF) {
                    float t = fdiv(p.ymin - o.y, d.y);
                    float x = fmaf(t, d.x, o.x), z = fmaf(t, d.z, o.z);
                    if (x * x + z * z <= 1.0f) { *tags |= 1u << (2 * n); ts[n++] = t; }
                }
                if (p.ymax < RL_INF) {
                    float t = fdiv(p.ymax - o.y, d.y);
                    float x = fmaf(t, d.x, o.x), z = fmaf(t, d.z, o.z);
                    if (x * x + z * z <= 1.0f) { *tags |= 2u << (2 * n); ts[n++] = t; }
                }
            }
            break;
        }
        case PK_RTC_CONE: {  // cone.rs:88-151, caps 29-63 (cap test `<= |y|` as written)
            float a = d.x * d.x - d.y * d.y + d.z * d.z;
            float b = 2.0f * (o.x * d.x - o.y * d.y + o.z * d.z);
            float c = o.x * o.x - o.y * o.y + o.z * o.z;
            bool a0 = fabsf(a) < RTC_EPS, b0 = fabsf(b) < RTC_EPS;
            if (a0 && b0) {
            } else if (a0) {
                ts[n++] = fdiv(-c, 2.0f * b);
            } else {
                float disc = b * b - 4.0f * a * c;
                if (disc >= 0.0f) {
                    float s = sqrtf(disc);
                    float t0 = fdiv(-b - s, 2.0f * a), t1 = fdiv(-b + s, 2.0f * a);
                    float y0 = fmaf(t0, d.y, o.y);
                    if (y0 > p.ymin && y0 < p.ymax) ts[n++] = t0;
                    float y1 = fmaf(t1, d.y, o.y);
                    if (y1 > p.ymin && y1 < p.ymax) ts[n++] = t1;
                }
            }
            if ((p.flags & 1) && !(fabsf(d.y) < RTC_EPS)) {
                if (p.ymin > -RL_INF) {
                    float t = fdiv(p.ymin - o.y, d.y);
                    float x = fmaf(t, d.x, o.x), z = fmaf(t, d.z, o.z);
                    if (x * x + z * z <= fabsf(p.ymin)) { *tags |= 1u << (2 * n); ts[n++] = t; }
                }
                if (p.ymax < RL_INF) {
                    float t = fdiv(p.ymax - o.y, d.y);
                    float x = fmaf(t, d.x, o.x), z = fmaf(t, d.z, o.z);
                    if (x * x + z * z <= fabsf(p.ymax)) { *tags |= 2u << (2 * n); ts[n++] = t; }
                }
            }
            break;
        }
    }
    return n;
}

// A triangle that is a leaf of a Csg (csg.rs:31-35 is generic over any Object) travels as an analytic primitive: its
// world-space vertices in inv[0..2], its normals in fwd[0..2] (flat: fwd[0] = the face normal), world -> pattern space in
// pat, flags bit0 = smooth.  Same watertight test as the LBVH triangles; one root, of either sign (triangle.rs:63-101).
__device__ __forceinline__ int tri_prim_roots(const RtcPrim& p, float3 o, float3 d, float ts[4], float* b1, float* b2) {
    RayPre pre = make_pre(o, d);
    float t, u, v;
    if (!tri_hit(pre, f3(p.inv[0]), f3(p.inv[1]), f3(p.inv[2]), &t, &u, &v)) return 0;
    ts[0] = t;
    *b1 = u;
    *b2 = v;
    return 1;
}

// local normal (PhysicalObject::normal_at) and the local hit point snapped back onto the surface, so the
// world-space point carries ~1 ulp of error instead of the error of t (keeps the 1e-5 bias meaningful)
__device__ __forceinline__ float3 prim_normal(const RtcPrim& p, float3& q, int tag) {
    switch (p.kind) {
        case PK_RTC_SPHERE: {  // sphere.rs:24-28
            float3 n = normalize_precise(q);
            q = n;
            return n;
        }
        case PK_RTC_PLANE:  // plane.rs:16-20
            q.y = 0.0f;
            return f3(0.0f, 1.0f, 0.0f);
        case PK_RTC_CUBE: {  // cube.rs:14-31
            float ax = fabsf(q.x), ay = fabsf(q.y), az = fabsf(q.z);
            float mc = fmaxf(ax, fmaxf(ay, az));
            if (mc == ax) { q.x = copysignf(1.0f, q.x); return f3(q.x, 0.0f, 0.0f); }
            if (mc == ay) { q.y = copysignf(1.0f, q.y); return f3(0.0f, q.y, 0.0f); }
            q.z = copysignf(1.0f, q.z);
            return f3(0.0f, 0.0f, q.z);
        }
        case PK_RTC_CYLINDER: {  // cylinder.rs:67-87
            float dist2 = q.x * q.x + q.z * q.z;
            if (tag == 2) { q.y = p.ymax; return f3(0.0f, 1.0f, 0.0f); }
            if (tag == 1) { q.y = p.ymin; return f3(0.0f, -1.0f, 0.0f); }
            float inv = rsqrtf(dist2);
            q.x *= inv;
            q.z *= inv;
            return f3(q.x, 0.0f, q.z);
        }
        case PK_RTC_CONE: {  // cone.rs:65-84
            float dist2 = q.x * q.x + q.z * q.z;
            // cap hits with dist2 >= y^2 fall through to the wall normal in the reference (cone.rs:69-79)
            if (tag == 2 && dist2 < p.ymax * p.ymax) { q.y = p.ymax; return f3(0.0f, 1.0f, 0.0f); }
            if (tag == 1 && dist2 < p.ymin * p.ymin) { q.y = p.ymin; return f3(0.0f, -1.0f, 0.0f); }
            float y = sqrtf(dist2);
            y = q.y > 0.0f ? -y : y;
            return normalize_precise(f3(q.x, y, q.z));
        }
    }
    return f3(0.0f, 1.0f, 0.0f);
}

// ---- patterns (pattern/*.rs) ---------------------------------------------------------------------------
__device__ __forceinline__ float3 pattern_at(const DevTexture& t, float3 p) {
    float3 a = f3(t.a), b = f3(t.b);
    int kind = __float_as_int(t.a.w);
    switch (kind) {
        case RL_TEX_RTC_STRIPE:  // stripe.rs:21-27
            return (((long long)floorf(p.x)) % 2 == 0) ? a : b;
        case RL_TEX_RTC_CHECKER3D:  // checker3d.rs:19-25
            return (((long long)(floorf(p.x) + floorf(p.y) + floorf(p.z))) % 2 == 0) ? a : b;
        case RL_TEX_RTC_GRADIENT: {  // gradient.rs:20-25
            float f = p.x - floorf(p.x);
            return fma3(b - a, f, a);
        }
        case RL_TEX_RTC_RING: {  // ring.rs:20-28
            float r = sqrtf(p.x * p.x + p.z * p.z);
            return (((long long)floorf(r)) % 2 == 0) ? a : b;
        }
    }
    return a;
}

// ---- hit record --------------------------------------------------------------------------------------
struct RtcHit {
    float t;
    int prim;   // >= 0 analytic prim index, < 0: ~triangle index, INT_MIN none
    int node;   // source node id (leaf order, breaks exact ties)
    float b1, b2;
    int tag;    // root tag of an analytic primitive (see prim_roots)
};
constexpr int NO_HIT = 0x7fffffff;

// CsgOperation::intersection_allowed (csg.rs:16-28)
__device__ __forceinline__ bool csg_allowed(int op, bool left, bool in_left, bool in_right) {
    if (op == RL_CSG_UNION) return left ? !in_right : !in_left;
    if (op == RL_CSG_INTERSECTION) return left ? in_right : in_left;
    return left ? !in_right : in_left;  // Difference
}
constexpr int CSG_CAP = 24;  // crossings one top-level Csg may produce along a ray (overflow is reported)

template <bool COUNT, bool CSG>
struct RtcTracer {
    const DevScene& sc;
    LocalCount<COUNT>& lc;

    __device__ __forceinline__ RtcTracer(const DevScene& s, LocalCount<COUNT>& l) : sc(s), lc(l) {}

    // Calls f(i, prim, ts, n, tags) for every analytic primitive with the roots that reach the World's list, i.e.
    // after every Csg above the leaf has filtered them (csg.rs:49-111); f returns true to stop.
    //
    // A top-level Csg owns a contiguous prim range.  Its leaves' crossings are gathered into a small per-thread
    // list, stably sorted by t (gather order = leaf DFS order = the reference's "left list, then right list,
    // sort_by t"), and the Csg nodes are applied bottom-up: each walks the surviving crossings of its range,
    // toggling in_left / in_right exactly like filter_intersections (csg.rs:51-76).
    template <class F>
    __device__ __forceinline__ void each_prim(float3 o, float3 d, F f) {
        int i = 0;
        while (i < sc.n_prims) {
            const RtcPrim& p = sc.prims[i];
            if (!CSG || p.csg_count == 0) {
                float3 lo = xf_point(p.inv, o), ld = xf_vec(p.inv, d);
                float ts[4];
                unsigned tags;
                int n = prim_roots(p, lo, ld, ts, &tags);
                if (COUNT) lc.prims++;
                if (f(i, p, ts, n, tags)) return;
                i++;
                continue;
            }
            RL_CHECK_OR(p.csg_first >= 0 && p.csg_first + p.csg_count <= sc.n_csg, lc, return);
            const int4 top = sc.csg[p.csg_first + p.csg_count - 1];
            RL_CHECK_OR(top.y >= 0 && top.w <= sc.n_prims && top.y <= top.w, lc, return);
            float ct[CSG_CAP];
            int ci[CSG_CAP];  // prim << 8 | tag
            int cn = 0;
            for (int j = top.y; j < top.w; j++) {
                const RtcPrim& q = sc.prims[j];
                float ts[4];
                unsigned tags = 0u;
                int n;
                if (q.kind == PK_TRIANGLE) {
                    float b1, b2;
                    n = tri_prim_roots(q, o, d, ts, &b1, &b2);
                    if (COUNT) lc.tris++;
                } else {
                    float3 lo = xf_point(q.inv, o), ld = xf_vec(q.inv, d);
                    n = prim_roots(q, lo, ld, ts, &tags);
                    if (COUNT) lc.prims++;
                }
                for (int k = 0; k < n; k++) {
                    if (cn >= CSG_CAP) { lc.overflow++; break; }
                    // stable insertion by t
                    int e = cn++;
                    while (e > 0 && ct[e - 1] > ts[k]) { ct[e] = ct[e - 1]; ci[e] = ci[e - 1]; e--; }
                    ct[e] = ts[k];
                    ci[e] = (j << 8) | (int)((tags >> (2 * k)) & 3u);
                }
            }
            unsigned alive = cn >= 32 ? 0xffffffffu : ((1u << cn) - 1u);
            for (int c = 0; c < p.csg_count; c++) {
                const int4 nd = sc.csg[p.csg_first + c];
                bool in_left = false, in_right = false;
                for (int e = 0; e < cn; e++) {
                    int j = ci[e] >> 8;
                    if (!((alive >> e) & 1u) || j < nd.y || j >= nd.w) continue;
                    bool left = j < nd.z;
                    if (!csg_allowed(nd.x, left, in_left, in_right)) alive &= ~(1u << e);
                    if (left) in_left = !in_left; else in_right = !in_right;
                }
            }
            for (int j = top.y; j < top.w; j++) {
                float ts[4];
                unsigned tags = 0u;
                int n = 0;
                for (int e = 0; e < cn; e++)
                    if (((alive >> e) & 1u) && (ci[e] >> 8) == j && n < 4) {
                        tags |= (unsigned)(ci[e] & 3) << (2 * n);
                        ts[n++] = ct[e];
                    }
                if (f(j, sc.prims[j], ts, n, tags)) return;
            }
            i = top.w;
        }
    }

    // closest hit per intersect::hit over World::intersect
    RL_TRACER_FN RtcHit closest(float3 o, float3 d) {
        if (COUNT) lc.rays++;
        RtcHit h;
        h.t = RL_INF;
        h.prim = NO_HIT;
        h.node = -1;
        h.b1 = h.b2 = 0.0f;
        h.tag = 0;
        each_prim(o, d, [&](int i, const RtcPrim& p, const float* ts, int n, unsigned tags) -> bool {
            for (int k = 0; k < n; k++) {
                float t = ts[k];
                if (t >= 0.0f && (t < h.t || (t == h.t && p.node >= h.node))) {
                    h.t = t;
                    h.prim = i;
                    h.node = p.node;
                    h.tag = (int)((tags >> (2 * k)) & 3u);
                }
            }
            return false;
        });
        if (sc.n_bvh_prims > 0) {
            RayPre pre = make_pre(o, d);
            const TriVerts* tv = sc.tri_verts;
            LocalCount<COUNT>& lcr = lc;
            RtcHit* hp = &h;
            bvh_traverse<COUNT>(sc.nodes, sc.n_bvh_prims, pre, 0.0f, h.t, lc, [&](int ref, float tmax) -> float {
                int ti = ref_index(ref);
                RL_CHECK_OR(ref_type(ref) == REF_TRI && ti < sc.n_tris, lcr, return tmax);
                float4 p0 = tv[ti].p0, p1 = tv[ti].p1, p2 = tv[ti].p2;
                if (COUNT) lcr.tris++;
                float t, b1, b2;
                if (tri_hit(pre, f3(p0), f3(p1), f3(p2), &t, &b1, &b2)) {
                    int node = __float_as_int(p1.w);
                    // ties: later leaf wins; the triangles of one device-ingested mesh share a node, their order is ti
                    if (t >= 0.0f && (t < hp->t || (t == hp->t && (node > hp->node || (node == hp->node && (hp->prim >= 0 || ti >= ~hp->prim)))))) {
                        hp->t = t;
                        hp->prim = ~ti;
                        hp->node = node;
                        hp->b1 = b1;
                        hp->b2 = b2;
                        return t;
                    }
                }
                return tmax;
            });
        }
        return h;
    }

    // is crossing (t, node, root) strictly before the hit crossing in the stable-sorted order?
    __device__ __forceinline__ static bool before(float t, int node, int root, float th, int nh, int rh) {
        return t < th || (t == th && (node < nh || (node == nh && root < rh)));
    }
    __device__ __forceinline__ static bool later(float t, int node, float tb, int nb) {
        return t > tb || (t == tb && node > nb);
    }

    // n1 / n2 of prepare_computations (intersect.rs:72-99) without building the list
    RL_TRACER_FN void refractive_indices(float3 o, float3 d, const RtcHit& h, float* n1, float* n2) {
        float best_t = -RL_INF;
        int best_node = -1;
        float best_ior = 1.0f;
        bool any_other = false;
        bool hit_open = false;
        float hit_last_t = -RL_INF;
        int hit_root = 0;
        float hit_ior = 1.0f;
        each_prim(o, d, [&](int i, const RtcPrim& p, const float* ts, int n, unsigned) -> bool {
            float th = h.t;
            if (i == h.prim) {
                // which root is the hit?  the one nearest to the recorded t (robust to re-evaluation)
                float bestd = RL_INF;
                for (int k = 0; k < n; k++) {
                    float dd = fabsf(ts[k] - h.t);
                    if (dd < bestd) { bestd = dd; hit_root = k; }
                }
                if (n > 0) th = ts[hit_root];
            }
            int cnt = 0;
            float last = -RL_INF;
            for (int k = 0; k < n; k++) {
                bool bf = (i == h.prim) ? (k != hit_root && before(ts[k], p.node, k, th, h.node, hit_root))
                                        : before(ts[k], p.node, 0, h.t, h.node, 0);
                if (bf) {
                    cnt++;
                    last = fmaxf(last, ts[k]);
                }
            }
            float ior = sc.materials[p.material].b.z;
            if (i == h.prim) {
                hit_open = cnt & 1;
                hit_last_t = last;
                hit_ior = ior;
            } else if (cnt & 1) {
                if (!any_other || later(last, p.node, best_t, best_node)) {
                    best_t = last;
                    best_node = p.node;
                    best_ior = ior;
                    any_other = true;
                }
            }
            return false;
        });
        if (h.prim < 0) hit_ior = sc.materials[__float_as_int(sc.tri_verts[~h.prim].p0.w)].b.z;
        if (sc.n_bvh_prims > 0) {
            // every triangle is its own object with a single crossing (triangle.rs:63-101)
            RayPre pre = make_pre(o, d);
            const TriVerts* tv = sc.tri_verts;
            const DevMaterial* mats = sc.materials;
            LocalCount<COUNT>& lcr = lc;
            int hit_tri = h.prim < 0 ? ~h.prim : -1;
            float th = h.t;
            int nh = h.node;
            bvh_traverse<COUNT>(sc.nodes, sc.n_bvh_prims, pre, -RL_INF, h.t, lc, [&](int ref, float tmax) -> float {
                int ti = ref_index(ref);
                if (ti == hit_tri) return tmax;
                float4 p0 = tv[ti].p0, p1 = tv[ti].p1, p2 = tv[ti].p2;
                if (COUNT) lcr.tris++;
                float t, b1, b2;
                if (tri_hit(pre, f3(p0), f3(p1), f3(p2), &t, &b1, &b2)) {
                    int node = __float_as_int(p1.w);
                    if (before(t, node, 0, th, nh, 0) && (!any_other || later(t, node, best_t, best_node))) {
                        best_t = t;
                        best_node = node;
                        best_ior = mats[__float_as_int(p0.w)].b.z;
                        any_other = true;
                    }
                }
                return tmax;
            });
        }
        float other = any_other ? best_ior : 1.0f;
        if (hit_open) {
            bool hit_is_last = !any_other || later(hit_last_t, h.node, best_t, best_node);
            *n1 = hit_is_last ? hit_ior : other;
            *n2 = other;  // the hit object is removed from the containers
        } else {
            *n1 = other;
            *n2 = hit_ior;  // the hit object is pushed last
        }
    }

    // World::shadow_attenuation (world.rs:104-126)
    RL_TRACER_FN float shadow(float3 point, float3 light_pos) {
        float3 v = light_pos - point;
        float dist2 = dot(v, v);
        if (dist2 == 0.0f) return 1.0f;
        float distance = sqrtf(dist2);
        float3 d = v * (1.0f / distance);
        if (COUNT) lc.rays++;
        if (!sc.has_transparency) {
            // every material is opaque: the first counted crossing already makes the product 0
            bool blocked = false;
            each_prim(point, d, [&](int, const RtcPrim&, const float* ts, int n, unsigned) -> bool {
                for (int k = 0; k < n; k++)
                    if (ts[k] > 0.0f && ts[k] < distance) blocked = true;
                return blocked;
            });
            if (blocked) return 0.0f;
            if (sc.n_bvh_prims > 0) {
                RayPre pre = make_pre(point, d);
                const TriVerts* tv = sc.tri_verts;
                LocalCount<COUNT>& lcr = lc;
                bvh_traverse<COUNT>(sc.nodes, sc.n_bvh_prims, pre, 0.0f, distance, lc, [&](int ref, float tmax) -> float {
                    int ti = ref_index(ref);
                    float4 p0 = tv[ti].p0, p1 = tv[ti].p1, p2 = tv[ti].p2;
                    if (COUNT) lcr.tris++;
                    float t, b1, b2;
                    if (tri_hit(pre, f3(p0), f3(p1), f3(p2), &t, &b1, &b2) && t > 0.0f && t < distance) {
                        blocked = true;
                        return -RL_INF;
                    }
                    return tmax;
                });
            }
            return blocked ? 0.0f : 1.0f;
        }
        // general case.  pass 1: the sorted walk stops at the first crossing of an object already seen,
        // i.e. at the earliest SECOND in-range crossing of any analytic primitive (triangles cross once)
        float stop_t = RL_INF;
        int stop_node = 0x7fffffff;
        each_prim(point, d, [&](int, const RtcPrim& p, const float* ts, int n, unsigned) -> bool {
            float first = RL_INF, second = RL_INF;
            for (int k = 0; k < n; k++) {
                float t = ts[k];
                if (t > 0.0f && t < distance) {
                    if (t < first) { second = first; first = t; }
                    else if (t < second) second = t;
                }
            }
            if (second < stop_t || (second == stop_t && second < RL_INF && p.node < stop_node)) {
                stop_t = second;
                stop_node = p.node;
            }
            return false;
        });
        // pass 2: product of transparency over first crossings that come before the stop
        float prod = 1.0f;
        each_prim(point, d, [&](int, const RtcPrim& p, const float* ts, int n, unsigned) -> bool {
            float first = RL_INF;
            for (int k = 0; k < n; k++)
                if (ts[k] > 0.0f && ts[k] < distance) first = fminf(first, ts[k]);
            if (first < RL_INF && (first < stop_t || (first == stop_t && p.node <= stop_node)))
                prod *= sc.materials[p.material].b.y;
            return false;
        });
        if (prod != 0.0f && sc.n_bvh_prims > 0) {
            RayPre pre = make_pre(point, d);
            const TriVerts* tv = sc.tri_verts;
            const DevMaterial* mats = sc.materials;
            LocalCount<COUNT>& lcr = lc;
            float lim = fminf(distance, stop_t);
            bvh_traverse<COUNT>(sc.nodes, sc.n_bvh_prims, pre, 0.0f, lim, lc, [&](int ref, float tmax) -> float {
                int ti = ref_index(ref);
                float4 p0 = tv[ti].p0, p1 = tv[ti].p1, p2 = tv[ti].p2;
                if (COUNT) lcr.tris++;
                float t, b1, b2;
                if (tri_hit(pre, f3(p0), f3(p1), f3(p2), &t, &b1, &b2) && t > 0.0f && t < distance &&
                    (t < stop_t || (t == stop_t && __float_as_int(p1.w) < stop_node))) {
                    prod *= mats[__float_as_int(p0.w)].b.y;
                    if (prod == 0.0f) return -RL_INF;
                }
                return tmax;
            });
        }
        return prod;
    }

    // material::lighting (material.rs:54-90)
    __device__ __forceinline__ float3 lighting(const DevMaterial& m, float3 point, float3 obj_color, const DevLight& l,
                                               float3 eyev, float3 normalv, float atten) {
        float3 li = f3(l.intensity);
        float3 effective = obj_color * li;
        float3 lightv = normalize_precise(f3(l.pos) - point);
        float3 col = effective * m.a.x;  // ambient is never shadowed
        float ldn = dot(lightv, normalv);
        if (!(ldn < 0.0f)) {
            col = fma3(effective, m.a.y * ldn * atten, col);
            float3 reflectv = -(lightv - normalv * (2.0f * ldn));
            float rde = dot(reflectv, eyev);
            if (rde > 0.0f) {
                float factor = powf(rde, m.a.w);
                col = fma3(li, m.a.z * factor * atten, col);
            }
        }
        return col;
    }

    // World::color_at (world.rs:89-102) with the recursion unrolled onto an explicit stack
    __device__ float3 color_at(float3 o0, float3 d0) {
        float3 so[RTC_STACK], sd[RTC_STACK];
        float sw[RTC_STACK];
        int sr[RTC_STACK];
        int sp = 0;
        so[0] = o0; sd[0] = d0; sw[0] = 1.0f; sr[0] = sc.max_reflection_depth;
        sp = 1;
        float3 result = f3(0.0f, 0.0f, 0.0f);
        const float3 void_c = f3(sc.void_color[0], sc.void_color[1], sc.void_color[2]);
        while (sp > 0) {
            sp--;
            float3 o = so[sp], d = sd[sp];
            float w = sw[sp];
            int remaining = sr[sp];
            RtcHit h = closest(o, d);
            if (h.prim == NO_HIT || sc.n_lights == 0) {  // no lights: shade_hit returns None (world.rs:57-87)
                result = fma3(void_c, w, result);
                continue;
            }
            if (COUNT) lc.shades++;
            // ---- prepare_computations (intersect.rs:47-70) ----
            float3 point, normal, pat_p;
            int mat_id;
            if (CSG && h.prim >= 0 && sc.prims[h.prim].kind == PK_TRIANGLE) {
                const RtcPrim& p = sc.prims[h.prim];
                float ts[4], b1 = 0.0f, b2 = 0.0f;
                tri_prim_roots(p, o, d, ts, &b1, &b2);  // the barycentrics of the hit (the crossing list keeps t only)
                const float b0 = 1.0f - b1 - b2;
                point = f3(p.inv[0]) * b0 + f3(p.inv[1]) * b1 + f3(p.inv[2]) * b2;
                normal = (p.flags & 1) ? normalize_precise(f3(p.fwd[1]) * b1 + f3(p.fwd[2]) * b2 + f3(p.fwd[0]) * b0) : f3(p.fwd[0]);
                mat_id = p.material;
                pat_p = xf_point(p.pat, point);
            } else if (h.prim >= 0) {
                const RtcPrim& p = sc.prims[h.prim];
                float3 lo = xf_point(p.inv, o), ld = xf_vec(p.inv, d);
                float3 q = fma3(ld, h.t, lo);
                float3 nl = prim_normal(p, q, h.tag);
                normal = normalize_precise(xf_normal(p.inv, nl));
                point = xf_point(p.fwd, q);
                mat_id = p.material;
                pat_p = xf_point(p.pat, q);  // patterns live in the leaf's object space (object/mod.rs:20-32)
            } else {
                int ti = ~h.prim;
                float4 p0 = sc.tri_verts[ti].p0, p1 = sc.tri_verts[ti].p1, p2 = sc.tri_verts[ti].p2;
                float b0 = 1.0f - h.b1 - h.b2;
                // hit point from the barycentrics: error ~ ulp(|p|), independent of t
                point = f3(p0) * b0 + f3(p1) * h.b1 + f3(p2) * h.b2;
                int flags = __float_as_int(p2.w);
                float4 s0 = sc.tri_shade[ti].s0;
                if (flags & 1) {  // smooth: n2*u + n3*v + n1*(1-u-v) (triangle.rs:87-92)
                    float4 s1 = sc.tri_shade[ti].s1, s2 = sc.tri_shade[ti].s2;
                    normal = normalize_precise(f3(s1) * h.b1 + f3(s2) * h.b2 + f3(s0) * b0);
                } else {
                    normal = f3(s0);
                }
                mat_id = __float_as_int(p0.w);
                int xf = flags >> 8;
                pat_p = xf > 0 ? xf_point(sc.xforms[xf - 1].r, point) : point;
            }
            RL_CHECK_OR(mat_id >= 0 && mat_id < sc.n_materials, lc, continue);
            const DevMaterial m = sc.materials[mat_id];
            int tex = __float_as_int(m.color.w);
            RL_CHECK_OR(tex < sc.n_textures, lc, continue);
            float3 obj_color = tex >= 0 ? pattern_at(sc.textures[tex], pat_p) : f3(m.color);
            float3 eyev = normalize_precise(-d);
            if (dot(normal, eyev) < 0.0f) normal = -normal;
            float bias = fmaxf(RTC_BIAS, 2e-6f * max_abs(point));
            float3 over_point = fma3(normal, bias, point);
            // ---- shade_hit (world.rs:57-87) ----
            float3 surface = f3(0.0f, 0.0f, 0.0f);
            for (int l = 0; l < sc.n_lights; l++) {
                const DevLight lt = sc.lights[l];
                float atten = shadow(over_point, f3(lt.pos));
                surface = surface + lighting(m, point, obj_color, lt, eyev, normal, atten);
            }
            result = fma3(surface, w, result);
            float refl = m.b.x, transp = m.b.y;
            if (remaining > 0 && (refl > 0.0f || transp > 0.0f)) {
                float wl = w * (float)sc.n_lights;  // the reference re-traces both rays once per light
                float n1 = 1.0f, n2 = 1.0f;
                if (transp > 0.0f) refractive_indices(o, d, h, &n1, &n2);
                float cos_i = dot(eyev, normal);
                float wr = wl * refl, wt = wl * transp;
                if (refl > 0.0f && transp > 0.0f) {  // Schlick (intersect.rs:140-156)
                    float n = n1 / n2;
                    float sin2_t = n * n * (1.0f - cos_i * cos_i);
                    float reflectance;
                    if (sin2_t > 1.0f && n > 1.0f) {
                        reflectance = 1.0f;
                    } else {
                        float c = n > 1.0f ? sqrtf(1.0f - sin2_t) : cos_i;
                        float r0 = (n1 - n2) / (n1 + n2);
                        r0 *= r0;
                        float x = 1.0f - c;
                        float x2 = x * x;
                        reflectance = r0 + (1.0f - r0) * (x2 * x2 * x);
                    }
                    wr *= reflectance;
                    wt *= 1.0f - reflectance;
                }
                if (transp > 0.0f) {  // refracted_color (world.rs:138-159)
                    float n_ratio = n1 / n2;
                    float sin2_t = n_ratio * n_ratio * (1.0f - cos_i * cos_i);
                    if (!(sin2_t > 1.0f)) {
                        float cos_t = sqrtf(1.0f - sin2_t);
                        float3 dir = normal * (n_ratio * cos_i - cos_t) - eyev * n_ratio;
                        if (sp < RTC_STACK) {
                            so[sp] = fma3(normal, -bias, point);
                            sd[sp] = dir;
                            sw[sp] = wt;
                            sr[sp] = remaining - 1;
                            sp++;
                        } else {
                            lc.overflow++;
                        }
                    }
                }
                if (refl > 0.0f) {  // reflected_color (world.rs:128-136)
                    float3 rv = normalize_precise(d - normal * (2.0f * dot(d, normal)));
                    if (sp < RTC_STACK) {
                        so[sp] = over_point;
                        sd[sp] = rv;
                        sw[sp] = wr;
                        sr[sp] = remaining - 1;
                        sp++;
                    } else {
                        lc.overflow++;
                    }
                }
            }
        }
        return result;
    }
};

// rays_for_pixel (camera.rs:63-91)
__device__ __forceinline__ void camera_ray(const RtcCam& c, int px, int py, int nx, int ny, float3* o, float3* d) {
    float so = 1.0f / (float)c.aa;
    float xoff = ((float)px + so * ((float)nx + 0.5f)) * c.pixel_size;
    float yoff = ((float)py + so * ((float)ny + 0.5f)) * c.pixel_size;
    float wx = c.half_width - xoff, wy = c.half_height - yoff;
    float3 pixel = xf_point(c.inv, f3(wx, wy, -1.0f));
    float3 origin = f3(c.inv[0].w, c.inv[1].w, c.inv[2].w);
    *o = origin;
    *d = normalize_precise(pixel - origin);
}

// one thread per pixel; warps walk 8x4 pixel micro-tiles of each job rectangle
// Resident CTAs per SM the render kernel is compiled for.  Unconstrained, ptxas takes 96 registers (5 CTAs = 20 warps per SM)
// and ncu shows the kernel waiting on instruction fetch (`no_instruction` 3.5 warps per issue on the mirror scene: 7 000
// SASS instructions of branchy code) with too few warps to hide it.  Measured at 4K (gpurun_out/abrtc1, abrtc2):
//   CTAs/SM (registers)   1 (96)    6 (80)    8 (64)    10 (48)   12 (40)
//   C2 mirror scene       6.40 ms   5.45      4.97      5.33      5.23
//   C3 teapot             0.676     0.628     0.603     0.618     0.635
// Out-of-line prim_roots / closest / shadow (smaller code) was slower: 6.65 - 7.28 ms on C2.
#ifndef RL_RTC_MINB
#define RL_RTC_MINB 8
#endif
#ifndef RL_RTC_MINB_CSG
#define RL_RTC_MINB_CSG 8
#endif
template <bool COUNT, bool CSG>
__global__ void __launch_bounds__(128, CSG ? RL_RTC_MINB_CSG : RL_RTC_MINB)
    k_rtc_render(DevScene sc, RtcCam cam, JobTable jt, float* __restrict__ out, Counters* counters) {
    LocalCount<COUNT> lc;
    long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (item < jt.n_items) {
        int j = find_job(jt, item);
        rl_job job = jt_job(jt, j);
        int x, y;
        tile_pixel(job.x1 - job.x0, job.y1 - job.y0, item - jt_prefix(jt, j), &x, &y);
        x += job.x0;
        y += job.y0;
        if (x < job.x1 && y < job.y1) {
            RtcTracer<COUNT, CSG> tr(sc, lc);
            float3 acc = f3(0.0f, 0.0f, 0.0f);
            for (int nx = 0; nx < cam.aa; nx++)
                for (int ny = 0; ny < cam.aa; ny++) {
                    float3 o, d;
                    camera_ray(cam, x, y, nx, ny, &o, &d);
                    acc = acc + tr.color_at(o, d);
                }
            float s = 1.0f / (float)(cam.aa * cam.aa);
            size_t idx = ((size_t)y * cam.hsize + x) * 3;
            out[idx + 0] = acc.x * s;
            out[idx + 1] = acc.y * s;
            out[idx + 2] = acc.z * s;
        }
    }
    lc.flush(counters);
}

template <bool COUNT, bool CSG>
__global__ void __launch_bounds__(128) k_rtc_trace(DevScene sc, const rl_ray* __restrict__ rays, unsigned long long n,
                                                   rl_hit* __restrict__ hits, Counters* counters) {
    LocalCount<COUNT> lc;
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        rl_ray r = rays[i];
        RtcTracer<COUNT, CSG> tr(sc, lc);
        RtcHit h = tr.closest(f3(r.origin[0], r.origin[1], r.origin[2]), f3(r.direction[0], r.direction[1], r.direction[2]));
        rl_hit o;
        o.node = (h.prim == NO_HIT || h.node < 0 || h.node >= sc.n_ord) ? -1 : sc.ord_node[h.node];  // DFS ordinal -> caller's id
        o.t = h.t;
        o.u = h.b1;
        o.v = h.b2;
        hits[i] = o;
    }
    lc.flush(counters);
}

}  // namespace

cudaError_t launch_rtc_render(const DevScene& sc, const rl_rtc_camera* cam, const double inv[12], uint32_t aa,
                              const JobTable& jt, float* d_out, Counters* d_counters, bool instrumented,
                              cudaStream_t stream) {
    RtcCam c;
    c.hsize = cam->hsize;
    c.vsize = cam->vsize;
    // Camera::new (camera.rs:35-57), in f64 on the host
    double half_view = tan(cam->fov / 2.0);
    double aspect = (double)cam->hsize / (double)cam->vsize;
    double hw, hh;
    if (aspect >= 1.0) { hw = half_view; hh = half_view / aspect; }
    else { hw = half_view * aspect; hh = half_view; }
    c.half_width = (float)hw;
    c.half_height = (float)hh;
    c.pixel_size = (float)(hw * 2.0 / (double)cam->hsize);
    for (int i = 0; i < 3; i++)
        c.inv[i] = make_float4((float)inv[i * 4 + 0], (float)inv[i * 4 + 1], (float)inv[i * 4 + 2], (float)inv[i * 4 + 3]);
    c.aa = (int)aa;
    if (jt.n_items <= 0) return cudaSuccess;
    unsigned blocks = (unsigned)((jt.n_items + 127) / 128);
    // scenes without a Csg run the instantiation that carries no crossing lists (fewer registers, no local memory)
    if (sc.n_csg > 0) {
        if (instrumented) k_rtc_render<true, true><<<blocks, 128, 0, stream>>>(sc, c, jt, d_out, d_counters);
        else k_rtc_render<false, true><<<blocks, 128, 0, stream>>>(sc, c, jt, d_out, d_counters);
    } else {
        if (instrumented) k_rtc_render<true, false><<<blocks, 128, 0, stream>>>(sc, c, jt, d_out, d_counters);
        else k_rtc_render<false, false><<<blocks, 128, 0, stream>>>(sc, c, jt, d_out, d_counters);
    }
    return cudaGetLastError();
}

cudaError_t launch_rtc_trace(const DevScene& sc, const rl_ray* d_rays, uint64_t n, rl_hit* d_hits,
                             Counters* d_counters, bool instrumented, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    unsigned blocks = (unsigned)((n + 127) / 128);
    if (sc.n_csg > 0) {
        if (instrumented) k_rtc_trace<true, true><<<blocks, 128, 0, stream>>>(sc, d_rays, n, d_hits, d_counters);
        else k_rtc_trace<false, true><<<blocks, 128, 0, stream>>>(sc, d_rays, n, d_hits, d_counters);
    } else {
        if (instrumented) k_rtc_trace<true, false><<<blocks, 128, 0, stream>>>(sc, d_rays, n, d_hits, d_counters);
        else k_rtc_trace<false, false><<<blocks, 128, 0, stream>>>(sc, d_rays, n, d_hits, d_counters);
    }
    return cudaGetLastError();
}

}  // namespace rl
