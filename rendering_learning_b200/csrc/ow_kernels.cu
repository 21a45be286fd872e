// OW (ray-tracing-one-weekend) render path: stochastic path tracer in f32.
//
// Replaces Camera::_render -> get_ray -> ray_color -> world.hit / Material::scatter / emitted
// (OW/src/camera.rs:145-260, bvh.rs:81-90, hittable/*.rs, material.rs:69-195, texture.rs:15-82).
//
// Execution model: persistent warps.  A work item is (pixel, sample chunk); every LANE owns one item at a
// time, walks its samples in order (so the per-item sum is deterministic) and regenerates a new camera
// path as soon as its current path dies, so lanes never idle behind a long path in the same warp.  Lanes
// that run out of samples refill from a global queue with one warp-aggregated atomic (ballot + popc
// compaction of the requesting lanes).  Partial sums are written per (chunk, pixel) with one 16-byte store and
// folded in chunk order by k_ow_reduce, so the image is bit-identical for any GPU count / schedule.
// k_ow_render5 is the production kernel (resumable per-lane traversal, ballot-scheduled node steps / leaf rounds /
// service rounds, instantiated per primitive mix); k_ow_render (v3) is kept as the A/B baseline of DESIGN.md §4.
//
// RNG: Philox4x32-10 keyed by the camera seed, counter = (pixel, absolute sample, bounce, dimension) —
// statistical parity with the reference's ChaCha8 streams, and absolute sample indices
// keep render_from_checkpoint streams disjoint like camera.rs:162-170.
#include "device.cuh"
#include "kernels.h"

namespace rl {
namespace {

struct OwCam {
    int width, height, spp, max_depth;
    int first_sample, n_chunks, defocus, pad;
    float3 lookfrom, pixel00, du, dv, disk_u, disk_v, background;
    unsigned seed_lo, seed_hi;
};

// ---- Philox4x32-10 ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox(uint4 c, uint2 k) {
#pragma unroll
    for (int i = 0; i < 10; i++) {
        unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ float u01(unsigned x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// uniform point on the unit sphere (any unbiased sampler gives statistical parity with UnitSphere)
// MUFU-based square root and sine / cosine for SAMPLING (random directions, lens points) and for the sphere test's root:
// 1-2 ulp instead of correctly rounded, ~2 instructions instead of ~10 / ~25 (sincospif was 2.6 % of the cover scene's
// warp instructions at 10-16 lanes).  The angle is kept in [-pi, pi), where sin.approx / cos.approx are accurate to 2^-21.
__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void fast_sincos_turn(float u, float* s, float* c) {  // angle = 2 pi u - pi, u in [0, 1)
    __sincosf(fmaf(u, 6.28318530717959f, -3.14159265358979f), s, c);
}
__device__ __forceinline__ float3 unit_vector(float u1, float u2) {
    float z = 1.0f - 2.0f * u1;
    float r = fast_sqrt(fmaxf(0.0f, 1.0f - z * z));
    float s, c;
    fast_sincos_turn(u2, &s, &c);
    return f3(r * c, r * s, z);
}

// PRIMS: what the scene holds (bit 0 spheres, bit 1 triangles, bit 2 quads, bit 3 constant media, bit 4 Noise
// textures).  The render kernel is instantiated for "spheres only", "no spheres", "every surface" and "everything", so
// a sphere scene carries no triangle / quad / image / medium / Perlin code (smaller code, fewer registers, fewer
// instruction-cache misses: ncu showed 8 % of issue slots lost to `no_instruction` on the one-size-fits-all build).
constexpr int PRIMS_SPHERES = 1, PRIMS_TRIS = 2, PRIMS_QUADS = 4, PRIMS_MEDIA = 8, PRIMS_NOISE = 16;
constexpr int PRIMS_ALL = 7, PRIMS_FULL = 31;

// ---- Perlin noise (perlin.rs:39-115) ---------------------------------------------------------------------
__device__ __forceinline__ float perlin_noise(const float4* __restrict__ vec, const int* __restrict__ perm, float3 p) {
    float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
    float u = p.x - fx, v = p.y - fy, w = p.z - fz;
    int i = (int)fx, j = (int)fy, k = (int)fz;
    float uu = u * u * (3.0f - 2.0f * u), vv = v * v * (3.0f - 2.0f * v), ww = w * w * (3.0f - 2.0f * w);
    float accum = 0.0f;
#pragma unroll
    for (int di = 0; di < 2; di++) {
        int px = __ldg(perm + ((i + di) & 255));
        float wi = di ? uu : 1.0f - uu;
#pragma unroll
        for (int dj = 0; dj < 2; dj++) {
            int py = __ldg(perm + 256 + ((j + dj) & 255));
            float wj = dj ? vv : 1.0f - vv;
#pragma unroll
            for (int dk = 0; dk < 2; dk++) {
                int pz = __ldg(perm + 512 + ((k + dk) & 255));
                float wk = dk ? ww : 1.0f - ww;
                float4 c = __ldg(vec + (px ^ py ^ pz));
                accum += wi * wj * wk * dot(f3(c), f3(u - (float)di, v - (float)dj, w - (float)dk));
            }
        }
    }
    return accum;
}
__device__ __forceinline__ float perlin_turb(const float4* vec, const int* perm, float3 p, int depth) {
    float accum = 0.0f, weight = 1.0f;
    for (int i = 0; i < depth; i++) {
        accum = fmaf(weight, perlin_noise(vec, perm, p), accum);
        weight *= 0.5f;
        p = p * 2.0f;
    }
    return fabsf(accum);
}

// ---- textures (texture.rs) -------------------------------------------------------------------------------
template <int PRIMS = PRIMS_ALL>
__device__ __forceinline__ float3 tex_value(const DevScene& sc, int tex, float u, float v, float3 p) {
    DevTexture t = sc.textures[tex];
    for (int guard = 0; guard < 8 && __float_as_int(t.a.w) == RL_TEX_OW_CHECKER; guard++) {
        float inv_scale = t.b.w;  // texture.rs:42-54
        // `floor() as i64`, summed, `% 2 == 0`: only the parity matters, and the parity of a wrapping 32-bit sum is the sum's
        // (one F2I with round-down per axis instead of the 64-bit conversions; exact below 2^31 cells from the origin)
        int xi = __float2int_rd(p.x * inv_scale);
        int yi = __float2int_rd(p.y * inv_scale);
        int zi = __float2int_rd(p.z * inv_scale);
        bool even = (((unsigned)xi + (unsigned)yi + (unsigned)zi) & 1u) == 0u;
        t = sc.textures[even ? t.idx.x : t.idx.y];
    }
    if (__float_as_int(t.a.w) == RL_TEX_OW_IMAGE) {  // texture.rs:63-81
        DevImage im = sc.images[t.idx.z];
        float uu = fminf(fmaxf(u, 0.0f), 1.0f);
        float vv = 1.0f - fminf(fmaxf(v, 0.0f), 1.0f);
        int i = (int)(uu * (float)(im.width - 1));
        int j = (int)(vv * (float)(im.height - 1));
        float4 px = __ldg(im.texels + (size_t)j * im.width + i);
        return f3(px);
    }
    if ((PRIMS & PRIMS_NOISE) && __float_as_int(t.a.w) == RL_TEX_OW_NOISE) {  // texture.rs:89-93
        const float4* vec = sc.perlin_vec + 256 * t.idx.z;
        const int* perm = sc.perlin_perm + 768 * t.idx.z;
        float g = 0.5f * (1.0f + sinf(fmaf(t.b.w, p.z, 10.0f * perlin_turb(vec, perm, p, 7))));
        return f3(g, g, g);
    }
    return f3(t.a);
}

struct OwHit {
    float t;
    int ref;  // leaf ref, -1 none
    float b1, b2;
};

// One primitive test of world.hit: updates `h` when the primitive is hit closer than h.t.
// `self_ref` is the primitive the ray starts on (never re-hit at t ~ 0; a sphere only at its far root).

__device__ __forceinline__ unsigned hash32(unsigned x) {  // lowbias32 (Wellons): decorrelates the per-medium draws
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <bool COUNT, int PRIMS>
__device__ __forceinline__ void ow_medium_test(const DevScene& sc, int ref, const RayPre& pre, float a_dd, float time,
                                               float tmin, unsigned ray_rnd, OwHit& h, LocalCount<COUNT>& lc);

constexpr float OW_TRI_EDGE_EPS = 2e-5f;
template <bool COUNT, int PRIMS = PRIMS_ALL>
__device__ __forceinline__ void ow_leaf_test(const DevScene& sc, int ref, const RayPre& pre, float a_dd, float time,
                                             int self_ref, float tmin, OwHit& h, LocalCount<COUNT>& lc, unsigned ray_rnd = 0u) {
    const float3 o = pre.o, d = pre.d;
    int type = ref_type(ref), idx = ref_index(ref);
    RL_CHECK_OR(ref >= 0 && idx < (type == REF_SPHERE ? sc.n_spheres : type == REF_TRI ? sc.n_tris : type == REF_QUAD ? sc.n_quads : sc.n_media),
                lc, return);
    if ((PRIMS & PRIMS_MEDIA) && type == REF_MEDIUM) {
        ow_medium_test<COUNT, PRIMS>(sc, ref, pre, a_dd, time, tmin, ray_rnd, h, lc);
        return;
    }
    if (PRIMS == PRIMS_SPHERES || ((PRIMS & PRIMS_SPHERES) && type == REF_SPHERE)) {  // sphere.rs:34-75
#if RL_LDG256
        float4 c, dc;
        ldg256(sc.spheres + idx, c, dc);
#else
        float4 c = sc.spheres[idx].c, dc = sc.spheres[idx].dc;
#endif
        if (COUNT) lc.prims++;
        float3 center = fma3(f3(dc), time, f3(c));
        float3 oc = o - center;
        float hb = dot(oc, d);
        float t;
        if (ref == self_ref) {
            // origin lies on this sphere: the roots are 0 and -2 hb / a; only the far one is a new hit
            t = -2.0f * hb / a_dd;
            if (!(t > 1e-4f * fabsf(c.w) * rsqrtf(a_dd))) return;
        } else if (fabsf(c.w) >= OW_BIG_RADIUS) {
            // a huge sphere seen from near its surface (the r = 1000 ground of the cover scene): r^2 - |perp|^2
            // cancels ~7 digits, more than f32 has.  One such primitive sits near the BVH root, so its
            // quadratic's cancelling terms are evaluated in f64 (the cost is one test per ray).
            // o - centre in f64 as well: in f32 the ray's height above an r = 1000 sphere keeps 6e-5 of absolute precision,
            // which a grazing ray turns into 4e-4 of relative error in t (measured on the reference's scattered rays)
            double ox = (double)o.x - fma((double)dc.x, (double)time, (double)c.x);
            double oy = (double)o.y - fma((double)dc.y, (double)time, (double)c.y);
            double oz = (double)o.z - fma((double)dc.z, (double)time, (double)c.z);
            double dx = (double)d.x, dy = (double)d.y, dz = (double)d.z;
            double a = dx * dx + dy * dy + dz * dz;
            double hbd = ox * dx + oy * dy + oz * dz;
            double cc = ox * ox + oy * oy + oz * oz - (double)c.w * (double)c.w;
            double disc = hbd * hbd - a * cc;
            if (disc < 0.0) return;
            // The cancellations (|oc|^2 - r^2, hbd^2 - a cc) are done, in f64; the roots follow in f32 through the form that
            // has none left: q = -(hb + sign(hb) sqrt(disc)), roots q / a and cc / q.  (The f64 square root and the two f64
            // divisions this replaces were 2.3 % of the cover scene's warp instructions.)
            const float sq = fast_sqrt((float)disc), hbf = (float)hbd;
            const float q = hbf < 0.0f ? sq - hbf : -(sq + hbf);
            const float ra = __fdividef(q, (float)a), rb = __fdividef((float)cc, q);
            t = fminf(ra, rb);
            if (!(t >= tmin && t <= h.t)) {
                t = fmaxf(ra, rb);
                if (!(t >= tmin && t <= h.t)) return;
            }
        } else {
            float tc = -hb / a_dd;
            float3 perp = fma3(d, tc, oc);
            float disc = a_dd * (c.w * c.w - dot(perp, perp));
            if (disc < 0.0f) return;
            float q = fast_sqrt(disc) / a_dd;
            t = tc - q;
            if (!(t >= tmin && t <= h.t)) {
                t = tc + q;
                if (!(t >= tmin && t <= h.t)) return;
            }
        }
        if (t < h.t) {
            h.t = t;
            h.ref = ref;
            return;
        }
        return;
    } else if ((PRIMS & PRIMS_TRIS) && type == REF_TRI) {  // flat/plane.rs:51-80 + flat/triangle.rs:60-67, the reference's own form
        // Round 1 ran the watertight ray-space test here (device.cuh tri_hit, still the RTC path): ~90 instructions with its
        // per-ray axis permutation selects, 20 % of the Cornell-box + spot warp instructions at 5 - 6 lanes
        // (profiles/r02_lines_ow_c5_4k_256_watertight_triangles.txt).  The plane form is the quad test with a different acceptance: ~35.
        if (ref == self_ref) return;
        const OwTriPlane& tp = sc.tri_plane[idx];
        float4 n4 = tp.n;
        if (COUNT) lc.tris++;
        float denom = dot(f3(n4), d);
        if (fabsf(denom) < 1e-8f) return;
        float t = __fdividef(n4.w - dot(f3(n4), o), denom);  // 2 ulp (|denom| >= 1e-8): far below the cancellation in the numerator
        if (!(t >= tmin && t < h.t)) return;
        float3 ip = fma3(d, t, o);
        float4 a4 = tp.a, b4 = tp.b;
        float alpha = dot(f3(a4), ip) - a4.w;
        float beta = dot(f3(b4), ip) - b4.w;
        // f32 evaluates alpha / beta to ~1e-5 for a unit-sized triangle a few hundred units from the origin; the acceptance is
        // widened by that much so that the triangles of a mesh overlap along shared edges instead of leaving cracks
        if (!(alpha >= -OW_TRI_EDGE_EPS && beta >= -OW_TRI_EDGE_EPS && alpha + beta <= 1.0f + OW_TRI_EDGE_EPS)) return;
        h.t = t;
        h.ref = ref;
        h.b1 = alpha;
        h.b2 = beta;
        return;
    } else if ((PRIMS & PRIMS_QUADS) && type == REF_QUAD) {  // quad: flat/plane.rs:51-80 + flat/quad.rs:37-42
        if (ref == self_ref) return;
        const OwQuad& qd = sc.quads[idx];
        float4 n4 = qd.n;
        if (COUNT) lc.prims++;
        float denom = dot(f3(n4), d);
        if (fabsf(denom) < 1e-8f) return;
        float t = __fdividef(n4.w - dot(f3(n4), o), denom);  // 2 ulp (|denom| >= 1e-8): far below the cancellation in the numerator
        if (!(t >= tmin && t < h.t)) return;
        float3 ip = fma3(d, t, o);
        float4 a4 = qd.a, b4 = qd.b;
        float alpha = dot(f3(a4), ip) - a4.w;
        float beta = dot(f3(b4), ip) - b4.w;
        if (!(0.0f <= alpha && alpha <= 1.0f && 0.0f <= beta && beta <= 1.0f)) return;
        h.t = t;
        h.ref = ref;
        h.b1 = alpha;
        h.b2 = beta;
        return;
    }
}

// ConstantMedium::hit (constant_medium.rs:28-83).  The boundary is a short list of ordinary primitives that are in
// neither the LBVH nor the big list; entry and exit are two closest-hit passes over it (`boundary.hit(r, universe)`
// and `boundary.hit(r, [t1 + 1e-4, inf))`).  The scattering distance is drawn from the ray's own counter-based draw
// (`ray_rnd`, hashed with the medium index) — seeded, where the reference uses the process-global RNG.
template <bool COUNT, int PRIMS>
__device__ __forceinline__ void ow_medium_test(const DevScene& sc, int ref, const RayPre& pre, float a_dd, float time,
                                               float tmin, unsigned ray_rnd, OwHit& h, LocalCount<COUNT>& lc) {
    const OwMedium m = sc.media[ref_index(ref)];
    constexpr int SURF = PRIMS & ~PRIMS_MEDIA;
    float t1 = RL_INF, t2 = RL_INF;
    for (int k = 0; k < m.ref_count; k++) {
        OwHit b;
        b.t = RL_INF; b.ref = -1; b.b1 = b.b2 = 0.0f;
        ow_leaf_test<COUNT, SURF>(sc, sc.medium_refs[m.ref_begin + k], pre, a_dd, time, -1, -RL_INF, b, lc);
        t1 = fminf(t1, b.t);
    }
    if (!(t1 < RL_INF)) return;
    const float lo2 = t1 + fmaxf(1e-4f, 1e-6f * fabsf(t1));  // the reference's 1e-4, kept above the f32 ulp of t1
    for (int k = 0; k < m.ref_count; k++) {
        OwHit b;
        b.t = RL_INF; b.ref = -1; b.b1 = b.b2 = 0.0f;
        ow_leaf_test<COUNT, SURF>(sc, sc.medium_refs[m.ref_begin + k], pre, a_dd, time, -1, lo2, b, lc);
        t2 = fminf(t2, b.t);
    }
    if (!(t2 < RL_INF)) return;
    t1 = fmaxf(t1, tmin);
    t2 = fminf(t2, h.t);
    if (t1 >= t2) return;
    t1 = fmaxf(t1, 0.0f);
    const float ray_length = sqrtf(a_dd);
    const float inside = (t2 - t1) * ray_length;
    const float u = u01(hash32(ray_rnd ^ (0x9E3779B9u * (unsigned)(ref_index(ref) + 1))));
    const float hit_distance = m.neg_inv_density * logf(u);  // u = 0: +inf, no scatter
    if (hit_distance > inside) return;
    h.t = t1 + hit_distance / ray_length;
    h.ref = ref;
    h.b1 = h.b2 = 0.0f;
}

// same test from (o, d) plus the watertight shear constants the v5 kernel computes once per ray (make_pre per triangle
// test was 15 % of the Cornell-box warp instructions, at 4-6 lanes: profiles/r01_ncu_k_ow_render_v5_c5.json)
struct TriShear {
    int k;  // kx | ky << 2 | kz << 4
    float Sx, Sy, Sz;
};
__device__ __forceinline__ TriShear make_shear(float3 o, float3 d) {
    RayPre r = make_pre(o, d);
    TriShear t;
    t.k = r.kx | (r.ky << 2) | (r.kz << 4);
    t.Sx = r.Sx; t.Sy = r.Sy; t.Sz = r.Sz;
    return t;
}
template <bool COUNT, int PRIMS>
__device__ __forceinline__ void ow_leaf_test_od(const DevScene& sc, int ref, float3 o, float3 d, const TriShear& sh, float time,
                                                int self_ref, float tmin, OwHit& h, LocalCount<COUNT>& lc, unsigned ray_rnd = 0u) {
    RayPre pre;
    pre.o = o;
    pre.d = d;
    (void)sh;  // the OW triangle test is the plane form since round 2: no per-ray shear constants
    // the v5 kernel normalises every ray direction when its traversal starts, so a = d.d is the CONSTANT 1 here and
    // the divisions by it in the sphere test fold away (they were ~16 instructions per test)
    ow_leaf_test<COUNT, PRIMS>(sc, ref, pre, 1.0f, time, self_ref, tmin, h, lc, ray_rnd);
}

// state of one path (per lane)
struct Path {
    float3 o, d;
    float3 thr;
    float time;
    int depth;
    int self_ref;
};

// t_min of a ray: the reference's f64 1e-10 (camera.rs:242) cannot hold for f32 origins; 1e-5 x max|origin_k| + 1e-6 in units of
// the NORMALISED direction (the kernels normalise every direction when its traversal starts).
// hit record + material evaluation for one bounce; returns false when the path ends
template <bool COUNT, int PRIMS = PRIMS_ALL>
__device__ __forceinline__ bool ow_shade(const DevScene& sc, const OwCam& cam, Path& p, const OwHit& h, uint4 rnd,
                                         float3& rad, LocalCount<COUNT>& lc);

// hit record + emitted + scatter for the closest hit `h` of path `p` (camera.rs:246-258).  Returns false when the path
// ends; `rad` is then the sample's colour.  (Only terminal events add radiance — a miss adds the background, a
// DiffuseLight emits and never scatters, every scattering material emits black — so the reference's
// `emitted + attenuation * ray_color(..)` recursion collapses to throughput x the single terminal term.)
template <bool COUNT, int PRIMS>
__device__ __forceinline__ bool ow_shade(const DevScene& sc, const OwCam& cam, Path& p, const OwHit& h, uint4 rnd,
                                         float3& rad, LocalCount<COUNT>& lc) {
    rad = f3(0.0f, 0.0f, 0.0f);
    if (h.ref < 0) {  // camera.rs:256-258
        rad = p.thr * cam.background;
        return false;
    }
    if (COUNT) lc.shades++;
    int type = ref_type(h.ref), idx = ref_index(h.ref);
    float3 pos = p.o, normal;
    float u = h.b1, v = h.b2;
    int mat_id = 0;
    bool uv_from_sphere = false;
    float3 outward = f3(0.0f, 1.0f, 0.0f);
    bool in_medium = false;
    if ((PRIMS & PRIMS_MEDIA) && type == REF_MEDIUM) {  // constant_medium.rs:68-75: p = r.at(t), the rest is arbitrary
        pos = fma3(p.d, h.t, p.o);
        outward = f3(1.0f, 0.0f, 0.0f);
        u = v = 0.0f;
        mat_id = sc.media[idx].material;
        in_medium = true;
    } else if (PRIMS == PRIMS_SPHERES || ((PRIMS & PRIMS_SPHERES) && type == REF_SPHERE)) {
        float4 c = sc.spheres[idx].c, dc = sc.spheres[idx].dc;
        float3 center = fma3(f3(dc), p.time, f3(c));
        float3 q = fma3(p.d, h.t, p.o) - center;
        outward = normalize_precise(q) * (c.w < 0.0f ? -1.0f : 1.0f);  // (p - c) / r keeps the sign of r
        pos = fma3(normalize_precise(q), fabsf(c.w), center);           // snapped back onto the surface
        mat_id = __float_as_int(dc.w);
        uv_from_sphere = true;
    } else if ((PRIMS & PRIMS_TRIS) && type == REF_TRI) {
        float4 p0 = sc.tri_verts[idx].p0, p1 = sc.tri_verts[idx].p1, p2 = sc.tri_verts[idx].p2;
        float b0 = 1.0f - h.b1 - h.b2;
        pos = f3(p0) * b0 + f3(p1) * h.b1 + f3(p2) * h.b2;
        int flags = __float_as_int(p2.w);
        float4 s0 = sc.tri_shade[idx].s0, s1 = sc.tri_shade[idx].s1, s2 = sc.tri_shade[idx].s2;
        if (flags & 1) outward = normalize_precise(f3(s1) * h.b1 + f3(s2) * h.b2 + f3(s0) * b0);
        else outward = f3(s0);
        if (flags & 2) {
            float4 s3 = sc.tri_shade[idx].s3;
            u = s0.w * b0 + s2.w * h.b1 + s3.y * h.b2;
            v = s1.w * b0 + s3.x * h.b1 + s3.z * h.b2;
        }
        mat_id = __float_as_int(p0.w);
    } else if ((PRIMS & PRIMS_QUADS) && type == REF_QUAD) {
        const OwQuad& qd = sc.quads[idx];
        outward = f3(qd.n);
        float3 ip = fma3(p.d, h.t, p.o);
        pos = ip - outward * (dot(outward, ip) - qd.n.w);  // snapped onto the plane
        mat_id = qd.m.x;
    }
    bool front = in_medium || dot(p.d, outward) <= 0.0f;  // hittable/mod.rs:32-38 (a medium hit is Face::Front)
    normal = front ? outward : -outward;
    RL_CHECK_OR(mat_id >= 0 && mat_id < sc.n_materials, lc, return false);
    const DevMaterial m = sc.materials[mat_id];
    int kind = __float_as_int(m.b.w);
    int tex = __float_as_int(m.color.w);
    if (uv_from_sphere && tex >= 0 && sc.n_images > 0) {  // get_sphere_uv (sphere.rs:91-99); only image textures read it
        float theta = acosf(fminf(fmaxf(-outward.y, -1.0f), 1.0f));
        float phi = atan2f(-outward.z, outward.x) + 3.14159265358979f;
        u = phi * (1.0f / 6.28318530717959f);
        v = theta * (1.0f / 3.14159265358979f);
    }
    if (kind == RL_MAT_OW_DIFFUSE_LIGHT) {  // material.rs:178-195
        rad = p.thr * tex_value<PRIMS>(sc, tex, u, v, pos);
        return false;
    }
    float3 dir;
    float3 atten;
    // Lambertian and Metal both draw a UnitSphere sample (Metal even with fuzz = 0, material.rs:113): evaluated once,
    // before the material branches, so the warp runs it converged
    const float3 uvec = kind == RL_MAT_OW_DIELECTRIC ? f3(0.0f, 0.0f, 0.0f) : unit_vector(u01(rnd.x), u01(rnd.y));
    if ((PRIMS & PRIMS_MEDIA) && kind == RL_MAT_OW_ISOTROPIC) {  // material.rs:201-216
        dir = uvec;
        atten = tex_value<PRIMS>(sc, tex, u, v, pos);
    } else if (kind == RL_MAT_OW_LAMBERTIAN) {  // material.rs:74-92
        dir = normal + uvec;
        if (fabsf(dir.x) <= 1e-8f && fabsf(dir.y) <= 1e-8f && fabsf(dir.z) <= 1e-8f) dir = normal;
        atten = tex_value<PRIMS>(sc, tex, u, v, pos);
    } else if (kind == RL_MAT_OW_METAL) {  // material.rs:105-122
        float3 refl = p.d - normal * (2.0f * dot(p.d, normal));
        dir = fma3(uvec, m.a.x, normalize(refl));
        if (!(dot(dir, normal) > 0.0f)) return false;
        atten = f3(m.color);
    } else {  // Dielectric, material.rs:139-176
        float ri = front ? 1.0f / m.a.y : m.a.y;
        float3 unit = normalize(p.d);
        float cos_theta = fminf(-dot(unit, normal), 1.0f);
        float sin_theta = sqrtf(fmaxf(0.0f, 1.0f - cos_theta * cos_theta));
        bool reflect = ri * sin_theta > 1.0f;
        if (!reflect) {
            float r0 = (1.0f - ri) / (1.0f + ri);
            r0 *= r0;
            float x = 1.0f - cos_theta;
            float x2 = x * x;
            reflect = (r0 + (1.0f - r0) * (x2 * x2 * x)) > u01(rnd.z);
        }
        if (reflect) {
            dir = unit - normal * (2.0f * dot(unit, normal));
        } else {  // vec3.rs:224-230
            float3 perp = (unit + normal * cos_theta) * ri;
            float3 par = normal * (-sqrtf(fabsf(1.0f - dot(perp, perp))));
            dir = perp + par;
        }
        atten = f3(1.0f, 1.0f, 1.0f);
    }
    p.thr = p.thr * atten;
    p.o = pos;
    p.d = dir;
    p.self_ref = h.ref;
    p.depth--;
    return p.depth > 0;  // depth exhausted: the remaining term is black (camera.rs:239-241)
}

// get_ray (camera.rs:203-230)
__device__ __forceinline__ void ow_camera_ray(const OwCam& cam, int i, int j, unsigned sample, Path& p) {
    uint2 key = make_uint2(cam.seed_lo, cam.seed_hi);
    unsigned pixel = (unsigned)(j * cam.width + i);
    // ONE Philox call per camera ray: the sub-pixel jitter takes the two 16-bit halves of one word (1/65536 of a pixel
    // is far below what a sample count can resolve), time and the two defocus-disc coordinates a word each
    uint4 r = philox(make_uint4(pixel, sample, 0u, 0u), key);
    float3 center = cam.pixel00 + cam.du * (float)i + cam.dv * (float)j;
    float px = -0.5f + (float)(r.x >> 16) * (1.0f / 65536.0f), py = -0.5f + (float)(r.x & 0xffffu) * (1.0f / 65536.0f);
    float3 sample_p = center + cam.du * px + cam.dv * py;
    float3 origin = cam.lookfrom;
    if (cam.defocus) {
        // uniform point in the unit disc (polar map; UnitDisc's rejection loop is statistically identical)
        float rr = fast_sqrt(u01(r.w));
        float s, c;
        fast_sincos_turn(u01(r.y), &s, &c);
        origin = cam.lookfrom + cam.disk_u * (rr * c) + cam.disk_v * (rr * s);
    }
    p.o = origin;
    p.d = sample_p - origin;
    p.time = u01(r.z);
    p.thr = f3(1.0f, 1.0f, 1.0f);
    p.depth = cam.max_depth;
    p.self_ref = -1;
}

// ---- v4: resumable traversal + threshold-triggered service ---------------------------------------------------------
// ncu on v3 (profiles/r01_ncu_k_ow_render_v3.json, r01_ncu_ow_c5.json): 10.8 of 32 lanes per issued instruction on the
// cover scene and 6.3 on the Cornell box — a bounce (traversal + shade) ends only when the slowest lane's traversal
// does.  Here the traversal state of a lane (node, stack, hit) survives across "service" rounds: the warp leaves the
// traversal loop as soon as `svc_min` lanes wait for service (shade + scatter, retire, refill, regenerate), services
// exactly those lanes, and drops back into the loop where the other lanes simply resume (Aila & Laine 2009's
// persistent while-while with dynamic fetch, with the fetch replaced by a full material evaluation).  The v2 state
// machine tried the same with one ballot-scheduled action per single step and lost to its own scheduling overhead;
// this version keeps v3's tight inner loops and pays two ballots per leaf round.  Per-lane arithmetic is unchanged,
// so the image is bit-identical to v3's.
//
// v5: the while-while loop still wastes lanes at SEGMENT level — every leaf round waits for the lane with the longest
// run of inner nodes (mean ~5, max over 32 lanes ~20 on the cover scene, which is the measured 10-12 of 32 lanes).  So
// every iteration is ONE node step for every lane at an inner node; lanes that reach a leaf park, and the leaf test
// runs for all parked lanes together once `leaf_min` of them wait (or nobody can step).  leaf_min = 32 degenerates to
// the while-while schedule (v4).  ncu (profiles/r01_ncu_k_ow_render_v4.json, _v5.json): node steps run at 20 instead of 12.5
// lanes, the whole kernel at 16 instead of 11.3.
// ---- CTA-level reserve of work items ------------------------------------------------------------------------------------
// Warps take their batches (<= 64 items) from a reserve the whole CTA shares; the reserve refills from the GLOBAL queue
// with one atomic per `qbatch` items (512), the next refill prefetched one reserve ahead so that its round trip — over
// NVLink when the counter is rank 0's — overlaps rendering.  Round 1 (and the first half of round 2) popped the global
// counter once per WARP batch: 790 k same-address atomics per cover-scene step, 56 M/s at 8 GPUs, all landing on one L2
// slice of GPU 0 — the ~1.7 ms per step that did not scale (14.0 ms at 8 GPUs against 12.3 ms = 98.5 / 8).
#ifndef RL_GUIDED_WCAP
#define RL_GUIDED_WCAP 64
#endif
struct ItemReserve {
    volatile int it_lock, it_dry, nx_size;
    volatile long long it_next, it_end;
    volatile unsigned long long nx_base;  // prefetched next refill (its atomic was issued one refill ago)
};
__device__ __forceinline__ void reserve_init(ItemReserve& r, unsigned long long* queue, int sys_queue, int qbatch) {
    r.it_lock = 0;
    r.it_dry = 0;
    r.it_next = r.it_end = 0;
    r.nx_size = qbatch;
    r.nx_base = sys_queue ? atomicAdd_system(queue, (unsigned long long)qbatch) : atomicAdd(queue, (unsigned long long)qbatch);
}
// `n` work items for one service batch, out of the CTA's reserved batch of the global queue (one lane calls).  The next
// batch's global atomic (system scope over NVLink when the counter is rank 0's) was issued when the current one was
// installed, so its round trip overlaps a whole batch of rendering.  Returns how many were granted from *start on.
__device__ __forceinline__ int items_take(ItemReserve& ctl, int n, long long n_items, unsigned long long* queue, int sys_queue, int qbatch,
                                          int qtail, long long q_guided, long long* start, int* dry) {
    while (atomicCAS((int*)&ctl.it_lock, 0, 1) != 0) __nanosleep(32);
    __threadfence_block();
    long long nx = ctl.it_next, en = ctl.it_end;
    int isdry = ctl.it_dry;
    if (nx >= en && !isdry) {
        nx = (long long)ctl.nx_base;
        const int got = ctl.nx_size;
        en = nx + got < n_items ? nx + got : n_items;
        if (nx >= n_items) {
            isdry = 1;
            en = nx;
            ctl.it_dry = 1;
        } else {
            // guided self-scheduling: full batches while the queue is long, small ones near its end
            const int want = n_items - nx > q_guided ? qbatch : qtail;
            ctl.nx_size = want;
            ctl.nx_base = sys_queue ? atomicAdd_system(queue, (unsigned long long)want) : atomicAdd(queue, (unsigned long long)want);
        }
        ctl.it_end = en;
    }
    const long long avail = en - nx;
    // in the guided phase (small refills near the end of the queue) a warp takes at most RL_GUIDED_WCAP items at a time, so
    // that what a lane still owns when the queue runs dry is one item, not two or three
    if (n > RL_GUIDED_WCAP && n_items - nx <= q_guided) n = RL_GUIDED_WCAP;
    const int take = (long long)n < avail ? n : (int)avail;
    *start = nx;
    ctl.it_next = nx + take;
    *dry = isdry;
    __threadfence_block();
    atomicExch((int*)&ctl.it_lock, 0);
    return take;
}

// TRACE = true runs caller-supplied rays through the SAME loop (work items = ray indices, "service" = write the rl_hit of
// the finished ray and load the next one): rl_trace_batch is this kernel, not a second traversal.
struct TraceIO {
    const rl_ray* rays;
    const int* self_refs;  // optional: the leaf ref each ray starts on (-1 none)
    rl_hit* hits;
};
#ifndef RL_OW_OPT  // node steps per ballot; experiment builds override it (tools/build_alt.py)
#define RL_OW_OPT 3
#endif
template <bool COUNT, int MINB, int PRIMS, bool TRACE = false, int OPT = RL_OW_OPT>
__global__ void __launch_bounds__(256, MINB) k_ow_render5(DevScene sc, OwCam cam, JobTable jt, float* __restrict__ partial,
                                                          unsigned long long* __restrict__ queue, Counters* counters,
                                                          int sys_queue, int qbatch, int wbatch, long long q_guided, int svc_min,
                                                          int leaf_min, TraceIO tio) {
    LocalCount<COUNT> lc;
    const unsigned lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const uint2 key = make_uint2(cam.seed_lo, cam.seed_hi);
#ifdef RL_TIMELINE  // experiment build (tools/build_alt.py): warp lifetimes into the statistics counters
    unsigned long long tl_start, tl_dry = 0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tl_start));
#endif
    // item.  Cold per-lane state (touched once per path or once per item, never inside the traversal loop) lives in
    // shared memory, [word][thread] so every access is conflict free; that is ~10 registers the traversal loop gets
    // back (ptxas: the 64-register build spilled 262 B with them in registers).
    __shared__ float sm_acc[3][256], sm_thr[3][256];
    __shared__ int sm_x[256], sm_y[256], sm_chunk[256], sm_send[256];
    const int tid = threadIdx.x;
    bool has_item = false, alive = false, done = false;
    int s = 0;
    unsigned pixel = 0, retired = 0;
    // path
    Path p;
    p.depth = 0;
    // traversal (resumable)
    int node = TRAV_END;
    __shared__ int sm_stack[STACK_SM * 256];  // [entry][thread]: the first STACK_SM entries of every lane's traversal stack
    TravStack st;
    stack_init<256>(st, sm_stack + threadIdx.x);
    StackSpill spill;
    float3 inv_d = f3(1.0f, 1.0f, 1.0f), oi = f3(0.0f, 0.0f, 0.0f);
    float tmin = 0.0f;
    TriShear shear;
    shear.k = 0; shear.Sx = shear.Sy = shear.Sz = 0.0f;
    unsigned ray_rnd = 0u;  // the ray's draw for ConstantMedium scattering distances
    OwHit hit;
    hit.t = RL_INF; hit.ref = -1; hit.b1 = hit.b2 = 0.0f;
    // work items: each warp owns a batch [next, end) taken from the CTA's reserve (ItemReserve above), decoded into a table
    __shared__ long long sm_qnext[8], sm_qend[8], sm_qfirst[8];
    __shared__ int sm_qdry[8];
    __shared__ ItemReserve sm_reserve;
    constexpr int OW_ITEM_TABLE = 64;  // the largest batch a warp takes
    __shared__ int sm_it_xy[8][OW_ITEM_TABLE], sm_it_chunk[8][OW_ITEM_TABLE], sm_it_s0[8][OW_ITEM_TABLE], sm_it_s1[8][OW_ITEM_TABLE];
    const int wid = threadIdx.x >> 5;
    if (lane == 0) {
        sm_qnext[wid] = sm_qend[wid] = sm_qfirst[wid] = 0;
        sm_qdry[wid] = 0;
    }
    if (threadIdx.x == 0) reserve_init(sm_reserve, queue, sys_queue, qbatch);
    __syncthreads();
    while (true) {
        // ================= service: every lane that is not mid-traversal =================
        const bool svc = node == TRAV_END && !done;
        if (TRACE && svc && alive) {  // a traced ray came back: its rl_hit (t in the caller's units: the ray ran with |d| = 1)
            rl_hit out;
            out.node = -1;
            out.t = hit.t * sm_thr[0][tid];
            out.u = hit.b1;
            out.v = hit.b2;
            if (hit.ref >= 0) {
                const int type = ref_type(hit.ref), idx = ref_index(hit.ref);
                out.node = type == REF_SPHERE ? sc.sphere_node[idx]
                         : type == REF_QUAD ? sc.quad_node[idx]
                         : type == REF_TRI ? __float_as_int(sc.tri_verts[idx].p1.w) : -1;
            }
            tio.hits[s] = out;
            alive = false;
            has_item = false;
        }
        if (!TRACE && svc && alive) {  // its traversal just finished: hit record + emitted + scatter
            unsigned bounce = (unsigned)(cam.max_depth - p.depth + 1);
            uint4 rnd = philox(make_uint4(pixel, (unsigned)(cam.first_sample + s), bounce, 0u), key);
            float3 rad;
            p.thr = f3(sm_thr[0][tid], sm_thr[1][tid], sm_thr[2][tid]);
            if (!ow_shade<COUNT, PRIMS>(sc, cam, p, hit, rnd, rad, lc)) {
                sm_acc[0][tid] += rad.x;  // samples are folded in order (camera.rs:174)
                sm_acc[1][tid] += rad.y;
                sm_acc[2][tid] += rad.z;
                alive = false;
                s++;
            } else {
                sm_thr[0][tid] = p.thr.x;
                sm_thr[1][tid] = p.thr.y;
                sm_thr[2][tid] = p.thr.z;
            }
        }
        if (!TRACE && svc && has_item && !alive && s == sm_send[tid]) {
            // ONE 16-byte store per finished item (over NVLink when the buffer is rank 0's: three 4-byte stores per
            // item were 134 M small remote writes per 8-GPU cover-scene step)
            size_t idx = ((size_t)sm_chunk[tid] * cam.height + sm_y[tid]) * cam.width + sm_x[tid];
            RL_CHECK(sm_chunk[tid] >= 0 && sm_chunk[tid] < cam.n_chunks && sm_x[tid] < cam.width && sm_y[tid] < cam.height, lc);
            reinterpret_cast<float4*>(partial)[idx] = make_float4(sm_acc[0][tid], sm_acc[1][tid], sm_acc[2][tid], 0.0f);
            has_item = false;
            retired++;
        }
        const bool need = svc && !has_item;
        const unsigned mask = __ballot_sync(FULL, need);
        if (mask) {  // warp-aggregated pop from the reserved batch; the next batch's atomic is already in flight
            long long cur_next = sm_qnext[wid], cur_end = sm_qend[wid];
            bool q_dry = sm_qdry[wid] != 0;
            __syncwarp();
            if (cur_next >= cur_end && !q_dry) {
                // guided self-scheduling lives in the reserve: full refills while the queue is long, small ones near its
                // end, so the last CTAs to finish hold little work
                long long start = 0;
                int take = 0, dry = 0;
                if (lane == 0) take = items_take(sm_reserve, wbatch, jt.n_items, queue, sys_queue, qbatch, qbatch < 128 ? qbatch : 128, q_guided, &start, &dry);
                take = __shfl_sync(FULL, take, 0);
                dry = __shfl_sync(FULL, dry, 0);
                start = __shfl_sync(FULL, start, 0);
                cur_next = start;
                cur_end = start + take;
                if (take == 0 && dry) q_dry = true;
#ifdef RL_TIMELINE
                if (q_dry && !tl_dry) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tl_dry));
#endif
                if (lane == 0) sm_qfirst[wid] = cur_next;
                // Decode the WHOLE batch now, with every lane: item -> (job, chunk, pixel, sample range) is ~300 instructions of
                // 64-bit divisions and a binary search, and round 1 ran it per item at ~2 of 32 lanes — whenever a service
                // round happened to include a lane that had just finished its item — for 8 % of all warp instructions
                // (profiles/r02_ncu_v5_c4_struct_stack.json: svc_queue + kernels.h + tile_pixel lines).
                for (int k = (int)lane; !TRACE && k < OW_ITEM_TABLE; k += 32) {
                    const long long item = cur_next + k;
                    int xy = -1, chunk = 0, s0 = 0, s1 = 0;
                    if (item < cur_end) {
                        const int j = find_job(jt, item);
                        const rl_job job = jt_job(jt, j);
                        const long long local = item - jt_prefix(jt, j);
                        const int w = job.x1 - job.x0, hgt = job.y1 - job.y0;
                        const long long pp = padded_pixels(w, hgt);
                        const int ck = (int)(local / pp);
                        int px, py;
                        tile_pixel(w, hgt, local - (long long)ck * pp, &px, &py);
                        if (px < w && py < hgt) {
                            chunk = job.chunk_begin + ck;
                            ow_chunk_range(cam.spp, cam.n_chunks, chunk, &s0, &s1);
                            if (cam.max_depth <= 0) s0 = s1;  // depth 0: every sample is black (camera.rs:239-241)
                            xy = ((job.y0 + py) << 16) | (job.x0 + px);
                        }
                    }
                    sm_it_xy[wid][k] = xy;
                    sm_it_chunk[wid][k] = chunk;
                    sm_it_s0[wid][k] = s0;
                    sm_it_s1[wid][k] = s1;
                }
                __syncwarp();
            }
            long long avail = cur_end - cur_next;
            int rank_in = __popc(mask & ((1u << lane) - 1u));
            if (need) {
                if (TRACE && rank_in < avail) {  // work item = ray index
                    const long long item = cur_next + rank_in;
                    const rl_ray r = tio.rays[item];
                    p.o = f3(r.origin[0], r.origin[1], r.origin[2]);
                    p.d = f3(r.direction[0], r.direction[1], r.direction[2]);
                    p.time = r.time;
                    p.self_ref = tio.self_refs ? tio.self_refs[item] : -1;
                    p.depth = 1;
                    sm_thr[0][tid] = rsqrtf(dot(p.d, p.d));  // t(caller) = t(unit direction) / |d|: the same factor normalises d below
                    s = (int)item;
                    has_item = true;
                    alive = true;
                } else if (rank_in < avail) {
                    // the batch was decoded when it was installed (all 32 lanes, two items each): three shared-memory reads
                    const int k = (int)(cur_next + rank_in - sm_qfirst[wid]);
                    const int xy = sm_it_xy[wid][k];
                    if (xy >= 0) {  // padded slots outside the rectangle are simply skipped
                        const int x = xy & 0xffff, y = xy >> 16;
                        s = sm_it_s0[wid][k];
                        pixel = (unsigned)(y * cam.width + x);
                        sm_x[tid] = x;
                        sm_y[tid] = y;
                        sm_chunk[tid] = sm_it_chunk[wid][k];
                        sm_send[tid] = sm_it_s1[wid][k];
                        sm_acc[0][tid] = sm_acc[1][tid] = sm_acc[2][tid] = 0.0f;
                        has_item = true;
                    } else {
                        retired++;  // nothing to render, but it IS a work item of the job list (rl_ow_job_items counts padded slots)
                    }
                } else if (q_dry) {
                    done = true;
                }
            }
            int taken = __popc(mask);
            cur_next += taken < avail ? taken : avail;
            if (lane == 0) {
                sm_qnext[wid] = cur_next;
                sm_qend[wid] = cur_end;
                sm_qdry[wid] = q_dry ? 1 : 0;
            }
            __syncwarp();
        }
        if (!TRACE && svc && has_item && !alive && s < sm_send[tid]) {  // path regeneration
            ow_camera_ray(cam, sm_x[tid], sm_y[tid], (unsigned)(cam.first_sample + s), p);
            sm_thr[0][tid] = sm_thr[1][tid] = sm_thr[2][tid] = 1.0f;
            alive = true;
        }
        if (svc && alive) {  // start the traversal of the lane's new ray
            // unit direction: t is internal to this kernel (hit point = o + d t either way), and with |d| = 1 the
            // sphere quadratic, the medium's ray length and the tmin scale lose their divisions by d.d
            p.d = p.d * rsqrtf(dot(p.d, p.d));
            inv_d = safe_inv_fast(p.d);
            oi = p.o * inv_d;
            tmin = fmaf(1e-5f, max_abs(p.o), 1e-6f);  // t_min, see Path
            if ((PRIMS & PRIMS_MEDIA) && !TRACE)
                ray_rnd = philox(make_uint4(pixel, (unsigned)(cam.first_sample + s), (unsigned)(cam.max_depth - p.depth + 1), 2u), key).x;
            hit.t = RL_INF; hit.ref = -1; hit.b1 = hit.b2 = 0.0f;
            stack_reset(st);
            if (COUNT) lc.rays++;
            for (int k = 0; k < sc.n_big; k++)  // the big list: once per ray, here, with the serviced lanes
                ow_leaf_test_od<COUNT, PRIMS>(sc, sc.big_refs[k], p.o, p.d, shear, p.time, p.self_ref, tmin, hit, lc, ray_rnd);
            node = sc.n_bvh_prims > 0 ? 0 : TRAV_END;
        }
        const unsigned m_done = __ballot_sync(FULL, done);
        if (m_done == FULL) break;
        // ================= traversal: until svc_min lanes wait for service =================
        const int n_done = __popc(m_done);
        while (true) {
            // leaf round: every lane parked at a leaf tests it and pops
            if (node < 0 && node != TRAV_END) {
                ow_leaf_test_od<COUNT, PRIMS>(sc, ~node, p.o, p.d, shear, p.time, p.self_ref, tmin, hit, lc, ray_rnd);
                node = stack_pop<256>(st, spill);
            }
            const int n_end = __popc(__ballot_sync(FULL, node == TRAV_END));
            if (n_end == 32 || n_end - n_done >= svc_min) break;
            // node steps, one per lane per iteration, until leaf_min lanes have parked (or ended) or none can step
            const int keep = 32 - n_end - leaf_min;
            int n_in;
            do {
                // OPT node steps per ballot.  Round 1 (58-instruction steps): 1 -> 12.44 / 67.4 ms (C4 / C5 at reduced spp),
                // 2 -> 12.11 / 64.7, 3 -> 12.39 / 65.6.  Round 2's step is 40 % shorter, so the ballot + loop control weighs
                // more (13 % of the warp instructions at two steps): 3 -> C4 98.4 -> 97.95 ms, C5 57.7 -> 55.6 ms; 4 -> 98.1 / 56.0
                // against 96.5 / 55.7 at 3.  Running the step for ALL lanes under predicates instead of branching around it
                // (no BSSY / BRA / BSYNC per step) was 8 % slower: 104.4 - 108.4 ms (gpurun_out/ab2).
#pragma unroll
                for (int k = 0; k < (OPT < 1 ? 1 : OPT); k++) {
                    if (node >= 0) bvh2_step<COUNT, 256>(sc.nodes, node, st, spill, inv_d, oi, tmin, hit.t, lc, sc.n_bvh_nodes);
                }
                n_in = __popc(__ballot_sync(FULL, node >= 0));
            } while (n_in > keep && n_in > 0);
        }
    }
    // completion accounting: items this GPU finished, added to the counter next to the queue (rank 0's over NVLink when the
    // queue is shared), so the owner can tell a finished render from one a peer dropped out of
    if (!TRACE) {
        for (int off = 16; off > 0; off >>= 1) retired += __shfl_xor_sync(FULL, retired, off);
        if (lane == 0 && retired) {
            if (sys_queue) atomicAdd_system(queue + 1, (unsigned long long)retired);
            else atomicAdd(queue + 1, (unsigned long long)retired);
        }
    }
#ifdef RL_TIMELINE
    if (!COUNT && !TRACE && lane == 0) {
        unsigned long long tl_end;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tl_end));
        atomicAdd(&counters->rays, 1ull);                                         // warps
        atomicAdd(&counters->node_visits, tl_end - tl_start);                     // sum of warp lifetimes (ns)
        atomicAdd(&counters->prim_tests, (tl_dry ? tl_dry : tl_end) - tl_start);  // sum of time until the warp found the queue dry
        atomicMax(&counters->shades, tl_end - tl_start);                          // longest warp lifetime
        atomicMax(&counters->tri_tests, tl_dry ? tl_end - tl_dry : 0ull);         // longest drain of one warp
    }
#endif
    lc.flush(counters);
}

// ---- v6: CTA-pooled paths — ready / done queues in shared memory, any warp traverses, any warp services ---------------
// ncu on v5 (profiles/r01_ncu_k_ow_render5_c4_500spp.json): 17 of 32 lanes per issued instruction.  The loss is
// structural: a warp's lanes are in one of three states (inner node, leaf, waiting for service) and only one state
// runs at a time; with the measured-best service threshold of 24, about 12 lanes idle through the node steps, and the
// service round itself runs at ~10 lanes because materials diverge.  v6 takes the path OUT of the lane:
//   * a CTA owns P path slots in shared memory (SoA, [word][slot]): ray, throughput, item bookkeeping, hit record;
//   * a lane that finishes its traversal writes the hit into the slot, pushes the slot id on the DONE ring and takes the
//     next ray from the READY ring at once (ballot / popc compaction, one shared-memory atomic per warp and ring), so
//     node steps run with nearly full warps whatever the other lanes' rays do;
//   * any warp that finds 32 ids on the DONE ring services them as ONE full batch (hit record + emitted + scatter,
//     retire, item refill from the global queue, camera-ray regeneration, big-list test) and pushes the new rays on the
//     READY ring.  Nothing ever waits for another warp to do a particular thing: every warp does whatever work exists.
// This is north_star's wavefront (stages separated by queues, warp-level compaction) at CTA scope: the queues hold
// 4-byte slot ids and never leave shared memory, where a global wavefront round-trips 128 B per ray-bounce through L2 /
// HBM.  Per-lane arithmetic — leaf tests, shading, RNG counters, the order samples are folded in — is
// v5's, so the image is bit-identical to v5's for any schedule (tests/test_gpu_ow.py).
// MODE_TRACE runs caller-supplied rays through the SAME queues, refill, big-list start, node steps and leaf rounds and
// writes rl_hit records: rl_trace_batch is the production traversal, not a second implementation.
namespace v6 {
constexpr int THREADS = 256, QCAP = 512, EMPTY = -1;
constexpr int MODE_RENDER = 0, MODE_TRACE = 1;
// slot words
enum { OX, OY, OZ, DX, DY, DZ, TIME, SELF, HT, HREF, THRX, THRY, THRZ, DEPTH, XY, S, SEND, CHUNK, ACCX, ACCY, ACCZ, FLAGS, N_BASE };
constexpr int HB1 = N_BASE, HB2 = N_BASE + 1, RND = N_BASE + 2;
__host__ __device__ constexpr int slot_words(int prims) {
    return (prims & PRIMS_MEDIA) ? N_BASE + 3 : ((prims & (PRIMS_TRIS | PRIMS_QUADS)) ? N_BASE + 2 : N_BASE);
}
__host__ __device__ constexpr size_t smem_bytes(int prims, int P) {
    return (size_t)slot_words(prims) * P * 4 + (size_t)STACK_SM * THREADS * 4 + 2 * (size_t)QCAP * 4;
}

struct Ctl {
    volatile unsigned r_head, r_tail, d_head, d_tail;  // READY / DONE rings
    volatile int live, abort;                         // slots not yet dead; watchdog
    ItemReserve items;                                // CTA-level item reserve (refilled from the global queue)
};

using rl::TraceIO;

__device__ __forceinline__ void ring_push(int* ring, volatile unsigned* tail, unsigned mask, bool pred, int id, unsigned lane) {
    if (!mask) return;
    const int n = __popc(mask), leader = __ffs(mask) - 1;
    unsigned base = 0;
    if ((int)lane == leader) base = atomicAdd((unsigned*)tail, (unsigned)n);
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred) {
        volatile int* e = ring + ((base + __popc(mask & ((1u << lane) - 1u))) & (QCAP - 1));
        while (*e != EMPTY) {}  // a consumer that reserved this entry a whole lap ago has not read it yet (never seen; kept for safety)
        *e = id;
    }
    __syncwarp();
}

// up to popc(mask) ids for the lanes in `mask`; -1 for lanes that got none
__device__ __forceinline__ int ring_pop(int* ring, volatile unsigned* head, volatile unsigned* tail, unsigned mask, bool pred,
                                        unsigned lane) {
    if (!mask) return -1;
    const int want = __popc(mask), leader = __ffs(mask) - 1;
    unsigned h = 0;
    int got = 0;
    if ((int)lane == leader) {
        h = *head;
        while (true) {
            const int avail = (int)(*tail - h);
            got = avail < want ? avail : want;
            if (got <= 0) { got = 0; break; }
            const unsigned old = atomicCAS((unsigned*)head, h, h + (unsigned)got);
            if (old == h) break;
            h = old;
        }
    }
    got = __shfl_sync(0xffffffffu, got, leader);
    h = __shfl_sync(0xffffffffu, h, leader);
    int id = -1;
    if (pred) {
        const int r = __popc(mask & ((1u << lane) - 1u));
        if (r < got) {
            volatile int* e = ring + ((h + (unsigned)r) & (QCAP - 1));
            while ((id = *e) == EMPTY) {}  // the producer bumped the tail and writes the entry right after
            *e = EMPTY;
        }
    }
    __threadfence_block();  // slot data written before the id was pushed is visible from here on
    __syncwarp();
    return id;
}

}  // namespace v6

template <bool COUNT, int MINB, int PRIMS, int MODE>
__global__ void __launch_bounds__(256, MINB) k_ow_render6(DevScene sc, OwCam cam, JobTable jt, float* __restrict__ partial,
                                                          unsigned long long* __restrict__ queue, Counters* counters,
                                                          int sys_queue, int qbatch, long long q_guided, int P, int svc_lo,
                                                          int exit_min, int leaf_min, TraceIO tio) {
    using namespace v6;
    constexpr bool HAS_B = (PRIMS & (PRIMS_TRIS | PRIMS_QUADS | PRIMS_MEDIA)) != 0;
    constexpr int NW = slot_words(PRIMS);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* const pool = reinterpret_cast<float*>(smem_raw);
    int* const pooli = reinterpret_cast<int*>(smem_raw);
    int* const stack_base = pooli + NW * P;
    int* const ring_r = stack_base + STACK_SM * THREADS;
    int* const ring_d = ring_r + QCAP;
    __shared__ Ctl ctl;
#define SL(w, id) pool[(w) * P + (id)]
#define SLI(w, id) pooli[(w) * P + (id)]
    LocalCount<COUNT> lc;
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x;
    const unsigned lane = threadIdx.x & 31;
    const uint2 key = make_uint2(cam.seed_lo, cam.seed_hi);
    const long long n_items = jt.n_items;
    // every slot starts on the DONE ring with no item: the first service rounds hand out the first items
    for (int i = tid; i < QCAP; i += THREADS) {
        ring_r[i] = EMPTY;
        ring_d[i] = i < P ? i : EMPTY;
    }
    for (int i = tid; i < P; i += THREADS) {
        SLI(CHUNK, i) = -1;
        SLI(FLAGS, i) = 0;
    }
    if (tid == 0) {
        ctl.r_head = ctl.r_tail = 0;
        ctl.d_head = 0;
        ctl.d_tail = (unsigned)P;
        ctl.live = P;
        ctl.abort = 0;
        reserve_init(ctl.items, queue, sys_queue, qbatch);
    }
    __syncthreads();

    // ---- lane state: the ray this lane is traversing ----
    int slot = -1, node = TRAV_END;
    TravStack st;
    stack_init<THREADS>(st, stack_base + tid);
    StackSpill spill;
    float3 o = f3(0.0f, 0.0f, 0.0f), d = f3(0.0f, 0.0f, 1.0f), inv_d = f3(1.0f, 1.0f, 1.0f), oi = f3(0.0f, 0.0f, 0.0f);
    float tmin = 0.0f, time = 0.0f;
    int self_ref = -1;
    TriShear shear;
    shear.k = 0; shear.Sx = shear.Sy = shear.Sz = 0.0f;
    unsigned ray_rnd = 0u;
    OwHit hit;
    hit.t = RL_INF; hit.ref = -1; hit.b1 = hit.b2 = 0.0f;
    unsigned spins = 0, retired = 0;

    while (true) {
        // ================= (A) lanes whose traversal ended: hit record into the slot, slot id onto the DONE ring =================
        const bool fin = slot >= 0 && node == TRAV_END;
        if (fin) {
            SL(HT, slot) = hit.t;
            SLI(HREF, slot) = hit.ref;
            if (HAS_B) {
                SL(HB1, slot) = hit.b1;
                SL(HB2, slot) = hit.b2;
            }
        }
        const unsigned m_fin = __ballot_sync(FULL, fin);
        if (m_fin) {
            __threadfence_block();
            ring_push(ring_d, &ctl.d_tail, m_fin, fin, slot, lane);
            if (fin) slot = -1;
        }
        // ================= (B) idle lanes take the next ray from the READY ring =================
        const unsigned m_idle = __ballot_sync(FULL, slot < 0);
        if (m_idle) {
            const int got = ring_pop(ring_r, &ctl.r_head, &ctl.r_tail, m_idle, slot < 0, lane);
            if (got >= 0) {
                slot = got;
                o = f3(SL(OX, got), SL(OY, got), SL(OZ, got));
                d = f3(SL(DX, got), SL(DY, got), SL(DZ, got));  // unit length (the service round normalised it)
                time = SL(TIME, got);
                self_ref = SLI(SELF, got);
                hit.t = SL(HT, got);  // the big list was tested when the ray was made: traversal starts with a finite tmax
                hit.ref = SLI(HREF, got);
                if (HAS_B) {
                    hit.b1 = SL(HB1, got);
                    hit.b2 = SL(HB2, got);
                }
                if (PRIMS & PRIMS_MEDIA) ray_rnd = (unsigned)SLI(RND, got);
                inv_d = safe_inv_fast(d);
                oi = o * inv_d;
                tmin = fmaf(1e-5f, max_abs(o), 1e-6f);  // t_min, see Path
                if (PRIMS & PRIMS_TRIS) shear = make_shear(o, d);
                stack_reset(st);
                node = sc.n_bvh_prims > 0 ? 0 : TRAV_END;
            }
        }
        const int n_act = __popc(__ballot_sync(FULL, slot >= 0));
        // ================= (C) service: a full batch whenever one is waiting; a partial one when this warp runs dry =================
        const int n_done = (int)(ctl.d_tail - ctl.d_head);
        const int do_svc = __shfl_sync(FULL, (n_done >= 32 || (n_done > 0 && n_act <= svc_lo)) ? 1 : 0, 0);
        if (do_svc) {
            const int id = ring_pop(ring_d, &ctl.d_head, &ctl.d_tail, FULL, true, lane);
            bool has_item = false, alive = false, requeue = false;
            int s = 0, s_end = 0, xy = 0;
            unsigned pixel = 0;
            Path p;
            p.depth = 0;
            if (id >= 0) {
                has_item = SLI(CHUNK, id) >= 0;
                alive = (SLI(FLAGS, id) & 1) != 0;
                if (has_item) {
                    s = SLI(S, id);
                    s_end = SLI(SEND, id);
                    xy = SLI(XY, id);
                    pixel = (unsigned)((xy >> 16) * cam.width + (xy & 0xffff));
                }
            }
            if (MODE == MODE_TRACE) {
                if (alive) {  // a traced ray came back: write its rl_hit (t in the caller's units: the ray ran with |d| = 1)
                    rl_hit out;
                    const int ref = SLI(HREF, id);
                    out.node = -1;
                    out.t = SL(HT, id) * SL(THRX, id);
                    out.u = HAS_B ? SL(HB1, id) : 0.0f;
                    out.v = HAS_B ? SL(HB2, id) : 0.0f;
                    if (ref >= 0) {
                        const int type = ref_type(ref), idx = ref_index(ref);
                        out.node = type == REF_SPHERE ? sc.sphere_node[idx]
                                 : type == REF_QUAD ? sc.quad_node[idx]
                                 : type == REF_TRI ? __float_as_int(sc.tri_verts[idx].p1.w) : -1;
                    }
                    tio.hits[s] = out;
                    alive = false;
                    has_item = false;
                }
            } else {
                if (alive) {  // its traversal finished: hit record + emitted + scatter
                    p.o = f3(SL(OX, id), SL(OY, id), SL(OZ, id));
                    p.d = f3(SL(DX, id), SL(DY, id), SL(DZ, id));
                    p.time = SL(TIME, id);
                    p.self_ref = SLI(SELF, id);
                    p.thr = f3(SL(THRX, id), SL(THRY, id), SL(THRZ, id));
                    p.depth = SLI(DEPTH, id);
                    OwHit h;
                    h.t = SL(HT, id);
                    h.ref = SLI(HREF, id);
                    h.b1 = HAS_B ? SL(HB1, id) : 0.0f;
                    h.b2 = HAS_B ? SL(HB2, id) : 0.0f;
                    const unsigned bounce = (unsigned)(cam.max_depth - p.depth + 1);
                    const uint4 rnd = philox(make_uint4(pixel, (unsigned)(cam.first_sample + s), bounce, 0u), key);
                    float3 rad;
                    if (!ow_shade<COUNT, PRIMS>(sc, cam, p, h, rnd, rad, lc)) {
                        SL(ACCX, id) += rad.x;  // samples are folded in order (camera.rs:174)
                        SL(ACCY, id) += rad.y;
                        SL(ACCZ, id) += rad.z;
                        alive = false;
                        s++;
                    }
                }
                if (has_item && !alive && s == s_end) {
                    // ONE 16-byte store per finished item (over NVLink when the buffer is rank 0's)
                    const size_t idx = ((size_t)SLI(CHUNK, id) * cam.height + (xy >> 16)) * cam.width + (xy & 0xffff);
                    reinterpret_cast<float4*>(partial)[idx] = make_float4(SL(ACCX, id), SL(ACCY, id), SL(ACCZ, id), 0.0f);
                    has_item = false;
                    retired++;
                }
            }
            // ---- slots without an item take one from the CTA's reserved batch ----
            const bool need = id >= 0 && !has_item;
            const unsigned m_need = __ballot_sync(FULL, need);
            bool dead = false;
            if (m_need) {
                const int leader = __ffs(m_need) - 1;
                long long start = 0;
                int take = 0, dry = 0;
                if ((int)lane == leader)
                    take = items_take(ctl.items, __popc(m_need), n_items, queue, sys_queue, qbatch, 32, q_guided, &start, &dry);
                take = __shfl_sync(FULL, take, leader);
                dry = __shfl_sync(FULL, dry, leader);
                start = __shfl_sync(FULL, start, leader);
                if (need) {
                    const int rank_in = __popc(m_need & ((1u << lane) - 1u));
                    if (rank_in < take) {
                        const long long item = start + rank_in;
                        if (MODE == MODE_TRACE) {
                            const rl_ray r = tio.rays[item];
                            p.o = f3(r.origin[0], r.origin[1], r.origin[2]);
                            p.d = f3(r.direction[0], r.direction[1], r.direction[2]);
                            p.time = r.time;
                            p.self_ref = tio.self_refs ? tio.self_refs[item] : -1;
                            p.thr = f3(rsqrtf(dot(p.d, p.d)), 0.0f, 0.0f);  // t(caller) = t(unit direction) / |d|
                            p.depth = 1;
                            s = (int)item;
                            s_end = s + 1;
                            SLI(CHUNK, id) = 0;
                            SLI(SEND, id) = s_end;
                            has_item = true;
                            alive = true;
                        } else {
                            const int j = find_job(jt, item);
                            const rl_job job = jt_job(jt, j);
                            const long long local = item - jt_prefix(jt, j);
                            const int w = job.x1 - job.x0, hgt = job.y1 - job.y0;
                            const long long pp = padded_pixels(w, hgt);
                            const int ck = (int)(local / pp);
                            int px, py;
                            tile_pixel(w, hgt, local - (long long)ck * pp, &px, &py);
                            if (px < w && py < hgt) {
                                const int x = job.x0 + px, y = job.y0 + py, chunk = job.chunk_begin + ck;
                                ow_chunk_range(cam.spp, cam.n_chunks, chunk, &s, &s_end);
                                if (cam.max_depth <= 0) s = s_end;  // depth 0: every sample is black (camera.rs:239-241)
                                xy = (y << 16) | x;
                                pixel = (unsigned)(y * cam.width + x);
                                SLI(XY, id) = xy;
                                SLI(SEND, id) = s_end;
                                SLI(CHUNK, id) = chunk;
                                SL(ACCX, id) = SL(ACCY, id) = SL(ACCZ, id) = 0.0f;
                                has_item = true;
                            } else {
                                requeue = true;  // a padded slot outside the rectangle: nothing to render, try again next round
                                retired++;       // ... but it IS a work item of the job list (rl_ow_job_items counts padded slots)
                            }
                        }
                    } else if (dry) {
                        dead = true;
                    } else {
                        requeue = true;  // the reserved batch ran out under this request: the next round installs a new one
                    }
                    if (!has_item) SLI(CHUNK, id) = -1;
                }
            }
            if (MODE == MODE_RENDER && has_item && !alive && s < s_end) {  // path regeneration
                ow_camera_ray(cam, xy & 0xffff, xy >> 16, (unsigned)(cam.first_sample + s), p);
                alive = true;
            }
            // ---- new ray: unit direction, big list, slot write-back ----
            if (alive) {
                p.d = p.d * rsqrtf(dot(p.d, p.d));
                const float tm = fmaf(1e-5f, max_abs(p.o), 1e-6f);
                TriShear sh;
                sh.k = 0; sh.Sx = sh.Sy = sh.Sz = 0.0f;
                if (PRIMS & PRIMS_TRIS) sh = make_shear(p.o, p.d);
                unsigned rr = 0u;
                if ((PRIMS & PRIMS_MEDIA) && MODE == MODE_RENDER)
                    rr = philox(make_uint4(pixel, (unsigned)(cam.first_sample + s), (unsigned)(cam.max_depth - p.depth + 1), 2u), key).x;
                OwHit h0;
                h0.t = RL_INF; h0.ref = -1; h0.b1 = h0.b2 = 0.0f;
                if (COUNT) lc.rays++;
                for (int k = 0; k < sc.n_big; k++)  // the big list: once per ray, here, with a full batch
                    ow_leaf_test_od<COUNT, PRIMS>(sc, sc.big_refs[k], p.o, p.d, sh, p.time, p.self_ref, tm, h0, lc, rr);
                SL(OX, id) = p.o.x; SL(OY, id) = p.o.y; SL(OZ, id) = p.o.z;
                SL(DX, id) = p.d.x; SL(DY, id) = p.d.y; SL(DZ, id) = p.d.z;
                SL(TIME, id) = p.time;
                SLI(SELF, id) = p.self_ref;
                SL(THRX, id) = p.thr.x; SL(THRY, id) = p.thr.y; SL(THRZ, id) = p.thr.z;
                SLI(DEPTH, id) = p.depth;
                SL(HT, id) = h0.t;
                SLI(HREF, id) = h0.ref;
                if (HAS_B) {
                    SL(HB1, id) = h0.b1;
                    SL(HB2, id) = h0.b2;
                }
                if (PRIMS & PRIMS_MEDIA) SLI(RND, id) = (int)rr;
            }
            if (id >= 0 && !dead) {
                SLI(S, id) = s;
                SLI(FLAGS, id) = alive ? 1 : 0;
                if (!alive) requeue = true;  // an item that is finished (or black) waits for the next round to retire
            }
            __threadfence_block();
            ring_push(ring_r, &ctl.r_tail, __ballot_sync(FULL, alive), alive, id, lane);
            ring_push(ring_d, &ctl.d_tail, __ballot_sync(FULL, requeue && !alive), requeue && !alive, id, lane);
            const unsigned m_dead = __ballot_sync(FULL, dead);
            if (m_dead && lane == 0) atomicSub((int*)&ctl.live, __popc(m_dead));
            spins = 0;
            continue;
        }
        if (n_act == 0) {  // nothing to traverse, nothing to service: done, or the other warps hold every live slot
            if (__shfl_sync(FULL, (ctl.live <= 0 || ctl.abort) ? 1 : 0, 0)) break;  // lane 0 decides for the warp
            __nanosleep(128);
            if (++spins > (1u << 24)) {  // watchdog: report instead of hanging the GPU (~ 2 s of polling)
                ctl.abort = 1;
                lc.overflow++;
                break;
            }
            continue;
        }
        spins = 0;
        // ================= (D) traversal: node steps / leaf rounds until exit_min lanes have finished their rays =================
        const int n_idle0 = 32 - n_act;
        while (true) {
            if (node < 0 && node != TRAV_END) {  // leaf round: every lane parked at a leaf tests it and pops
                ow_leaf_test_od<COUNT, PRIMS>(sc, ~node, o, d, shear, time, self_ref, tmin, hit, lc, ray_rnd);
                node = stack_pop<THREADS>(st, spill);
            }
            const int n_end = __popc(__ballot_sync(FULL, node == TRAV_END));  // idle lanes count as ended
            if (n_end == 32 || n_end - n_idle0 >= exit_min) break;
            if (n_idle0 >= 4 && __shfl_sync(FULL, ctl.r_tail != ctl.r_head ? 1 : 0, 0)) break;  // rays appeared for the idle lanes
            // node steps, until leaf_min lanes have parked (or ended) or none can step
            const int keep = 32 - n_end - leaf_min;
            int n_in;
            do {
#pragma unroll
                for (int k = 0; k < 2; k++) {  // two steps per ballot (measured in round 1: 1 / 2 / 3 / 4 -> 12.44 / 12.11 / 12.39 / 12.63 ms)
                    if (node >= 0) bvh2_step<COUNT, THREADS>(sc.nodes, node, st, spill, inv_d, oi, tmin, hit.t, lc, sc.n_bvh_nodes);
                }
                n_in = __popc(__ballot_sync(FULL, node >= 0));
            } while (n_in > keep && n_in > 0);
        }
    }
#undef SL
#undef SLI
    // completion accounting: items this GPU stored, added to the counter next to the queue (rank 0's over NVLink when the
    // queue is shared), so the owner can tell a finished render from one a peer dropped out of
    if (MODE == MODE_RENDER) {
        for (int off = 16; off > 0; off >>= 1) retired += __shfl_xor_sync(FULL, retired, off);
        if (lane == 0 && retired) {
            if (sys_queue) atomicAdd_system(queue + 1, (unsigned long long)retired);
            else atomicAdd(queue + 1, (unsigned long long)retired);
        }
    }
    lc.flush(counters);
}

// ---- v7: GLOBAL wavefront — the design north_star names, built once for the measured A/B ----------------
// Path state lives in global memory (SoA, [word][slot], S slots), the render alternates two kernels:
//   k_wf_logic  one thread per slot: hit record + emitted + scatter of the ray that came back, retire / refill the item,
//               regenerate the camera ray, big-list test; the slots that have a ray to trace are compacted into
//               `ray_list` (ballot / popc inside the warp, one shared-memory and one global atomic per CTA);
//   k_wf_trace  persistent warps: lanes fetch slot ids from `ray_list` dynamically (warp-aggregated atomic), run the SAME
//               node steps / leaf rounds as the production kernel, write the hit into the slot and fetch again.
// Per-path arithmetic, RNG counters and the order samples are folded in are the production kernel's, so the frame is
// bit-identical; what differs is where the state lives (L2 / HBM instead of registers + shared memory) and that every
// stage starts with full warps.
namespace wf {
enum { OX, OY, OZ, DX, DY, DZ, TIME, SELF, HT, HREF, THRX, THRY, THRZ, DEPTH, XY, S, SEND, CHUNK, ACCX, ACCY, ACCZ, FLAGS, HB1, HB2, RND, N_WORDS };
static_assert(N_WORDS == OW_WF_WORDS, "kernels.h sizes the state buffer");
struct Ctr {
    unsigned long long items;     // next work item
    unsigned long long retired;   // items finished (completion accounting)
    unsigned ray_count[2];        // rays to trace, ping-pong by iteration parity
    unsigned fetch;               // next ray_list entry the trace kernel hands out
    unsigned live;                // slots that are not dead yet
};
}  // namespace wf

template <int PRIMS>
__global__ void __launch_bounds__(256) k_wf_logic(DevScene sc, OwCam cam, JobTable jt, float* __restrict__ partial, float* __restrict__ st,
                                                  int* __restrict__ ray_list, wf::Ctr* __restrict__ ctr, Counters* counters,
                                                  int n_slots, int iter) {
    using namespace wf;
    constexpr bool HAS_B = (PRIMS & (PRIMS_TRIS | PRIMS_QUADS | PRIMS_MEDIA)) != 0;
    LocalCount<false> lc;
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t NS = (size_t)n_slots;
    int* sti = reinterpret_cast<int*>(st);
#define WS(w) st[(size_t)(w) * NS + id]
#define WI(w) sti[(size_t)(w) * NS + id]
    const uint2 key = make_uint2(cam.seed_lo, cam.seed_hi);
    const int par = iter & 1;
    if (id == 0) {  // nobody touches these during this kernel
        ctr->fetch = 0u;
        ctr->ray_count[par ^ 1] = 0u;
    }
    __shared__ int sm_need[8], sm_emit[8];
    __shared__ long long sm_item_base;
    __shared__ unsigned sm_ray_base;
    const bool valid = id < n_slots;
    int flags = valid ? WI(FLAGS) : 2;
    bool alive = (flags & 1) != 0, dead = (flags & 2) != 0;
    bool has_item = valid && !dead && WI(CHUNK) >= 0;
    int s = 0, s_end = 0, xy = 0;
    unsigned pixel = 0;
    Path p;
    p.depth = 0;
    if (has_item) {
        s = WI(S);
        s_end = WI(SEND);
        xy = WI(XY);
        pixel = (unsigned)((xy >> 16) * cam.width + (xy & 0xffff));
    }
    if (alive) {  // its ray came back from the trace kernel
        p.o = f3(WS(OX), WS(OY), WS(OZ));
        p.d = f3(WS(DX), WS(DY), WS(DZ));
        p.time = WS(TIME);
        p.self_ref = WI(SELF);
        p.thr = f3(WS(THRX), WS(THRY), WS(THRZ));
        p.depth = WI(DEPTH);
        OwHit h;
        h.t = WS(HT);
        h.ref = WI(HREF);
        h.b1 = HAS_B ? WS(HB1) : 0.0f;
        h.b2 = HAS_B ? WS(HB2) : 0.0f;
        const unsigned bounce = (unsigned)(cam.max_depth - p.depth + 1);
        const uint4 rnd = philox(make_uint4(pixel, (unsigned)(cam.first_sample + s), bounce, 0u), key);
        float3 rad;
        if (!ow_shade<false, PRIMS>(sc, cam, p, h, rnd, rad, lc)) {
            WS(ACCX) += rad.x;
            WS(ACCY) += rad.y;
            WS(ACCZ) += rad.z;
            alive = false;
            s++;
        }
    }
    unsigned retired = 0;
    if (has_item && !alive && s == s_end) {
        const size_t idx = ((size_t)WI(CHUNK) * cam.height + (xy >> 16)) * cam.width + (xy & 0xffff);
        reinterpret_cast<float4*>(partial)[idx] = make_float4(WS(ACCX), WS(ACCY), WS(ACCZ), 0.0f);
        has_item = false;
        retired++;
    }
    // ---- work items: one global atomic per CTA ----
    const bool need = valid && !dead && !has_item;
    const unsigned m_need = __ballot_sync(FULL, need);
    if (lane == 0) sm_need[wid] = __popc(m_need);
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < 8; w++) { int c = sm_need[w]; sm_need[w] = tot; tot += c; }
        sm_item_base = tot ? (long long)atomicAdd(&ctr->items, (unsigned long long)tot) : 0ll;
    }
    __syncthreads();
    if (need) {
        const long long item = sm_item_base + sm_need[wid] + __popc(m_need & ((1u << lane) - 1u));
        if (item < jt.n_items) {
            const int j = find_job(jt, item);
            const rl_job job = jt_job(jt, j);
            const long long local = item - jt_prefix(jt, j);
            const int w = job.x1 - job.x0, hgt = job.y1 - job.y0;
            const long long pp = padded_pixels(w, hgt);
            const int ck = (int)(local / pp);
            int px, py;
            tile_pixel(w, hgt, local - (long long)ck * pp, &px, &py);
            if (px < w && py < hgt) {
                const int x = job.x0 + px, y = job.y0 + py, chunk = job.chunk_begin + ck;
                ow_chunk_range(cam.spp, cam.n_chunks, chunk, &s, &s_end);
                if (cam.max_depth <= 0) s = s_end;
                xy = (y << 16) | x;
                pixel = (unsigned)(y * cam.width + x);
                WI(XY) = xy;
                WI(SEND) = s_end;
                WI(CHUNK) = chunk;
                WS(ACCX) = WS(ACCY) = WS(ACCZ) = 0.0f;
                has_item = true;
            } else {
                retired++;  // a padded slot: a work item with nothing to render; the slot asks again next iteration
            }
        } else {
            dead = true;  // the queue is dry
        }
        if (!has_item) WI(CHUNK) = -1;
    }
    if (has_item && !alive && s < s_end) {
        ow_camera_ray(cam, xy & 0xffff, xy >> 16, (unsigned)(cam.first_sample + s), p);
        alive = true;
    }
    if (alive) {
        p.d = p.d * rsqrtf(dot(p.d, p.d));
        const float tm = fmaf(1e-5f, max_abs(p.o), 1e-6f);
        TriShear sh;
        sh.k = 0; sh.Sx = sh.Sy = sh.Sz = 0.0f;
        if (PRIMS & PRIMS_TRIS) sh = make_shear(p.o, p.d);
        unsigned rr = 0u;
        if (PRIMS & PRIMS_MEDIA)
            rr = philox(make_uint4(pixel, (unsigned)(cam.first_sample + s), (unsigned)(cam.max_depth - p.depth + 1), 2u), key).x;
        OwHit h0;
        h0.t = RL_INF; h0.ref = -1; h0.b1 = h0.b2 = 0.0f;
        for (int k = 0; k < sc.n_big; k++)
            ow_leaf_test_od<false, PRIMS>(sc, sc.big_refs[k], p.o, p.d, sh, p.time, p.self_ref, tm, h0, lc, rr);
        WS(OX) = p.o.x; WS(OY) = p.o.y; WS(OZ) = p.o.z;
        WS(DX) = p.d.x; WS(DY) = p.d.y; WS(DZ) = p.d.z;
        WS(TIME) = p.time;
        WI(SELF) = p.self_ref;
        WS(THRX) = p.thr.x; WS(THRY) = p.thr.y; WS(THRZ) = p.thr.z;
        WI(DEPTH) = p.depth;
        WS(HT) = h0.t;
        WI(HREF) = h0.ref;
        if (HAS_B) { WS(HB1) = h0.b1; WS(HB2) = h0.b2; }
        if (PRIMS & PRIMS_MEDIA) WI(RND) = (int)rr;
    }
    if (valid) {
        WI(S) = s;
        WI(FLAGS) = (alive ? 1 : 0) | (dead ? 2 : 0);
    }
    // ---- compaction of the slots that have a ray: ballot / popc, one shared and one global atomic per CTA ----
    const unsigned m_emit = __ballot_sync(FULL, alive);
    if (lane == 0) sm_emit[wid] = __popc(m_emit);
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < 8; w++) { int c = sm_emit[w]; sm_emit[w] = tot; tot += c; }
        sm_ray_base = tot ? atomicAdd(&ctr->ray_count[par], (unsigned)tot) : 0u;
    }
    __syncthreads();
    if (alive) ray_list[sm_ray_base + sm_emit[wid] + __popc(m_emit & ((1u << lane) - 1u))] = id;
    for (int off = 16; off > 0; off >>= 1) retired += __shfl_xor_sync(FULL, retired, off);
    if (lane == 0 && retired) atomicAdd(&ctr->retired, (unsigned long long)retired);
    lc.flush(counters);
#undef WS
#undef WI
}

template <int PRIMS>
__global__ void __launch_bounds__(256, 4) k_wf_trace(DevScene sc, float* __restrict__ st, const int* __restrict__ ray_list,
                                                     wf::Ctr* __restrict__ ctr, Counters* counters, int n_slots, int iter,
                                                     int exit_min, int leaf_min) {
    using namespace wf;
    constexpr bool HAS_B = (PRIMS & (PRIMS_TRIS | PRIMS_QUADS | PRIMS_MEDIA)) != 0;
    LocalCount<false> lc;
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31;
    const size_t NS = (size_t)n_slots;
    int* sti = reinterpret_cast<int*>(st);
    __shared__ int sm_stack[STACK_SM * 256];
    TravStack stk;
    stack_init<256>(stk, sm_stack + threadIdx.x);
    StackSpill spill;
    const unsigned n_rays = ctr->ray_count[iter & 1];
    int slot = -1, node = TRAV_END;
    float3 o = f3(0.0f, 0.0f, 0.0f), d = f3(0.0f, 0.0f, 1.0f), inv_d = f3(1.0f, 1.0f, 1.0f), oi = f3(0.0f, 0.0f, 0.0f);
    float tmin = 0.0f, time = 0.0f;
    int self_ref = -1;
    TriShear shear;
    shear.k = 0; shear.Sx = shear.Sy = shear.Sz = 0.0f;
    unsigned ray_rnd = 0u;
    OwHit hit;
    hit.t = RL_INF; hit.ref = -1; hit.b1 = hit.b2 = 0.0f;
    bool dry = false;
    while (true) {
        // lanes whose ray is finished: hit into the slot
        if (slot >= 0 && node == TRAV_END) {
            st[(size_t)HT * NS + slot] = hit.t;
            sti[(size_t)HREF * NS + slot] = hit.ref;
            if (HAS_B) {
                st[(size_t)HB1 * NS + slot] = hit.b1;
                st[(size_t)HB2 * NS + slot] = hit.b2;
            }
            slot = -1;
        }
        // idle lanes fetch the next ray (warp-aggregated)
        const unsigned m_idle = __ballot_sync(FULL, slot < 0);
        if (m_idle && !dry) {
            const int want = __popc(m_idle), leader = __ffs(m_idle) - 1;
            unsigned base = 0;
            if ((int)lane == leader) base = atomicAdd(&ctr->fetch, (unsigned)want);
            base = __shfl_sync(FULL, base, leader);
            if (base + (unsigned)want >= n_rays) dry = true;  // warp-uniform: nothing (more) beyond this batch
            if (slot < 0) {
                const unsigned k = base + (unsigned)__popc(m_idle & ((1u << lane) - 1u));
                if (k < n_rays) {
                    slot = ray_list[k];
                    o = f3(st[(size_t)OX * NS + slot], st[(size_t)OY * NS + slot], st[(size_t)OZ * NS + slot]);
                    d = f3(st[(size_t)DX * NS + slot], st[(size_t)DY * NS + slot], st[(size_t)DZ * NS + slot]);
                    time = st[(size_t)TIME * NS + slot];
                    self_ref = sti[(size_t)SELF * NS + slot];
                    hit.t = st[(size_t)HT * NS + slot];
                    hit.ref = sti[(size_t)HREF * NS + slot];
                    if (HAS_B) {
                        hit.b1 = st[(size_t)HB1 * NS + slot];
                        hit.b2 = st[(size_t)HB2 * NS + slot];
                    }
                    if (PRIMS & PRIMS_MEDIA) ray_rnd = (unsigned)sti[(size_t)RND * NS + slot];
                    inv_d = safe_inv_fast(d);
                    oi = o * inv_d;
                    tmin = fmaf(1e-5f, max_abs(o), 1e-6f);
                    if (PRIMS & PRIMS_TRIS) shear = make_shear(o, d);
                    stack_reset(stk);
                    node = sc.n_bvh_prims > 0 ? 0 : TRAV_END;
                }
            }
        }
        const int n_act = __popc(__ballot_sync(FULL, slot >= 0));
        if (n_act == 0) break;  // dry and nothing in flight
        const int n_idle0 = 32 - n_act;
        while (true) {
            if (node < 0 && node != TRAV_END) {
                ow_leaf_test_od<false, PRIMS>(sc, ~node, o, d, shear, time, self_ref, tmin, hit, lc, ray_rnd);
                node = stack_pop<256>(stk, spill);
            }
            const int n_end = __popc(__ballot_sync(FULL, node == TRAV_END));
            if (n_end == 32 || (!dry && n_end - n_idle0 >= exit_min)) break;
            const int keep = 32 - n_end - leaf_min;
            int n_in;
            do {
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    if (node >= 0) bvh2_step<false, 256>(sc.nodes, node, stk, spill, inv_d, oi, tmin, hit.t, lc, sc.n_bvh_nodes);
                }
                n_in = __popc(__ballot_sync(FULL, node >= 0));
            } while (n_in > keep && n_in > 0);
        }
    }
    lc.flush(counters);
}

// fold the per-chunk partial sums ([chunk][pixel] float4) in chunk order into [pixel][3]
__global__ void k_ow_reduce(const float4* __restrict__ partial, float* __restrict__ out, size_t n_pixels, int n_chunks) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pixels) return;
    float r = 0.0f, g = 0.0f, b = 0.0f;
    for (int c = 0; c < n_chunks; c++) {
        float4 v = partial[(size_t)c * n_pixels + i];
        r += v.x;
        g += v.y;
        b += v.z;
    }
    out[3 * i + 0] = r;
    out[3 * i + 1] = g;
    out[3 * i + 2] = b;
}

// ---- 8-bit output encoders ------------------------------------------------------------------------------------
// f64 on purpose: these must agree bit for bit with the reference's encoders applied to the f32 framebuffer.
__global__ void k_encode_rtc_u8(const float* __restrict__ rgb, uint8_t* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = (double)rgb[i] * 255.0;  // canvas.rs:53-56: `(color * 255).round() as i32`, clamp(0, 255)
    double r = round(v);                // half away from zero, like f64::round
    out[i] = (uint8_t)(r != r ? 0.0 : fmin(fmax(r, 0.0), 255.0));  // `NaN as i32` is 0
}
__global__ void k_encode_ow_u8(const float* __restrict__ sum, uint8_t* __restrict__ out, size_t n, double inv_samples) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = (double)sum[i] * inv_samples;  // `c / samples` is `c * (1.0 / samples)` (vec3.rs:178-184)
    double s = v <= 0.0031308 ? 12.92 * v : 1.055 * pow(v, 1.0 / 2.4) - 0.055;  // color.rs:130-136
    double f = floor(s * 255.999);                                              // color.rs:47-57, `as i16` saturates
    out[i] = (uint8_t)(f != f ? 0.0 : fmin(fmax(f, 0.0), 255.0));
}

}  // namespace

// Samples are cut into chunks of >= 8 (at most 64 chunks) plus the small tail chunks of ow_chunk_range (kernels.h): the
// chunk is the unit of work a path slot owns, so it bounds both the load-balancing tail (a 32-sample item was 1.7 ms of
// lane time, 10 % of an 8-GPU cover-scene step) and the size of the partial-sum buffer (<= 64 frames).  A function of
// spp ALONE: every rank of a multi-GPU render computes the same partition (round 1 read two environment variables here).
namespace {
// OwTriPlane from the f32 vertices the scene holds (flat/plane.rs:23-40: n = u x v, normal = n / |n|, d = normal . q,
// w = n / (n . n); alpha = w . ((p - q) x v) = (v x w) . p - (v x w) . q and beta = w . (u x (p - q)) = (w x u) . p - (w x u) . q by
// the cyclic symmetry of the triple product).  f64, roundings pinned, one thread per triangle.  A degenerate triangle
// (u parallel to v: `Plane::new` panics in the reference) gets NaN vectors and is never hit.
__global__ void k_tri_planes(const TriVerts* __restrict__ tv, int n, OwTriPlane* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p0 = tv[i].p0, p1 = tv[i].p1, p2 = tv[i].p2;
    const double Q[3] = {p0.x, p0.y, p0.z};
    const double U[3] = {__dsub_rn(p1.x, p0.x), __dsub_rn(p1.y, p0.y), __dsub_rn(p1.z, p0.z)};
    const double V[3] = {__dsub_rn(p2.x, p0.x), __dsub_rn(p2.y, p0.y), __dsub_rn(p2.z, p0.z)};
    auto cross = [](const double* a, const double* b, double* c) {
        c[0] = __dsub_rn(__dmul_rn(a[1], b[2]), __dmul_rn(a[2], b[1]));
        c[1] = __dsub_rn(__dmul_rn(a[2], b[0]), __dmul_rn(a[0], b[2]));
        c[2] = __dsub_rn(__dmul_rn(a[0], b[1]), __dmul_rn(a[1], b[0]));
    };
    auto dot3 = [](const double* a, const double* b) {
        return __dadd_rn(__dadd_rn(__dmul_rn(a[0], b[0]), __dmul_rn(a[1], b[1])), __dmul_rn(a[2], b[2]));
    };
    double N[3], W[3], A[3], B[3], un[3];
    cross(U, V, N);
    const double nn = dot3(N, N), len = __dsqrt_rn(nn);
    for (int k = 0; k < 3; k++) { un[k] = __ddiv_rn(N[k], len); W[k] = __ddiv_rn(N[k], nn); }
    cross(V, W, A);
    cross(W, U, B);
    OwTriPlane o;
    o.n = make_float4((float)un[0], (float)un[1], (float)un[2], (float)dot3(un, Q));
    o.a = make_float4((float)A[0], (float)A[1], (float)A[2], (float)dot3(A, Q));
    o.b = make_float4((float)B[0], (float)B[1], (float)B[2], (float)dot3(B, Q));
    out[i] = o;
}

}  // namespace

cudaError_t launch_tri_planes(const TriVerts* d_verts, int n, OwTriPlane* d_planes, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    k_tri_planes<<<(n + 127) / 128, 128, 0, stream>>>(d_verts, n, d_planes);
    return cudaGetLastError();
}

#ifndef RL_OW_MIN_CHUNK  // overridable for tools/build_alt.py experiments only
#define RL_OW_MIN_CHUNK 8
#define RL_OW_MAX_CHUNKS 64
#endif
int ow_num_chunks(int spp) {
    if (spp <= 0) return 1;
    constexpr int min_chunk = RL_OW_MIN_CHUNK, max_chunks = RL_OW_MAX_CHUNKS;
    const int t = ow_tail_chunks(spp), body = spp - t * OW_TAIL_SIZE;
    int per_chunk = (body + (max_chunks - t) - 1) / (max_chunks - t);
    if (per_chunk < min_chunk) per_chunk = min_chunk;
    return (body + per_chunk - 1) / per_chunk + t;
}

int ow_image_height(const rl_ow_camera* c) {
    double h = (double)c->image_width / c->aspect_ratio;
    long long hh = h > 0 ? (long long)h : 0;
    return (int)(hh < 1 ? 1 : hh);
}

static OwCam make_cam(const rl_ow_camera* p, uint32_t first_sample) {
    // Camera::new (camera.rs:72-118) in f64 on the host
    OwCam c{};
    c.width = p->image_width;
    c.height = ow_image_height(p);
    c.spp = p->samples_per_pixel;
    c.max_depth = p->max_depth;
    c.first_sample = (int)first_sample;
    c.n_chunks = ow_num_chunks(p->samples_per_pixel);
    c.defocus = p->defocus_angle > 0.0 ? 1 : 0;
    const double PI = 3.14159265358979323846;
    auto sub = [](const double* a, const double* b, double* o) { for (int k = 0; k < 3; k++) o[k] = a[k] - b[k]; };
    auto norm = [](double* v) { double m = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); for (int k = 0; k < 3; k++) v[k] /= m; };
    auto crs = [](const double* a, const double* b, double* o) {
        o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
    };
    double theta = p->vfov * PI / 180.0;
    double h = tan(theta / 2.0);
    double vh = 2.0 * h * p->focus_dist;
    double vw = vh * ((double)c.width / (double)c.height);
    double w[3], u[3], v[3];
    sub(p->lookfrom, p->lookat, w);
    norm(w);
    crs(p->vup, w, u);
    norm(u);
    crs(w, u, v);
    norm(v);
    double du[3], dv[3], p00[3];
    for (int k = 0; k < 3; k++) {
        double vu = vw * u[k], vv = vh * -v[k];
        du[k] = vu / c.width;
        dv[k] = vv / c.height;
        double ul = p->lookfrom[k] - p->focus_dist * w[k] - vu / 2.0 - vv / 2.0;
        p00[k] = ul + 0.5 * (du[k] + dv[k]);
    }
    double dr = p->focus_dist * tan((p->defocus_angle / 2.0) * PI / 180.0);
    c.lookfrom = make_float3((float)p->lookfrom[0], (float)p->lookfrom[1], (float)p->lookfrom[2]);
    c.pixel00 = make_float3((float)p00[0], (float)p00[1], (float)p00[2]);
    c.du = make_float3((float)du[0], (float)du[1], (float)du[2]);
    c.dv = make_float3((float)dv[0], (float)dv[1], (float)dv[2]);
    c.disk_u = make_float3((float)(u[0] * dr), (float)(u[1] * dr), (float)(u[2] * dr));
    c.disk_v = make_float3((float)(v[0] * dr), (float)(v[1] * dr), (float)(v[2] * dr));
    c.background = make_float3((float)p->background[0], (float)p->background[1], (float)p->background[2]);
    c.seed_lo = (unsigned)(p->seed & 0xffffffffu);
    c.seed_hi = (unsigned)(p->seed >> 32);
    return c;
}


namespace {
constexpr int PRIMS_FLAT = PRIMS_TRIS | PRIMS_QUADS;
typedef void (*K5)(DevScene, OwCam, JobTable, float*, unsigned long long*, Counters*, int, int, int, long long, int, int, TraceIO);
typedef void (*K6)(DevScene, OwCam, JobTable, float*, unsigned long long*, Counters*, int, int, long long, int, int, int, int,
                   TraceIO);

// which instantiation a scene runs: spheres only / no spheres / every surface / everything (media, Noise)
int scene_prims(const DevScene& sc) {
    if (sc.n_media > 0 || sc.n_perlins > 0) return PRIMS_FULL;
    if (sc.n_tris == 0 && sc.n_quads == 0 && sc.n_images == 0) return PRIMS_SPHERES;
    if (sc.n_spheres == 0) return PRIMS_FLAT;
    return PRIMS_ALL;
}

template <int MODE>
K6 pick_k6(int prims, int minb, bool instrumented) {
#define RL_K6(P_)                                                                                                   \
    (minb >= 4 ? (instrumented ? (K6)k_ow_render6<true, 4, P_, MODE> : (K6)k_ow_render6<false, 4, P_, MODE>)       \
               : (instrumented ? (K6)k_ow_render6<true, 3, P_, MODE> : (K6)k_ow_render6<false, 3, P_, MODE>))
    switch (prims) {
        case PRIMS_SPHERES: return RL_K6(PRIMS_SPHERES);
        case PRIMS_FLAT: return RL_K6(PRIMS_FLAT);
        case PRIMS_ALL: return RL_K6(PRIMS_ALL);
        default: return instrumented ? (K6)k_ow_render6<true, 3, PRIMS_FULL, MODE> : (K6)k_ow_render6<false, 3, PRIMS_FULL, MODE>;
    }
#undef RL_K6
}

#ifndef RL_K5_HI  // resident CTAs per SM the production kernel is compiled for (experiment builds: 5 = 48 registers)
#define RL_K5_HI 4
#endif
template <bool TRACE>
K5 pick_k5(int prims, int minb, bool instrumented) {
#define RL_K5(P_)                                                                                                     \
    (minb >= 4 ? (instrumented ? (K5)k_ow_render5<true, RL_K5_HI, P_, TRACE> : (K5)k_ow_render5<false, RL_K5_HI, P_, TRACE>) \
               : (instrumented ? (K5)k_ow_render5<true, 3, P_, TRACE> : (K5)k_ow_render5<false, 3, P_, TRACE>))
    switch (prims) {
        case PRIMS_SPHERES: return RL_K5(PRIMS_SPHERES);
        case PRIMS_FLAT: return RL_K5(PRIMS_FLAT);
        case PRIMS_ALL: return RL_K5(PRIMS_ALL);
        default: return instrumented ? (K5)k_ow_render5<true, 3, PRIMS_FULL, TRACE> : (K5)k_ow_render5<false, 3, PRIMS_FULL, TRACE>;
    }
#undef RL_K5
}

// one launch of the production kernel (render or trace): occupancy, grid, batch sizes, thresholds
cudaError_t launch_k5(K5 k5, int prims, const DevScene& sc, const OwCam& c, const JobTable& jt, float* d_partial,
                      unsigned long long* d_queue, Counters* d_counters, int sm_count, cudaStream_t stream, bool shared_queue,
                      const OwTuning& tune, const TraceIO& tio) {
    // service threshold: sphere scenes shade cheaply and prefer fuller service rounds; parked lanes per leaf round
    // (gpurun_out sweeps of round 2, cover scene / Cornell box + spot: 24 / 8 -> 96.3 ms, 20 / 8 -> 55.9 ms)
    const int svc_min = tune.svc_min > 0 ? tune.svc_min : (prims == PRIMS_SPHERES ? 24 : 20);
    const int leaf_min = tune.leaf_min > 0 ? tune.leaf_min : 8;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k5, 256, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (tune.ctas_per_sm > 0 && per_sm > tune.ctas_per_sm) per_sm = tune.ctas_per_sm;
    long long want = (jt.n_items + 255) / 256;
    long long grid = (long long)sm_count * per_sm;  // persistent: every SM full, a multiple of the SM count
    if (grid > want && !shared_queue) grid = want;
    if (grid < 1) grid = 1;
    // a warp takes up to 64 items at a time from its CTA's reserve; the reserve refills from the global queue 512 at a
    // time (ONE atomic on the global counter per 512 items), never so many that CTAs starve; below `q_guided` remaining
    // items the refills shrink to one warp batch (a shared queue feeds up to 8 GPUs)
    long long per_warp = jt.n_items / (grid * 8 * 4);
    const int wbatch = (int)(per_warp < 32 ? 32 : (per_warp > 64 ? 64 : per_warp));
    long long per_cta = jt.n_items / (grid * 4);
    const int qbatch = (int)(per_cta < wbatch ? wbatch : (per_cta > 512 ? 512 : per_cta));
    const long long q_guided = grid * (long long)qbatch * 2 * (shared_queue ? 8 : 1);
    k5<<<(unsigned)grid, 256, 0, stream>>>(sc, c, jt, d_partial, d_queue, d_counters, shared_queue ? 1 : 0, qbatch, wbatch, q_guided,
                                           svc_min, leaf_min, tio);
    return cudaGetLastError();
}

// one v6 launch (render or trace): shared-memory size, occupancy, grid, batch sizes
cudaError_t launch_k6(K6 k, int prims, const DevScene& sc, const OwCam& c, const JobTable& jt, float* d_partial,
                      unsigned long long* d_queue, Counters* d_counters, int sm_count, cudaStream_t stream, bool shared_queue,
                      const OwTuning& tune, const TraceIO& tio) {
    int P = tune.slots > 0 ? tune.slots : 384;
    P = (P + 31) / 32 * 32;
    if (P < 256) P = 256;  // at least one slot per lane, or lanes starve by construction
    if (P > v6::QCAP) P = v6::QCAP;
    const size_t smem = v6::smem_bytes(prims, P);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, 256, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (tune.ctas_per_sm > 0 && per_sm > tune.ctas_per_sm) per_sm = tune.ctas_per_sm;
    long long want = (jt.n_items + 255) / 256;
    long long grid = (long long)sm_count * per_sm;  // persistent: every SM full, a multiple of the SM count
    if (grid > want && !shared_queue) grid = want;
    if (grid < 1) grid = 1;
    // items a CTA reserves per global atomic: plenty while there is plenty of work, never so many that CTAs starve
    long long per_cta = jt.n_items / (grid * 8);
    const int qbatch = (int)(per_cta < 32 ? 32 : (per_cta > 256 ? 256 : per_cta));
    // below this many remaining items a CTA reserves 32 instead of qbatch (a shared queue feeds up to 8 GPUs)
    const long long q_guided = grid * (long long)qbatch * 2 * (shared_queue ? 8 : 1);
    const int svc_lo = tune.svc_lo > 0 ? tune.svc_lo : 16;
    const int exit_min = tune.exit_min > 0 ? tune.exit_min : 8;
    const int leaf_min = tune.leaf_min > 0 ? tune.leaf_min : (prims == PRIMS_SPHERES ? 8 : 12);
    k<<<(unsigned)grid, 256, smem, stream>>>(sc, c, jt, d_partial, d_queue, d_counters, shared_queue ? 1 : 0, qbatch, q_guided, P,
                                             svc_lo, exit_min, leaf_min, tio);
    return cudaGetLastError();
}
}  // namespace

cudaError_t launch_ow_render(const DevScene& sc, const rl_ow_camera* cam, uint32_t first_sample, const JobTable& jt,
                             float* d_partial, unsigned long long* d_queue, Counters* d_counters, bool instrumented,
                             int sm_count, cudaStream_t stream, bool shared_queue, const OwTuning& tune) {
    if (jt.n_items <= 0) return cudaSuccess;
    OwCam c = make_cam(cam, first_sample);
    cudaError_t e = cudaSuccess;
    if (!shared_queue) e = cudaMemsetAsync(d_queue, 0, 2 * sizeof(unsigned long long), stream);  // the owner resets a shared queue
    if (e != cudaSuccess) return e;
    const int prims = scene_prims(sc);
    // resident CTAs per SM the kernel is compiled for (register budget): the spheres-only and no-spheres builds fit 64
    // registers (4 CTAs/SM); the builds that carry every primitive kind are better at 80 (3 CTAs/SM)
    const int minb = tune.minb ? tune.minb : ((prims == PRIMS_SPHERES || prims == PRIMS_FLAT) ? 4 : 3);
    TraceIO tio{nullptr, nullptr, nullptr};
    if (tune.variant == 6)  // the pooled-paths experiment (measured 1.6-1.9x SLOWER; kept selectable for the A/B)
        return launch_k6(pick_k6<v6::MODE_RENDER>(prims, minb, instrumented), prims, sc, c, jt, d_partial, d_queue, d_counters,
                         sm_count, stream, shared_queue, tune, tio);
    return launch_k5(pick_k5<false>(prims, minb, instrumented), prims, sc, c, jt, d_partial, d_queue, d_counters, sm_count, stream,
                     shared_queue, tune, tio);
}

// The global-wavefront render (ow.variant = 7, single GPU): alternate k_wf_logic / k_wf_trace until no slot has a ray left
// and the item queue is dry; the host looks at the counters every 16 iterations.
cudaError_t launch_ow_wavefront(const DevScene& sc, const rl_ow_camera* cam, uint32_t first_sample, const JobTable& jt,
                                float* d_partial, unsigned long long* d_queue, Counters* d_counters, const WavefrontBuffers& wb,
                                int sm_count, cudaStream_t stream, const OwTuning& tune, int* launches) {
    using namespace wf;
    if (jt.n_items <= 0) return cudaSuccess;
    const OwCam c = make_cam(cam, first_sample);
    const int prims = scene_prims(sc);
    const int S = wb.slots;
    Ctr* ctr = reinterpret_cast<Ctr*>(wb.ctr);
    cudaError_t e = cudaMemsetAsync(ctr, 0, sizeof(Ctr), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(wb.state + (size_t)FLAGS * S, 0, sizeof(float) * (size_t)S, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(wb.state + (size_t)CHUNK * S, 0xFF, sizeof(float) * (size_t)S, stream);  // -1: no item
    if (e != cudaSuccess) return e;
    typedef void (*KL)(DevScene, OwCam, JobTable, float*, float*, int*, Ctr*, Counters*, int, int);
    typedef void (*KT)(DevScene, float*, const int*, Ctr*, Counters*, int, int, int, int);
    KL kl;
    KT kt;
    switch (prims) {
        case PRIMS_SPHERES: kl = k_wf_logic<PRIMS_SPHERES>; kt = k_wf_trace<PRIMS_SPHERES>; break;
        case PRIMS_FLAT: kl = k_wf_logic<PRIMS_FLAT>; kt = k_wf_trace<PRIMS_FLAT>; break;
        case PRIMS_ALL: kl = k_wf_logic<PRIMS_ALL>; kt = k_wf_trace<PRIMS_ALL>; break;
        default: kl = k_wf_logic<PRIMS_FULL>; kt = k_wf_trace<PRIMS_FULL>; break;
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kt, 256, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const unsigned tgrid = (unsigned)(sm_count * per_sm), lgrid = (unsigned)((S + 255) / 256);
    const int exit_min = tune.exit_min > 0 ? tune.exit_min : 8;
    const int leaf_min = tune.leaf_min > 0 ? tune.leaf_min : 8;
    int iter = 0, n_launch = 0;
    for (;;) {
        for (int k = 0; k < 16; k++, iter++) {
            kl<<<lgrid, 256, 0, stream>>>(sc, c, jt, d_partial, wb.state, wb.ray_list, ctr, d_counters, S, iter);
            kt<<<tgrid, 256, 0, stream>>>(sc, wb.state, wb.ray_list, ctr, d_counters, S, iter, exit_min, leaf_min);
            n_launch += 2;
        }
        Ctr h;
        e = cudaMemcpyAsync(&h, ctr, sizeof(h), cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return e;
        if (h.ray_count[(iter - 1) & 1] == 0u && (long long)h.items >= jt.n_items) break;
        if (iter > (1 << 22)) return cudaErrorLaunchTimeout;  // cannot happen: every iteration retires rays
    }
    e = cudaMemcpyAsync(d_queue + 1, &ctr->retired, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, stream);  // completion accounting
    if (launches) *launches = n_launch;
    return e == cudaSuccess ? cudaGetLastError() : e;
}

cudaError_t launch_ow_reduce(const rl_ow_camera* cam, const float* d_partial, float* d_out, cudaStream_t stream) {
    size_t n = (size_t)cam->image_width * ow_image_height(cam);
    int nc = ow_num_chunks(cam->samples_per_pixel);
    k_ow_reduce<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const float4*>(d_partial), d_out, n, nc);
    return cudaGetLastError();
}

cudaError_t launch_encode_rtc_u8(const float* d_rgb, uint8_t* d_out, size_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    k_encode_rtc_u8<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_rgb, d_out, n);
    return cudaGetLastError();
}

cudaError_t launch_encode_ow_u8(const float* d_rgb_sum, uint8_t* d_out, size_t n, int samples, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    k_encode_ow_u8<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_rgb_sum, d_out, n, 1.0 / (double)samples);
    return cudaGetLastError();
}

// rl_trace_batch for OW scenes: the rays run through the RENDER kernel itself (TRACE instantiation of the variant the
// ctx renders with): same work queue, service / refill rounds, unit directions, t_min rule, start-on-surface rule, big
// list at ray start, node steps and leaf rounds.  d_self_refs (optional) = the leaf ref each ray starts on.
cudaError_t launch_ow_trace(const DevScene& sc, const rl_ray* d_rays, const int* d_self_refs, uint64_t n, rl_hit* d_hits,
                            unsigned long long* d_queue, Counters* d_counters, bool instrumented, int sm_count,
                            cudaStream_t stream, const OwTuning& tune) {
    if (n == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(d_queue, 0, 2 * sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    OwCam c{};
    c.width = 1;
    c.max_depth = 1;
    JobTable jt{};
    jt.inline_jobs = 1;
    jt.n_items = (long long)n;
    const int prims = scene_prims(sc);
    const int minb = tune.minb ? tune.minb : ((prims == PRIMS_SPHERES || prims == PRIMS_FLAT) ? 4 : 3);
    TraceIO tio{d_rays, d_self_refs, d_hits};
    if (tune.variant == 6)
        return launch_k6(pick_k6<v6::MODE_TRACE>(prims, minb, instrumented), prims, sc, c, jt, nullptr, d_queue, d_counters, sm_count,
                         stream, false, tune, tio);
    return launch_k5(pick_k5<true>(prims, minb, instrumented), prims, sc, c, jt, nullptr, d_queue, d_counters, sm_count, stream, false,
                     tune, tio);
}

}  // namespace rl
