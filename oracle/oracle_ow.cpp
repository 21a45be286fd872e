// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement (C++17, f64) of the ray-tracing-one-weekend (OW) per-pixel ray loop of
// marcantony/rendering-learning, including the un-vendored RNG stack it depends on:
//   rand_core 0.6.4  SeedableRng::seed_from_u64 (PCG32 seed expansion), BlockRng u32->u64 pairing
//   rand_chacha 0.3.1 ChaCha8Rng (8 rounds, 64-bit counter + 64-bit stream, 4-block buffer,
//                     set_stream keeps the word position)
//   rand 0.8.5       Standard f64 = (u64 >> 11) * 2^-53;  Uniform<f64>(-1,1) = ([1,2) mantissa trick - 1)*2 - 1
//   rand_distr 0.4.3 UnitSphere (Marsaglia), UnitDisc (rejection)
// Those crates' sources are NOT under /root/reference (Cargo.lock only); the recipes are restated from
// their published algorithms and pinned by reproducing OW/tests/expectations/test.ppm byte for byte
// (tests/test_oracle_golden.py) — that image exercises Lambertian, fuzzy Metal, Dielectric and the
// defocus disk, i.e. Standard f64, UnitSphere and UnitDisc.
// PARITY UNPINNED for the parts no reference test renders: Perlin / Noise (perlin.rs, texture.rs:84-94), ConstantMedium
// and Isotropic (constant_medium.rs, material.rs:197-221).  They are pinned by tests/test_host_logic.py instead: a host
// mirror of Perlin::noise agrees to 1e-12, and a unit slab of density 1.5 transmits exp(-1.5) (Beer-Lambert).
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/ray-tracing-one-weekend/src/).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>

#include "../include/rl_b200.h"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr double INF = std::numeric_limits<double>::infinity();
constexpr double PI = 3.14159265358979323846264338327950288;

// ------------------------------------------------------------------------------------------------
// RNG stack
// ------------------------------------------------------------------------------------------------
inline uint32_t rotl32(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }

struct ChaCha8Rng {
    uint32_t key[8];
    uint64_t block_pos = 0;  // counter of the NEXT block to generate (words 12-13)
    uint64_t stream = 0;     // words 14-15
    uint32_t results[64];    // 4 blocks, like BlockRng<ChaCha8Core>
    int index = 64;          // 64 == empty

    // rand_core 0.6.4 SeedableRng::seed_from_u64 (OW/src/camera.rs:161)
    explicit ChaCha8Rng(uint64_t seed) {
        uint64_t state = seed;
        for (int i = 0; i < 8; i++) {
            state = state * 6364136223846793005ULL + 11634580027462260723ULL;
            uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
            uint32_t rot = (uint32_t)(state >> 59);
            key[i] = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));  // little-endian word
        }
    }

    static void block(const uint32_t key[8], uint64_t counter, uint64_t stream, uint32_t out[16]) {
        uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                          key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                          (uint32_t)counter, (uint32_t)(counter >> 32),
                          (uint32_t)stream, (uint32_t)(stream >> 32)};
        uint32_t x[16];
        std::memcpy(x, s, sizeof(x));
#define QR(a, b, c, d)                                  \
    x[a] += x[b]; x[d] ^= x[a]; x[d] = rotl32(x[d], 16); \
    x[c] += x[d]; x[b] ^= x[c]; x[b] = rotl32(x[b], 12); \
    x[a] += x[b]; x[d] ^= x[a]; x[d] = rotl32(x[d], 8);  \
    x[c] += x[d]; x[b] ^= x[c]; x[b] = rotl32(x[b], 7);
        for (int r = 0; r < 4; r++) {  // 8 rounds = 4 double rounds
            QR(0, 4, 8, 12) QR(1, 5, 9, 13) QR(2, 6, 10, 14) QR(3, 7, 11, 15)
            QR(0, 5, 10, 15) QR(1, 6, 11, 12) QR(2, 7, 8, 13) QR(3, 4, 9, 14)
        }
#undef QR
        for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
    }

    void generate() {  // ChaCha8Core::generate — 4 consecutive blocks, counter += 4
        for (int b = 0; b < 4; b++) block(key, block_pos + (uint64_t)b, stream, results + 16 * b);
        block_pos += 4;
    }
    void generate_and_set(int idx) {
        generate();
        index = idx;
    }
    // ChaCha8Rng::set_stream (rand_chacha 0.3.1): new nonce; if the buffer is live, regenerate it at
    // the same word position (get_word_pos -> set_word_pos)
    void set_stream(uint64_t s) {
        stream = s;
        if (index != 64) {
            uint64_t buf_start_block = block_pos - 4;
            uint64_t pos_block = buf_start_block + (uint64_t)(index / 16);
            int words = index % 16;
            block_pos = pos_block;
            generate_and_set(words);
        }
    }
    // BlockRng::next_u64
    uint64_t next_u64() {
        const int len = 64;
        if (index < len - 1) {
            uint64_t v = ((uint64_t)results[index + 1] << 32) | results[index];
            index += 2;
            return v;
        } else if (index >= len) {
            generate_and_set(2);
            return ((uint64_t)results[1] << 32) | results[0];
        } else {
            uint64_t x = results[len - 1];
            generate_and_set(1);
            uint64_t y = results[0];
            return (y << 32) | x;
        }
    }
    // rand 0.8.5 Standard for f64
    double gen_f64() { return (double)(next_u64() >> 11) * (1.0 / 9007199254740992.0); }
    // rand 0.8.5 UniformFloat<f64>::new(-1, 1).sample: value1_2 - 1.0, * scale(2) + low(-1)
    double uniform_m1_1() {
        uint64_t bits = (next_u64() >> 12) | 0x3FF0000000000000ULL;
        double v12;
        std::memcpy(&v12, &bits, 8);
        return (v12 - 1.0) * 2.0 + -1.0;
    }
};

// ------------------------------------------------------------------------------------------------
// vec3.rs
// ------------------------------------------------------------------------------------------------
struct V3 {
    double x, y, z;
};
inline V3 v3(double x, double y, double z) { return {x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(double s, V3 a) { return a * s; }                   // vec3.rs:170-176
inline V3 operator/(V3 a, double s) { return a * (1.0 / s); }           // vec3.rs:178-184
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double length_squared(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline double length(V3 a) { return std::sqrt(length_squared(a)); }
inline V3 normalize(V3 a) { return a / length(a); }                     // vec3.rs:56-58
// NormalizedVec3::try_from (vec3.rs:235-248); float_cmp approx_eq(0, eps=1e-16, ulps=0)
inline bool try_normalize(V3 a, V3* out) {
    double m = length_squared(a);
    if (m == 0.0 || std::fabs(m - 0.0) <= 1e-16) return false;
    *out = normalize(a);
    return true;
}
inline bool near_zero(V3 a) {                                           // vec3.rs:60-65
    return std::fabs(a.x) <= 1e-8 && std::fabs(a.y) <= 1e-8 && std::fabs(a.z) <= 1e-8;
}
inline V3 reflect(V3 v, V3 n) { return v - 2.0 * dot(v, n) * n; }       // vec3.rs:67-69
inline V3 refract(V3 uv, V3 n, double etai_over_etat) {                 // vec3.rs:224-230
    double cos_theta = std::fmin(dot(-uv, n), 1.0);
    V3 r_out_perp = etai_over_etat * (uv + cos_theta * n);
    V3 r_out_parallel = -std::sqrt(std::fabs(1.0 - length_squared(r_out_perp))) * n;
    return r_out_perp + r_out_parallel;
}
inline double powi5(double x) { double x2 = x * x; double x4 = x2 * x2; return x * x4; }  // f64::powi(5)

inline V3 random_unit_vector(ChaCha8Rng& rng) {  // vec3.rs:72-75 -> rand_distr::UnitSphere
    for (;;) {
        double x1 = rng.uniform_m1_1();
        double x2 = rng.uniform_m1_1();
        double sum = x1 * x1 + x2 * x2;
        if (sum >= 1.0) continue;
        double factor = 2.0 * std::sqrt(1.0 - sum);
        return {x1 * factor, x2 * factor, 1.0 - 2.0 * sum};
    }
}

struct Ray {  // ray.rs
    V3 origin, direction;
    double time;
    V3 at(double t) const { return origin + direction * t; }
};
struct Interval {  // interval.rs
    double min, max;
    double size() const { return max - min; }
    bool contains(double x) const { return min <= x && x <= max; }
    Interval expand(double delta) const { double p = delta / 2.0; return {min - p, max + p}; }
    Interval merge(const Interval& o) const { return {std::fmin(min, o.min), std::fmax(max, o.max)}; }
};
struct AABB {  // aabb.rs
    Interval x, y, z;
    static AABB make(Interval x, Interval y, Interval z) {  // AABB::new (14-27)
        const double DELTA = 1e-4;
        return {x.size() < DELTA ? x.expand(DELTA) : x, y.size() < DELTA ? y.expand(DELTA) : y,
                z.size() < DELTA ? z.expand(DELTA) : z};
    }
    static AABB from_extrema(V3 a, V3 b) {  // 30-64
        Interval x = a.x <= b.x ? Interval{a.x, b.x} : Interval{b.x, a.x};
        Interval y = a.y <= b.y ? Interval{a.y, b.y} : Interval{b.y, a.y};
        Interval z = a.z <= b.z ? Interval{a.z, b.z} : Interval{b.z, a.z};
        return make(x, y, z);
    }
    static AABB empty() { return {{INF, -INF}, {INF, -INF}, {INF, -INF}}; }
    AABB merge(const AABB& o) const { return {x.merge(o.x), y.merge(o.y), z.merge(o.z)}; }
    static void axis(const Interval& i, double origin, double speed, double* a, double* b) {  // 143-152
        double t0 = (i.min - origin) / speed, t1 = (i.max - origin) / speed;
        if (t0 < t1) { *a = t0; *b = t1; } else { *a = t1; *b = t0; }
    }
    bool hit(const Ray& r, const Interval& rt) const {  // 123-132
        double x0, x1, y0, y1, z0, z1;
        axis(x, r.origin.x, r.direction.x, &x0, &x1);
        axis(y, r.origin.y, r.direction.y, &y0, &y1);
        axis(z, r.origin.z, r.direction.z, &z0, &z1);
        double tmin = std::fmax(std::fmax(std::fmax(x0, y0), z0), rt.min);
        double tmax = std::fmin(std::fmin(std::fmin(x1, y1), z1), rt.max);
        return tmin < tmax;
    }
};

struct HitRecord {  // hittable/mod.rs:24-30
    V3 p, normal;
    double t, u, v;
    bool front;
    int material;
    int node;
};

inline void face_normal(const Ray& r, V3 outward, V3* n, bool* front) {  // hittable/mod.rs:32-38
    if (dot(r.direction, outward) <= 0.0) { *n = outward; *front = true; }
    else { *n = -outward; *front = false; }
}

struct M3 { double m[3][3]; };
inline V3 mul(const M3& a, V3 v) {  // matrix.rs:42-60
    double d[3] = {v.x, v.y, v.z}, o[3];
    for (int n = 0; n < 3; n++) {
        double sum = 0.0;
        for (int k = 0; k < 3; k++) sum += a.m[n][k] * d[k];
        o[n] = sum;
    }
    return {o[0], o[1], o[2]};
}

// ConstantMedium::hit draws from the process-global RNG in the reference (constant_medium.rs:52-55, its own TODO: not
// repeatable).  The oracle draws from the path's seeded stream instead, handed over through this pointer by ray_color;
// with no stream attached (orc_ow_trace) a medium is never hit.
static thread_local ChaCha8Rng* tl_medium_rng = nullptr;

// TEST HOOK (orc_ow_trace_self only; -1 everywhere else, which leaves the restatement untouched).  The reference keeps a
// scattered ray from re-hitting the surface it starts on with f64 and t_min = 1e-10 (camera.rs:242).  The f32 device
// path cannot (its origin is ~1e-7 off the surface), so it states the same intent as a rule: a ray never re-hits the
// PLANAR primitive it starts on, and takes only the FAR root of its own sphere, beyond 1e-4 radii.  When the parity
// tests hand f32-rounded scattered rays to this oracle, the oracle must apply that rule too — otherwise it would
// "hit" the start surface at t ~ 1e-8 whenever rounding pushed the origin inside.
static thread_local int tl_self_node = -1;

struct Scene;
struct Obj {
    int kind, node, material = -1;
    AABB bbox;
    // sphere
    V3 c1, c2; double radius; bool moving;
    // plane basis (quad / triangle) — flat/plane.rs:23-40
    V3 q, u, v, w, normal; double d;
    bool has_uv = false, has_n = false; V3 n1, n2, n3; double uv[6];
    // transform / translate
    M3 M, Minv, MinvT; V3 offset;
    // constant medium
    double neg_inv_density = 0.0;
    // children (list / bvh node / the boundary of a medium)
    std::vector<std::unique_ptr<Obj>> kids;
    bool is_bvh_leaf = false;
};

struct Scene {
    const rl_scene_desc* d;
    std::unique_ptr<Obj> root;
    bool ok = true;

    explicit Scene(const rl_scene_desc* desc) : d(desc) {
        if (d->n_roots != 1) { ok = false; return; }
        root = build(d->roots[0]);
    }

    void plane_init(Obj& o, V3 q, V3 u, V3 v) {  // Plane::new (flat/plane.rs:23-40)
        V3 n = cross(u, v);
        if (!try_normalize(n, &o.normal)) ok = false;
        o.d = dot(o.normal, q);
        o.w = n / dot(n, n);
        o.q = q; o.u = u; o.v = v;
    }

    std::unique_ptr<Obj> build(int id) {
        const rl_node& nd = d->nodes[id];
        auto o = std::make_unique<Obj>();
        o->kind = nd.kind; o->node = id; o->material = nd.material;
        const double* p = nd.param >= 0 ? d->params + nd.param : nullptr;
        switch (nd.kind) {
            case RL_OW_SPHERE: {  // sphere.rs:77-87
                o->c1 = v3(p[0], p[1], p[2]); o->c2 = v3(p[3], p[4], p[5]); o->radius = p[6];
                o->moving = nd.flags & 1;
                V3 rv = v3(o->radius, o->radius, o->radius);
                AABB b = AABB::from_extrema(o->c1 - rv, o->c1 + rv);
                if (o->moving) b = b.merge(AABB::from_extrema(o->c2 - rv, o->c2 + rv));
                o->bbox = b;
                break;
            }
            case RL_OW_QUAD: {  // flat/quad.rs:26-34
                V3 q = v3(p[0], p[1], p[2]), u = v3(p[3], p[4], p[5]), v = v3(p[6], p[7], p[8]);
                AABB d1 = AABB::from_extrema(q, q + u + v), d2 = AABB::from_extrema(q + u, q + v);
                o->bbox = d1.merge(d2);
                plane_init(*o, q, u, v);
                break;
            }
            case RL_OW_TRIANGLE: {  // flat/triangle.rs:32-56, AABB::from_points (aabb.rs:68-90)
                V3 p1 = v3(p[0], p[1], p[2]), p2 = v3(p[3], p[4], p[5]), p3 = v3(p[6], p[7], p[8]);
                V3 mn = p1, mx = p1;
                for (V3 q : {p1, p2, p3}) {
                    mn = {std::fmin(mn.x, q.x), std::fmin(mn.y, q.y), std::fmin(mn.z, q.z)};
                    mx = {std::fmax(mx.x, q.x), std::fmax(mx.y, q.y), std::fmax(mx.z, q.z)};
                }
                o->bbox = AABB::from_extrema(mn, mx);
                plane_init(*o, p1, p2 - p1, p3 - p1);
                o->has_uv = nd.flags & 1; o->has_n = nd.flags & 2;
                for (int k = 0; k < 6; k++) o->uv[k] = p[9 + k];
                o->n1 = v3(p[15], p[16], p[17]); o->n2 = v3(p[18], p[19], p[20]); o->n3 = v3(p[21], p[22], p[23]);
                break;
            }
            case RL_OW_TRANSFORM: {  // hittable/transform.rs:88-140
                for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
                    o->M.m[i][j] = p[i * 3 + j]; o->Minv.m[i][j] = p[9 + i * 3 + j];
                }
                for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o->MinvT.m[i][j] = o->Minv.m[j][i];
                o->kids.push_back(build(nd.child_begin));
                const AABB& b = o->kids[0]->bbox;
                double mnx = INF, mny = INF, mnz = INF, mxx = -INF, mxy = -INF, mxz = -INF;
                for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) {
                    double fi = i, fj = j, fk = k;
                    double x = fi * b.x.max + (1.0 - fi) * b.x.min;
                    double y = fj * b.y.max + (1.0 - fj) * b.y.min;
                    double z = fk * b.z.max + (1.0 - fk) * b.z.min;
                    V3 t = mul(o->M, v3(x, y, z));
                    mnx = std::fmin(mnx, t.x); mxx = std::fmax(mxx, t.x);
                    mny = std::fmin(mny, t.y); mxy = std::fmax(mxy, t.y);
                    mnz = std::fmin(mnz, t.z); mxz = std::fmax(mxz, t.z);
                }
                o->bbox = AABB::from_extrema(v3(mnx, mny, mnz), v3(mxx, mxy, mxz));
                break;
            }
            case RL_OW_TRANSLATE: {  // hittable/translate.rs:23-25, aabb.rs:160-166
                o->offset = v3(p[0], p[1], p[2]);
                o->kids.push_back(build(nd.child_begin));
                const AABB& b = o->kids[0]->bbox;
                o->bbox = AABB::make({b.x.min + o->offset.x, b.x.max + o->offset.x},
                                     {b.y.min + o->offset.y, b.y.max + o->offset.y},
                                     {b.z.min + o->offset.z, b.z.max + o->offset.z});
                break;
            }
            case RL_OW_LIST: {  // hittable/mod.rs:107-110
                AABB b = AABB::empty();
                for (int k = nd.child_begin; k < nd.child_end; k++) {
                    o->kids.push_back(build(d->children[k]));
                    b = b.merge(o->kids.back()->bbox);
                }
                o->bbox = b;
                break;
            }
            case RL_OW_BVH: {  // bvh.rs:22-61
                std::vector<std::unique_ptr<Obj>> hs;
                for (int k = nd.child_begin; k < nd.child_end; k++) hs.push_back(build(d->children[k]));
                if (hs.empty()) { ok = false; break; }
                o = bvh_new(std::move(hs));
                o->node = id;
                break;
            }
            case RL_OW_CONSTANT_MEDIUM: {  // constant_medium.rs:14-22, bounding_box 85-87
                if (!p || nd.child_begin < 0) { ok = false; break; }
                o->neg_inv_density = -1.0 / p[0];
                o->kids.push_back(build(nd.child_begin));
                o->bbox = o->kids[0]->bbox;
                break;
            }
            default: ok = false;
        }
        return o;
    }

    static int longest_axis(const AABB& b) {  // bvh.rs:63-77
        if (b.x.size() > b.y.size()) return b.x.size() > b.z.size() ? 0 : 2;
        return b.y.size() > b.z.size() ? 1 : 2;
    }
    // f64::total_cmp
    static bool total_less(double a, double b) {
        int64_t x, y;
        std::memcpy(&x, &a, 8); std::memcpy(&y, &b, 8);
        x ^= (int64_t)((uint64_t)(x >> 63) >> 1);
        y ^= (int64_t)((uint64_t)(y >> 63) >> 1);
        return x < y;
    }
    std::unique_ptr<Obj> bvh_new(std::vector<std::unique_ptr<Obj>> hs) {
        auto o = std::make_unique<Obj>();
        o->kind = RL_OW_BVH; o->node = -1;
        if (hs.size() == 1) {
            o->bbox = hs[0]->bbox;
            o->kids.push_back(std::move(hs[0]));
            o->is_bvh_leaf = true;
        } else if (hs.size() == 2) {
            // swap_remove(0) twice: left = hs[0], then the last element moved to 0 -> right = hs[1]
            o->bbox = hs[0]->bbox.merge(hs[1]->bbox);
            o->kids.push_back(std::move(hs[0]));
            o->kids.push_back(std::move(hs[1]));
            o->is_bvh_leaf = true;
        } else {
            AABB b = AABB::empty();
            for (auto& h : hs) b = b.merge(h->bbox);
            int axis = longest_axis(b);
            auto key = [axis](const std::unique_ptr<Obj>& h) {
                return axis == 0 ? h->bbox.x.min : axis == 1 ? h->bbox.y.min : h->bbox.z.min;
            };
            // sort_unstable_by(total_cmp): order of equal keys is unspecified in the reference too
            std::stable_sort(hs.begin(), hs.end(), [&](const std::unique_ptr<Obj>& l, const std::unique_ptr<Obj>& r) {
                return total_less(key(l), key(r));
            });
            size_t mid = hs.size() / 2;
            std::vector<std::unique_ptr<Obj>> ls, rs;
            for (size_t i = 0; i < hs.size(); i++) (i < mid ? ls : rs).push_back(std::move(hs[i]));
            o->kids.push_back(bvh_new(std::move(ls)));
            o->kids.push_back(bvh_new(std::move(rs)));
            o->bbox = b;
        }
        return o;
    }

    // `[H]::hit` (hittable/mod.rs:88-105)
    bool hit_slice(const std::vector<std::unique_ptr<Obj>>& hs, const Ray& r, const Interval& rt, HitRecord* rec) const {
        bool any = false;
        double closest = rt.max;
        for (const auto& h : hs) {
            HitRecord tmp;
            if (hit(*h, r, {rt.min, closest}, &tmp)) {
                any = true;
                closest = tmp.t;
                *rec = tmp;
            }
        }
        return any;
    }

    // Plane::hit_ab (flat/plane.rs:51-80) + Plane::hit (86-100)
    bool hit_plane(const Obj& o, const Ray& r, const Interval& rt, HitRecord* rec) const {
        if (o.node == tl_self_node) return false;  // test hook, see tl_self_node
        double denom = dot(o.normal, r.direction);
        if (std::fabs(denom) < 1e-8) return false;
        double t = (o.d - dot(o.normal, r.origin)) / denom;
        if (!rt.contains(t)) return false;
        V3 ip = r.at(t);
        V3 ph = ip - o.q;
        double alpha = dot(o.w, cross(ph, o.v));
        double beta = dot(o.w, cross(o.u, ph));
        face_normal(r, o.normal, &rec->normal, &rec->front);
        rec->p = ip; rec->t = t; rec->u = alpha; rec->v = beta;
        rec->material = o.material; rec->node = o.node;
        return true;
    }

    bool hit(const Obj& o, const Ray& r, const Interval& rt, HitRecord* rec) const {
        switch (o.kind) {
            case RL_OW_SPHERE: {  // sphere.rs:34-75
                V3 center = o.moving ? o.c1 + r.time * (o.c2 - o.c1) : o.c1;
                V3 oc = r.origin - center;
                double a = length_squared(r.direction);
                double half_b = dot(oc, r.direction);
                double c = length_squared(oc) - o.radius * o.radius;
                double disc = half_b * half_b - a * c;
                if (disc < 0.0) return false;
                double s = std::sqrt(disc);
                double r_l = (-half_b - s) / a, r_u = (-half_b + s) / a;
                double t;
                if (o.node == tl_self_node) {  // test hook, see tl_self_node
                    r_u = -2.0 * half_b / a;
                    if (!(r_u > 1e-4 * std::fabs(o.radius) / std::sqrt(a)) || !rt.contains(r_u)) return false;
                    t = r_u;
                } else
                if (rt.contains(r_l)) t = r_l;
                else if (rt.contains(r_u)) t = r_u;
                else return false;
                V3 p = r.at(t);
                V3 outward = (p - center) / o.radius;
                face_normal(r, outward, &rec->normal, &rec->front);
                rec->p = p; rec->t = t;
                // get_sphere_uv (sphere.rs:91-99)
                double theta = std::acos(-outward.y);
                double phi = std::atan2(-outward.z, outward.x) + PI;
                rec->u = phi / (2.0 * PI); rec->v = theta / PI;
                rec->material = o.material; rec->node = o.node;
                return true;
            }
            case RL_OW_QUAD: {  // flat/quad.rs:37-42
                HitRecord h;
                if (!hit_plane(o, r, rt, &h)) return false;
                if (!(0.0 <= h.u && h.u <= 1.0 && 0.0 <= h.v && h.v <= 1.0)) return false;
                *rec = h;
                return true;
            }
            case RL_OW_TRIANGLE: {  // flat/triangle.rs:60-95
                HitRecord h;
                if (!hit_plane(o, r, rt, &h)) return false;
                if (!(0.0 <= h.u && 0.0 <= h.v && h.u + h.v <= 1.0)) return false;
                double frac2 = h.u, frac3 = h.v, frac1 = 1.0 - h.u - h.v;
                if (o.has_n) {
                    V3 n{0, 0, 0};
                    try_normalize(o.n2 * frac2 + o.n3 * frac3 + o.n1 * frac1, &n);
                    face_normal(r, n, &h.normal, &h.front);
                }
                if (o.has_uv) {
                    double nu = o.uv[0] * frac1 + o.uv[2] * frac2 + o.uv[4] * frac3;
                    double nv = o.uv[1] * frac1 + o.uv[3] * frac2 + o.uv[5] * frac3;
                    h.u = nu; h.v = nv;
                }
                *rec = h;
                return true;
            }
            case RL_OW_TRANSFORM: {  // hittable/transform.rs:145-164
                Ray tr{mul(o.Minv, r.origin), mul(o.Minv, r.direction), r.time};
                if (!hit(*o.kids[0], tr, rt, rec)) return false;
                rec->p = mul(o.M, rec->p);
                V3 n{0, 0, 0};
                try_normalize(mul(o.MinvT, rec->normal), &n);
                rec->normal = n;
                return true;
            }
            case RL_OW_TRANSLATE: {  // hittable/translate.rs:14-21
                Ray tr{r.origin - o.offset, r.direction, r.time};
                if (!hit(*o.kids[0], tr, rt, rec)) return false;
                rec->p = rec->p + o.offset;
                return true;
            }
            case RL_OW_CONSTANT_MEDIUM: {  // constant_medium.rs:28-83
                if (!tl_medium_rng) return false;
                const Obj& b = *o.kids[0];
                HitRecord rec1, rec2;
                if (!hit(b, r, {-INF, INF}, &rec1)) return false;
                if (!hit(b, r, {rec1.t + 1e-4, INF}, &rec2)) return false;
                rec1.t = std::fmax(rec1.t, rt.min);
                rec2.t = std::fmin(rec2.t, rt.max);
                if (rec1.t >= rec2.t) return false;
                rec1.t = std::fmax(rec1.t, 0.0);
                double ray_length = length(r.direction);
                double distance_inside_boundary = (rec2.t - rec1.t) * ray_length;
                double hit_distance = o.neg_inv_density * std::log(tl_medium_rng->gen_f64());
                if (hit_distance > distance_inside_boundary) return false;
                double t = rec1.t + hit_distance / ray_length;
                rec->p = r.at(t);
                rec->normal = v3(1.0, 0.0, 0.0);  // arbitrary
                rec->t = t; rec->u = 0.0; rec->v = 0.0; rec->front = true;
                rec->material = o.material; rec->node = o.node;
                return true;
            }
            case RL_OW_LIST:
                return hit_slice(o.kids, r, rt, rec);
            case RL_OW_BVH:  // bvh.rs:81-90
                if (!o.bbox.hit(r, rt)) return false;
                return hit_slice(o.kids, r, rt, rec);
        }
        return false;
    }

    // perlin.rs:39-63 (noise), 91-115 (perlin_interp), 65-78 (turb)
    double perlin_noise(const rl_perlin& pn, V3 p) const {
        double fx = std::floor(p.x), fy = std::floor(p.y), fz = std::floor(p.z);
        double u = p.x - fx, v = p.y - fy, w = p.z - fz;
        int i = (int)fx, j = (int)fy, k = (int)fz;
        double uu = u * u * (3.0 - 2.0 * u), vv = v * v * (3.0 - 2.0 * v), ww = w * w * (3.0 - 2.0 * w);
        double accum = 0.0;
        for (int di = 0; di < 2; di++)
            for (int dj = 0; dj < 2; dj++)
                for (int dk = 0; dk < 2; dk++) {
                    const double* c = pn.randvec[pn.perm_x[(i + di) & 255] ^ pn.perm_y[(j + dj) & 255] ^ pn.perm_z[(k + dk) & 255]];
                    double i_f = di, j_f = dj, k_f = dk;
                    V3 weight_v = v3(u - i_f, v - j_f, w - k_f);
                    accum += (i_f * uu + (1.0 - i_f) * (1.0 - uu)) * (j_f * vv + (1.0 - j_f) * (1.0 - vv)) *
                             (k_f * ww + (1.0 - k_f) * (1.0 - ww)) * dot(v3(c[0], c[1], c[2]), weight_v);
                }
        return accum;
    }
    double perlin_turb(const rl_perlin& pn, V3 p, int depth) const {
        double accum = 0.0, weight = 1.0;
        V3 temp_p = p;
        for (int i = 0; i < depth; i++) {
            accum += weight * perlin_noise(pn, temp_p);
            weight *= 0.5;
            temp_p = temp_p * 2.0;
        }
        return std::fabs(accum);
    }

    // texture.rs
    V3 tex_value(int tex, double u, double v, V3 p) const {
        const rl_texture& t = d->textures[tex];
        switch (t.kind) {
            case RL_TEX_OW_SOLID: return v3(t.a[0], t.a[1], t.a[2]);
            case RL_TEX_OW_CHECKER: {  // 42-54
                double inv_scale = 1.0 / t.scale;
                int64_t xi = (int64_t)std::floor(p.x * inv_scale);
                int64_t yi = (int64_t)std::floor(p.y * inv_scale);
                int64_t zi = (int64_t)std::floor(p.z * inv_scale);
                bool even = (xi + yi + zi) % 2 == 0;
                return tex_value(even ? t.tex_a : t.tex_b, u, v, p);
            }
            case RL_TEX_OW_IMAGE: {  // 63-81
                const rl_image& im = d->images[t.image];
                double uu = std::fmin(std::fmax(u, 0.0), 1.0);
                double vv = 1.0 - std::fmin(std::fmax(v, 0.0), 1.0);
                uint32_t i = (uint32_t)(uu * (double)(im.width - 1));
                uint32_t j = (uint32_t)(vv * (double)(im.height - 1));
                const float* px = im.rgb + ((size_t)j * im.width + i) * 3;
                return v3((double)px[0], (double)px[1], (double)px[2]);
            }
            case RL_TEX_OW_NOISE: {  // 89-93
                const rl_perlin& pn = d->perlins[t.image];
                return v3(0.5, 0.5, 0.5) * (1.0 + std::sin(t.scale * p.z + 10.0 * perlin_turb(pn, p, 7)));
            }
        }
        return v3(0, 0, 0);
    }

    // material.rs: emitted
    V3 emitted(int m, double u, double v, V3 p) const {
        const rl_material& mt = d->materials[m];
        if (mt.kind == RL_MAT_OW_DIFFUSE_LIGHT) return tex_value(mt.texture, u, v, p);  // 192-194
        return v3(0, 0, 0);
    }
    // material.rs: scatter
    bool scatter(int m, ChaCha8Rng& rng, const Ray& ray, const HitRecord& h, V3* atten, Ray* out) const {
        const rl_material& mt = d->materials[m];
        switch (mt.kind) {
            case RL_MAT_OW_LAMBERTIAN: {  // 74-92
                V3 dir = h.normal + random_unit_vector(rng);
                if (near_zero(dir)) dir = h.normal;
                *out = {h.p, dir, ray.time};
                *atten = tex_value(mt.texture, h.u, h.v, h.p);
                return true;
            }
            case RL_MAT_OW_METAL: {  // 105-122
                V3 reflected = reflect(ray.direction, h.normal);
                V3 fuzzed = normalize(reflected) + (mt.fuzz * random_unit_vector(rng));
                *out = {h.p, fuzzed, ray.time};
                if (dot(fuzzed, h.normal) > 0.0) {
                    *atten = v3(mt.color[0], mt.color[1], mt.color[2]);
                    return true;
                }
                return false;
            }
            case RL_MAT_OW_DIELECTRIC: {  // 139-165, reflectance 173-176
                double ri = h.front ? 1.0 / mt.refractive_index : mt.refractive_index;
                V3 unit{0, 0, 0};
                try_normalize(ray.direction, &unit);
                double cos_theta = std::fmin(dot(-unit, h.normal), 1.0);
                double sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
                bool cannot_refract = ri * sin_theta > 1.0;
                bool refl = cannot_refract;
                if (!refl) {
                    double r0 = (1.0 - ri) / (1.0 + ri);
                    r0 = r0 * r0;
                    double reflectance = r0 + (1.0 - r0) * powi5(1.0 - cos_theta);
                    refl = reflectance > rng.gen_f64();
                }
                V3 dir = refl ? reflect(unit, h.normal) : refract(unit, h.normal, ri);
                *out = {h.p, dir, ray.time};
                *atten = v3(1, 1, 1);
                return true;
            }
            case RL_MAT_OW_ISOTROPIC: {  // 201-216
                *out = {h.p, random_unit_vector(rng), ray.time};
                *atten = tex_value(mt.texture, h.u, h.v, h.p);
                return true;
            }
            default: return false;  // DiffuseLight never scatters (182-189)
        }
    }
};

// camera.rs
struct Camera {
    rl_ow_camera p;
    int width, height;
    V3 lookfrom, pixel00, du, dv, disk_u, disk_v, background;
    explicit Camera(const rl_ow_camera* c) : p(*c) {  // Camera::new (72-118)
        width = p.image_width;
        size_t hh = (size_t)((double)width / p.aspect_ratio);
        height = (int)std::max<size_t>(hh, 1);
        lookfrom = v3(p.lookfrom[0], p.lookfrom[1], p.lookfrom[2]);
        V3 lookat = v3(p.lookat[0], p.lookat[1], p.lookat[2]), vup = v3(p.vup[0], p.vup[1], p.vup[2]);
        double theta = p.vfov * PI / 180.0;  // utility.rs:1-3
        double h = std::tan(theta / 2.0);
        double viewport_height = 2.0 * h * p.focus_dist;
        double viewport_width = viewport_height * ((double)width / (double)height);
        V3 w{0, 0, 0}, u{0, 0, 0}, v{0, 0, 0};
        try_normalize(lookfrom - lookat, &w);
        try_normalize(cross(vup, w), &u);
        try_normalize(cross(w, u), &v);
        V3 viewport_u = viewport_width * u;
        V3 viewport_v = viewport_height * -v;
        du = viewport_u / (double)width;
        dv = viewport_v / (double)height;
        V3 upper_left = lookfrom - (p.focus_dist * w) - viewport_u / 2.0 - viewport_v / 2.0;
        pixel00 = upper_left + 0.5 * (du + dv);
        double defocus_radius = p.focus_dist * std::tan((p.defocus_angle / 2.0) * PI / 180.0);
        disk_u = u * defocus_radius;
        disk_v = v * defocus_radius;
        background = v3(p.background[0], p.background[1], p.background[2]);
    }
    Ray get_ray(ChaCha8Rng& rng, int i, int j) const {  // 203-230
        V3 pixel_center = pixel00 + ((double)i * du) + ((double)j * dv);
        double px = -0.5 + rng.gen_f64();
        double py = -0.5 + rng.gen_f64();
        V3 pixel_sample = pixel_center + ((px * du) + (py * dv));
        V3 origin = lookfrom;
        if (!(p.defocus_angle <= 0.0)) {
            double x1, x2;
            for (;;) {  // rand_distr::UnitDisc
                x1 = rng.uniform_m1_1();
                x2 = rng.uniform_m1_1();
                if (x1 * x1 + x2 * x2 <= 1.0) break;
            }
            origin = lookfrom + (x1 * disk_u) + (x2 * disk_v);
        }
        V3 dir = pixel_sample - origin;
        double time = rng.gen_f64();
        return {origin, dir, time};
    }
    V3 ray_color(ChaCha8Rng& rng, const Ray& r, const Scene& sc, int depth, uint64_t* rays) const {  // 232-260
        if (depth == 0) return v3(0, 0, 0);
        if (rays) ++*rays;
        HitRecord h;
        tl_medium_rng = &rng;
        bool was_hit = sc.hit(*sc.root, r, {1e-10, INF}, &h);
        tl_medium_rng = nullptr;
        if (was_hit) {
            V3 em = sc.emitted(h.material, h.u, h.v, h.p);
            V3 att;
            Ray sr;
            if (sc.scatter(h.material, rng, r, h, &att, &sr)) {
                V3 sc_col = att * ray_color(rng, sr, sc, depth - 1, rays);
                return em + sc_col;
            }
            return em;
        }
        return background;
    }
};

}  // namespace

extern "C" {

int orc_ow_image_height(const rl_ow_camera* cam) { return Camera(cam).height; }

// Camera::_render (camera.rs:145-199): out_sum = H*W*3 f64 SUMS, row-major. rows [y0,y1) only (bounded
// CPU-baseline samples); rays (optional) receives the number of rays cast.
int orc_ow_render(const rl_scene_desc* desc, const rl_ow_camera* cam, uint32_t first_sample, double* out_sum,
                  int y0, int y1, int threads, uint64_t* rays_out) {
    if (!desc || desc->flavor != RL_FLAVOR_OW) return -1;
    Scene sc(desc);
    if (!sc.ok) return -1;
    Camera c(cam);
    if (y1 < 0 || y1 > c.height) y1 = c.height;
    if (y0 < 0) y0 = 0;
    uint64_t total_rays = 0;
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total_rays) num_threads(threads > 0 ? threads : omp_get_max_threads())
#endif
    for (int y = y0; y < y1; y++) {
        for (int x = 0; x < c.width; x++) {
            int i = x, j = y;  // camera.rs:155: `(x, y)` destructured as `(i, j)`
            ChaCha8Rng rng(c.p.seed);
            V3 sum = v3(0, 0, 0);
            uint64_t rays = 0;
            for (uint32_t n = 0; n < (uint32_t)c.p.samples_per_pixel; n++) {
                uint64_t sample_index = (uint64_t)n + first_sample;
                uint64_t stream = sample_index * (uint64_t)c.width * (uint64_t)c.height +
                                  (uint64_t)i * (uint64_t)c.width + (uint64_t)j;
                rng.set_stream(stream);
                Ray r = c.get_ray(rng, i, j);
                sum = sum + c.ray_color(rng, r, sc, c.p.max_depth, &rays);
            }
            total_rays += rays;
            double* o = out_sum + ((size_t)y * c.width + x) * 3;
            o[0] = sum.x; o[1] = sum.y; o[2] = sum.z;
        }
    }
    if (rays_out) *rays_out = total_rays;
    return 0;
}

// world.hit(r, [1e-10, inf)) for a batch of rays (n*7 doubles: origin, direction, time)
int orc_ow_trace(const rl_scene_desc* desc, const double* rays, uint64_t n, int32_t* node, double* t, double* uv,
                 int threads) {
    if (!desc || desc->flavor != RL_FLAVOR_OW) return -1;
    Scene sc(desc);
    if (!sc.ok) return -1;
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 256) num_threads(threads > 0 ? threads : omp_get_max_threads())
#endif
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const double* r = rays + 7 * i;
        Ray ray{v3(r[0], r[1], r[2]), v3(r[3], r[4], r[5]), r[6]};
        HitRecord h;
        if (sc.hit(*sc.root, ray, {1e-10, INF}, &h)) {
            node[i] = h.node; t[i] = h.t;
            if (uv) { uv[2 * i] = h.u; uv[2 * i + 1] = h.v; }
        } else {
            node[i] = -1; t[i] = INF;
            if (uv) { uv[2 * i] = 0; uv[2 * i + 1] = 0; }
        }
    }
    return 0;
}

// orc_ow_trace for rays that start on a surface: self_nodes[i] = the leaf node ray i starts on (-1 none); see tl_self_node
int orc_ow_trace_self(const rl_scene_desc* desc, const double* rays, const int32_t* self_nodes, uint64_t n, int32_t* node,
                      double* t, double* uv, int threads) {
    if (!desc || desc->flavor != RL_FLAVOR_OW) return -1;
    Scene sc(desc);
    if (!sc.ok) return -1;
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 256) num_threads(threads > 0 ? threads : omp_get_max_threads())
#endif
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const double* r = rays + 7 * i;
        Ray ray{v3(r[0], r[1], r[2]), v3(r[3], r[4], r[5]), r[6]};
        HitRecord h;
        tl_self_node = self_nodes ? self_nodes[i] : -1;
        const bool was_hit = sc.hit(*sc.root, ray, {1e-10, INF}, &h);
        tl_self_node = -1;
        if (was_hit) {
            node[i] = h.node; t[i] = h.t;
            if (uv) { uv[2 * i] = h.u; uv[2 * i + 1] = h.v; }
        } else {
            node[i] = -1; t[i] = INF;
            if (uv) { uv[2 * i] = 0; uv[2 * i + 1] = 0; }
        }
    }
    return 0;
}

// The reference's OWN ray at bounce `bounce` of every pixel's first sample (bounce 0 = the camera ray, 1 = the first
// scattered ray, ...): get_ray, then `bounce` rounds of world.hit + Material::scatter with the reference's RNG stream
// (camera.rs:161-170, 232-260).  rays: n*7 doubles; self_nodes[i] = the leaf the ray starts on (-1 for camera rays),
// or -2 when the path ended before that bounce (miss, absorbed, light).
int orc_ow_bounce_rays(const rl_scene_desc* desc, const rl_ow_camera* cam, int bounce, double* rays, int32_t* self_nodes) {
    if (!desc || desc->flavor != RL_FLAVOR_OW || bounce < 0) return -1;
    Scene sc(desc);
    if (!sc.ok) return -1;
    Camera c(cam);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4)
#endif
    for (int y = 0; y < c.height; y++)
        for (int x = 0; x < c.width; x++) {
            ChaCha8Rng rng(c.p.seed);
            rng.set_stream((uint64_t)x * (uint64_t)c.width + (uint64_t)y);
            Ray r = c.get_ray(rng, x, y);
            int self = -1;
            for (int b = 0; b < bounce && self != -2; b++) {
                HitRecord h;
                tl_medium_rng = &rng;
                const bool was_hit = sc.hit(*sc.root, r, {1e-10, INF}, &h);
                tl_medium_rng = nullptr;
                V3 att;
                Ray sr;
                if (was_hit && sc.scatter(h.material, rng, r, h, &att, &sr)) {
                    r = sr;
                    self = h.node;
                } else {
                    self = -2;
                }
            }
            const size_t k = (size_t)y * c.width + x;
            double* o = rays + 7 * k;
            o[0] = r.origin.x; o[1] = r.origin.y; o[2] = r.origin.z;
            o[3] = r.direction.x; o[4] = r.direction.y; o[5] = r.direction.z; o[6] = r.time;
            self_nodes[k] = self;
        }
    return 0;
}

// first-sample camera rays (jittered exactly like the reference) for identical-ray-batch tests
int orc_ow_camera_rays(const rl_ow_camera* cam, double* rays) {
    Camera c(cam);
    size_t k = 0;
    for (int y = 0; y < c.height; y++)
        for (int x = 0; x < c.width; x++) {
            ChaCha8Rng rng(c.p.seed);
            rng.set_stream((uint64_t)x * (uint64_t)c.width + (uint64_t)y);
            Ray r = c.get_ray(rng, x, y);
            double* o = rays + 7 * k++;
            o[0] = r.origin.x; o[1] = r.origin.y; o[2] = r.origin.z;
            o[3] = r.direction.x; o[4] = r.direction.y; o[5] = r.direction.z; o[6] = r.time;
        }
    return 0;
}

// raw RNG words for known-answer tests: seed, stream, n u64 outputs
// Texture::value at n points (u, v, x, y, z) — lets the tests pin the Perlin / Noise restatement against the host mirror
int orc_ow_tex_value(const rl_scene_desc* d, int tex, uint64_t n, const double* uvp, double* rgb) {
    if (!d || tex < 0 || tex >= d->n_textures) return -1;
    Scene sc(d);
    for (uint64_t i = 0; i < n; i++) {
        V3 c = sc.tex_value(tex, uvp[5 * i], uvp[5 * i + 1], v3(uvp[5 * i + 2], uvp[5 * i + 3], uvp[5 * i + 4]));
        rgb[3 * i] = c.x; rgb[3 * i + 1] = c.y; rgb[3 * i + 2] = c.z;
    }
    return 0;
}

int orc_chacha8_u64(uint64_t seed, uint64_t stream, int n, uint64_t* out) {
    ChaCha8Rng rng(seed);
    rng.set_stream(stream);
    for (int i = 0; i < n; i++) out[i] = rng.next_u64();
    return 0;
}

}  // extern "C"
