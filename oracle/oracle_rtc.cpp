// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement (C++17, f64) of the ray-tracer-challenge (RTC) per-pixel ray loop of
// marcantony/rendering-learning.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs may load this library.  It deliberately keeps the reference's algorithmic
// choices (a Vec of all roots per object, stable sort per group / world, brute-force triangle lists
// behind one AABB, recursion) and its operation order, so that it reproduces the reference's golden
// PPMs byte for byte (tests/test_oracle_golden.py) and is a representative CPU baseline.
//
// Parity pinned by: RTC/tests/expectations/test_{mirror,obj,csg}_scene.ppm and the unit known-answer
// vectors of SURVEY.md §4 (tests/test_oracle_rtc_kat.py).
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/ray-tracer-challenge/src/).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../include/rl_b200.h"

#ifdef _OPENMP
#include <omp.h>
static inline int omp_get_max_threads_compat() { return omp_get_max_threads(); }
#endif

namespace {

// f64::powi(5) as LLVM expands it / compiler-rt __powidf2 computes it: x * ((x*x)*(x*x))
inline double powi5(double x) {
    double x2 = x * x;
    double x4 = x2 * x2;
    return x * x4;
}

constexpr double INF = std::numeric_limits<double>::infinity();

// ---- math/vector.rs, math/point.rs ----------------------------------------------------------
struct V3 {
    double x, y, z;
};
inline V3 v3(double x, double y, double z) { return {x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 mulc(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }  // draw/color.rs:53-59
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // vector.rs:45-47
inline V3 cross(V3 a, V3 b) {                                                // vector.rs:49-55
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double mag(V3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }  // vector.rs:32-34
// Vec3d::norm (vector.rs:36-43) and NormalizedVec3d::try_from (vector.rs:141-156): divide by mag.
inline bool normalize(V3 a, V3* out) {
    double m = mag(a);
    if (m == 0.0) return false;
    *out = {a.x / m, a.y / m, a.z / m};
    return true;
}
inline V3 normalized(V3 a) {
    V3 r = a;
    normalize(a, &r);  // the reference unwraps; a zero vector would panic there
    return r;
}
// Vec3d::reflect (vector.rs:57-59): self - (normal*2.0)*self.dot(normal)
inline V3 reflect(V3 v, V3 n) { return v - (n * 2.0) * dot(v, n); }

// ---- math/matrix.rs ---------------------------------------------------------------------------
struct M4 {
    double m[4][4];
};
inline M4 m4_from(const double* p) {
    M4 r;
    std::memcpy(r.m, p, sizeof(r.m));
    return r;
}
// determinant / cofactor / minor by recursive first-row expansion (matrix.rs:88-131)
double det_n(const double* d, int n);
double minor_n(const double* d, int n, int r, int c) {
    double sub[9];
    int k = 0;
    for (int i = 0; i < n; i++) {
        if (i == r) continue;
        for (int j = 0; j < n; j++) {
            if (j == c) continue;
            sub[k++] = d[i * n + j];
        }
    }
    return det_n(sub, n - 1);
}
double cofactor_n(const double* d, int n, int r, int c) {
    double mi = minor_n(d, n, r, c);
    return ((r + c) % 2 == 0) ? mi : -mi;
}
double det_n(const double* d, int n) {
    if (n == 2) return d[0] * d[3] - d[1] * d[2];
    double sum = 0.0;
    for (int i = 0; i < n; i++) sum += d[i] * cofactor_n(d, n, 0, i);
    return sum;
}
// SquareMatrix::invert (matrix.rs:68-86)
bool invert(const M4& a, M4* out) {
    const double* d = &a.m[0][0];
    double det = det_n(d, 4);
    if (det == 0.0) return false;
    for (int n = 0; n < 4; n++)
        for (int m = 0; m < 4; m++) out->m[m][n] = cofactor_n(d, 4, n, m) / det;
    return true;
}
inline M4 transpose(const M4& a) {
    M4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) r.m[j][i] = a.m[i][j];
    return r;
}
// &SquareMatrix<4> * &Point3d (point.rs:89-96): 4x4 * 4x1 with w = 1, sum from 0.0 in column order
inline V3 mul_point(const M4& a, V3 p) {
    double v[4] = {p.x, p.y, p.z, 1.0}, o[3];
    for (int n = 0; n < 3; n++) {
        double sum = 0.0;
        for (int i = 0; i < 4; i++) sum += a.m[n][i] * v[i];
        o[n] = sum;
    }
    return {o[0], o[1], o[2]};
}
// &SquareMatrix<4> * &Vec3d (vector.rs:115-122): w = 0
inline V3 mul_vec(const M4& a, V3 p) {
    double v[4] = {p.x, p.y, p.z, 0.0}, o[3];
    for (int n = 0; n < 3; n++) {
        double sum = 0.0;
        for (int i = 0; i < 4; i++) sum += a.m[n][i] * v[i];
        o[n] = sum;
    }
    return {o[0], o[1], o[2]};
}

// math/util.rs:1-22
bool are_equal(double a, double b) {
    if (std::isnan(a) || std::isnan(b)) return false;
    if (std::isinf(a) && std::isinf(b)) return a == b;
    double abs_diff = std::fabs(a - b);
    if (abs_diff <= std::numeric_limits<double>::epsilon() * 2.0) return true;
    uint64_t au, bu;
    std::memcpy(&au, &a, 8);
    std::memcpy(&bu, &b, 8);
    uint64_t ulps = au > bu ? au - bu : bu - au;
    return ulps <= 8;
}

// Rust f64::max / f64::min ignore a NaN operand; fmax/fmin have the same rule.
inline double rmax(double a, double b) { return std::fmax(a, b); }
inline double rmin(double a, double b) { return std::fmin(a, b); }

// ---- scene/ray.rs --------------------------------------------------------------------------------
struct Ray {
    V3 origin, direction;
    V3 position(double t) const { return origin + direction * t; }  // ray.rs:14-16
};

struct Bounds {
    V3 mn, mx;
};

// scene/intersect.rs:10-16
struct Isect {
    double t;
    int object;  // leaf node id (object identity = address of the leaf in the reference)
    V3 color;
    V3 normal;
};

struct Scene {
    const rl_scene_desc* d;
    std::vector<M4> inv, inv_t;        // per TRANSFORMED node (indexed by node id)
    std::vector<Bounds> bounds;        // per BOUNDED node
    std::vector<M4> tex_inv;           // per pattern
    std::vector<V3> tri_e1, tri_e2, tri_n;  // per TRIANGLE node
    bool ok = true;

    explicit Scene(const rl_scene_desc* desc) : d(desc) {
        int n = d->n_nodes;
        inv.resize(n);
        inv_t.resize(n);
        bounds.resize(n);
        tri_e1.resize(n);
        tri_e2.resize(n);
        tri_n.resize(n);
        tex_inv.resize(d->n_textures);
        for (int i = 0; i < d->n_textures; i++) {
            M4 m = m4_from(d->textures[i].transform);
            if (!invert(m, &tex_inv[i])) ok = false;
        }
        for (int i = 0; i < n; i++) {
            const rl_node& nd = d->nodes[i];
            if (nd.kind == RL_RTC_TRANSFORMED) {
                // Transformed::new (object/transformed.rs:19-26)
                M4 m = m4_from(d->params + nd.param);
                if (!invert(m, &inv[i])) ok = false;
                inv_t[i] = transpose(inv[i]);
            } else if (nd.kind == RL_RTC_TRIANGLE) {
                // Triangle::flat / smooth (object/triangle.rs:30-55)
                const double* p = d->params + nd.param;
                V3 p1 = v3(p[0], p[1], p[2]), p2 = v3(p[3], p[4], p[5]), p3 = v3(p[6], p[7], p[8]);
                tri_e1[i] = p2 - p1;
                tri_e2[i] = p3 - p1;
                if (!(nd.flags & 1)) tri_n[i] = normalized(cross(tri_e2[i], tri_e1[i]));
            }
        }
        // Bounded::new (object/bounded.rs:92-98) captures child.bounds() at construction
        for (int i = 0; i < n; i++)
            if (d->nodes[i].kind == RL_RTC_BOUNDED) bounds[i] = node_bounds(d->nodes[i].child_begin);
    }

    // Bounds::from_points (bounded.rs:34-59)
    static Bounds from_points(const std::vector<V3>& pts) {
        V3 mn = pts[0], mx = pts[0];
        for (const V3& p : pts) {
            mn = {rmin(mn.x, p.x), rmin(mn.y, p.y), rmin(mn.z, p.z)};
            mx = {rmax(mx.x, p.x), rmax(mx.y, p.y), rmax(mx.z, p.z)};
        }
        return {mn, mx};
    }
    // Bounds::from_bounds (bounded.rs:61-76)
    static Bounds from_bounds(const std::vector<Bounds>& bs) {
        std::vector<V3> pts;
        for (const Bounds& b : bs) {
            pts.push_back(b.mn);
            pts.push_back(b.mx);
        }
        if (pts.empty()) return {v3(0, 0, 0), v3(0, 0, 0)};
        return from_points(pts);
    }

    Bounds node_bounds(int id) const {
        const rl_node& nd = d->nodes[id];
        switch (nd.kind) {
            case RL_RTC_SPHERE:  // sphere.rs:61-66
            case RL_RTC_CUBE:    // cube.rs:58-63
                return {v3(-1, -1, -1), v3(1, 1, 1)};
            case RL_RTC_PLANE:  // plane.rs:41-47
                return {v3(-INF, -1e8, -INF), v3(INF, 1e8, INF)};
            case RL_RTC_CYLINDER: {  // cylinder.rs:141-146
                const double* p = d->params + nd.param;
                return {v3(-1, p[0], -1), v3(1, p[1], 1)};
            }
            case RL_RTC_CONE: {  // cone.rs:153-163
                const double* p = d->params + nd.param;
                double radius = rmax(std::fabs(p[1]), std::fabs(p[0]));
                return {v3(-radius, p[0], -radius), v3(radius, p[1], radius)};
            }
            case RL_RTC_TRIANGLE: {  // triangle.rs:103-105
                const double* p = d->params + nd.param;
                return from_points({v3(p[0], p[1], p[2]), v3(p[3], p[4], p[5]), v3(p[6], p[7], p[8])});
            }
            case RL_RTC_TRANSFORMED: {  // transformed.rs:53-57 (uses the FORWARD matrix)
                Bounds b = node_bounds(nd.child_begin);
                M4 m = m4_from(d->params + nd.param);
                // Bounds::enumerate (bounded.rs:17-32)
                V3 c[8] = {v3(b.mn.x, b.mn.y, b.mn.z), v3(b.mn.x, b.mn.y, b.mx.z),
                           v3(b.mn.x, b.mx.y, b.mn.z), v3(b.mn.x, b.mx.y, b.mx.z),
                           v3(b.mx.x, b.mn.y, b.mn.z), v3(b.mx.x, b.mn.y, b.mx.z),
                           v3(b.mx.x, b.mx.y, b.mn.z), v3(b.mx.x, b.mx.y, b.mx.z)};
                std::vector<V3> pts;
                for (const V3& p : c) pts.push_back(mul_point(m, p));
                return from_points(pts);
            }
            case RL_RTC_GROUP: {  // group.rs:44-47
                std::vector<Bounds> bs;
                for (int k = nd.child_begin; k < nd.child_end; k++)
                    bs.push_back(node_bounds(d->children[k]));
                return from_bounds(bs);
            }
            case RL_RTC_BOUNDED:  // bounded.rs:154-156
                return node_bounds(nd.child_begin);
            case RL_RTC_CSG:  // csg.rs:108-110
                return from_bounds({node_bounds(nd.child_begin), node_bounds(nd.child_end)});
        }
        return {v3(0, 0, 0), v3(0, 0, 0)};
    }

    // ---- patterns / surface ------------------------------------------------------------------
    V3 pattern_at(int tex, V3 point) const {
        const rl_texture& t = d->textures[tex];
        V3 p = mul_point(tex_inv[tex], point);  // pattern/mod.rs:9-11
        V3 a = v3(t.a[0], t.a[1], t.a[2]), b = v3(t.b[0], t.b[1], t.b[2]);
        switch (t.kind) {
            case RL_TEX_RTC_STRIPE:  // stripe.rs:21-27
                return ((int64_t)std::floor(p.x) % 2 == 0) ? a : b;
            case RL_TEX_RTC_CHECKER3D:  // checker3d.rs:19-25
                return ((int64_t)(std::floor(p.x) + std::floor(p.y) + std::floor(p.z)) % 2 == 0) ? a : b;
            case RL_TEX_RTC_GRADIENT: {  // gradient.rs:20-25
                V3 distance = b - a;
                double fraction = p.x - std::floor(p.x);
                return a + distance * fraction;
            }
            case RL_TEX_RTC_RING: {  // ring.rs:20-28
                double radius = std::sqrt(p.x * p.x + p.z * p.z);
                return ((int64_t)std::floor(radius) % 2 == 0) ? a : b;
            }
        }
        return a;
    }
    // Surface::color_at (material.rs:13-20)
    V3 surface_color(int mat, V3 p) const {
        const rl_material& m = d->materials[mat];
        if (m.texture < 0) return v3(m.color[0], m.color[1], m.color[2]);
        return pattern_at(m.texture, p);
    }

    // ---- normals (PhysicalObject::normal_at) ---------------------------------------------------
    V3 normal_at(int id, V3 p) const {
        const rl_node& nd = d->nodes[id];
        switch (nd.kind) {
            case RL_RTC_SPHERE:  // sphere.rs:24-28
                return normalized(p - v3(0, 0, 0));
            case RL_RTC_PLANE:  // plane.rs:16-20
                return normalized(v3(0, 1, 0));
            case RL_RTC_CUBE: {  // cube.rs:14-31
                double mc = rmax(rmax(std::fabs(p.x), std::fabs(p.y)), std::fabs(p.z));
                if (mc == std::fabs(p.x)) return normalized(v3(p.x, 0, 0));
                if (mc == std::fabs(p.y)) return normalized(v3(0, p.y, 0));
                return normalized(v3(0, 0, p.z));
            }
            case RL_RTC_CYLINDER: {  // cylinder.rs:67-87
                const double* q = d->params + nd.param;
                bool has_min = q[0] != -INF, has_max = q[1] != INF;
                double dist2 = p.x * p.x + p.z * p.z;
                if (dist2 < 1.0 && has_max && p.y >= q[1] - 1e-8) return normalized(v3(0, 1, 0));
                if (dist2 < 1.0 && has_min && p.y <= q[0] + 1e-8) return normalized(v3(0, -1, 0));
                return normalized(v3(p.x, 0, p.z));
            }
            case RL_RTC_CONE: {  // cone.rs:65-84
                const double* q = d->params + nd.param;
                bool has_min = q[0] != -INF, has_max = q[1] != INF;
                double dist2 = p.x * p.x + p.z * p.z;
                if (has_max && dist2 < q[1] * q[1] && p.y >= q[1] - 1e-8) return normalized(v3(0, 1, 0));
                if (has_min && dist2 < q[0] * q[0] && p.y <= q[0] + 1e-8) return normalized(v3(0, -1, 0));
                double y = std::sqrt(p.x * p.x + p.z * p.z);
                y = p.y > 0.0 ? -y : y;
                return normalized(v3(p.x, y, p.z));
            }
        }
        return v3(0, 0, 0);
    }

    // object/mod.rs:20-32
    Isect basic(const Ray& ray, double t, int id) const {
        V3 p = ray.position(t);
        return {t, id, surface_color(d->nodes[id].material, p), normal_at(id, p)};
    }

    static bool in_bounds(const double* q, double y) {  // cylinder.rs:21-28 / cone.rs:20-27
        bool has_min = q[0] != -INF, has_max = q[1] != INF;
        if (has_min && has_max) return y > q[0] && y < q[1];
        if (has_min) return y > q[0];
        if (has_max) return y < q[1];
        return true;
    }

    // intersect::sort (intersect.rs:170-172) — stable, by t
    static void sort_xs(std::vector<Isect>& xs) {
        std::stable_sort(xs.begin(), xs.end(), [](const Isect& a, const Isect& b) { return a.t < b.t; });
    }

    static void check_axis_cube(double origin, double direction, double* tmin, double* tmax) {  // cube.rs:66-79
        double a = (-1.0 - origin) / direction, b = (1.0 - origin) / direction;
        if (a > b) std::swap(a, b);
        *tmin = a;
        *tmax = b;
    }
    static void check_axis_bounds(double mn, double mx, double origin, double speed, double* tmin,
                                  double* tmax) {  // bounded.rs:132-144
        double a = (mn - origin) / speed, b = (mx - origin) / speed;
        if (a > b) std::swap(a, b);
        *tmin = a;
        *tmax = b;
    }

    // Object::intersect for every node kind
    std::vector<Isect> intersect(int id, const Ray& ray) const {
        const rl_node& nd = d->nodes[id];
        std::vector<Isect> out;
        switch (nd.kind) {
            case RL_RTC_SPHERE: {  // sphere.rs:35-59
                V3 sphere_to_ray = ray.origin - v3(0, 0, 0);
                double a = dot(ray.direction, ray.direction);
                double b = 2.0 * dot(ray.direction, sphere_to_ray);
                double c = dot(sphere_to_ray, sphere_to_ray) - 1.0;
                double disc = b * b - 4.0 * a * c;
                if (disc < 0.0) return out;
                double s = std::sqrt(disc);
                double t1 = (-b - s) / (2.0 * a), t2 = (-b + s) / (2.0 * a);
                out.push_back(basic(ray, t1, id));
                out.push_back(basic(ray, t2, id));
                return out;
            }
            case RL_RTC_PLANE: {  // plane.rs:26-39
                if (std::fabs(ray.direction.y) < 1e-8) return out;
                out.push_back(basic(ray, -ray.origin.y / ray.direction.y, id));
                return out;
            }
            case RL_RTC_CUBE: {  // cube.rs:38-56
                double x0, x1, y0, y1, z0, z1;
                check_axis_cube(ray.origin.x, ray.direction.x, &x0, &x1);
                check_axis_cube(ray.origin.y, ray.direction.y, &y0, &y1);
                check_axis_cube(ray.origin.z, ray.direction.z, &z0, &z1);
                double tmin = rmax(rmax(x0, y0), z0), tmax = rmin(rmin(x1, y1), z1);
                if (tmin > tmax) return out;
                out.push_back(basic(ray, tmin, id));
                out.push_back(basic(ray, tmax, id));
                return out;
            }
            case RL_RTC_CYLINDER: {  // cylinder.rs:94-139, caps 30-64
                const double* q = d->params + nd.param;
                bool closed = nd.flags & 1;
                std::vector<double> ts;
                double a = ray.direction.x * ray.direction.x + ray.direction.z * ray.direction.z;
                if (!(std::fabs(a) < 1e-8)) {
                    double b = 2.0 * ray.origin.x * ray.direction.x + 2.0 * ray.origin.z * ray.direction.z;
                    double c = ray.origin.x * ray.origin.x + ray.origin.z * ray.origin.z - 1.0;
                    double disc = b * b - 4.0 * a * c;
                    if (!(disc < 0.0)) {
                        double t0 = (-b - std::sqrt(disc)) / (2.0 * a);
                        double t1 = (-b + std::sqrt(disc)) / (2.0 * a);
                        double y0 = ray.origin.y + t0 * ray.direction.y;
                        if (in_bounds(q, y0)) ts.push_back(t0);
                        double y1 = ray.origin.y + t1 * ray.direction.y;
                        if (in_bounds(q, y1)) ts.push_back(t1);
                    }
                }
                if (closed && !(std::fabs(ray.direction.y) < 1e-8)) {
                    for (int k = 0; k < 2; k++) {
                        if ((k == 0 && q[0] == -INF) || (k == 1 && q[1] == INF)) continue;
                        double t = (q[k] - ray.origin.y) / ray.direction.y;
                        double x = ray.origin.x + t * ray.direction.x;
                        double z = ray.origin.z + t * ray.direction.z;
                        if (x * x + z * z <= 1.0) ts.push_back(t);
                    }
                }
                for (double t : ts) out.push_back(basic(ray, t, id));
                return out;
            }
            case RL_RTC_CONE: {  // cone.rs:88-151, caps 29-63 (cap radius test is `<= |y|`, as written)
                const double* q = d->params + nd.param;
                bool closed = nd.flags & 1;
                const V3 &o = ray.origin, &dr = ray.direction;
                double a = dr.x * dr.x - dr.y * dr.y + dr.z * dr.z;
                double b = 2.0 * o.x * dr.x - 2.0 * o.y * dr.y + 2.0 * o.z * dr.z;
                double c = o.x * o.x - o.y * o.y + o.z * o.z;
                bool a0 = std::fabs(a) < 1e-8, b0 = std::fabs(b) < 1e-8;
                std::vector<double> ts;
                if (a0 && b0) {
                } else if (a0 && !b0) {
                    ts.push_back(-c / (2.0 * b));
                } else {
                    double disc = b * b - 4.0 * a * c;
                    if (!(disc < 0.0)) {
                        double t0 = (-b - std::sqrt(disc)) / (2.0 * a);
                        double t1 = (-b + std::sqrt(disc)) / (2.0 * a);
                        double y0 = o.y + t0 * dr.y;
                        if (in_bounds(q, y0)) ts.push_back(t0);
                        double y1 = o.y + t1 * dr.y;
                        if (in_bounds(q, y1)) ts.push_back(t1);
                    }
                }
                if (closed && !(std::fabs(dr.y) < 1e-8)) {
                    for (int k = 0; k < 2; k++) {
                        if ((k == 0 && q[0] == -INF) || (k == 1 && q[1] == INF)) continue;
                        double t = (q[k] - o.y) / dr.y;
                        double x = o.x + t * dr.x, z = o.z + t * dr.z;
                        if (x * x + z * z <= std::fabs(q[k])) ts.push_back(t);
                    }
                }
                for (double t : ts) out.push_back(basic(ray, t, id));
                return out;
            }
            case RL_RTC_TRIANGLE: {  // triangle.rs:63-101
                const double* p = d->params + nd.param;
                V3 p1 = v3(p[0], p[1], p[2]);
                V3 e1 = tri_e1[id], e2 = tri_e2[id];
                V3 dir_cross_e2 = cross(ray.direction, e2);
                double det = dot(e1, dir_cross_e2);
                if (std::fabs(det) < 1e-8) return out;
                double f = 1.0 / det;
                V3 p1_to_origin = ray.origin - p1;
                double u = f * dot(p1_to_origin, dir_cross_e2);
                if (!(u >= 0.0 && u <= 1.0)) return out;
                V3 origin_cross_e1 = cross(p1_to_origin, e1);
                double v = f * dot(ray.direction, origin_cross_e1);
                if (v < 0.0 || (u + v) > 1.0) return out;
                double t = f * dot(e2, origin_cross_e1);
                V3 pt = ray.position(t);
                V3 color = surface_color(nd.material, pt);
                V3 normal;
                if (nd.flags & 1) {
                    V3 n1 = v3(p[9], p[10], p[11]), n2 = v3(p[12], p[13], p[14]), n3 = v3(p[15], p[16], p[17]);
                    normal = normalized((n2 * u + n3 * v) + n1 * (1.0 - u - v));
                } else {
                    normal = tri_n[id];
                }
                out.push_back({t, id, color, normal});
                return out;
            }
            case RL_RTC_TRANSFORMED: {  // transformed.rs:39-51, ray.rs:18-20
                Ray local{mul_point(inv[id], ray.origin), mul_vec(inv[id], ray.direction)};
                out = intersect(nd.child_begin, local);
                for (Isect& x : out) x.normal = normalized(mul_vec(inv_t[id], x.normal));
                return out;
            }
            case RL_RTC_GROUP: {  // group.rs:29-42
                for (int k = nd.child_begin; k < nd.child_end; k++) {
                    std::vector<Isect> c = intersect(d->children[k], ray);
                    out.insert(out.end(), c.begin(), c.end());
                }
                sort_xs(out);
                return out;
            }
            case RL_RTC_BOUNDED: {  // bounded.rs:100-152
                const Bounds& b = bounds[id];
                double x0, x1, y0, y1, z0, z1;
                check_axis_bounds(b.mn.x, b.mx.x, ray.origin.x, ray.direction.x, &x0, &x1);
                check_axis_bounds(b.mn.y, b.mx.y, ray.origin.y, ray.direction.y, &y0, &y1);
                check_axis_bounds(b.mn.z, b.mx.z, ray.origin.z, ray.direction.z, &z0, &z1);
                double tmin = rmax(rmax(x0, y0), z0), tmax = rmin(rmin(x1, y1), z1);
                if (tmin <= tmax) return intersect(nd.child_begin, ray);
                return out;
            }
            case RL_RTC_CSG: {  // csg.rs:81-106, filter 49-72, truth table 15-28
                std::vector<Isect> l = intersect(nd.child_begin, ray), r = intersect(nd.child_end, ray);
                struct Sided {
                    Isect i;
                    bool left;
                };
                std::vector<Sided> all;
                for (const Isect& i : l) all.push_back({i, true});
                for (const Isect& i : r) all.push_back({i, false});
                std::stable_sort(all.begin(), all.end(),
                                 [](const Sided& a, const Sided& b) { return a.i.t < b.i.t; });
                bool in_l = false, in_r = false;
                for (const Sided& s : all) {
                    bool allowed;
                    switch (nd.flags) {
                        case RL_CSG_UNION: allowed = (s.left && !in_r) || (!s.left && !in_l); break;
                        case RL_CSG_INTERSECTION: allowed = (s.left && in_r) || (!s.left && in_l); break;
                        default: allowed = (s.left && !in_r) || (!s.left && in_l); break;
                    }
                    if (s.left) in_l = !in_l; else in_r = !in_r;
                    if (allowed) out.push_back(s.i);
                }
                return out;
            }
        }
        return out;
    }

    // World::intersect (world.rs:46-55)
    std::vector<Isect> world_intersect(const Ray& ray) const {
        std::vector<Isect> xs;
        for (int k = 0; k < d->n_roots; k++) {
            std::vector<Isect> c = intersect(d->roots[k], ray);
            xs.insert(xs.end(), c.begin(), c.end());
        }
        sort_xs(xs);
        return xs;
    }

    // intersect::hit (intersect.rs:159-168): lowest t >= 0; on equal t the LATER element wins
    static int hit(const std::vector<Isect>& xs) {
        int acc = -1;
        for (int i = 0; i < (int)xs.size(); i++) {
            if (xs[i].t >= 0.0) {
                if (acc < 0) acc = i;
                else acc = (xs[acc].t < xs[i].t) ? acc : i;
            }
        }
        return acc;
    }

    struct Comps {  // intersect.rs:123-137
        double t;
        int object;
        V3 point, eye_v, normal_v;
        bool inside;
        V3 over_point, under_point, reflect_v;
        double n1, n2;
        V3 object_color;
    };

    // prepare_computations_helper (intersect.rs:47-115)
    Comps prepare(const Isect& h, const Ray& ray, const std::vector<Isect>& xs) const {
        Comps c;
        c.t = h.t;
        c.object = h.object;
        c.point = ray.position(h.t);
        c.eye_v = normalized(-ray.direction);
        V3 normal_v = h.normal;
        double nde = dot(normal_v, c.eye_v);
        if (nde < 0.0) {
            c.normal_v = -normal_v;
            c.inside = true;
        } else {
            c.normal_v = normal_v;
            c.inside = false;
        }
        c.over_point = c.point + c.normal_v * 1e-5;
        c.under_point = c.point - c.normal_v * 1e-5;
        c.reflect_v = normalized(reflect(ray.direction, c.normal_v));
        std::vector<int> containers;
        double n1 = 1.0, n2 = 1.0;
        for (const Isect& i : xs) {
            bool is_hit = are_equal(i.t, h.t) && i.object == h.object;  // intersect.rs:117-121
            if (is_hit) n1 = containers.empty() ? 1.0 : mat(containers.back()).refractive_index;
            auto it = std::find(containers.begin(), containers.end(), i.object);
            if (it != containers.end()) containers.erase(it);
            else containers.push_back(i.object);
            if (is_hit) {
                n2 = containers.empty() ? 1.0 : mat(containers.back()).refractive_index;
                break;
            }
        }
        c.n1 = n1;
        c.n2 = n2;
        c.object_color = h.color;
        return c;
    }

    const rl_material& mat(int object) const { return d->materials[d->nodes[object].material]; }

    // Precomputation::schlick (intersect.rs:140-156)
    static double schlick(const Comps& c) {
        double cosv = dot(c.eye_v, c.normal_v);
        double n = c.n1 / c.n2;
        double sin2_t = n * n * (1.0 - cosv * cosv);
        double cos_t = std::sqrt(1.0 - sin2_t);
        double cos_adj = n > 1.0 ? cos_t : cosv;
        if (sin2_t > 1.0 && n > 1.0) return 1.0;
        double r0 = (c.n1 - c.n2) / (c.n1 + c.n2);
        r0 = r0 * r0;
        return r0 + (1.0 - r0) * powi5(1.0 - cos_adj);
    }

    // material::lighting (material.rs:54-90)
    V3 lighting(const rl_material& m, V3 point, V3 object_color, const rl_light& light, V3 eyev,
                V3 normalv, double shadow_attenuation) const {
        V3 li = v3(light.intensity[0], light.intensity[1], light.intensity[2]);
        V3 lp = v3(light.position[0], light.position[1], light.position[2]);
        V3 effective = mulc(object_color, li);
        V3 lightv = normalized(lp - point);
        V3 ambient = effective * m.ambient;
        double ldn = dot(lightv, normalv);
        V3 diffuse = v3(0, 0, 0), specular = v3(0, 0, 0);
        if (!(ldn < 0.0)) {
            V3 diff = (effective * m.diffuse) * ldn;
            V3 reflectv = -reflect(lightv, normalv);
            double rde = dot(reflectv, eyev);
            diffuse = diff * shadow_attenuation;
            if (!(rde <= 0.0)) {
                double factor = std::pow(rde, m.shininess);
                specular = li * (m.specular * factor * shadow_attenuation);
            }
        }
        return (ambient + diffuse) + specular;
    }

    // World::shadow_attenuation (world.rs:104-126)
    double shadow_attenuation(V3 point, const rl_light& light) const {
        V3 v = v3(light.position[0], light.position[1], light.position[2]) - point;
        double distance = mag(v);
        V3 dir;
        if (!normalize(v, &dir)) return 1.0;
        Ray r{point, dir};
        std::vector<Isect> xs = world_intersect(r);
        std::vector<int> seen;
        double prod = 1.0;
        for (const Isect& i : xs) {
            if (!(i.t > 0.0 && i.t < distance)) continue;
            if (std::find(seen.begin(), seen.end(), i.object) != seen.end()) break;  // take_while
            seen.push_back(i.object);
            prod = prod * mat(i.object).transparency;
        }
        return prod;
    }

    V3 void_color() const { return v3(d->void_color[0], d->void_color[1], d->void_color[2]); }

    // World::reflected_color (world.rs:128-136)
    V3 reflected_color(const Comps& c, int remaining) const {
        const rl_material& m = mat(c.object);
        if (remaining == 0 || m.reflectivity == 0.0) return v3(0, 0, 0);
        Ray r{c.over_point, c.reflect_v};
        return color_at_internal(r, remaining - 1) * m.reflectivity;
    }
    // World::refracted_color (world.rs:138-159)
    V3 refracted_color(const Comps& c, int remaining) const {
        const rl_material& m = mat(c.object);
        if (remaining == 0 || m.transparency == 0.0) return v3(0, 0, 0);
        double n_ratio = c.n1 / c.n2;
        double cos_i = dot(c.eye_v, c.normal_v);
        double sin2_t = n_ratio * n_ratio * (1.0 - cos_i * cos_i);
        if (sin2_t > 1.0) return v3(0, 0, 0);
        double cos_t = std::sqrt(1.0 - sin2_t);
        V3 direction = c.normal_v * (n_ratio * cos_i - cos_t) - c.eye_v * n_ratio;
        Ray r{c.under_point, direction};
        return color_at_internal(r, remaining - 1) * m.transparency;
    }

    // World::shade_hit (world.rs:57-87); returns false for `None` (no lights)
    bool shade_hit(const Comps& c, int remaining, V3* out) const {
        if (d->n_lights == 0) return false;
        V3 acc = v3(0, 0, 0);
        const rl_material& m = mat(c.object);
        for (int l = 0; l < d->n_lights; l++) {
            double sa = shadow_attenuation(c.over_point, d->lights[l]);
            V3 surface = lighting(m, c.point, c.object_color, d->lights[l], c.eye_v, c.normal_v, sa);
            V3 reflected = reflected_color(c, remaining);
            V3 refracted = refracted_color(c, remaining);
            V3 col;
            if (m.reflectivity > 0.0 && m.transparency > 0.0) {
                double reflectance = schlick(c);
                col = surface + (reflected * reflectance + refracted * (1.0 - reflectance));
            } else {
                col = surface + (reflected + refracted);
            }
            acc = (l == 0) ? col : acc + col;
        }
        *out = acc;
        return true;
    }

    // World::color_at_internal (world.rs:89-98)
    V3 color_at_internal(const Ray& ray, int remaining) const {
        std::vector<Isect> xs = world_intersect(ray);
        int h = hit(xs);
        if (h < 0) return void_color();
        Comps c = prepare(xs[h], ray, xs);
        V3 out;
        if (!shade_hit(c, remaining, &out)) return void_color();
        return out;
    }
    V3 color_at(const Ray& ray) const { return color_at_internal(ray, d->max_reflection_depth); }
};

// Camera (scene/camera.rs:35-124)
struct Camera {
    int hsize, vsize;
    double pixel_size, half_width, half_height;
    M4 inverse;
    bool ok;
    explicit Camera(const rl_rtc_camera* c) {
        hsize = c->hsize;
        vsize = c->vsize;
        double half_view = std::tan(c->fov / 2.0);
        double aspect = (double)hsize / (double)vsize;
        if (aspect >= 1.0) {
            half_width = half_view;
            half_height = half_view / aspect;
        } else {
            half_width = half_view * aspect;
            half_height = half_view;
        }
        pixel_size = half_width * 2.0 / (double)hsize;
        ok = invert(m4_from(c->transform), &inverse);
    }
    // rays_for_pixel (camera.rs:63-91), sample (nx, ny)
    Ray ray(int px, int py, int samples, int nx, int ny) const {
        double sample_offset = 1.0 / (double)samples;
        double xoffset = ((double)px + sample_offset * ((double)nx + 0.5)) * pixel_size;
        double yoffset = ((double)py + sample_offset * ((double)ny + 0.5)) * pixel_size;
        double world_x = half_width - xoffset, world_y = half_height - yoffset;
        V3 pixel = mul_point(inverse, v3(world_x, world_y, -1.0));
        V3 origin = mul_point(inverse, v3(0, 0, 0));
        return {origin, normalized(pixel - origin)};
    }
};

}  // namespace

extern "C" {

// Camera::render (camera.rs:93-124). out_rgb: vsize*hsize*3 f64, idx = width*y + x (canvas.rs:42-48).
// threads <= 0: all cores.  OpenMP dynamic schedule over columns x = the rayon analogue (camera.rs:96-98).
int orc_rtc_render(const rl_scene_desc* desc, const rl_rtc_camera* cam, uint32_t aa, double* out_rgb,
                   int threads) {
    if (!desc || !cam || !out_rgb || aa < 1 || desc->flavor != RL_FLAVOR_RTC) return -1;
    Scene sc(desc);
    Camera c(cam);
    if (!sc.ok || !c.ok) return -1;
    int s = (int)aa;
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : omp_get_max_threads_compat())
#endif
    for (int x = 0; x < c.hsize; x++) {
        for (int y = 0; y < c.vsize; y++) {
            V3 acc = v3(0, 0, 0);
            bool first = true;
            for (int nx = 0; nx < s; nx++)
                for (int ny = 0; ny < s; ny++) {
                    V3 col = sc.color_at(c.ray(x, y, s, nx, ny));
                    acc = first ? col : acc + col;
                    first = false;
                }
            V3 px = acc * (1.0 / (double)(s * s));
            double* o = out_rgb + ((size_t)c.hsize * y + x) * 3;
            o[0] = px.x;
            o[1] = px.y;
            o[2] = px.z;
        }
    }
    return 0;
}

// camera rays of a pixel range, for identical-ray-batch parity tests. rays: n*6 (origin, direction)
int orc_rtc_camera_rays(const rl_rtc_camera* cam, uint32_t aa, double* rays) {
    Camera c(cam);
    if (!c.ok) return -1;
    size_t k = 0;
    for (int y = 0; y < c.vsize; y++)
        for (int x = 0; x < c.hsize; x++)
            for (int nx = 0; nx < (int)aa; nx++)
                for (int ny = 0; ny < (int)aa; ny++) {
                    Ray r = c.ray(x, y, (int)aa, nx, ny);
                    double* o = rays + 6 * k++;
                    o[0] = r.origin.x; o[1] = r.origin.y; o[2] = r.origin.z;
                    o[3] = r.direction.x; o[4] = r.direction.y; o[5] = r.direction.z;
                }
    return 0;
}

// closest hit per intersect::hit over World::intersect for a batch of rays (n*6 doubles).
// margin[i] (optional) = distance in t to the runner-up candidate / nearest rejected root, used by the
// parity test to recognise rays that are degenerate in f32 (grazing / exact ties).
int orc_rtc_trace(const rl_scene_desc* desc, const double* rays, uint64_t n, int32_t* node, double* t,
                  double* second_t, int threads) {
    if (!desc || desc->flavor != RL_FLAVOR_RTC) return -1;
    Scene sc(desc);
    if (!sc.ok) return -1;
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 256) num_threads(threads > 0 ? threads : omp_get_max_threads_compat())
#endif
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const double* r = rays + 6 * i;
        Ray ray{v3(r[0], r[1], r[2]), v3(r[3], r[4], r[5])};
        std::vector<Isect> xs = sc.world_intersect(ray);
        int h = Scene::hit(xs);
        node[i] = h < 0 ? -1 : xs[h].object;
        t[i] = h < 0 ? INF : xs[h].t;
        if (second_t) {
            double best = INF;  // nearest other candidate with t >= 0
            for (int k = 0; k < (int)xs.size(); k++)
                if (k != h && xs[k].t >= 0.0) best = std::fmin(best, xs[k].t);
            second_t[i] = best;
        }
    }
    return 0;
}

// World::intersect for one ray: sorted t's and leaf ids (unit known-answer tests). returns count.
int orc_rtc_intersect(const rl_scene_desc* desc, const double* ray6, int cap, double* ts, int32_t* nodes,
                      double* normals, double* colors) {
    Scene sc(desc);
    if (!sc.ok) return -1;
    Ray ray{v3(ray6[0], ray6[1], ray6[2]), v3(ray6[3], ray6[4], ray6[5])};
    std::vector<Isect> xs = sc.world_intersect(ray);
    int n = (int)xs.size();
    for (int i = 0; i < n && i < cap; i++) {
        ts[i] = xs[i].t;
        nodes[i] = xs[i].object;
        if (normals) { normals[3 * i] = xs[i].normal.x; normals[3 * i + 1] = xs[i].normal.y; normals[3 * i + 2] = xs[i].normal.z; }
        if (colors) { colors[3 * i] = xs[i].color.x; colors[3 * i + 1] = xs[i].color.y; colors[3 * i + 2] = xs[i].color.z; }
    }
    return n;
}

// World::color_at for one ray
int orc_rtc_color_at(const rl_scene_desc* desc, const double* ray6, int remaining, double* out3) {
    Scene sc(desc);
    if (!sc.ok) return -1;
    Ray ray{v3(ray6[0], ray6[1], ray6[2]), v3(ray6[3], ray6[4], ray6[5])};
    V3 c = remaining < 0 ? sc.color_at(ray) : sc.color_at_internal(ray, remaining);
    out3[0] = c.x; out3[1] = c.y; out3[2] = c.z;
    return 0;
}

// prepare_computations + schlick for the index-th intersection (index < 0: the hit) of one ray.
// out[0..27]: t, object, point3, eye3, normal3, inside, over3, under3, reflect3, n1, n2, schlick, shadow(light0 @ over)
int orc_rtc_prepare(const rl_scene_desc* desc, const double* ray6, int index, double* out) {
    Scene sc(desc);
    if (!sc.ok) return -1;
    Ray ray{v3(ray6[0], ray6[1], ray6[2]), v3(ray6[3], ray6[4], ray6[5])};
    std::vector<Isect> xs = sc.world_intersect(ray);
    int h = index < 0 ? Scene::hit(xs) : index;
    if (h < 0 || h >= (int)xs.size()) return -2;
    Scene::Comps c = sc.prepare(xs[h], ray, xs);
    int k = 0;
    out[k++] = c.t; out[k++] = c.object;
    for (V3 v : {c.point, c.eye_v, c.normal_v}) { out[k++] = v.x; out[k++] = v.y; out[k++] = v.z; }
    out[k++] = c.inside ? 1.0 : 0.0;
    for (V3 v : {c.over_point, c.under_point, c.reflect_v}) { out[k++] = v.x; out[k++] = v.y; out[k++] = v.z; }
    out[k++] = c.n1; out[k++] = c.n2; out[k++] = Scene::schlick(c);
    out[k++] = desc->n_lights > 0 ? sc.shadow_attenuation(c.over_point, desc->lights[0]) : 1.0;
    return 0;
}

// material::lighting known-answer hook: point3, color3, eye3, normal3, attenuation -> rgb
int orc_rtc_lighting(const rl_scene_desc* desc, int material, int light, const double* in13, double* out3) {
    Scene sc(desc);
    V3 c = sc.lighting(desc->materials[material], v3(in13[0], in13[1], in13[2]), v3(in13[3], in13[4], in13[5]),
                       desc->lights[light], v3(in13[6], in13[7], in13[8]), v3(in13[9], in13[10], in13[11]), in13[12]);
    out3[0] = c.x; out3[1] = c.y; out3[2] = c.z;
    return 0;
}

// World::shadow_attenuation at an arbitrary point
double orc_rtc_shadow(const rl_scene_desc* desc, const double* p3, int light) {
    Scene sc(desc);
    return sc.shadow_attenuation(v3(p3[0], p3[1], p3[2]), desc->lights[light]);
}

// cofactor inverse (matrix.rs:68-86); returns 0 and fills out16, or -1 when singular
int orc_rtc_invert(const double* m16, double* out16) {
    M4 o;
    if (!invert(m4_from(m16), &o)) return -1;
    std::memcpy(out16, o.m, sizeof(o.m));
    return 0;
}

}  // extern "C"
