// Scene flattener: lowers the reference's object tree (rl_scene_desc) to the SoA device layout.
//
//  * RTC: every leaf gets ONE composed, pre-inverted affine transform
//         inv_total = inv_leaf * ... * inv_root   (Transformed::intersect, RTC/src/scene/object/transformed.rs:39-51,
//         applies inv per nesting level; the normal matrix is inv_total^T — proved equivalent by the
//         reference's own test at transformed.rs:226-278).  Triangles are baked to world space
//         (t is preserved because ray directions are never renormalised, ray.rs:18-20).
//         Group / Bounded only order or skip work and are flattened away (group.rs:29-42, bounded.rs:146-152).
//  * OW : Transform (3x3 rotate / uniform scale) and Translate chains are baked into world-space
//         triangles, quads and spheres (hittable/transform.rs:145-164, translate.rs:14-21); Bvh / list
//         nodes are replaced by one LBVH over all bounded primitives (closest-hit semantics, bvh.rs:81-90).
//
// All composition / inversion is done in f64 and rounded to f32 once, at the end.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "scene.h"

namespace rl {
namespace {

struct Aff {  // 3x4 affine, f64
    double m[3][4];
};
Aff aff_identity() {
    Aff a{};
    a.m[0][0] = a.m[1][1] = a.m[2][2] = 1.0;
    return a;
}
Aff aff_mul(const Aff& a, const Aff& b) {  // a * b
    Aff r{};
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 4; j++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += a.m[i][k] * b.m[k][j];
            if (j == 3) s += a.m[i][3];
            r.m[i][j] = s;
        }
    }
    return r;
}
// inverse of an affine map [A | t]: [A^-1 | -A^-1 t], A^-1 by the adjugate
bool aff_inverse(const Aff& a, Aff* out) {
    const double(*m)[4] = a.m;
    double c00 = m[1][1] * m[2][2] - m[1][2] * m[2][1];
    double c01 = m[1][2] * m[2][0] - m[1][0] * m[2][2];
    double c02 = m[1][0] * m[2][1] - m[1][1] * m[2][0];
    double det = m[0][0] * c00 + m[0][1] * c01 + m[0][2] * c02;
    if (det == 0.0 || !std::isfinite(det)) return false;
    double id = 1.0 / det;
    double inv[3][3];
    inv[0][0] = c00 * id;
    inv[1][0] = c01 * id;
    inv[2][0] = c02 * id;
    inv[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * id;
    inv[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * id;
    inv[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * id;
    inv[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * id;
    inv[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * id;
    inv[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * id;
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) out->m[i][j] = inv[i][j];
        out->m[i][3] = -(inv[i][0] * m[0][3] + inv[i][1] * m[1][3] + inv[i][2] * m[2][3]);
    }
    return true;
}
bool aff_from_4x4(const double* p, Aff* out, std::string* err) {
    if (p[12] != 0.0 || p[13] != 0.0 || p[14] != 0.0 || p[15] != 1.0) {
        *err = "projective 4x4 transform (bottom row != 0 0 0 1) is not supported";
        return false;
    }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 4; j++) out->m[i][j] = p[i * 4 + j];
    return true;
}
void aff_point(const Aff& a, const double p[3], double o[3]) {
    for (int i = 0; i < 3; i++) o[i] = a.m[i][0] * p[0] + a.m[i][1] * p[1] + a.m[i][2] * p[2] + a.m[i][3];
}
void aff_vec(const Aff& a, const double p[3], double o[3]) {
    for (int i = 0; i < 3; i++) o[i] = a.m[i][0] * p[0] + a.m[i][1] * p[1] + a.m[i][2] * p[2];
}
// n_world = inv^T * n_local
void aff_normal(const Aff& inv, const double n[3], double o[3]) {
    for (int i = 0; i < 3; i++) o[i] = inv.m[0][i] * n[0] + inv.m[1][i] * n[1] + inv.m[2][i] * n[2];
}
void store_rows(const Aff& a, float4 r[3]) {
    for (int i = 0; i < 3; i++) r[i] = make_float4((float)a.m[i][0], (float)a.m[i][1], (float)a.m[i][2], (float)a.m[i][3]);
}
inline float as_f(int v) {
    float f;
    std::memcpy(&f, &v, 4);
    return f;
}
bool normalize3(double v[3]) {
    double m = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (!(m > 0.0)) return false;
    v[0] /= m; v[1] /= m; v[2] /= m;
    return true;
}

// conservative f32 box around f64 extents (so rounding to f32 never shrinks a box)
void push_aabb(FlatScene* fs, const double lo[3], const double hi[3], int ref, int node) {
    for (int k = 0; k < 3; k++) {
        float f = (float)lo[k];
        if ((double)f > lo[k]) f = std::nextafterf(f, -std::numeric_limits<float>::infinity());
        // thin / touching geometry: pad by 4 ulp so slab tests stay conservative in f32
        float pad = 4.0f * 1.1920929e-7f * std::fmax(std::fabs(f), 1e-3f);
        fs->bvh_aabb.push_back(f - pad);
    }
    for (int k = 0; k < 3; k++) {
        float f = (float)hi[k];
        if ((double)f < hi[k]) f = std::nextafterf(f, std::numeric_limits<float>::infinity());
        float pad = 4.0f * 1.1920929e-7f * std::fmax(std::fabs(f), 1e-3f);
        fs->bvh_aabb.push_back(f + pad);
    }
    fs->bvh_ref.push_back(ref);
    fs->bvh_node_id.push_back(node);
}

constexpr int MESH_MARK = 0x7000000;  // leaf-ref index of the placeholder an OW mesh use leaves in the LBVH input

struct Flattener {
    const rl_scene_desc* d;
    FlatScene* fs;
    std::string* err;
    const MeshMeta* mesh = nullptr;
    int rc = RL_OK;
    int csg_depth = 0;
    // while lowering the boundary of a ConstantMedium, leaves go to medium_refs instead of the LBVH input
    bool in_boundary = false;
    double blo[3], bhi[3];

    void emit(const double lo[3], const double hi[3], int ref, int node) {
        if (in_boundary) {
            fs->medium_refs.push_back(ref);
            for (int k = 0; k < 3; k++) {
                blo[k] = std::fmin(blo[k], lo[k]);
                bhi[k] = std::fmax(bhi[k], hi[k]);
            }
        } else {
            push_aabb(fs, lo, hi, ref, node);
        }
    }

    bool fail(int code, const std::string& msg) {
        if (rc == RL_OK) {
            rc = code;
            *err = msg;
        }
        return false;
    }
    bool check_node(int id) {
        if (id < 0 || id >= d->n_nodes) return fail(RL_E_INVALID, "node id out of range");
        return true;
    }
    const double* params(const rl_node& nd, int need) {
        if (nd.param < 0 || (int64_t)nd.param + need > d->n_params) {
            fail(RL_E_INVALID, "node params out of range");
            return nullptr;
        }
        return d->params + nd.param;
    }

    // ---- materials / textures ------------------------------------------------------------------
    bool lower_tables() {
        for (int i = 0; i < d->n_textures; i++) {
            const rl_texture& t = d->textures[i];
            const bool rtc_tex = t.kind >= RL_TEX_RTC_STRIPE && t.kind <= RL_TEX_RTC_RING;
            const bool ow_tex = t.kind >= RL_TEX_OW_SOLID && t.kind <= RL_TEX_OW_NOISE;
            if (!(d->flavor == RL_FLAVOR_RTC ? rtc_tex : ow_tex))
                return fail(RL_E_INVALID, "unknown texture kind (or a kind of the other renderer)");
            DevTexture o{};
            o.a = make_float4((float)t.a[0], (float)t.a[1], (float)t.a[2], as_f(t.kind));
            float inv_scale = t.kind == RL_TEX_OW_CHECKER ? (float)(1.0 / t.scale) : 0.0f;
            o.b = make_float4((float)t.b[0], (float)t.b[1], (float)t.b[2], inv_scale);
            o.idx = make_int4(t.tex_a, t.tex_b, t.image, 0);
            if (t.kind == RL_TEX_OW_CHECKER &&
                (t.tex_a < 0 || t.tex_a >= d->n_textures || t.tex_b < 0 || t.tex_b >= d->n_textures))
                return fail(RL_E_INVALID, "checker sub-texture out of range");
            if (t.kind == RL_TEX_OW_IMAGE && (t.image < 0 || t.image >= d->n_images))
                return fail(RL_E_INVALID, "image index out of range");
            if (t.kind == RL_TEX_OW_NOISE) {  // texture.rs:84-94: b.w = scale, idx.z = Perlin table
                if (t.image < 0 || t.image >= d->n_perlins) return fail(RL_E_INVALID, "perlin index out of range");
                o.b.w = (float)t.scale;
            }
            fs->textures.push_back(o);
        }
        for (int i = 0; i < d->n_images; i++) {
            const rl_image& im = d->images[i];
            if (im.width <= 0 || im.height <= 0 || !im.rgb) return fail(RL_E_INVALID, "Image has no data");
            FlatScene::Img o;
            o.w = im.width;
            o.h = im.height;
            o.texels.resize((size_t)im.width * im.height);
            for (size_t k = 0; k < o.texels.size(); k++)
                o.texels[k] = make_float4(im.rgb[3 * k], im.rgb[3 * k + 1], im.rgb[3 * k + 2], 0.0f);
            fs->images.push_back(std::move(o));
        }
        for (int i = 0; i < d->n_perlins; i++) {  // perlin.rs:9-14
            const rl_perlin& pn = d->perlins[i];
            for (int k = 0; k < 256; k++)
                fs->perlin_vec.push_back(make_float4((float)pn.randvec[k][0], (float)pn.randvec[k][1], (float)pn.randvec[k][2], 0.0f));
            for (int k = 0; k < 256; k++) fs->perlin_perm.push_back(pn.perm_x[k] & 255);
            for (int k = 0; k < 256; k++) fs->perlin_perm.push_back(pn.perm_y[k] & 255);
            for (int k = 0; k < 256; k++) fs->perlin_perm.push_back(pn.perm_z[k] & 255);
        }
        for (int i = 0; i < d->n_materials; i++) {
            const rl_material& m = d->materials[i];
            const bool ow_mat = m.kind >= RL_MAT_OW_LAMBERTIAN && m.kind <= RL_MAT_OW_ISOTROPIC;
            if (!(d->flavor == RL_FLAVOR_RTC ? m.kind == RL_MAT_RTC_PHONG : ow_mat))
                return fail(RL_E_INVALID, "unknown material kind (or a kind of the other renderer)");
            DevMaterial o{};
            if (m.texture >= d->n_textures) return fail(RL_E_INVALID, "material texture out of range");
            o.color = make_float4((float)m.color[0], (float)m.color[1], (float)m.color[2], as_f(m.texture));
            if (m.kind == RL_MAT_RTC_PHONG) {
                o.a = make_float4((float)m.ambient, (float)m.diffuse, (float)m.specular, (float)m.shininess);
                o.b = make_float4((float)m.reflectivity, (float)m.transparency, (float)m.refractive_index,
                                  as_f(m.kind));
                if (m.transparency > 0.0) fs->has_transparency = 1;
            } else {
                o.a = make_float4((float)m.fuzz, (float)m.refractive_index, 0.0f, 0.0f);
                o.b = make_float4(0.0f, 0.0f, 0.0f, as_f(m.kind));
                if ((m.kind == RL_MAT_OW_LAMBERTIAN || m.kind == RL_MAT_OW_DIFFUSE_LIGHT || m.kind == RL_MAT_OW_ISOTROPIC) && m.texture < 0)
                    return fail(RL_E_INVALID, "OW material without a texture");
            }
            fs->materials.push_back(o);
        }
        for (int i = 0; i < d->n_lights; i++) {
            const rl_light& l = d->lights[i];
            DevLight o;
            o.pos = make_float4((float)l.position[0], (float)l.position[1], (float)l.position[2], 0.0f);
            o.intensity = make_float4((float)l.intensity[0], (float)l.intensity[1], (float)l.intensity[2], 0.0f);
            fs->lights.push_back(o);
        }
        return true;
    }

    bool check_material(int m) {
        if (m < 0 || m >= d->n_materials) return fail(RL_E_INVALID, "material index out of range");
        return true;
    }

    // world -> pattern space for a leaf whose material has a pattern (pattern.inv * inv_total)
    bool pattern_xform(int material, const Aff& inv_total, Aff* out, bool* has) {
        int tex = d->materials[material].texture;
        *has = false;
        *out = inv_total;
        if (tex < 0) return true;
        Aff fwd, pinv;
        if (!aff_from_4x4(d->textures[tex].transform, &fwd, err)) return fail(RL_E_UNSUPPORTED, *err);
        if (!aff_inverse(fwd, &pinv)) return fail(RL_E_INVALID, "Matrix is not invertible.");
        *out = aff_mul(pinv, inv_total);
        *has = true;
        return true;
    }

    // ---- RTC -----------------------------------------------------------------------------------
    // fwd_total: object -> world, inv_total: world -> object
    // RTC ties at equal t ("later object wins", intersect.rs:159-168; the before() / later() order of n1 / n2 and shadows)
    // follow the reference's World / Group order = the order this DFS meets the leaves.  Each leaf gets its DFS ordinal,
    // the device orders by it, and ord_node maps it back to the caller's node id for rl_hit — so the caller's ids may be
    // in any order, and a subtree referenced twice gets two ordinals.
    int leaf_ordinal(int id) {
        fs->ord_node.push_back(id);
        return (int)fs->ord_node.size() - 1;
    }
    bool rtc_node(int id, const Aff& fwd_total, const Aff& inv_total, int depth) {
        if (!check_node(id)) return false;
        if (depth > 64) return fail(RL_E_INVALID, "object tree too deep (cycle?)");
        const rl_node& nd = d->nodes[id];
        switch (nd.kind) {
            case RL_RTC_SPHERE: case RL_RTC_PLANE: case RL_RTC_CUBE: case RL_RTC_CYLINDER: case RL_RTC_CONE: {
                if (!check_material(nd.material)) return false;
                RtcPrim p{};
                store_rows(inv_total, p.inv);
                store_rows(fwd_total, p.fwd);
                Aff pat;
                bool has;
                if (!pattern_xform(nd.material, aff_identity(), &pat, &has)) return false;
                store_rows(pat, p.pat);
                p.ymin = -INFINITY;
                p.ymax = INFINITY;
                if (nd.kind == RL_RTC_CYLINDER || nd.kind == RL_RTC_CONE) {
                    const double* q = params(nd, 2);
                    if (!q) return false;
                    p.ymin = (float)q[0];
                    p.ymax = (float)q[1];
                }
                p.kind = nd.kind == RL_RTC_SPHERE ? PK_RTC_SPHERE : nd.kind == RL_RTC_PLANE ? PK_RTC_PLANE
                       : nd.kind == RL_RTC_CUBE ? PK_RTC_CUBE : nd.kind == RL_RTC_CYLINDER ? PK_RTC_CYLINDER
                                                                                            : PK_RTC_CONE;
                p.flags = nd.flags & 1;
                p.material = nd.material;
                p.node = leaf_ordinal(id);
                fs->prims.push_back(p);
                return true;
            }
            case RL_RTC_TRIANGLE: {
                if (!check_material(nd.material)) return false;
                const double* q = params(nd, 18);
                if (!q) return false;
                double P[3][3], N[3][3];
                for (int k = 0; k < 3; k++) aff_point(fwd_total, q + 3 * k, P[k]);
                bool smooth = nd.flags & 1;
                if (smooth) {
                    // normalize(inv^T * normalize(sum b_i n_i)) == normalize(sum b_i (inv^T n_i))
                    for (int k = 0; k < 3; k++) aff_normal(inv_total, q + 9 + 3 * k, N[k]);
                } else {
                    // Triangle::flat (triangle.rs:30-42): normal = normalize(e2 x e1) in object space
                    double e1[3], e2[3], n[3];
                    for (int k = 0; k < 3; k++) { e1[k] = q[3 + k] - q[k]; e2[k] = q[6 + k] - q[k]; }
                    n[0] = e2[1] * e1[2] - e2[2] * e1[1];
                    n[1] = e2[2] * e1[0] - e2[0] * e1[2];
                    n[2] = e2[0] * e1[1] - e2[1] * e1[0];
                    if (!normalize3(n)) return fail(RL_E_INVALID, "degenerate triangle cannot be normalized");
                    aff_normal(inv_total, n, N[0]);
                    if (!normalize3(N[0])) return fail(RL_E_INVALID, "degenerate triangle normal");
                    for (int k = 0; k < 3; k++) N[1][k] = N[2][k] = N[0][k];
                }
                Aff pat;
                bool has;
                if (!pattern_xform(nd.material, inv_total, &pat, &has)) return false;
                int xf = 0;
                if (has) {
                    Xform x;
                    store_rows(pat, x.r);
                    fs->xforms.push_back(x);
                    xf = (int)fs->xforms.size();  // index + 1
                }
                if (csg_depth > 0) {
                    // a leaf of a Csg (csg.rs:31-35 takes any Object): an analytic primitive inside the Csg's contiguous leaf
                    // range — world-space vertices, normals, world -> pattern space (rtc_kernels.cu tri_prim_roots)
                    if (has) fs->xforms.pop_back();
                    RtcPrim p{};
                    for (int k = 0; k < 3; k++) {
                        p.inv[k] = make_float4((float)P[k][0], (float)P[k][1], (float)P[k][2], 0.0f);
                        p.fwd[k] = make_float4((float)N[k][0], (float)N[k][1], (float)N[k][2], 0.0f);
                    }
                    store_rows(has ? pat : aff_identity(), p.pat);
                    p.ymin = -INFINITY;
                    p.ymax = INFINITY;
                    p.kind = PK_TRIANGLE;
                    p.flags = smooth ? 1 : 0;
                    p.material = nd.material;
                    p.node = leaf_ordinal(id);
                    fs->prims.push_back(p);
                    return true;
                }
                push_triangle(P, N, nullptr, nd.material, leaf_ordinal(id), (smooth ? 1 : 0) | (xf << 8));
                return true;
            }
            case RL_RTC_MESH: {  // io/wavefront_obj.rs:73-76: Bounded<Group<Triangle>> of the ctx's parsed mesh
                if (csg_depth > 0) return fail(RL_E_UNSUPPORTED, "a device mesh under a Csg is not lowered yet");
                if (!mesh) return fail(RL_E_INVALID, "RL_RTC_MESH needs a ctx that holds a parsed mesh (rl_obj_parse)");
                if (!check_material(nd.material)) return false;
                if (mesh->n_triangles == 0) return true;
                Aff pat;
                bool has;
                if (!pattern_xform(nd.material, inv_total, &pat, &has)) return false;
                const int ord = leaf_ordinal(id);  // one ordinal for the mesh: its triangles tie-break by index
                FlatScene::MeshUse* u = mesh_use(fwd_total, inv_total, nd.material, ord);
                if (has) {
                    Xform x;
                    store_rows(pat, x.r);
                    fs->xforms.push_back(x);
                    u->xf = (int)fs->xforms.size();
                }
                u->bvh_first = (int)fs->bvh_ref.size();
                fs->bvh_aabb.resize(fs->bvh_aabb.size() + 6 * (size_t)mesh->n_triangles, 0.0f);
                fs->bvh_ref.resize(fs->bvh_ref.size() + (size_t)mesh->n_triangles, 0);
                fs->bvh_node_id.resize(fs->bvh_node_id.size() + (size_t)mesh->n_triangles, ord);
                return true;
            }
            case RL_RTC_TRANSFORMED: {
                const double* q = params(nd, 16);
                if (!q) return false;
                Aff fwd, inv;
                if (!aff_from_4x4(q, &fwd, err)) return fail(RL_E_UNSUPPORTED, *err);
                if (!aff_inverse(fwd, &inv)) return fail(RL_E_INVALID, "Matrix is not invertible.");
                return rtc_node(nd.child_begin, aff_mul(fwd_total, fwd), aff_mul(inv, inv_total), depth + 1);
            }
            case RL_RTC_GROUP: {
                if (nd.child_begin < 0 || nd.child_end > d->n_children || nd.child_begin > nd.child_end)
                    return fail(RL_E_INVALID, "group children out of range");
                for (int k = nd.child_begin; k < nd.child_end; k++)
                    if (!rtc_node(d->children[k], fwd_total, inv_total, depth + 1)) return false;
                return true;
            }
            case RL_RTC_BOUNDED:
                return rtc_node(nd.child_begin, fwd_total, inv_total, depth + 1);
            case RL_RTC_CSG: {
                // csg.rs:31-35.  Leaves are emitted in DFS order, so a CSG node's operands are the contiguous prim
                // ranges [lo, mid) and [mid, hi); nodes are stored post-order (children before parents), which is the
                // order the device filter applies them in (rtc_kernels.cu each_prim).
                if (nd.flags < RL_CSG_UNION || nd.flags > RL_CSG_DIFFERENCE) return fail(RL_E_INVALID, "unknown CsgOperation");
                const bool outermost = csg_depth == 0;
                const int first = (int)fs->csg.size();
                const int lo = (int)fs->prims.size();
                csg_depth++;
                bool ok = rtc_node(nd.child_begin, fwd_total, inv_total, depth + 1);
                const int mid = (int)fs->prims.size();
                ok = ok && rtc_node(nd.child_end, fwd_total, inv_total, depth + 1);
                csg_depth--;
                if (!ok) return false;
                const int hi = (int)fs->prims.size();
                fs->csg.push_back(make_int4(nd.flags, lo, mid, hi));
                if (outermost) {
                    const int count = (int)fs->csg.size() - first;
                    for (int k = lo; k < hi; k++) {
                        fs->prims[k].csg_first = first;
                        fs->prims[k].csg_count = count;
                    }
                }
                return true;
            }
            default:
                return fail(RL_E_INVALID, "unknown RTC node kind");
        }
    }

    // a use of the ctx's device-resident mesh: reserve its triangle slots, remember the transform.  Returns the use.
    FlatScene::MeshUse* mesh_use(const Aff& fwd, const Aff& inv, int material, int node) {
        FlatScene::MeshUse u{};
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 4; j++) { u.fwd[i][j] = fwd.m[i][j]; u.inv[i][j] = inv.m[i][j]; }
        u.material = material;
        u.node = node;
        u.tri_first = (int)fs->tri_verts.size();
        u.bvh_first = -1;
        u.xf = 0;
        fs->tri_verts.resize(fs->tri_verts.size() + (size_t)mesh->n_triangles, TriVerts{});
        fs->tri_shade.resize(fs->tri_shade.size() + (size_t)mesh->n_triangles, TriShade{});
        fs->meshes.push_back(u);
        return &fs->meshes.back();
    }
    // world-space box of the mesh: the 8 transformed corners of its object-space box
    void mesh_world_box(const Aff& fwd, double lo[3], double hi[3]) const {
        for (int k = 0; k < 3; k++) { lo[k] = INFINITY; hi[k] = -INFINITY; }
        for (int c = 0; c < 8; c++) {
            double p[3] = {mesh->bounds[(c & 1) ? 3 : 0], mesh->bounds[(c & 2) ? 4 : 1], mesh->bounds[(c & 4) ? 5 : 2]}, w[3];
            aff_point(fwd, p, w);
            for (int k = 0; k < 3; k++) { lo[k] = std::fmin(lo[k], w[k]); hi[k] = std::fmax(hi[k], w[k]); }
        }
    }

    void push_triangle(const double P[3][3], const double N[3][3], const double* uv, int material, int node,
                       int flags) {
        TriVerts v;
        v.p0 = make_float4((float)P[0][0], (float)P[0][1], (float)P[0][2], as_f(material));
        v.p1 = make_float4((float)P[1][0], (float)P[1][1], (float)P[1][2], as_f(node));
        v.p2 = make_float4((float)P[2][0], (float)P[2][1], (float)P[2][2], as_f(flags));
        TriShade s;
        double u[6] = {0, 0, 0, 0, 0, 0};
        if (uv) std::memcpy(u, uv, sizeof(u));
        s.s0 = make_float4((float)N[0][0], (float)N[0][1], (float)N[0][2], (float)u[0]);
        s.s1 = make_float4((float)N[1][0], (float)N[1][1], (float)N[1][2], (float)u[1]);
        s.s2 = make_float4((float)N[2][0], (float)N[2][1], (float)N[2][2], (float)u[2]);
        s.s3 = make_float4((float)u[3], (float)u[4], (float)u[5], 0.0f);
        int idx = (int)fs->tri_verts.size();
        fs->tri_verts.push_back(v);
        fs->tri_shade.push_back(s);
        double lo[3], hi[3];
        for (int k = 0; k < 3; k++) {
            // bound the f32-rounded vertices the device will actually intersect
            float a = (&v.p0.x)[k], b = (&v.p1.x)[k], c = (&v.p2.x)[k];
            lo[k] = std::fmin(a, std::fmin(b, c));
            hi[k] = std::fmax(a, std::fmax(b, c));
        }
        emit(lo, hi, make_ref(REF_TRI, idx), node);
    }

    // ---- OW ------------------------------------------------------------------------------------
    // M: object -> world linear part + offset (p_world = M p + t); s = uniform scale factor of M
    bool ow_node(int id, const Aff& fwd, const Aff& inv, bool rotated, int depth) {
        if (!check_node(id)) return false;
        if (depth > 64) return fail(RL_E_INVALID, "hittable tree too deep (cycle?)");
        const rl_node& nd = d->nodes[id];
        switch (nd.kind) {
            case RL_OW_SPHERE: {
                if (!check_material(nd.material)) return false;
                const double* q = params(nd, 7);
                if (!q) return false;
                double c1[3], c2[3];
                aff_point(fwd, q, c1);
                aff_point(fwd, q + 3, c2);
                // Transform only offers rotations and uniform scale, so a sphere stays a sphere
                double sx = std::sqrt(fwd.m[0][0] * fwd.m[0][0] + fwd.m[1][0] * fwd.m[1][0] + fwd.m[2][0] * fwd.m[2][0]);
                double r = q[6] * sx;
                const rl_material& m = d->materials[nd.material];
                if (rotated && m.texture >= 0 && texture_uses_uv(m.texture))
                    return fail(RL_E_UNSUPPORTED, "image-textured sphere under a rotation is not lowered yet");
                OwSphere s;
                s.c = make_float4((float)c1[0], (float)c1[1], (float)c1[2], (float)r);
                bool moving = nd.flags & 1;
                s.dc = make_float4(moving ? (float)(c2[0] - c1[0]) : 0.0f, moving ? (float)(c2[1] - c1[1]) : 0.0f,
                                   moving ? (float)(c2[2] - c1[2]) : 0.0f, as_f(nd.material));
                int idx = (int)fs->spheres.size();
                fs->spheres.push_back(s);
                fs->sphere_node.push_back(id);
                // swept bounds (sphere.rs:77-87), from the f32 values the device uses
                double lo[3], hi[3];
                double ar = std::fabs((double)s.c.w);
                for (int k = 0; k < 3; k++) {
                    double a = (&s.c.x)[k], b = a + (&s.dc.x)[k];
                    lo[k] = std::fmin(a, b) - ar;
                    hi[k] = std::fmax(a, b) + ar;
                }
                emit(lo, hi, make_ref(REF_SPHERE, idx), id);  // select_big_prims decides later whether it leaves the LBVH
                return true;
            }
            case RL_OW_QUAD: {
                if (!check_material(nd.material)) return false;
                const double* q = params(nd, 9);
                if (!q) return false;
                double Q[3], U[3], V[3];
                aff_point(fwd, q, Q);
                aff_vec(fwd, q + 3, U);
                aff_vec(fwd, q + 6, V);
                double n[3] = {U[1] * V[2] - U[2] * V[1], U[2] * V[0] - U[0] * V[2], U[0] * V[1] - U[1] * V[0]};
                double nn = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
                if (!(nn > 1e-16)) return fail(RL_E_INVALID, "Failed to find normal because u and v were parallel");
                double un[3] = {n[0], n[1], n[2]};
                normalize3(un);
                // plane.rs:51-80 evaluates alpha = w . (h x v), beta = w . (u x h) with h = p - q, w = n / (n . n).  By the
                // cyclic symmetry of the triple product alpha = h . (v x w) and beta = h . (w x u): two dot products with
                // vectors precomputed here in f64 (the device test is 3 loads + ~25 instructions instead of 5 + ~50)
                double W[3] = {n[0] / nn, n[1] / nn, n[2] / nn};
                double A[3] = {V[1] * W[2] - V[2] * W[1], V[2] * W[0] - V[0] * W[2], V[0] * W[1] - V[1] * W[0]};
                double B[3] = {W[1] * U[2] - W[2] * U[1], W[2] * U[0] - W[0] * U[2], W[0] * U[1] - W[1] * U[0]};
                OwQuad o;
                o.n = make_float4((float)un[0], (float)un[1], (float)un[2], (float)(un[0] * Q[0] + un[1] * Q[1] + un[2] * Q[2]));
                o.a = make_float4((float)A[0], (float)A[1], (float)A[2], (float)(A[0] * Q[0] + A[1] * Q[1] + A[2] * Q[2]));
                o.b = make_float4((float)B[0], (float)B[1], (float)B[2], (float)(B[0] * Q[0] + B[1] * Q[1] + B[2] * Q[2]));
                o.m = make_int4(nd.material, 0, 0, 0);
                int idx = (int)fs->quads.size();
                fs->quads.push_back(o);
                fs->quad_node.push_back(id);
                double lo[3], hi[3];
                for (int k = 0; k < 3; k++) {
                    double c[4] = {Q[k], Q[k] + U[k], Q[k] + V[k], Q[k] + U[k] + V[k]};
                    lo[k] = std::fmin(std::fmin(c[0], c[1]), std::fmin(c[2], c[3]));
                    hi[k] = std::fmax(std::fmax(c[0], c[1]), std::fmax(c[2], c[3]));
                }
                emit(lo, hi, make_ref(REF_QUAD, idx), id);
                return true;
            }
            case RL_OW_TRIANGLE: {
                if (!check_material(nd.material)) return false;
                const double* q = params(nd, 24);
                if (!q) return false;
                double P[3][3], N[3][3];
                for (int k = 0; k < 3; k++) aff_point(fwd, q + 3 * k, P[k]);
                bool has_uv = nd.flags & 1, has_n = nd.flags & 2;
                if (has_n) {
                    for (int k = 0; k < 3; k++) aff_normal(inv, q + 15 + 3 * k, N[k]);
                } else {
                    // Plane::new (flat/plane.rs:23-28): n = normalize(u x v)
                    double u[3], v[3], n[3];
                    for (int k = 0; k < 3; k++) { u[k] = P[1][k] - P[0][k]; v[k] = P[2][k] - P[0][k]; }
                    n[0] = u[1] * v[2] - u[2] * v[1];
                    n[1] = u[2] * v[0] - u[0] * v[2];
                    n[2] = u[0] * v[1] - u[1] * v[0];
                    if (!normalize3(n)) return fail(RL_E_INVALID, "Failed to find normal because u and v were parallel");
                    for (int k = 0; k < 3; k++) N[0][k] = N[1][k] = N[2][k] = n[k];
                }
                push_triangle(P, N, has_uv ? q + 9 : nullptr, nd.material, id, (has_n ? 1 : 0) | (has_uv ? 2 : 0));
                return true;
            }
            case RL_OW_MESH: {  // io/wavefront_obj.rs:88-103: Bvh<Triangle<&M>> of the ctx's parsed mesh
                if (in_boundary) return fail(RL_E_UNSUPPORTED, "a device mesh as the boundary of a ConstantMedium is not lowered yet");
                if (!mesh) return fail(RL_E_INVALID, "RL_OW_MESH needs a ctx that holds a parsed mesh (rl_obj_parse)");
                if (!check_material(nd.material)) return false;
                if (mesh->n_triangles == 0) return fail(RL_E_INVALID, "Cannot make a BVH node without hittables.");
                mesh_use(fwd, inv, nd.material, id);
                // ONE placeholder in the LBVH input, with the mesh's world box, so that select_big_prims weighs the walls
                // around a mesh against the mesh; flatten_scene then swaps it for the mesh's triangle slots
                double lo[3], hi[3];
                mesh_world_box(fwd, lo, hi);
                push_aabb(fs, lo, hi, make_ref(REF_TRI, MESH_MARK + (int)fs->meshes.size() - 1), id);
                return true;
            }
            case RL_OW_TRANSFORM: {
                const double* q = params(nd, 18);
                if (!q) return false;
                Aff m{}, mi{};
                for (int i = 0; i < 3; i++)
                    for (int j = 0; j < 3; j++) { m.m[i][j] = q[i * 3 + j]; mi.m[i][j] = q[9 + i * 3 + j]; }
                bool rot = rotated || m.m[0][1] != 0.0 || m.m[0][2] != 0.0 || m.m[1][0] != 0.0 ||
                           m.m[1][2] != 0.0 || m.m[2][0] != 0.0 || m.m[2][1] != 0.0;
                return ow_node(nd.child_begin, aff_mul(fwd, m), aff_mul(mi, inv), rot, depth + 1);
            }
            case RL_OW_TRANSLATE: {
                const double* q = params(nd, 3);
                if (!q) return false;
                Aff m = aff_identity(), mi = aff_identity();
                for (int i = 0; i < 3; i++) { m.m[i][3] = q[i]; mi.m[i][3] = -q[i]; }
                return ow_node(nd.child_begin, aff_mul(fwd, m), aff_mul(mi, inv), rotated, depth + 1);
            }
            case RL_OW_BVH:
            case RL_OW_LIST: {
                if (nd.child_begin < 0 || nd.child_end > d->n_children || nd.child_begin > nd.child_end)
                    return fail(RL_E_INVALID, "children out of range");
                if (nd.kind == RL_OW_BVH && nd.child_begin == nd.child_end)
                    return fail(RL_E_INVALID, "Cannot make a BVH node without hittables.");
                for (int k = nd.child_begin; k < nd.child_end; k++)
                    if (!ow_node(d->children[k], fwd, inv, rotated, depth + 1)) return false;
                return true;
            }
            case RL_OW_CONSTANT_MEDIUM: {  // hittable/constant_medium.rs:14-22
                if (!check_material(nd.material)) return false;
                const double* q = params(nd, 1);
                if (!q) return false;
                if (in_boundary) return fail(RL_E_UNSUPPORTED, "a ConstantMedium inside the boundary of another one");
                if (d->materials[nd.material].kind != RL_MAT_OW_ISOTROPIC)
                    return fail(RL_E_UNSUPPORTED, "ConstantMedium is only intended to be used with Isotropic material");
                OwMedium m;
                m.ref_begin = (int)fs->medium_refs.size();
                m.neg_inv_density = (float)(-1.0 / q[0]);
                m.material = nd.material;
                in_boundary = true;
                for (int k = 0; k < 3; k++) { blo[k] = INFINITY; bhi[k] = -INFINITY; }
                bool ok = ow_node(nd.child_begin, fwd, inv, rotated, depth + 1);
                in_boundary = false;
                if (!ok) return false;
                m.ref_count = (int)fs->medium_refs.size() - m.ref_begin;
                if (m.ref_count == 0) return fail(RL_E_INVALID, "ConstantMedium with an empty boundary");
                int idx = (int)fs->media.size();
                fs->media.push_back(m);
                push_aabb(fs, blo, bhi, make_ref(REF_MEDIUM, idx), id);  // bounding_box = the boundary's (85-87)
                return true;
            }
            default:
                return fail(RL_E_INVALID, "unknown OW node kind");
        }
    }

    // a checker can nest checkers; the device follows at most 8 levels (tex_value), and a cyclic description must not
    // recurse forever here
    bool texture_uses_uv(int tex, int depth = 0) {
        if (depth > 8) return false;
        const rl_texture& t = d->textures[tex];
        if (t.kind == RL_TEX_OW_IMAGE) return true;
        if (t.kind == RL_TEX_OW_CHECKER) return texture_uses_uv(t.tex_a, depth + 1) || texture_uses_uv(t.tex_b, depth + 1);
        return false;
    }
};

}  // namespace

// inverse of an affine 4x4 (row-major) as 3x4 rows — used for the RTC camera (Camera::transform.inverse())
bool invert_affine_4x4(const double* m16, double* inv12, std::string* err) {
    Aff fwd, inv;
    if (!aff_from_4x4(m16, &fwd, err)) return false;
    if (!aff_inverse(fwd, &inv)) {
        *err = "Matrix is not invertible.";
        return false;
    }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 4; j++) inv12[i * 4 + j] = inv.m[i][j];
    return true;
}

// Primitives that are large against the rest of the scene (the walls and the light of a Cornell box around a mesh)
// leave the LBVH for the brute-force "big" list (scene.h OW_MAX_BIG): inside the hierarchy their boxes overlap
// everything, so every ray walks them anyway, while outside it a ray that misses the remaining geometry's root box
// ends its traversal after one node.  Rule: of the OW_MAX_BIG largest boxes (by surface area), those whose area is at
// least a quarter of the area of the box around all the OTHER primitives.
static double box_area(const float* b) {
    double x = (double)b[3] - b[0], y = (double)b[4] - b[1], z = (double)b[5] - b[2];
    return 2.0 * (x * y + y * z + x * z);
}
static void select_big_prims(FlatScene* fs, int mesh_triangles) {
    const int n = (int)fs->bvh_ref.size();
    const int room = OW_MAX_BIG - (int)fs->big_refs.size();
    // a device-resident mesh is ONE placeholder here but stands for all of its triangles
    const long long n_effective = (long long)n + (long long)fs->meshes.size() * (mesh_triangles > 0 ? mesh_triangles - 1 : 0);
    if (n_effective < 2 * OW_MAX_BIG || room <= 0) return;
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    auto bigger = [&](int a, int b) {
        double sa = box_area(&fs->bvh_aabb[6 * a]), sb = box_area(&fs->bvh_aabb[6 * b]);
        return sa > sb || (sa == sb && a < b);
    };
    auto is_mesh = [&](int i) { return ref_type(fs->bvh_ref[i]) == REF_TRI && ref_index(fs->bvh_ref[i]) >= MESH_MARK; };
    order.erase(std::remove_if(order.begin(), order.end(), is_mesh), order.end());  // a mesh placeholder is never "big"
    const int take = std::min(room, (int)order.size());
    if (take == 0) return;
    std::partial_sort(order.begin(), order.begin() + take, order.end(), bigger);
    std::vector<char> cand(n, 0);
    for (int k = 0; k < take; k++) cand[order[k]] = 1;
    float rest[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int i = 0; i < n; i++) {
        if (cand[i]) continue;
        for (int k = 0; k < 3; k++) {
            rest[k] = std::fmin(rest[k], fs->bvh_aabb[6 * i + k]);
            rest[3 + k] = std::fmax(rest[3 + k], fs->bvh_aabb[6 * i + 3 + k]);
        }
    }
    const double limit = 0.25 * box_area(rest);
    std::vector<char> big(n, 0);
    bool any = false;
    for (int k = 0; k < take; k++)
        if (box_area(&fs->bvh_aabb[6 * order[k]]) >= limit) big[order[k]] = 1, any = true;
    if (!any) return;
    std::vector<float> aabb;
    std::vector<int> ref, node;
    for (int i = 0; i < n; i++) {
        if (big[i]) {
            fs->big_refs.push_back(fs->bvh_ref[i]);
            continue;
        }
        aabb.insert(aabb.end(), fs->bvh_aabb.begin() + 6 * i, fs->bvh_aabb.begin() + 6 * i + 6);
        ref.push_back(fs->bvh_ref[i]);
        node.push_back(fs->bvh_node_id[i]);
    }
    fs->bvh_aabb.swap(aabb);
    fs->bvh_ref.swap(ref);
    fs->bvh_node_id.swap(node);
}

// OW: swap every mesh placeholder of the LBVH input for the mesh's triangle slots (filled on the device)
static void expand_mesh_placeholders(FlatScene* fs, const MeshMeta* mesh) {
    if (fs->meshes.empty()) return;
    std::vector<float> aabb;
    std::vector<int> ref, node;
    for (size_t i = 0; i < fs->bvh_ref.size(); i++) {
        const int r = fs->bvh_ref[i];
        if (ref_type(r) == REF_TRI && ref_index(r) >= MESH_MARK) continue;
        aabb.insert(aabb.end(), fs->bvh_aabb.begin() + 6 * i, fs->bvh_aabb.begin() + 6 * i + 6);
        ref.push_back(r);
        node.push_back(fs->bvh_node_id[i]);
    }
    for (FlatScene::MeshUse& u : fs->meshes) {
        u.bvh_first = (int)ref.size();
        aabb.resize(aabb.size() + 6 * (size_t)mesh->n_triangles, 0.0f);
        ref.resize(ref.size() + (size_t)mesh->n_triangles, 0);
        node.resize(node.size() + (size_t)mesh->n_triangles, u.node);
    }
    fs->bvh_aabb.swap(aabb);
    fs->bvh_ref.swap(ref);
    fs->bvh_node_id.swap(node);
}

int flatten_scene(const rl_scene_desc* d, FlatScene* out, std::string* err, const MeshMeta* mesh) {
    if (!d || d->abi_version != RL_B200_ABI_VERSION) {
        *err = "scene description missing or ABI version mismatch";
        return RL_E_INVALID;
    }
    if (d->flavor != RL_FLAVOR_RTC && d->flavor != RL_FLAVOR_OW) {
        *err = "unknown scene flavor";
        return RL_E_INVALID;
    }
    Flattener f{d, out, err};
    f.mesh = mesh;
    out->flavor = d->flavor;
    out->max_reflection_depth = d->max_reflection_depth;
    for (int k = 0; k < 3; k++) out->void_color[k] = (float)d->void_color[k];
    if (!f.lower_tables()) return f.rc;
    if (d->flavor == RL_FLAVOR_RTC) {
        if (d->max_reflection_depth < 0) {  // World.max_reflection_depth is a usize (world.rs:29): any non-negative value
            *err = "max_reflection_depth must be >= 0";
            return RL_E_INVALID;
        }
        for (int k = 0; k < d->n_roots; k++)
            if (!f.rtc_node(d->roots[k], aff_identity(), aff_identity(), 0)) return f.rc;
        for (int& v : out->bvh_node_id) v = out->ord_node[v];  // rl_scene_info reports the caller's node ids
    } else {
        if (d->n_roots != 1) {
            *err = "an OW scene has exactly one root hittable";
            return RL_E_INVALID;
        }
        if (!f.ow_node(d->roots[0], aff_identity(), aff_identity(), false, 0)) return f.rc;
        select_big_prims(out, mesh ? mesh->n_triangles : 0);
        expand_mesh_placeholders(out, mesh);
    }
    return RL_OK;
}

}  // namespace rl
