//! Drop into `ray-tracing-one-weekend/src/gpu.rs` (+ `pub mod gpu;` in `lib.rs`, `rl-b200` in Cargo.toml).
//!
//! `Hittable`, `Material` and `Texture` each gain ONE method, `fn lower(&self, out: &mut SceneBuilder) -> i32`
//! (node id / material id / texture id).  Private fields read below (make them `pub(crate)` or paste each `impl`
//! beside its type):
//!     flat::plane::Plane { q, u, v, material }     hittable/flat/plane.rs:11-19
//!     Quad { plane }, Triangle { plane, normals, texture_coords }   flat/quad.rs:13-16, flat/triangle.rs:13-18
//!     Transform { object, transformation, inv_transformation }      hittable/transform.rs:13-19
//!     Bvh { children }, enum Children              bvh.rs:10-19
//!     Checker { inv_scale }                        texture.rs:25-29
//!     Camera { params }                            camera.rs:60-68
use rl_b200::{ow::LowerOw, SceneBuilder};
use rl_b200_sys as sys;

use crate::{
    bvh::{Bvh, Children},
    camera::{Camera, Canvas},
    color::Color,
    hittable::{flat::{quad::Quad, triangle::Triangle}, sphere::{Center, Sphere}, transform::Transform, translate::Translate, Hittable},
    material::{Dielectric, DiffuseLight, Lambertian, Material, Metal},
    texture::{Checker, Image, SolidColor, Texture},
    vec3::Vec3,
};

fn v3(v: &Vec3) -> [f64; 3] { [v.x(), v.y(), v.z()] }
fn tex(kind: i32) -> sys::rl_texture {
    sys::rl_texture { kind, tex_a: -1, tex_b: -1, image: -1, a: [0.0; 3], b: [0.0; 3], scale: 1.0, transform: [0.0; 16] }
}
fn mat(kind: i32) -> sys::rl_material {
    sys::rl_material { kind, texture: -1, ..Default::default() }
}

// ---- textures (texture.rs) ------------------------------------------------------------------------------------------
impl SolidColor { pub fn lower(&self, out: &mut SceneBuilder) -> i32 { out.texture_id(self, |_| sys::rl_texture { a: v3(&self.albedo), ..tex(sys::RL_TEX_OW_SOLID) }) } }
impl<A: Texture, B: Texture> Checker<A, B> {
    pub fn lower(&self, out: &mut SceneBuilder) -> i32 {
        out.texture_id(self, |out| {
            let (a, b) = (self.even.lower(out), self.odd.lower(out));
            sys::rl_texture { tex_a: a, tex_b: b, scale: 1.0 / self.inv_scale, ..tex(sys::RL_TEX_OW_CHECKER) }
        })
    }
}
impl Image {
    pub fn lower(&self, out: &mut SceneBuilder) -> i32 {
        out.texture_id(self, |out| {
            assert!(self.image.width() > 0 && self.image.height() > 0, "Image has no data"); // texture.rs:64-67
            let id = out.add_image(self.image.width() as i32, self.image.height() as i32, self.image.as_raw().clone());
            sys::rl_texture { image: id, ..tex(sys::RL_TEX_OW_IMAGE) }
        })
    }
}

impl crate::texture::Noise {
    /// Perlin { randvec, perm_x, perm_y, perm_z } (perlin.rs:9-14) are private: read them through a `pub(crate)` view
    pub fn lower(&self, out: &mut SceneBuilder) -> i32 {
        out.texture_id(self, |out| {
            let mut p = sys::rl_perlin { randvec: [[0.0; 3]; 256], perm_x: [0; 256], perm_y: [0; 256], perm_z: [0; 256] };
            for i in 0..256 {
                p.randvec[i] = v3(&self.noise.randvec[i]);
                p.perm_x[i] = self.noise.perm_x[i] as i32;
                p.perm_y[i] = self.noise.perm_y[i] as i32;
                p.perm_z[i] = self.noise.perm_z[i] as i32;
            }
            let id = out.add_perlin(p);
            sys::rl_texture { image: id, scale: self.scale, ..tex(sys::RL_TEX_OW_NOISE) }
        })
    }
}

// ---- materials (material.rs) ----------------------------------------------------------------------------------------
impl<T: Texture> Lambertian<T>   { pub fn lower(&self, out: &mut SceneBuilder) -> i32 { out.material_id(self, |out| sys::rl_material { texture: self.texture.lower(out), ..mat(sys::RL_MAT_OW_LAMBERTIAN) }) } }
impl<T: Texture> DiffuseLight<T> { pub fn lower(&self, out: &mut SceneBuilder) -> i32 { out.material_id(self, |out| sys::rl_material { texture: self.texture.lower(out), ..mat(sys::RL_MAT_OW_DIFFUSE_LIGHT) }) } }
impl<T: Texture> crate::material::Isotropic<T> { pub fn lower(&self, out: &mut SceneBuilder) -> i32 { out.material_id(self, |out| sys::rl_material { texture: self.texture.lower(out), ..mat(sys::RL_MAT_OW_ISOTROPIC) }) } }
impl Metal      { pub fn lower(&self, out: &mut SceneBuilder) -> i32 { out.material_id(self, |_| sys::rl_material { color: v3(&self.albedo), fuzz: self.fuzz, ..mat(sys::RL_MAT_OW_METAL) }) } }
impl Dielectric { pub fn lower(&self, out: &mut SceneBuilder) -> i32 { out.material_id(self, |_| sys::rl_material { refractive_index: self.refraction_index, ..mat(sys::RL_MAT_OW_DIELECTRIC) }) } }

// ---- hittables ------------------------------------------------------------------------------------------------------
impl<M: Material> LowerOw for Sphere<M> {
    fn lower(&self, out: &mut SceneBuilder) -> i32 {
        let m = self.material.lower(out);
        let (c1, c2, moving) = match &self.center {
            Center::Stationary(p) => (v3(p), v3(p), 0),
            Center::Moving(p1, p2) => (v3(p1), v3(p2), 1),
        };
        let p = out.add_params(&[c1[0], c1[1], c1[2], c2[0], c2[1], c2[2], self.radius]);
        out.add_node(sys::RL_OW_SPHERE, m, moving, p)
    }
}
impl<M: Material> LowerOw for Quad<M> {
    fn lower(&self, out: &mut SceneBuilder) -> i32 {
        let m = self.plane.material.lower(out);
        let (q, u, v) = (v3(&self.plane.q), v3(&self.plane.u), v3(&self.plane.v));
        let p = out.add_params(&[q[0], q[1], q[2], u[0], u[1], u[2], v[0], v[1], v[2]]);
        out.add_node(sys::RL_OW_QUAD, m, 0, p)
    }
}
impl<M: Material> LowerOw for Triangle<M> {
    fn lower(&self, out: &mut SceneBuilder) -> i32 {
        let m = self.plane.material.lower(out);
        let (q, u, v) = (&self.plane.q, &self.plane.u, &self.plane.v);
        let mut vals = vec![q.x(), q.y(), q.z(), q.x() + u.x(), q.y() + u.y(), q.z() + u.z(), q.x() + v.x(), q.y() + v.y(), q.z() + v.z()];
        let mut flags = 0;
        match &self.texture_coords { Some(t) => { flags |= 1; for (a, b) in t { vals.extend_from_slice(&[*a, *b]); } } None => vals.extend_from_slice(&[0.0; 6]) }
        match &self.normals { Some(n) => { flags |= 2; for x in n { vals.extend_from_slice(&v3(x)); } } None => vals.extend_from_slice(&[0.0; 9]) }
        let p = out.add_params(&vals);
        out.add_node(sys::RL_OW_TRIANGLE, m, flags, p)
    }
}
impl<H: Hittable + LowerOw> LowerOw for Transform<H> {
    fn lower(&self, out: &mut SceneBuilder) -> i32 {
        let mut vals = Vec::with_capacity(18);
        for r in &self.transformation.0 { vals.extend_from_slice(r); }
        for r in &self.inv_transformation.0 { vals.extend_from_slice(r); }
        let p = out.add_params(&vals);
        let me = out.add_node(sys::RL_OW_TRANSFORM, -1, 0, p);
        let c = self.object.lower(out);
        out.set_node_children(me, c, c + 1);
        me
    }
}
impl<H: Hittable + LowerOw> LowerOw for Translate<H> {
    fn lower(&self, out: &mut SceneBuilder) -> i32 {
        let p = out.add_params(&v3(&self.offset));
        let me = out.add_node(sys::RL_OW_TRANSLATE, -1, 0, p);
        let c = self.object.lower(out);
        out.set_node_children(me, c, c + 1);
        me
    }
}
impl<M: Material, H: Hittable + LowerOw> LowerOw for crate::hittable::constant_medium::ConstantMedium<M, H> {
    /// fields boundary / neg_inv_density / phase_function are private (constant_medium.rs:8-12)
    fn lower(&self, out: &mut SceneBuilder) -> i32 {
        let m = self.phase_function.lower(out);
        let p = out.add_params(&[-1.0 / self.neg_inv_density]); // density
        let me = out.add_node(sys::RL_OW_CONSTANT_MEDIUM, m, 0, p);
        let c = self.boundary.lower(out);
        out.set_node_children(me, c, c + 1);
        me
    }
}
impl<H: Hittable + LowerOw> LowerOw for Bvh<H> {
    /// The median-split tree (bvh.rs:22-61) is replaced by the device LBVH; only the leaves matter (the closest hit
    /// is BVH-independent).  Leaves are emitted in the tree's left-to-right order, which is the order the
    /// reference's fold visits them in (bvh.rs:81-90), so exact-t ties resolve the same way.
    fn lower(&self, out: &mut SceneBuilder) -> i32 {
        fn leaves<H: Hittable + LowerOw>(b: &Bvh<H>, out: &mut SceneBuilder, ids: &mut Vec<i32>) {
            match &b.children {
                Children::Leaf(hs) => for h in hs { ids.push(h.lower(out)); },
                Children::Inner(bs) => for c in bs { leaves(c, out, ids); },
            }
        }
        let me = out.add_node(sys::RL_OW_BVH, -1, 0, -1);
        let mut ids = vec![];
        leaves(self, out, &mut ids);
        let (b, e) = out.add_children(&ids);
        out.set_node_children(me, b, e);
        me
    }
}
impl<H: Hittable + LowerOw> LowerOw for [H] {
    fn lower(&self, out: &mut SceneBuilder) -> i32 {
        let me = out.add_node(sys::RL_OW_LIST, -1, 0, -1);
        let ids: Vec<i32> = self.iter().map(|h| h.lower(out)).collect();
        let (b, e) = out.add_children(&ids);
        out.set_node_children(me, b, e);
        me
    }
}

// ---- the drop-in: Camera::render_gpu / render_from_checkpoint_gpu (camera.rs:122-143) ----------------------------------
impl Camera {
    fn abi(&self) -> sys::rl_ow_camera {
        let p = &self.params;
        sys::rl_ow_camera {
            aspect_ratio: p.aspect_ratio, image_width: p.image_width as i32, samples_per_pixel: p.samples_per_pixel as i32,
            max_depth: p.max_depth as i32, _pad: 0, vfov: p.vfov, lookfrom: v3(&p.lookfrom), lookat: v3(&p.lookat), vup: v3(&p.vup),
            defocus_angle: p.defocus_angle, focus_dist: p.focus_dist, background: v3(&p.background), seed: p.seed,
        }
    }
    fn _render_gpu<H: Hittable + LowerOw + ?Sized>(&self, ctx: &mut rl_b200::Ctx, first_sample: usize, world: &H) -> Canvas {
        let scene = rl_b200::ow::lower_world(world);
        let sums = rl_b200::ow::render_sums(ctx, &scene, &self.abi(), first_sample).expect("rl_render_ow");
        let data = sums.chunks_exact(3).map(|c| Color::new(c[0] as f64, c[1] as f64, c[2] as f64)).collect();
        Canvas::from_sums(self.params.samples_per_pixel, self.params.image_width, self.image_height, data) // 3-line ctor beside `merge`
    }
    pub fn render_gpu<H: Hittable + LowerOw + ?Sized>(&self, ctx: &mut rl_b200::Ctx, world: &H) -> Canvas {
        self._render_gpu(ctx, 0, world)
    }
    pub fn render_from_checkpoint_gpu<H: Hittable + LowerOw + ?Sized>(&self, ctx: &mut rl_b200::Ctx, world: &H, checkpoint: &Canvas) -> Canvas {
        self._render_gpu(ctx, checkpoint.samples, world).merge(checkpoint)
    }
}
