//! Drop into `ray-tracer-challenge/src/scene/gpu.rs` (+ `pub mod gpu;` in `scene/mod.rs`, `rl-b200` in Cargo.toml).
//!
//! One `lower` body per concrete type behind the new trait method `Object::lower` (and `Pattern::lower`).
//! The bodies touch private fields, so either paste each `impl` block next to its type or relax the handful of
//! fields they read to `pub(crate)`:
//!     Transformed { child, transform }            scene/object/transformed.rs:12-16
//!     Group { children }                           scene/object/group.rs:14-16
//!     Bounded { child }                            scene/object/bounded.rs:87-90
//!     Triangle { points, normal, material }        scene/object/triangle.rs:22-27 (+ enum TriangleNormal 17-20)
//! Everything else they read is already `pub`.  Nothing here computes: composing / inverting transforms, bounds and
//! the LBVH are the library's job (csrc/flatten.cpp), so these bodies stay one-to-one with the reference's fields.
use rl_b200::{rtc::LowerRtc, SceneBuilder};
use rl_b200_sys as sys;

use crate::{
    draw::color::Color,
    math::matrix::InvertibleMatrix,
    scene::{
        material::{Material, Surface},
        object::{
            bounded::Bounded, cone::Cone, csg::{Csg, CsgOperation}, cube::Cube, cylinder::Cylinder, group::Group,
            plane::Plane, sphere::Sphere, transformed::Transformed, triangle::{Triangle, TriangleNormal}, Object,
        },
        pattern::{checker3d::Checker3d, gradient::Gradient, ring::Ring, stripe::Stripe},
    },
};

fn rgb(c: &Color) -> [f64; 3] {
    [c.r(), c.g(), c.b()]
}

fn flat4(m: &InvertibleMatrix<4>) -> [f64; 16] {
    let mut out = [0.0; 16];
    for n in 0..4 {
        for k in 0..4 {
            out[n * 4 + k] = m.at(n, k); // Deref to the forward matrix (math/matrix.rs:272-278)
        }
    }
    out
}

// ---- patterns (Pattern gains `fn lower(&self, out: &mut SceneBuilder) -> i32`) -------------------------------------
fn lower_pattern<T>(key: &T, kind: i32, a: &Color, b: &Color, transform: &InvertibleMatrix<4>, out: &mut SceneBuilder) -> i32 {
    out.texture_id(key, |_| sys::rl_texture {
        kind, tex_a: -1, tex_b: -1, image: -1, a: rgb(a), b: rgb(b), scale: 1.0, transform: flat4(transform),
    })
}
impl Stripe    { pub fn lower(&self, out: &mut SceneBuilder) -> i32 { lower_pattern(self, sys::RL_TEX_RTC_STRIPE, &self.a, &self.b, &self.transform, out) } }
impl Checker3d { pub fn lower(&self, out: &mut SceneBuilder) -> i32 { lower_pattern(self, sys::RL_TEX_RTC_CHECKER3D, &self.a, &self.b, &self.transform, out) } }
impl Gradient  { pub fn lower(&self, out: &mut SceneBuilder) -> i32 { lower_pattern(self, sys::RL_TEX_RTC_GRADIENT, &self.a, &self.b, &self.transform, out) } }
impl Ring      { pub fn lower(&self, out: &mut SceneBuilder) -> i32 { lower_pattern(self, sys::RL_TEX_RTC_RING, &self.a, &self.b, &self.transform, out) } }

// ---- material (scene/material.rs:22-31) ------------------------------------------------------------------------------
fn lower_material(m: &Material, out: &mut SceneBuilder) -> i32 {
    out.material_id(m, |out| {
        let (texture, color) = match &m.surface {
            Surface::Color(c) => (-1, rgb(c)),
            Surface::Pattern(p) => (p.lower(out), [0.0; 3]), // `Pattern::lower`, the one new trait method
        };
        sys::rl_material {
            kind: sys::RL_MAT_RTC_PHONG, texture, color,
            ambient: m.ambient, diffuse: m.diffuse, specular: m.specular, shininess: m.shininess,
            reflectivity: m.reflectivity, transparency: m.transparency, refractive_index: m.refractive_index, fuzz: 0.0,
        }
    })
}

// ---- leaves -----------------------------------------------------------------------------------------------------------
impl LowerRtc for Sphere { fn lower(&self, out: &mut SceneBuilder) -> i32 { let m = lower_material(&self.material, out); out.add_node(sys::RL_RTC_SPHERE, m, 0, -1) } }
impl LowerRtc for Plane  { fn lower(&self, out: &mut SceneBuilder) -> i32 { let m = lower_material(&self.material, out); out.add_node(sys::RL_RTC_PLANE, m, 0, -1) } }
impl LowerRtc for Cube   { fn lower(&self, out: &mut SceneBuilder) -> i32 { let m = lower_material(&self.material, out); out.add_node(sys::RL_RTC_CUBE, m, 0, -1) } }

fn lower_truncated(kind: i32, material: &Material, minimum: Option<f64>, maximum: Option<f64>, closed: bool, out: &mut SceneBuilder) -> i32 {
    let m = lower_material(material, out);
    let p = out.add_params(&[minimum.unwrap_or(f64::NEG_INFINITY), maximum.unwrap_or(f64::INFINITY)]);
    out.add_node(kind, m, closed as i32, p)
}
impl LowerRtc for Cylinder { fn lower(&self, out: &mut SceneBuilder) -> i32 { lower_truncated(sys::RL_RTC_CYLINDER, &self.material, self.minimum, self.maximum, self.closed, out) } }
impl LowerRtc for Cone     { fn lower(&self, out: &mut SceneBuilder) -> i32 { lower_truncated(sys::RL_RTC_CONE, &self.material, self.minimum, self.maximum, self.closed, out) } }

impl LowerRtc for Triangle {
    fn lower(&self, out: &mut SceneBuilder) -> i32 {
        let m = lower_material(&self.material, out);
        let mut v = Vec::with_capacity(18);
        for p in &self.points { v.extend_from_slice(&[p.x(), p.y(), p.z()]); }
        let smooth = match &self.normal {
            TriangleNormal::Smooth(ns) => { for n in ns { v.extend_from_slice(&[n.x(), n.y(), n.z()]); } 1 }
            TriangleNormal::Flat(_) => { v.extend_from_slice(&[0.0; 9]); 0 } // the library rebuilds normalize(e2 x e1)
        };
        let p = out.add_params(&v);
        out.add_node(sys::RL_RTC_TRIANGLE, m, smooth, p)
    }
}

// ---- wrappers -----------------------------------------------------------------------------------------------------------
impl<T: Object + LowerRtc> LowerRtc for Transformed<T> {
    fn lower(&self, out: &mut SceneBuilder) -> i32 {
        let p = out.add_params(&flat4(&self.transform));
        let me = out.add_node(sys::RL_RTC_TRANSFORMED, -1, 0, p);
        let c = self.child.lower(out);
        out.set_node_children(me, c, c + 1);
        me
    }
}
impl<T: Object + LowerRtc> LowerRtc for Group<T> {
    fn lower(&self, out: &mut SceneBuilder) -> i32 {
        let me = out.add_node(sys::RL_RTC_GROUP, -1, 0, -1);
        let ids: Vec<i32> = self.children.iter().map(|c| c.lower(out)).collect();
        let (b, e) = out.add_children(&ids);
        out.set_node_children(me, b, e);
        me
    }
}
impl<T: Object + LowerRtc> LowerRtc for Bounded<T> {
    fn lower(&self, out: &mut SceneBuilder) -> i32 {
        let me = out.add_node(sys::RL_RTC_BOUNDED, -1, 0, -1);
        let c = self.child.lower(out);
        out.set_node_children(me, c, c + 1);
        me
    }
}
impl<T: Object + LowerRtc> LowerRtc for Csg<T> {
    fn lower(&self, out: &mut SceneBuilder) -> i32 {
        let op = match self.operation {
            CsgOperation::Union => sys::RL_CSG_UNION,
            CsgOperation::Intersection => sys::RL_CSG_INTERSECTION,
            CsgOperation::Difference => sys::RL_CSG_DIFFERENCE,
        };
        let me = out.add_node(sys::RL_RTC_CSG, -1, op, -1);
        let l = self.left.lower(out);
        let r = self.right.lower(out);
        out.set_node_children(me, l, r);
        me
    }
}
impl LowerRtc for Box<dyn Object> { fn lower(&self, out: &mut SceneBuilder) -> i32 { (**self).lower(out) } } // via `Object::lower`

// ---- the drop-in: Camera::render_gpu, same signature and result as Camera::render (scene/camera.rs:93-124) -----------
impl crate::scene::camera::Camera {
    pub fn render_gpu(&self, ctx: &mut rl_b200::Ctx, world: &crate::scene::world::World,
                      opts: &crate::scene::camera::RenderOpts) -> crate::draw::canvas::Canvas {
        let lights: Vec<_> = world.lights.iter()
            .map(|l| ([l.position.x(), l.position.y(), l.position.z()], rgb(&l.intensity))).collect();
        let scene = rl_b200::rtc::lower_world(world.objects.iter().map(|o| o as &dyn LowerRtc), &lights,
                                              world.max_reflection_depth, rgb(&world.void_color));
        let px = rl_b200::rtc::render(ctx, &scene, self.hsize, self.vsize, self.fov, flat4(&self.transform),
                                      opts.anti_aliasing_samples).expect("rl_render_rtc"); // reference panics on bad input too
        let mut canvas = crate::draw::canvas::Canvas::new(self.hsize, self.vsize);
        for y in 0..self.vsize {
            for x in 0..self.hsize {
                let i = (self.hsize * y + x) * 3; // draw/canvas.rs:42-48 indexing
                canvas.write((x, y), Color::new(px[i] as f64, px[i + 1] as f64, px[i + 2] as f64));
            }
        }
        canvas
    }
}
