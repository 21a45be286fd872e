//! ray-tracing-one-weekend side: lowering traits and the drop-in render calls.
//!
//! `Hittable` (OW/src/hittable/mod.rs:40-43), `Material` (OW/src/material.rs:11-20) and `Texture`
//! (OW/src/texture.rs:5-7) are behaviour-only traits; each gets ONE additive method:
//!
//! ```ignore
//! pub trait Hittable { ...; fn lower(&self, out: &mut rl_b200::SceneBuilder) -> i32; }   // node id
//! pub trait Material { ...; fn lower(&self, out: &mut rl_b200::SceneBuilder) -> i32; }   // material id
//! pub trait Texture  { ...; fn lower(&self, out: &mut rl_b200::SceneBuilder) -> i32; }   // texture id
//! ```
//!
//! The per-type bodies are in `rust/reference-impls/ow_lower.rs`.
use crate::{sys_reexport as sys, Ctx, Result, SceneBuilder};

pub trait LowerOw {
    fn lower(&self, out: &mut SceneBuilder) -> i32;
}

pub fn lower_world(world: &dyn LowerOw) -> SceneBuilder {
    let mut sb = SceneBuilder::new(sys::RL_FLAVOR_OW);
    let id = world.lower(&mut sb);
    sb.roots.push(id);
    sb
}

/// Drop-in body of `Camera::_render(first_sample, world)` (OW/src/camera.rs:145-199): per-pixel colour SUMS
/// over samples `[first_sample, first_sample + samples_per_pixel)`, row-major, so that `Canvas::merge` and
/// `render_from_checkpoint` (camera.rs:136-143, 273-291) keep working unchanged.
pub fn render_sums(ctx: &mut Ctx, scene: &SceneBuilder, cam: &sys::rl_ow_camera, first_sample: usize) -> Result<Vec<f32>> {
    ctx.scene_upload(scene)?;
    let (sums, _stats) = ctx.render_ow(cam, first_sample as u32)?;
    Ok(sums)
}
