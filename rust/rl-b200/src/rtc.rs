//! ray-tracer-challenge side: the lowering trait and the drop-in render call.
//!
//! `Object` (RTC/src/scene/object/mod.rs:10-14) exposes only `material / intersect / bounds`, and wrapper fields
//! are private, so a flattener needs ONE additive method on the trait.  The reference crate adds
//!
//! ```ignore
//! pub trait Object: Sync + Send {
//!     fn material(&self) -> &Material;
//!     fn intersect(&self, ray: &Ray) -> Vec<Intersection<&dyn Object, Color, NormalizedVec3d>>;
//!     fn bounds(&self) -> Bounds;
//!     fn lower(&self, out: &mut rl_b200::SceneBuilder) -> i32;      // <- new, returns the node id
//! }
//! ```
//!
//! and one `lower` body per concrete type — all of them are in `rust/reference-impls/rtc_lower.rs`.
use crate::{sys_reexport as sys, Ctx, Result, SceneBuilder};

/// What `Object::lower`, `Pattern::lower` and `Material` lowering write through.
pub trait LowerRtc {
    /// Record `self` (and its subtree) and return its node id.
    fn lower(&self, out: &mut SceneBuilder) -> i32;
}

/// `World` → scene description.  `objects`: `World.objects` in order (hit ties are resolved by this order,
/// RTC/src/scene/intersect.rs:159-168).
pub fn lower_world<'a>(objects: impl Iterator<Item = &'a dyn LowerRtc>, lights: &[([f64; 3], [f64; 3])],
                       max_reflection_depth: usize, void_color: [f64; 3]) -> SceneBuilder {
    let mut sb = SceneBuilder::new(sys::RL_FLAVOR_RTC);
    for o in objects {
        let id = o.lower(&mut sb);
        sb.roots.push(id);
    }
    for (p, i) in lights {
        sb.lights.push(sys::rl_light { position: *p, intensity: *i });
    }
    sb.max_reflection_depth = max_reflection_depth as i32;
    sb.void_color = void_color;
    sb
}

/// Drop-in body of `Camera::render(&self, &World, &RenderOpts) -> Canvas` (RTC/src/scene/camera.rs:93-124):
/// returns W*H*3 f32 means, row-major (`width * y + x`), which the caller writes into its `Canvas`.
/// Panics like the reference does on invalid input (`unwrap`) — errors of the device path are returned.
pub fn render(ctx: &mut Ctx, scene: &SceneBuilder, hsize: usize, vsize: usize, fov: f64, transform: [f64; 16],
              anti_aliasing_samples: usize) -> Result<Vec<f32>> {
    ctx.scene_upload(scene)?;
    let cam = sys::rl_rtc_camera { hsize: hsize as i32, vsize: vsize as i32, fov, transform };
    let (rgb, _stats) = ctx.render_rtc(&cam, anti_aliasing_samples as u32)?;
    Ok(rgb)
}
