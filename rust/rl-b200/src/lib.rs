//! Safe wrapper over `rl-b200-sys`.
//!
//! * [`Ctx`] — RAII over `rl_ctx*`; one per GPU; `Send` but not `Sync` (the C ABI wants one caller thread at a
//!   time per ctx).  `Ctx::new` fails with `Error::NoDevice` when no sm_100 GPU is visible: there is **no CPU
//!   fallback** in this crate — callers that want one keep calling the reference's own `Camera::render`.
//! * [`SceneBuilder`] — records the reference's object tree in the POD form of `include/rl_b200.h`.  The
//!   reference types write themselves into it through `LowerRtc::lower` / `LowerOw::lower` (see
//!   `rtc.rs`, `ow.rs` and INTEGRATION.md §2 — one additive method per reference trait).  The builder only
//!   *records*; composing / pre-inverting transforms and building the LBVH happen inside the library.
//!
//! The executable twin of this file is `rendering_learning_b200/{desc,context}.py`.
pub mod ow;
pub mod rtc;

use rl_b200_sys as sys;
pub use rl_b200_sys as sys_reexport;
use std::{collections::HashMap, ffi::CStr, ptr};

#[derive(Debug)]
pub enum Error {
    Invalid(String),
    NoDevice(String),
    Cuda(String),
    Unsupported(String),
    NoScene,
    Overflow(String),
}

pub type Result<T> = std::result::Result<T, Error>;

pub struct Ctx {
    raw: *mut sys::rl_ctx,
}

// one caller thread at a time (include/rl_b200.h "Conventions"): movable, not shareable
unsafe impl Send for Ctx {}

impl Ctx {
    pub fn new(device_id: i32) -> Result<Self> {
        let mut raw = ptr::null_mut();
        let rc = unsafe { sys::rl_create(device_id, &mut raw) };
        if rc != sys::RL_OK {
            return Err(Self::error_of(ptr::null(), rc));
        }
        assert_eq!(unsafe { sys::rl_abi_version() }, sys::RL_B200_ABI_VERSION);
        Ok(Ctx { raw })
    }

    /// One context over several GPUs of one node (`rl_create_multi`): `render_ow` / `render_rtc` then use all of them
    /// — dynamic (pixel x sample-chunk) queue in GPU 0's HBM, partial sums stored over NVLink, bit-identical image.
    pub fn new_multi(device_ids: &[i32]) -> Result<Self> {
        let mut raw = ptr::null_mut();
        let rc = unsafe { sys::rl_create_multi(device_ids.as_ptr(), device_ids.len() as i32, &mut raw) };
        if rc != sys::RL_OK {
            return Err(Self::error_of(ptr::null(), rc));
        }
        assert_eq!(unsafe { sys::rl_abi_version() }, sys::RL_B200_ABI_VERSION);
        Ok(Ctx { raw })
    }

    pub fn device_count(&self) -> i32 {
        unsafe { sys::rl_device_count(self.raw) }
    }

    fn error_of(raw: *const sys::rl_ctx, rc: i32) -> Error {
        let msg = unsafe { CStr::from_ptr(sys::rl_last_error(raw)) }.to_string_lossy().into_owned();
        match rc {
            sys::RL_E_INVALID => Error::Invalid(msg),
            sys::RL_E_NO_DEVICE => Error::NoDevice(msg),
            sys::RL_E_UNSUPPORTED => Error::Unsupported(msg),
            sys::RL_E_NO_SCENE => Error::NoScene,
            sys::RL_E_OVERFLOW => Error::Overflow(msg),
            _ => Error::Cuda(msg),
        }
    }

    fn check(&self, rc: i32) -> Result<()> {
        if rc == sys::RL_OK {
            Ok(())
        } else {
            Err(Self::error_of(self.raw, rc))
        }
    }

    /// Flatten + upload + LBVH build.  Replaces walking `World.objects` / `world.hit` per ray.
    pub fn scene_upload(&mut self, scene: &SceneBuilder) -> Result<()> {
        let images: Vec<sys::rl_image> = scene
            .images
            .iter()
            .map(|(w, h, px)| sys::rl_image { width: *w, height: *h, rgb: px.as_ptr() })
            .collect();
        let d = scene.desc(&images);
        self.check(unsafe { sys::rl_scene_upload(self.raw, &d) })
    }

    pub fn render_rtc(&mut self, cam: &sys::rl_rtc_camera, aa: u32) -> Result<(Vec<f32>, sys::rl_stats)> {
        let mut out = vec![0f32; cam.hsize as usize * cam.vsize as usize * 3];
        let mut st = sys::rl_stats::default();
        self.check(unsafe { sys::rl_render_rtc(self.raw, cam, aa, out.as_mut_ptr(), &mut st) })?;
        Ok((out, st))
    }

    pub fn render_ow(&mut self, cam: &sys::rl_ow_camera, first_sample: u32) -> Result<(Vec<f32>, sys::rl_stats)> {
        let h = unsafe { sys::rl_ow_image_height(cam) } as usize;
        let mut out = vec![0f32; cam.image_width as usize * h * 3];
        let mut st = sys::rl_stats::default();
        self.check(unsafe { sys::rl_render_ow(self.raw, cam, first_sample, out.as_mut_ptr(), &mut st) })?;
        Ok((out, st))
    }

    /// `rl_render_rtc_u8`: render + `Canvas::ppm`'s 8-bit `translate` on the device (W*H*3 bytes, row-major).
    pub fn render_rtc_u8(&mut self, cam: &sys::rl_rtc_camera, aa: u32) -> Result<Vec<u8>> {
        let mut out = vec![0u8; cam.hsize as usize * cam.vsize as usize * 3];
        self.check(unsafe { sys::rl_render_rtc_u8(self.raw, cam, aa, out.as_mut_ptr(), ptr::null_mut()) })?;
        Ok(out)
    }

    /// `rl_render_ow_u8`: render + `pixel_data` / `linear_to_srgb` / `to_u8` on the device.
    pub fn render_ow_u8(&mut self, cam: &sys::rl_ow_camera, first_sample: u32) -> Result<Vec<u8>> {
        let h = unsafe { sys::rl_ow_image_height(cam) } as usize;
        let mut out = vec![0u8; cam.image_width as usize * h * 3];
        self.check(unsafe { sys::rl_render_ow_u8(self.raw, cam, first_sample, out.as_mut_ptr(), ptr::null_mut()) })?;
        Ok(out)
    }

    pub fn trace_batch(&mut self, rays: &[sys::rl_ray]) -> Result<Vec<sys::rl_hit>> {
        let mut hits = vec![sys::rl_hit::default(); rays.len()];
        self.check(unsafe { sys::rl_trace_batch(self.raw, rays.as_ptr(), rays.len() as u64, hits.as_mut_ptr()) })?;
        Ok(hits)
    }

    pub fn raw(&mut self) -> *mut sys::rl_ctx {
        self.raw
    }
}

impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { sys::rl_destroy(self.raw) }
    }
}

/// Append-only recorder of a scene tree (`rl_scene_desc`).  Node ids are indices into `nodes`.
pub struct SceneBuilder {
    pub flavor: i32,
    pub nodes: Vec<sys::rl_node>,
    pub children: Vec<i32>,
    pub params: Vec<f64>,
    pub roots: Vec<i32>,
    pub materials: Vec<sys::rl_material>,
    pub textures: Vec<sys::rl_texture>,
    pub images: Vec<(i32, i32, Vec<f32>)>,
    pub lights: Vec<sys::rl_light>,
    pub perlins: Vec<sys::rl_perlin>,
    pub max_reflection_depth: i32,
    pub void_color: [f64; 3],
    // materials / textures are shared by ADDRESS, like `&Material` / `Box<dyn Material>` in the reference
    mat_ids: HashMap<usize, i32>,
    tex_ids: HashMap<usize, i32>,
}

impl SceneBuilder {
    pub fn new(flavor: i32) -> Self {
        SceneBuilder {
            flavor,
            nodes: vec![],
            children: vec![],
            params: vec![],
            roots: vec![],
            materials: vec![],
            textures: vec![],
            images: vec![],
            lights: vec![],
            perlins: vec![],
            max_reflection_depth: 5,
            void_color: [0.0; 3],
            mat_ids: HashMap::new(),
            tex_ids: HashMap::new(),
        }
    }

    pub fn add_params(&mut self, values: &[f64]) -> i32 {
        let off = self.params.len() as i32;
        self.params.extend_from_slice(values);
        off
    }

    pub fn add_node(&mut self, kind: i32, material: i32, flags: i32, param: i32) -> i32 {
        self.nodes.push(sys::rl_node { kind, material, child_begin: -1, child_end: -1, flags, param });
        self.nodes.len() as i32 - 1
    }

    pub fn set_node_children(&mut self, node: i32, begin: i32, end: i32) {
        let n = &mut self.nodes[node as usize];
        n.child_begin = begin;
        n.child_end = end;
    }

    pub fn add_children(&mut self, ids: &[i32]) -> (i32, i32) {
        let b = self.children.len() as i32;
        self.children.extend_from_slice(ids);
        (b, self.children.len() as i32)
    }

    pub fn material_id<T: ?Sized>(&mut self, key: &T, make: impl FnOnce(&mut Self) -> sys::rl_material) -> i32 {
        let k = key as *const T as *const u8 as usize;
        if let Some(id) = self.mat_ids.get(&k) {
            return *id;
        }
        let m = make(self);
        self.materials.push(m);
        let id = self.materials.len() as i32 - 1;
        self.mat_ids.insert(k, id);
        id
    }

    pub fn texture_id<T: ?Sized>(&mut self, key: &T, make: impl FnOnce(&mut Self) -> sys::rl_texture) -> i32 {
        let k = key as *const T as *const u8 as usize;
        if let Some(id) = self.tex_ids.get(&k) {
            return *id;
        }
        let t = make(self); // may register sub-textures first
        self.textures.push(t);
        let id = self.textures.len() as i32 - 1;
        self.tex_ids.insert(k, id);
        id
    }

    /// `rl_scene_check`: validate + flatten on the host only (no GPU needed); what `Ctx::scene_upload` would accept.
    pub fn check(&self) -> std::result::Result<sys::rl_scene_info, (i32, String)> {
        let images: Vec<sys::rl_image> =
            self.images.iter().map(|(w, h, px)| sys::rl_image { width: *w, height: *h, rgb: px.as_ptr() }).collect();
        let d = self.desc(&images);
        let mut info = sys::rl_scene_info::default();
        let mut err = [0 as std::os::raw::c_char; 512];
        let rc = unsafe { sys::rl_scene_check(&d, &mut info, err.as_mut_ptr(), 512) };
        if rc == sys::RL_OK {
            Ok(info)
        } else {
            Err((rc, unsafe { CStr::from_ptr(err.as_ptr()) }.to_string_lossy().into_owned()))
        }
    }

    /// the `rl_scene_desc` view of this builder; `images` must outlive the returned struct
    pub fn desc(&self, images: &[sys::rl_image]) -> sys::rl_scene_desc {
        sys::rl_scene_desc {
            abi_version: sys::RL_B200_ABI_VERSION,
            flavor: self.flavor,
            nodes: self.nodes.as_ptr(),
            n_nodes: self.nodes.len() as i32,
            children: self.children.as_ptr(),
            n_children: self.children.len() as i32,
            params: self.params.as_ptr(),
            n_params: self.params.len() as i64,
            roots: self.roots.as_ptr(),
            n_roots: self.roots.len() as i32,
            materials: self.materials.as_ptr(),
            n_materials: self.materials.len() as i32,
            textures: self.textures.as_ptr(),
            n_textures: self.textures.len() as i32,
            images: images.as_ptr(),
            n_images: images.len() as i32,
            lights: self.lights.as_ptr(),
            n_lights: self.lights.len() as i32,
            max_reflection_depth: self.max_reflection_depth,
            void_color: self.void_color,
            perlins: self.perlins.as_ptr(),
            n_perlins: self.perlins.len() as i32,
        }
    }

    pub fn add_perlin(&mut self, p: sys::rl_perlin) -> i32 {
        self.perlins.push(p);
        self.perlins.len() as i32 - 1
    }

    pub fn add_image(&mut self, width: i32, height: i32, rgb: Vec<f32>) -> i32 {
        assert_eq!(rgb.len(), (width * height * 3) as usize);
        self.images.push((width, height, rgb));
        self.images.len() as i32 - 1
    }
}
