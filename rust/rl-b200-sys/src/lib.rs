//! `extern "C"` surface of `librl_b200.so` — a field-for-field transcription of `include/rl_b200.h`
//! (ABI version 2).  `tests/test_abi.py::test_rust_sys_matches_header` keeps this file and the header in
//! step.  Nothing here is safe; the safe wrapper is the `rl-b200` crate.
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_int, c_void};

pub const RL_B200_ABI_VERSION: i32 = 3;
pub const RL_QUEUE_SLOTS: i32 = 2;

pub const RL_OK: c_int = 0;
pub const RL_E_INVALID: c_int = -1;
pub const RL_E_NO_DEVICE: c_int = -2;
pub const RL_E_CUDA: c_int = -3;
pub const RL_E_UNSUPPORTED: c_int = -4;
pub const RL_E_NO_SCENE: c_int = -5;
pub const RL_E_OVERFLOW: c_int = -6;

pub const RL_FLAVOR_RTC: i32 = 1;
pub const RL_FLAVOR_OW: i32 = 2;

pub const RL_RTC_SPHERE: i32 = 1;
pub const RL_RTC_PLANE: i32 = 2;
pub const RL_RTC_CUBE: i32 = 3;
pub const RL_RTC_CYLINDER: i32 = 4;
pub const RL_RTC_CONE: i32 = 5;
pub const RL_RTC_TRIANGLE: i32 = 6;
pub const RL_RTC_TRANSFORMED: i32 = 7;
pub const RL_RTC_GROUP: i32 = 8;
pub const RL_RTC_BOUNDED: i32 = 9;
pub const RL_RTC_CSG: i32 = 10;
pub const RL_RTC_MESH: i32 = 11;
pub const RL_OW_SPHERE: i32 = 32;
pub const RL_OW_QUAD: i32 = 33;
pub const RL_OW_TRIANGLE: i32 = 34;
pub const RL_OW_TRANSFORM: i32 = 35;
pub const RL_OW_TRANSLATE: i32 = 36;
pub const RL_OW_BVH: i32 = 37;
pub const RL_OW_LIST: i32 = 38;
pub const RL_OW_CONSTANT_MEDIUM: i32 = 39;
pub const RL_OW_MESH: i32 = 40;

pub const RL_CSG_UNION: i32 = 0;
pub const RL_CSG_INTERSECTION: i32 = 1;
pub const RL_CSG_DIFFERENCE: i32 = 2;

pub const RL_MAT_RTC_PHONG: i32 = 1;
pub const RL_MAT_OW_LAMBERTIAN: i32 = 16;
pub const RL_MAT_OW_METAL: i32 = 17;
pub const RL_MAT_OW_DIELECTRIC: i32 = 18;
pub const RL_MAT_OW_DIFFUSE_LIGHT: i32 = 19;
pub const RL_MAT_OW_ISOTROPIC: i32 = 20;

pub const RL_TEX_RTC_STRIPE: i32 = 1;
pub const RL_TEX_RTC_CHECKER3D: i32 = 2;
pub const RL_TEX_RTC_GRADIENT: i32 = 3;
pub const RL_TEX_RTC_RING: i32 = 4;
pub const RL_TEX_OW_SOLID: i32 = 16;
pub const RL_TEX_OW_CHECKER: i32 = 17;
pub const RL_TEX_OW_IMAGE: i32 = 18;
pub const RL_TEX_OW_NOISE: i32 = 19;

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct rl_node {
    pub kind: i32,
    pub material: i32,
    pub child_begin: i32,
    pub child_end: i32,
    pub flags: i32,
    pub param: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct rl_material {
    pub kind: i32,
    pub texture: i32,
    pub color: [f64; 3],
    pub ambient: f64,
    pub diffuse: f64,
    pub specular: f64,
    pub shininess: f64,
    pub reflectivity: f64,
    pub transparency: f64,
    pub refractive_index: f64,
    pub fuzz: f64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct rl_texture {
    pub kind: i32,
    pub tex_a: i32,
    pub tex_b: i32,
    pub image: i32,
    pub a: [f64; 3],
    pub b: [f64; 3],
    pub scale: f64,
    pub transform: [f64; 16],
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct rl_image {
    pub width: i32,
    pub height: i32,
    pub rgb: *const f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct rl_perlin {
    pub randvec: [[f64; 3]; 256],
    pub perm_x: [i32; 256],
    pub perm_y: [i32; 256],
    pub perm_z: [i32; 256],
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct rl_light {
    pub position: [f64; 3],
    pub intensity: [f64; 3],
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct rl_scene_desc {
    pub abi_version: i32,
    pub flavor: i32,
    pub nodes: *const rl_node,
    pub n_nodes: i32,
    pub children: *const i32,
    pub n_children: i32,
    pub params: *const f64,
    pub n_params: i64,
    pub roots: *const i32,
    pub n_roots: i32,
    pub materials: *const rl_material,
    pub n_materials: i32,
    pub textures: *const rl_texture,
    pub n_textures: i32,
    pub images: *const rl_image,
    pub n_images: i32,
    pub lights: *const rl_light,
    pub n_lights: i32,
    pub max_reflection_depth: i32,
    pub void_color: [f64; 3],
    pub perlins: *const rl_perlin,
    pub n_perlins: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct rl_rtc_camera {
    pub hsize: i32,
    pub vsize: i32,
    pub fov: f64,
    pub transform: [f64; 16],
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct rl_ow_camera {
    pub aspect_ratio: f64,
    pub image_width: i32,
    pub samples_per_pixel: i32,
    pub max_depth: i32,
    pub _pad: i32,
    pub vfov: f64,
    pub lookfrom: [f64; 3],
    pub lookat: [f64; 3],
    pub vup: [f64; 3],
    pub defocus_angle: f64,
    pub focus_dist: f64,
    pub background: [f64; 3],
    pub seed: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct rl_ray {
    pub origin: [f32; 3],
    pub direction: [f32; 3],
    pub time: f32,
    pub _pad: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct rl_hit {
    pub node: i32,
    pub t: f32,
    pub u: f32,
    pub v: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct rl_stats {
    pub rays: u64,
    pub node_visits: u64,
    pub prim_tests: u64,
    pub tri_tests: u64,
    pub shades: u64,
    pub samples: u64,
    pub overflow: u64,
    pub kernel_ms: f32,
    pub upload_ms: f32,
    pub kernel_launches: i32,
    pub _pad: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct rl_scene_info {
    pub flavor: i32,
    pub n_prims: i32,
    pub n_bvh_prims: i32,
    pub n_bvh_nodes: i32,
    pub n_materials: i32,
    pub n_textures: i32,
    pub n_lights: i32,
    pub has_transparency: i32,
    pub device_bytes: i64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct rl_lbvh_host {
    pub prim_aabb: *mut f32,
    pub prim_node: *mut i32,
    pub morton: *mut u64,
    pub sorted_prim: *mut i32,
    pub left: *mut i32,
    pub right: *mut i32,
    pub parent: *mut i32,
    pub node_aabb: *mut f32,
    pub scene_lo: [f32; 3],
    pub scene_hi: [f32; 3],
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct rl_job {
    pub x0: i32,
    pub y0: i32,
    pub x1: i32,
    pub y1: i32,
    pub chunk_begin: i32,
    pub chunk_end: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct rl_obj_info {
    pub n_vertices: i32,
    pub n_normals: i32,
    pub n_texcoords: i32,
    pub n_triangles: i32,
    pub n_groups: i32,
    pub ignored: i32,
    pub kernel_launches: i32,
    pub _pad: i32,
    pub bounds: [f64; 6],
}

#[repr(C)]
pub struct rl_ctx {
    _opaque: [u8; 0],
}

extern "C" {
    pub fn rl_create(device_id: c_int, out: *mut *mut rl_ctx) -> c_int;
    pub fn rl_create_multi(device_ids: *const i32, n: i32, out: *mut *mut rl_ctx) -> c_int;
    pub fn rl_device_count(ctx: *const rl_ctx) -> c_int;
    pub fn rl_destroy(ctx: *mut rl_ctx);
    pub fn rl_last_error(ctx: *const rl_ctx) -> *const c_char;
    pub fn rl_abi_version() -> c_int;
    pub fn rl_device_info(ctx: *mut rl_ctx, sm_count: *mut c_int, cc_major: *mut c_int, cc_minor: *mut c_int,
                          hbm_bytes: *mut i64) -> c_int;
    pub fn rl_synchronize(ctx: *mut rl_ctx) -> c_int;
    pub fn rl_measure_peaks(ctx: *mut rl_ctx, fp32_tflops: *mut f64, l2_gbs: *mut f64, hbm_gbs: *mut f64) -> c_int;

    pub fn rl_scene_upload(ctx: *mut rl_ctx, scene: *const rl_scene_desc) -> c_int;
    pub fn rl_scene_info_get(ctx: *mut rl_ctx, out: *mut rl_scene_info) -> c_int;
    pub fn rl_scene_check(scene: *const rl_scene_desc, out: *mut rl_scene_info, err: *mut c_char, err_cap: i32) -> c_int;
    pub fn rl_lbvh_download(ctx: *mut rl_ctx, out: *mut rl_lbvh_host) -> c_int;

    pub fn rl_obj_parse(ctx: *mut rl_ctx, text: *const c_char, len: u64, flavor: i32, info: *mut rl_obj_info) -> c_int;
    pub fn rl_obj_download(ctx: *mut rl_ctx, tri_p: *mut f64, tri_n: *mut f64, tri_uv: *mut f64, flags: *mut u8) -> c_int;

    pub fn rl_trace_batch(ctx: *mut rl_ctx, rays: *const rl_ray, n: u64, out: *mut rl_hit) -> c_int;
    pub fn rl_trace_batch_ex(ctx: *mut rl_ctx, rays: *const rl_ray, self_nodes: *const i32, n: u64, out: *mut rl_hit) -> c_int;

    pub fn rl_render_rtc(ctx: *mut rl_ctx, cam: *const rl_rtc_camera, anti_aliasing_samples: u32, out_rgb: *mut f32,
                         stats: *mut rl_stats) -> c_int;
    pub fn rl_render_ow(ctx: *mut rl_ctx, cam: *const rl_ow_camera, first_sample: u32, out_rgb_sum: *mut f32,
                        stats: *mut rl_stats) -> c_int;
    pub fn rl_render_rtc_u8(ctx: *mut rl_ctx, cam: *const rl_rtc_camera, anti_aliasing_samples: u32, out_rgb8: *mut u8,
                            stats: *mut rl_stats) -> c_int;
    pub fn rl_render_ow_u8(ctx: *mut rl_ctx, cam: *const rl_ow_camera, first_sample: u32, out_rgb8: *mut u8,
                           stats: *mut rl_stats) -> c_int;
    pub fn rl_ow_image_height(cam: *const rl_ow_camera) -> c_int;
    pub fn rl_ow_num_chunks(cam: *const rl_ow_camera) -> c_int;

    pub fn rl_render_rtc_device(ctx: *mut rl_ctx, cam: *const rl_rtc_camera, anti_aliasing_samples: u32,
                                jobs: *const rl_job, n_jobs: i32, d_out_rgb: *mut c_void, stream: *mut c_void,
                                stats: *mut rl_stats) -> c_int;
    pub fn rl_render_ow_device(ctx: *mut rl_ctx, cam: *const rl_ow_camera, first_sample: u32, jobs: *const rl_job,
                               n_jobs: i32, d_partial: *mut c_void, stream: *mut c_void, stats: *mut rl_stats) -> c_int;
    pub fn rl_ow_reduce_device(ctx: *mut rl_ctx, cam: *const rl_ow_camera, d_partial: *const c_void,
                               d_out_rgb_sum: *mut c_void, stream: *mut c_void) -> c_int;

    pub fn rl_queue_export(ctx: *mut rl_ctx, handle64: *mut c_void) -> c_int;
    pub fn rl_queue_import(ctx: *mut rl_ctx, handle64: *const c_void) -> c_int;
    pub fn rl_queue_reset(ctx: *mut rl_ctx, stream: *mut c_void, slot: i32) -> c_int;
    pub fn rl_queue_completed(ctx: *mut rl_ctx, stream: *mut c_void, slot: i32, items: *mut u64) -> c_int;
    pub fn rl_partial_export(ctx: *mut rl_ctx, bytes_per_slot: u64, n_slots: i32, handle64: *mut c_void) -> c_int;
    pub fn rl_partial_import(ctx: *mut rl_ctx, handle64: *const c_void, bytes_per_slot: u64, n_slots: i32) -> c_int;
    pub fn rl_render_ow_shared(ctx: *mut rl_ctx, cam: *const rl_ow_camera, first_sample: u32, jobs: *const rl_job,
                               n_jobs: i32, d_partial: *mut c_void, stream: *mut c_void, slot: i32) -> c_int;
    pub fn rl_ow_reduce_shared(ctx: *mut rl_ctx, cam: *const rl_ow_camera, slot: i32, d_out_rgb_sum: *mut c_void,
                               stream: *mut c_void) -> c_int;
    pub fn rl_ow_job_items(cam: *const rl_ow_camera, jobs: *const rl_job, n_jobs: i32) -> i64;
    pub fn rl_set_option(ctx: *mut rl_ctx, name: *const c_char, value: i32) -> c_int;

    pub fn rl_set_instrumented(ctx: *mut rl_ctx, enabled: c_int) -> c_int;
}
