"""ctypes image of include/rl_b200.h and the loader for librl_b200.so.

There is no CPU fallback: if the CUDA library is missing or no sm_100 device is visible the
product path raises (`RlError`) instead of routing anywhere else.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# RL_B200_DEBUG=1 loads the bounds-asserting build (librl_b200_debug.so, `__graft_entry__.build(debug=True)`): the same
# CUDA code with every index checked against the scene's counts — still the device path, never a CPU one
LIB_PATH = os.path.join(HERE, "librl_b200_debug.so" if os.environ.get("RL_B200_DEBUG") == "1" else "librl_b200.so")

RL_B200_ABI_VERSION = 3
RL_QUEUE_SLOTS = 2

RL_OK = 0
RL_E_INVALID = -1
RL_E_NO_DEVICE = -2
RL_E_CUDA = -3
RL_E_UNSUPPORTED = -4
RL_E_NO_SCENE = -5
RL_E_OVERFLOW = -6

RL_FLAVOR_RTC = 1
RL_FLAVOR_OW = 2

# node kinds
RL_RTC_SPHERE = 1
RL_RTC_PLANE = 2
RL_RTC_CUBE = 3
RL_RTC_CYLINDER = 4
RL_RTC_CONE = 5
RL_RTC_TRIANGLE = 6
RL_RTC_TRANSFORMED = 7
RL_RTC_GROUP = 8
RL_RTC_BOUNDED = 9
RL_RTC_CSG = 10
RL_RTC_MESH = 11
RL_OW_SPHERE = 32
RL_OW_QUAD = 33
RL_OW_TRIANGLE = 34
RL_OW_TRANSFORM = 35
RL_OW_TRANSLATE = 36
RL_OW_BVH = 37
RL_OW_LIST = 38
RL_OW_CONSTANT_MEDIUM = 39
RL_OW_MESH = 40

RL_CSG_UNION = 0
RL_CSG_INTERSECTION = 1
RL_CSG_DIFFERENCE = 2

RL_MAT_RTC_PHONG = 1
RL_MAT_OW_LAMBERTIAN = 16
RL_MAT_OW_METAL = 17
RL_MAT_OW_DIELECTRIC = 18
RL_MAT_OW_DIFFUSE_LIGHT = 19
RL_MAT_OW_ISOTROPIC = 20

RL_TEX_RTC_STRIPE = 1
RL_TEX_RTC_CHECKER3D = 2
RL_TEX_RTC_GRADIENT = 3
RL_TEX_RTC_RING = 4
RL_TEX_OW_SOLID = 16
RL_TEX_OW_CHECKER = 17
RL_TEX_OW_IMAGE = 18
RL_TEX_OW_NOISE = 19


class rl_node(C.Structure):
    _fields_ = [("kind", C.c_int32), ("material", C.c_int32), ("child_begin", C.c_int32),
                ("child_end", C.c_int32), ("flags", C.c_int32), ("param", C.c_int32)]


class rl_material(C.Structure):
    _fields_ = [("kind", C.c_int32), ("texture", C.c_int32), ("color", C.c_double * 3),
                ("ambient", C.c_double), ("diffuse", C.c_double), ("specular", C.c_double),
                ("shininess", C.c_double), ("reflectivity", C.c_double),
                ("transparency", C.c_double), ("refractive_index", C.c_double),
                ("fuzz", C.c_double)]


class rl_texture(C.Structure):
    _fields_ = [("kind", C.c_int32), ("tex_a", C.c_int32), ("tex_b", C.c_int32),
                ("image", C.c_int32), ("a", C.c_double * 3), ("b", C.c_double * 3),
                ("scale", C.c_double), ("transform", C.c_double * 16)]


class rl_image(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("rgb", C.POINTER(C.c_float))]


class rl_perlin(C.Structure):
    _fields_ = [("randvec", (C.c_double * 3) * 256), ("perm_x", C.c_int32 * 256),
                ("perm_y", C.c_int32 * 256), ("perm_z", C.c_int32 * 256)]


class rl_light(C.Structure):
    _fields_ = [("position", C.c_double * 3), ("intensity", C.c_double * 3)]


class rl_scene_desc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("flavor", C.c_int32),
                ("nodes", C.POINTER(rl_node)), ("n_nodes", C.c_int32),
                ("children", C.POINTER(C.c_int32)), ("n_children", C.c_int32),
                ("params", C.POINTER(C.c_double)), ("n_params", C.c_int64),
                ("roots", C.POINTER(C.c_int32)), ("n_roots", C.c_int32),
                ("materials", C.POINTER(rl_material)), ("n_materials", C.c_int32),
                ("textures", C.POINTER(rl_texture)), ("n_textures", C.c_int32),
                ("images", C.POINTER(rl_image)), ("n_images", C.c_int32),
                ("lights", C.POINTER(rl_light)), ("n_lights", C.c_int32),
                ("max_reflection_depth", C.c_int32), ("void_color", C.c_double * 3),
                ("perlins", C.POINTER(rl_perlin)), ("n_perlins", C.c_int32)]


class rl_rtc_camera(C.Structure):
    _fields_ = [("hsize", C.c_int32), ("vsize", C.c_int32), ("fov", C.c_double),
                ("transform", C.c_double * 16)]


class rl_ow_camera(C.Structure):
    _fields_ = [("aspect_ratio", C.c_double), ("image_width", C.c_int32),
                ("samples_per_pixel", C.c_int32), ("max_depth", C.c_int32), ("_pad", C.c_int32),
                ("vfov", C.c_double), ("lookfrom", C.c_double * 3), ("lookat", C.c_double * 3),
                ("vup", C.c_double * 3), ("defocus_angle", C.c_double), ("focus_dist", C.c_double),
                ("background", C.c_double * 3), ("seed", C.c_uint64)]


class rl_ray(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("direction", C.c_float * 3), ("time", C.c_float),
                ("_pad", C.c_float)]


class rl_hit(C.Structure):
    _fields_ = [("node", C.c_int32), ("t", C.c_float), ("u", C.c_float), ("v", C.c_float)]


class rl_stats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("node_visits", C.c_uint64), ("prim_tests", C.c_uint64),
                ("tri_tests", C.c_uint64), ("shades", C.c_uint64), ("samples", C.c_uint64),
                ("overflow", C.c_uint64), ("kernel_ms", C.c_float), ("upload_ms", C.c_float),
                ("kernel_launches", C.c_int32), ("_pad", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("_")}


class rl_scene_info(C.Structure):
    _fields_ = [("flavor", C.c_int32), ("n_prims", C.c_int32), ("n_bvh_prims", C.c_int32),
                ("n_bvh_nodes", C.c_int32), ("n_materials", C.c_int32), ("n_textures", C.c_int32),
                ("n_lights", C.c_int32), ("has_transparency", C.c_int32),
                ("device_bytes", C.c_int64)]


class rl_lbvh_host(C.Structure):
    _fields_ = [("prim_aabb", C.POINTER(C.c_float)), ("prim_node", C.POINTER(C.c_int32)),
                ("morton", C.POINTER(C.c_uint64)), ("sorted_prim", C.POINTER(C.c_int32)),
                ("left", C.POINTER(C.c_int32)), ("right", C.POINTER(C.c_int32)),
                ("parent", C.POINTER(C.c_int32)), ("node_aabb", C.POINTER(C.c_float)),
                ("scene_lo", C.c_float * 3), ("scene_hi", C.c_float * 3)]


class rl_job(C.Structure):
    _fields_ = [("x0", C.c_int32), ("y0", C.c_int32), ("x1", C.c_int32), ("y1", C.c_int32),
                ("chunk_begin", C.c_int32), ("chunk_end", C.c_int32)]


class rl_obj_info(C.Structure):
    _fields_ = [("n_vertices", C.c_int32), ("n_normals", C.c_int32), ("n_texcoords", C.c_int32), ("n_triangles", C.c_int32),
                ("n_groups", C.c_int32), ("ignored", C.c_int32), ("kernel_launches", C.c_int32), ("_pad", C.c_int32),
                ("bounds", C.c_double * 6)]


# every symbol include/rl_b200.h declares, with its ctypes signature
_P = C.c_void_p
SYMBOLS = {
    "rl_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "rl_create_multi": (C.c_int, [C.POINTER(C.c_int32), C.c_int32, C.POINTER(_P)]),
    "rl_device_count": (C.c_int, [_P]),
    "rl_destroy": (None, [_P]),
    "rl_last_error": (C.c_char_p, [_P]),
    "rl_abi_version": (C.c_int, []),
    "rl_device_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                 C.POINTER(C.c_int64)]),
    "rl_synchronize": (C.c_int, [_P]),
    "rl_measure_peaks": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "rl_scene_upload": (C.c_int, [_P, C.POINTER(rl_scene_desc)]),
    "rl_scene_info_get": (C.c_int, [_P, C.POINTER(rl_scene_info)]),
    "rl_scene_check": (C.c_int, [C.POINTER(rl_scene_desc), C.POINTER(rl_scene_info), C.c_char_p, C.c_int32]),
    "rl_lbvh_download": (C.c_int, [_P, C.POINTER(rl_lbvh_host)]),
    "rl_obj_parse": (C.c_int, [_P, C.c_char_p, C.c_uint64, C.c_int32, C.POINTER(rl_obj_info)]),
    "rl_obj_download": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint8)]),
    "rl_trace_batch": (C.c_int, [_P, C.POINTER(rl_ray), C.c_uint64, C.POINTER(rl_hit)]),
    "rl_trace_batch_ex": (C.c_int, [_P, C.POINTER(rl_ray), C.POINTER(C.c_int32), C.c_uint64, C.POINTER(rl_hit)]),
    "rl_render_rtc": (C.c_int, [_P, C.POINTER(rl_rtc_camera), C.c_uint32, C.POINTER(C.c_float),
                                C.POINTER(rl_stats)]),
    "rl_render_ow": (C.c_int, [_P, C.POINTER(rl_ow_camera), C.c_uint32, C.POINTER(C.c_float),
                               C.POINTER(rl_stats)]),
    "rl_render_rtc_u8": (C.c_int, [_P, C.POINTER(rl_rtc_camera), C.c_uint32, C.POINTER(C.c_uint8), C.POINTER(rl_stats)]),
    "rl_render_ow_u8": (C.c_int, [_P, C.POINTER(rl_ow_camera), C.c_uint32, C.POINTER(C.c_uint8), C.POINTER(rl_stats)]),
    "rl_ow_image_height": (C.c_int, [C.POINTER(rl_ow_camera)]),
    "rl_ow_num_chunks": (C.c_int, [C.POINTER(rl_ow_camera)]),
    "rl_render_rtc_device": (C.c_int, [_P, C.POINTER(rl_rtc_camera), C.c_uint32,
                                       C.POINTER(rl_job), C.c_int32, _P, _P,
                                       C.POINTER(rl_stats)]),
    "rl_render_ow_device": (C.c_int, [_P, C.POINTER(rl_ow_camera), C.c_uint32, C.POINTER(rl_job),
                                      C.c_int32, _P, _P, C.POINTER(rl_stats)]),
    "rl_ow_reduce_device": (C.c_int, [_P, C.POINTER(rl_ow_camera), _P, _P, _P]),
    "rl_set_instrumented": (C.c_int, [_P, C.c_int]),
    "rl_queue_export": (C.c_int, [_P, C.c_void_p]),
    "rl_queue_import": (C.c_int, [_P, C.c_void_p]),
    "rl_queue_reset": (C.c_int, [_P, _P, C.c_int32]),
    "rl_queue_completed": (C.c_int, [_P, _P, C.c_int32, C.POINTER(C.c_uint64)]),
    "rl_partial_export": (C.c_int, [_P, C.c_uint64, C.c_int32, C.c_void_p]),
    "rl_partial_import": (C.c_int, [_P, C.c_void_p, C.c_uint64, C.c_int32]),
    "rl_render_ow_shared": (C.c_int, [_P, C.POINTER(rl_ow_camera), C.c_uint32, C.POINTER(rl_job), C.c_int32, _P, _P,
                                      C.c_int32]),
    "rl_ow_reduce_shared": (C.c_int, [_P, C.POINTER(rl_ow_camera), C.c_int32, _P, _P]),
    "rl_ow_job_items": (C.c_int64, [C.POINTER(rl_ow_camera), C.POINTER(rl_job), C.c_int32]),
    "rl_set_option": (C.c_int, [_P, C.c_char_p, C.c_int32]),
}


class RlError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"rl_b200 error {code}: {message}")
        self.code = code
        self.message = message


_lib = None


def load_library(path: str | None = None):
    """dlopen librl_b200.so (no GPU needed for this) and attach signatures."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RlError(RL_E_NO_DEVICE,
                      f"{p} is missing — build it with `python -c 'import __graft_entry__ as g; "
                      f"g.build()'`. There is no CPU fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib
