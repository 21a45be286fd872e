"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/liboracle.so (the f64 CPU restatement of the reference's ray loop).
Only tests/, __graft_entry__.smoke() and bench.py's `cpu_baseline` / `--impl reference` legs may
import this module; nothing under rendering_learning_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from rendering_learning_b200 import _abi as A

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
_lib = None


def build(force: bool = False):
    srcs = [os.path.join(HERE, f) for f in ("oracle_rtc.cpp", "oracle_ow.cpp", "oracle_lbvh.cpp")]
    hdr = os.path.join(HERE, "..", "include", "rl_b200.h")
    if not force and os.path.exists(LIB):
        newest = max(os.path.getmtime(p) for p in srcs + [hdr] if os.path.exists(p))
        if os.path.getmtime(LIB) >= newest:
            return LIB
    subprocess.check_call(["make", "-C", HERE, "-B", "liboracle.so"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.orc_rtc_shadow.restype = C.c_double
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


# ---- RTC -----------------------------------------------------------------------------------------

def rtc_render(desc, cam: A.rl_rtc_camera, aa: int = 1, threads: int = 0) -> np.ndarray:
    d = desc.freeze()
    out = np.zeros((cam.vsize, cam.hsize, 3), np.float64)
    rc = lib().orc_rtc_render(C.byref(d), C.byref(cam), C.c_uint32(aa), _dp(out), C.c_int(threads))
    if rc != 0:
        raise RuntimeError(f"orc_rtc_render failed: {rc}")
    return out


def rtc_camera_rays(cam: A.rl_rtc_camera, aa: int = 1) -> np.ndarray:
    rays = np.zeros((cam.vsize * cam.hsize * aa * aa, 6), np.float64)
    rc = lib().orc_rtc_camera_rays(C.byref(cam), C.c_uint32(aa), _dp(rays))
    if rc != 0:
        raise RuntimeError("orc_rtc_camera_rays failed")
    return rays


def rtc_trace(desc, rays: np.ndarray, threads: int = 0):
    d = desc.freeze()
    rays = np.ascontiguousarray(rays, np.float64).reshape(-1, 6)
    n = rays.shape[0]
    node = np.zeros(n, np.int32)
    t = np.zeros(n, np.float64)
    second = np.zeros(n, np.float64)
    rc = lib().orc_rtc_trace(C.byref(d), _dp(rays), C.c_uint64(n), _ip(node), _dp(t), _dp(second),
                             C.c_int(threads))
    if rc != 0:
        raise RuntimeError("orc_rtc_trace failed")
    return node, t, second


def rtc_intersect(desc, ray, cap: int = 64):
    d = desc.freeze()
    r = np.asarray(ray, np.float64).reshape(6)
    ts = np.zeros(cap)
    nodes = np.zeros(cap, np.int32)
    normals = np.zeros((cap, 3))
    colors = np.zeros((cap, 3))
    n = lib().orc_rtc_intersect(C.byref(d), _dp(r), cap, _dp(ts), _ip(nodes), _dp(normals), _dp(colors))
    if n < 0:
        raise RuntimeError("orc_rtc_intersect failed")
    return ts[:n], nodes[:n], normals[:n], colors[:n]


def rtc_color_at(desc, ray, remaining: int = -1):
    d = desc.freeze()
    r = np.asarray(ray, np.float64).reshape(6)
    out = np.zeros(3)
    rc = lib().orc_rtc_color_at(C.byref(d), _dp(r), C.c_int(remaining), _dp(out))
    if rc != 0:
        raise RuntimeError("orc_rtc_color_at failed")
    return out


def rtc_prepare(desc, ray, index: int = -1) -> dict:
    d = desc.freeze()
    r = np.asarray(ray, np.float64).reshape(6)
    o = np.zeros(32)
    rc = lib().orc_rtc_prepare(C.byref(d), _dp(r), C.c_int(index), _dp(o))
    if rc != 0:
        raise RuntimeError(f"orc_rtc_prepare failed: {rc}")
    return {"t": o[0], "object": int(o[1]), "point": o[2:5], "eye_v": o[5:8], "normal_v": o[8:11],
            "inside": bool(o[11]), "over_point": o[12:15], "under_point": o[15:18],
            "reflect_v": o[18:21], "n1": o[21], "n2": o[22], "schlick": o[23], "shadow": o[24]}


def rtc_lighting(desc, material, light, point, color, eye, normal, attenuation):
    d = desc.freeze()
    i = np.asarray(list(point) + list(color) + list(eye) + list(normal) + [attenuation], np.float64)
    out = np.zeros(3)
    lib().orc_rtc_lighting(C.byref(d), C.c_int(material), C.c_int(light), _dp(i), _dp(out))
    return out


def rtc_shadow(desc, point, light: int = 0) -> float:
    d = desc.freeze()
    p = np.asarray(point, np.float64)
    return float(lib().orc_rtc_shadow(C.byref(d), _dp(p), C.c_int(light)))


def rtc_invert(m16):
    m = np.asarray(m16, np.float64).reshape(16)
    out = np.zeros(16)
    rc = lib().orc_rtc_invert(_dp(m), _dp(out))
    return None if rc != 0 else out.reshape(4, 4)


# ---- OW ------------------------------------------------------------------------------------------

def ow_image_height(cam: A.rl_ow_camera) -> int:
    return int(lib().orc_ow_image_height(C.byref(cam)))


def ow_render(desc, cam: A.rl_ow_camera, first_sample: int = 0, rows=None, threads: int = 0):
    """Returns (sums [H,W,3] f64, rays cast).  rows=(y0,y1) renders only those rows (others stay 0)."""
    d = desc.freeze()
    h = ow_image_height(cam)
    out = np.zeros((h, cam.image_width, 3), np.float64)
    y0, y1 = rows if rows is not None else (0, h)
    rays = C.c_uint64(0)
    rc = lib().orc_ow_render(C.byref(d), C.byref(cam), C.c_uint32(first_sample), _dp(out), C.c_int(y0),
                             C.c_int(y1), C.c_int(threads), C.byref(rays))
    if rc != 0:
        raise RuntimeError(f"orc_ow_render failed: {rc}")
    return out, rays.value


def ow_trace(desc, rays: np.ndarray, threads: int = 0):
    d = desc.freeze()
    rays = np.ascontiguousarray(rays, np.float64).reshape(-1, 7)
    n = rays.shape[0]
    node = np.zeros(n, np.int32)
    t = np.zeros(n, np.float64)
    uv = np.zeros((n, 2), np.float64)
    rc = lib().orc_ow_trace(C.byref(d), _dp(rays), C.c_uint64(n), _ip(node), _dp(t), _dp(uv), C.c_int(threads))
    if rc != 0:
        raise RuntimeError("orc_ow_trace failed")
    return node, t, uv


def ow_trace_self(desc, rays: np.ndarray, self_nodes: np.ndarray, threads: int = 0):
    """ow_trace for rays that start on a surface (self_nodes[i] = that leaf, -1 none): the oracle then applies the
    device path's start-on-surface rule (oracle_ow.cpp tl_self_node) instead of relying on f64 + t_min = 1e-10."""
    d = desc.freeze()
    rays = np.ascontiguousarray(rays, np.float64).reshape(-1, 7)
    sn = np.ascontiguousarray(self_nodes, np.int32).reshape(-1)
    n = rays.shape[0]
    node = np.zeros(n, np.int32)
    t = np.zeros(n, np.float64)
    uv = np.zeros((n, 2), np.float64)
    rc = lib().orc_ow_trace_self(C.byref(d), _dp(rays), _ip(sn), C.c_uint64(n), _ip(node), _dp(t), _dp(uv), C.c_int(threads))
    if rc != 0:
        raise RuntimeError("orc_ow_trace_self failed")
    return node, t, uv


def ow_bounce_rays(desc, cam: A.rl_ow_camera, bounce: int):
    """The reference's own ray at bounce `bounce` of every pixel's first sample: (rays [n,7] f64, self_nodes [n] —
    -1 camera ray, -2 the path ended earlier)."""
    d = desc.freeze()
    h = ow_image_height(cam)
    n = h * cam.image_width
    rays = np.zeros((n, 7), np.float64)
    sn = np.zeros(n, np.int32)
    if lib().orc_ow_bounce_rays(C.byref(d), C.byref(cam), C.c_int(bounce), _dp(rays), _ip(sn)) != 0:
        raise RuntimeError("orc_ow_bounce_rays failed")
    return rays, sn


def ow_tex_value(desc, tex: int, uvp: np.ndarray) -> np.ndarray:
    """Texture::value at rows (u, v, x, y, z)"""
    d = desc.freeze()
    uvp = np.ascontiguousarray(uvp, np.float64).reshape(-1, 5)
    out = np.zeros((uvp.shape[0], 3), np.float64)
    if lib().orc_ow_tex_value(C.byref(d), C.c_int(tex), C.c_uint64(uvp.shape[0]), _dp(uvp), _dp(out)) != 0:
        raise RuntimeError("orc_ow_tex_value failed")
    return out


def ow_camera_rays(cam: A.rl_ow_camera) -> np.ndarray:
    h = ow_image_height(cam)
    rays = np.zeros((h * cam.image_width, 7), np.float64)
    lib().orc_ow_camera_rays(C.byref(cam), _dp(rays))
    return rays


def chacha8_u64(seed: int, stream: int, n: int) -> np.ndarray:
    out = np.zeros(n, np.uint64)
    lib().orc_chacha8_u64(C.c_uint64(seed), C.c_uint64(stream), C.c_int(n),
                          out.ctypes.data_as(C.POINTER(C.c_uint64)))
    return out


# ---- LBVH host rebuild ---------------------------------------------------------------------------

def lbvh_build(prim_aabb: np.ndarray) -> dict:
    a = np.ascontiguousarray(prim_aabb, np.float32).reshape(-1, 6)
    n = a.shape[0]
    m = max(n - 1, 0)
    out = {"morton": np.zeros(max(n, 1), np.uint64), "sorted_prim": np.zeros(max(n, 1), np.int32),
           "left": np.zeros(max(m, 1), np.int32), "right": np.zeros(max(m, 1), np.int32),
           "parent": np.zeros(max(n + m, 1), np.int32), "node_aabb": np.zeros((max(m, 1), 6), np.float32),
           "bounds": np.zeros(6, np.float32)}
    fp = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
    lib().orc_lbvh_build(fp(a), C.c_int(n), out["morton"].ctypes.data_as(C.POINTER(C.c_uint64)),
                         _ip(out["sorted_prim"]), _ip(out["left"]), _ip(out["right"]), _ip(out["parent"]),
                         fp(out["node_aabb"]), fp(out["bounds"]))
    out["morton"], out["sorted_prim"] = out["morton"][:n], out["sorted_prim"][:n]
    out["left"], out["right"], out["node_aabb"] = out["left"][:m], out["right"][:m], out["node_aabb"][:m]
    out["parent"] = out["parent"][:n + m]
    return out
