"""Thin object wrapper over the C ABI (include/rl_b200.h).  One Context = one rl_ctx: one GPU, or — Context([0, 1, ...])
— several GPUs of one node behind rl_create_multi."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi as A
from .desc import SceneDesc


class Context:
    def __init__(self, device_id=0):
        self.lib = A.load_library()
        h = C.c_void_p()
        if isinstance(device_id, (list, tuple)):
            ids = (C.c_int32 * len(device_id))(*[int(d) for d in device_id])
            rc = self.lib.rl_create_multi(ids, len(device_id), C.byref(h))
            device_id = int(device_id[0]) if len(device_id) else 0
        else:
            rc = self.lib.rl_create(int(device_id), C.byref(h))
        if rc != A.RL_OK:
            msg = self.lib.rl_last_error(None)
            raise A.RlError(rc, (msg or b"").decode())
        self.h = h
        self.device_id = device_id
        self._scene = None
        # True: frames returned by render_* live in page-locked host memory (torch's caching host allocator recycles the
        # block when the array is dropped), so the device -> host copy is one DMA instead of a staged copy into fresh
        # pageable pages — at 4K that copy, not the kernel, is the RTC frame time.  Default False: what a plain caller has.
        self.pinned_frames = False

    def _frame(self, shape, dtype):
        if self.pinned_frames:
            import torch
            t = torch.empty(shape, dtype=torch.float32 if dtype == np.float32 else torch.uint8, pin_memory=True)
            return t.numpy()  # shares the tensor's storage and keeps it alive
        return np.empty(shape, dtype)

    def close(self):
        if getattr(self, "h", None):
            self.lib.rl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != A.RL_OK:
            raise A.RlError(rc, (self.lib.rl_last_error(self.h) or b"").decode())

    # ---- device / scene ----------------------------------------------------------------------
    def device_info(self) -> dict:
        sm, ma, mi, hb = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        self._check(self.lib.rl_device_info(self.h, C.byref(sm), C.byref(ma), C.byref(mi),
                                            C.byref(hb)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "hbm_bytes": hb.value}

    def synchronize(self):
        self._check(self.lib.rl_synchronize(self.h))

    def device_count(self) -> int:
        return int(self.lib.rl_device_count(self.h))

    def set_option(self, name: str, value: int):
        """Scheduling parameters of the OW kernel (rl_set_option): never change the image."""
        self._check(self.lib.rl_set_option(self.h, name.encode(), int(value)))

    def measure_peaks(self) -> dict:
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._check(self.lib.rl_measure_peaks(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"fp32_tflops": a.value, "l2_gbs": b.value, "hbm_gbs": c.value}

    def scene_upload(self, desc: SceneDesc):
        """rl_scene_upload.  A frozen SceneDesc is an immutable snapshot, so uploading the SAME description object again
        is a no-op: a caller that keeps its lowered scene (`Camera.render(scene_desc)`) pays for flattening, H2D and the
        LBVH build once."""
        if self._scene is desc and desc._frozen is not None:
            return
        self._scene = None
        d = desc.freeze()
        self._check(self.lib.rl_scene_upload(self.h, C.byref(d)))
        self._scene = desc

    def scene_info(self) -> A.rl_scene_info:
        info = A.rl_scene_info()
        self._check(self.lib.rl_scene_info_get(self.h, C.byref(info)))
        return info

    def set_instrumented(self, enabled: bool):
        self._check(self.lib.rl_set_instrumented(self.h, int(bool(enabled))))

    def lbvh_download(self) -> dict:
        info = self.scene_info()
        n, m = info.n_bvh_prims, info.n_bvh_nodes
        out = {
            "prim_aabb": np.zeros((max(n, 1), 6), np.float32),
            "prim_node": np.zeros(max(n, 1), np.int32),
            "morton": np.zeros(max(n, 1), np.uint64),
            "sorted_prim": np.zeros(max(n, 1), np.int32),
            "left": np.zeros(max(m, 1), np.int32),
            "right": np.zeros(max(m, 1), np.int32),
            "parent": np.zeros(max(n + m, 1), np.int32),
            "node_aabb": np.zeros((max(m, 1), 6), np.float32),
        }
        h = A.rl_lbvh_host()
        h.prim_aabb = out["prim_aabb"].ctypes.data_as(C.POINTER(C.c_float))
        h.prim_node = out["prim_node"].ctypes.data_as(C.POINTER(C.c_int32))
        h.morton = out["morton"].ctypes.data_as(C.POINTER(C.c_uint64))
        h.sorted_prim = out["sorted_prim"].ctypes.data_as(C.POINTER(C.c_int32))
        h.left = out["left"].ctypes.data_as(C.POINTER(C.c_int32))
        h.right = out["right"].ctypes.data_as(C.POINTER(C.c_int32))
        h.parent = out["parent"].ctypes.data_as(C.POINTER(C.c_int32))
        h.node_aabb = out["node_aabb"].ctypes.data_as(C.POINTER(C.c_float))
        self._check(self.lib.rl_lbvh_download(self.h, C.byref(h)))
        out = {k: (v[:n] if k in ("prim_aabb", "prim_node", "morton", "sorted_prim") else
                   v[:m] if k in ("left", "right", "node_aabb") else v[:n + m])
               for k, v in out.items()}
        out["scene_lo"] = np.array(list(h.scene_lo), np.float32)
        out["scene_hi"] = np.array(list(h.scene_hi), np.float32)
        out["n_prims"], out["n_nodes"] = n, m
        return out

    # ---- OBJ ingest on the device ---------------------------------------------------
    def obj_parse(self, text, flavor: int) -> A.rl_obj_info:
        """rl_obj_parse: parse Wavefront OBJ text on the GPU; the triangles stay in HBM and RL_*_MESH nodes instance them"""
        if isinstance(text, str):
            text = text.encode()
        info = A.rl_obj_info()
        self._check(self.lib.rl_obj_parse(self.h, text, C.c_uint64(len(text)), int(flavor), C.byref(info)))
        self._scene = None  # a resident scene may reference the previous mesh
        return info

    def obj_download(self, n_triangles: int) -> dict:
        p = np.zeros((n_triangles, 3, 3), np.float64)
        nrm = np.zeros((n_triangles, 3, 3), np.float64)
        uv = np.zeros((n_triangles, 3, 2), np.float64)
        fl = np.zeros(n_triangles, np.uint8)
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        self._check(self.lib.rl_obj_download(self.h, dp(p), dp(nrm), dp(uv), fl.ctypes.data_as(C.POINTER(C.c_uint8))))
        return {"tri_p": p, "tri_n": nrm, "tri_uv": uv, "flags": fl}

    # ---- ray batches -------------------------------------------------------------------------
    def trace_batch(self, origins, directions, times=None, self_nodes=None):
        """closest hits of a ray batch; OW rays run through the render kernel's own traversal.  self_nodes (OW): the
        leaf node each ray starts on (-1 none), for scattered rays (rl_trace_batch_ex)."""
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, np.float32).reshape(-1, 3)
        n = o.shape[0]
        rays = np.zeros((n, 8), np.float32)
        rays[:, 0:3] = o
        rays[:, 3:6] = d
        if times is not None:
            rays[:, 6] = np.asarray(times, np.float32)
        hits = np.zeros(n, dtype=np.dtype([("node", np.int32), ("t", np.float32),
                                           ("u", np.float32), ("v", np.float32)]))
        if self_nodes is not None:
            sn = np.ascontiguousarray(self_nodes, np.int32).reshape(-1)
            assert sn.shape[0] == n
            self._check(self.lib.rl_trace_batch_ex(self.h, rays.ctypes.data_as(C.POINTER(A.rl_ray)),
                                                   sn.ctypes.data_as(C.POINTER(C.c_int32)), C.c_uint64(n),
                                                   hits.ctypes.data_as(C.POINTER(A.rl_hit))))
            return hits
        self._check(self.lib.rl_trace_batch(self.h, rays.ctypes.data_as(C.POINTER(A.rl_ray)),
                                            C.c_uint64(n),
                                            hits.ctypes.data_as(C.POINTER(A.rl_hit))))
        return hits

    # ---- renders (host buffers) ----------------------------------------------------------------
    def render_rtc(self, cam: A.rl_rtc_camera, aa_samples: int = 1, out: np.ndarray | None = None):
        if out is None:
            out = self._frame((cam.vsize, cam.hsize, 3), np.float32)
        stats = A.rl_stats()
        self._check(self.lib.rl_render_rtc(self.h, C.byref(cam), C.c_uint32(aa_samples),
                                           out.ctypes.data_as(C.POINTER(C.c_float)),
                                           C.byref(stats)))
        return out, stats

    def render_rtc_u8(self, cam: A.rl_rtc_camera, aa_samples: int = 1):
        """render + Canvas::ppm's 8-bit `translate` on the device: [H][W][3] uint8"""
        out = self._frame((cam.vsize, cam.hsize, 3), np.uint8)
        stats = A.rl_stats()
        self._check(self.lib.rl_render_rtc_u8(self.h, C.byref(cam), C.c_uint32(aa_samples),
                                              out.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(stats)))
        return out, stats

    def render_ow_u8(self, cam: A.rl_ow_camera, first_sample: int = 0):
        """render + pixel_data / linear_to_srgb / to_u8 on the device: [H][W][3] uint8"""
        out = self._frame((self.ow_image_height(cam), cam.image_width, 3), np.uint8)
        stats = A.rl_stats()
        self._check(self.lib.rl_render_ow_u8(self.h, C.byref(cam), C.c_uint32(first_sample),
                                             out.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(stats)))
        return out, stats

    def ow_image_height(self, cam: A.rl_ow_camera) -> int:
        return int(self.lib.rl_ow_image_height(C.byref(cam)))

    def ow_num_chunks(self, cam: A.rl_ow_camera) -> int:
        return int(self.lib.rl_ow_num_chunks(C.byref(cam)))

    def render_ow(self, cam: A.rl_ow_camera, first_sample: int = 0, out: np.ndarray | None = None):
        h = self.ow_image_height(cam)
        if out is None:
            out = self._frame((h, cam.image_width, 3), np.float32)
        stats = A.rl_stats()
        self._check(self.lib.rl_render_ow(self.h, C.byref(cam), C.c_uint32(first_sample),
                                          out.ctypes.data_as(C.POINTER(C.c_float)),
                                          C.byref(stats)))
        return out, stats

    # ---- renders (device buffers; multi-GPU plumbing) -----------------------------------------------
    @staticmethod
    def _jobs(jobs):
        arr = (A.rl_job * len(jobs))()
        for i, j in enumerate(jobs):
            arr[i] = A.rl_job(*j)
        return arr

    def render_rtc_device(self, cam, aa_samples, jobs, d_out_ptr: int, stream: int = 0, sync: bool = True):
        """sync=False launches and returns (no stats); call synchronize() to wait and surface errors."""
        stats = A.rl_stats() if sync else None
        arr = self._jobs(jobs)
        self._check(self.lib.rl_render_rtc_device(self.h, C.byref(cam), C.c_uint32(aa_samples), arr,
                                                  len(jobs), C.c_void_p(d_out_ptr), C.c_void_p(stream),
                                                  C.byref(stats) if sync else None))
        return stats

    def render_ow_device(self, cam, first_sample, jobs, d_partial_ptr: int, stream: int = 0, sync: bool = True):
        stats = A.rl_stats() if sync else None
        arr = self._jobs(jobs)
        self._check(self.lib.rl_render_ow_device(self.h, C.byref(cam), C.c_uint32(first_sample), arr,
                                                 len(jobs), C.c_void_p(d_partial_ptr), C.c_void_p(stream),
                                                 C.byref(stats) if sync else None))
        return stats

    # ---- cross-GPU queue (CUDA IPC + NVLink atomics) -----------------------------------------------
    def queue_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._check(self.lib.rl_queue_export(self.h, buf))
        return buf.raw

    def queue_import(self, handle: bytes):
        buf = C.create_string_buffer(handle, 64)
        self._check(self.lib.rl_queue_import(self.h, buf))

    def partial_export(self, bytes_per_slot: int, n_slots: int = A.RL_QUEUE_SLOTS) -> bytes:
        buf = C.create_string_buffer(64)
        self._check(self.lib.rl_partial_export(self.h, C.c_uint64(bytes_per_slot), n_slots, buf))
        return buf.raw

    def partial_import(self, handle: bytes, bytes_per_slot: int, n_slots: int = A.RL_QUEUE_SLOTS):
        buf = C.create_string_buffer(handle, 64)
        self._check(self.lib.rl_partial_import(self.h, buf, C.c_uint64(bytes_per_slot), n_slots))

    def queue_reset(self, stream: int = 0, slot: int = 0):
        self._check(self.lib.rl_queue_reset(self.h, C.c_void_p(stream), slot))

    def queue_completed(self, stream: int = 0, slot: int = 0) -> int:
        v = C.c_uint64()
        self._check(self.lib.rl_queue_completed(self.h, C.c_void_p(stream), slot, C.byref(v)))
        return int(v.value)

    def ow_job_items(self, cam, jobs) -> int:
        arr = self._jobs(jobs)
        return int(self.lib.rl_ow_job_items(C.byref(cam), arr, len(jobs)))

    def render_ow_shared(self, cam, first_sample, jobs, d_partial_ptr: int, stream: int = 0, slot: int = 0):
        arr = self._jobs(jobs)
        self._check(self.lib.rl_render_ow_shared(self.h, C.byref(cam), C.c_uint32(first_sample), arr, len(jobs),
                                                 C.c_void_p(d_partial_ptr), C.c_void_p(stream), slot))

    def ow_reduce_device(self, cam, d_partial_ptr: int, d_out_ptr: int, stream: int = 0):
        self._check(self.lib.rl_ow_reduce_device(self.h, C.byref(cam), C.c_void_p(d_partial_ptr),
                                                 C.c_void_p(d_out_ptr), C.c_void_p(stream)))

    def ow_reduce_shared(self, cam, slot: int, d_out_ptr: int, stream: int = 0):
        self._check(self.lib.rl_ow_reduce_shared(self.h, C.byref(cam), slot, C.c_void_p(d_out_ptr), C.c_void_p(stream)))


_default_ctx: Context | None = None


def default_context() -> Context:
    """The context `Camera.render` uses (cuda:LOCAL_RANK or cuda:0)."""
    global _default_ctx
    if _default_ctx is None:
        import os
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx
