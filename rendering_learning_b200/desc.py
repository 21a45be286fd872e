"""Scene-description builder: the Python twin of the Rust `FlatSceneBuilder` the `lower()` trait
methods would write into.  It only *records* the reference's object tree in the POD
form of include/rl_b200.h; flattening / transform composition / inversion happen inside the library.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi as A


class SceneDesc:
    """Owns the numpy/ctypes buffers behind an `rl_scene_desc` so they outlive the call."""

    def __init__(self, flavor: int):
        self.flavor = flavor
        self.nodes: list[tuple[int, int, int, int, int, int]] = []
        self.children: list[int] = []
        self.params: list[float] = []
        self.roots: list[int] = []
        self.materials: list[A.rl_material] = []
        self.textures: list[A.rl_texture] = []
        self.images: list[np.ndarray] = []
        self.lights: list[A.rl_light] = []
        self.perlins: list[A.rl_perlin] = []
        self.max_reflection_depth = 5
        self.void_color = (0.0, 0.0, 0.0)
        self._mat_ids: dict[int, int] = {}
        self._tex_ids: dict[int, int] = {}
        self._keep: list[object] = []
        self._frozen = None

    # ---- recording -------------------------------------------------------------------------
    def add_params(self, values) -> int:
        off = len(self.params)
        self.params.extend(float(v) for v in values)
        return off

    def add_node(self, kind, material=-1, child_begin=-1, child_end=-1, flags=0, param=-1) -> int:
        self.nodes.append((kind, material, child_begin, child_end, flags, param))
        return len(self.nodes) - 1

    def set_node_children(self, node_id: int, child_begin: int, child_end: int):
        k, m, _, _, f, p = self.nodes[node_id]
        self.nodes[node_id] = (k, m, child_begin, child_end, f, p)

    def add_children(self, ids) -> tuple[int, int]:
        b = len(self.children)
        self.children.extend(int(i) for i in ids)
        return b, len(self.children)

    def material_id(self, key: object, make) -> int:
        """Materials are shared by identity, like `&Material` / `Box<dyn Material>` in the reference."""
        k = id(key)
        if k not in self._mat_ids:
            self._keep.append(key)
            self.materials.append(make())
            self._mat_ids[k] = len(self.materials) - 1
        return self._mat_ids[k]

    def texture_id(self, key: object, make) -> int:
        k = id(key)
        if k not in self._tex_ids:
            self._keep.append(key)
            t = make()  # may recursively register sub-textures first
            self.textures.append(t)
            self._tex_ids[k] = len(self.textures) - 1
        return self._tex_ids[k]

    def add_image(self, rgb: np.ndarray) -> int:
        arr = np.ascontiguousarray(rgb, dtype=np.float32)
        assert arr.ndim == 3 and arr.shape[2] == 3
        self.images.append(arr)
        return len(self.images) - 1

    def add_perlin(self, randvec, perm_x, perm_y, perm_z) -> int:
        p = A.rl_perlin()
        for i in range(256):
            for k in range(3):
                p.randvec[i][k] = float(randvec[i][k])
            p.perm_x[i], p.perm_y[i], p.perm_z[i] = int(perm_x[i]), int(perm_y[i]), int(perm_z[i])
        self.perlins.append(p)
        return len(self.perlins) - 1

    # ---- freezing --------------------------------------------------------------------------
    def freeze(self) -> A.rl_scene_desc:
        if self._frozen is not None:
            return self._frozen
        d = A.rl_scene_desc()
        d.abi_version = A.RL_B200_ABI_VERSION
        d.flavor = self.flavor
        n = len(self.nodes)
        self._nodes_arr = (A.rl_node * max(n, 1))()
        for i, t in enumerate(self.nodes):
            self._nodes_arr[i] = A.rl_node(*t)
        d.nodes = C.cast(self._nodes_arr, C.POINTER(A.rl_node))
        d.n_nodes = n
        self._children_np = np.asarray(self.children if self.children else [0], dtype=np.int32)
        d.children = self._children_np.ctypes.data_as(C.POINTER(C.c_int32))
        d.n_children = len(self.children)
        self._params_np = np.asarray(self.params if self.params else [0.0], dtype=np.float64)
        d.params = self._params_np.ctypes.data_as(C.POINTER(C.c_double))
        d.n_params = len(self.params)
        self._roots_np = np.asarray(self.roots if self.roots else [0], dtype=np.int32)
        d.roots = self._roots_np.ctypes.data_as(C.POINTER(C.c_int32))
        d.n_roots = len(self.roots)
        self._mats_arr = (A.rl_material * max(len(self.materials), 1))(*self.materials)
        d.materials = C.cast(self._mats_arr, C.POINTER(A.rl_material))
        d.n_materials = len(self.materials)
        self._tex_arr = (A.rl_texture * max(len(self.textures), 1))(*self.textures)
        d.textures = C.cast(self._tex_arr, C.POINTER(A.rl_texture))
        d.n_textures = len(self.textures)
        self._img_arr = (A.rl_image * max(len(self.images), 1))()
        for i, im in enumerate(self.images):
            self._img_arr[i].width = im.shape[1]
            self._img_arr[i].height = im.shape[0]
            self._img_arr[i].rgb = im.ctypes.data_as(C.POINTER(C.c_float))
        d.images = C.cast(self._img_arr, C.POINTER(A.rl_image))
        d.n_images = len(self.images)
        self._lights_arr = (A.rl_light * max(len(self.lights), 1))(*self.lights)
        d.lights = C.cast(self._lights_arr, C.POINTER(A.rl_light))
        d.n_lights = len(self.lights)
        self._perlin_arr = (A.rl_perlin * max(len(self.perlins), 1))(*self.perlins)
        d.perlins = C.cast(self._perlin_arr, C.POINTER(A.rl_perlin))
        d.n_perlins = len(self.perlins)
        d.max_reflection_depth = int(self.max_reflection_depth)
        d.void_color = (C.c_double * 3)(*self.void_color)
        self._frozen = d
        return d

    def check(self) -> A.rl_scene_info:
        """rl_scene_check: validate + flatten on the host (no GPU needed); raises RlError like scene_upload would"""
        info = A.rl_scene_info()
        err = C.create_string_buffer(512)
        rc = A.load_library().rl_scene_check(C.byref(self.freeze()), C.byref(info), err, 512)
        if rc != A.RL_OK:
            raise A.RlError(rc, err.value.decode(errors="replace"))
        return info

    def nbytes(self) -> int:
        """Host bytes that cross to the library on rl_scene_upload (for e2e h2d accounting)."""
        self.freeze()
        n = C.sizeof(A.rl_node) * len(self.nodes) + 4 * len(self.children) + 8 * len(self.params)
        n += C.sizeof(A.rl_material) * len(self.materials)
        n += C.sizeof(A.rl_texture) * len(self.textures)
        n += C.sizeof(A.rl_light) * len(self.lights)
        n += sum(im.nbytes for im in self.images)
        n += C.sizeof(A.rl_perlin) * len(self.perlins)
        return n
