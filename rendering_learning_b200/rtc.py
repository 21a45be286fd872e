"""Host-side mirror of the `ray-tracer-challenge` (RTC) scene API.

Same type names, field names and error behaviour as the reference crate, so a scene written against
the reference reads the same here; `Camera.render` lowers the object tree into an `rl_scene_desc`
and renders it on the B200 through the C ABI (there is no CPU path in this module).

Reference files mirrored (all under ray-tracer-challenge/src/):
  math/matrix.rs, math/point.rs, math/vector.rs          -> Matrix helpers, Point3d, Vec3d
  scene/transformation.rs:9-86                            -> translation .. view_transform, sequence
  scene/object/*.rs                                        -> Sphere .. Csg
  scene/material.rs:8-52, scene/pattern/*.rs, scene/light.rs
  scene/world.rs:26-31, scene/camera.rs:11-124, scene/mod.rs:18-27
  draw/canvas.rs:3-97, draw/color.rs
  io/wavefront_obj.rs:21-187
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Iterable, Sequence

import numpy as np

from . import _abi as A
from .desc import SceneDesc

Matrix4 = list  # 4x4 list of lists of float, row-major

# --------------------------------------------------------------------------------------------
# math (f64, same operation order as the reference so matrices are bit-identical)
# --------------------------------------------------------------------------------------------


def Point3d(x, y, z):
    return (float(x), float(y), float(z))


def Vec3d(x, y, z):
    return (float(x), float(y), float(z))


def identity() -> Matrix4:
    return [[1.0 if i == j else 0.0 for j in range(4)] for i in range(4)]


def matmul(a: Matrix4, b: Matrix4) -> Matrix4:
    """math/matrix.rs:196-213 — `sum += a[n][i] * b[i][m]`, i ascending, sum starts at 0.0."""
    out = [[0.0] * 4 for _ in range(4)]
    for n in range(4):
        for m in range(4):
            s = 0.0
            for i in range(4):
                s += a[n][i] * b[i][m]
            out[n][m] = s
    return out


def _det(d):
    if len(d) == 2:
        return d[0][0] * d[1][1] - d[0][1] * d[1][0]
    s = 0.0
    for i in range(len(d)):
        s += d[0][i] * _cofactor(d, 0, i)
    return s


def _minor(d, n, m):
    sub = [[v for j, v in enumerate(row) if j != m] for i, row in enumerate(d) if i != n]
    return _det(sub)


def _cofactor(d, n, m):
    mi = _minor(d, n, m)
    return mi if (n + m) % 2 == 0 else -mi


def invert(mat: Matrix4):
    """math/matrix.rs:68-86 — cofactor inverse; None when det == 0."""
    det = _det(mat)
    if det == 0.0:
        return None
    n = len(mat)
    out = [[0.0] * n for _ in range(n)]
    for i in range(n):
        for j in range(n):
            out[j][i] = _cofactor(mat, i, j) / det
    return out


class InvertibleMatrix:
    """math/matrix.rs:226-284.  `InvertibleMatrix.try_from(m)` raises like the reference's Err."""

    def __init__(self, matrix: Matrix4):
        inv = invert(matrix)
        if inv is None:
            raise ValueError("Matrix is not invertible.")
        self.matrix = [list(map(float, r)) for r in matrix]
        self._inverse = inv

    @classmethod
    def try_from(cls, matrix: Matrix4) -> "InvertibleMatrix":
        return cls(matrix)

    @classmethod
    def identity(cls) -> "InvertibleMatrix":
        return cls(identity())

    def inverse(self) -> Matrix4:
        return self._inverse

    def __mul__(self, rhs: "InvertibleMatrix") -> "InvertibleMatrix":
        return InvertibleMatrix(matmul(self.matrix, rhs.matrix))

    def flat(self):
        return [v for r in self.matrix for v in r]


def _as_invertible(m) -> InvertibleMatrix:
    return m if isinstance(m, InvertibleMatrix) else InvertibleMatrix(m)


class transformation:
    """scene/transformation.rs:9-86 (module mirrored as a namespace class)."""

    @staticmethod
    def translation(x, y, z) -> Matrix4:
        return [[1.0, 0.0, 0.0, float(x)], [0.0, 1.0, 0.0, float(y)], [0.0, 0.0, 1.0, float(z)],
                [0.0, 0.0, 0.0, 1.0]]

    @staticmethod
    def scaling(x, y, z) -> Matrix4:
        return [[float(x), 0.0, 0.0, 0.0], [0.0, float(y), 0.0, 0.0], [0.0, 0.0, float(z), 0.0],
                [0.0, 0.0, 0.0, 1.0]]

    @staticmethod
    def rotation_x(r) -> Matrix4:
        s, c = math.sin(r), math.cos(r)
        return [[1.0, 0.0, 0.0, 0.0], [0.0, c, -s, 0.0], [0.0, s, c, 0.0], [0.0, 0.0, 0.0, 1.0]]

    @staticmethod
    def rotation_y(r) -> Matrix4:
        s, c = math.sin(r), math.cos(r)
        return [[c, 0.0, s, 0.0], [0.0, 1.0, 0.0, 0.0], [-s, 0.0, c, 0.0], [0.0, 0.0, 0.0, 1.0]]

    @staticmethod
    def rotation_z(r) -> Matrix4:
        s, c = math.sin(r), math.cos(r)
        return [[c, -s, 0.0, 0.0], [s, c, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0]]

    @staticmethod
    def shearing(x_y, x_z, y_x, y_z, z_x, z_y) -> Matrix4:
        return [[1.0, x_y, x_z, 0.0], [y_x, 1.0, y_z, 0.0], [z_x, z_y, 1.0, 0.0],
                [0.0, 0.0, 0.0, 1.0]]

    @staticmethod
    def sequence(ts: Sequence[Matrix4]) -> Matrix4:
        acc = identity()
        for t in ts:
            acc = matmul(t, acc)
        return acc

    @staticmethod
    def view_transform(frm, to, up) -> Matrix4:
        def sub(a, b):
            return (a[0] - b[0], a[1] - b[1], a[2] - b[2])

        def norm(v):
            m = math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
            if m == 0.0:
                raise ValueError("called `Option::unwrap()` on a `None` value")
            return (v[0] / m, v[1] / m, v[2] / m)

        def cross(a, b):
            return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])

        forward = norm(sub(to, frm))
        upn = norm(up)
        left = cross(forward, upn)
        true_up = cross(left, forward)
        orientation = [[left[0], left[1], left[2], 0.0],
                       [true_up[0], true_up[1], true_up[2], 0.0],
                       [-forward[0], -forward[1], -forward[2], 0.0],
                       [0.0, 0.0, 0.0, 1.0]]
        return matmul(orientation, transformation.translation(-frm[0], -frm[1], -frm[2]))


# --------------------------------------------------------------------------------------------
# colour, patterns, material, light
# --------------------------------------------------------------------------------------------


def Color(r, g, b):
    return (float(r), float(g), float(b))


class color:
    """draw/color.rs:64-86 named colours."""
    black = staticmethod(lambda: Color(0.0, 0.0, 0.0))
    white = staticmethod(lambda: Color(1.0, 1.0, 1.0))
    red = staticmethod(lambda: Color(1.0, 0.0, 0.0))
    green = staticmethod(lambda: Color(0.0, 1.0, 0.0))
    blue = staticmethod(lambda: Color(0.0, 0.0, 1.0))


@dataclass
class _Pattern:
    a: tuple = (1.0, 1.0, 1.0)
    b: tuple = (0.0, 0.0, 0.0)
    transform: InvertibleMatrix = field(default_factory=InvertibleMatrix.identity)
    _kind = 0

    def _lower(self, sd: SceneDesc) -> int:
        def make():
            t = A.rl_texture()
            t.kind = self._kind
            t.tex_a = t.tex_b = t.image = -1
            t.a = (A.C.c_double * 3)(*self.a)
            t.b = (A.C.c_double * 3)(*self.b)
            t.scale = 1.0
            t.transform = (A.C.c_double * 16)(*_as_invertible(self.transform).flat())
            return t
        return sd.texture_id(self, make)


class Stripe(_Pattern):
    _kind = A.RL_TEX_RTC_STRIPE


class Checker3d(_Pattern):
    _kind = A.RL_TEX_RTC_CHECKER3D


class Gradient(_Pattern):
    _kind = A.RL_TEX_RTC_GRADIENT


class Ring(_Pattern):
    _kind = A.RL_TEX_RTC_RING


class Surface:
    """scene/material.rs:8-20 — `Surface::Color(c)` / `Surface::Pattern(p)`."""

    @staticmethod
    def Color(c):
        return tuple(map(float, c))

    @staticmethod
    def Pattern(p: _Pattern):
        return p


@dataclass(eq=False)
class Material:
    """scene/material.rs:22-52 (defaults identical)."""
    surface: object = (1.0, 1.0, 1.0)
    ambient: float = 0.1
    diffuse: float = 0.9
    specular: float = 0.9
    shininess: float = 200.0
    reflectivity: float = 0.0
    transparency: float = 0.0
    refractive_index: float = 1.0

    def _lower(self, sd: SceneDesc) -> int:
        def make():
            m = A.rl_material()
            m.kind = A.RL_MAT_RTC_PHONG
            if isinstance(self.surface, _Pattern):
                m.texture = self.surface._lower(sd)
                m.color = (A.C.c_double * 3)(0.0, 0.0, 0.0)
            else:
                m.texture = -1
                m.color = (A.C.c_double * 3)(*map(float, self.surface))
            m.ambient, m.diffuse, m.specular = self.ambient, self.diffuse, self.specular
            m.shininess, m.reflectivity = self.shininess, self.reflectivity
            m.transparency, m.refractive_index = self.transparency, self.refractive_index
            m.fuzz = 0.0
            return m
        return sd.material_id(self, make)


@dataclass
class PointLight:
    position: tuple
    intensity: tuple


# --------------------------------------------------------------------------------------------
# objects
# --------------------------------------------------------------------------------------------


class Object:
    def _lower(self, sd: SceneDesc) -> int:  # pragma: no cover - interface
        raise NotImplementedError


class _Leaf(Object):
    _kind = 0

    def __init__(self, material: Material | None = None):
        self.material = material if material is not None else Material()

    def _lower(self, sd):
        return sd.add_node(self._kind, material=self.material._lower(sd))


class Sphere(_Leaf):
    _kind = A.RL_RTC_SPHERE

    @classmethod
    def unit(cls):
        return cls()


class Plane(_Leaf):
    _kind = A.RL_RTC_PLANE


class Cube(_Leaf):
    _kind = A.RL_RTC_CUBE


class Cylinder(_Leaf):
    _kind = A.RL_RTC_CYLINDER

    def __init__(self, material=None, minimum=None, maximum=None, closed=False):
        super().__init__(material)
        self.minimum, self.maximum, self.closed = minimum, maximum, closed

    def _lower(self, sd):
        p = sd.add_params([-math.inf if self.minimum is None else self.minimum,
                           math.inf if self.maximum is None else self.maximum])
        return sd.add_node(self._kind, material=self.material._lower(sd),
                           flags=1 if self.closed else 0, param=p)


class Cone(Cylinder):
    _kind = A.RL_RTC_CONE


class Triangle(Object):
    """scene/object/triangle.rs:29-55."""

    def __init__(self, points, normals, material):
        self.points, self.normals = points, normals
        self.material = material if material is not None else Material()

    @classmethod
    def flat(cls, points, material=None):
        p1, p2, p3 = points
        e1 = (p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2])
        e2 = (p3[0] - p1[0], p3[1] - p1[1], p3[2] - p1[2])
        n = (e2[1] * e1[2] - e2[2] * e1[1], e2[2] * e1[0] - e2[0] * e1[2],
             e2[0] * e1[1] - e2[1] * e1[0])
        if math.sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]) == 0.0:
            raise ValueError("cannot be normalized.")  # NormalizedVec3d::try_from(...).unwrap()
        return cls(list(points), None, material)

    @classmethod
    def smooth(cls, vertices, material=None):
        return cls([v[0] for v in vertices], [v[1] for v in vertices], material)

    def _lower(self, sd):
        vals = [c for p in self.points for c in p]
        if self.normals is not None:
            vals += [c for n in self.normals for c in n]
        else:
            vals += [0.0] * 9
        p = sd.add_params(vals)
        return sd.add_node(A.RL_RTC_TRIANGLE, material=self.material._lower(sd),
                           flags=1 if self.normals is not None else 0, param=p)


class Transformed(Object):
    def __init__(self, child: Object, transform):
        self.child, self.transform = child, _as_invertible(transform)

    @classmethod
    def new(cls, child, transform):
        return cls(child, transform)

    def _lower(self, sd):
        me = sd.add_node(A.RL_RTC_TRANSFORMED, param=sd.add_params(self.transform.flat()))
        c = self.child._lower(sd)
        sd.set_node_children(me, c, c + 1)
        return me


class Group(Object):
    def __init__(self, children: Iterable[Object]):
        self.children = list(children)

    @classmethod
    def new(cls, children):
        return cls(children)

    def _lower(self, sd):
        me = sd.add_node(A.RL_RTC_GROUP)
        ids = [c._lower(sd) for c in self.children]
        b, e = sd.add_children(ids)
        sd.set_node_children(me, b, e)
        return me


class Bounded(Object):
    def __init__(self, child: Object):
        self.child = child

    @classmethod
    def new(cls, child):
        return cls(child)

    def _lower(self, sd):
        me = sd.add_node(A.RL_RTC_BOUNDED)
        c = self.child._lower(sd)
        sd.set_node_children(me, c, c + 1)
        return me


class DeviceMesh(Object):
    """`WavefrontObj::parse(reader).to_object()` (io/wavefront_obj.rs:22-76) with the OBJ text parsed ON THE GPU
    (Context.obj_parse): a Bounded<Group<Triangle>> whose triangles never visit the host.  Every triangle
    gets Material::default() (wavefront_obj.rs:164-166) unless another material is given."""

    def __init__(self, info, material=None):
        self.info = info
        self.material = material if material is not None else Material()

    @classmethod
    def parse(cls, text, ctx=None, material=None) -> "DeviceMesh":
        from .context import default_context
        ctx = ctx or default_context()
        return cls(ctx.obj_parse(text, A.RL_FLAVOR_RTC), material)

    def _lower(self, sd):
        return sd.add_node(A.RL_RTC_MESH, material=self.material._lower(sd))


class CsgOperation:
    Union = A.RL_CSG_UNION
    Intersection = A.RL_CSG_INTERSECTION
    Difference = A.RL_CSG_DIFFERENCE


class Csg(Object):
    def __init__(self, left: Object, right: Object, operation: int):
        self.left, self.right, self.operation = left, right, operation

    def _lower(self, sd):
        me = sd.add_node(A.RL_RTC_CSG, flags=int(self.operation))
        l = self.left._lower(sd)
        r = self.right._lower(sd)
        sd.set_node_children(me, l, r)
        return me


# --------------------------------------------------------------------------------------------
# world, camera, canvas
# --------------------------------------------------------------------------------------------


@dataclass
class World:
    """scene/world.rs:26-31, Default at 162-171."""
    objects: list = field(default_factory=list)
    lights: list = field(default_factory=list)
    max_reflection_depth: int = 5
    void_color: tuple = (0.0, 0.0, 0.0)

    def lower(self) -> SceneDesc:
        sd = SceneDesc(A.RL_FLAVOR_RTC)
        for o in self.objects:
            sd.roots.append(o._lower(sd))
        for l in self.lights:
            rl = A.rl_light()
            rl.position = (A.C.c_double * 3)(*map(float, l.position))
            rl.intensity = (A.C.c_double * 3)(*map(float, l.intensity))
            sd.lights.append(rl)
        sd.max_reflection_depth = int(self.max_reflection_depth)
        sd.void_color = tuple(map(float, self.void_color))
        return sd


@dataclass
class RenderOpts:
    anti_aliasing_samples: int = 1


class Canvas:
    """draw/canvas.rs:3-97.  `data` is [height][width][3] f64, row-major (idx = width*y + x)."""

    def __init__(self, width: int, height: int, data: np.ndarray | None = None):
        self.width, self.height = int(width), int(height)
        if data is None:
            self._data = np.zeros((self.height, self.width, 3), np.float64)
        elif isinstance(data, np.ndarray) and data.dtype == np.float32:
            # a frame straight from the device: held as the f32 it was computed in and widened to the reference's f64
            # `Color` on first pixel access (`data`) — widening a 4K frame costs more than rendering it
            self._data = data.reshape(self.height, self.width, 3)
        else:
            self._data = np.asarray(data, np.float64).reshape(self.height, self.width, 3)

    @property
    def data(self) -> np.ndarray:
        if self._data.dtype != np.float64:
            self._data = self._data.astype(np.float64)
        return self._data

    @data.setter
    def data(self, value):
        self._data = np.asarray(value, np.float64).reshape(self.height, self.width, 3)

    def at(self, x, y):
        if x >= self.width or y >= self.height:
            return None
        return tuple(self.data[y, x])

    def write(self, coords, c):
        x, y = coords
        if x >= self.width or y >= self.height:
            return None
        self.data[y, x] = c
        return ()

    def to_u8(self) -> np.ndarray:
        """canvas.rs:53-56 — `(c*255).round() as i32` (half away from zero) clamped to [0,255]."""
        v = self.data * 255.0
        av = np.abs(v)
        fl = np.floor(av)
        r = np.copysign(fl + ((av - fl) >= 0.5), v)
        r = np.where(np.isnan(v), 0.0, r)  # `NaN as i32` == 0
        return np.clip(r, 0, 255).astype(np.int64)

    def ppm(self) -> str:
        """canvas.rs:50-97 — P3, rows wrapped so no line exceeds 70 characters."""
        return ppm_from_u8(self.to_u8())


def ppm_from_u8(u8) -> str:
    """Canvas::ppm's text (canvas.rs:58-97) from [H][W][3] 8-bit channels: header, one image row per wrapped group of
    lines, greedy wrap so that no line exceeds 70 characters, trailing newline.  Vectorised over the rows (the greedy
    wrap is sequential along a row only), so a 4K frame formats in well under a second."""
    u8 = np.asarray(u8)
    height, width = u8.shape[0], u8.shape[1]
    v = u8.reshape(height, width * 3).astype(np.int64)
    n_tok = v.shape[1]
    length = 1 + (v >= 10) + (v >= 100)                      # digits per token
    len_t = np.ascontiguousarray(length.T)                   # [token][row]: the loop below walks contiguous rows
    sep_t = np.full((n_tok, height), ord(" "), np.uint8)     # separator BEFORE each token
    cur = len_t[0].copy()
    for k in range(1, n_tok):
        nxt = cur + 1 + len_t[k]
        fits = nxt <= 70
        sep_t[k][~fits] = ord("\n")
        cur = np.where(fits, nxt, len_t[k])
    sep_t[0] = ord("\n")                                     # rows are joined by newlines; the very first one is dropped below
    sep = sep_t.T
    size = (length + 1).ravel()
    end = np.cumsum(size)
    start = end - size
    buf = np.empty(int(end[-1]) if end.size else 0, np.uint8)
    fv, fl = v.ravel(), length.ravel()
    buf[start] = sep.ravel()
    ones, tens, hund = fv % 10, (fv // 10) % 10, fv // 100
    buf[end - 1] = ones + 48
    m2 = fl >= 2
    buf[end[m2] - 2] = tens[m2] + 48
    m3 = fl == 3
    buf[end[m3] - 3] = hund[m3] + 48
    return f"P3\n{width} {height}\n255\n" + buf[1:].tobytes().decode("ascii") + "\n"


class Camera:
    """scene/camera.rs:11-124."""

    def __init__(self, hsize: int, vsize: int, fov: float, transform=None):
        self.hsize, self.vsize, self.fov = int(hsize), int(vsize), float(fov)
        self.transform = _as_invertible(transform if transform is not None else identity())
        half_view = math.tan(self.fov / 2.0)
        aspect = self.hsize / self.vsize
        if aspect >= 1.0:
            self.half_width, self.half_height = half_view, half_view / aspect
        else:
            self.half_width, self.half_height = half_view * aspect, half_view
        self.pixel_size = self.half_width * 2.0 / self.hsize

    @classmethod
    def new(cls, hsize, vsize, fov, transform):
        return cls(hsize, vsize, fov, transform)

    @classmethod
    def default(cls, hsize, vsize, fov):
        return cls(hsize, vsize, fov, InvertibleMatrix.identity())

    def abi(self) -> A.rl_rtc_camera:
        c = A.rl_rtc_camera()
        c.hsize, c.vsize, c.fov = self.hsize, self.vsize, self.fov
        c.transform = (A.C.c_double * 16)(*self.transform.flat())
        return c

    def _resident(self, world, opts, ctx):
        from .context import default_context
        opts = opts or RenderOpts()
        if opts.anti_aliasing_samples < 1:
            # `.reduce(..).unwrap()` on zero rays panics in the reference
            raise ValueError("called `Option::unwrap()` on a `None` value")
        ctx = ctx or default_context()
        ctx.scene_upload(world if isinstance(world, SceneDesc) else world.lower())
        return ctx, opts

    def render(self, world, opts: RenderOpts | None = None, ctx=None) -> Canvas:
        """Drop-in for `Camera::render` (camera.rs:93-124): same inputs, same Canvas.  `world` is a World, or the
        SceneDesc a caller lowered once (`world.lower()`) and keeps across renders."""
        ctx, opts = self._resident(world, opts, ctx)
        rgb, _ = ctx.render_rtc(self.abi(), opts.anti_aliasing_samples)
        return Canvas(self.hsize, self.vsize, rgb)

    def render_ppm(self, world, opts: RenderOpts | None = None, ctx=None) -> str:
        """`camera.render(&world, &opts).ppm()` (draw_scene.rs:42-47) for a caller that only wants the file: the 8-bit
        `translate` of canvas.rs:53-56 runs on the device (rl_render_rtc_u8, bit-identical to Canvas.ppm of the f32
        frame) and 3 bytes per pixel cross PCIe instead of 12 + an f64 Canvas on the host."""
        ctx, opts = self._resident(world, opts, ctx)
        u8, _ = ctx.render_rtc_u8(self.abi(), opts.anti_aliasing_samples)
        return ppm_from_u8(u8)


@dataclass
class Scene:
    """scene/mod.rs:18-27."""
    camera: Camera
    world: World

    def render(self, opts: RenderOpts | None = None, ctx=None) -> Canvas:
        return self.camera.render(self.world, opts, ctx=ctx)

    def render_ppm(self, opts: RenderOpts | None = None, ctx=None) -> str:
        return self.camera.render_ppm(self.world, opts, ctx=ctx)


# --------------------------------------------------------------------------------------------
# OBJ ingest (io/wavefront_obj.rs) — unchanged input stage; produces the same triangle list
# --------------------------------------------------------------------------------------------


def _parse_f64(s: str):
    try:
        return float(s)
    except ValueError:
        return None


class WavefrontObj:
    def __init__(self):
        self.ignored = 0
        self.groups: dict[str | None, list[Triangle]] = {}
        self.vertices: list[tuple] = []
        self.normals: list[tuple] = []
        # every OBJ triangle gets `Material::default()` (wavefront_obj.rs:164-166); one shared
        # instance lowers to one material record instead of one per triangle
        self._default_material = Material()

    @classmethod
    def parse(cls, text) -> "WavefrontObj":
        if isinstance(text, bytes):
            text = text.decode()
        obj = cls()
        cur_name: str | None = None
        cur: list[Triangle] = []
        for line in text.splitlines():
            head, sep, tail = line.partition(" ")
            ok = False
            if sep:
                trimmed = tail.strip()
                if head == "v":
                    nums = [_parse_f64(s) for s in trimmed.split()]
                    if len(nums) == 3 and None not in nums:
                        obj.vertices.append(tuple(nums))
                        ok = True
                elif head == "vn":
                    nums = [_parse_f64(s) for s in trimmed.split()]
                    if len(nums) == 3 and None not in nums:
                        obj.normals.append(tuple(nums))
                        ok = True
                elif head == "f":
                    tris = obj._parse_face(trimmed)
                    if tris is not None:
                        cur.extend(tris)
                        ok = True
                elif head == "g":
                    obj.groups[cur_name] = cur
                    cur_name, cur = trimmed, []
                    ok = True
            if not ok:
                obj.ignored += 1
        obj.groups[cur_name] = cur
        return obj

    def _parse_face(self, tail: str):
        idx = []
        for tok in tail.split():
            parts = tok.split("/")
            if len(parts) in (1, 2):
                v, n = parts[0], None
            elif len(parts) == 3:
                v, n = parts[0], parts[2]
            else:
                return None
            if not v.isdigit():
                return None
            if n is not None:
                if not n.isdigit():
                    return None
                idx.append((int(v), int(n)))
            else:
                idx.append((int(v), None))
        if len(idx) < 3:
            return None
        verts = [(self.vertices[v - 1], None if n is None else self.normals[n - 1])
                 for v, n in idx]
        tris = []
        for i in range(2, len(verts)):
            a, b, c = verts[0], verts[i - 1], verts[i]
            if a[1] is not None and b[1] is not None and c[1] is not None:
                tris.append(Triangle.smooth([a, b, c], self._default_material))
            else:
                tris.append(Triangle.flat([a[0], b[0], c[0]], self._default_material))
        return tris

    def triangles(self) -> list[Triangle]:
        # the reference iterates a HashMap (arbitrary group order); file order is one valid order
        return [t for g in self.groups.values() for t in g]

    def to_object(self) -> Object:
        return Bounded(Group(self.triangles()))
