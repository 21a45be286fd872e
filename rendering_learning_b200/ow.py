"""Host-side mirror of the `ray-tracing-one-weekend` (OW) scene API.

Same type and field names as the reference crate; `Camera.render` / `render_from_checkpoint` lower
the hittable tree to an `rl_scene_desc` and render on the B200 through the C ABI, returning the same
`Canvas` of per-pixel colour SUMS the reference returns (so `merge` / checkpoints keep working).

Reference files mirrored (all under ray-tracing-one-weekend/src/):
  camera.rs:24-59 CameraParams, 72-143 Camera, 263-296 Canvas
  hittable/{sphere,flat/quad,flat/triangle,transform,translate}.rs, hittable/mod.rs:40-85, bvh.rs
  material.rs:69-195, texture.rs:15-82, color.rs:22-57 + 114-137, output.rs:5-14
  io/wavefront_obj.rs:32-242
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import _abi as A
from .desc import SceneDesc


def Vec3(x, y, z):
    return (float(x), float(y), float(z))


Point3 = Vec3
Color = Vec3

# ---- textures (texture.rs) ---------------------------------------------------------------------


class Texture:
    def _lower(self, sd: SceneDesc) -> int:  # pragma: no cover
        raise NotImplementedError


def _new_tex(kind):
    t = A.rl_texture()
    t.kind = kind
    t.tex_a = t.tex_b = t.image = -1
    t.scale = 1.0
    return t


@dataclass(eq=False)
class SolidColor(Texture):
    albedo: tuple

    def _lower(self, sd):
        def make():
            t = _new_tex(A.RL_TEX_OW_SOLID)
            t.a = (A.C.c_double * 3)(*map(float, self.albedo))
            return t
        return sd.texture_id(self, make)


class Checker(Texture):
    def __init__(self, scale, even: Texture, odd: Texture):
        self.scale, self.even, self.odd = float(scale), even, odd
        self.inv_scale = 1.0 / self.scale

    @classmethod
    def new(cls, scale, even, odd):
        return cls(scale, even, odd)

    def _lower(self, sd):
        def make():
            t = _new_tex(A.RL_TEX_OW_CHECKER)
            t.tex_a = self.even._lower(sd)
            t.tex_b = self.odd._lower(sd)
            t.scale = self.scale
            return t
        return sd.texture_id(self, make)


class Image(Texture):
    """texture.rs:58-82 — `image` is linear-light float32 [H][W][3] (an `Rgb32FImage`)."""

    def __init__(self, image: np.ndarray):
        self.image = np.ascontiguousarray(image, np.float32)
        if self.image.ndim != 3 or self.image.shape[2] != 3:
            raise ValueError("Image texture wants [H][W][3] float data")

    def _lower(self, sd):
        def make():
            if self.image.shape[0] == 0 or self.image.shape[1] == 0:
                raise ValueError("Image has no data")  # texture.rs:64-67 assert
            t = _new_tex(A.RL_TEX_OW_IMAGE)
            t.image = sd.add_image(self.image)
            return t
        return sd.texture_id(self, make)


class Perlin:
    """perlin.rs:9-36 — the tables `Perlin::new(rng)` draws: 256 unit vectors, then three permutations.
    `rng` offers `unit_sphere()` (rand_distr::UnitSphere, vec3.rs:72-75) and `gen_range_usize(n)`
    (`rng.gen_range(0..n)`, perlin.rs:84); `scenes.Xoshiro256PlusPlus` is the generator the reference's examples use."""

    def __init__(self, randvec, perm_x, perm_y, perm_z):
        self.randvec, self.perm_x, self.perm_y, self.perm_z = randvec, perm_x, perm_y, perm_z

    @classmethod
    def new(cls, rng):
        randvec = [rng.unit_sphere() for _ in range(256)]

        def permutation():  # permute (perlin.rs:81-89): target drawn from 0..i, EXCLUSIVE of i
            p = list(range(256))
            for i in range(255, 0, -1):
                t = rng.gen_range_usize(i)
                p[i], p[t] = p[t], p[i]
            return p
        return cls(randvec, permutation(), permutation(), permutation())

    def noise(self, p) -> float:
        """perlin.rs:39-63 + perlin_interp 91-115 (host restatement, used by tests)"""
        import math
        fl = [math.floor(c) for c in p]
        u, v, w = (p[0] - fl[0], p[1] - fl[1], p[2] - fl[2])
        i, j, k = int(fl[0]), int(fl[1]), int(fl[2])
        uu, vv, ww = u * u * (3.0 - 2.0 * u), v * v * (3.0 - 2.0 * v), w * w * (3.0 - 2.0 * w)
        acc = 0.0
        for di in range(2):
            for dj in range(2):
                for dk in range(2):
                    c = self.randvec[self.perm_x[(i + di) & 255] ^ self.perm_y[(j + dj) & 255] ^ self.perm_z[(k + dk) & 255]]
                    wv = (u - di, v - dj, w - dk)
                    acc += ((di * uu + (1.0 - di) * (1.0 - uu)) * (dj * vv + (1.0 - dj) * (1.0 - vv)) *
                            (dk * ww + (1.0 - dk) * (1.0 - ww)) * (c[0] * wv[0] + c[1] * wv[1] + c[2] * wv[2]))
        return acc

    def turb(self, p, depth: int) -> float:
        """perlin.rs:65-78"""
        acc, weight, q = 0.0, 1.0, tuple(p)
        for _ in range(depth):
            acc += weight * self.noise(q)
            weight *= 0.5
            q = (q[0] * 2.0, q[1] * 2.0, q[2] * 2.0)
        return abs(acc)


@dataclass(eq=False)
class Noise(Texture):
    """texture.rs:84-94"""
    noise: Perlin
    scale: float

    def value(self, u, v, p):
        import math
        g = 0.5 * (1.0 + math.sin(self.scale * p[2] + 10.0 * self.noise.turb(p, 7)))
        return (g, g, g)

    def _lower(self, sd):
        def make():
            t = _new_tex(A.RL_TEX_OW_NOISE)
            t.image = sd.add_perlin(self.noise.randvec, self.noise.perm_x, self.noise.perm_y, self.noise.perm_z)
            t.scale = float(self.scale)
            return t
        return sd.texture_id(self, make)


# ---- materials (material.rs) --------------------------------------------------------------------


class Material:
    def _lower(self, sd: SceneDesc) -> int:  # pragma: no cover
        raise NotImplementedError


def _new_mat(kind):
    m = A.rl_material()
    m.kind = kind
    m.texture = -1
    return m


@dataclass(eq=False)
class Lambertian(Material):
    texture: Texture

    def _lower(self, sd):
        def make():
            m = _new_mat(A.RL_MAT_OW_LAMBERTIAN)
            m.texture = self.texture._lower(sd)
            return m
        return sd.material_id(self, make)


@dataclass(eq=False)
class Metal(Material):
    albedo: tuple
    fuzz: float

    def _lower(self, sd):
        def make():
            m = _new_mat(A.RL_MAT_OW_METAL)
            m.color = (A.C.c_double * 3)(*map(float, self.albedo))
            m.fuzz = float(self.fuzz)
            return m
        return sd.material_id(self, make)


@dataclass(eq=False)
class Dielectric(Material):
    refraction_index: float

    def _lower(self, sd):
        def make():
            m = _new_mat(A.RL_MAT_OW_DIELECTRIC)
            m.refractive_index = float(self.refraction_index)
            return m
        return sd.material_id(self, make)


@dataclass(eq=False)
class DiffuseLight(Material):
    texture: Texture

    def _lower(self, sd):
        def make():
            m = _new_mat(A.RL_MAT_OW_DIFFUSE_LIGHT)
            m.texture = self.texture._lower(sd)
            return m
        return sd.material_id(self, make)


@dataclass(eq=False)
class Isotropic(Material):
    """material.rs:197-221 — the phase function of a ConstantMedium"""
    texture: Texture

    def _lower(self, sd):
        def make():
            m = _new_mat(A.RL_MAT_OW_ISOTROPIC)
            m.texture = self.texture._lower(sd)
            return m
        return sd.material_id(self, make)


# ---- hittables -------------------------------------------------------------------------------------


class Hittable:
    """hittable/mod.rs:40-85 incl. the builder methods."""

    def _lower(self, sd: SceneDesc) -> int:  # pragma: no cover
        raise NotImplementedError

    def translate(self, offset):
        return Translate(self, offset)

    def rotate_x(self, degrees):
        return Transform.rotate_x(self, degrees)

    def rotate_y(self, degrees):
        return Transform.rotate_y(self, degrees)

    def rotate_z(self, degrees):
        return Transform.rotate_z(self, degrees)

    def scale(self, s):
        return Transform.scale(self, s)


class Center:
    @staticmethod
    def Stationary(p):
        return ("stationary", tuple(map(float, p)))

    @staticmethod
    def Moving(p1, p2):
        return ("moving", tuple(map(float, p1)), tuple(map(float, p2)))


class Sphere(Hittable):
    def __init__(self, center, radius, material: Material):
        self.center, self.radius, self.material = center, float(radius), material

    def _lower(self, sd):
        if self.center[0] == "moving":
            c1, c2, fl = self.center[1], self.center[2], 1
        else:
            c1, c2, fl = self.center[1], self.center[1], 0
        p = sd.add_params(list(c1) + list(c2) + [self.radius])
        return sd.add_node(A.RL_OW_SPHERE, material=self.material._lower(sd), flags=fl, param=p)


def _check_plane(u, v):
    """Plane::new (flat/plane.rs:23-28): `NormalizedVec3::try_from(&n).expect(..)`."""
    n = (u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0])
    m = n[0] * n[0] + n[1] * n[1] + n[2] * n[2]
    if m == 0.0 or abs(m) <= 1e-16:
        raise ValueError("Failed to find normal because u and v were parallel")


class Quad(Hittable):
    def __init__(self, q, u, v, material: Material):
        _check_plane(u, v)
        self.q, self.u, self.v, self.material = q, u, v, material

    @classmethod
    def new(cls, q, u, v, material):
        return cls(q, u, v, material)

    def _lower(self, sd):
        p = sd.add_params(list(self.q) + list(self.u) + list(self.v))
        return sd.add_node(A.RL_OW_QUAD, material=self.material._lower(sd), param=p)


class Triangle(Hittable):
    def __init__(self, points, texture_coords, normals, material: Material):
        p1, p2, p3 = points
        u = (p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2])
        v = (p3[0] - p1[0], p3[1] - p1[1], p3[2] - p1[2])
        _check_plane(u, v)
        self.points, self.texture_coords, self.normals = points, texture_coords, normals
        self.material = material

    @classmethod
    def new(cls, q, u, v, material):
        p2 = (q[0] + u[0], q[1] + u[1], q[2] + u[2])
        p3 = (q[0] + v[0], q[1] + v[1], q[2] + v[2])
        return cls([q, p2, p3], None, None, material)

    @classmethod
    def from_model(cls, points, texture_coords, normals, material):
        return cls(points, texture_coords, normals, material)

    def _lower(self, sd):
        vals = [c for p in self.points for c in p]
        fl = 0
        if self.texture_coords is not None:
            vals += [c for t in self.texture_coords for c in t]
            fl |= 1
        else:
            vals += [0.0] * 6
        if self.normals is not None:
            vals += [c for n in self.normals for c in n]
            fl |= 2
        else:
            vals += [0.0] * 9
        return sd.add_node(A.RL_OW_TRIANGLE, material=self.material._lower(sd), flags=fl,
                           param=sd.add_params(vals))


class _ctor_or_builder:
    """`Transform::rotate_y(object, deg)` on the class, `hittable.rotate_y(deg)` on an instance
    (the reference has both: transform.rs:22-86 and the builder methods of hittable/mod.rs:56-85)."""

    def __init__(self, fn):
        self.fn = fn

    def __get__(self, inst, owner):
        if inst is None:
            return lambda obj, arg: self.fn(owner, obj, arg)
        return lambda arg: self.fn(owner, inst, arg)


class Transform(Hittable):
    """hittable/transform.rs:21-86 — forward and inverse 3x3 are both written out by the ctor."""

    def __init__(self, obj: Hittable, m, minv):
        self.object, self.m, self.minv = obj, m, minv

    @_ctor_or_builder
    def rotate_x(cls, obj, degrees):
        r = math.radians(degrees)
        s, c = math.sin(r), math.cos(r)
        return cls(obj, [[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]],
                   [[1.0, 0.0, 0.0], [0.0, c, s], [0.0, -s, c]])

    @_ctor_or_builder
    def rotate_y(cls, obj, degrees):
        r = math.radians(degrees)
        s, c = math.sin(r), math.cos(r)
        return cls(obj, [[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]],
                   [[c, 0.0, -s], [0.0, 1.0, 0.0], [s, 0.0, c]])

    @_ctor_or_builder
    def rotate_z(cls, obj, degrees):
        r = math.radians(degrees)
        s, c = math.sin(r), math.cos(r)
        return cls(obj, [[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]],
                   [[c, s, 0.0], [-s, c, 0.0], [0.0, 0.0, 1.0]])

    @_ctor_or_builder
    def scale(cls, obj, scale):
        s = float(scale)
        i = 1.0 / s
        return cls(obj, [[s, 0.0, 0.0], [0.0, s, 0.0], [0.0, 0.0, s]],
                   [[i, 0.0, 0.0], [0.0, i, 0.0], [0.0, 0.0, i]])

    def _lower(self, sd):
        vals = [v for r in self.m for v in r] + [v for r in self.minv for v in r]
        me = sd.add_node(A.RL_OW_TRANSFORM, param=sd.add_params(vals))
        c = self.object._lower(sd)
        sd.set_node_children(me, c, c + 1)
        return me


class Translate(Hittable):
    def __init__(self, obj: Hittable, offset):
        self.object, self.offset = obj, tuple(map(float, offset))

    def _lower(self, sd):
        me = sd.add_node(A.RL_OW_TRANSLATE, param=sd.add_params(self.offset))
        c = self.object._lower(sd)
        sd.set_node_children(me, c, c + 1)
        return me


class ConstantMedium(Hittable):
    """hittable/constant_medium.rs:8-23.  The reference draws the scattering distance from the process-global
    RNG (its own TODO at 52-55: renders with a medium are not repeatable); the device path and the oracle draw it
    from the path's seeded stream instead, so renders are deterministic — parity is statistical either way."""

    def __init__(self, boundary: Hittable, density: float, material: Material):
        self.boundary, self.density, self.phase_function = boundary, float(density), material
        # f64 division as in Rust: a zero density gives -inf (the medium never scatters), not a panic
        self.neg_inv_density = -1.0 / self.density if self.density != 0.0 else float("-inf")

    @classmethod
    def new(cls, boundary, density, material):
        return cls(boundary, density, material)

    def _lower(self, sd):
        me = sd.add_node(A.RL_OW_CONSTANT_MEDIUM, material=self.phase_function._lower(sd),
                         param=sd.add_params([self.density]))
        c = self.boundary._lower(sd)
        sd.set_node_children(me, c, c + 1)
        return me


class Bvh(Hittable):
    def __init__(self, hs):
        hs = list(hs)
        if not hs:
            raise ValueError("Cannot make a BVH node without hittables.")  # bvh.rs:23-25 panic
        self.children = hs

    @classmethod
    def new(cls, hs):
        return cls(hs)

    def _lower(self, sd):
        me = sd.add_node(A.RL_OW_BVH)
        ids = [c._lower(sd) for c in self.children]
        b, e = sd.add_children(ids)
        sd.set_node_children(me, b, e)
        return me


class DeviceMesh(Hittable):
    """`WavefrontObj::parse(reader).to_object(material)` (io/wavefront_obj.rs:32-104) with the OBJ text parsed ON THE GPU
    (Context.obj_parse): the scene node only names the mesh the ctx holds, its triangles never visit the host."""

    def __init__(self, info, material: Material):
        self.info, self.material = info, material

    @classmethod
    def parse(cls, text, material: Material, ctx=None) -> "DeviceMesh":
        from .context import default_context
        ctx = ctx or default_context()
        return cls(ctx.obj_parse(text, A.RL_FLAVOR_OW), material)

    def _lower(self, sd):
        return sd.add_node(A.RL_OW_MESH, material=self.material._lower(sd))


class HittableList(Hittable):
    """A slice / array of hittables (`impl Hittable for [H]`, hittable/mod.rs:86-111)."""

    def __init__(self, hs):
        self.children = list(hs)

    def _lower(self, sd):
        me = sd.add_node(A.RL_OW_LIST)
        ids = [c._lower(sd) for c in self.children]
        b, e = sd.add_children(ids)
        sd.set_node_children(me, b, e)
        return me


def lower_world(world) -> SceneDesc:
    sd = SceneDesc(A.RL_FLAVOR_OW)
    if isinstance(world, (list, tuple)):
        world = HittableList(world)
    sd.roots.append(world._lower(sd))
    return sd


# ---- colour / output (color.rs, output.rs) ---------------------------------------------------------


class srgb:
    U, V, A_, C_, GAMMA = 0.04045, 0.0031308, 12.92, 0.055, 2.4

    @staticmethod
    def srgb_to_linear(u):
        u = np.asarray(u, np.float64)
        return np.where(u <= srgb.U, u / srgb.A_, np.power((u + srgb.C_) / (1.0 + srgb.C_), srgb.GAMMA))

    @staticmethod
    def linear_to_srgb(v):
        v = np.asarray(v, np.float64)
        with np.errstate(invalid="ignore"):
            return np.where(v <= srgb.V, srgb.A_ * v,
                            (1.0 + srgb.C_) * np.power(v, 1.0 / srgb.GAMMA) - srgb.C_)


def to_u8(rgb):
    """color.rs:47-57 — `(val*255.999).floor() as i16` clamped to [0,255]."""
    n = np.floor(np.asarray(rgb, np.float64) * 255.999)
    n = np.where(np.isnan(n), 0.0, n)
    n = np.clip(n, -32768, 32767)  # `as i16` saturates
    return np.clip(n, 0, 255).astype(np.int64)


class Canvas:
    """camera.rs:263-296: per-pixel colour SUMS over `samples` samples, row-major."""

    def __init__(self, samples, width, height, data):
        self.samples, self.width, self.height = int(samples), int(width), int(height)
        if isinstance(data, np.ndarray) and data.dtype == np.float32:
            # sums straight from the device: held as the f32 they were accumulated in, widened to the reference's f64
            # `Color` on first access (`data`)
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

    def merge(self, other: "Canvas") -> "Canvas":
        assert self.width == other.width
        assert self.height == other.height
        assert self.data.size == other.data.size
        return Canvas(self.samples + other.samples, self.width, self.height, self.data + other.data)

    # ---- bincode 1.3.3 image of `#[derive(Serialize, Deserialize)] struct Canvas` (camera.rs:263-270), the
    # checkpoint format of examples/common/mod.rs:24-56: usize fields as u64 LE, `Vec<Color>` as a u64 length
    # followed by the elements, a `Color` = `Vec3([f64; 3])` as three f64 LE (fixed-size arrays carry no length)
    def to_bincode(self) -> bytes:
        import struct
        head = struct.pack("<QQQQ", self.samples, self.width, self.height, self.width * self.height)
        return head + np.ascontiguousarray(self.data, "<f8").tobytes()

    @classmethod
    def from_bincode(cls, blob: bytes) -> "Canvas":
        import struct
        if len(blob) < 32:
            raise ValueError("io error: unexpected end of file")  # bincode's ErrorKind::Io
        samples, width, height, n = struct.unpack_from("<QQQQ", blob, 0)
        if len(blob) < 32 + 24 * n:
            raise ValueError("io error: unexpected end of file")
        data = np.frombuffer(blob, "<f8", count=3 * n, offset=32)
        if n != width * height:
            raise ValueError("checkpoint pixel count does not match width x height")
        return cls(samples, width, height, data.copy())

    def pixel_data(self) -> np.ndarray:
        # `c / samples` is `c * (1.0 / samples)` (vec3.rs:178-184)
        return self.data * (1.0 / float(self.samples))

    def to_u8(self) -> np.ndarray:
        return to_u8(srgb.linear_to_srgb(self.pixel_data()))

    def __eq__(self, other):
        return (isinstance(other, Canvas) and self.samples == other.samples and
                self.width == other.width and self.height == other.height and
                np.array_equal(self.data, other.data))


class output:
    @staticmethod
    def output_ppm(canvas: Canvas) -> str:
        """output.rs:5-14 + color.rs:22-29: one `r g b` line per pixel."""
        u8 = canvas.to_u8().reshape(-1, 3)
        lines = [f"P3\n{canvas.width} {canvas.height}\n255\n"]
        lines.append("".join(f"{r} {g} {b}\n" for r, g, b in u8.tolist()))
        return "".join(lines)


# ---- camera ----------------------------------------------------------------------------------------


@dataclass
class CameraParams:
    """camera.rs:24-59 (defaults identical)."""
    aspect_ratio: float = 1.0
    image_width: int = 100
    samples_per_pixel: int = 10
    max_depth: int = 10
    vfov: float = 90.0
    lookfrom: tuple = (0.0, 0.0, 0.0)
    lookat: tuple = (0.0, 0.0, -1.0)
    vup: tuple = (0.0, 1.0, 0.0)
    defocus_angle: float = 0.0
    focus_dist: float = 10.0
    background: tuple = (0.7, 0.8, 1.0)
    seed: int = 0

    def abi(self) -> A.rl_ow_camera:
        c = A.rl_ow_camera()
        c.aspect_ratio = float(self.aspect_ratio)
        c.image_width = int(self.image_width)
        c.samples_per_pixel = int(self.samples_per_pixel)
        c.max_depth = int(self.max_depth)
        c.vfov = float(self.vfov)
        c.lookfrom = (A.C.c_double * 3)(*map(float, self.lookfrom))
        c.lookat = (A.C.c_double * 3)(*map(float, self.lookat))
        c.vup = (A.C.c_double * 3)(*map(float, self.vup))
        c.defocus_angle = float(self.defocus_angle)
        c.focus_dist = float(self.focus_dist)
        c.background = (A.C.c_double * 3)(*map(float, self.background))
        c.seed = int(self.seed) & 0xFFFFFFFFFFFFFFFF
        return c


class Camera:
    def __init__(self, params: CameraParams):
        self.params = params
        # camera.rs:75 — `((image_width as f64 / aspect_ratio) as usize).max(1)`
        self.image_height = max(int(params.image_width / params.aspect_ratio), 1)
        d = (params.lookfrom[0] - params.lookat[0], params.lookfrom[1] - params.lookat[1],
             params.lookfrom[2] - params.lookat[2])
        if d[0] * d[0] + d[1] * d[1] + d[2] * d[2] <= 1e-16:
            raise ValueError("cannot normalize vector with magnitude 0")  # camera.rs:87 unwrap

    @classmethod
    def new(cls, params):
        return cls(params)

    def _render(self, samples_already_rendered: int, world, ctx=None) -> Canvas:
        from .context import default_context
        ctx = ctx or default_context()
        sd = world if isinstance(world, SceneDesc) else lower_world(world)
        ctx.scene_upload(sd)
        sums, _ = ctx.render_ow(self.params.abi(), samples_already_rendered)
        return Canvas(self.params.samples_per_pixel, self.params.image_width, self.image_height, sums)

    def render(self, world, ctx=None) -> Canvas:
        """Drop-in for `Camera::render` (camera.rs:122-124)."""
        return self._render(0, world, ctx)

    def render_from_checkpoint(self, world, checkpoint: Canvas, ctx=None) -> Canvas:
        """camera.rs:136-143 — sample indices continue from `checkpoint.samples`."""
        return self._render(checkpoint.samples, world, ctx).merge(checkpoint)


# ---- OBJ ingest (io/wavefront_obj.rs) ----------------------------------------------------------------


def _f(s):
    try:
        return float(s)
    except ValueError:
        return None


class WavefrontObj:
    def __init__(self):
        self.ignored = 0
        self.groups: dict = {}
        self.vertices, self.normals, self.texture_coords = [], [], []

    @classmethod
    def parse(cls, text) -> "WavefrontObj":
        if isinstance(text, bytes):
            text = text.decode()
        obj = cls()
        name, cur = None, []
        for line in text.splitlines():
            head, sep, tail = line.partition(" ")
            ok = False
            if sep:
                t = tail.strip()
                if head in ("v", "vn"):
                    nums = [_f(s) for s in t.split()]
                    if len(nums) == 3 and None not in nums:
                        (obj.vertices if head == "v" else obj.normals).append(tuple(nums))
                        ok = True
                elif head == "vt":
                    nums = [_f(s) for s in t.split()]
                    if nums and None not in nums:
                        obj.texture_coords.append((nums[0], 0.0) if len(nums) == 1 else (nums[0], nums[1]))
                        ok = True
                elif head == "f":
                    tris = obj._face(t)
                    if tris is not None:
                        cur.extend(tris)
                        ok = True
                elif head == "g":
                    obj.groups[name] = cur
                    name, cur = t, []
                    ok = True
            if not ok:
                obj.ignored += 1
        obj.groups[name] = cur
        return obj

    def _face(self, tail):
        idx = []
        for tok in tail.split():
            parts = tok.split("/")
            if not 1 <= len(parts) <= 3:
                return None
            if not parts[0].isdigit():
                return None
            ti = int(parts[1]) if len(parts) >= 2 and parts[1].isdigit() else None
            ni = int(parts[2]) if len(parts) == 3 and parts[2].isdigit() else None
            idx.append((int(parts[0]), ti, ni))
        if len(idx) < 3:
            return None
        verts = [(self.vertices[v - 1], None if t is None else self.texture_coords[t - 1],
                  None if n is None else self.normals[n - 1]) for v, t, n in idx]
        out = []
        for i in range(2, len(verts)):
            a, b, c = verts[0], verts[i - 1], verts[i]
            pts = [a[0], b[0], c[0]]
            uv = [a[1], b[1], c[1]] if None not in (a[1], b[1], c[1]) else None
            ns = [a[2], b[2], c[2]] if None not in (a[2], b[2], c[2]) else None
            out.append((pts, uv, ns))
        return out

    def tris(self):
        return [t for g in self.groups.values() for t in g]

    def to_object(self, material: Material) -> Bvh:
        return Bvh.new([Triangle.from_model(p, uv, n, material) for p, uv, n in self.tris()])
