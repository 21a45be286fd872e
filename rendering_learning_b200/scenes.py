"""Scene builders written against the mirrored reference API.

* the reference's own golden-test scenes (RTC/tests/ray_tracer.rs:56-368,
  OW/tests/ray_tracing_one_weekend.rs:14-75), and
* the five BASELINE.json configs: C1 three spheres on a plane, C2 mirror scene,
  C3 teapot, C4 RTIOW cover scene (OW/examples/bouncing_spheres.rs:16-134), C5 Cornell box + spot
  (OW/examples/cow.rs:17-139).

Mesh / texture inputs come from rendering_learning_b200/assets/*.npz (parsed from the reference's
objs/ by tests/golden/make_fixtures.py, because /root/reference does not exist on the GPU box).
"""
from __future__ import annotations

import math
import os

import numpy as np

from . import rtc

ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")

# std::f64::consts — decimal literals as in Rust's core (FRAC_PI_3 / FRAC_PI_6 are NOT equal to
# math.pi/3, math.pi/6: the literals round to the nearest double of the true quotient)
FRAC_PI_2 = 1.57079632679489661923132169163975144
FRAC_PI_3 = 1.04719755119659774615421446109316763
FRAC_PI_4 = 0.785398163397448309615660845819875721
FRAC_PI_6 = 0.52359877559829887307710723054658381


def _inv(m):
    return rtc.InvertibleMatrix.try_from(m)


# --------------------------------------------------------------------------------------------------
# RTC
# --------------------------------------------------------------------------------------------------

def rtc_mirror_world() -> rtc.World:
    """RTC/tests/ray_tracer.rs:56-225 (`test_mirror_scene`, world part)."""
    T = rtc.transformation
    gs1 = rtc.Transformed.new(rtc.Sphere.unit(), _inv(T.translation(-0.5, 0.0, 0.0)))
    gs2 = rtc.Transformed.new(rtc.Sphere.unit(), _inv(T.translation(0.5, 0.0, 0.0)))
    sphere_group = rtc.Bounded.new(rtc.Transformed.new(
        rtc.Group.new([gs1, gs2]),
        _inv(T.sequence([T.rotation_z(FRAC_PI_2), T.translation(-2.0, 2.0, 0.0)]))))
    floor = rtc.Plane(rtc.Material(
        surface=rtc.Surface.Pattern(rtc.Checker3d(a=rtc.color.white(), b=rtc.color.black(),
                                                  transform=_inv(T.translation(0.0, -0.01, 0.0)))),
        specular=0.0, reflectivity=0.02))
    left_wall = rtc.Transformed.new(
        rtc.Plane(rtc.Material(surface=rtc.color.white(), specular=1.0, reflectivity=0.9,
                               shininess=400.0, diffuse=0.0)),
        _inv(T.sequence([T.rotation_x(FRAC_PI_2), T.rotation_y(-FRAC_PI_3),
                         T.translation(-8.0, 0.0, 0.0)])))
    right_wall = rtc.Transformed.new(
        rtc.Plane(rtc.Material(surface=rtc.color.white(), specular=1.0, reflectivity=1.0,
                               shininess=400.0, diffuse=0.0)),
        _inv(T.sequence([T.rotation_x(FRAC_PI_2), T.rotation_y(FRAC_PI_4),
                         T.translation(10.0, 0.0, 0.0)])))
    middle_wall = rtc.Transformed.new(
        rtc.Plane(rtc.Material(surface=rtc.Color(0.945, 0.788, 0.647), specular=0.1,
                               shininess=50.0)),
        _inv(T.sequence([T.rotation_x(FRAC_PI_2), T.translation(0.0, 0.0, 7.0)])))
    ball = rtc.Transformed.new(
        rtc.Sphere(rtc.Material(surface=rtc.Color(0.059, 0.322, 0.729), diffuse=0.3, specular=1.0,
                                reflectivity=0.9, transparency=0.75, refractive_index=1.52)),
        _inv(T.translation(0.0, 2.0, 0.0)))
    inner_air_pocket = rtc.Transformed.new(
        rtc.Sphere(rtc.Material(surface=rtc.color.white(), ambient=0.0, diffuse=0.0, specular=0.0,
                                transparency=1.0, refractive_index=1.0, reflectivity=1.0)),
        _inv(T.sequence([T.scaling(0.5, 0.5, 0.5), T.translation(0.0, 2.0, 0.0)])))
    behind_cube = rtc.Transformed.new(
        rtc.Cube(rtc.Material(surface=rtc.Surface.Pattern(rtc.Stripe(
            a=rtc.Color(0.545, 0.0, 0.0), b=rtc.Color(0.0, 0.392, 0.0),
            transform=_inv(T.scaling(0.2, 1.0, 1.0)))))),
        _inv(T.translation(3.0, 0.0, -10.0)))
    behind_wall = rtc.Transformed.new(
        rtc.Plane(rtc.Material(surface=rtc.Color(0.678, 0.847, 0.902), specular=0.1,
                               shininess=50.0)),
        _inv(T.sequence([T.rotation_x(FRAC_PI_2), T.translation(0.0, 0.0, -100.0)])))
    light = rtc.PointLight(position=rtc.Point3d(-10.0, 10.0, -10.0), intensity=rtc.color.white())
    return rtc.World(objects=[floor, left_wall, right_wall, middle_wall, ball, inner_air_pocket,
                              behind_cube, behind_wall, sphere_group], lights=[light])


def rtc_mirror_scene(hsize=300, vsize=200) -> rtc.Scene:
    """RTC/tests/ray_tracer.rs:227-240 camera; BASELINE config C2 at 3840x2160."""
    T = rtc.transformation
    cam = rtc.Camera.new(hsize, vsize, FRAC_PI_3, _inv(T.view_transform(
        rtc.Point3d(0.0, 2.0, -7.0), rtc.Point3d(0.0, 1.5, 0.0), rtc.Vec3d(0.0, 1.0, 0.0))))
    return rtc.Scene(camera=cam, world=rtc_mirror_world())


def load_mesh(name: str) -> dict:
    return dict(np.load(os.path.join(ASSETS, name + ".npz")))


def rtc_teapot_object() -> rtc.Object:
    """`WavefrontObj::parse(teapot-low.obj).to_object()` rebuilt from the parsed fixture."""
    m = load_mesh("teapot_low")
    mat = rtc.Material()
    tris = []
    P, N = m["tri_p"], m["tri_n"]
    for i in range(P.shape[0]):
        pts = [tuple(P[i, k]) for k in range(3)]
        if m["tri_smooth"][i]:
            tris.append(rtc.Triangle.smooth([(pts[k], tuple(N[i, k])) for k in range(3)], mat))
        else:
            tris.append(rtc.Triangle.flat(pts, mat))
    return rtc.Bounded.new(rtc.Group.new(tris))


def rtc_obj_scene(hsize=300, vsize=200, obj: rtc.Object | None = None) -> rtc.Scene:
    """RTC/tests/ray_tracer.rs:242-275 (`test_obj_scene`); BASELINE config C3 at 3840x2160."""
    T = rtc.transformation
    o = rtc.Transformed.new(obj if obj is not None else rtc_teapot_object(),
                            _inv(T.sequence([T.rotation_x(-FRAC_PI_2)])))
    light = rtc.PointLight(position=rtc.Point3d(-2.0, 20.0, -30.0), intensity=rtc.color.white())
    world = rtc.World(objects=[o], lights=[light])
    cam = rtc.Camera.new(hsize, vsize, FRAC_PI_3, _inv(T.view_transform(
        rtc.Point3d(0.0, 15.0, -30.0), rtc.Point3d(0.0, 5.0, 0.0), rtc.Vec3d(0.0, 1.0, 0.0))))
    return rtc.Scene(camera=cam, world=world)


def rtc_csg_scene(hsize=300, vsize=200) -> rtc.Scene:
    """RTC/tests/ray_tracer.rs:277-368 (`test_csg_scene`)."""
    T = rtc.transformation
    room = rtc.Transformed.new(
        rtc.Cube(rtc.Material(
            surface=rtc.Surface.Pattern(rtc.Checker3d(
                a=rtc.Color(0.6, 0.6, 0.6), b=rtc.Color(0.7, 0.7, 0.7),
                transform=_inv(T.sequence([T.translation(0.01, 0.01, 0.01),
                                           T.scaling(0.02, 0.02, 0.02)])))),
            reflectivity=0.0, ambient=0.5, shininess=10.0, diffuse=0.3, specular=0.3)),
        _inv(T.scaling(50.0, 50.0, 50.0)))
    hollow_circle = rtc.Csg(
        left=rtc.Sphere(rtc.Material(surface=rtc.color.green())),
        right=rtc.Transformed.new(rtc.Sphere(rtc.Material(surface=rtc.color.blue())),
                                  _inv(T.scaling(0.7, 0.7, 0.7))),
        operation=rtc.CsgOperation.Difference)
    obj = rtc.Csg(
        left=hollow_circle,
        right=rtc.Transformed.new(rtc.Cube(rtc.Material(surface=rtc.color.red())),
                                  _inv(T.translation(1.0, 0.0, 0.0))),
        operation=rtc.CsgOperation.Difference)
    obj_t = rtc.Transformed.new(obj, _inv(T.sequence([T.rotation_y(FRAC_PI_6),
                                                      T.scaling(7.0, 7.0, 7.0)])))
    l1 = rtc.PointLight(position=rtc.Point3d(-2.0, 20.0, -30.0), intensity=rtc.Color(0.5, 0.5, 0.5))
    l2 = rtc.PointLight(position=rtc.Point3d(10.0, 20.0, -30.0), intensity=rtc.Color(0.5, 0.5, 0.5))
    world = rtc.World(objects=[room, obj_t], lights=[l1, l2])
    cam = rtc.Camera.new(hsize, vsize, FRAC_PI_3, _inv(T.view_transform(
        rtc.Point3d(0.0, 0.0, -30.0), rtc.Point3d(0.0, 0.0, 0.0), rtc.Vec3d(0.0, 1.0, 0.0))))
    return rtc.Scene(camera=cam, world=world)


def rtc_three_spheres_scene(hsize=1920, vsize=1080) -> rtc.Scene:
    """BASELINE config C1: book ch. 9 poses assembled from the reference API."""
    T = rtc.transformation
    floor = rtc.Plane(rtc.Material(surface=rtc.Color(1.0, 0.9, 0.9), specular=0.0))
    middle = rtc.Transformed.new(
        rtc.Sphere(rtc.Material(surface=rtc.Color(0.1, 1.0, 0.5), diffuse=0.7, specular=0.3)),
        _inv(T.translation(-0.5, 1.0, 0.5)))
    right = rtc.Transformed.new(
        rtc.Sphere(rtc.Material(surface=rtc.Color(0.5, 1.0, 0.1), diffuse=0.7, specular=0.3)),
        _inv(matmul_seq(T.translation(1.5, 0.5, -0.5), T.scaling(0.5, 0.5, 0.5))))
    left = rtc.Transformed.new(
        rtc.Sphere(rtc.Material(surface=rtc.Color(1.0, 0.8, 0.1), diffuse=0.7, specular=0.3)),
        _inv(matmul_seq(T.translation(-1.5, 0.33, -0.75), T.scaling(0.33, 0.33, 0.33))))
    light = rtc.PointLight(position=rtc.Point3d(-10.0, 10.0, -10.0), intensity=rtc.color.white())
    world = rtc.World(objects=[floor, middle, right, left], lights=[light])
    cam = rtc.Camera.new(hsize, vsize, FRAC_PI_3, _inv(T.view_transform(
        rtc.Point3d(0.0, 1.5, -5.0), rtc.Point3d(0.0, 1.0, 0.0), rtc.Vec3d(0.0, 1.0, 0.0))))
    return rtc.Scene(camera=cam, world=world)


def matmul_seq(*ms):
    """a * b * c ... (left to right matrix product, `&a * &b` in the reference)."""
    acc = ms[0]
    for m in ms[1:]:
        acc = rtc.matmul(acc, m)
    return acc


# --------------------------------------------------------------------------------------------------
# OW
# --------------------------------------------------------------------------------------------------

def ow_test_scene():
    """OW/tests/ray_tracing_one_weekend.rs:14-75 (`test_scene`): returns (world, CameraParams)."""
    from . import ow
    world = [
        ow.Sphere(ow.Center.Stationary(ow.Point3(0.0, -100.5, -1.0)), 100.0,
                  ow.Lambertian(ow.SolidColor(ow.Color(0.8, 0.8, 0.0)))),
        ow.Sphere(ow.Center.Stationary(ow.Point3(0.0, 0.0, -1.2)), 0.5,
                  ow.Lambertian(ow.SolidColor(ow.Color(0.1, 0.2, 0.5)))),
        ow.Sphere(ow.Center.Stationary(ow.Point3(-1.0, 0.0, -1.0)), 0.5, ow.Dielectric(1.5)),
        ow.Sphere(ow.Center.Stationary(ow.Point3(-1.0, 0.0, -1.0)), 0.4, ow.Dielectric(1.0 / 1.5)),
        ow.Sphere(ow.Center.Stationary(ow.Point3(1.0, 0.0, -1.0)), 0.5,
                  ow.Metal(ow.Color(0.8, 0.6, 0.2), 1.0)),
    ]
    params = ow.CameraParams(aspect_ratio=16.0 / 9.0, image_width=300, samples_per_pixel=10,
                             max_depth=10, vfov=20.0, lookfrom=ow.Point3(-2.0, 2.0, 1.0),
                             lookat=ow.Point3(0.0, 0.0, -1.0), vup=ow.Vec3(0.0, 1.0, 0.0),
                             defocus_angle=10.0, focus_dist=3.4,
                             background=ow.Color(0.7, 0.8, 1.0), seed=0)
    return world, params


class _Xoshiro256PlusPlus:
    """rand_xoshiro 0.6.0 `Xoshiro256PlusPlus::seed_from_u64` (SplitMix64 seeding) + rand 0.8.5 float
    sampling.  Used only to lay out the cover scene like OW/examples/bouncing_spheres.rs:16; the crate
    source is not vendored, so the exact layout is "parity unpinned"  — the oracle and the
    GPU consume the SAME generated spheres, which is all the parity tests need."""
    M = (1 << 64) - 1

    def __init__(self, seed: int):
        s = seed & self.M
        st = []
        for _ in range(4):
            s = (s + 0x9E3779B97F4A7C15) & self.M
            z = s
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & self.M
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & self.M
            st.append(z ^ (z >> 31))
        self.s = st

    @staticmethod
    def _rotl(x, k):
        return ((x << k) | (x >> (64 - k))) & _Xoshiro256PlusPlus.M

    def next_u64(self) -> int:
        s = self.s
        r = (self._rotl((s[0] + s[3]) & self.M, 23) + s[0]) & self.M
        t = (s[1] << 17) & self.M
        s[2] ^= s[0]
        s[3] ^= s[1]
        s[1] ^= s[2]
        s[0] ^= s[3]
        s[2] ^= t
        s[3] = self._rotl(s[3], 45)
        return r

    def gen(self) -> float:
        return (self.next_u64() >> 11) * (1.0 / 9007199254740992.0)

    def gen_range(self, low: float, high: float) -> float:
        import struct
        scale = high - low
        while True:
            bits = (self.next_u64() >> 12) | 0x3FF0000000000000
            v12 = struct.unpack("<d", struct.pack("<Q", bits))[0]
            res = v12 * scale + (low - scale)
            if res < high:
                return res

    def uniform_m1_1(self) -> float:
        """rand 0.8.5 `Uniform::<f64>::new(-1.0, 1.0)` as rand_distr samples it: 52-bit v in [0, 1), v * 2 - 1"""
        import struct
        bits = (self.next_u64() >> 12) | 0x3FF0000000000000
        return (struct.unpack("<d", struct.pack("<Q", bits))[0] - 1.0) * 2.0 + -1.0

    def unit_sphere(self):
        """rand_distr 0.4.3 UnitSphere (Marsaglia 1972) — `Vec3::random_unit_vector` (vec3.rs:72-75)"""
        import math
        while True:
            x1, x2 = self.uniform_m1_1(), self.uniform_m1_1()
            s2 = x1 * x1 + x2 * x2
            if s2 >= 1.0:
                continue
            f = 2.0 * math.sqrt(1.0 - s2)
            return (x1 * f, x2 * f, 1.0 - 2.0 * s2)

    def gen_range_usize(self, n: int) -> int:
        """rand 0.8.5 `gen_range(0..n)` for usize: widening multiply + zone rejection"""
        zone = ((n << (64 - n.bit_length())) - 1) & self.M
        while True:
            m = self.next_u64() * n
            hi, lo = m >> 64, m & self.M
            if lo <= zone:
                return hi


Xoshiro256PlusPlus = _Xoshiro256PlusPlus


def ow_cover_world():
    """OW/examples/bouncing_spheres.rs:16-117 — the RTIOW cover scene (BASELINE config C4)."""
    from . import ow
    rng = _Xoshiro256PlusPlus(1)
    world = []
    checker = ow.Checker.new(0.32, ow.SolidColor(ow.Color(0.2, 0.23, 0.1)), ow.SolidColor(ow.Color(0.9, 0.9, 0.9)))
    world.append(ow.Sphere(ow.Center.Stationary(ow.Point3(0.0, -1000.0, 0.0)), 1000.0, ow.Lambertian(checker)))
    for a in range(-11, 11):
        for b in range(-11, 11):
            choose_mat = rng.gen()
            cx = a + 0.9 * rng.gen()
            cz = b + 0.9 * rng.gen()
            center = ow.Point3(cx, 0.2, cz)
            dx, dy, dz = cx - 4.0, 0.2 - 0.2, cz - 0.0
            if math.sqrt(dx * dx + dy * dy + dz * dz) > 0.9:
                if choose_mat < 0.8:
                    c2 = ow.Point3(cx + 0.0, 0.2 + rng.gen_range(0.0, 0.5), cz + 0.0)
                    c_a = (rng.gen(), rng.gen(), rng.gen())
                    c_b = (rng.gen(), rng.gen(), rng.gen())
                    albedo = ow.Color(c_a[0] * c_b[0], c_a[1] * c_b[1], c_a[2] * c_b[2])
                    world.append(ow.Sphere(ow.Center.Moving(center, c2), 0.2, ow.Lambertian(ow.SolidColor(albedo))))
                elif choose_mat < 0.95:
                    albedo = ow.Color(rng.gen_range(0.5, 1.0), rng.gen_range(0.5, 1.0), rng.gen_range(0.5, 1.0))
                    fuzz = rng.gen()
                    world.append(ow.Sphere(ow.Center.Stationary(center), 0.2, ow.Metal(albedo, fuzz)))
                else:
                    world.append(ow.Sphere(ow.Center.Stationary(center), 0.2, ow.Dielectric(1.5)))
    world.append(ow.Sphere(ow.Center.Stationary(ow.Point3(0.0, 1.0, 0.0)), 1.0, ow.Dielectric(1.5)))
    world.append(ow.Sphere(ow.Center.Stationary(ow.Point3(-4.0, 1.0, 0.0)), 1.0,
                           ow.Lambertian(ow.SolidColor(ow.Color(0.4, 0.2, 0.1)))))
    world.append(ow.Sphere(ow.Center.Stationary(ow.Point3(4.0, 1.0, 0.0)), 1.0, ow.Metal(ow.Color(0.7, 0.6, 0.5), 0.0)))
    return ow.Bvh.new(world)


def ow_cover_params(image_width=1200, samples_per_pixel=500, max_depth=50, seed=0):
    """bouncing_spheres.rs:121-131 camera at the BASELINE C4 size (1200x675, 500 spp, depth 50)."""
    from . import ow
    return ow.CameraParams(aspect_ratio=16.0 / 9.0, image_width=image_width,
                           samples_per_pixel=samples_per_pixel, max_depth=max_depth, vfov=20.0,
                           lookfrom=ow.Point3(13.0, 2.0, 3.0), lookat=ow.Point3(0.0, 0.0, 0.0),
                           vup=ow.Vec3(0.0, 1.0, 0.0), defocus_angle=0.6, focus_dist=10.0, seed=seed)


def ow_perlin_spheres():
    """OW/examples/perlin_spheres.rs:14-57 — two spheres with the marble Noise texture"""
    from . import ow
    rng = _Xoshiro256PlusPlus(1)
    mat = ow.Lambertian(ow.Noise(ow.Perlin.new(rng), 4.0))
    world = [ow.Sphere(ow.Center.Stationary((0.0, -1000.0, 0.0)), 1000.0, mat),
             ow.Sphere(ow.Center.Stationary((0.0, 2.0, 0.0)), 2.0, mat)]
    params = ow.CameraParams(aspect_ratio=16.0 / 9.0, image_width=400, samples_per_pixel=100, max_depth=50, vfov=20.0,
                             lookfrom=(13.0, 2.0, 3.0), lookat=(0.0, 0.0, 0.0), vup=(0.0, 1.0, 0.0), defocus_angle=0.0)
    return world, params


def _ow_box(a, b, material):
    """examples/common/mod.rs make_box: the six quads of the axis-aligned box with opposite corners a, b"""
    from . import ow
    lo = tuple(min(x, y) for x, y in zip(a, b))
    hi = tuple(max(x, y) for x, y in zip(a, b))
    dx, dy, dz = (hi[0] - lo[0], 0.0, 0.0), (0.0, hi[1] - lo[1], 0.0), (0.0, 0.0, hi[2] - lo[2])
    neg = lambda v: (-v[0], -v[1], -v[2])
    return [ow.Quad.new((lo[0], lo[1], hi[2]), dx, dy, material), ow.Quad.new((hi[0], lo[1], hi[2]), neg(dz), dy, material),
            ow.Quad.new((hi[0], lo[1], lo[2]), neg(dx), dy, material), ow.Quad.new((lo[0], lo[1], lo[2]), dz, dy, material),
            ow.Quad.new((lo[0], hi[1], hi[2]), dx, neg(dz), material), ow.Quad.new((lo[0], lo[1], lo[2]), dx, dz, material)]


def ow_cornell_smoke():
    """OW/examples/cornell_smoke.rs — Cornell box with two boxes of smoke (ConstantMedium + Isotropic)"""
    from . import ow
    red = ow.Lambertian(ow.SolidColor((0.65, 0.05, 0.05)))
    white = ow.Lambertian(ow.SolidColor((0.73, 0.73, 0.73)))
    green = ow.Lambertian(ow.SolidColor((0.12, 0.45, 0.15)))
    light = ow.DiffuseLight(ow.SolidColor((7.0, 7.0, 7.0)))
    world = [ow.Quad.new((555.0, 0.0, 0.0), (0.0, 555.0, 0.0), (0.0, 0.0, 555.0), green),
             ow.Quad.new((0.0, 0.0, 0.0), (0.0, 555.0, 0.0), (0.0, 0.0, 555.0), red),
             ow.Quad.new((113.0, 554.0, 127.0), (330.0, 0.0, 0.0), (0.0, 0.0, 305.0), light),
             ow.Quad.new((0.0, 555.0, 0.0), (555.0, 0.0, 0.0), (0.0, 0.0, 555.0), white),
             ow.Quad.new((0.0, 0.0, 0.0), (555.0, 0.0, 0.0), (0.0, 0.0, 555.0), white),
             ow.Quad.new((0.0, 0.0, 555.0), (555.0, 0.0, 0.0), (0.0, 555.0, 0.0), white)]
    box1 = ow.HittableList(_ow_box((0.0, 0.0, 0.0), (165.0, 330.0, 165.0), white)).rotate_y(15.0).translate((265.0, 0.0, 295.0))
    box2 = ow.HittableList(_ow_box((0.0, 0.0, 0.0), (165.0, 165.0, 165.0), white)).rotate_y(-18.0).translate((130.0, 0.0, 65.0))
    world.append(ow.ConstantMedium.new(box1, 0.01, ow.Isotropic(ow.SolidColor((0.0, 0.0, 0.0)))))
    world.append(ow.ConstantMedium.new(box2, 0.01, ow.Isotropic(ow.SolidColor((1.0, 1.0, 1.0)))))
    params = ow.CameraParams(aspect_ratio=1.0, image_width=600, samples_per_pixel=200, max_depth=50, vfov=40.0,
                             lookfrom=(278.0, 278.0, -800.0), lookat=(278.0, 278.0, 0.0), vup=(0.0, 1.0, 0.0),
                             defocus_angle=0.0, background=(0.0, 0.0, 0.0))
    return world, params


def ow_final_scene(image_width=400, samples_per_pixel=250, max_depth=4):
    """OW/examples/final_scene.rs:35-258 — the "next week" final scene: 400 ground boxes in a Bvh, a light, a moving
    sphere, glass, fuzzy metal, a glass sphere filled with a blue ConstantMedium, a global fog (radius-5000 boundary),
    an image-textured globe, a Perlin sphere and a rotated + translated Bvh of 1000 small spheres.  The reference embeds
    examples/files/earthmap.jpg; this builder uses a procedural 256 x 128 map instead (no binary asset travels), which
    is irrelevant to parity: oracle and device consume the same description."""
    from . import ow
    rng = _Xoshiro256PlusPlus(0)
    ground = ow.Lambertian(ow.SolidColor((0.48, 0.83, 0.53)))
    light = ow.DiffuseLight(ow.SolidColor((7.0, 7.0, 7.0)))
    sphere_material = ow.Lambertian(ow.SolidColor((0.7, 0.3, 0.1)))
    glass = ow.Dielectric(1.5)
    metal = ow.Metal((0.8, 0.8, 0.9), 1.0)
    subsurface = ow.Isotropic(ow.SolidColor((0.2, 0.4, 0.9)))
    fog = ow.Isotropic(ow.SolidColor((1.0, 1.0, 1.0)))
    v, u = np.meshgrid(np.linspace(0, 1, 128), np.linspace(0, 1, 256), indexing="ij")
    srgb_map = np.stack([0.5 + 0.5 * np.sin(12 * np.pi * u), 0.5 + 0.5 * np.cos(6 * np.pi * v), 0.3 + 0.4 * u * v], axis=2)
    earth = ow.Lambertian(ow.Image(ow.srgb.srgb_to_linear(srgb_map).astype(np.float32)))
    perlin = ow.Lambertian(ow.Noise(ow.Perlin.new(rng), 0.2))
    white = ow.Lambertian(ow.SolidColor((0.73, 0.73, 0.73)))
    world = []
    boxes = []
    for i in range(20):
        for j in range(20):
            w = 100.0
            x0, z0, y0 = -1000.0 + i * w, -1000.0 + j * w, 0.0
            y1 = rng.gen_range(1.0, 101.0)
            boxes.append(ow.HittableList(_ow_box((x0, y0, z0), (x0 + w, y1, z0 + w), ground)))
    world.append(ow.Bvh.new(boxes))
    world.append(ow.Quad.new((123.0, 554.0, 147.0), (300.0, 0.0, 0.0), (0.0, 0.0, 265.0), light))
    world.append(ow.Sphere(ow.Center.Moving((400.0, 400.0, 200.0), (430.0, 400.0, 200.0)), 50.0, sphere_material))
    world.append(ow.Sphere(ow.Center.Stationary((260.0, 150.0, 45.0)), 50.0, glass))
    world.append(ow.Sphere(ow.Center.Stationary((0.0, 150.0, 145.0)), 50.0, metal))
    boundary = lambda: ow.Sphere(ow.Center.Stationary((360.0, 150.0, 145.0)), 70.0, glass)
    world.append(boundary())
    world.append(ow.ConstantMedium.new(boundary(), 0.2, subsurface))
    world.append(ow.ConstantMedium.new(ow.Sphere(ow.Center.Stationary((0.0, 0.0, 0.0)), 5000.0, glass), 0.0001, fog))
    world.append(ow.Sphere(ow.Center.Stationary((400.0, 200.0, 400.0)), 100.0, earth))
    world.append(ow.Sphere(ow.Center.Stationary((220.0, 280.0, 300.0)), 80.0, perlin))
    small = [ow.Sphere(ow.Center.Stationary((rng.gen_range(0.0, 165.0), rng.gen_range(0.0, 165.0), rng.gen_range(0.0, 165.0))),
                       10.0, white) for _ in range(1000)]
    world.append(ow.Bvh.new(small).rotate_y(15.0).translate((-100.0, 270.0, 395.0)))
    params = ow.CameraParams(aspect_ratio=1.0, image_width=image_width, samples_per_pixel=samples_per_pixel,
                             max_depth=max_depth, vfov=40.0, lookfrom=(478.0, 278.0, -600.0), lookat=(278.0, 278.0, 0.0),
                             vup=(0.0, 1.0, 0.0), defocus_angle=0.0, background=(0.0, 0.0, 0.0))
    return world, params


def ow_spot_texture() -> np.ndarray:
    """cow.rs:19-30: decode -> into_rgb32f (u8/255 as f32) -> srgb_to_linear in f64 -> f32."""
    from . import ow
    rgb8 = np.load(os.path.join(ASSETS, "spot_texture.npz"))["rgb8"]
    f = (rgb8.astype(np.float32) / np.float32(255.0)).astype(np.float64)
    return ow.srgb.srgb_to_linear(f).astype(np.float32)


def ow_cow_world(cow=None):
    """OW/examples/cow.rs:17-117 — Cornell box + textured spot (BASELINE config C5).  `cow`: the mesh object (e.g. an
    ow.DeviceMesh parsed on the GPU) instead of the host-parsed fixture."""
    from . import ow
    if cow is None:
        m = load_mesh("spot")
        cow_surface = ow.Lambertian(ow.Image(ow_spot_texture()))
        tris = []
        P, UV, N = m["tri_p"], m["tri_uv"], m["tri_n"]
        for i in range(P.shape[0]):
            pts = [tuple(P[i, k]) for k in range(3)]
            uv = [tuple(UV[i, k]) for k in range(3)] if m["has_uv"][i] else None
            ns = [tuple(N[i, k]) for k in range(3)] if m["has_n"][i] else None
            tris.append(ow.Triangle.from_model(pts, uv, ns, cow_surface))
        cow = ow.Bvh.new(tris)
    cow = cow.scale(200.0).rotate_y(45.0).translate(ow.Vec3(240.0, 165.0, 240.0))
    red = ow.Lambertian(ow.SolidColor(ow.Color(0.65, 0.05, 0.05)))
    white = ow.Lambertian(ow.SolidColor(ow.Color(0.73, 0.73, 0.73)))
    green = ow.Lambertian(ow.SolidColor(ow.Color(0.12, 0.45, 0.15)))
    light = ow.DiffuseLight(ow.SolidColor(ow.Color(5.0, 5.0, 5.0)))
    world = [
        ow.Quad.new(ow.Point3(555.0, 0.0, 0.0), ow.Vec3(0.0, 555.0, 0.0), ow.Vec3(0.0, 0.0, 555.0), green),
        ow.Quad.new(ow.Point3(0.0, 0.0, 0.0), ow.Vec3(0.0, 555.0, 0.0), ow.Vec3(0.0, 0.0, 555.0), red),
        ow.Quad.new(ow.Point3(113.0, 554.0, 127.0), ow.Vec3(330.0, 0.0, 0.0), ow.Vec3(0.0, 0.0, 305.0), light),
        ow.Quad.new(ow.Point3(0.0, 0.0, 0.0), ow.Vec3(555.0, 0.0, 0.0), ow.Vec3(0.0, 0.0, 555.0), white),
        ow.Quad.new(ow.Point3(555.0, 555.0, 555.0), ow.Vec3(-555.0, 0.0, 0.0), ow.Vec3(0.0, 0.0, -555.0), white),
        ow.Quad.new(ow.Point3(0.0, 0.0, 555.0), ow.Vec3(555.0, 0.0, 0.0), ow.Vec3(0.0, 555.0, 0.0), white),
        cow,
    ]
    return ow.Bvh.new(world)


def ow_cow_params(image_width=3840, samples_per_pixel=256, max_depth=40, seed=0, aspect_ratio=16.0 / 9.0):
    """cow.rs:121-136 camera at the BASELINE C5 size (3840x2160, 256 spp, depth 40)."""
    from . import ow
    return ow.CameraParams(aspect_ratio=aspect_ratio, image_width=image_width,
                           samples_per_pixel=samples_per_pixel, max_depth=max_depth,
                           background=ow.Color(0.0, 0.0, 0.0), vfov=40.0,
                           lookfrom=ow.Point3(278.0, 278.0, -800.0), lookat=ow.Point3(278.0, 278.0, 0.0),
                           vup=ow.Vec3(0.0, 1.0, 0.0), defocus_angle=0.0, seed=seed)
