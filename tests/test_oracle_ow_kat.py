"""The reference's OW unit known-answer vectors (hittable/sphere.rs:102-203, hittable/mod.rs:113-203) replayed on the
oracle through `world.hit(r, [1e-10, inf))`."""
import numpy as np

from rendering_learning_b200 import ow

FLAT = ow.Lambertian(ow.SolidColor((0.5, 0.5, 0.5)))  # `Flat` in the reference's tests: the material is never asked


def unit_sphere(c=(0.0, 0.0, 0.0)):
    return ow.Sphere(ow.Center.Stationary(c), 1.0, FLAT)


def trace(oracle, world, o, d):
    desc = ow.lower_world(world)
    node, t, uv = oracle.ow_trace(desc, np.array([list(o) + list(d) + [0.0]]))
    return int(node[0]), float(t[0]), uv[0], desc


def test_sphere_hit_vectors(oracle):
    # a_ray_misses_a_sphere
    assert trace(oracle, [unit_sphere()], (0, 2, 5), (0, 0, -1))[0] == -1
    # a_ray_is_tangent_to_a_sphere: t == 5 exactly
    n, t, _, _ = trace(oracle, [unit_sphere()], (0, 1, 5), (0, 0, -1))
    assert n >= 0 and t == 5.0
    # a_ray_goes_through_a_sphere: nearest root
    assert trace(oracle, [unit_sphere()], (0, 0, 5), (0, 0, -1))[1] == 4.0
    # a_ray_starts_inside_a_sphere: the far root
    assert trace(oracle, [unit_sphere()], (0, 0, 0), (0, 0, -1))[1] == 1.0


def test_sphere_uv_vectors(oracle):
    # get_sphere_uv (sphere.rs:91-99) at the six axis points, via rays that hit the sphere exactly there
    cases = {(1, 0, 0): (0.5, 0.5), (0, 1, 0): (0.5, 1.0), (0, 0, 1): (0.25, 0.5),
             (-1, 0, 0): (0.0, 0.5), (0, -1, 0): (0.5, 0.0), (0, 0, -1): (0.75, 0.5)}
    for p, (eu, ev) in cases.items():
        o = tuple(2.0 * c for c in p)
        d = tuple(-float(c) for c in p)
        n, t, uv, _ = trace(oracle, [unit_sphere()], o, d)
        assert n >= 0 and abs(t - 1.0) < 1e-12
        u = uv[0] % 1.0  # u = 0 and u = 1 are the same meridian (the reference's epsilon = 0.01 comparison at -x)
        assert min(abs(u - eu), abs(u - eu - 1.0), abs(u - eu + 1.0)) < 0.01 and abs(uv[1] - ev) < 0.01


def test_slice_hit_vectors(oracle):
    back, middle, front = unit_sphere((0, 0, -10)), unit_sphere((0, 0, -5)), unit_sphere((0, 0, 0))
    # hitting_nothing
    assert trace(oracle, [back], (0, 0, 5), (0, 1, 0))[0] == -1
    # hitting_a_hittable
    assert trace(oracle, [back], (0, 0, 5), (0, 0, -1))[1] == 14.0
    # hitting_the_closest_hittable: order in the slice does not matter
    n, t, _, desc = trace(oracle, [back, front, middle], (0, 0, 5), (0, 0, -1))
    assert t == 4.0
    spheres = [i for i, nd in enumerate(desc.nodes) if nd[0] == 32]  # RL_OW_SPHERE
    assert n == spheres[1]  # the hit reports `front`, the second element


def test_moving_sphere_center_is_a_lerp_in_time(oracle):
    # sphere.rs:24-29: center(time) = c1 + time * (c2 - c1)
    s = ow.Sphere(ow.Center.Moving((0.0, 0.0, 0.0), (0.0, 2.0, 0.0)), 1.0, FLAT)
    desc = ow.lower_world([s])
    rays = np.array([[0, 0, 5, 0, 0, -1, 0.0], [0, 0, 5, 0, 0, -1, 0.5], [0, 0, 5, 0, 0, -1, 1.0]], dtype=np.float64)
    node, t, _ = oracle.ow_trace(desc, rays)
    assert t[0] == 4.0 and abs(t[1] - 5.0) < 1e-12 and node[2] == -1  # centre at y = 0, 1 (tangent), 2 (miss)
