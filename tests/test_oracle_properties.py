"""Properties of the REFERENCE's semantics that the device path relies on, checked on the oracle (CPU):
the flattener composes nested transforms into one per leaf, treats Bounded / Group as transparent and bakes OW
Transform / Translate chains — all of that must be result-preserving in the reference's own arithmetic."""
import math

import numpy as np
import pytest

from rendering_learning_b200 import ow, rtc, scenes

T = rtc.transformation


def _render(oracle, world, w=60, h=40):
    cam = rtc.Camera.new(w, h, 1.0, T.view_transform((0.5, 2.5, -6.0), (0, 0.8, 0), (0, 1, 0)))
    return oracle.rtc_render(world.lower(), cam.abi(), 1)


def _lights():
    return [rtc.PointLight((-6, 8, -6), (0.7, 0.7, 0.7)), rtc.PointLight((5, 6, -4), (0.3, 0.3, 0.3))]


def test_nested_transforms_equal_their_product(oracle):
    """transformed.rs:226-278: Transformed(Transformed(s, A), B) behaves like Transformed(s, B * A)"""
    A = T.sequence([T.scaling(0.5, 1.2, 0.8), T.rotation_z(0.4)])
    B = T.sequence([T.rotation_y(0.9), T.translation(0.7, 0.9, 0.3)])
    mat = rtc.Material(surface=rtc.Stripe(a=(1, 0, 0), b=(0, 0, 1), transform=T.scaling(0.2, 1, 1)), reflectivity=0.2)
    floor = rtc.Plane(rtc.Material(surface=(0.8, 0.8, 0.8)))
    nested = rtc.World(objects=[floor, rtc.Transformed.new(rtc.Transformed.new(rtc.Cube(mat), A), B)], lights=_lights())
    product = np.array(B) @ np.array(A)
    flat = rtc.World(objects=[floor, rtc.Transformed.new(rtc.Cube(mat), product.tolist())], lights=_lights())
    a, b = _render(oracle, nested), _render(oracle, flat)
    assert np.abs(a - b).max() < 1e-9 and a.std() > 0.05


def test_bounded_and_group_are_transparent(oracle):
    """bounded.rs:100-152 only skips work; a Group's list is its children's lists, sorted (group.rs:29-42)"""
    s1 = rtc.Transformed.new(rtc.Sphere(rtc.Material(surface=(0.9, 0.3, 0.1))), T.translation(-1.0, 1.0, 0.0))
    s2 = rtc.Transformed.new(rtc.Sphere(rtc.Material(surface=(0.1, 0.3, 0.9), transparency=0.6, refractive_index=1.3)),
                             T.translation(0.4, 1.0, -0.5))
    cyl = rtc.Transformed.new(rtc.Cylinder(rtc.Material(surface=(0.2, 0.8, 0.3)), minimum=0.0, maximum=1.5, closed=True),
                              T.translation(2.0, 0.0, 1.0))
    floor = rtc.Plane(rtc.Material(surface=(0.8, 0.8, 0.8), reflectivity=0.1))
    plain = rtc.World(objects=[floor, s1, s2, cyl], lights=_lights())
    wrapped = rtc.World(objects=[floor, rtc.Bounded.new(rtc.Group.new([s1, rtc.Bounded.new(s2)])), rtc.Group.new([cyl])],
                        lights=_lights())
    a, b = _render(oracle, plain), _render(oracle, wrapped)
    assert np.array_equal(a, b) and a.std() > 0.05


def test_group_transform_distributes_over_children(oracle):
    """Transformed<Group<..>> (draw_scene.rs:84-113): one outer transform == the same transform on every child"""
    M = T.sequence([T.rotation_y(0.5), T.scaling(0.8, 0.8, 0.8), T.translation(0.3, 0.2, 0.4)])
    kids = lambda: [rtc.Transformed.new(rtc.Sphere(rtc.Material(surface=(0.9, 0.6, 0.1))), T.translation(-1.2, 1.0, 0.0)),
                    rtc.Transformed.new(rtc.Cone(rtc.Material(surface=(0.3, 0.2, 0.8)), minimum=-1.0, maximum=0.0, closed=True),
                                        T.translation(1.0, 1.0, 0.0))]
    floor = rtc.Plane(rtc.Material(surface=(0.8, 0.8, 0.8)))
    outer = rtc.World(objects=[floor, rtc.Transformed.new(rtc.Group.new(kids()), M)], lights=_lights())
    inner = rtc.World(objects=[floor] + [rtc.Transformed.new(k, M) for k in kids()], lights=_lights())
    assert np.abs(_render(oracle, outer) - _render(oracle, inner)).max() < 1e-9


def test_ow_translate_and_transform_are_baked_exactly(oracle):
    """translate.rs:14-21, transform.rs:145-164: a translated / scaled sphere is the sphere with the moved centre and
    scaled radius; t is preserved because ray directions are never renormalised"""
    m = ow.Lambertian(ow.SolidColor((0.5, 0.5, 0.5)))
    moved = ow.lower_world([ow.Sphere(ow.Center.Stationary((0.0, 0.0, 0.0)), 0.5, m).scale(2.0).translate((1.0, 2.0, -3.0))])
    plain = ow.lower_world([ow.Sphere(ow.Center.Stationary((1.0, 2.0, -3.0)), 1.0, m)])
    rng = np.random.default_rng(2)
    o = rng.uniform(-6, 6, size=(2000, 3))
    d = np.array([1.0, 2.0, -3.0]) + rng.normal(scale=0.7, size=(2000, 3)) - o
    rays = np.concatenate([o, d, np.zeros((2000, 1))], axis=1)
    n1, t1, _ = oracle.ow_trace(moved, rays)
    n2, t2, _ = oracle.ow_trace(plain, rays)
    assert ((n1 >= 0) == (n2 >= 0)).all() and (n1 >= 0).mean() > 0.3
    hit = n1 >= 0
    assert np.abs(t1[hit] - t2[hit]).max() < 1e-9


def test_ow_bvh_is_result_independent(oracle):
    """bvh.rs:81-90: the closest hit does not depend on how the hittables are grouped (what lets the LBVH replace it)"""
    world, params = scenes.ow_test_scene()
    items = world.children if hasattr(world, "children") else list(world)
    flat = ow.lower_world(ow.HittableList(items))
    bvh = ow.lower_world(ow.Bvh.new(items))
    rays = oracle.ow_camera_rays(params.abi())[::7]
    n1, t1, _ = oracle.ow_trace(flat, rays)
    n2, t2, _ = oracle.ow_trace(bvh, rays)
    assert np.array_equal(t1, t2) and ((n1 >= 0) == (n2 >= 0)).all()
