"""The flattener (csrc/flatten.cpp) through the host-only `rl_scene_check` — no GPU: counts, the big list, CSG lowering,
constant media, and every RL_E_* the device path answers with instead of rendering something else."""
import math

import numpy as np
import pytest

from rendering_learning_b200 import RlError, ow, rtc, scenes
from rendering_learning_b200 import _abi as A

T = rtc.transformation


def test_baseline_scene_counts():
    i = scenes.rtc_three_spheres_scene(64, 36).world.lower().check()
    assert (i.flavor, i.n_prims, i.n_bvh_prims, i.n_lights, i.has_transparency) == (A.RL_FLAVOR_RTC, 4, 0, 1, 0)
    i = scenes.rtc_mirror_scene(30, 20).world.lower().check()
    assert i.n_prims == 10 and i.has_transparency == 1
    i = scenes.rtc_obj_scene(30, 20).world.lower().check()  # teapot: 240 triangles under the LBVH, no analytic prims
    assert (i.n_prims, i.n_bvh_prims, i.n_bvh_nodes) == (0, 240, 239)
    i = scenes.rtc_csg_scene(30, 20).world.lower().check()  # room cube + 3 Csg leaves
    assert i.n_prims == 4 and i.n_lights == 2
    # cover scene: the r = 1000 ground sphere leaves the LBVH for the big list
    i = ow.lower_world(scenes.ow_cover_world()).check()
    assert i.flavor == A.RL_FLAVOR_OW and i.n_prims == 1 and 400 < i.n_bvh_prims <= 487
    # Cornell box + spot: the six quads are large against the mesh -> big list; 5856 triangles stay
    i = ow.lower_world(scenes.ow_cow_world()).check()
    assert (i.n_prims, i.n_bvh_prims) == (6, 5856)
    assert i.device_bytes > 1024 * 1024 * 16  # the 1024^2 float4 texture


def test_big_list_rule():
    m = ow.Lambertian(ow.SolidColor((0.5, 0.5, 0.5)))
    rng = np.random.default_rng(0)
    small = [ow.Sphere(ow.Center.Stationary(tuple(c)), 0.2, m) for c in rng.uniform(-5, 5, size=(40, 3))]
    # uniformly sized primitives: nobody is big
    assert ow.lower_world(list(small)).check().n_prims == 0
    # a wall that dwarfs them is
    wall = ow.Quad.new((-50.0, -50.0, -6.0), (100.0, 0.0, 0.0), (0.0, 100.0, 0.0), m)
    i = ow.lower_world(small + [wall]).check()
    assert (i.n_prims, i.n_bvh_prims) == (1, 40)
    # fewer than 16 primitives: the LBVH is kept whole (a brute-force list would not pay)
    assert ow.lower_world(small[:5] + [wall]).check().n_prims == 0
    # at most OW_MAX_BIG = 8 primitives go on the list
    walls = [ow.Quad.new((-50.0, -50.0, -6.0 - k), (100.0, 0.0, 0.0), (0.0, 100.0, 0.0), m) for k in range(12)]
    i = ow.lower_world(small + walls).check()
    assert i.n_prims == 8 and i.n_bvh_prims == 44


def test_media_and_noise_lowering():
    i = ow.lower_world(scenes.ow_cornell_smoke()[0]).check()
    # 6 walls -> big list; the two media are ONE LBVH leaf each, their 2 x 6 boundary quads are in neither structure
    assert (i.n_prims, i.n_bvh_prims) == (0, 8) or (i.n_prims + i.n_bvh_prims) == 8
    i = ow.lower_world(scenes.ow_final_scene()[0]).check()
    assert i.n_prims + i.n_bvh_prims == 2400 + 1 + 4 + 2 + 2 + 1000  # boxes, light, spheres, media, globe + perlin, small spheres
    assert i.n_prims == 1  # the radius-5000 fog, and nothing else
    white = ow.Lambertian(ow.SolidColor((1.0, 1.0, 1.0)))
    box = ow.HittableList(scenes._ow_box((0.0, 0.0, 0.0), (1.0, 1.0, 1.0), white))
    with pytest.raises(RlError) as e:  # the reference only intends Isotropic as a phase function
        ow.lower_world([ow.ConstantMedium.new(box, 1.0, white)]).check()
    assert e.value.code == A.RL_E_UNSUPPORTED
    iso = ow.Isotropic(ow.SolidColor((1.0, 1.0, 1.0)))
    with pytest.raises(RlError) as e:
        ow.lower_world([ow.ConstantMedium.new(ow.HittableList([ow.ConstantMedium.new(box, 1.0, iso)]), 1.0, iso)]).check()
    assert e.value.code == A.RL_E_UNSUPPORTED


def test_error_codes():
    # unsupported constructs say so (RL_E_UNSUPPORTED) instead of rendering something else
    # round 1 refused triangles under a Csg and max_reflection_depth > 16; both lower now (csg.rs:31-35 is generic over
    # any Object, world.rs:29 is a usize)
    tri = rtc.Triangle.flat([(0, 0, 0), (1, 0, 0), (0, 1, 0)])
    info = rtc.World(objects=[rtc.Csg(rtc.Sphere(), tri, rtc.CsgOperation.Union)]).lower().check()
    assert info.n_prims == 2 and info.n_bvh_prims == 0  # the triangle is a leaf of the Csg's range, not an LBVH primitive
    w = scenes.rtc_mirror_world()
    w.max_reflection_depth = 99
    assert w.lower().check().n_prims > 0
    w.max_reflection_depth = -1
    with pytest.raises(RlError) as e:
        w.lower().check()
    assert e.value.code == A.RL_E_INVALID
    img = ow.Lambertian(ow.Image(np.ones((4, 8, 3), np.float32)))
    with pytest.raises(RlError) as e:  # sphere uv is defined in the sphere's own frame
        ow.lower_world([ow.Sphere(ow.Center.Stationary((0.0, 0.0, 0.0)), 1.0, img).rotate_y(30.0)]).check()
    assert e.value.code == A.RL_E_UNSUPPORTED
    assert ow.lower_world([ow.Sphere(ow.Center.Stationary((0.0, 0.0, 0.0)), 1.0, img).translate((1.0, 0.0, 0.0))]).check().n_bvh_prims == 1
    # malformed descriptions are RL_E_INVALID
    d = ow.lower_world([ow.Sphere(ow.Center.Stationary((0.0, 0.0, 0.0)), 1.0, ow.Dielectric(1.5))])
    d.nodes[-1] = (A.RL_OW_SPHERE, 7, -1, -1, 0, d.nodes[-1][5])  # material index out of range
    with pytest.raises(RlError) as e:
        d.check()
    assert e.value.code == A.RL_E_INVALID
    d = scenes.rtc_three_spheres_scene(8, 8).world.lower()
    d.freeze().abi_version = 1  # an ABI-1 caller (no perlins field) is refused, not misread
    with pytest.raises(RlError) as e:
        d.check()
    assert e.value.code == A.RL_E_INVALID
    with pytest.raises(RlError) as e:  # a non-affine matrix cannot be a Transformed
        m = [[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 1.0, 0], [0.1, 0, 0, 1.0]]
        rtc.World(objects=[rtc.Transformed.new(rtc.Sphere(), m)]).lower().check()
    assert e.value.code in (A.RL_E_UNSUPPORTED, A.RL_E_INVALID)


def test_csg_nesting_is_lowered_to_contiguous_ranges():
    a, b, c, d = (rtc.Sphere() for _ in range(4))
    inner1 = rtc.Csg(a, rtc.Transformed.new(b, T.translation(0.5, 0, 0)), rtc.CsgOperation.Union)
    inner2 = rtc.Csg(rtc.Group.new([c]), rtc.Bounded.new(d), rtc.CsgOperation.Intersection)
    w = rtc.World(objects=[rtc.Plane(), rtc.Csg(inner1, inner2, rtc.CsgOperation.Difference), rtc.Cube()],
                  lights=[rtc.PointLight((0, 5, -5), (1, 1, 1))])
    i = w.lower().check()
    assert i.n_prims == 6 and i.n_bvh_prims == 0


# ---- random trees (hypothesis): whatever nesting the caller builds, every leaf is lowered exactly once ------------------
from hypothesis import given, settings, strategies as st  # noqa: E402


def _ow_tree(draw, depth, counter):
    m = ow.Lambertian(ow.SolidColor((0.5, 0.5, 0.5)))
    kind = draw(st.integers(0, 5 if depth < 3 else 1))
    if kind == 0:
        counter[0] += 1
        c = tuple(draw(st.floats(-5, 5)) for _ in range(3))
        return ow.Sphere(ow.Center.Stationary(c), draw(st.floats(0.1, 2.0)), m)
    if kind == 1:
        counter[0] += 1
        return ow.Quad.new((draw(st.floats(-3, 3)), 0.0, 0.0), (1.0, 0.0, 0.0), (0.0, 1.0 + draw(st.floats(0, 2)), 0.0), m)
    if kind == 2:
        return _ow_tree(draw, depth + 1, counter).translate((draw(st.floats(-2, 2)), 0.5, -1.0))
    if kind == 3:
        return _ow_tree(draw, depth + 1, counter).rotate_y(draw(st.floats(-90, 90))).scale(draw(st.floats(0.5, 2.0)))
    kids = [_ow_tree(draw, depth + 1, counter) for _ in range(draw(st.integers(1, 4)))]
    return ow.Bvh.new(kids) if kind == 4 else ow.HittableList(kids)


@settings(max_examples=60, deadline=None)
@given(st.data())
def test_random_ow_trees_lower_every_leaf_once(data):
    counter = [0]
    world = _ow_tree(data.draw, 0, counter)
    info = ow.lower_world(world).check()
    assert info.n_prims + info.n_bvh_prims == counter[0]
    assert info.n_bvh_nodes == max(info.n_bvh_prims - 1, 0)


def _rtc_tree(draw, depth, counter, in_csg):
    kind = draw(st.integers(0, 7 if depth < 3 else 2))
    if kind <= 1:
        counter[0] += 1
        return [rtc.Sphere, rtc.Cube][kind]()
    if kind == 2:
        if in_csg:  # triangles under a Csg are the one unsupported construct
            counter[0] += 1
            return rtc.Cylinder(minimum=0.0, maximum=1.0, closed=True)
        counter[1] += 1
        return rtc.Triangle.flat([(0, 0, 0), (1, 0, 0), (0, 1 + draw(st.floats(0, 1)), 0)])
    if kind == 3:
        return rtc.Transformed.new(_rtc_tree(draw, depth + 1, counter, in_csg), T.translation(draw(st.floats(-2, 2)), 0.0, 1.0))
    if kind == 4:
        return rtc.Bounded.new(_rtc_tree(draw, depth + 1, counter, in_csg))
    if kind == 5:
        return rtc.Group.new([_rtc_tree(draw, depth + 1, counter, in_csg) for _ in range(draw(st.integers(1, 3)))])
    return rtc.Csg(_rtc_tree(draw, depth + 1, counter, True), _rtc_tree(draw, depth + 1, counter, True),
                   draw(st.sampled_from([rtc.CsgOperation.Union, rtc.CsgOperation.Intersection, rtc.CsgOperation.Difference])))


@settings(max_examples=60, deadline=None)
@given(st.data())
def test_random_rtc_trees_lower_every_leaf_once(data):
    counter = [0, 0]
    objs = [_rtc_tree(data.draw, 0, counter, False) for _ in range(data.draw(st.integers(1, 3)))]
    info = rtc.World(objects=objs, lights=[rtc.PointLight((0, 5, -5), (1, 1, 1))]).lower().check()
    assert (info.n_prims, info.n_bvh_prims) == (counter[0], counter[1])


def test_device_mesh_nodes_need_a_ctx_that_holds_a_mesh():
    """RL_RTC_MESH / RL_OW_MESH name the mesh rl_obj_parse left on a ctx; the host-only check has none"""
    from rendering_learning_b200 import RlError, ow, rtc
    sd = ow.lower_world([ow.DeviceMesh(None, ow.Lambertian(ow.SolidColor((0.5, 0.5, 0.5))))])
    with pytest.raises(RlError, match="parsed mesh"):
        sd.check()
    w = rtc.World(objects=[rtc.DeviceMesh(None)], lights=[])
    with pytest.raises(RlError, match="parsed mesh"):
        w.lower().check()
