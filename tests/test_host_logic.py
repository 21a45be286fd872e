"""Host-side mirror logic (no GPU): matrices, lowering, PPM encoders, OBJ parsers, error behaviour."""
import math

import numpy as np
import pytest

from rendering_learning_b200 import _abi as A
from rendering_learning_b200 import ow, rtc, scenes

T = rtc.transformation


def test_matrix_inverse_matches_reference_vectors(oracle):
    # RTC/src/math/matrix.rs tests: inverse of a known matrix (book ch. 3)
    m = [[-5.0, 2.0, 6.0, -8.0], [1.0, -5.0, 1.0, 8.0], [7.0, 7.0, -6.0, -7.0], [1.0, -3.0, 7.0, 4.0]]
    inv = rtc.invert(m)
    exp = [[0.21805, 0.45113, 0.24060, -0.04511], [-0.80827, -1.45677, -0.44361, 0.52068],
           [-0.07895, -0.22368, -0.05263, 0.19737], [-0.52256, -0.81391, -0.30075, 0.30639]]
    assert np.allclose(inv, exp, atol=1e-5)
    # the Python mirror, and the oracle's cofactor inverse are the same arithmetic
    assert np.array_equal(np.array(inv), oracle.rtc_invert([v for r in m for v in r]))
    with pytest.raises(ValueError, match="not invertible"):
        rtc.InvertibleMatrix.try_from([[0.0] * 4] * 4)


def test_view_transform_known_answers():
    # RTC/src/scene/transformation.rs tests (book ch. 7)
    vt = T.view_transform((1.0, 3.0, 2.0), (4.0, -2.0, 8.0), (1.0, 1.0, 0.0))
    exp = [[-0.50709, 0.50709, 0.67612, -2.36643], [0.76772, 0.60609, 0.12122, -2.82843],
           [-0.35857, 0.59761, -0.71714, 0.0], [0.0, 0.0, 0.0, 1.0]]
    assert np.allclose(vt, exp, atol=1e-5)
    assert np.allclose(T.view_transform((0, 0, 0), (0, 0, -1), (0, 1, 0)), rtc.identity())
    s = T.sequence([T.rotation_x(math.pi / 2), T.scaling(5, 5, 5), T.translation(10, 5, 7)])
    p = np.array(s) @ np.array([1.0, 0.0, 1.0, 1.0])
    assert np.allclose(p[:3], [15.0, 0.0, 7.0])


def test_rtc_canvas_ppm_format():
    # RTC/src/draw/canvas.rs:131-191
    c = rtc.Canvas(5, 3)
    c.write((0, 0), (1.5, 0.0, 0.0))
    c.write((2, 1), (0.0, 0.5, 0.0))
    c.write((4, 2), (-0.5, 0.0, 1.0))
    lines = c.ppm().split("\n")
    assert lines[:3] == ["P3", "5 3", "255"]
    assert lines[3] == "255 0 0 0 0 0 0 0 0 0 0 0 0 0 0"
    assert lines[4] == "0 0 0 0 0 0 0 128 0 0 0 0 0 0 0"
    assert lines[5] == "0 0 0 0 0 0 0 0 0 0 0 0 0 0 255"
    c = rtc.Canvas(10, 2, np.tile(np.array([1.0, 0.8, 0.6]), (2, 10, 1)))
    lines = c.ppm().split("\n")
    assert lines[3] == "255 204 153 255 204 153 255 204 153 255 204 153 255 204 153 255 204"
    assert lines[4] == "153 255 204 153 255 204 153 255 204 153 255 204 153"
    assert c.ppm().endswith("\n") and all(len(l) <= 70 for l in lines)
    assert c.at(10, 0) is None and c.write((0, 2), (0, 0, 0)) is None


def test_ow_colour_output():
    # OW/src/color.rs:86-109: (0, 0.5, 1) -> "0 188 255", clamping
    cv = ow.Canvas(1, 2, 1, np.array([[[0.0, 0.5, 1.0], [-1.0, 0.5, 2.0]]]))
    assert ow.output.output_ppm(cv) == "P3\n2 1\n255\n0 188 255\n0 188 255\n"
    a = ow.Canvas(2, 1, 1, [[[1.0, 2.0, 3.0]]])
    b = ow.Canvas(3, 1, 1, [[[0.5, 0.5, 0.5]]])
    m = a.merge(b)  # camera.rs:302-327
    assert m.samples == 5 and np.allclose(m.data, [[[1.5, 2.5, 3.5]]])
    assert np.allclose(m.pixel_data(), np.array([[[1.5, 2.5, 3.5]]]) * (1.0 / 5.0))


def test_camera_construction():
    c = rtc.Camera.new(200, 125, math.pi / 2, rtc.InvertibleMatrix.identity())
    assert abs(c.pixel_size - 0.01) < 1e-12  # camera.rs:153-163
    c = rtc.Camera.new(125, 200, math.pi / 2, rtc.InvertibleMatrix.identity())
    assert abs(c.pixel_size - 0.01) < 1e-12
    assert ow.Camera.new(ow.CameraParams(aspect_ratio=16.0 / 9.0, image_width=1200)).image_height == 675
    assert ow.Camera.new(ow.CameraParams(aspect_ratio=1000.0, image_width=10)).image_height == 1
    with pytest.raises(ValueError):
        ow.Camera.new(ow.CameraParams(lookfrom=(0, 0, 0), lookat=(0, 0, 0)))


def test_lowering_shapes():
    sd = scenes.rtc_mirror_scene().world.lower()
    kinds = [n[0] for n in sd.nodes]
    assert kinds.count(A.RL_RTC_SPHERE) == 4 and kinds.count(A.RL_RTC_PLANE) == 5
    assert kinds.count(A.RL_RTC_TRANSFORMED) == 10 and kinds.count(A.RL_RTC_GROUP) == 1
    assert len(sd.roots) == 9 and len(sd.lights) == 1 and len(sd.textures) == 2
    d = sd.freeze()
    assert d.n_nodes == len(sd.nodes) and d.flavor == A.RL_FLAVOR_RTC and sd.nbytes() > 0
    w = scenes.ow_cover_world()
    s2 = ow.lower_world(w)
    assert 400 < len(s2.nodes) <= 489 and s2.nodes[0][0] == A.RL_OW_BVH
    with pytest.raises(ValueError, match="without hittables"):
        ow.Bvh.new([])
    with pytest.raises(ValueError, match="parallel"):
        ow.Quad.new((0, 0, 0), (1, 0, 0), (2, 0, 0), ow.Dielectric(1.5))


def test_obj_parsers():
    # RTC/src/io/wavefront_obj.rs tests: gibberish ignored, fan triangulation, groups, normals
    o = rtc.WavefrontObj.parse("There was a young lady named Bright\nwho traveled much faster than light.\n")
    assert o.ignored == 2 and o.triangles() == []
    txt = "v -1 1 0\nv -1 0 0\nv 1 0 0\nv 1 1 0\nv 0 2 0\n\nf 1 2 3 4 5\n"
    o = rtc.WavefrontObj.parse(txt)
    tris = o.triangles()
    assert len(tris) == 3 and tris[2].points == [(-1.0, 1.0, 0.0), (1.0, 1.0, 0.0), (0.0, 2.0, 0.0)]
    txt = "v 0 1 0\nv -1 0 0\nv 1 0 0\nvn -1 0 0\nvn 1 0 0\nvn 0 1 0\nf 1//3 2//1 3//2\nf 1/0/3 2/102/1 3/14/2\n"
    o = rtc.WavefrontObj.parse(txt)
    t = o.triangles()
    assert len(t) == 2 and t[0].normals == [(0.0, 1.0, 0.0), (-1.0, 0.0, 0.0), (1.0, 0.0, 0.0)]
    txt = "v -1 1 0\nv -1 0 0\nv 1 0 0\nv 1 1 0\ng FirstGroup\nf 1 2 3\ng SecondGroup\nf 1 3 4\n"
    o = rtc.WavefrontObj.parse(txt)
    assert len(o.groups["FirstGroup"]) == 1 and len(o.groups["SecondGroup"]) == 1
    # OW parser keeps vt (OW/src/io/wavefront_obj.rs:248-306)
    txt = "v 0 1 0\nv -1 0 0\nv 1 0 0\nvt 0 0\nvt 1 0\nvt 0.5 1\nvn 0 0 1\nf 1/1/1 2/2/1 3/3/1\nf 1/1 2/2 3/3\nf 1 2 3\n"
    o2 = ow.WavefrontObj.parse(txt)
    tr = o2.tris()
    assert tr[0][1] == [(0.0, 0.0), (1.0, 0.0), (0.5, 1.0)] and tr[0][2] is not None
    assert tr[1][1] is not None and tr[1][2] is None and tr[2][1] is None


def test_fixture_meshes_present():
    m = scenes.load_mesh("teapot_low")
    assert m["tri_p"].shape == (240, 3, 3) and m["tri_smooth"].all()
    s = scenes.load_mesh("spot")
    assert s["tri_p"].shape == (5856, 3, 3) and s["has_uv"].all() and not s["has_n"].any()
    assert scenes.ow_spot_texture().shape == (1024, 1024, 3)
