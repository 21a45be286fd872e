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


def test_perlin_tables_and_noise(oracle):
    """perlin.rs: the tables Perlin::new draws, and noise / turb / Noise::value — host mirror vs oracle restatement"""
    world, _ = scenes.ow_perlin_spheres()
    noise = world[0].material.texture
    pn = noise.noise
    assert len(pn.randvec) == 256 and all(abs(sum(c * c for c in v) - 1.0) < 1e-12 for v in pn.randvec)
    for perm in (pn.perm_x, pn.perm_y, pn.perm_z):
        assert sorted(perm) == list(range(256))
        # `gen_range(0..i)` excludes i (perlin.rs:84): the shuffle is Sattolo's — one cycle, no fixed point
        assert all(perm[i] != i for i in range(256))
    assert pn.perm_x != pn.perm_y != pn.perm_z
    # lattice points: every weight vector is a lattice offset and the corner's own weight is 1 -> noise = c . 0 = 0
    assert pn.noise((3.0, -2.0, 7.0)) == 0.0
    rng = np.random.default_rng(3)
    pts = rng.uniform(-30, 30, size=(500, 3))
    vals = np.array([pn.noise(tuple(p)) for p in pts])
    assert np.abs(vals).max() <= 1.0 and vals.std() > 0.1  # "in the range [-1, 1]" (perlin.rs:38)
    desc = ow.lower_world(world)
    tex = desc.materials[0].texture
    uvp = np.concatenate([np.zeros((500, 2)), pts], axis=1)
    got = oracle.ow_tex_value(desc, tex, uvp)
    exp = np.array([noise.value(0.0, 0.0, tuple(p)) for p in pts])
    assert np.allclose(got, exp, rtol=0, atol=1e-12)
    assert (got >= 0).all() and (got <= 1).all() and got.std() > 0.05


def test_constant_medium_lowering_and_oracle_statistics(oracle):
    """constant_medium.rs: a unit-density slab of thickness L transmits exp(-L) of the straight-through rays"""
    black = ow.Isotropic(ow.SolidColor((0.0, 0.0, 0.0)))  # absorbs: a scattered path carries nothing
    wall = ow.Lambertian(ow.SolidColor((0.5, 0.5, 0.5)))
    box = ow.HittableList(scenes._ow_box((-50.0, -50.0, -1.0), (50.0, 50.0, 0.0), wall))
    world = [ow.ConstantMedium.new(box, 1.5, black)]
    desc = ow.lower_world(world)
    kinds = [n[0] for n in desc.nodes]
    assert kinds.count(A.RL_OW_CONSTANT_MEDIUM) == 1 and kinds.count(A.RL_OW_QUAD) == 6
    assert desc.materials[desc.nodes[kinds.index(A.RL_OW_CONSTANT_MEDIUM)][1]].kind == A.RL_MAT_OW_ISOTROPIC
    params = ow.CameraParams(aspect_ratio=1.0, image_width=8, samples_per_pixel=4000, max_depth=10, vfov=1.0,
                             lookfrom=(0.0, 0.0, 5.0), lookat=(0.0, 0.0, 0.0), vup=(0.0, 1.0, 0.0),
                             background=(1.0, 1.0, 1.0), seed=5)
    sums, _ = oracle.ow_render(desc, params.abi())
    mean = sums.mean() / 4000
    assert abs(mean - math.exp(-1.5)) < 0.01, mean
    # zero density: -1/0 = -inf in f64, the medium never scatters (no panic in the reference either)
    sums, _ = oracle.ow_render(ow.lower_world([ow.ConstantMedium.new(box, 0.0, black)]), params.abi())
    assert np.allclose(sums / 4000, 1.0)


def test_ow_checkpoint_bincode_layout():
    """the reference's checkpoint files (examples/common/mod.rs:24-56) are bincode 1.3.3 of `Canvas` (camera.rs:263-270)"""
    import struct
    data = np.arange(2 * 3 * 3, dtype=np.float64).reshape(2, 3, 3) * 0.25
    cv = ow.Canvas(7, 3, 2, data)
    blob = cv.to_bincode()
    assert len(blob) == 4 * 8 + 6 * 24
    assert struct.unpack_from("<QQQQ", blob) == (7, 3, 2, 6)  # samples, width, height, Vec<Color> length
    assert struct.unpack_from("<ddd", blob, 32 + 24 * 4) == (3.0, 3.25, 3.5)  # pixel (x=1, y=1), row-major
    back = ow.Canvas.from_bincode(blob)
    assert back == cv and back.samples == 7
    merged = back.merge(cv)  # resume semantics: sums add, samples add (camera.rs:273-291)
    assert merged.samples == 14 and np.array_equal(merged.data, 2 * data)
    with pytest.raises(ValueError):
        ow.Canvas.from_bincode(blob[:-1])


def test_canvases_hold_the_device_f32_frame_and_widen_on_access():
    """`Camera.render` hands the Canvas the f32 frame the device wrote; the reference's `Color` is f64, so every accessor
    sees f64 — widened once, on first access — and the 8-bit / PPM / bincode results are those of the widened frame."""
    from rendering_learning_b200 import ow, rtc
    rng = np.random.default_rng(3)
    f32 = rng.uniform(-0.2, 1.4, size=(6, 8, 3)).astype(np.float32)
    f32[0, 0] = (np.nan, 0.5, 2.0)
    c = rtc.Canvas(8, 6, f32)
    assert c._data.dtype == np.float32 and c._data is not None          # no copy made at construction
    wide = rtc.Canvas(8, 6, f32.astype(np.float64))
    assert np.array_equal(c.to_u8(), wide.to_u8()) and c.ppm() == wide.ppm()
    assert c.data.dtype == np.float64 and c._data.dtype == np.float64   # widened once, kept
    assert c.at(1, 2) == tuple(f32[2, 1].astype(np.float64))
    c.write((1, 2), (0.25, 0.5, 0.75))
    assert c.at(1, 2) == (0.25, 0.5, 0.75) and rtc.Canvas(2, 2).data.dtype == np.float64

    sums = np.abs(f32[1:]).astype(np.float32) * 10.0
    a = ow.Canvas(10, 8, 5, sums)
    b = ow.Canvas(10, 8, 5, sums.astype(np.float64))
    assert a._data.dtype == np.float32
    assert a == b and np.array_equal(a.to_u8(), b.to_u8()) and a.to_bincode() == b.to_bincode()
    m = a.merge(b)
    assert m.samples == 20 and m.data.dtype == np.float64 and np.array_equal(m.data, 2.0 * sums.astype(np.float64))
    assert ow.Canvas.from_bincode(a.to_bincode()) == b
