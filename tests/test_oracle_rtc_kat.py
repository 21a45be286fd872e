"""Known-answer vectors from the reference's own unit tests, replayed on the oracle (tier 1).
Each test cites the reference test it restates (paths under ray-tracer-challenge/src/scene/)."""
import math

import numpy as np
import pytest

from rendering_learning_b200 import rtc

T = rtc.transformation
SQ2 = math.sqrt(2.0)


def world(objs, lights=None, **kw):
    if lights is None:
        lights = [rtc.PointLight((-10.0, 10.0, -10.0), (1.0, 1.0, 1.0))]
    return rtc.World(objects=objs, lights=lights, **kw).lower()


def basic_spheres():
    # world.rs:180-197
    s1 = rtc.Transformed.new(rtc.Sphere(rtc.Material(surface=(0.8, 1.0, 0.6), diffuse=0.7, specular=0.2)), rtc.identity())
    s2 = rtc.Transformed.new(rtc.Sphere(), T.scaling(0.5, 0.5, 0.5))
    return [s1, s2]


def glass(ior=1.52):  # sphere::glass_sphere (object/sphere.rs:75-83)
    return rtc.Sphere(rtc.Material(transparency=1.0, refractive_index=ior))


def ts(oracle, obj, o, d):
    return list(oracle.rtc_intersect(world([obj]), list(o) + list(d))[0])


def test_sphere_roots(oracle):  # object/sphere.rs:98-146
    s = rtc.Sphere()
    assert ts(oracle, s, (0, 0, -5), (0, 0, 1)) == [4.0, 6.0]
    assert ts(oracle, s, (0, 1, -5), (0, 0, 1)) == [5.0, 5.0]
    assert ts(oracle, s, (0, 2, -5), (0, 0, 1)) == []
    assert ts(oracle, s, (0, 0, 0), (0, 0, 1)) == [-1.0, 1.0]
    assert ts(oracle, s, (0, 0, 5), (0, 0, 1)) == [-6.0, -4.0]
    # scaled / translated (transformed.rs tests)
    assert ts(oracle, rtc.Transformed.new(rtc.Sphere(), T.scaling(2, 2, 2)), (0, 0, -5), (0, 0, 1)) == [3.0, 7.0]
    assert ts(oracle, rtc.Transformed.new(rtc.Sphere(), T.translation(5, 0, 0)), (0, 0, -5), (0, 0, 1)) == []


def test_plane(oracle):  # object/plane.rs:57-122
    p = rtc.Plane()
    assert ts(oracle, p, (0, 10, 0), (0, 0, 1)) == []
    assert ts(oracle, p, (0, 0, 0), (0, 0, 1)) == []
    assert ts(oracle, p, (0, 1, 0), (0, -1, 0)) == [1.0]
    assert ts(oracle, p, (0, -1, 0), (0, 1, 0)) == [1.0]


CUBE = [((5, 0.5, 0), (-1, 0, 0), [4, 6]), ((-5, 0.5, 0), (1, 0, 0), [4, 6]), ((0.5, 5, 0), (0, -1, 0), [4, 6]),
        ((0.5, -5, 0), (0, 1, 0), [4, 6]), ((0.5, 0, 5), (0, 0, -1), [4, 6]), ((0.5, 0, -5), (0, 0, 1), [4, 6]),
        ((0, 0.5, 0), (0, 0, 1), [-1, 1]),
        ((-2, 0, 0), (0.2673, 0.5345, 0.8018), []), ((0, -2, 0), (0.8018, 0.2673, 0.5345), []),
        ((0, 0, -2), (0.5345, 0.8018, 0.2673), []), ((2, 0, 2), (0, 0, -1), []), ((0, 2, 2), (0, -1, 0), []),
        ((2, 2, 0), (-1, 0, 0), [])]


@pytest.mark.parametrize("o,d,exp", CUBE)
def test_cube_table(oracle, o, d, exp):  # object/cube.rs:109-126
    assert ts(oracle, rtc.Cube(), o, d) == [float(e) for e in exp]


def test_cube_normals(oracle):  # object/cube.rs:165-174 (via a ray that lands on the point)
    for p, n in [((1, 0.5, -0.8), (1, 0, 0)), ((-1, -0.2, 0.9), (-1, 0, 0)), ((-0.4, 1, -0.1), (0, 1, 0)),
                 ((0.3, -1, -0.7), (0, -1, 0)), ((-0.6, 0.3, 1), (0, 0, 1)), ((0.4, 0.4, -1), (0, 0, -1))]:
        o = tuple(3.0 * np.array(n) + np.array(p) - np.array(n) * np.abs(np.array(p) * np.array(n)).sum() + np.array(n))
        d = tuple(-np.array(n, dtype=float))
        t, _, normals, _ = oracle.rtc_intersect(world([rtc.Cube()]), list(o) + list(d))
        assert np.allclose(normals[0], n)


def test_cylinder_tables(oracle):  # object/cylinder.rs:179-184, 267-272, 312-316
    c = rtc.Cylinder()
    n = lambda v: tuple(np.array(v, float) / np.linalg.norm(v))
    assert ts(oracle, c, (1, 0, 0), n((0, 1, 0))) == []
    assert ts(oracle, c, (0, 0, 0), n((0, 1, 0))) == []
    assert ts(oracle, c, (0, 0, -5), n((1, 1, 1))) == []
    assert ts(oracle, c, (1, 0, -5), n((0, 0, 1))) == [5.0, 5.0]
    assert ts(oracle, c, (0, 0, -5), n((0, 0, 1))) == [4.0, 6.0]
    got = ts(oracle, c, (0.5, 0, -5), n((0.1, 1, 1)))
    assert np.allclose(got, [6.80798191702732, 7.088723439378861], rtol=0, atol=1e-12)
    tc = rtc.Cylinder(minimum=1.0, maximum=2.0)
    for o, d, cnt in [((0, 1.5, 0), (0.1, 1, 0), 0), ((0, 3, -5), (0, 0, 1), 0), ((0, 0, -5), (0, 0, 1), 0),
                      ((0, 2, -5), (0, 0, 1), 0), ((0, 1, -5), (0, 0, 1), 0), ((0, 1.5, -2), (0, 0, 1), 2)]:
        assert len(ts(oracle, tc, o, n(d))) == cnt
    cc = rtc.Cylinder(minimum=1.0, maximum=2.0, closed=True)
    for o, d in [((0, 3, 0), (0, -1, 0)), ((0, 3, -2), (0, -1, 2)), ((0, 4, -2), (0, -1, 1)),
                 ((0, 0, -2), (0, 1, 2)), ((0, -1, -2), (0, 1, 1))]:
        assert len(ts(oracle, cc, o, n(d))) == 2


def test_cone_tables(oracle):  # object/cone.rs:184-186, 224-226
    c = rtc.Cone()
    n = lambda v: tuple(np.array(v, float) / np.linalg.norm(v))
    assert np.allclose(ts(oracle, c, (0, 1e-6, -5), n((0, 0, 1))), [4.999999000844085, 5.000000999155915], atol=1e-9)
    assert np.allclose(ts(oracle, c, (0, 0, -5), n((1, 1, 1))), [8.660254037844386] * 2, atol=1e-6)
    assert np.allclose(ts(oracle, c, (1, 1, -5), n((-0.5, -1, 1))), [4.550055679356349, 49.449944320643645], atol=1e-9)
    cc = rtc.Cone(minimum=-0.5, maximum=0.5, closed=True)
    for o, d, cnt in [((0, 0, -5), (0, 1, 0), 0), ((0, 0, -0.25), (0, 1, 1), 2), ((0, 0, -0.25), (0, 1, 0), 4)]:
        assert len(ts(oracle, cc, o, n(d))) == cnt


def test_triangle(oracle):  # object/triangle.rs:169-251
    t = rtc.Triangle.flat([(0, 1, 0), (-1, 0, 0), (1, 0, 0)])
    assert ts(oracle, t, (0, -1, -2), (0, 1, 0)) == []
    assert ts(oracle, t, (1, 1, -2), (0, 0, 1)) == []
    assert ts(oracle, t, (-1, 1, -2), (0, 0, 1)) == []
    assert ts(oracle, t, (0, -1, -2), (0, 0, 1)) == []
    assert ts(oracle, t, (0, 0.5, -2), (0, 0, 1)) == [2.0]
    st = rtc.Triangle.smooth([((0, 1, 0), (0, 1, 0)), ((-1, 0, 0), (-1, 0, 0)), ((1, 0, 0), (1, 0, 0))])
    _, _, normals, _ = oracle.rtc_intersect(world([st]), [-0.2, 0.3, -2, 0, 0, 1])
    assert np.allclose(normals[0], (-0.5547, 0.83205, 0.0), atol=1e-5)


def test_group_ordering_and_transform(oracle):  # object/group.rs:83-103, transformed.rs:147-167
    s1 = rtc.Sphere()
    s2 = rtc.Transformed.new(rtc.Sphere(), T.translation(0, 0, -3))
    s3 = rtc.Transformed.new(rtc.Sphere(), T.translation(5, 0, 0))
    assert ts(oracle, rtc.Group.new([s1, s2, s3]), (0, 0, -5), (0, 0, 1)) == [1.0, 3.0, 4.0, 6.0]
    g = rtc.Transformed.new(rtc.Group.new([rtc.Transformed.new(rtc.Sphere(), T.translation(5, 0, 0))]), T.scaling(2, 2, 2))
    assert len(ts(oracle, g, (10, 0, -10), (0, 0, 1))) == 2
    # normal on a transformed sphere: n_world = normalize(inv^T n_local) (transformed.rs:39-51, 147-167),
    # checked against independent linear algebra by hitting the world point from outside
    m = T.sequence([T.rotation_z(math.pi / 5), T.scaling(1, 0.5, 1)])
    tr = rtc.Transformed.new(rtc.Sphere(), m)
    q = np.array([0.0, SQ2 / 2, -SQ2 / 2])
    M = np.array(m)
    pw = (M @ np.append(q, 1.0))[:3]
    n = (np.linalg.inv(M).T @ np.append(q, 0.0))[:3]
    n /= np.linalg.norm(n)
    tsv, _, normals, _ = oracle.rtc_intersect(world([tr]), list(pw + 3 * n) + list(-n))
    assert abs(tsv[0] - 3.0) < 1e-9 and np.allclose(normals[0], n, atol=1e-9)
    tr2 = rtc.Transformed.new(rtc.Sphere(), T.translation(0, 1, 0))
    pw = np.array([0.0, 1.70711, -0.70711])
    n = np.array([0.0, 0.70711, -0.70711])
    _, _, normals, _ = oracle.rtc_intersect(world([tr2]), list(pw + 3 * n) + list(-n))
    assert np.allclose(normals[0], n, atol=1e-4)


def test_hit_rules(oracle):  # intersect.rs:211-263 via two concentric / offset spheres
    a = rtc.Sphere()
    w = world([a])
    p = oracle.rtc_prepare(w, [0, 0, -5, 0, 0, 1])
    assert p["t"] == 4.0 and not p["inside"]
    assert np.allclose(p["point"], (0, 0, -1)) and np.allclose(p["eye_v"], (0, 0, -1)) and np.allclose(p["normal_v"], (0, 0, -1))
    p = oracle.rtc_prepare(w, [0, 0, 0, 0, 0, 1])  # inside: normal flipped (intersect.rs:306-322)
    assert p["t"] == 1.0 and p["inside"] and np.allclose(p["normal_v"], (0, 0, -1))
    # over / under point offsets (intersect.rs:325-377)
    w2 = world([rtc.Transformed.new(rtc.Sphere(), T.translation(0, 0, 1))])
    p = oracle.rtc_prepare(w2, [0, 0, -5, 0, 0, 1])
    assert p["over_point"][2] < -1e-5 / 2 and p["point"][2] > p["over_point"][2]
    assert p["under_point"][2] > 1e-5 / 2 - 1e-9 - 0.0 or p["under_point"][2] > p["point"][2]
    # reflection vector (intersect.rs:354-365)
    w3 = world([rtc.Plane()])
    p = oracle.rtc_prepare(w3, [0, 1, -1, 0, -SQ2 / 2, SQ2 / 2])
    assert np.allclose(p["reflect_v"], (0, SQ2 / 2, SQ2 / 2))
    # all hits behind -> no hit
    node, t, _ = oracle.rtc_trace(w, np.array([[0, 0, 5, 0, 0, 1.0]]))
    assert node[0] == -1


@pytest.mark.parametrize("index,n1,n2", [(0, 1.0, 1.5), (1, 1.5, 2.0), (2, 2.0, 2.5), (3, 2.5, 2.5), (4, 2.5, 1.5), (5, 1.5, 1.0)])
def test_n1_n2_table(oracle, index, n1, n2):  # intersect.rs:380-456
    a = rtc.Transformed.new(glass(1.5), T.scaling(2, 2, 2))
    b = rtc.Transformed.new(glass(2.0), T.translation(0, 0, -0.25))
    c = rtc.Transformed.new(glass(2.5), T.translation(0, 0, 0.25))
    w = world([a, b, c])
    tsv = list(oracle.rtc_intersect(w, [0, 0, -4, 0, 0, 1])[0])
    assert tsv == [2.0, 2.75, 3.25, 4.75, 5.25, 6.0]
    p = oracle.rtc_prepare(w, [0, 0, -4, 0, 0, 1], index)
    assert (p["n1"], p["n2"]) == (n1, n2)


def test_schlick(oracle):  # intersect.rs:462-505
    w = world([glass()])
    assert oracle.rtc_prepare(w, [0, 0, SQ2 / 2, 0, 1, 0], 1)["schlick"] == 1.0
    assert abs(oracle.rtc_prepare(w, [0, 0, 0, 0, 1, 0], 1)["schlick"] - 0.04) < 1e-2
    assert abs(oracle.rtc_prepare(w, [0, 0.99, -2, 0, 0, 1], 0)["schlick"] - 0.49018) < 1e-5


def test_lighting(oracle):  # material.rs:136-273
    w = world([rtc.Sphere()], lights=[rtc.PointLight((0, 0, -10), (1, 1, 1))])
    L = lambda eye, normal, att=1.0, light=0, wd=w: oracle.rtc_lighting(wd, 0, light, (0, 0, 0), (1, 1, 1), eye, normal, att)
    assert np.allclose(L((0, 0, -1), (0, 0, -1)), 1.9)
    assert np.allclose(L((0, SQ2 / 2, -SQ2 / 2), (0, 0, -1)), 1.0)
    w45 = world([rtc.Sphere()], lights=[rtc.PointLight((0, 10, -10), (1, 1, 1))])
    assert np.allclose(L((0, 0, -1), (0, 0, -1), wd=w45), 0.7364, atol=1e-4)
    assert np.allclose(L((0, -SQ2 / 2, -SQ2 / 2), (0, 0, -1), wd=w45), 1.6364, atol=1e-4)
    wb = world([rtc.Sphere()], lights=[rtc.PointLight((0, 0, 10), (1, 1, 1))])
    assert np.allclose(L((0, 0, -1), (0, 0, -1), wd=wb), 0.1)
    assert np.allclose(L((0, 0, -1), (0, 0, -1), att=0.0), 0.1)


def test_world_colours(oracle):  # world.rs:240-287, 344-350, 376-423
    w = world(basic_spheres())
    assert np.allclose(oracle.rtc_color_at(w, [0, 0, -5, 0, 0, 1]), (0.38066, 0.47583, 0.2855), atol=1e-5)
    assert np.allclose(oracle.rtc_color_at(w, [0, 0, -5, 0, 1, 0]), (0, 0, 0))
    wi = world(basic_spheres(), lights=[rtc.PointLight((0, 0.25, 0), (1, 1, 1))])
    assert np.allclose(oracle.rtc_color_at(wi, [0, 0, 0, 0, 0, 1]), 0.90498, atol=1e-5)
    wn = world(basic_spheres(), lights=[], void_color=(0.25, 0.5, 0.75))
    assert np.allclose(oracle.rtc_color_at(wn, [0, 0, -5, 0, 0, 1]), (0.25, 0.5, 0.75))  # shade_hit -> None
    for p, exp in [((0, 10, 0), 1.0), ((10, -10, 10), 0.0), ((-20, 20, -20), 1.0), ((-2, 2, -2), 1.0)]:
        assert oracle.rtc_shadow(w, p) == exp


def test_reflection_and_refraction_colours(oracle):  # world.rs:455-514, 639-748
    plane = rtc.Transformed.new(rtc.Plane(rtc.Material(reflectivity=0.5)), T.translation(0, -1, 0))
    w = world(basic_spheres() + [plane])
    ray = [0, 0, -3, 0, -SQ2 / 2, SQ2 / 2]
    assert np.allclose(oracle.rtc_color_at(w, ray), (0.87675, 0.92434, 0.82917), atol=1e-4)
    floor = rtc.Transformed.new(rtc.Plane(rtc.Material(transparency=0.5, refractive_index=1.5)), T.translation(0, -1, 0))
    ball = rtc.Transformed.new(rtc.Sphere(rtc.Material(surface=(1, 0, 0), ambient=0.5)), T.translation(0, -3.5, -0.5))
    w = world(basic_spheres() + [floor, ball])
    assert np.allclose(oracle.rtc_color_at(w, ray), (1.12546, 0.68642, 0.68642), atol=1e-5)
    floor2 = rtc.Transformed.new(rtc.Plane(rtc.Material(transparency=0.5, refractive_index=1.5, reflectivity=0.5)),
                                 T.translation(0, -1, 0))
    ball2 = rtc.Transformed.new(rtc.Sphere(rtc.Material(surface=(1, 0, 0), ambient=0.5)), T.translation(0, -3.5, -0.5))
    w = world(basic_spheres() + [floor2, ball2])
    assert np.allclose(oracle.rtc_color_at(w, ray), (1.115, 0.69643, 0.69243), atol=1e-5)
    # mutually reflective surfaces terminate (world.rs:517-548)
    lower = rtc.Transformed.new(rtc.Plane(rtc.Material(reflectivity=1.0)), T.translation(0, -1, 0))
    upper = rtc.Transformed.new(rtc.Plane(rtc.Material(reflectivity=1.0)), T.translation(0, 1, 0))
    w = world([lower, upper], lights=[rtc.PointLight((0, 0, 0), (1, 1, 1))])
    assert np.isfinite(oracle.rtc_color_at(w, [0, 0, 0, 0, 1, 0])).all()


def test_camera_rays_and_render(oracle):  # camera.rs:176-273
    cam = rtc.Camera.new(201, 101, math.pi / 2, rtc.InvertibleMatrix.identity())
    rays = oracle.rtc_camera_rays(cam.abi(), 1).reshape(101, 201, 6)
    assert np.allclose(rays[50, 100], (0, 0, 0, 0, 0, -1))
    assert np.allclose(rays[0, 0], (0, 0, 0, 0.66519, 0.33259, -0.66851), atol=1e-5)
    cam2 = rtc.Camera.new(201, 101, math.pi / 2, rtc.matmul(T.rotation_y(math.pi / 4), T.translation(0, -2, 5)))
    rays = oracle.rtc_camera_rays(cam2.abi(), 1).reshape(101, 201, 6)
    assert np.allclose(rays[50, 100], (0, 2, -5, SQ2 / 2, 0, -SQ2 / 2), atol=1e-9)
    cam3 = rtc.Camera.new(11, 11, math.pi / 2, T.view_transform((0, 0, -5), (0, 0, 0), (0, 1, 0)))
    img = oracle.rtc_render(world(basic_spheres()), cam3.abi(), 1)
    assert np.allclose(img[5, 5], (0.38066, 0.47583, 0.2855), atol=1e-5)
    # AA sub-sample order: nx-major (camera.rs:227-254)
    r2 = oracle.rtc_camera_rays(rtc.Camera.default(2, 2, math.pi / 2).abi(), 2).reshape(2, 2, 4, 6)
    assert r2[0, 0, 0, 3] > r2[0, 0, 2, 3] and r2[0, 0, 0, 4] > r2[0, 0, 1, 4]
