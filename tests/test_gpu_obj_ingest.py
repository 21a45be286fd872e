"""OBJ ingest on the device.  Parity = the triangle arrays rl_obj_parse leaves in HBM are IDENTICAL (f64,
bit for bit) to what the host parsers — the mirrors of RTC/src/io/wavefront_obj.rs:22-187 and OW/src/io/wavefront_obj.rs:
32-252 — produce from the same text, on the reference's own meshes and on the records the reference's unit tests use; and
a scene that instances the device mesh renders the bit-identical image of the scene built from host-parsed triangles."""
import gzip
import os

import numpy as np
import pytest

from rendering_learning_b200 import RlError, ow, rtc, scenes
from rendering_learning_b200 import _abi as A

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def obj_text(name):
    return gzip.open(os.path.join(GOLD, name + ".gz"), "rb").read()


def host_arrays_rtc(text):
    o = rtc.WavefrontObj.parse(text)
    tris = o.triangles()
    P = np.array([[list(p) for p in t.points] for t in tris], np.float64).reshape(-1, 3, 3)
    N = np.array([[list(n) for n in (t.normals or [(0, 0, 0)] * 3)] for t in tris], np.float64).reshape(-1, 3, 3)
    F = np.array([1 if t.normals is not None else 0 for t in tris], np.uint8)
    return o, P, N, np.zeros((len(tris), 3, 2)), F


def host_arrays_ow(text):
    o = ow.WavefrontObj.parse(text)
    T = o.tris()
    P = np.array([t[0] for t in T], np.float64).reshape(-1, 3, 3)
    UV = np.array([t[1] if t[1] is not None else [(0, 0)] * 3 for t in T], np.float64).reshape(-1, 3, 2)
    N = np.array([t[2] if t[2] is not None else [(0, 0, 0)] * 3 for t in T], np.float64).reshape(-1, 3, 3)
    F = np.array([(1 if t[2] is not None else 0) | (2 if t[1] is not None else 0) for t in T], np.uint8)
    return o, P, N, UV, F


def check(ctx, text, flavor):
    host, P, N, UV, F = (host_arrays_rtc if flavor == A.RL_FLAVOR_RTC else host_arrays_ow)(text)
    info = ctx.obj_parse(text, flavor)
    assert info.n_triangles == len(P)
    assert info.n_vertices == len(host.vertices) and info.n_normals == len(host.normals)
    if flavor == A.RL_FLAVOR_OW:
        assert info.n_texcoords == len(host.texture_coords)
    assert info.ignored == host.ignored and info.n_groups >= 1
    if len(P):
        d = ctx.obj_download(info.n_triangles)
        assert np.array_equal(d["flags"], F)
        assert np.array_equal(d["tri_p"].view(np.uint64), P.view(np.uint64))  # bit for bit: the decimal -> f64 conversion is exact
        assert np.array_equal(d["tri_n"].view(np.uint64), N.view(np.uint64))
        if flavor == A.RL_FLAVOR_OW:
            assert np.array_equal(d["tri_uv"].view(np.uint64), UV.view(np.uint64))
        lo, hi = P.reshape(-1, 3).min(axis=0), P.reshape(-1, 3).max(axis=0)
        assert np.array_equal(np.array(list(info.bounds)), np.concatenate([lo, hi]))
    return info


def test_reference_meshes_parse_identically(ctx):
    info = check(ctx, obj_text("teapot-low.obj"), A.RL_FLAVOR_RTC)
    assert info.n_triangles == 240 and info.n_groups == 2  # 128 quads / triangles fanned out, `g Teapot001`
    info = check(ctx, obj_text("spot_triangulated.obj"), A.RL_FLAVOR_OW)
    assert info.n_triangles == 5856 and info.n_vertices == 2930 and info.n_texcoords == 3225
    check(ctx, obj_text("teapot-low.obj"), A.RL_FLAVOR_OW)      # the other parser's rules on the same text
    check(ctx, obj_text("spot_triangulated.obj"), A.RL_FLAVOR_RTC)


# the reference's own unit-test inputs (RTC wavefront_obj.rs tests: gibberish, vertex records, triangle faces, polygons,
# named groups, normals; OW: faces with texture coordinates) plus the record shapes the parsers distinguish
CASES = {
    "gibberish": "There was a young lady named Bright\nwho traveled much faster than light.\nShe set out one day\n"
                 "in a relative way,\nand came back the previous night.\n",
    "vertices": "v -1 1 0\nv -1.0000 0.5000 0.0000\nv 1 0 0\nv 1 1 0\n",
    "faces": "v -1 1 0\nv -1 0 0\nv 1 0 0\nv 1 1 0\n\nf 1 2 3\nf 1 3 4\n",
    "polygon": "v -1 1 0\nv -1 0 0\nv 1 0 0\nv 1 1 0\nv 0 2 0\n\nf 1 2 3 4 5\n",
    "groups": "v -1 1 0\nv -1 0 0\nv 1 0 0\nv 1 1 0\n\ng FirstGroup\nf 1 2 3\ng SecondGroup\nf 1 3 4\n",
    "normals": "v 0 1 0\nv -1 0 0\nv 1 0 0\n\nvn -1 0 0\nvn 1 0 0\nvn 0 1 0\n\nf 1//3 2//1 3//2\nf 1/0/3 2/102/1 3/14/2\n",
    "texcoords": "v 0 1 0\nv -1 0 0\nv 1 0 0\nvn -1 0 0\nvn 1 0 0\nvn 0 1 0\nvt 1 2\nvt 4 5\nvt 7 8 9\nvt 0.25\n"
                 "f 1/1/3 2/2/1 3/3/2\nf 1/4 2/1 3/2\nf 1/1 2 3/2\nf 1//1 2//2 3\n",
    "numbers": "v 1e2 -2.5E-3 .5\nv 5. 0 -0\nv 12345.678901234 0.000001 100000\nv 1 2\nv 1 2 3 4\nv 1 2 x\nvn 0.1 0.2 0.3\n"
               "v 0.1 0.2 0.3\nf 1 2 3\nf 1 2\nf 1 2 3/1/1/1\nf a b c\nf 3 2 1 \n",
    "whitespace_crlf": "v  -1   1\t0  \r\nv -1 0 0\r\nv 1 0 0\r\n\r\nf   1  2\t3  \r\n# comment\r\ng  padded name  \r\nf 3 2 1\r\nv 9 9 9",
    "group_replaced": "v 0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\nf 1 2 3\ng a\nf 1 2 4\ng b\nf 1 3 4\ng a\nf 2 3 4\nf 4 3 2\ng\nf 1 2 3\n",
    "no_trailing_newline": "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3",
    "empty": "",
    "only_blank": "\n\n\n",
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("flavor", [A.RL_FLAVOR_RTC, A.RL_FLAVOR_OW])
def test_record_shapes(ctx, name, flavor):
    if name == "normals" and flavor == A.RL_FLAVOR_OW:
        # the RTC unit test's `f 1/0/3 2/102/1 3/14/2`: RTC never reads the vt index; the OW parser does and
        # `read_texcoords[102 - 1]` is out of bounds (panic in the reference, IndexError in the host mirror)
        with pytest.raises(IndexError):
            ow.WavefrontObj.parse(CASES[name])
        with pytest.raises(RlError, match="out of bounds"):
            ctx.obj_parse(CASES[name].encode(), flavor)
        return
    check(ctx, CASES[name].encode(), flavor)


def test_numbers_beyond_the_fast_path_are_still_exact(ctx):
    """|exp10| > 22 or more than 15-16 digits: the big-integer path must round exactly like str::parse::<f64> (= Python float)"""
    toks = ["-4.33681e-19", "1e23", "8.5e-30", "123456789012345678", "0.000000000000000000000000000123456789012345678",
            "9007199254740993", "9007199254740992.5e3", "1.7976931348623157e40", "2.2250738585072014e-40",
            "4.9e-50", "3.141592653589793238", "0.1e-25", "72057594037927945"]
    lines = "".join(f"v {t} 0 1\n" for t in toks) + "f 1 2 3\n"
    info = ctx.obj_parse(lines.encode(), A.RL_FLAVOR_RTC)
    assert info.n_vertices == len(toks)
    text2 = "".join(f"v {t} 0 1\nv 0 {t} 1\nv 1 0 {t}\nf {3 * i + 1} {3 * i + 2} {3 * i + 3}\n" for i, t in enumerate(toks))
    info = ctx.obj_parse(text2.encode(), A.RL_FLAVOR_RTC)
    d = ctx.obj_download(info.n_triangles)
    for i, t in enumerate(toks):
        assert d["tri_p"][i, 0, 0] == float(t) and d["tri_p"][i, 1, 1] == float(t) and d["tri_p"][i, 2, 2] == float(t), t


def test_large_synthetic_mesh_and_scan_levels(ctx):
    """200 k triangles (8 MB of text, > one scan tile of tiles): counts and arrays against the host parser on a slice"""
    rng = np.random.default_rng(5)
    nv = 100_000
    V = np.round(rng.uniform(-50, 50, (nv, 3)), 5)
    F = rng.integers(1, nv + 1, (200_000, 3))
    lines = [f"v {a:.5f} {b:.5f} {c:.5f}" for a, b, c in V] + [f"f {a} {b} {c}" for a, b, c in F]
    text = ("\n".join(lines) + "\n").encode()
    info = ctx.obj_parse(text, A.RL_FLAVOR_OW)
    assert info.n_vertices == nv and info.n_triangles == 200_000 and info.ignored == 0
    d = ctx.obj_download(info.n_triangles)
    exp = np.array([[float(f"{x:.5f}") for x in V[i - 1]] for i in F[:2000].ravel()]).reshape(-1, 3, 3)
    assert np.array_equal(d["tri_p"][:2000], exp)
    last = np.array([[float(f"{x:.5f}") for x in V[i - 1]] for i in F[-1]]).reshape(3, 3)
    assert np.array_equal(d["tri_p"][-1], last)


def test_errors(ctx):
    with pytest.raises(RlError, match="out of bounds"):  # read_vertices[vi - 1] panics in the reference
        ctx.obj_parse(b"v 0 0 0\nv 1 0 0\nf 1 2 3\nv 0 1 0\n", A.RL_FLAVOR_RTC)
    with pytest.raises(RlError, match="out of bounds"):
        ctx.obj_parse(b"v 0 0 0\nv 1 0 0\nv 0 1 0\nf 0 1 2\n", A.RL_FLAVOR_OW)
    with pytest.raises(RlError, match="significant digits") as e:  # 19+ significant digits or |exp10| > 60: refused, not rounded
        ctx.obj_parse(b"v 0.12345678901234567890123 0 0\n", A.RL_FLAVOR_RTC)
    assert e.value.code == A.RL_E_UNSUPPORTED
    with pytest.raises(RlError, match="significant digits"):
        ctx.obj_parse(b"v 1e-200 0 0\n", A.RL_FLAVOR_RTC)
    # a scene that names a mesh the ctx does not hold is refused
    ctx.obj_parse(b"", A.RL_FLAVOR_OW)
    with pytest.raises(RlError):
        ow.Camera.new(ow.CameraParams(image_width=8)).render(ow.DeviceMesh(None, ow.Lambertian(ow.SolidColor((0.5, 0.5, 0.5)))), ctx=ctx)


def test_rtc_teapot_scene_from_device_mesh_is_bit_identical(ctx, oracle):
    """test_obj_scene (RTC/tests/ray_tracer.rs:242-275) with the teapot parsed on the GPU vs lowered triangle by triangle"""
    host_scene = scenes.rtc_obj_scene(300, 200)
    ref = host_scene.render(ctx=ctx)
    mesh = rtc.DeviceMesh.parse(obj_text("teapot-low.obj"), ctx=ctx)
    dev_scene = scenes.rtc_obj_scene(300, 200, obj=mesh)
    got = dev_scene.render(ctx=ctx)
    assert np.array_equal(ref.data, got.data)
    # and against the oracle / golden like any other RTC scene
    gold = np.load(os.path.join(GOLD, "rtc_obj.npz"))["pixels"]
    d = np.abs(got.to_u8().astype(np.int64) - gold.astype(np.int64))
    assert (d > 1).any(axis=2).mean() <= 1e-3
    info = ctx.scene_info()
    assert info.n_bvh_prims == 240


def test_ow_cow_scene_from_device_mesh_is_bit_identical(ctx):
    """examples/cow.rs with spot parsed on the GPU: same hits, same frame as the host-lowered mesh"""
    params = scenes.ow_cow_params(image_width=200, samples_per_pixel=8)
    host_world = scenes.ow_cow_world()
    a = ow.Camera.new(params).render(host_world, ctx=ctx)
    cow_surface = ow.Lambertian(ow.Image(scenes.ow_spot_texture()))
    mesh = ow.DeviceMesh.parse(obj_text("spot_triangulated.obj"), cow_surface, ctx=ctx)
    dev_world = scenes.ow_cow_world(cow=mesh)
    b = ow.Camera.new(params).render(dev_world, ctx=ctx)
    assert np.array_equal(a.data, b.data)
    assert ctx.scene_info().n_bvh_prims == 5856
