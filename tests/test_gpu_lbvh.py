"""GPU LBVH (Morton codes -> radix sort -> Karras hierarchy -> refit) is bit-exact against the host rebuild."""
import numpy as np
import pytest

from rendering_learning_b200 import ow, scenes

pytestmark = pytest.mark.gpu

KEYS = ("morton", "sorted_prim", "left", "right", "parent", "node_aabb")


def _bits(a):
    return a.view(np.uint32) if a.dtype == np.float32 else a


def _check(ctx, oracle, desc, n_expected=None):
    ctx.scene_upload(desc)
    g = ctx.lbvh_download()
    if n_expected is not None:
        assert g["n_prims"] == n_expected, g["n_prims"]
    h = oracle.lbvh_build(g["prim_aabb"])
    for k in KEYS:
        assert np.array_equal(_bits(g[k]), _bits(h[k])), k
    assert np.array_equal(_bits(np.concatenate([g["scene_lo"], g["scene_hi"]])), _bits(h["bounds"]))
    # structural sanity: sorted keys, a permutation, every node has exactly one parent
    assert (np.diff(g["morton"].astype(np.uint64)) >= 0).all() if g["n_prims"] > 1 else True
    assert sorted(g["sorted_prim"].tolist()) == list(range(g["n_prims"]))
    return g


def test_teapot(ctx, oracle):
    _check(ctx, oracle, scenes.rtc_obj_scene().world.lower(), 240)


def test_cover_scene(ctx, oracle):
    g = _check(ctx, oracle, ow.lower_world(scenes.ow_cover_world()))
    assert 400 < g["n_prims"] <= 488


def test_spot_in_cornell_box(ctx, oracle):
    # the six Cornell quads are large against the mesh: they go on the brute-force "big" list, outside the LBVH
    _check(ctx, oracle, ow.lower_world(scenes.ow_cow_world()), 5856)
    assert ctx.scene_info().n_prims == 6


@pytest.mark.parametrize("n", [1, 2, 3, 5, 33, 1025, 4097, 16383, 16384, 40000])  # >= 16384: the multi-CTA radix sort
def test_sizes_and_duplicates(ctx, oracle, n):
    """ragged sizes, identical centroids (key ties broken by position), one-primitive trees"""
    rng = np.random.default_rng(n)
    m = ow.Lambertian(ow.SolidColor((0.5, 0.5, 0.5)))
    cs = rng.uniform(-5, 5, size=(n, 3))
    cs[n // 3:] = cs[n // 3]  # many exact duplicates
    world = ow.HittableList([ow.Sphere(ow.Center.Stationary(tuple(c)), 0.25, m) for c in cs])
    _check(ctx, oracle, ow.lower_world(world), n)
