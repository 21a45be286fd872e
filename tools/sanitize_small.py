"""Small runs of every kernel family; prints one md5 per case.  Run twice — with the release library and with
RL_B200_DEBUG=1 (the bounds-asserting build, every index checked against the scene's counts) — and compare: the debug
build must report no failed check (overflow == 0) and the same bits.  compute-sanitizer is closed on this pool; this is
the substitute (tests/test_gpu_debug_build.py drives it)."""
import gzip, hashlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rendering_learning_b200 import Context, ow, rtc, scenes
from rendering_learning_b200 import _abi as A

md5 = lambda a: hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()[:12]
ctx = Context(0)
print("lib", os.path.basename(A.LIB_PATH), flush=True)


def ow_case(name, world, p, **opts):
    for k, v in opts.items():
        ctx.set_option(k, v)
    ctx.scene_upload(ow.lower_world(world))
    img, st = ctx.render_ow(p.abi())
    assert st.overflow == 0, (name, st.overflow)
    print(name, md5(img), flush=True)
    for k in opts:
        ctx.set_option(k, 5 if k == "ow.variant" else 0)


w, p = scenes.ow_cover_world(), scenes.ow_cover_params(image_width=64, samples_per_pixel=70, max_depth=12)
ow_case("cover", w, p)
ow_case("cover_pooled", w, p, **{"ow.variant": 6})
ow_case("cow", scenes.ow_cow_world(), scenes.ow_cow_params(image_width=48, samples_per_pixel=8, max_depth=8))
w, p = scenes.ow_cornell_smoke()
p.image_width, p.samples_per_pixel, p.max_depth = 32, 8, 8
ow_case("smoke", w, p)
w, p = scenes.ow_final_scene(image_width=40, samples_per_pixel=8, max_depth=8)
ow_case("final_scene", w, p)
# trace mode of the production kernel
w, p = scenes.ow_test_scene()
ctx.scene_upload(ow.lower_world(w))
rng = np.random.default_rng(1)
o = rng.uniform(-2, 2, (4096, 3)).astype(np.float32); o[:, 1] = np.abs(o[:, 1]) + 0.2
d = rng.normal(size=(4096, 3)).astype(np.float32)
print("trace", md5(ctx.trace_batch(o, d, rng.uniform(0, 1, 4096))), flush=True)
for name, sc in (("csg", scenes.rtc_csg_scene(48, 32)), ("teapot", scenes.rtc_obj_scene(48, 32)), ("mirror", scenes.rtc_mirror_scene(48, 32))):
    ctx.scene_upload(sc.world.lower())
    img, st = ctx.render_rtc(sc.camera.abi(), 1)
    assert st.overflow == 0, (name, st.overflow)
    print(name, md5(img), flush=True)
# OBJ ingest + device mesh instancing
text = gzip.open(os.path.join(ROOT, "tests", "golden", "teapot-low.obj.gz"), "rb").read()
mesh = rtc.DeviceMesh.parse(text, ctx=ctx)
sc = scenes.rtc_obj_scene(48, 32, obj=mesh)
print("teapot_device_mesh", md5(sc.render(ctx=ctx).data), flush=True)
ctx.close()
print("done", flush=True)
