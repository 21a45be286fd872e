"""LBVH build at a size far beyond the BASELINE scenes: time + bit-exactness vs the host rebuild."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc
from rendering_learning_b200 import Context, ow
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
rng = np.random.default_rng(1)
m = ow.Lambertian(ow.SolidColor((0.5, 0.5, 0.5)))
cs = rng.uniform(-100, 100, size=(n, 3))
world = ow.HittableList([ow.Sphere(ow.Center.Stationary(tuple(c)), 0.3, m) for c in cs])
t0 = time.perf_counter(); desc = ow.lower_world(world); desc.freeze(); t1 = time.perf_counter()
ctx = Context(0)
ctx.scene_upload(desc); t2 = time.perf_counter()
ctx.scene_upload(desc)
g = ctx.lbvh_download()
h = orc.lbvh_build(g["prim_aabb"])
ok = all(np.array_equal(g[k].view(np.uint32) if g[k].dtype == np.float32 else g[k], h[k].view(np.uint32) if h[k].dtype == np.float32 else h[k])
         for k in ("morton", "sorted_prim", "left", "right", "parent", "node_aabb"))
params = ow.CameraParams(aspect_ratio=1.0, image_width=512, samples_per_pixel=8, max_depth=8, vfov=60.0,
                         lookfrom=(0.0, 0.0, 260.0), lookat=(0.0, 0.0, 0.0), vup=(0.0, 1.0, 0.0))
img, st = ctx.render_ow(params.abi())
print(f"n={n}: lower {1e3*(t1-t0):.0f} ms, first upload {1e3*(t2-t1):.0f} ms, device upload+LBVH {st.upload_ms:.2f} ms, bit-exact {ok}, "
      f"render 512x512x8 {st.kernel_ms:.2f} ms, overflow {st.overflow}, mean {img.mean()/8:.4f}", flush=True)
