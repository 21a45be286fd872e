"""Where the end-to-end time of Camera.render goes (lower / upload / render + D2H) for C4 or C5."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rendering_learning_b200 import Context, ow, scenes
wl = sys.argv[1] if len(sys.argv) > 1 else "C5"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ctx = Context(0)
world, params = (scenes.ow_cover_world(), scenes.ow_cover_params()) if wl == "C4" else (scenes.ow_cow_world(), scenes.ow_cow_params())
if spp: params.samples_per_pixel = spp
cam = params.abi()
out = np.empty((ctx.ow_image_height(cam), cam.image_width, 3), np.float32)
for i in range(3):
    t0 = time.perf_counter(); d = ow.lower_world(world); d.freeze()
    t1 = time.perf_counter(); ctx.scene_upload(d)
    t2 = time.perf_counter(); _, st = ctx.render_ow(cam, 0, out=out)
    t3 = time.perf_counter()
    print(f"{wl} lower {1e3*(t1-t0):.1f} ms  upload {1e3*(t2-t1):.1f} ms (device part {st.upload_ms:.2f})  render+D2H {1e3*(t3-t2):.1f} ms (kernel {st.kernel_ms:.1f})", flush=True)
