"""Time the OW kernel on a workload (env RL_OW_KERNEL_V / RL_OW_SVC / RL_OW_LEAF select the variant)."""
import hashlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rendering_learning_b200 import Context, ow, scenes
wl = sys.argv[1] if len(sys.argv) > 1 else "C4"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 50
ctx = Context(0)
if wl == "C4":
    world, params = scenes.ow_cover_world(), scenes.ow_cover_params(samples_per_pixel=spp)
elif wl == "C5":
    world, params = scenes.ow_cow_world(), scenes.ow_cow_params(image_width=1920, samples_per_pixel=spp)
else:
    world, params = scenes.ow_test_scene()
    params.samples_per_pixel = spp
ctx.scene_upload(ow.lower_world(world))
cam = params.abi()
ts = []
for i in range(4):
    img, st = ctx.render_ow(cam)
    ts.append(st.kernel_ms)
tag = f"V={os.environ.get('RL_OW_KERNEL_V','2')} SVC={os.environ.get('RL_OW_SVC','-')} LEAF={os.environ.get('RL_OW_LEAF','-')}"
print(f"{wl} spp={spp} {tag}: best {min(ts[1:]):.3f} ms  mean_radiance {img.mean()/spp:.6f}  md5 {hashlib.md5(img.tobytes()).hexdigest()[:8]}", flush=True)
