"""Time the OW kernel on a workload: python tools/time_ow.py [C4|C5|test] [spp] [opt=value ...]   (rl_set_option names)"""
import hashlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rendering_learning_b200 import _abi as _A
_alt = [a for a in sys.argv if a.startswith('lib=')]
if _alt:  # an experiment build made by tools/build_alt.py
    sys.argv.remove(_alt[0]); _A.LIB_PATH = os.path.join(os.path.dirname(_A.LIB_PATH), f'librl_b200_{_alt[0][4:]}.so')
from rendering_learning_b200 import Context, ow, scenes
wl = sys.argv[1] if len(sys.argv) > 1 else "C4"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 50
opts = dict(a.split("=") for a in sys.argv[3:])
ctx = Context(0)
for k, v in opts.items():
    ctx.set_option(k, int(v))
if wl == "C4":
    world, params = scenes.ow_cover_world(), scenes.ow_cover_params(samples_per_pixel=spp)
elif wl == "C5full":  # the bench configuration: 3840 x 2160
    world, params = scenes.ow_cow_world(), scenes.ow_cow_params(image_width=3840, samples_per_pixel=spp)
elif wl == "C5":
    world, params = scenes.ow_cow_world(), scenes.ow_cow_params(image_width=1920, samples_per_pixel=spp)
else:
    world, params = scenes.ow_test_scene()
    params.samples_per_pixel = spp
ctx.scene_upload(ow.lower_world(world))
cam = params.abi()
ts = []
for i in range(3):
    img, st = ctx.render_ow(cam)
    ts.append(st.kernel_ms)
print(f"{wl} spp={spp} {opts}: best {min(ts[1:]):.3f} ms  mean_radiance {img.mean()/spp:.6f}  md5 {hashlib.md5(img.tobytes()).hexdigest()[:8]}", flush=True)
