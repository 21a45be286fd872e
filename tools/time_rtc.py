"""Time the RTC kernel on a workload (C1 | C2 | C3); one upload (LBVH build) + 4 renders — the ncu target for k_rtc_render / lbvh."""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rendering_learning_b200 import _abi as _A
_alt = [a for a in sys.argv if a.startswith('lib=')]
if _alt:  # an experiment build made by tools/build_alt.py
    sys.argv.remove(_alt[0]); _A.LIB_PATH = os.path.join(os.path.dirname(_A.LIB_PATH), f'librl_b200_{_alt[0][4:]}.so')
from rendering_learning_b200 import Context, scenes
wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
sc = {"C1": lambda: scenes.rtc_three_spheres_scene(1920, 1080), "C2": lambda: scenes.rtc_mirror_scene(3840, 2160),
      "C3": lambda: scenes.rtc_obj_scene(3840, 2160), "CSG": lambda: scenes.rtc_csg_scene(3840, 2160)}[wl]()
ctx = Context(0)
ctx.scene_upload(sc.world.lower())
ts = []
for i in range(4):
    img, st = ctx.render_rtc(sc.camera.abi(), 1)
    ts.append(st.kernel_ms)
print(f"{wl}: best {min(ts[1:]):.3f} ms upload {st.upload_ms:.3f} ms md5 {hashlib.md5(img.tobytes()).hexdigest()[:8]}", flush=True)
