"""Kernel time against samples per pixel on ONE GPU (cover scene, full frame): the intercept is what a render costs
beyond its rays (ramp-up + drain tail) — the part that does not shrink when the work is split over more GPUs."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
from rendering_learning_b200 import _abi as _A
_alt = [a for a in sys.argv if a.startswith('lib=')]
if _alt:  # an experiment build made by tools/build_alt.py
    sys.argv.remove(_alt[0]); _A.LIB_PATH = os.path.join(os.path.dirname(_A.LIB_PATH), f'librl_b200_{_alt[0][4:]}.so')
from rendering_learning_b200 import Context, ow, scenes
ctx = Context(0)
world = scenes.ow_cover_world()
ctx.scene_upload(ow.lower_world(world))
import torch
rows = []
quick = 'quick' in sys.argv
for spp in (() if quick else (8, 16, 32, 63, 64, 125, 250, 500)):
    params = scenes.ow_cover_params(samples_per_pixel=spp)
    cam = params.abi()
    W, H, nc = cam.image_width, ctx.ow_image_height(cam), ctx.ow_num_chunks(cam)
    partial = torch.zeros((nc, H, W, 4), dtype=torch.float32, device="cuda")
    ts = [ctx.render_ow_device(cam, 0, [(0, 0, W, H, 0, nc)], partial.data_ptr()).kernel_ms for _ in range(4)]
    rows.append((spp, min(ts[1:]), nc))
    print(f"spp {spp:4d}  chunks {nc:3d}  kernel {min(ts[1:]):8.3f} ms", flush=True)
if rows:
    x = np.array([r[0] for r in rows], float); y = np.array([r[1] for r in rows])
    a, b = np.polyfit(x[3:], y[3:], 1)
    print(f"fit over spp >= 63: {a:.4f} ms/spp + {b:.3f} ms")

# one eighth of the frame (rows 3H/8 .. 4H/8): the share one GPU of eight gets; intercept of kernel time against spp
pts = []
for spp in (125, 250, 375, 500):
    params = scenes.ow_cover_params(samples_per_pixel=spp)
    cam = params.abi()
    W, H, nc = cam.image_width, ctx.ow_image_height(cam), ctx.ow_num_chunks(cam)
    partial = torch.zeros((nc, H, W, 4), dtype=torch.float32, device="cuda")
    y0 = (H // 8) * 3
    t = min(ctx.render_ow_device(cam, 0, [(0, y0, W, y0 + H // 8, 0, nc)], partial.data_ptr()).kernel_ms for _ in range(5))
    pts.append((spp, t))
    print(f"strip rows {y0}..{y0 + H // 8}: spp {spp:4d} chunks {nc:3d} kernel {t:8.3f} ms", flush=True)
a, b = np.polyfit([p[0] for p in pts], [p[1] for p in pts], 1)
print(f"strip fit: {a:.5f} ms/spp + {b:.3f} ms  (500 spp: {pts[-1][1]:.3f} ms, of which {b / pts[-1][1] * 100:.1f} % fixed)")
if _alt and _alt[0].startswith(("lib=timeline", "lib=tl")):  # the RL_TIMELINE experiment build reports warp lifetimes in the counters
    for spp, rows in ((500, None), (500, 8), (125, 8)):
        params = scenes.ow_cover_params(samples_per_pixel=spp)
        cam = params.abi()
        W, H, nc = cam.image_width, ctx.ow_image_height(cam), ctx.ow_num_chunks(cam)
        partial = torch.zeros((nc, H, W, 4), dtype=torch.float32, device="cuda")
        y0, y1 = (0, H) if rows is None else ((H // 8) * 3, (H // 8) * 4)
        for _ in range(3):
            st = ctx.render_ow_device(cam, 0, [(0, y0, W, y1, 0, nc)], partial.data_ptr())
        n = max(st.rays, 1)
        print(f"timeline spp {spp} rows {y0}..{y1}: kernel {st.kernel_ms:.3f} ms, warps {st.rays}, mean warp life {st.node_visits / n / 1e6:.3f} ms, "
              f"longest {st.shades / 1e6:.3f} ms, mean time to queue-dry {st.prim_tests / n / 1e6:.3f} ms, longest single-warp drain {st.tri_tests / 1e6:.3f} ms")
