"""Kernel time against samples per pixel on ONE GPU (cover scene, full frame): the intercept is what a render costs
beyond its rays (ramp-up + drain tail) — the part that does not shrink when the work is split over more GPUs."""
import sys
import numpy as np
sys.path.insert(0, ".")
from rendering_learning_b200 import Context, ow, scenes
ctx = Context(0)
world = scenes.ow_cover_world()
ctx.scene_upload(ow.lower_world(world))
import torch
rows = []
for spp in (8, 16, 32, 63, 64, 125, 250, 500):
    params = scenes.ow_cover_params(samples_per_pixel=spp)
    cam = params.abi()
    W, H, nc = cam.image_width, ctx.ow_image_height(cam), ctx.ow_num_chunks(cam)
    partial = torch.zeros((nc, H, W, 4), dtype=torch.float32, device="cuda")
    ts = [ctx.render_ow_device(cam, 0, [(0, 0, W, H, 0, nc)], partial.data_ptr()).kernel_ms for _ in range(4)]
    rows.append((spp, min(ts[1:]), nc))
    print(f"spp {spp:4d}  chunks {nc:3d}  kernel {min(ts[1:]):8.3f} ms", flush=True)
x = np.array([r[0] for r in rows], float); y = np.array([r[1] for r in rows])
a, b = np.polyfit(x[3:], y[3:], 1)
print(f"fit over spp >= 63: {a:.4f} ms/spp + {b:.3f} ms")
