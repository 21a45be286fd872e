"""Time the OW render kernel over a grid of scheduling options (rl_set_option) on the B200 and check that every setting
renders the bit-identical frame.  python tools/sweep_ow.py [quick|full] > gpurun_out/sweep.jsonl"""
import hashlib
import itertools
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from rendering_learning_b200 import Context, ow, scenes  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
ctx = Context(0)
dev = torch.device("cuda", 0)
RESET = {"ow.variant": 5, "ow.slots": 0, "ow.minb": 0, "ow.ctas_per_sm": 0, "ow.svc_lo": 0, "ow.exit_min": 0, "ow.leaf_min": 0,
         "ow.svc_min": 0}


def setopts(o):
    for k, v in {**RESET, **o}.items():
        ctx.set_option(k, v)


def run(name, world, params, grid, reps=3):
    desc = ow.lower_world(world)
    ctx.scene_upload(desc)
    cam = params.abi()
    W, H, nc = cam.image_width, ctx.ow_image_height(cam), ctx.ow_num_chunks(cam)
    partial = torch.zeros((nc, H, W, 4), dtype=torch.float32, device=dev)
    frame = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
    ref = None
    for o in grid:
        setopts(o)
        try:
            ts = []
            for _ in range(reps):
                torch.cuda.synchronize()
                st = ctx.render_ow_device(cam, 0, [(0, 0, W, H, 0, nc)], partial.data_ptr())
                ts.append(st.kernel_ms)
            ctx.ow_reduce_device(cam, partial.data_ptr(), frame.data_ptr())
            ctx.synchronize()
            md5 = hashlib.md5(frame.cpu().numpy().tobytes()).hexdigest()
            if ref is None:
                ref = md5
            print(json.dumps({"scene": name, "opts": o, "ms_min": min(ts), "ms": ts, "md5": md5, "same_bits": md5 == ref}), flush=True)
        except Exception as e:  # keep sweeping: one bad setting must not hide the others
            print(json.dumps({"scene": name, "opts": o, "error": str(e)[:300]}), flush=True)
    setopts({})


v5 = [{}]
if mode == "quick":
    g6 = [{"ow.variant": 6}, {"ow.variant": 6, "ow.exit_min": 16}, {"ow.variant": 6, "ow.exit_min": 24, "ow.minb": 3}]
    g5 = [{"ow.svc_min": a, "ow.leaf_min": b} for a in (16, 20, 24) for b in (6, 8, 12)]
    g5 += [{"ow.minb": 3}, {"ow.minb": 4}]
    g6 = g5 + g6
else:
    g6 = [dict(zip(("ow.minb", "ow.slots", "ow.exit_min", "ow.svc_lo", "ow.leaf_min"), v))
          for v in itertools.product((3, 4), (256, 320, 384, 512), (6, 8, 12), (8, 16, 24), (6, 8, 12))]

t0 = time.time()
if mode == "c4":  # thresholds of the spheres-only instantiation only
    grid = [{}] + [{"ow.svc_min": a, "ow.leaf_min": b} for a in (20, 22, 24, 26, 28) for b in (6, 8, 10, 12)]
    run("C4_1200x675_500spp", scenes.ow_cover_world(), scenes.ow_cover_params(), grid, reps=3)
    print(json.dumps({"seconds": time.time() - t0}), flush=True)
    sys.exit(0)
if mode == "c5":  # thresholds of the triangle-scene instantiation only
    grid = [{}] + [{"ow.svc_min": a, "ow.leaf_min": b} for a in (12, 16, 20, 24, 28) for b in (4, 8, 12, 16)] + [{"ow.minb": 3}]
    run("C5_1920x1080_64spp", scenes.ow_cow_world(), scenes.ow_cow_params(image_width=1920, samples_per_pixel=64), grid, reps=3)
    print(json.dumps({"seconds": time.time() - t0}), flush=True)
    sys.exit(0)
w, p = scenes.ow_test_scene()
p.samples_per_pixel = 32
run("test_scene_300x168_32spp", w, p, v5 + g6[:6], reps=2)
run("C4_1200x675_500spp", scenes.ow_cover_world(), scenes.ow_cover_params(), v5 + g6)
run("C5_1920x1080_64spp", scenes.ow_cow_world(), scenes.ow_cow_params(image_width=1920, samples_per_pixel=64), v5 + g6, reps=2)
print(json.dumps({"seconds": time.time() - t0}), flush=True)
