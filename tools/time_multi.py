"""rl_create_multi: one ctx over all visible GPUs vs one GPU, C4 / C5 through the host-buffer entry points.
python tools/time_multi.py [n_gpus]"""
import hashlib, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from rendering_learning_b200 import Context, ow, scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
for name, world, params in (("C4", scenes.ow_cover_world(), scenes.ow_cover_params()),
                            ("C5_1920_64spp", scenes.ow_cow_world(), scenes.ow_cow_params(image_width=1920, samples_per_pixel=64))):
    desc = ow.lower_world(world)
    res = {}
    for ids in ([0], list(range(n))):
        ctx = Context(ids)
        ctx.scene_upload(desc)
        ts, wall = [], []
        for i in range(4):
            t0 = time.perf_counter()
            img, st = ctx.render_ow(params.abi())
            wall.append((time.perf_counter() - t0) * 1e3)
            ts.append(st.kernel_ms)
        res[len(ids)] = (min(ts[1:]), min(wall[1:]), hashlib.md5(img.tobytes()).hexdigest()[:10])
        ctx.close()
    a, b = res[1], res[n]
    print(f"{name}: 1 GPU {a[0]:.2f} ms device / {a[1]:.2f} ms call, {n} GPUs {b[0]:.2f} / {b[1]:.2f} ms, speed-up {a[0]/b[0]:.2f}x device {a[1]/b[1]:.2f}x call, md5 {a[2]} {b[2]} same={a[2]==b[2]}", flush=True)
