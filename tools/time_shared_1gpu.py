"""Shared-queue mode on ONE GPU (the ctx pops its own exported counter with system-scope atomics and stores into its own
exported partial buffer) against the plain single-GPU launch: isolates what the multi-GPU MODE costs from what several
GPUs cost.  python tools/time_shared_1gpu.py [spp]"""
import sys
import torch
sys.path.insert(0, ".")
from rendering_learning_b200 import Context, ow, scenes
from rendering_learning_b200 import dist as rd
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 500
ctx = Context(0)
world, params = scenes.ow_cover_world(), scenes.ow_cover_params(samples_per_pixel=spp)
ctx.scene_upload(ow.lower_world(world))
cam = params.abi()
W, H, nc = cam.image_width, ctx.ow_image_height(cam), ctx.ow_num_chunks(cam)
partial = torch.zeros((nc, H, W, 4), dtype=torch.float32, device="cuda")
stream = rd.current_stream_handle()
jobs = [(0, 0, W, H, 0, nc)]
ctx.queue_export(); ctx.partial_export(nc * H * W * 16)
def timed(fn, n=4):
    ts = []
    for i in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); fn(i); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts[1:])
plain = timed(lambda i: ctx.render_ow_device(cam, 0, jobs, partial.data_ptr(), stream, sync=False))
def shared(i):
    ctx.queue_reset(stream, i & 1)
    ctx.render_ow_shared(cam, 0, jobs, 0, stream, i & 1)
sh = timed(shared)
def shared_local(i):
    ctx.queue_reset(stream, i & 1)
    ctx.render_ow_shared(cam, 0, jobs, partial.data_ptr(), stream, i & 1)
sl = timed(shared_local)
print(f"C4 {spp} spp: plain {plain:.3f} ms, shared mode (own counter + own exported partial) {sh:.3f} ms, shared counter + plain partial {sl:.3f} ms")
