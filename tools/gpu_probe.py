"""Print parity metrics and timings for every config (run on the GPU box; writes gpurun_out/probe.json)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402
from rendering_learning_b200 import Context, ow, rtc, scenes  # noqa: E402

OUT = {}


def u8_rtc(a):
    return rtc.Canvas(a.shape[1], a.shape[0], a).to_u8()


def rtc_case(ctx, name, scene, full=None):
    desc = scene.world.lower()
    t0 = time.time()
    ctx.scene_upload(desc)
    up = time.time() - t0
    cam = scene.camera.abi()
    img, st = ctx.render_rtc(cam, 1)
    t0 = time.time()
    ref = orc.rtc_render(desc, cam, 1)
    cpu = time.time() - t0
    d = np.abs(u8_rtc(img.astype(np.float64)) - u8_rtc(ref))
    rays = orc.rtc_camera_rays(cam, 1).astype(np.float32)
    node, t, second = orc.rtc_trace(desc, rays.astype(np.float64))
    hits = ctx.trace_batch(rays[:, 0:3], rays[:, 3:6])
    mism = hits["node"] != node
    both = (~mism) & (node >= 0)
    rel = np.abs(hits["t"][both].astype(np.float64) - t[both]) / np.maximum(np.abs(t[both]), 1e-30)
    ctx.set_instrumented(True)
    _, sti = ctx.render_rtc(cam, 1)
    ctx.set_instrumented(False)
    r = {"size": [cam.hsize, cam.vsize], "upload_s": up, "kernel_ms": st.kernel_ms, "cpu_s": cpu,
         "px_gt1": float((d > 1).any(axis=2).mean()), "px_ne": float((d > 0).any(axis=2).mean()),
         "max_diff": int(d.max()), "id_mismatch": float(mism.mean()), "n_rays": int(len(node)),
         "t_rel_max": float(rel.max()) if rel.size else 0.0, "t_rel_gt1e-4": float((rel > 1e-4).mean()) if rel.size else 0.0,
         "stats": sti.as_dict()}
    if full:
        sc2 = full()
        cam2 = sc2.camera.abi()
        ctx.scene_upload(sc2.world.lower())
        _, st2 = ctx.render_rtc(cam2, 1)
        _, st2 = ctx.render_rtc(cam2, 1)
        r["full_size"] = [cam2.hsize, cam2.vsize]
        r["full_kernel_ms"] = st2.kernel_ms
    OUT[name] = r
    print(name, json.dumps(r), flush=True)


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 10 * np.log10(255.0 ** 2 / max(mse, 1e-12))


def ow_case(ctx, name, world, params, spp, ref_mult=8):
    desc = ow.lower_world(world)
    t0 = time.time()
    ctx.scene_upload(desc)
    up = time.time() - t0
    params.samples_per_pixel = spp
    cam = params.abi()
    rays = orc.ow_camera_rays(cam).astype(np.float32)
    node, t, _ = orc.ow_trace(desc, rays.astype(np.float64))
    hits = ctx.trace_batch(rays[:, 0:3], rays[:, 3:6], rays[:, 6])
    mism = hits["node"] != node
    both = (~mism) & (node >= 0)
    rel = np.abs(hits["t"][both].astype(np.float64) - t[both]) / np.maximum(np.abs(t[both]), 1e-30)
    sums, st = ctx.render_ow(cam)
    ctx.set_instrumented(True)
    _, sti = ctx.render_ow(cam)
    ctx.set_instrumented(False)
    h = sums.shape[0]
    gpu = ow.Canvas(spp, cam.image_width, h, sums).to_u8()
    t0 = time.time()
    o1, rays1 = orc.ow_render(desc, cam)
    cpu = time.time() - t0
    p2 = params.abi()
    p2.seed = 12345
    o2, _ = orc.ow_render(desc, p2)
    p3 = params.abi()
    p3.seed = 777
    p3.samples_per_pixel = spp * ref_mult
    o3, _ = orc.ow_render(desc, p3)
    ref_hi = ow.Canvas(spp * ref_mult, cam.image_width, h, o3).to_u8()
    c1 = ow.Canvas(spp, cam.image_width, h, o1).to_u8()
    c2 = ow.Canvas(spp, cam.image_width, h, o2).to_u8()
    r = {"size": [cam.image_width, h], "spp": spp, "upload_s": up, "kernel_ms": st.kernel_ms, "cpu_s": cpu,
         "cpu_rays": int(rays1), "id_mismatch": float(mism.mean()), "n_rays": int(len(node)),
         "t_rel_max": float(rel.max()) if rel.size else 0.0, "t_rel_gt1e-4": float((rel > 1e-4).mean()) if rel.size else 0.0,
         "psnr_gpu_vs_hi": psnr(gpu, ref_hi), "psnr_cpu1_vs_hi": psnr(c1, ref_hi), "psnr_cpu2_vs_hi": psnr(c2, ref_hi),
         "mean_gpu": float(sums.mean() / spp), "mean_cpu": float(o1.mean() / spp), "mean_hi": float(o3.mean() / (spp * ref_mult)),
         "stats": sti.as_dict()}
    OUT[name] = r
    print(name, json.dumps(r), flush=True)


def lbvh_case(ctx, name, desc):
    ctx.scene_upload(desc)
    g = ctx.lbvh_download()
    h = orc.lbvh_build(g["prim_aabb"])
    res = {k: bool(np.array_equal(g[k].view(np.uint32) if g[k].dtype == np.float32 else g[k],
                                  h[k].view(np.uint32) if h[k].dtype == np.float32 else h[k]))
           for k in ("morton", "sorted_prim", "left", "right", "parent", "node_aabb")}
    res["n"] = int(g["n_prims"])
    OUT["lbvh_" + name] = res
    print("lbvh", name, res, flush=True)


def main():
    ctx = Context(0)
    print(ctx.device_info(), flush=True)
    rtc_case(ctx, "C1_three_spheres", scenes.rtc_three_spheres_scene(480, 270), lambda: scenes.rtc_three_spheres_scene(1920, 1080))
    rtc_case(ctx, "C2_mirror", scenes.rtc_mirror_scene(300, 200), lambda: scenes.rtc_mirror_scene(3840, 2160))
    rtc_case(ctx, "C3_teapot", scenes.rtc_obj_scene(300, 200), lambda: scenes.rtc_obj_scene(3840, 2160))
    lbvh_case(ctx, "teapot", scenes.rtc_obj_scene(300, 200).world.lower())
    cover = scenes.ow_cover_world()
    lbvh_case(ctx, "cover", ow.lower_world(cover))
    cow = scenes.ow_cow_world()
    lbvh_case(ctx, "cow", ow.lower_world(cow))
    w, p = scenes.ow_test_scene()
    ow_case(ctx, "OW_test", w, p, 16)
    ow_case(ctx, "C4_cover_small", cover, scenes.ow_cover_params(image_width=300, samples_per_pixel=16), 16)
    ow_case(ctx, "C5_cow_small", cow, scenes.ow_cow_params(image_width=160, samples_per_pixel=16), 16)
    # full-size timings
    ctx.scene_upload(ow.lower_world(cover))
    pc = scenes.ow_cover_params(samples_per_pixel=50)
    _, st = ctx.render_ow(pc.abi())
    OUT["C4_cover_1200x675_50spp_ms"] = st.kernel_ms
    print("C4 1200x675 50spp kernel ms", st.kernel_ms, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(OUT, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
