"""bench.py — the headline measurement of the B200-native ray loop.

    python bench.py --gpus N --steps K --warmup W [--workload C4|C5|C1|C2|C3] [--impl reference]

A "step" is one complete render of the workload.  Default workload = BASELINE.json's headline: the RTIOW
cover scene at 1200x675, 500 spp, depth 50 (config C4); the teapot (C3, 3840x2160) is measured beside it at
N = 1 and reported under "secondary".  For N > 1 the driver launches this file under torchrun (one rank per
GPU); the image is cut into tile x sample-chunk jobs pulled from a dynamic queue and the framebuffer is
gathered to rank 0 with NCCL.  Rank 0 prints ONE JSON line.

  value      whole-job Mrays/s with the scene resident in HBM, device-timed (CUDA events on the launching
             stream), max over ranks.  rays = every ray cast (camera + secondary + shadow), counted by device
             atomics in a separate instrumented pass with the same seed.
  e2e        the same metric through the public API (`Camera.render`): lowering, H2D scene upload, LBVH build,
             render, D2H of the framebuffer into pinned host memory, every step.
  roofline   algorithmic FP32 flops (SURVEY.md §8d constants x device counters) / kernel time vs the FP32
             peak measured live on this GPU; L2 and HBM fractions beside it.  bound = "fp32": no stage of this
             path is a dense contraction and the working set is L1/L2 resident.  `traffic` =
             dram__bytes_read + dram__bytes_write of the same kernel at the same config from the committed
             `ncu --set full` capture (profiles/ncu_<workload>.json, written by tools/ncu_traffic.py).
  cpu_baseline  the oracle (a C++ f64 restatement of the reference; the Rust reference cannot be built in
             this image) on all host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# algorithmic work per unit (definitions, fixed so numbers compare across commits)
FLOPS = {"node": 28, "sphere": 36, "tri": 45, "quad": 40, "rtc_prim": 50, "rtc_shade": 110, "ow_scatter": 45, "camera": 40}
BYTES_L2 = {"node": 64, "sphere": 32, "tri": 48, "quad": 64, "rtc_prim": 64, "rtc_shade": 48, "ow_scatter": 36}


def workloads():
    from rendering_learning_b200 import ow, scenes
    return {
        "C4": dict(kind="ow", name="RTIOW cover scene (bouncing_spheres), 1200x675, 500 spp, depth 50",
                   world=scenes.ow_cover_world, params=lambda: scenes.ow_cover_params()),
        "C5": dict(kind="ow", name="Cornell box + textured spot (cow), 3840x2160, 256 spp, depth 40",
                   world=scenes.ow_cow_world, params=lambda: scenes.ow_cow_params()),
        "C1": dict(kind="rtc", name="RTC three spheres on a plane, 1920x1080, 1 spp",
                   scene=lambda: scenes.rtc_three_spheres_scene(1920, 1080)),
        "C2": dict(kind="rtc", name="RTC mirror scene (reflect/refract, depth 5), 3840x2160",
                   scene=lambda: scenes.rtc_mirror_scene(3840, 2160)),
        "C3": dict(kind="rtc", name="RTC teapot-low.obj (240 triangles), Phong + shadows, 3840x2160",
                   scene=lambda: scenes.rtc_obj_scene(3840, 2160)),
    }


# ---- clocks -------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---- helpers --------------------------------------------------------------------------------------------------
def algorithmic(kind: str, st: dict, samples: int) -> tuple[float, float]:
    """(flops, L2-level bytes) of one step from the instrumented counters."""
    if kind == "ow":
        fl = (st["node_visits"] * FLOPS["node"] + st["prim_tests"] * FLOPS["sphere"] + st["tri_tests"] * FLOPS["tri"] +
              st["shades"] * FLOPS["ow_scatter"] + samples * FLOPS["camera"])
        by = (st["node_visits"] * BYTES_L2["node"] + st["prim_tests"] * BYTES_L2["sphere"] +
              st["tri_tests"] * BYTES_L2["tri"] + st["shades"] * BYTES_L2["ow_scatter"])
    else:
        fl = (st["node_visits"] * FLOPS["node"] + st["prim_tests"] * FLOPS["rtc_prim"] + st["tri_tests"] * FLOPS["tri"] +
              st["shades"] * FLOPS["rtc_shade"] + samples * FLOPS["camera"])
        by = (st["node_visits"] * BYTES_L2["node"] + st["prim_tests"] * BYTES_L2["rtc_prim"] +
              st["tri_tests"] * BYTES_L2["tri"] + st["shades"] * BYTES_L2["rtc_shade"])
    return float(fl), float(by)


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---- reference arm / cpu baseline: the oracle on the host cores ---------------------------------------------
def cpu_run(wl: dict, budget_s: float):
    """Time the CPU restatement on a bounded sample; returns (Mrays/s, samples/s, description, seconds)."""
    from oracle import oracle as orc
    from rendering_learning_b200 import ow
    if wl["kind"] == "ow":
        desc = ow.lower_world(wl["world"]())
        params = wl["params"]()
        full_spp = params.samples_per_pixel
        params.samples_per_pixel = 1
        t0 = time.perf_counter()
        orc.ow_render(desc, params.abi())
        t1 = time.perf_counter() - t0
        spp = int(max(2, min(full_spp, budget_s / max(t1, 1e-3))))
        params.samples_per_pixel = spp
        cam = params.abi()
        t0 = time.perf_counter()
        sums, rays = orc.ow_render(desc, cam)
        dt = time.perf_counter() - t0
        samples = sums.shape[0] * sums.shape[1] * spp
        what = (f"full {sums.shape[1]}x{sums.shape[0]} frame at {spp} spp of {full_spp} "
                f"(rays/s is spp-independent; scale time linearly in spp)")
        return rays / dt / 1e6, samples / dt, what, dt
    sc = wl["scene"]()
    desc = sc.world.lower()
    cam = sc.camera.abi()
    rows = cam.vsize
    full = (cam.hsize, cam.vsize)
    # bounded: shrink the frame (same scene, same aspect) until one render fits the budget
    from rendering_learning_b200 import rtc
    scale = 1
    while True:
        c2 = rtc.Camera.new(full[0] // scale, full[1] // scale, sc.camera.fov, sc.camera.transform).abi()
        t0 = time.perf_counter()
        orc.rtc_render(desc, c2, 1)
        dt = time.perf_counter() - t0
        if dt <= budget_s or scale >= 8:
            break
        scale *= 2
    rays = orc.rtc_camera_rays(c2, 1)
    # rays per pixel of the oracle = the GPU's instrumented count per pixel (same algorithm); report pixels/s too
    what = f"{c2.hsize}x{c2.vsize} frame (1/{scale} linear size of {full[0]}x{full[1]}), 1 spp"
    return None, c2.hsize * c2.vsize / dt, what, dt


def reference_arm(args, wl_key: str):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workloads()[wl_key]
    from oracle import oracle as orc
    orc.build()
    per_step = args.cpu_seconds or (8.0 if wl["kind"] == "ow" else 20.0)
    vals, what, secs, sps = [], "", [], []
    for i in range(args.warmup + args.steps):
        v, s, what, dt = cpu_run(wl, per_step)
        if i >= args.warmup:
            vals.append(v)
            secs.append(dt)
            sps.append(s)
        if i == 0 and dt * (args.warmup + args.steps) > 240:
            per_step = max(1.0, per_step / 2)
    cores = host_cores()
    if vals[0] is None:  # RTC: express as Mrays/s with the GPU-countable rays/pixel unavailable -> use primary rays
        value = float(np.mean(sps)) / 1e6
        unit = "Mpixels/s"
    else:
        value, unit = float(np.mean(vals)), "Mrays/s"
    line = {"metric": f"Mrays/s, {wl['name']}", "impl": "reference", "value": value, "unit": unit,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "sample": what},
            "samples_per_s": float(np.mean(sps)),
            "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": what},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU restatement (C++ f64, OpenMP over rows/columns like the reference's rayon loop) of the "
                    "Rust reference, which cannot be compiled in this image (no cargo/rustc)"}
    print(json.dumps(line), flush=True)


# ---- our arm --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C4", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (invalidates the headline)")
    ap.add_argument("--cpu-seconds", type=float, default=0.0, help="CPU budget per reference-arm step (default 8 s OW / 20 s RTC)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        return reference_arm(args, args.workload)

    # Libraries chat on stdout (NCCL prints its version there); the contract is ONE JSON line, so everything
    # else goes to stderr and the line is written to the saved descriptor at the end.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line: dict):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    import torch
    import torch.distributed as dist
    from rendering_learning_b200 import Context, ow, rtc
    from rendering_learning_b200 import dist as rd

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = Context(local)
    wl = workloads()[args.workload]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world_size == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world_size == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    stream = rd.current_stream_handle()
    if wl["kind"] == "ow":
        world = wl["world"]()
        params = wl["params"]()
        if args.spp:
            params.samples_per_pixel = args.spp
        desc = ow.lower_world(world)
        cam = params.abi()
        ctx.scene_upload(desc)
        W, H, nc = cam.image_width, ctx.ow_image_height(cam), ctx.ow_num_chunks(cam)
        partial = torch.zeros((nc, H, W, 4), dtype=torch.float32, device=dev)
        frame = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
        jobs = rd.jobs_for(W, H, nc, world_size) if world_size > 1 else [(0, 0, W, H, 0, nc)]
        samples = W * H * params.samples_per_pixel
        out_bytes = H * W * 3 * 4
        # instrumented pass (same seed): rays and unit counts of one step
        ctx.set_instrumented(True)
        torch.cuda.synchronize()
        st_i = ctx.render_ow_device(cam, 0, [(0, 0, W, H, 0, nc)], partial.data_ptr(), stream).as_dict() if rank == 0 else None
        ctx.set_instrumented(False)
        counter = [0]

        # RL_MGPU: "fused"  = NVLink-atomic queue + partial sums stored straight into rank 0's HBM (default)
        #          "device" = NVLink-atomic queue, NCCL sum-gather of the partial sums
        #          "store"  = c10d-store job queue, NCCL sum-gather
        mgpu = os.environ.get("RL_MGPU", "fused")
        shared = world_size > 1 and mgpu in ("device", "fused") and rd.setup_shared_queue(
            ctx, partial.numel() * 4 if mgpu == "fused" else 0)

        def step(i):
            counter[0] += 1
            if shared and mgpu == "fused":
                rd.render_ow_fused(ctx, cam, 0, frame, nc, H, W)
                return [0]
            if shared:
                rd.render_ow_shared_queue(ctx, cam, 0, partial, frame, nc, H, W)
                return [0]
            return rd.render_ow_distributed(ctx, cam, 0, jobs, partial, frame, f"s{i}_{counter[0]}")

        def kernel_probe():
            # the dominant kernel alone, whole frame, on this stream (roofline numerator / denominator)
            return ctx.render_ow_device(cam, 0, [(0, 0, W, H, 0, nc)], partial.data_ptr(), stream)

        def e2e_step():
            out = e2e_out
            t0 = time.perf_counter()
            d2 = ow.lower_world(world)  # the public call lowers the tree every time
            d2.freeze()
            t1 = time.perf_counter()
            ctx.scene_upload(d2)
            t2 = time.perf_counter()
            if world_size == 1:
                ctx.render_ow(cam, 0, out=out)
            else:
                step(10_000 + counter[0])
                if rank == 0:
                    e2e_pinned.copy_(frame, non_blocking=False)
            t3 = time.perf_counter()
            e2e_phases.append([(t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3])
            return desc.nbytes(), out_bytes
    else:
        scene = wl["scene"]()
        desc = scene.world.lower()
        cam = scene.camera.abi()
        ctx.scene_upload(desc)
        W, H = cam.hsize, cam.vsize
        frame = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
        jobs = rd.make_jobs(W, H, 1, rows_per_job=max(4, H // (world_size * 4))) if world_size > 1 else [(0, 0, W, H, 0, 1)]
        samples = W * H
        out_bytes = H * W * 3 * 4
        ctx.set_instrumented(True)
        torch.cuda.synchronize()
        st_i = ctx.render_rtc_device(cam, 1, [(0, 0, W, H, 0, 1)], frame.data_ptr(), stream).as_dict() if rank == 0 else None
        ctx.set_instrumented(False)
        counter = [0]
        shared, mgpu = False, "static"

        def step(i):
            counter[0] += 1
            return rd.render_rtc_distributed(ctx, cam, 1, jobs, frame, f"s{i}_{counter[0]}")

        def kernel_probe():
            return ctx.render_rtc_device(cam, 1, [(0, 0, W, H, 0, 1)], frame.data_ptr(), stream)

        def e2e_step():
            t0 = time.perf_counter()
            d2 = scene.world.lower()
            d2.freeze()
            t1 = time.perf_counter()
            ctx.scene_upload(d2)
            t2 = time.perf_counter()
            if world_size == 1:
                ctx.render_rtc(cam, 1, out=e2e_out)
            else:
                step(10_000 + counter[0])
                if rank == 0:
                    e2e_pinned.copy_(frame, non_blocking=False)
            t3 = time.perf_counter()
            e2e_phases.append([(t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3])
            return desc.nbytes(), out_bytes

    e2e_pinned = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()
    e2e_phases = []
    e2e_out = e2e_pinned.numpy()

    # ---- warm-up ----
    for i in range(args.warmup):
        step(-1 - i)
        ctx.synchronize()
    # ---- timed region: EXACTLY K steps, device-timed per step (the L2 flush between steps is outside) ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches = 0
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)  # evict L2 between timed iterations
        if world_size > 1:
            dist.barrier()
        evs[i][0].record()
        mine = step(i)
        evs[i][1].record()
        launches += len(mine) + (1 if (rank == 0 and wl["kind"] == "ow") else 0)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    ctx.synchronize()
    step_ms = [max_over_ranks(a.elapsed_time(b)) for a, b in evs]
    ms_per_step = float(np.mean(step_ms))
    clocks = sampler.stop() if rank == 0 else None
    total_launches = int(sum_over_ranks(float(launches)))

    # ---- dominant kernel alone (roofline) ----
    peaks = ctx.measure_peaks() if rank == 0 else None
    k_ms = []
    if rank == 0:
        for _ in range(3):
            flush.fill_(1)
            torch.cuda.synchronize()
            k_ms.append(kernel_probe().kernel_ms)
    barrier()

    # ---- e2e through the public API ----
    e2e_n = max(1, min(args.steps, 3))
    e2e_step()
    ctx.synchronize()
    barrier()
    e2e_phases.clear()
    t0 = time.perf_counter()
    for _ in range(e2e_n):
        h2d, d2h = e2e_step()
    ctx.synchronize()
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_n)

    if rank != 0:
        if world_size > 1:
            dist.destroy_process_group()
        return

    # the image of the last timed step: identical for every GPU count / tile schedule (compare across --gpus runs)
    import hashlib
    torch.cuda.synchronize()
    frame_md5 = hashlib.md5(frame.detach().cpu().numpy().tobytes()).hexdigest()
    rays = st_i["rays"]
    value = rays / (ms_per_step * 1e-3) / 1e6
    fl, by = algorithmic(wl["kind"], st_i, samples)
    kms = float(np.mean(k_ms))
    measured = {}
    try:
        measured = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = measured.get("hbm_gbs", 6650.0)
    ach_tf = fl / (kms * 1e-3) / 1e12
    hbm_alg = samples * 0 + out_bytes + (partial.numel() * 4 if wl["kind"] == "ow" else 0)
    prof = {}
    pj = os.path.join(ROOT, "profiles", f"ncu_{args.workload}.json")
    if os.path.exists(pj):
        prof = json.load(open(pj))
    roofline = {"bound": "fp32", "achieved": ach_tf, "peak": peaks["fp32_tflops"], "unit": "TFLOP/s",
                "frac": ach_tf / peaks["fp32_tflops"], "traffic": prof.get("dram_bytes_per_launch"),
                "kernel": "k_ow_render5" if wl["kind"] == "ow" else "k_rtc_render", "kernel_ms": kms,
                "traffic_source": prof.get("source"),
                "flops_per_launch": fl, "peak_source": "measured live (rl_measure_peaks: FMA chains, all SMs)",
                "l2": {"achieved": by / (kms * 1e-3) / 1e9, "peak": peaks["l2_gbs"], "unit": "GB/s",
                       "frac": by / (kms * 1e-3) / 1e9 / peaks["l2_gbs"], "bytes_per_launch": by},
                "hbm": {"achieved": hbm_alg / (kms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": hbm_alg / (kms * 1e-3) / 1e9 / hbm_peak, "bytes_per_launch": hbm_alg,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in measured else "fallback 6650"},
                "counters": {k: st_i[k] for k in ("rays", "node_visits", "prim_tests", "tri_tests", "shades")}}
    line = {"metric": f"Mrays/s (device-timed), {wl['name']}", "value": value, "unit": "Mrays/s",
            "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "image": [W, H], "l2": "flushed between timed steps (256 MiB write)",
                       "parallelism": (("one persistent launch per GPU; warps pop (pixel x sample-chunk) items from ONE counter in "
                                        "rank 0's HBM with system-scope atomics over NVLink (CUDA IPC); " +
                                        ("partial sums stored straight into rank 0's HBM over NVLink (fused gather)"
                                         if mgpu == "fused" else "NCCL sum-gather to rank 0")
                                        if (wl["kind"] == "ow" and shared)
                                        else f"{len(jobs)} tile x sample-chunk jobs from a c10d-store queue, NCCL sum-gather to rank 0") +
                                       f", {world_size} ranks") if world_size > 1 else "1 rank, persistent warps"},
            "samples_per_s": samples / (ms_per_step * 1e-3), "rays_per_step": rays, "samples_per_step": samples,
            "wall_s_timed_region": t_wall, "step_ms": step_ms, "frame_md5": frame_md5,
            "clocks": clocks, "gpu_launches": total_launches,
            "e2e": {"value": rays / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3, "steps": e2e_n,
                    "phases_ms_rank0": {"lower_tree": float(np.mean([p[0] for p in e2e_phases])),
                                        "scene_upload": float(np.mean([p[1] for p in e2e_phases])),
                                        "render_and_d2h": float(np.mean([p[2] for p in e2e_phases]))},
                    "path": "Camera.render: lower tree -> rl_scene_upload (flatten, H2D, LBVH build) -> render -> D2H (pinned)"},
            "roofline": roofline}

    if world_size == 1 and not args.no_secondary and args.workload == "C4":
        # the metric names the teapot beside the cover scene: measure C3 the same way
        sc3 = workloads()["C3"]["scene"]()
        ctx.scene_upload(sc3.world.lower())
        c3 = sc3.camera.abi()
        f3 = torch.zeros((c3.vsize, c3.hsize, 3), dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        ctx.set_instrumented(True)
        s3 = ctx.render_rtc_device(c3, 1, [(0, 0, c3.hsize, c3.vsize, 0, 1)], f3.data_ptr(), stream).as_dict()
        ctx.set_instrumented(False)
        t3 = []
        for i in range(args.warmup + args.steps):
            flush.fill_(i & 0xFF)
            torch.cuda.synchronize()
            t3.append(ctx.render_rtc_device(c3, 1, [(0, 0, c3.hsize, c3.vsize, 0, 1)], f3.data_ptr(), stream).kernel_ms)
        t3 = float(np.mean(t3[args.warmup:]))
        fl3, by3 = algorithmic("rtc", s3, c3.hsize * c3.vsize)
        line["secondary"] = {"workload": workloads()["C3"]["name"], "value": s3["rays"] / (t3 * 1e-3) / 1e6,
                             "unit": "Mrays/s", "ms_per_step": t3, "samples_per_s": c3.hsize * c3.vsize / (t3 * 1e-3),
                             "roofline": {"bound": "fp32", "achieved": fl3 / (t3 * 1e-3) / 1e12, "peak": peaks["fp32_tflops"],
                                          "unit": "TFLOP/s", "frac": fl3 / (t3 * 1e-3) / 1e12 / peaks["fp32_tflops"]}}

    if world_size == 1 and not args.no_cpu:
        v, sps, what, dt = cpu_run(wl, 15.0)
        line["cpu_baseline"] = {"value": v if v is not None else sps / 1e6, "unit": "Mrays/s" if v is not None else "Mpixels/s",
                                "cores": host_cores(), "kind": "port", "sample": what, "seconds": dt,
                                "samples_per_s": sps}
    emit(line)
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
