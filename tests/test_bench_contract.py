"""bench.py's reference arm runs without a GPU: it must print ONE JSON line carrying the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-seconds", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mrays/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("RTIOW cover scene") and d["data"] == "synthetic" and d["dtype"] == "f64"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "spp" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the driver divides the two arms' lines only when their `metric` strings are identical (round 1: they were not)
    sys.path.insert(0, ROOT)
    import bench
    assert d["metric"] == bench.metric_name(bench.workloads()["C4"])
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"metric": metric_name(wl)') == 2  # both arms build the string with the same function
    # OpenMP threads are set explicitly (torchrun exports OMP_NUM_THREADS=1) and reported
    assert cb["threads"] == cb["cores"]


def test_reference_arm_uses_every_core_under_torchrun_env():
    """torchrun exports OMP_NUM_THREADS=1; the reference arm must not inherit it (round 1: 1 thread reported as 32 cores)"""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-seconds", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    d1 = json.loads([l for l in out.stdout.splitlines() if l.strip()][0])
    cores = d1["cpu_baseline"]["cores"]
    if cores >= 4:
        import numpy as np  # one thread renders the 1-spp probe frame far slower than `cores` threads do
        from oracle import oracle as orc
        from rendering_learning_b200 import ow, scenes
        desc = ow.lower_world(scenes.ow_cover_world())
        p = scenes.ow_cover_params(image_width=300, samples_per_pixel=2)
        import time
        t0 = time.perf_counter(); _, rays = orc.ow_render(desc, p.abi(), threads=1); one = rays / (time.perf_counter() - t0) / 1e6
        assert d1["value"] > 2.0 * one, (d1["value"], one)


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
