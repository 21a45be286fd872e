"""N > 1 host logic on CPU: world_size-2 gloo processes pull from the dynamic tile queue, every job is
rendered exactly once, and the SUM gather reproduces the single-process framebuffer bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rendering_learning_b200 import dist as rd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_partial(job, shape):
    """what a rank would write for `job`: a deterministic function of (chunk, y, x)"""
    nc, H, W, _ = shape
    x0, y0, x1, y1, c0, c1 = job
    c, y, x = np.meshgrid(np.arange(c0, c1), np.arange(y0, y1), np.arange(x0, x1), indexing="ij")
    v = (np.sin(c * 12.9898 + y * 78.233 + x * 37.719) * 43758.5453) % 1.0
    return (c0, c1, y0, y1, x0, x1), np.stack([v, v * 0.5, v * 0.25], axis=-1).astype(np.float32)


def _worker(rank, world, port, shape, jobs, dynamic, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    partial = torch.zeros(shape, dtype=torch.float32)
    queue = (rd.TileQueue(rd.default_store(), "rl_q_test", len(jobs)) if dynamic
             else rd.StaticQueue(rank, world, len(jobs)))

    def launch(j):
        (c0, c1, y0, y1, x0, x1), v = _fake_partial(jobs[j], shape)
        partial[c0:c1, y0:y1, x0:x1] = torch.from_numpy(v)

    mine = rd.drain(queue, launch)
    dist.reduce(partial, dst=0, op=dist.ReduceOp.SUM)
    counts = [None] * world
    dist.all_gather_object(counts, mine)
    if rank == 0:
        q.put((partial.numpy().copy(), counts))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("dynamic", [True, False])
def test_two_rank_queue_and_gather(dynamic):
    shape = (3, 30, 40, 3)
    jobs = rd.make_jobs(40, 30, 3, rows_per_job=8)
    assert len(jobs) == 3 * 4 and jobs[0] == (0, 0, 40, 8, 0, 1) and jobs[-1] == (0, 24, 40, 30, 2, 3)
    ref = np.zeros(shape, np.float32)
    for j in jobs:
        (c0, c1, y0, y1, x0, x1), v = _fake_partial(j, shape)
        ref[c0:c1, y0:y1, x0:x1] = v
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, shape, jobs, dynamic, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, counts = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    allj = sorted(j for c in counts for j in c)
    assert allj == list(range(len(jobs)))  # every job exactly once
    assert np.array_equal(got, ref)  # x + 0 == x: the gather is exact
    if not dynamic:
        assert counts[0] == list(range(0, len(jobs), 2))


def test_job_grids():
    jobs = rd.jobs_for(1200, 675, 16, world_size=8)
    px = sum((j[2] - j[0]) * (j[3] - j[1]) * (j[5] - j[4]) for j in jobs)
    assert px == 1200 * 675 * 16 and 100 <= len(jobs) <= 400
    assert all(j[1] % 4 == 0 for j in jobs)
    jobs = rd.jobs_for(3840, 2160, 1, world_size=1)
    assert sum((j[2] - j[0]) * (j[3] - j[1]) for j in jobs) == 3840 * 2160
