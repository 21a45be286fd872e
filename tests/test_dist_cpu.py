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


# ---- the fused multi-GPU protocol (alternating queue slots, one closing rendezvous) on CPU ----------------------------
class _FakeCtx:
    """records the C-ABI calls dist.render_ow_fused makes; `completed` plays the owner's completion counter"""

    def __init__(self, rank, log):
        self.rank, self.log, self.completed = rank, log, {0: 0, 1: 0}

    def queue_reset(self, stream, slot):
        self.log.append(("reset", slot))
        self.completed[slot] = 0

    def render_ow_shared(self, cam, first, jobs, d_partial, stream, slot):
        assert d_partial == 0  # fused gather: the library picks the owner's buffer of this slot
        self.log.append(("render", slot))
        self.completed[slot] += 100

    def ow_reduce_shared(self, cam, slot, out_ptr, stream):
        self.log.append(("fold", slot))

    def ow_job_items(self, cam, jobs):
        return 100

    def queue_completed(self, stream, slot):
        return self.completed[slot]


def _fused_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    log = []
    ctx = _FakeCtx(rank, log)
    rd._fused_state["render"] = 0
    out = torch.zeros(4)
    slots = []
    for _ in range(3):
        slot = rd.render_ow_fused(ctx, None, 0, out, 2, 8, 8)
        slots.append(slot)
        if rank == 0:
            rd.check_fused_complete(ctx, None, slot, 2, 8, 8)
    bad = None
    if rank == 0:
        ctx.completed[slots[-1]] = 99  # a rank that died: one item short
        try:
            rd.check_fused_complete(ctx, None, slots[-1], 2, 8, 8)
        except RuntimeError as e:
            bad = str(e)
    q.put((rank, log, slots, bad))
    dist.barrier()
    dist.destroy_process_group()


def test_fused_protocol_alternates_slots_and_detects_missing_items():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fused_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict((r, (log, slots, bad)) for r, log, slots, bad in (q.get(timeout=120) for _ in range(2)))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    log0, slots0, bad0 = res[0]
    log1, slots1, _ = res[1]
    # rank 0: before rendering on slot s it resets the OTHER slot (for the next render), then folds slot s
    assert log0 == [("reset", 1), ("render", 0), ("fold", 0), ("reset", 0), ("render", 1), ("fold", 1),
                    ("reset", 1), ("render", 0), ("fold", 0)]
    assert slots0 == [0, 1, 0]
    # the peer only launches: no reset, no fold, same slot sequence
    assert log1 == [("render", 0), ("render", 1), ("render", 0)] and slots1 == [None, None, None]
    assert bad0 is not None and "99 of 100" in bad0
