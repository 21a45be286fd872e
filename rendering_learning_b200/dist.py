"""Multi-GPU sharding of the per-pixel ray loop: one process per GPU, a dynamic tile queue, and a
framebuffer gather over NVLink with NCCL (torch.distributed is the plumbing).

The path shards naturally (pixels and samples are independent): the scene + LBVH are
replicated on every GPU (<= 5 MB), the image is cut into jobs = pixel rectangles x sample-chunk ranges, and
ranks pull jobs from ONE shared counter (an atomic fetch-add served by the c10d store, so it works across
processes) because per-tile cost varies > 10x (sky vs glass).  There is no data-path collective during
rendering; the only exchange is the final gather: every (chunk, pixel) slot is written by exactly one
rank and is zero elsewhere, so a SUM reduce to rank 0 is exact (x + 0 == x) and the folded image is
bit-identical for any GPU count and any schedule.

`launch` callables keep this module testable on CPU with the gloo backend (tests/test_dist_cpu.py).
"""
from __future__ import annotations

from typing import Callable, Sequence

Job = tuple  # (x0, y0, x1, y1, chunk_begin, chunk_end)


def make_jobs(width: int, height: int, n_chunks: int, rows_per_job: int, chunks_per_job: int = 1) -> list[Job]:
    """Row bands x chunk groups, chunk-major so neighbouring jobs touch neighbouring memory."""
    rows_per_job = max(4, (rows_per_job + 3) // 4 * 4)  # keep the 8x4 micro-tiles whole
    jobs = []
    for c0 in range(0, n_chunks, chunks_per_job):
        for y0 in range(0, height, rows_per_job):
            jobs.append((0, y0, width, min(y0 + rows_per_job, height), c0, min(c0 + chunks_per_job, n_chunks)))
    return jobs


def jobs_for(width: int, height: int, n_chunks: int, world_size: int, jobs_per_rank: int = 16) -> list[Job]:
    """About `jobs_per_rank` jobs per GPU: fine enough to balance, coarse enough to amortise a launch."""
    want = max(1, world_size * jobs_per_rank)
    bands = max(1, -(-want // n_chunks))
    rows = max(4, -(-height // bands))
    return make_jobs(width, height, n_chunks, rows)


class TileQueue:
    """Dynamic job queue: `pop` is an atomic fetch-add on a counter every rank shares."""

    def __init__(self, store, key: str, n_jobs: int):
        self.store, self.key, self.n_jobs = store, key, n_jobs

    def pop(self):
        v = self.store.add(self.key, 1) - 1
        return v if v < self.n_jobs else None


class StaticQueue:
    """Round-robin assignment (used for the millisecond-scale RTC configs)."""

    def __init__(self, rank: int, world_size: int, n_jobs: int):
        self.it = iter(range(rank, n_jobs, world_size))

    def pop(self):
        return next(self.it, None)


def drain(queue, launch: Callable[[int], None]) -> list[int]:
    """Pull jobs until the queue is empty; `launch(j)` must be asynchronous so the next pop overlaps it."""
    mine = []
    while True:
        j = queue.pop()
        if j is None:
            return mine
        launch(j)
        mine.append(j)


def _rank_world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def current_stream_handle() -> int:
    """cudaStream_t of torch's current stream.  The C ABI reads 0 as "the ctx's own stream", so the legacy
    default stream is passed as cudaStreamLegacy (0x1) to keep the launch on the stream torch events see."""
    import torch
    if not torch.cuda.is_available():
        return 0  # CPU-side tests of the protocol (gloo): there is no stream
    h = torch.cuda.current_stream().cuda_stream
    return h if h else 1


def default_store():
    import torch.distributed as dist
    return dist.distributed_c10d._get_default_store()


def render_ow_distributed(ctx, cam, first_sample: int, jobs: Sequence[Job], partial, out, step_key: str,
                          store=None, static: bool = False):
    """One OW render sharded over the ranks of the default process group.

    partial: [n_chunks, H, W, 4] f32 device tensor (rgb + pad) on every rank; out: [H, W, 3] f32 (written on rank 0).
    Returns the job ids this rank rendered."""
    import torch
    import torch.distributed as dist

    rank, world = _rank_world()
    stream = current_stream_handle()
    partial.zero_()
    if static or world == 1:
        q = StaticQueue(rank, world, len(jobs))
    else:
        q = TileQueue(store or default_store(), f"rl_q_{step_key}", len(jobs))
    mine = drain(q, lambda j: ctx.render_ow_device(cam, first_sample, [jobs[j]], partial.data_ptr(), stream, sync=False))
    if world > 1:
        dist.reduce(partial, dst=0, op=dist.ReduceOp.SUM)  # NCCL over NVLink; zeros elsewhere => exact
    if rank == 0:
        ctx.ow_reduce_device(cam, partial.data_ptr(), out.data_ptr(), stream)
    return mine


def setup_shared_queue(ctx, partial_bytes: int = 0) -> bool:
    """Map rank 0's control block (and, with partial_bytes > 0, its RL_QUEUE_SLOTS partial-sum buffers of that size)
    into every rank with CUDA IPC.  The size travels WITH the handle, so a rank cannot launch a camera that does not
    fit the owner's buffer.  Returns False when there is a single rank."""
    import torch.distributed as dist
    rank, world = _rank_world()
    if world == 1:
        return False
    box = [None, None, 0]
    if rank == 0:
        box = [ctx.queue_export(), ctx.partial_export(partial_bytes) if partial_bytes else None, int(partial_bytes)]
        ctx.queue_reset(0, 0)  # slot 0 is ready for the first render; every render resets the OTHER slot for the next one
        ctx.synchronize()
    dist.broadcast_object_list(box, src=0)
    if rank != 0:
        ctx.queue_import(box[0])
        if box[1] is not None:
            ctx.partial_import(box[1], box[2])
    dist.barrier()
    _fused_state["render"] = 0
    return True


_fused_state = {"render": 0}


def render_ow_fused(ctx, cam, first_sample: int, out, n_chunks: int, height: int, width: int, events=None):
    """OW render, fully device-driven: every GPU's persistent warps pop items from rank 0's counter and store the
    finished partial sums straight into rank 0's buffer, both over NVLink peer memory.

    Renders alternate between two {counter, partial buffer} slots.  Rank 0 resets the slot of render i + 1 on its stream
    BEFORE its own kernel of render i, i.e. before it joins render i's closing rendezvous — which every rank must pass
    before it launches render i + 1.  So a rank launches without waiting for anybody (a rank whose host thread is late
    simply pops fewer items) and the only collective is ONE 4-byte NCCL all-reduce at the end of the render, after
    which rank 0 checks the completion counter and folds."""
    import torch.distributed as dist
    rank, world = _rank_world()
    stream = current_stream_handle()
    slot = _fused_state["render"] & 1
    _fused_state["render"] += 1
    jobs = [(0, 0, width, height, 0, n_chunks)]
    if rank == 0:
        ctx.queue_reset(stream, slot ^ 1)
    if events:
        events[0].record()
    ctx.render_ow_shared(cam, first_sample, jobs, 0, stream, slot)
    if events:
        events[1].record()
    dist.all_reduce(_token(out.device))  # every rank's kernel has finished: its stores have landed in rank 0's HBM
    if rank == 0:
        ctx.ow_reduce_shared(cam, slot, out.data_ptr(), stream)
        return slot
    return None


def check_fused_complete(ctx, cam, slot: int, n_chunks: int, height: int, width: int):
    """rank 0, after render_ow_fused: the completion counter must equal the number of work items — a rank that died
    (or a kernel that aborted) leaves it short instead of leaving silent zeros in the frame.  Synchronises the stream."""
    want = ctx.ow_job_items(cam, [(0, 0, width, height, 0, n_chunks)])
    got = ctx.queue_completed(current_stream_handle(), slot)
    if got != want:
        raise RuntimeError(f"multi-GPU render incomplete: {got} of {want} work items were stored")


def render_ow_shared_queue(ctx, cam, first_sample: int, partial, out, n_chunks: int, height: int, width: int, events=None):
    """OW render with the cross-GPU device queue: one persistent launch per GPU, warps of every GPU pop
    (pixel x sample-chunk) items from rank 0's counter over NVLink; NCCL sum-gathers the partial sums (the comparison
    point for the fused gather above)."""
    import torch
    import torch.distributed as dist
    rank, world = _rank_world()
    stream = current_stream_handle()
    slot = _fused_state["render"] & 1
    _fused_state["render"] += 1
    partial.zero_()
    if rank == 0:
        ctx.queue_reset(stream, slot ^ 1)
    if events:
        events[0].record()
    ctx.render_ow_shared(cam, first_sample, [(0, 0, width, height, 0, n_chunks)], partial.data_ptr(), stream, slot)
    if events:
        events[1].record()
    dist.reduce(partial, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        ctx.ow_reduce_device(cam, partial.data_ptr(), out.data_ptr(), stream)


_tokens = {}


def _token(device):
    import torch
    if device not in _tokens:
        _tokens[device] = torch.zeros(1, dtype=torch.float32, device=device)
    return _tokens[device]


def render_rtc_distributed(ctx, cam, aa: int, jobs: Sequence[Job], frame, step_key: str, store=None,
                           static: bool = True):
    """One RTC render sharded over the ranks; frame: [H, W, 3] f32 device tensor (complete on rank 0)."""
    import torch
    import torch.distributed as dist

    rank, world = _rank_world()
    stream = current_stream_handle()
    frame.zero_()
    if static or world == 1:
        q = StaticQueue(rank, world, len(jobs))
    else:
        q = TileQueue(store or default_store(), f"rl_q_{step_key}", len(jobs))
    mine = drain(q, lambda j: ctx.render_rtc_device(cam, aa, [jobs[j]], frame.data_ptr(), stream, sync=False))
    if world > 1:
        dist.reduce(frame, dst=0, op=dist.ReduceOp.SUM)
    return mine


# ---- the drop-in call for one-process-per-GPU programs ---------------------------------------------------------------
def camera_render_ow(camera, world, ctx, buffers):
    """`Camera::render(world)` (OW/src/camera.rs:122-124) as a COLLECTIVE: every rank of an SPMD program calls it with the
    same camera and world (each rank built the same scene), the ranks render it together, rank 0 gets the Canvas (the
    others None).  Per call: lower the object tree (unless `world` is already a SceneDesc), rl_scene_upload on every rank,
    one fused multi-GPU render, device -> pageable host copy and f64 Canvas on rank 0.
    `buffers`: an object with .frame ([H, W, 3] f32 device tensor), .nc, .H, .W whose rank 0 ctx exported the queue and
    partial-sum slots (setup_shared_queue)."""
    from . import ow
    from .desc import SceneDesc
    rank, world_size = _rank_world()
    sd = world if isinstance(world, SceneDesc) else ow.lower_world(world)
    ctx.scene_upload(sd)
    if getattr(buffers, "partial", None) is not None:  # the NCCL sum-gather comparison mode keeps its own partial buffer
        render_ow_shared_queue(ctx, camera.params.abi(), 0, buffers.partial, buffers.frame, buffers.nc, buffers.H, buffers.W)
        if rank != 0:
            return None
        ctx.synchronize()
    else:
        slot = render_ow_fused(ctx, camera.params.abi(), 0, buffers.frame, buffers.nc, buffers.H, buffers.W)
        if rank != 0:
            return None
        check_fused_complete(ctx, camera.params.abi(), slot, buffers.nc, buffers.H, buffers.W)
    sums = buffers.frame.cpu().numpy()  # pageable
    return ow.Canvas(camera.params.samples_per_pixel, buffers.W, buffers.H, sums)


def camera_render_rtc(camera, world, ctx, buffers):
    """`Camera::render(&world, &opts)` (RTC/src/scene/camera.rs:93) as a collective; see camera_render_ow."""
    from . import rtc
    from .desc import SceneDesc
    rank, world_size = _rank_world()
    sd = world if isinstance(world, SceneDesc) else world.lower()
    ctx.scene_upload(sd)
    buffers.counter += 1
    render_rtc_distributed(ctx, camera.abi(), 1, buffers.jobs, buffers.frame, f"e{buffers.counter}")
    if rank != 0:
        return None
    ctx.synchronize()
    rgb = buffers.frame.cpu().numpy()
    return rtc.Canvas(camera.hsize, camera.vsize, rgb)
