"""Multi-GPU sharding of the DrawMesh path across the GPUs of one box, one
process per GPU (torch.distributed is the plumbing).

Two ways the path shards:

1. **Frame sharding** (animation batches, examples/animate.go): the mesh is
   replicated, frame k is rendered by rank k mod N.  No collective.

2. **Sort-last by triangle range** (large meshes): rank r draws triangles
   [r*T/N, (r+1)*T/N) into its own full-frame buffers; the buffers are
   depth-composited by a min-reduction over packed keys
   ``(depth32 << 32 | R<<24 | G<<16 | B<<8 | A) ^ 2^63`` stored as int64
   (fgl_composite_pack / fgl_composite_unpack).  Only valid for the
   order-independent state: ReadDepth and WriteDepth on, DepthBias 0, opaque
   output.  Differences against a single-GPU render are confined to depth ties
   (smaller colour wins instead of the later triangle) and to the 32-bit depth
   quantisation; tests report the count.

Three ways to run the reduction, all collective (every rank calls once per frame):

* ``NcclComposite`` -- inside the library (fgl_comm_init / fgl_composite): stripe
  reduce-scatter of the keys with ncclMin on ncclUint64, then a gather to the
  presenting rank or an all-gather.  A Go host needs nothing else.
* ``PeerGroup`` -- inside the library (fgl_peer_*): exact float64 depth, sparse (only
  strips a rank has drawn), P2P loads/stores in one kernel per rank, device-side flags.
* ``composite_min`` -- ``torch.distributed.all_reduce(MIN)`` on the biased int64
  keys: the plain baseline (gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple


def triangle_range(ntriangles: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [first, first+count) of rank `rank`."""
    first = rank * ntriangles // world
    last = (rank + 1) * ntriangles // world
    return first, last - first


def frame_shard(nframes: int, rank: int, world: int) -> List[int]:
    """Frames k with k % world == rank."""
    return list(range(rank, nframes, world))


def sort_last_valid(ctx) -> bool:
    """The packed-key composite is only defined for the order-independent state."""
    return bool(ctx.ReadDepth and ctx.WriteDepth and ctx.DepthBias == 0)


def composite_min(keys, group=None):
    """In-place min-reduction of a packed-key tensor (int64) over all ranks."""
    import torch
    import torch.distributed as dist
    assert keys.dtype == torch.int64
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(keys, op=dist.ReduceOp.MIN, group=group)
    return keys


def triangle_blocks(ntriangles: int, rank: int, world: int, block: int = 4096):
    """Indices of the triangles of rank `rank` when blocks of `block` consecutive triangles are dealt
    round-robin: the same total work as contiguous ranges, but every rank gets a share of every part of the
    mesh (contiguous latitude bands leave the ranks that hold the far side of a closed surface idle).  The
    composite does not care which rank drew a triangle (order-independent state)."""
    import numpy as np
    idx = np.arange(ntriangles, dtype=np.int64)
    return idx[(idx // block) % world == rank]


def rebalance_ranges(bounds, times, damping: float = 0.8):
    """Contiguous triangle ranges of equal measured COST instead of equal count.  ``bounds`` are the N+1 range
    boundaries of the last frame, ``times[r]`` the time rank r took to draw [bounds[r], bounds[r+1]).  The cost
    density is taken as constant inside each old range; the new boundaries cut the cumulative cost into N equal
    parts (moved ``damping`` of the way, which keeps the iteration from oscillating).  A frame's cost is part
    per-triangle (geometry) and part per-pixel (ranges that face the camera cover more pixels), so equal counts
    leave the ranks up to 2x apart; a few frames of feedback bring them within a few percent."""
    n = len(times)
    assert len(bounds) == n + 1
    total = float(sum(times))
    if total <= 0:
        return list(bounds)
    new = [bounds[0]]
    r, done = 0, 0.0        # walking through the old ranges: `done` = cost of the ranges before r
    for k in range(1, n):
        target = total * k / n
        while r < n - 1 and done + times[r] < target:
            done += times[r]
            r += 1
        width = bounds[r + 1] - bounds[r]
        frac = 0.0 if times[r] <= 0 else (target - done) / times[r]
        cut = bounds[r] + frac * width
        old = bounds[k]
        cut = old + damping * (cut - old)
        new.append(int(min(max(round(cut), new[-1]), bounds[-1])))
    new.append(bounds[-1])
    return new


class NcclComposite:
    """fgl_comm_init / fgl_composite: the packed-key min-reduce with NCCL inside the library.  The unique id is
    created on rank 0 and broadcast with torch.distributed (any transport would do)."""

    def __init__(self, ctx, rank: int, world: int, group=None):
        import ctypes as C
        import torch.distributed as dist
        from .context import COMM_ID_BYTES, _check, capi
        self.ctx, self.rank, self.world = ctx, rank, world
        ident = [None]
        if rank == 0:
            buf = C.create_string_buffer(COMM_ID_BYTES)
            _check(capi().fgl_comm_unique_id(buf))
            ident[0] = buf.raw
        if world > 1:
            dist.broadcast_object_list(ident, src=0, group=group)
        self.handle = C.c_void_p()
        _check(capi().fgl_comm_init(ctx._h, world, rank, ident[0], C.byref(self.handle)), ctx._h)

    def composite(self, root: int = 0):
        """Enqueue pack -> reduce-scatter(min) -> gather to `root` (all-gather if root < 0) -> unpack."""
        from .context import _check, capi
        _check(capi().fgl_composite(self.ctx._h, self.handle, int(root)), self.ctx._h)

    def stage_times(self):
        """{pack_ms, reduce_scatter_ms, gather_ms, unpack_ms} per composite issued while profiling was on."""
        import ctypes as C
        from .context import _check, capi
        ms, n = (C.c_float * 4)(), C.c_uint32(0)
        _check(capi().fgl_comm_stage_times(self.ctx._h, self.handle, ms, C.byref(n)), self.ctx._h)
        k = max(int(n.value), 1)
        return {"pack_ms": ms[0] / k, "reduce_scatter_ms": ms[1] / k, "gather_ms": ms[2] / k, "unpack_ms": ms[3] / k,
                "composites": int(n.value)}

    def close(self):
        from .context import capi
        if self.handle:
            self.ctx.Sync()
            capi().fgl_comm_destroy(self.handle)
            self.handle = None


class PeerGroup:
    """fgl_peer_group: the exact, sparse, fused composite over peer memory.  ``records`` (one fgl_peer_export
    record per rank, in rank order) are exchanged with ``all_gather_object`` when not given -- pass them
    explicitly for ranks that live in one process (``PeerGroup.local``)."""

    def __init__(self, ctx, rank: int, world: int, group=None, records=None):
        import ctypes as C
        from .context import _check, capi
        self.ctx, self.rank, self.world = ctx, rank, world
        if records is None:
            import torch.distributed as dist
            mine = ctx.PeerExport()
            records = [None] * world
            if world > 1:
                dist.all_gather_object(records, mine, group=group)
            else:
                records[0] = mine
        blob = b"".join(records)
        self.handle = C.c_void_p()
        _check(capi().fgl_peer_group_create(ctx._h, rank, world, blob, C.byref(self.handle)), ctx._h)

    @classmethod
    def local(cls, ctxs):
        """One group object per context for ranks that are contexts of THIS process (any mix of devices)."""
        records = [c.PeerExport() for c in ctxs]
        return [cls(c, r, len(ctxs), records=records) for r, c in enumerate(ctxs)]

    def composite(self, root: int = -1, color_only: bool = False):
        """Enqueue signal -> wait -> sparse composite -> signal -> wait on the context's stream (no host sync).
        Afterwards rank `root` (every rank if root < 0) holds the frame; with ``color_only`` only its colours
        (FGL_COMPOSITE_COLOR_ONLY: enough to present the frame, a third of the bytes)."""
        from .context import _check, capi
        _check(capi().fgl_peer_composite(self.ctx._h, self.handle, int(root), 1 if color_only else 0), self.ctx._h)

    @staticmethod
    def composite_local(groups, root: int = -1, color_only: bool = False):
        """Composite for ranks that are contexts of THIS process, submitted by this one thread: phase by phase over all
        ranks (fgl_peer_composite_phase), so that no waiting kernel is ever submitted ahead of the signal it waits for."""
        from .context import _check, capi
        for phase in (1, 2, 3):
            for g in groups:
                _check(capi().fgl_peer_composite_phase(g.ctx._h, g.handle, int(root), 1 if color_only else 0, phase), g.ctx._h)

    def stage_times(self):
        """{wait_all_drawn_ms, composite_ms, wait_all_done_ms} per composite issued while profiling was on."""
        import ctypes as C
        from .context import _check, capi
        ms, n = (C.c_float * 4)(), C.c_uint32(0)
        _check(capi().fgl_peer_stage_times(self.ctx._h, self.handle, ms, C.byref(n)), self.ctx._h)
        k = max(int(n.value), 1)
        return {"wait_all_drawn_ms": ms[0] / k, "composite_ms": ms[1] / k, "wait_all_done_ms": ms[2] / k,
                "composites": int(n.value)}

    def status(self):
        """Wait for the stream; raises if a rank never arrived."""
        from .context import _check, capi
        _check(capi().fgl_peer_status(self.ctx._h, self.handle), self.ctx._h)

    def close(self):
        from .context import capi
        if self.handle:
            capi().fgl_peer_group_destroy(self.handle)
            self.handle = None


def sort_last_draw(ctx, mesh, keys, rank: int, world: int, group=None):
    """Draw this rank's triangle range and composite: afterwards every rank's
    colour buffer holds the full image.  `keys` is a CUDA int64 tensor of
    width*height elements that lives on ctx's device.  Returns the summed
    RasterizeInfo of this rank's draws (TotalPixels adds up over ranks;
    UpdatedPixels is per-rank and not comparable with a single-GPU run)."""
    import torch
    if not sort_last_valid(ctx):
        raise ValueError("sort-last composite needs ReadDepth, WriteDepth and DepthBias == 0")
    dm = ctx.device_mesh(mesh)
    first, count = triangle_range(dm.num_triangles, rank, world)
    info = ctx.DrawTriangles(dm, first, count)
    stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=keys.device)
    ctx.CompositePack(keys.data_ptr())
    with torch.cuda.stream(stream):
        composite_min(keys, group)
    ctx.CompositeUnpack(keys.data_ptr())
    return info
