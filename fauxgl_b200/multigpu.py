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

The reduction itself is ``torch.distributed.all_reduce(MIN)`` on int64 -- NCCL
over NVLink/NVSwitch on the GPUs, gloo in the CPU tests.
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


class PeerComposite:
    """Sort-last composite over peer memory: ONE kernel per rank does the depth compare and the exchange
    with P2P loads/stores through NVLink (fgl_composite_peer) instead of pack -> NCCL all-reduce -> unpack.
    It keeps the float64 depth and breaks ties towards the higher rank (the later triangle range), so the
    result equals a single-GPU render bit for bit in the order-independent state.

    Buffers are shared through CUDA IPC handles exchanged once with ``all_gather_object``; host barriers
    fence the kernel (all ranks drawn before anyone reads; all composited before anyone draws again)."""

    def __init__(self, ctx, rank: int, world: int, group=None):
        import torch.distributed as dist
        self.ctx, self.rank, self.world, self.group = ctx, rank, world, group
        mine = ctx.IpcExport()
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, mine, group=group)
        else:
            handles[0] = mine
        self.color, self.depth, self._opened = [], [], []
        for r, (hc, hd) in enumerate(handles):
            if r == rank:
                self.color.append(ctx.color_ptr)
                self.depth.append(ctx.depth_ptr)
            else:
                pc, pd = ctx.IpcOpen(hc, hd)
                self._opened.append((pc, pd))
                self.color.append(pc)
                self.depth.append(pd)

    def _barrier(self):
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier(group=self.group)

    def composite(self):
        """Call after this rank's draws of the frame; returns when every rank holds the full frame."""
        info = self.ctx.Sync()               # (RasterizeInfo of this rank's async draws, if any)
        self._barrier()                      # every rank's buffers are final
        self.ctx.CompositePeer(self.rank, self.color, self.depth)
        self.ctx.Sync()
        self._barrier()                      # every stripe has been written everywhere
        return info

    def close(self):
        for pc, pd in self._opened:
            self.ctx.IpcClose(pc, pd)
        self._opened = []


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
