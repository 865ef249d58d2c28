"""Frame pipeline over the streaming entry points of the C ABI: the host -> device copy of frame
i+1's mesh overlaps the draw and the read-back of frame i (separate CUDA streams, events between
them), which is how an examples/animate.go-style loop -- re-pose the mesh on the host, draw, save
the image -- keeps PCIe and the SMs busy at the same time.

Each slot owns a device mesh, a pinned image and a fence; ``submit`` enqueues a whole frame without
blocking (unless ``depth`` frames are already in flight), ``collect`` returns the oldest frame's
image and RasterizeInfo.  Results are identical to the synchronous calls (same kernels, same order).
"""
from __future__ import annotations

from collections import deque
from typing import Optional

import numpy as np

from .context import Context, DeviceMesh, Fence, RasterizeInfo, pinned_empty
from .mesh import Mesh


class FramePipeline:
    def __init__(self, ctx: Context, template: Mesh, depth: int = 2, attributes=("position", "normal")):
        assert depth >= 1
        self.ctx = ctx
        self.attributes = tuple(attributes)
        self.slots = [{"mesh": DeviceMesh(ctx, template, self.attributes),
                       "image": pinned_empty((ctx.Height, ctx.Width, 4), np.uint8),
                       "fence": Fence(ctx)} for _ in range(depth)]
        # one synchronous draw sizes the work buffers (async draws cannot regrow them)
        ctx.DrawMesh(self.slots[0]["mesh"])
        self.inflight = deque()
        self.next = 0

    def submit(self, mesh: Mesh, clear_color=None, clear_depth: bool = True):
        """Enqueue upload + clears + DrawMesh + read-back of one frame.  ``mesh``'s arrays must stay
        unchanged until the frame is collected (keep them in ``pinned_empty`` arrays for a real overlap)."""
        if len(self.inflight) == len(self.slots):
            raise RuntimeError("pipeline full: collect() a frame first")
        slot = self.slots[self.next]
        self.next = (self.next + 1) % len(self.slots)
        slot["mesh"].update_async(mesh, self.attributes)
        if clear_depth:
            self.ctx.ClearDepthBuffer()
        if clear_color is not None:
            self.ctx.ClearColorBufferWith(clear_color)
        self.ctx.DrawMeshAsync(slot["mesh"])
        self.ctx.FrameEnd(slot["image"], slot["fence"])
        self.inflight.append(slot)

    def collect(self):
        """(image, RasterizeInfo) of the oldest frame in flight; the image is the slot's pinned buffer and
        is overwritten when the slot is reused ``depth`` submits later."""
        slot = self.inflight.popleft()
        info = slot["fence"].wait()
        return slot["image"], info

    def __len__(self):
        return len(self.inflight)


class IndexedFramePipeline:
    """The same pipeline for a mesh kept as shared-vertex tables (``Mesh.indexed()``): the corner indices stay on
    the device, a frame uploads only its v / vn tables (fgl_mesh_update_indexed_async) -- 48 bytes per shared vertex
    instead of 144 per triangle -- and the expansion to the device layout runs on the copy stream."""

    def __init__(self, ctx: Context, v: np.ndarray, vn: np.ndarray, corners: np.ndarray, depth: int = 2):
        assert depth >= 1
        self.ctx = ctx
        vt = np.zeros((1, 3), dtype=np.float64)
        self.slots = [{"mesh": DeviceMesh.FromIndexed(ctx, v, vt, vn, corners),
                       "image": pinned_empty((ctx.Height, ctx.Width, 4), np.uint8),
                       "fence": Fence(ctx)} for _ in range(depth)]
        ctx.DrawMesh(self.slots[0]["mesh"])   # one synchronous draw sizes the work buffers
        self.inflight = deque()
        self.next = 0

    def submit(self, v: np.ndarray, vn: np.ndarray, clear_color=None, clear_depth: bool = True):
        """``v`` / ``vn``: this frame's tables (float64 [n][3]; pinned for a real overlap), unchanged until collected."""
        if len(self.inflight) == len(self.slots):
            raise RuntimeError("pipeline full: collect() a frame first")
        slot = self.slots[self.next]
        self.next = (self.next + 1) % len(self.slots)
        slot["mesh"].update_indexed_async(v=v, vn=vn)
        if clear_depth:
            self.ctx.ClearDepthBuffer()
        if clear_color is not None:
            self.ctx.ClearColorBufferWith(clear_color)
        self.ctx.DrawMeshAsync(slot["mesh"])
        self.ctx.FrameEnd(slot["image"], slot["fence"])
        self.inflight.append(slot)

    def collect(self):
        slot = self.inflight.popleft()
        info = slot["fence"].wait()
        return slot["image"], info

    def __len__(self):
        return len(self.inflight)
