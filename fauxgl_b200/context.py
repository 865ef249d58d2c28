"""Python mirror of the reference's ``Context`` (context.go:40-145, 351-439) over
the C ABI of ``include/fauxgl_b200.h``.

This is the host-side stand-in for the Go package (go/fauxgl is the cgo shim a
Go program would import; the build image has no Go toolchain).  Names, argument
meaning and defaults follow the reference so parity tests read like its
examples.  Every call goes to ``libfauxgl_b200.so`` (hand-written sm_100a CUDA
kernels); there is no CPU fallback: if the library or a CUDA device is
missing, construction raises ``FauxglError``.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import weakref
from typing import NamedTuple, Optional

import numpy as np

from .color import Color, Transparent
from .matrix import Identity, Matrix
from .mesh import Box, Mesh
from .vector import Vector
from .shader import (SHADER_PHONG, SHADER_SOLID, SHADER_TEXTURE, ImageTexture, PhongShader,
                     SolidColorShader, TextureShader)

COMM_ID_BYTES, PEER_EXPORT_BYTES = 128, 512   # FGL_COMM_ID_BYTES, FGL_PEER_EXPORT_BYTES
COMPOSITE_COLOR_ONLY = 1                       # FGL_COMPOSITE_COLOR_ONLY
FaceCW, FaceCCW = 1, 2
CullNone, CullFront, CullBack = 1, 2, 3

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfauxgl_b200.so")


class FauxglError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__("fauxgl_b200 error %d: %s" % (status, message))
        self.status = status


class RasterizeInfo(NamedTuple):
    """context.go:28-38"""
    TotalPixels: int
    UpdatedPixels: int

    def Add(self, other):
        return RasterizeInfo(self.TotalPixels + other.TotalPixels, self.UpdatedPixels + other.UpdatedPixels)


class _State(C.Structure):
    _fields_ = [("read_depth", C.c_int32), ("write_depth", C.c_int32), ("write_color", C.c_int32),
                ("alpha_blend", C.c_int32), ("wireframe", C.c_int32), ("front_face", C.c_int32),
                ("cull", C.c_int32), ("x_guard", C.c_int32), ("line_width", C.c_double), ("depth_bias", C.c_double)]


class _Shader(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("matrix", C.c_double * 16),
                ("light", C.c_double * 3), ("camera", C.c_double * 3), ("object", C.c_double * 4),
                ("ambient", C.c_double * 4), ("diffuse", C.c_double * 4), ("specular", C.c_double * 4),
                ("specular_power", C.c_double), ("color", C.c_double * 4), ("texture", C.c_void_p)]


class _MeshDesc(C.Structure):
    _fields_ = [("ntriangles", C.c_uint64), ("position", C.c_void_p), ("normal", C.c_void_p),
                ("texture", C.c_void_p), ("color", C.c_void_p), ("nlines", C.c_uint64),
                ("lposition", C.c_void_p), ("lnormal", C.c_void_p), ("ltexture", C.c_void_p),
                ("lcolor", C.c_void_p)]


class _Info(C.Structure):
    _fields_ = [("total_pixels", C.c_uint64), ("updated_pixels", C.c_uint64)]


class DrawStats(C.Structure):
    _fields_ = [("prims_in", C.c_uint64), ("records", C.c_uint64), ("pairs", C.c_uint64),
                ("clip_triangles", C.c_uint64), ("tiles_x", C.c_uint32), ("tiles_y", C.c_uint32),
                ("tile_w", C.c_uint32), ("tile_h", C.c_uint32), ("kernel_launches", C.c_uint32),
                ("retries", C.c_uint32)]


class StageTimes(C.Structure):
    _fields_ = [("geometry_ms", C.c_float), ("spans_ms", C.c_float), ("sort_ms", C.c_float),
                ("raster_ms", C.c_float), ("draws", C.c_uint32)]


# every symbol include/fauxgl_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p


class IndexedDesc(C.Structure):
    """fgl_indexed_desc (include/fauxgl_b200.h)."""
    _fields_ = [("v", _P), ("vt", _P), ("vn", _P), ("nv", C.c_uint64), ("nvt", C.c_uint64), ("nvn", C.c_uint64),
                ("corners", _P), ("ntriangles", C.c_uint64)]


ABI = [
    ("fgl_abi_version", C.c_int, []),
    ("fgl_last_error", C.c_char_p, [_P]),
    ("fgl_device_count", C.c_int, []),
    ("fgl_host_alloc", C.c_int, [C.c_size_t, C.POINTER(_P)]),
    ("fgl_host_free", C.c_int, [_P]),
    ("fgl_context_create", C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    ("fgl_context_destroy", C.c_int, [_P]),
    ("fgl_context_size", C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ("fgl_clear_color", C.c_int, [_P, C.POINTER(C.c_uint8)]),
    ("fgl_clear_depth", C.c_int, [_P, C.c_double]),
    ("fgl_mesh_create", C.c_int, [_P, C.POINTER(_MeshDesc), C.POINTER(_P)]),
    ("fgl_mesh_update", C.c_int, [_P, _P, C.POINTER(_MeshDesc)]),
    ("fgl_mesh_create_stl", C.c_int, [_P, _P, C.c_uint64, C.POINTER(_P)]),
    ("fgl_mesh_bounds", C.c_int, [_P, _P, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    ("fgl_mesh_update_async", C.c_int, [_P, _P, C.POINTER(_MeshDesc)]),
    ("fgl_mesh_upload_wait", C.c_int, [_P, _P]),
    ("fgl_mesh_destroy", C.c_int, [_P]),
    ("fgl_mesh_counts", C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("fgl_mesh_transform", C.c_int, [_P, _P, C.POINTER(C.c_double)]),
    ("fgl_mesh_read", C.c_int, [_P, _P, _P, _P, _P, _P]),
    ("fgl_mesh_create_indexed", C.c_int, [_P, _P, C.POINTER(_P)]),
    ("fgl_mesh_update_indexed_async", C.c_int, [_P, _P, _P]),
    ("fgl_mesh_smooth_normals", C.c_int, [_P, _P]),
    ("fgl_mesh_smooth_normals_threshold", C.c_int, [_P, _P, C.c_double]),
    ("fgl_texture_create", C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    ("fgl_texture_destroy", C.c_int, [_P]),
    ("fgl_draw_triangles", C.c_int, [_P, C.POINTER(_State), C.POINTER(_Shader), _P, C.c_uint64, C.c_uint64, C.POINTER(_Info)]),
    ("fgl_draw_lines", C.c_int, [_P, C.POINTER(_State), C.POINTER(_Shader), _P, C.c_uint64, C.c_uint64, C.POINTER(_Info)]),
    ("fgl_draw_triangles_each", C.c_int, [_P, C.POINTER(_State), C.POINTER(_Shader), _P, C.c_uint64, C.c_uint64, _P, C.POINTER(_Info)]),
    ("fgl_draw_lines_each", C.c_int, [_P, C.POINTER(_State), C.POINTER(_Shader), _P, C.c_uint64, C.c_uint64, _P, C.POINTER(_Info)]),
    ("fgl_draw_triangles_async", C.c_int, [_P, C.POINTER(_State), C.POINTER(_Shader), _P, C.c_uint64, C.c_uint64]),
    ("fgl_draw_lines_async", C.c_int, [_P, C.POINTER(_State), C.POINTER(_Shader), _P, C.c_uint64, C.c_uint64]),
    ("fgl_sync", C.c_int, [_P, C.POINTER(_Info)]),
    ("fgl_frame_end", C.c_int, [_P, _P, C.c_size_t, C.POINTER(_P)]),
    ("fgl_fence_wait", C.c_int, [_P, _P, C.POINTER(_Info)]),
    ("fgl_fence_destroy", C.c_int, [_P]),
    ("fgl_get_draw_stats", C.c_int, [_P, C.POINTER(DrawStats)]),
    ("fgl_graph_begin", C.c_int, [_P]),
    ("fgl_graph_end", C.c_int, [_P, C.POINTER(_P)]),
    ("fgl_graph_launch", C.c_int, [_P, _P]),
    ("fgl_graph_destroy", C.c_int, [_P]),
    ("fgl_set_profiling", C.c_int, [_P, C.c_int]),
    ("fgl_get_stage_times", C.c_int, [_P, C.POINTER(StageTimes)]),
    ("fgl_read_color", C.c_int, [_P, _P, C.c_size_t]),
    ("fgl_read_depth", C.c_int, [_P, _P]),
    ("fgl_depth_image", C.c_int, [_P, _P]),
    ("fgl_write_color", C.c_int, [_P, _P, C.c_size_t]),
    ("fgl_write_depth", C.c_int, [_P, _P]),
    ("fgl_resolve", C.c_int, [_P, C.c_int, _P]),
    ("fgl_resolve_device", C.c_int, [_P, C.c_int]),
    ("fgl_read_resolved", C.c_int, [_P, _P]),
    ("fgl_composite_pack", C.c_int, [_P, _P]),
    ("fgl_composite_unpack", C.c_int, [_P, _P]),
    ("fgl_composite_min", C.c_int, [_P, _P, _P, C.c_uint64]),
    ("fgl_ipc_export", C.c_int, [_P, _P, _P]),
    ("fgl_ipc_open", C.c_int, [_P, _P, _P, C.POINTER(_P), C.POINTER(_P)]),
    ("fgl_ipc_close", C.c_int, [_P, _P, _P]),
    ("fgl_composite_peer", C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_P), C.POINTER(_P)]),
    ("fgl_comm_unique_id", C.c_int, [_P]),
    ("fgl_comm_init", C.c_int, [_P, C.c_int, C.c_int, _P, C.POINTER(_P)]),
    ("fgl_comm_destroy", C.c_int, [_P]),
    ("fgl_composite", C.c_int, [_P, _P, C.c_int]),
    ("fgl_comm_stage_times", C.c_int, [_P, _P, C.POINTER(C.c_float), C.POINTER(C.c_uint32)]),
    ("fgl_peer_stage_times", C.c_int, [_P, _P, C.POINTER(C.c_float), C.POINTER(C.c_uint32)]),
    ("fgl_peer_export", C.c_int, [_P, _P]),
    ("fgl_peer_group_create", C.c_int, [_P, C.c_int, C.c_int, _P, C.POINTER(_P)]),
    ("fgl_peer_group_destroy", C.c_int, [_P]),
    ("fgl_peer_composite", C.c_int, [_P, _P, C.c_int, C.c_int]),
    ("fgl_peer_composite_phase", C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int]),
    ("fgl_peer_status", C.c_int, [_P, _P]),
    ("fgl_debug_tile_cycles", C.c_int, [_P, _P, C.c_uint64]),
    ("fgl_probe_atomic_rate", C.c_int, [_P, C.c_uint64, C.POINTER(C.c_double)]),
    ("fgl_debug_div_check", C.c_int, [_P, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("fgl_stream", _P, [_P]),
    ("fgl_color_device_ptr", _P, [_P]),
    ("fgl_depth_device_ptr", _P, [_P]),
]

_lib = None


def capi():
    """Load libfauxgl_b200.so (built by fauxgl_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is None:
        path = os.environ.get("FGL_LIB", LIB_PATH)  # tuning aid: a variant built by tools/build_variant.py
        if not os.path.exists(path):
            raise FauxglError(-3, "libfauxgl_b200.so has not been built (python -m fauxgl_b200.build); "
                                  "there is no CPU fallback")
        L = C.CDLL(path)
        for name, res, args in ABI:
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc: int, ctx=None):
    if rc != 0:
        msg = capi().fgl_last_error(ctx)
        raise FauxglError(rc, msg.decode() if msg else "?")


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


def pinned_empty(shape, dtype) -> np.ndarray:
    """A numpy array in page-locked host memory (fgl_host_alloc), freed with the array."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = _P()
    _check(capi().fgl_host_alloc(n, C.byref(p)))
    buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(buf, capi().fgl_host_free, p)
    return arr


class Fence:
    """fgl_fence: completion of one pipelined frame (Context.FrameEnd)."""

    def __init__(self, ctx: "Context"):
        self.ctx = ctx
        self.handle = _P()
        self._fin = None

    def wait(self) -> RasterizeInfo:
        info = _Info()
        if self.handle:
            _check(capi().fgl_fence_wait(self.ctx._h, self.handle, C.byref(info)), self.ctx._h)
        return RasterizeInfo(info.total_pixels, info.updated_pixels)


class FrameGraph:
    """fgl_graph: a recorded frame (CUDA graph) of one context; ``launch()`` replays it with one call."""

    def __init__(self, ctx: "Context", handle):
        self.ctx, self.handle = ctx, handle
        self._fin = weakref.finalize(self, capi().fgl_graph_destroy, handle)

    def launch(self):
        _check(capi().fgl_graph_launch(self.ctx._h, self.handle), self.ctx._h)

    def Close(self):
        self._fin()


class DeviceTexture:
    def __init__(self, ctx: "Context", tex: ImageTexture):
        self.handle = _P()
        px = np.ascontiguousarray(tex.pixels)   # uint8 texels, or uint16 for TEX_RGBA64
        _check(capi().fgl_texture_create(ctx._h, px.ctypes.data, tex.Width, tex.Height, tex.format,
                                         C.byref(self.handle)), ctx._h)
        self._fin = weakref.finalize(self, capi().fgl_texture_destroy, self.handle)


class DeviceMesh:
    """Device-resident copy of a Mesh (fgl_mesh_create)."""

    def __init__(self, ctx: "Context", mesh: Mesh, attributes=("position", "normal", "texture", "color")):
        self.ctx = ctx
        self.generation = mesh.generation
        self.attributes = tuple(attributes)
        self.num_triangles, self.num_lines = mesh.num_triangles, mesh.num_lines
        d, keep = self._desc(mesh, self.attributes)
        self.handle = _P()
        _check(capi().fgl_mesh_create(ctx._h, C.byref(d), C.byref(self.handle)), ctx._h)
        self._fin = weakref.finalize(self, capi().fgl_mesh_destroy, self.handle)

    @classmethod
    def FromSTL(cls, ctx: "Context", source) -> "DeviceMesh":
        """LoadSTL (stl.go:23-57) of a BINARY STL straight to the device layout: only the 50-byte
        records cross PCIe (fgl_mesh_create_stl).  ``source`` is a path or the file's bytes."""
        if isinstance(source, (str, os.PathLike)):
            with open(source, "rb") as f:
                source = f.read()
        data = bytes(source)
        count = int.from_bytes(data[80:84], "little") if len(data) >= 84 else -1
        if count < 0 or len(data) != 84 + 50 * count:
            raise FauxglError(-1, "not a binary STL (ASCII files go through mesh.LoadSTL)")
        rec = np.frombuffer(data, dtype=np.uint8, count=50 * count, offset=84)
        self = cls.__new__(cls)
        self.ctx = ctx
        self.generation = 0
        self.attributes = ("position", "normal")
        self.num_triangles, self.num_lines = count, 0
        self.handle = _P()
        _check(capi().fgl_mesh_create_stl(ctx._h, rec.ctypes.data if count else None, count, C.byref(self.handle)), ctx._h)
        self._fin = weakref.finalize(self, capi().fgl_mesh_destroy, self.handle)
        return self

    @classmethod
    def FromOBJ(cls, ctx: "Context", path) -> "DeviceMesh":
        """LoadOBJ (obj.go:19-79) with the expansion on the device: the text is parsed on the host
        (``mesh.ParseOBJ``), the v / vt / vn tables and 36 B of indices per triangle cross PCIe, and
        ``fgl_mesh_create_indexed`` gathers them into the planes and applies Triangle.FixNormals."""
        from .mesh import ParseOBJ
        vs, vts, vns, corners = ParseOBJ(path)
        keep = [np.ascontiguousarray(vs, dtype=np.float64), np.ascontiguousarray(vts, dtype=np.float64),
                np.ascontiguousarray(vns, dtype=np.float64), np.ascontiguousarray(corners, dtype=np.int32)]
        d = IndexedDesc()
        d.v, d.vt, d.vn = (C.cast(a.ctypes.data, _P) for a in keep[:3])
        d.nv, d.nvt, d.nvn = len(keep[0]), len(keep[1]), len(keep[2])
        d.corners = C.cast(keep[3].ctypes.data, _P)
        d.ntriangles = len(keep[3])
        self = cls.__new__(cls)
        self.ctx = ctx
        self.generation = 0
        self.attributes = ("position", "normal", "texture")
        self.num_triangles, self.num_lines = int(d.ntriangles), 0
        self.table_sizes = (int(d.nv), int(d.nvt), int(d.nvn))
        self.handle = _P()
        _check(capi().fgl_mesh_create_indexed(ctx._h, C.byref(d), C.byref(self.handle)), ctx._h)
        self._fin = weakref.finalize(self, capi().fgl_mesh_destroy, self.handle)
        return self

    @classmethod
    def FromIndexed(cls, ctx: "Context", v, vt, vn, corners) -> "DeviceMesh":
        """fgl_mesh_create_indexed from tables already in memory: v / vt / vn [n][3] float64 (entry 0 of each is the
        reference's 1-based tables' unused zero entry, or any entry -- the indices decide) and corners [T][3][3]
        int32 (v, vt, vn index per triangle corner).  The indices stay on the device: ``update_indexed_async``
        re-poses the mesh by sending only the tables."""
        keep = [np.ascontiguousarray(v, dtype=np.float64), np.ascontiguousarray(vt, dtype=np.float64),
                np.ascontiguousarray(vn, dtype=np.float64), np.ascontiguousarray(corners, dtype=np.int32)]
        d = IndexedDesc()
        d.v, d.vt, d.vn = (C.cast(a.ctypes.data, _P) for a in keep[:3])
        d.nv, d.nvt, d.nvn = len(keep[0]), len(keep[1]), len(keep[2])
        d.corners = C.cast(keep[3].ctypes.data, _P)
        d.ntriangles = len(keep[3])
        self = cls.__new__(cls)
        self.ctx = ctx
        self.generation = 0
        self.attributes = ("position", "normal", "texture")
        self.num_triangles, self.num_lines = int(d.ntriangles), 0
        self.table_sizes = (int(d.nv), int(d.nvt), int(d.nvn))
        self.handle = _P()
        _check(capi().fgl_mesh_create_indexed(ctx._h, C.byref(d), C.byref(self.handle)), ctx._h)
        self._fin = weakref.finalize(self, capi().fgl_mesh_destroy, self.handle)
        return self

    def update_indexed_async(self, v=None, vt=None, vn=None):
        """fgl_mesh_update_indexed_async: enqueue new v / vt / vn tables (float64 [n][3], same sizes; pinned memory
        for a real overlap) on the copy stream and return; the arrays must stay unchanged until ``upload_wait``."""
        d = IndexedDesc()
        keep = []
        for name, a, n in (("v", v, self.table_sizes[0]), ("vt", vt, self.table_sizes[1]), ("vn", vn, self.table_sizes[2])):
            if a is None:
                continue
            assert a.dtype == np.float64 and a.flags.c_contiguous and a.shape == (n, 3), (name, a.shape, n)
            keep.append(a)
            setattr(d, name, C.cast(a.ctypes.data, _P))
        d.nv, d.nvt, d.nvn = self.table_sizes
        d.ntriangles = self.num_triangles
        self._pending = keep
        _check(capi().fgl_mesh_update_indexed_async(self.ctx._h, self.handle, C.byref(d)), self.ctx._h)

    def BoundingBox(self) -> Box:
        """Mesh.BoundingBox (mesh.go:153-165) of the device copy."""
        mn, mx = (C.c_double * 3)(), (C.c_double * 3)()
        _check(capi().fgl_mesh_bounds(self.ctx._h, self.handle, mn, mx), self.ctx._h)
        return Box(Vector(*mn), Vector(*mx))

    def FitInside(self, box: Box, anchor: Vector) -> Matrix:
        """Mesh.FitInside (mesh.go:142-151) without the mesh leaving the device."""
        bb = self.BoundingBox()
        scale = box.Size().Div(bb.Size()).MinComponent()
        extra = box.Size().Sub(bb.Size().MulScalar(scale))
        matrix = Identity()
        matrix = matrix.Translate(bb.Min.Negate())
        matrix = matrix.Scale(Vector(scale, scale, scale))
        matrix = matrix.Translate(box.Min.Add(extra.Mul(anchor)))
        self.Transform(matrix)
        return matrix

    def BiUnitCube(self) -> Matrix:  # mesh.go:127-130
        return self.FitInside(Box(Vector(-1.0, -1.0, -1.0), Vector(1.0, 1.0, 1.0)), Vector(0.5, 0.5, 0.5))

    def UnitCube(self) -> Matrix:  # mesh.go:122-125
        return self.FitInside(Box(Vector(-0.5, -0.5, -0.5), Vector(0.5, 0.5, 0.5)), Vector(0.5, 0.5, 0.5))

    def update(self, mesh: Mesh, attributes=None):
        """fgl_mesh_update: re-upload (a subset of) the attributes into the same device buffers."""
        assert (mesh.num_triangles, mesh.num_lines) == (self.num_triangles, self.num_lines)
        attributes = self.attributes if attributes is None else tuple(attributes)
        d, keep = self._desc(mesh, attributes, position="position" in attributes)
        _check(capi().fgl_mesh_update(self.ctx._h, self.handle, C.byref(d)), self.ctx._h)
        self.generation = mesh.generation

    @staticmethod
    def _desc(mesh: Mesh, attributes, position=True):
        d = _MeshDesc()
        keep = []

        def arr(a, want):
            if not want or a is None or len(a) == 0:
                return None
            a = np.ascontiguousarray(a, dtype=np.float64)
            keep.append(a)
            return a.ctypes.data
        d.ntriangles = mesh.num_triangles
        d.position = arr(mesh.position, position)
        d.normal = arr(mesh.normal, "normal" in attributes)
        d.texture = arr(mesh.texture, "texture" in attributes)
        d.color = arr(mesh.color, "color" in attributes)
        d.nlines = mesh.num_lines
        d.lposition = arr(mesh.lposition, position)
        d.lnormal = arr(mesh.lnormal, "normal" in attributes)
        d.ltexture = arr(mesh.ltexture, "texture" in attributes)
        d.lcolor = arr(mesh.lcolor, "color" in attributes)
        return d, keep

    def update_async(self, mesh: Mesh, attributes=None):
        """fgl_mesh_update_async: enqueue the re-upload on the copy stream and return; ``mesh``'s arrays
        (pinned for a real overlap, see ``pinned_empty``) must stay unchanged until ``upload_wait``."""
        assert (mesh.num_triangles, mesh.num_lines) == (self.num_triangles, self.num_lines)
        attributes = self.attributes if attributes is None else tuple(attributes)
        d, keep = self._desc(mesh, attributes, position="position" in attributes)
        self._pending = keep
        _check(capi().fgl_mesh_update_async(self.ctx._h, self.handle, C.byref(d)), self.ctx._h)
        self.generation = mesh.generation

    def upload_wait(self):
        _check(capi().fgl_mesh_upload_wait(self.ctx._h, self.handle), self.ctx._h)
        self._pending = None

    def Transform(self, matrix: Matrix):
        """Mesh.Transform on the device (mesh.go:167-175)."""
        m = (C.c_double * 16)(*matrix)
        _check(capi().fgl_mesh_transform(self.ctx._h, self.handle, m), self.ctx._h)

    def SmoothNormals(self):
        """Mesh.SmoothNormals on the device (mesh.go:105-120), bit-identical to the host mirror."""
        _check(capi().fgl_mesh_smooth_normals(self.ctx._h, self.handle), self.ctx._h)

    def SmoothNormalsThreshold(self, radians: float):
        """Mesh.SmoothNormalsThreshold on the device (mesh.go:80-103); the cosine is taken here, on the host."""
        _check(capi().fgl_mesh_smooth_normals_threshold(self.ctx._h, self.handle, math.cos(radians)), self.ctx._h)

    def read(self):
        pos = np.empty((self.num_triangles, 3, 3)); nrm = np.empty((self.num_triangles, 3, 3))
        lpos = np.empty((self.num_lines, 2, 3)); lnrm = np.empty((self.num_lines, 2, 3))
        _check(capi().fgl_mesh_read(self.ctx._h, self.handle, _ptr(pos), _ptr(nrm), _ptr(lpos), _ptr(lnrm)), self.ctx._h)
        return pos, nrm, lpos, lnrm

    def Close(self):
        self._fin()


class Context:
    """context.go:40-58.  Exported fields keep the reference's names and
    defaults (NewContext, context.go:60-81).  The device owns the colour and
    depth buffers between draws; ``Image()`` / ``DepthBuffer`` read them back."""

    def __init__(self, width: int, height: int, device: int = 0):
        self._h = _P()
        _check(capi().fgl_context_create(int(width), int(height), int(device), C.byref(self._h)))
        self._fin = weakref.finalize(self, capi().fgl_context_destroy, self._h)
        self.Width, self.Height = int(width), int(height)
        self.ClearColor = Transparent
        self.Shader = SolidColorShader(Identity(), Color(1.0, 0.0, 1.0, 1.0))
        self.ReadDepth = True
        self.WriteDepth = True
        self.WriteColor = True
        self.AlphaBlend = True
        self.Wireframe = False
        self.FrontFace = FaceCCW
        self.Cull = CullBack
        self.LineWidth = 2.0
        self.DepthBias = 0.0
        # Not a field of the reference: False keeps its index rule (context.go:223-228: only i = y*W + x is
        # range-checked, so a fat line leaving the screen sideways aliases into the neighbouring row); True drops
        # every fragment with x outside [0, Width) instead.
        self.XGuard = False
        self._meshes = weakref.WeakKeyDictionary()   # Mesh -> DeviceMesh
        self._textures = weakref.WeakKeyDictionary()  # ImageTexture -> DeviceTexture
        self._keep = None
        # which vertex attributes DrawMesh(host_mesh) uploads; a Phong+ObjectColor scene needs
        # only position+normal (the reference copies all 17 doubles per vertex regardless)
        self.upload_attributes = ("position", "normal", "texture", "color")

    # -- lifetime ------------------------------------------------------------------
    def Close(self):
        self._meshes.clear()
        self._textures.clear()
        self._fin()

    # -- buffers -------------------------------------------------------------------
    def Image(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        """context.go:83: ColorBuffer as (H,W,4) uint8, NRGBA.  ``out`` may be a
        caller-owned (e.g. pinned) buffer to read into."""
        if out is None:
            out = np.empty((self.Height, self.Width, 4), dtype=np.uint8)
        assert out.shape == (self.Height, self.Width, 4) and out.dtype == np.uint8 and out.flags.c_contiguous
        _check(capi().fgl_read_color(self._h, out.ctypes.data, 0), self._h)
        return out

    @property
    def ColorBuffer(self) -> np.ndarray:
        return self.Image()

    @property
    def DepthBuffer(self) -> np.ndarray:
        out = np.empty((self.Height, self.Width), dtype=np.float64)
        _check(capi().fgl_read_depth(self._h, out.ctypes.data), self._h)
        return out

    def DepthImage(self) -> np.ndarray:
        """context.go:87-117: the depth buffer as a Gray16 image, (H,W) uint16, normalised on the device."""
        out = np.empty((self.Height, self.Width), dtype=np.uint16)
        _check(capi().fgl_depth_image(self._h, out.ctypes.data), self._h)
        return out

    def UploadColorBuffer(self, pix: np.ndarray):
        pix = np.ascontiguousarray(pix, dtype=np.uint8)
        assert pix.shape == (self.Height, self.Width, 4)
        _check(capi().fgl_write_color(self._h, pix.ctypes.data, 0), self._h)

    def UploadDepthBuffer(self, depth: np.ndarray):
        depth = np.ascontiguousarray(depth, dtype=np.float64)
        assert depth.shape == (self.Height, self.Width)
        _check(capi().fgl_write_depth(self._h, depth.ctypes.data), self._h)

    def ClearColorBufferWith(self, color: Color):  # context.go:119
        c = (C.c_uint8 * 4)(*Color(*color).NRGBA())
        _check(capi().fgl_clear_color(self._h, c), self._h)

    def ClearColorBuffer(self):  # context.go:133
        self.ClearColorBufferWith(self.ClearColor)

    def ClearDepthBufferWith(self, value: float):  # context.go:137
        _check(capi().fgl_clear_depth(self._h, float(value)), self._h)

    def ClearDepthBuffer(self):  # context.go:143
        self.ClearDepthBufferWith(np.finfo(np.float64).max)

    # -- draw ---------------------------------------------------------------------------
    def _state(self) -> _State:
        s = _State()
        s.read_depth, s.write_depth, s.write_color = int(self.ReadDepth), int(self.WriteDepth), int(self.WriteColor)
        s.alpha_blend, s.wireframe = int(self.AlphaBlend), int(self.Wireframe)
        s.front_face, s.cull = int(self.FrontFace), int(self.Cull)
        s.x_guard = int(self.XGuard)
        s.line_width, s.depth_bias = float(self.LineWidth), float(self.DepthBias)
        return s

    def _shader(self) -> _Shader:
        sh = self.Shader
        # the closed set the Go shim resolves by type switch; anything else has no device
        # equivalent and is an error, never a CPU fallback
        if not isinstance(sh, (SolidColorShader, TextureShader, PhongShader)):
            raise FauxglError(-4, "Shader %r has no device implementation (only SolidColorShader, "
                                  "TextureShader and PhongShader do); there is no CPU fallback" % type(sh).__name__)
        d = sh.describe()
        s = _Shader()
        s.kind = d["kind"]
        s.matrix[:] = d["matrix"]
        for key in ("light", "camera", "object", "ambient", "diffuse", "specular", "color"):
            if key in d:
                getattr(s, key)[:] = d[key]
        s.specular_power = d.get("specular_power", 0.0)
        tex = d.get("texture")
        if tex is not None:
            if not isinstance(tex, ImageTexture):
                raise FauxglError(-4, "Texture %r has no device implementation" % type(tex).__name__)
            dt = self._textures.get(tex)
            if dt is None:
                dt = DeviceTexture(self, tex)
                self._textures[tex] = dt
            s.texture = dt.handle
            self._keep = dt
        return s

    def device_mesh(self, mesh) -> DeviceMesh:
        if isinstance(mesh, DeviceMesh):
            return mesh
        dm = self._meshes.get(mesh)
        want = self._needed_attributes()
        if dm is not None and not set(want) <= set(dm.attributes):
            # first drawn with fewer attributes than this draw's shader reads (e.g. position+normal, now a
            # TextureShader): the device copy's missing planes are zero -- upload the union instead
            want = tuple(a for a in ("position", "normal", "texture", "color") if a in set(want) | set(dm.attributes))
            dm = None
        if dm is None or (dm.num_triangles, dm.num_lines) != (mesh.num_triangles, mesh.num_lines):
            dm = DeviceMesh(self, mesh, want)
            self._meshes[mesh] = dm
        elif dm.generation != mesh.generation:
            dm.update(mesh)   # mutated on the host since the last draw: re-upload in place
        return dm

    def _needed_attributes(self):
        """``upload_attributes`` plus whatever the current shader reads beyond it (shader.go:44,75-96)."""
        need = list(self.upload_attributes)
        sh = self.Shader
        if isinstance(sh, TextureShader) or (isinstance(sh, PhongShader) and sh.Texture is not None):
            need.append("texture")
        if isinstance(sh, PhongShader):
            need.append("normal")
            if sh.Texture is None and tuple(sh.ObjectColor) == (0.0, 0.0, 0.0, 0.0):
                need.append("color")
        return tuple(a for a in ("position", "normal", "texture", "color") if a in need)

    def DrawTriangles(self, mesh, first: int = 0, count: Optional[int] = None) -> RasterizeInfo:  # context.go:413
        dm = self.device_mesh(mesh)
        count = dm.num_triangles - first if count is None else count
        info = _Info()
        st, sh = self._state(), self._shader()
        _check(capi().fgl_draw_triangles(self._h, C.byref(st), C.byref(sh), dm.handle, first, count, C.byref(info)), self._h)
        return RasterizeInfo(info.total_pixels, info.updated_pixels)

    def DrawLines(self, mesh, first: int = 0, count: Optional[int] = None) -> RasterizeInfo:  # context.go:391
        dm = self.device_mesh(mesh)
        count = dm.num_lines - first if count is None else count
        info = _Info()
        st, sh = self._state(), self._shader()
        _check(capi().fgl_draw_lines(self._h, C.byref(st), C.byref(sh), dm.handle, first, count, C.byref(info)), self._h)
        return RasterizeInfo(info.total_pixels, info.updated_pixels)

    def _draw_each(self, fn, dm, first, count):
        infos = np.zeros((count, 2), dtype=np.uint64)
        info = _Info()
        st, sh = self._state(), self._shader()
        _check(fn(self._h, C.byref(st), C.byref(sh), dm.handle, first, count, infos.ctypes.data if count else None,
                  C.byref(info)), self._h)
        return infos

    def DrawLinesEach(self, mesh, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        """One (TotalPixels, UpdatedPixels) row per line, as if each were drawn with Context.DrawLine
        (context.go:351-368) in index order -- examples/silhouette.go:163-166 -- in one launch sequence."""
        dm = self.device_mesh(mesh)
        count = dm.num_lines - first if count is None else count
        return self._draw_each(capi().fgl_draw_lines_each, dm, first, count)

    def DrawTrianglesEach(self, mesh, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        """One (TotalPixels, UpdatedPixels) row per triangle (Context.DrawTriangle, context.go:370-389)."""
        dm = self.device_mesh(mesh)
        count = dm.num_triangles - first if count is None else count
        return self._draw_each(capi().fgl_draw_triangles_each, dm, first, count)

    def DrawMesh(self, mesh) -> RasterizeInfo:  # context.go:435
        info1 = self.DrawTriangles(mesh)
        info2 = self.DrawLines(mesh)
        return info1.Add(info2)

    def DrawMeshAsync(self, mesh, first: int = 0, count: Optional[int] = None):
        """Enqueue DrawTriangles + DrawLines without waiting (Sync() returns the summed info)."""
        dm = self.device_mesh(mesh)
        st, sh = self._state(), self._shader()
        tcount = dm.num_triangles - first if count is None else count
        _check(capi().fgl_draw_triangles_async(self._h, C.byref(st), C.byref(sh), dm.handle, first, tcount), self._h)
        if dm.num_lines and count is None:
            _check(capi().fgl_draw_lines_async(self._h, C.byref(st), C.byref(sh), dm.handle, 0, dm.num_lines), self._h)

    def Sync(self) -> RasterizeInfo:
        info = _Info()
        _check(capi().fgl_sync(self._h, C.byref(info)), self._h)
        return RasterizeInfo(info.total_pixels, info.updated_pixels)

    def FrameEnd(self, out: Optional[np.ndarray] = None, fence: Optional[Fence] = None) -> Fence:
        """fgl_frame_end: enqueue the read-back of the colour buffer into ``out`` ((H,W,4) uint8, pinned for
        a real overlap) and of the frame's RasterizeInfo, and return a fence; ``fence.wait()`` blocks until
        exactly this frame is in host memory while later frames keep running."""
        if out is not None:
            assert out.shape == (self.Height, self.Width, 4) and out.dtype == np.uint8 and out.flags.c_contiguous
        fence = Fence(self) if fence is None else fence
        _check(capi().fgl_frame_end(self._h, None if out is None else out.ctypes.data, 0, C.byref(fence.handle)), self._h)
        if fence._fin is None:
            fence._fin = weakref.finalize(fence, capi().fgl_fence_destroy, fence.handle)
        return fence

    def GraphBegin(self):
        """fgl_graph_begin: record the following clears / async draws / resolves instead of running them."""
        _check(capi().fgl_graph_begin(self._h), self._h)

    def GraphEnd(self) -> FrameGraph:
        h = _P()
        _check(capi().fgl_graph_end(self._h, C.byref(h)), self._h)
        return FrameGraph(self, h)

    def SetProfiling(self, enabled: bool):
        _check(capi().fgl_set_profiling(self._h, int(enabled)), self._h)

    def StageTimes(self) -> StageTimes:
        s = StageTimes()
        _check(capi().fgl_get_stage_times(self._h, C.byref(s)), self._h)
        return s

    def DrawStats(self) -> DrawStats:
        s = DrawStats()
        _check(capi().fgl_get_draw_stats(self._h, C.byref(s)), self._h)
        return s

    def ProbeAtomicRate(self, ops: int = 1 << 26) -> float:
        """64-bit atomicMin operations per second on a depth-buffer-sized array of this device: the
        denominator of the fragment-rate bound (see include/fauxgl_b200.h)."""
        r = C.c_double(0)
        _check(capi().fgl_probe_atomic_rate(self._h, int(ops), C.byref(r)), self._h)
        return float(r.value)

    def DivCheck(self, seed: int, pairs: int):
        """(mismatches, fast-path results) of the device's branch-free division helpers against the operator
        (fgl_debug_div_check, include/fauxgl_b200.h)."""
        bad, fast = C.c_uint64(0), C.c_uint64(0)
        _check(capi().fgl_debug_div_check(self._h, int(seed), int(pairs), C.byref(bad), C.byref(fast)), self._h)
        return int(bad.value), int(fast.value)

    # -- SSAA resolve (resize.Resize(..., resize.Bilinear) in the examples) -------------
    def Resolve(self, factor: int) -> np.ndarray:
        out = np.empty((self.Height // factor, self.Width // factor, 4), dtype=np.uint8)
        _check(capi().fgl_resolve(self._h, int(factor), out.ctypes.data), self._h)
        return out

    def ResolveDevice(self, factor: int):
        _check(capi().fgl_resolve_device(self._h, int(factor)), self._h)

    # -- sort-last composite -----------------------------------------------------------------
    def CompositePack(self, keys_dev_ptr: int):
        _check(capi().fgl_composite_pack(self._h, keys_dev_ptr), self._h)

    def CompositeUnpack(self, keys_dev_ptr: int):
        _check(capi().fgl_composite_unpack(self._h, keys_dev_ptr), self._h)

    def CompositeMin(self, inout_ptr: int, other_ptr: int, count: int):
        _check(capi().fgl_composite_min(self._h, inout_ptr, other_ptr, count), self._h)

    def IpcExport(self):
        """(color_handle, depth_handle): 64-byte CUDA IPC handles of this context's buffers."""
        hc, hd = C.create_string_buffer(64), C.create_string_buffer(64)
        _check(capi().fgl_ipc_export(self._h, hc, hd), self._h)
        return hc.raw, hd.raw

    def IpcOpen(self, color_handle: bytes, depth_handle: bytes):
        pc, pd = _P(), _P()
        _check(capi().fgl_ipc_open(self._h, color_handle, depth_handle, C.byref(pc), C.byref(pd)), self._h)
        return pc.value, pd.value

    def IpcClose(self, color_ptr: int, depth_ptr: int):
        _check(capi().fgl_ipc_close(self._h, color_ptr, depth_ptr), self._h)

    def CompositePeer(self, rank: int, color_ptrs, depth_ptrs):
        """One fused kernel: min-depth composite of this rank's stripe over all ranks' buffers (P2P)."""
        n = len(color_ptrs)
        ca = (_P * n)(*color_ptrs)
        da = (_P * n)(*depth_ptrs)
        _check(capi().fgl_composite_peer(self._h, int(rank), n, ca, da), self._h)

    # -- in-library composite (fgl_comm.cu) ---------------------------------------------------------
    def PeerExport(self) -> bytes:
        """fgl_peer_export: this context's record for fgl_peer_group_create (FGL_PEER_EXPORT_BYTES)."""
        rec = C.create_string_buffer(PEER_EXPORT_BYTES)
        _check(capi().fgl_peer_export(self._h, rec), self._h)
        return rec.raw

    # -- interop ---------------------------------------------------------------------------------
    @property
    def stream_ptr(self) -> int:
        return capi().fgl_stream(self._h) or 0

    @property
    def color_ptr(self) -> int:
        return capi().fgl_color_device_ptr(self._h) or 0

    @property
    def depth_ptr(self) -> int:
        return capi().fgl_depth_device_ptr(self._h) or 0


def NewContext(width: int, height: int, device: int = 0) -> Context:
    """context.go:60"""
    return Context(width, height, device)
