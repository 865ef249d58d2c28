"""ctypes front end of oracle/libfauxgl_oracle.so (the C restatement of the
reference's DrawMesh path).  TEST INFRASTRUCTURE ONLY -- see fauxgl_oracle.h.

``OracleContext`` mirrors the slice of the reference's Context API the parity
tests drive (context.go:40-145, 391-439) so a test can run the same script
against the oracle and against the GPU back end.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfauxgl_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "fauxgl_oracle.c")
    hdr = os.path.join(_HERE, "fauxgl_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)
             or any(os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr)))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libfauxgl_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


class OShader(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("has_texture", C.c_int32),
        ("matrix", C.c_double * 16), ("light", C.c_double * 3), ("camera", C.c_double * 3),
        ("object", C.c_double * 4), ("ambient", C.c_double * 4), ("diffuse", C.c_double * 4),
        ("specular", C.c_double * 4), ("specular_power", C.c_double), ("color", C.c_double * 4),
        ("tex", C.c_void_p), ("tex_w", C.c_int32), ("tex_h", C.c_int32),
        ("tex_format", C.c_int32), ("_pad", C.c_int32),
    ]


class OCtx(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32),
        ("color", C.c_void_p), ("depth", C.c_void_p),
        ("read_depth", C.c_int32), ("write_depth", C.c_int32), ("write_color", C.c_int32),
        ("alpha_blend", C.c_int32), ("wireframe", C.c_int32),
        ("front_face", C.c_int32), ("cull", C.c_int32), ("x_guard", C.c_int32),
        ("line_width", C.c_double), ("depth_bias", C.c_double),
    ]


class OInfo(C.Structure):
    _fields_ = [("total_pixels", C.c_uint64), ("updated_pixels", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_clear_color.argtypes = [C.POINTER(OCtx), C.POINTER(C.c_double)]
        L.oracle_clear_depth.argtypes = [C.POINTER(OCtx), C.c_double]
        L.oracle_draw_triangles.argtypes = [C.POINTER(OCtx), C.POINTER(OShader), C.c_void_p, C.c_size_t,
                                            C.c_int, C.POINTER(OInfo)]
        L.oracle_draw_lines.argtypes = L.oracle_draw_triangles.argtypes
        L.oracle_setup_triangle.argtypes = [C.POINTER(OCtx), C.POINTER(OShader), C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_size_t]
        L.oracle_setup_triangle.restype = C.c_size_t
        L.oracle_fragment.argtypes = [C.POINTER(OShader), C.c_void_p, C.POINTER(C.c_double)]
        L.oracle_pow.argtypes = [C.c_double, C.c_double]
        L.oracle_pow.restype = C.c_double
        L.oracle_resolve.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.oracle_pack_key.argtypes = [C.c_double, C.c_void_p]
        L.oracle_pack_key.restype = C.c_uint64
        L.oracle_draw_each.argtypes = [C.POINTER(OCtx), C.POINTER(OShader), C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        L.oracle_depth_image.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.oracle_stl_triangles.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        _lib = L
    return _lib


def make_oshader(shader) -> tuple:
    """Translate a fauxgl_b200.shader object (duck-typed: .describe()) into an
    OShader.  Returns (OShader, keepalive)."""
    d = shader.describe()
    s = OShader()
    s.kind = d["kind"]
    s.matrix[:] = d["matrix"]
    s.light[:] = d.get("light", (0, 0, 0))
    s.camera[:] = d.get("camera", (0, 0, 0))
    s.object[:] = d.get("object", (0, 0, 0, 0))
    s.ambient[:] = d.get("ambient", (0, 0, 0, 0))
    s.diffuse[:] = d.get("diffuse", (0, 0, 0, 0))
    s.specular[:] = d.get("specular", (0, 0, 0, 0))
    s.specular_power = d.get("specular_power", 0.0)
    s.color[:] = d.get("color", (0, 0, 0, 0))
    tex = d.get("texture")
    keep = None
    if tex is not None:
        keep = np.ascontiguousarray(tex.pixels)   # uint8 texels, or uint16 for O_TEX_RGBA64
        s.has_texture = 1
        s.tex = keep.ctypes.data
        s.tex_h, s.tex_w = keep.shape[0], keep.shape[1]
        s.tex_format = tex.format
    return s, keep


class OracleContext:
    """Reference-shaped context running on the CPU oracle (context.go:40-81)."""

    def __init__(self, width: int, height: int, x_guard: bool = False, threads: int = 1):
        self.Width, self.Height = width, height
        self.ColorBuffer = np.zeros((height, width, 4), dtype=np.uint8)   # image.NewNRGBA: zeroed
        self.DepthBuffer = np.empty((height, width), dtype=np.float64)
        self.ClearColor = (0.0, 0.0, 0.0, 0.0)
        self.Shader = None
        self.ReadDepth = True
        self.WriteDepth = True
        self.WriteColor = True
        self.AlphaBlend = True
        self.Wireframe = False
        self.FrontFace = 2   # FaceCCW
        self.Cull = 3        # CullBack
        self.LineWidth = 2.0
        self.DepthBias = 0.0
        self.XGuard = x_guard           # False: the reference's rule (context.go:223-228); True: drop x outside [0, W)
        self.threads = threads
        self.ClearDepthBuffer()

    def _ctx(self) -> OCtx:
        c = OCtx()
        c.width, c.height = self.Width, self.Height
        c.color = self.ColorBuffer.ctypes.data
        c.depth = self.DepthBuffer.ctypes.data
        c.read_depth, c.write_depth, c.write_color = int(self.ReadDepth), int(self.WriteDepth), int(self.WriteColor)
        c.alpha_blend, c.wireframe = int(self.AlphaBlend), int(self.Wireframe)
        c.front_face, c.cull = int(self.FrontFace), int(self.Cull)
        c.x_guard = int(self.XGuard)
        c.line_width, c.depth_bias = float(self.LineWidth), float(self.DepthBias)
        return c

    def ClearColorBufferWith(self, color):
        c = self._ctx()
        lib().oracle_clear_color(C.byref(c), (C.c_double * 4)(*color))

    def ClearColorBuffer(self):
        self.ClearColorBufferWith(self.ClearColor)

    def ClearDepthBufferWith(self, value: float):
        c = self._ctx()
        lib().oracle_clear_depth(C.byref(c), float(value))

    def ClearDepthBuffer(self):
        self.ClearDepthBufferWith(np.finfo(np.float64).max)

    def _draw(self, fn, verts: np.ndarray):
        c = self._ctx()
        s, keep = make_oshader(self.Shader)
        info = OInfo()
        verts = np.ascontiguousarray(verts, dtype=np.float64)
        fn(C.byref(c), C.byref(s), verts.ctypes.data, len(verts), int(self.threads), C.byref(info))
        del keep
        return (info.total_pixels, info.updated_pixels)

    def DrawTriangles(self, mesh, first=0, count=None):
        v = mesh.triangle_vertices()
        count = len(v) - first if count is None else count
        return self._draw(lib().oracle_draw_triangles, v[first:first + count])

    def DrawLines(self, mesh, first=0, count=None):
        v = mesh.line_vertices()
        count = len(v) - first if count is None else count
        return self._draw(lib().oracle_draw_lines, v[first:first + count])

    def DrawMesh(self, mesh):
        a = self.DrawTriangles(mesh)
        b = self.DrawLines(mesh)
        return (a[0] + b[0], a[1] + b[1])

    def _draw_each(self, verts: np.ndarray, is_lines: bool) -> np.ndarray:
        c = self._ctx()
        s, keep = make_oshader(self.Shader)
        verts = np.ascontiguousarray(verts, dtype=np.float64)
        infos = np.zeros((len(verts), 2), dtype=np.uint64)
        lib().oracle_draw_each(C.byref(c), C.byref(s), verts.ctypes.data, len(verts), int(is_lines), infos.ctypes.data)
        del keep
        return infos

    def DrawLinesEach(self, mesh, first=0, count=None) -> np.ndarray:
        """[(TotalPixels, UpdatedPixels)] of Context.DrawLine per line, context.go:351-368."""
        v = mesh.line_vertices()
        count = len(v) - first if count is None else count
        return self._draw_each(v[first:first + count], True)

    def DrawTrianglesEach(self, mesh, first=0, count=None) -> np.ndarray:
        """[(TotalPixels, UpdatedPixels)] of Context.DrawTriangle per triangle, context.go:370-389."""
        v = mesh.triangle_vertices()
        count = len(v) - first if count is None else count
        return self._draw_each(v[first:first + count], False)

    def Image(self) -> np.ndarray:
        return self.ColorBuffer

    def DepthImage(self) -> np.ndarray:  # context.go:87-117
        out = np.empty((self.Height, self.Width), dtype=np.uint16)
        d = np.ascontiguousarray(self.DepthBuffer, dtype=np.float64)
        lib().oracle_depth_image(d.ctypes.data, self.Width, self.Height, out.ctypes.data)
        return out

    def Resolve(self, factor: int) -> np.ndarray:
        dw, dh = self.Width // factor, self.Height // factor
        out = np.empty((dh, dw, 4), dtype=np.uint8)
        lib().oracle_resolve(self.ColorBuffer.ctypes.data, self.Width, self.Height, dw, dh, out.ctypes.data)
        return out


def resolve(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    src = np.ascontiguousarray(src, dtype=np.uint8)
    out = np.empty((dh, dw, 4), dtype=np.uint8)
    lib().oracle_resolve(src.ctypes.data, src.shape[1], src.shape[0], dw, dh, out.ctypes.data)
    return out


def depth_image(depth: np.ndarray) -> np.ndarray:
    d = np.ascontiguousarray(depth, dtype=np.float64)
    out = np.empty(d.shape, dtype=np.uint16)
    lib().oracle_depth_image(d.ctypes.data, d.shape[1], d.shape[0], out.ctypes.data)
    return out


def stl_triangles(records: bytes):
    """(position [n,3,3], normal [n,3,3]) of loadSTLB (stl.go:86-154) for n 50-byte records."""
    n = len(records) // 50
    buf = np.frombuffer(records, dtype=np.uint8, count=n * 50).copy()
    out = np.zeros((n * 3, 17), dtype=np.float64)
    lib().oracle_stl_triangles(buf.ctypes.data, n, out.ctypes.data)
    out = out.reshape(n, 3, 17)
    return out[:, :, 0:3].copy(), out[:, :, 3:6].copy()


def pack_key(depth: float, rgba) -> int:
    a = (C.c_uint8 * 4)(*rgba)
    return int(lib().oracle_pack_key(float(depth), a))
