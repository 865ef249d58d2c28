"""Host-side mesh container and loaders, mirroring the reference's ``mesh.go``,
``triangle.go``, ``stl.go`` and ``obj.go`` for the inputs the hot path consumes.

The reference keeps ``[]*Triangle`` (408-byte structs behind pointers,
mesh.go:9-13, triangle.go:3-5).  Here a mesh is attribute-major numpy float64:
``position[T,3,3]``, ``normal[T,3,3]``, ``texture[T,3,3]``, ``color[T,3,4]``
(lines: ``[L,2,k]``) -- the layout ``fgl_mesh_create`` takes; the device
transposes it into planar SoA.  Mesh preparation (BiUnitCube, normal smoothing)
happens once per mesh on the host and is not part of the hot path, but it is
restated exactly (same evaluation order, same list order) so that the oracle and
the GPU back end see the inputs the reference's examples would produce.
"""
from __future__ import annotations

import math
import struct
from typing import Optional

import numpy as np

from .matrix import Identity, Matrix
from .vector import Vector
from .color import Color

_F = np.float64


def _normalize_rows(v: np.ndarray) -> np.ndarray:
    """vector.go:83-86 on an (...,3) array: r = 1/sqrt((x*x+y*y)+z*z); v*r."""
    x, y, z = v[..., 0], v[..., 1], v[..., 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        r = 1.0 / np.sqrt(x * x + y * y + z * z)
        return np.stack([x * r, y * r, z * r], axis=-1)


def _face_normals(position: np.ndarray) -> np.ndarray:
    """triangle.go:33-37 for every triangle: normalize((p2-p1) x (p3-p1))."""
    e1 = position[:, 1] - position[:, 0]
    e2 = position[:, 2] - position[:, 0]
    n = np.stack([
        e1[:, 1] * e2[:, 2] - e1[:, 2] * e2[:, 1],
        e1[:, 2] * e2[:, 0] - e1[:, 0] * e2[:, 2],
        e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]], axis=-1)
    return _normalize_rows(n)


def _mul_position(m: Matrix, p: np.ndarray) -> np.ndarray:
    """matrix.go:209-214 on an (...,3) array."""
    x, y, z = p[..., 0], p[..., 1], p[..., 2]
    return np.stack([
        m[0] * x + m[1] * y + m[2] * z + m[3],
        m[4] * x + m[5] * y + m[6] * z + m[7],
        m[8] * x + m[9] * y + m[10] * z + m[11]], axis=-1)


def _mul_direction(m: Matrix, p: np.ndarray) -> np.ndarray:
    """matrix.go:224-229 on an (...,3) array."""
    x, y, z = p[..., 0], p[..., 1], p[..., 2]
    return _normalize_rows(np.stack([
        m[0] * x + m[1] * y + m[2] * z,
        m[4] * x + m[5] * y + m[6] * z,
        m[8] * x + m[9] * y + m[10] * z], axis=-1))


class Box:
    """box.go:7-9"""

    def __init__(self, mn: Vector, mx: Vector):
        self.Min, self.Max = mn, mx

    def Size(self) -> Vector:  # box.go:45
        return self.Max.Sub(self.Min)


class Mesh:
    """mesh.go:9-13 (Triangles + Lines), attribute-major."""

    def __init__(self, position=None, normal=None, texture=None, color=None,
                 lposition=None, lnormal=None, ltexture=None, lcolor=None):
        def arr(a, n, k, shape_mid):
            if a is None:
                return np.zeros((n, shape_mid, k), dtype=_F)
            a = np.ascontiguousarray(a, dtype=_F)
            assert a.shape == (n, shape_mid, k), (a.shape, (n, shape_mid, k))
            return a
        T = 0 if position is None else len(position)
        L = 0 if lposition is None else len(lposition)
        self.position = arr(position, T, 3, 3)
        self.normal = arr(normal, T, 3, 3)
        self.texture = arr(texture, T, 3, 3)
        self.color = arr(color, T, 4, 3)
        self.lposition = arr(lposition, L, 3, 2)
        self.lnormal = arr(lnormal, L, 3, 2)
        self.ltexture = arr(ltexture, L, 3, 2)
        self.lcolor = arr(lcolor, L, 4, 2)
        self.generation = 0  # bumped by every mutator; device copies key on it

    # -- sizes -----------------------------------------------------------------
    @property
    def num_triangles(self) -> int:
        return len(self.position)

    @property
    def num_lines(self) -> int:
        return len(self.lposition)

    def indexed(self):
        """(v [n][3], vn [n][3], corners [T][3][3] int32): the triangle soup as shared-vertex tables, the form
        fgl_mesh_create_indexed / fgl_mesh_update_indexed_async take (obj.go:19-79 parses such tables from a file).
        Two corners share an entry when position AND normal are bit-identical, so expanding the tables gives the
        soup back exactly.  The vt index of every corner is 0 (pass a single zero row as the vt table)."""
        T = self.num_triangles
        rec = np.ascontiguousarray(np.concatenate([self.position.reshape(T * 3, 3), self.normal.reshape(T * 3, 3)], axis=1))
        key = rec.view(np.dtype((np.void, rec.dtype.itemsize * 6))).ravel()
        _, first, inverse = np.unique(key, return_index=True, return_inverse=True)
        table = rec[first]
        corners = np.zeros((T, 3, 3), dtype=np.int32)
        corners[:, :, 0] = inverse.reshape(T, 3)
        corners[:, :, 2] = inverse.reshape(T, 3)
        return np.ascontiguousarray(table[:, :3]), np.ascontiguousarray(table[:, 3:]), corners

    def Invalidate(self):
        """Call after poking the arrays directly (the reference lets users write
        ``mesh.Triangles[i].V1.Color = ...``; here that needs a re-upload)."""
        self.generation += 1

    # -- oracle interchange: the Go struct layout -----------------------------------
    def triangle_vertices(self) -> np.ndarray:
        """(T,3,17) float64 in the reference's Vertex layout (vertex.go:3-12)."""
        T = self.num_triangles
        out = np.zeros((T, 3, 17), dtype=_F)
        out[:, :, 0:3] = self.position
        out[:, :, 3:6] = self.normal
        out[:, :, 6:9] = self.texture
        out[:, :, 9:13] = self.color
        return out

    def line_vertices(self) -> np.ndarray:
        L = self.num_lines
        out = np.zeros((L, 2, 17), dtype=_F)
        out[:, :, 0:3] = self.lposition
        out[:, :, 3:6] = self.lnormal
        out[:, :, 6:9] = self.ltexture
        out[:, :, 9:13] = self.lcolor
        return out

    # -- mesh.go ---------------------------------------------------------------------
    def Copy(self) -> "Mesh":  # mesh.go:34-46
        return Mesh(self.position.copy(), self.normal.copy(), self.texture.copy(), self.color.copy(),
                    self.lposition.copy(), self.lnormal.copy(), self.ltexture.copy(), self.lcolor.copy())

    def Add(self, b: "Mesh"):  # mesh.go:48-52
        for name in ("position", "normal", "texture", "color",
                     "lposition", "lnormal", "ltexture", "lcolor"):
            setattr(self, name, np.concatenate([getattr(self, name), getattr(b, name)], axis=0))
        self.generation += 1

    def SetColor(self, c: Color):  # mesh.go:54-58 (triangles only)
        self.color[:, :, :] = np.array(c, dtype=_F)
        self.generation += 1

    def BoundingBox(self) -> Box:  # mesh.go:153-165
        pts = [self.position.reshape(-1, 3), self.lposition.reshape(-1, 3)]
        p = np.concatenate(pts, axis=0)
        if len(p) == 0:
            return Box(Vector(0.0, 0.0, 0.0), Vector(0.0, 0.0, 0.0))
        mn, mx = p.min(axis=0), p.max(axis=0)
        return Box(Vector(*map(float, mn)), Vector(*map(float, mx)))

    def Transform(self, matrix: Matrix):  # mesh.go:167-175, triangle.go:66-73, line.go:23-28
        if self.num_triangles:
            self.position = _mul_position(matrix, self.position)
            self.normal = _mul_direction(matrix, self.normal)
        if self.num_lines:
            self.lposition = _mul_position(matrix, self.lposition)
            self.lnormal = _mul_direction(matrix, self.lnormal)
        self.generation += 1

    def FitInside(self, box: Box, anchor: Vector) -> Matrix:  # mesh.go:142-151
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
        r = 1.0
        return self.FitInside(Box(Vector(-r, -r, -r), Vector(r, r, r)), Vector(0.5, 0.5, 0.5))

    def UnitCube(self) -> Matrix:  # mesh.go:122-125
        r = 0.5
        return self.FitInside(Box(Vector(-r, -r, -r), Vector(r, r, r)), Vector(0.5, 0.5, 0.5))

    def _corner_groups(self):
        """Group the 3T corners by exact position (Go map[Vector] semantics:
        +0 == -0), preserving corner order (t0.V1, t0.V2, t0.V3, t1.V1, ...)."""
        pos = self.position.reshape(-1, 3) + 0.0  # -0 -> +0
        key = np.ascontiguousarray(pos).view(np.dtype((np.void, 24))).ravel()
        _, inverse, counts = np.unique(key, return_inverse=True, return_counts=True)
        order = np.argsort(inverse, kind="stable")
        starts = np.concatenate([[0], np.cumsum(counts)[:-1]])
        return inverse, counts, order, starts

    def SmoothNormalsThreshold(self, radians: float):  # mesh.go:80-103
        threshold = math.cos(radians)
        nrm = self.normal.reshape(-1, 3)
        inverse, counts, order, starts = self._corner_groups()
        out = np.empty_like(nrm)
        for k in np.unique(counts):
            groups = np.nonzero(counts == k)[0]
            # members[g, j] = corner index of the j-th list entry of group g
            members = order[starts[groups][:, None] + np.arange(k)[None, :]]
            mn = nrm[members]  # (G,k,3)
            for i in range(k):
                ni = mn[:, i]
                acc = np.zeros((len(groups), 3), dtype=_F)
                for j in range(k):
                    x = mn[:, j]
                    dot = x[:, 0] * ni[:, 0] + x[:, 1] * ni[:, 1] + x[:, 2] * ni[:, 2]
                    keep = dot >= threshold
                    acc = np.where(keep[:, None], acc + x, acc)
                out[members[:, i]] = _normalize_rows(acc)
        # a position with a NaN component is a map key that is never found again: the corner's list is nil and
        # Vector{}.Normalize() = 0 * (1/0) = NaN (mesh.go:80-88)
        nanpos = np.isnan(self.position.reshape(-1, 3)).any(axis=1)
        if nanpos.any():
            out[nanpos] = _normalize_rows(np.zeros((int(nanpos.sum()), 3), dtype=_F))  # 0 * Inf, as computed
        self.normal = out.reshape(-1, 3, 3)
        self.generation += 1

    def SmoothNormals(self):  # mesh.go:105-120
        nrm = self.normal.reshape(-1, 3)
        inverse, counts, order, starts = self._corner_groups()
        out = np.empty_like(nrm)
        for k in np.unique(counts):
            groups = np.nonzero(counts == k)[0]
            members = order[starts[groups][:, None] + np.arange(k)[None, :]]
            mn = nrm[members]
            acc = np.zeros((len(groups), 3), dtype=_F)
            for j in range(k):
                acc = acc + mn[:, j]
            acc = _normalize_rows(acc)
            for i in range(k):
                out[members[:, i]] = acc
        # a position with a NaN component is a map key that is never found again (NaN != NaN): the final
        # lookup[t.V1.Position] returns the zero Vector (mesh.go:115-119)
        out[np.isnan(self.position.reshape(-1, 3)).any(axis=1)] = 0.0
        self.normal = out.reshape(-1, 3, 3)
        self.generation += 1


def NewTriangleMesh(position, normal=None, texture=None, color=None, fix_normals=True) -> Mesh:
    """mesh.go:19-21; zero normals are replaced by the face normal, as
    Triangle.FixNormals does on every loader path (triangle.go:46-58)."""
    position = np.ascontiguousarray(position, dtype=_F)
    T = len(position)
    if normal is None:
        normal = np.zeros((T, 3, 3), dtype=_F)
    normal = np.array(normal, dtype=_F)
    if fix_normals and T:
        zero = np.all(normal == 0, axis=2)  # (T,3)
        if zero.any():
            fn = _face_normals(position)
            normal = np.where(zero[:, :, None], fn[:, None, :], normal)
    return Mesh(position, normal, texture, color)


def NewLineMesh(lposition, lnormal=None, ltexture=None, lcolor=None) -> Mesh:
    """mesh.go:23-25"""
    return Mesh(lposition=lposition, lnormal=lnormal, ltexture=ltexture, lcolor=lcolor)


# -- stl.go ---------------------------------------------------------------------------

def LoadSTL(path: str) -> Mesh:  # stl.go:23-57
    with open(path, "rb") as f:
        data = f.read()
    if len(data) >= 84:
        count = struct.unpack_from("<I", data, 80)[0]
        if len(data) == count * 50 + 84:
            return _load_stl_binary(data, count)
    return _load_stl_ascii(data.decode("utf-8", errors="replace"))


def _load_stl_ascii(text: str) -> Mesh:  # stl.go:59-80
    verts = []
    for line in text.splitlines():
        fields = line.split()
        if len(fields) == 4 and fields[0] == "vertex":
            verts.append((_parse_float(fields[1]), _parse_float(fields[2]), _parse_float(fields[3])))
    n = len(verts) // 3
    position = np.array(verts[:3 * n], dtype=_F).reshape(n, 3, 3)
    return NewTriangleMesh(position)


def _load_stl_binary(data: bytes, count: int) -> Mesh:  # stl.go:86-154
    rec = np.frombuffer(data, dtype=np.uint8, count=count * 50, offset=84).reshape(count, 50)
    f32 = np.ascontiguousarray(rec[:, 12:48]).view("<f4").reshape(count, 3, 3)
    position = f32.astype(_F)  # float32 widened to float64, stl.go:82-84
    fn = _face_normals(position)
    normal = np.repeat(fn[:, None, :], 3, axis=1)
    return Mesh(position, normal)


def _parse_float(s: str) -> float:  # util.go:65-72 (errors -> 0)
    try:
        return float(s)
    except ValueError:
        return 0.0


# -- obj.go ---------------------------------------------------------------------------

def ParseOBJ(path: str):  # obj.go:19-79, the text part
    """Parse an OBJ file into the loader's tables: ``(vs, vts, vns, corners)`` with vs/vts/vns float64 (n,3)
    arrays whose entry 0 is the zero vector of the reference's 1-based tables (obj.go:26-28) and corners an
    int32 (T,3,3) array of (v, vt, vn) table indices per triangle corner, polygons triangulated as fans
    (obj.go:58-60).  ``LoadOBJ`` expands it on the host, ``DeviceMesh.FromOBJ`` on the device."""
    vs = [(0.0, 0.0, 0.0)]
    vts = [(0.0, 0.0, 0.0)]
    vns = [(0.0, 0.0, 0.0)]
    corners = []

    def parse_index(value: str, length: int) -> int:  # obj.go:10-17
        try:
            n = int(value, 0) if value else 0
        except ValueError:
            n = 0
        if n < 0:
            n += length
        return n

    with open(path, "r", errors="replace") as f:
        for line in f:
            fields = line.split()
            if not fields:
                continue
            kw, args = fields[0], fields[1:]
            if kw == "v":
                fl = [_parse_float(a) for a in args]
                vs.append((fl[0], fl[1], fl[2]))
            elif kw == "vt":
                fl = [_parse_float(a) for a in args]
                vts.append((fl[0], fl[1], 0.0))
            elif kw == "vn":
                fl = [_parse_float(a) for a in args]
                vns.append((fl[0], fl[1], fl[2]))
            elif kw == "f":
                fv, ft, fn = [], [], []
                for arg in args:
                    vertex = (arg + "//").split("/")
                    fv.append(parse_index(vertex[0], len(vs)))
                    ft.append(parse_index(vertex[1], len(vts)))
                    fn.append(parse_index(vertex[2], len(vns)))
                for i in range(1, len(fv) - 1):
                    corners.append([(fv[k], ft[k], fn[k]) for k in (0, i, i + 1)])
    for tri in corners:  # Go panics on an index outside its table; so does this loader
        for iv, it, inn in tri:
            if not (0 <= iv < len(vs) and 0 <= it < len(vts) and 0 <= inn < len(vns)):
                raise IndexError("OBJ face index out of range")
    return (np.array(vs, dtype=_F).reshape(-1, 3), np.array(vts, dtype=_F).reshape(-1, 3),
            np.array(vns, dtype=_F).reshape(-1, 3), np.array(corners, dtype=np.int32).reshape(-1, 3, 3))


def LoadOBJ(path: str) -> Mesh:  # obj.go:19-79
    vs, vts, vns, corners = ParseOBJ(path)
    position = vs[corners[:, :, 0]]
    texture = vts[corners[:, :, 1]]
    normal = vns[corners[:, :, 2]]
    return NewTriangleMesh(position, normal, texture)


def LoadMesh(path: str) -> Mesh:  # util.go:31-44 (.ply/.3ds loaders are out of scope)
    ext = path.lower().rsplit(".", 1)[-1]
    if ext == "stl":
        return LoadSTL(path)
    if ext == "obj":
        return LoadOBJ(path)
    raise ValueError("unrecognized mesh extension: ." + ext)
