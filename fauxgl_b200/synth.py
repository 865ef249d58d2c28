"""Synthetic meshes for the benchmark configurations and the parity tests.

The reference's headline mesh (Stanford dragon, 871 306 triangles,
README.md:36) is not in its repository (examples/dragon.go:11-12 is a download
URL), so the benchmark uses the deterministic stand-ins SURVEY.md 8(d) defines.
A few of the reference's procedural builders (shapes.go) are mirrored because
its multi-pass examples are built from them.
"""
from __future__ import annotations

import math

import numpy as np

from .matrix import Scale
from .mesh import Mesh, NewLineMesh, NewTriangleMesh
from .vector import Vector

M871K_TRIANGLES = 871306  # README.md:36


def NewCube() -> Mesh:
    """shapes.go:17-39"""
    v = np.array([[-1, -1, -1], [-1, -1, 1], [-1, 1, -1], [-1, 1, 1],
                  [1, -1, -1], [1, -1, 1], [1, 1, -1], [1, 1, 1]], dtype=np.float64)
    idx = [(3, 5, 7), (5, 3, 1), (0, 6, 4), (6, 0, 2), (0, 5, 1), (5, 0, 4),
           (5, 6, 7), (6, 5, 4), (6, 3, 7), (3, 6, 2), (0, 3, 2), (3, 0, 1)]
    mesh = NewTriangleMesh(v[np.array(idx)])
    mesh.Transform(Scale(Vector(0.5, 0.5, 0.5)))
    return mesh


def NewCubeOutline(x0, y0, z0, x1, y1, z1) -> Mesh:
    """shapes.go:50-72 (NewCubeOutlineForBox)"""
    e = [((x0, y0, z0), (x0, y0, z1)), ((x0, y1, z0), (x0, y1, z1)), ((x1, y0, z0), (x1, y0, z1)),
         ((x1, y1, z0), (x1, y1, z1)), ((x0, y0, z0), (x0, y1, z0)), ((x0, y0, z1), (x0, y1, z1)),
         ((x1, y0, z0), (x1, y1, z0)), ((x1, y0, z1), (x1, y1, z1)), ((x0, y0, z0), (x1, y0, z0)),
         ((x0, y1, z0), (x1, y1, z0)), ((x0, y0, z1), (x1, y0, z1)), ((x0, y1, z1), (x1, y1, z1))]
    return NewLineMesh(np.array(e, dtype=np.float64))


def _latlng(lat, lng):
    """util.go:23-29"""
    lat, lng = lat * math.pi / 180, lng * math.pi / 180
    return (math.cos(lat) * math.cos(lng), math.cos(lat) * math.sin(lng), math.sin(lat))


def NewLatLngSphere(latStep: int, lngStep: int) -> Mesh:
    """shapes.go:74-110 (positions + texture coordinates; face normals via FixNormals)."""
    P, U = [], []
    for lat0 in range(-90, 90, latStep):
        lat1 = lat0 + latStep
        v0, v1 = (lat0 + 90) / 180, (lat1 + 90) / 180
        for lng0 in range(-180, 180, lngStep):
            lng1 = lng0 + lngStep
            u0, u1 = (lng0 + 180) / 360, (lng1 + 180) / 360
            if lng1 >= 180:
                lng1 -= 360
            p00, p01 = _latlng(lat0, lng0), _latlng(lat0, lng1)
            p10, p11 = _latlng(lat1, lng0), _latlng(lat1, lng1)
            if lat0 != -90:
                P.append((p00, p01, p11)); U.append(((u0, v0, 0), (u1, v0, 0), (u1, v1, 0)))
            if lat1 != 90:
                P.append((p00, p11, p10)); U.append(((u0, v0, 0), (u1, v1, 0), (u0, v1, 0)))
    return NewTriangleMesh(np.array(P, dtype=np.float64), texture=np.array(U, dtype=np.float64))


def _grid_surface(radius_fn, nu: int, nv: int, scale=(1.0, 1.0, 1.0)):
    """Closed surface r(u,v) on a nu x nv longitude/latitude grid, shared grid
    vertices bit-identical, pole rows collapsed to single triangles, counter-clockwise
    seen from outside.  Returns position[T,3,3] with T = nu*(2*nv-2)."""
    u = np.arange(nu, dtype=np.float64) * (2 * math.pi / nu)
    v = np.arange(nv + 1, dtype=np.float64) * (math.pi / nv)
    uu, vv = np.meshgrid(u, v, indexing="ij")          # (nu, nv+1)
    r = radius_fn(uu, vv)
    sv = np.sin(vv)
    sv[:, 0] = 0.0
    sv[:, nv] = 0.0                                     # exact poles: one shared vertex each
    grid = np.stack([r * sv * np.cos(uu) * scale[0], r * np.cos(vv) * scale[1], r * sv * np.sin(uu) * scale[2]], axis=-1)
    grid[:, 0] = grid[0, 0]
    grid[:, nv] = grid[0, nv]
    i0 = np.arange(nu)
    i1 = (i0 + 1) % nu
    tris = []
    for j in range(nv):
        a, b = grid[i0, j], grid[i1, j]                  # row j
        c, d = grid[i0, j + 1], grid[i1, j + 1]          # row j+1
        if j == 0:
            tris.append(np.stack([a, d, c], axis=1))
        elif j == nv - 1:
            tris.append(np.stack([a, b, c], axis=1))
        else:
            quad = np.empty((nu, 2, 3, 3), dtype=np.float64)
            quad[:, 0] = np.stack([a, b, d], axis=1)
            quad[:, 1] = np.stack([a, d, c], axis=1)
            tris.append(quad.reshape(-1, 3, 3))
    return np.concatenate(tris, axis=0)


def bumpy_surface(triangles: int = M871K_TRIANGLES, nu: int = 661, nv: int = 661, smooth: bool = True) -> Mesh:
    """M871k: bumpy closed surface, 661 x 660 quads -> 872 520
    triangles, first 871 306 kept; BiUnitCube; SmoothNormals.  Deterministic."""
    def radius(u, v):
        return (1 + 0.15 * np.sin(7 * u) * np.sin(5 * v) + 0.05 * np.sin(23 * u + 1) * np.sin(17 * v)
                + 0.02 * np.sin(61 * u) * np.sin(47 * v))
    pos = _grid_surface(radius, nu, nv, scale=(1.6, 0.9, 0.7))
    assert len(pos) >= triangles, (len(pos), triangles)
    mesh = NewTriangleMesh(pos[:triangles])
    mesh.BiUnitCube()
    if smooth:
        mesh.SmoothNormals()
    return mesh


def uv_sphere(nu: int, nv: int) -> Mesh:
    """M10M-style unit sphere with analytic (position) normals; nu*(2*nv-2) triangles."""
    pos = _grid_surface(lambda u, v: np.ones_like(u), nu, nv)
    with np.errstate(invalid="ignore", divide="ignore"):
        nrm = pos / np.sqrt((pos * pos).sum(axis=-1, keepdims=True))
    return Mesh(pos, nrm)
