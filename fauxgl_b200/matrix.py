"""Host-side 4x4 float64 matrices, mirroring the reference's ``matrix.go``.

Builders (LookAt, Perspective, Rotate, ...) are evaluated once per frame on the
host and handed to the device as 16 row-major doubles (X00..X33,
matrix.go:5-10).  Expressions keep the reference's evaluation order.
"""
from __future__ import annotations

import math
from typing import Tuple

from .vector import Vector


class Matrix(tuple):
    """Row-major 16-tuple X00..X33 (matrix.go:5-10)."""

    def __new__(cls, *vals):
        if len(vals) == 1:
            vals = tuple(vals[0])
        if len(vals) != 16:
            raise ValueError("Matrix needs 16 values")
        return super().__new__(cls, (float(v) for v in vals))

    # -- products ---------------------------------------------------------
    def Mul(a, b):  # matrix.go:188-207
        m = [0.0] * 16
        for r in range(4):
            for c in range(4):
                m[4 * r + c] = (a[4 * r] * b[c] + a[4 * r + 1] * b[4 + c]
                                + a[4 * r + 2] * b[8 + c] + a[4 * r + 3] * b[12 + c])
        return Matrix(m)

    def MulPosition(a, b: Vector) -> Vector:  # matrix.go:209
        x = a[0] * b.X + a[1] * b.Y + a[2] * b.Z + a[3]
        y = a[4] * b.X + a[5] * b.Y + a[6] * b.Z + a[7]
        z = a[8] * b.X + a[9] * b.Y + a[10] * b.Z + a[11]
        return Vector(x, y, z)

    def MulPositionW(a, b: Vector) -> Tuple[float, float, float, float]:  # matrix.go:216
        x = a[0] * b.X + a[1] * b.Y + a[2] * b.Z + a[3]
        y = a[4] * b.X + a[5] * b.Y + a[6] * b.Z + a[7]
        z = a[8] * b.X + a[9] * b.Y + a[10] * b.Z + a[11]
        w = a[12] * b.X + a[13] * b.Y + a[14] * b.Z + a[15]
        return (x, y, z, w)

    def MulDirection(a, b: Vector) -> Vector:  # matrix.go:224
        x = a[0] * b.X + a[1] * b.Y + a[2] * b.Z
        y = a[4] * b.X + a[5] * b.Y + a[6] * b.Z
        z = a[8] * b.X + a[9] * b.Y + a[10] * b.Z
        return Vector(x, y, z).Normalize()

    # -- method forms left-multiply (matrix.go:143-177) ------------------------
    def Translate(m, v):
        return Translate(v).Mul(m)

    def Scale(m, v):
        return Scale(v).Mul(m)

    def Rotate(m, v, a):
        return Rotate(v, a).Mul(m)

    def Frustum(m, l, r, b, t, n, f):
        return Frustum(l, r, b, t, n, f).Mul(m)

    def Orthographic(m, l, r, b, t, n, f):
        return Orthographic(l, r, b, t, n, f).Mul(m)

    def Perspective(m, fovy, aspect, near, far):
        return Perspective(fovy, aspect, near, far).Mul(m)

    def LookAt(m, eye, center, up):
        return LookAt(eye, center, up).Mul(m)


def Identity() -> Matrix:  # matrix.go:12
    return Matrix(1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1)


def Translate(v: Vector) -> Matrix:  # matrix.go:20
    return Matrix(1, 0, 0, v.X, 0, 1, 0, v.Y, 0, 0, 1, v.Z, 0, 0, 0, 1)


def Scale(v: Vector) -> Matrix:  # matrix.go:28
    return Matrix(v.X, 0, 0, 0, 0, v.Y, 0, 0, 0, 0, v.Z, 0, 0, 0, 0, 1)


def Rotate(v: Vector, a: float) -> Matrix:  # matrix.go:36
    v = v.Normalize()
    s = math.sin(a)
    c = math.cos(a)
    m = 1 - c
    return Matrix(
        m * v.X * v.X + c, m * v.X * v.Y + v.Z * s, m * v.Z * v.X - v.Y * s, 0,
        m * v.X * v.Y - v.Z * s, m * v.Y * v.Y + c, m * v.Y * v.Z + v.X * s, 0,
        m * v.Z * v.X + v.Y * s, m * v.Y * v.Z - v.X * s, m * v.Z * v.Z + c, 0,
        0, 0, 0, 1)


def Frustum(l, r, b, t, n, f) -> Matrix:  # matrix.go:69
    t1 = 2 * n
    t2 = r - l
    t3 = t - b
    t4 = f - n
    return Matrix(
        t1 / t2, 0, (r + l) / t2, 0,
        0, t1 / t3, (t + b) / t3, 0,
        0, 0, (-f - n) / t4, (-t1 * f) / t4,
        0, 0, -1, 0)


def Orthographic(l, r, b, t, n, f) -> Matrix:  # matrix.go:81
    return Matrix(
        2 / (r - l), 0, 0, -(r + l) / (r - l),
        0, 2 / (t - b), 0, -(t + b) / (t - b),
        0, 0, -2 / (f - n), -(f + n) / (f - n),
        0, 0, 0, 1)


def Perspective(fovy, aspect, near, far) -> Matrix:  # matrix.go:89
    ymax = near * math.tan(fovy * math.pi / 360)
    xmax = ymax * aspect
    return Frustum(-xmax, xmax, -ymax, ymax, near, far)


def LookAt(eye: Vector, center: Vector, up: Vector) -> Matrix:  # matrix.go:95
    z = eye.Sub(center).Normalize()
    x = up.Cross(z).Normalize()
    y = z.Cross(x)
    return Matrix(
        x.X, x.Y, x.Z, -x.Dot(eye),
        y.X, y.Y, y.Z, -y.Dot(eye),
        z.X, z.Y, z.Z, -z.Dot(eye),
        0, 0, 0, 1)


def Screen(w: int, h: int) -> Matrix:  # matrix.go:119
    w2 = float(w) / 2
    h2 = float(h) / 2
    return Matrix(w2, 0, 0, w2, 0, -h2, 0, h2, 0, 0, 0.5, 0.5, 0, 0, 0, 1)
