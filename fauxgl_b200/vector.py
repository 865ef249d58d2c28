"""Host-side float64 vector helpers, mirroring the subset of the reference's
``vector.go`` the examples use to set a scene up (V, Normalize, Sub, Cross, Dot).

Python floats are IEEE doubles and every operator rounds once, so writing the
expressions in the reference's order reproduces Go/amd64's unfused results.
These run once per frame on the host; nothing here is on the hot path.
"""
from __future__ import annotations

import math
from typing import NamedTuple


class Vector(NamedTuple):
    """vector.go:8-10"""
    X: float
    Y: float
    Z: float

    def Add(a, b):  # vector.go:96
        return Vector(a.X + b.X, a.Y + b.Y, a.Z + b.Z)

    def Sub(a, b):  # vector.go:100
        return Vector(a.X - b.X, a.Y - b.Y, a.Z - b.Z)

    def Mul(a, b):  # vector.go:104
        return Vector(a.X * b.X, a.Y * b.Y, a.Z * b.Z)

    def Div(a, b):  # vector.go:108
        return Vector(a.X / b.X, a.Y / b.Y, a.Z / b.Z)

    def MulScalar(a, b):  # vector.go:128
        return Vector(a.X * b, a.Y * b, a.Z * b)

    def Dot(a, b):  # vector.go:72
        return a.X * b.X + a.Y * b.Y + a.Z * b.Z

    def Cross(a, b):  # vector.go:76
        return Vector(a.Y * b.Z - a.Z * b.Y, a.Z * b.X - a.X * b.Z, a.X * b.Y - a.Y * b.X)

    def Length(a):  # vector.go:38
        return math.sqrt(a.X * a.X + a.Y * a.Y + a.Z * a.Z)

    def Normalize(a):  # vector.go:83
        r = 1 / math.sqrt(a.X * a.X + a.Y * a.Y + a.Z * a.Z)
        return Vector(a.X * r, a.Y * r, a.Z * r)

    def Negate(a):  # vector.go:88
        return Vector(-a.X, -a.Y, -a.Z)

    def MinComponent(a):  # vector.go:167
        return min(min(a.X, a.Y), a.Z)

    def Perpendicular(a):  # vector.go:179
        if a.X == 0 and a.Y == 0:
            if a.Z == 0:
                return Vector(0.0, 0.0, 0.0)
            return Vector(0.0, 1.0, 0.0)
        return Vector(-a.Y, a.X, 0.0).Normalize()


def V(x, y, z) -> Vector:
    """vector.go:12"""
    return Vector(float(x), float(y), float(z))


def Radians(degrees: float) -> float:
    """util.go:15"""
    return degrees * math.pi / 180
