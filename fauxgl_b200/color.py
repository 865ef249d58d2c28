"""Host-side colour helpers mirroring the reference's ``color.go``."""
from __future__ import annotations

from typing import NamedTuple


class Color(NamedTuple):
    """color.go:17-19 -- float64 RGBA, non-premultiplied."""
    R: float
    G: float
    B: float
    A: float

    def Alpha(a, alpha):  # color.go:69
        return Color(a.R, a.G, a.B, float(alpha))

    def Opaque(a):  # color.go:65
        return Color(a.R, a.G, a.B, 1.0)

    def MulScalar(a, b):  # color.go:101
        return Color(a.R * b, a.G * b, a.B * b, a.A * b)

    def NRGBA(c):  # color.go:56-63 (clamp, truncating *255)
        def q(x):
            x = 0.0 if x < 0 else (1.0 if x > 1 else x)
            return int(x * 255) & 0xFF
        return (q(c.R), q(c.G), q(c.B), q(c.A))


Discard = Color(0.0, 0.0, 0.0, 0.0)      # color.go:11
Transparent = Color(0.0, 0.0, 0.0, 0.0)  # color.go:12
Black = Color(0.0, 0.0, 0.0, 1.0)        # color.go:13
White = Color(1.0, 1.0, 1.0, 1.0)        # color.go:14


def Gray(x: float) -> Color:  # color.go:21
    return Color(float(x), float(x), float(x), 1.0)


def HexColor(x: str) -> Color:  # color.go:31-54
    x = x.strip("#")
    r = g = b = 0
    a = 255
    if len(x) == 3:
        r, g, b = (int(ch, 16) for ch in x)
        r, g, b = (r << 4) | r, (g << 4) | g, (b << 4) | b
    elif len(x) == 4:
        r, g, b, a = (int(ch, 16) for ch in x)
        r, g, b, a = (r << 4) | r, (g << 4) | g, (b << 4) | b, (a << 4) | a
    elif len(x) == 6:
        r, g, b = int(x[0:2], 16), int(x[2:4], 16), int(x[4:6], 16)
    elif len(x) == 8:
        r, g, b, a = int(x[0:2], 16), int(x[2:4], 16), int(x[4:6], 16), int(x[6:8], 16)
    d = 255.0
    return Color(r / d, g / d, b / d, a / d)


def ycbcr_to_rgba64(y, cb, cr):
    """What Go's ``color.YCbCr{Y, Cb, Cr}.RGBA()`` returns (image/color/ycbcr.go, the 16-bit variant of
    YCbCrToRGB: fixed-point BT.601 with 0x10101 / 91881 / 22554 / 46802 / 116130), per texel, as an
    (H, W, 4) uint16 array -- the TEX_RGBA64 upload of a texture that Go decoded to *image.YCbCr (every JPEG;
    examples/capsule.go:35).  A Go host calls At(x, y).RGBA() itself; this numpy restatement (Go stdlib is not in
    the reference tree: unpinned) only feeds the parity scene, oracle and device alike."""
    import numpy as np
    yy1 = np.asarray(y, dtype=np.int32) * 0x10101
    cb1 = np.asarray(cb, dtype=np.int32) - 128
    cr1 = np.asarray(cr, dtype=np.int32) - 128

    def fix(v):
        v = v.astype(np.int32)
        ok = (v.view(np.uint32) & 0xff000000) == 0
        return np.where(ok, v >> 8, ~(v >> 31) & 0xffff).astype(np.uint16)
    r = fix(yy1 + 91881 * cr1)
    g = fix(yy1 - 22554 * cb1 - 46802 * cr1)
    b = fix(yy1 + 116130 * cb1)
    a = np.full(r.shape, 0xffff, dtype=np.uint16)
    return np.stack([r, g, b, a], axis=-1)
