"""Per-tile cycle counts of the tile kernel on the benchmark frame (tuning aid).

    FGL_TILE_CLOCK=1 python tools/tile_cycles.py [--scale 4]
"""
import argparse
import ctypes as C
import os
import sys

os.environ["FGL_TILE_CLOCK"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import bench  # noqa: E402
from fauxgl_b200 import synth  # noqa: E402
from fauxgl_b200.context import Context, DeviceMesh, _check, capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=1)
args = ap.parse_args()

mesh = synth.bumpy_surface()
shader, bg = bench.scene_setup()
ctx = Context(bench.W1 * args.scale, bench.H1 * args.scale)
ctx.Shader = shader
dm = DeviceMesh(ctx, mesh, ("position", "normal"))
for _ in range(3):
    ctx.ClearDepthBuffer()
    ctx.ClearColorBufferWith(bg)
    info = ctx.DrawMesh(dm)
st = ctx.DrawStats()
nt = st.tiles_x * st.tiles_y
full = np.zeros((nt + 8, 2), dtype=np.uint64)
_check(capi().fgl_debug_tile_cycles(ctx._h, full.ctypes.data, nt + 8), ctx._h)
buf, dbg = full[:nt], full[nt:].ravel()
print("chunks by path (3 draws): disjoint %d staged %d generic %d; fragments %d / %d / %d; generic rounds %d; heavy strips %d" % (
    dbg[0], dbg[1], dbg[2], dbg[4], dbg[5], dbg[6], dbg[8], dbg[9]))
if dbg[10]:  # variant built with -DFGL_STRIP_PHASES=1: staged chunks of heavy strips
    n = float(dbg[10])
    print("heavy-strip staged chunks %d: cycles per chunk  pairs(c+2) %.0f  records wait+take %.0f  issue next gather %.0f  overlap test + stage %.0f  replay %.0f" % (
        dbg[10], dbg[11] / n, dbg[12] / n, dbg[13] / n, dbg[14] / n, dbg[15] / n))
cyc = buf[:, 0].astype(np.int64)
smid = (buf[:, 1] >> np.uint64(32)).astype(np.int64)
segs = (buf[:, 1] & np.uint64(0xffffffff)).astype(np.int64)
busy = cyc > 0
print("info", info, "tiles", nt, "busy", int(busy.sum()), "segments", int(segs.sum()))
print("cycles: sum %.0f  max %d  mean(busy) %.0f" % (cyc.sum(), cyc.max(), cyc[busy].mean()))
order = np.argsort(-cyc)[:12]
for t in order:
    print("  tile %5d (x=%3d y=%3d) sm %3d segs %6d cycles %8d  (%.1f us)  cyc/seg %.0f" % (
        t, t % st.tiles_x, t // st.tiles_x, smid[t], segs[t], cyc[t], cyc[t] / 1965.0, cyc[t] / max(segs[t], 1)))
per_sm = np.bincount(smid[busy], weights=cyc[busy], minlength=148)
print("per-SM busy cycles: max %.0f mean %.0f  (2 CTAs/SM run concurrently)" % (per_sm.max(), per_sm.mean()))
# cycles vs segments fit
A = np.vstack([segs[busy], np.ones(busy.sum())]).T
coef, *_ = np.linalg.lstsq(A, cyc[busy], rcond=None)
print("fit: cycles ~= %.1f * segs + %.0f" % (coef[0], coef[1]))
