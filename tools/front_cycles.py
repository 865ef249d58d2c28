"""Where a k_front block spends its cycles on the benchmark frame (tuning aid).  Needs a variant library built with
-DFGL_FRONT_CLOCK=1:

    python tools/build_variant.py fclock -DFGL_FRONT_CLOCK=1
    FGL_LIB=fauxgl_b200/libfauxgl_b200.fclock.so python tools/front_cycles.py [--scale 4]
"""
import argparse
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
DRAWS = 3
for _ in range(DRAWS):
    ctx.ClearDepthBuffer()
    ctx.ClearColorBufferWith(bg)
    info = ctx.DrawMesh(dm)
st = ctx.DrawStats()
nt = st.tiles_x * st.tiles_y
full = np.zeros((nt + 8, 2), dtype=np.uint64)
_check(capi().fgl_debug_tile_cycles(ctx._h, full.ctypes.data, nt + 8), ctx._h)
dbg = full[nt:].ravel().astype(np.float64)
nb = max(dbg[5], 1.0)
names = ["phase 1 (thread 0: loads, transform, divisions, set-up)", "barrier behind phase 1", "scans, item ranges, first lookup",
         "walk of warp 0's items", "publish"]
print("k_front, %d draws, %d fast blocks: cycles of warp 0 per block" % (DRAWS, int(nb)))
tot = 0.0
for k, nme in enumerate(names):
    print("  %8.0f  %s" % (dbg[k] / nb, nme))
    tot += dbg[k] / nb
print("  %8.0f  total (%.2f us at 1965 MHz)" % (tot, tot / 1965.0))
print("mean lifetime of a warp: %.0f cycles over %d warps" % (dbg[6] / max(dbg[7], 1.0), int(dbg[7])))
