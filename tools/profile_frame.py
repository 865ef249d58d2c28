"""Render a few frames of a benchmark configuration, for use under ncu:

    ncu --set full -k regex:k_raster -c 1 python tools/profile_frame.py --frames 2
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from fauxgl_b200 import synth  # noqa: E402
from fauxgl_b200.context import Context, DeviceMesh  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=2)
ap.add_argument("--scale", type=int, default=1)
ap.add_argument("--resolve", type=int, default=0)
args = ap.parse_args()

mesh = synth.bumpy_surface()
shader, bg = bench.scene_setup()
ctx = Context(bench.W1 * args.scale, bench.H1 * args.scale)
ctx.Shader = shader
dm = DeviceMesh(ctx, mesh, ("position", "normal"))
ctx.DrawMesh(dm)  # synchronous first draw sizes the work buffers
for _ in range(args.frames):
    ctx.ClearDepthBuffer()
    ctx.ClearColorBufferWith(bg)
    ctx.DrawMeshAsync(dm)
    if args.resolve:
        ctx.ResolveDevice(args.resolve)
print(ctx.Sync())
