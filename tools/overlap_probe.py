"""Throughput of independent frames rendered concurrently from several contexts (one stream each) that share
one device-resident mesh: python tools/overlap_probe.py [--ctx 3] [--frames 60] [--scale 1]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from fauxgl_b200 import synth  # noqa: E402
from fauxgl_b200.context import Context, DeviceMesh  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ctx", type=int, default=3)
ap.add_argument("--frames", type=int, default=60)
ap.add_argument("--scale", type=int, default=1)
args = ap.parse_args()
mesh = synth.bumpy_surface()
shader, bg = bench.scene_setup()
for nctx in range(1, args.ctx + 1):
    ctxs = [Context(bench.W1 * args.scale, bench.H1 * args.scale) for _ in range(nctx)]
    for c in ctxs:
        c.Shader = shader
    dm = DeviceMesh(ctxs[0], mesh, ("position", "normal"))
    streams = [torch.cuda.ExternalStream(c.stream_ptr) for c in ctxs]
    for c in ctxs:
        c.DrawMesh(dm)

    def frame(c):
        c.ClearDepthBuffer()
        c.ClearColorBufferWith(bg)
        c.DrawMeshAsync(dm)
    for i in range(3 * nctx):
        frame(ctxs[i % nctx])
    for c in ctxs:
        c.Sync()
    torch.cuda.synchronize()
    start = torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event(enable_timing=True) for _ in ctxs]
    start.record(streams[0])
    for s in streams[1:]:
        s.wait_event(start)
    for i in range(args.frames):
        frame(ctxs[i % nctx])
    for e, s in zip(ends, streams):
        e.record(s)
    infos = [c.Sync() for c in ctxs]
    torch.cuda.synchronize()
    ms = max(start.elapsed_time(e) for e in ends)
    print("contexts %d: %d frames in %.3f ms -> %.4f ms/frame, %.0f Mtri/s  (%s)" % (
        nctx, args.frames, ms, ms / args.frames, bench.T_TRIANGLES * args.frames / ms / 1e3, infos[0]))
    del dm
    for c in ctxs:
        c.Close()
