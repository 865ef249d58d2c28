"""Time a spread of workloads (ms per frame, device-timed) to spot pathologies outside the
headline config: big triangles, inline shading (texture / blending), wireframe, lines, 10 M triangles.

    python tools/perf_scenes.py
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import scenes  # noqa: E402
from fauxgl_b200 import synth  # noqa: E402
from fauxgl_b200.context import Context  # noqa: E402


class Timed:
    """Wraps a Context: every DrawMesh/DrawTriangles/DrawLines is timed with CUDA events."""

    def __init__(self, ctx):
        self.__dict__["c"] = ctx
        self.__dict__["ms"] = 0.0
        self.__dict__["ext"] = torch.cuda.ExternalStream(ctx.stream_ptr)

    def __getattr__(self, k):
        return getattr(self.c, k)

    def __setattr__(self, k, v):
        setattr(self.c, k, v)

    def _timed(self, fn, *a):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(self.ext)
        r = fn(*a)
        e.record(self.ext)
        torch.cuda.synchronize()
        self.__dict__["ms"] += s.elapsed_time(e)
        return r

    def DrawMesh(self, m):
        return self._timed(self.c.DrawMesh, m)

    def DrawTriangles(self, m, *a):
        return self._timed(self.c.DrawTriangles, m, *a)

    def DrawLines(self, m, *a):
        return self._timed(self.c.DrawLines, m, *a)


def run(name, scene, reps=3):
    ctx = Context(scene.width, scene.height)
    best = None
    for _ in range(reps):
        ctx.ClearDepthBuffer()
        ctx.Wireframe, ctx.DepthBias, ctx.LineWidth = False, 0.0, 2.0
        ctx.Cull, ctx.FrontFace = 3, 2
        ctx.ReadDepth = ctx.WriteDepth = ctx.WriteColor = ctx.AlphaBlend = True
        t = Timed(ctx)
        infos = scene.run(t)
        best = t.ms if best is None else min(best, t.ms)
    tot = sum(i[0] for i in infos)
    print("%-34s %5dx%-5d draws ms %8.3f   TotalPixels %11d   %7.1f Mpix/s" % (
        name, scene.width, scene.height, best, tot, tot / best / 1e3))
    ctx.Close()


def main():
    for name in ("hello", "bowser", "bowser_close", "capsule_texture", "capsule_phong_texture", "shapes_multipass",
                 "lines", "state_wireframe_solid"):
        run(name, scenes.SCENES[name]())
    run("bowser 7680x4320", scenes._bowser_scene(7680, 4320))
    m = synth.bumpy_surface()
    run("M871k 1080p", scenes.dragon_scene(m, 1920, 1080))
    run("M871k 8K", scenes.dragon_scene(m, 7680, 4320))

    def wire(mesh, W, H):
        sc = scenes.dragon_scene(mesh, W, H)
        inner = sc._run

        def r(ctx):
            ctx.Wireframe = True
            ctx.LineWidth = 1.0
            return inner(ctx)
        sc._run = r
        return sc
    run("M871k wireframe 1080p", wire(m, 1920, 1080))
    sphere = synth.uv_sphere(2237, 2237)
    run("M10M sphere 8K", scenes.dragon_scene(sphere, 7680, 4320), reps=2)


if __name__ == "__main__":
    main()
