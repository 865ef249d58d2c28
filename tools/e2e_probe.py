"""What bounds the end-to-end frame (tuning aid): raw PCIe rates of the box for the e2e transfer sizes, alone and both
directions at once, then the indexed-upload pipeline at several depths.   python tools/e2e_probe.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from fauxgl_b200 import synth  # noqa: E402
from fauxgl_b200.context import Context  # noqa: E402
from fauxgl_b200.pipeline import IndexedFramePipeline  # noqa: E402

dev = torch.device("cuda:0")
up = [torch.empty(435986 * 3, dtype=torch.float64, pin_memory=True) for _ in range(2)]
dup = [torch.empty(435986 * 3, dtype=torch.float64, device=dev) for _ in range(2)]
img = torch.empty(1080 * 1920 * 4, dtype=torch.uint8, pin_memory=True)
dimg = torch.empty(1080 * 1920 * 4, dtype=torch.uint8, device=dev)
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, n=20):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / n


def h2d():
    with torch.cuda.stream(s_up):
        for a, b in zip(up, dup):
            b.copy_(a, non_blocking=True)


def d2h():
    with torch.cuda.stream(s_dn):
        img.copy_(dimg, non_blocking=True)


def both():
    h2d()
    d2h()


t_up, t_dn, t_both = timed(h2d), timed(d2h), timed(both)
print("H2D 2 x %.1f MB: %.3f ms (%.1f GB/s)   D2H %.1f MB: %.3f ms (%.1f GB/s)   both at once: %.3f ms" % (
    up[0].numel() * 8 / 1e6, t_up, 2 * up[0].numel() * 8 / t_up / 1e6, img.numel() / 1e6, t_dn, img.numel() / t_dn / 1e6, t_both))

mesh = synth.bumpy_surface()
shader, bg = bench.scene_setup()
tv, tvn, corners = mesh.indexed()
pv = torch.empty(tv.shape, dtype=torch.float64, pin_memory=True)
pvn = torch.empty(tvn.shape, dtype=torch.float64, pin_memory=True)
pv.numpy()[...] = tv
pvn.numpy()[...] = tvn
for depth in (1, 2, 3, 4):
    ctx = Context(bench.W1, bench.H1)
    ctx.Shader = shader
    pipe = IndexedFramePipeline(ctx, pv.numpy(), pvn.numpy(), corners, depth=depth)

    def run(n):
        for _ in range(n):
            if len(pipe) == depth:
                pipe.collect()
            pipe.submit(pv.numpy(), pvn.numpy(), bg)
        while len(pipe):
            pipe.collect()
    run(4)
    t0 = time.perf_counter()
    run(20)
    torch.cuda.synchronize()
    print("indexed pipeline depth %d: %.3f ms/frame" % (depth, (time.perf_counter() - t0) * 1e3 / 20))
    del pipe
    ctx.Close()
