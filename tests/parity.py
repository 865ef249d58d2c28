"""Shared comparison helpers for the parity tests."""
from __future__ import annotations

import numpy as np


def compare_buffers(ocolor, odepth, gcolor, gdepth):
    """Return a dict of mismatch statistics between oracle and GPU buffers.

    coverage/depth-test outcome: a pixel 'matches' when its float64 depth is
    bit-identical (covered-by-the-same-winner implies identical depth; a cleared
    pixel holds MaxFloat64 on both sides)."""
    od = odepth.view(np.uint64)
    gd = gdepth.view(np.uint64)
    depth_mis = od != gd
    color_mis = (ocolor != gcolor).any(axis=-1)
    max_err = int(np.abs(ocolor.astype(np.int16) - gcolor.astype(np.int16)).max()) if ocolor.size else 0
    # colour error on pixels whose depth outcome matches (north_star's tolerance)
    ok = ~depth_mis
    err_ok = int(np.abs(ocolor[ok].astype(np.int16) - gcolor[ok].astype(np.int16)).max()) if ok.any() else 0
    first = None
    bad = depth_mis | color_mis
    if bad.any():
        y, x = np.argwhere(bad)[0]
        first = (int(x), int(y), odepth[y, x], gdepth[y, x], ocolor[y, x].tolist(), gcolor[y, x].tolist())
    return {
        "pixels": int(depth_mis.size),
        "depth_mismatch": int(depth_mis.sum()),
        "color_mismatch": int(color_mis.sum()),
        "max_color_err": max_err,
        "max_color_err_on_matching_depth": err_ok,
        "first": first,
    }


def run_both(scene, oracle_mod, gpu_ctx_cls, device=0, x_guard=False):
    """x_guard=False (default) is the reference's own index rule on both sides (context.go:223-228)."""
    octx = oracle_mod.OracleContext(scene.width, scene.height, x_guard=x_guard)
    oinfo = scene.run(octx)
    gctx = gpu_ctx_cls(scene.width, scene.height, device)
    gctx.XGuard = x_guard
    ginfo = scene.run(gctx)
    gcolor, gdepth = gctx.Image(), gctx.DepthBuffer
    stats = compare_buffers(octx.ColorBuffer, octx.DepthBuffer, gcolor, gdepth)
    stats["oracle_info"] = oinfo
    stats["gpu_info"] = ginfo
    gctx.Close()
    return stats
