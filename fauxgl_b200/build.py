"""Build libfauxgl_b200.so in-tree with nvcc for sm_100a.

``-fmad=false`` is part of the product's semantics, not a tuning flag: the
reference's float64 arithmetic is unfused (Go/amd64), and coverage, depth and
colour must come out bit-identical.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfauxgl_b200.so")
SOURCES = ["fgl_api.cu", "fgl_geom.cu", "fgl_scan_sort.cu", "fgl_span.cu", "fgl_order.cu", "fgl_raster.cu", "fgl_post.cu", "fgl_ingest.cu", "fgl_comm.cu"]
HEADERS = ["fgl_internal.h", "fgl_ctx.h", "fgl_math.cuh", "fgl_block.cuh", "fgl_shade.cuh", "fgl_walk.cuh", "../../include/fauxgl_b200.h"]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--shared",
    "-ldl",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; fauxgl_b200 cannot be built (there is no CPU fallback)")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + proc.stdout)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed (see %s)" % log)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
