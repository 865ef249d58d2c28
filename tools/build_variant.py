"""Build a tuning variant of the library next to the real one:

    python tools/build_variant.py NAME -DFGL_FRONT_MINB=6 ...   ->  fauxgl_b200/libfauxgl_b200.NAME.so
    FGL_LIB=fauxgl_b200/libfauxgl_b200.NAME.so python bench.py ...
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fauxgl_b200 import build  # noqa: E402

name, extra = sys.argv[1], sys.argv[2:]
out = os.path.join(build.HERE, "libfauxgl_b200.%s.so" % name)
cmd = [build.nvcc_path()] + build.NVCC_FLAGS + extra + ["-o", out] + [os.path.join(build.CSRC, s) for s in build.SOURCES]
proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
if proc.returncode:
    sys.stderr.write(proc.stdout)
    sys.exit(1)
print(out)
