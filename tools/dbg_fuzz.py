"""Debug: per-primitive RasterizeInfo of one fuzz seed, oracle vs device."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import test_fuzz_gpu as t
from oracle import pyoracle
from fauxgl_b200.context import Context
seed = int(sys.argv[1]); front = sys.argv[2] if len(sys.argv) > 2 else "fused"
if front != "auto":
    os.environ["FGL_FRONT"] = front
W, H, clear, draws, xg = t._script(seed)
print("seed", seed, W, H, "x_guard", xg, "draws", len(draws))
o = pyoracle.OracleContext(W, H, x_guard=xg); g = Context(W, H); g.XGuard = xg
for c in (o, g):
    if clear is not None:
        c.ClearColorBufferWith(clear)
for di, (mesh, shader, state) in enumerate(draws):
    for c in (o, g):
        c.Shader = shader
        for k, v in state.items():
            setattr(c, k, v)
    print("draw", di, type(shader).__name__, state, mesh.num_triangles, mesh.num_lines)
    oe, ge = o.DrawTrianglesEach(mesh), g.DrawTrianglesEach(mesh)
    bad = np.nonzero((oe != ge).any(axis=1))[0]
    for b in bad[:10]:
        print("  tri", b, "oracle", oe[b], "gpu", ge[b]); print(mesh.position[b])
    if mesh.num_lines:
        oe, ge = o.DrawLinesEach(mesh), g.DrawLinesEach(mesh)
        bad = np.nonzero((oe != ge).any(axis=1))[0]
        for b in bad[:10]:
            print("  line", b, "oracle", oe[b], "gpu", ge[b]); print(mesh.lposition[b])
    print("  depth mismatches", int((o.DepthBuffer.view(np.uint64) != g.DepthBuffer.view(np.uint64)).sum()),
          "colour", int((o.ColorBuffer != g.Image()).any(axis=-1).sum()))
