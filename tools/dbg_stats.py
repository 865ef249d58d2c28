"""Debug: draw statistics (records, segments, retries) of parity scenes in one process."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes
from fauxgl_b200.context import Context
for name in sys.argv[2:]:
    sc = scenes.SCENES[name]()
    g = Context(sc.width, sc.height); g.XGuard = sys.argv[1] == "1"
    class Spy:
        def __init__(s, c): s.__dict__["c"] = c
        def __getattr__(s, k): return getattr(s.c, k)
        def __setattr__(s, k, v): setattr(s.c, k, v)
        def DrawMesh(s, m):
            r = s.c.DrawMesh(m); st = s.c.DrawStats()
            print("   ", name, "DrawMesh", tuple(r), "prims", st.prims_in, "records", st.records, "segs", st.pairs, "clip", st.clip_triangles, "launches", st.kernel_launches, "retries", st.retries, flush=True)
            return r
    sc.run(Spy(g)); g.Close()
