import os, sys
os.environ["FGL_TILE_CLOCK"]="1"
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import test_features_gpu as t
from fauxgl_b200.context import Context, capi, _check
for kind in ("wide","narrow"):
    sc=t._stacked_scene(kind,"default")
    ctx=Context(sc.width, sc.height)
    print(kind, sc.run(ctx))
    st=ctx.DrawStats(); nt=st.tiles_x*st.tiles_y
    full=np.zeros((nt+8,2),dtype=np.uint64)
    _check(capi().fgl_debug_tile_cycles(ctx._h, full.ctypes.data, nt+8), ctx._h)
    dbg=full[nt:].ravel(); segs=(full[:nt,1] & np.uint64(0xffffffff)).astype(np.int64)
    print(" heavy strips (last draw)", int(dbg[9]), "strips", nt, "max segs/strip", int(segs.max()), "strips >= 96:", int((segs>=96).sum()))
