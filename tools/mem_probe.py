import os, sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, scenes
from fauxgl_b200.context import Context
for front in ("fused","split"):
    os.environ["FGL_FRONT"]=front
    for name in ("shapes_multipass","bowser","hello","lines_scene","capsule_texture"):
        sc=getattr(scenes,name)()
        torch.cuda.synchronize(); f0=torch.cuda.mem_get_info()[0]
        ctx=Context(sc.width, sc.height); sc.run(ctx); ctx.Sync()
        f1=torch.cuda.mem_get_info()[0]
        print(front, name, sc.width, sc.height, "device MB used by the context: %.1f"%((f0-f1)/1e6))
        ctx.Close()
