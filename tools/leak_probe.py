import os, sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, scenes
from fauxgl_b200.context import Context
torch.cuda.synchronize(); base=torch.cuda.mem_get_info()[0]
for it in range(12):
    for front in ("fused","split"):
        os.environ["FGL_FRONT"]=front
        sc=scenes.shapes_multipass()
        ctx=Context(sc.width, sc.height); sc.run(ctx); ctx.Sync(); ctx.Close()
    if it % 3 == 2:
        print(it, "MB not returned since start: %.1f"%((base-torch.cuda.mem_get_info()[0])/1e6))
