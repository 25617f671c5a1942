import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch, mray_b200
from mray_b200 import capi, scenes
ctx = mray_b200.Context(0); ctx.set_stream(torch.cuda.current_stream())
p, i = scenes.arcade_mesh()
dp, di = torch.from_numpy(p).cuda(), torch.from_numpy(i.view(np.int32)).cuda()
for _ in range(3):
    a = capi.Accelerator(ctx, dp, di); print("build ms", a.info.buildMs, flush=True); a.close()
torch.cuda.synchronize()
