"""ncu driver: one warm + one measured build of an N-triangle random soup."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mray_b200
from mray_b200 import capi, scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
ctx = mray_b200.Context(0); ctx.set_stream(torch.cuda.current_stream())
p, i = scenes.random_soup(n)
dp, di = torch.from_numpy(p).cuda(), torch.from_numpy(i.view(np.int32)).cuda()
for _ in range(2):
    a = capi.Accelerator(ctx, dp, di); print(a.info.buildMs); a.close()
