"""Short driver for ncu captures: config-2 scene, primary closest / AO closest / AO any-hit casts."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mray_b200 import capi, scenes
import mray_b200

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ctx = mray_b200.Context(0); ctx.set_stream(torch.cuda.current_stream())
p, i = scenes.arcade_mesh()
acc = capi.Accelerator(ctx, torch.from_numpy(p).cuda(), torch.from_numpy(i.view(np.int32)).cuda())
rays = scenes.pinhole_rays(1920, 1080, **scenes.ARCADE_CAMERA)
n = rays.shape[0]
d0 = torch.from_numpy(rays).cuda()
k = torch.full((n, 4), -1, dtype=torch.int32, device="cuda"); h = torch.zeros((n, 2), device="cuda")
w = d0.clone()
acc.cast_rays(k, h, w, None, capi.MRB_TRACE_WIDE); torch.cuda.synchronize()
prim = k.cpu().numpy().view(np.uint32)[:, 0]; tp = w.cpu().numpy()[:, 7]
e = acc.export_lbvh(); diam = float(np.linalg.norm(e["accel_aabb"][3:] - e["accel_aabb"][:3]))
ao = scenes.ao_rays(rays, prim, tp, p, i, 0.15 * diam); da = torch.from_numpy(ao).cuda()
bits = torch.full(((n + 31) // 32,), -1, dtype=torch.int32, device="cuda")
for _ in range(reps):
    w.copy_(d0); acc.cast_rays(k, h, w, None, capi.MRB_TRACE_WIDE)
    w.copy_(da); acc.cast_rays(k, h, w, None, capi.MRB_TRACE_WIDE)
    acc.cast_visibility_rays(bits, da, None, capi.MRB_TRACE_WIDE)
torch.cuda.synchronize()
print("done")
