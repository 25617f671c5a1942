"""Diagnostic (GPU box): build time, per-cast times and exact-fallback statistics on config 2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from mray_b200 import capi, scenes
import mray_b200

ctx = mray_b200.Context(0); ctx.set_stream(torch.cuda.current_stream())
p, i = scenes.arcade_mesh()
dp, di = torch.from_numpy(p).cuda(), torch.from_numpy(i.view(np.int32)).cuda()
for k in range(3):
    acc = capi.Accelerator(ctx, dp, di)
    print("build ms", acc.info.buildMs, "wide nodes", acc.info.wideNodeCount)
    if k < 2: acc.close()
rays = scenes.pinhole_rays(1920, 1080, **scenes.ARCADE_CAMERA)
n = rays.shape[0]
d0 = torch.from_numpy(rays).cuda()
def cast(dr, any_hit=False, reps=5):
    ts = []
    for _ in range(reps):
        w = dr.clone(); k = torch.full((n, 4), -1, dtype=torch.int32, device="cuda"); h = torch.zeros((n, 2), device="cuda")
        bits = torch.full(((n + 31) // 32,), -1, dtype=torch.int32, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if any_hit: acc.cast_visibility_rays(bits, w, None, capi.MRB_TRACE_WIDE)
        else: acc.cast_rays(k, h, w, None, capi.MRB_TRACE_WIDE)
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), ctx.last_fallback_stats, k, w
t, fb, k, w = cast(d0); print("primary closest ms", t, "Mrays/s", n / t / 1e3, "fallback", fb)
prim = k.cpu().numpy().view(np.uint32)[:, 0]; tp = w.cpu().numpy()[:, 7]
e = acc.export_lbvh(); diam = float(np.linalg.norm(e["accel_aabb"][3:] - e["accel_aabb"][:3]))
ao = scenes.ao_rays(rays, prim, tp, p, i, 0.15 * diam); da = torch.from_numpy(ao).cuda()
t, fb, _, _ = cast(da); print("ao closest ms", t, "Mrays/s", n / t / 1e3, "fallback", fb)
t, fb, _, _ = cast(da, True); print("ao any ms", t, "Mrays/s", n / t / 1e3, "fallback", fb)
for sz in (100_000, 1_000_000, 10_000_000):
    sp, si = scenes.random_soup(sz)
    a2 = capi.Accelerator(ctx, torch.from_numpy(sp).cuda(), torch.from_numpy(si.view(np.int32)).cuda())
    a2.close(); a2 = capi.Accelerator(ctx, torch.from_numpy(sp).cuda(), torch.from_numpy(si.view(np.int32)).cuda())
    print("soup", sz, "build ms", a2.info.buildMs, "Mtris/s", sz / a2.info.buildMs / 1e3, "wide", a2.info.wideNodeCount); a2.close()
