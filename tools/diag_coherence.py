"""Diagnostic (GPU box): closest/any-hit throughput of the config-2 AO rays in natural, shuffled and
spatially sorted order (direct and through rayIndices)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mray_b200
from mray_b200 import capi, scenes
ctx = mray_b200.Context(0); ctx.set_stream(torch.cuda.current_stream())
p, i = scenes.arcade_mesh()
acc = capi.Accelerator(ctx, torch.from_numpy(p).cuda(), torch.from_numpy(i.view(np.int32)).cuda())
rays = scenes.pinhole_rays(1920, 1080, **scenes.ARCADE_CAMERA)
n = rays.shape[0]
k = torch.full((n, 4), -1, dtype=torch.int32, device="cuda"); h = torch.zeros((n, 2), device="cuda")
w = torch.from_numpy(rays).cuda()
acc.cast_rays(k, h, w, None, capi.MRB_TRACE_WIDE); torch.cuda.synchronize()
prim = k.cpu().numpy().view(np.uint32)[:, 0]; tp = w.cpu().numpy()[:, 7]
e = acc.export_lbvh(); lo, hi = e["accel_aabb"][:3], e["accel_aabb"][3:]
ao = scenes.ao_rays(rays, prim, tp, p, i, 0.15 * float(np.linalg.norm(hi - lo)))
# second-bounce-like rays: origins = AO hit points, random directions, unbounded
def timed(dr, idx=None, any_hit=False, reps=5):
    ts = []
    for _ in range(reps):
        ww = dr.clone(); kk = torch.full((n, 4), -1, dtype=torch.int32, device="cuda"); hh = torch.zeros((n, 2), device="cuda")
        bits = torch.full(((n + 31) // 32,), -1, dtype=torch.int32, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if any_hit: acc.cast_visibility_rays(bits, ww, idx, capi.MRB_TRACE_WIDE)
        else: acc.cast_rays(kk, hh, ww, idx, capi.MRB_TRACE_WIDE)
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
def morton_key(r):
    q = np.clip(((r[:, :3] - lo) / (hi - lo) * 1023).astype(np.uint64), 0, 1023)
    def spread(v):
        v = (v | (v << 16)) & 0x030000FF; v = (v | (v << 8)) & 0x0300F00F; v = (v | (v << 4)) & 0x030C30C3; v = (v | (v << 2)) & 0x09249249; return v
    key = (spread(q[:, 0]) << 2) | (spread(q[:, 1]) << 1) | spread(q[:, 2])
    octant = ((r[:, 4] < 0).astype(np.uint64) << 2) | ((r[:, 5] < 0).astype(np.uint64) << 1) | (r[:, 6] < 0).astype(np.uint64)
    return key, octant
rng = np.random.default_rng(0)
for name, base in (("AO(0.15 diam)", ao), ("AO unbounded", np.concatenate([ao[:, :7], np.full((n, 1), 3e38, np.float32)], axis=1))):
    perm = rng.permutation(n)
    shuffled = np.ascontiguousarray(base[perm])
    key, octant = morton_key(shuffled)
    order_m = np.argsort(key, kind="stable")
    order_mo = np.argsort((key << np.uint64(3)) | octant, kind="stable")
    order_om = np.argsort((octant << np.uint64(30)) | key, kind="stable")
    for any_hit in (False, True):
        res = {}
        res["natural"] = timed(torch.from_numpy(base).cuda(), None, any_hit)
        ds = torch.from_numpy(shuffled).cuda()
        res["shuffled"] = timed(ds, None, any_hit)
        for nm, od in (("sorted(morton) direct", order_m), ("sorted(morton,octant) direct", order_mo), ("sorted(octant,morton) direct", order_om)):
            res[nm] = timed(torch.from_numpy(np.ascontiguousarray(shuffled[od])).cuda(), None, any_hit)
        res["sorted(morton,octant) via rayIndices"] = timed(ds, torch.from_numpy(order_mo.astype(np.int32)).cuda(), any_hit)
        print(name, "any" if any_hit else "closest", {a: round(n / b / 1e3, 0) for a, b in res.items()})
