"""Diagnostic (GPU box): classify closest-hit mismatches between the wide-BVH path and the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle_lib as O
from helpers import oracle_trace_mt, full_hashes
from mray_b200 import capi, scenes
import mray_b200

ctx = mray_b200.Context(0); ctx.set_stream(torch.cuda.current_stream())
p, i = scenes.arcade_mesh()
acc = capi.Accelerator(ctx, p, i)
b = O.oracle_build(p, i)
rank = np.empty(b.n, np.uint32); rank[b.sorted_idx] = np.arange(b.n, dtype=np.uint32)
rays = scenes.pinhole_rays(1920, 1080, **scenes.ARCADE_CAMERA)
prim, t, bary, _ = oracle_trace_mt(p, i, b, rays)
ao = scenes.ao_rays(rays, prim, t, p, i, 0.15 * full_hashes()["scene_diameter"])

def gpu(rays_np, mode):
    n = rays_np.shape[0]
    r = torch.from_numpy(rays_np.copy()).cuda()
    k = torch.full((n, 4), -1, dtype=torch.int32, device="cuda"); h = torch.zeros((n, 2), device="cuda")
    acc.cast_rays(k, h, r, None, mode); torch.cuda.synchronize()
    return k.cpu().numpy().view(np.uint32)[:, 0], r.cpu().numpy()[:, 7]

gprim, gt = gpu(ao, capi.MRB_TRACE_WIDE)
bprim, bt = gpu(ao, capi.MRB_TRACE_BINARY_EXACT)
aprim, at, _, _ = oracle_trace_mt(p, i, b, ao)
mm = np.nonzero(gprim != aprim)[0]
print("mismatch wide vs oracle:", mm.size, "of", ao.shape[0], "; binary-exact vs oracle:", int((bprim != aprim).sum()),
      "t mismatches (binary):", int((bt != at).sum()))
sub = np.ascontiguousarray(ao[mm[:200]])
brp, brt = O.oracle_brute(p, i, sub, rank)
print("wide == brute on mismatches:", int((gprim[mm[:200]] == brp).sum()), "of", sub.shape[0])
for k in mm[:12]:
    print(k, "ray", ao[k], "gpu", gprim[k], gt[k], "oracle", aprim[k], at[k])
