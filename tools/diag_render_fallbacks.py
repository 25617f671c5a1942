"""Diagnostic (GPU box): exact-fallback statistics of the closest-hit casts inside the path tracer (Pure mode, so
the closest cast is the last cast of every iteration)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mray_b200
from mray_b200 import capi, scenes
ctx = mray_b200.Context(0); ctx.set_stream(torch.cuda.current_stream())
p, i = scenes.arcade_mesh()
pidx, pranges, pkeys, palb, prad, _ = scenes.arcade_materials(p, i)
acc = mray_b200.Accelerator(ctx, torch.from_numpy(p).cuda(), torch.from_numpy(pidx.view(np.int32)).cuda(), prim_ranges=pranges, light_or_mat_keys=pkeys)
r = mray_b200.Renderer(ctx, acc, p.shape[0], pidx.shape[0], palb, prad, scenes.ARCADE_CAMERA, 1920, 1080, 64, sample_mode="Pure", rr_range=(3, 8), seed=0)
for it in range(8):
    r.iterate(1)
    print("iteration", it, "(uncertified, near-tie, leaf-cert-fail, full binary):", ctx.last_fallback_stats)
