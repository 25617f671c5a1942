"""Diagnostic (GPU box): 1080p path tracer ms/spp with / without the material-key ray sort."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mray_b200
from mray_b200 import capi, scenes
ctx = mray_b200.Context(0); stream = torch.cuda.current_stream(); ctx.set_stream(stream)
p, i = scenes.arcade_mesh()
pidx, pranges, pkeys, palb, prad, _ = scenes.arcade_materials(p, i)
acc = mray_b200.Accelerator(ctx, torch.from_numpy(p).cuda(), torch.from_numpy(pidx.view(np.int32)).cuda(), prim_ranges=pranges, light_or_mat_keys=pkeys)
for part in (True, False, True, False):
    spp = 8
    r = mray_b200.Renderer(ctx, acc, p.shape[0], pidx.shape[0], palb, prad, scenes.ARCADE_CAMERA, 1920, 1080, spp, sample_mode="WithNEEAndMIS",
                           rr_range=(3, 8), seed=0, partition_rays=part)
    r.iterate(2); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    while True:
        r.iterate(8); st = r.stats()
        if st.finished: break
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("partition", part, "ms/spp", round(ms / spp, 3), "Mrays/s", round((st.closestRays + st.shadowRays) / ms / 1e3, 1), "iters", st.iterations)
    r.close()
