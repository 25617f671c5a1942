"""ncu driver: a few iterations of the 1080p path tracer on the arcade mesh (config-3 flavour)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mray_b200
from mray_b200 import capi, scenes, spectral
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 6
ctx = mray_b200.Context(0); ctx.set_stream(torch.cuda.current_stream())
p, i = scenes.arcade_mesh()
pidx, pranges, pkeys, palb, prad, _ = scenes.arcade_materials(p, i)
acc = mray_b200.Accelerator(ctx, torch.from_numpy(p).cuda(), torch.from_numpy(pidx.view(np.int32)).cuda(), prim_ranges=pranges, light_or_mat_keys=pkeys)
spec = mray_b200.Spectrum(ctx, spectral.load()) if (len(sys.argv) > 2 and sys.argv[2] == "spectral") else None
r = mray_b200.Renderer(ctx, acc, p.shape[0], pidx.shape[0], palb, prad, scenes.ARCADE_CAMERA, 1920, 1080, 64, sample_mode="WithNEEAndMIS",
                       rr_range=(3, 8), seed=0, partition_rays=("sort" in sys.argv), spectrum=spec)
r.iterate(iters); torch.cuda.synchronize()
print("done", r.stats().pathsCompleted)
