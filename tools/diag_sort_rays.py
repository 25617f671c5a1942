"""Diagnostic (GPU box): spatially sorted ray order for the two casts of the path tracer (env MRB_SORT_RAYS = Morton bits per axis of the
ray origin, + direction octant; MRB_SORT_WHICH 1 = closest-hit cast, 2 = shadow cast, 3 = both; read once per process) on the config-3
flavour (1080p, RGB, 16 spp). usage: python tools/diag_sort_rays.py   (parent: one child per setting, prints one JSON line)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    import mray_b200
    from mray_b200 import scenes
    ctx = mray_b200.Context(0); stream = torch.cuda.current_stream(); ctx.set_stream(stream)
    p, i = scenes.arcade_mesh()
    pidx, pranges, pkeys, palb, prad, _ = scenes.arcade_materials(p, i)
    acc = mray_b200.Accelerator(ctx, torch.from_numpy(p).cuda(), torch.from_numpy(pidx.view(np.int32)).cuda(), prim_ranges=pranges, light_or_mat_keys=pkeys)
    best = 1e30; mean = None
    for rep in range(2):
        r = mray_b200.Renderer(ctx, acc, p.shape[0], pidx.shape[0], palb, prad, scenes.ARCADE_CAMERA, 1920, 1080, 16, sample_mode="WithNEEAndMIS",
                               rr_range=(3, 8), seed=7)
        r.iterate(2); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        while True:
            r.iterate(8)
            if r.stats().finished: break
        e1.record(stream); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 16)
        rgb, w = r.read_film()
        mean = float((rgb / np.maximum(w, 1e-20)[..., None]).mean())
        r.close()
    print("MS_PER_SPP", round(best, 4), round(mean, 6))
    sys.exit(0)
out = {}
for bits, which in ((0, 3), (5, 3), (5, 1), (5, 2), (3, 3), (7, 3), (9, 3)):
    env = dict(os.environ, MRB_SORT_RAYS=str(bits), MRB_SORT_WHICH=str(which))
    o = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], capture_output=True, text=True, env=env)
    ms = [l.split()[1:] for l in o.stdout.splitlines() if l.startswith("MS_PER_SPP")]
    out["bits%d_which%d" % (bits, which)] = [float(x) for x in ms[0]] if ms else o.stderr[-300:]
    print(json.dumps(out), flush=True)
