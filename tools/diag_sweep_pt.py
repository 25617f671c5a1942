"""Diagnostic (GPU box): the wide-traversal scheduling parameters (env MRB_TRI_DIV, MRB_FETCH_THR; read once per process)
on the PATH TRACER workload (config-3 flavour, 1080p, 4 spp) — the bench sweep of round 1 used config 2's rays only.
usage: python tools/diag_sweep_pt.py            (parent: runs one child per parameter pair, prints one JSON line)"""
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
    best = 1e30
    for rep in range(2):
        r = mray_b200.Renderer(ctx, acc, p.shape[0], pidx.shape[0], palb, prad, scenes.ARCADE_CAMERA, 1920, 1080, 4, sample_mode="WithNEEAndMIS",
                               rr_range=(3, 8), seed=rep)
        r.iterate(2); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        while True:
            r.iterate(8)
            if r.stats().finished: break
        e1.record(stream); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 4)
        r.close()
    print("MS_PER_SPP", round(best, 4))
    sys.exit(0)
out = {}
for td, ft in ((8, 24), (4, 24), (16, 24), (8, 20), (8, 28), (4, 20), (16, 28)):
    env = dict(os.environ, MRB_TRI_DIV=str(td), MRB_FETCH_THR=str(ft))
    o = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], capture_output=True, text=True, env=env).stdout
    ms = [float(l.split()[1]) for l in o.splitlines() if l.startswith("MS_PER_SPP")]
    out["triDiv%d_fetchThr%d" % (td, ft)] = ms[0] if ms else None
    print(json.dumps(out), flush=True)
