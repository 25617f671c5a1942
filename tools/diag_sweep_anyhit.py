"""Diagnostic (GPU box): the scheduling knobs of the ANY-HIT wide kernel on their own (env MRB_TRI_DIV_ANY, MRB_FETCH_THR_ANY; the
closest-hit kernel keeps 8 / 24) on the path-tracer workload (config-3 flavour, 1080p, spectral, 16 spp), with the per-kernel profile.
usage: python tools/diag_sweep_anyhit.py   (parent: one child per pair, prints one JSON line)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    import mray_b200
    from mray_b200 import scenes, spectral
    ctx = mray_b200.Context(0); stream = torch.cuda.current_stream(); ctx.set_stream(stream)
    p, i = scenes.arcade_mesh()
    pidx, pranges, pkeys, palb, prad, _ = scenes.arcade_materials(p, i)
    acc = mray_b200.Accelerator(ctx, torch.from_numpy(p).cuda(), torch.from_numpy(pidx.view(np.int32)).cuda(), prim_ranges=pranges, light_or_mat_keys=pkeys)
    spec = mray_b200.Spectrum(ctx, spectral.load())
    best = 1e30
    for rep in range(2):
        r = mray_b200.Renderer(ctx, acc, p.shape[0], pidx.shape[0], palb, prad, scenes.ARCADE_CAMERA, 1920, 1080, 16, sample_mode="WithNEEAndMIS",
                               rr_range=(3, 8), seed=7, spectrum=spec)
        r.iterate(2); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        while True:
            r.iterate(8)
            if r.stats().finished: break
        e1.record(stream); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 16)
        r.close()
    print("MS_PER_SPP", round(best, 4))
    sys.exit(0)
out = {}
for td, ft in [tuple(map(int, a.split(","))) for a in sys.argv[1:]] or ((8, 24), (8, 28), (8, 32), (8, 20), (4, 24), (16, 24), (16, 28), (32, 28), (4, 28)):
    env = dict(os.environ, MRB_TRI_DIV_ANY=str(td), MRB_FETCH_THR_ANY=str(ft))
    o = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], capture_output=True, text=True, env=env)
    ms = [float(l.split()[1]) for l in o.stdout.splitlines() if l.startswith("MS_PER_SPP")]
    out["any_triDiv%d_fetchThr%d" % (td, ft)] = ms[0] if ms else o.stderr[-200:]
    print(json.dumps(out), flush=True)
