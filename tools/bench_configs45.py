"""BASELINE configs 4 and 5 on one B200 (device-timed, CUDA events), JSON to stdout.
  config 4: ~10 M-triangle instanced scene (two-level BVH, 64 materials), 3840x2160 path tracing
  config 5: BVH build sweep 100 K .. 50 M triangles, incoherent traversal vs bounce depth
usage: python tools/bench_configs45.py [max_build_tris]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mray_b200
from mray_b200 import capi, scenes

ctx = mray_b200.Context(0); stream = torch.cuda.current_stream(); ctx.set_stream(stream)
out = {}
ev = lambda: torch.cuda.Event(enable_timing=True)

# ---------------- config 5a: build sweep ----------------
max_tris = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
sweep = []
for n in (100_000, 300_000, 1_000_000, 3_000_000, 10_000_000, 30_000_000, 50_000_000):
    if n > max_tris: break
    p, i = scenes.random_soup(n)
    dp, di = torch.from_numpy(p).cuda(), torch.from_numpy(i.view(np.int32)).cuda()
    best = 1e30
    try:
        for _ in range(3):
            a = capi.Accelerator(ctx, dp, di); best = min(best, a.info.buildMs); wide = a.info.wideNodeCount; byts = a.info.deviceBytes; a.close()
    except Exception as e:   # report, keep sweeping
        sweep.append({"triangles": n, "error": str(e)[:200]}); continue
    sweep.append({"triangles": n, "build_ms": round(best, 3), "mtris_s": round(n / best / 1e3, 1), "wide_nodes": int(wide),
                  "device_mb": round(byts / 2**20, 1)})
    del dp, di
    torch.cuda.empty_cache()
out["config5_build_sweep"] = sweep

# ---------------- config 5b: traversal vs bounce depth (rays of the path tracer, per iteration) ----------------
p, i = scenes.arcade_mesh()
pidx, pranges, pkeys, palb, prad, _ = scenes.arcade_materials(p, i)
dp = torch.from_numpy(p).cuda()
acc = mray_b200.Accelerator(ctx, dp, torch.from_numpy(pidx.view(np.int32)).cuda(), prim_ranges=pranges, light_or_mat_keys=pkeys)
W, H = 1920, 1080
r = mray_b200.Renderer(ctx, acc, p.shape[0], pidx.shape[0], palb, prad, scenes.ARCADE_CAMERA, W, H, 1, sample_mode="WithNEEAndMIS",
                       rr_range=(8, 8), seed=0, partition_rays=True)
depth = []
prev = r.stats()
for it in range(8):   # with spp = 1 every iteration is one bounce of the same wave of paths
    e0, e1 = ev(), ev()
    e0.record(stream); r.iterate(1); e1.record(stream); torch.cuda.synchronize()
    st = r.stats()
    rays = (st.closestRays - prev.closestRays) + (st.shadowRays - prev.shadowRays)
    depth.append({"bounce": it, "closest_rays": int(st.closestRays - prev.closestRays), "shadow_rays": int(st.shadowRays - prev.shadowRays),
                  "iteration_ms": round(e0.elapsed_time(e1), 3), "mrays_s": round(rays / e0.elapsed_time(e1) / 1e3, 1)})
    prev = st
out["config5_bounce_depth"] = {"workload": "arcade 264K tris, 1080p, 1 spp wave, NEE+MIS, whole iteration (reload+trace+sort+shade+shadow trace+film)",
                               "per_bounce": depth}
r.close()

# ---------------- config 4: 1000 instances of 10 meshes (1 K .. 100 K triangles), ~10 M instanced triangles, 64 materials
# round-robin per instance (mrb_instance_desc.lightOrMatKeys), 3840x2160, spectral NEE+MIS ----------------
acc.close()
from mray_b200 import spectral
f = scenes.instanced_field()
e0, e1 = ev(), ev(); e0.record(stream)
accs = [mray_b200.Accelerator(ctx, torch.from_numpy(mp).cuda(), torch.from_numpy(mi.view(np.int32)).cuda(),
                              prim_ranges=[[0, mi.shape[0]]], light_or_mat_keys=[0]) for mp, mi in f["meshes"]]
insts = [(accs[m], M, [capi.light_key(0) if mat < 0 else mat]) for m, M, mat in f["instances"]]
scene = mray_b200.Scene(ctx, insts)
e1.record(stream); torch.cuda.synchronize()
build_ms = e0.elapsed_time(e1)
W4, H4 = 3840, 2160
spp = 2
sp = capi.Spectrum(ctx, spectral.load(), "HyperbolicPBRT") if spectral.available() else None
cfg4 = {}
for label, part in (("fused_shading", False), ("material_key_sort", True)):
    r4 = mray_b200.Renderer(ctx, scene, 0, 0, f["albedo"], f["radiance"], f["camera"], W4, H4, spp, sample_mode="WithNEEAndMIS",
                            rr_range=(3, 8), seed=1, partition_rays=part, spectrum=sp)
    r4.iterate(2); torch.cuda.synchronize()
    e0, e1 = ev(), ev(); e0.record(stream)
    while True:
        r4.iterate(8); st = r4.stats()
        if st.finished: break
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    cfg4[label] = {"ms_per_spp": round(ms / spp, 2), "mrays_s": round((st.closestRays + st.shadowRays) / ms / 1e3, 1),
                   "mpaths_s": round(st.pathsCompleted / ms / 1e3, 1), "iterations": int(st.iterations)}
    r4.close()
out["config4_instanced_4k"] = {"instances": len(insts), "distinct_meshes": len(accs), "triangles_instanced": f["triangles_instanced"],
                               "triangles_stored": int(sum(m[1].shape[0] for m in f["meshes"])), "materials": int(len(f["albedo"])),
                               "renderer": "PathTracerSpectral" if sp is not None else "PathTracerRGB", "resolution": [W4, H4], "spp": spp,
                               "paths_in_flight": W4 * H4, "scene_build_ms_incl_upload": round(build_ms, 2), "tlas_build_ms": round(scene.build_ms, 3) if hasattr(scene, "build_ms") else None,
                               **cfg4, "used_device_mb": round(ctx.used_device_memory / 2**20, 1)}
scene.close()
for a in accs: a.close()
print(json.dumps(out))
