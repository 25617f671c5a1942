"""BASELINE config 4 with a REAL 64-material mix: 1 017 instances of 10 meshes (9.87 M instanced triangles), 64
materials round-robin per instance — all (Mt)Lambert, or 40 Lambert + 8 (Mt)Reflect + 8 (Mt)Refract + 8 (Mt)Unreal —, 3840x2160,
(R)PathTracerSpectral NEE+MIS rr [3, 8], with the material-key ray sort off (one fused shading kernel) and on. Device-timed; JSON."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mray_b200
from mray_b200 import capi, scenes, spectral

ctx = mray_b200.Context(0); stream = torch.cuda.current_stream(); ctx.set_stream(stream)
ev = lambda: torch.cuda.Event(enable_timing=True)
f = scenes.instanced_field()
accs = [mray_b200.Accelerator(ctx, torch.from_numpy(mp).cuda(), torch.from_numpy(mi.view(np.int32)).cuda(),
                              prim_ranges=[[0, mi.shape[0]]], light_or_mat_keys=[0]) for mp, mi in f["meshes"]]
insts = [(accs[m], M, [capi.light_key(0) if mat < 0 else mat]) for m, M, mat in f["instances"]]
scene = mray_b200.Scene(ctx, insts)
n_mat = len(f["albedo"])
rng = np.random.default_rng(4)
mixed_type = np.zeros(n_mat, np.uint8)
mixed_type[5::8] = 1; mixed_type[6::8] = 2; mixed_type[7::8] = 3          # every 8th material: Reflect / Refract / Unreal
params = np.zeros((n_mat, 8), np.float32)
params[mixed_type == 2, 0] = 1.0; params[mixed_type == 2, 4] = 1.5; params[mixed_type == 2, 5] = 0.004     # air -> Cauchy glass
params[mixed_type == 3, 0] = rng.uniform(0.1, 0.6, (mixed_type == 3).sum()); params[mixed_type == 3, 1] = 0.5; params[mixed_type == 3, 2] = rng.integers(0, 2, (mixed_type == 3).sum())
W4, H4, spp = 3840, 2160, 2
sp = capi.Spectrum(ctx, spectral.load(), "HyperbolicPBRT")
out = {"workload": "config 4: 1017 instances of 10 meshes, 9.87 M instanced triangles, 64 materials, 3840x2160, (R)PathTracerSpectral NEE+MIS rr[3,8]",
       "material_mix": {"Lambert": int((mixed_type == 0).sum()), "Reflect": int((mixed_type == 1).sum()), "Refract": int((mixed_type == 2).sum()),
                        "Unreal": int((mixed_type == 3).sum())}}
for mix_label, mt in (("all_lambert", None), ("mixed", mixed_type)):
    for label, part in (("fused_shading", False), ("material_key_sort", True)):
        kw = {} if mt is None else dict(material_type=mt, material_params=params)
        r4 = mray_b200.Renderer(ctx, scene, 0, 0, f["albedo"], f["radiance"], f["camera"], W4, H4, spp, sample_mode="WithNEEAndMIS",
                                rr_range=(3, 8), seed=1, partition_rays=part, spectrum=sp, **kw)
        r4.iterate(2); torch.cuda.synchronize()
        e0, e1 = ev(), ev(); e0.record(stream)
        while True:
            r4.iterate(8); st = r4.stats()
            if st.finished: break
        e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out[f"{mix_label}.{label}"] = {"ms_per_spp": round(ms / spp, 2), "mrays_s": round((st.closestRays + st.shadowRays) / ms / 1e3, 1),
                                       "mpaths_s": round(st.pathsCompleted / ms / 1e3, 1), "iterations": int(st.iterations),
                                       "rays_per_path": round((st.closestRays + st.shadowRays) / max(1, st.pathsCompleted), 2)}
        r4.close()
out["used_device_mb"] = round(ctx.used_device_memory / 2**20, 1)
print(json.dumps(out))
