"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck): build, casts, two-level scene,
partitioner, spectral + Z-Sobol render. Keep it tiny: the sanitizer slows kernels by 10-100x."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mray_b200
from mray_b200 import capi, scenes, spectral

ctx = mray_b200.Context(0)
p, i = scenes.arcade_mesh(3000)
acc = capi.Accelerator(ctx, p, i)
rays = scenes.pinhole_rays(48, 27, **scenes.ARCADE_CAMERA); n = rays.shape[0]
k = np.full((n, 4), 0xFFFFFFFF, np.uint32); h = np.zeros((n, 2), np.float32)
acc.cast_rays(k, h, rays.copy(), None, capi.MRB_TRACE_WIDE)
acc.cast_rays(k, h, rays.copy(), None, capi.MRB_TRACE_BINARY_EXACT)
bits = np.full(((n + 31) // 32,), 0xFFFFFFFF, np.uint32)
acc.cast_visibility_rays(bits, rays.copy(), None, capi.MRB_TRACE_WIDE)
idx = np.arange(0, n, 3, dtype=np.uint32)
acc.cast_rays(k, h, rays.copy(), idx, capi.MRB_TRACE_WIDE)
M = np.array([[0, 0, 1, 0.5], [0, 1, 0, 0.0], [-1, 0, 0, 0.25]], np.float64)
scene = capi.Scene(ctx, [(acc, None), (acc, M), (acc, M * 0.5)])
scene.cast_rays(k, h, rays.copy()); scene.cast_visibility_rays(bits, rays.copy())
keys = np.random.default_rng(0).integers(0, 1 << 10, size=5000).astype(np.uint32); vals = np.arange(5000, dtype=np.uint32)
ctx.radix_sort_pairs(keys, vals)
c = scenes.cornell_box()
order = np.argsort(c["material"], kind="stable"); tri = np.ascontiguousarray(c["indices"][order]); mat = c["material"][order]
rg, ky = [], []
for m in np.unique(mat):
    w = np.nonzero(mat == m)[0]; rg.append([w[0], w[-1] + 1]); ky.append(capi.light_key(0) if m == 3 else int(m))
cacc = capi.Accelerator(ctx, c["positions"], tri, prim_ranges=rg, light_or_mat_keys=ky)
spec = capi.Spectrum(ctx, spectral.load()) if spectral.available() else None
for sampler, part in (("Independent", True), ("ZSobol", False), ("Sobol", False)):
    r = capi.Renderer(ctx, cacc, c["positions"].shape[0], tri.shape[0], c["albedo"][:3], c["radiance"], c["camera"], 16, 16, 4,
                      seed=1, partition_rays=part, spectrum=spec, sampler=sampler)
    img, st = r.render(batch=4); r.close()
    assert st.finished and np.isfinite(img).all()
cs = capi.Scene(ctx, [(cacc, None), (cacc, np.array([[1, 0, 0, 3.0], [0, 1, 0, 0], [0, 0, 1, 0]], np.float64))])
r = capi.Renderer(ctx, cs, 0, 0, c["albedo"][:3], c["radiance"], c["camera"], 16, 16, 4, seed=2)
img, st = r.render(batch=4); r.close()
print("sanitize workload done", float(img.mean()))
