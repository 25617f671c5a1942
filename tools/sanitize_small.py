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
# round 2: glossy materials + skysphere, alpha maps, normal map, mip chains (generated, clamped, explicit) with ray cones, the texture taps,
# 2-D distributions, spectral LUT columns, the 8-lane and the serial collapse
g = scenes.cornell_glossy()
go = np.argsort(g["material"], kind="stable"); gtri = np.ascontiguousarray(g["indices"][go]); gmat = g["material"][go]
rg, ky = [], []
for m in np.unique(gmat):
    w = np.nonzero(gmat == m)[0]; rg.append([w[0], w[-1] + 1]); ky.append(capi.light_key(0) if m == 3 else int(m))
gacc = capi.Accelerator(ctx, g["positions"], gtri, prim_ranges=rg, light_or_mat_keys=ky)
sky = scenes.sky_texture()
r = capi.Renderer(ctx, gacc, g["positions"].shape[0], gtri.shape[0], g["albedo"], g["radiance"], g["camera"], 16, 16, 4, seed=3, spectrum=spec,
                  material_type=g["material_type"], material_params=g["material_params"], textures=[sky], albedo_texture=np.full(g["albedo"].shape[0], -1),
                  boundary=dict(type="Skysphere_CoOcta", texture=0))
img, st = r.render(batch=4); r.close(); gacc.close()
assert st.finished and np.isfinite(img).all()
for kind in ("explicit", "gen_glossy", "sphere_mirror"):
    m = scenes.cornell_mips(kind)
    mo = np.argsort(m["material"], kind="stable"); mtri = np.ascontiguousarray(m["indices"][mo]); mmat = m["material"][mo]
    rg, ky = [], []
    for v in np.unique(mmat):
        w = np.nonzero(mmat == v)[0]; rg.append([w[0], w[-1] + 1]); ky.append(capi.light_key(0) if v == 3 else int(v))
    macc = capi.Accelerator(ctx, m["positions"], mtri, prim_ranges=rg, light_or_mat_keys=ky)
    tex = [dict(t, gen_mips=m.get("gen_mips"), clamp_res=(32 if kind == "gen_glossy" else None)) for t in m["textures"]]
    kw = {}
    if "material_type" in m: kw["material_type"] = m["material_type"]
    if "material_params" in m: kw["material_params"] = m["material_params"]
    r = capi.Renderer(ctx, macc, m["positions"].shape[0], mtri.shape[0], m["albedo"], m["radiance"], m["camera"], 16, 16, 4, seed=4,
                      textures=tex, albedo_texture=m["albedo_texture"], vertex_uvs=m["uvs"], texture_lod_mode=(1 if kind == "explicit" else 0), **kw)
    img, st = r.render(batch=4); r.close(); macc.close()
    assert st.finished and np.isfinite(img).all()
a = scenes.cornell_alpha()
ao = np.argsort(a["material"], kind="stable"); atri = np.ascontiguousarray(a["indices"][ao]); amat = a["material"][ao]
rg, ky, am = [], [], []
for v in np.unique(amat):
    w = np.nonzero(amat == v)[0]; rg.append([w[0], w[-1] + 1]); ky.append(capi.light_key(0) if v == 3 else int(v)); am.append(int(a["alpha_map"][int(v)]))
aacc = capi.Accelerator(ctx, a["positions"], atri, prim_ranges=rg, light_or_mat_keys=ky, vertex_uvs=a["uvs"], alpha_textures=[a["alpha_texture"]], range_alpha_map=am)
r = capi.Renderer(ctx, aacc, a["positions"].shape[0], atri.shape[0], a["albedo"], a["radiance"], a["camera"], 16, 16, 4, seed=5)
img, st = r.render(batch=4); r.close(); aacc.close()
rng = np.random.default_rng(2)
t = dict(data=rng.random((20, 12, 4), dtype=np.float32), gen_mips=("Mitchell-Netravali", 2.0), clamp_res=8)
capi.texture_mip_chain(ctx, t)
uv = rng.random((100, 2)).astype(np.float32)
capi.texture_sample_lod(ctx, t, uv, lod=rng.random(100).astype(np.float32) * 5)
capi.texture_sample_lod(ctx, t, uv, dpdx=rng.standard_normal((100, 2)).astype(np.float32), dpdy=rng.standard_normal((100, 2)).astype(np.float32), lod_mode=1)
capi.texture_convert(ctx, dict(data=(rng.random((9, 7, 4)) * 255).astype(np.uint8), gamma=2.2))
f = rng.random((13, 29)).astype(np.float32)
cx, cy = capi.dist2d_build(ctx, f)
capi.dist2d_sample(ctx, cx, cy, rng.random((64, 2)).astype(np.float32))
b1 = capi.Accelerator(ctx, p, i, flags=capi.MRB_BUILD_SERIAL_COLLAPSE); b1.export_wide(); b1.close()
print("sanitize workload done", float(img.mean()))
