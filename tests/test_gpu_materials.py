"""GPU (B200): (Mt)Refract and (Mt)Unreal (Tracer/MaterialsDefault.hpp:L232-760) in the wavefront path tracer — closed
forms, the estimator oracle, and images rendered by the unmodified reference (RGB and spectral, the latter with
dispersion) through the TracerI plugin."""
import os

import numpy as np
import pytest

import oracle_lib as O
from mray_b200 import capi, scenes
from test_oracle_pt import glossy

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "mray_b200", "lib", "libTracerDLL_B200.so")
bm = lambda x, k: x.reshape(x.shape[0] // k, k, x.shape[1] // k, k, 3).mean(axis=(1, 3))
rel = lambda a, b: float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))


def glossy_accel(ctx):
    c, tm, alb, mtype, mparams = glossy()
    order = np.argsort(np.where(tm < 0, 99, tm), kind="stable")
    idx = np.ascontiguousarray(c["indices"][order]); m = tm[order]
    ranges, keys = [], []
    for k in np.unique(m):
        w = np.nonzero(m == k)[0]
        ranges.append([w[0], w[-1] + 1]); keys.append(capi.light_key(0) if k < 0 else int(k))
    acc = capi.Accelerator(ctx, c["positions"], idx, prim_ranges=ranges, light_or_mat_keys=keys)
    return c, idx, m, acc, alb, mtype, mparams


def test_refract_closed_form_and_unknown_type(gpu_ctx):
    L = 3.0
    pane1 = np.array([[-5, -5, 1], [5, -5, 1], [5, 5, 1], [-5, 5, 1]], np.float32)
    pane2 = pane1 + np.array([0, 0, -1], np.float32)
    light = np.array([[-30, -30, -3], [30, -30, -3], [30, 30, -3], [-30, 30, -3]], np.float32)
    pos = np.ascontiguousarray(np.concatenate([pane1, pane2, light]))
    idx = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7], [8, 9, 10], [8, 10, 11]], np.uint32)
    acc = capi.Accelerator(gpu_ctx, pos, idx, prim_ranges=[[0, 4], [4, 6]], light_or_mat_keys=[0, capi.light_key(0)])
    cam = dict(eye=(0.0, 0.0, 4.0), gaze=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fov_y_deg=20.0)
    mp = np.zeros((1, 8), np.float32); mp[0, 0] = 1.3; mp[0, 4] = 1.3
    for mode in ("Pure", "WithNextEventEstimation", "WithNEEAndMIS"):
        r = capi.Renderer(gpu_ctx, acc, 12, 6, [[0.5, 0.5, 0.5]], [L, L, L], cam, 8, 8, 16, sample_mode=mode, material_type=[2], material_params=mp)
        img, st = r.render(); r.close()
        assert np.allclose(img, L, rtol=1e-5), (mode, img.min(), img.max())
    with pytest.raises(capi.MrbError):
        capi.Renderer(gpu_ctx, acc, 12, 6, [[0.5, 0.5, 0.5]], [L, L, L], cam, 8, 8, 1, material_type=[2])       # no material_params
    with pytest.raises(capi.MrbError):
        capi.Renderer(gpu_ctx, acc, 12, 6, [[0.5, 0.5, 0.5]], [L, L, L], cam, 8, 8, 1, material_type=[9], material_params=mp)
    acc.close()


def test_glossy_matches_oracle_in_every_sample_mode(gpu_ctx):
    c, idx, m, acc, alb, mtype, mparams = glossy_accel(gpu_ctx)
    res = 32
    ref = O.oracle_render(c["positions"], idx, m, alb, c["radiance"], c["camera"], res, res, 8192, sample_mode=2, seed=3,
                          material_type=mtype, material_params=mparams)
    mask = ref.max(axis=-1) < 5.0
    for mode, spp in (("WithNEEAndMIS", 65536), ("WithNextEventEstimation", 32768), ("Pure", 65536)):
        r = capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], alb, c["radiance"], c["camera"], res, res, spp,
                          sample_mode=mode, seed=12, material_type=mtype, material_params=mparams)
        img, st = r.render(batch=64); r.close()
        assert st.finished
        assert np.allclose(img[mask].mean(axis=0), ref[mask].mean(axis=0), rtol=0.03), (mode, img[mask].mean(axis=0), ref[mask].mean(axis=0))
        if mode == "WithNEEAndMIS":
            assert rel(bm(img, 4), bm(ref, 4)) <= 2e-3, rel(bm(img, 4), bm(ref, 4))
    acc.close()


@pytest.mark.skipif(not (os.path.exists(PLUGIN) and O.driver_available()), reason="plugin / driver were not prebuilt")
@pytest.mark.parametrize("renderer,name", [("PathTracerRGB", "cornell64_glossy_spp16384"), ("PathTracerSpectral", "cornell64_glossy_spectral_spp16384")])
def test_glossy_through_tracer_interface_against_reference(renderer, name):
    """CreateMaterialGroup("(Mt)Unreal" / "(Mt)Refract") + their attribute pushes through TracerI, against the reference's
    own render of the same calls. The spectral render exercises dispersion (a refracted path keeps one wavelength)."""
    path = os.path.join(ROOT, "tests", "golden", f"render_{name}.npz")
    if not os.path.exists(path):
        pytest.skip("golden image was not generated")
    if renderer == "PathTracerSpectral":
        from mray_b200 import spectral
        if not spectral.available():
            pytest.skip("spectral LUT was not generated")
    ref = np.load(path)["img"].astype(np.float32)
    c = scenes.cornell_glossy()
    b = O.batched_scene(c["positions"], c["indices"], c["material"])
    img, w, st = O.driver_render(PLUGIN, b, c["albedo"], 3, c["radiance"], c["camera"], 64, 64, 32768, renderer=renderer, seed=33,
                                 burst_size=64, material_kind=c["material_type"], material_params=c["material_params"])
    assert np.allclose(w, 32768, rtol=1e-3)
    err = rel(bm(img, 8), bm(ref, 8))
    assert err <= 2e-3, err
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.015), (img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)))


def test_smooth_normals_against_reference(gpu_ctx):
    """a19: shading normals from interpolated tangent frames. C-ABI with vertexTBN (quaternions) and through TracerI
    (the NORMAL attribute pushed as quaternions), both against the reference's render of scenes.cornell_sphere."""
    path = os.path.join(ROOT, "tests", "golden", "render_cornell64_sphere_spp16384.npz")
    if not os.path.exists(path):
        pytest.skip("golden image was not generated")
    ref = np.load(path)["img"].astype(np.float32)
    c = scenes.cornell_sphere()
    order = np.argsort(c["material"], kind="stable")
    idx = np.ascontiguousarray(c["indices"][order]); mat = c["material"][order]
    ranges, keys = [], []
    for m in np.unique(mat):
        w = np.nonzero(mat == m)[0]
        ranges.append([w[0], w[-1] + 1]); keys.append(capi.light_key(0) if m == 3 else int(m))
    acc = capi.Accelerator(gpu_ctx, c["positions"], idx, prim_ranges=ranges, light_or_mat_keys=keys)
    imgs = {}
    for name, kw in (("tbn", dict(vertex_tbn=O.normals_to_tbn(c["normals"]))), ("linear", dict(vertex_normals=c["normals"])), ("flat", dict())):
        r = capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"], 64, 64, 16384, seed=51, **kw)
        imgs[name], st = r.render(batch=32); r.close()
    acc.close()
    assert rel(bm(imgs["tbn"], 2), bm(ref, 2)) <= 1e-3, rel(bm(imgs["tbn"], 2), bm(ref, 2))
    assert rel(bm(imgs["linear"], 2), bm(ref, 2)) <= 1e-3          # linear interpolation of the normals: indistinguishable at this tessellation
    assert rel(bm(imgs["flat"], 2), bm(ref, 2)) > 2 * rel(bm(imgs["tbn"], 2), bm(ref, 2))   # ... but flat shading is not
    if os.path.exists(PLUGIN) and O.driver_available():
        b = O.batched_scene(c["positions"], c["indices"], c["material"], normals=c["normals"])
        img, w, st = O.driver_render(PLUGIN, b, c["albedo"], 3, c["radiance"], c["camera"], 64, 64, 16384, seed=52, burst_size=64)
        assert np.allclose(w, 16384, rtol=1e-3)
        assert rel(bm(img, 2), bm(ref, 2)) <= 1e-3, rel(bm(img, 2), bm(ref, 2))
        assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.01)


def test_normal_map_against_reference(gpu_ctx):
    """Normal maps (mrb_render_desc.normalTexture; through TracerI the optional "normalMap" attribute of (Mt)Lambert) against the
    reference's render of scenes.cornell_normal_map."""
    path = os.path.join(ROOT, "tests", "golden", "render_cornell64_normalmap_spp16384.npz")
    if not os.path.exists(path):
        pytest.skip("golden image was not generated")
    ref = np.load(path)["img"].astype(np.float32)
    c = scenes.cornell_normal_map()
    order = np.argsort(c["material"], kind="stable")
    idx = np.ascontiguousarray(c["indices"][order]); mat = c["material"][order]
    ranges, keys = [], []
    for m in np.unique(mat):
        w = np.nonzero(mat == m)[0]
        ranges.append([w[0], w[-1] + 1]); keys.append(capi.light_key(0) if m == 3 else int(m))
    acc = capi.Accelerator(gpu_ctx, c["positions"], idx, prim_ranges=ranges, light_or_mat_keys=keys)
    args = (gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"], 64, 64)
    kw = dict(textures=[c["normal_texture"]], albedo_texture=[-1, -1, -1], vertex_uvs=c["uvs"], vertex_tbn=O.normals_to_tbn(c["normals"]))
    r = capi.Renderer(*args, 16384, seed=61, normal_texture=c["normal_map"][:3], **kw)
    img, st = r.render(batch=32); r.close()
    assert rel(bm(img, 2), bm(ref, 2)) <= 1e-3, rel(bm(img, 2), bm(ref, 2))
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.01)
    r = capi.Renderer(*args, 1024, seed=62, **kw)           # the same scene without the map is a visibly different image
    flat, _ = r.render(batch=32); r.close()
    assert rel(bm(flat, 4), bm(ref, 4)) > 5 * rel(bm(img, 4), bm(ref, 4))
    with pytest.raises(capi.MrbError):                      # a normal map needs the tangent frames
        capi.Renderer(*args, 1, normal_texture=c["normal_map"][:3], textures=[c["normal_texture"]], albedo_texture=[-1, -1, -1], vertex_uvs=c["uvs"])
    acc.close()
    if os.path.exists(PLUGIN) and O.driver_available():
        b = O.batched_scene(c["positions"], c["indices"], c["material"], normals=c["normals"], uvs=c["uvs"])
        pimg, w, pst = O.driver_render(PLUGIN, b, c["albedo"], 3, c["radiance"], c["camera"], 64, 64, 16384, seed=63, burst_size=64,
                                       textures=[c["normal_texture"]], normal_map=c["normal_map"])
        assert np.allclose(w, 16384, rtol=1e-3)
        assert rel(bm(pimg, 2), bm(ref, 2)) <= 1e-3, rel(bm(pimg, 2), bm(ref, 2))
        assert np.allclose(pimg.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.01)
