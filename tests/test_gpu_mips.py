"""GPU (B200): texture mip chains and level-of-detail reads — mip generation (KGenerateMipLevel), the level /
gradient reads of the shading kernel (SampleTextureLod / SampleTextureGrad) against the oracle and against golden vectors made by the
reference's own TextureMemory + TracerTexView (tests/golden/texture_mips.npz), and renders whose textures carry mip levels."""
import numpy as np
import pytest

import oracle_lib as O
from mray_b200 import capi
from test_oracle_mips import CASES

pytestmark = pytest.mark.gpu

# generated levels whose filter weight needs exp(): the device's expf may differ from the host libm's by an ulp
EXP_FILTERS = ("Gaussian",)


@pytest.mark.parametrize("name", sorted(CASES))
def test_mip_chain_matches_the_oracle(gpu_ctx, name):
    t, _ = CASES[name]
    chain, count = capi.texture_mip_chain(gpu_ctx, t)
    ref, ref_count = O.mip_chain(t)
    assert count == ref_count and chain.shape == ref.shape and chain.dtype == ref.dtype
    exact = not (t.get("gen_mips") and t["gen_mips"][0] in EXP_FILTERS)
    if exact or chain.dtype == np.uint8:
        same = (chain == ref).mean()
        assert same == 1.0 if exact else same > 0.995, same     # unorm8: an ulp of a weight can flip a rounding on a rare texel
    else:
        assert np.allclose(chain, ref, rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("name", sorted(CASES))
def test_lod_reads_match_the_reference_goldens(gpu_ctx, name):
    t, v = CASES[name]
    got_lod = capi.texture_sample_lod(gpu_ctx, t, v["uv"], lod=v["lod"])
    got_grad = capi.texture_sample_lod(gpu_ctx, t, v["uv"], dpdx=v["dpdx"], dpdy=v["dpdy"], lod_mode=0)
    exact = not (t.get("gen_mips") and t["gen_mips"][0] in EXP_FILTERS)
    if exact:
        assert np.array_equal(got_lod, v["rgb_lod"])
        # gradient reads go through log2f: the device's differs from the host libm's by an ulp on a few arguments, which moves the
        # blend weight of the two levels by 2^-23 there
        assert np.allclose(got_grad, v["rgb_grad"], rtol=0, atol=1e-6) and (got_grad == v["rgb_grad"]).mean() > 0.9
    else:
        tol = dict(rtol=0, atol=3e-3) if t["data"].dtype == np.uint8 else dict(rtol=4e-6, atol=2e-7)
        assert np.allclose(got_lod, v["rgb_lod"], **tol) and np.allclose(got_grad, v["rgb_grad"], **tol)
        if t["data"].dtype == np.uint8:
            assert (got_lod == v["rgb_lod"]).mean() > 0.99


def test_device_lod_mode_scales_the_gradients_by_the_texture_size(gpu_ctx):
    t, v = CASES["explicit3_f32"]
    h, w, _ = t["data"].shape
    size = np.array([w, h], np.float32)
    host = capi.texture_sample_lod(gpu_ctx, t, v["uv"], dpdx=v["dpdx"] * size, dpdy=v["dpdy"] * size, lod_mode=0)
    dev = capi.texture_sample_lod(gpu_ctx, t, v["uv"], dpdx=v["dpdx"], dpdy=v["dpdy"], lod_mode=1)
    assert np.array_equal(host, dev)
    ref = O.oracle_texture_sample_lod(t, v["uv"], dpdx=v["dpdx"], dpdy=v["dpdy"], lod_mode=1)
    assert np.allclose(dev, ref, rtol=0, atol=1e-6) and (dev == ref).mean() > 0.9


def test_resolution_clamp(gpu_ctx):
    """TracerParameters.clampedTexRes: final sizes, the filtered level 0 against the oracle (Mitchell-Netravali too), kept levels."""
    rng = np.random.default_rng(4)
    g = rng.random((40, 64, 4), dtype=np.float32)
    assert capi.texture_final_extent(gpu_ctx, dict(data=g, clamp_res=16)) == (16, 10, 1)
    assert capi.texture_final_extent(gpu_ctx, dict(data=g, clamp_res=20, gen_mips=("Tent", 1.0))) == (16, 10, 5)
    assert capi.texture_final_extent(gpu_ctx, dict(data=g, clamp_res=4096)) == (64, 40, 1)
    for filt in [("Mitchell-Netravali", 2.0), ("Tent", 1.0), ("Box", 0.75)]:
        t = dict(data=g, clamp_res=16, gen_mips=filt)
        chain, count = capi.texture_mip_chain(gpu_ctx, t)
        ref, ref_count = O.mip_chain(t)
        assert count == ref_count == 5 and np.allclose(chain, ref, rtol=0, atol=2e-6), filt
    lv = [rng.random((8, 8, 4), dtype=np.float32), rng.random((4, 4, 4), dtype=np.float32), rng.random((2, 2, 4), dtype=np.float32)]
    chain, count = capi.texture_mip_chain(gpu_ctx, dict(data=lv[0], mips=lv[1:], clamp_res=4))     # the levels that fit are kept
    assert count == 2 and np.array_equal(chain[:16].reshape(4, 4, 4), lv[1]) and np.array_equal(chain[16:].reshape(2, 2, 4), lv[2])
    conv = capi.texture_convert(gpu_ctx, dict(data=g, clamp_res=16))
    assert conv.shape == (10, 16, 4) and np.array_equal(conv.reshape(-1, 4), O.mip_chain(dict(data=g, clamp_res=16))[0])


def test_large_chain_generation(gpu_ctx):
    """A 1024 x 512 RGBA8 texture: 11 levels, every one equal to the oracle's on all but a stray texel."""
    rng = np.random.default_rng(8)
    t = dict(data=rng.integers(0, 256, size=(512, 1024, 4), dtype=np.uint8), gen_mips=("Gaussian", 2.0))
    chain, count = capi.texture_mip_chain(gpu_ctx, t)
    ref, ref_count = O.mip_chain(t)
    assert count == ref_count == 11
    assert (chain == ref).mean() > 0.999 and np.abs(chain.astype(np.int32) - ref.astype(np.int32)).max() <= 1


def test_bad_mip_descriptors_are_refused(gpu_ctx):
    base = np.zeros((4, 4, 4), np.float32)
    with pytest.raises(capi.MrbError):
        capi.texture_mip_chain(gpu_ctx, dict(data=base, mips=[np.zeros((2, 2, 4), np.float32), np.zeros((1, 1, 4), np.float32), np.zeros((1, 1, 4), np.float32)]))
    t, keep = capi._texture_desc(dict(data=base))
    t.generateMips, t.mipFilterType, t.mipFilterRadius = 1, 9, 1.0
    out = np.zeros((64, 4), np.float32); n = capi.C.c_uint32(0)
    assert gpu_ctx.lib.mrb_texture_mip_chain(gpu_ctx.handle, capi.C.byref(t), out.ctypes.data, capi.C.byref(n)) != 0


# ---------------------------------------------------------------------------------------------------------------------
# renders: ray cones + level selection in KShade<full>, against the reference's own images of scenes.cornell_mips
# ---------------------------------------------------------------------------------------------------------------------
import os

from mray_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "mray_b200", "lib", "libTracerDLL_B200.so")


def rel(a, b):
    return float(((a - b) ** 2).mean() / (b ** 2).mean())


def bm(img, k):
    h, w, c = img.shape
    return img.reshape(h // k, k, w // k, k, c).mean(axis=(1, 3))


def golden(kind):
    path = os.path.join(ROOT, "tests", "golden", f"render_cornell64_mips_{kind}_spp16384.npz")
    if not os.path.exists(path):
        pytest.skip("golden image was not generated")
    return np.load(path)["img"].astype(np.float32)


def mip_scene_renderer(ctx, kind, res, spp, seed, strip_mips=False, **kw):
    c = scenes.cornell_mips(kind)
    order = np.argsort(c["material"], kind="stable")
    idx = np.ascontiguousarray(c["indices"][order]); mat = c["material"][order]
    ranges, keys = [], []
    for m in np.unique(mat):
        w = np.nonzero(mat == m)[0]
        ranges.append([w[0], w[-1] + 1]); keys.append(capi.light_key(0) if m == 3 else int(m))
    acc = capi.Accelerator(ctx, c["positions"], idx, prim_ranges=ranges, light_or_mat_keys=keys)
    textures = [dict(t, gen_mips=c.get("gen_mips")) for t in c["textures"]]
    if strip_mips:
        textures = [dict(t, mips=None, gen_mips=None) for t in textures]
    extra = {}
    if "material_type" in c:
        extra["material_type"] = c["material_type"]
    if "material_params" in c:
        extra["material_params"] = c["material_params"]
    if "normals" in c:
        extra["vertex_tbn"] = O.normals_to_tbn(c["normals"])
    r = capi.Renderer(ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"], c["radiance"], c["camera"], res, res, spp, seed=seed,
                      textures=textures, albedo_texture=c["albedo_texture"], vertex_uvs=c["uvs"], **extra, **kw)
    img, st = r.render(batch=64)
    assert st.finished
    r.close(); acc.close()
    return img


@pytest.mark.parametrize("kind,tol", [("explicit", 1e-3), ("sphere_mirror", 1e-3), ("gen_glossy", 2e-3)])
def test_mip_mapped_render_against_reference(gpu_ctx, kind, tol):
    ref = golden(kind)
    img = mip_scene_renderer(gpu_ctx, kind, 64, 32768, seed=71)
    e = rel(bm(img, 2), bm(ref, 2))
    assert e <= tol, e
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.012), (img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)))
    if kind != "gen_glossy":
        # the levels matter: reading level 0 everywhere gives another image (a generated chain averages to the same colours)
        base = mip_scene_renderer(gpu_ctx, kind, 64, 2048, seed=72, strip_mips=True)
        assert rel(bm(base, 2), bm(ref, 2)) > 20 * max(e, 1e-4)


def test_mip_mapped_render_is_repeatable_and_lod_modes_differ(gpu_ctx):
    a = mip_scene_renderer(gpu_ctx, "explicit", 32, 64, seed=5)
    b = mip_scene_renderer(gpu_ctx, "explicit", 32, 64, seed=5)
    assert np.allclose(a, b, rtol=2e-5, atol=1e-6)       # the order of the film's float additions is free
    dev = mip_scene_renderer(gpu_ctx, "explicit", 32, 64, seed=5, texture_lod_mode=1)     # texel-space gradients: 5 levels coarser
    assert rel(dev, a) > 1e-2
    # ... and the oracle agrees with the device mode too
    c = scenes.cornell_mips("explicit")
    tm = np.where(c["material"] == 3, -1, c["material"].astype(np.int32))
    ref = O.oracle_render(c["positions"], c["indices"], tm, c["albedo"], c["radiance"], c["camera"], 32, 32, 4096, seed=9, textures=c["textures"],
                          albedo_texture=c["albedo_texture"], vertex_uvs=c["uvs"], texture_lod_mode=1)
    dev = mip_scene_renderer(gpu_ctx, "explicit", 32, 16384, seed=6, texture_lod_mode=1)
    assert rel(bm(dev, 2), bm(ref, 2)) <= 1.5e-3, rel(bm(dev, 2), bm(ref, 2))


@pytest.mark.skipif(not (os.path.exists(PLUGIN) and O.driver_available()), reason="plugin / driver were not prebuilt")
@pytest.mark.parametrize("kind,tol", [("explicit", 1e-3), ("gen_glossy", 2e-3)])
def test_mip_mapped_render_through_tracer_interface(kind, tol):
    """CreateTexture2D(size, mipCount) + PushTextureData per level, and TracerParameters.genMips / mipGenFilter, through the plugin."""
    ref = golden(kind)
    c = scenes.cornell_mips(kind)
    b = O.batched_scene(c["positions"], c["indices"], c["material"], normals=c.get("normals"), uvs=c["uvs"])
    kw = dict(textures=c["textures"], material_texture=c["albedo_texture"], gen_mips=c.get("gen_mips"))
    if "material_type" in c:
        kw.update(material_kind=c["material_type"], material_params=c["material_params"])
    img, w, st = O.driver_render(PLUGIN, b, c["albedo"], 3, c["radiance"], c["camera"], 64, 64, 32768, seed=73, burst_size=64, **kw)
    assert np.allclose(w, 32768, rtol=1e-3)
    e = rel(bm(img, 2), bm(ref, 2))
    assert e <= tol, e
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.012)


def test_mip_mapped_two_level_scene_matches_flat(gpu_ctx):
    """The same mip-mapped scene as instances under rigid (T)Single transforms: the ray cone's footprint, curvature and texture
    gradients are formed from world-space positions, so the levels — hence the image — do not move with the instances' local frames."""
    from test_gpu_render import _rigid
    c = scenes.cornell_mips("explicit")
    ref = golden("explicit")
    rng = np.random.default_rng(15)
    mat = c["material"]
    instances, accels, uvs = [], [], []
    for k, m in enumerate(np.unique(mat)):
        tri = c["indices"][mat == m]
        wpos = c["positions"][tri.reshape(-1)].astype(np.float64)            # unwelded world-space vertices
        M = None if k == 1 else _rigid(rng)
        if M is None:
            lpos = wpos
        else:
            inv = np.linalg.inv(np.vstack([M, [0, 0, 0, 1]]))
            lpos = wpos @ inv[:3, :3].T + inv[:3, 3]
        lpos = np.ascontiguousarray(lpos, np.float32)
        lidx = np.arange(lpos.shape[0], dtype=np.uint32).reshape(-1, 3)
        key = capi.light_key(0) if m == 3 else int(m)
        a = capi.Accelerator(gpu_ctx, lpos, lidx, prim_ranges=[[0, lidx.shape[0]]], light_or_mat_keys=[key])
        accels.append(a); instances.append((a, M)); uvs.append(np.ascontiguousarray(c["uvs"][tri.reshape(-1)], np.float32))
    scene = capi.Scene(gpu_ctx, instances)
    r = capi.Renderer(gpu_ctx, scene, 0, 0, c["albedo"], c["radiance"], c["camera"], 64, 64, 32768, seed=81,
                      textures=c["textures"], albedo_texture=c["albedo_texture"], instance_vertex_uvs=uvs)
    img, st = r.render(batch=64)
    assert st.finished
    r.close(); scene.close()
    for a in accels:
        a.close()
    e = rel(bm(img, 2), bm(ref, 2))
    assert e <= 1e-3, e
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.012), (img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)))


def test_mip_mapped_spectral_render_matches_oracle(gpu_ctx):
    """Ray cones under (R)PathTracerSpectral (the refracted cone uses the first wavelength's indices of refraction)."""
    from mray_b200 import spectral
    if not spectral.available():
        pytest.skip("spectral LUT was not generated")
    sp = capi.Spectrum(gpu_ctx, spectral.load(), "HyperbolicPBRT")
    img = mip_scene_renderer(gpu_ctx, "gen_glossy", 32, 32768, seed=91, spectrum=sp)
    sp.close()
    c = scenes.cornell_mips("gen_glossy")
    tm = np.where(c["material"] == 3, -1, c["material"].astype(np.int32))
    textures = [dict(t, gen_mips=c["gen_mips"]) for t in c["textures"]]
    ref = O.oracle_render(c["positions"], c["indices"], tm, c["albedo"], c["radiance"], c["camera"], 32, 32, 8192, seed=92, textures=textures,
                          albedo_texture=c["albedo_texture"], vertex_uvs=c["uvs"], material_type=c["material_type"], material_params=c["material_params"],
                          spectral_data=spectral.load())
    e = rel(bm(img, 2), bm(ref, 2))
    assert e <= 3e-3, e
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.02), (img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)))


def test_degenerate_sizes_and_empty_reads(gpu_ctx):
    one = dict(data=np.full((1, 1, 4), 0.25, np.float32), gen_mips=("Gaussian", 2.0))
    chain, count = capi.texture_mip_chain(gpu_ctx, one)
    assert count == 1 and np.array_equal(chain, O.mip_chain(one)[0])
    strip = dict(data=np.arange(7 * 4, dtype=np.float32).reshape(7, 1, 4) / 28.0, gen_mips=("Box", 0.5))
    chain, count = capi.texture_mip_chain(gpu_ctx, strip)
    ref, ref_count = O.mip_chain(strip)
    assert count == ref_count == 3 and np.array_equal(chain, ref)
    assert capi.texture_final_extent(gpu_ctx, strip) == (1, 7, 3)
    uv = np.array([[0.5, 0.5], [0.1, 0.9], [-3.25, 7.5]], np.float32)
    lod = np.array([0.0, 1.5, 2.0], np.float32)
    assert np.array_equal(capi.texture_sample_lod(gpu_ctx, strip, uv, lod=lod), O.oracle_texture_sample_lod(strip, uv, lod=lod))
    assert capi.texture_sample_lod(gpu_ctx, strip, np.zeros((0, 2), np.float32), lod=np.zeros(0, np.float32)).shape == (0, 3)
