"""GPU (B200): textured Lambert albedo (first slice: single-level 2-D textures read through
ParamVaryingData) — the kernel's texture filter bit for bit against the oracle's restatement of the reference's
host-backend view, and textured renders (RGB + spectral, flat + two-level) against the estimator oracle."""
import os

import numpy as np
import pytest

import oracle_lib as O
from mray_b200 import capi, scenes, spectral
from test_gpu_render import cornell_accel, rel_mse, REL_MSE_TOL

pytestmark = pytest.mark.gpu


def test_texture_filter_bit_exact(gpu_ctx):
    rng = np.random.default_rng(2)
    uv = np.concatenate([rng.uniform(-3.0, 4.0, size=(20000, 2)), rng.uniform(0, 1, size=(20000, 2)),
                         [[0, 0], [1, 1], [0.5, 0.5], [-1e-4, 1 - 1e-4], [0.0625, 0.9375]]]).astype(np.float32)
    for tex in (rng.random((5, 7, 3)).astype(np.float32), rng.random((16, 16, 4)).astype(np.float32),
                rng.integers(0, 256, size=(4, 4, 4), dtype=np.uint8), rng.integers(0, 256, size=(9, 3, 3), dtype=np.uint8)):
        for interp in ("Nearest", "Linear"):
            for edge in ("Wrap", "Clamp", "Mirror"):
                t = dict(data=tex, interp=interp, edge=edge)
                got = capi.texture_sample(gpu_ctx, t, uv)
                ref = O.oracle_texture_sample(t, uv[::16])
                assert np.array_equal(got[::16], ref), (tex.shape, tex.dtype, interp, edge)
                assert np.isfinite(got).all()


def _textured(gpu_ctx, res, spp, seed, spectrum=None):
    c, idx, tm, acc = cornell_accel(gpu_ctx)
    uvs, textures, at = scenes.cornell_textures()
    r = capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"],
                      res, res, spp, seed=seed, textures=textures, albedo_texture=at[:3], vertex_uvs=uvs, spectrum=spectrum)
    img, st = r.render(batch=64)
    assert st.finished
    r.close(); acc.close()
    return c, idx, tm, uvs, textures, at, img


def test_textured_cornell_matches_oracle(gpu_ctx):
    res = 32
    c, idx, tm, uvs, textures, at, img = _textured(gpu_ctx, res, 65536, 3)
    ref = O.oracle_render(c["positions"], idx, tm, c["albedo"][:3], c["radiance"], c["camera"], res, res, 16384, sample_mode=2, seed=4,
                          textures=textures, albedo_texture=at[:3], vertex_uvs=uvs)
    err = rel_mse(img, ref)
    assert err <= REL_MSE_TOL, err
    # and the textures matter: the untextured image is far away
    plain = O.oracle_render(c["positions"], idx, tm, c["albedo"][:3], c["radiance"], c["camera"], res, res, 1024, sample_mode=2, seed=5)
    assert rel_mse(img, plain) > 20 * REL_MSE_TOL


@pytest.mark.skipif(not spectral.available(), reason="spectral LUT was not generated")
def test_textured_cornell_spectral_matches_oracle(gpu_ctx):
    res = 32
    sp = capi.Spectrum(gpu_ctx, spectral.load(), "HyperbolicPBRT")
    c, idx, tm, uvs, textures, at, img = _textured(gpu_ctx, res, 131072, 6, spectrum=sp)
    sp.close()
    ref = O.oracle_render(c["positions"], idx, tm, c["albedo"][:3], c["radiance"], c["camera"], res, res, 16384, sample_mode=2, seed=7,
                          textures=textures, albedo_texture=at[:3], vertex_uvs=uvs, spectral_data=spectral.load(), wavelength_mode=2)
    err = rel_mse(img, ref)
    assert err <= REL_MSE_TOL, err


# ------------------------------------------------------------------------------------------------
# colour-space / gamma conversion at upload (TextureMemory::ConvertColorspaces -> KCConvertColor)
# ------------------------------------------------------------------------------------------------
def test_texture_colour_conversion_tap(gpu_ctx):
    """mrb_texture_convert against the numpy restatement of KCConvertColor (Tracer/ColorConverter.cu:L306-398): gamma to linear,
    then ColorspaceTransfer<REC_709, ACES_CG>::RGBToRGBMatrix; fp32 within pow()'s ulps, unorm8 to the quantisation step."""
    import oracle_lib as O
    rng = np.random.default_rng(12)
    m = O.rgb_to_rgb_matrix("REC_709", "ACES_CG")
    f = rng.random((19, 23, 4), dtype=np.float32)
    got = capi.texture_convert(gpu_ctx, dict(data=f, gamma=2.2, color_matrix=m))
    want = O.convert_texture_color(f, 2.2, m)
    assert np.allclose(got[..., :3], want[..., :3], rtol=3e-6, atol=1e-7), np.abs(got - want).max()
    assert np.array_equal(got[..., 3], f[..., 3])                       # the fourth channel is not a colour
    assert np.array_equal(capi.texture_convert(gpu_ctx, dict(data=f)), f)   # nothing to convert: untouched
    only_m = capi.texture_convert(gpu_ctx, dict(data=f, color_matrix=m))
    assert np.allclose(only_m[..., :3], O.convert_texture_color(f, 1.0, m)[..., :3], rtol=2e-6, atol=1e-7)
    u = rng.integers(0, 256, size=(16, 16, 4), dtype=np.uint8)
    got8 = capi.texture_convert(gpu_ctx, dict(data=u, gamma=2.2))
    want8 = O.convert_texture_color(u, 2.2, None)
    assert np.abs(got8.astype(int) - want8.astype(int)).max() <= 1 and (got8 == want8).mean() > 0.99
    # a white texel: the row sums of the matrix
    w = capi.texture_convert(gpu_ctx, dict(data=np.ones((1, 1, 4), np.float32), color_matrix=m))
    assert np.allclose(w[0, 0, :3], m.sum(axis=1), rtol=1e-6)


def test_converted_textures_render_like_the_reference(gpu_ctx):
    """The textured-albedo Cornell box with the fp32 texture declared REC_709 + gamma 2.2 and the unorm8 one gamma 2.2, against the
    reference's render (its ConvertColorspaces ran at load): C-ABI renderer with the conversion fields, and the plugin through
    TracerI with MRayTextureParameters.colorSpace / gamma."""
    import oracle_lib as O
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "render_cornell64_srgbtex_spp16384.npz")
    if not os.path.exists(path):
        pytest.skip("golden image was not generated")
    ref = np.load(path)["img"].astype(np.float32)
    bm = lambda x, k: x.reshape(x.shape[0] // k, k, x.shape[1] // k, k, 3).mean(axis=(1, 3))
    rel = lambda a, b: float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))
    c = scenes.cornell_box()
    uvs, textures, at = scenes.cornell_textures()
    order = np.argsort(c["material"], kind="stable")
    idx = np.ascontiguousarray(c["indices"][order]); mat = c["material"][order]
    ranges, keys = [], []
    for m in np.unique(mat):
        w = np.nonzero(mat == m)[0]
        ranges.append([w[0], w[-1] + 1]); keys.append(capi.light_key(0) if m == 3 else int(m))
    acc = capi.Accelerator(gpu_ctx, c["positions"], idx, prim_ranges=ranges, light_or_mat_keys=keys)
    tex = [dict(textures[0], gamma=2.2, color_matrix=O.rgb_to_rgb_matrix("REC_709", "ACES_CG")), dict(textures[1], gamma=2.2)]
    r = capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"], 64, 64, 16384, seed=71,
                      textures=tex, albedo_texture=at[:3], vertex_uvs=uvs)
    img, st = r.render(batch=32); r.close(); acc.close()
    assert rel(bm(img, 2), bm(ref, 2)) <= 1e-3, rel(bm(img, 2), bm(ref, 2))
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.01)
    plugin = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mray_b200", "lib", "libTracerDLL_B200.so")
    if os.path.exists(plugin) and O.driver_available():
        b = O.batched_scene(c["positions"], c["indices"], c["material"], uvs=uvs)
        ptex = [dict(textures[0], color_space="REC_709", gamma=2.2), dict(textures[1], gamma=2.2)]
        pimg, w, _ = O.driver_render(plugin, b, c["albedo"], 3, c["radiance"], c["camera"], 64, 64, 16384, seed=72, burst_size=64,
                                     textures=ptex, material_texture=at)
        assert np.allclose(w, 16384, rtol=1e-3)
        assert rel(bm(pimg, 2), bm(ref, 2)) <= 1e-3, rel(bm(pimg, 2), bm(ref, 2))
        assert np.allclose(pimg.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.01)
