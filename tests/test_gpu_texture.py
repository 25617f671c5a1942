"""GPU (B200): textured Lambert albedo (first slice: single-level 2-D textures read through
ParamVaryingData) — the kernel's texture filter bit for bit against the oracle's restatement of the reference's
host-backend view, and textured renders (RGB + spectral, flat + two-level) against the estimator oracle."""
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
