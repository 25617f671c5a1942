"""GPU (B200): hero-wavelength spectral transport through the C-ABI against the
reference's own outputs (tests/golden/spectrum_mode*.npz, produced by running the unmodified reference),
the C restatement, the reference's round-trip property test, and a spectral render against the spectral
estimator oracle and its RGB twin."""
import os
import numpy as np
import pytest
import torch

import oracle_lib as O
from mray_b200 import capi, scenes, spectral
from test_gpu_render import cornell_accel, rel_mse, REL_MSE_TOL
from test_oracle_pt import rect_form_factor

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not spectral.available(), reason="spectral LUT was not generated")]
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODES = {0: "Uniform", 1: "GaussianMIS", 2: "HyperbolicPBRT"}


@pytest.fixture(scope="module")
def spec_data():
    return spectral.load()


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_kernels_match_reference_goldens(gpu_ctx, spec_data, mode):
    g = np.load(os.path.join(GOLDEN, f"spectrum_mode{mode}.npz"))
    sp = capi.Spectrum(gpu_ctx, spec_data, MODES[mode])
    n = g["randoms"].shape[0]
    waves = np.zeros((n, 4), np.float32); pdfs = np.zeros((n, 4), np.float32)
    sp.sample_wavelengths(waves, pdfs, np.ascontiguousarray(g["randoms"]))
    # Uniform is plain arithmetic; the other modes go through device atanhf/coshf/erfinvf vs the reference's
    # host libm / its erfinv polynomial
    wtol, ptol = {0: (0.0, 0.0), 1: (2e-5, 4e-4), 2: (1e-6, 1e-5)}[mode]
    assert np.allclose(waves, g["waves"], rtol=wtol, atol=0), np.abs(waves / g["waves"] - 1).max()
    assert np.allclose(pdfs, g["pdfs"], rtol=ptol, atol=1e-30), np.nanmax(np.abs(pdfs / g["pdfs"] - 1))
    # conversions at the REFERENCE's wavelengths so that only the converter is compared
    gw, gp = np.ascontiguousarray(g["waves"]), np.ascontiguousarray(g["pdfs"])
    worst = 0.0
    for c, rgb in enumerate(g["colors"]):
        alb = np.zeros((n, 4), np.float32); rad = np.zeros((n, 4), np.float32)
        sp.upsample(alb, np.ascontiguousarray(rgb, np.float32), gw, is_radiance=False)
        sp.upsample(rad, np.ascontiguousarray(rgb * g["radiance_scale"], np.float32), gw, is_radiance=True)
        # device sinf/asinf (inverse smoothstep) and rsqrtf differ from host libm by ulps; the polynomial
        # c0 l^2 + c1 l + c2 amplifies coefficient rounding by l^2 ~ 5e5, hence 2e-5 absolute on [0,1] values
        worst = max(worst, float(np.abs(alb - g["albedo_spec"][c]).max()))
        assert np.allclose(alb, g["albedo_spec"][c], rtol=1e-5, atol=2e-5), (c, np.abs(alb - g["albedo_spec"][c]).max())
        assert np.allclose(rad, g["radiance_spec"][c], rtol=3e-5, atol=1e-4), (c, np.abs(rad - g["radiance_spec"][c]).max())
        ok = np.isfinite(g["rgb_radiance"][c]).all(axis=1)          # Gaussian tails: pdf 0 -> the reference divides to 0 too
        rgb_r = rad.copy()
        sp.convert_to_rgb(rgb_r, gw, gp)
        scale = np.abs(g["rgb_radiance"][c][ok]).max() + 1e-6
        assert np.allclose(rgb_r[ok], g["rgb_radiance"][c][ok], rtol=1e-4, atol=1e-5 * scale), c
        # per-element rgb array variant == uniform variant
        alb2 = np.zeros((n, 4), np.float32)
        sp.upsample(alb2, np.ascontiguousarray(np.tile(rgb, (n, 1)), np.float32), gw, is_radiance=False)
        assert np.array_equal(alb, alb2)
    print("worst albedo-spectrum deviation from the reference:", worst)
    sp.close()


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_round_trip_like_reference_test(gpu_ctx, spec_data, mode):
    """Tests/Tracer/T_Spectrum.cu: colour -> spectrum x illuminant -> RGB over 1024 equally spaced random
    numbers averages back to the colour within 1e-1 — on device arrays."""
    sp = capi.Spectrum(gpu_ctx, spec_data, MODES[mode])
    n = 1024
    rn = torch.from_numpy(((np.arange(n, dtype=np.uint64) * ((1 << 24) // n)) << 8).astype(np.uint32).view(np.int32)).cuda()
    waves = torch.zeros((n, 4), device="cuda"); pdfs = torch.zeros((n, 4), device="cuda")
    sp.sample_wavelengths(waves, pdfs, rn)
    ow, op = O.oracle_sample_wavelengths(mode, rn.cpu().numpy().view(np.uint32))
    assert np.allclose(waves.cpu().numpy(), ow, rtol=2e-5)
    rng = np.random.default_rng(0)
    colors = [[0.00368, 0.00304, 0.01033], [0, 0, 0], [0.5, 0.5, 0.5], [1, 1, 1], [0.85, 0.15, 0.15],
              [0.15, 0.85, 0.15], [0.15, 0.15, 0.85]] + rng.uniform(0.15, 0.85, size=(5, 3)).tolist()
    white = torch.ones(3, device="cuda") * 0.5       # ConvertRadiance(0.5) = sigmoid(...)(~1) * illuminant * 1
    for rgb in colors:
        c = torch.tensor(rgb, dtype=torch.float32, device="cuda")
        alb = torch.zeros((n, 4), device="cuda"); ill = torch.zeros((n, 4), device="cuda")
        sp.upsample(alb, c, waves, is_radiance=False)
        sp.upsample(ill, white, waves, is_radiance=True)
        one = torch.zeros((n, 4), device="cuda"); sp.upsample(one, white, waves, is_radiance=False)
        val = alb * (ill / one)                      # albedo spectrum x illuminant SPD
        sp.convert_to_rgb(val, waves, pdfs)
        got = val[:, :3].mean(dim=0).cpu().numpy()
        assert np.allclose(got, rgb, atol=1e-1), (rgb, got)
    sp.close()


def test_spectral_cornell_matches_spectral_oracle(gpu_ctx, spec_data):
    c, idx, tm, acc = cornell_accel(gpu_ctx)
    sp = capi.Spectrum(gpu_ctx, spec_data, "HyperbolicPBRT")
    res, spp = 32, 131072
    r = capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"],
                      res, res, spp, seed=21, spectrum=sp)
    img, st = r.render(batch=64)
    assert st.finished
    r.close()
    ref = O.oracle_render(c["positions"], idx, tm, c["albedo"][:3], c["radiance"], c["camera"], res, res, 16384,
                          sample_mode=2, seed=5, spectral_data=spec_data, wavelength_mode=2)
    # two independent spectral estimates of 131072 and 16384 spp (colour noise adds to the RGB figure of
    # test_gpu_render.py): converged images agree to the north-star tolerance
    err = rel_mse(img, ref)
    assert err <= REL_MSE_TOL, err
    # and the spectral image stays close to its RGB twin (upsampling round trip, T_Spectrum's 1e-1 per colour)
    rgb_r = capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"],
                          res, res, 16384, seed=22)
    img_rgb, _ = rgb_r.render(batch=64); rgb_r.close()
    mask = img_rgb.max(axis=-1) < 5.0
    assert np.allclose(img[mask].mean(axis=0), img_rgb[mask].mean(axis=0), rtol=0.1), (img[mask].mean(axis=0), img_rgb[mask].mean(axis=0))
    sp.close(); acc.close()


def test_spectral_direct_lighting_closed_form(gpu_ctx, spec_data):
    """Grey floor under a white square light: E = rho L F for every wavelength, so the spectral estimator
    must return (rho L F) x RGB(illuminant-white) = the closed form within the upsampling round trip."""
    half, h, L, rho = 0.5, 1.5, 10.0, 0.6
    floor = np.array([[-50, 0, 50], [50, 0, 50], [50, 0, -50], [-50, 0, -50]], np.float32)
    light = np.array([[-half, h, -half], [half, h, -half], [half, h, half], [-half, h, half]], np.float32)
    pos = np.ascontiguousarray(np.concatenate([floor, light]))
    idx = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7]], np.uint32)
    acc = capi.Accelerator(gpu_ctx, pos, idx, prim_ranges=[[0, 2], [2, 4]], light_or_mat_keys=[0, capi.light_key(0)])
    cam = dict(eye=(0.0, 1.0, 0.0), gaze=(0.0, 0.0, 0.0), up=(0.0, 0.0, -1.0), fov_y_deg=2.0)
    expect = rho * L * 4 * rect_form_factor(half, half, h)
    sp = capi.Spectrum(gpu_ctx, spec_data, "HyperbolicPBRT")
    r = capi.Renderer(gpu_ctx, acc, 8, 4, [[rho, rho, rho]], [L, L, L], cam, 16, 16, 16384, sample_mode="WithNextEventEstimation",
                      rr_range=(2, 2), spectrum=sp)
    img, st = r.render()
    got = img[4:12, 4:12].mean(axis=(0, 1))
    assert np.allclose(got, expect, rtol=0.03), (got, expect)
    r.close(); sp.close(); acc.close()
