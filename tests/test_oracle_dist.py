"""CPU: the piecewise-constant 2-D distribution / skysphere converter restatement (oracle/dist_oracle.c) against golden
vectors produced by the unmodified reference (oracle/gen_golden_dist.py -> tests/golden/dist2d_*.npz), plus the
reference's own tests Dist_PiecewiseConstant2D.Uniform / ZeroVariance (Tests/Tracer/T_Distributions.cu:L110-256)."""
import glob
import os
import numpy as np
import pytest

import oracle_lib as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[7:-4] for p in glob.glob(os.path.join(GOLDEN, "dist2d_*.npz")))


@pytest.mark.parametrize("case", CASES)
def test_build_and_sample_match_reference_bit_for_bit(case):
    g = np.load(os.path.join(GOLDEN, f"dist2d_{case}.npz"))
    cx, cy = O.oracle_dist2d_build(g["function"])
    assert np.array_equal(cx.view(np.uint32), g["cdf_x"].view(np.uint32))
    assert np.array_equal(cy.view(np.uint32), g["cdf_y"].view(np.uint32))
    s = O.oracle_dist2d_sample(cx, cy, g["xi"])
    assert np.array_equal(s.view(np.uint32), g["samples"].view(np.uint32))


@pytest.mark.parametrize("mode", [1, 2])
def test_converters_match_reference(mode):
    g = np.load(os.path.join(GOLDEN, "dist2d_hdr.npz"))
    ours = O.oracle_sky_converters(mode, g["dirs"])
    ref = g["converters"][mode - 1]
    # libm on both sides: atan2f / acosf / sinf / cosf agree to an ulp or two
    assert np.allclose(ours, ref, rtol=2e-6, atol=2e-7), np.abs(ours - ref).max()
    # the mapping round-trips (DirToUV -> UVToDir gives the direction back)
    # (not at the poles: the spherical map loses the azimuth there, and the reference's co-octahedral DirToUV answers
    # uv = 0 for the exact +-Y axis — GraphicsFunctions.h:L325 — which UVToDir maps to -Y)
    ok = np.abs(g["dirs"][:, 1]) < 0.999
    assert np.allclose(ours[ok, 3:6], g["dirs"][ok], atol=2e-5)


def test_reference_unit_test_uniform():
    """Dist_PiecewiseConstant2D.Uniform: a constant function gives pdf 1 everywhere and uv == xi."""
    rng = np.random.default_rng(332)
    xi = rng.random((4096, 2), dtype=np.float32)
    cx, cy = O.oracle_dist2d_build(np.full((216, 384), 12.0, np.float32))
    s = O.oracle_dist2d_sample(cx, cy, xi)
    assert np.allclose(s[:, 2], 1.0, atol=1e-4) and np.allclose(s[:, 3], 1.0, atol=1e-4)
    assert np.allclose(s[:, :2], xi, atol=2e-4)


def test_reference_unit_test_zero_variance():
    """Dist_PiecewiseConstant2D.ZeroVariance: f(sample) / pdf(sample) is the same for every sample (= the mean of f)."""
    rng = np.random.default_rng(123)
    f = (rng.random((64, 96), dtype=np.float32) * 9.0 + 1.0).astype(np.float32)
    cx, cy = O.oracle_dist2d_build(f)
    xi = rng.random((8192, 2), dtype=np.float32)
    s = O.oracle_dist2d_sample(cx, cy, xi)
    ix = np.minimum((s[:, 0] * 96).astype(int), 95); iy = np.minimum((s[:, 1] * 64).astype(int), 63)
    est = f[iy, ix] / s[:, 2]
    assert np.allclose(est, f.mean(), rtol=2e-3), (est.min(), est.max(), f.mean())
    assert np.allclose(s[:, 2], s[:, 3], rtol=1e-5)


def test_luminance_is_the_y_row_dot():
    rng = np.random.default_rng(1)
    rgb = rng.random((5, 7, 4), dtype=np.float32)
    y = np.array([0.271564007, 0.673637331, 0.0577730648], np.float32)
    lum = O.oracle_luminance(rgb, y)
    assert np.allclose(lum, rgb[..., :3] @ y, rtol=1e-6)
