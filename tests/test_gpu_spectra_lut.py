"""GPU (B200): mrb_spectra_lut_generate — the reference's SpectraLUTGen (Source/SpectraLUTGen/main.cpp) on the device — against
the oracle's restatement column by column, and the WHOLE 64^3 x 9 table against the file the unmodified reference tool wrote
(mray_b200/data/ACES_CG.mrspectra), through the spectra the coefficients encode and bit for bit."""
import os
import subprocess
import numpy as np
import pytest

import oracle_lib as O
from mray_b200 import capi, spectral
from test_oracle_spectra_lut import INPUTS, spectra

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "mray_b200", "lib", "mray_b200_spectra_lut_gen")


def inputs():
    a = np.fromfile(INPUTS, np.float32)
    n = 471
    return dict(cie_xyz=a[:3 * n].reshape(n, 3), illuminant_spd=a[3 * n:4 * n], illuminant_norm=float(a[4 * n]),
                rgb_to_xyz=a[4 * n + 1:4 * n + 10], xyz_to_rgb=a[4 * n + 10:4 * n + 19]), a


def test_small_table_matches_the_oracle(gpu_ctx):
    kw, raw = inputs()
    lut, wp = capi.spectra_lut_generate(gpu_ctx, resolution=16, **kw)
    lut = lut.reshape(3, 3, 16, 16, 16)
    rng = np.random.default_rng(3)
    exact = []
    for _ in range(40):
        l, j, i = int(rng.integers(3)), int(rng.integers(16)), int(rng.integers(16))
        col = O.oracle_spectra_lut_column(raw, l, j, i, res=16)
        got = lut[l, :, :, j, i].T
        assert np.abs(spectra(got) - spectra(col)).max() < 1e-3, (l, j, i)
        exact.append((got.view(np.uint32) == col.view(np.uint32)).mean())
    assert np.mean(exact) > 0.95, np.mean(exact)
    assert np.allclose(wp, [0.9526, 1.0, 1.0088], atol=2e-3)      # ACES white (D60-like) with Y normalised to 1
    with pytest.raises(capi.MrbError):
        capi.spectra_lut_generate(gpu_ctx, resolution=2, **kw)


@pytest.mark.skipif(not os.path.exists(spectral.lut_path()), reason="mray_b200/data/ACES_CG.mrspectra was not generated")
def test_full_table_matches_the_reference_tools_file(gpu_ctx, tmp_path):
    kw, _ = inputs()
    lut, wp = capi.spectra_lut_generate(gpu_ctx, resolution=64, **kw)
    ref = spectral.read_mrspectra(spectral.lut_path())
    assert np.isfinite(lut).all()
    # compared through the spectra the coefficients encode (L1 over 360..830 nm): saturated cells hold step-like spectra whose edge
    # moves by a fraction of a nanometre with the last bits, so a point-wise comparison would be meaningless there
    lam = np.arange(360.0, 831.0, 2.0)
    unstable = []
    for l in range(3):
        a, b = np.moveaxis(lut.reshape(3, 3, 64, 64, 64)[l], 0, -1), np.moveaxis(ref.reshape(3, 3, 64, 64, 64)[l], 0, -1)
        m = np.abs(spectra(a, lam) - spectra(b, lam)).mean(axis=-1)                   # [z, y, x]
        unstable += [(l, int(z), int(y), int(x)) for z, y, x in np.argwhere(m > 1e-4)]
    # In ~90 cells of the blue-dominant table with (almost) no red — pure, fully saturated blues — Gauss-Newton does not settle: it
    # alternates between two polynomials from one brightness cell to the next, and WHICH of the two a cell gets flips with the
    # last bits of the iteration (in the reference itself the white point is summed by racing atomics; the C oracle flips elsewhere
    # again). Everywhere else the two tables encode the same spectra to 1e-4.
    assert len(unstable) < 400, len(unstable)
    assert all(l == 2 and x <= 4 for l, z, y, x in unstable), unstable[:10]
    assert (lut.view(np.uint32) == ref.view(np.uint32)).mean() > 0.95
    # the command-line tool writes the same bytes the reference's loader expects
    if os.path.exists(TOOL):
        r = subprocess.run([TOOL, "64", "ACES_CG", str(tmp_path)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        mine = spectral.read_mrspectra(str(tmp_path / "ACES_CG.mrspectra"))
        assert np.array_equal(mine.view(np.uint32), lut.view(np.uint32))
