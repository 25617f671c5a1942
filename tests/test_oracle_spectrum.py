"""CPU: the spectral restatement (oracle/spectrum_oracle.c) against golden vectors produced by the
unmodified reference (oracle/gen_golden_spectrum.py -> tests/golden/spectrum_mode*.npz), plus the
reference's own round-trip test (Tests/Tracer/T_Spectrum.cu: colour -> spectrum x illuminant -> RGB
averages back to the colour within 1e-1)."""
import os
import numpy as np
import pytest

import oracle_lib as O
from mray_b200 import spectral

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_lut = pytest.mark.skipif(not spectral.available(), reason="mray_b200/data/ACES_CG.mrspectra was not generated")


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_wavelength_sampling_matches_reference(mode):
    g = np.load(os.path.join(GOLDEN, f"spectrum_mode{mode}.npz"))
    w, p = O.oracle_sample_wavelengths(mode, g["randoms"])
    # mode 0 is pure arithmetic (exact); 1/2 go through libm transcendentals on both sides, the reference
    # additionally through its own expf/logf polynomials in Gaussian mode
    tol = {0: 0.0, 1: 2e-5, 2: 1e-6}[mode]
    assert np.allclose(w, g["waves"], rtol=tol, atol=0), np.abs(w / g["waves"] - 1).max()
    assert np.allclose(p, g["pdfs"], rtol=max(tol * 20, 1e-6) if mode else 0.0, atol=0)


@needs_lut
@pytest.mark.parametrize("mode", [0, 2])
def test_conversions_match_reference(mode):
    g = np.load(os.path.join(GOLDEN, f"spectrum_mode{mode}.npz"))
    data = spectral.load()
    for c, rgb in enumerate(g["colors"]):
        alb, rad, rgb_a, rgb_r = O.oracle_convert_batch(data, rgb, g["waves"], g["pdfs"], float(g["radiance_scale"]))
        assert np.allclose(alb, g["albedo_spec"][c], rtol=2e-6, atol=1e-7), (c, np.abs(alb - g["albedo_spec"][c]).max())
        assert np.allclose(rad, g["radiance_spec"][c], rtol=2e-6, atol=1e-7), c
        assert np.allclose(rgb_a, g["rgb_albedo_illum"][c], rtol=1e-5, atol=1e-6), c
        assert np.allclose(rgb_r, g["rgb_radiance"][c], rtol=1e-5, atol=1e-5), c


@needs_lut
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_round_trip_like_reference_test(mode):
    """T_Spectrum.cu:L52-215 with 1024 equally spaced random numbers."""
    data = spectral.load()
    n = 1024
    rn = ((np.arange(n, dtype=np.uint64) * ((1 << 24) // n)) << 8).astype(np.uint32)
    w, p = O.oracle_sample_wavelengths(mode, rn)
    rng = np.random.default_rng(0)
    colors = [[0.00368, 0.00304, 0.01033], [0, 0, 0], [0.5, 0.5, 0.5], [1, 1, 1], [0.85, 0.15, 0.15],
              [0.15, 0.85, 0.15], [0.15, 0.15, 0.85]] + rng.uniform(0.15, 0.85, size=(5, 3)).tolist()   # saturated primaries do not round-trip (T_Spectrum.cu:L70-77)
    for rgb in colors:
        _, _, rgb_a, _ = O.oracle_convert_batch(data, rgb, w, p, 1.0)
        assert np.allclose(rgb_a[:, :3].mean(axis=0), rgb, atol=1e-1), (rgb, rgb_a[:, :3].mean(axis=0))


def test_mrspectra_header_errors(tmp_path):
    """ReadMRSpectraFileHeader's checks (SpectrumContext.cu:L298-352)."""
    p = tmp_path / "x.mrspectra"
    p.write_bytes(b"ARTCEPS_RM" + b"\0" * 8)
    with pytest.raises(ValueError, match="character code"):
        spectral.read_mrspectra(str(p))
    p.write_bytes(b"MR_SPECTRA" + np.array([32, 1], np.uint32).tobytes())
    with pytest.raises(ValueError, match="Wrong size"):
        spectral.read_mrspectra(str(p))
    p.write_bytes(b"MR_SPECTRA" + np.array([64, 0], np.uint32).tobytes())
    with pytest.raises(ValueError, match="Wrong mode"):
        spectral.read_mrspectra(str(p))
