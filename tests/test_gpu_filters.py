"""GPU (B200): the film filters of TracerParameters.filmFilter (Tracer/Filters.h) — the kernel's sampler against the
oracle's restatement and the reference's own filter tests (Tests/Tracer/T_Filters.cu), and renders with every filter
against images rendered by the unmodified reference (tests/golden/render_cornell64_{box,tent,mitchell}_spp16384.npz)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O
from mray_b200 import capi, scenes
from test_gpu_render import cornell_accel
from test_oracle_reference_unit_tests import reference_filter_test

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILTERS = ["Box", "Tent", "Gaussian", "Mitchell-Netravali"]


@pytest.mark.parametrize("name", FILTERS)
def test_filter_sampler_matches_oracle_and_reference_properties(gpu_ctx, name):
    L = O.lib()
    L.orc_pt_filter_sample_typed.argtypes = [C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_void_p]
    rng = np.random.default_rng(7)
    ftype = FILTERS.index(name)
    out = np.zeros(5, np.float32)
    for radius in (0.05, 1.0, 2.5, 13.0):
        xi = rng.random((4096, 2)).astype(np.float32)
        xi[0] = 0.0; xi[1] = np.nextafter(np.float32(1), np.float32(0))
        got = capi.filter_sample(gpu_ctx, name, radius, xi)
        assert np.isfinite(got).all()
        ref = np.zeros((xi.shape[0], 4), np.float32)
        for k in range(xi.shape[0]):
            L.orc_pt_filter_sample_typed(ftype, radius, float(xi[k, 0]), float(xi[k, 1]), out.ctypes.data)
            ref[k] = out[[0, 1, 2, 4]]
        # offsets: the device's erfinvf vs a Newton iteration on erf, single-precision tails
        assert np.allclose(got[:, :2], ref[:, :2], rtol=2e-4, atol=2e-5 * radius), np.abs(got[:, :2] - ref[:, :2]).max()
        w_got, w_ref = got[:, 3] / got[:, 2], ref[:, 3] / ref[:, 2]
        assert np.allclose(w_got, w_ref, rtol=5e-3, atol=5e-3), np.abs(w_got - w_ref).max()
    # the reference's own test on the GPU sampler (Pdf(offset) is the oracle's: the kernel only needs Sample and Evaluate)
    def sample(r, x0, x1):
        g = capi.filter_sample(gpu_ctx, name, r, np.array([[x0, x1]], np.float32))[0]
        L.orc_pt_filter_sample_typed(ftype, r, x0, x1, out.ctypes.data)
        return np.array([g[0], g[1], g[2], out[3], g[3]], np.float32)
    reference_filter_test(sample, name != "Mitchell-Netravali")


def test_unknown_filter_is_an_error(gpu_ctx):
    c, idx, tm, acc = cornell_accel(gpu_ctx)
    with pytest.raises(capi.MrbError):
        capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"], 8, 8, 1, film_filter=7)
    with pytest.raises(capi.MrbError):
        capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"], 8, 8, 1,
                      film_filter="Box", film_filter_radius=0.0)
    with pytest.raises(capi.MrbError):
        capi.filter_sample(gpu_ctx, 9, 1.0, np.zeros((1, 2), np.float32))
    acc.close()


@pytest.mark.parametrize("name,golden", [("Box", "box"), ("Tent", "tent"), ("Mitchell-Netravali", "mitchell")])
def test_filter_render_against_reference(gpu_ctx, name, golden):
    """The film holds plain radiance sums and FILTER-WEIGHT sums (KCSetImagePixelsIndirect, Tracer/TextureFilter.cu:L642-668:
    val += radiance, weight += Evaluate / pdf), so with Mitchell-Netravali the weight plane itself is part of the contract."""
    path = os.path.join(ROOT, "tests", "golden", f"render_cornell64_{golden}_spp16384.npz")
    if not os.path.exists(path):
        pytest.skip("golden image was not generated")
    g = np.load(path)
    ref, ref_w = g["img"].astype(np.float32), g["weight"]
    radius, spp = float(g["film_filter_radius"]), 16384
    c, idx, tm, acc = cornell_accel(gpu_ctx)
    r = capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"], 64, 64, spp,
                      seed=31, film_filter=name, film_filter_radius=radius)
    st = r.run_pass(16)
    assert st.finished
    rgb, w = r.read_film(); r.close(); acc.close()
    img = rgb / w[..., None]
    bm = lambda x, k: x.reshape(x.shape[0] // k, k, x.shape[1] // k, k, -1).mean(axis=(1, 3))
    # mean film weight per sample: 1 for the filters sampled from their own shape, E[f / pdf] = 1 for Mitchell too
    assert abs(w.mean() / spp - 1.0) < 2e-3 and abs(float(ref_w.mean()) / spp - 1.0) < 2e-3, (w.mean() / spp, ref_w.mean() / spp)
    if name != "Mitchell-Netravali":
        assert np.allclose(w, spp, rtol=1e-3)
    else:   # per-pixel weight sums scatter like the reference's: std of f / pdf is 0.70 per axis-pair sample
        assert abs(w.std() / ref_w.std() - 1.0) < 0.1, (w.std(), ref_w.std())
    err = float(np.mean((bm(img, 2) - bm(ref, 2)) ** 2 / (bm(ref, 2) ** 2 + 1e-2)))
    assert err <= (1e-3 if name != "Mitchell-Netravali" else 2e-3), err
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.01), (img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)))
