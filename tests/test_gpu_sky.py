"""GPU (B200): SURVEY.md §8f rank 3 — the piecewise-constant 2-D distribution (Tracer/Distributions.cu / .h), the skysphere
coordinate converters and the skysphere boundary light (Tracer/LightsDefault.hpp:L173-443) in the wavefront path tracer.
Pinned by the unmodified reference: DistributionGroupPwC2D outputs (tests/golden/dist2d_*.npz, oracle/gen_golden_dist.py)
and images rendered through TracerI (tests/golden/render_cornell64_sky_*.npz, oracle/gen_golden_render.py)."""
import glob
import os

import numpy as np
import pytest

import oracle_lib as O
from mray_b200 import capi, scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
PLUGIN = os.path.join(ROOT, "mray_b200", "lib", "libTracerDLL_B200.so")
CASES = sorted(os.path.basename(p)[7:-4] for p in glob.glob(os.path.join(GOLDEN, "dist2d_*.npz")))
bm = lambda x, k: x.reshape(x.shape[0] // k, k, x.shape[1] // k, k, 3).mean(axis=(1, 3))
rel = lambda a, b: float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))
SKY_CONSTANT = (1.5, 1.8, 2.5)
SKY_ROTATION = [[0.0, 0.0, 1.0, 0.0], [0.0, 1.0, 0.0, 0.0], [-1.0, 0.0, 0.0, 0.0]]


def ulp_diff(a, b):
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


@pytest.mark.parametrize("case", CASES)
def test_dist2d_matches_reference_vectors(gpu_ctx, case):
    """mrb_dist2d_build / mrb_dist2d_sample against DistributionGroupPwC2D of the unmodified reference: bit for bit on
    these sizes (one scan step per row, so the fp64 additions happen in the reference's order up to associativity of a
    warp scan; every fp32 result here rounds the same way)."""
    g = np.load(os.path.join(GOLDEN, f"dist2d_{case}.npz"))
    cx, cy = capi.dist2d_build(gpu_ctx, g["function"])
    assert ulp_diff(cx, g["cdf_x"]).max() <= 1 and (cx.view(np.uint32) == g["cdf_x"].view(np.uint32)).mean() >= 0.999
    assert np.array_equal(cy.view(np.uint32), g["cdf_y"].view(np.uint32)) or ulp_diff(cy, g["cdf_y"]).max() <= 1
    # sampling on the REFERENCE's tables: binary search, interpolation and pdf are exact
    s = capi.dist2d_sample(gpu_ctx, g["cdf_x"], g["cdf_y"], g["xi"])
    assert np.array_equal(s.view(np.uint32), g["samples"].view(np.uint32))


def test_dist2d_large_against_oracle_and_reference_unit_tests(gpu_ctx):
    """A 2048 x 1024 HDR-like function (the multi-step row scan) against the C oracle, then the reference's own tests
    Dist_PiecewiseConstant2D.Uniform / ZeroVariance (Tests/Tracer/T_Distributions.cu:L110-256) on the GPU."""
    rng = np.random.default_rng(8)
    f = (rng.random((1024, 2048), dtype=np.float32) ** 6 * 300.0).astype(np.float32)
    f[700:704, 300:306] = 2.0e5
    cx, cy = capi.dist2d_build(gpu_ctx, f)
    ox, oy = O.oracle_dist2d_build(f)
    d = ulp_diff(cx, ox)
    assert d.max() <= 1 and (d == 0).mean() >= 0.9999, (d.max(), (d == 0).mean())
    assert ulp_diff(cy, oy).max() <= 1
    xi = rng.random((1 << 16, 2), dtype=np.float32)
    s = capi.dist2d_sample(gpu_ctx, ox, oy, xi)
    assert np.array_equal(s.view(np.uint32), O.oracle_dist2d_sample(ox, oy, xi).view(np.uint32))
    # ZeroVariance: f / pdf is the same for every sample (the mean of f)
    ix = np.minimum((s[:, 0] * 2048).astype(int), 2047); iy = np.minimum((s[:, 1] * 1024).astype(int), 1023)
    est = f[iy, ix] / s[:, 2]
    ok = f[iy, ix] > 1.0           # darker texels: their CDF step is a handful of fp32 ulps (and u * width may round into the neighbour)
    assert np.allclose(est[ok], f.mean(dtype=np.float64), rtol=0.02), (est[ok].min(), est[ok].max(), f.mean())
    # PdfUV(SampleUV(xi)) re-derives the texel from u * width: equal to the sample's own pdf except where that product
    # rounds into the neighbouring texel
    assert np.isclose(s[:, 2], s[:, 3], rtol=1e-5).mean() > 0.995
    # Uniform: pdf 1, uv = xi
    ux, uy = capi.dist2d_build(gpu_ctx, np.full((2160, 3840), 12.0, np.float32))
    xi = np.random.default_rng(332).random((4096, 2), dtype=np.float32)
    u = capi.dist2d_sample(gpu_ctx, ux, uy, xi)
    assert np.allclose(u[:, 2], 1.0, atol=1e-3) and np.allclose(u[:, 3], 1.0, atol=1e-3)
    assert np.allclose(u[:, :2], xi, atol=3e-4)


@pytest.mark.parametrize("mode", ["Skysphere_Spherical", "Skysphere_CoOcta"])
def test_converters_match_reference(gpu_ctx, mode):
    g = np.load(os.path.join(GOLDEN, "dist2d_hdr.npz"))
    ref = g["converters"][capi.BOUNDARY_TYPES[mode] - 1]
    got = capi.skysphere_convert(gpu_ctx, mode, g["dirs"])
    # transcendental functions of two different math libraries: a few ulps
    assert np.allclose(got, ref, rtol=4e-6, atol=4e-7), np.abs(got - ref).max()
    rng = np.random.default_rng(3)
    d = rng.standard_normal((20000, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d[np.abs(d[:, 1]) < 0.999].astype(np.float32)
    out = capi.skysphere_convert(gpu_ctx, mode, d)
    assert np.allclose(out[:, 3:6], d, atol=3e-5)              # DirToUV -> UVToDir round trip
    assert out[:, :2].min() >= 0.0 and out[:, :2].max() <= 1.0
    assert np.allclose(out[:, 2], out[:, 6], rtol=2e-3, atol=1e-6)   # the two ToSolidAnglePdf overloads agree


def test_texture_luminance_is_bit_exact(gpu_ctx):
    t = scenes.sky_texture()
    lum = capi.texture_luminance(gpu_ctx, t)
    assert np.array_equal(lum.view(np.uint32), O.oracle_luminance(t["data"], capi.ACES_CG_LUMINANCE_ROW).view(np.uint32))
    t8 = dict(data=np.random.default_rng(1).integers(0, 256, size=(9, 13, 4), dtype=np.uint8))
    lum8 = capi.texture_luminance(gpu_ctx, t8)
    ref8 = O.oracle_luminance(t8["data"].astype(np.float32) * np.float32(1.0 / 255.0), capi.ACES_CG_LUMINANCE_ROW)
    assert np.array_equal(lum8.view(np.uint32), ref8.view(np.uint32))


def open_cornell_accel(ctx, keep_light):
    c = scenes.cornell_open(keep_light=keep_light)
    order = np.argsort(c["material"], kind="stable")
    idx = np.ascontiguousarray(c["indices"][order]); mat = c["material"][order]
    ranges, keys = [], []
    for m in np.unique(mat):
        w = np.nonzero(mat == m)[0]
        ranges.append([w[0], w[-1] + 1]); keys.append(capi.light_key(0) if m == 3 else int(m))
    return c, idx, mat, capi.Accelerator(ctx, c["positions"], idx, prim_ranges=ranges, light_or_mat_keys=keys)


def test_white_furnace_under_a_constant_sky(gpu_ctx):
    """Closed form: inside a uniform environment of radiance L a white (albedo 1) Lambert scene shows L everywhere, in every
    sample mode and on both maps — emission, NEE sampling, pdf conversion and MIS weights all have to cancel."""
    c, idx, mat, acc = open_cornell_accel(gpu_ctx, False)
    L = 0.75
    for kind in ("Skysphere_Spherical", "Skysphere_CoOcta"):
        for mode in ("Pure", "WithNextEventEstimation", "WithNEEAndMIS"):
            r = capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], np.ones((3, 3), np.float32), np.zeros((0, 3), np.float32),
                              c["camera"], 16, 16, 4096, sample_mode=mode, rr_range=(40, 40), seed=9,
                              boundary=dict(type=kind, radiance=(L, L, L)))
            img, st = r.render(batch=32); r.close()
            # paths are cut at depth 40: what is missing is below 1e-3 in this open box
            assert abs(img.mean() / L - 1.0) < 0.02, (kind, mode, img.mean())
    acc.close()


def sky_boundary(flavour, kind):
    if flavour == "const":
        return None, dict(type=kind, radiance=SKY_CONSTANT)
    b = dict(type=kind, texture=0)
    if flavour == "tex":
        b["transform"] = SKY_ROTATION
    return [scenes.sky_texture()], b


SKY_ITEMS = [("cornell64_sky_const_spp16384", "PathTracerRGB", "WithNEEAndMIS", "Skysphere_Spherical", "const"),
             ("cornell64_sky_tex_spp16384", "PathTracerRGB", "WithNEEAndMIS", "Skysphere_Spherical", "tex"),
             ("cornell64_sky_nee_spp16384", "PathTracerRGB", "WithNextEventEstimation", "Skysphere_Spherical", "tex"),
             ("cornell64_sky_coocta_spectral_spp16384", "PathTracerSpectral", "WithNEEAndMIS", "Skysphere_CoOcta", "tex+light")]


@pytest.mark.parametrize("name,renderer,mode,kind,flavour", SKY_ITEMS)
def test_sky_renders_match_the_reference(gpu_ctx, name, renderer, mode, kind, flavour):
    """The C-ABI renderer and (when prebuilt) the plugin through TracerI against the reference's own render of the same scene."""
    path = os.path.join(GOLDEN, f"render_{name}.npz")
    if not os.path.exists(path):
        pytest.skip("golden image was not generated")
    spectrum = None
    if renderer == "PathTracerSpectral":
        from mray_b200 import spectral
        if not spectral.available():
            pytest.skip("spectral LUT was not generated")
        spectrum = capi.Spectrum(gpu_ctx, spectral.load())
    ref = np.load(path)["img"].astype(np.float32)
    keep_light = flavour == "tex+light"
    c, idx, mat, acc = open_cornell_accel(gpu_ctx, keep_light)
    textures, boundary = sky_boundary(flavour, kind)
    rad = c["radiance"] if keep_light else np.zeros((0, 3), np.float32)
    r = capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], rad, c["camera"], 64, 64, 16384,
                      sample_mode=mode, seed=77, textures=textures, boundary=boundary, spectrum=spectrum)
    img, st = r.render(batch=32); r.close(); acc.close()
    if spectrum is not None:
        spectrum.close()
    # the sun makes these scenes noisier than the diffuse box: 4x4 block means of two 16384-spp images
    # (NEE-only is the noisiest estimator under a sun: two independent 16384-spp renders of the ORACLE differ by 1.2e-3)
    tol = 2.5e-3 if mode == "WithNextEventEstimation" else 1e-3
    e = rel(bm(img, 4), bm(ref, 4))
    assert e <= tol, e
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.01), (img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)))
    if os.path.exists(PLUGIN) and O.driver_available():
        b = O.batched_scene(c["positions"], c["indices"], c["material"])
        kw = dict(boundary=boundary)
        if textures:
            kw["textures"] = textures
        pimg, w, pst = O.driver_render(PLUGIN, b, c["albedo"], 3, c["radiance"], c["camera"], 64, 64, 16384, renderer=renderer, sample_mode=mode,
                                       seed=78, burst_size=64, **kw)
        assert np.allclose(w, 16384, rtol=1e-3)
        pe = rel(bm(pimg, 4), bm(ref, 4))
        assert pe <= tol, pe
        assert np.allclose(pimg.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.01)


def test_bad_boundary_descriptors_are_refused(gpu_ctx):
    c, idx, mat, acc = open_cornell_accel(gpu_ctx, False)
    args = (gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], np.zeros((0, 3), np.float32), c["camera"], 8, 8, 1)
    with pytest.raises(capi.MrbError):
        capi.Renderer(*args, boundary=dict(type="Skysphere_Spherical", texture=0))          # no such texture
    with pytest.raises(capi.MrbError):
        capi.Renderer(*args, boundary=dict(type="Skysphere_CoOcta", radiance=(1, 1, 1), transform=np.zeros((3, 4), np.float32)))
    acc.close()
