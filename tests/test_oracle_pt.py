"""CPU: the path-tracing estimator oracle (oracle/pt_oracle.c) against IMAGES RENDERED BY THE UNMODIFIED
REFERENCE (tests/golden/render_*.npz, produced by oracle/gen_golden_render.py: the reference's CPU backend
driven through TracerI), plus self-consistency and closed-form checks."""
import os

import numpy as np
import pytest

import oracle_lib as O
from mray_b200 import scenes


def cornell():
    c = scenes.cornell_box()
    tm = np.where(c["material"] == 3, -1, c["material"]).astype(np.int32)
    return c, tm


def test_pure_nee_and_mis_agree_in_expectation():
    c, tm = cornell()
    imgs = [O.oracle_render(c["positions"], c["indices"], tm, c["albedo"][:3], c["radiance"], c["camera"], 32, 32, 512,
                            sample_mode=m, seed=m) for m in (0, 1, 2)]
    # exclude pixels that see the light directly (huge values dominate the mean)
    mask = imgs[2].max(axis=-1) < 5.0
    means = [im[mask].mean(axis=0) for im in imgs]
    for m in means[1:]:
        assert np.allclose(m, means[0], rtol=0.04), (means[0], m)


def rect_form_factor(x, y, h):
    """dA -> parallel rectangle with one corner above dA, sides x,y, distance h."""
    a, b = np.sqrt(x * x + h * h), np.sqrt(y * y + h * h)
    return (x / a * np.arctan(y / a) + y / b * np.arctan(x / b)) / (2 * np.pi)


def test_direct_lighting_matches_closed_form():
    """Floor + square one-sided light facing down. With sampleMode NEE and rrRange (2,2) a pixel shows
    exactly the direct term: L_o = albedo * L * F(dA -> light)."""
    half, h, L, rho = 0.5, 1.5, 10.0, 0.6
    floor = np.array([[-50, 0, 50], [50, 0, 50], [50, 0, -50], [-50, 0, -50]], np.float32)
    light = np.array([[-half, h, -half], [half, h, -half], [half, h, half], [-half, h, half]], np.float32)
    pos = np.concatenate([floor, light])
    idx = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7]], np.uint32)
    tm = np.array([0, 0, -1, -1], np.int32)
    cam = dict(eye=(0.0, 1.0, 0.0), gaze=(0.0, 0.0, 0.0), up=(0.0, 0.0, -1.0), fov_y_deg=2.0)
    img = O.oracle_render(pos, idx, tm, [[rho, rho, rho]], [L, L, L], cam, 8, 8, 4096, sample_mode=1, rr_range=(2, 2))
    expect = rho * L * 4 * rect_form_factor(half, half, h)
    got = img[2:6, 2:6].mean()
    assert abs(got - expect) / expect < 0.02, (got, expect)
    # and with MIS (bxdf + light strategies combined) the same direct term must come out
    img2 = O.oracle_render(pos, idx, tm, [[rho, rho, rho]], [L, L, L], cam, 8, 8, 4096, sample_mode=2, rr_range=(2, 2))
    got2 = img2[2:6, 2:6].mean()
    assert abs(got2 - expect) / expect < 0.03, (got2, expect)


# ------------------------------------------------------------------------------------------------
# pinned by reference execution
# ------------------------------------------------------------------------------------------------
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def ref_image(name):
    return np.load(os.path.join(GOLDEN, f"render_{name}.npz"))["img"].astype(np.float32)


def rel_mse(a, b):
    return float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))


def block_mean(img, k=2):
    h, w, c = img.shape
    return img.reshape(h // k, k, w // k, k, c).mean(axis=(1, 3))


def test_oracle_converged_image_matches_reference_render():
    """WithNEEAndMIS, 64x64, compared as 2x2 block means: oracle at 4096 spp (16384 samples per block) vs the
    reference's 16384-spp image (65536 per block). Two independent estimates of N and M samples differ by relMSE
    ~ 7.7 (1/N + 1/M) ~ 5.9e-4 here; the north-star tolerance is 1e-3."""
    c, tm = cornell()
    ref = ref_image("cornell64_spp16384")
    img = O.oracle_render(c["positions"], c["indices"], tm, c["albedo"][:3], c["radiance"], c["camera"], 64, 64, 4096,
                          sample_mode=2, seed=41)
    err = rel_mse(block_mean(img), block_mean(ref))
    assert err <= 1e-3, err
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=5e-3), (img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)))


@pytest.mark.parametrize("mode,name", [(1, "cornell64_nee_spp16384"), (0, "cornell64_pure_spp16384")])
def test_oracle_other_sample_modes_match_reference_render(mode, name):
    c, tm = cornell()
    ref = ref_image(name)
    img = O.oracle_render(c["positions"], c["indices"], tm, c["albedo"][:3], c["radiance"], c["camera"], 64, 64, 1024,
                          sample_mode=mode, seed=42 + mode)
    mask = ref.max(axis=-1) < 5.0
    assert np.allclose(img[mask].mean(axis=0), ref[mask].mean(axis=0), rtol=0.02), (img[mask].mean(axis=0), ref[mask].mean(axis=0))


def test_reference_sample_modes_and_two_level_render_agree():
    """Sanity of the fixtures themselves: the reference's three sample modes and its two-level ((T)Single per batch)
    render share one expectation."""
    mis = ref_image("cornell64_spp16384")
    mask = mis.max(axis=-1) < 5.0
    for name, tol in (("cornell64_nee_spp16384", 0.01), ("cornell64_pure_spp16384", 0.02), ("cornell64_single_spp16384", 0.005)):
        im = ref_image(name)
        assert np.allclose(im[mask].mean(axis=0), mis[mask].mean(axis=0), rtol=tol), (name, im[mask].mean(axis=0), mis[mask].mean(axis=0))
    assert rel_mse(block_mean(ref_image("cornell64_single_spp16384")), block_mean(mis)) <= 1e-3


def test_spectral_oracle_matches_reference_spectral_render():
    from mray_b200 import spectral
    if not spectral.available():
        pytest.skip("spectral LUT was not generated")
    c, tm = cornell()
    ref = ref_image("cornell64_spectral_spp16384")
    img = O.oracle_render(c["positions"], c["indices"], tm, c["albedo"][:3], c["radiance"], c["camera"], 64, 64, 2048,
                          sample_mode=2, seed=43, spectral_data=spectral.load(), wavelength_mode=2)
    # 2048 spp (8192 per 2x2 block) keeps the CPU suite short: noise ~ 1.2e-3; the converged comparison runs on the GPU
    err = rel_mse(block_mean(img), block_mean(ref))
    assert err <= 2.5e-3, err
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.01), (img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)))


def test_textured_oracle_matches_reference_render():
    """Textured Lambert albedo (single-level 2-D textures, bilinear / wrap and nearest / clamp): the oracle's
    restatement of the reference's host-backend texture view against the reference's own textured render."""
    from mray_b200 import scenes
    c, tm = cornell()
    uvs, textures, at = scenes.cornell_textures()
    ref = ref_image("cornell64_textured_spp16384")
    img = O.oracle_render(c["positions"], c["indices"], tm, c["albedo"][:3], c["radiance"], c["camera"], 64, 64, 4096,
                          sample_mode=2, seed=44, textures=textures, albedo_texture=at[:3], vertex_uvs=uvs)
    err = rel_mse(block_mean(img), block_mean(ref))
    assert err <= 1e-3, err
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=5e-3), (img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)))
    # the textures matter: the untextured reference image is far away
    assert rel_mse(block_mean(ref), block_mean(ref_image("cornell64_spp16384"))) > 2e-2


def test_mirror_shows_the_light_exactly():
    """(Mt)Reflect closed form: a camera looking at a mirror floor that reflects a one-sided light sees exactly the
    light's radiance (reflectance 1, pdf 1, SPECULAR_RAY hits count in full in every sample mode)."""
    L = 7.0
    floor = np.array([[-5, 0, 5], [5, 0, 5], [5, 0, -5], [-5, 0, -5]], np.float32)
    light = np.array([[-30, 4, -30], [30, 4, -30], [30, 4, 30], [-30, 4, 30]], np.float32)   # faces down
    pos = np.concatenate([floor, light]); idx = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7]], np.uint32)
    tm = np.array([0, 0, -1, -1], np.int32)
    cam = dict(eye=(0.0, 1.0, 2.0), gaze=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fov_y_deg=10.0)
    for mode in (0, 1, 2):
        img = O.oracle_render(pos, idx, tm, [[0.5, 0.5, 0.5]], [L, L, L], cam, 8, 8, 16, sample_mode=mode, material_type=[1])
        assert np.allclose(img, L, rtol=1e-5), (mode, img.min(), img.max())


def test_mirror_box_changes_the_image_but_not_the_energy_scale():
    from mray_b200 import scenes
    c = scenes.cornell_mirror()
    tm = np.where(c["material"] == 3, -1, np.where(c["material"] == 4, 3, c["material"])).astype(np.int32)
    alb = c["albedo"][[0, 1, 2, 4]]
    img = O.oracle_render(c["positions"], c["indices"], tm, alb, c["radiance"], c["camera"], 32, 32, 512, sample_mode=2, seed=3,
                          material_type=[0, 0, 0, 1])
    plain = ref_image("cornell64_spp16384")
    plain32 = block_mean(plain)
    assert np.isfinite(img).all()
    assert rel_mse(img, plain32) > 1e-2                       # the mirror is visible
    assert 0.7 < img.mean() / plain32.mean() < 1.4


def test_mirror_oracle_matches_reference_render():
    """(Mt)Reflect against the reference's render of the mirror-box Cornell. The scene is ~5x noisier than the diffuse
    one (reference-vs-reference relMSE ~ 43 (1/N + 1/M)), so 8x8 block means carry the converged comparison."""
    from mray_b200 import scenes
    c = scenes.cornell_mirror()
    tm = np.where(c["material"] == 3, -1, np.where(c["material"] == 4, 3, c["material"])).astype(np.int32)
    ref = ref_image("cornell64_mirror_spp16384")
    img = O.oracle_render(c["positions"], c["indices"], tm, c["albedo"][[0, 1, 2, 4]], c["radiance"], c["camera"], 64, 64, 1024,
                          sample_mode=2, seed=45, material_type=[0, 0, 0, 1])
    err = rel_mse(block_mean(img, 8), block_mean(ref, 8))
    assert err <= 1e-3, err
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.01), (img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)))
    assert rel_mse(block_mean(ref, 8), block_mean(ref_image("cornell64_spp16384"), 8)) > 5e-3     # the mirror is visible


def test_two_sided_light_oracle_matches_reference_render():
    """(L)Prim(P)Triangle's isTwoSided attribute against the reference's render with it switched on."""
    c, tm = cornell()
    ref = ref_image("cornell64_twosided_spp16384")
    img = O.oracle_render(c["positions"], c["indices"], tm, c["albedo"][:3], c["radiance"], c["camera"], 64, 64, 2048,
                          sample_mode=2, seed=46, light_two_sided=[1])
    # the ceiling 2 cm above the light makes this scene several times noisier than the one-sided one (like the mirror
    # scene): the converged comparison runs on 8x8 block means
    err = rel_mse(block_mean(img, 8), block_mean(ref, 8))
    assert err <= 1e-3, err
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=5e-3), (img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)))
    # the attribute matters: brighter than the one-sided image
    assert ref.mean() > 1.02 * ref_image("cornell64_spp16384").mean()


def glossy():
    """scenes.cornell_glossy() in oracle shape: flat material table without the light row."""
    from mray_b200 import scenes
    c = scenes.cornell_glossy()
    order = {0: 0, 1: 1, 2: 2, 4: 3, 5: 4}
    tm = np.array([-1 if m == 3 else order[int(m)] for m in c["material"]], np.int32)
    rows = [0, 1, 2, 4, 5]
    return c, tm, c["albedo"][rows], c["material_type"][rows], c["material_params"][rows]


def test_refract_closed_forms():
    """(Mt)Refract: equal indices of refraction on both sides make the interface invisible (Fresnel 0, straight-through
    refraction): a camera looking at a one-sided light through two such panes sees exactly its radiance."""
    L = 3.0
    pane1 = np.array([[-5, -5, 1], [5, -5, 1], [5, 5, 1], [-5, 5, 1]], np.float32)
    pane2 = pane1 + np.array([0, 0, -1], np.float32)
    light = np.array([[-30, -30, -3], [30, -30, -3], [30, 30, -3], [-30, 30, -3]], np.float32)   # faces +z
    pos = np.concatenate([pane1, pane2, light])
    idx = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7], [8, 9, 10], [8, 10, 11]], np.uint32)
    tm = np.array([0, 0, 0, 0, -1, -1], np.int32)
    cam = dict(eye=(0.0, 0.0, 4.0), gaze=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fov_y_deg=20.0)
    mp = np.zeros((1, 8), np.float32); mp[0, 0] = 1.3; mp[0, 4] = 1.3
    for mode in (0, 1, 2):
        img = O.oracle_render(pos, idx, tm, [[0.5, 0.5, 0.5]], [L, L, L], cam, 8, 8, 16, sample_mode=mode, material_type=[2], material_params=mp)
        assert np.allclose(img, L, rtol=1e-5), (mode, img.min(), img.max())


def test_glossy_oracle_matches_reference_render():
    """(Mt)Unreal + (Mt)Refract against the reference's render of scenes.cornell_glossy (RGB renderer). The glass box
    makes the scene noisy (caustic paths): 8x8 block means carry the converged comparison."""
    path = os.path.join(GOLDEN, "render_cornell64_glossy_spp16384.npz")
    if not os.path.exists(path):
        pytest.skip("golden image was not generated")
    c, tm, alb, mtype, mparams = glossy()
    ref = ref_image("cornell64_glossy_spp16384")
    img = O.oracle_render(c["positions"], c["indices"], tm, alb, c["radiance"], c["camera"], 64, 64, 1024,
                          sample_mode=2, seed=47, material_type=mtype, material_params=mparams)
    err = rel_mse(block_mean(img, 8), block_mean(ref, 8))
    assert err <= 2e-3, err
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.02), (img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)))
    assert rel_mse(block_mean(ref, 8), block_mean(ref_image("cornell64_spp16384"), 8)) > 5e-3     # the materials are visible


def test_smooth_normals_oracle_matches_reference_render():
    """Shading normals that differ from the geometric ones: the 80-triangle sphere with radial vertex normals
    (scenes.cornell_sphere) against the reference's render — the interpolated tangent frames of Triangle::GenerateSurface
    (Quaternion::BarySLerp of the vertex quaternions, Z axis as the shading normal)."""
    path = os.path.join(GOLDEN, "render_cornell64_sphere_spp16384.npz")
    if not os.path.exists(path):
        pytest.skip("golden image was not generated")
    from mray_b200 import scenes
    c = scenes.cornell_sphere()
    tm = np.where(c["material"] == 3, -1, c["material"]).astype(np.int32)
    ref = ref_image("cornell64_sphere_spp16384")
    img = O.oracle_render(c["positions"], c["indices"], tm, c["albedo"][:3], c["radiance"], c["camera"], 64, 64, 1024,
                          sample_mode=2, seed=48, vertex_normals=c["normals"])
    err = rel_mse(block_mean(img, 4), block_mean(ref, 4))
    assert err <= 1e-3, err
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.01), (img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)))
    # the smooth normals matter: the same mesh shaded with its face normals is visibly different on the sphere
    flat = O.oracle_render(c["positions"], c["indices"], tm, c["albedo"][:3], c["radiance"], c["camera"], 64, 64, 1024, sample_mode=2, seed=49)
    sphere = (slice(4, 24), slice(36, 60))          # rows / columns covering the sphere (row 0 = bottom)
    assert rel_mse(img[sphere], ref[sphere]) < 0.5 * rel_mse(flat[sphere], ref[sphere])


# ------------------------------------------------------------------------------------------------
# skysphere boundary lights, pinned by reference renders of scenes.cornell_open
# ------------------------------------------------------------------------------------------------
SKY_ROTATION = [[0.0, 0.0, 1.0, 0.0], [0.0, 1.0, 0.0, 0.0], [-1.0, 0.0, 0.0, 0.0]]


def open_cornell(keep_light=False):
    c = scenes.cornell_open(keep_light=keep_light)
    return c, np.where(c["material"] == 3, -1, c["material"]).astype(np.int32)


@pytest.mark.parametrize("name,mode,spp,kw", [
    ("cornell64_sky_const_spp16384", 2, 256, dict(boundary=dict(type="Skysphere_Spherical", radiance=(1.5, 1.8, 2.5)))),
    ("cornell64_sky_tex_spp16384", 2, 1024, dict(boundary=dict(type="Skysphere_Spherical", texture=0, transform=SKY_ROTATION))),
    ("cornell64_sky_nee_spp16384", 1, 1024, dict(boundary=dict(type="Skysphere_Spherical", texture=0, transform=SKY_ROTATION))),
])
def test_oracle_skysphere_matches_reference_render(name, mode, spp, kw):
    """LightSkysphere (constant: uniform uv sampling; textured: the luminance PwC distribution, under a (T)Single rotation)
    as the boundary light, NEE+MIS and NEE-only: oracle at `spp` vs the reference's 16384-spp image, 8x8 block means."""
    path = os.path.join(GOLDEN, f"render_{name}.npz")
    if not os.path.exists(path):
        pytest.skip("golden image was not generated")
    ref = np.load(path)["img"].astype(np.float32)
    c, tm = open_cornell()
    if "texture" in kw["boundary"]:
        kw = dict(kw, textures=[scenes.sky_texture()], albedo_texture=[-1, -1, -1])
    img = O.oracle_render(c["positions"], c["indices"], tm, c["albedo"][:3], c["radiance"], c["camera"], 64, 64, spp,
                          sample_mode=mode, seed=5, **kw)
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.01), (img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)))
    e = rel_mse(block_mean(img, 8), block_mean(ref, 8))
    # NEE-only under a "sun" is the noisiest estimator here: measured 1.9e-3 at 1024 spp, 7.8e-4 at 4096 spp (means within 0.05 %)
    assert e <= (3e-3 if mode == 1 else 1e-3), e


def test_oracle_alpha_map_matches_reference_render():
    """SurfaceParams.alphaMaps: the stochastic alpha test of IntersectionCheck (AcceleratorLBVH.hpp:L263-282) on a pane with
    transparent / opaque / fractional texels (scenes.cornell_alpha), closest-hit and shadow rays alike."""
    path = os.path.join(GOLDEN, "render_cornell64_alpha_spp16384.npz")
    if not os.path.exists(path):
        pytest.skip("golden image was not generated")
    ref = np.load(path)["img"].astype(np.float32)
    c = scenes.cornell_alpha()
    tm = np.where(c["material"] == 3, -1, np.where(c["material"] == 4, 3, c["material"])).astype(np.int32)
    tri_alpha = np.where(c["material"] == 4, 0, -1).astype(np.int32)
    img = O.oracle_render(c["positions"], c["indices"], tm, c["albedo"][[0, 1, 2, 4]], c["radiance"], c["camera"], 64, 64, 1024,
                          sample_mode=2, seed=9, textures=[c["alpha_texture"]], vertex_uvs=c["uvs"], tri_alpha=tri_alpha)
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.01), (img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)))
    e = rel_mse(block_mean(img, 4), block_mean(ref, 4))
    assert e <= 1e-3, e
    # and the alpha map matters: without it the pane is opaque and the image differs by far more than the noise
    opaque = O.oracle_render(c["positions"], c["indices"], tm, c["albedo"][[0, 1, 2, 4]], c["radiance"], c["camera"], 64, 64, 256, sample_mode=2, seed=9)
    assert rel_mse(block_mean(opaque, 4), block_mean(ref, 4)) > 20 * e


def test_oracle_normal_map_matches_reference_render():
    """The optional "normalMap" of (Mt)Lambert: Triangle::GenerateSurface re-aims the hit's tangent frame at the texture's
    tangent-space normal (PrimitiveDefaultTriangle.hpp:L571-575). scenes.cornell_normal_map, vs the reference's render."""
    path = os.path.join(GOLDEN, "render_cornell64_normalmap_spp16384.npz")
    if not os.path.exists(path):
        pytest.skip("golden image was not generated")
    ref = np.load(path)["img"].astype(np.float32)
    c = scenes.cornell_normal_map()
    tm = np.where(c["material"] == 3, -1, c["material"]).astype(np.int32)
    kw = dict(sample_mode=2, seed=4, textures=[c["normal_texture"]], vertex_uvs=c["uvs"], vertex_normals=c["normals"])
    img = O.oracle_render(c["positions"], c["indices"], tm, c["albedo"][:3], c["radiance"], c["camera"], 64, 64, 1024,
                          normal_texture=c["normal_map"][:3], **kw)
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.01), (img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)))
    e = rel_mse(block_mean(img, 4), block_mean(ref, 4))
    assert e <= 1e-3, e
    flat = O.oracle_render(c["positions"], c["indices"], tm, c["albedo"][:3], c["radiance"], c["camera"], 64, 64, 256, **kw)
    assert rel_mse(block_mean(flat, 4), block_mean(ref, 4)) > 5 * e      # the bump is visible: without the map the image differs


def test_oracle_converted_textures_match_reference_render():
    """TextureMemory::ConvertColorspaces: textures declared REC_709 + gamma 2.2 (fp32) and gamma 2.2 (unorm8) are converted to the
    tracer's ACES_CG space at load; the oracle renders with the numpy restatement of that conversion (oracle_lib.convert_texture_color)."""
    path = os.path.join(GOLDEN, "render_cornell64_srgbtex_spp16384.npz")
    if not os.path.exists(path):
        pytest.skip("golden image was not generated")
    ref = np.load(path)["img"].astype(np.float32)
    c, tm = cornell()
    uvs, textures, at = scenes.cornell_textures()
    conv = [dict(textures[0], data=O.convert_texture_color(textures[0]["data"], 2.2, O.rgb_to_rgb_matrix("REC_709", "ACES_CG"))),
            dict(textures[1], data=O.convert_texture_color(textures[1]["data"], 2.2, None))]
    img = O.oracle_render(c["positions"], c["indices"], tm, c["albedo"][:3], c["radiance"], c["camera"], 64, 64, 1024, sample_mode=2, seed=6,
                          textures=conv, albedo_texture=at[:3], vertex_uvs=uvs)
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.01), (img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)))
    e = rel_mse(block_mean(img, 4), block_mean(ref, 4))
    assert e <= 1e-3, e
    plain = O.oracle_render(c["positions"], c["indices"], tm, c["albedo"][:3], c["radiance"], c["camera"], 64, 64, 256, sample_mode=2, seed=6,
                            textures=textures, albedo_texture=at[:3], vertex_uvs=uvs)
    assert rel_mse(block_mean(plain, 4), block_mean(ref, 4)) > 5 * e    # unconverted textures give a visibly different image
