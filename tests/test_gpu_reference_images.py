"""GPU (B200): the wavefront path tracer against IMAGES RENDERED BY THE UNMODIFIED REFERENCE (its CPU
backend driven through TracerI, oracle/gen_golden_render.py -> tests/golden/render_*.npz). This is BASELINE
config 1's acceptance test: Cornell box, (R)PathTracerRGB / (R)PathTracerSpectral,
WithNEEAndMIS, rrRange [2,20], Gaussian film filter r=1, independent sampler.

Statistical parity, north-star tolerance relMSE <= 1e-3 on converged images. relMSE(a, b) =
mean((a-b)^2 / (b^2 + 1e-2)); between two independent estimates of N and M spp of this scene it is
~ 7.7 (1/N + 1/M) (measured reference-vs-reference and oracle-vs-oracle)."""
import os

import numpy as np
import pytest

import oracle_lib as O
from mray_b200 import capi, scenes, spectral
from test_gpu_render import cornell_accel, rel_mse, REL_MSE_TOL

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def ref_image(name):
    z = np.load(os.path.join(GOLDEN, f"render_{name}.npz"))
    return z["img"].astype(np.float32), int(z["spp"])


def block_mean(img, k):
    h, w, c = img.shape
    return img.reshape(h // k, k, w // k, k, c).mean(axis=(1, 3))


def render(ctx, acc, c, idx, res, spp, seed, **kw):
    r = capi.Renderer(ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"],
                      res, res, spp, sample_mode="WithNEEAndMIS", rr_range=(2, 20), seed=seed, **kw)
    img, st = r.render(batch=64)
    assert st.finished and st.pathsCompleted == spp * res * res
    r.close()
    return img


def test_config1_cornell_512_64spp_against_reference(gpu_ctx):
    """Config 1 as stated: 512x512, 64 spp. Our 64-spp image and the reference's own 64-spp image are both
    compared with the reference's 1024-spp image: the noise floors must match, and the block-averaged
    (16x16 pixels -> 16384 samples per block) images must agree to the converged tolerance."""
    ref64, _ = ref_image("cornell512_spp64")
    refhi, _ = ref_image("cornell512_spp1024")
    c, idx, tm, acc = cornell_accel(gpu_ctx)
    ours = render(gpu_ctx, acc, c, idx, 512, 64, seed=0)
    e_ref, e_ours = rel_mse(ref64, refhi), rel_mse(ours, refhi)
    assert abs(e_ours / e_ref - 1.0) < 0.05, (e_ours, e_ref)
    # no bias: the difference image averaged over blocks is pure noise of the expected size
    b_ours, b_ref64, b_refhi = block_mean(ours, 16), block_mean(ref64, 16), block_mean(refhi, 16)
    eb_ours, eb_ref = rel_mse(b_ours, b_refhi), rel_mse(b_ref64, b_refhi)
    assert eb_ours <= REL_MSE_TOL, eb_ours
    assert eb_ours < 1.5 * eb_ref + 1e-4, (eb_ours, eb_ref)
    assert np.allclose(ours.mean(axis=(0, 1)), refhi.mean(axis=(0, 1)), rtol=5e-3), (ours.mean(axis=(0, 1)), refhi.mean(axis=(0, 1)))
    acc.close()


def test_cornell_converged_against_reference(gpu_ctx):
    """128x128: reference at 16384 spp, ours at 65536 spp -> expected relMSE 7.7 (1/16384 + 1/65536) ~ 6e-4."""
    ref, spp_ref = ref_image("cornell128_spp16384")
    c, idx, tm, acc = cornell_accel(gpu_ctx)
    ours = render(gpu_ctx, acc, c, idx, 128, 65536, seed=5)
    err = rel_mse(ours, ref)
    assert err <= REL_MSE_TOL, err
    # channel means over the whole image: a 0.3 % test of the estimator's expectation
    assert np.allclose(ours.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=3e-3), (ours.mean(axis=(0, 1)), ref.mean(axis=(0, 1)))
    acc.close()


@pytest.mark.skipif(not spectral.available(), reason="spectral LUT was not generated")
def test_spectral_cornell_converged_against_reference(gpu_ctx):
    """(R)PathTracerSpectral (hero wavelengths, HyperbolicPBRT sampling, Jakob-2019 upsampling, ACES_CG) against the
    reference's spectral renderer."""
    ref, spp_ref = ref_image("cornell128_spectral_spp16384")
    c, idx, tm, acc = cornell_accel(gpu_ctx)
    sp = capi.Spectrum(gpu_ctx, spectral.load(), "HyperbolicPBRT")
    ours = render(gpu_ctx, acc, c, idx, 128, 65536, seed=6, spectrum=sp)
    err = rel_mse(ours, ref)
    # colour noise of the 4-wavelength estimator adds to the RGB constant (reference-vs-reference ~ 1.3x)
    assert err <= REL_MSE_TOL, err
    assert np.allclose(ours.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=5e-3), (ours.mean(axis=(0, 1)), ref.mean(axis=(0, 1)))
    sp.close(); acc.close()


@pytest.mark.parametrize("mode,name", [("WithNextEventEstimation", "cornell64_nee_spp16384"), ("Pure", "cornell64_pure_spp16384")])
def test_other_sample_modes_against_reference(gpu_ctx, mode, name):
    """NEE without MIS and pure path tracing: same expectation as the reference's images of those modes."""
    ref, spp_ref = ref_image(name)
    c, idx, tm, acc = cornell_accel(gpu_ctx)
    r = capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"],
                      64, 64, 65536, sample_mode=mode, rr_range=(2, 20), seed=8)
    ours, st = r.render(batch=64)
    r.close()
    mask = ref.max(axis=-1) < 5.0                      # away from the directly visible light
    assert np.allclose(ours[mask].mean(axis=0), ref[mask].mean(axis=0), rtol=0.01), (mode, ours[mask].mean(axis=0), ref[mask].mean(axis=0))
    if mode != "Pure":                                # pure path tracing is far from converged at these counts
        assert rel_mse(block_mean(ours, 2), block_mean(ref, 2)) <= 2 * REL_MSE_TOL
    acc.close()


PLUGIN = os.path.join(os.path.dirname(GOLDEN), "..", "mray_b200", "lib", "libTracerDLL_B200.so")


@pytest.mark.skipif(not (os.path.exists(PLUGIN) and O.driver_available()), reason="plugin / driver were not prebuilt")
def test_two_level_cornell_through_tracer_interface_against_reference():
    """The SAME TracerI call sequence (per-batch (T)Single transforms, local-space vertices) that produced the
    reference's image, replayed on the B200 plugin. Transform family: translation + axis-permutation rotations,
    the one for which the reference's own two-level image equals its flat image (oracle/gen_golden_render.py)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "oracle"))
    import gen_golden_render as G
    ref, spp_ref = ref_image("cornell64_single_spp16384")
    c = scenes.cornell_box()
    b = O.batched_scene(c["positions"], c["indices"], c["material"])
    bt = G.localise(b)
    img, w, st = O.driver_render(os.path.abspath(PLUGIN), b, c["albedo"], 3, c["radiance"], c["camera"], 64, 64, 32768,
                                 sample_mode="WithNEEAndMIS", rr_range=(2, 20), seed=9, batch_transforms=bt)
    assert np.allclose(w, 32768, rtol=1e-3)
    err = rel_mse(block_mean(img, 2), block_mean(ref, 2))      # 131072 vs 65536 samples per block
    assert err <= REL_MSE_TOL, err
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=5e-3)


def instanced_cornell():
    """The Cornell box as the reference's own documentation builds it (Docs/markdown/scene/mrayScene.md:L310-553):
    walls are INSTANCES of one unit plane under (T)Single transforms, each with its own material; boxes, light and
    the back wall keep their own batches. (The back wall needs a rotation about X, for which the reference's affine
    inverse is wrong — Core/Matrix.hpp:L892 has +s1 for -s1, s1 = m00 m12 - m02 m10 — so it stays a plain batch
    and the same scene can be rendered by the reference. The B200 plugin has no such restriction.)
    Returns (batched dict, per-batch 3x4 transforms, instance_of)."""
    c = scenes.cornell_box()
    tri_is_wall = np.isin(np.arange(c["indices"].shape[0]), [0, 1, 2, 3, 6, 7, 8, 9])
    rest = O.batched_scene(c["positions"], c["indices"][~tri_is_wall], c["material"][~tri_is_wall])   # back + boxes (0), light (3)
    plane_p = np.array([[-1, 0, 1], [1, 0, 1], [1, 0, -1], [-1, 0, -1]], np.float32)
    plane_n = np.tile(np.array([[0, 1, 0]], np.float32), (4, 1))
    plane_i = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)
    walls = [  # (material, rotation, translation): floor, ceiling, left (red), right (green)
        (0, np.eye(3), (0, 0, 0)),
        (0, np.diag([1.0, -1.0, -1.0]), (0, 2, 0)),
        (1, np.array([[0.0, 1, 0], [-1, 0, 0], [0, 0, 1]]), (-1, 1, 0)),
        (2, np.array([[0.0, -1, 0], [1, 0, 0], [0, 0, 1]]), (1, 1, 0)),
    ]
    nb0 = len(rest["materials"])
    mats = list(rest["materials"]) + [w[0] for w in walls]
    vo, to = list(rest["vertex_offsets"]), list(rest["tri_offsets"])
    P, N, I = [rest["positions"]], [rest["normals"]], [rest["indices"]]
    for _ in walls:     # every wall batch carries the plane (only the first one's geometry is used)
        P.append(plane_p); N.append(plane_n); I.append(plane_i)
        vo.append(vo[-1] + 4); to.append(to[-1] + 2)
    b = dict(materials=np.array(mats), vertex_offsets=np.array(vo, np.uint32), tri_offsets=np.array(to, np.uint32),
             positions=np.ascontiguousarray(np.concatenate(P), np.float32), normals=np.ascontiguousarray(np.concatenate(N), np.float32),
             indices=np.ascontiguousarray(np.concatenate(I), np.uint32))
    ident = np.hstack([np.eye(3), np.zeros((3, 1))])
    bt = np.stack([ident] * nb0 + [np.hstack([R, np.array(t, np.float64)[:, None]]) for _, R, t in walls])
    inst = np.array([-1] * nb0 + [-1] + [nb0] * (len(walls) - 1), np.int32)
    return c, b, bt, inst


@pytest.mark.skipif(not (os.path.exists(PLUGIN) and O.driver_available()), reason="plugin / driver were not prebuilt")
def test_instanced_cornell_through_tracer_interface_against_reference():
    """Surfaces that share a primitive batch: the plugin builds ONE accelerator for the unit plane and four instances
    (floor, ceiling, left, right) with their own transform and material key (mrb_instance_desc.lightOrMatKeys); the image equals the reference's
    image of the same world."""
    ref, spp_ref = ref_image("cornell64_spp16384")
    c, b, bt, inst = instanced_cornell()
    img, w, st = O.driver_render(os.path.abspath(PLUGIN), b, c["albedo"], 3, c["radiance"], c["camera"], 64, 64, 32768,
                                 sample_mode="WithNEEAndMIS", rr_range=(2, 20), seed=10, batch_transforms=bt, instance_of=inst)
    assert np.allclose(w, 32768, rtol=1e-3)
    box = np.array(st["aabb"])
    assert np.allclose(box, [-1, 0, -1, 1, 2, 1], atol=1e-5)
    err = rel_mse(block_mean(img, 2), block_mean(ref, 2))
    assert err <= REL_MSE_TOL, err
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=5e-3)


def test_textured_cornell_against_reference(gpu_ctx):
    """Textured Lambert albedo through the C-ABI renderer against the reference's textured render."""
    ref, spp_ref = ref_image("cornell64_textured_spp16384")
    c, idx, tm, acc = cornell_accel(gpu_ctx)
    uvs, textures, at = scenes.cornell_textures()
    ours = render(gpu_ctx, acc, c, idx, 64, 65536, seed=12, textures=textures, albedo_texture=at[:3], vertex_uvs=uvs)
    err = rel_mse(block_mean(ours, 2), block_mean(ref, 2))
    assert err <= REL_MSE_TOL, err
    assert np.allclose(ours.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=5e-3), (ours.mean(axis=(0, 1)), ref.mean(axis=(0, 1)))
    acc.close()


@pytest.mark.skipif(not (os.path.exists(PLUGIN) and O.driver_available()), reason="plugin / driver were not prebuilt")
def test_textured_cornell_through_tracer_interface_against_reference():
    """CreateTexture2D -> CommitTextures -> PushTextureData -> PushMatAttribute(albedo, texture ids) + UV0, the call
    sequence that produced the reference's image, replayed on the B200 plugin."""
    ref, spp_ref = ref_image("cornell64_textured_spp16384")
    c = scenes.cornell_box()
    uvs, textures, at = scenes.cornell_textures()
    b = O.batched_scene(c["positions"], c["indices"], c["material"], uvs=uvs)
    img, w, st = O.driver_render(os.path.abspath(PLUGIN), b, c["albedo"], 3, c["radiance"], c["camera"], 64, 64, 32768,
                                 sample_mode="WithNEEAndMIS", rr_range=(2, 20), seed=13, textures=textures, material_texture=at)
    assert np.allclose(w, 32768, rtol=1e-3)
    err = rel_mse(block_mean(img, 2), block_mean(ref, 2))
    assert err <= REL_MSE_TOL, err
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=5e-3)


def test_image_rendered_as_regions_against_reference(gpu_ctx):
    """RenderImageParams{resolution, regionMin, regionMax}: four renderers, one 32x32 region each, assemble the
    reference's 64x64 image (the hook tile- and GPU-sharding of one image uses)."""
    ref, spp_ref = ref_image("cornell64_spp16384")
    c, idx, tm, acc = cornell_accel(gpu_ctx)
    ours = np.zeros((64, 64, 3), np.float32)
    for k, (x0, y0) in enumerate([(0, 0), (32, 0), (0, 32), (32, 32)]):
        ours[y0:y0 + 32, x0:x0 + 32] = render(gpu_ctx, acc, c, idx, 32, 65536, seed=20 + k, full_resolution=(64, 64), region_min=(x0, y0))
    err = rel_mse(block_mean(ours, 2), block_mean(ref, 2))
    assert err <= REL_MSE_TOL, err
    assert np.allclose(ours.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=5e-3)
    acc.close()


@pytest.mark.skipif(not (os.path.exists(PLUGIN) and O.driver_available()), reason="plugin / driver were not prebuilt")
def test_render_regions_through_tracer_interface_against_reference():
    """StartRender with a sub-region: the left and right halves, delivered as RenderImageSections placed at
    pixelMin / pixelMax of the full image."""
    ref, spp_ref = ref_image("cornell64_spp16384")
    c = scenes.cornell_box()
    b = O.batched_scene(c["positions"], c["indices"], c["material"])
    img = np.zeros((64, 64, 3), np.float32)
    for k, reg in enumerate([(0, 0, 24, 64), (24, 0, 64, 64)]):
        part, w, st = O.driver_render(os.path.abspath(PLUGIN), b, c["albedo"], 3, c["radiance"], c["camera"], 64, 64, 32768,
                                      sample_mode="WithNEEAndMIS", rr_range=(2, 20), seed=30 + k, region=reg)
        inside = np.zeros((64, 64), bool); inside[reg[1]:reg[3], reg[0]:reg[2]] = True
        assert np.allclose(w[inside], 32768, rtol=1e-3) and np.all(w[~inside] == 0)
        img[inside] = part[inside]
    err = rel_mse(block_mean(img, 2), block_mean(ref, 2))
    assert err <= REL_MSE_TOL, err


def test_mirror_cornell_against_reference(gpu_ctx):
    """(Mt)Reflect through the C-ABI renderer against the reference's render (4x4 block means: the scene is ~5x noisier)."""
    from test_gpu_render import mirror_accel
    ref, spp_ref = ref_image("cornell64_mirror_spp16384")
    c, idx, tm, acc, alb, mtype = mirror_accel(gpu_ctx)
    r = capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], alb, c["radiance"], c["camera"], 64, 64, 65536,
                      sample_mode="WithNEEAndMIS", rr_range=(2, 20), seed=14, material_type=mtype)
    ours, st = r.render(batch=64); r.close(); acc.close()
    err = rel_mse(block_mean(ours, 4), block_mean(ref, 4))
    assert err <= REL_MSE_TOL, err
    assert np.allclose(ours.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=5e-3), (ours.mean(axis=(0, 1)), ref.mean(axis=(0, 1)))


@pytest.mark.skipif(not (os.path.exists(PLUGIN) and O.driver_available()), reason="plugin / driver were not prebuilt")
def test_mirror_cornell_through_tracer_interface_against_reference():
    """CreateMaterialGroup("(Mt)Reflect") through TracerI on the B200 plugin."""
    ref, spp_ref = ref_image("cornell64_mirror_spp16384")
    c = scenes.cornell_mirror()
    b = O.batched_scene(c["positions"], c["indices"], c["material"])
    img, w, st = O.driver_render(os.path.abspath(PLUGIN), b, c["albedo"], 3, c["radiance"], c["camera"], 64, 64, 32768,
                                 sample_mode="WithNEEAndMIS", rr_range=(2, 20), seed=15, material_kind=c["material_type"])
    assert np.allclose(w, 32768, rtol=1e-3)
    err = rel_mse(block_mean(img, 4), block_mean(ref, 4))
    assert err <= REL_MSE_TOL, err
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=5e-3)
