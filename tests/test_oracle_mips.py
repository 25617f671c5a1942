"""Mip chains and level-of-detail reads: the C restatement (oracle/pt_oracle.c: orc_texture_generate_mips,
orc_texture_sample_lod / _grad) against golden vectors produced by the reference's own TextureMemory + TracerTexView on its CPU backend
(oracle/gen_golden_texture.py -> tests/golden/texture_mips.npz). CPU only."""
import os

import numpy as np
import pytest

import oracle_lib as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "texture_mips.npz")


def load_cases():
    z = np.load(GOLDEN)
    for name in z["names"]:
        interp, edge, filt, radius, nmips, clamp = z[name + "_params"]
        t = dict(data=z[name + "_data"], interp=str(interp), edge=str(edge))
        if int(clamp):
            t["clamp_res"] = int(clamp)
        if int(nmips):
            t["mips"] = [z[f"{name}_mip{k + 1}"] for k in range(int(nmips))]
        if str(filt):
            t["gen_mips"] = (str(filt), float(radius))
        yield str(name), t, {k: z[f"{name}_{k}"] for k in ["uv", "lod", "dpdx", "dpdy", "rgb_lod", "rgb_grad"]}


CASES = {name: (t, v) for name, t, v in load_cases()}


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_lod_reads_match_the_reference_bit_for_bit(name):
    t, v = CASES[name]
    got = O.oracle_texture_sample_lod(t, v["uv"], lod=v["lod"])
    assert np.array_equal(got, v["rgb_lod"])
    got = O.oracle_texture_sample_lod(t, v["uv"], dpdx=v["dpdx"], dpdy=v["dpdy"], lod_mode=0)
    assert np.array_equal(got, v["rgb_grad"])


def test_mip_chain_layout_and_counts():
    assert [O.full_mip_count(w, h) for w, h in [(1, 1), (2, 1), (32, 16), (20, 12), (64, 64), (65, 3)]] == [1, 2, 6, 5, 7, 7]
    assert O.mip_dims(20, 12, 3) == (2, 1) and O.mip_dims(20, 12, 4) == (1, 1)
    t = CASES["gauss2_u8_npot"][0]
    chain, count = O.mip_chain(t)
    assert count == 5 and chain.shape == (20 * 12 + 10 * 6 + 5 * 3 + 2 * 1 + 1, 4) and chain.dtype == np.uint8
    assert np.array_equal(O.mip_level(chain, 12, 20, 0), t["data"])     # (w, h) = (12, 20): data is [h, w, C]


def test_clamped_sizes_follow_the_reference_rule():
    """ceil(log2(ceil(maxDim / clamp))) levels dropped (TextureMemory::CreateTexture)"""
    assert O.clamped_size(64, 40, 16) == (16, 10) and O.clamped_size(64, 40, 20) == (16, 10) and O.clamped_size(64, 40, 32) == (32, 20)
    assert O.clamped_size(64, 40, 64) == (64, 40) and O.clamped_size(64, 40, 1000) == (64, 40) and O.clamped_size(64, 40, 1) == (1, 1)
    chain, count = O.mip_chain(CASES["clamp16_gen_f32"][0])
    assert count == 5 and chain.shape[0] == 16 * 10 + 8 * 5 + 4 * 2 + 2 * 1 + 1
    # enough levels supplied: the ones that fit are kept as they are (no filtering)
    rng = np.random.default_rng(1)
    lv = [rng.random((8, 8, 4), dtype=np.float32), rng.random((4, 4, 4), dtype=np.float32), rng.random((2, 2, 4), dtype=np.float32)]
    chain, count = O.mip_chain(dict(data=lv[0], mips=lv[1:], clamp_res=4))
    assert count == 2 and np.array_equal(chain[:16].reshape(4, 4, 4), lv[1]) and np.array_equal(chain[16:].reshape(2, 2, 4), lv[2])


def test_box_half_radius_is_the_2x2_average():
    """A known answer: Box with radius 0.5 weighs the four parents of an even-sized level equally."""
    rng = np.random.default_rng(3)
    base = rng.random((8, 8, 4), dtype=np.float32)
    chain, count = O.mip_chain(dict(data=base, gen_mips=("Box", 0.5)))
    lvl1 = O.mip_level(chain, 8, 8, 1)
    expect = base.reshape(4, 2, 4, 2, 4).mean(axis=(1, 3))
    assert count == 4 and np.allclose(lvl1, expect, rtol=0, atol=2e-6)   # 64 weighted terms per texel


def test_constant_texture_stays_constant_through_the_chain():
    base = np.full((16, 16, 4), 0.37, np.float32)
    for filt in [("Gaussian", 2.0), ("Tent", 1.5), ("Mitchell-Netravali", 2.0), ("Box", 1.0)]:
        chain, _ = O.mip_chain(dict(data=base, gen_mips=filt))
        assert np.allclose(chain, 0.37, rtol=0, atol=1e-6), filt


def test_device_lod_mode_is_the_host_mode_with_texel_space_gradients():
    t, v = CASES["gauss2_f32"]
    h, w, _ = t["data"].shape
    size = np.array([w, h], np.float32)
    host = O.oracle_texture_sample_lod(t, v["uv"], dpdx=v["dpdx"] * size, dpdy=v["dpdy"] * size, lod_mode=0)
    dev = O.oracle_texture_sample_lod(t, v["uv"], dpdx=v["dpdx"], dpdy=v["dpdy"], lod_mode=1)
    assert np.array_equal(host, dev)


def test_single_level_reads_ignore_the_gradients():
    t, v = CASES["gauss2_f32"]
    single = dict(data=t["data"])
    assert np.array_equal(O.oracle_texture_sample_lod(single, v["uv"], dpdx=v["dpdx"], dpdy=v["dpdy"]), O.oracle_texture_sample(single, v["uv"]))
    assert np.array_equal(O.oracle_texture_sample_lod(single, v["uv"], lod=v["lod"]), O.oracle_texture_sample(single, v["uv"]))


# ---- the estimator oracle with ray cones against the reference's renders of scenes.cornell_mips ----
from mray_b200 import scenes  # noqa: E402


def _rel(a, b):
    return float(((a - b) ** 2).mean() / (b ** 2).mean())


def _bm(img, k):
    h, w, c = img.shape
    return img.reshape(h // k, k, w // k, k, c).mean(axis=(1, 3))


def _oracle_mip_render(kind, spp, seed, strip=False):
    c = scenes.cornell_mips(kind)
    tm = np.where(c["material"] == 3, -1, c["material"].astype(np.int32))
    textures = [dict(t, gen_mips=c.get("gen_mips")) for t in c["textures"]]
    if strip:
        textures = [dict(t, mips=None, gen_mips=None) for t in textures]
    kw = {}
    if "material_type" in c:
        kw["material_type"] = c["material_type"]
    if "material_params" in c:
        kw["material_params"] = c["material_params"]
    if "normals" in c:
        kw["vertex_normals"] = c["normals"]
    return O.oracle_render(c["positions"], c["indices"], tm, c["albedo"], c["radiance"], c["camera"], 64, 64, spp, seed=seed,
                           textures=textures, albedo_texture=c["albedo_texture"], vertex_uvs=c["uvs"], **kw)


@pytest.mark.parametrize("kind", ["explicit", "sphere_mirror", "gen_glossy"])
def test_oracle_ray_cones_match_reference_render(kind):
    """RayCone::Advance / Project, the curvature and texture-gradient half of Triangle::GenerateSurface, ConeAfterScatter and
    RefractMaterial::RefractRayCone restated in pt_oracle.c: the level every textured read takes decides the colours of these images."""
    path = os.path.join(os.path.dirname(__file__), "golden", f"render_cornell64_mips_{kind}_spp16384.npz")
    if not os.path.exists(path):
        pytest.skip("golden image was not generated")
    ref = np.load(path)["img"].astype(np.float32)
    img = _oracle_mip_render(kind, 1024, 3)
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.01), (img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)))
    e = _rel(_bm(img, 4), _bm(ref, 4))
    assert e <= 1e-3, e
    if kind != "gen_glossy":
        flat = _oracle_mip_render(kind, 256, 4, strip=True)
        assert _rel(_bm(flat, 4), _bm(ref, 4)) > 20 * e


def test_degenerate_sizes():
    """1 x 1 and 1 x N textures: the chain stops at 1 x 1 and a level never drops below one texel per axis."""
    one = dict(data=np.full((1, 1, 4), 0.25, np.float32), gen_mips=("Gaussian", 2.0))
    chain, count = O.mip_chain(one)
    assert count == 1 and chain.shape == (1, 4)
    strip = dict(data=np.arange(7 * 4, dtype=np.float32).reshape(7, 1, 4) / 28.0, gen_mips=("Box", 0.5))
    chain, count = O.mip_chain(strip)
    assert count == 3 and chain.shape == (7 + 3 + 1, 4)
    assert [O.mip_dims(1, 7, k) for k in range(3)] == [(1, 7), (1, 3), (1, 1)]
    uv = np.array([[0.5, 0.5], [0.1, 0.9]], np.float32)
    assert np.allclose(O.oracle_texture_sample_lod(one, uv, lod=[0.0, 5.0]), 0.25)
    top = O.oracle_texture_sample_lod(strip, uv, lod=[2.0, 2.0])
    assert np.allclose(top[0], top[1])           # the 1 x 1 level is one colour everywhere


def test_ray_cone_projection_edge_cases():
    """The two inputs the reference keeps in Tests/Tracer/T_RayCone.cu ("value from an assert"): a ray exactly against the normal
    (h1 would vanish: RayCone::Project nudges d) and a runaway negative width. The footprint axes stay finite; for a regular cone
    they are perpendicular to the normal and, head-on, as long as the cone's radius."""
    L = O.lib()
    L.orc_ray_cone_project.argtypes = [O.C.c_float, O.C.c_float, O.C.c_void_p, O.C.c_void_p, O.C.c_void_p]

    def project(aperture, width, f, d):
        f = np.asarray(f, np.float32); d = np.asarray(d, np.float32); out = np.zeros(6, np.float32)
        L.orc_ray_cone_project(aperture, width, f.ctypes.data, d.ctypes.data, out.ctypes.data)
        return out[:3], out[3:]
    a1, a2 = project(0.000318206352, 0.00251944619, [-0.0, -0.0, -1.0], [-0.0, -0.0, 1.0])
    assert np.isfinite(a1).all() and np.isfinite(a2).all()
    a1, a2 = project(0.0, -5.33283590e+19, [-0.867630541, -0.418694645, 0.268164247], [-0.167027116, -0.138377547, -0.976193428])
    assert np.isfinite(a1).all() and np.isfinite(a2).all()
    f = np.array([0.0, 1.0, 0.0], np.float32)
    d = np.array([0.6, -0.8, 0.0], np.float32)                  # 53 degrees off the normal
    a1, a2 = project(0.01, 0.2, f, d)
    assert abs(float(a1 @ f)) < 1e-6 and abs(float(a2 @ f)) < 1e-6 and abs(float(a1 @ a2)) < 1e-6
    assert np.isclose(np.linalg.norm(a2), 0.1, rtol=1e-5)       # across the plane of incidence: the cone's radius
    assert np.isclose(np.linalg.norm(a1), 0.1 / 0.8, rtol=1e-5)  # along it: stretched by 1 / cos(theta)


def _oracle_refract_cone(aperture, width, beta, wo, n, e0, e1):
    L = O.lib()
    L.orc_refract_ray_cone.argtypes = [O.C.c_float] * 3 + [O.C.c_void_p, O.C.c_void_p, O.C.c_float, O.C.c_float, O.C.c_void_p]
    wo = np.ascontiguousarray(wo, np.float32); n = np.ascontiguousarray(n, np.float32); out = np.zeros(2, np.float32)
    L.orc_refract_ray_cone(aperture, width, beta, wo.ctypes.data, n.ctypes.data, e0, e1, out.ctypes.data)
    return out


def test_refracted_ray_cone_matches_the_reference():
    """pt_oracle.c::refract_ray_cone + cone_after_scatter against RefractMaterial::RefractRayCone + ConeAfterScatter run by the
    reference itself (oracle/gen_golden_raycone.py -> tests/golden/refract_ray_cone.npz, 3 000 random surface cones / directions /
    index pairs). Widths agree to rounding; apertures come out of acos(dot) of two nearly parallel unit vectors — ill-conditioned
    near 1, where one ulp of the dot is ~1e-4 rad — so they are compared at that resolution."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "refract_ray_cone.npz"))
    inp, ref = z["inputs"], z["cones"]
    got = np.zeros_like(ref)
    for i in range(inp.shape[0]):
        e0, e1 = (inp[i, 9], inp[i, 10]) if inp[i, 11] == 0 else (inp[i, 10], inp[i, 9])
        got[i] = _oracle_refract_cone(inp[i, 0], inp[i, 1], inp[i, 2], inp[i, 3:6], inp[i, 6:9], e0, e1)
    assert np.isfinite(got).all()
    assert np.allclose(got[:, 1], ref[:, 1], rtol=1e-3, atol=1e-6)      # measured: 99.9 % within 1e-5, worst (a grazing ray) 2.7e-4
    assert np.allclose(got[:, 0], ref[:, 0], rtol=0, atol=1e-3)         # measured: worst 5.9e-4
    assert (np.abs(got[:, 0] - ref[:, 0]) <= 2e-5).mean() > 0.95 and (got == ref).all(axis=1).mean() > 0.5
    # what the reference does with a diverging cone on a FLAT surface: its 2-D frame keeps y along the normal, the signed angle between the refracted
    # edge rays is negative, and the aperture falls to the Epsilon clamp (+ betaN, minus betaN again in ConeAfterScatter)
    diverging = (inp[:, 0] > 1e-3) & (inp[:, 1] > 1e-3) & (inp[:, 2] == 0) & (np.abs(ref[:, 0] - inp[:, 0]) > 1e-9)   # flat surface, refraction exists
    assert (np.abs(ref[diverging, 0] - 1.0e-5) < 1e-6).mean() > 0.95


def test_refracted_ray_cone_closed_forms():
    """Widths: kept at (near-)normal incidence and for equal indices; under total internal reflection the surface cone comes back
    untouched (ConeAfterScatter then subtracts betaN from the back cone)."""
    n = [0.0, 1.0, 0.0]
    unit = lambda v: np.asarray(v, np.float32) / np.linalg.norm(v)
    assert np.isclose(_oracle_refract_cone(0.02, 0.3, 0.0, unit([1e-4, 1.0, 0.0]), n, 1.0, 1.5)[1], 0.3, rtol=2e-3)
    assert np.isclose(_oracle_refract_cone(0.02, 0.3, 0.0, unit([0.5, 0.8, 0.0]), n, 1.3, 1.3)[1], 0.3, rtol=2e-3)
    assert _oracle_refract_cone(0.02, 0.3, 0.0, unit([0.8, 0.6, 0.0]), n, 1.0, 1.5)[1] > 0.3     # oblique into the denser medium: wider
    ap, w = _oracle_refract_cone(0.02, 0.3, 0.004, unit([0.8, 0.6, 0.0]), n, 1.5, 1.0)     # sin(theta_t) would be 1.2
    assert np.isclose(ap, 0.02 - 0.004, atol=1e-7) and np.isclose(w, 0.3)
