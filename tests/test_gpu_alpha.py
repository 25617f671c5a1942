"""GPU (B200): alpha maps — SurfaceParams.alphaMaps and the stochastic alpha test inside IntersectionCheck
(Tracer/AcceleratorLBVH.hpp:L263-282), SURVEY.md §8f rank 1. Closed forms on the casts (single accelerator and two-level
scene, wide and exact-binary kernels, closest and visibility), determinism of the per-(ray, triangle) decisions, and the
reference's own render of scenes.cornell_alpha through the C-ABI renderer and the TracerI plugin."""
import os

import numpy as np
import pytest
import torch

import oracle_lib as O
from mray_b200 import capi, scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
PLUGIN = os.path.join(ROOT, "mray_b200", "lib", "libTracerDLL_B200.so")
bm = lambda x, k: x.reshape(x.shape[0] // k, k, x.shape[1] // k, k, 3).mean(axis=(1, 3))
rel = lambda a, b: float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))

# two parallel unit quads facing +z: FRONT at z = 1 (alpha mapped), BACK at z = 0 (opaque)
QUADS = np.array([[0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1], [0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
QIDX = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7]], np.uint32)
QUV = np.array([[0, 0], [1, 0], [1, 1], [0, 1]] * 2, np.float32)


def grid_rays(n, z0=3.0):
    """n x n rays along -z over the quads' interior"""
    g = (np.arange(n, dtype=np.float32) + 0.5) / n
    x, y = np.meshgrid(g, g)
    r = np.zeros((n * n, 8), np.float32)
    r[:, 0], r[:, 1], r[:, 2], r[:, 3] = x.ravel(), y.ravel(), z0, 0.0
    r[:, 6], r[:, 7] = -1.0, 1.0e30
    return r


def cast(obj, rays, mode, any_hit=False):
    n = rays.shape[0]
    w = torch.from_numpy(rays.copy()).cuda()
    if any_hit:
        bits = torch.full(((n + 31) // 32,), -1, dtype=torch.int32, device="cuda")
        obj.cast_visibility_rays(bits, w, None, mode)
        b = bits.cpu().numpy().view(np.uint32)
        return ((b[np.arange(n) >> 5] >> (np.arange(n) & 31)) & 1) == 0        # occluded
    k = torch.full((n, 4), -1, dtype=torch.int32, device="cuda"); h = torch.zeros((n, 2), dtype=torch.float32, device="cuda")
    obj.cast_rays(k, h, w, None, mode)
    return k.cpu().numpy().view(np.uint32)[:, 0], w.cpu().numpy()[:, 7]


def two_quads(ctx, alpha_data, interp="Nearest"):
    return capi.Accelerator(ctx, QUADS, QIDX, prim_ranges=[[0, 2], [2, 4]], light_or_mat_keys=[0, 1], vertex_uvs=QUV,
                            alpha_textures=[dict(data=alpha_data, interp=interp, edge="Clamp")], range_alpha_map=[0, -1])


def test_constant_alpha_keeps_that_fraction_of_the_hits(gpu_ctx):
    rays = grid_rays(256)
    n = rays.shape[0]
    for alpha in (0.0, 0.3, 1.0):
        acc = two_quads(gpu_ctx, np.full((1, 1), alpha, np.float32))
        res = {}
        for mode in (capi.MRB_TRACE_WIDE, capi.MRB_TRACE_BINARY_EXACT):
            gpu_ctx.set_alpha_seed(1234)
            prim, t = cast(acc, rays, mode)
            front = (prim & 0x0FFFFFFF) < 2
            assert np.all(t[front] == 2.0) and np.all(t[~front] == 3.0)      # what the front pane lets through hits the back one
            assert abs(front.mean() - alpha) <= 4 * np.sqrt(max(alpha * (1 - alpha), 1e-9) / n) + 1e-9, (alpha, front.mean())
            res[mode] = prim
        # the decision is a function of (seed, ray, triangle): the wide kernel and the reference's binary algorithm agree ray by ray
        assert np.array_equal(res[capi.MRB_TRACE_WIDE], res[capi.MRB_TRACE_BINARY_EXACT])
        acc.close()


def test_visibility_casts_and_seeds(gpu_ctx):
    # only the front pane in the way of a bounded ray: occlusion probability = alpha
    rays = grid_rays(256)
    rays[:, 7] = 2.5                      # stops between the panes (front at t = 2, back at t = 3)
    n = rays.shape[0]
    acc = two_quads(gpu_ctx, np.full((1, 1), 0.6, np.float32))
    gpu_ctx.set_alpha_seed(77)
    occ_w = cast(acc, rays, capi.MRB_TRACE_WIDE, any_hit=True)
    gpu_ctx.set_alpha_seed(77)
    occ_b = cast(acc, rays, capi.MRB_TRACE_BINARY_EXACT, any_hit=True)
    assert np.array_equal(occ_w, occ_b)
    assert abs(occ_w.mean() - 0.6) <= 4 * np.sqrt(0.24 / n)
    gpu_ctx.set_alpha_seed(77)
    assert np.array_equal(cast(acc, rays, capi.MRB_TRACE_WIDE, any_hit=True), occ_w)            # same seed, same decisions
    nxt = cast(acc, rays, capi.MRB_TRACE_WIDE, any_hit=True)                                      # the seed advanced: independent decisions
    assert abs((nxt & occ_w).mean() - 0.36) <= 4 * np.sqrt(0.36 * 0.64 / n)
    acc.close()


def test_alpha_follows_the_texture_and_the_uvs(gpu_ctx):
    """A 2 x 1 texture: left texel transparent, right texel opaque; nearest and bilinear filtering."""
    rays = grid_rays(128)
    acc = two_quads(gpu_ctx, np.array([[0.0, 1.0]], np.float32), interp="Nearest")
    prim, t = cast(acc, rays, capi.MRB_TRACE_WIDE)
    front = (prim & 0x0FFFFFFF) < 2
    assert np.array_equal(front, rays[:, 0] > 0.5)
    acc.close()
    # bilinear + clamp: alpha ramps from 0 at u = 0.25 to 1 at u = 0.75 -> the kept fraction of a column follows the ramp
    acc = two_quads(gpu_ctx, np.array([[0.0, 1.0]], np.float32), interp="Linear")
    prim, t = cast(acc, rays, capi.MRB_TRACE_WIDE)
    front = ((prim & 0x0FFFFFFF) < 2).reshape(128, 128)
    u = (np.arange(128) + 0.5) / 128
    expect = np.clip((u - 0.25) / 0.5, 0.0, 1.0)
    assert np.abs(front.mean(axis=0) - expect).max() < 0.2 and abs(front.mean() - expect.mean()) < 0.02
    acc.close()


def test_two_level_scene_alpha(gpu_ctx):
    """The same panes as an INSTANCE under a rigid transform: the alpha test runs in the instance's local space; two
    instances of one accelerator decide independently."""
    acc = two_quads(gpu_ctx, np.full((1, 1), 0.5, np.float32))
    t1 = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]], np.float32)
    t2 = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 5]], np.float32)       # a second copy 5 units closer to the ray origins
    sc = capi.Scene(gpu_ctx, [(acc, None), (acc, t2)])
    rays = grid_rays(256, z0=9.0)
    n = rays.shape[0]
    res = {}
    for mode in (capi.MRB_TRACE_WIDE, capi.MRB_TRACE_BINARY_EXACT):
        gpu_ctx.set_alpha_seed(5)
        prim, t = cast(sc, rays, mode)
        res[mode] = t
    assert np.array_equal(res[capi.MRB_TRACE_WIDE], res[capi.MRB_TRACE_BINARY_EXACT])
    t = res[capi.MRB_TRACE_WIDE]
    # first obstacle: the near instance's front pane (t = 3, kept with probability 1/2), else its opaque back pane (t = 4)
    assert set(np.unique(t)) == {3.0, 4.0}
    assert abs((t == 3.0).mean() - 0.5) <= 4 * np.sqrt(0.25 / n)
    # visibility of a ray that stops between the near instance's panes: its front pane alone decides
    r2 = rays.copy(); r2[:, 7] = 3.5
    occ = cast(sc, r2, capi.MRB_TRACE_WIDE, any_hit=True)
    assert abs(occ.mean() - 0.5) <= 4 * np.sqrt(0.25 / n)
    sc.close(); acc.close()


def test_bad_alpha_descriptors_are_refused(gpu_ctx):
    with pytest.raises(capi.MrbError):   # no UVs
        capi.Accelerator(gpu_ctx, QUADS, QIDX, prim_ranges=[[0, 2], [2, 4]], alpha_textures=[dict(data=np.ones((1, 1), np.float32))], range_alpha_map=[0, -1])
    with pytest.raises(capi.MrbError):   # index out of the table
        capi.Accelerator(gpu_ctx, QUADS, QIDX, prim_ranges=[[0, 2], [2, 4]], vertex_uvs=QUV, alpha_textures=[dict(data=np.ones((1, 1), np.float32))],
                         range_alpha_map=[3, -1])


def alpha_cornell_accel(ctx):
    c = scenes.cornell_alpha()
    order = np.argsort(c["material"], kind="stable")
    idx = np.ascontiguousarray(c["indices"][order]); mat = c["material"][order]
    ranges, keys, amap = [], [], []
    flat = {0: 0, 1: 1, 2: 2, 4: 3}              # material id -> index into the renderer's albedo table (3 is the light)
    for m in np.unique(mat):
        w = np.nonzero(mat == m)[0]
        ranges.append([w[0], w[-1] + 1]); keys.append(capi.light_key(0) if m == 3 else flat[int(m)]); amap.append(int(c["alpha_map"][m]))
    acc = capi.Accelerator(ctx, c["positions"], idx, prim_ranges=ranges, light_or_mat_keys=keys, vertex_uvs=c["uvs"],
                           alpha_textures=[c["alpha_texture"]], range_alpha_map=amap)
    return c, idx, acc, c["albedo"][[0, 1, 2, 4]]


def test_alpha_render_matches_the_reference(gpu_ctx):
    path = os.path.join(GOLDEN, "render_cornell64_alpha_spp16384.npz")
    if not os.path.exists(path):
        pytest.skip("golden image was not generated")
    ref = np.load(path)["img"].astype(np.float32)
    c, idx, acc, alb = alpha_cornell_accel(gpu_ctx)
    r = capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], alb, c["radiance"], c["camera"], 64, 64, 16384, seed=91)
    img, st = r.render(batch=32)
    # a fixed seed gives a fixed image, alpha decisions included
    r2 = capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], alb, c["radiance"], c["camera"], 64, 64, 64, seed=91)
    a1, _ = r2.render(); r2.close()
    r3 = capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], alb, c["radiance"], c["camera"], 64, 64, 64, seed=91)
    a2, _ = r3.render(); r3.close()
    r.close(); acc.close()
    assert np.allclose(a1, a2, rtol=1e-5, atol=1e-6)
    e = rel(bm(img, 2), bm(ref, 2))
    assert e <= 1e-3, e
    assert np.allclose(img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.01), (img.mean(axis=(0, 1)), ref.mean(axis=(0, 1)))
    if os.path.exists(PLUGIN) and O.driver_available():
        b = O.batched_scene(c["positions"], c["indices"], c["material"], uvs=c["uvs"])
        pimg, w, pst = O.driver_render(PLUGIN, b, c["albedo"], 3, c["radiance"], c["camera"], 64, 64, 16384, seed=92, burst_size=64,
                                       textures=[c["alpha_texture"]], alpha_map=c["alpha_map"])
        assert np.allclose(w, 16384, rtol=1e-3)
        pe = rel(bm(pimg, 2), bm(ref, 2))
        assert pe <= 1e-3, pe
        assert np.allclose(pimg.mean(axis=(0, 1)), ref.mean(axis=(0, 1)), rtol=0.01)
