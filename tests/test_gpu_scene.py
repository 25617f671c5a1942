"""GPU (B200): two-level scenes (instances + top-level LBVH) against the oracle's restatement of
BaseAcceleratorLBVH (top-level build, left-first instance order, local-space bottom-level casts)."""
import numpy as np
import pytest

import oracle_lib as O
from mray_b200 import capi, scenes
from test_gpu_parity import dev, host

pytestmark = pytest.mark.gpu


def trs(rng, scale=(0.5, 1.5), extent=12.0):
    a = rng.normal(size=3); a /= np.linalg.norm(a)
    th = rng.uniform(0, 2 * np.pi)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
    S = np.diag(rng.uniform(*scale, size=3))
    t = rng.uniform(-extent, extent, size=3)
    return np.concatenate([R @ S, t[:, None]], 1)


def build_scene(ctx, n_inst, seed, identity_only=False):
    rng = np.random.default_rng(seed)
    meshes = [scenes.arcade_mesh(1500), scenes.random_soup(400, seed=3, size=0.6), (scenes.cornell_box()["positions"], scenes.cornell_box()["indices"])]
    meshes = [(np.ascontiguousarray(p * (1.0 if k == 0 else 6.0), np.float32), i) for k, (p, i) in enumerate(meshes)]
    accs = [capi.Accelerator(ctx, p, i, light_or_mat_keys=[k + 10], prim_ranges=[[0, i.shape[0]]]) for k, (p, i) in enumerate(meshes)]
    obv = [O.oracle_build(p, i) for p, i in meshes]
    inst, oinst, aabbs = [], [], []
    for k in range(n_inst):
        m = k % len(meshes)
        M = None if (identity_only or k == 0) else trs(rng)
        inst.append((accs[m], M))
    sc = capi.Scene(ctx, inst)
    for k, (m32, i32, ident) in enumerate(sc.transforms):
        m = k % len(meshes)
        oinst.append((meshes[m][0], meshes[m][1], obv[m], i32, ident))
        aabbs.append(obv[m].accel_aabb if ident else O.oracle_transform_aabb(m32, obv[m].accel_aabb))
    return sc, accs, oinst, np.array(aabbs, np.float32)


def scene_rays(n, seed, extent=20.0):
    rng = np.random.default_rng(seed)
    o = rng.uniform(-extent, extent, size=(n, 3)); tgt = rng.uniform(-extent * 0.6, extent * 0.6, size=(n, 3))
    d = tgt - o; d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.ascontiguousarray(np.concatenate([o, np.zeros((n, 1)), d, np.full((n, 1), 3.0e38)], 1), np.float32)


def gpu_scene_cast(sc, rays_np, mode):
    import torch
    n = rays_np.shape[0]
    rays = dev(rays_np)
    keys = torch.full((n, 4), -1, dtype=torch.int32, device="cuda")
    hits = torch.zeros((n, 2), dtype=torch.float32, device="cuda")
    sc.cast_rays(keys, hits, rays, None, mode)
    bits = torch.full(((n + 31) // 32,), -1, dtype=torch.int32, device="cuda")
    sc.cast_visibility_rays(bits, dev(rays_np), None, mode)
    torch.cuda.synchronize()
    w = host(bits, np.uint32)
    vis = ((w[np.arange(n) // 32] >> (np.arange(n) % 32).astype(np.uint32)) & 1).astype(bool)
    return host(keys, np.uint32), hits.cpu().numpy(), rays.cpu().numpy(), vis


@pytest.mark.parametrize("n_inst,identity_only", [(1, True), (7, True), (40, False), (600, False)])
def test_two_level_scene(gpu_ctx, n_inst, identity_only):
    sc, accs, oinst, aabbs = build_scene(gpu_ctx, n_inst, seed=n_inst, identity_only=identity_only)
    e = sc.export_tlas()
    assert np.array_equal(e["instance_aabb"], aabbs)                     # world AABBs (FMA-chain transform)
    t = O.oracle_tlas_build(aabbs)
    assert np.array_equal(e["scene_aabb"], t.accel_aabb)
    assert np.array_equal(e["morton"], t.morton) and np.array_equal(e["sorted_idx"], t.sorted_idx)
    if n_inst > 1 and len(np.unique(t.morton)) == n_inst:
        assert np.array_equal(e["nodes"], t.nodes) and np.array_equal(e["boxes"], t.boxes)
    rays = scene_rays(30000, seed=5)
    oi, op, ot, ob = O.oracle_scene_trace(oinst, t, rays, mode=0)
    vi, _, _, _ = O.oracle_scene_trace(oinst, t, rays, mode=1)
    for mode in (capi.MRB_TRACE_WIDE, capi.MRB_TRACE_BINARY_EXACT):
        keys, hits, rout, vis = gpu_scene_cast(sc, rays, mode)
        hit = oi != O.INVALID
        assert np.array_equal(keys[:, 0] != O.INVALID, hit)
        assert np.array_equal(keys[hit, 3], oi[hit])                       # accelKey = instance
        assert np.array_equal(keys[hit, 0], op[hit])                       # primitive id
        assert np.array_equal(rout[:, 7], ot)                              # t bit exact
        assert np.array_equal(hits[hit], ob[hit])
        assert np.array_equal(keys[hit, 1] - 10, oi[hit] % 3)              # lightOrMatKey of the instance's accelerator
        assert np.array_equal(~vis, vi != O.INVALID)
    assert hit.mean() > 0.2
    sc.close()
    for a in accs:
        a.close()


def test_instances_of_one_accelerator_with_their_own_material_keys(gpu_ctx):
    """mrb_instance_desc.lightOrMatKeys: the reference keeps material keys per INSTANCE while surfaces with the same
    primitive batches share one concrete accelerator (Tracer/AcceleratorC.h:L780-905)."""
    p, i = scenes.arcade_mesh(1500)
    half = i.shape[0] // 2
    acc = capi.Accelerator(gpu_ctx, p, i, light_or_mat_keys=[7, 8], prim_ranges=[[0, half], [half, i.shape[0]]])
    rng = np.random.default_rng(3)
    n_inst = 24
    inst = [(acc, None if k == 0 else trs(rng), None if k % 4 == 0 else [100 + 2 * k, 101 + 2 * k]) for k in range(n_inst)]
    sc = capi.Scene(gpu_ctx, inst)
    rays = scene_rays(20000, seed=9)
    for mode in (capi.MRB_TRACE_WIDE, capi.MRB_TRACE_BINARY_EXACT):
        keys, hits, rout, vis = gpu_scene_cast(sc, rays, mode)
        hit = keys[:, 0] != O.INVALID
        assert hit.mean() > 0.2
        k_inst, prim = keys[hit, 3].astype(np.int64), (keys[hit, 0] & 0x0FFFFFFF).astype(np.int64)
        second = (prim >= half).astype(np.int64)
        expect = np.where(k_inst % 4 == 0, 7 + second, 100 + 2 * k_inst + second)
        assert np.array_equal(keys[hit, 1].astype(np.int64), expect)
        assert len(np.unique(k_inst)) > n_inst // 2
    sc.close(); acc.close()
