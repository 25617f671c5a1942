"""GPU (B200): parity of the CUDA path — called through the C-ABI (mray_b200.capi -> libmray_b200.so)
— with the oracle and with golden artefacts of the reference. Bit-exact for Morton codes, sort
permutations, hierarchy, boxes and closest-hit primitive ids; hit distance / barycentrics are held
bit-exact as well (north_star only asks 1e-5 relative — REL_T_TOL — for t)."""
import numpy as np
import pytest

import oracle_lib as O
from helpers import SMALL_CASES, digest, full_hashes, load_small, oracle_trace_mt
from mray_b200 import capi, scenes

pytestmark = pytest.mark.gpu
REL_T_TOL = 1e-5


def _torch():
    import torch
    return torch


def dev(a):
    torch = _torch()
    if a.dtype == np.uint32:
        return torch.from_numpy(a.view(np.int32).copy()).cuda()
    if a.dtype == np.uint64:
        return torch.from_numpy(a.view(np.int64).copy()).cuda()
    return torch.from_numpy(a.copy()).cuda()


def host(t, dtype):
    return t.cpu().numpy().view(dtype)


def gpu_cast(acc, rays_np, mode, indices=None):
    torch = _torch()
    n = rays_np.shape[0]
    rays = dev(rays_np)
    keys = torch.full((n, 4), -1, dtype=torch.int32, device="cuda")
    hits = torch.zeros((n, 2), dtype=torch.float32, device="cuda")
    idx = None if indices is None else dev(indices)
    acc.cast_rays(keys, hits, rays, idx, mode)
    torch.cuda.synchronize()
    return host(keys, np.uint32), hits.cpu().numpy(), rays.cpu().numpy()


def gpu_visibility(acc, rays_np, mode, indices=None):
    torch = _torch()
    n = rays_np.shape[0]
    words = (n + 31) // 32
    bits = torch.full((words,), -1, dtype=torch.int32, device="cuda")
    idx = None if indices is None else dev(indices)
    acc.cast_visibility_rays(bits, dev(rays_np), idx, mode)
    torch.cuda.synchronize()
    w = host(bits, np.uint32)
    return ((w[np.arange(n) // 32] >> (np.arange(n) % 32).astype(np.uint32)) & 1).astype(bool)


# ---------------------------------------------------------------------------------------------
# radix sort (RayPartitioner / LBVH build primitive)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.uint32, np.uint64])
@pytest.mark.parametrize("n", [1, 31, 1111, 3072, 4096, 4097, 100_003, 3_000_000])
def test_radix_sort_full_range(gpu_ctx, dtype, n):
    rng = np.random.default_rng(n)
    if n == 1111:  # Tests/Device/T_AlgRadixSort.cu:L76-117 — shuffled iota
        keys = rng.permutation(n).astype(dtype)
    else:
        keys = rng.integers(0, np.iinfo(dtype).max, size=n, dtype=dtype)
    vals = np.arange(n, dtype=np.uint32)
    ek, ev = keys.copy(), vals.copy()
    (O.lib().orc_radix_sort_u64 if dtype == np.uint64 else O.lib().orc_radix_sort_u32)(ek, ev, n, 0, 8 * keys.itemsize)
    dk, dv = dev(keys), dev(vals)
    gpu_ctx.radix_sort_pairs(dk, dv)
    _torch().cuda.synchronize()
    assert np.array_equal(host(dk, dtype), ek)
    assert np.array_equal(host(dv, np.uint32), ev)


@pytest.mark.parametrize("bits", [(0, 4), (3, 11), (20, 32), (0, 17)])
def test_radix_sort_bit_range_is_stable(gpu_ctx, bits):
    n = 50_000
    rng = np.random.default_rng(1)
    keys = rng.integers(0, 2 ** 32 - 1, size=n, dtype=np.uint32)
    vals = np.arange(n, dtype=np.uint32)
    sub = (keys >> bits[0]) & ((1 << (bits[1] - bits[0])) - 1)
    expect = np.argsort(sub, kind="stable").astype(np.uint32)
    hk, hv = keys.copy(), vals.copy()
    gpu_ctx.radix_sort_pairs(hk, hv, bits[0], bits[1])  # host-memory variant of the C-ABI
    assert np.array_equal(hv, expect)
    assert np.array_equal(hk, keys[expect])


def test_radix_sort_few_distinct_keys_and_empty(gpu_ctx):
    keys = np.zeros(0, np.uint32)
    gpu_ctx.radix_sort_pairs(keys, np.zeros(0, np.uint32))
    rng = np.random.default_rng(3)
    keys = rng.integers(0, 3, size=200_000).astype(np.uint64) << np.uint64(40)
    vals = np.arange(keys.shape[0], dtype=np.uint32)
    expect = np.argsort(keys, kind="stable").astype(np.uint32)
    gpu_ctx.radix_sort_pairs(keys, vals)
    assert np.array_equal(vals, expect)


# ---------------------------------------------------------------------------------------------
# LBVH build
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", SMALL_CASES)
def test_build_matches_reference_golden(gpu_ctx, name):
    g = load_small(name)
    acc = capi.Accelerator(gpu_ctx, g["positions"], g["indices"])
    e = acc.export_lbvh()
    assert acc.info.duplicateCodes == 0
    for key in ["leaf_aabb", "accel_aabb", "morton", "sorted_morton", "sorted_idx", "boxes"]:
        assert np.array_equal(e[key], g[key]), key
    if name != "single":
        assert np.array_equal(e["nodes"], g["nodes"])
        assert np.array_equal(e["leaf_parent"], g["leaf_parent"])
    else:
        assert list(e["nodes"][0]) == [O.LEAF_FLAG, O.INVALID, O.INVALID]
    acc.close()


def test_build_device_pointers_and_prim_ranges(gpu_ctx):
    """Leaf list from two prim ranges of the group (KCGeneratePrimitiveKeys) — device-resident inputs."""
    p, i = scenes.arcade_mesh(6000)
    n = i.shape[0]
    ranges = np.array([[100, 1500], [3000, n - 7]], np.uint32)
    sel = np.concatenate([np.arange(a, b) for a, b in ranges])
    acc = capi.Accelerator(gpu_ctx, dev(p), dev(i), prim_ranges=ranges, light_or_mat_keys=[7, 9], prim_group_id=2)
    e = acc.export_lbvh()
    b = O.oracle_build(p, np.ascontiguousarray(i[sel]))
    for key in ["morton", "sorted_idx", "nodes", "boxes"]:
        assert np.array_equal(e[key], getattr(b, key)), key
    rays = scenes.pinhole_rays(64, 36, **scenes.ARCADE_CAMERA)
    keys, hits, rout = gpu_cast(acc, rays, capi.MRB_TRACE_WIDE)
    prim, t, bary, _ = O.oracle_trace(p, np.ascontiguousarray(i[sel]), b, rays)
    hit = prim != O.INVALID
    assert np.array_equal(keys[:, 0] != O.INVALID, hit)
    # PrimitiveKey = group(4) | index(28) of the ORIGINAL group index list
    assert np.array_equal(keys[hit, 0], (2 << 28) | sel[prim[hit]].astype(np.uint32))
    assert np.array_equal(keys[hit, 1], np.where(prim[hit] < 1400, 7, 9).astype(np.uint32))
    assert np.array_equal(rout[hit, 7], t[hit])
    acc.close()


def test_many_prim_ranges_flattened_surfaces(gpu_ctx):
    """More than 8 prim ranges in one accelerator (identity-transform surfaces flattened): every range
    keeps its own lightOrMatKey / cull flag; hits report the ORIGINAL prim index of the group."""
    p, i = scenes.arcade_mesh(6000)
    n = i.shape[0]
    cuts = np.linspace(0, n, 41).astype(np.uint32)
    ranges = np.stack([cuts[:-1], cuts[1:]], 1)
    lm = (np.arange(40) * 3 + 1).astype(np.uint32)
    acc = capi.Accelerator(gpu_ctx, p, i, prim_ranges=ranges, light_or_mat_keys=lm)
    b = O.oracle_build(p, i)
    assert np.array_equal(acc.export_lbvh()["nodes"], b.nodes)
    rays = scenes.pinhole_rays(96, 54, **scenes.ARCADE_CAMERA)
    for mode in (capi.MRB_TRACE_WIDE, capi.MRB_TRACE_BINARY_EXACT):
        keys, hits, rout = gpu_cast(acc, rays, mode)
        prim, t, bary, _ = O.oracle_trace(p, i, b, rays)
        assert np.array_equal(keys[:, 0], prim)
        hit = prim != O.INVALID
        expect_lm = lm[np.searchsorted(cuts, prim[hit], side="right") - 1]
        assert np.array_equal(keys[hit, 1], expect_lm)
    acc.close()


def test_build_large_soup_matches_oracle(gpu_ctx):
    p, i = scenes.random_soup(1_000_000)
    acc = capi.Accelerator(gpu_ctx, p, i)
    e = acc.export_lbvh()
    b = O.oracle_build(p, i, robust=1)
    for key in ["morton", "sorted_idx", "boxes"]:
        assert np.array_equal(e[key], getattr(b, key)), key
    assert np.array_equal(e["nodes"], b.nodes)
    assert acc.info.wideNodeCount > 0
    acc.close()


def _wide_signature(acc):
    """The wide tree without what depends on allocation order inside a level (child / triangle base offsets, node order)."""
    nodes, tris = acc.export_wide()
    sig = nodes.copy(); sig[:, 4] = 0; sig[:, 5] = 0                 # q1.x = childBase, q1.y = triBase
    sig = sig[np.lexsort(sig.T[::-1])]
    t = tris.view(np.uint32)[:, [3, 7, 11]]                           # leaf, rank, flags | range
    return sig, t[np.lexsort(t.T[::-1])]


@pytest.mark.parametrize("mesh", ["cornell", "arcade20k", "arcade264k", "soup300k", "duplicates"])
def test_group_collapse_builds_the_serial_tree(gpu_ctx, mesh):
    """KCollapseGroups (eight lanes per wide node, the default) against the one-thread-per-node audit kernel: same nodes, same
    triangle records, and the same hits through the wide traversal."""
    if mesh == "cornell":
        c = scenes.cornell_box(); p, i = c["positions"], c["indices"]
    elif mesh == "arcade20k": p, i = scenes.arcade_mesh(20000)
    elif mesh == "arcade264k": p, i = scenes.arcade_mesh()
    elif mesh == "soup300k": p, i = scenes.random_soup(300_000, seed=11)
    else:
        p, i = scenes.arcade_mesh(3000); i = np.ascontiguousarray(np.concatenate([i, i, i]))
    a = capi.Accelerator(gpu_ctx, p, i)
    b = capi.Accelerator(gpu_ctx, p, i, flags=capi.MRB_BUILD_SERIAL_COLLAPSE)
    assert a.info.wideNodeCount == b.info.wideNodeCount > 0
    (na, ta), (nb, tb) = _wide_signature(a), _wide_signature(b)
    assert np.array_equal(na, nb) and np.array_equal(ta, tb)
    rays = scenes.pinhole_rays(160, 90, **scenes.ARCADE_CAMERA)
    ka, _, ra = gpu_cast(a, rays, capi.MRB_TRACE_WIDE)
    kb, _, rb = gpu_cast(b, rays, capi.MRB_TRACE_WIDE)
    assert np.array_equal(ka, kb) and np.array_equal(ra, rb)
    a.close(); b.close()


def test_build_duplicate_codes_reference_delta_audit(gpu_ctx):
    """Repeated codes: REFERENCE_DELTA reproduces the reference's (possibly ill-formed) node list node
    for node; the default build uses the augmented key and still traces correctly."""
    p, i = scenes.arcade_mesh(3000)
    i2 = np.ascontiguousarray(np.concatenate([i, i]))
    acc = capi.Accelerator(gpu_ctx, p, i2, flags=capi.MRB_BUILD_REFERENCE_DELTA)
    assert acc.info.duplicateCodes == 1 and acc.info.wideNodeCount == 0
    assert np.array_equal(acc.export_lbvh()["nodes"], O.oracle_build(p, i2, robust=0).nodes)
    acc.close()
    acc = capi.Accelerator(gpu_ctx, p, i2)
    b = O.oracle_build(p, i2, robust=1)
    assert np.array_equal(acc.export_lbvh()["nodes"], b.nodes)
    rays = scenes.pinhole_rays(64, 36, **scenes.ARCADE_CAMERA)
    keys, _, rout = gpu_cast(acc, rays, capi.MRB_TRACE_WIDE)
    prim, t, _, _ = O.oracle_trace(p, i2, b, rays)
    assert np.array_equal(keys[:, 0], prim)  # exact-t ties between the two copies -> first in Morton order
    assert np.array_equal(rout[:, 7], t)
    acc.close()


# ---------------------------------------------------------------------------------------------
# ray casting
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [capi.MRB_TRACE_WIDE, capi.MRB_TRACE_BINARY_EXACT])
@pytest.mark.parametrize("name", SMALL_CASES)
def test_cast_rays_matches_reference_golden(gpu_ctx, name, mode):
    g = load_small(name)
    acc = capi.Accelerator(gpu_ctx, g["positions"], g["indices"])
    keys, hits, rout = gpu_cast(acc, g["rays"], mode)
    assert np.array_equal(keys[:, 0], g["hit_prim"])                       # closest-hit ids: bit exact
    hit = g["hit_prim"] != O.INVALID
    assert np.allclose(rout[hit, 7], g["hit_t"][hit], rtol=REL_T_TOL, atol=0)  # north_star tolerance
    assert np.array_equal(rout[:, 7], g["hit_t"])                          # ... and in fact bit exact
    assert np.array_equal(hits[hit], g["hit_bary"][hit])
    assert np.all(keys[~hit] == O.INVALID)                                  # misses untouched
    vis = gpu_visibility(acc, g["rays"], mode)
    assert np.array_equal(~vis, g["any_hit"])
    acc.close()


def test_cast_rays_indirect_ragged_and_empty(gpu_ctx):
    g = load_small("arcade")
    acc = capi.Accelerator(gpu_ctx, g["positions"], g["indices"])
    n = g["rays"].shape[0]
    idx = np.random.default_rng(2).permutation(n)[: n // 3].astype(np.uint32)
    keys, hits, rout = gpu_cast(acc, g["rays"], capi.MRB_TRACE_WIDE, idx)
    mask = np.zeros(n, bool); mask[idx] = True
    assert np.array_equal(keys[mask, 0], g["hit_prim"][mask])
    assert np.all(keys[~mask] == O.INVALID)
    assert np.array_equal(rout[~mask], g["rays"][~mask])
    vis = gpu_visibility(acc, g["rays"], capi.MRB_TRACE_WIDE, idx)
    assert np.array_equal(~vis[mask], g["any_hit"][mask]) and np.all(vis[~mask])
    # empty batch is a no-op
    k0, _, _ = gpu_cast(acc, g["rays"], capi.MRB_TRACE_WIDE, np.zeros(0, np.uint32))
    assert np.all(k0 == O.INVALID)
    acc.close()


def test_cast_rays_host_buffers_and_culling(gpu_ctx):
    """Host-pointer variant of the C-ABI (the e2e path) + back-face culling flag."""
    c = scenes.cornell_box()
    p, i = c["positions"], c["indices"]
    rays = scenes.pinhole_rays(48, 48, **c["camera"])
    for cull in (0, 1):
        acc = capi.Accelerator(gpu_ctx, p, i, prim_ranges=[[0, i.shape[0]]], cull_backface=[cull])
        b = O.oracle_build(p, i)
        keys = np.full((rays.shape[0], 4), 0xFFFFFFFF, np.uint32)
        hits = np.zeros((rays.shape[0], 2), np.float32)
        r = rays.copy()
        acc.cast_rays(keys, hits, r, None, capi.MRB_TRACE_WIDE)
        prim, t, bary, _ = O.oracle_trace(p, i, b, rays, cull=cull)
        assert np.array_equal(keys[:, 0], prim) and np.array_equal(r[:, 7], t)
        acc.close()


def test_full_size_config2_parity(gpu_ctx):
    """BASELINE config 2 at full size: 264 K triangles, 1920x1080 primary rays, AO closest + any hit."""
    h = full_hashes()
    p, i = scenes.arcade_mesh()
    acc = capi.Accelerator(gpu_ctx, p, i)
    e = acc.export_lbvh()
    b = O.oracle_build(p, i)
    for key in ["morton", "sorted_idx", "nodes", "boxes"]:
        assert np.array_equal(e[key], getattr(b, key)), key
    same_scene = digest(p) == h["positions"] and digest(i) == h["indices"]
    if same_scene:  # digests of the reference's own outputs
        assert digest(e["morton"]) == h["morton"] and digest(e["sorted_idx"]) == h["sorted_idx"]
        assert digest(e["nodes"]) == h["nodes"] and digest(e["boxes"]) == h["boxes"]
    rays = scenes.pinhole_rays(1920, 1080, **scenes.ARCADE_CAMERA)
    keys, hits, rout = gpu_cast(acc, rays, capi.MRB_TRACE_WIDE)
    prim, t, bary, _ = oracle_trace_mt(p, i, b, rays)
    assert np.array_equal(keys[:, 0], prim)
    assert np.array_equal(rout[:, 7], t)
    hit = prim != O.INVALID
    assert np.array_equal(hits[hit], bary[hit])
    if same_scene:
        assert digest(keys[:, 0].copy()) == h["primary_prim"] and digest(rout[:, 7].copy()) == h["primary_t"]
    ao = scenes.ao_rays(rays, prim, t, p, i, 0.15 * h["scene_diameter"])
    akeys, _, aout = gpu_cast(acc, ao, capi.MRB_TRACE_WIDE)
    aprim, at, _, _ = oracle_trace_mt(p, i, b, ao)
    assert np.array_equal(akeys[:, 0], aprim) and np.array_equal(aout[:, 7], at)
    vis = gpu_visibility(acc, ao, capi.MRB_TRACE_WIDE)
    vprim, _, _, _ = oracle_trace_mt(p, i, b, ao, mode=1)
    assert np.array_equal(~vis, vprim != O.INVALID)
    if same_scene:
        assert digest(akeys[:, 0].copy()) == h["ao_prim"]
        assert digest((~vis).astype(np.uint8)) == h["ao_any"]
    # size-independent property: the audit path (reference algorithm on the GPU) agrees as well
    bkeys, _, bout = gpu_cast(acc, rays, capi.MRB_TRACE_BINARY_EXACT)
    assert np.array_equal(bkeys[:, 0], prim) and np.array_equal(bout[:, 7], t)
    acc.close()


def test_host_pointer_casts_are_pipelined_and_identical(gpu_ctx):
    """Host-buffer casts large enough to take the chunked upload / trace / download pipeline return exactly
    what the device-pointer call returns (ragged last chunk, pre-set miss values untouched, visibility words)."""
    p, i = scenes.arcade_mesh(60_000)
    acc = capi.Accelerator(gpu_ctx, p, i)
    rays = scenes.pinhole_rays(801, 517, **scenes.ARCADE_CAMERA)        # 414 117 rays: not a multiple of anything
    n = rays.shape[0]
    rays[::7, 7] = 0.5                                                   # short rays: plenty of misses
    dkeys, dhits, drays = gpu_cast(acc, rays, capi.MRB_TRACE_WIDE)
    hkeys = np.full((n, 4), 0xFFFFFFFF, np.uint32); hhits = np.zeros((n, 2), np.float32); hrays = rays.copy()
    acc.cast_rays(hkeys, hhits, hrays, None, capi.MRB_TRACE_WIDE)
    assert np.array_equal(hkeys, dkeys) and np.array_equal(hhits, dhits) and np.array_equal(hrays, drays)
    assert (hkeys[:, 0] == 0xFFFFFFFF).any() and (hkeys[:, 0] != 0xFFFFFFFF).any()
    dvis = gpu_visibility(acc, rays, capi.MRB_TRACE_WIDE)
    bits = np.full(((n + 31) // 32,), 0xFFFFFFFF, np.uint32)
    acc.cast_visibility_rays(bits, rays.copy(), None, capi.MRB_TRACE_WIDE)
    hvis = ((bits[np.arange(n) // 32] >> (np.arange(n) % 32).astype(np.uint32)) & 1).astype(bool)
    assert np.array_equal(hvis, dvis)
    # MRB_TRACE_FRESH_OUTPUTS: whatever the output buffers held is ignored; misses come back INVALID / zero / visible
    fkeys = np.full((n, 4), 0x12345678, np.uint32); fhits = np.full((n, 2), 7.0, np.float32); frays = rays.copy()
    acc.cast_rays(fkeys, fhits, frays, None, capi.MRB_TRACE_WIDE | capi.MRB_TRACE_FRESH_OUTPUTS)
    assert np.array_equal(fkeys, dkeys) and np.array_equal(fhits, dhits) and np.array_equal(frays, drays)
    fbits = np.zeros(((n + 31) // 32,), np.uint32)
    acc.cast_visibility_rays(fbits, rays.copy(), None, capi.MRB_TRACE_WIDE | capi.MRB_TRACE_FRESH_OUTPUTS)
    tail = (1 << (n % 32)) - 1 if n % 32 else 0xFFFFFFFF
    assert np.array_equal(fbits[:-1], bits[:-1]) and (fbits[-1] & tail) == (bits[-1] & tail)
    acc.close()


# ---------------------------------------------------------------------------------------------
# RayPartitioner
# ---------------------------------------------------------------------------------------------
def test_multi_partition_like_reference_test(gpu_ctx):
    """Tests/Tracer/T_RayPartitioner.cu:L12-188 (SimulateBasicPathTracer): 500 000 rays, 16 batches x 256 data
    values, seed 333 — every partition non-empty, all rays visited, order = stable sort by (batch, data)."""
    import random
    rng = random.Random(333)
    n, batch_bits, data_bits = 500_000, 4, 8
    keys = np.array([(rng.randrange(16) << data_bits) | rng.randrange(256) for _ in range(n)], np.uint32)
    idx = np.arange(n, dtype=np.uint32)
    k, i = keys.copy(), idx.copy()
    count, ofs, pk = gpu_ctx.multi_partition(k, i, (0, data_bits), (data_bits, data_bits + batch_bits), 64)
    expect = np.argsort(keys, kind="stable").astype(np.uint32)
    assert np.array_equal(i, expect) and np.array_equal(k, keys[expect])        # permutation: bit exact
    assert count == 16 and ofs[0] == 0 and ofs[-1] == n
    assert np.all(np.diff(ofs.astype(np.int64)) > 0)                             # every partition non-empty
    assert np.array_equal(pk >> data_bits, np.arange(16, dtype=np.uint32))
    visited = np.zeros(n, bool)
    for p in range(count):
        part = i[ofs[p]:ofs[p + 1]]
        assert np.all((keys[part] >> data_bits) == (pk[p] >> data_bits))
        visited[part] = True
    assert visited.all()


def test_multi_partition_noncontiguous_ranges_and_batch_only(gpu_ctx):
    rng = np.random.default_rng(4)
    n = 70_000
    data = rng.integers(0, 32, n).astype(np.uint32); batch = rng.choice([1, 5, 6, 200], n).astype(np.uint32)
    keys = (batch << 20) | (data << 3) | rng.integers(0, 8, n).astype(np.uint32)  # junk in bits 0..2 and 8..19
    k, i = keys.copy(), np.arange(n, dtype=np.uint32)
    count, ofs, pk = gpu_ctx.multi_partition(k, i, (3, 8), (20, 28), 16)
    expect = np.lexsort((np.arange(n), data, batch)).astype(np.uint32)
    assert np.array_equal(i, expect)
    assert count == 4 and list(pk >> 20) == [1, 5, 6, 200]
    k, i = keys.copy(), np.arange(n, dtype=np.uint32)
    count, ofs, pk = gpu_ctx.multi_partition(k, i, (3, 8), (20, 28), 16, only_batches=True)
    assert np.array_equal(i, np.argsort(batch, kind="stable").astype(np.uint32))
    # empty input
    count, ofs, pk = gpu_ctx.multi_partition(np.zeros(0, np.uint32), np.zeros(0, np.uint32), (0, 4), (4, 8), 4)
    assert count == 0 and ofs[0] == 0


def test_binary_partition_is_stable(gpu_ctx):
    rng = np.random.default_rng(8)
    slots = 300_000
    flags = (rng.random(slots) < 0.3).astype(np.uint8)
    indices = rng.permutation(slots)[:200_000].astype(np.uint32)
    out, left = gpu_ctx.binary_partition(indices, flags)
    alive = flags[indices] != 0
    assert left == int(alive.sum())
    assert np.array_equal(out[:left], indices[alive]) and np.array_equal(out[left:], indices[~alive])
