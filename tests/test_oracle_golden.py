"""CPU: pins the oracle (oracle/mray_oracle.c) against
  * the known-answer vectors of the reference's own tests, and
  * golden artefacts produced by the reference itself (tests/golden, made by oracle/gen_golden.py).
"""
import random

import numpy as np
import pytest

import oracle_lib as O
from helpers import SMALL_CASES, digest, full_hashes, load_small, oracle_trace_mt
from mray_b200 import scenes


def test_morton_known_answers():
    # Tests/Core/T_GraphicsFunctions.cpp:L182-200 (32 bit) and L263-283 (64 bit)
    L = O.lib()
    assert L.orc_morton_compose32(0b1111111111, 0, 0) == 0b001001001001001001001001001001
    assert L.orc_morton_compose32(0, 0b1111111111, 0) == 0b010010010010010010010010010010
    assert L.orc_morton_compose32(0, 0, 0b1111111111) == 0b100100100100100100100100100100
    assert L.orc_morton_compose64(0xFFFFF, 0, 0) == 0x249249249249249
    assert L.orc_morton_compose64(0, 0xFFFFF, 0) == 0x492492492492492
    assert L.orc_morton_compose64(0, 0, 0xFFFFF) == 0x924924924924924
    # 21-bit inputs use all 63 bits
    assert L.orc_morton_compose64(0x1FFFFF, 0x1FFFFF, 0x1FFFFF) == 0x7FFFFFFFFFFFFFFF


@pytest.mark.parametrize("dtype", [np.uint32, np.uint64])
def test_radix_sort_like_reference_test(dtype):
    # Tests/Device/T_AlgRadixSort.cu:L76-117: 1111 shuffled iota keys, values follow keys, result = iota
    n = 1111
    keys = np.arange(n, dtype=dtype)
    rng = random.Random(123)
    perm = list(range(n))
    rng.shuffle(perm)
    keys = keys[perm].copy()
    vals = keys.astype(np.uint32).copy()
    fn = O.lib().orc_radix_sort_u64 if dtype == np.uint64 else O.lib().orc_radix_sort_u32
    fn(keys, vals, n, 0, 8 * keys.itemsize)
    assert np.array_equal(keys, np.arange(n, dtype=dtype))
    assert np.array_equal(vals, np.arange(n, dtype=np.uint32))


def test_radix_sort_is_stable():
    rng = np.random.default_rng(5)
    keys = rng.integers(0, 16, size=5000).astype(np.uint64)
    vals = np.arange(5000, dtype=np.uint32)
    ref = np.argsort(keys, kind="stable").astype(np.uint32)
    O.lib().orc_radix_sort_u64(keys, vals, 5000, 0, 64)
    assert np.array_equal(vals, ref)


@pytest.mark.parametrize("name", SMALL_CASES)
def test_build_matches_reference_golden(name):
    g = load_small(name)
    b = O.oracle_build(g["positions"], g["indices"])
    for key in ["leaf_aabb", "accel_aabb", "morton", "sorted_morton", "sorted_idx", "boxes"]:
        assert np.array_equal(getattr(b, key), g[key]), key
    if name == "single":
        # single leaf: one node {leaf 0, invalid, invalid} (AcceleratorLBVH.cu:L219-226)
        assert b.nodes[0, 0] == O.LEAF_FLAG and b.nodes[0, 1] == O.INVALID and b.nodes[0, 2] == O.INVALID
    else:
        assert np.array_equal(b.nodes, g["nodes"])
        assert np.array_equal(b.leaf_parent, g["leaf_parent"])


@pytest.mark.parametrize("name", SMALL_CASES)
def test_trace_matches_reference_golden(name):
    g = load_small(name)
    b = O.oracle_build(g["positions"], g["indices"])
    prim, t, bary, back = O.oracle_trace(g["positions"], g["indices"], b, g["rays"], mode=0)
    assert np.array_equal(prim, g["hit_prim"])
    assert np.array_equal(t, g["hit_t"])            # bit exact, not merely 1e-5
    assert np.array_equal(bary, g["hit_bary"])
    assert np.array_equal(back, g["hit_back"])
    aprim, _, _, _ = O.oracle_trace(g["positions"], g["indices"], b, g["rays"], mode=1)
    assert np.array_equal(aprim != O.INVALID, g["any_hit"])


@pytest.mark.parametrize("name", ["arcade", "cornell", "soup"])
def test_rank_tiebreak_rule_equals_left_first_traversal(name):
    """The topology-independent rule the CUDA path uses (min over (t, Morton rank)) gives the
    reference's left-first answer."""
    g = load_small(name)
    b = O.oracle_build(g["positions"], g["indices"])
    rank = np.empty(b.n, np.uint32)
    rank[b.sorted_idx] = np.arange(b.n, dtype=np.uint32)
    rays = np.ascontiguousarray(g["rays"][:1500])
    prim, t = O.oracle_brute(g["positions"], g["indices"], rays, rank)
    assert np.array_equal(prim, g["hit_prim"][:1500])
    assert np.array_equal(t, g["hit_t"][:1500])


def test_full_size_scene_against_reference_digests():
    """Config 2 at full size (264 K triangles, 1080p primary + AO batch): the oracle reproduces the
    digests of the reference's outputs."""
    h = full_hashes()
    p, i = scenes.arcade_mesh()
    if digest(p) != h["positions"] or digest(i) != h["indices"]:
        pytest.skip("procedural mesh differs on this host's libm; digests not comparable")
    b = O.oracle_build(p, i)
    assert digest(b.morton) == h["morton"]
    assert digest(b.sorted_idx) == h["sorted_idx"]
    assert digest(b.nodes) == h["nodes"]
    assert digest(b.boxes) == h["boxes"]
    assert len(np.unique(b.morton)) == b.n  # parity precondition: distinct codes
    rays = scenes.pinhole_rays(1920, 1080, **scenes.ARCADE_CAMERA)
    assert digest(rays) == h["primary_rays"]
    prim, t, bary, _ = oracle_trace_mt(p, i, b, rays)
    assert digest(prim) == h["primary_prim"]
    assert digest(t) == h["primary_t"]
    assert digest(bary) == h["primary_bary"]
    ao = scenes.ao_rays(rays, prim, t, p, i, 0.15 * h["scene_diameter"])
    assert digest(ao) == h["ao_rays"]
    aprim, at, _, _ = oracle_trace_mt(p, i, b, ao)
    assert digest(aprim) == h["ao_prim"] and digest(at) == h["ao_t"]
    vprim, _, _, _ = oracle_trace_mt(p, i, b, ao, mode=1)
    assert digest((vprim != O.INVALID).astype(np.uint8)) == h["ao_any"]


def test_reference_delta_vs_robust_delta_on_duplicates():
    """With repeated Morton codes the reference's Delta (AcceleratorLBVH.cu:L85-90) and Karras'
    augmented key differ; with distinct codes they are the same function."""
    p, i = scenes.arcade_mesh(3000)
    a = O.oracle_build(p, i, robust=0)
    b = O.oracle_build(p, i, robust=1)
    assert np.array_equal(a.nodes, b.nodes)
    # duplicate every triangle -> every code appears twice
    i2 = np.ascontiguousarray(np.concatenate([i, i]))
    c = O.oracle_build(p, i2, robust=1)
    # robust tree is a valid binary tree: every node has exactly one parent, root has none
    parents = c.nodes[:, 2]
    assert parents[0] == O.INVALID and np.all(parents[1:] < c.nodes.shape[0])
    kids = c.nodes[:, :2].ravel()
    inner = kids[(kids & O.LEAF_FLAG) == 0]
    assert len(np.unique(inner)) == c.nodes.shape[0] - 1
