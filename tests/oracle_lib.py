"""ctypes bindings of the CPU oracle (oracle/liboracle.so) and, when it was built in this
container, of the reference taps (oracle/_ref/libref_taps.so). TEST INFRASTRUCTURE ONLY —
imported by tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke(), never by mray_b200."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
INVALID = 0xFFFFFFFF
LEAF_FLAG = 0x80000000

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build_oracle():
    src = os.path.join(ORACLE_DIR, "mray_oracle.c")
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        L.orc_morton_compose64.restype = C.c_uint64
        L.orc_morton_compose64.argtypes = [C.c_uint32] * 3
        L.orc_morton_compose32.restype = C.c_uint32
        L.orc_morton_compose32.argtypes = [C.c_uint32] * 3
        L.orc_tri_aabb_center.argtypes = [_f32p, _u32p, C.c_uint32, _f32p, _f32p]
        L.orc_aabb_union.argtypes = [_f32p, C.c_uint32, _f32p]
        L.orc_morton63.argtypes = [_f32p, C.c_uint32, _f32p, _u64p]
        L.orc_radix_sort_u64.argtypes = [_u64p, _u32p, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_radix_sort_u32.argtypes = [_u32p, _u32p, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_karras.argtypes = [_u64p, _u32p, C.c_uint32, _u32p, _u32p, C.c_int]
        L.orc_union_boxes.argtypes = [_u32p, _u32p, _f32p, C.c_uint32, _f32p]
        L.orc_lbvh_build.argtypes = [_f32p, _u32p, C.c_uint32, C.c_int, _f32p, _f32p, _u64p, _u64p, _u32p,
                                     _u32p, _u32p, _f32p]
        L.orc_lbvh_trace.argtypes = [_f32p, _u32p, _u32p, _f32p, _f32p, C.c_uint32, C.c_int, C.c_int,
                                     _u32p, _f32p, _f32p, _u8p]
        L.orc_brute_trace.argtypes = [_f32p, _u32p, C.c_uint32, C.c_void_p, _f32p, C.c_uint32, C.c_int,
                                      _u32p, _f32p]
        _lib = L
    return _lib


class LBVH:
    """Plain container of the binary LBVH artefacts of ONE accelerator (reference layout:
    LBVHNode {left,right,parent} u32x3, LBVHBoundingBox {min[3],max[3]} — AcceleratorLBVH.h:L62-77)."""

    def __init__(self, n):
        nn = max(1, n - 1)
        self.n = n
        self.leaf_aabb = np.zeros((n, 6), np.float32)
        self.accel_aabb = np.zeros(6, np.float32)
        self.morton = np.zeros(n, np.uint64)
        self.sorted_morton = np.zeros(n, np.uint64)
        self.sorted_idx = np.zeros(n, np.uint32)
        self.nodes = np.zeros((nn, 3), np.uint32)
        self.leaf_parent = np.zeros(n, np.uint32)
        self.boxes = np.zeros((nn, 6), np.float32)


def oracle_build(positions, indices, robust=0) -> LBVH:
    n = indices.shape[0]
    b = LBVH(n)
    lib().orc_lbvh_build(positions, indices, n, robust, b.leaf_aabb, b.accel_aabb, b.morton,
                         b.sorted_morton, b.sorted_idx, b.nodes, b.leaf_parent, b.boxes)
    return b


def oracle_trace(positions, indices, bvh: LBVH, rays, mode=0, cull=0):
    n = rays.shape[0]
    prim = np.zeros(n, np.uint32)
    t = np.zeros(n, np.float32)
    bary = np.zeros((n, 2), np.float32)
    back = np.zeros(n, np.uint8)
    lib().orc_lbvh_trace(positions, indices, bvh.nodes, bvh.boxes, np.ascontiguousarray(rays), n, mode, cull,
                         prim, t, bary, back)
    return prim, t, bary, back


def oracle_brute(positions, indices, rays, rank=None, cull=0):
    n = rays.shape[0]
    prim = np.zeros(n, np.uint32)
    t = np.zeros(n, np.float32)
    rp = None if rank is None else rank.ctypes.data_as(C.c_void_p)
    lib().orc_brute_trace(positions, indices, indices.shape[0], rp, np.ascontiguousarray(rays), n, cull, prim, t)
    return prim, t


# ------------------------------------------------------------------------------------------------
# reference taps (only where oracle/_ref was built, i.e. in the authoring container)
# ------------------------------------------------------------------------------------------------
_ref = None


def ref_available():
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libref_taps.so"))


def ref():
    global _ref
    if _ref is None:
        L = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libref_taps.so"))
        L.ref_morton_compose64.restype = C.c_uint64
        L.ref_morton_compose64.argtypes = [C.c_uint32] * 3
        L.ref_morton_compose32.restype = C.c_uint32
        L.ref_morton_compose32.argtypes = [C.c_uint32] * 3
        L.ref_tri_aabb_center.argtypes = [_f32p, C.c_uint32, _u32p, C.c_uint32, _f32p, _f32p]
        L.ref_lbvh_build.argtypes = [_f32p, _f32p, _u32p, C.c_uint32, _f32p, _u64p, _u64p, _u32p, _u32p, _u32p, _f32p]
        L.ref_lbvh_trace.argtypes = [_f32p, C.c_uint32, _u32p, C.c_uint32, _u32p, _f32p, C.c_uint32,
                                     _f32p, C.c_uint32, C.c_int, C.c_int, _u32p, _f32p, _f32p, _u8p]
        L.ref_linear_trace.argtypes = [_f32p, C.c_uint32, _u32p, C.c_uint32, _f32p, C.c_uint32, C.c_int, _u32p, _f32p]
        _ref = L
    return _ref


def ref_build(positions, indices) -> LBVH:
    n = indices.shape[0]
    b = LBVH(n)
    centers = np.zeros((n, 3), np.float32)
    ref().ref_tri_aabb_center(positions, positions.shape[0], indices, n, b.leaf_aabb, centers)
    ranges = np.array([0, n], np.uint32)
    accel = np.zeros((1, 6), np.float32)
    ref().ref_lbvh_build(b.leaf_aabb, centers, ranges, 1, accel, b.morton, b.sorted_morton, b.sorted_idx,
                         b.nodes, b.leaf_parent, b.boxes)
    b.accel_aabb = accel[0]
    b.centers = centers
    return b


def ref_trace(positions, indices, bvh: LBVH, rays, mode=0, cull=0):
    n = rays.shape[0]
    prim = np.zeros(n, np.uint32)
    t = np.zeros(n, np.float32)
    bary = np.zeros((n, 2), np.float32)
    back = np.zeros(n, np.uint8)
    ref().ref_lbvh_trace(positions, positions.shape[0], indices, indices.shape[0], bvh.nodes, bvh.boxes,
                         bvh.nodes.shape[0], np.ascontiguousarray(rays), n, mode, cull, prim, t, bary, back)
    return prim, t, bary, back
