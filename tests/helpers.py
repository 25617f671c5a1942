"""Shared helpers of the parity tests (test infrastructure)."""
import hashlib
import json
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

import oracle_lib as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_small(name):
    return dict(np.load(os.path.join(GOLD, f"lbvh_small_{name}.npz")))


def full_hashes():
    with open(os.path.join(GOLD, "lbvh_full_hashes.json")) as f:
        return json.load(f)


def oracle_trace_mt(positions, indices, bvh, rays, mode=0, cull=0, threads=None):
    """oracle_trace split over host threads (ctypes releases the GIL)."""
    threads = threads or min(16, os.cpu_count() or 1)
    n = rays.shape[0]
    bounds = np.linspace(0, n, threads + 1).astype(np.int64)
    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(lambda k: O.oracle_trace(positions, indices, bvh, np.ascontiguousarray(rays[bounds[k]:bounds[k + 1]]), mode, cull),
                            range(threads)))
    return tuple(np.concatenate([p[j] for p in parts]) for j in range(4))


SMALL_CASES = ["arcade", "cornell", "soup", "single"]
