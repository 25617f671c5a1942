"""Generates tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libref_taps.so, i.e. the
unmodified reference kernels running on the reference's CPU device backend). Run in the authoring
container only (needs /root/reference to have been built by oracle/ref_build/build_ref.sh):

    python oracle/gen_golden.py

Fixtures:
  lbvh_small_*.npz : full artefacts (Morton codes, sort permutation, LBVHNode list, node boxes, closest
                     / any hit results on a fixed ray batch) for small meshes — compared array by array.
  lbvh_full_hashes.json : sha256 of the same artefacts for the full-size config-2 scene (264 K triangles,
                     1920x1080 primary rays + AO rays) — compared by digest on the GPU box, where the
                     reference does not exist.
TEST INFRASTRUCTURE ONLY.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
from mray_b200 import scenes  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def small_case(name, positions, indices, rays):
    b = O.ref_build(positions, indices)
    prim, t, bary, back = O.ref_trace(positions, indices, b, rays, mode=0)
    aprim, _, _, _ = O.ref_trace(positions, indices, b, rays, mode=1)
    np.savez_compressed(os.path.join(GOLD, f"lbvh_small_{name}.npz"),
                        positions=positions, indices=indices, rays=rays,
                        leaf_aabb=b.leaf_aabb, accel_aabb=b.accel_aabb, morton=b.morton,
                        sorted_morton=b.sorted_morton, sorted_idx=b.sorted_idx, nodes=b.nodes,
                        leaf_parent=b.leaf_parent, boxes=b.boxes,
                        hit_prim=prim, hit_t=t, hit_bary=bary, hit_back=back,
                        any_hit=(aprim != O.INVALID))
    print(name, "tris", indices.shape[0], "rays", rays.shape[0], "hit", float((prim != O.INVALID).mean()))


def main():
    os.makedirs(GOLD, exist_ok=True)
    # small meshes
    p, i = scenes.arcade_mesh(3000)
    small_case("arcade", p, i, scenes.pinhole_rays(96, 54, **scenes.ARCADE_CAMERA))
    c = scenes.cornell_box()
    small_case("cornell", c["positions"], c["indices"], scenes.pinhole_rays(64, 64, **c["camera"]))
    p, i = scenes.random_soup(1500, seed=7, size=0.08)
    rng = np.random.default_rng(11)
    org = rng.random((4000, 3)).astype(np.float32) * 1.4 - 0.2
    d = rng.normal(size=(4000, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([org, np.full((4000, 1), 1e-4, np.float32), d.astype(np.float32),
                           np.full((4000, 1), 0.9, np.float32)], 1).astype(np.float32)
    small_case("soup", p, i, np.ascontiguousarray(rays))
    p1, i1 = scenes.random_soup(1, seed=3, size=0.5)
    small_case("single", p1, i1, np.ascontiguousarray(rays[:256]))

    # full-size config 2
    p, i = scenes.arcade_mesh()
    b = O.ref_build(p, i)
    rays = scenes.pinhole_rays(1920, 1080, **scenes.ARCADE_CAMERA)
    prim, t, bary, back = O.ref_trace(p, i, b, rays, mode=0)
    diam = float(np.linalg.norm(b.accel_aabb[3:] - b.accel_aabb[:3]))
    ao = scenes.ao_rays(rays, prim, t, p, i, 0.15 * diam)
    aprim, at, abary, _ = O.ref_trace(p, i, b, ao, mode=0)
    vprim, _, _, _ = O.ref_trace(p, i, b, ao, mode=1)
    hashes = dict(
        scene="arcade_mesh(target_tris=262144, seed=1234)", triangles=int(i.shape[0]), vertices=int(p.shape[0]),
        positions=digest(p), indices=digest(i),
        morton=digest(b.morton), sorted_idx=digest(b.sorted_idx), nodes=digest(b.nodes), boxes=digest(b.boxes),
        primary_rays=digest(rays), primary_prim=digest(prim), primary_t=digest(t), primary_bary=digest(bary),
        primary_hit_fraction=float((prim != O.INVALID).mean()),
        ao_rays=digest(ao), ao_prim=digest(aprim), ao_t=digest(at),
        ao_any=digest((vprim != O.INVALID).astype(np.uint8)),
        ao_occluded_fraction=float((vprim != O.INVALID).mean()),
        scene_diameter=diam,
    )
    with open(os.path.join(GOLD, "lbvh_full_hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1)
    print(json.dumps(hashes, indent=1))


if __name__ == "__main__":
    main()
