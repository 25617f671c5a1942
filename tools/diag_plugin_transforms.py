"""Diagnostic (GPU box): Cornell through the TracerI plugin with per-batch (T)Single transforms; dumps images."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from mray_b200 import scenes
from test_gpu_render import _rigid
PLUGIN = os.path.join(ROOT, "mray_b200", "lib", "libTracerDLL_B200.so")
c = scenes.cornell_box()
out = {}
for variant in ("identity", "translate", "rigid", "rigid_nonormals"):
    b = O.batched_scene(c["positions"], c["indices"], c["material"])
    rng = np.random.default_rng(17)
    nb = len(b["materials"])
    if variant == "identity": mats34 = np.stack([np.eye(4)[:3] for _ in range(nb)])
    elif variant == "translate": mats34 = np.stack([np.concatenate([np.eye(3), rng.uniform(-3, 3, size=(3, 1))], axis=1) for _ in range(nb)])
    else: mats34 = np.stack([_rigid(rng) for _ in range(nb)])
    pos = b["positions"].astype(np.float64).copy()
    for k in range(nb):
        lo, hi = int(b["vertex_offsets"][k]), int(b["vertex_offsets"][k + 1])
        inv = np.linalg.inv(np.vstack([mats34[k], [0, 0, 0, 1]]))
        pos[lo:hi] = pos[lo:hi] @ inv[:3, :3].T + inv[:3, 3]
        n = b["normals"][lo:hi].astype(np.float64) @ mats34[k][:, :3]
        b["normals"][lo:hi] = (n / np.linalg.norm(n, axis=1, keepdims=True)).astype(np.float32)
    if variant == "rigid_nonormals": b["normals"][:] = 0
    b["positions"] = np.ascontiguousarray(pos, np.float32)
    img, w, st = O.driver_render(PLUGIN, b, c["albedo"], 3, c["radiance"], c["camera"], 32, 32, 1024,
                                 sample_mode="WithNEEAndMIS", rr_range=(2, 20), seed=2, batch_transforms=mats34)
    print(variant, "mean", img.mean(axis=(0, 1)), "aabb", st["aabb"])
    out[variant] = img
np.savez(os.path.join(ROOT, "gpurun_out", "plugin_transforms.npz"), **out)
