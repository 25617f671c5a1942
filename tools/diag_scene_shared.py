"""Diagnostic (GPU box): two-level scene whose accelerators share one vertex/index array (subset ranges)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle_lib as O
import mray_b200
from mray_b200 import scenes, capi
from test_gpu_render import _rigid
ctx = mray_b200.Context(0)
c = scenes.cornell_box()
b = O.batched_scene(c["positions"], c["indices"], c["material"])
rng = np.random.default_rng(17)
nb = len(b["materials"])
mats34 = np.stack([_rigid(rng) for _ in range(nb)])
pos = b["positions"].astype(np.float64).copy()
gidx = b["indices"].copy()
for k in range(nb):
    lo, hi = int(b["vertex_offsets"][k]), int(b["vertex_offsets"][k + 1])
    inv = np.linalg.inv(np.vstack([mats34[k], [0, 0, 0, 1]]))
    pos[lo:hi] = pos[lo:hi] @ inv[:3, :3].T + inv[:3, 3]
    gidx[b["tri_offsets"][k]:b["tri_offsets"][k + 1]] += lo
pos = np.ascontiguousarray(pos, np.float32)
wpos = np.ascontiguousarray(b["positions"]); 
flat = capi.Accelerator(ctx, wpos, gidx)
accs = [capi.Accelerator(ctx, pos, gidx, prim_ranges=[[int(b["tri_offsets"][k]), int(b["tri_offsets"][k + 1])]], light_or_mat_keys=[k]) for k in range(nb)]
scene = capi.Scene(ctx, [(accs[k], mats34[k]) for k in range(nb)])
cam = c["camera"]
rays = scenes.pinhole_rays(64, 64, eye=cam["eye"], gaze=cam["gaze"], up=cam["up"], fov_y_deg=cam["fov_y_deg"])
n = rays.shape[0]
def run(obj):
    k = np.full((n, 4), 0xFFFFFFFF, np.uint32); h = np.zeros((n, 2), np.float32); r = rays.copy()
    obj.cast_rays(k, h, r)
    return k, h, r
k0, h0, r0 = run(flat); k1, h1, r1 = run(scene)
print("prim equal", (k0[:, 0] == k1[:, 0]).mean(), "t close", np.isclose(r0[:, 7], r1[:, 7], rtol=1e-4).mean())
bad = np.nonzero(k0[:, 0] != k1[:, 0])[0]
print("bad", bad.size, "examples", [(int(k0[i, 0]), int(k1[i, 0]), float(r0[i, 7]), float(r1[i, 7])) for i in bad[:8]])
for k in range(nb):
    m = (k0[:, 0] >= b["tri_offsets"][k]) & (k0[:, 0] < b["tri_offsets"][k + 1])
    print("batch", k, "rays", m.sum(), "mismatch", (k0[m, 0] != k1[m, 0]).sum())
print(scene.export_tlas()["instance_aabb"])
