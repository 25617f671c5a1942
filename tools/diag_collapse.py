"""Diagnostic (GPU box): the wide-tree collapse with eight lanes per node (default) against the one-thread-per-node audit kernel
(MRB_COLLAPSE_SERIAL=1, read once per process: run this script twice). Prints per mesh: wide node count, build ms (best of 5),
and a signature of the tree that does not depend on allocation order (sorted multiset of node contents without the child /
triangle base offsets; sorted (leaf, rank, flags) of the triangle records)."""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mray_b200
from mray_b200 import capi, scenes
ctx = mray_b200.Context(0); ctx.set_stream(torch.cuda.current_stream())
out = {"serial": os.environ.get("MRB_COLLAPSE_SERIAL", "0")}
meshes = {"arcade264k": scenes.arcade_mesh(), "arcade20k": scenes.arcade_mesh(20000), "soup1m": scenes.random_soup(1_000_000, seed=3), "cornell": None}
c = scenes.cornell_box(); meshes["cornell"] = (c["positions"], c["indices"])
for name, (p, i) in meshes.items():
    dp, di = torch.from_numpy(np.ascontiguousarray(p)).cuda(), torch.from_numpy(np.ascontiguousarray(i).view(np.int32)).cuda()
    best = 1e30
    for _ in range(5):
        a = capi.Accelerator(ctx, dp, di); best = min(best, a.info.buildMs)
        nodes, tris = a.export_wide(); wide = int(a.info.wideNodeCount); a.close()
    sig = nodes.copy(); sig[:, 4] = 0; sig[:, 5] = 0                       # q1.x = childBase, q1.y = triBase
    sig = sig[np.lexsort(sig.T[::-1])]
    t = tris.view(np.uint32)[:, [3, 7, 11]]; t = t[np.lexsort(t.T[::-1])]
    out[name] = {"wide_nodes": wide, "build_ms": round(best, 4), "node_sig": hashlib.sha256(sig.tobytes()).hexdigest()[:16],
                 "tri_sig": hashlib.sha256(t.tobytes()).hexdigest()[:16]}
print(json.dumps(out))
