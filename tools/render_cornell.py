"""GPU box: render the Cornell box (config 1) through the C-ABI renderer; save PNG + npy to gpurun_out/."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mray_b200
from mray_b200 import capi, scenes

res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
mode = sys.argv[3] if len(sys.argv) > 3 else "WithNEEAndMIS"
ctx = mray_b200.Context(0); ctx.set_stream(torch.cuda.current_stream())
c = scenes.cornell_box()
# sort triangles by material so each material is one prim range (<= 8 ranges per accelerator)
order = np.argsort(c["material"], kind="stable")
idx = np.ascontiguousarray(c["indices"][order]); mat = c["material"][order]
ranges, keys = [], []
for m in np.unique(mat):
    w = np.nonzero(mat == m)[0]
    ranges.append([w[0], w[-1] + 1])
    keys.append(capi.light_key(0) if m == 3 else int(m))
acc = capi.Accelerator(ctx, c["positions"], idx, prim_ranges=ranges, light_or_mat_keys=keys)
r = capi.Renderer(ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"], res, res, spp,
                  sample_mode=mode, rr_range=(2, 20), seed=0)
t0 = time.time()
img, st = r.render()
dt = time.time() - t0
print(f"{res}x{res} {spp}spp {mode}: {dt*1e3:.1f} ms wall, paths {st.pathsCompleted}, closest rays {st.closestRays}, shadow rays {st.shadowRays}, "
      f"iterations {st.iterations}, mean {img.mean(axis=(0,1))}, Mrays/s(wall) {(st.closestRays+st.shadowRays)/dt/1e6:.1f}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", f"cornell_{res}_{spp}_{mode}.npy"), img.astype(np.float16))
try:
    from PIL import Image
    ldr = np.clip((img / (1 + img)) ** (1 / 2.2), 0, 1)[::-1]
    Image.fromarray((ldr * 255).astype(np.uint8)).save(os.path.join(ROOT, "gpurun_out", f"cornell_{res}_{spp}_{mode}.png"))
except Exception as e:
    print("no PNG:", e)
