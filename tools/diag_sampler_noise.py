"""relMSE of 4096-spp Cornell renders (64x64) against the reference's converged image: the reference's own Sobol / Z-Sobol
images (tests/golden) vs the B200 plugin's Independent / Sobol / Z-Sobol. usage: python tools/diag_sampler_noise.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from mray_b200 import scenes
g = lambda n: np.load(os.path.join(ROOT, "tests", "golden", f"render_{n}.npz"))["img"].astype(np.float32)
rel = lambda a, b: float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))
conv = g("cornell64_spp16384")
out = {"reference_Sobol": rel(g("cornell64_sobol_spp4096"), conv), "reference_ZSobol": rel(g("cornell64_zsobol_spp4096"), conv),
       "floor_of_the_converged_image": 8.3 / 16384}
c = scenes.cornell_box(); b = O.batched_scene(c["positions"], c["indices"], c["material"])
plugin = os.path.join(ROOT, "mray_b200", "lib", "libTracerDLL_B200.so")
for s in ("Independent", "Sobol", "ZSobol"):
    img, w, st = O.driver_render(plugin, b, c["albedo"], 3, c["radiance"], c["camera"], 64, 64, 4096, seed=16, sampler=s)
    out["b200_" + s] = rel(img, conv)
print(json.dumps(out))
