"""Diagnostic (GPU box): config-2 cast timings for library variants under mray_b200/lib/variants/."""
import os, sys, glob, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1:
    sys.path.insert(0, ROOT)
    from mray_b200 import capi
    capi.LIB_PATH = sys.argv[1]
    sys.argv = [sys.argv[0]]
    exec(open(os.path.join(ROOT, "tools", "diag_timing.py")).read().split("for sz in")[0])
else:
    for lib in [os.path.join(ROOT, "mray_b200", "lib", "libmray_b200.so")] + sorted(glob.glob(os.path.join(ROOT, "mray_b200", "lib", "variants", "*.so"))):
        out = subprocess.run([sys.executable, __file__, lib], capture_output=True, text=True).stdout
        print(os.path.basename(lib), " | ".join(l.split("fallback")[0].strip() for l in out.splitlines() if "Mrays" in l))
