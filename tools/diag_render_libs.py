"""Diagnostic (GPU box): 1080p path-tracer ms/spp for library variants under mray_b200/lib/variants/."""
import os, sys, glob, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1:
    sys.path.insert(0, ROOT)
    from mray_b200 import capi
    capi.LIB_PATH = sys.argv[1]
    sys.argv = [sys.argv[0]]
    exec(open(os.path.join(ROOT, "tools", "diag_render_variants.py")).read().replace("(True, False, True, False)", "(False, False)"))
else:
    for lib in [os.path.join(ROOT, "mray_b200", "lib", "libmray_b200.so")] + sorted(glob.glob(os.path.join(ROOT, "mray_b200", "lib", "variants", "*.so"))):
        out = subprocess.run([sys.executable, __file__, lib], capture_output=True, text=True)
        print(os.path.basename(lib), " | ".join(l.strip() for l in out.stdout.splitlines() if "ms/spp" in l), out.stderr[-200:] if out.returncode else "", flush=True)
