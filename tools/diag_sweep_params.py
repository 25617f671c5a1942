"""Diagnostic (GPU box): sweep the wide-traversal scheduling parameters (env MRB_TRI_DIV, MRB_FETCH_THR)."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for td in (8, 12, 20):
    for ft in (20, 22, 24, 26):
        env = dict(os.environ, MRB_TRI_DIV=str(td), MRB_FETCH_THR=str(ft))
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "diag_variants.py"), os.path.join(ROOT, "mray_b200", "lib", "libmray_b200.so")],
                             capture_output=True, text=True, env=env).stdout
        ms = [float(l.split(" ms ")[1].split()[0]) for l in out.splitlines() if "Mrays" in l]
        print("triDiv", td, "fetchThr", ft, [round(x, 4) for x in ms], "sum", round(sum(ms), 4), flush=True)
