"""Runs the same short measurement (bench.py at reduced spp, config-2 sub-leg included) on several builds of the library:
    python tools/bench_variants.py [--spp 64] name=path/to/lib.so ...   ("default" = mray_b200/lib/libmray_b200.so)
Experiment builds come from `make -C mray_b200/csrc VARIANT=name EXTRA=-D...`. One JSON line per variant on stdout."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spp, items, extra_env = 64, [], {}
args = sys.argv[1:]
while args:
    a = args.pop(0)
    if a == "--spp": spp = int(args.pop(0))
    elif a.startswith("env:"): k, v = a[4:].split("=", 1); extra_env[k] = v
    else: items.append(a)
for it in items or ["default"]:
    name, _, path = it.partition("=")
    env = dict(os.environ, MRB_BENCH_SKIP_E2E="1", MRB_BENCH_SKIP_CPU="1", **extra_env)
    for kv in name.split("+")[1:]:            # name+KEY:VALUE adds an environment variable for that variant
        k, v = kv.split(":", 1); env[k] = v
    if path: env["MRB_LIB_PATH"] = os.path.join(ROOT, path)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3", "--spp", str(spp)],
                       env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        c = d["config"]
        print(json.dumps({"variant": name, "ms_per_spp": c["ms_per_spp_1080p"], "mrays_s": d["value"], "kernel_ms": c["kernel_ms_per_iteration"],
                          "config2": {k: c["config2_traversal"][k] for k in ("mrays_s", "mrays_primary", "mrays_ao_closest", "mrays_ao_anyhit")},
                          "fallback": c["exact_fallback_rays_last_cast"], "film_w": c["film_weight_min_max"]}), flush=True)
    except Exception as e:
        print(json.dumps({"variant": name, "error": str(e), "stderr": r.stderr[-800:], "stdout": r.stdout[-300:]}), flush=True)
