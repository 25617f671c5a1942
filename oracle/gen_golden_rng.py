"""Generates the sampler golden vectors by RUNNING THE UNMODIFIED REFERENCE (oracle/_ref/ref_rng_tap, built by
oracle/ref_build/build_spectral_data.sh from /root/reference): tests/golden/rng_{sobol,zsobol}.npz and the Joe-Kuo
generator matrices the product consumes as data, mray_b200/data/sobol_matrices.bin (u32[256*52], the table the
reference uploads in RNGGroupSobol's constructor, Tracer/Random.cu:L911). Authoring container only."""
import os, subprocess, tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
TAP = os.path.join(HERE, "_ref", "ref_rng_tap")
W, H, SEED, MAX_SPP, INCREMENTS = 12, 5, 0x1234567800000042, 4, 13     # 13 increments cross three ZSobol rounds (4, 8, 16 spp)
CONFIGS = [(0, [2, 1, 3, 2, 1]), (37, [3, 2, 1]), (249, [2, 2, 3]), (300, [1, 2])]


def run(kind):
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "o.bin")
        subprocess.run([TAP, str(kind), str(W), str(H), str(SEED), str(MAX_SPP), str(INCREMENTS), out], check=True, timeout=600)
        o = np.fromfile(out, np.uint32)
    n = W * H
    k = n
    seeds = o[:n]
    numbers = {}
    for inc in range(INCREMENTS):
        for ci, (_, req) in enumerate(CONFIGS):
            cnt = sum(req) * n
            numbers[f"inc{inc}_cfg{ci}"] = o[k:k + cnt].reshape(sum(req), n); k += cnt
    matrices = o[k:] if kind == 1 else None
    assert (kind == 1 and matrices.size == 256 * 52) or (kind == 2 and k == o.size)
    return seeds, numbers, matrices


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "mray_b200", "data"), exist_ok=True)
    for kind, name in ((1, "sobol"), (2, "zsobol")):
        seeds, numbers, matrices = run(kind)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"rng_{name}.npz"), seeds=seeds, width=W, height=H, seed=np.uint64(SEED),
                            initial_max_spp=MAX_SPP, increments=INCREMENTS,
                            config_dim_start=np.array([c[0] for c in CONFIGS]), config_requests=np.array([c[1] + [0] * (5 - len(c[1])) for c in CONFIGS]),
                            **numbers)
        if matrices is not None:
            matrices.astype(np.uint32).tofile(os.path.join(ROOT, "mray_b200", "data", "sobol_matrices.bin"))
    print("ok")
