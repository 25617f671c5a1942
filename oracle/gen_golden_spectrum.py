"""Generates the spectral golden vectors by RUNNING THE UNMODIFIED REFERENCE (oracle/_ref/ref_spectrum_tap,
built by oracle/ref_build/build_ref.sh from /root/reference): tests/golden/spectrum_mode{0,1,2}.npz and the
colour-space tables the product consumes, mray_b200/data/spectral_tables_ACES_CG.bin
(f32: observerXYZ[471*3] normalised, illuminant[471] normalised, xyzToRGB[9]) — the data
SpectrumContextJakob2019's constructor uploads next to the LUT (Tracer/SpectrumContext.cu:L416-438).
Only runs in the authoring container (needs oracle/_ref)."""
import os, subprocess, tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
TAP = os.path.join(HERE, "_ref", "ref_spectrum_tap")
N = 96
# Tests/Tracer/T_Spectrum.cu:L61-88 colours + seeded random ones
COLORS = np.array([[0.00368, 0.00304, 0.01033], [0, 0, 0], [0.5, 0.5, 0.5], [1, 1, 1],
                   [0.85, 0.15, 0.15], [0.15, 0.85, 0.15], [0.15, 0.15, 0.85]], np.float32)
rng = np.random.default_rng(2019)
COLORS = np.concatenate([COLORS, rng.uniform(0, 1, size=(5, 3)).astype(np.float32)])


def run(mode):
    # equally spaced (the reference test's inverse of ToFloat01) + random + edge random numbers
    strat = ((np.arange(N // 2, dtype=np.uint64) * ((1 << 24) // (N // 2))) << 8).astype(np.uint32)
    rnd = rng.integers(0, 1 << 32, size=N - N // 2 - 2, dtype=np.uint64).astype(np.uint32)
    rn = np.concatenate([strat, rnd, np.array([0xFFFFFFFF, 0x80000000], np.uint32)])
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        with open(fin, "wb") as f:
            f.write(np.array([N, len(COLORS)], np.uint32).tobytes()); f.write(rn.tobytes()); f.write(COLORS.tobytes())
        subprocess.run([TAP, fin, fout, str(mode)], check=True, timeout=600)
        o = np.fromfile(fout, np.float32)
    k = 0
    def take(cnt, shape):
        nonlocal k
        a = o[k:k + cnt].reshape(shape); k += cnt; return a
    obs = take(471 * 3, (471, 3)); ill = take(471, (471,)); M = take(9, (3, 3))
    waves = take(N * 4, (N, 4)); pdfs = take(N * 4, (N, 4))
    alb, rad, rgbA, rgbR = [], [], [], []
    for _ in range(len(COLORS)):
        alb.append(take(N * 4, (N, 4))); rad.append(take(N * 4, (N, 4)))
        rgbA.append(take(N * 4, (N, 4))); rgbR.append(take(N * 4, (N, 4)))
    assert k == o.size
    return dict(randoms=rn, colors=COLORS, waves=waves, pdfs=pdfs, albedo_spec=np.stack(alb), radiance_spec=np.stack(rad),
                rgb_albedo_illum=np.stack(rgbA), rgb_radiance=np.stack(rgbR), radiance_scale=np.float32(4.5)), (obs, ill, M)


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    os.makedirs(os.path.join(ROOT, "mray_b200", "data"), exist_ok=True)
    for mode in (0, 1, 2):
        g, (obs, ill, M) = run(mode)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"spectrum_mode{mode}.npz"), **g)
    with open(os.path.join(ROOT, "mray_b200", "data", "spectral_tables_ACES_CG.bin"), "wb") as f:
        f.write(obs.astype(np.float32).tobytes()); f.write(ill.astype(np.float32).tobytes()); f.write(M.astype(np.float32).tobytes())
    print("wrote goldens; tables", obs.shape, ill.shape, M)
