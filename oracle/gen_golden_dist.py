"""Generates the piecewise-constant 2-D distribution / skysphere-converter golden vectors by RUNNING THE UNMODIFIED
REFERENCE (oracle/_ref/ref_dist_tap: DistributionGroupPwC2D on the CPU backend, Tracer/Distributions.cu, and the
Spherical / CoOcta coordinate converters of Tracer/LightsDefault.hpp): tests/golden/dist2d_*.npz. Authoring container only."""
import os, subprocess, tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
TAP = os.path.join(HERE, "_ref", "ref_dist_tap")


def functions():
    rng = np.random.default_rng(77)
    yield "uniform", np.full((24, 48), 12.0, np.float32)                      # T_Distributions.cu Dist_PiecewiseConstant2D.Uniform (smaller)
    f = rng.random((33, 61), dtype=np.float32) ** 4 * 50.0                     # HDR-like: mostly dark, a few bright texels
    f[7, 11] = 4000.0                                                          # a "sun"
    yield "hdr", f
    g = rng.standard_normal((16, 20)).astype(np.float32)                       # negative values: |f| is what counts
    g[3, :5] = 0.0                                                             # leading zero-probability texels in a row
    yield "signed", g
    yield "row", (rng.random((1, 97), dtype=np.float32) + 0.01).astype(np.float32)
    yield "column", (rng.random((53, 1), dtype=np.float32) + 0.01).astype(np.float32)


def run(f, xi, dirs):
    h, w = f.shape
    with tempfile.TemporaryDirectory() as d:
        i, o = os.path.join(d, "i.bin"), os.path.join(d, "o.bin")
        with open(i, "wb") as fh:
            np.array([w, h, xi.shape[0], dirs.shape[0]], np.uint32).tofile(fh)
            f.astype(np.float32).tofile(fh); xi.astype(np.float32).tofile(fh); dirs.astype(np.float32).tofile(fh)
        subprocess.run([TAP, i, o], check=True, timeout=600)
        out = np.fromfile(o, np.float32)
    k = 0
    cdf_x = out[k:k + w * h].reshape(h, w); k += w * h
    cdf_y = out[k:k + h]; k += h
    samples = out[k:k + 4 * xi.shape[0]].reshape(-1, 4); k += 4 * xi.shape[0]
    conv = out[k:].reshape(2, dirs.shape[0], 8)
    return cdf_x, cdf_y, samples, conv


if __name__ == "__main__":
    rng = np.random.default_rng(5)
    xi = rng.random((512, 2), dtype=np.float32)
    xi[:4] = [[0.0, 0.0], [0.99999994, 0.99999994], [0.5, 0.0], [0.0, 0.5]]
    dirs = rng.standard_normal((256, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    dirs[:6] = [[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]]
    dirs = dirs.astype(np.float32)
    for name, f in functions():
        cdf_x, cdf_y, samples, conv = run(f, xi, dirs)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"dist2d_{name}.npz"), function=f, xi=xi, dirs=dirs,
                            cdf_x=cdf_x, cdf_y=cdf_y, samples=samples, converters=conv)
        print(name, f.shape, "cdfY last", cdf_y[-1], "pdf mean", samples[:, 2].mean())
