"""Generates the refracted-ray-cone golden vectors by RUNNING THE UNMODIFIED REFERENCE (oracle/_ref/libref_taps.so::
ref_refract_ray_cone: RefractMatDetail::RefractMaterial::RefractRayCone + RayConeSurface::ConeAfterScatter, Tracer/MaterialsDefault.hpp
L355-462): tests/golden/refract_ray_cone.npz. Authoring container only."""
import os, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402


def main():
    rng = np.random.default_rng(3)
    n = 3000
    inp = np.zeros((n, 12), np.float32)
    inp[:, 0] = rng.uniform(-0.03, 0.05, n)                     # aperture (negative = converging cones too)
    inp[:, 1] = rng.uniform(0.0, 0.5, n); inp[:50, 1] = 0.0     # width (some exactly 0: camera rays)
    inp[:, 2] = rng.uniform(-0.02, 0.02, n); inp[::3, 2] = 0.0  # curvature term betaN (a third flat)
    nrm = rng.standard_normal((n, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    wo = rng.standard_normal((n, 3)); wo /= np.linalg.norm(wo, axis=1, keepdims=True)
    flip = (wo * nrm).sum(axis=1) < 0; wo[flip] *= -1           # wO on the side of the (already flipped) geometric normal
    inp[:, 3:6] = wo; inp[:, 6:9] = nrm
    inp[:, 9] = rng.uniform(1.0, 1.8, n); inp[:, 10] = rng.uniform(1.0, 1.8, n)   # front / back index of refraction
    inp[:, 11] = rng.integers(0, 2, n)                          # backSide: swaps them
    out = np.zeros((n, 2), np.float32)
    R = O.ref()
    R.ref_refract_ray_cone.argtypes = [O.C.c_void_p, O.C.c_uint32, O.C.c_void_p]
    R.ref_refract_ray_cone(inp.ctypes.data, n, out.ctypes.data)
    path = os.path.join(ROOT, "tests", "golden", "refract_ray_cone.npz")
    np.savez_compressed(path, inputs=inp, cones=out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
