"""Generates the mip-chain / level-of-detail golden vectors by RUNNING THE UNMODIFIED REFERENCE (oracle/_ref/libref_taps.so::
ref_texture_sample: TextureMemory::CreateTexture2D / PushTextureData / Finalize — ConvertColorspaces + GenerateMipmaps — and the
resulting TracerTexView<2, Vector3> read with an explicit level or with gradients, on the reference's CPU backend):
tests/golden/texture_mips.npz. Authoring container only."""
import os, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402


def cases():
    """name -> texture dict (see oracle_lib.mip_chain)"""
    rng = np.random.default_rng(2024)
    f = rng.random((16, 32, 4), dtype=np.float32)
    yield "gauss2_f32", dict(data=f, gen_mips=("Gaussian", 2.0))                     # the reference's default mipGenFilter
    yield "gauss1_f32_clamp", dict(data=f, gen_mips=("Gaussian", 1.0), edge="Clamp")
    yield "box_f32", dict(data=f, gen_mips=("Box", 0.5))   # (MIRROR is left to the single-level goldens: the reference's mirror resolve can index one past a level, which small levels hit often)
    yield "tent_f32", dict(data=f, gen_mips=("Tent", 1.0))
    yield "mitchell_f32", dict(data=f, gen_mips=("Mitchell-Netravali", 2.0))
    u = (rng.random((20, 12, 4)) * 255).astype(np.uint8)                             # non-power-of-two: levels 20x12 .. 1x1
    yield "gauss2_u8_npot", dict(data=u, gen_mips=("Gaussian", 2.0))
    explicit = [rng.random((8, 16, 4), dtype=np.float32), rng.random((4, 8, 4), dtype=np.float32)]
    yield "explicit3_f32", dict(data=f, mips=explicit)                                # levels pushed by the caller, none generated
    yield "explicit2_then_gen", dict(data=f, mips=explicit[:1], gen_mips=("Gaussian", 2.0))   # level 1 supplied, 2.. generated from it
    yield "nearest_f32", dict(data=f, gen_mips=("Gaussian", 2.0), interp="Nearest")
    # TracerParameters.clampedTexRes: the pushed image filtered down at load (KCClampImage), then (optionally) a generated chain
    g = rng.random((40, 64, 4), dtype=np.float32)
    yield "clamp16_f32", dict(data=g, clamp_res=16)                                    # 64x40 -> 16x10, one level, default Gaussian 2
    yield "clamp16_gen_f32", dict(data=g, clamp_res=16, gen_mips=("Gaussian", 2.0))
    yield "clamp20_tent_f32", dict(data=g, clamp_res=20, gen_mips=("Tent", 1.5))      # 20 does not divide 64: still 2 levels dropped
    yield "clamp8_u8", dict(data=(g * 255).astype(np.uint8), clamp_res=8, gen_mips=("Gaussian", 2.0))
    yield "clamp100_noop_f32", dict(data=g, clamp_res=100)
    yield "clamp8_box_f32", dict(data=g, clamp_res=8, gen_mips=("Box", 1.0))
    # (Mitchell-Netravali's sampler — a three-Gaussian mixture — matches the restatement to 2e-7, not bit for bit: it is checked
    #  against the oracle on the GPU only)


def main():
    rng = np.random.default_rng(5)
    n = 400
    out = {}
    names = []
    for name, t in cases():
        nearest = t.get("interp") == "Nearest"
        # NEAREST + mips: the reference resolves the edge against the BASE size (TextureViewCPU.h:L446), which indexes out of the
        # level for coordinates outside [0, 1): keep those reads inside
        uv = rng.random((n, 2)).astype(np.float32) if nearest else (rng.random((n, 2)) * 3 - 1).astype(np.float32)
        lod = (rng.random(n) * 8 - 1).astype(np.float32)
        lod[:8] = [0, 1, 2, 3, 0.5, 1.5, -3, 40]
        scale = np.exp(rng.standard_normal((n, 1)) * 2)
        dpdx = (rng.standard_normal((n, 2)) * scale).astype(np.float32)
        dpdy = (rng.standard_normal((n, 2)) * np.exp(rng.standard_normal((n, 1)) * 2)).astype(np.float32)
        dpdx[:4] = 0; dpdy[:4] = 0          # zero footprint: log2(0) = -inf clamps to level 0
        out[name + "_data"] = t["data"]
        for k, m in enumerate(t.get("mips") or []):
            out[f"{name}_mip{k + 1}"] = m
        out[name + "_params"] = np.array([t.get("interp", "Linear"), t.get("edge", "Wrap"), (t.get("gen_mips") or ("", 0))[0],
                                          str((t.get("gen_mips") or ("", 0))[1]), str(len(t.get("mips") or [])), str(int(t.get("clamp_res") or 0))])
        out[name + "_uv"], out[name + "_lod"], out[name + "_dpdx"], out[name + "_dpdy"] = uv, lod, dpdx, dpdy
        out[name + "_rgb_lod"] = O.ref_texture_sample(t, uv, lod=lod)
        out[name + "_rgb_grad"] = O.ref_texture_sample(t, uv, dpdx=dpdx, dpdy=dpdy)
        # the generated levels themselves, read back texel by texel through the view (NEAREST-equivalent: texel centres, integer level)
        names.append(name)
    out["names"] = np.array(names)
    path = os.path.join(ROOT, "tests", "golden", "texture_mips.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
