"""CPU: the texture filter of the estimator oracle (oracle/pt_oracle.c::orc_texture_sample) against an independent
numpy restatement of the reference's host-backend view (Device/CPU/TextureViewCPU.h:L196-385) and hand-computed
known answers; the textured render is pinned by a reference image in tests/test_oracle_pt.py."""
import numpy as np

import oracle_lib as O


def np_sample(tex, interp, edge, uv):
    h, w = tex.shape[:2]
    data = tex[..., :3].astype(np.float32) * (np.float32(1.0) / np.float32(255.0)) if tex.dtype == np.uint8 else tex[..., :3]

    def resolve(i, n):
        if edge == "Clamp":
            return min(max(i, 0), n - 1)
        dim = int(i / n)                     # C++ truncation
        r = int(np.fmod(i, n))
        if r < 0:
            r += n
        if edge == "Mirror" and (dim & 1) == 1:
            r = n - r
        return min(r, n - 1)
    out = []
    for u, v in uv:
        tu, tv = np.float32(u) * np.float32(w), np.float32(v) * np.float32(h)
        if interp == "Nearest":
            x = int(np.copysign(np.floor(abs(tu - np.float32(0.5)) + np.float32(0.5)), tu - np.float32(0.5)))
            y = int(np.copysign(np.floor(abs(tv - np.float32(0.5)) + np.float32(0.5)), tv - np.float32(0.5)))
            out.append(data[resolve(y, h), resolve(x, w)]); continue
        fx, bx = np.modf(np.float32(tu - np.float32(0.5))); fy, by = np.modf(np.float32(tv - np.float32(0.5)))
        x0, y0 = int(bx), int(by)
        if fx < 0: x0 -= 1; fx = -fx
        if fy < 0: y0 -= 1; fy = -fy
        fx, fy = np.float32(fx), np.float32(fy)
        lerp = lambda a, b, t: a * (np.float32(1) - t) + b * t
        p0 = lerp(data[resolve(y0, h), resolve(x0, w)], data[resolve(y0, h), resolve(x0 + 1, w)], fx)
        p1 = lerp(data[resolve(y0 + 1, h), resolve(x0, w)], data[resolve(y0 + 1, h), resolve(x0 + 1, w)], fx)
        out.append(lerp(p0, p1, fy))
    return np.array(out, np.float32)


def test_texture_filter_matches_numpy_restatement():
    rng = np.random.default_rng(1)
    uv = np.concatenate([rng.uniform(-2.5, 3.5, size=(400, 2)), [[0, 0], [1, 1], [0.5, 0.5], [-1e-4, 1 - 1e-4], [0.0625, 0.9375]]]).astype(np.float32)
    for tex in (rng.random((5, 7, 3)).astype(np.float32), rng.integers(0, 256, size=(4, 4, 4), dtype=np.uint8)):
        for interp in ("Nearest", "Linear"):
            for edge in ("Wrap", "Clamp", "Mirror"):
                got = O.oracle_texture_sample(dict(data=tex, interp=interp, edge=edge), uv)
                assert np.array_equal(got, np_sample(tex, interp, edge, uv)), (tex.dtype, interp, edge)


def test_texture_filter_known_answers():
    tex = np.zeros((2, 2, 3), np.float32)
    tex[0, 0] = 1.0; tex[0, 1] = 2.0; tex[1, 0] = 3.0; tex[1, 1] = 4.0
    t = dict(data=tex, interp="Linear", edge="Clamp")
    # texel centres return the texel, the middle their mean; row 0 is v in [0, 0.5)
    got = O.oracle_texture_sample(t, [[0.25, 0.25], [0.75, 0.25], [0.25, 0.75], [0.75, 0.75], [0.5, 0.5], [0.5, 0.25]])
    assert np.allclose(got[:, 0], [1, 2, 3, 4, 2.5, 1.5])
    n = dict(data=tex, interp="Nearest", edge="Wrap")
    got = O.oracle_texture_sample(n, [[0.1, 0.1], [0.6, 0.1], [1.1, 0.6], [-0.4, 0.6]])
    assert np.allclose(got[:, 0], [1, 2, 3, 4])
    u8 = dict(data=np.full((1, 1, 4), 255, np.uint8), interp="Linear", edge="Wrap")
    assert np.array_equal(O.oracle_texture_sample(u8, [[0.3, 0.8]]), np.ones((1, 3), np.float32))
