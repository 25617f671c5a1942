"""GPU (B200): the wavefront path tracer behind the C-ABI (mrb_renderer_*) against the estimator oracle
(oracle/pt_oracle.c) and the closed-form direct-lighting answer. Statistical parity: relMSE <= 1e-3 on
converged images (north_star); per-sample equality is impossible even between the reference's own
backends."""
import numpy as np
import pytest

import oracle_lib as O
from mray_b200 import capi, scenes
from test_oracle_pt import rect_form_factor

pytestmark = pytest.mark.gpu
REL_MSE_TOL = 1e-3


def rel_mse(a, b):
    return float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))


def cornell_accel(ctx):
    c = scenes.cornell_box()
    order = np.argsort(c["material"], kind="stable")
    idx = np.ascontiguousarray(c["indices"][order]); mat = c["material"][order]
    ranges, keys = [], []
    for m in np.unique(mat):
        w = np.nonzero(mat == m)[0]
        ranges.append([w[0], w[-1] + 1])
        keys.append(capi.light_key(0) if m == 3 else int(m))
    acc = capi.Accelerator(ctx, c["positions"], idx, prim_ranges=ranges, light_or_mat_keys=keys)
    tm = np.where(mat == 3, -1, mat).astype(np.int32)
    return c, idx, tm, acc


def test_cornell_matches_oracle_and_modes_agree(gpu_ctx):
    c, idx, tm, acc = cornell_accel(gpu_ctx)
    res = 32
    spp = {"Pure": 16384, "WithNextEventEstimation": 16384, "WithNEEAndMIS": 131072}
    imgs = {}
    for mode in spp:
        r = capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"],
                          res, res, spp[mode], sample_mode=mode, seed=7)
        imgs[mode], st = r.render(batch=64)
        assert st.finished and st.pathsCompleted == spp[mode] * res * res  # triggerSave condition
        assert st.closestRays >= st.pathsCompleted
        assert (st.shadowRays > 0) == (mode != "Pure")
        rgb, w = r.read_film()
        assert np.allclose(w, spp[mode], rtol=1e-3)                        # Gaussian filter: weight 1 per path
        r.close()
    # oracle at 16384 spp; relMSE between two independent estimates of N and M spp is ~ 7.7 (1/N + 1/M)
    # for this scene (measured oracle-vs-oracle), i.e. ~5e-4 here: the converged images agree to 1e-3.
    ref = O.oracle_render(c["positions"], idx, tm, c["albedo"][:3], c["radiance"], c["camera"], res, res, 16384, sample_mode=2, seed=3)
    err = rel_mse(imgs["WithNEEAndMIS"], ref)
    assert err <= REL_MSE_TOL, err
    # the three estimators share one expectation (compare means away from the directly visible light)
    mask = ref.max(axis=-1) < 5.0
    m_ref = ref[mask].mean(axis=0)
    for mode, im in imgs.items():
        assert np.allclose(im[mask].mean(axis=0), m_ref, rtol=0.02), (mode, im[mask].mean(axis=0), m_ref)
    acc.close()


def test_direct_lighting_closed_form(gpu_ctx):
    half, h, L, rho = 0.5, 1.5, 10.0, 0.6
    floor = np.array([[-50, 0, 50], [50, 0, 50], [50, 0, -50], [-50, 0, -50]], np.float32)
    light = np.array([[-half, h, -half], [half, h, -half], [half, h, half], [-half, h, half]], np.float32)
    pos = np.ascontiguousarray(np.concatenate([floor, light]))
    idx = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7]], np.uint32)
    acc = capi.Accelerator(gpu_ctx, pos, idx, prim_ranges=[[0, 2], [2, 4]], light_or_mat_keys=[0, capi.light_key(0)])
    cam = dict(eye=(0.0, 1.0, 0.0), gaze=(0.0, 0.0, 0.0), up=(0.0, 0.0, -1.0), fov_y_deg=2.0)
    expect = rho * L * 4 * rect_form_factor(half, half, h)
    for mode, tol in (("WithNextEventEstimation", 0.01), ("WithNEEAndMIS", 0.015)):
        r = capi.Renderer(gpu_ctx, acc, 8, 4, [[rho, rho, rho]], [L, L, L], cam, 16, 16, 8192, sample_mode=mode, rr_range=(2, 2))
        img, st = r.render()
        got = img[4:12, 4:12].mean()
        assert abs(got - expect) / expect < tol, (mode, got, expect)
        r.close()
    acc.close()


def test_film_delta_protocol_and_partial_path_pool(gpu_ctx):
    """read_film(clear=True) hands out per-call deltas that sum to the full film (the reference's
    RenderImageSection protocol); a path pool smaller than the tile still completes every sample."""
    c, idx, tm, acc = cornell_accel(gpu_ctx)
    res, spp = 32, 64
    r = capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"],
                      res, res, spp, max_path_count=300, partition_rays=True)  # + material-key ray sort
    total_rgb = np.zeros((res, res, 3)); total_w = np.zeros((res, res))
    for _ in range(100000):
        r.iterate(16)
        rgb, w = r.read_film(clear=True)
        total_rgb += rgb; total_w += w
        if r.stats().finished:
            break
    assert np.allclose(total_w, spp, atol=0.5)
    img = total_rgb / total_w[..., None]
    ref = O.oracle_render(c["positions"], idx, tm, c["albedo"][:3], c["radiance"], c["camera"], res, res, 1024, sample_mode=2, seed=5)
    mask = ref.max(axis=-1) < 5.0
    assert np.allclose(img[mask].mean(axis=0), ref[mask].mean(axis=0), rtol=0.05)
    r.close(); acc.close()


def _rigid(rng, scale=True):
    """Random rotation * uniform scale + translation as a 3x4 float64 matrix."""
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    s = rng.uniform(0.5, 2.0) if scale else 1.0
    return np.concatenate([R * s, rng.uniform(-3, 3, size=(3, 1))], axis=1)


def test_two_level_scene_render_matches_flat_and_oracle(gpu_ctx):
    """The Cornell box cut into per-material instances, each stored in its own local space under a random
    similarity transform (and with per-vertex shading normals on the transformed instances), renders the
    same converged image as the flat mesh: exercises TransformContextSingle on hit positions, geometric
    and shading normals, instance-space emissive triangles and the two-level any-hit shadow rays."""
    c, idx, tm, acc = cornell_accel(gpu_ctx)
    res, spp = 32, 65536
    flat = capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"],
                         res, res, spp, seed=11)
    img_flat, _ = flat.render(batch=64); flat.close()
    rng = np.random.default_rng(5)
    instances, accels, normals = [], [], []
    mats = np.unique(tm)
    for k, m in enumerate(mats):
        tri = idx[tm == m]
        wpos = c["positions"][tri.reshape(-1)].astype(np.float64)            # unwelded world-space vertices
        M = None if k == 1 else _rigid(rng)       # the emissive instance (m == -1, k == 0) is transformed too
        if M is None: lpos = wpos
        else:
            inv = np.linalg.inv(np.vstack([M, [0, 0, 0, 1]]))
            lpos = wpos @ inv[:3, :3].T + inv[:3, 3]
        lpos = np.ascontiguousarray(lpos, np.float32)
        lidx = np.arange(lpos.shape[0], dtype=np.uint32).reshape(-1, 3)
        key = capi.light_key(0) if m == -1 else int(m)
        a = capi.Accelerator(gpu_ctx, lpos, lidx, prim_ranges=[[0, lidx.shape[0]]], light_or_mat_keys=[key])
        fn = np.cross(lpos[1::3] - lpos[0::3], lpos[2::3] - lpos[0::3]); fn /= np.linalg.norm(fn, axis=1, keepdims=True)
        normals.append(None if M is None else np.repeat(fn, 3, axis=0).astype(np.float32))
        accels.append(a); instances.append((a, M))
    scene = capi.Scene(gpu_ctx, instances)
    r = capi.Renderer(gpu_ctx, scene, 0, 0, c["albedo"][:3], c["radiance"], c["camera"], res, res, spp, seed=12,
                      instance_vertex_normals=normals)
    img_scene, st = r.render(batch=64)
    assert st.finished and st.shadowRays > 0
    r.close()
    assert rel_mse(img_scene, img_flat) <= REL_MSE_TOL, rel_mse(img_scene, img_flat)
    ref = O.oracle_render(c["positions"], idx, tm, c["albedo"][:3], c["radiance"], c["camera"], res, res, 16384, sample_mode=2, seed=3)
    assert rel_mse(img_scene, ref) <= REL_MSE_TOL, rel_mse(img_scene, ref)
    scene.close()
    for a in accels: a.close()
    acc.close()


def mirror_accel(ctx):
    """cornell_mirror() as one accelerator: prim ranges by material id (4 = (Mt)Reflect, flat material index 3)."""
    c = scenes.cornell_mirror()
    order = np.argsort(c["material"], kind="stable")
    idx = np.ascontiguousarray(c["indices"][order]); mat = c["material"][order]
    flat = {0: 0, 1: 1, 2: 2, 4: 3}
    ranges, keys = [], []
    for m in np.unique(mat):
        w = np.nonzero(mat == m)[0]
        ranges.append([w[0], w[-1] + 1])
        keys.append(capi.light_key(0) if m == 3 else flat[int(m)])
    acc = capi.Accelerator(ctx, c["positions"], idx, prim_ranges=ranges, light_or_mat_keys=keys)
    tm = np.where(mat == 3, -1, np.vectorize(lambda v: flat.get(int(v), 0))(mat)).astype(np.int32)
    return c, idx, tm, acc, c["albedo"][[0, 1, 2, 4]], np.array([0, 0, 0, 1], np.uint8)


def test_reflect_material_matches_oracle_and_closed_form(gpu_ctx):
    """(Mt)Reflect: the mirror box Cornell against the estimator oracle in the three sample modes, and the closed form
    (a mirror floor under a large one-sided light shows exactly the light's radiance)."""
    c, idx, tm, acc, alb, mtype = mirror_accel(gpu_ctx)
    res = 32
    ref = O.oracle_render(c["positions"], idx, tm, alb, c["radiance"], c["camera"], res, res, 16384, sample_mode=2, seed=3, material_type=mtype)
    mask = ref.max(axis=-1) < 5.0
    for mode, spp in (("WithNEEAndMIS", 65536), ("WithNextEventEstimation", 16384), ("Pure", 65536)):
        r = capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], alb, c["radiance"], c["camera"], res, res, spp,
                          sample_mode=mode, seed=11, material_type=mtype)
        img, st = r.render(batch=64); r.close()
        assert st.finished
        assert np.allclose(img[mask].mean(axis=0), ref[mask].mean(axis=0), rtol=0.02), (mode, img[mask].mean(axis=0), ref[mask].mean(axis=0))
        if mode == "WithNEEAndMIS":
            # the mirror scene is ~5x noisier than the diffuse one (oracle-vs-oracle relMSE ~ 43 (1/N + 1/M)):
            # the converged comparison runs on 4x4 block means
            bm = lambda x: x.reshape(res // 4, 4, res // 4, 4, 3).mean(axis=(1, 3))
            assert rel_mse(bm(img), bm(ref)) <= REL_MSE_TOL, rel_mse(bm(img), bm(ref))
    acc.close()
    L = 7.0
    floor = np.array([[-5, 0, 5], [5, 0, 5], [5, 0, -5], [-5, 0, -5]], np.float32)
    light = np.array([[-30, 4, -30], [30, 4, -30], [30, 4, 30], [-30, 4, 30]], np.float32)
    pos = np.ascontiguousarray(np.concatenate([floor, light])); tri = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7]], np.uint32)
    acc = capi.Accelerator(gpu_ctx, pos, tri, prim_ranges=[[0, 2], [2, 4]], light_or_mat_keys=[0, capi.light_key(0)])
    cam = dict(eye=(0.0, 1.0, 2.0), gaze=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fov_y_deg=10.0)
    for mode in ("Pure", "WithNextEventEstimation", "WithNEEAndMIS"):
        r = capi.Renderer(gpu_ctx, acc, 8, 4, [[0.5, 0.5, 0.5]], [L, L, L], cam, 8, 8, 16, sample_mode=mode, material_type=[1])
        img, st = r.render(); r.close()
        assert np.allclose(img, L, rtol=1e-5), (mode, img.min(), img.max())
    acc.close()


def test_spp_limit_passes(gpu_ctx):
    """mrb_renderer_set_spp_limit (DoLatencyRender): every pass completes exactly its samples of every pixel, and the
    pass-mode film has the expectation of the reference's own render (tests/golden/render_cornell64_spp16384.npz) and of
    the throughput-mode film. 1024 spp at 64x64 on 8x8 block means = 65536 samples per block; the mask comes from the
    converged golden image, never from a noisy render."""
    import os
    golden = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                                  "render_cornell64_spp16384.npz"))["img"].astype(np.float32)
    bm = lambda x: x.reshape(8, 8, 8, 8, 3).mean(axis=(1, 3))
    c, idx, tm, acc = cornell_accel(gpu_ctx)
    res, total, step = 64, 1024, 256
    r = capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"], res, res, total, seed=2)
    r.set_spp_limit(0)
    assert r.stats().finished                                     # an empty pass is complete
    for k in range(1, total // step + 1):
        r.set_spp_limit(step * k)
        while True:
            r.iterate(4)
            st = r.stats()
            if st.finished:
                break
        assert st.pathsCompleted == step * k * res * res
        rgb, w = r.read_film()
        assert np.allclose(w, step * k, rtol=1e-4), (k, w.min(), w.max())
    with pytest.raises(capi.MrbError):
        r.set_spp_limit(total + 1)
    img_pass = rgb / w[..., None]
    r.close()
    r2 = capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"], res, res, total, seed=3)
    img_thr, st2 = r2.render(); r2.close()
    assert st2.finished and st2.pathsCompleted == total * res * res
    # relMSE of 8x8 block means between independent 1024-spp estimates of this scene is ~ 7.7 / 65536 * 2 = 2.4e-4
    rel = lambda a, b: float(np.mean((bm(a) - bm(b)) ** 2 / (bm(b) ** 2 + 1e-2)))
    assert rel(img_pass, golden) <= REL_MSE_TOL, rel(img_pass, golden)
    assert rel(img_thr, golden) <= REL_MSE_TOL, rel(img_thr, golden)
    mask = golden.max(axis=-1) < 5.0
    m_gold = golden[mask].mean(axis=0)
    # channel means over ~4000 pixels x 1024 spp: sigma ~ 0.15 %; 1 % is > 6 sigma
    assert np.allclose(img_pass[mask].mean(axis=0), m_gold, rtol=0.01), (img_pass[mask].mean(axis=0), m_gold)
    assert np.allclose(img_thr[mask].mean(axis=0), m_gold, rtol=0.01), (img_thr[mask].mean(axis=0), m_gold)
    acc.close()


def test_fixed_seed_gives_fixed_image(gpu_ctx):
    """Random numbers are a function of (seed, pixel, sample index): the image does not depend on which slot picks a
    sample up, on the size of the path pool, on how the samples are cut into passes, sample ranges (GPUs) or tiles.
    What remains is the order of the film's float additions (atomics): a few ulp of the accumulated sums."""
    c, idx, tm, acc = cornell_accel(gpu_ctx)
    res, spp = 48, 64
    mk = lambda **kw: capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"],
                                    res, res, kw.pop("spp", spp), seed=kw.pop("seed", 5), **kw)

    def film(r):
        st = r.run_pass(4); assert st.finished
        rgb, w = r.read_film(); r.close()
        return np.concatenate([rgb, w[..., None]], axis=-1).astype(np.float64)
    a = film(mk())
    b = film(mk())                                  # same seed, scheduling free to differ
    close = lambda x, y: np.allclose(x, y, rtol=2e-5, atol=1e-4)
    assert close(a, b), np.abs(a - b).max()
    assert close(a, film(mk(max_path_count=777)))   # a small path pool: different slot <-> sample assignment
    assert not close(a, film(mk(seed=6)))
    # sample ranges: [0, 24) + [24, 64) rendered by two renderers (two GPUs in production) add up to the same film
    lo = film(mk(spp=24, sample_offset=0, job_spp=spp)); hi = film(mk(spp=40, sample_offset=24, job_spp=spp))
    assert close(a, lo + hi), np.abs(a - lo - hi).max()
    # passes over four tiles of a 2 x 2 tiling, two bursts each
    r = capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"],
                      res // 2, res // 2, spp, seed=5, full_resolution=(res, res), region_min=(0, 0))
    r.set_spp_limit(0)
    tiled = np.zeros_like(a)
    for s0, cnt in ((0, 40), (40, 24)):
        for ty in range(2):
            for tx in range(2):
                r.begin_pass((tx * res // 2, ty * res // 2), (res // 2, res // 2), s0, cnt)
                assert r.run_pass(4).finished
                rgb, w = r.read_film(clear=True)
                tiled[ty * res // 2:(ty + 1) * res // 2, tx * res // 2:(tx + 1) * res // 2] += np.concatenate([rgb, w[..., None]], axis=-1)
    r.close()
    assert close(a, tiled), np.abs(a - tiled).max()
    acc.close()


def test_async_film_handoff_and_poll(gpu_ctx):
    """mrb_renderer_film_handoff: deltas copied on the copy stream while rendering continues into the second film
    buffer; their sum is the full film. mrb_renderer_poll_stats reports completion without a stream drain."""
    import time
    import torch
    c, idx, tm, acc = cornell_accel(gpu_ctx)
    res, spp = 32, 96
    r = capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"], res, res, spp, seed=9)
    bufs = [torch.zeros((4, res, res), dtype=torch.float32).pin_memory() for _ in range(2)]
    done = set()
    total = np.zeros((4, res, res), np.float64)

    def consume(i):
        t0 = time.time()
        while i not in done and time.time() - t0 < 20.0:
            time.sleep(0.0005)
        assert i in done, "hand-off callback never ran"
        total[:] += bufs[i % 2].numpy()
    issued, finished = 0, False
    while issued < 100000:
        r.iterate(3)
        st = r.poll_stats()
        r.film_handoff(bufs[issued % 2], on_complete=lambda _u, i=issued: done.add(i))
        issued += 1
        if issued >= 2:
            consume(issued - 2)                     # the previous hand-off is read while this one is in flight
        if finished:                                # one more delta after the snapshot that reported the end
            break
        finished = bool(st.finished)
    consume(issued - 1)
    assert issued > 3
    assert np.allclose(total[3], spp, rtol=1e-4), (total[3].min(), total[3].max())
    r2 = capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"], res, res, spp, seed=9)
    st2 = r2.run_pass(4)
    rgb, w = r2.read_film(); r2.close(); r.close(); acc.close()
    assert np.allclose(np.moveaxis(total[:3], 0, -1), rgb, rtol=2e-5, atol=1e-4)
