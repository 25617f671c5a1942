"""GPU (B200): multi-GPU inside the product and the ImageTiler (Tracer/RenderImage.cpp:L20-136).

One TracerB200 drives MRB_DEVICES devices: scene replicated, every pass's sample range split over the devices, films
summed into device 0's over peer memory, ONE RenderImageSection handed to the caller. Random numbers are a function of
(seed, pixel, sample index), so the multi-device and the tiled renders must equal the single-device single-tile render up
to the order of the film's float additions. With one physical GPU the same code paths run with two contexts on device 0
("0,0"); with two or more GPUs the peer-memory path is exercised as well."""
import os

import numpy as np
import pytest

import oracle_lib as O
from mray_b200 import capi, scenes
from test_gpu_render import cornell_accel

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "mray_b200", "lib", "libTracerDLL_B200.so")
needs_plugin = pytest.mark.skipif(not (os.path.exists(PLUGIN) and O.driver_available()), reason="plugin / driver were not prebuilt")


def device_count():
    import torch
    return torch.cuda.device_count()


def close(a, b):
    return np.allclose(a, b, rtol=3e-5, atol=2e-4)


@pytest.mark.parametrize("second_device", [0, 1])
def test_reduce_peers_equals_one_renderer(gpu_ctx, second_device):
    """mrb_renderer_reduce_peers: two renderers with complementary sample ranges, films added over peer memory."""
    if second_device >= device_count():
        pytest.skip("needs a second GPU")
    res, spp = 40, 48
    ctx1 = capi.Context(second_device)
    accs, rs = [], []
    for ctx, (s0, n) in ((gpu_ctx, (0, 20)), (ctx1, (20, 28))):
        c, idx, tm, acc = cornell_accel(ctx)
        r = capi.Renderer(ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"], res, res, n,
                          seed=4, sample_offset=s0, job_spp=spp)
        assert r.run_pass(4).finished
        accs.append(acc); rs.append(r)
    rs[0].reduce_peers([rs[1]])
    rgb, w = rs[0].read_film()
    rgb1, w1 = rs[1].read_film()
    assert not rgb1.any() and not w1.any()                       # the peer's film was cleared by the reduction
    c, idx, tm, acc = cornell_accel(gpu_ctx)
    one = capi.Renderer(gpu_ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"], res, res, spp, seed=4)
    assert one.run_pass(4).finished
    rgb0, w0 = one.read_film()
    assert np.array_equal(w, w0) and close(rgb, rgb0), np.abs(rgb - rgb0).max()
    for r in rs + [one]:
        r.close()
    for a in accs + [acc]:
        a.close()
    ctx1.close()


def _plugin_render(env, **kw):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        c = scenes.cornell_box()
        b = O.batched_scene(c["positions"], c["indices"], c["material"])
        return O.driver_render(PLUGIN, b, c["albedo"], 3, c["radiance"], c["camera"], kw.pop("res", 48), kw.pop("res2", 48), kw.pop("spp", 96),
                               sample_mode="WithNEEAndMIS", rr_range=(2, 20), seed=21, **kw)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@needs_plugin
def test_plugin_multi_device_equals_single_device():
    devs = "0,1" if device_count() >= 2 else "0,0"
    for kw in (dict(burst_size=32), dict(latency=True, spp=12), dict()):       # burst passes, latency passes, throughput (1 bounce / call)
        img1, w1, st1 = _plugin_render({"MRB_DEVICES": "1"}, **dict(kw))
        img2, w2, st2 = _plugin_render({"MRB_DEVICES": devs}, **dict(kw))
        assert np.array_equal(w1, w2) and np.allclose(w1, kw.get("spp", 96)), (w1.min(), w2.min())
        assert close(img1, img2), (kw, np.abs(img1 - img2).max())
        if kw:
            assert st1["iterations"] == st2["iterations"]                     # same number of DoRenderWork calls / sections
    # three "devices" and a sample count that does not divide evenly
    img3, w3, _ = _plugin_render({"MRB_DEVICES": "0,0,0"}, burst_size=32, spp=50)
    img1, w1, _ = _plugin_render({"MRB_DEVICES": "1"}, burst_size=32, spp=50)
    assert np.array_equal(w1, w3) and close(img1, img3)


@needs_plugin
def test_plugin_sample_shards_add_up():
    """MRB_SPP_SHARD = rank/world: the processes of a multi-process job render disjoint sample ranges whose films add up."""
    full, wf, _ = _plugin_render({}, burst_size=16, spp=64)
    acc = np.zeros_like(full, dtype=np.float64); wsum = np.zeros_like(wf, dtype=np.float64)
    for rank in range(3):
        img, w, st = _plugin_render({"MRB_SPP_SHARD": f"{rank}/3"}, burst_size=16, spp=64)
        acc += img.astype(np.float64) * w[..., None]; wsum += w
    assert np.allclose(wsum, 64)
    assert close(acc / wsum[..., None], full)


@needs_plugin
def test_image_tiler_sections_cover_the_image():
    """parallelizationHint smaller than the image: ImageTiler::FindOptimumTileSize cuts it into equal tiles, every
    DoRenderWork renders burstSize samples of ONE tile and hands over that tile's section (pixelMin / pixelMax); the
    assembled image equals the single-tile render."""
    one, w1, st1 = _plugin_render({}, res=48, res2=32, spp=32, burst_size=16)
    til, w2, st2 = _plugin_render({}, res=48, res2=32, spp=32, burst_size=16, parallel_hint=400)    # 48x32 -> 24x16 tiles (2 x 2)
    assert st1["iterations"] == 2 and st2["iterations"] == 2 * 4
    assert np.array_equal(w1, w2) and close(one, til), np.abs(one - til).max()
    # a hint that does not divide the image: 16 x 11 tiles, 3 x 3 of them, the last row only 10 pixels high
    til2, w3, st3 = _plugin_render({}, res=48, res2=32, spp=32, burst_size=16, parallel_hint=200)
    assert st3["iterations"] == 2 * 9
    assert np.array_equal(w1, w3) and close(one, til2)
    # Throughput with burstSize 1 on a multi-tile image runs burst passes too (PathTracerRendererBase::DoRender, L515-533)
    til3, w4, st4 = _plugin_render({}, res=48, res2=32, spp=4, burst_size=1, parallel_hint=400)
    one3, w5, st5 = _plugin_render({}, res=48, res2=32, spp=4, burst_size=4)
    assert st4["iterations"] == 4 * 4 and np.array_equal(w4, w5) and close(til3, one3)
