"""torchrun entry (one rank per GPU): sample-sharded path tracing with an NCCL film all-reduce.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/render_multi_gpu.py [res] [spp] [cornell|arcade]
"arcade" = BASELINE config 3: the 264 K-triangle mesh, 64 Lambert + 200 emissive triangles, (R)PathTracerSpectral,
WithNEEAndMIS, rrRange [3,8], 1920x1080 (res is ignored), total spp sharded over the ranks; prints one JSON line.
Every rank builds the (replicated) BVH, renders its sample range with its own seed, the planar film
(R,G,B,W) is summed over NVLink, rank 0 resolves and reports device-timed ms/spp and Mrays/s."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import mray_b200
from mray_b200 import capi, scenes, sharding

res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 256
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = mray_b200.Context(local); ctx.set_stream(torch.cuda.current_stream())
scene_name = sys.argv[3] if len(sys.argv) > 3 else "cornell"
b, e = sharding.shard_samples(spp, world, rank)
spec = None
if scene_name == "arcade":
    from mray_b200 import spectral
    p, i = scenes.arcade_mesh()
    pidx, pranges, pkeys, palb, prad, _ = scenes.arcade_materials(p, i)
    acc = capi.Accelerator(ctx, p, pidx, prim_ranges=pranges, light_or_mat_keys=pkeys)
    W, H = 1920, 1080
    spec = capi.Spectrum(ctx, spectral.load(), "HyperbolicPBRT") if spectral.available() else None
    mk = lambda n: capi.Renderer(ctx, acc, p.shape[0], pidx.shape[0], palb, prad, scenes.ARCADE_CAMERA, W, H, n, sample_mode="WithNEEAndMIS",
                                 rr_range=(3, 8), seed=sharding.rank_seed(0, rank), spectrum=spec)
    warm = mk(1); warm.iterate(16); torch.cuda.synchronize(); warm.close()      # untimed warm-up on a throw-away renderer
    r = mk(e - b)
else:
    c = scenes.cornell_box()
    order = np.argsort(c["material"], kind="stable")
    idx = np.ascontiguousarray(c["indices"][order]); mat = c["material"][order]
    ranges = [[np.nonzero(mat == m)[0][0], np.nonzero(mat == m)[0][-1] + 1] for m in np.unique(mat)]
    keys = [capi.light_key(0) if m == 3 else int(m) for m in np.unique(mat)]
    acc = capi.Accelerator(ctx, c["positions"], idx, prim_ranges=ranges, light_or_mat_keys=keys)
    W = H = res
    r = capi.Renderer(ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"], res, res, e - b,
                      seed=sharding.rank_seed(0, rank))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if world > 1:   # NCCL builds its communicator on the first collective: keep that out of the timed region
    sharding.reduce_film(torch.zeros(1024, device="cuda"))
    dist.barrier()
torch.cuda.synchronize()
ev0.record()
while True:
    r.iterate(32)
    if r.stats().finished: break
film = torch.empty((4, H, W), dtype=torch.float32, device="cuda")
ctx.check(ctx.lib.mrb_renderer_read_film(ctx.handle, r.handle, film.data_ptr(), capi.MRB_MEM_DEVICE, 0))
sharding.reduce_film(film)
ev1.record(); torch.cuda.synchronize()
ms = sharding.max_over_ranks(ev0.elapsed_time(ev1), device="cuda")
st = r.stats()
rays = torch.tensor([st.closestRays + st.shadowRays], dtype=torch.float64, device="cuda")
if world > 1: dist.all_reduce(rays)
if rank == 0:
    img = sharding.resolve(film.cpu().numpy())
    if scene_name == "arcade":
        import json
        print(json.dumps({"config": "3: arcade 264K tris, 64 Lambert + 200 emissive, %s, WithNEEAndMIS rr[3,8], 1920x1080" % ("PathTracerSpectral" if spec is not None else "PathTracerRGB"),
                          "n_gpus": world, "total_spp": spp, "spp_per_gpu": e - b, "ms_total": round(ms, 3), "ms_per_spp_1080p": round(ms / spp, 4),
                          "mrays_s": round(rays.item() / ms / 1e3, 1), "film_weight_min_max": [film[3].min().item(), film[3].max().item()],
                          "sharding": "sample ranges per rank, BVH replicated, one NCCL all-reduce of the 33 MB film inside the timed region (max over ranks)"}))
    print(f"ranks {world}: {W}x{H} {spp} spp, {ms:.2f} ms total, {ms/spp:.4f} ms/spp, {rays.item()/ms/1e3:.1f} Mrays/s, "
          f"weight min/max {film[3].min().item():.1f}/{film[3].max().item():.1f}, mean {img.mean(axis=(0,1))}")
r.close(); acc.close(); ctx.close()
if world > 1: dist.destroy_process_group()
