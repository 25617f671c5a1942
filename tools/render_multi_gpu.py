"""torchrun entry (one rank per GPU): sample-sharded path tracing with an NCCL film all-reduce.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/render_multi_gpu.py [res] [spp]
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
c = scenes.cornell_box()
order = np.argsort(c["material"], kind="stable")
idx = np.ascontiguousarray(c["indices"][order]); mat = c["material"][order]
ranges = [[np.nonzero(mat == m)[0][0], np.nonzero(mat == m)[0][-1] + 1] for m in np.unique(mat)]
keys = [capi.light_key(0) if m == 3 else int(m) for m in np.unique(mat)]
acc = capi.Accelerator(ctx, c["positions"], idx, prim_ranges=ranges, light_or_mat_keys=keys)
b, e = sharding.shard_samples(spp, world, rank)
r = capi.Renderer(ctx, acc, c["positions"].shape[0], idx.shape[0], c["albedo"][:3], c["radiance"], c["camera"], res, res, e - b,
                  seed=sharding.rank_seed(0, rank))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if world > 1: dist.barrier()
torch.cuda.synchronize()
ev0.record()
while True:
    r.iterate(32)
    if r.stats().finished: break
film = torch.empty((4, res, res), dtype=torch.float32, device="cuda")
ctx.check(ctx.lib.mrb_renderer_read_film(ctx.handle, r.handle, film.data_ptr(), capi.MRB_MEM_DEVICE, 0))
sharding.reduce_film(film)
ev1.record(); torch.cuda.synchronize()
ms = sharding.max_over_ranks(ev0.elapsed_time(ev1), device="cuda")
st = r.stats()
rays = torch.tensor([st.closestRays + st.shadowRays], dtype=torch.float64, device="cuda")
if world > 1: dist.all_reduce(rays)
if rank == 0:
    img = sharding.resolve(film.cpu().numpy())
    print(f"ranks {world}: {res}x{res} {spp} spp, {ms:.2f} ms total, {ms/spp:.4f} ms/spp, {rays.item()/ms/1e3:.1f} Mrays/s, "
          f"weight min/max {film[3].min().item():.1f}/{film[3].max().item():.1f}, mean {img.mean(axis=(0,1))}")
r.close(); acc.close(); ctx.close()
if world > 1: dist.destroy_process_group()
