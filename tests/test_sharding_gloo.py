"""CPU, world_size 2 over gloo: the multi-GPU host logic (sample sharding, film reduction, max-over-ranks
timing) without GPUs. The per-rank film comes from the estimator oracle, so the reduced image must equal
a single-process render of the full sample count in expectation."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mray_b200 import sharding


def test_shard_samples_partition():
    for total in (1, 7, 64, 1024):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_samples(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.shard_tiles(4, 2, 1) == [1, 3]
    assert len({sharding.rank_seed(0, r) for r in range(8)}) == 8


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import oracle_lib as O
    from mray_b200 import scenes
    c = scenes.cornell_box()
    tm = np.where(c["material"] == 3, -1, c["material"]).astype(np.int32)
    res, total_spp = 16, 256
    b, e = sharding.shard_samples(total_spp, world, rank)
    img = O.oracle_render(c["positions"], c["indices"], tm, c["albedo"][:3], c["radiance"], c["camera"], res, res, e - b,
                          sample_mode=2, seed=sharding.rank_seed(11, rank), threads=2)
    film = np.concatenate([np.moveaxis(img, -1, 0) * (e - b), np.full((1, res, res), float(e - b), np.float32)]).astype(np.float32)
    t = torch.from_numpy(film.copy())
    sharding.reduce_film(t)
    slow = sharding.max_over_ranks(1.0 + rank)
    if rank == 0:
        np.save(os.path.join(out_dir, "film.npy"), t.numpy())
        np.save(os.path.join(out_dir, "slow.npy"), np.array([slow]))
    dist.destroy_process_group()


def test_two_rank_film_reduce(tmp_path):
    world, port = 2, 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    film = np.load(tmp_path / "film.npy")
    assert np.allclose(film[3], 256.0)                       # weights of both ranks arrived
    assert float(np.load(tmp_path / "slow.npy")[0]) == 2.0   # max over ranks
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import oracle_lib as O
    from mray_b200 import scenes
    c = scenes.cornell_box()
    tm = np.where(c["material"] == 3, -1, c["material"]).astype(np.int32)
    ref = O.oracle_render(c["positions"], c["indices"], tm, c["albedo"][:3], c["radiance"], c["camera"], 16, 16, 256, sample_mode=2, seed=99)
    img = sharding.resolve(film)
    mask = ref.max(axis=-1) < 5.0
    assert np.allclose(img[mask].mean(axis=0), ref[mask].mean(axis=0), rtol=0.08)


def test_shard_region_partition():
    for res in ((1920, 1080), (64, 64), (5, 7)):
        for world in (1, 2, 3, 8):
            if world > res[1]:
                continue
            regs = [sharding.shard_region(res, world, r) for r in range(world)]
            assert regs[0][1] == 0 and regs[-1][3] == res[1]
            assert all(a[3] == b[1] for a, b in zip(regs, regs[1:]))                 # contiguous, no overlap
            assert all(r[0] == 0 and r[2] == res[0] for r in regs)
            rows = [r[3] - r[1] for r in regs]
            assert max(rows) - min(rows) <= 1


def _region_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = (12, 9)
    reg = sharding.shard_region(res, world, rank)
    band = np.full((4, reg[3] - reg[1], reg[2] - reg[0]), float(rank + 1), np.float32)   # this rank's region film
    film = sharding.place_region(np.zeros((4, res[1], res[0]), np.float32), band, reg)
    t = torch.from_numpy(film)
    sharding.reduce_film(t)            # disjoint regions: the sum over ranks is the gather
    if rank == 0:
        np.save(os.path.join(out_dir, "regions.npy"), t.numpy())
    dist.destroy_process_group()


def test_two_rank_region_gather(tmp_path):
    world, port = 2, 31000 + os.getpid() % 2000
    mp.spawn(_region_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    film = np.load(tmp_path / "regions.npy")
    r0, r1 = sharding.shard_region((12, 9), 2, 0), sharding.shard_region((12, 9), 2, 1)
    assert np.all(film[:, r0[1]:r0[3]] == 1.0) and np.all(film[:, r1[1]:r1[3]] == 2.0)
