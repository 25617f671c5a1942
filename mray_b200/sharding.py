"""Multi-GPU sharding of the path: paths are independent, the scene + BVH are
replicated, work is split by sample range (and by tile for multi-tile images); the only shared state
is the film, summed over ranks (NCCL all-reduce over NVLink on GPUs, gloo in CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_samples(total_spp: int, world: int, rank: int) -> tuple[int, int]:
    """[begin, end) of the sample indices rank renders: contiguous, sizes differ by at most one."""
    base, extra = divmod(total_spp, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_tiles(tile_count: int, world: int, rank: int) -> list[int]:
    """Round-robin tile ownership (4K = four 1080p tiles, Tracer/RenderImage.cpp:L20-52)."""
    return list(range(rank, tile_count, world))


def shard_region(resolution: tuple[int, int], world: int, rank: int) -> tuple[int, int, int, int]:
    """(minX, minY, maxX, maxY) of the horizontal band of the image rank renders — the reference's
    RenderImageParams{resolution, regionMin, regionMax} hook (Core/TracerI.h:L38-43; mrb_render_desc.fullResolution /
    regionMin): bands are contiguous, cover the image exactly once and differ by at most one row. With region
    sharding every rank renders ALL samples of its own pixels, so the film needs a gather, not a sum."""
    w, h = resolution
    base, extra = divmod(h, world)
    y0 = rank * base + min(rank, extra)
    return 0, y0, w, y0 + base + (1 if rank < extra else 0)


def place_region(full_film: np.ndarray, region_film: np.ndarray, region: tuple[int, int, int, int]) -> np.ndarray:
    """Writes a rank's (4, h, w) region film into the (4, H, W) image film (RenderImageSection.pixelMin / pixelMax)."""
    x0, y0, x1, y1 = region
    full_film[:, y0:y1, x0:x1] = region_film
    return full_film


def rank_seed(seed: int, rank: int) -> int:
    """Per-rank RNG seed: ranks must draw independent streams (statistical, not bitwise, equivalence
    across world sizes)."""
    return (seed * 0x9E3779B97F4A7C15 + rank * 0xD1B54A32D192ED03 + 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF


def reduce_film(film, group=None):
    """Sum the planar (R,G,B,W) film over ranks in place. `film` is a torch tensor (CUDA -> NCCL,
    CPU -> gloo). No-op without an initialised process group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(film, op=dist.ReduceOp.SUM, group=group)
    return film


def resolve(film_rgbw: np.ndarray) -> np.ndarray:
    """(4,h,w) sums -> (h,w,3) image = sum radiance / sum filter weight (MRay/RunCommand.cpp:L293-345)."""
    return np.moveaxis(film_rgbw[:3], 0, -1) / np.maximum(film_rgbw[3], 1e-20)[..., None]


def max_over_ranks(value: float, device=None, group=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
