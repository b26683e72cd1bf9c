"""Multi-GPU sharding of the two hot paths: one process per GPU, no data-path collective.

The reference is single-process, single-device; this is new.  Rays are
independent (cpu_traverse_single, src/traversal/mapping_cpu.impala:412-417) and every camera
sample is a pure function of (sample, iter, x, y) (src/render/renderer.impala:28-33), so

  * a traversal job is cut into contiguous ray ranges, one per rank, BVH replicated;
  * a render job deals the image rows out in bands of `band` rows, round robin, each rank runs
    its own wavefront loop, scene replicated;

and the only communication is ONE collective at the end: the hit records are gathered, or the
films -- disjoint rows, zeros elsewhere -- are summed onto rank 0.  torch.distributed is the
plumbing (NCCL over NVLink on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def ray_range(rank: int, world: int, num_rays: int) -> tuple[int, int]:
    """[begin, end) of the rays rank `rank` traces; ranges tile [0, num_rays) in rank order."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return (num_rays * rank) // world, (num_rays * (rank + 1)) // world


def rows_of(part: int, num_parts: int, height: int, band: int = 8) -> np.ndarray:
    """Image rows owned by `part`: (y // band) % num_parts == part -- the rule of
    rodent_b200_renderer_create (include/rodent_b200.h)."""
    if not (0 <= part < num_parts) or band <= 0:
        raise ValueError("bad partition")
    y = np.arange(height)
    return y[(y // band) % num_parts == part]


def gather_hits(local_hits: np.ndarray, rank: int, world: int, num_rays: int, device=None):
    """Gathers the per-rank hit records (a HIT1 array covering ray_range(rank)) on rank 0, in ray
    order; other ranks get None.  One gather of fixed-size, padded buffers."""
    import torch
    import torch.distributed as dist
    begin, end = ray_range(rank, world, num_rays)
    assert len(local_hits) == end - begin
    longest = max(ray_range(r, world, num_rays)[1] - ray_range(r, world, num_rays)[0] for r in range(world))
    buf = np.zeros((longest, 4), np.int32)
    buf[:len(local_hits)] = np.ascontiguousarray(local_hits).view(np.int32).reshape(-1, 4)
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, out, dst=0)
    if rank != 0:
        return None
    parts = []
    for r in range(world):
        b, e = ray_range(r, world, num_rays)
        parts.append(out[r][: e - b].cpu().numpy())
    return np.concatenate(parts).view(local_hits.dtype).reshape(-1)


def reduce_film(film, dst: int = 0):
    """Sums the per-rank films (torch tensors, rows outside a rank's bands are zero) onto rank `dst`,
    in place.  ncclReduce on the GPUs; gloo in the CPU tests."""
    import torch.distributed as dist
    dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
    return film
