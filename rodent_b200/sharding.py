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


def pin_to_gpu_numa_node(device_index: int) -> dict:
    """Binds the calling process to the CPUs of the NUMA node the GPU hangs off (Linux sysfs; best effort), so that the
    host threads that feed the GPU run next to it and the pinned staging memory they allocate afterwards is local to it
    (first touch).  With eight ranks on one box, leaving everything on node 0 sends every PCIe transfer through one
    socket's memory controllers (round 1: 8 ranks x 96 MB per step topped out at 157 GB/s).  Returns what was done."""
    import os
    info = {"numa_node": None, "cpus": None, "bound": False}
    try:
        import torch
        prop = torch.cuda.get_device_properties(device_index)
        bdf = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        info["numa_node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"], info["bound"] = len(allowed), True
    except Exception as exc:                      # no sysfs entry, container without the node files, ...
        info["error"] = f"{type(exc).__name__}: {exc}"
    return info


def gather_hits_device(local_hits, out_list, dst: int = 0, async_op: bool = False):
    """The same gather on device tensors (NCCL): `local_hits` an (n, 4) int32 tensor holding this rank's Hit1 records,
    `out_list` on rank `dst` a list of world tensors of that shape (None elsewhere).  Every rank passes the same n."""
    import torch.distributed as dist
    return dist.gather(local_hits, out_list, dst=dst, async_op=async_op)


class PeerBuffer:
    """A device buffer on rank `dst` that every rank can write: allocated there, exported with CUDA IPC, opened on the
    other ranks as NVLink peer memory (rodent_b200_ipc_*).  `ptr` is the address of the buffer in THIS process; None on
    every rank when IPC / peer access is not available (the caller then gathers with a collective instead)."""

    def __init__(self, local_device: int, rank: int, nbytes: int, dst: int = 0):
        import ctypes
        import torch.distributed as dist
        from . import lib
        L = lib.load()
        self.L, self.dev, self.rank, self.dst, self.nbytes = L, local_device, rank, dst, nbytes
        self.ptr, self.owner = None, rank == dst
        handle = None
        if self.owner:
            self.ptr = L.rodent_b200_alloc_device(local_device, nbytes)
            buf = ctypes.create_string_buffer(64)
            handle = bytes(buf.raw) if L.rodent_b200_ipc_export(local_device, self.ptr, buf) else None
        box = [handle]
        dist.broadcast_object_list(box, src=dst)
        ok = box[0] is not None
        if ok and not self.owner:
            self.ptr = L.rodent_b200_ipc_open(local_device, ctypes.create_string_buffer(box[0], 64))
            ok = bool(self.ptr)
        flags = [None] * dist.get_world_size()
        dist.all_gather_object(flags, ok)
        if not all(flags):
            self.close()
            self.ptr = None

    def close(self):
        if self.ptr:
            if self.owner:
                self.L.rodent_b200_free_device(self.dev, self.ptr)
            else:
                self.L.rodent_b200_ipc_close(self.dev, self.ptr)
            self.ptr = None
