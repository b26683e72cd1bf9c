"""Host-side mirror of the reference's bench_traversal interface
(tools/bench_traversal/bench_traversal.cpp:44-135, tools/common/load_bvh.h,
load_rays.h): device arrays, BVH/ray upload, and the intersect/occluded calls.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import formats, lib


class DeviceArray:
    """A typed device allocation (the role anydsl::Array<T>(platform, device, n) plays
    in tools/common/load_bvh.h:64-65)."""

    def __init__(self, dev: int, dtype: np.dtype, count: int):
        self.dev, self.dtype, self.count = dev, np.dtype(dtype), int(count)
        self.nbytes = self.dtype.itemsize * self.count
        self.ptr = lib.load().rodent_b200_alloc_device(dev, self.nbytes)

    @classmethod
    def from_host(cls, dev: int, host: np.ndarray) -> "DeviceArray":
        host = np.ascontiguousarray(host)
        arr = cls(dev, host.dtype, len(host))
        if arr.nbytes:
            lib.load().rodent_b200_copy_to_device(dev, arr.ptr, host.ctypes.data, arr.nbytes)
        return arr

    def to_host(self) -> np.ndarray:
        out = np.empty(self.count, self.dtype)
        if self.nbytes:
            lib.load().rodent_b200_copy_to_host(self.dev, out.ctypes.data, self.ptr, self.nbytes)
        return out

    def free(self) -> None:
        if self.ptr:
            lib.load().rodent_b200_free_device(self.dev, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Bvh8:
    """BVH8/Tri4 -- or BVH4/Tri4, or the reference GPU path's BVH2/Tri1 -- resident on one device
    (load_bvh<Node, Tri>, load_bvh.h:46-74)."""

    def __init__(self, dev: int, nodes: np.ndarray, tris: np.ndarray):
        assert (nodes.dtype in (formats.NODE8, formats.NODE4) and tris.dtype == formats.TRI4) or \
               (nodes.dtype == formats.NODE2 and tris.dtype == formats.TRI1)
        self.arity = {formats.NODE8: 8, formats.NODE4: 4, formats.NODE2: 2}[nodes.dtype]
        self.dev = dev
        self.nodes = DeviceArray.from_host(dev, nodes)
        self.tris = DeviceArray.from_host(dev, tris)

    @classmethod
    def load(cls, dev: int, path) -> "Bvh8":
        return cls(dev, *formats.load_bvh(path, formats.BVH8_TRI4))


def intersect(bvh: Bvh8, rays: DeviceArray, hits: DeviceArray, any_hit: bool = False, count: int | None = None) -> float:
    """One synchronous traversal pass over device-resident rays; returns the kernel
    time in ms (bench_gpu, bench_traversal.cpp:124-135)."""
    L = lib.load()
    n = rays.count if count is None else count
    fn = getattr(L, f"cuda_{'occluded' if any_hit else 'intersect'}_single_ray1_bvh{bvh.arity}_tri{1 if bvh.arity == 2 else 4}")
    fn(bvh.dev, bvh.nodes.ptr, bvh.tris.ptr, rays.ptr, hits.ptr, n)
    return L.rodent_b200_last_kernel_ms(bvh.dev)


def intersect_async(bvh: Bvh8, rays: DeviceArray, hits: DeviceArray, stream: int, work_counter: int, any_hit: bool = False,
                    count: int | None = None) -> None:
    """Enqueue one traversal pass on `stream` (a cudaStream_t as an integer) without synchronising
    (cuda_*_single_ray1_bvh8_tri4_async).  `work_counter`: device address of an int32 owned by the caller, one per stream
    in flight."""
    assert bvh.arity == 8
    L = lib.load()
    n = rays.count if count is None else count
    fn = getattr(L, f"cuda_{'occluded' if any_hit else 'intersect'}_single_ray1_bvh8_tri4_async")
    fn(bvh.dev, bvh.nodes.ptr, bvh.tris.ptr, rays.ptr, hits.ptr, n, stream, work_counter)


def intersect_host(nodes: np.ndarray, tris: np.ndarray, rays: np.ndarray, hits: np.ndarray | None = None,
                   any_hit: bool = False) -> np.ndarray:
    """The host-buffer drop-in for cpu_{intersect,occluded}_single_ray1_bvh8_tri4
    (bench_cpu_single, bench_traversal.cpp:76-82): numpy in, numpy out."""
    L = lib.load()
    if hits is None:
        hits = np.zeros(len(rays), formats.HIT1)
    arity = 8 if nodes.dtype == formats.NODE8 else 4
    fn = getattr(L, f"b200_{'occluded' if any_hit else 'intersect'}_single_ray1_bvh{arity}_tri4")
    fn(nodes.ctypes.data, tris.ctypes.data, rays.ctypes.data, hits.ctypes.data, len(rays))
    return hits


def pin_host(arr: np.ndarray) -> bool:
    """Page-lock a numpy array in place (rodent_b200_pin_host): host-pointer calls on it then take the direct path."""
    return lib.load().rodent_b200_pin_host(arr.ctypes.data, arr.nbytes) == 0


def unpin_host(arr: np.ndarray) -> bool:
    return lib.load().rodent_b200_unpin_host(arr.ctypes.data) == 0


class PinnedArray:
    """Page-locked host array viewed as numpy (for the end-to-end path)."""

    def __init__(self, dtype: np.dtype, count: int):
        self.dtype, self.count = np.dtype(dtype), int(count)
        self.nbytes = max(self.dtype.itemsize * self.count, 16)
        self.ptr = lib.load().rodent_b200_alloc_host(self.nbytes)
        buf = (ctypes.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, self.dtype, self.count)

    def free(self) -> None:
        if self.ptr:
            self.array = None
            lib.load().rodent_b200_free_host(self.ptr)
            self.ptr = None


def intersect_host_packets(nodes: np.ndarray, tris: np.ndarray, packets: np.ndarray, kind: str = "hybrid", any_hit: bool = False,
                           hits: np.ndarray | None = None) -> np.ndarray:
    """The host-buffer drop-in for cpu_{intersect,occluded}_{packet,hybrid}_ray{4,8}_bvh{4,8}_tri4
    (bench_cpu_packet / bench_cpu_hybrid, bench_traversal.cpp:44-74,84-122)."""
    L = lib.load()
    width = packets.dtype["tmin"].shape[0]
    arity = 8 if nodes.dtype == formats.NODE8 else 4
    if hits is None:
        hits = np.zeros(len(packets), formats.packet_dtypes(width)[1])
    fn = getattr(L, f"b200_{'occluded' if any_hit else 'intersect'}_{kind}_ray{width}_bvh{arity}_tri4")
    fn(nodes.ctypes.data, tris.ctypes.data, packets.ctypes.data, hits.ctypes.data, len(packets))
    return hits
