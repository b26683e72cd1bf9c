"""Developer helper: both Sponza ray sets as two launches -- one after the other, or on two streams so that the second
launch fills the tail of the first -- and the host-pointer entry points called from one or two host threads."""
import sys
import threading
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from rodent_b200 import formats, lib, testdata, traversal

L = lib.load()
nodes, tris = formats.load_bvh(testdata.sponza_bvh8())
bvh = traversal.Bvh8(0, nodes, tris)
rays, d_rays, d_hits = {}, {}, {}
for name in ("primary", "random"):
    tmin, tmax = testdata.RAY_SETS[name]
    rays[name] = formats.load_rays(testdata.rays(name), tmin, tmax)
    d_rays[name] = traversal.DeviceArray.from_host(0, rays[name])
    d_hits[name] = traversal.DeviceArray(0, formats.HIT1, len(rays[name]))
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
counters = torch.zeros(64, dtype=torch.int32, device="cuda")
n_rays = sum(len(r) for r in rays.values())


def serial():
    return sum(traversal.intersect(bvh, d_rays[n], d_hits[n]) for n in ("primary", "random"))


def overlapped(order, priorities=(0, 0)):
    streams = [torch.cuda.Stream(priority=p) for p in priorities]
    ev0, ev = torch.cuda.Event(enable_timing=True), [torch.cuda.Event(enable_timing=True) for _ in order]

    def run():
        ev0.record(streams[0])
        streams[1].wait_event(ev0)
        for k, name in enumerate(order):
            traversal.intersect_async(bvh, d_rays[name], d_hits[name], streams[k].cuda_stream, counters.data_ptr() + 32 * k)
            ev[k].record(streams[k])
        torch.cuda.synchronize()
        return max(ev0.elapsed_time(e) for e in ev)
    return run


class Slice:
    """A contiguous part of a device array, for launching a set in pieces."""
    def __init__(self, arr, first, count):
        self.ptr, self.count = arr.ptr + first * arr.dtype.itemsize, count


def pieces(plan):
    """plan: [(set name, first fraction, last fraction), ...], one stream per entry, launched in this order."""
    streams = [torch.cuda.Stream() for _ in plan]
    ev0, ev = torch.cuda.Event(enable_timing=True), [torch.cuda.Event(enable_timing=True) for _ in plan]

    def run():
        ev0.record(streams[0])
        for st in streams[1:]:
            st.wait_event(ev0)
        for k, (name, a, b) in enumerate(plan):
            n = d_rays[name].count
            first, last = int(n * a), int(n * b)
            traversal.intersect_async(bvh, Slice(d_rays[name], first, last - first), Slice(d_hits[name], first, last - first),
                                      streams[k].cuda_stream, counters.data_ptr() + 32 * k)
            ev[k].record(streams[k])
        torch.cuda.synchronize()
        return max(ev0.elapsed_time(e) for e in ev)
    return run


def measure(label, fn, reps=12):
    times = []
    for i in range(reps + 3):
        flush.zero_()
        torch.cuda.synchronize()
        ms = fn()
        if i >= 3:
            times.append(ms)
    print(f"{label:44s} median {np.median(times):.4f} ms  min {min(times):.4f}  -> {n_rays / np.median(times) / 1e3:.0f} Mrays/s", flush=True)


measure("one after the other (sync entry points)", serial)
measure("two streams, random first", overlapped(("random", "primary")))
measure("two streams, primary first", overlapped(("primary", "random")))
measure("random, primary halves", pieces([("random", 0, 1), ("primary", 0, 0.5), ("primary", 0.5, 1)]))
measure("random halves, primary", pieces([("random", 0, 0.5), ("random", 0.5, 1), ("primary", 0, 1)]))
measure("random, primary 3/4 + 1/4", pieces([("random", 0, 1), ("primary", 0, 0.75), ("primary", 0.75, 1)]))
measure("random 1/2, primary, random 1/2", pieces([("random", 0, 0.5), ("primary", 0, 1), ("random", 0.5, 1)]))
measure("random, primary 1/4 x 4", pieces([("random", 0, 1)] + [("primary", k / 4, (k + 1) / 4) for k in range(4)]))
want = {n: d_hits[n].to_host().copy() for n in rays}

# host-pointer entry points
pin_r = {n: traversal.PinnedArray(formats.RAY1, len(rays[n])) for n in rays}
pin_h = {n: traversal.PinnedArray(formats.HIT1, len(rays[n])) for n in rays}
for n in rays:
    pin_r[n].array[:] = rays[n]


def host_serial():
    t0 = time.perf_counter()
    for n in ("primary", "random"):
        traversal.intersect_host(nodes, tris, pin_r[n].array, pin_h[n].array)
    return (time.perf_counter() - t0) * 1e3


def host_threads(order):
    def run():
        ts = [threading.Thread(target=traversal.intersect_host, args=(nodes, tris, pin_r[n].array, pin_h[n].array)) for n in order]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return (time.perf_counter() - t0) * 1e3
    return run


for chunks in (3,):
    lib.tune("host_chunks", chunks)
    measure(f"host pointers, one thread, {chunks} pieces", host_serial)
    measure(f"host pointers, two threads (random first), {chunks}", host_threads(("random", "primary")))
    measure(f"host pointers, two threads (primary first), {chunks}", host_threads(("primary", "random")))
    for n in rays:
        assert pin_h[n].array.tobytes() == want[n].tobytes(), n
print("results identical")
