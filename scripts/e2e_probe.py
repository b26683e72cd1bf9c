"""Developer helper: times the host-pointer entry point (pinned host buffers) for several chunk counts."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from rodent_b200 import formats, lib, testdata, traversal
L = lib.load()
nodes, tris = formats.load_bvh(testdata.sponza_bvh8())
sets = {}
for name, (tmin, tmax) in testdata.RAY_SETS.items():
    rays = formats.load_rays(testdata.rays(name), tmin, tmax)
    pr = traversal.PinnedArray(formats.RAY1, len(rays)); pr.array[:] = rays
    ph = traversal.PinnedArray(formats.HIT1, len(rays))
    sets[name] = (pr, ph)
for chunks in (3, 4, 5, 6, 8):
    lib.tune("host_chunks", chunks)
    for _ in range(3):
        for pr, ph in sets.values():
            traversal.intersect_host(nodes, tris, pr.array, ph.array)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter()
        for pr, ph in sets.values():
            traversal.intersect_host(nodes, tris, pr.array, ph.array)
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"chunks {chunks:2d}: median {np.median(ts):.3f} ms, min {min(ts):.3f} ms -> {2 * (1 << 20) / np.median(ts) / 1e3:.0f} Mrays/s", flush=True)
