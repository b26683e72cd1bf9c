"""Developer helper: occupancy / refill / streak sweep of the BVH2 kernel on both Sponza sets."""
import sys; sys.path.insert(0, ".")
import numpy as np
from rodent_b200 import formats as F, testdata, traversal, lib
nodes, tris = F.load_bvh(testdata.sponza_bvh2(), F.BVH2_TRI1)
bvh = traversal.Bvh8(0, nodes, tris)
sets = {}
for name, (tmin, tmax) in testdata.RAY_SETS.items():
    rays = F.load_rays(testdata.rays(name), tmin, tmax)
    sets[name] = (traversal.DeviceArray.from_host(0, rays), traversal.DeviceArray(0, F.HIT1, len(rays)), len(rays))
for mb in (8, 10, 12):
    for refill in (24, 16, 8):
        for streak in (8, 4, 16):
            lib.tune("bvh2_min_blocks", mb); lib.tune("refill_min", refill); lib.tune("bvh2_streak_min", streak)
            out = []
            for name, (r, h, n) in sets.items():
                ms = sorted(traversal.intersect(bvh, r, h) for _ in range(9))[4]
                out.append(f"{name} {n/ms/1e3:.0f}")
            print(mb, refill, streak, out, flush=True)
