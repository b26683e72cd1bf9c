"""Developer probe: the packet / hybrid entry points in the reference's packet order (rodent_b200_set_packet_order(1))
against the default (single-ray records), ms per call of 1 Mi rays, host arrays pageable."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from rodent_b200 import formats, lib, testdata, traversal
for arity, loader, block in ((4, testdata.sponza_bvh4, formats.BVH4_TRI4), (8, testdata.sponza_bvh8, formats.BVH8_TRI4)):
    nodes, tris = formats.load_bvh(loader(), block)
    for name, (tmin, tmax) in testdata.RAY_SETS.items():
        rays = formats.load_rays(testdata.rays(name), tmin, tmax)
        for kind, width in (("hybrid", 8), ("packet", 8)):
            packets = formats.pack_rays(rays, width)
            row = []
            for mode in (0, 1):
                lib.load().rodent_b200_set_packet_order(mode)
                for _ in range(2):
                    traversal.intersect_host_packets(nodes, tris, packets, kind)
                ts = []
                for _ in range(6):
                    t0 = time.perf_counter(); traversal.intersect_host_packets(nodes, tris, packets, kind); ts.append((time.perf_counter() - t0) * 1e3)
                row.append(f"{np.median(ts):.2f} ms = {len(rays) / np.median(ts) / 1e3:.0f} Mrays/s")
            lib.load().rodent_b200_set_packet_order(0)
            print(f"bvh{arity} {name:8s} {kind:6s} ray{width}: single-ray records {row[0]}; packet order {row[1]}", flush=True)
