"""Developer helper: kernel time of small launches (as the host-pointer path issues them) vs tuning."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from rodent_b200 import formats, lib, testdata, traversal
L = lib.load()
nodes, tris = formats.load_bvh(testdata.sponza_bvh8())
bvh = traversal.Bvh8(0, nodes, tris)
for name, (tmin, tmax) in testdata.RAY_SETS.items():
    rays = formats.load_rays(testdata.rays(name), tmin, tmax)
    for n in (1 << 20, 1 << 19, 1 << 18, 1 << 17):
        d_rays = traversal.DeviceArray.from_host(0, np.ascontiguousarray(rays[:n]))
        d_hits = traversal.DeviceArray(0, formats.HIT1, n)
        line = [f"{name} n={n:8d}"]
        for cfg in (dict(refill_min=24, blocks_per_sm=0), dict(refill_min=16, blocks_per_sm=0), dict(refill_min=8, blocks_per_sm=0),
                    dict(refill_min=24, blocks_per_sm=3), dict(refill_min=8, blocks_per_sm=3), dict(refill_min=8, blocks_per_sm=2)):
            for k, v in cfg.items():
                lib.tune(k, v)
            for _ in range(3):
                traversal.intersect(bvh, d_rays, d_hits)
            ms = float(np.median([traversal.intersect(bvh, d_rays, d_hits) for _ in range(10)]))
            line.append(f"rm{cfg['refill_min']}/b{cfg['blocks_per_sm']}: {ms*1e3:6.0f} us ({n/ms/1e3:6.0f} M/s)")
        print(" | ".join(line), flush=True)
