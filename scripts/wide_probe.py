"""Developer helper: 128-bit against 256-bit record loads in the default BVH8 kernel, both Sponza sets, closest and any hit."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from oracle import oracle
from rodent_b200 import formats, lib, testdata, traversal
nodes, tris = formats.load_bvh(testdata.sponza_bvh8())
bvh = traversal.Bvh8(0, nodes, tris)
for name, (tmin, tmax) in testdata.RAY_SETS.items():
    rays = formats.load_rays(testdata.rays(name), tmin, tmax)
    want = oracle.traverse(nodes, tris, rays)
    d_rays = traversal.DeviceArray.from_host(0, rays); d_hits = traversal.DeviceArray(0, formats.HIT1, len(rays))
    for wide in (24, 16, 12, 24, 16, 12):
        lib.tune("vote_smem_depth", wide)
        out = []
        for any_hit in (False, True):
            ts = sorted(traversal.intersect(bvh, d_rays, d_hits, any_hit=any_hit) for _ in range(13))
            out.append(f"{'any' if any_hit else 'closest'} {len(rays) / ts[6] / 1e3:.0f} Mrays/s")
        traversal.intersect(bvh, d_rays, d_hits)
        ok = d_hits.to_host().tobytes() == want.tobytes()
        print(name, "smem depth", wide, out, "bit-exact" if ok else "DIFFERENT", flush=True)
