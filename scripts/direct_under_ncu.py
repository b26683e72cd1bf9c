"""Developer helper for ncu: the host-pointer entry point with page-locked arrays (traverse_direct), both Sponza sets, a few calls."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rodent_b200 import formats, lib, testdata, traversal
nodes, tris = formats.load_bvh(testdata.sponza_bvh8())
for name, (tmin, tmax) in testdata.RAY_SETS.items():
    rays = formats.load_rays(testdata.rays(name), tmin, tmax)
    pr, ph = traversal.PinnedArray(formats.RAY1, len(rays)), traversal.PinnedArray(formats.HIT1, len(rays))
    pr.array[:] = rays
    for _ in range(3):
        traversal.intersect_host(nodes, tris, pr.array, ph.array)
    print(name, lib.load().rodent_b200_last_kernel_name(0).decode(), int((ph.array["tri_id"] >= 0).sum()), flush=True)
