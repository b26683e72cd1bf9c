"""Developer helper: a few traversal passes with given tuning (for ncu).  usage: one_pass.py key=value ... [sets=primary,random]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rodent_b200 import formats, lib, testdata, traversal

lib.load()
opts = dict(a.split("=") for a in sys.argv[1:])
for k, v in opts.items():
    if k not in ("passes", "any", "sets"):
        lib.tune(k, int(v))
nodes, tris = formats.load_bvh(testdata.sponza_bvh8())
bvh = traversal.Bvh8(0, nodes, tris)
for name in opts.get("sets", "primary,random").split(","):
    tmin, tmax = testdata.RAY_SETS[name]
    rays = formats.load_rays(testdata.rays(name), tmin, tmax)
    d_rays = traversal.DeviceArray.from_host(0, rays)
    d_hits = traversal.DeviceArray(0, formats.HIT1, len(rays))
    for _ in range(int(opts.get("passes", 3))):
        ms = traversal.intersect(bvh, d_rays, d_hits, any_hit=bool(int(opts.get("any", 0))))
    print(name, f"{len(rays) / ms / 1e3:.1f} Mrays/s")
