"""Developer helper: a few traversal passes with given tuning (for ncu).  usage: one_pass.py key=value ... [sets=primary,random] [bvh=8|4|2]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rodent_b200 import formats, lib, testdata, traversal

lib.load()
opts = dict(a.split("=") for a in sys.argv[1:])
for k, v in opts.items():
    if k not in ("passes", "any", "sets", "bvh"):
        lib.tune(k, int(v))
width = int(opts.get("bvh", 8))
nodes, tris = formats.load_bvh({8: testdata.sponza_bvh8, 4: testdata.sponza_bvh4, 2: testdata.sponza_bvh2}[width](),
                               {8: formats.BVH8_TRI4, 4: formats.BVH4_TRI4, 2: formats.BVH2_TRI1}[width])
bvh = traversal.Bvh8(0, nodes, tris)
for name in opts.get("sets", "primary,random").split(","):
    tmin, tmax = testdata.RAY_SETS[name]
    rays = formats.load_rays(testdata.rays(name), tmin, tmax)
    d_rays = traversal.DeviceArray.from_host(0, rays)
    d_hits = traversal.DeviceArray(0, formats.HIT1, len(rays))
    for _ in range(int(opts.get("passes", 3))):
        ms = traversal.intersect(bvh, d_rays, d_hits, any_hit=bool(int(opts.get("any", 0))))
    print(name, f"{len(rays) / ms / 1e3:.1f} Mrays/s")
