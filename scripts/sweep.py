"""Developer sweep: times the traversal kernel variants on both Sponza ray sets."""
import sys, itertools, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from rodent_b200 import formats, lib, testdata, traversal

def main():
    L = lib.load()
    nodes, tris = formats.load_bvh(testdata.sponza_bvh8())
    bvh = traversal.Bvh8(0, nodes, tris)
    sets = {}
    for name, (tmin, tmax) in testdata.RAY_SETS.items():
        rays = formats.load_rays(testdata.rays(name), tmin, tmax)
        sets[name] = (traversal.DeviceArray.from_host(0, rays), traversal.DeviceArray(0, formats.HIT1, len(rays)))
    configs = []
    for ns in (33, 8):
        configs.append(dict(mapping=2, refill_min=24, node_streak_min=ns))
    extra = [a for a in sys.argv[1:]]
    for cfg in configs:
        for k, v in cfg.items():
            lib.tune(k, v)
        line = [f"{cfg}"]
        for any_hit in (False, True):
            for name, (d_rays, d_hits) in sets.items():
                for _ in range(3):
                    traversal.intersect(bvh, d_rays, d_hits, any_hit=any_hit)
                ts = [traversal.intersect(bvh, d_rays, d_hits, any_hit=any_hit) for _ in range(10)]
                ms = float(np.median(ts))
                line.append(f"{name}{'-any' if any_hit else ''}: {d_rays.count / ms / 1e3:8.1f} Mrays/s ({ms:.3f} ms, min {min(ts):.3f})")
        print(" | ".join(line), flush=True)

if __name__ == "__main__":
    main()
