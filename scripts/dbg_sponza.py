"""Developer helper: small Sponza render through the BVH2 and the BVH8 closest-hit kernels, films saved for comparison."""
import sys; sys.path.insert(0, ".")
import numpy as np
from rodent_b200 import render as R, workloads, lib
scene = workloads.load_scene("sponza")
W, H, spp = 192, 108, 2
cam = workloads.camera("sponza", W, H)
out = {}
for use in (1, 0):
    lib.tune("render_bvh2", use)
    for depth in (3, 4, 8):
        for lanes in (1, 3):
            lib.tune("render_lanes", lanes)
            r = R.Renderer(scene, 0, W, H, spp, depth)
            r.render(cam, 0)
            out[f"bvh2_{use}_depth{depth}_lanes{lanes}"] = r.film().copy()
            r.free()
np.savez_compressed("gpurun_out/dbg_sponza_films.npz", **out)
print({k: float(v.mean()) for k, v in out.items()})
