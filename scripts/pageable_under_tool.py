"""Developer helper (run under ncu / compute-sanitizer): one host-pointer call with pageable arrays; prints the kernel it took."""
import sys; sys.path.insert(0, '.')
import numpy as np
from rodent_b200 import formats, lib, testdata, traversal
nodes, tris = formats.load_bvh(testdata.sponza_bvh8())
rays = formats.load_rays(testdata.rays("random"), 0.0, 1.0)[:200000].copy()
h = traversal.intersect_host(nodes, tris, rays)
print("pageable call under the tool:", lib.load().rodent_b200_last_kernel_name(0).decode(), int((h["tri_id"] >= 0).sum()), flush=True)
