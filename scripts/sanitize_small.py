"""Developer helper (run under compute-sanitizer): small traversals through every kernel family and a tiny textured render."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from rodent_b200 import formats as F, lib, render as R, testdata, traversal, workloads
rng = np.random.default_rng(0)
for loader, typ in ((testdata.sponza_bvh8, F.BVH8_TRI4), (testdata.sponza_bvh4, F.BVH4_TRI4), (testdata.sponza_bvh2, F.BVH2_TRI1)):
    nodes, tris = F.load_bvh(loader(), typ)
    bvh = traversal.Bvh8(0, nodes, tris)
    for name, (tmin, tmax) in testdata.RAY_SETS.items():
        rays = np.ascontiguousarray(F.load_rays(testdata.rays(name), tmin, tmax)[::97])
        d_rays = traversal.DeviceArray.from_host(0, rays); d_hits = traversal.DeviceArray(0, F.HIT1, len(rays))
        for any_hit in (False, True):
            traversal.intersect(bvh, d_rays, d_hits, any_hit=any_hit)
        print(typ, name, int((d_hits.to_host()["tri_id"] >= 0).sum()), flush=True)
    if typ == F.BVH8_TRI4:
        sub = np.ascontiguousarray(F.load_rays(testdata.rays("random"), 0.0, 1.0)[:40000])
        traversal.intersect_host(nodes, tris, sub)
        # pinned buffers: the direct path (armed slots, records sent home by groups), twice through the same context,
        # with the copy-engine pieces in between
        pr, ph = traversal.PinnedArray(F.RAY1, len(sub)), traversal.PinnedArray(F.HIT1, len(sub))
        pr.array[:] = sub
        want = traversal.intersect_host(nodes, tris, sub)
        for n in (len(sub), 1000, 17):
            ph.array[:] = 0
            traversal.intersect_host(nodes, tris, pr.array[:n], ph.array[:n])
            assert ph.array[:n].tobytes() == want[:n].tobytes()
            traversal.intersect_host(nodes, tris, sub[:n])
        print("direct host-pointer path", lib.load().rodent_b200_last_kernel_name(0).decode(), flush=True)
        for mapping in (1, 3, 4):
            lib.tune("mapping", mapping)
            traversal.intersect(bvh, d_rays, d_hits)
        lib.tune("mapping", 2)
scene = workloads.load_scene("sponza")
cam = workloads.camera("sponza", 96, 64)
r = R.Renderer(scene, 0, 96, 64, 2, 6)
r.render(cam, 0)
print("sponza film", float(r.film().mean()))
r.free()
cornell = workloads.load_scene("cornell")
r = R.Renderer(cornell, 0, 64, 64, 2, 6)
r.render(workloads.camera("cornell", 64, 64), 0)
print("cornell film", float(r.film().mean()))
r.free()
# several pipelines (device-driven wavefront loop, two-level scatter) and a stream smaller than the image: many wavefronts
lib.tune("render_capacity", 4096)
r = R.Renderer(cornell, 0, 96, 128, 3, 5)
for it in range(2):
    r.render(workloads.camera("cornell", 96, 128), it)
print("cornell film, 3 pipelines, 4096-ray streams", float(r.film().mean()), r.stats())
r.free()
lib.tune("render_capacity", 1 << 21)
