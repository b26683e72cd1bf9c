"""Developer helper: a few recorded path-tracing rays through the BVH2 and BVH8 device kernels and the oracles."""
import sys; sys.path.insert(0, ".")
import numpy as np
from oracle import oracle
from rodent_b200 import formats as F, testdata, traversal
H = """c46800c5 43f1fb23 c1fc5c5d 00000000 3f7d2376 3df5b464 bdb54cc5
44a0b0a4 443c21ca c365b6bd 3a83126f bddb90aa be9935c5 3f72b910
44a0b0a3 443c21c4 c365b671 3a83126f 3eb30db0 3f6f9e5a bd231a4d
44a2b011 4446d374 c3678850 3a83126f bf0e99c7 3f52697a 3df3aed5
44796628 449b7745 c3269dc8 3a83126f 3ee27084 3f65587a 3d2db834"""
rows = np.array([[int(w, 16) for w in line.split()] for line in H.splitlines()], np.uint32).view(np.float32)
rays = np.zeros(len(rows), F.RAY1)
rays["org"], rays["tmin"], rays["dir"], rays["tmax"] = rows[:, :3], rows[:, 3], rows[:, 4:7], np.float32(3.4028234664e+38)
rays = np.tile(rays, 40)           # more than a warp, ragged
n2, t1 = F.load_bvh(testdata.sponza_bvh2(), F.BVH2_TRI1)
n8, t4 = F.load_bvh(testdata.sponza_bvh8(), F.BVH8_TRI4)
for label, nodes, tris, ref in (("bvh2", n2, t1, oracle.traverse_bvh2(n2, t1, rays)), ("bvh8", n8, t4, oracle.traverse(n8, t4, rays))):
    bvh = traversal.Bvh8(0, nodes, tris)
    d_rays = traversal.DeviceArray.from_host(0, rays); d_hits = traversal.DeviceArray(0, F.HIT1, len(rays))
    traversal.intersect(bvh, d_rays, d_hits)
    got = d_hits.to_host()
    print(label, "equal" if got.tobytes() == ref.tobytes() else "DIFFERENT")
    for k in range(5):
        print("   ", got[k], ref[k])
    bad = np.nonzero(got["tri_id"] != ref["tri_id"])[0]
    print("   differing rays:", bad[:20])
