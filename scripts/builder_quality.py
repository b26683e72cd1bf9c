"""Developer tool (CPU): quality of this repository's BVH builder against the reference's own tree.
Builds a BVH4 over the Sponza triangles (recovered from the reference's sponza.bvh) with rodent_b200_scene_bvh4 and
counts, with the oracle, the inner nodes and Tri4 packets a ray visits -- next to the same counters on the BVH4 block of
the reference's file."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from oracle import oracle
from rodent_b200 import formats, testdata, workloads


def visits(nodes, tris, rays):
    _, st = oracle.traverse(nodes, tris, rays, want_stats=True)
    return st.nodes / len(rays), st.tri4 / len(rays)


def main():
    scene = workloads.load_scene("sponza")
    t0 = time.perf_counter()
    mine = scene.bvh4()
    dt = time.perf_counter() - t0
    ref = formats.load_bvh(testdata.sponza_bvh4(), formats.BVH4_TRI4)
    print(f"built in {dt:.2f} s: {len(mine[0])} Node4, {len(mine[1])} Tri4 (reference file: {len(ref[0])}, {len(ref[1])})")
    for name, (tmin, tmax) in testdata.RAY_SETS.items():
        rays = formats.load_rays(testdata.rays(name), tmin, tmax)[::4].copy()
        a, b = visits(*mine, rays), visits(*ref, rays)
        want = oracle.traverse(*ref, rays)
        got = oracle.traverse(*mine, rays)
        same_t = (got["t"] == want["t"]).mean()
        print(f"{name:8s} nodes/ray {a[0]:6.2f} vs {b[0]:6.2f} ({(a[0] / b[0] - 1) * 100:+.1f} %)   Tri4/ray {a[1]:5.2f} vs {b[1]:5.2f} ({(a[1] / b[1] - 1) * 100:+.1f} %)"
              f"   cost 128*nodes+224*tri4: {(128 * a[0] + 224 * a[1]) / (128 * b[0] + 224 * b[1]) - 1:+.1%}   same t: {same_t:.6f}")


if __name__ == "__main__":
    main()
