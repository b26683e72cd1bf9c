"""Regenerates the committed golden fixtures from the reference tree (run once, in
the dev container where /root/reference exists):

    python tests/golden/make_fixtures.py

* sponza_bvh8.bvh.xz  -- the BVH8_TRI4 block of testing/sponza.bvh re-wrapped as a
                         single-block .bvh file and xz-compressed
* sponza_bvh4.bvh.xz  -- the same for the BVH4_TRI4 block (the reference's default
                         --bvh-width)
* sponza_bvh2.bvh.xz  -- the same for the BVH2_TRI1 block, the layout of the reference's own GPU path
* ref-primary.png, ref-random.png, ref-cornell.png -- the reference's golden images
* cornell_box.obj/.mtl -- the reference's Cornell scene (test input data)
* sponza_hits_sample.npz -- oracle hit records for a fixed sample of rays, so the
                         GPU parity tests also have a frozen known-answer set

The two ray sets are NOT stored: tools/ray_gen regenerates them bit for bit
(tests/test_fixtures.py checks that against the reference files when present).
"""
import lzma
import shutil
import struct
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference/testing")


def main():
    from rodent_b200 import formats as F
    data = (REF / "sponza.bvh").read_bytes()
    for typ, n_nodes, n_tris, off in F.bvh_blocks(data):
        if typ == F.BVH8_TRI4:
            payload = data[off - 20: off + n_nodes * 256 + n_tris * 224]
            blob = struct.pack("<I", F.BVH_MAGIC) + payload
            (HERE / "sponza_bvh8.bvh.xz").write_bytes(lzma.compress(blob, preset=9 | lzma.PRESET_EXTREME))
            print("sponza_bvh8.bvh.xz", n_nodes, n_tris, len(blob))
        if typ == F.BVH4_TRI4:
            payload = data[off - 20: off + n_nodes * 128 + n_tris * 224]
            blob = struct.pack("<I", F.BVH_MAGIC) + payload
            (HERE / "sponza_bvh4.bvh.xz").write_bytes(lzma.compress(blob, preset=9 | lzma.PRESET_EXTREME))
            print("sponza_bvh4.bvh.xz", n_nodes, n_tris, len(blob))
        if typ == F.BVH2_TRI1:
            payload = data[off - 20: off + n_nodes * 64 + n_tris * 48]
            blob = struct.pack("<I", F.BVH_MAGIC) + payload
            (HERE / "sponza_bvh2.bvh.xz").write_bytes(lzma.compress(blob, preset=9 | lzma.PRESET_EXTREME))
            print("sponza_bvh2.bvh.xz", n_nodes, n_tris, len(blob))
    for name in ("ref-primary.png", "ref-random.png", "ref-cornell.png", "cornell_box.obj", "cornell_box.mtl"):
        shutil.copyfile(REF / name, HERE / name)

    # frozen known-answer sample: every 257th ray of both sets, oracle records
    from oracle import oracle
    from rodent_b200 import testdata
    nodes, tris = F.load_bvh(testdata.sponza_bvh8())
    out = {}
    for name, (tmin, tmax) in testdata.RAY_SETS.items():
        rays = np.ascontiguousarray(F.load_rays(testdata.rays(name), tmin, tmax)[::257])
        out[f"{name}_rays"] = rays
        out[f"{name}_hits"] = oracle.traverse(nodes, tris, rays, threads=8)
        out[f"{name}_any"] = oracle.traverse(nodes, tris, rays, any_hit=True, threads=8)["tri_id"] >= 0
    np.savez_compressed(HERE / "sponza_hits_sample.npz", **out)


if __name__ == "__main__":
    main()
