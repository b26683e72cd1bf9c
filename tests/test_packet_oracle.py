"""The reference's packet / hybrid traversal (cpu_traverse_hybrid_helper, src/traversal/mapping_cpu.impala:259-384) restated
in oracle/traversal_oracle.c: traverse_packet.  Pinned the way the reference's own CTest pins those variants
(tools/CMakeLists.txt:26-31 compare every one of them with the same ref-*.png); and the measure of what the packet order
changes against the single-ray kernel -- the records the b200_*_{packet,hybrid}_* entry points return."""
from pathlib import Path

import numpy as np
import pytest
from PIL import Image

from oracle import oracle
from rodent_b200 import formats

GOLDEN = Path(__file__).parent / "golden"
HIT_COUNTS = {"primary": 1_026_430, "random": 959_359}


@pytest.fixture(scope="module")
def sponza4():
    from rodent_b200 import testdata
    return formats.load_bvh(testdata.sponza_bvh4(), formats.BVH4_TRI4)


def ulps(a, b):
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


@pytest.mark.parametrize("kind,width", [("hybrid", 8), ("packet", 8), ("hybrid", 4), ("packet", 4)])
@pytest.mark.parametrize("arity", [8, 4])
def test_packet_oracle_against_the_golden_images_and_the_single_ray_records(kind, width, arity, sponza, sponza4, ray_sets, oracle_hits):
    nodes, tris = sponza if arity == 8 else sponza4
    for name in ("primary", "random"):
        rays = ray_sets[name]
        single = oracle_hits[name] if arity == 8 else oracle.traverse(nodes, tris, rays)
        got = formats.unpack_hits(oracle.traverse_packets(nodes, tris, formats.pack_rays(rays, width), kind))
        assert len(got) == len(rays)
        # the golden image of the reference's CTest for this variant
        ref = np.array(Image.open(GOLDEN / f"ref-{name}.png"))[..., 0]
        assert int((ref != formats.fbuf_to_gray(got["t"]).reshape(1024, 1024)).sum()) <= 2
        assert int((got["tri_id"] >= 0).sum()) == HIT_COUNTS[name]
        # against the single-ray kernel: the same rays hit, at the same distance to the last bit or the next one (a
        # coincident triangle tested later is accepted by `t <= tmax` and stays: mapping_cpu.impala:362-372), another
        # triangle of a tie on up to ~2 % of the incoherent rays
        assert ((got["tri_id"] >= 0) == (single["tri_id"] >= 0)).all()
        assert int(ulps(got["t"], single["t"]).max()) <= 2
        other = float((got["tri_id"] != single["tri_id"]).mean())
        assert other <= (0.001 if name == "primary" else 0.025), other
        same = got["tri_id"] == single["tri_id"]
        assert (got["u"][same] == single["u"][same]).all() and (got["v"][same] == single["v"][same]).all()


def test_packet_oracle_any_hit(sponza, ray_sets, oracle_hits):
    nodes, tris = sponza
    rays = np.ascontiguousarray(ray_sets["random"][:200_000])
    for kind, width in (("hybrid", 8), ("packet", 4)):
        occl = oracle.traverse_packets(nodes, tris, formats.pack_rays(rays, width), kind, any_hit=True)
        n = len(occl) * width
        assert ((formats.unpack_hits(occl)["tri_id"] >= 0) == (oracle_hits["random"][:n]["tri_id"] >= 0)).all()
        assert (occl["t"] == 0).all()            # make_cpu_hit4/8 with any_hit: tri_id only


@pytest.mark.gpu
@pytest.mark.parametrize("kind,width,arity", [("hybrid", 8, 4), ("packet", 8, 8), ("hybrid", 4, 8)])
def test_entry_points_against_the_packet_oracle(kind, width, arity, sponza, sponza4, ray_sets):
    """b200_*_{packet,hybrid}_* return the single-ray kernel's records (tests/test_gpu_traversal.py: bit-exact); against
    the reference's packet order that means: same rays hit, distances within 2 ulp, another triangle only where hits tie."""
    from rodent_b200 import traversal
    nodes, tris = sponza if arity == 8 else sponza4
    for name, limit in (("primary", 0.001), ("random", 0.025)):
        packets = formats.pack_rays(np.ascontiguousarray(ray_sets[name][:400_000]), width)
        want = formats.unpack_hits(oracle.traverse_packets(nodes, tris, packets, kind))
        got = formats.unpack_hits(traversal.intersect_host_packets(nodes, tris, packets, kind))
        assert ((got["tri_id"] >= 0) == (want["tri_id"] >= 0)).all()
        assert int(ulps(got["t"], want["t"]).max()) <= 2
        assert float((got["tri_id"] != want["tri_id"]).mean()) <= limit


@pytest.mark.gpu
@pytest.mark.parametrize("kind,width,arity", [("hybrid", 8, 4), ("packet", 8, 4), ("hybrid", 4, 4), ("packet", 4, 8), ("hybrid", 8, 8), ("packet", 8, 8), ("hybrid", 4, 8)])
def test_packet_order_mode_is_the_packet_oracle_to_the_bit(kind, width, arity, sponza, sponza4, ray_sets):
    """rodent_b200_set_packet_order(1): traverse_packets_ordered walks in the reference's packet order -- every record
    of both sets' first 300 000 rays identical to the restated packet kernel, closest and any hit, degenerate rays too."""
    from rodent_b200 import lib, traversal
    nodes, tris = sponza if arity == 8 else sponza4
    lib.load().rodent_b200_set_packet_order(1)
    try:
        for name in ("primary", "random"):
            rays = np.ascontiguousarray(ray_sets[name][:300_000])
            if name == "random":          # axis-parallel and clamped directions: the NaN paths of the slab test
                rays["dir"][::97, 0] = 0.0; rays["dir"][::89, 1] = -0.0; rays["dir"][::83, 2] = 1e-12
            packets = formats.pack_rays(rays, width)
            want = oracle.traverse_packets(nodes, tris, packets, kind)
            got = traversal.intersect_host_packets(nodes, tris, packets, kind)
            assert lib.load().rodent_b200_last_kernel_name(0).decode() == "traverse_packets_ordered"
            assert got.tobytes() == want.tobytes(), (name, int((formats.unpack_hits(got)["tri_id"] != formats.unpack_hits(want)["tri_id"]).sum()))
            pre = np.zeros(len(packets), formats.packet_dtypes(width)[1])
            pre["t"] = 5.0
            occl = traversal.intersect_host_packets(nodes, tris, packets, kind, any_hit=True, hits=pre)
            want_any = oracle.traverse_packets(nodes, tris, packets, kind, any_hit=True)
            assert np.array_equal(occl["tri_id"], want_any["tri_id"]) and (occl["t"] == 5.0).all()
    finally:
        lib.load().rodent_b200_set_packet_order(0)
