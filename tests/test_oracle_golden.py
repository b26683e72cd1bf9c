"""Pins the CPU oracle (oracle/traversal_oracle.c) against the reference's golden
vectors for the traversal path and against its own cross-checks."""
import hashlib
from pathlib import Path

import numpy as np
import pytest
from PIL import Image

from conftest import REFERENCE
from oracle import oracle
from rodent_b200 import formats, testdata

GOLDEN = Path(__file__).parent / "golden"

# Known answers of the reference commands README.md:33-36 (tmin 0; primary tmax 5000, random tmax 1)
HIT_COUNTS = {"primary": 1_026_430, "random": 959_359}
# pixels (of 1 Mi) allowed to differ from the golden PNG: the reference binary is built
# -ffast-math, so a couple of rays sit on a rounding boundary of the 8-bit quantisation
PNG_TOLERANCE = {"primary": 2, "random": 1}


@pytest.mark.parametrize("name", ["primary", "random"])
def test_hit_distance_matches_golden_png(name, oracle_hits):
    hits = oracle_hits[name]
    ref = np.array(Image.open(GOLDEN / f"ref-{name}.png"))
    assert ref.shape == (1024, 1024, 4)
    gray = formats.fbuf_to_gray(hits["t"]).reshape(1024, 1024)   # fbuf2png -n
    differ = int((ref[..., 0] != gray).sum())
    assert differ <= PNG_TOLERANCE[name], f"{differ} pixels differ from ref-{name}.png"
    assert (ref[..., 3] == 255).all()
    assert int((hits["tri_id"] >= 0).sum()) == HIT_COUNTS[name]


def test_ctest_settings_tmin(sponza, ray_sets):
    # cmake/test/run_traversal.cmake:1 uses --tmin 0.01 --tmax 5000 for the same PNG
    nodes, tris = sponza
    rays = ray_sets["primary"].copy()
    rays["tmin"] = np.float32(0.01)
    hits = oracle.traverse(nodes, tris, rays)
    ref = np.array(Image.open(GOLDEN / "ref-primary.png"))[..., 0]
    assert int((ref != formats.fbuf_to_gray(hits["t"]).reshape(1024, 1024)).sum()) <= 2


@pytest.mark.parametrize("name", ["primary", "random"])
def test_frozen_sample(name, sponza):
    """Committed known-answer records (tests/golden/make_fixtures.py)."""
    nodes, tris = sponza
    z = np.load(GOLDEN / "sponza_hits_sample.npz")
    rays = z[f"{name}_rays"]
    hits = oracle.traverse(nodes, tris, rays)
    assert hits.tobytes() == z[f"{name}_hits"].tobytes()
    occl = oracle.traverse(nodes, tris, rays, any_hit=True)
    assert ((occl["tri_id"] >= 0) == z[f"{name}_any"]).all()


@pytest.mark.parametrize("name", ["primary", "random"])
def test_any_hit_agrees_with_closest_hit(name, sponza, ray_sets, oracle_hits):
    nodes, tris = sponza
    occl = oracle.traverse(nodes, tris, ray_sets[name], any_hit=True)
    assert ((occl["tri_id"] >= 0) == (oracle_hits[name]["tri_id"] >= 0)).all()
    # occluded writes tri_id only (make_cpu_hit1 any_hit, bench_traversal.impala:121-131)
    assert (occl["t"] == 0).all() and (occl["u"] == 0).all() and (occl["v"] == 0).all()


def test_misses_report_tmax(oracle_hits, ray_sets):
    for name, hits in oracle_hits.items():
        miss = hits["tri_id"] < 0
        assert (hits["tri_id"][miss] == -1).all()
        assert (hits["t"][miss] == ray_sets[name]["tmax"][miss]).all()
        hit = ~miss
        assert (hits["t"][hit] <= ray_sets[name]["tmax"][hit]).all()
        assert (hits["u"][hit] >= 0).all() and (hits["v"][hit] >= 0).all()
        assert (hits["u"][hit] + hits["v"][hit] <= 1.0 + 1e-6).all()


def test_brute_force_agrees(sponza, ray_sets, oracle_hits):
    """All-triangles search on a sample: same t everywhere; same triangle unless tied."""
    nodes, tris = sponza
    for name in ("primary", "random"):
        idx = np.arange(0, 1 << 20, 4099)
        rays = np.ascontiguousarray(ray_sets[name][idx])
        bf = oracle.brute_force(tris, rays)
        tr = oracle_hits[name][idx]
        assert (bf["t"] == tr["t"]).all()
        same = bf["tri_id"] == tr["tri_id"]
        assert same.mean() > 0.99
        assert ((bf["tri_id"] >= 0) == (tr["tri_id"] >= 0)).all()


def test_threads_do_not_change_results(sponza, ray_sets, oracle_hits):
    nodes, tris = sponza
    rays = np.ascontiguousarray(ray_sets["random"][:50_000])
    one = oracle.traverse(nodes, tris, rays, threads=1)
    assert one.tobytes() == oracle_hits["random"][:50_000].tobytes()


def test_work_counters(sponza, ray_sets):
    """Node / Tri4 visit counts behind the roofline's algorithmic bytes."""
    nodes, tris = sponza
    expect = {"primary": (19.5546, 3.9438), "random": (8.7082, 2.6271)}
    for name, (n_exp, t_exp) in expect.items():
        _, st = oracle.traverse(nodes, tris, ray_sets[name], want_stats=True)
        n = len(ray_sets[name])
        assert abs(st.nodes / n - n_exp) < 1e-3 and abs(st.tri4 / n - t_exp) < 1e-3
        assert st.max_stack < 64


def test_sorting_networks():
    """Comparator sequences of src/core/sort.impala:3-66 as unrolled by the reference."""
    assert oracle.network(8, 3) == [(0, 1), (0, 2), (1, 2)]
    assert oracle.network(8, 4) == [(0, 1), (2, 3), (0, 2), (1, 3), (1, 2)]
    assert oracle.network(8, 8) == [(0, 1), (2, 3), (0, 2), (1, 3), (1, 2), (4, 5), (6, 7), (4, 6), (5, 7), (5, 6),
                                    (0, 4), (2, 6), (2, 4), (1, 5), (3, 7), (3, 5), (1, 2), (3, 4), (5, 6)]
    assert oracle.network(4, 4) == [(0, 1), (2, 3), (0, 2), (1, 3), (1, 2)]
    for n in range(3, 9):   # every network sorts (descending, as the traversal uses it)
        net = oracle.network(8, n)
        rng = np.random.default_rng(n)
        for _ in range(200):
            k = list(rng.integers(0, 5, n))
            for i, j in net:
                if k[i] < k[j]:
                    k[i], k[j] = k[j], k[i]
            assert k == sorted(k, reverse=True)


def test_bvh4_block_gives_same_records(ray_sets, oracle_hits):
    """The BVH4 block of the reference file carries the same triangles.  Records agree
    with the BVH8 traversal up to visit order: coplanar overlapping triangles (Sponza has
    many) are accepted through `t <= |det| * tmax` and the survivor depends on which was
    met first, which moves t by at most 1 ulp ("Ties")."""
    from rodent_b200 import testdata
    nodes4, tris4 = formats.load_bvh(testdata.sponza_bvh4(), formats.BVH4_TRI4)
    for name in ("primary", "random"):
        h4 = oracle.traverse(nodes4, tris4, ray_sets[name])
        h8 = oracle_hits[name]
        assert ((h4["tri_id"] >= 0) == (h8["tri_id"] >= 0)).all()
        assert (np.abs(h4["t"] - h8["t"]) <= np.spacing(np.maximum(h4["t"], h8["t"]))).all()
        assert (h4["t"] == h8["t"]).mean() > 0.998
        assert (h4["tri_id"] == h8["tri_id"]).mean() > 0.99


def test_bvh4_hit_distance_matches_golden_png(ray_sets):
    """The reference's CTest runs single_bvh4 against the same golden image (tools/CMakeLists.txt:26-31)."""
    from PIL import Image
    from rodent_b200 import testdata
    nodes4, tris4 = formats.load_bvh(testdata.sponza_bvh4(), formats.BVH4_TRI4)
    for name in ("primary", "random"):
        h4 = oracle.traverse(nodes4, tris4, ray_sets[name])
        ref = np.array(Image.open(GOLDEN / f"ref-{name}.png"))[..., 0]
        assert int((ref != formats.fbuf_to_gray(h4["t"]).reshape(1024, 1024)).sum()) <= 2


def test_vector_and_scalar_builds_agree(tmp_path, sponza, ray_sets):
    """The oracle's AVX2 node test / 4-lane triangle test (what the shipped liboracle.so runs, and what the reference's
    vectorised CPU build does) against the scalar code of the same file, compiled without AVX2: same bits, degenerate rays
    included."""
    import ctypes
    import subprocess
    src = Path(oracle.__file__).resolve().parent
    so = tmp_path / "liboracle_scalar.so"
    subprocess.run(["gcc", "-O2", "-march=x86-64", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-std=c11", "-D_GNU_SOURCE", "-shared",
                    "-o", str(so), str(src / "render_oracle.c"), str(src / "shading_bench_oracle.c"), str(src / "traversal_bvh2_oracle.c"),
                    "-lpthread", "-lm"], check=True)
    scalar = ctypes.CDLL(str(so))
    scalar.oracle_traverse.argtypes = oracle.lib().oracle_traverse.argtypes
    scalar.oracle_traverse.restype = None
    nodes, tris = sponza
    nodes4, tris4 = formats.load_bvh(testdata.sponza_bvh4(), formats.BVH4_TRI4)
    rng = np.random.default_rng(5)
    n = 40000
    d = rng.normal(size=(n, 3))
    d[np.arange(n), rng.integers(0, 3, n)] = 0.0
    d[n // 2:][np.abs(d[n // 2:]) < 0.3] = 1e-9
    d[(d == 0).all(axis=1), 0] = 1.0
    od = np.concatenate([rng.uniform([-1900, -100, -1100], [1800, 1400, 1100], (n, 3)), d], axis=1).astype(np.float32)
    sets = [formats.make_rays(od, 0.0, 1e30), np.ascontiguousarray(ray_sets["primary"][::16]), np.ascontiguousarray(ray_sets["random"][::16])]
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    for rays in sets:
        for arity, (nn, tt) in ((8, (nodes, tris)), (4, (nodes4, tris4))):
            for any_hit in (False, True):
                want = np.zeros(len(rays), formats.HIT1)
                scalar.oracle_traverse(arity, int(any_hit), ptr(nn), ptr(tt), ptr(rays), ptr(want), len(rays), 4, None)
                got = oracle.traverse(nn, tt, rays, any_hit=any_hit)
                if any_hit:
                    assert np.array_equal(got["tri_id"] >= 0, want["tri_id"] >= 0)
                else:
                    assert got.tobytes() == want.tobytes()
