"""Parity of the CUDA traversal (through the C ABI) with the oracle: bit-exact hit
records on both full Sponza ray sets, the golden PNGs, edge cases, and
size-independent properties.  Runs on the B200 box (`-m gpu`)."""
from pathlib import Path

import numpy as np
import pytest
from PIL import Image

from rodent_b200 import formats

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"


@pytest.fixture(scope="module")
def gpu(sponza):
    from rodent_b200 import lib, traversal
    L = lib.load()
    assert L.rodent_b200_device_count() >= 1, "no CUDA device"
    nodes, tris = sponza
    return traversal.Bvh8(0, nodes, tris)


def run_gpu(bvh, rays, any_hit=False, prefill=None):
    from rodent_b200 import traversal
    d_rays = traversal.DeviceArray.from_host(0, rays)
    init = np.zeros(len(rays), formats.HIT1) if prefill is None else prefill
    d_hits = traversal.DeviceArray.from_host(0, init)
    traversal.intersect(bvh, d_rays, d_hits, any_hit=any_hit)
    return d_hits.to_host()


def assert_records_equal(got, want):
    """Bit-exact on every field (t/u/v compared as bit patterns)."""
    assert got.dtype == want.dtype and len(got) == len(want)
    if got.tobytes() != want.tobytes():
        bad = np.nonzero((got["tri_id"] != want["tri_id"]) | (got["t"].view("i4") != want["t"].view("i4")) |
                         (got["u"].view("i4") != want["u"].view("i4")) | (got["v"].view("i4") != want["v"].view("i4")))[0]
        raise AssertionError(f"{len(bad)} of {len(got)} records differ, first: ray {bad[0]} got {got[bad[0]]} want {want[bad[0]]}")


# kernel variants: the default vote-scheduled kernel, the while-while thread-per-ray kernels and the quad-per-ray kernel
VARIANTS = {"vote": {"mapping": 2}, "vote-refill1": {"mapping": 2, "refill_min": 1}, "vote-no-streaks": {"mapping": 2, "node_streak_min": 33},
            "vote-long-streaks": {"mapping": 2, "node_streak_min": 1}, "pool": {"mapping": 3}, "pool-refill1": {"mapping": 3, "pool_refill_min": 1},
            "quad": {"mapping": 4},
            # out-of-range knobs are clamped by rodent_b200_tune (0 would never leave the streak loop / elect lane -1)
            "vote-clamped-knobs": {"mapping": 2, "node_streak_min": 0, "refill_min": 0},
            "thread-persistent": {"mapping": 1, "persistent": 1}, "thread-grid": {"mapping": 1, "persistent": 0}}
DEFAULTS = {"mapping": 2, "persistent": 1, "refill_min": 24, "pool_refill_min": 24, "node_streak_min": 8}


@pytest.fixture(params=list(VARIANTS))
def variant(request):
    from rodent_b200 import lib
    for k, v in VARIANTS[request.param].items():
        lib.tune(k, v)
    yield request.param
    for k, v in DEFAULTS.items():
        lib.tune(k, v)


@pytest.mark.parametrize("name", ["primary", "random"])
def test_closest_hit_bit_exact(name, variant, gpu, ray_sets, oracle_hits):
    assert_records_equal(run_gpu(gpu, ray_sets[name]), oracle_hits[name])


@pytest.mark.parametrize("name", ["primary", "random"])
def test_golden_png(name, gpu, ray_sets):
    """cmake/test/run_traversal.cmake: bench_traversal -o x.fbuf ; fbuf2png -n ; compare."""
    got = run_gpu(gpu, ray_sets[name])
    ref = np.array(Image.open(GOLDEN / f"ref-{name}.png"))[..., 0]
    assert int((ref != formats.fbuf_to_gray(got["t"]).reshape(1024, 1024)).sum()) <= 2


@pytest.mark.parametrize("name", ["primary", "random"])
def test_any_hit(name, variant, gpu, sponza, ray_sets, oracle_hits):
    from oracle import oracle
    nodes, tris = sponza
    want = oracle.traverse(nodes, tris, ray_sets[name], any_hit=True)
    prefill = np.zeros(len(want), formats.HIT1)
    prefill["tri_id"] = -2
    prefill["t"] = 7.0
    got = run_gpu(gpu, ray_sets[name], any_hit=True, prefill=prefill)
    assert np.array_equal(got["tri_id"], want["tri_id"])          # same first-found triangle, same order
    assert (got["t"] == 7.0).all() and (got["u"] == 0).all()      # occluded writes tri_id only
    assert ((got["tri_id"] >= 0) == (oracle_hits[name]["tri_id"] >= 0)).all()


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 127, 129, 1000, 4097])
def test_ragged_sizes(n, gpu, ray_sets, oracle_hits):
    rays = np.ascontiguousarray(ray_sets["random"][:n])
    sentinel = np.zeros(n + 8, formats.HIT1)
    sentinel["tri_id"] = -77
    from rodent_b200 import traversal
    d_rays = traversal.DeviceArray.from_host(0, rays)
    d_hits = traversal.DeviceArray.from_host(0, sentinel)
    traversal.intersect(gpu, d_rays, d_hits, count=n)
    out = d_hits.to_host()
    assert_records_equal(out[:n], oracle_hits["random"][:n])
    assert (out[n:]["tri_id"] == -77).all(), "wrote past the end"


def test_frozen_known_answers(gpu):
    z = np.load(GOLDEN / "sponza_hits_sample.npz")
    for name in ("primary", "random"):
        assert_records_equal(run_gpu(gpu, z[f"{name}_rays"]), z[f"{name}_hits"])
        occl = run_gpu(gpu, z[f"{name}_rays"], any_hit=True, prefill=np.full(len(z[f"{name}_rays"]), -2, "i4").astype(formats.HIT1))
        assert ((occl["tri_id"] >= 0) == z[f"{name}_any"]).all()


def test_degenerate_rays(variant, gpu, sponza):
    """Axis-parallel / zero direction components (safe_rcp), tmin > 0, tmax < tmin, far origin."""
    from oracle import oracle
    nodes, tris = sponza
    rng = np.random.default_rng(7)
    n = 20_000
    od = np.empty((n, 6), np.float32)
    od[:, :3] = rng.uniform([-1900, -100, -1100], [1800, 1400, 1100], (n, 3))
    od[:, 3:] = rng.normal(size=(n, 3)) * 300
    od[::7, 3] = 0.0
    od[::11, 4] = 0.0
    od[::13, 5] = -0.0
    od[::77, 3:] = 0.0                      # null direction
    od[::5, 3:] *= 1e-12                    # |d| below the safe_rcp threshold
    od[::101, :3] += 1e6                    # origin far outside
    rays = formats.make_rays(od, 0.0, 10.0)
    rays["tmin"][::3] = 0.25
    rays["tmax"][::17] = 0.1                # tmax < tmin on some
    want = oracle.traverse(nodes, tris, rays)
    assert_records_equal(run_gpu(gpu, rays), want)
    occl = run_gpu(gpu, rays, any_hit=True)
    assert ((occl["tri_id"] >= 0) == (oracle.traverse(nodes, tris, rays, any_hit=True)["tri_id"] >= 0)).all()


def test_host_pointer_entry_points(sponza, ray_sets, oracle_hits):
    """b200_* : the drop-in for the cpu_* call sites, host buffers in and out."""
    from rodent_b200 import traversal
    nodes, tris = sponza
    for name in ("primary", "random"):
        got = traversal.intersect_host(nodes, tris, ray_sets[name])
        assert_records_equal(got, oracle_hits[name])
    pre = np.zeros(5000, formats.HIT1)
    pre["t"] = 3.0
    occl = traversal.intersect_host(nodes, tris, np.ascontiguousarray(ray_sets["random"][:5000]), hits=pre, any_hit=True)
    assert ((occl["tri_id"] >= 0) == (oracle_hits["random"][:5000]["tri_id"] >= 0)).all()
    assert (occl["t"] == 3.0).all()


def test_host_bvh_cache_follows_the_content(sponza, ray_sets, oracle_hits):
    """The cpu_* functions read whatever the arrays hold at call time.  One allocation that first holds the Sponza BVH8 and
    is then overwritten in place -- by a BVH over other triangles, then by the original again -- must be traced as it
    is at each call (the upload is keyed on sampled content, not on the address)."""
    import ctypes
    from oracle import oracle
    from rodent_b200 import lib, render, traversal
    L = lib.load()
    nodes, tris = sponza
    rays = np.ascontiguousarray(ray_sets["random"][:20000])
    buf_n, buf_t = nodes.copy(), tris.copy()
    stats = (ctypes.c_int64 * 3)()
    L.rodent_b200_bvh_cache_stats(stats)
    uploads0, reuploads0 = stats[0], stats[1]
    assert_records_equal(traversal.intersect_host(buf_n, buf_t, rays), oracle_hits["random"][:20000])
    assert_records_equal(traversal.intersect_host(buf_n, buf_t, rays), oracle_hits["random"][:20000])
    L.rodent_b200_bvh_cache_stats(stats)
    assert stats[0] == uploads0 + 1 and stats[1] == reuploads0, "the second call must reuse the upload"
    # another BVH8 in the same allocation: the Cornell box (36 triangles, scaled into the ray set's neighbourhood)
    cornell = render.Scene.load_obj(GOLDEN / "cornell_box.obj")
    cn, ct = cornell.array("nodes").copy(), cornell.array("tris").copy()
    buf_n[:len(cn)] = cn
    buf_t[:len(ct)] = ct
    crays = formats.make_rays(np.random.default_rng(5).uniform(-1, 1, (5000, 6)).astype(np.float32) + np.array([0, 1, 0, 0, 0, 0], np.float32), 0.0, 10.0)
    want = oracle.traverse(cn, ct, crays)
    assert (want["tri_id"] >= 0).sum() > 1000
    assert_records_equal(traversal.intersect_host(buf_n, buf_t, crays), want)
    buf_n[:], buf_t[:] = nodes, tris
    assert_records_equal(traversal.intersect_host(buf_n, buf_t, rays), oracle_hits["random"][:20000])
    L.rodent_b200_bvh_cache_stats(stats)
    assert stats[1] == reuploads0 + 2
    L.rodent_b200_forget_bvh(buf_n.ctypes.data, buf_t.ctypes.data)
    cornell.free()


def test_host_buffers_pageable_and_pinned(sponza, ray_sets, oracle_hits):
    """Pinned caller buffers are used as they are, pageable ones go through the library's staging memory (or, with the
    staging switched off, to the driver): same records every way -- with and without non-temporal staging copies, one
    buffer pinned and the other not --, closest and any hit."""
    from rodent_b200 import lib, traversal
    nodes, tris = sponza
    n = 300001
    rays = np.ascontiguousarray(ray_sets["random"][:n])
    want = oracle_hits["random"][:n]
    assert_records_equal(traversal.intersect_host(nodes, tris, rays), want)                      # pageable, staged
    lib.tune("host_stream_stores", 0)
    try:
        assert_records_equal(traversal.intersect_host(nodes, tris, rays[1:]), want[1:])          # (and an odd address)
    finally:
        lib.tune("host_stream_stores", 1)
    assert_records_equal(traversal.intersect_host(nodes, tris, rays[1:]), want[1:])
    lib.tune("host_staging", 0)
    try:
        assert_records_equal(traversal.intersect_host(nodes, tris, rays), want)                  # pageable, driver-staged
    finally:
        lib.tune("host_staging", 1)
    pin_r, pin_h = traversal.PinnedArray(formats.RAY1, n), traversal.PinnedArray(formats.HIT1, n)
    pin_r.array[:] = rays
    assert_records_equal(traversal.intersect_host(nodes, tris, pin_r.array, pin_h.array).copy(), want)     # pinned
    mixed = traversal.intersect_host(nodes, tris, pin_r.array)                                    # pinned in, pageable out
    assert_records_equal(mixed, want)
    pin_h.array[:] = 0
    assert_records_equal(traversal.intersect_host(nodes, tris, rays, pin_h.array).copy(), want)   # pageable in, pinned out
    pre = np.zeros(n, formats.HIT1)
    pre["t"] = 7.0
    occl = traversal.intersect_host(nodes, tris, rays, hits=pre, any_hit=True)
    assert ((occl["tri_id"] >= 0) == (want["tri_id"] >= 0)).all() and (occl["t"] == 7.0).all()
    pin_r.free(); pin_h.free()


def test_pinned_buffers_direct_path(sponza, ray_sets, oracle_hits):
    """Pinned caller buffers, closest hit: one launch that starts while a copy engine is still bringing the rays in
    (armed slots) and sends the records home itself (traverse_direct).  Bit-exact against the oracle at ragged sizes
    (group of 16 records: full, ragged, single), in every ray / record mode, from sub-ranges of one allocation, from
    several threads at once, interleaved with calls that overwrite the same context's ray array, with rays whose
    tmin / tmax look like an armed slot, and with the path switched off (copy-engine pieces)."""
    import threading
    from rodent_b200 import lib, traversal
    L = lib.load()
    nodes, tris = sponza
    full = len(ray_sets["random"])
    pin_r, pin_h = traversal.PinnedArray(formats.RAY1, full), traversal.PinnedArray(formats.HIT1, full)
    pin_r.array[:] = ray_sets["random"]
    for n in (1, 15, 16, 17, 33, 4097, 300001, full):
        pin_h.array[:n] = 0
        got = traversal.intersect_host(nodes, tris, pin_r.array[:n], pin_h.array[:n])
        assert L.rodent_b200_last_kernel_name(0).decode() == "traverse_direct<false, 8>"
        assert_records_equal(got.copy(), oracle_hits["random"][:n])
    try:
        for by_copy_engine in (0, 1):
            for mode in (0, 2, 1):
                lib.tune("host_direct_rays", by_copy_engine); lib.tune("host_direct_push", mode)
                pin_h.array[:] = 0
                assert_records_equal(traversal.intersect_host(nodes, tris, pin_r.array[:100003], pin_h.array[:100003]).copy(), oracle_hits["random"][:100003])
    finally:
        lib.tune("host_direct_push", 1); lib.tune("host_direct_rays", 1)
    # the same context serves a pageable call (copy-engine pieces into the same device array) in between: no stale ray
    # may be taken for an arrived one afterwards
    pageable = np.ascontiguousarray(ray_sets["primary"][:100003])
    for _ in range(2):
        assert_records_equal(traversal.intersect_host(nodes, tris, pageable), oracle_hits["primary"][:100003])
        pin_h.array[:] = 0
        assert_records_equal(traversal.intersect_host(nodes, tris, pin_r.array[:100003], pin_h.array[:100003]).copy(), oracle_hits["random"][:100003])
        assert_records_equal(traversal.intersect_host(nodes, tris, pin_r.array[5:50005], pin_h.array[5:50005]).copy(), oracle_hits["random"][5:50005])
    # arrays the host allocated itself, page-locked in place (rodent_b200_pin_host): the same path; a second registration
    # of the same range is refused, and after the release the staged path serves them again
    own_r, own_h = np.ascontiguousarray(ray_sets["primary"][:70001]), np.zeros(70001, formats.HIT1)
    assert traversal.pin_host(own_r) and traversal.pin_host(own_h) and not traversal.pin_host(own_r)
    assert_records_equal(traversal.intersect_host(nodes, tris, own_r, own_h), oracle_hits["primary"][:70001])
    assert L.rodent_b200_last_kernel_name(0).decode() == "traverse_direct<false, 8>"
    assert traversal.unpin_host(own_r) and traversal.unpin_host(own_h) and not traversal.unpin_host(own_h)
    # pageable again: the copy-engine pieces (pageable rays would have to be staged before a single launch could start) --
    # or, forced, the single launch fed through the library's staging arrays by helper threads
    for staged, kernel in ((1, "traverse_bvh8_vote"), (0, "traverse_bvh8_vote"), (2, "traverse_direct<false, 8>")):
        lib.tune("host_staged_direct", staged)
        try:
            for n in (70001, 17, 40000):
                own_h[:] = 0
                assert_records_equal(traversal.intersect_host(nodes, tris, own_r[:n], own_h[:n]), oracle_hits["primary"][:n])
                assert L.rodent_b200_last_kernel_name(0).decode().startswith(kernel)
                assert not own_h[n:].view(np.uint8).any()
        finally:
            lib.tune("host_staged_direct", 1)
    # rays whose tmin or tmax is the all-ones NaN look like slots that have not arrived: the call still ends, with the
    # records the device-pointer entry point gives for the same rays
    odd = traversal.PinnedArray(formats.RAY1, 40000)
    odd.array[:] = ray_sets["random"][:40000]
    odd.array["tmax"].view(np.uint32)[::7] = 0xFFFFFFFF
    odd.array["tmin"].view(np.uint32)[3::11] = 0xFFFFFFFF
    d_rays, d_hits = traversal.DeviceArray.from_host(0, odd.array), traversal.DeviceArray(0, formats.HIT1, 40000)
    traversal.intersect(traversal.Bvh8(0, nodes, tris), d_rays, d_hits)
    pin_h.array[:] = 0
    got = traversal.intersect_host(nodes, tris, odd.array, pin_h.array[:40000]).copy()
    assert got.tobytes() == d_hits.to_host().tobytes()
    odd.free()
    # a range in the middle of the allocation (what rodent_b200_set_devices hands every device); its neighbours stay untouched
    pin_h.array[:] = 0
    traversal.intersect_host(nodes, tris, pin_r.array[1001:70001], pin_h.array[1001:70001])
    assert_records_equal(pin_h.array[1001:70001].copy(), oracle_hits["random"][1001:70001])
    assert not pin_h.array[:1001].view(np.uint8).any() and not pin_h.array[70001:].view(np.uint8).any()
    # two sets from two threads, repeatedly
    pin_r2, pin_h2 = traversal.PinnedArray(formats.RAY1, full), traversal.PinnedArray(formats.HIT1, full)
    pin_r2.array[:] = ray_sets["primary"]

    def work(r, h):
        for _ in range(4):
            h.array[:] = 0
            traversal.intersect_host(nodes, tris, r.array, h.array)
    threads = [threading.Thread(target=work, args=a) for a in ((pin_r, pin_h), (pin_r2, pin_h2))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert_records_equal(pin_h.array.copy(), oracle_hits["random"])
    assert_records_equal(pin_h2.array.copy(), oracle_hits["primary"])
    # any hit: the same single launch; only tri_id comes home (a dense id array the helper threads scatter into the
    # records), the caller's t, u, v stay -- page-locked or pageable records, ragged sizes on one context (the tail of the
    # last group of ids must not leak into a later, longer call), ids equal to the device-pointer entry point's
    bvh_dev = traversal.Bvh8(0, nodes, tris)
    d_rays = traversal.DeviceArray.from_host(0, ray_sets["random"])
    d_hits = traversal.DeviceArray(0, formats.HIT1, full)
    traversal.intersect(bvh_dev, d_rays, d_hits, any_hit=True)
    want_ids = d_hits.to_host()["tri_id"]
    for n in (50001, 3, 64, 65, 200003, 50002):
        pin_h.array[:] = 0
        pin_h.array["t"] = 7.0
        occl = traversal.intersect_host(nodes, tris, pin_r.array[:n], pin_h.array[:n], any_hit=True)
        assert L.rodent_b200_last_kernel_name(0).decode() == "traverse_direct<true, 8>"
        assert (occl["tri_id"] == want_ids[:n]).all() and (occl["t"] == 7.0).all() and (occl["u"] == 0.0).all()
        assert (pin_h.array[n:]["tri_id"] == 0).all()
        own = np.zeros(n, formats.HIT1)
        own["v"] = 2.5
        occl = traversal.intersect_host(nodes, tris, np.ascontiguousarray(ray_sets["random"][:n]), own, any_hit=True)      # pageable both: pieces
        assert (occl["tri_id"] == want_ids[:n]).all() and (occl["v"] == 2.5).all()
        own["tri_id"] = 0
        occl = traversal.intersect_host(nodes, tris, pin_r.array[:n], own, any_hit=True)       # page-locked rays, pageable records: the ids are scattered into them
        assert L.rodent_b200_last_kernel_name(0).decode() == "traverse_direct<true, 8>"
        assert (occl["tri_id"] == want_ids[:n]).all() and (occl["v"] == 2.5).all()
    lib.tune("host_direct", 0)
    try:
        pin_h.array[:] = 0
        assert_records_equal(traversal.intersect_host(nodes, tris, pin_r.array[:4097], pin_h.array[:4097]).copy(), oracle_hits["random"][:4097])
        assert L.rodent_b200_last_kernel_name(0).decode().startswith("traverse_bvh8_vote")
    finally:
        lib.tune("host_direct", 1)
    for a in (pin_r, pin_h, pin_r2, pin_h2):
        a.free()


def test_reference_names_shim(sponza, ray_sets, oracle_hits):
    """cpu_intersect_single_ray1_bvh8_tri4 / cpu_occluded_... of librodent_b200_refnames.so: the reference's own symbol, host
    pointers, the oracle's records."""
    import ctypes
    from rodent_b200 import build
    S = ctypes.CDLL(str(build.build_shim()))
    nodes, tris = sponza
    n = 50001
    rays = np.ascontiguousarray(ray_sets["random"][:n])
    hits = np.zeros(n, formats.HIT1)
    args = [ctypes.c_void_p(a.ctypes.data) for a in (nodes, tris, rays, hits)] + [ctypes.c_int32(n)]
    S.cpu_intersect_single_ray1_bvh8_tri4.restype = None
    S.cpu_intersect_single_ray1_bvh8_tri4(*args)
    assert_records_equal(hits, oracle_hits["random"][:n])
    hits[:] = 0
    hits["t"] = 3.0
    S.cpu_occluded_single_ray1_bvh8_tri4.restype = None
    S.cpu_occluded_single_ray1_bvh8_tri4(*args)
    assert ((hits["tri_id"] >= 0) == (oracle_hits["random"][:n]["tri_id"] >= 0)).all() and (hits["t"] == 3.0).all()


def test_host_entry_points_are_reentrant(sponza, ray_sets, oracle_hits):
    """The cpu_* functions are pure; their drop-ins may be called from several host threads at once -- different sets,
    sizes and closest / any hit mixed, repeatedly (contexts are reused)."""
    import threading
    from rodent_b200 import traversal
    nodes, tris = sponza
    jobs = [("random", slice(None), False), ("primary", slice(None), False), ("random", slice(1000, 301000), False),
            ("primary", slice(7, 50007), True), ("random", slice(0, 3), False), ("primary", slice(500000, 1048576), False)]
    results = {}

    def work(k, name, sl, any_hit):
        rays = np.ascontiguousarray(ray_sets[name][sl])
        for rep in range(3):
            results[k, rep] = traversal.intersect_host(nodes, tris, rays, any_hit=any_hit)

    threads = [threading.Thread(target=work, args=(k, *job)) for k, job in enumerate(jobs)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for k, (name, sl, any_hit) in enumerate(jobs):
        for rep in range(3):
            if any_hit:
                assert ((results[k, rep]["tri_id"] >= 0) == (oracle_hits[name][sl]["tri_id"] >= 0)).all()
            else:
                assert_records_equal(results[k, rep], oracle_hits[name][sl])


def test_async_launches_overlap_on_two_streams(gpu, ray_sets, oracle_hits):
    """cuda_*_async: both sets in flight at once on two streams with their own work counters (what bench.py times)."""
    import torch
    from rodent_b200 import traversal
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    counters = torch.zeros(32, dtype=torch.int32, device="cuda")
    d_rays = {n: traversal.DeviceArray.from_host(0, ray_sets[n]) for n in ("random", "primary")}
    d_hits = {n: traversal.DeviceArray.from_host(0, np.zeros(len(ray_sets[n]), formats.HIT1)) for n in d_rays}
    for rep in range(3):
        for k, n in enumerate(d_rays):
            traversal.intersect_async(gpu, d_rays[n], d_hits[n], streams[k].cuda_stream, counters.data_ptr() + 64 * k)
        torch.cuda.synchronize()
    for n in d_rays:
        assert_records_equal(d_hits[n].to_host(), oracle_hits[n])


def test_properties_full_size(gpu, ray_sets):
    """Size-independent properties on the full 1 Mi sets: order independence (a
    permutation of the rays permutes the records), idempotence, tmax clipping
    (re-tracing with tmax just beyond the found t finds the same surface)."""
    rays = ray_sets["random"]
    base = run_gpu(gpu, rays)
    perm = np.random.default_rng(3).permutation(len(rays))
    shuffled = run_gpu(gpu, np.ascontiguousarray(rays[perm]))
    assert_records_equal(shuffled, base[perm])
    assert_records_equal(run_gpu(gpu, rays), base)
    clipped = rays.copy()
    # (slack: the slab test is not conservative for a tmax within rounding error of the hit)
    clipped["tmax"] = np.minimum(base["t"] * np.float32(1.01) + np.float32(1e-3), rays["tmax"])
    again = run_gpu(gpu, clipped)
    hit = base["tri_id"] >= 0
    assert (again["tri_id"][hit] >= 0).all()
    # same surface; t may move by 1 ulp where coplanar triangles overlap (accept order changes with tmax)
    assert (np.abs(again["t"][hit] - base["t"][hit]) <= np.spacing(base["t"][hit])).all()
    assert (again["tri_id"][~hit] == -1).all()


def test_kernel_launches_are_counted(gpu, ray_sets):
    from rodent_b200 import lib
    L = lib.load()
    before = L.rodent_b200_launch_count()
    run_gpu(gpu, np.ascontiguousarray(ray_sets["primary"][:1024]))
    assert L.rodent_b200_launch_count() == before + 1
    assert L.rodent_b200_last_kernel_ms(0) > 0


# ---- BVH4 input (the reference's default --bvh-width) ---------------------------------------------------
@pytest.fixture(scope="module")
def sponza4():
    from rodent_b200 import testdata
    return formats.load_bvh(testdata.sponza_bvh4(), formats.BVH4_TRI4)


@pytest.fixture(scope="module")
def gpu4(sponza4):
    from rodent_b200 import traversal
    return traversal.Bvh8(0, *sponza4)


@pytest.mark.parametrize("name", ["primary", "random"])
def test_bvh4_closest_and_any_hit_bit_exact(name, gpu4, sponza4, ray_sets):
    from oracle import oracle
    nodes4, tris4 = sponza4
    want = oracle.traverse(nodes4, tris4, ray_sets[name])
    assert_records_equal(run_gpu(gpu4, ray_sets[name]), want)
    occl = run_gpu(gpu4, ray_sets[name], any_hit=True)
    assert np.array_equal(occl["tri_id"], oracle.traverse(nodes4, tris4, ray_sets[name], any_hit=True)["tri_id"])
    ref = np.array(Image.open(GOLDEN / f"ref-{name}.png"))[..., 0]
    assert int((ref != formats.fbuf_to_gray(want["t"]).reshape(1024, 1024)).sum()) <= 2


def test_bvh4_degenerate_rays_and_host_entry_point(gpu4, sponza4, ray_sets):
    from oracle import oracle
    from rodent_b200 import traversal
    nodes4, tris4 = sponza4
    rng = np.random.default_rng(11)
    n = 10_000
    od = np.empty((n, 6), np.float32)
    od[:, :3] = rng.uniform([-1900, -100, -1100], [1800, 1400, 1100], (n, 3))
    od[:, 3:] = rng.normal(size=(n, 3)) * 300
    od[::7, 3] = 0.0; od[::11, 4] = 0.0; od[::13, 5] = -0.0; od[::77, 3:] = 0.0; od[::5, 3:] *= 1e-12
    rays = formats.make_rays(od, 0.0, 10.0)
    rays["tmin"][::3] = 0.25
    rays["tmax"][::17] = 0.1
    assert_records_equal(run_gpu(gpu4, rays), oracle.traverse(nodes4, tris4, rays))
    some = np.ascontiguousarray(ray_sets["random"][:20_001])
    want = oracle.traverse(nodes4, tris4, some)
    assert_records_equal(traversal.intersect_host(nodes4, tris4, some), want)
    # page-locked arrays: the direct path over the BVH4 (the reference's default --bvh-width), the degenerate rays too
    from rodent_b200 import lib
    for r, w in ((some, want), (rays, oracle.traverse(nodes4, tris4, rays))):
        pr, ph = traversal.PinnedArray(formats.RAY1, len(r)), traversal.PinnedArray(formats.HIT1, len(r))
        pr.array[:] = r
        assert_records_equal(traversal.intersect_host(nodes4, tris4, pr.array, ph.array).copy(), w)
        assert lib.load().rodent_b200_last_kernel_name(0).decode() == "traverse_direct<false, 4>"
        pr.free(); ph.free()


# ---- packet / hybrid entry points ------------------------------------------------------------------------
@pytest.mark.parametrize("kind,width,arity", [("hybrid", 8, 4), ("packet", 4, 4), ("hybrid", 4, 8), ("packet", 8, 8)])
def test_packet_entry_points(kind, width, arity, sponza, sponza4, ray_sets):
    """Every ray of a packet gets the record the single-ray kernel (and oracle) gives it; occluded writes tri_id only."""
    from oracle import oracle
    from rodent_b200 import traversal
    nodes, tris = sponza if arity == 8 else sponza4
    rays = np.ascontiguousarray(ray_sets["random"][:100_003])
    packets = formats.pack_rays(rays, width)
    n = len(packets) * width
    assert n == 100_003 // width * width
    want = oracle.traverse(nodes, tris, np.ascontiguousarray(rays[:n]))
    got = formats.unpack_hits(traversal.intersect_host_packets(nodes, tris, packets, kind))
    assert_records_equal(got, want)
    pre = np.zeros(len(packets), formats.packet_dtypes(width)[1])
    pre["t"] = 3.0
    occl = traversal.intersect_host_packets(nodes, tris, packets, kind, any_hit=True, hits=pre)
    assert np.array_equal(formats.unpack_hits(occl)["tri_id"], oracle.traverse(nodes, tris, np.ascontiguousarray(rays[:n]), any_hit=True)["tri_id"])
    assert (occl["t"] == 3.0).all()
    # packets transposed by the helper threads on the way in and out (copy-engine pieces for closest hit, the single
    # launch with its dense id array for any hit), and the plain copy / launch / copy of the packet kernel; one packet, a
    # few, many
    from rodent_b200 import lib
    want_any = oracle.traverse(nodes, tris, np.ascontiguousarray(rays[:n]), any_hit=True)["tri_id"]
    for staged in (1, 0):
        lib.tune("host_staged_direct", staged)
        try:
            for count in (1, 5, 4099, len(packets)):
                got = formats.unpack_hits(traversal.intersect_host_packets(nodes, tris, packets[:count], kind))
                assert_records_equal(got, want[:count * width])
                pre = np.zeros(count, formats.packet_dtypes(width)[1])
                pre["u"] = 0.25
                occl = traversal.intersect_host_packets(nodes, tris, packets[:count], kind, any_hit=True, hits=pre)
                assert np.array_equal(formats.unpack_hits(occl)["tri_id"], want_any[:count * width]) and (occl["u"] == 0.25).all()
        finally:
            lib.tune("host_staged_direct", 1)
