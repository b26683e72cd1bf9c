"""Round-2 developer probe (one GPU): knob sweeps of the traversal kernel and of the render loop.
usage: python scripts/r02_probe.py [trav] [render]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from rodent_b200 import formats, lib, testdata, traversal


def trav():
    nodes, tris = formats.load_bvh(testdata.sponza_bvh8())
    bvh = traversal.Bvh8(0, nodes, tris)
    sets = {}
    for name, (tmin, tmax) in testdata.RAY_SETS.items():
        rays = formats.load_rays(testdata.rays(name), tmin, tmax)
        sets[name] = (traversal.DeviceArray.from_host(0, rays), traversal.DeviceArray(0, formats.HIT1, len(rays)))

    def run(name):
        d_rays, d_hits = sets[name]
        for _ in range(3):
            traversal.intersect(bvh, d_rays, d_hits)
        ts = [traversal.intersect(bvh, d_rays, d_hits) for _ in range(12)]
        return d_rays.count / float(np.median(ts)) / 1e3

    print("refill_min x node_streak_min -> primary / random Mrays/s (one launch at a time)")
    for refill in (12, 16, 20, 24, 28):
        row = []
        for streak in (4, 8, 12, 16, 33):
            lib.tune("refill_min", refill); lib.tune("node_streak_min", streak)
            row.append(f"{run('primary'):6.0f}/{run('random'):5.0f}")
        print(f"refill {refill:2d}: " + "  ".join(f"s{s}: {r}" for s, r in zip((4, 8, 12, 16, 33), row)), flush=True)
    lib.tune("refill_min", 24); lib.tune("node_streak_min", 8)


def render():
    from rodent_b200 import render as R, workloads
    for name, spp in (("cornell", 64), ("sponza", 32)):
        cfg = workloads.RENDER_CONFIGS[name]
        W, H, depth = cfg["width"], cfg["height"], cfg["max_path_len"]
        scene = workloads.load_scene(name)
        cam = workloads.camera(name)
        for fma in (1, 0):
            for lanes in (1, 2, 3, 4):
                lib.tune("render_fma", fma); lib.tune("render_lanes", lanes)
                r = R.Renderer(scene, 0, W, H, spp, depth)
                r.render(cam, 0, present=False)
                t0 = time.perf_counter()
                ms = [r.render(cam, it, present=False) for it in range(1, 3)]
                wall = (time.perf_counter() - t0) / 2 * 1e3
                st = r.stats()
                r.free()
                print(f"{name} {W}x{H} {spp} spp fma {fma} lanes {lanes}: {W * H * spp / np.mean(ms) / 1e3:8.1f} Msamples/s device "
                      f"({np.mean(ms):.2f} ms, wall {wall:.2f} ms, {st['wavefronts']} wavefronts, {st['kernels']} kernels)", flush=True)
    lib.tune("render_fma", 1); lib.tune("render_lanes", 3)


def e2e():
    """Host-pointer entry points, both sets from two host threads (what bench.py's e2e does): piece count and schedule."""
    from concurrent.futures import ThreadPoolExecutor
    nodes, tris = formats.load_bvh(testdata.sponza_bvh8())
    sets = {}
    for name, (tmin, tmax) in testdata.RAY_SETS.items():
        rays = formats.load_rays(testdata.rays(name), tmin, tmax)
        pr = traversal.PinnedArray(formats.RAY1, len(rays)); pr.array[:] = rays
        sets[name] = (pr, traversal.PinnedArray(formats.HIT1, len(rays)))
    pool = ThreadPoolExecutor(2)

    def step():
        jobs = [pool.submit(traversal.intersect_host, nodes, tris, sets[n][0].array, sets[n][1].array) for n in ("random", "primary")]
        for j in jobs:
            j.result()

    for ramp in (0, 1):
        for chunks in (2, 3, 4, 5, 6, 8):
            lib.tune("host_ramp", ramp); lib.tune("host_chunks", chunks)
            for _ in range(3):
                step()
            ts = []
            for _ in range(15):
                t0 = time.perf_counter(); step(); ts.append((time.perf_counter() - t0) * 1e3)
            print(f"ramp {ramp} chunks {chunks}: median {np.median(ts):.3f} ms, min {min(ts):.3f} ms -> {2 * (1 << 20) / np.median(ts) / 1e3:.0f} Mrays/s", flush=True)
    lib.tune("host_ramp", 0); lib.tune("host_chunks", 5)


def e2e_direct():
    """One launch per host-pointer call, rays and records over PCIe by the kernel itself (run_host_direct), against the copy-engine pieces."""
    from concurrent.futures import ThreadPoolExecutor
    nodes, tris = formats.load_bvh(testdata.sponza_bvh8())
    bvh = traversal.Bvh8(0, nodes, tris)
    sets, want = {}, {}
    for name, (tmin, tmax) in testdata.RAY_SETS.items():
        rays = formats.load_rays(testdata.rays(name), tmin, tmax)
        pr = traversal.PinnedArray(formats.RAY1, len(rays)); pr.array[:] = rays
        sets[name] = (pr, traversal.PinnedArray(formats.HIT1, len(rays)))
        d_rays, d_hits = traversal.DeviceArray.from_host(0, rays), traversal.DeviceArray(0, formats.HIT1, len(rays))
        traversal.intersect(bvh, d_rays, d_hits)
        want[name] = d_hits.to_host().tobytes()
    pool = ThreadPoolExecutor(2)

    def step(names):
        jobs = [pool.submit(traversal.intersect_host, nodes, tris, sets[n][0].array, sets[n][1].array) for n in names]
        for j in jobs:
            j.result()

    def measure(label):
        for names in (("random", "primary"), ("primary", "random"), ("primary",), ("random",)):
            for n in names:
                sets[n][1].array[:] = 0
            for _ in range(3):
                step(names)
            ok = all(sets[n][1].array.tobytes() == want[n] for n in names)
            ts = []
            for _ in range(20):
                t0 = time.perf_counter(); step(names); ts.append((time.perf_counter() - t0) * 1e3)
            ok &= all(sets[n][1].array.tobytes() == want[n] for n in names)
            print(f"{label:28s} {'+'.join(names):15s}: median {np.median(ts):.3f} ms, min {min(ts):.3f} ms -> "
                  f"{len(names) * (1 << 20) / np.median(ts) / 1e3:.0f} Mrays/s, records equal to the device path: {ok}, "
                  f"kernel {lib.load().rodent_b200_last_kernel_name(0).decode()}", flush=True)

    lib.tune("host_direct", 0); measure("copy-engine pieces")
    lib.tune("host_direct", 1)
    for rays_by_ce, push in ((1, 1), (0, 1), (1, 2), (1, 1)):
        lib.tune("host_direct_rays", rays_by_ce); lib.tune("host_direct_push", push); measure(f"direct, rays by copy engine {rays_by_ce}, push {push}")
        lib.tune("host_trace", 1); step(("primary",)); step(("random",)); step(("random", "primary")); lib.tune("host_trace", 0)
    lib.tune("host_direct_push", 1); lib.tune("host_direct_rays", 1)
    # any hit: the pieces round-trip the caller's records, the direct kernel sends a dense id array home
    for direct in (0, 1):
        lib.tune("host_direct", direct)
        for n in ("primary", "random"):
            for _ in range(3):
                traversal.intersect_host(nodes, tris, sets[n][0].array, sets[n][1].array, any_hit=True)
            ts = []
            for _ in range(12):
                t0 = time.perf_counter(); traversal.intersect_host(nodes, tris, sets[n][0].array, sets[n][1].array, any_hit=True); ts.append((time.perf_counter() - t0) * 1e3)
            print(f"any hit, direct {direct}, {n}: median {np.median(ts):.3f} ms", flush=True)
    lib.tune("host_direct", 1)
    # ragged sizes and any-hit
    rays = sets["random"][0].array
    for n in (1 << 12, (1 << 12) + 1, 100_003, 777_777):
        for any_hit in (False, True):
            outs = []
            for direct in (1, 0):
                lib.tune("host_direct", direct)
                h = traversal.PinnedArray(formats.HIT1, n); h.array[:] = 0
                traversal.intersect_host(nodes, tris, rays[:n], h.array, any_hit=any_hit)
                outs.append(h.array.tobytes())
            print(f"n {n} any {any_hit}: direct == pieces: {outs[0] == outs[1]}", flush=True)
    lib.tune("host_direct", 1)


def zero_copy():
    """The device-pointer entry point handed PINNED HOST pointers (UVA): the kernel reads its rays and writes its records
    over PCIe itself, no copy engine.  ms per launch (CUDA events), records against the device-resident run."""
    import types
    nodes, tris = formats.load_bvh(testdata.sponza_bvh8())
    bvh = traversal.Bvh8(0, nodes, tris)
    for name, (tmin, tmax) in testdata.RAY_SETS.items():
        rays = formats.load_rays(testdata.rays(name), tmin, tmax)
        n = len(rays)
        pr = traversal.PinnedArray(formats.RAY1, n); pr.array[:] = rays
        ph = traversal.PinnedArray(formats.HIT1, n)
        d_rays, d_hits = traversal.DeviceArray.from_host(0, rays), traversal.DeviceArray(0, formats.HIT1, n)
        traversal.intersect(bvh, d_rays, d_hits)
        want = d_hits.to_host().tobytes()
        h_rays = types.SimpleNamespace(ptr=pr.ptr, count=n); h_hits = types.SimpleNamespace(ptr=ph.ptr, count=n)
        for label, r, h in (("device rays, device records", d_rays, d_hits), ("host rays, device records", h_rays, d_hits),
                            ("device rays, host records", d_rays, h_hits), ("host rays, host records", h_rays, h_hits)):
            ph.array[:] = 0
            for _ in range(3):
                traversal.intersect(bvh, r, h)
            ts = [traversal.intersect(bvh, r, h) for _ in range(12)]
            got = ph.array.tobytes() if h is h_hits else d_hits.to_host().tobytes()
            print(f"{name:8s} {label:30s}: {np.median(ts):.3f} ms -> {n / np.median(ts) / 1e3:.0f} Mrays/s, equal {got == want}", flush=True)


def pageable():
    """Host-pointer calls with pageable (numpy) buffers: threads per staging memcpy, pieces per call; one call at a time and two at once."""
    from concurrent.futures import ThreadPoolExecutor
    import os
    nodes, tris = formats.load_bvh(testdata.sponza_bvh8())
    sets = {}
    for name, (tmin, tmax) in testdata.RAY_SETS.items():
        rays = formats.load_rays(testdata.rays(name), tmin, tmax)
        sets[name] = (rays, np.zeros(len(rays), formats.HIT1))
    pool = ThreadPoolExecutor(2)
    print("host cores:", os.cpu_count(), flush=True)

    def step(names):
        jobs = [pool.submit(traversal.intersect_host, nodes, tris, sets[n][0], sets[n][1]) for n in names]
        for j in jobs:
            j.result()

    want = {}
    for n in sets:
        lib.tune("host_staged_direct", 0)
        want[n] = traversal.intersect_host(nodes, tris, sets[n][0]).tobytes()
    for staged, parts, nt in ((0, 4, 0), (0, 4, 1), (0, 8, 1), (1, 4, 0), (1, 4, 1), (1, 8, 1)):
        if True:
            lib.tune("host_staged_direct", staged); lib.tune("host_copy_parts", parts); lib.tune("host_stream_stores", nt)
            row = []
            for names in (("primary",), ("random",), ("random", "primary")):
                for n in names:
                    sets[n][1][:] = 0
                for _ in range(3):
                    step(names)
                ok = all(sets[n][1].tobytes() == want[n] for n in names)
                ts = []
                for _ in range(12):
                    t0 = time.perf_counter(); step(names); ts.append((time.perf_counter() - t0) * 1e3)
                ok &= all(sets[n][1].tobytes() == want[n] for n in names)
                row.append(f"{'+'.join(names)} {np.median(ts):.3f} ms ({'ok' if ok else 'WRONG'})")
            print(f"staged direct {staged}, {parts} threads per copy, non-temporal stores {nt}: " + ", ".join(row) + f"  [{lib.load().rodent_b200_last_kernel_name(0).decode()}]", flush=True)
    lib.tune("host_staged_direct", 1); lib.tune("host_stream_stores", 1)
    lib.tune("host_copy_parts", 8); lib.tune("host_chunks", 5)


def sbvh():
    """Sponza render through the reference file's BVH2 block against a BVH2 from this repository's split-BVH builder."""
    from rodent_b200 import render as R, workloads
    cfg = workloads.RENDER_CONFIGS["sponza"]
    W, H, depth, spp = cfg["width"], cfg["height"], cfg["max_path_len"], 32
    cam = workloads.camera("sponza")
    nodes, tris = formats.load_bvh(testdata.sponza_bvh8(), formats.BVH8_TRI4)
    for which in ("reference BVH2 block", "own split BVH2"):
        scene = R.Scene.from_bvh8(nodes, tris, workloads.sponza_materials(), workloads.sponza_material_of_prim(tris))
        t0 = time.perf_counter()
        if which.startswith("ref"):
            scene.set_bvh2(*formats.load_bvh(testdata.sponza_bvh2(), formats.BVH2_TRI1))
        else:
            scene.build_bvh2()
        build_s = time.perf_counter() - t0
        r = R.Renderer(scene, 0, W, H, spp, depth)
        r.render(cam, 0, present=False)
        ms = [r.render(cam, it, present=False) for it in range(1, 3)]
        film_mean = float(np.mean(r.film())) if False else 0.0
        st = r.stats()
        r.free()
        print(f"sponza {W}x{H} {spp} spp, {which} ({scene.view.num_nodes2} Node2, {scene.view.num_tri1} Tri1, {build_s:.1f} s): "
              f"{W * H * spp / np.mean(ms) / 1e3:8.1f} Msamples/s ({np.mean(ms):.2f} ms, {st['primary_rays']} + {st['shadow_rays']} rays)", flush=True)
        scene.free()


if __name__ == "__main__":
    what = sys.argv[1:] or ["trav", "render"]
    if "trav" in what:
        trav()
    if "render" in what:
        render()
    if "e2e" in what:
        e2e()
    if "e2e_direct" in what:
        e2e_direct()
    if "zero_copy" in what:
        zero_copy()
    if "pageable" in what:
        pageable()
    if "sbvh" in what:
        sbvh()
