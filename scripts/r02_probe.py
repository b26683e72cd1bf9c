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
    if "sbvh" in what:
        sbvh()
