#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json on B200.

Workload (BASELINE.json configs[1]): `bench_traversal` on the Sponza BVH8 block with
both reference ray sets, closest hit, single-ray semantics:
    sponza-primary.rays  1 048 576 rays, tmin 0, tmax 5000   (README.md:33-34 of the reference)
    sponza-random.rays   1 048 576 rays, tmin 0, tmax 1      (README.md:35-36)
One "step" traces both sets once (2 kernel launches, 2 097 152 rays).  The two launches go to two streams through
the asynchronous entry points, the incoherent set first, so that the second launch fills the SMs the first one's
stragglers leave idle (profiles/r01_experiments.md, "The tail of a launch"); `per_set` repeats the passes one
launch at a time, as bench_traversal does.
Metric: Mrays/sec = rays / (1000 * ms), as tools/bench_traversal/bench_traversal.cpp:386-387.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line on rank 0.  `value`: rays resident in HBM, CUDA-event time of the
steps (first launch's start to the last launch's end).  `path_trace`: the render configs of BASELINE.json (configs[2..4]) through
the wavefront path tracer, image rows dealt out across the ranks, one NCCL reduce of the film
inside the timed region (samples/s = spp*w*h / t, src/driver/driver.cpp:300 of the reference).  `e2e`: the same passes through the host-pointer C ABI
(b200_intersect_single_ray1_bvh8_tri4) with pinned HOST buffers, copies in the timed
region; the two sets are submitted from two host threads (the entry points are reentrant, like the cpu_* functions
they replace).  `roofline`: algorithmic bytes of the reference layout / kernel time against
the measured HBM copy bandwidth.  `cpu_baseline` / `--impl reference`: the oracle
(restated reference algorithm, oracle/traversal_oracle.c) on the box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

METRIC = "Mrays/sec on Sponza primary+random"
UNIT = "Mrays/s"
# Inner-node / Tri4 visits per ray under the reference's traversal order, measured by the
# oracle's counters and pinned in tests/test_oracle_golden.py::test_work_counters.
VISITS = {"primary": (19.554641, 3.943832), "random": (8.708188, 2.627138)}
L2_FLUSH_BYTES = 512 << 20   # > the 126 MB L2


def bytes_per_ray(name: str, visits=None) -> float:
    """32 (Ray1) + 16 (Hit1) + 256 * inner nodes + 224 * Tri4 packets."""
    nodes, tri4 = (visits or VISITS)[name]
    return 32.0 + 16.0 + 256.0 * nodes + 224.0 * tri4


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per launch of the traversal kernel from the committed ncu capture."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get("dram_bytes_per_launch_avg")
        except Exception:
            return None
    return None


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.window = [None, None]          # timed region, perf_counter seconds

    def mark_start(self):
        self.window[0] = time.perf_counter()

    def mark_end(self):
        self.window[1] = time.perf_counter()

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.perf_counter()] + [c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.12)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        """Samples taken inside the timed region (nvidia-smi is started before the warm-up so that it is up by then);
        when the region was shorter than the sampling period, the sample nearest to it."""
        rows = [r for r in self.rows if len(r) >= 7 and r[1].isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        t0, t1 = self.window
        inside = [r for r in rows if t0 is not None and t1 is not None and t0 <= r[0] <= t1 + 0.06]
        if not inside and t0 is not None:
            inside = [min(rows, key=lambda r: abs(r[0] - t0))]
        rows = inside or rows
        sm = [int(r[1]) for r in rows]
        mx = [int(r[2]) for r in rows if r[2].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k] == "Active" for r in rows)]
        return {"sm_mhz": int(statistics.median(sm)), "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def load_workload():
    from rodent_b200 import formats, testdata
    nodes, tris = formats.load_bvh(testdata.sponza_bvh8(), formats.BVH8_TRI4)
    rays = {name: formats.load_rays(testdata.rays(name), tmin, tmax) for name, (tmin, tmax) in testdata.RAY_SETS.items()}
    return nodes, tris, rays


def cpu_pass(nodes, tris, rays, threads, min_seconds: float):
    """Times the oracle on whole passes over both ray sets until `min_seconds` elapsed."""
    from oracle import oracle
    n_rays = sum(len(r) for r in rays.values())
    oracle.traverse(nodes, tris, rays["random"][:65536].copy(), threads=threads)   # warm
    t0 = time.perf_counter()
    passes, visits = 0, {}
    while True:
        for name, r in rays.items():
            _, st = oracle.traverse(nodes, tris, r, threads=threads, want_stats=True)
            visits[name] = (st.nodes / len(r), st.tri4 / len(r))
        passes += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            break
    return passes * n_rays / dt / 1e6, passes, dt, visits


def config_dict(n_gpus: int):
    return {"workload": "bench_traversal: Sponza BVH8/Tri4 (15054 Node8 + 71115 Tri4), sponza-primary.rays "
                        "(tmax 5000) + sponza-random.rays (tmax 1), 1048576 rays each, closest hit, single-ray order",
            "rays_per_step_per_gpu": 2 << 20, "parallelism": f"rays replicated per GPU x{n_gpus}, BVH replicated, no collective",
            "l2": f"flushed between steps ({L2_FLUSH_BYTES >> 20} MiB memset outside the timed events)"}


def path_trace_bench(local: int, rank: int, world: int, barrier):
    """BASELINE.json configs[2..4] through the wavefront path tracer.  Every rank owns the row bands
    (y // 8) % world == rank, renders into a torch tensor bound as the renderer's film, and one
    ncclReduce puts the image together on rank 0 inside the timed region.  Returns a dict on every rank."""
    import torch
    import torch.distributed as dist
    from rodent_b200 import render as R, sharding, workloads
    out = {}
    names = ["cornell", "sponza"] + (["sponza4k"] if world >= 8 else [])
    for name in names:
        cfg = workloads.RENDER_CONFIGS[name]
        W, H, spp, depth = cfg["width"], cfg["height"], cfg["spp"], cfg["max_path_len"]
        scene = workloads.load_scene(name)
        cam = workloads.camera(name)
        film = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
        r = R.Renderer(scene, local, W, H, spp, depth, part=rank, num_parts=world, band=8)
        r.bind_film(film.data_ptr())
        iters = 3 if name == "cornell" else 1
        if name != "sponza4k":                      # warm-up render (the 4K config is warmed by the 1080p one)
            r.render(cam, 0, present=False)
            film.zero_()
        barrier()
        t0 = time.perf_counter()
        dev_ms = 0.0
        for it in range(iters):
            dev_ms += r.render(cam, it, present=False)
        if world > 1:
            sharding.reduce_film(film)
        barrier()
        ms = (time.perf_counter() - t0) * 1e3
        st = r.stats()
        stats = torch.tensor([ms, dev_ms, float(st["primary_rays"]), float(st["shadow_rays"]), float(st["kernels"])],
                             dtype=torch.float64, device="cuda")
        if world > 1:
            mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            ms, dev_ms = float(mx[0]), float(mx[1])
            rays = float(sm[2] + sm[3])
            kernels = float(sm[4])
        else:
            rays, kernels = float(stats[2] + stats[3]), float(stats[4])
        samples = float(W) * H * spp * iters
        mean = float(film.mean()) / iters
        out[name] = {"msamples_s": round(samples / ms / 1e3, 2), "ms_per_render": round(ms / iters, 2),
                     "device_ms_per_render_max_rank": round(dev_ms / iters, 2),
                     "mrays_s": round(rays * iters / ms / 1e3, 1),     # stats are those of the last render call
                     "width": W, "height": H, "spp": spp, "max_path_len": depth, "renders": iters,
                     "kernels_per_render_all_ranks": int(kernels), "film_mean": round(mean, 6),
                     "sharding": f"row bands of 8 over {world} rank(s), film summed with one reduce"}
        r.bind_film(None)
        r.free()
        del film
    return out


def cpu_path_trace_sample(threads: int):
    """The CPU path-tracing oracle on a bounded Cornell sample (same scene and camera as configs[2])."""
    from oracle import oracle
    from rodent_b200 import workloads
    scene = workloads.load_scene("cornell")
    W, H, spp, depth = 1024, 1024, 32, 4                  # half of configs[2]'s samples: ~1 s on 16 threads
    cam = workloads.camera("cornell", W, H)
    oracle.render(scene.view, cam, 64, 64, 1, depth, 0, threads=threads)
    t0 = time.perf_counter()
    oracle.render(scene.view, cam, W, H, spp, depth, 0, threads=threads)
    dt = time.perf_counter() - t0
    return {"msamples_s": round(W * H * spp / dt / 1e6, 3), "sample": f"cornell {W}x{H}, {spp} spp, {depth} bounces, {dt:.2f} s on {threads} threads"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nodes, tris, rays = load_workload()
    threads = os.cpu_count() or 1
    n_rays = sum(len(r) for r in rays.values())
    from oracle import oracle
    for _ in range(args.warmup):
        for r in rays.values():
            oracle.traverse(nodes, tris, r, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for r in rays.values():
            oracle.traverse(nodes, tris, r, threads=threads)
    dt = time.perf_counter() - t0
    value = args.steps * n_rays / dt / 1e6
    sample = f"{args.steps} full passes over both ray sets (2097152 rays per pass), {threads} threads, contiguous ray ranges"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.gpus),
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "oracle/traversal_oracle.c: restated reference single-ray BVH8 algorithm (the AnyDSL build "
                                 "cannot be produced here), gcc -O3 -march=x86-64-v3 -ffp-contract=off, AVX2 node test and 4-lane triangle test"},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "path_trace": {"cornell_cpu_sample": cpu_path_trace_sample(threads)},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)     # the reference protocol: --bench 50 --warmup 10
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-path-trace", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (rodent_b200 has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from rodent_b200 import formats, lib, traversal
    L = lib.load()
    L.rodent_b200_set_device(local)
    nodes, tris, rays = load_workload()
    bvh = traversal.Bvh8(local, nodes, tris)
    names = list(rays)
    d_rays = {n: traversal.DeviceArray.from_host(local, rays[n]) for n in names}
    d_hits = {n: traversal.DeviceArray(local, formats.HIT1, len(rays[n])) for n in names}
    n_rays = sum(len(rays[n]) for n in names)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    order = sorted(names, key=lambda n: n != "random")          # the incoherent set first: its tail is the long one
    streams = [torch.cuda.Stream() for _ in names]
    counters = torch.zeros(16 * len(names), dtype=torch.int32, device="cuda")
    ev0 = torch.cuda.Event(enable_timing=True)
    ev_end = [torch.cuda.Event(enable_timing=True) for _ in names]

    def step():
        """One pass over both ray sets, the launches overlapping on two streams; returns the device time in ms from the
        start of the first launch to the end of the last one (CUDA events on the launch streams)."""
        flush.zero_()
        torch.cuda.synchronize()
        ev0.record(streams[0])
        for st in streams[1:]:
            st.wait_event(ev0)
        for k, n in enumerate(order):
            traversal.intersect_async(bvh, d_rays[n], d_hits[n], streams[k].cuda_stream, counters.data_ptr() + 64 * k)
            ev_end[k].record(streams[k])
        torch.cuda.synchronize()
        return max(ev0.elapsed_time(e) for e in ev_end)

    def serial_step():
        """The same passes one launch at a time; returns per-set kernel ms (bench_gpu, bench_traversal.cpp:124-135)."""
        flush.zero_()
        torch.cuda.synchronize()
        return [traversal.intersect(bvh, d_rays[n], d_hits[n]) for n in names]

    with ClockSampler(local) as clocks:
        for _ in range(args.warmup):
            step()
        barrier()
        launches0 = L.rodent_b200_launch_count()
        step_ms = []
        t_wall0 = time.perf_counter()
        clocks.mark_start()
        for _ in range(args.steps):
            step_ms.append(step())
        barrier()
        clocks.mark_end()
        wall_ms = (time.perf_counter() - t_wall0) * 1e3
    launches = L.rodent_b200_launch_count() - launches0
    kernel_ms = float(sum(step_ms))                  # timed region on the device: K steps
    hits_found = {n: int((d_hits[n].to_host()["tri_id"] >= 0).sum()) for n in names}
    for _ in range(3):
        serial_step()
    per_set = np.array([serial_step() for _ in range(max(3, min(args.steps, 20)))])     # steps x sets, outside the timed region

    # ---- end to end: host buffers through the host-pointer C ABI -----------------------------
    pin_rays = {n: traversal.PinnedArray(formats.RAY1, len(rays[n])) for n in names}
    pin_hits = {n: traversal.PinnedArray(formats.HIT1, len(rays[n])) for n in names}
    for n in names:
        pin_rays[n].array[:] = rays[n]
    e2e_steps = max(3, min(args.steps, 20))
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(len(names))

    def e2e_step():
        """Both sets through the host-pointer entry point, one host thread per set (ctypes drops the GIL in the call)."""
        jobs = [pool.submit(traversal.intersect_host, nodes, tris, pin_rays[n].array, pin_hits[n].array) for n in order]
        for j in jobs:
            j.result()

    for _ in range(3):
        e2e_step()
    barrier()
    e2e_ms = 0.0
    for _ in range(e2e_steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_step()
        e2e_ms += (time.perf_counter() - t0) * 1e3
    pool.shutdown()
    e2e_ok = all(int((pin_hits[n].array["tri_id"] >= 0).sum()) == hits_found[n] for n in names)

    # ---- the other surfaces of the same tool, one launch at a time, outside the timed region ------------------
    # `-any` and the reference GPU path's own BVH2 / Tri1 block through cuda_*_bvh2_tri1
    def median_ms(b, n, any_hit):
        scratch = traversal.DeviceArray(local, formats.HIT1, len(rays[n]))
        ms = []
        for _ in range(7):
            flush.zero_()
            torch.cuda.synchronize()
            ms.append(traversal.intersect(b, d_rays[n], scratch, any_hit=any_hit))
        scratch.free()
        return float(np.median(ms[2:]))

    variants = {"bvh8_tri4_any_hit": {n: round(len(rays[n]) / median_ms(bvh, n, True) / 1e3, 1) for n in names}}
    from rodent_b200 import testdata
    nodes2, tris1 = formats.load_bvh(testdata.sponza_bvh2(), formats.BVH2_TRI1)
    bvh2 = traversal.Bvh8(local, nodes2, tris1)
    variants["bvh2_tri1"] = {n: round(len(rays[n]) / median_ms(bvh2, n, False) / 1e3, 1) for n in names}
    variants["bvh2_tri1_any_hit"] = {n: round(len(rays[n]) / median_ms(bvh2, n, True) / 1e3, 1) for n in names}
    variants["note"] = ("Mrays/s per set; bvh2_tri1 = the block and traversal semantics of the reference's own GPU path "
                        "(nvvm_*_single_ray1_bvh2_tri1), whose records differ from the CPU single-ray path's in ties and rounding")

    # ---- path tracing (configs[2..4]) ------------------------------------------------------------
    for pa in list(pin_rays.values()) + list(pin_hits.values()):
        pa.free()
    path_trace = None if args.no_path_trace else path_trace_bench(local, rank, world, barrier)

    # ---- max over ranks ------------------------------------------------------------------------
    if world > 1:
        t = torch.tensor([kernel_ms, e2e_ms, wall_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kernel_ms, e2e_ms, wall_ms = (float(x) for x in t.tolist())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * args.steps * n_rays / kernel_ms / 1e3
    e2e_value = world * e2e_steps * n_rays / e2e_ms / 1e3
    per_set_ms = per_set.mean(axis=0)
    peak, peak_src = hbm_peak()
    algo_bytes = {n: bytes_per_ray(n) * len(rays[n]) for n in names}
    achieved = sum(algo_bytes.values()) / (kernel_ms / args.steps * 1e-3) / 1e9            # both launches / the step they share
    achieved_serial = sum(algo_bytes.values()) / (float(per_set_ms.sum()) * 1e-3) / 1e9
    traffic = ncu_traffic()
    out = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(kernel_ms / args.steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(world),
        "per_set_note": "one launch at a time, outside the timed region (bench_traversal's procedure)",
        "per_set": {n: {"mrays_s": round(len(rays[n]) / float(per_set_ms[k]) / 1e3, 2), "ms": round(float(per_set_ms[k]), 4),
                        "min_ms": round(float(per_set[:, k].min()), 4), "hits": hits_found[n],
                        "bytes_per_ray": round(bytes_per_ray(n), 1)} for k, n in enumerate(names)},
        "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": int(sum(r.nbytes for r in rays.values())),
                "d2h_bytes_per_step": int(16 * n_rays), "steps": e2e_steps, "results_match_device_path": e2e_ok,
                "api": "b200_intersect_single_ray1_bvh8_tri4 (host pointers, pinned; BVH upload cached), "
                       "the two sets submitted from two host threads"},
        "gpu_launches": int(launches),
        "wall_ms_per_step_incl_l2_flush": round(wall_ms / args.steps, 3),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                     "kernel": "traverse_bvh8_vote<false, 5, true, 24>",
                     "launches_per_step": len(names), "achieved_one_launch_at_a_time": round(achieved_serial, 1),
                     "note": "the step's two launches of this kernel overlap on two streams: achieved = algorithmic bytes of "
                             "both / the step's device time (first start to last end); one launch at a time (per_set, what "
                             "the ncu launch list serialises) gives achieved_one_launch_at_a_time.  "
                             "algorithmic bytes of the reference layout (32+16+256*nodes+224*Tri4 per ray); the 19.8 MB BVH "
                             "is L2-resident, so frac > 1 means served from L2, not faster than HBM"},
        "clocks": clocks.summary(),
    }
    out["variants"] = variants
    if path_trace is not None:
        out["path_trace"] = path_trace
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cpu_value, passes, secs, visits = cpu_pass(nodes, tris, rays, threads, 4.0)
        out["cpu_baseline"] = {"value": round(cpu_value, 3), "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": f"{passes} full passes over both ray sets in {secs:.1f} s wall on {threads} threads",
                               "visits_per_ray": {n: [round(v, 4) for v in visits[n]] for n in names}}
        # configs[0]: the reference's CPU-runnable case, `bench_traversal -s` on the primary set, one thread
        from oracle import oracle
        t0 = time.perf_counter()
        oracle.traverse(nodes, tris, rays["primary"], threads=1)
        dt1 = time.perf_counter() - t0
        out["cpu_baseline"]["single_thread_primary"] = {"value": round(len(rays["primary"]) / dt1 / 1e6, 3), "unit": UNIT,
                                                        "sample": f"one pass over sponza-primary.rays in {dt1:.2f} s on 1 thread (configs[0])"}
        if path_trace is not None:
            out["cpu_baseline"]["path_trace"] = cpu_path_trace_sample(threads)
        for n in names:   # the algorithmic bytes must come from the oracle's counters, not from a stale constant
            assert abs(bytes_per_ray(n, visits) - bytes_per_ray(n)) < 1.0, "VISITS constants are stale"
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
