"""Developer helper: times the wavefront path tracer on the BASELINE.json render configs.
usage: render_probe.py [cornell|sponza] [width height spp depth iters] [lanes] [render_bvh2]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from rodent_b200 import lib, render as R, workloads

name = sys.argv[1] if len(sys.argv) > 1 else "cornell"
cfg = workloads.RENDER_CONFIGS[name]
W, H, spp, depth = (int(x) for x in sys.argv[2:6]) if len(sys.argv) >= 6 else (cfg["width"], cfg["height"], cfg["spp"], cfg["max_path_len"])
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 3
if len(sys.argv) > 7:
    lib.tune("render_lanes", int(sys.argv[7]))
if len(sys.argv) > 8:
    lib.tune("render_bvh2", int(sys.argv[8]))
if len(sys.argv) > 9:
    lib.tune("render_shadow_bvh2", int(sys.argv[9]))
if len(sys.argv) > 11:
    lib.tune("render_wide", int(sys.argv[11]))
scene = workloads.load_scene(name)
if len(sys.argv) > 10 and int(sys.argv[10]):
    scene.build_bvh2()
cam = workloads.camera(name, W, H)
r = R.Renderer(scene, 0, W, H, spp, depth)
for it in range(iters):
    t0 = time.perf_counter()
    ms = r.render(cam, it, present=False)
    wall = (time.perf_counter() - t0) * 1e3
    st = r.stats()
    print(f"{name} {W}x{H} spp {spp} depth {depth} iter {it}: {ms:.1f} ms device ({wall:.1f} wall), {W*H*spp/ms/1e3:.1f} Msamples/s, "
          f"{(st['primary_rays']+st['shadow_rays'])/ms/1e3:.1f} Mrays/s, {st}", flush=True)
r.present()
print("film mean", float(r.film().mean()) / iters)
