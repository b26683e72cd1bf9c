"""Developer helper: one knob of the BVH2 stream kernels on the Sponza render.  usage: render_sweep.py key v1 v2 ..."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rodent_b200 import lib, render as R, workloads
scene = workloads.load_scene("sponza")
W, H, spp, depth = 1920, 1080, 16, 8
cam = workloads.camera("sponza", W, H)
key = sys.argv[1]
for value in map(int, sys.argv[2:]):
    lib.tune(key, value)
    r = R.Renderer(scene, 0, W, H, spp, depth)
    ms = [r.render(cam, it, present=False) for it in range(4)]
    mean = float(r.film().mean()) if False else None
    r.present()
    print(key, value, f"{sorted(ms[1:])[1]:.1f} ms  {W * H * spp / sorted(ms[1:])[1] / 1e3:.1f} Msamples/s  film mean {float(r.film().mean()) / 4:.9f}", flush=True)
    r.free()
