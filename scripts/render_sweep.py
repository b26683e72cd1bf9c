"""Developer helper: refill / streak thresholds of the BVH2 stream kernels on the Sponza render."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rodent_b200 import lib, render as R, workloads
scene = workloads.load_scene("sponza")
W, H, spp, depth = 1920, 1080, 16, 8
cam = workloads.camera("sponza", W, H)
for refill in (20,):
    for streak in (32, 24, 16, 8):          # here: stack levels in shared memory
        lib.tune("render_bvh2_stack", streak)
        r = R.Renderer(scene, 0, W, H, spp, depth)
        ms = [r.render(cam, it, present=False) for it in range(4)]
        r.free()
        print(refill, streak, f"{sorted(ms[1:])[1]:.1f} ms  {W * H * spp / sorted(ms[1:])[1] / 1e3:.1f} Msamples/s", flush=True)
