"""Developer helper: how much of a launch is tail?  Times the random set in its own order, longest rays first,
and without the rays that take more than 64 steps (step counts from scripts/ray_steps.c)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from rodent_b200 import formats, lib, testdata, traversal
lib.load()
nodes, tris = formats.load_bvh(testdata.sponza_bvh8())
bvh = traversal.Bvh8(0, nodes, tris)
rays = formats.load_rays(testdata.rays("random"), 0.0, 1.0)
import subprocess, tempfile
root = Path(__file__).resolve().parent.parent
exe, out = Path(tempfile.gettempdir()) / "ray_steps", Path(tempfile.gettempdir()) / "steps_random.u32"
subprocess.run(["gcc", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-o", str(exe), str(root / "scripts" / "ray_steps.c"), "-lpthread", "-lm"], check=True)
subprocess.run([str(exe), str(testdata.sponza_bvh8()), str(testdata.rays("random")), "0", "1", str(out)], check=True)
steps = np.fromfile(out, np.uint32)
def timed(r, label):
    d_rays = traversal.DeviceArray.from_host(0, np.ascontiguousarray(r)); d_hits = traversal.DeviceArray(0, formats.HIT1, len(r))
    for _ in range(3): traversal.intersect(bvh, d_rays, d_hits)
    ms = float(np.median([traversal.intersect(bvh, d_rays, d_hits) for _ in range(10)]))
    print(f"{label:44s} {len(r):8d} rays {ms*1e3:7.0f} us {len(r)/ms/1e3:7.0f} Mrays/s", flush=True)
timed(rays, "file order")
order = np.argsort(-steps.astype(np.int64), kind="stable")
timed(rays[order], "longest rays first")
long_first = np.concatenate([np.nonzero(steps > 64)[0], np.nonzero(steps <= 64)[0]])
timed(rays[long_first], "rays > 64 steps first, rest in file order")
timed(rays[steps <= 64], "without the rays > 64 steps")
timed(rays[steps <= 32], "without the rays > 32 steps")
print("rays > 64 steps:", int((steps > 64).sum()), " > 128:", int((steps > 128).sum()), " max", int(steps.max()))
