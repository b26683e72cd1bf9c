"""Developer helper: a launch made of the longest rays only (the tail of a real launch in isolation)."""
import subprocess, sys, tempfile
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from rodent_b200 import formats, lib, testdata, traversal
lib.load()
root = Path(__file__).resolve().parent.parent
exe, out = Path(tempfile.gettempdir()) / "ray_steps", Path(tempfile.gettempdir()) / "steps_random.u32"
subprocess.run(["gcc", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-o", str(exe), str(root / "scripts" / "ray_steps.c"), "-lpthread", "-lm"], check=True)
subprocess.run([str(exe), str(testdata.sponza_bvh8()), str(testdata.rays("random")), "0", "1", str(out)], check=True)
steps = np.fromfile(out, np.uint32)
nodes, tris = formats.load_bvh(testdata.sponza_bvh8())
bvh = traversal.Bvh8(0, nodes, tris)
rays = formats.load_rays(testdata.rays("random"), 0.0, 1.0)
for k, v in (a.split("=") for a in sys.argv[1:]):
    lib.tune(k, int(v))
for thr in (128, 300):
    sel = np.nonzero(steps > thr)[0]
    r = np.ascontiguousarray(rays[sel])
    # one ray per warp: pad with rays that finish at once (tmax < tmin) so that every monster sits alone in its warp
    padded = np.repeat(r, 32)
    padded["tmax"][np.arange(len(padded)) % 32 != 0] = -1.0
    for label, rr in (("packed", r), ("one per warp", padded)):
        d_rays = traversal.DeviceArray.from_host(0, rr); d_hits = traversal.DeviceArray(0, formats.HIT1, len(rr))
        for _ in range(3): traversal.intersect(bvh, d_rays, d_hits)
        ms = float(np.median([traversal.intersect(bvh, d_rays, d_hits) for _ in range(10)]))
        print(f"rays > {thr} steps: {len(sel):4d} rays, max {int(steps[sel].max())} steps, {label:13s}: {ms*1e3:7.1f} us -> {ms*1e3/steps[sel].max():.3f} us per step of the longest", flush=True)
