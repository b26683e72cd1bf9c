"""Developer helper: K renderers over interleaved row bands of one film on ONE GPU, each driven by its own host
thread, so that the kernels of one pipeline fill the tails and sync gaps of the others.  usage: render_dual_probe.py scene K [spp]"""
import sys, threading, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from rodent_b200 import render as R, workloads, traversal, formats

name = sys.argv[1]; K = int(sys.argv[2])
cfg = workloads.RENDER_CONFIGS[name]
W, H, depth = cfg["width"], cfg["height"], cfg["max_path_len"]
spp = int(sys.argv[3]) if len(sys.argv) > 3 else cfg["spp"]
scene = workloads.load_scene(name); cam = workloads.camera(name, W, H)
film = traversal.DeviceArray(0, np.float32, W * H * 3)
rs = [R.Renderer(scene, 0, W, H, spp, depth, part=k, num_parts=K, band=8) for k in range(K)]
for r in rs: r.bind_film(film.ptr)
def run(it):
    ths = [threading.Thread(target=lambda r=r: r.render(cam, it, present=False)) for r in rs]
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    return (time.perf_counter() - t0) * 1e3
run(0)
for it in range(1, 4):
    ms = run(it)
    print(f"{name} K={K} spp {spp}: {ms:.1f} ms wall, {W*H*spp/ms/1e3:.1f} Msamples/s", flush=True)
