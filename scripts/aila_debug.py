"""Developer tool: how the Aila-Laine comparator's distances differ from the BVH2 kernel's."""
import subprocess, sys, tempfile
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from rodent_b200 import formats, testdata, traversal
ROOT = Path(__file__).resolve().parent.parent
exe = ROOT / "baseline" / "_ref" / "aila" / "bench_aila"
nodes, tris = formats.load_bvh(testdata.sponza_bvh2(), formats.BVH2_TRI1)
bvh = traversal.Bvh8(0, nodes, tris)
for name, (tmin, tmax) in testdata.RAY_SETS.items():
    out = Path(tempfile.mkdtemp()) / "a.fbuf"
    subprocess.run([str(exe), "-bvh", str(testdata.sponza_bvh2()), "-ray", str(testdata.rays(name)), "--tmin", str(tmin), "--tmax", str(tmax), "-o", str(out)], check=True, capture_output=True, timeout=60)
    ta = np.fromfile(out, "<f4")
    rays = formats.load_rays(testdata.rays(name), tmin, tmax)
    d_rays = traversal.DeviceArray.from_host(0, rays); d_hits = traversal.DeviceArray.from_host(0, np.zeros(len(rays), formats.HIT1))
    traversal.intersect(bvh, d_rays, d_hits)
    tm = d_hits.to_host()["t"]
    print(name, "equal", (ta == tm).mean(), "close", (np.abs(ta - tm) <= 1e-4 * np.abs(tm)).mean(), "ta>tm", (ta > tm * 1.0001).mean(), "ta<tm", (ta < tm * 0.9999).mean())
    print("  sorted multisets close:", (np.abs(np.sort(ta) - np.sort(tm)) <= 1e-4 * np.abs(np.sort(tm))).mean())
    bad = np.nonzero(np.abs(ta - tm) > 1e-4 * np.abs(tm))[0]
    print("  first bad rays", bad[:10], "ta", ta[bad[:10]], "tm", tm[bad[:10]])
    print("  bad per 128-ray block histogram:", np.bincount((bad // 128) % 8, minlength=8), " bad lanes", np.bincount(bad % 32, minlength=32))
