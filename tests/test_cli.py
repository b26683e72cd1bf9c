"""bench_traversal / fbuf2png command lines (the reference's CTest procedure,
cmake/test/run_traversal.cmake:1-12, README.md:33-40)."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest
from PIL import Image

from rodent_b200 import build, formats, testdata

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"


@pytest.fixture(scope="module")
def tools():
    build.build_cuda()
    build.build_tools()
    return ROOT / "tools" / "bin"


def run(*cmd):
    return subprocess.run([str(c) for c in cmd], capture_output=True, text=True)


def test_bench_traversal_argument_errors(tools):
    exe = tools / "bench_traversal"
    assert run(exe, "--help").returncode == 0
    for args, msg in ((["-ray", "x"], "No BVH file specified"), (["-bvh", "x"], "No ray file specified"),
                      (["-bvh"], "Missing argument for -bvh"), (["--frobnicate"], "Unknown option '--frobnicate'"),
                      (["stray"], "Invalid argument 'stray'"), (["-gpu", "opencl", "-bvh", "a", "-ray", "b"], "Unknown GPU platform 'opencl'"),
                      (["-bvh", "a", "-ray", "b", "-gpu", "cuda", "-s"], "Options '--gpu' and '--single' are incompatible"),
                      (["-bvh", "a", "-ray", "b", "-p", "-s"], "Options '--packet' and '--single' are incompatible"),
                      (["-bvh", "a", "-ray", "b", "--bvh-width", "3"], "Invalid BVH width"),
                      (["-bvh", "a", "-ray", "b", "-s", "--bvh-width", "2"], "Invalid BVH width"),
                      (["-bvh", "a", "-ray", "b", "--ray-width", "5"], "Invalid ray width"),
                      (["-bvh", "/nonexistent", "-ray", "b", "-s", "--bvh-width", "8"], "Cannot load BVH file")):
        r = run(exe, *args)
        assert r.returncode == 1 and msg in r.stderr, (args, r.stderr)
    # every CPU variant of the reference has a drop-in now: the default (hybrid, ray 8, BVH 4) gets as far as the file
    r = run(exe, "-bvh", "/nonexistent", "-ray", "b")
    assert r.returncode == 1 and "Cannot load BVH file" in r.stderr


def test_fbuf2png_matches_reference_quantisation(tools, tmp_path, oracle_hits):
    for name in ("primary", "random"):
        formats.save_fbuf(tmp_path / f"{name}.fbuf", oracle_hits[name])
        assert run(tools / "fbuf2png", "-n", tmp_path / f"{name}.fbuf", tmp_path / f"{name}.png").returncode == 0
        got = np.array(Image.open(tmp_path / f"{name}.png"))
        ref = np.array(Image.open(GOLDEN / f"ref-{name}.png"))
        assert got.shape == ref.shape and int((got[..., 0] != ref[..., 0]).sum()) <= 2
        assert (got[..., 3] == 255).all()
    assert run(tools / "fbuf2png", tmp_path / "primary.fbuf").returncode == 1
    (tmp_path / "short.fbuf").write_bytes(b"\0" * 100)
    assert run(tools / "fbuf2png", tmp_path / "short.fbuf", tmp_path / "x.png").returncode == 1


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [["-gpu", "cuda"], ["-s", "--bvh-width", "8"], ["-s"], ["-gpu", "cuda", "--bvh-width", "4"],
                                  ["-gpu", "cuda", "--bvh-width", "2"], ["-s", "--pinned", "--bvh-width", "8"], ["-s", "--pinned"]])
@pytest.mark.parametrize("name,tmin,tmax,hits", [("primary", "0.01", "5000", None), ("random", "0", "1", 959359)])
def test_ctest_procedure_on_gpu(tools, tmp_path, mode, name, tmin, tmax, hits):
    """single_bvh8 / single_bvh4 of tools/CMakeLists.txt:26-31 (`-s` alone is the reference's default width, 4); width 2 with
    -gpu is the block and the semantics of the reference's `-gpu nvvm`."""
    fbuf = tmp_path / "out.fbuf"
    bvh = testdata.sponza_bvh4() if mode in (["-s"], ["-s", "--pinned"], ["-gpu", "cuda", "--bvh-width", "4"]) else \
        testdata.sponza_bvh2() if mode[-1] == "2" else testdata.sponza_bvh8()      # (--pinned: the direct host-pointer path)
    r = run(tools / "bench_traversal", "-bvh", bvh, "-ray", testdata.rays(name), "--bench", "3", "--warmup", "1",
            "--tmin", tmin, "--tmax", tmax, "-o", fbuf, *mode)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    assert "1048576 ray(s) in the distribution file." in out
    assert re.search(r"ms for 3 iteration\(s\)\n[\d.e+]+ Mrays/sec\n# Average: .* ms\n# Median: .* ms\n# Min: .* ms\n(\d+) intersection\(s\)", out)
    if hits is not None:
        assert f"{hits} intersection(s)" in out
    assert run(tools / "fbuf2png", "-n", fbuf, tmp_path / "out.png").returncode == 0
    got = np.array(Image.open(tmp_path / "out.png"))[..., 0]
    ref = np.array(Image.open(GOLDEN / f"ref-{name}.png"))[..., 0]
    assert int((got != ref).sum()) <= 2      # ImageMagick `compare -metric MSE` in the reference's CTest


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [[], ["-p"], ["--ray-width", "4"], ["-p", "--ray-width", "4", "--bvh-width", "8"], ["--bvh-width", "8"],
                                  ["--packet-order"], ["--packet-order", "-p"]])
def test_packet_and_hybrid_variants_on_gpu(tools, tmp_path, mode):
    """hybrid_bvh4 (the reference's default command line), packet_bvh4, ..., of tools/CMakeLists.txt:26-31."""
    fbuf = tmp_path / "out.fbuf"
    bvh = testdata.sponza_bvh8() if "8" in mode[-1:] and "--bvh-width" in mode else testdata.sponza_bvh4()
    r = run(tools / "bench_traversal", "-bvh", bvh, "-ray", testdata.rays("primary"), "--bench", "2", "--tmin", "0.01", "--tmax", "5000", "-o", fbuf, *mode)
    assert r.returncode == 0, r.stderr
    assert "1048576 ray(s) in the distribution file." in r.stdout and "1026430 intersection(s)" in r.stdout
    assert run(tools / "fbuf2png", "-n", fbuf, tmp_path / "out.png").returncode == 0
    got = np.array(Image.open(tmp_path / "out.png"))[..., 0]
    ref = np.array(Image.open(GOLDEN / "ref-primary.png"))[..., 0]
    assert int((got != ref).sum()) <= 2


@pytest.mark.gpu
def test_any_hit_cli(tools):
    r = run(tools / "bench_traversal", "-bvh", testdata.sponza_bvh8(), "-ray", testdata.rays("random"), "--tmax", "1", "-gpu", "cuda", "-any")
    assert r.returncode == 0 and "959359 intersection(s)" in r.stdout


def test_rodent_argument_errors(tools, tmp_path):
    """tools/rodent: the option handling of src/driver/driver.cpp:164-232 (messages on stderr, exit code 1)."""
    exe = tools / "rodent"
    r = run(exe, "--help")
    assert r.returncode == 0 and "--bench  iterations" in r.stdout and "--eye    x y z" in r.stdout
    for args, msg in ((["--frobnicate"], "Unknown option '--frobnicate'"), (["stray"], "Unexpected argument 'stray'"),
                      (["--eye", "1", "2"], "Option '--eye' expects 3 arguments"), (["--width"], "Option '--width' expects 1 arguments"),
                      ([], "No scene"), (["--scene", "x.obj"], "pass --bench"),
                      (["--scene", tmp_path / "missing.obj", "--bench", "1"], "missing.obj")):
        r = run(exe, *args)
        assert r.returncode == 1 and msg in r.stderr, (args, r.stderr)


@pytest.mark.gpu
def test_rodent_ctest_procedure_on_gpu(tools, tmp_path):
    """cmake/test/run_rodent.cmake:1-8: rodent --bench 50 -o out.png --eye 0 1 2.7 --dir 0 0 -1 --up 0 1 0, compared
    with testing/ref-cornell.png by MSE (defaults 1080x720, fov 60, SPP 4, MAX_PATH_LEN 64)."""
    out = tmp_path / "cornell.png"
    r = run(tools / "rodent", "--scene", GOLDEN / "cornell_box.obj", "--bench", "50", "-o", out,
            "--eye", "0", "1", "2.7", "--dir", "0", "0", "-1", "--up", "0", "1", "0")
    assert r.returncode == 0, r.stderr
    m = re.search(r"# ([\d.e+-]+)/([\d.e+-]+)/([\d.e+-]+) \(min/med/max Msamples/s\)", r.stdout)
    assert m and 0 < float(m.group(1)) <= float(m.group(2)) <= float(m.group(3))
    img = np.array(Image.open(out))[..., :3].astype(np.int32)
    ref = np.array(Image.open(GOLDEN / "ref-cornell.png"))[..., :3].astype(np.int32)
    assert img.shape == ref.shape
    assert ((img - ref) ** 2).mean() < 0.5


def test_bvh_extractor_writes_both_blocks(tools, tmp_path):
    """tools/bvh_extractor (the reference's extract_bvh4_8): an OBJ scene becomes a .bvh with a BVH8 and a BVH4 block that
    the readers accept and that give the same closest hits as a brute-force search over the triangles."""
    from oracle import oracle
    out = tmp_path / "cornell.bvh"
    r = run(tools / "bvh_extractor", GOLDEN / "cornell_box.obj", out)
    assert r.returncode == 0 and "36 triangles" in r.stdout, r.stderr
    n8, t8 = formats.load_bvh(out, formats.BVH8_TRI4)
    n4, t4 = formats.load_bvh(out, formats.BVH4_TRI4)
    assert len(n8) >= 1 and len(n4) >= len(n8) and len(t8) == len(t4)
    prims = lambda t: sorted((t["prim_id"][t["prim_id"] != -1] & 0x7FFFFFFF).tolist())
    assert prims(t8) == prims(t4) == list(range(36))
    rng = np.random.default_rng(0)
    od = np.concatenate([rng.uniform(-1, 2, (5000, 3)), rng.normal(size=(5000, 3))], 1).astype(np.float32)
    rays = formats.make_rays(od, 0.0, 100.0)
    h8, h4, hb = oracle.traverse(n8, t8, rays), oracle.traverse(n4, t4, rays), oracle.brute_force(t8, rays)
    assert (h8["tri_id"] >= 0).sum() > 1000
    assert np.array_equal(h8["t"], hb["t"]) and np.array_equal(h4["t"], hb["t"])
    assert run(tools / "bvh_extractor", tmp_path / "missing.obj", out).returncode == 1
    assert run(tools / "bvh_extractor").returncode == 1


@pytest.mark.gpu
def test_obj_to_bvh_to_bench_on_gpu(tools, tmp_path):
    """The reference's tool chain on a scene of one's own: bvh_extractor -> ray_gen -> bench_traversal, both BVH widths."""
    from oracle import oracle
    bvh, rays = tmp_path / "cornell.bvh", tmp_path / "cornell.rays"
    assert run(tools / "bvh_extractor", GOLDEN / "cornell_box.obj", bvh).returncode == 0
    assert run(tools / "ray_gen", "random", bvh, "50000", "7", rays).returncode == 0
    want = int((oracle.traverse(*formats.load_bvh(bvh, formats.BVH8_TRI4), formats.load_rays(rays, 0.0, 1.0))["tri_id"] >= 0).sum())
    for mode in (["-gpu", "cuda"], ["-gpu", "cuda", "--bvh-width", "4"], ["-s"], []):
        r = run(tools / "bench_traversal", "-bvh", bvh, "-ray", rays, "--tmax", "1", *mode)
        assert r.returncode == 0 and f"{want} intersection(s)" in r.stdout, (mode, r.stdout, r.stderr)      # 50000 rays = 6250 whole packets


def test_shading_tools_argument_errors(tools):
    for tool in ("bench_interface", "bench_shading"):
        r = run(tools / tool, "--bogus")
        assert r.returncode == 1 and "Invalid argument '--bogus'" in r.stderr


@pytest.mark.gpu
def test_bench_interface_and_bench_shading_tools(tools):
    """tools/bench_interface and tools/bench_shading of the reference: no arguments, one line "<x> Mrays/s"
    (bench_interface.cpp:188, bench_shading.cpp:227).  The reference's workload has constant images, so every hit's colour is
    kd / pi = (0.1, 0.2, 0.3) / pi."""
    r = run(tools / "bench_interface", "--iters", "20", "--check")
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert re.fullmatch(r"[\d.e+]+ Mrays/s", lines[0]) and float(lines[0].split()[0]) > 100
    for line in lines[1:5]:
        assert np.allclose([float(x) for x in line.split()], np.array([0.1, 0.2, 0.3]) / np.pi, rtol=1e-5)
    r = run(tools / "bench_shading", "--bench", "5", "--iters", "50")
    assert r.returncode == 0, r.stderr
    assert re.fullmatch(r"[\d.e+]+ Mrays/s", r.stdout.strip()) and float(r.stdout.split()[0]) > 1
