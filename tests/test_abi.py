"""The C-ABI library loads and exports every symbol include/rodent_b200.h declares
(no compute: this runs without a GPU)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "rodent_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:cuda_|b200_|rodent_b200_|bench_interface|render|get_spp|get_pixels|clear_pixels|setup_interface|cleanup_interface)\w*)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from rodent_b200 import build, lib
    build.build_cuda()
    L = lib.load()
    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/rodent_b200.h but not exported"
    from rodent_b200 import render, shading_bench
    assert set(names) == set(lib.SIGNATURES) | set(render.SIGNATURES) | set(shading_bench.SIGNATURES), "python bindings and header disagree"
    assert L.rodent_b200_version().decode().startswith("rodent_b200")


def test_struct_sizes_match_header():
    from rodent_b200 import formats
    text = (ROOT / "include" / "rodent_b200.h").read_text()
    for name, dt in (("Node8", formats.NODE8), ("Node4", formats.NODE4), ("Tri4", formats.TRI4),
                     ("Ray1", formats.RAY1), ("Hit1", formats.HIT1)):
        assert f"typedef struct {name}" in text
    assert formats.NODE8.itemsize == 256 and formats.TRI4.itemsize == 224
    assert formats.RAY1.itemsize == 32 and formats.HIT1.itemsize == 16


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import pytest
    from rodent_b200 import lib
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lib.load()


def reference_names():
    """What the reference's objects export for this path (`extern fn` of tools/bench_traversal/bench_traversal.impala:159-493,
    tools/bench_shading/bench_shading.impala:22)."""
    names = [f"cpu_{op}_{kind}_ray{w}_bvh{a}_tri4" for a in (4, 8) for kind in ("packet", "hybrid") for w in (4, 8) for op in ("intersect", "occluded")]
    names += [f"cpu_{op}_single_ray1_bvh{a}_tri4" for a in (4, 8) for op in ("intersect", "occluded")]
    names += ["nvvm_intersect_single_ray1_bvh2_tri1", "nvvm_occluded_single_ray1_bvh2_tri1", "cpu_bench_shading"]
    return names


def test_shim_exports_the_reference_names():
    """librodent_b200_refnames.so: linked instead of the reference's generated traversal object, it gives the reference's
    own programs their symbols; it forwards to librodent_b200.so (loaded through its rpath)."""
    import subprocess
    from rodent_b200 import build
    build.build_cuda()
    shim = build.build_shim()
    S = ctypes.CDLL(str(shim))
    names = reference_names()
    assert len(names) == 23
    for name in names:
        assert hasattr(S, name), f"{name} missing from the shim"
    exported = subprocess.run(["nm", "-D", "--defined-only", str(shim)], capture_output=True, text=True, check=True).stdout
    assert sorted(line.split()[-1] for line in exported.splitlines() if " T " in line) == sorted(names)


def test_host_side_of_the_pageable_path_without_a_device():
    """The helper-thread pool and the staging copies (plain and non-temporal stores, odd sizes and alignments, tasks that
    ask to be run again, two callers at once) need no GPU: the library checks them itself."""
    from rodent_b200 import lib
    assert lib.load().rodent_b200_selftest_host_copies() == 0
