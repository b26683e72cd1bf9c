"""Build recipes: the sm_100a shared library (C ABI of include/rodent_b200.h) and the
C++ command-line tools.  Everything is built in-tree so the binaries travel with a
snapshot of the repo; nothing is JIT-compiled at run time.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "rodent_b200"
CSRC = PKG / "csrc"
LIB = PKG / "librodent_b200.so"
TOOLS = ROOT / "tools"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",            # parity contract: no FMA contraction anywhere
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def _newer(target: Path, sources) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(s).stat().st_mtime <= t for s in sources)


def cuda_sources():
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cpp"))


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    srcs = cuda_sources()
    deps = srcs + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [ROOT / "include" / "rodent_b200.h"]
    if not force and _newer(LIB, deps):
        return LIB
    cmd = [NVCC, *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", str(LIB), *map(str, srcs), "-lz"]
    print("+", " ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB


SHIM = PKG / "librodent_b200_refnames.so"


def build_shim(force: bool = False) -> Path:
    """librodent_b200_refnames.so: the entry points under the reference's own exported names (cpu_*, nvvm_*), forwarding to librodent_b200.so."""
    src = PKG / "shim" / "reference_names.cpp"
    if not force and _newer(SHIM, [src, ROOT / "include" / "rodent_b200.h"]):
        return SHIM
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-Wall", "-fPIC", "-shared", "-o", str(SHIM), str(src),
           f"-L{PKG}", "-lrodent_b200", "-Wl,-rpath,$ORIGIN"]
    print("+", " ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return SHIM


def build_tools(force: bool = False) -> None:
    """ray_gen, fbuf2png, and -- linked against librodent_b200.so -- bench_traversal, bench_interface, bench_shading, bvh_extractor, converter, rodent."""
    (TOOLS / "bin").mkdir(exist_ok=True)
    cxx = os.environ.get("CXX", "g++")
    common = [cxx, "-O2", "-std=c++17", "-ffp-contract=off", "-Wall"]
    jobs = [("ray_gen", ["ray_gen.cpp"], [])]
    if (TOOLS / "bench_traversal.cpp").exists():
        jobs.append(("bench_traversal", ["bench_traversal.cpp"],
                     [f"-L{PKG}", "-lrodent_b200", f"-Wl,-rpath,{PKG}", "-Wl,-rpath,$ORIGIN/../../rodent_b200"]))
    if (TOOLS / "fbuf2png.cpp").exists():
        jobs.append(("fbuf2png", ["fbuf2png.cpp"], ["-lz"]))
    if (TOOLS / "bvh_extractor.cpp").exists():
        jobs.append(("bvh_extractor", ["bvh_extractor.cpp"],
                     [f"-L{PKG}", "-lrodent_b200", f"-Wl,-rpath,{PKG}", "-Wl,-rpath,$ORIGIN/../../rodent_b200"]))
    for name in ("bench_interface", "bench_shading", "converter"):
        if (TOOLS / f"{name}.cpp").exists():
            jobs.append((name, [f"{name}.cpp"],
                         [f"-L{PKG}", "-lrodent_b200", f"-Wl,-rpath,{PKG}", "-Wl,-rpath,$ORIGIN/../../rodent_b200"]))
    if (TOOLS / "rodent.cpp").exists():
        jobs.append(("rodent", ["rodent.cpp"],
                     [f"-L{PKG}", "-lrodent_b200", "-lz", f"-Wl,-rpath,{PKG}", "-Wl,-rpath,$ORIGIN/../../rodent_b200"]))
    for name, srcs, extra in jobs:
        out = TOOLS / "bin" / name
        deps = [TOOLS / s for s in srcs] + [TOOLS / "formats.h", ROOT / "include" / "rodent_b200.h"]
        if not force and _newer(out, deps):
            continue
        cmd = [*common, "-o", str(out), *[str(TOOLS / s) for s in srcs], *extra]
        print("+", " ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_cuda(force, verbose)
    build_shim(force)
    build_tools(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
