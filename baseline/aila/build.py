"""Builds the same-box Aila-Laine comparator from the reference's own sources (see README.md).
    python baseline/aila/build.py           -> baseline/_ref/aila/bench_aila
Nothing from /root/reference is copied into the repository: the rewritten sources exist in a temporary directory for the
duration of the compile; what stays under baseline/_ref/ (git-ignored) is the binary."""
from __future__ import annotations

import os
import re
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path(os.environ.get("RODENT_REFERENCE", "/root/reference")) / "tools" / "bench_aila"
OUT = ROOT / "baseline" / "_ref" / "aila"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
BINARY = OUT / "bench_aila"


def patch_header(text: str) -> str:
    text, n = re.subn(r"texture<float4,\s*1>\s*(t_\w+);", r"__device__ const float4* \1;", text)
    text, m = re.subn(r"texture<int,\s*1>\s*(t_\w+);", r"__device__ const int* \1;", text)
    assert n == 8 and m == 1, f"expected 8 + 1 texture references, found {n} + {m}"
    # after the CUDA headers: every tex1Dfetch(t_X, i) of the kernel becomes a read-only global load
    return text.replace("#define FETCH_GLOBAL", "#define tex1Dfetch(TEX, IDX) __ldg(&(TEX)[IDX])\n#define FETCH_GLOBAL", 1)


def patch_kernel(text: str) -> str:
    text, n = re.subn(r"cudaBindTexture\(nullptr,\s*(t_\w+),\s*(\w+),[^;]*\)\);", r"cudaMemcpyToSymbol(\1, &\2, sizeof(\2)));", text)
    text, m = re.subn(r"^\s*CHECK_CUDA_CALL\(cudaUnbindTexture\(t_\w+\)\);\n", "", text, flags=re.M)
    assert n == 3 and m == 3, f"expected 3 binds and 3 unbinds, found {n} and {m}"
    # The ray fetch broadcasts the warp's base index through a shared-memory word, written by one lane and read by the
    # others with no barrier in between (:121-124) -- warp-synchronous code from before independent thread scheduling.
    # On sm_70+ the lanes of a warp reach this point in groups, the groups race on the word, and ranges of rays are
    # handed out and never traced (measured on B200: 21 % / 29 % of the two Sponza sets left unwritten).  The same
    # broadcast as a shuffle among the lanes that fetch together:
    text, k = re.subn(r"if \(idxTerminated == 0\)\s*\n\s*rayBase = atomicAdd\(&g_warpCounter, numTerminated\);\s*\n\s*rayidx = rayBase \+ idxTerminated;",
                      "int fetchBase = 0;\n            if (idxTerminated == 0)\n                fetchBase = atomicAdd(&g_warpCounter, numTerminated);\n"
                      "            fetchBase = __shfl_sync(maskTerminated, fetchBase, __ffs(maskTerminated) - 1);\n            rayidx = fetchBase + idxTerminated;", text)
    assert k == 1, "ray fetch not found"
    # Both warp votes name the full warp (`__ballot_sync(all_mask, ...)`, a blind translation of Kepler's `__ballot`): one at
    # the head of the persistent loop, one inside the traversal loop where only the lanes still traversing arrive.  The
    # hardware pairs the two sites, so the head's vote also sees the `true` of lanes that are in the traversal loop, counts
    # them as terminated and hands out rays for them.  Kepler's `__ballot` voted among the ACTIVE lanes; so do these:
    text, a = re.subn(r"__ballot_sync\(all_mask, terminated\)", "__ballot_sync(__activemask(), terminated)", text)
    text, b = re.subn(r"__popc\(__ballot_sync\(all_mask, true\)\)", "__popc(__activemask())", text)
    assert a == 1 and b == 1, f"expected the two warp votes, found {a} and {b}"
    return text


def build(force: bool = False) -> Path | None:
    if not REF.exists():
        return BINARY if BINARY.exists() else None
    srcs = [REF / "CudaTracerKernels.hpp", REF / "kepler_dynamic_fetch.cu", HERE / "bench_aila_main.cpp", HERE / "traversal.h", Path(__file__)]
    if not force and BINARY.exists() and all(s.stat().st_mtime <= BINARY.stat().st_mtime for s in srcs):
        return BINARY
    OUT.mkdir(parents=True, exist_ok=True)
    import tempfile
    # the rewritten sources exist only for the duration of the compile, outside the repository: what stays is the binary
    with tempfile.TemporaryDirectory(prefix="aila_build_") as tmp:
        tmp = Path(tmp)
        (tmp / "CudaTracerKernels.hpp").write_text(patch_header((REF / "CudaTracerKernels.hpp").read_text()))
        (tmp / "kepler_dynamic_fetch.cu").write_text(patch_kernel((REF / "kepler_dynamic_fetch.cu").read_text()))
        cmd = [NVCC, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-w",
               f"-I{tmp}", f"-I{HERE}", f"-I{ROOT / 'include'}", f"-I{ROOT / 'tools'}",
               "-o", str(BINARY), str(tmp / "kepler_dynamic_fetch.cu"), str(HERE / "bench_aila_main.cpp")]
        print("+", " ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
    return BINARY


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
