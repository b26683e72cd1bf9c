"""Materialises the benchmark inputs of BASELINE.json (Sponza BVH8 + the two ray
sets) under ``_data/`` from what the repo commits, without touching the reference
tree: the BVH8 block is unpacked from tests/golden/sponza_bvh8.bvh.xz and the ray
files are regenerated with tools/ray_gen (bit-identical to the reference's
testing/sponza-{primary,random}.rays).
"""
from __future__ import annotations

import lzma
import os
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
DATA = ROOT / "_data"
GOLDEN = ROOT / "tests" / "golden"

# name -> (tmin, tmax) of the reference's README commands (README.md:33-36)
RAY_SETS = {"primary": (0.0, 5000.0), "random": (0.0, 1.0)}

_PRIMARY_ARGS = "-928.012 483.962 -31.5451 1 0 0 0 1 0 60 1024 1024".split()
_RANDOM_ARGS = ["1048576", "42"]


def _atomic_write(path: Path, producer) -> None:
    tmp = path.with_name(f".{path.name}.{os.getpid()}.tmp")
    producer(tmp)
    os.replace(tmp, path)


def ray_gen_bin() -> Path:
    exe = ROOT / "tools" / "bin" / "ray_gen"
    if not exe.exists():
        from . import build
        build.build_tools()
    return exe


def sponza_bvh8() -> Path:
    out = DATA / "sponza_bvh8.bvh"
    if not out.exists():
        DATA.mkdir(exist_ok=True)
        blob = lzma.decompress((GOLDEN / "sponza_bvh8.bvh.xz").read_bytes())
        _atomic_write(out, lambda p: p.write_bytes(blob))
    return out


def sponza_bvh4() -> Path:
    out = DATA / "sponza_bvh4.bvh"
    if not out.exists():
        DATA.mkdir(exist_ok=True)
        blob = lzma.decompress((GOLDEN / "sponza_bvh4.bvh.xz").read_bytes())
        _atomic_write(out, lambda p: p.write_bytes(blob))
    return out


def sponza_bvh2() -> Path:
    """The BVH2 / Tri1 block of the reference's testing/sponza.bvh (the layout of its GPU path)."""
    out = DATA / "sponza_bvh2.bvh"
    if not out.exists():
        DATA.mkdir(exist_ok=True)
        blob = lzma.decompress((GOLDEN / "sponza_bvh2.bvh.xz").read_bytes())
        _atomic_write(out, lambda p: p.write_bytes(blob))
    return out


def rays(name: str) -> Path:
    out = DATA / f"sponza-{name}.rays"
    if not out.exists():
        DATA.mkdir(exist_ok=True)
        if name == "primary":
            args = ["primary", *_PRIMARY_ARGS]
        elif name == "random":
            args = ["random", str(sponza_bvh8()), *_RANDOM_ARGS]
        else:
            raise KeyError(name)
        _atomic_write(out, lambda p: subprocess.run([str(ray_gen_bin()), *args, str(p)], check=True))
    return out
