"""Host-side mirror of tools/bench_interface/bench_interface.cpp: the quad mesh, the three textures, the
1 Mi random hits, and the call of `bench_interface` (include/rodent_b200.h)."""
from __future__ import annotations

import ctypes
from ctypes import POINTER, c_float, c_int32, c_uint32, c_void_p

import numpy as np

from . import lib

TRI_HIT = np.dtype([("id", "<i4"), ("uv", "<f4", (2,))])
assert TRI_HIT.itemsize == 12


class Color(ctypes.Structure):
    _fields_ = [("r", c_float), ("g", c_float), ("b", c_float)]


class Tex(ctypes.Structure):
    """bench_interface.impala:8-15"""
    _fields_ = [("pixels", c_void_p), ("border_color", Color), ("border", c_uint32), ("sampler", c_uint32),
                ("width", c_int32), ("height", c_int32)]


class ShadedMesh(ctypes.Structure):
    """bench_interface.impala:17-26"""
    _fields_ = [("vertices", c_void_p), ("indices", c_void_p), ("normals", c_void_p), ("texcoords", c_void_p),
                ("tex_kd", Tex), ("tex_ks", Tex), ("tex_ns", Tex)]


assert ctypes.sizeof(Tex) == 40 and ctypes.sizeof(ShadedMesh) == 152

SIGNATURES = {"bench_interface": (None, [POINTER(ShadedMesh), c_void_p, c_void_p, c_void_p, c_void_p, c_int32])}

BORDER_CLAMP, BORDER_REPEAT, BORDER_CONSTANT = 0, 1, 2
SAMPLER_NEAREST, SAMPLER_BILINEAR = 0, 1


def _bind(L):
    fn = L.bench_interface
    fn.restype, fn.argtypes = SIGNATURES["bench_interface"]
    return L


def quad_arrays():
    """The quad of bench_interface.cpp:63-95."""
    vertices = np.array([(-1, 1, 0), (-1, -1, 0), (1, -1, 0), (1, 1, 0)], np.float32)
    normals = np.array([(0, 0, 1)] * 4, np.float32)
    texcoords = np.array([(-1, 1), (-1, -1), (1, -1), (1, 1)], np.float32)
    indices = np.array([0, 1, 2, -1, 2, 3, 0, -1], np.int32)
    return vertices, normals, texcoords, indices


def reference_textures(width: int = 1024, height: int = 1024):
    """(pixels, border colour, border, sampler) x 3 of bench_interface.cpp:97-131."""
    fill = lambda c: np.tile(np.array(c, np.float32), (width * height, 1))
    return ((fill((0.1, 0.2, 0.3)), (0.0, 0.0, 0.0), BORDER_CLAMP, SAMPLER_BILINEAR),
            (fill((1.0, 0.5, 0.1)), (0.5, 1.0, 0.2), BORDER_CONSTANT, SAMPLER_NEAREST),
            (fill((0.1, 0.5, 1.0)), (0.0, 0.0, 0.0), BORDER_REPEAT, SAMPLER_BILINEAR))


def reference_hits(n: int = 1 << 20, seed: int = 42):
    """Random hits and directions in the manner of bench_interface.cpp:148-170 (numpy's generator, not mt19937)."""
    rng = np.random.default_rng(seed)
    hits = np.zeros(n, TRI_HIT)
    hits["id"] = np.arange(n) % 2
    hits["uv"] = rng.random((n, 2), np.float32)
    unit = lambda a: (a / np.linalg.norm(a, axis=1, keepdims=True)).astype(np.float32)
    return hits, unit(rng.random((n, 3), np.float32) + 1e-3), unit(rng.random((n, 3), np.float32) + 1e-3)


def make_mesh(ptr, arrays, textures, width: int, height: int) -> ShadedMesh:
    """`ptr(array) -> address` decides where the data lives: numpy addresses for the oracle, device uploads for CUDA."""
    vertices, normals, texcoords, indices = arrays
    tex = [Tex(ptr(px), Color(*bc), border, sampler, width, height) for px, bc, border, sampler in textures]
    return ShadedMesh(ptr(vertices), ptr(indices), ptr(normals), ptr(texcoords), *tex)


def run_cuda(arrays, textures, width, height, hits, in_dirs, out_dirs, repeat: int = 1):
    """Uploads everything to device 0, calls bench_interface `repeat` times, returns (colors, seconds per call)."""
    import time
    from .traversal import DeviceArray
    L = _bind(lib.load())
    keep = []

    def upload(a):
        d = DeviceArray.from_host(0, np.ascontiguousarray(a).view(np.uint8).reshape(-1))
        keep.append(d)
        return d.ptr

    mesh = make_mesh(upload, arrays, textures, width, height)
    d_hits, d_in, d_out = upload(hits), upload(in_dirs), upload(out_dirs)
    colors = DeviceArray(0, np.float32, 3 * len(hits))
    L.bench_interface(ctypes.byref(mesh), d_hits, d_in, d_out, colors.ptr, len(hits))
    t0 = time.perf_counter()
    for _ in range(repeat):
        L.bench_interface(ctypes.byref(mesh), d_hits, d_in, d_out, colors.ptr, len(hits))
    dt = (time.perf_counter() - t0) / repeat
    return colors.to_host().reshape(-1, 3), dt
