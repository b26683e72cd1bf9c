"""Host-side mirrors of tools/bench_interface/bench_interface.cpp (the quad mesh, the three textures, the 1 Mi random
hits, the call of `bench_interface`) and of tools/bench_shading/bench_shading.cpp (streams, workload, the call of
`b200_bench_shading`), over include/rodent_b200.h."""
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


# ---- bench_shading (tools/bench_shading) ----------------------------------------------------------------
class RayStream(ctypes.Structure):
    """src/render/driver.impala:24-34"""
    _fields_ = [(n, c_void_p) for n in ("id", "org_x", "org_y", "org_z", "dir_x", "dir_y", "dir_z", "tmin", "tmax")]


class PrimaryStream(ctypes.Structure):
    """src/render/driver.impala:36-52"""
    _fields_ = [("rays", RayStream)] + [(n, c_void_p) for n in ("geom_id", "prim_id", "t", "u", "v", "rnd", "mis",
                                                                  "contrib_r", "contrib_g", "contrib_b", "depth")] + \
               [("size", c_int32), ("pad", c_int32)]


assert ctypes.sizeof(PrimaryStream) == 20 * 8 + 8

_SHADING_ARGS = [POINTER(PrimaryStream), POINTER(PrimaryStream)] + [c_void_p] * 6 + [c_int32, c_int32, c_void_p, c_void_p, c_int32, c_int32]
SIGNATURES["b200_bench_shading"] = (None, _SHADING_ARGS)

_FIELDS = ["id", "org_x", "org_y", "org_z", "dir_x", "dir_y", "dir_z", "tmin", "tmax", "geom_id", "prim_id", "t", "u", "v",
           "rnd", "mis", "contrib_r", "contrib_g", "contrib_b", "depth"]
_INT_FIELDS = {"id", "geom_id", "prim_id", "depth"}


class HostStream:
    """A primary stream carved out of one 20 * capacity float buffer (get_primary_stream, bench_shading.cpp:40-54)."""

    def __init__(self, capacity: int):
        self.capacity = capacity
        self.data = np.zeros(20 * capacity, np.float32)
        self.struct = PrimaryStream()
        for k, name in enumerate(_FIELDS):
            addr = self.data.ctypes.data + 4 * k * capacity
            setattr(self.struct.rays if k < 9 else self.struct, name, addr)
        self.struct.size = 0

    def field(self, name: str) -> np.ndarray:
        k = _FIELDS.index(name)
        view = self.data[k * self.capacity:(k + 1) * self.capacity]
        return view.view(np.int32) if name in _INT_FIELDS else view.view(np.uint32) if name == "rnd" else view


def shading_workload(num_rays: int = 4096, seed: int = 42, width: int = 1024, height: int = 1024):
    """The benchmark's input (bench_shading.cpp:61-160): a quad, a checkerboard, rays from (0, 0, -1) onto the quad in
    four geometry ranges.  numpy's generator instead of mt19937, same distribution."""
    vertices, normals, texcoords, indices = quad_arrays()
    face_normals = np.array([(0, 0, 1), (0, 0, 1)], np.float32)
    yy, xx = np.mgrid[0:height, 0:width]
    pixels = np.where((xx + yy) % 2 != 0, np.uint32(0xFFFFFFFF), np.uint32(0)).astype(np.uint32).reshape(-1)
    rng = np.random.default_rng(seed)
    s = HostStream(num_rays)
    per = num_rays // 4
    begins = np.arange(4, dtype=np.int32) * per
    ends = begins + per
    prim = (rng.random(num_rays) >= 0.5).astype(np.int32)
    u, v = rng.random(num_rays, np.float32), rng.random(num_rays, np.float32)
    flip = u + v > 1.0
    u[flip], v[flip] = 1.0 - u[flip], 1.0 - v[flip]
    tri = indices.reshape(-1, 4)[prim][:, :3]
    p = ((1.0 - u - v)[:, None] * vertices[tri[:, 0]] + u[:, None] * vertices[tri[:, 1]] + v[:, None] * vertices[tri[:, 2]]).astype(np.float32)
    org = np.array((0.0, 0.0, -1.0), np.float32)
    d = p - org
    geom = np.repeat(np.arange(4, dtype=np.int32), per)
    s.field("id")[:] = np.arange(num_rays)
    for k, name in enumerate(("org_x", "org_y", "org_z")):
        s.field(name)[:] = org[k]
    for k, name in enumerate(("dir_x", "dir_y", "dir_z")):
        s.field(name)[:] = d[:, k]
    s.field("tmin")[:] = 0.0
    s.field("tmax")[:] = np.finfo(np.float32).max
    s.field("geom_id")[:] = geom
    s.field("prim_id")[:] = prim
    s.field("t")[:] = 1.0
    s.field("u")[:] = u
    s.field("v")[:] = v
    s.field("rnd")[:] = (33 * geom + np.tile(np.arange(per), 4)).astype(np.uint32)
    s.field("mis")[:] = 0.5
    for name in ("contrib_r", "contrib_g", "contrib_b"):
        s.field(name)[:] = 1.0
    s.field("depth")[:] = 0
    s.struct.size = num_rays
    mesh = dict(vertices=vertices, normals=normals, face_normals=face_normals, texcoords=texcoords, indices=indices, pixels=pixels,
                width=width, height=height, begins=begins, ends=ends)
    return s, mesh


def call_bench_shading(fn, stream_in: HostStream, stream_out: HostStream, mesh: dict, num_iters: int) -> None:
    """Calls `fn` (b200_bench_shading or the oracle's twin) with the reference's argument list (bench_shading.cpp:207-222)."""
    fn.restype, fn.argtypes = None, _SHADING_ARGS
    a = lambda name: mesh[name].ctypes.data
    fn(ctypes.byref(stream_in.struct), ctypes.byref(stream_out.struct), a("vertices"), a("normals"), a("face_normals"), a("texcoords"),
       a("indices"), a("pixels"), mesh["width"], mesh["height"], a("begins"), a("ends"), 2, num_iters)
