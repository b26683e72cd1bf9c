"""ctypes binding of the CPU oracle (oracle/traversal_oracle.c).

TEST INFRASTRUCTURE: importable only from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg.  The product package rodent_b200
never imports this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_LIB = None


class OracleStats(ctypes.Structure):
    _fields_ = [("nodes", ctypes.c_uint64), ("tri4", ctypes.c_uint64), ("max_stack", ctypes.c_uint64)]


def build(force: bool = False) -> Path:
    so = _DIR / "liboracle.so"
    newest = max((_DIR / f).stat().st_mtime for f in ("traversal_oracle.c", "traversal_bvh2_oracle.c", "render_oracle.c", "shading_bench_oracle.c", "Makefile"))
    if force or not so.exists() or so.stat().st_mtime < newest:
        subprocess.run(["make", "-C", str(_DIR), "-B" if force else "-s"] + (["-s"] if force else []), check=True)
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(str(build()))
        _LIB.oracle_traverse.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int,
                                         ctypes.POINTER(OracleStats)]
        _LIB.oracle_traverse.restype = None
        _LIB.oracle_brute_force.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]
        _LIB.oracle_brute_force.restype = None
        _LIB.oracle_network.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _LIB.oracle_network.restype = ctypes.c_int
    return _LIB


def _ptr(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p)


def traverse(nodes: np.ndarray, tris: np.ndarray, rays: np.ndarray, any_hit: bool = False,
             threads: int | None = None, want_stats: bool = False):
    """Run the oracle; returns hits (and OracleStats when want_stats)."""
    from rodent_b200.formats import HIT1
    arity = nodes.dtype["child"].shape[0]
    hits = np.zeros(len(rays), HIT1)
    if any_hit:
        hits["tri_id"] = -2   # make "written" visible; t/u/v stay 0 (untouched by occluded)
    stats = OracleStats()
    if threads is None:
        threads = os.cpu_count() or 1
    lib().oracle_traverse(arity, int(any_hit), _ptr(nodes), _ptr(tris), _ptr(rays), _ptr(hits), len(rays),
                          threads, ctypes.byref(stats) if want_stats else None)
    return (hits, stats) if want_stats else hits


def traverse_packets(nodes: np.ndarray, tris: np.ndarray, packets: np.ndarray, kind: str = "hybrid", any_hit: bool = False,
                     threads: int | None = None) -> np.ndarray:
    """The reference's packet / hybrid kernel (cpu_traverse_hybrid_helper, oracle/traversal_oracle.c: traverse_packet) on
    Ray4 / Ray8 packets; returns Hit4 / Hit8 packets (any hit: only tri_id is written, the rest stays 0)."""
    from rodent_b200.formats import packet_dtypes
    width = packets.dtype["tmin"].shape[0]
    arity = nodes.dtype["child"].shape[0]
    hits = np.zeros(len(packets), packet_dtypes(width)[1])
    fn = lib().oracle_traverse_packets
    fn.restype = None
    fn.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p] * 4 + [ctypes.c_int32, ctypes.c_int]
    fn(arity, width, int(kind == "hybrid"), int(any_hit), _ptr(nodes), _ptr(tris), _ptr(packets), _ptr(hits), len(packets),
       threads if threads is not None else (os.cpu_count() or 1))
    return hits


def traverse_bvh2(nodes: np.ndarray, tris: np.ndarray, rays: np.ndarray, any_hit: bool = False, threads: int | None = None,
                  want_counters: bool = False):
    """The reference's GPU traversal semantics on its BVH2 / Tri1 layout (oracle/traversal_bvh2_oracle.c)."""
    from rodent_b200.formats import HIT1, NODE2, TRI1
    assert nodes.dtype == NODE2 and tris.dtype == TRI1
    hits = np.zeros(len(rays), HIT1)
    counters = (ctypes.c_uint64 * 2)()
    fn = lib().oracle_traverse_bvh2
    fn.restype = None
    fn.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int, ctypes.c_void_p]
    fn(int(any_hit), _ptr(nodes), _ptr(tris), _ptr(rays), _ptr(hits), len(rays), threads or (os.cpu_count() or 1), counters)
    return (hits, (int(counters[0]), int(counters[1]))) if want_counters else hits


def brute_force(tris: np.ndarray, rays: np.ndarray) -> np.ndarray:
    from rodent_b200.formats import HIT1
    hits = np.zeros(len(rays), HIT1)
    lib().oracle_brute_force(_ptr(tris), len(tris), _ptr(rays), _ptr(hits), len(rays))
    return hits


def network(arity: int, n: int):
    a = np.zeros(32, np.int8); b = np.zeros(32, np.int8)
    k = lib().oracle_network(arity, n, _ptr(a), _ptr(b))
    return [(int(a[i]), int(b[i])) for i in range(k)]


class OracleRenderStats(ctypes.Structure):
    _fields_ = [("samples", ctypes.c_uint64), ("primary_rays", ctypes.c_uint64), ("shadow_rays", ctypes.c_uint64), ("trav", OracleStats)]


def render(scene_view, settings, width: int, height: int, spp: int, max_path_len: int, iteration: int,
           film: np.ndarray | None = None, threads: int | None = None):
    """One render(settings, iter) call of the CPU path tracer; returns (film, stats).  `scene_view` is a
    rodent_b200.render.SceneView (host arrays of a loaded scene), `settings` a rodent_b200.render.Settings."""
    L = lib()
    L.oracle_render.restype = None
    L.oracle_render.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 5 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    if film is None:
        film = np.zeros((height, width, 3), np.float32)
    stats = OracleRenderStats()
    L.oracle_render(ctypes.byref(scene_view), ctypes.byref(settings), width, height, spp, max_path_len, iteration,
                    film.ctypes.data, threads or (os.cpu_count() or 1), ctypes.byref(stats))
    return film, stats


def set_poly_trig(on: bool) -> None:
    """sin / cos of the path tracer from rodent_b200/csrc/poly_trig.h (the device has the same switch) instead of libm."""
    lib().oracle_set_poly_trig(int(on))


def bench_interface(mesh, tri_hits: np.ndarray, in_dirs: np.ndarray, out_dirs: np.ndarray) -> np.ndarray:
    """oracle_bench_interface: `mesh` is a rodent_b200.shading_bench.ShadedMesh whose pointers are HOST pointers."""
    L = lib()
    L.oracle_bench_interface.restype = None
    L.oracle_bench_interface.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int32]
    colors = np.zeros((len(tri_hits), 3), np.float32)
    L.oracle_bench_interface(ctypes.byref(mesh), _ptr(tri_hits), _ptr(in_dirs), _ptr(out_dirs), _ptr(colors), len(tri_hits))
    return colors


def bench_shading_fn():
    """oracle_bench_shading (render_oracle.c): same argument list as cpu_bench_shading."""
    return lib().oracle_bench_shading
