"""On-disk formats at the boundary of the traversal path, byte-compatible with the
reference's readers/writers:

* ``.bvh``  multi-block container  -- tools/common/load_bvh.h:8-74
* ``.rays`` raw 6 x f32 per ray     -- tools/common/load_rays.h:58-92
* ``.fbuf`` raw f32 ``t`` per ray   -- tools/bench_traversal/bench_traversal.cpp:350-354

Arrays are numpy structured arrays whose dtypes mirror ``include/rodent_b200.h``.
"""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

BVH_MAGIC = 0x95CBED1F          # load_bvh.h:24
BVH2_TRI1, BVH4_TRI4, BVH8_TRI4 = 1, 2, 3   # load_bvh.h:8-12

NODE8 = np.dtype([("bounds", "<f4", (6, 8)), ("child", "<i4", (8,)), ("pad", "<i4", (8,))])
NODE4 = np.dtype([("bounds", "<f4", (6, 4)), ("child", "<i4", (4,)), ("pad", "<i4", (4,))])
TRI4 = np.dtype([("v0", "<f4", (3, 4)), ("e1", "<f4", (3, 4)), ("e2", "<f4", (3, 4)), ("n", "<f4", (3, 4)),
                 ("prim_id", "<i4", (4,)), ("geom_id", "<i4", (4,))])
NODE2 = np.dtype([("bounds", "<f4", (12,)), ("child", "<i4", (2,)), ("pad", "<i4", (2,))])
TRI1 = np.dtype([("v0", "<f4", (3,)), ("pad", "<i4"), ("e1", "<f4", (3,)), ("geom_id", "<i4"),
                 ("e2", "<f4", (3,)), ("prim_id", "<i4")])
RAY1 = np.dtype([("org", "<f4", (3,)), ("tmin", "<f4"), ("dir", "<f4", (3,)), ("tmax", "<f4")])
HIT1 = np.dtype([("tri_id", "<i4"), ("t", "<f4"), ("u", "<f4"), ("v", "<f4")])

assert NODE8.itemsize == 256 and NODE4.itemsize == 128 and TRI4.itemsize == 224
assert NODE2.itemsize == 64 and TRI1.itemsize == 48 and RAY1.itemsize == 32 and HIT1.itemsize == 16

_BLOCK_TYPES = {BVH2_TRI1: (NODE2, TRI1), BVH4_TRI4: (NODE4, TRI4), BVH8_TRI4: (NODE8, TRI4)}


def bvh_blocks(data: bytes):
    """Yield (type, node_count, tri_count, payload_offset) for each block (load_bvh.h:27-42)."""
    (magic,) = struct.unpack_from("<I", data, 0)
    if magic != BVH_MAGIC:
        raise ValueError("not a .bvh file (bad magic)")
    pos = 4
    while pos + 20 <= len(data):
        size, typ, n_nodes, n_tris = struct.unpack_from("<QIII", data, pos)
        yield typ, n_nodes, n_tris, pos + 20
        pos += 8 + size


def load_bvh(path, bvh_type: int = BVH8_TRI4):
    """Return (nodes, tris) of the first block of ``bvh_type`` (load_bvh.h:46-74)."""
    data = Path(path).read_bytes()
    node_dt, tri_dt = _BLOCK_TYPES[bvh_type]
    for typ, n_nodes, n_tris, off in bvh_blocks(data):
        if typ == bvh_type:
            nodes = np.frombuffer(data, node_dt, n_nodes, off).copy()
            tris = np.frombuffer(data, tri_dt, n_tris, off + n_nodes * node_dt.itemsize).copy()
            return nodes, tris
    raise ValueError(f"{path}: no block of type {bvh_type}")


def save_bvh(path, blocks) -> None:
    """Write ``blocks`` = [(type, nodes, tris), ...] in the container load_bvh.h reads
    (writer side: tools/bvh_extractor/extract_bvh2.cpp:108-138)."""
    with open(path, "wb") as f:
        f.write(struct.pack("<I", BVH_MAGIC))
        for typ, nodes, tris in blocks:
            payload = nodes.tobytes() + tris.tobytes()
            f.write(struct.pack("<QIII", len(payload) + 12, typ, len(nodes), len(tris)))
            f.write(payload)


def load_rays(path, tmin: float, tmax: float) -> np.ndarray:
    """Ray1 array with the CLI's constant tmin/tmax (load_rays.h:58-92, RayTraits<Ray1>)."""
    raw = np.fromfile(path, "<f4")
    if raw.size % 6:
        raise ValueError("ray file size is not a multiple of 24 bytes")
    return make_rays(raw.reshape(-1, 6), tmin, tmax)


def make_rays(org_dir: np.ndarray, tmin: float, tmax: float) -> np.ndarray:
    rays = np.empty(len(org_dir), RAY1)
    rays["org"] = org_dir[:, 0:3]
    rays["dir"] = org_dir[:, 3:6]
    rays["tmin"] = np.float32(tmin)
    rays["tmax"] = np.float32(tmax)
    return rays


def save_fbuf(path, hits: np.ndarray) -> None:
    np.ascontiguousarray(hits["t"], "<f4").tofile(path)


def fbuf_to_gray(t: np.ndarray, normalize: bool = True) -> np.ndarray:
    """8-bit grey value per ray exactly as tools/fbuf2png/fbuf2png.cpp:82,104:
    ``uint8(255.0f * t / max)`` in fp32 with C truncation."""
    t = np.asarray(t, np.float32)
    tmax = t.max() if normalize else np.float32(1.0)
    return ((np.float32(255.0) * t) / np.float32(tmax)).astype(np.uint8)


def packet_dtypes(width: int):
    """(RayW, HitW) of tools/bench_traversal/bench_traversal.impala:32-65."""
    ray = np.dtype([("org", "<f4", (3, width)), ("dir", "<f4", (3, width)), ("tmin", "<f4", (width,)), ("tmax", "<f4", (width,))])
    hit = np.dtype([("tri_id", "<i4", (width,)), ("t", "<f4", (width,)), ("u", "<f4", (width,)), ("v", "<f4", (width,))])
    assert ray.itemsize == 32 * width and hit.itemsize == 16 * width
    return ray, hit


def pack_rays(rays: np.ndarray, width: int) -> np.ndarray:
    """Ray1 array -> packets, whole packets only (load_rays<RayW>, tools/common/load_rays.h:58-92)."""
    ray_dt, _ = packet_dtypes(width)
    n = len(rays) // width
    r = rays[: n * width].reshape(n, width)
    out = np.zeros(n, ray_dt)
    out["org"] = r["org"].transpose(0, 2, 1)
    out["dir"] = r["dir"].transpose(0, 2, 1)
    out["tmin"] = r["tmin"]
    out["tmax"] = r["tmax"]
    return out


def unpack_hits(hits: np.ndarray) -> np.ndarray:
    """HitW packets -> Hit1 array in ray order."""
    out = np.zeros(hits["tri_id"].size, HIT1)
    for name in ("tri_id", "t", "u", "v"):
        out[name] = hits[name].reshape(-1)
    return out


# ---- the reference's data containers (src/driver/buffer.h, data/bvh.bin) --------------------------------------
def lz4_compress(data: bytes) -> bytes:
    """One LZ4 block (what LZ4_compress_default writes)."""
    import ctypes
    from . import lib
    L = lib.load()
    cap = L.rodent_b200_lz4_compress_bound(len(data))
    out = ctypes.create_string_buffer(cap)
    n = L.rodent_b200_lz4_compress(data, len(data), out, cap)
    if n < 0:
        raise ValueError("LZ4 compression failed")
    return out.raw[:n]


def lz4_decompress(block: bytes, raw_size: int) -> bytes:
    """LZ4_decompress_safe: raises on malformed input or when the block does not decode to exactly raw_size bytes."""
    import ctypes
    from . import lib
    out = ctypes.create_string_buffer(max(raw_size, 1))
    n = lib.load().rodent_b200_lz4_decompress(block, len(block), out, raw_size)
    if n != raw_size:
        raise ValueError("malformed LZ4 block")
    return out.raw[:raw_size]


def load_buffer(path, dtype=np.uint8) -> np.ndarray:
    """A `data/*.bin` buffer of the reference's converter (read_buffer, buffer.h:22-37)."""
    import ctypes
    from . import lib
    L = lib.load()
    size = ctypes.c_int64()
    p = L.rodent_b200_load_buffer(str(path).encode(), ctypes.byref(size))
    if not p:
        raise ValueError(f"{path}: not a buffer file")
    try:
        return np.frombuffer(ctypes.string_at(p, size.value), dtype).copy()
    finally:
        L.rodent_b200_free_buffer(p)


def write_buffer(path, array: np.ndarray) -> None:
    """write_buffer, buffer.h:46-61."""
    from . import lib
    a = np.ascontiguousarray(array)
    if not lib.load().rodent_b200_write_buffer(str(path).encode(), a.ctypes.data, a.nbytes):
        raise OSError(f"cannot write {path}")


def load_bvh_bin(path, bvh_type: int = BVH8_TRI4):
    """(nodes, tris) of the entry of `data/bvh.bin` with this layout (load_bvh<Node, Tri>, interface.cpp:432-454)."""
    import ctypes
    from . import lib
    L = lib.load()
    node_dt, tri_dt = _BLOCK_TYPES[bvh_type]
    nodes, tris, nn, nt = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_int64()
    if not L.rodent_b200_load_bvh_bin(str(path).encode(), node_dt.itemsize, tri_dt.itemsize, ctypes.byref(nodes), ctypes.byref(nn),
                                      ctypes.byref(tris), ctypes.byref(nt)):
        raise ValueError(f"{path}: no BVH entry of type {bvh_type}")
    try:
        return (np.frombuffer(ctypes.string_at(nodes.value, nn.value * node_dt.itemsize), node_dt).copy(),
                np.frombuffer(ctypes.string_at(tris.value, nt.value * tri_dt.itemsize), tri_dt).copy())
    finally:
        L.rodent_b200_free_buffer(nodes)
        L.rodent_b200_free_buffer(tris)


def append_bvh_bin(path, nodes: np.ndarray, tris: np.ndarray) -> None:
    """write_bvh, converter.cpp:428-438 (the file is opened for appending there too)."""
    from . import lib
    nodes, tris = np.ascontiguousarray(nodes), np.ascontiguousarray(tris)
    if not lib.load().rodent_b200_append_bvh_bin(str(path).encode(), nodes.dtype.itemsize, tris.dtype.itemsize, nodes.ctypes.data, len(nodes),
                                                 tris.ctypes.data, len(tris)):
        raise OSError(f"cannot write {path}")
