"""ctypes binding of librodent_b200.so (the C ABI of include/rodent_b200.h).

Loading fails loudly when the library is missing or does not export a declared
symbol; nothing here falls back to a CPU implementation.
"""
from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_double, c_int32, c_int64, c_size_t, c_void_p
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "librodent_b200.so"

_TRAVERSE_DEV = [c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int32]
_TRAVERSE_HOST = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32]

# symbol -> (restype, argtypes); must list every declaration of include/rodent_b200.h
SIGNATURES = {
    "cuda_intersect_single_ray1_bvh8_tri4": (None, _TRAVERSE_DEV),
    "cuda_occluded_single_ray1_bvh8_tri4": (None, _TRAVERSE_DEV),
    "cuda_intersect_single_ray1_bvh8_tri4_async": (None, _TRAVERSE_DEV + [c_void_p, c_void_p]),
    "cuda_occluded_single_ray1_bvh8_tri4_async": (None, _TRAVERSE_DEV + [c_void_p, c_void_p]),
    "cuda_intersect_single_ray1_bvh2_tri1": (None, _TRAVERSE_DEV),
    "cuda_occluded_single_ray1_bvh2_tri1": (None, _TRAVERSE_DEV),
    "cuda_intersect_single_ray1_bvh4_tri4": (None, _TRAVERSE_DEV),
    "cuda_occluded_single_ray1_bvh4_tri4": (None, _TRAVERSE_DEV),
    "b200_intersect_single_ray1_bvh4_tri4": (None, _TRAVERSE_HOST),
    "b200_occluded_single_ray1_bvh4_tri4": (None, _TRAVERSE_HOST),
    "b200_intersect_single_ray1_bvh8_tri4": (None, _TRAVERSE_HOST),
    "b200_occluded_single_ray1_bvh8_tri4": (None, _TRAVERSE_HOST),
    "rodent_b200_forget_bvh": (None, [c_void_p, c_void_p]),
    "rodent_b200_bvh_cache_stats": (None, [ctypes.POINTER(c_int64)]),
    "rodent_b200_lz4_decompress": (ctypes.c_int64, [c_void_p, ctypes.c_int64, c_void_p, ctypes.c_int64]),
    "rodent_b200_lz4_compress_bound": (ctypes.c_int64, [ctypes.c_int64]),
    "rodent_b200_lz4_compress": (ctypes.c_int64, [c_void_p, ctypes.c_int64, c_void_p, ctypes.c_int64]),
    "rodent_b200_load_buffer": (c_void_p, [c_char_p, ctypes.POINTER(ctypes.c_int64)]),
    "rodent_b200_free_buffer": (None, [c_void_p]),
    "rodent_b200_write_buffer": (c_int32, [c_char_p, c_void_p, ctypes.c_int64]),
    "rodent_b200_load_bvh_bin": (c_int32, [c_char_p, c_int32, c_int32, ctypes.POINTER(c_void_p), ctypes.POINTER(ctypes.c_int64),
                                          ctypes.POINTER(c_void_p), ctypes.POINTER(ctypes.c_int64)]),
    "rodent_b200_append_bvh_bin": (c_int32, [c_char_p, c_int32, c_int32, c_void_p, ctypes.c_int64, c_void_p, ctypes.c_int64]),
    "rodent_b200_device_count": (c_int32, []),
    "rodent_b200_set_device": (None, [c_int32]),
    "rodent_b200_set_devices": (None, [ctypes.POINTER(c_int32), c_int32]),
    "rodent_b200_ipc_export": (c_int32, [c_int32, c_void_p, c_void_p]),
    "rodent_b200_ipc_open": (c_void_p, [c_int32, c_void_p]),
    "rodent_b200_ipc_close": (None, [c_int32, c_void_p]),
    "rodent_b200_alloc_device": (c_void_p, [c_int32, c_size_t]),
    "rodent_b200_free_device": (None, [c_int32, c_void_p]),
    "rodent_b200_alloc_host": (c_void_p, [c_size_t]),
    "rodent_b200_free_host": (None, [c_void_p]),
    "rodent_b200_pin_host": (c_int32, [c_void_p, c_size_t]),
    "rodent_b200_unpin_host": (c_int32, [c_void_p]),
    "rodent_b200_selftest_host_copies": (c_int32, []),
    "rodent_b200_copy_to_device": (None, [c_int32, c_void_p, c_void_p, c_size_t]),
    "rodent_b200_copy_to_host": (None, [c_int32, c_void_p, c_void_p, c_size_t]),
    "rodent_b200_sync": (None, [c_int32]),
    "rodent_b200_last_kernel_ms": (c_double, [c_int32]),
    "rodent_b200_last_kernel_name": (c_char_p, [c_int32]),
    "rodent_b200_launch_count": (c_int64, []),
    "rodent_b200_set_packet_order": (None, [c_int32]),
    "rodent_b200_version": (c_char_p, []),
}
for _kind in ("packet", "hybrid"):
    for _w in (4, 8):
        for _b in (4, 8):
            for _op in ("intersect", "occluded"):
                SIGNATURES[f"b200_{_op}_{_kind}_ray{_w}_bvh{_b}_tri4"] = (None, _TRAVERSE_HOST)
# experiment knobs, not part of the drop-in surface
_EXTRA = {"rodent_b200_tune": (None, [c_char_p, c_int32])}

_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m rodent_b200.build` "
                "(rodent_b200 has no CPU fallback)")
        lib = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in {**SIGNATURES, **_EXTRA}.items():
            fn = getattr(lib, name)      # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def tune(key: str, value: int) -> None:
    load().rodent_b200_tune(key.encode(), int(value))
