"""Host-side mirror of the reference's `rodent` driver (src/driver/driver.cpp): scene
loading, camera set-up, the render loop and tone mapping, over the C ABI of
include/rodent_b200.h (scene + wavefront path tracer)."""
from __future__ import annotations

import ctypes
import math
from ctypes import POINTER, c_float, c_int32, c_int64, c_void_p

import numpy as np

from . import lib

MATERIAL = np.dtype([("bsdf", "<i4"), ("is_emissive", "<i4"), ("ns", "<f4"), ("ni", "<f4"), ("kd", "<f4", (3,)), ("mix_k", "<f4"),
                     ("ks", "<f4", (3,)), ("map_kd", "<i4"), ("tf", "<f4", (3,)), ("map_ks", "<i4"), ("ke", "<f4", (3,)), ("map_ke", "<i4")])
TEXTURE = np.dtype([("width", "<i4"), ("height", "<i4"), ("offset", "<i8")])
LIGHT = np.dtype([("v0", "<f4", (3,)), ("inv_area", "<f4"), ("v1", "<f4", (3,)), ("prim", "<i4"), ("v2", "<f4", (3,)), ("map_ke", "<i4"),
                  ("n", "<f4", (3,)), ("pad2", "<f4"), ("color", "<f4", (3,)), ("pad3", "<f4")])
assert MATERIAL.itemsize == 80 and LIGHT.itemsize == 80
BSDF_BLACK, BSDF_DIFFUSE, BSDF_PHONG, BSDF_MIX, BSDF_MIRROR, BSDF_GLASS = range(6)


class Vec3(ctypes.Structure):
    _fields_ = [("x", c_float), ("y", c_float), ("z", c_float)]


class Settings(ctypes.Structure):
    """src/dummy_main.impala:3-13"""
    _fields_ = [("eye", Vec3), ("dir", Vec3), ("up", Vec3), ("right", Vec3), ("width", c_float), ("height", c_float)]


class SceneView(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in ("num_tris", "num_vertices", "num_materials", "num_lights", "num_nodes", "num_tri4")] + \
               [(n, c_void_p) for n in ("vertices", "normals", "face_normals", "texcoords", "indices", "light_ids",
                                        "materials", "lights", "nodes", "tris", "textures", "texture_pixels")] + \
               [("num_texture_pixels", c_int64), ("num_textures", c_int32), ("pad", c_int32)] + \
               [("nodes2", c_void_p), ("tris1", c_void_p), ("num_nodes2", c_int32), ("num_tri1", c_int32)]


# symbol -> (restype, argtypes) of the scene / renderer part of include/rodent_b200.h
SIGNATURES = {
    "rodent_b200_scene_load_obj": (c_void_p, [ctypes.c_char_p]),
    "rodent_b200_scene_from_bvh8": (c_void_p, [c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_int32]),
    "rodent_b200_scene_load_data": (c_void_p, [ctypes.c_char_p, ctypes.c_char_p]),
    "rodent_b200_scene_write_data": (c_int32, [c_void_p, ctypes.c_char_p, c_int32, c_int32, ctypes.c_char_p]),
    "rodent_b200_scene_view": (None, [c_void_p, POINTER(SceneView)]),
    "rodent_b200_scene_free": (None, [c_void_p]),
    "rodent_b200_scene_add_texture": (c_int32, [c_void_p, c_void_p, c_int32, c_int32]),
    "rodent_b200_scene_add_png": (c_int32, [c_void_p, ctypes.c_char_p]),
    "rodent_b200_scene_add_jpg": (c_int32, [c_void_p, ctypes.c_char_p]),
    "rodent_b200_scene_add_tga": (c_int32, [c_void_p, ctypes.c_char_p]),
    "rodent_b200_scene_build_bvh2": (None, [c_void_p]),
    "rodent_b200_scene_rebuild_bvh8": (None, [c_void_p]),
    "rodent_b200_scene_set_bvh2": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int32]),
    "rodent_b200_scene_bvh4": (None, [c_void_p, POINTER(c_void_p), POINTER(c_int32), POINTER(c_void_p), POINTER(c_int32)]),
    "rodent_b200_renderer_create": (c_void_p, [c_void_p] + [c_int32] * 8),
    "rodent_b200_renderer_create_multi": (c_void_p, [c_void_p, POINTER(c_int32), c_int32] + [c_int32] * 5),
    "rodent_b200_bind_multi": (None, [c_void_p, POINTER(c_int32), c_int32, c_int32, c_int32]),
    "rodent_b200_renderer_free": (None, [c_void_p]),
    "rodent_b200_render": (None, [c_void_p, POINTER(Settings), c_int32]),
    "rodent_b200_render_device": (None, [c_void_p, POINTER(Settings), c_int32]),
    "rodent_b200_present": (None, [c_void_p]),
    "rodent_b200_film": (POINTER(c_float), [c_void_p]),
    "rodent_b200_film_device": (c_void_p, [c_void_p]),
    "rodent_b200_renderer_bind_film": (None, [c_void_p, c_void_p]),
    "rodent_b200_clear": (None, [c_void_p]),
    "rodent_b200_render_stats": (None, [c_void_p, POINTER(c_int64)]),
    "rodent_b200_render_last_ms": (ctypes.c_double, [c_void_p]),
    "rodent_b200_bind": (None, [c_void_p, c_int32, c_int32, c_int32]),
    "setup_interface": (None, [ctypes.c_size_t, ctypes.c_size_t]),
    "cleanup_interface": (None, []),
    "get_pixels": (POINTER(c_float), []),
    "clear_pixels": (None, []),
    "get_spp": (c_int32, []),
    "render": (None, [POINTER(Settings), c_int32]),
}


def _bind(L):
    if not getattr(L, "_render_bound", False):
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        L._render_bound = True
    return L


def camera(eye, direction, up, fov: float, width: int, height: int) -> Settings:
    """Camera::Camera of src/driver/driver.cpp:31-38 in fp32."""
    f = np.float32
    d = np.asarray(direction, f)
    u = np.asarray(up, f)

    def norm(v):
        return (v * (f(1.0) / np.sqrt((v * v).sum(dtype=f), dtype=f))).astype(f)

    d = norm(d)
    r = norm(np.cross(d, u).astype(f))
    u = norm(np.cross(r, d).astype(f))
    w = f(math.tan(float(f(fov) * f(3.14159265359) / f(360.0))))
    h = f(w / f(f(width) / f(height)))
    return Settings(Vec3(*map(float, eye)), Vec3(*map(float, d)), Vec3(*map(float, u)), Vec3(*map(float, r)), float(w), float(h))


class Scene:
    def __init__(self, handle):
        if not handle:
            raise RuntimeError("scene could not be created (see stderr)")
        self.handle = c_void_p(handle)
        self._view = SceneView()
        _bind(lib.load()).rodent_b200_scene_view(self.handle, ctypes.byref(self._view))

    @classmethod
    def load_obj(cls, path) -> "Scene":
        return cls(_bind(lib.load()).rodent_b200_scene_load_obj(str(path).encode()))

    @classmethod
    def load_data(cls, data_dir, obj_file=None) -> "Scene":
        """A scene from the reference converter's data/ directory (rodent_b200_scene_load_data); materials come from the
        OBJ / MTL named by `obj_file` or by data/bvh.stamp."""
        return cls(_bind(lib.load()).rodent_b200_scene_load_data(str(data_dir).encode(), str(obj_file).encode() if obj_file else None))

    def write_data(self, data_dir, bvh_arity: int = 8, padded: bool = False, obj_file="") -> None:
        """Writes the converter's data/ directory for this scene (rodent_b200_scene_write_data)."""
        if not _bind(lib.load()).rodent_b200_scene_write_data(self.handle, str(data_dir).encode(), bvh_arity, int(padded), str(obj_file).encode()):
            raise RuntimeError(f"cannot write {data_dir} (see stderr)")

    @classmethod
    def from_bvh8(cls, nodes: np.ndarray, tris: np.ndarray, materials: np.ndarray, material_of_prim: np.ndarray) -> "Scene":
        materials = np.ascontiguousarray(materials, MATERIAL)
        mop = np.ascontiguousarray(material_of_prim, np.int32)
        return cls(_bind(lib.load()).rodent_b200_scene_from_bvh8(nodes.ctypes.data, len(nodes), tris.ctypes.data, len(tris),
                                                                  materials.ctypes.data, len(materials), mop.ctypes.data, len(mop)))

    def build_bvh2(self) -> None:
        """A BVH2 / Tri1 over the scene's triangles from the scene's own builder; renderers created afterwards trace their
        closest-hit rays through it (rodent_b200_scene_build_bvh2)."""
        L = _bind(lib.load())
        L.rodent_b200_scene_build_bvh2(self.handle)
        L.rodent_b200_scene_view(self.handle, ctypes.byref(self._view))

    def rebuild_bvh8(self) -> None:
        """Replaces the scene's BVH8 by one from this library's builder over its own triangles."""
        L = _bind(lib.load())
        L.rodent_b200_scene_rebuild_bvh8(self.handle)
        L.rodent_b200_scene_view(self.handle, ctypes.byref(self._view))

    def set_bvh2(self, nodes: np.ndarray, tris: np.ndarray) -> None:
        """Adopts a BVH2 / Tri1 over the same triangles, e.g. the BVH2 block of the .bvh file the scene came from."""
        from . import formats
        nodes, tris = np.ascontiguousarray(nodes, formats.NODE2), np.ascontiguousarray(tris, formats.TRI1)
        L = _bind(lib.load())
        if not L.rodent_b200_scene_set_bvh2(self.handle, nodes.ctypes.data, len(nodes), tris.ctypes.data, len(tris)):
            raise RuntimeError("BVH2 rejected (see stderr)")
        L.rodent_b200_scene_view(self.handle, ctypes.byref(self._view))

    def add_texture(self, rgba: np.ndarray) -> int:
        """Appends an (height, width) uint32 image (RodentTexture layout: gamma-corrected, bottom row first); returns the
        value for a material's map_kd / map_ks."""
        rgba = np.ascontiguousarray(rgba, np.uint32)
        L = _bind(lib.load())
        tid = L.rodent_b200_scene_add_texture(self.handle, rgba.ctypes.data, rgba.shape[1], rgba.shape[0])
        L.rodent_b200_scene_view(self.handle, ctypes.byref(self._view))
        return tid

    def add_png(self, path) -> int:
        """Decodes a .png (or .jpg / .jpeg / .tga) file the way the reference's loaders do and appends it."""
        L = _bind(lib.load())
        name = str(path).lower()
        fn = L.rodent_b200_scene_add_jpg if name.endswith((".jpg", ".jpeg")) else L.rodent_b200_scene_add_tga if name.endswith(".tga") \
            else L.rodent_b200_scene_add_png
        tid = fn(self.handle, str(path).encode())
        if not tid:
            raise RuntimeError(f"cannot load {path} (see stderr)")
        L.rodent_b200_scene_view(self.handle, ctypes.byref(self._view))
        return tid

    @property
    def view(self) -> SceneView:
        return self._view

    def array(self, name: str) -> np.ndarray:
        from . import formats
        v = self._view
        table = {"vertices": ((v.num_vertices, 4), np.float32), "normals": ((v.num_vertices, 4), np.float32),
                 "face_normals": ((v.num_tris, 4), np.float32), "texcoords": ((v.num_vertices, 4), np.float32),
                 "indices": ((v.num_tris, 4), np.int32), "light_ids": ((v.num_tris,), np.int32),
                 "materials": ((v.num_materials,), MATERIAL), "lights": ((v.num_lights,), LIGHT),
                 "nodes": ((v.num_nodes,), formats.NODE8), "tris": ((v.num_tri4,), formats.TRI4),
                 "textures": ((v.num_textures,), TEXTURE), "texture_pixels": ((v.num_texture_pixels,), np.uint32),
                 "nodes2": ((v.num_nodes2,), formats.NODE2), "tris1": ((v.num_tri1,), formats.TRI1)}
        shape, dt = table[name]
        n = int(np.prod(shape)) * np.dtype(dt).itemsize
        if n == 0:
            return np.zeros(shape, dt)
        buf = (ctypes.c_char * n).from_address(getattr(v, name))
        return np.frombuffer(buf, dt).reshape(shape)

    def bvh4(self):
        """(nodes, tris) of the scene's triangles under a BVH4 (rodent_b200_scene_bvh4), as numpy copies."""
        from . import formats
        nodes, tris, nn, nt = c_void_p(), c_void_p(), c_int32(), c_int32()
        _bind(lib.load()).rodent_b200_scene_bvh4(self.handle, ctypes.byref(nodes), ctypes.byref(nn), ctypes.byref(tris), ctypes.byref(nt))
        view = lambda ptr, n, dt: np.frombuffer((ctypes.c_char * (n * dt.itemsize)).from_address(ptr), dt).copy()
        return view(nodes.value, nn.value, formats.NODE4), view(tris.value, nt.value, formats.TRI4)

    def free(self):
        if self.handle:
            lib.load().rodent_b200_scene_free(self.handle)
            self.handle = None


class Renderer:
    """Wavefront path tracer on one device (rodent_b200_renderer_*)."""

    def __init__(self, scene: Scene, dev, width: int, height: int, spp: int, max_path_len: int,
                 part: int = 0, num_parts: int = 1, band: int = 8):
        """`dev`: a device index, or a list of them (one process driving several devices: row bands dealt out over them,
        one ncclReduce per render call, rodent_b200_renderer_create_multi)."""
        self.L = _bind(lib.load())
        self.scene, self.width, self.height, self.spp = scene, width, height, spp
        if isinstance(dev, (list, tuple)):
            # One libnccl.so.2 per process (csrc/nccl_dyn.h binds whichever is loaded, else the system's): a process that
            # imports PyTorch AFTER the first multi-device call would find the system library under PyTorch's soname and
            # fail to import.  So PyTorch, where installed, goes first.
            import importlib.util
            if importlib.util.find_spec("torch") is not None:
                import torch  # noqa: F401
            devs = (c_int32 * len(dev))(*dev)
            self.handle = c_void_p(self.L.rodent_b200_renderer_create_multi(scene.handle, devs, len(dev), width, height, spp, max_path_len, band))
        else:
            self.handle = c_void_p(self.L.rodent_b200_renderer_create(scene.handle, dev, width, height, spp, max_path_len, part, num_parts, band))
        if not self.handle:
            raise RuntimeError("renderer could not be created")

    def render(self, settings: Settings, iteration: int, present: bool = True) -> float:
        (self.L.rodent_b200_render if present else self.L.rodent_b200_render_device)(self.handle, ctypes.byref(settings), iteration)
        return self.L.rodent_b200_render_last_ms(self.handle)

    def present(self):
        self.L.rodent_b200_present(self.handle)

    def clear(self):
        self.L.rodent_b200_clear(self.handle)

    def film(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.L.rodent_b200_film(self.handle), (self.height, self.width, 3))

    def film_device_ptr(self) -> int:
        return self.L.rodent_b200_film_device(self.handle)

    def bind_film(self, device_ptr: int | None) -> None:
        """Accumulate into caller-owned device memory (e.g. a torch tensor's data_ptr()); None: own film."""
        self.L.rodent_b200_renderer_bind_film(self.handle, c_void_p(device_ptr) if device_ptr else None)

    def stats(self) -> dict:
        out = (c_int64 * 5)()
        self.L.rodent_b200_render_stats(self.handle, out)
        return dict(zip(("samples", "primary_rays", "shadow_rays", "wavefronts", "kernels"), map(int, out)))

    def free(self):
        if self.handle:
            self.L.rodent_b200_renderer_free(self.handle)
            self.handle = None


def tonemap(film: np.ndarray, iterations: int) -> np.ndarray:
    """save_image of src/driver/driver.cpp:138-162: clamp((film / iter) ^ (1/2.2)) * 255, truncated to 8 bit."""
    f = np.float32
    x = np.power(np.maximum(film.astype(f) * f(1.0 / iterations), f(0)), f(1.0 / 2.2), dtype=f)
    return (np.clip(x, f(0), f(1)) * f(255.0)).astype(np.uint8)
