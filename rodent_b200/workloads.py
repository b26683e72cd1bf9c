"""The render workloads of BASELINE.json (configs[2..4]) as scenes + cameras.

`cornell` is the reference's own testing/cornell_box.obj (committed under tests/golden).
`sponza` has no scene description in the reference -- only the geometry inside
testing/sponza.bvh, every geom_id 0, no normals / uv / materials / lights (SURVEY.md 8d,
config 4) -- so its materials and lights are SYNTHETIC and frozen here:

  * material of a triangle = fnv_hash(prim_id) % 16 over the palette below: diffuse, and
    diffuse + Phong mixes with the mix weight of src/driver/converter.cpp:897-902; only
    BSDF types the reference's converter can emit (converter.cpp:870-913);
  * emitters: the down-facing ceiling triangles (centroid y > 1300, n_y < -0.7) whose
    fnv_hash(prim_id) % 8 == 0, emitting Ke = (17, 12, 4) like the Cornell box light
    (material 16);
  * flat shading (vertex normals = face normals), uv = 0;
  * camera: the one sponza-primary.rays was generated with (README.md:33 of the reference,
    tools/ray_gen/ray_gen.cpp:20-58): eye (-928.012, 483.962, -31.5451), dir +x, up +y, fov 60.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from . import formats, render, testdata

ROOT = Path(__file__).resolve().parent.parent

RENDER_CONFIGS = {
    # BASELINE.json configs[2]: cornell_box.obj path trace 1024x1024, 64 spp, 4 bounces
    "cornell": dict(width=1024, height=1024, spp=64, max_path_len=4,
                    eye=(0.0, 1.0, 2.7), dir=(0.0, 0.0, -1.0), up=(0.0, 1.0, 0.0), fov=60.0),
    # configs[3]: Sponza path trace 1920x1080, 256 spp, 8 bounces
    "sponza": dict(width=1920, height=1080, spp=256, max_path_len=8,
                   eye=(-928.012, 483.962, -31.5451), dir=(1.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fov=60.0),
    # configs[4]: the same scene at 3840x2160, 1024 spp, pixel rows dealt out across the GPUs
    "sponza4k": dict(width=3840, height=2160, spp=1024, max_path_len=8,
                     eye=(-928.012, 483.962, -31.5451), dir=(1.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fov=60.0),
}

_PALETTE_KD = [(0.80, 0.80, 0.80), (0.63, 0.06, 0.05), (0.14, 0.45, 0.09), (0.75, 0.71, 0.62),
               (0.55, 0.50, 0.45), (0.35, 0.38, 0.55), (0.70, 0.55, 0.35), (0.45, 0.45, 0.45),
               (0.60, 0.60, 0.55), (0.50, 0.30, 0.20), (0.25, 0.40, 0.45), (0.65, 0.62, 0.50),
               (0.72, 0.72, 0.72), (0.40, 0.25, 0.30), (0.30, 0.50, 0.30), (0.58, 0.48, 0.40)]
LIGHT_KE = (17.0, 12.0, 4.0)          # testing/cornell_box.mtl `light`


def fnv_hash_u32(values: np.ndarray) -> np.ndarray:
    """fnv_hash(fnv_init, v) of src/core/random.impala:116-126: four FNV-1a rounds over the bytes of v."""
    h = np.full(values.shape, 0x811C9DC5, np.uint64)
    v = values.astype(np.uint64)
    for shift in (0, 8, 16, 24):
        h = ((h * np.uint64(0x01000193)) & np.uint64(0xFFFFFFFF)) ^ ((v >> np.uint64(shift)) & np.uint64(0xFF))
    return h.astype(np.uint32)


def sponza_materials() -> np.ndarray:
    mats = np.zeros(17, render.MATERIAL)
    for i, kd in enumerate(_PALETTE_KD):
        m = mats[i]
        m["kd"] = kd
        m["ns"], m["ni"] = 1.0, 1.0
        m["tf"] = (1.0, 1.0, 1.0)
        if i % 4 == 3:                       # every fourth: diffuse + Phong mix (converter.cpp:897-902)
            ks = (0.25, 0.25, 0.25)
            m["ks"] = ks
            m["ns"] = 32.0
            lum = lambda c: c[0] * 0.2126 + c[1] * 0.7152 + c[2] * 0.0722
            m["mix_k"] = np.float32(lum(ks)) / (np.float32(lum(ks)) + np.float32(lum(kd)))
            m["bsdf"] = render.BSDF_MIX
        else:
            m["bsdf"] = render.BSDF_DIFFUSE
    light = mats[16]
    light["bsdf"], light["is_emissive"] = render.BSDF_DIFFUSE, 1
    light["kd"], light["ke"], light["tf"], light["ns"], light["ni"] = (0.78, 0.78, 0.78), LIGHT_KE, (1.0, 1.0, 1.0), 1.0, 1.0
    return mats


def sponza_material_of_prim(tris: np.ndarray) -> np.ndarray:
    pid = tris["prim_id"].reshape(-1)
    valid = pid != -1
    prim = (pid[valid] & 0x7FFFFFFF).astype(np.int64)
    num_prims = int(prim.max()) + 1
    comp = lambda field: np.stack([tris[field][:, c, :].reshape(-1) for c in range(3)], 1)[valid].astype(np.float64)
    v0, e1, e2 = comp("v0"), comp("e1"), comp("e2")
    v1, v2 = v0 - e1, v0 + e2                                 # make_tri, src/traversal/intersection.impala:110-119
    n = np.cross(v1 - v0, v2 - v0)
    ny = n[:, 1] / np.maximum(np.linalg.norm(n, axis=1), 1e-30)
    cy = (v0[:, 1] + v1[:, 1] + v2[:, 1]) / 3.0
    h = fnv_hash_u32(np.arange(num_prims, dtype=np.uint32))
    material = (h % 16).astype(np.int32)
    ceiling = np.zeros(num_prims, bool)
    ceiling[prim[(cy > 1300.0) & (ny < -0.7)]] = True
    material[ceiling & (h % 8 == 0)] = 16
    return material


_scene_cache: dict = {}


def load_scene(name: str) -> render.Scene:
    key = "sponza" if name.startswith("sponza") else name
    if key not in _scene_cache:
        # Sponza carries a BVH2 / Tri1 as well -- the BVH2 block of the reference's own file, the layout its GPU device
        # renders from: the renderer then traces through it (1.7x the samples/s of the BVH8 walk).
        # The 36-triangle Cornell box gains nothing from it (2.01 .. 2.06 Gsamples/s either way) and keeps the BVH8.
        if key == "cornell":
            _scene_cache[key] = render.Scene.load_obj(ROOT / "tests" / "golden" / "cornell_box.obj")
        elif key == "sponza":
            nodes, tris = formats.load_bvh(testdata.sponza_bvh8(), formats.BVH8_TRI4)
            _scene_cache[key] = render.Scene.from_bvh8(nodes, tris, sponza_materials(), sponza_material_of_prim(tris))
            _scene_cache[key].set_bvh2(*formats.load_bvh(testdata.sponza_bvh2(), formats.BVH2_TRI1))
        else:
            raise KeyError(name)
    return _scene_cache[key]


def camera(name: str, width: int | None = None, height: int | None = None) -> render.Settings:
    cfg = RENDER_CONFIGS[name]
    return render.camera(cfg["eye"], cfg["dir"], cfg["up"], cfg["fov"], width or cfg["width"], height or cfg["height"])
