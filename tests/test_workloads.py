"""The frozen synthetic Sponza scene and the render configs of BASELINE.json (rodent_b200/workloads.py)."""
import numpy as np

from rodent_b200 import render, workloads


def c_fnv(h, d):
    """fnv_hash of src/core/random.impala:116-126, spelled out."""
    for shift in (0, 8, 16, 24):
        h = ((h * 16777619) & 0xFFFFFFFF) ^ ((d >> shift) & 0xFF)
    return h


def test_fnv_hash_matches_the_reference_definition():
    values = np.array([0, 1, 2, 255, 256, 65535, 262266, 0xFFFFFFFF], np.uint32)
    got = workloads.fnv_hash_u32(values)
    assert [int(x) for x in got] == [c_fnv(0x811C9DC5, int(v)) for v in values]


def test_render_configs_are_the_baseline_ones():
    c = workloads.RENDER_CONFIGS
    assert (c["cornell"]["width"], c["cornell"]["height"], c["cornell"]["spp"], c["cornell"]["max_path_len"]) == (1024, 1024, 64, 4)
    assert (c["sponza"]["width"], c["sponza"]["height"], c["sponza"]["spp"], c["sponza"]["max_path_len"]) == (1920, 1080, 256, 8)
    assert (c["sponza4k"]["width"], c["sponza4k"]["height"], c["sponza4k"]["spp"], c["sponza4k"]["max_path_len"]) == (3840, 2160, 1024, 8)


def test_sponza_scene_is_frozen():
    scene = workloads.load_scene("sponza")
    v = scene.view
    assert (v.num_tris, v.num_materials, v.num_lights, v.num_nodes, v.num_tri4) == (262267, 17, 271, 15054, 71115)
    mats = scene.array("materials")
    assert set(mats["bsdf"][:16]) == {render.BSDF_DIFFUSE, render.BSDF_MIX} and (mats["bsdf"][3::4][:4] == render.BSDF_MIX).all()
    assert mats["is_emissive"].tolist() == [0] * 16 + [1] and tuple(mats["ke"][16]) == workloads.LIGHT_KE
    assert ((mats["mix_k"][3:16:4] > 0.2) & (mats["mix_k"][3:16:4] < 0.5)).all()
    material = scene.array("indices")[:, 3]
    counts = np.bincount(material, minlength=17)
    assert counts[16] == 271 and counts[:16].min() > 16000 and counts.sum() == 262267
    # the lights are the chosen ceiling triangles: high up, facing down
    lights = scene.array("lights")
    assert (lights["v0"][:, 1] > 1250).all() and (lights["n"][:, 1] < -0.7).all() and (lights["inv_area"] > 0).all()
    # same scene object on a second request
    assert workloads.load_scene("sponza4k") is scene


def test_camera_of_the_sponza_config_is_the_primary_ray_camera():
    cam = workloads.camera("sponza")
    assert (round(cam.eye.x, 3), round(cam.eye.y, 3), round(cam.eye.z, 4)) == (-928.012, 483.962, -31.5451)
    assert (cam.dir.x, cam.dir.y, cam.dir.z) == (1.0, 0.0, 0.0) and (cam.up.x, cam.up.y, cam.up.z) == (0.0, 1.0, 0.0)
    assert abs(cam.width - np.tan(np.pi / 6)) < 1e-6 and abs(cam.height - cam.width * 1080 / 1920) < 1e-6


def test_scene_from_bvh8_rejects_material_ids_outside_the_table(capfd):
    """The ids index device tables and shared-memory bins: one bad id must be an error, not an out-of-bounds write."""
    import pytest
    from rodent_b200 import formats, render, testdata, workloads
    nodes, tris = formats.load_bvh(testdata.sponza_bvh8(), formats.BVH8_TRI4)
    mats = workloads.sponza_materials()
    good = workloads.sponza_material_of_prim(tris)
    for bad_value in (len(mats), -1, 1 << 20):
        bad = good.copy()
        bad[12345] = bad_value
        with pytest.raises(RuntimeError):
            render.Scene.from_bvh8(nodes, tris, mats, bad)
    assert "material_of_prim[12345]" in capfd.readouterr().err


def test_rebuild_bvh8_gives_the_same_hits_with_fewer_visits():
    """rodent_b200_scene_rebuild_bvh8: the split-BVH builder over the triangles of a scene that came with its own tree
    (the reference's Sponza block) -- same hits, `t` to rounding (the edges are recomputed), fewer nodes and packets visited."""
    from oracle import oracle
    from rodent_b200 import formats, render, testdata, workloads
    nodes, tris = formats.load_bvh(testdata.sponza_bvh8(), formats.BVH8_TRI4)
    scene = render.Scene.from_bvh8(nodes, tris, workloads.sponza_materials(), workloads.sponza_material_of_prim(tris))
    scene.rebuild_bvh8()
    mine = (scene.array("nodes").copy(), scene.array("tris").copy())
    assert len(mine[0]) != len(nodes)
    rays = formats.load_rays(testdata.rays("random"), 0.0, 1.0)[::16].copy()
    a, sa = oracle.traverse(*mine, rays, want_stats=True)
    b, sb = oracle.traverse(nodes, tris, rays, want_stats=True)
    assert ((a["tri_id"] >= 0) == (b["tri_id"] >= 0)).all()
    hit = b["tri_id"] >= 0
    assert np.abs(a["t"][hit] - b["t"][hit]).max() <= 1e-6 * np.abs(b["t"][hit]).max()
    assert sa.nodes < 0.8 * sb.nodes and sa.tri4 < 0.7 * sb.tri4
    scene.free()
