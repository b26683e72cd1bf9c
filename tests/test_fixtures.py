"""The benchmark inputs are regenerated, not stored: check they are the reference's."""
import hashlib
import subprocess

import numpy as np
import pytest

from conftest import REFERENCE
from rodent_b200 import formats, testdata

SHA256 = {
    "sponza-primary.rays": "960faf3389dbaccdfcbc4f835da91d5df59c2a3c38388ec56087f2c0414d5c2f",
    "sponza-random.rays": "13aac2626524e118f75302f0a3916a331dc94839e9bdfd36e509d5e5c83064e5",
    "sponza_bvh8.bvh": "c2a74143e777ebfdc9aaadb74052bb87cccc8239420c2b0dac0b5df7f6d6c646",
}


def _sha(path):
    return hashlib.sha256(path.read_bytes()).hexdigest()


def test_generated_inputs_are_stable():
    assert _sha(testdata.rays("primary")) == SHA256["sponza-primary.rays"]
    assert _sha(testdata.rays("random")) == SHA256["sponza-random.rays"]
    assert _sha(testdata.sponza_bvh8()) == SHA256["sponza_bvh8.bvh"]


@pytest.mark.skipif(not (REFERENCE / "testing/sponza.bvh").exists(), reason="reference tree not present")
def test_generated_inputs_equal_reference_files():
    for name in ("primary", "random"):
        assert testdata.rays(name).read_bytes() == (REFERENCE / f"testing/sponza-{name}.rays").read_bytes()
    n8, t8 = formats.load_bvh(REFERENCE / "testing/sponza.bvh", formats.BVH8_TRI4)
    m8, u8 = formats.load_bvh(testdata.sponza_bvh8(), formats.BVH8_TRI4)
    assert n8.tobytes() == m8.tobytes() and t8.tobytes() == u8.tobytes()


def test_ray_gen_cli_errors(tmp_path):
    exe = str(testdata.ray_gen_bin())
    assert subprocess.run([exe], capture_output=True).returncode == 1
    assert subprocess.run([exe, "bogus"], capture_output=True).returncode == 1
    assert subprocess.run([exe, "primary", "1"], capture_output=True).returncode == 1
    assert subprocess.run([exe, "--help"], capture_output=True).returncode == 0


def test_shadow_mode(tmp_path):
    """shadow: rays from the light to the primary hit points (tools/ray_gen/ray_gen.cpp:60-85)."""
    exe = str(testdata.ray_gen_bin())
    org_dir = np.array([[0, 0, 0, 1, 0, 0], [1, 2, 3, 0, 2, 0]], "<f4")
    org_dir.tofile(tmp_path / "p.rays")
    np.array([2.0, 0.5], "<f4").tofile(tmp_path / "p.fbuf")
    r = subprocess.run([exe, "shadow", "0", "10", "0", str(tmp_path / "p.rays"), str(tmp_path / "p.fbuf"), "2", "1",
                        str(tmp_path / "s.rays")], capture_output=True)
    assert r.returncode == 0
    out = np.fromfile(tmp_path / "s.rays", "<f4").reshape(-1, 6)
    assert np.array_equal(out, np.array([[0, 10, 0, 2, -10, 0], [0, 10, 0, 1, -7, 3]], "<f4"))


def test_bvh_container_round_trip(tmp_path, sponza):
    nodes, tris = sponza
    formats.save_bvh(tmp_path / "x.bvh", [(formats.BVH8_TRI4, nodes[:10], tris[:7])])
    n, t = formats.load_bvh(tmp_path / "x.bvh", formats.BVH8_TRI4)
    assert n.tobytes() == nodes[:10].tobytes() and t.tobytes() == tris[:7].tobytes()
    with pytest.raises(ValueError):
        formats.load_bvh(tmp_path / "x.bvh", formats.BVH4_TRI4)
    (tmp_path / "bad.bvh").write_bytes(b"\0" * 64)
    with pytest.raises(ValueError):
        formats.load_bvh(tmp_path / "bad.bvh")


def test_fbuf_to_gray_matches_fbuf2png():
    t = np.array([0.0, 1.0, 2.5, 5000.0], np.float32)
    assert formats.fbuf_to_gray(t).tolist() == [0, 0, 0, 255]
    assert formats.fbuf_to_gray(np.array([0.2, 0.999, 1.0], np.float32), normalize=False).tolist() == [51, 254, 255]
