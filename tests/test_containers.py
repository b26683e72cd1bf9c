"""The reference's data containers: LZ4 blocks, `data/*.bin` buffers (src/driver/buffer.h) and `data/bvh.bin`
(converter.cpp:428-438, interface.cpp:432-454).  liblz4 is not in this image; the codec is checked against blocks written by
hand from the format description, against itself, and for refusing malformed input."""
import struct

import numpy as np
import pytest

from rodent_b200 import formats, testdata


def test_lz4_known_blocks():
    """Hand-assembled blocks: literals only; a match that overlaps its own output (run-length); extended lengths."""
    assert formats.lz4_decompress(bytes([0x50]) + b"hello", 5) == b"hello"
    # "abcd" literals, then a match of length 12 at offset 4 (copies what it writes), then the 5 literal bytes "vwxyz"
    block = bytes([0x48]) + b"abcd" + struct.pack("<H", 4) + bytes([0x50]) + b"vwxyz"
    assert formats.lz4_decompress(block, 21) == b"abcd" * 4 + b"vwxyz"
    # 1 literal 'x', match offset 1 length 4 + 15 + 255 + 3 = 277, then 5 literals
    block = bytes([0x1F]) + b"x" + struct.pack("<H", 1) + bytes([255, 3]) + bytes([0x50]) + b"12345"
    assert formats.lz4_decompress(block, 1 + 277 + 5) == b"x" * 278 + b"12345"
    # 300 literals: 15 + 255 + 30
    lit = bytes(range(256)) + bytes(44)
    assert formats.lz4_decompress(bytes([0xF0, 255, 30]) + lit, 300) == lit
    assert formats.lz4_decompress(bytes([0x00]), 0) == b""


@pytest.mark.parametrize("bad", [b"", bytes([0x50]) + b"hell", bytes([0x40]) + b"abcd" + b"\x00\x00" + bytes([0x50]) + b"vwxyz",
                                 bytes([0x44]) + b"abcd" + struct.pack("<H", 9) + bytes([0x50]) + b"vwxyz",
                                 bytes([0x4F]) + b"abcd" + struct.pack("<H", 4) + bytes([255]), bytes([0xF0, 255])])
def test_lz4_rejects_malformed_blocks(bad):
    with pytest.raises(ValueError):
        formats.lz4_decompress(bad, 21)


def test_lz4_round_trips_and_compresses():
    rng = np.random.default_rng(0)
    cases = [b"", b"a", b"abc" * 3, bytes(12), bytes(13), bytes(100000), rng.integers(0, 256, 70000, dtype=np.uint8).tobytes(),
             (b"the quick brown fox " * 5000)[:77777], np.arange(50000, dtype=np.float32).tobytes(),
             rng.integers(0, 4, 200000, dtype=np.uint8).tobytes()]
    for raw in cases:
        block = formats.lz4_compress(raw)
        assert formats.lz4_decompress(block, len(raw)) == raw
        if len(raw) > 12:
            # the end-of-block rules every LZ4 decoder relies on: the last sequence is literals only and at least 5 bytes long
            assert len(block) >= 6
    assert len(formats.lz4_compress(bytes(100000))) < 500
    assert len(formats.lz4_compress((b"the quick brown fox " * 5000)[:77777])) < 1000
    assert len(formats.lz4_compress(cases[6])) <= len(cases[6]) + len(cases[6]) // 255 + 16         # LZ4_compressBound
    with pytest.raises(ValueError):
        formats.lz4_decompress(formats.lz4_compress(cases[7]), len(cases[7]) - 1)                   # destination too small


def test_buffer_files(tmp_path):
    """[u32 raw size][u32 compressed size][block], buffer.h:46-55."""
    a = np.arange(10000, dtype=np.float32).reshape(-1, 4)
    formats.write_buffer(tmp_path / "vertices.bin", a)
    blob = (tmp_path / "vertices.bin").read_bytes()
    raw, comp = struct.unpack_from("<II", blob)
    assert raw == a.nbytes and comp == len(blob) - 8
    assert formats.lz4_decompress(blob[8:], raw) == a.tobytes()
    assert np.array_equal(formats.load_buffer(tmp_path / "vertices.bin", np.float32).reshape(-1, 4), a)
    formats.write_buffer(tmp_path / "empty.bin", np.zeros(0, np.int32))
    assert len(formats.load_buffer(tmp_path / "empty.bin")) == 0
    (tmp_path / "cut.bin").write_bytes(blob[:len(blob) // 2])
    (tmp_path / "lying.bin").write_bytes(struct.pack("<II", raw + 4, comp) + blob[8:])
    for name in ("cut.bin", "lying.bin", "missing.bin"):
        with pytest.raises(ValueError):
            formats.load_buffer(tmp_path / name)


def test_bvh_bin_with_several_layouts(tmp_path, sponza):
    """data/bvh.bin holds one entry per layout the converter was asked for; the loader takes the one whose record sizes match
    and skips the others (interface.cpp:438-451)."""
    nodes8, tris4 = sponza
    nodes8, tris4 = nodes8[:2000], tris4[:5000]
    nodes2, tris1 = formats.load_bvh(testdata.sponza_bvh2(), formats.BVH2_TRI1)
    nodes2, tris1 = nodes2[:3000], tris1[:7000]
    path = tmp_path / "bvh.bin"
    formats.append_bvh_bin(path, nodes2, tris1)
    formats.append_bvh_bin(path, nodes8, tris4)
    assert struct.unpack_from("<II", path.read_bytes()) == (64, 48)
    n, t = formats.load_bvh_bin(path, formats.BVH8_TRI4)
    assert n.tobytes() == nodes8.tobytes() and t.tobytes() == tris4.tobytes()
    n, t = formats.load_bvh_bin(path, formats.BVH2_TRI1)
    assert n.tobytes() == nodes2.tobytes() and t.tobytes() == tris1.tobytes()
    with pytest.raises(ValueError):
        formats.load_bvh_bin(path, formats.BVH4_TRI4)
    assert path.stat().st_size < 0.8 * (nodes8.nbytes + tris4.nbytes + nodes2.nbytes + tris1.nbytes)     # it does compress


def test_lz4_round_trip_property():
    """Any byte string survives compress -> decompress; a flipped bit in the block is either rejected or decodes to exactly
    raw_size bytes (never past the destination)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=300, deadline=None)
    @given(st.binary(max_size=4000), st.integers(0, 7), st.integers(0, 1 << 30))
    def check(raw, bit, where):
        # make part of the input repetitive so that matches occur
        raw = raw + raw[: len(raw) // 2] * 3
        block = formats.lz4_compress(raw)
        assert formats.lz4_decompress(block, len(raw)) == raw
        damaged = bytearray(block)
        damaged[where % len(damaged)] ^= 1 << bit
        try:
            out = formats.lz4_decompress(bytes(damaged), len(raw))
            assert len(out) == len(raw)
        except ValueError:
            pass
    check()
