"""A small PNG encoder for the tests: every colour type, bit depth, row filter and Adam7 interlacing, so the loader
(rodent_b200/csrc/image.cpp) is exercised on more than what PIL chooses to write."""
import struct
import zlib

import numpy as np

CHANNELS = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}
ADAM7 = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]


def _chunk(kind: bytes, body: bytes) -> bytes:
    return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body) & 0xFFFFFFFF)


def _pack_rows(samples: np.ndarray, depth: int) -> list[bytes]:
    """samples: (rows, width * channels) integers at `depth` bits -> one bytes object per row."""
    rows = []
    for r in samples:
        if depth == 8:
            rows.append(r.astype(np.uint8).tobytes())
        elif depth == 16:
            rows.append(r.astype(">u2").tobytes())
        else:
            bits = np.zeros(((len(r) * depth + 7) // 8) * 8, np.uint8)
            for k in range(depth):                                     # most significant bit first
                bits[np.arange(len(r)) * depth + k] = (r >> (depth - 1 - k)) & 1
            rows.append(np.packbits(bits).tobytes())
    return rows


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if pa <= pb and pa <= pc else (b if pb <= pc else c)


def _filter(rows: list[bytes], bpp: int, first_filter: int) -> bytes:
    out = bytearray()
    prev = bytes(len(rows[0])) if rows else b""
    for y, cur in enumerate(rows):
        f = (first_filter + y) % 5
        line = bytearray(len(cur))
        for i, x in enumerate(cur):
            a = cur[i - bpp] if i >= bpp else 0
            b = prev[i]
            c = prev[i - bpp] if i >= bpp else 0
            pred = (0, a, b, (a + b) >> 1, _paeth(a, b, c))[f]
            line[i] = (x - pred) & 0xFF
        out.append(f)
        out += line
        prev = cur
    return bytes(out)


def write_png(path, samples: np.ndarray, color: int, depth: int, *, interlace: bool = False, first_filter: int = 0,
              palette: np.ndarray | None = None, trns: bytes | None = None, idat_pieces: int = 1) -> None:
    """samples: (height, width, channels) integers (palette indices for colour type 3), top row first."""
    h, w, ch = samples.shape
    assert ch == CHANNELS[color]
    bpp = max(1, ch * depth // 8)
    raw = b""
    for x0, y0, dx, dy in (ADAM7 if interlace else [(0, 0, 1, 1)]):
        sub = samples[y0::dy, x0::dx]
        if sub.shape[0] == 0 or sub.shape[1] == 0:
            continue
        raw += _filter(_pack_rows(sub.reshape(sub.shape[0], -1), depth), bpp, first_filter)
    data = zlib.compress(raw, 6)
    out = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, color, 0, 0, int(interlace)))
    if palette is not None:
        out += _chunk(b"PLTE", np.asarray(palette, np.uint8).tobytes())
    if trns is not None:
        out += _chunk(b"tRNS", trns)
    out += _chunk(b"tEXt", b"Comment\0written by tests/png_writer.py")          # an ancillary chunk the loader must skip
    step = (len(data) + idat_pieces - 1) // idat_pieces
    for k in range(0, len(data), step):
        out += _chunk(b"IDAT", data[k:k + step])
    out += _chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(out)


def gamma_lut() -> np.ndarray:
    """gamma_correct, src/driver/image.cpp:10-18: pow(v / 255, 2.2) * 255 in float, truncated to a byte."""
    v = np.arange(256, dtype=np.float32) * np.float32(1.0 / 255.0)
    return (np.power(v, np.float32(2.2), dtype=np.float32) * np.float32(255.0)).astype(np.uint8)


def expected_pixels(rgba8: np.ndarray) -> np.ndarray:
    """(height, width, 4) uint8, top row first -> the (height, width) uint32 image load_png produces (bottom row first)."""
    lut = gamma_lut().astype(np.uint32)
    px = rgba8[::-1].astype(np.uint32)
    return lut[px[..., 0]] | lut[px[..., 1]] << 8 | lut[px[..., 2]] << 16 | px[..., 3] << 24
