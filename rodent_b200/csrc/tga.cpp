// TGA -> the RGBA8 image layout of the other loaders (bottom row first, gamma 2.2 on r, g, b as gamma_correct does,
// src/driver/image.cpp:10-18).  The reference's converter emits `device.load_tga(...)` for .tga / .tiff images
// (src/driver/converter.cpp:759-762), but no device defines it (src/render/mapping_*.impala have load_png / load_jpg only), so
// such scenes do not build there; here they load.  Truevision TGA 2.0: types 1 / 2 / 3 (colour-mapped, true colour, grey)
// and their run-length forms 9 / 10 / 11; 8, 15 / 16, 24 and 32 bits; either row and column order.
#include "scene.h"

#include <cmath>
#include <cstring>
#include <fstream>
#include <iterator>
#include <new>

namespace rb200 {

static bool decode_tga(const std::string& path, int& width, int& height, std::vector<uint32_t>& pixels, std::string& why) {
    std::ifstream in(path, std::ios::binary);
    if (!in) { why = "cannot open file"; return false; }
    const std::vector<uint8_t> f((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    if (f.size() < 18) { why = "not a TGA file"; return false; }
    const int id_len = f[0], map_type = f[1], type = f[2];
    const int map_first = f[3] | f[4] << 8, map_len = f[5] | f[6] << 8, map_bits = f[7];
    const int w = f[12] | f[13] << 8, h = f[14] | f[15] << 8, bits = f[16], desc = f[17];
    const bool rle = type >= 9 && type <= 11;
    const int kind = rle ? type - 8 : type;                                  // 1 colour-mapped, 2 true colour, 3 grey
    if (kind < 1 || kind > 3 || map_type > 1 || w == 0 || h == 0) { why = "unsupported TGA image type"; return false; }
    if ((kind == 1 && (bits != 8 || !map_type)) || (kind == 3 && bits != 8) ||
        (kind == 2 && bits != 15 && bits != 16 && bits != 24 && bits != 32)) { why = "unsupported TGA pixel depth"; return false; }
    if (map_type && map_bits != 15 && map_bits != 16 && map_bits != 24 && map_bits != 32) { why = "unsupported TGA colour map"; return false; }
    const size_t map_bytes = map_type ? size_t(map_len) * ((map_bits + 7) / 8) : 0;
    size_t pos = 18 + size_t(id_len);
    if (pos + map_bytes > f.size()) { why = "truncated TGA file"; return false; }
    const uint8_t* map = f.data() + pos;
    pos += map_bytes;

    auto colour = [](const uint8_t* p, int nbits, uint8_t out[4]) {          // little-endian B, G, R(, A) or 5-5-5
        if (nbits == 24 || nbits == 32) { out[0] = p[2]; out[1] = p[1]; out[2] = p[0]; out[3] = nbits == 32 ? p[3] : 255; }
        else {
            const int v = p[0] | p[1] << 8;
            const int r = (v >> 10) & 31, g = (v >> 5) & 31, b = v & 31;
            out[0] = uint8_t(r << 3 | r >> 2); out[1] = uint8_t(g << 3 | g >> 2); out[2] = uint8_t(b << 3 | b >> 2); out[3] = 255;
        }
    };
    uint32_t gamma_lut[256];
    for (int v = 0; v < 256; v++) gamma_lut[v] = uint32_t(uint8_t(std::pow(float(v) * (1.0f / 255.0f), 2.2f) * 255.0f));
    const size_t bpp = size_t(bits + 7) / 8, count = size_t(w) * h;
    // what is left of the file must be able to hold that many pixels (a run-length packet yields at most 128 from 1 + bpp bytes)
    if (rle ? count > (f.size() - pos) * 128 : pos + count * bpp > f.size()) { why = "truncated TGA file"; return false; }
    width = w; height = h;
    pixels.assign(count, 0);
    const bool top_first = (desc & 0x20) != 0, right_first = (desc & 0x10) != 0;
    size_t k = 0;                                                            // pixel number in file order
    auto put = [&](const uint8_t* p) -> bool {
        uint8_t c[4];
        if (kind == 3) { c[0] = c[1] = c[2] = p[0]; c[3] = 255; }
        else if (kind == 1) {
            const int idx = int(p[0]) - map_first;
            if (idx < 0 || idx >= map_len) return false;
            colour(map + size_t(idx) * ((map_bits + 7) / 8), map_bits, c);
        } else colour(p, bits, c);
        const size_t y = k / w, x = k % w;
        const size_t row = top_first ? size_t(h) - 1 - y : y, col = right_first ? size_t(w) - 1 - x : x;
        pixels[row * w + col] = gamma_lut[c[0]] | gamma_lut[c[1]] << 8 | gamma_lut[c[2]] << 16 | uint32_t(c[3]) << 24;
        k++;
        return true;
    };
    if (!rle) {
        if (pos + count * bpp > f.size()) { why = "truncated TGA file"; return false; }
        for (size_t i = 0; i < count; i++)
            if (!put(f.data() + pos + i * bpp)) { why = "colour index out of range"; return false; }
        return true;
    }
    while (k < count) {
        if (pos >= f.size()) { why = "truncated TGA file"; return false; }
        const int head = f[pos++], n = (head & 127) + 1;
        if (size_t(n) > count - k) { why = "corrupt TGA run"; return false; }
        if (head & 128) {
            if (pos + bpp > f.size()) { why = "truncated TGA file"; return false; }
            for (int i = 0; i < n; i++)
                if (!put(f.data() + pos)) { why = "colour index out of range"; return false; }
            pos += bpp;
        } else {
            if (pos + size_t(n) * bpp > f.size()) { why = "truncated TGA file"; return false; }
            for (int i = 0; i < n; i++)
                if (!put(f.data() + pos + size_t(i) * bpp)) { why = "colour index out of range"; return false; }
            pos += size_t(n) * bpp;
        }
    }
    return true;
}

bool load_tga(const std::string& path, int& width, int& height, std::vector<uint32_t>& pixels, std::string& why) {
    try { return decode_tga(path, width, height, pixels, why); }
    catch (const std::bad_alloc&) { why = "out of memory"; return false; }
}

}  // namespace rb200
