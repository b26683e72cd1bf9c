// PNG -> the RGBA8 image the reference's load_png produces (src/driver/image.cpp:25-93).
//
// The reference decodes with libpng and asks it for: palette and grey expanded to RGB (:63-68), 16 bit stripped to
// 8 (:71-72), tRNS turned into alpha, otherwise an opaque alpha byte appended (:75-80); it then stores the rows
// bottom-up (:84) and applies gamma 2.2 to r, g, b in place, truncating to a byte (:10-18).  libpng is not part of
// this image, zlib is: the container format (signature, chunks, IDAT inflate, the five row filters, Adam7) is
// decoded here directly from the PNG specification.  Alpha is carried along but no shader reads it
// (make_image_rgba32, src/render/image.impala:24-38).
#include "scene.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <new>

#include <zlib.h>

namespace rb200 {
namespace {

uint32_t be32(const uint8_t* p) { return uint32_t(p[0]) << 24 | uint32_t(p[1]) << 16 | uint32_t(p[2]) << 8 | uint32_t(p[3]); }

int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// Undoes the row filters of one (sub-)image in place; `data` holds rows of 1 + stride bytes.  Returns false on a bad filter id.
bool unfilter(uint8_t* data, size_t rows, size_t stride, size_t bpp) {
    std::vector<uint8_t> zero(stride, 0);
    const uint8_t* prev = zero.data();
    for (size_t y = 0; y < rows; y++) {
        uint8_t* row = data + y * (stride + 1);
        const int type = row[0];
        uint8_t* cur = row + 1;
        for (size_t i = 0; i < stride; i++) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int pred;
            switch (type) {
                case 0: pred = 0; break;
                case 1: pred = a; break;
                case 2: pred = b; break;
                case 3: pred = (a + b) >> 1; break;
                case 4: pred = paeth(a, b, c); break;
                default: return false;
            }
            cur[i] = uint8_t(cur[i] + pred);
        }
        prev = cur;
    }
    return true;
}

}  // namespace

static bool decode_png(const std::string& path, int& width, int& height, std::vector<uint32_t>& pixels, std::string& why);
bool load_png(const std::string& path, int& width, int& height, std::vector<uint32_t>& pixels, std::string& why) {
    try { return decode_png(path, width, height, pixels, why); }
    catch (const std::bad_alloc&) { why = "out of memory"; return false; }
}
static bool decode_png(const std::string& path, int& width, int& height, std::vector<uint32_t>& pixels, std::string& why) {
    std::ifstream in(path, std::ios::binary);
    if (!in) { why = "cannot open file"; return false; }
    const std::vector<uint8_t> file((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (file.size() < 8 || std::memcmp(file.data(), sig, 8) != 0) { why = "not a PNG file"; return false; }

    uint32_t w = 0, h = 0; int depth = 0, color = 0, interlace = 0;
    std::vector<uint8_t> idat, palette, trns;
    bool have_header = false, have_trns = false, ended = false;
    for (size_t pos = 8; pos + 12 <= file.size() && !ended;) {
        const uint32_t len = be32(&file[pos]);
        const char* type = reinterpret_cast<const char*>(&file[pos + 4]);
        if (size_t(len) + 12 > file.size() - pos) { why = "truncated chunk"; return false; }
        const uint8_t* body = &file[pos + 8];
        // critical chunks (upper-case first letter) must pass their CRC, as in libpng's default setting
        if (!(type[0] & 0x20) && be32(body + len) != uint32_t(crc32(0, &file[pos + 4], len + 4))) { why = "CRC error"; return false; }
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len != 13) { why = "bad IHDR"; return false; }
            w = be32(body); h = be32(body + 4); depth = body[8]; color = body[9]; interlace = body[12];
            if (body[10] != 0 || body[11] != 0 || interlace > 1) { why = "unknown compression / filter / interlace method"; return false; }
            have_header = true;
        } else if (!std::memcmp(type, "PLTE", 4)) palette.assign(body, body + len);
        else if (!std::memcmp(type, "tRNS", 4)) { trns.assign(body, body + len); have_trns = true; }
        else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!std::memcmp(type, "IEND", 4)) ended = true;
        pos += size_t(len) + 12;
    }
    if (!have_header || idat.empty()) { why = "missing IHDR or IDAT"; return false; }
    if (w == 0 || h == 0 || uint64_t(w) * h > (uint64_t(1) << 28)) { why = "unreasonable image size"; return false; }
    const int channels = color == 0 ? 1 : color == 2 ? 3 : color == 3 ? 1 : color == 4 ? 2 : color == 6 ? 4 : 0;
    const bool depth_ok = color == 0 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)
                        : color == 3 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8) : (depth == 8 || depth == 16);
    if (!channels || !depth_ok) { why = "invalid colour type / bit depth"; return false; }
    if (color == 3 && palette.size() < 3) { why = "paletted image without PLTE"; return false; }
    const size_t bits_pp = size_t(channels) * depth, bpp = (bits_pp + 7) / 8;

    // the (up to seven) passes: start and step in x and y
    struct Pass { int x0, y0, dx, dy; };
    static const Pass adam7[7] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
    static const Pass whole = {0, 0, 1, 1};
    const Pass* passes = interlace ? adam7 : &whole;
    const int num_passes = interlace ? 7 : 1;
    size_t raw_size = 0;
    for (int p = 0; p < num_passes; p++) {
        const size_t pw = (w - passes[p].x0 + passes[p].dx - 1) / passes[p].dx, ph = (h - passes[p].y0 + passes[p].dy - 1) / passes[p].dy;
        if (int(w) <= passes[p].x0 || int(h) <= passes[p].y0) continue;
        raw_size += ph * (1 + (pw * bits_pp + 7) / 8);
    }
    if (raw_size / 1100 > idat.size() + 16) { why = "image data too short for the stated size"; return false; }   // deflate expands at most ~1032 : 1
    std::vector<uint8_t> raw(raw_size);
    uLongf got = uLongf(raw_size);
    const int zr = uncompress(raw.data(), &got, idat.data(), uLong(idat.size()));
    if (zr != Z_OK || got != raw_size) { why = "corrupt image data (zlib " + std::to_string(zr) + ")"; return false; }

    width = int(w); height = int(h);
    pixels.assign(size_t(w) * h, 0);
    auto sample = [&](const uint8_t* row, size_t index) -> int {       // index-th sample of a row, at the file's bit depth
        if (depth == 8) return row[index];
        if (depth == 16) return row[2 * index];                          // png_set_strip_16: the high byte
        const size_t bit = index * depth;
        return (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
    };
    auto gamma = [](int v) { return uint32_t(uint8_t(std::pow(float(v) * (1.0f / 255.0f), 2.2f) * 255.0f)); };   // image.cpp:10-18
    uint32_t gamma_lut[256];
    for (int v = 0; v < 256; v++) gamma_lut[v] = gamma(v);

    size_t offset = 0;
    for (int p = 0; p < num_passes; p++) {
        if (int(w) <= passes[p].x0 || int(h) <= passes[p].y0) continue;
        const size_t pw = (w - passes[p].x0 + passes[p].dx - 1) / passes[p].dx, ph = (h - passes[p].y0 + passes[p].dy - 1) / passes[p].dy;
        const size_t stride = (pw * bits_pp + 7) / 8;
        if (!unfilter(raw.data() + offset, ph, stride, bpp)) { why = "unknown row filter"; return false; }
        for (size_t py = 0; py < ph; py++) {
            const uint8_t* row = raw.data() + offset + py * (stride + 1) + 1;
            const size_t y = passes[p].y0 + py * passes[p].dy;
            uint32_t* dst = pixels.data() + size_t(h - 1 - y) * w;       // bottom row first, image.cpp:84
            for (size_t px = 0; px < pw; px++) {
                const size_t x = passes[p].x0 + px * passes[p].dx;
                int r, g, b, a = 255;
                if (color == 3) {
                    const int idx = sample(row, px);
                    if (size_t(idx) * 3 + 2 >= palette.size()) { r = g = b = 0; }
                    else { r = palette[idx * 3]; g = palette[idx * 3 + 1]; b = palette[idx * 3 + 2]; }
                    if (have_trns && size_t(idx) < trns.size()) a = trns[idx];
                } else if (color == 0 || color == 4) {
                    const int v = sample(row, px * channels);
                    // grey below 8 bit is scaled to the full range when it is expanded
                    r = g = b = depth < 8 ? v * 255 / ((1 << depth) - 1) : v;
                    if (color == 4) a = sample(row, px * 2 + 1);
                    else if (have_trns && trns.size() >= 2) {
                        const int key = depth == 16 ? (trns[0] << 8 | trns[1]) : trns[1];
                        const int full = depth == 16 ? (row[2 * px] << 8 | row[2 * px + 1]) : v;
                        if (full == key) a = 0;
                    }
                } else {
                    r = sample(row, px * channels); g = sample(row, px * channels + 1); b = sample(row, px * channels + 2);
                    if (color == 6) a = sample(row, px * 4 + 3);
                    else if (have_trns && trns.size() >= 6) {
                        auto full = [&](int c) { return depth == 16 ? (row[2 * (px * 3 + c)] << 8 | row[2 * (px * 3 + c) + 1]) : sample(row, px * 3 + c); };
                        auto key = [&](int c) { return depth == 16 ? (trns[2 * c] << 8 | trns[2 * c + 1]) : trns[2 * c + 1]; };
                        if (full(0) == key(0) && full(1) == key(1) && full(2) == key(2)) a = 0;
                    }
                }
                dst[x] = gamma_lut[r] | gamma_lut[g] << 8 | gamma_lut[b] << 16 | uint32_t(a) << 24;
            }
        }
        offset += ph * (stride + 1);
    }
    return true;
}

}  // namespace rb200
