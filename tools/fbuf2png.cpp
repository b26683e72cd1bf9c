// fbuf2png -- converts a .fbuf (raw f32 per pixel) into an 8-bit RGBA PNG exactly as the
// reference's tools/fbuf2png/fbuf2png.cpp:76-117 does (c = uint8(255.0f * t / max), alpha 255),
// written with zlib only (libpng headers are not available in this image).
//   fbuf2png [-n] [-w width] [-h height] input.fbuf output.png     (defaults 1024 x 1024)
#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "formats.h"

namespace {
void put32(std::vector<uint8_t>& v, uint32_t x) { for (int s = 24; s >= 0; s -= 8) v.push_back(uint8_t(x >> s)); }
void chunk(rb200::File& out, const char type[4], const std::vector<uint8_t>& data) {
    std::vector<uint8_t> buf;
    put32(buf, uint32_t(data.size()));
    buf.insert(buf.end(), type, type + 4);
    buf.insert(buf.end(), data.begin(), data.end());
    put32(buf, uint32_t(crc32(0, buf.data() + 4, uInt(buf.size() - 4))));
    out.write(buf.data(), buf.size());
}
}  // namespace

int main(int argc, char** argv) {
    bool normalize = false;
    long width = 1024, height = 1024;
    std::vector<std::string> files;
    for (int i = 1; i < argc; i++) {
        if (!std::strcmp(argv[i], "-n")) normalize = true;
        else if (!std::strcmp(argv[i], "-w") && i + 1 < argc) width = std::strtol(argv[++i], nullptr, 10);
        else if (!std::strcmp(argv[i], "-h") && i + 1 < argc) height = std::strtol(argv[++i], nullptr, 10);
        else if (!std::strcmp(argv[i], "--help")) { std::cout << "fbuf2png [-n] [-w width] [-h height] input.fbuf output.png\n"; return 0; }
        else files.push_back(argv[i]);
    }
    if (files.size() < 2) { std::cerr << "Missing input or output file" << std::endl; return 1; }
    if (files.size() > 2) { std::cerr << "Too many arguments" << std::endl; return 1; }
    rb200::File in(files[0], "rb"), out(files[1], "wb");
    if (!in || !out) return 1;
    std::vector<float> image(size_t(width) * size_t(height));
    if (!in.read(image.data(), image.size() * 4)) { std::cerr << "Not enough data in the float buffer" << std::endl; return 1; }
    const float tmax = normalize ? *std::max_element(image.begin(), image.end()) : 1.0f;

    std::vector<uint8_t> raw;
    raw.reserve(size_t(height) * (size_t(width) * 4 + 1));
    for (long y = 0; y < height; y++) {
        raw.push_back(0);   // filter: none
        for (long x = 0; x < width; x++) {
            const uint8_t c = uint8_t(255.0f * image[size_t(y) * width + x] / tmax);
            raw.insert(raw.end(), {c, c, c, uint8_t(255)});
        }
    }
    uLongf zlen = compressBound(uLong(raw.size()));
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), uLong(raw.size()), 6) != Z_OK) return 1;
    z.resize(zlen);

    const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    out.write(sig, 8);
    std::vector<uint8_t> ihdr;
    put32(ihdr, uint32_t(width)); put32(ihdr, uint32_t(height));
    ihdr.insert(ihdr.end(), {8, 6, 0, 0, 0});   // 8 bit RGBA
    chunk(out, "IHDR", ihdr);
    chunk(out, "IDAT", z);
    chunk(out, "IEND", {});
    return 0;
}
