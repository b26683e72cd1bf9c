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
    rb200::File in(files[0], "rb");
    if (!in) { std::cerr << "Cannot open " << files[0] << std::endl; return 1; }
    std::vector<float> image(size_t(width) * size_t(height));
    if (!in.read(image.data(), image.size() * 4)) { std::cerr << "Not enough data in the float buffer" << std::endl; return 1; }
    const float tmax = normalize ? *std::max_element(image.begin(), image.end()) : 1.0f;

    std::vector<uint8_t> rgba;
    rgba.reserve(image.size() * 4);
    for (float t : image) {
        const uint8_t c = uint8_t(255.0f * t / tmax);
        rgba.insert(rgba.end(), {c, c, c, uint8_t(255)});
    }
    if (!rb200::write_png_rgba(files[1], rgba, uint32_t(width), uint32_t(height))) { std::cerr << "Cannot write " << files[1] << std::endl; return 1; }
    return 0;
}
