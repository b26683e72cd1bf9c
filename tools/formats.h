// Readers/writers for the reference's traversal file formats (host side, C++17).
//   .bvh  : u32 magic 0x95CBED1F, then blocks [u64 size][u32 type][u32 nodes][u32 tris][nodes][tris]
//           (reference reader: tools/common/load_bvh.h:21-74; `size` counts from the type field)
//   .rays : 6 x f32 per ray, org then dir (tools/common/load_rays.h:58-92)
//   .fbuf : f32 hit distance per ray (tools/bench_traversal/bench_traversal.cpp:350-354)
#pragma once

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../include/rodent_b200.h"

namespace rb200 {

enum BlockType : uint32_t { kBvh2Tri1 = 1, kBvh4Tri4 = 2, kBvh8Tri4 = 3 };
constexpr uint32_t kBvhMagic = 0x95CBED1Fu;

struct File {
    FILE* f = nullptr;
    File(const std::string& path, const char* mode) : f(std::fopen(path.c_str(), mode)) {}
    ~File() { if (f) std::fclose(f); }
    explicit operator bool() const { return f != nullptr; }
    bool read(void* p, size_t n) { return std::fread(p, 1, n, f) == n; }
    bool write(const void* p, size_t n) { return std::fwrite(p, 1, n, f) == n; }
};

// Finds the first block of `type` and reads its arrays.  NodeT/TriT must be the
// PODs of include/rodent_b200.h matching `type`.
template <typename NodeT, typename TriT>
bool read_bvh(const std::string& path, BlockType type, std::vector<NodeT>& nodes, std::vector<TriT>& tris) {
    File in(path, "rb");
    uint32_t magic = 0;
    if (!in || !in.read(&magic, 4) || magic != kBvhMagic) return false;
    for (;;) {
        uint64_t size = 0; uint32_t block = 0;
        if (!in.read(&size, 8) || !in.read(&block, 4)) return false;
        if (block == uint32_t(type)) break;
        if (std::fseek(in.f, long(size - 4), SEEK_CUR) != 0) return false;
    }
    uint32_t counts[2];
    if (!in.read(counts, 8)) return false;
    nodes.resize(counts[0]);
    tris.resize(counts[1]);
    return in.read(nodes.data(), sizeof(NodeT) * nodes.size()) && in.read(tris.data(), sizeof(TriT) * tris.size());
}

// Ray1 array with constant tmin/tmax, as RayTraits<Ray1>::write_ray does.
inline bool read_rays(const std::string& path, float tmin, float tmax, std::vector<Ray1>& rays) {
    File in(path, "rb");
    if (!in) return false;
    std::fseek(in.f, 0, SEEK_END);
    const long bytes = std::ftell(in.f);
    std::fseek(in.f, 0, SEEK_SET);
    if (bytes < 0 || bytes % 24 != 0) return false;
    std::vector<float> raw(size_t(bytes) / 4);
    if (!raw.empty() && !in.read(raw.data(), size_t(bytes))) return false;
    rays.resize(raw.size() / 6);
    for (size_t i = 0; i < rays.size(); i++) {
        const float* s = &raw[6 * i];
        rays[i] = Ray1{{s[0], s[1], s[2]}, tmin, {s[3], s[4], s[5]}, tmax};
    }
    return true;
}

inline bool write_fbuf(const std::string& path, const std::vector<Hit1>& hits) {
    File out(path, "wb");
    if (!out) return false;
    for (const Hit1& h : hits)
        if (!out.write(&h.t, 4)) return false;
    return true;
}

// 8-bit RGBA PNG through zlib only (libpng headers are not available in this image); the translation unit
// must link -lz and include <zlib.h> before this header to get it.
#ifdef ZLIB_H
inline bool write_png_rgba(const std::string& path, const std::vector<uint8_t>& rgba, uint32_t width, uint32_t height) {
    auto put32 = [](std::vector<uint8_t>& v, uint32_t x) { for (int s = 24; s >= 0; s -= 8) v.push_back(uint8_t(x >> s)); };
    File out(path, "wb");
    if (!out || rgba.size() != size_t(width) * height * 4) return false;
    auto chunk = [&](const char type[4], const std::vector<uint8_t>& data) {
        std::vector<uint8_t> buf;
        put32(buf, uint32_t(data.size()));
        buf.insert(buf.end(), type, type + 4);
        buf.insert(buf.end(), data.begin(), data.end());
        put32(buf, uint32_t(crc32(0, buf.data() + 4, uInt(buf.size() - 4))));
        return out.write(buf.data(), buf.size());
    };
    std::vector<uint8_t> raw;
    raw.reserve(size_t(height) * (size_t(width) * 4 + 1));
    for (uint32_t y = 0; y < height; y++) {
        raw.push_back(0);   // filter: none
        raw.insert(raw.end(), rgba.begin() + size_t(y) * width * 4, rgba.begin() + size_t(y + 1) * width * 4);
    }
    uLongf zlen = compressBound(uLong(raw.size()));
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), uLong(raw.size()), 6) != Z_OK) return false;
    z.resize(zlen);
    const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    if (!out.write(sig, 8)) return false;
    std::vector<uint8_t> ihdr;
    put32(ihdr, width); put32(ihdr, height);
    ihdr.insert(ihdr.end(), {8, 6, 0, 0, 0});   // 8 bit RGBA
    return chunk("IHDR", ihdr) && chunk("IDAT", z) && chunk("IEND", {});
}
#endif

}  // namespace rb200
