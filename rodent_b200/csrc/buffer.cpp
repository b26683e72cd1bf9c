// The reference's `data/*.bin` container (src/driver/buffer.h:10-61) and `data/bvh.bin` (src/driver/converter.cpp:428-438
// writes it, src/driver/interface.cpp:432-454 reads it):
//   buffer   = [u32 raw size][u32 compressed size][LZ4 block]
//   bvh.bin  = repeated { [u32 sizeof(Node)][u32 sizeof(Tri)] buffer(nodes) buffer(tris) }, one entry per layout
// The reference links liblz4, which is not part of this image; the LZ4 *block* format (what LZ4_compress_default writes
// and LZ4_decompress_safe reads) is small and public, so both directions are written here from its description:
//   sequence = token (literal length : 4 | match length - 4 : 4), [length bytes 255...], literals, u16 offset, [length bytes]
//   the last sequence has literals only; the last 5 bytes of a block are literals, the last match starts at least 12 bytes
//   before the end of the block.
#include "scene.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

namespace rb200 {
namespace {

// LZ4_decompress_safe: returns the number of bytes written, or -1 on malformed input / insufficient room.
int64_t lz4_decompress(const uint8_t* src, int64_t n, uint8_t* dst, int64_t cap) {
    int64_t ip = 0, op = 0;
    if (n == 0) return cap == 0 ? 0 : -1;
    while (ip < n) {
        const int token = src[ip++];
        int64_t lit = token >> 4;
        if (lit == 15) { int b; do { if (ip >= n) return -1; b = src[ip++]; lit += b; } while (b == 255); }
        if (lit > n - ip || lit > cap - op) return -1;
        if (lit) std::memcpy(dst + op, src + ip, size_t(lit));
        ip += lit; op += lit;
        if (ip == n) return op;                                              // the last sequence stops after its literals
        if (n - ip < 2) return -1;
        const int64_t offset = src[ip] | src[ip + 1] << 8;
        ip += 2;
        if (offset == 0 || offset > op) return -1;
        int64_t len = (token & 15) + 4;
        if ((token & 15) == 15) { int b; do { if (ip >= n) return -1; b = src[ip++]; len += b; } while (b == 255); }
        if (len > cap - op) return -1;
        for (int64_t k = 0; k < len; k++, op++) dst[op] = dst[op - offset];  // byte by byte: matches may overlap their own output
    }
    return -1;                                                               // ran out of input inside a sequence
}

int64_t lz4_bound(int64_t n) { return n + n / 255 + 16; }                    // LZ4_compressBound

// A greedy single-probe hash-table compressor producing a valid block (any LZ4 decoder reads it).
int64_t lz4_compress(const uint8_t* src, int64_t n, uint8_t* dst, int64_t cap) {
    if (cap < lz4_bound(n)) return -1;
    constexpr int kHashBits = 16;
    std::vector<int64_t> table(size_t(1) << kHashBits, -1);
    auto read32 = [&](int64_t i) { uint32_t v; std::memcpy(&v, src + i, 4); return v; };
    auto hash = [&](uint32_t v) { return (v * 2654435761u) >> (32 - kHashBits); };
    int64_t op = 0, anchor = 0, ip = 0;
    auto emit = [&](int64_t lit_len, int64_t match_len, int64_t offset) {    // match_len 0: the final, literal-only sequence
        const int64_t ml = match_len ? match_len - 4 : 0;
        dst[op++] = uint8_t((lit_len >= 15 ? 15 : lit_len) << 4 | (match_len ? (ml >= 15 ? 15 : ml) : 0));
        if (lit_len >= 15) { int64_t r = lit_len - 15; for (; r >= 255; r -= 255) dst[op++] = 255; dst[op++] = uint8_t(r); }
        if (lit_len) std::memcpy(dst + op, src + anchor, size_t(lit_len));
        op += lit_len;
        if (match_len) {
            dst[op++] = uint8_t(offset & 255); dst[op++] = uint8_t(offset >> 8);
            if (ml >= 15) { int64_t r = ml - 15; for (; r >= 255; r -= 255) dst[op++] = 255; dst[op++] = uint8_t(r); }
        }
    };
    const int64_t match_limit = n - 12;                                      // no match may start after this, none may reach the last 5 bytes
    while (ip < match_limit) {
        const uint32_t h = hash(read32(ip));
        const int64_t cand = table[h];
        table[h] = ip;
        if (cand >= 0 && ip - cand <= 65535 && read32(cand) == read32(ip)) {
            int64_t len = 4;
            while (ip + len < n - 5 && src[cand + len] == src[ip + len]) len++;
            emit(ip - anchor, len, ip - cand);
            ip += len; anchor = ip;
        } else ip++;
    }
    emit(n - anchor, 0, 0);
    return op;
}

bool read_u32(std::FILE* f, uint32_t& v) { return std::fread(&v, 4, 1, f) == 1; }

// read_buffer, buffer.h:22-31.  Returns malloc'ed bytes (size in `raw`), or nullptr.
uint8_t* read_one_buffer(std::FILE* f, int64_t& raw) {
    uint32_t in_size = 0, out_size = 0;
    if (!read_u32(f, in_size) || !read_u32(f, out_size)) return nullptr;
    std::vector<uint8_t> comp(out_size);
    if (out_size && std::fread(comp.data(), 1, out_size, f) != out_size) return nullptr;
    uint8_t* out = static_cast<uint8_t*>(std::malloc(in_size ? in_size : 1));
    if (!out) return nullptr;
    if (lz4_decompress(comp.data(), out_size, out, in_size) != int64_t(in_size)) { std::free(out); return nullptr; }
    raw = in_size;
    return out;
}
bool skip_one_buffer(std::FILE* f) {                                         // skip_buffer, buffer.h:10-15
    uint32_t in_size = 0, out_size = 0;
    return read_u32(f, in_size) && read_u32(f, out_size) && std::fseek(f, long(out_size), SEEK_CUR) == 0;
}
bool write_one_buffer(std::FILE* f, const void* data, int64_t bytes) {       // write_buffer, buffer.h:46-55
    if (bytes < 0 || bytes > 0x7FFFFFFF) return false;
    std::vector<uint8_t> comp(size_t(lz4_bound(bytes)));
    const int64_t c = lz4_compress(static_cast<const uint8_t*>(data), bytes, comp.data(), int64_t(comp.size()));
    if (c < 0) return false;
    const uint32_t in_size = uint32_t(bytes), out_size = uint32_t(c);
    return std::fwrite(&in_size, 4, 1, f) == 1 && std::fwrite(&out_size, 4, 1, f) == 1 && (c == 0 || std::fwrite(comp.data(), 1, size_t(c), f) == size_t(c));
}

}  // namespace
}  // namespace rb200

using namespace rb200;

extern "C" {

int64_t rodent_b200_lz4_decompress(const void* src, int64_t src_size, void* dst, int64_t dst_capacity) {
    if (src_size < 0 || dst_capacity < 0) return -1;
    return lz4_decompress(static_cast<const uint8_t*>(src), src_size, static_cast<uint8_t*>(dst), dst_capacity);
}
int64_t rodent_b200_lz4_compress_bound(int64_t n) { return lz4_bound(n); }
int64_t rodent_b200_lz4_compress(const void* src, int64_t n, void* dst, int64_t dst_capacity) {
    if (n < 0) return -1;
    try { return lz4_compress(static_cast<const uint8_t*>(src), n, static_cast<uint8_t*>(dst), dst_capacity); }
    catch (const std::bad_alloc&) { return -1; }
}

void* rodent_b200_load_buffer(const char* file, int64_t* size) {
    std::FILE* f = std::fopen(file, "rb");
    if (!f) { std::fprintf(stderr, "rodent_b200: cannot open buffer '%s'\n", file); return nullptr; }
    int64_t raw = 0;
    uint8_t* out = nullptr;
    try { out = read_one_buffer(f, raw); } catch (const std::bad_alloc&) { out = nullptr; }
    std::fclose(f);
    if (!out) { std::fprintf(stderr, "rodent_b200: invalid buffer file '%s'\n", file); return nullptr; }
    if (size) *size = raw;
    return out;
}
void rodent_b200_free_buffer(void* p) { std::free(p); }
int32_t rodent_b200_write_buffer(const char* file, const void* data, int64_t bytes) {
    std::FILE* f = std::fopen(file, "wb");
    if (!f) return 0;
    bool ok = false;
    try { ok = write_one_buffer(f, data, bytes); } catch (const std::bad_alloc&) { ok = false; }
    return (std::fclose(f) == 0 && ok) ? 1 : 0;
}

int32_t rodent_b200_load_bvh_bin(const char* file, int32_t node_size, int32_t tri_size,
                                 void** nodes, int64_t* num_nodes, void** tris, int64_t* num_tris) {
    std::FILE* f = std::fopen(file, "rb");
    if (!f) { std::fprintf(stderr, "rodent_b200: cannot open BVH '%s'\n", file); return 0; }
    int32_t found = 0;
    try {
        for (;;) {
            uint32_t ns = 0, ts = 0;
            if (!read_u32(f, ns) || !read_u32(f, ts)) break;
            if (int32_t(ns) == node_size && int32_t(ts) == tri_size) {
                int64_t nb = 0, tb = 0;
                uint8_t* n = read_one_buffer(f, nb);
                uint8_t* t = n ? read_one_buffer(f, tb) : nullptr;
                if (!n || !t || nb % node_size || tb % tri_size) { std::free(n); std::free(t); break; }
                *nodes = n; *num_nodes = nb / node_size; *tris = t; *num_tris = tb / tri_size;
                found = 1;
                break;
            }
            if (!skip_one_buffer(f) || !skip_one_buffer(f)) break;
        }
    } catch (const std::bad_alloc&) { found = 0; }
    std::fclose(f);
    if (!found) std::fprintf(stderr, "rodent_b200: invalid BVH file '%s' (no entry with node size %d, triangle size %d)\n", file, node_size, tri_size);
    return found;
}
int32_t rodent_b200_append_bvh_bin(const char* file, int32_t node_size, int32_t tri_size,
                                   const void* nodes, int64_t num_nodes, const void* tris, int64_t num_tris) {
    std::FILE* f = std::fopen(file, "ab");
    if (!f) return 0;
    const uint32_t ns = uint32_t(node_size), ts = uint32_t(tri_size);
    bool ok = false;
    try {
        ok = std::fwrite(&ns, 4, 1, f) == 1 && std::fwrite(&ts, 4, 1, f) == 1 &&
             write_one_buffer(f, nodes, num_nodes * node_size) && write_one_buffer(f, tris, num_tris * tri_size);
    } catch (const std::bad_alloc&) { ok = false; }
    return (std::fclose(f) == 0 && ok) ? 1 : 0;
}

}  // extern "C"
