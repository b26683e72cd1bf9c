// JPEG -> the RGBA8 image the reference's load_jpg produces (src/driver/image.cpp:186-238).
//
// The reference decodes with libjpeg at its defaults and copies `output_components` bytes per pixel into a zero-filled
// RGBA image, bottom row first, then applies gamma 2.2 (:10-18): colour images come out as (r, g, b, 0), grey ones as
// (grey, 0, 0, 0).  libjpeg is not part of this image, so the decoder is written here from the JPEG specification
// (ITU-T T.81: baseline / extended sequential and progressive Huffman, 8 bit, interleaved or one scan per component,
// restart intervals) with libjpeg's default choices for the
// parts the standard leaves open, so that pixels come out the same: the slow-but-accurate integer IDCT (jidctint.c),
// "fancy" triangle-filter chroma upsampling for 2x1 and 2x2 subsampling (jdsample.c) and the fixed-point YCbCr -> RGB
// tables (jdcolor.c).  tests/test_textures.py checks it pixel for pixel against PIL, which wraps libjpeg-turbo.
// Arithmetic-coded, lossless and 12-bit files are reported as unsupported.
#include "scene.h"

#include <cmath>
#include <cstring>
#include <fstream>
#include <iterator>
#include <new>

namespace rb200 {
namespace {

struct Huffman {
    uint8_t bits[17] = {0}, vals[256] = {0};
    int mincode[17], maxcode[18], valptr[17];
    bool present = false;
    void build() {
        int code = 0, k = 0;
        for (int l = 1; l <= 16; l++) {
            valptr[l] = k; mincode[l] = code;
            code += bits[l]; k += bits[l];
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7FFFFFFF;
    }
};

struct Component { int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0, dc_pred = 0; int blocks_w = 0, blocks_h = 0; std::vector<int16_t> coef; std::vector<uint8_t> plane; int stride = 0, rows = 0; };

struct BitReader {
    const uint8_t* p; const uint8_t* end;
    uint32_t acc = 0; int count = 0; bool hit_marker = false;
    void fill() {
        while (count <= 24) {
            int b = 0;
            if (!hit_marker && p < end) {
                b = *p;
                if (b == 0xFF) {
                    if (p + 1 < end && p[1] == 0x00) p += 2;             // stuffed zero
                    else { hit_marker = true; b = 0; }                   // a marker: feed zeros, as libjpeg does
                } else p++;
            }
            acc |= uint32_t(b) << (24 - count);
            count += 8;
        }
    }
    int bit() { if (count == 0) fill(); const int b = acc >> 31; acc <<= 1; count--; return b; }
    int bits(int n) { if (n == 0) return 0; if (count < n) fill(); const int v = int(acc >> (32 - n)); acc <<= n; count -= n; return v; }
    void reset() { acc = 0; count = 0; hit_marker = false; }
};

int decode_symbol(BitReader& br, const Huffman& h, bool& ok) {
    int code = 0;
    for (int l = 1; l <= 16; l++) {
        code = (code << 1) | br.bit();
        if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) return h.vals[h.valptr[l] + code - h.mincode[l]];
    }
    ok = false;
    return 0;
}
int extend(int v, int n) { return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v; }

const uint8_t kZigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                             35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

inline uint8_t range_limit(int x) {                      // libjpeg's IDCT range-limit table: clamp(x + 128), indexed with x & RANGE_MASK (10 bits)
    x = ((x + 512) & 1023) - 512;
    x += 128;
    return uint8_t(x < 0 ? 0 : x > 255 ? 255 : x);
}
inline int64_t descale(int64_t x, int n) { return (x + (int64_t(1) << (n - 1))) >> n; }

// jpeg_idct_islow (jidctint.c): 13-bit constants, 2 extra bits kept between the passes.
void idct_islow(const int* in /* dequantised, natural order */, uint8_t* out, int stride) {
    constexpr int C = 13, P = 2;
    constexpr int F_0_298 = 2446, F_0_390 = 3196, F_0_541 = 4433, F_0_765 = 6270, F_0_899 = 7373, F_1_175 = 9633,
                  F_1_501 = 12299, F_1_847 = 15137, F_1_961 = 16069, F_2_053 = 16819, F_2_562 = 20995, F_3_072 = 25172;
    int ws[64];
    // 64-bit intermediates, as libjpeg's JLONG (long) is on the reference's platform: corrupt coefficients cannot overflow
    using L = int64_t;
    auto pass = [&](L d0, L d1, L d2, L d3, L d4, L d5, L d6, L d7, L* o) {
        L z2 = d2, z3 = d6;
        L z1 = (z2 + z3) * F_0_541;
        L tmp2 = z1 + z3 * (-F_1_847), tmp3 = z1 + z2 * F_0_765;
        z2 = d0; z3 = d4;
        L tmp0 = (z2 + z3) * (1 << C), tmp1 = (z2 - z3) * (1 << C);
        const L tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = d7; tmp1 = d5; tmp2 = d3; tmp3 = d1;
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2; L z4 = tmp1 + tmp3;
        const L z5 = (z3 + z4) * F_1_175;
        tmp0 *= F_0_298; tmp1 *= F_2_053; tmp2 *= F_3_072; tmp3 *= F_1_501;
        z1 *= -F_0_899; z2 *= -F_2_562; z3 *= -F_1_961; z4 *= -F_0_390;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        o[0] = tmp10 + tmp3; o[7] = tmp10 - tmp3; o[1] = tmp11 + tmp2; o[6] = tmp11 - tmp2;
        o[2] = tmp12 + tmp1; o[5] = tmp12 - tmp1; o[3] = tmp13 + tmp0; o[4] = tmp13 - tmp0;
    };
    for (int c = 0; c < 8; c++) {                        // pass 1: columns
        L o[8];
        pass(in[c], in[8 + c], in[16 + c], in[24 + c], in[32 + c], in[40 + c], in[48 + c], in[56 + c], o);
        for (int r = 0; r < 8; r++) ws[8 * r + c] = int(descale(o[r], C - P));
    }
    for (int r = 0; r < 8; r++) {                        // pass 2: rows
        L o[8];
        const int* w = ws + 8 * r;
        pass(w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], o);
        for (int c = 0; c < 8; c++) out[r * stride + c] = range_limit(int(descale(o[c], C + P + 3) & 0x3FF));
    }
}

uint16_t be16(const uint8_t* p) { return uint16_t(p[0] << 8 | p[1]); }

}  // namespace

static bool decode_jpg(const std::string& path, int& width, int& height, std::vector<uint32_t>& pixels, std::string& why);
bool load_jpg(const std::string& path, int& width, int& height, std::vector<uint32_t>& pixels, std::string& why) {
    try { return decode_jpg(path, width, height, pixels, why); }
    catch (const std::bad_alloc&) { why = "out of memory"; return false; }
}
static bool decode_jpg(const std::string& path, int& width, int& height, std::vector<uint32_t>& pixels, std::string& why) {
    std::ifstream in(path, std::ios::binary);
    if (!in) { why = "cannot open file"; return false; }
    const std::vector<uint8_t> file((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    if (file.size() < 4 || file[0] != 0xFF || file[1] != 0xD8) { why = "not a JPEG file"; return false; }

    uint16_t qt[4][64] = {};
    Huffman dc[4], ac[4];
    std::vector<Component> comps;
    int w = 0, h = 0, restart_interval = 0, hmax = 1, vmax = 1, mcus_x = 0, mcus_y = 0;
    bool have_frame = false, progressive = false, adobe_rgb = false, saw_scan = false;
    size_t pos = 2;
    while (pos + 4 <= file.size()) {
        if (file[pos] != 0xFF) { pos++; continue; }                          // (garbage between segments is skipped, as libjpeg does)
        const int m = file[pos + 1];
        if (m == 0xFF) { pos++; continue; }                                  // fill bytes
        pos += 2;
        if (m == 0x00 || m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
        if (m == 0xD9) break;
        if (pos + 2 > file.size()) { why = "truncated segment"; return false; }
        const size_t len = be16(&file[pos]);
        if (len < 2 || pos + len > file.size()) { why = "truncated segment"; return false; }
        const uint8_t* s = &file[pos + 2];
        const size_t n = len - 2;
        pos += len;
        if (m == 0xDB) {                                                     // DQT
            for (size_t i = 0; i < n;) {
                const int pq = s[i] >> 4, tq = s[i] & 15;
                if (tq > 3 || i + 1 + (pq ? 128 : 64) > n) { why = "bad DQT"; return false; }
                for (int k = 0; k < 64; k++) qt[tq][kZigzag[k]] = pq ? be16(&s[i + 1 + 2 * k]) : s[i + 1 + k];
                i += 1 + (pq ? 128 : 64);
            }
        } else if (m == 0xC4) {                                              // DHT
            for (size_t i = 0; i < n;) {
                const int tc = s[i] >> 4, th = s[i] & 15;
                if (tc > 1 || th > 3 || i + 17 > n) { why = "bad DHT"; return false; }
                Huffman& t = tc ? ac[th] : dc[th];
                int total = 0;
                for (int l = 1; l <= 16; l++) { t.bits[l] = s[i + l]; total += t.bits[l]; }
                if (total > 256 || i + 17 + total > n) { why = "bad DHT"; return false; }
                std::memcpy(t.vals, &s[i + 17], total);
                t.build(); t.present = true;
                i += 17 + total;
            }
        } else if (m == 0xC0 || m == 0xC1 || m == 0xC2) {                    // SOF0 / SOF1: sequential, SOF2: progressive (Huffman)
            if (have_frame) { why = "more than one frame"; return false; }
            if (n < 6 || s[0] != 8) { why = "only 8-bit samples are supported"; return false; }
            progressive = m == 0xC2;
            h = be16(&s[1]); w = be16(&s[3]);
            const int nc = s[5];
            if ((nc != 1 && nc != 3) || n < size_t(6 + 3 * nc) || w == 0 || h == 0) { why = "unsupported number of components"; return false; }
            if (uint64_t(w) * uint64_t(h) > (uint64_t(1) << 28)) { why = "unreasonable image size"; return false; }
            comps.resize(nc);
            for (int c = 0; c < nc; c++) {
                comps[c].id = s[6 + 3 * c]; comps[c].h = s[7 + 3 * c] >> 4; comps[c].v = s[7 + 3 * c] & 15; comps[c].tq = s[8 + 3 * c] & 3;
            }
            if (nc == 3) {
                const bool ok = comps[1].h == 1 && comps[1].v == 1 && comps[2].h == 1 && comps[2].v == 1 &&
                                (comps[0].h == 1 || comps[0].h == 2) && (comps[0].v == 1 || comps[0].v == 2) && !(comps[0].h == 1 && comps[0].v == 2);
                if (!ok) { why = "unsupported chroma subsampling"; return false; }
            } else comps[0].h = comps[0].v = 1;                              // a single component is never interleaved
            for (auto& c : comps) { hmax = std::max(hmax, c.h); vmax = std::max(vmax, c.v); }
            mcus_x = (w + 8 * hmax - 1) / (8 * hmax); mcus_y = (h + 8 * vmax - 1) / (8 * vmax);
            {   // every block costs at least one bit of entropy-coded data: refuse sizes the file cannot hold
                uint64_t blocks = 0;
                for (auto& c : comps) blocks += uint64_t(mcus_x) * c.h * uint64_t(mcus_y) * c.v;
                if (blocks > uint64_t(file.size()) * 8) { why = "image data too short for the stated size"; return false; }
            }
            for (auto& c : comps) {                                          // planes and coefficients padded to whole MCUs
                c.blocks_w = mcus_x * c.h; c.blocks_h = mcus_y * c.v;
                c.stride = c.blocks_w * 8; c.rows = c.blocks_h * 8;
                c.coef.assign(size_t(c.blocks_w) * c.blocks_h * 64, 0);
            }
            have_frame = true;
        } else if (m >= 0xC5 && m <= 0xCF && m != 0xC8 && m != 0xCC) {
            why = "this JPEG coding process is not supported";
            return false;
        } else if (m == 0xDD) {
            if (n >= 2) restart_interval = be16(s);
        } else if (m == 0xEE) {                                              // Adobe: transform 0 with three components = RGB
            if (n >= 12 && !std::memcmp(s, "Adobe", 5)) adobe_rgb = s[11] == 0;
        } else if (m == 0xDA) {                                              // SOS: one scan
            if (!have_frame) { why = "SOS before SOF"; return false; }
            const int ns = s[0];
            if (ns < 1 || ns > int(comps.size()) || n < size_t(1 + 2 * ns + 3)) { why = "bad SOS"; return false; }
            Component* sc[3];
            for (int k = 0; k < ns; k++) {
                sc[k] = nullptr;
                for (auto& cc : comps) if (cc.id == s[1 + 2 * k]) sc[k] = &cc;
                if (!sc[k]) { why = "bad SOS"; return false; }
                sc[k]->td = (s[2 + 2 * k] >> 4) & 3; sc[k]->ta = s[2 + 2 * k] & 3;
            }
            const int Ss = s[1 + 2 * ns], Se = s[2 + 2 * ns], Ah = s[3 + 2 * ns] >> 4, Al = s[3 + 2 * ns] & 15;
            if (!progressive && (Ss != 0 || Se != 63 || Ah != 0 || Al != 0)) { why = "bad SOS"; return false; }
            if (progressive && (Ss > Se || Se > 63 || Al > 13 || (Ss == 0 && Se != 0) || (Ss > 0 && ns != 1))) { why = "bad progressive scan"; return false; }
            const bool need_dc = Ss == 0 && Ah == 0, need_ac = Se > 0;
            for (int k = 0; k < ns; k++)
                if ((need_dc && !dc[sc[k]->td].present) || (need_ac && !ac[sc[k]->ta].present)) { why = "missing Huffman table"; return false; }

            // blocks of this scan: interleaved MCUs, or (one component) its own blocks row by row, T.81 A.2.2 / A.2.3
            const int scan_w = ns > 1 ? mcus_x : ((w * sc[0]->h + hmax - 1) / hmax + 7) / 8;
            const int scan_h = ns > 1 ? mcus_y : ((h * sc[0]->v + vmax - 1) / vmax + 7) / 8;
            BitReader br{&file[pos], file.data() + file.size()};
            int restarts_left = restart_interval, eobrun = 0;
            for (auto& c : comps) c.dc_pred = 0;
            for (int my = 0; my < scan_h; my++)
                for (int mx = 0; mx < scan_w; mx++) {
                    if (restart_interval && restarts_left == 0) {            // skip to the RSTn marker, reset the predictors
                        const uint8_t* q = br.p;
                        while (q + 1 < br.end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) q++;
                        if (q + 1 >= br.end) { why = "missing restart marker"; return false; }
                        br.p = q + 2; br.reset();
                        for (auto& c : comps) c.dc_pred = 0;
                        eobrun = 0;
                        restarts_left = restart_interval;
                    }
                    for (int k = 0; k < ns; k++) {
                        Component& c = *sc[k];
                        const int nbx = ns > 1 ? c.h : 1, nby = ns > 1 ? c.v : 1;
                        for (int by = 0; by < nby; by++)
                            for (int bx = 0; bx < nbx; bx++) {
                                int16_t* blk = c.coef.data() + (size_t(my * nby + by) * c.blocks_w + (mx * nbx + bx)) * 64;
                                bool ok = true;
                                if (!progressive) {                          // sequential: T.81 F.2.2
                                    const int t = decode_symbol(br, dc[c.td], ok);
                                    if (!ok || t > 11) { why = "corrupt entropy-coded data"; return false; }
                                    c.dc_pred = int(uint32_t(c.dc_pred) + uint32_t(t ? extend(br.bits(t), t) : 0));
                                    blk[0] = int16_t(c.dc_pred);             // a coefficient is a JCOEF (short) in libjpeg
                                    for (int kk = 1; kk < 64;) {
                                        const int rs = decode_symbol(br, ac[c.ta], ok);
                                        if (!ok) { why = "corrupt entropy-coded data"; return false; }
                                        const int r = rs >> 4, sz = rs & 15;
                                        if (sz == 0) { if (r == 15) { kk += 16; continue; } break; }
                                        kk += r;
                                        if (kk > 63) { why = "corrupt entropy-coded data"; return false; }
                                        blk[kZigzag[kk]] = int16_t(extend(br.bits(sz), sz));
                                        kk++;
                                    }
                                } else if (Ss == 0) {                        // progressive DC: G.1.2.1
                                    if (Ah == 0) {
                                        const int t = decode_symbol(br, dc[c.td], ok);
                                        if (!ok || t > 11) { why = "corrupt entropy-coded data"; return false; }
                                        c.dc_pred = int(uint32_t(c.dc_pred) + uint32_t(t ? extend(br.bits(t), t) : 0));
                                        blk[0] = int16_t(c.dc_pred * (1 << Al));
                                    } else if (br.bit()) blk[0] = int16_t(blk[0] | (1 << Al));
                                } else if (Ah == 0) {                        // progressive AC, first pass: G.1.2.2
                                    if (eobrun > 0) { eobrun--; continue; }
                                    for (int kk = Ss; kk <= Se; kk++) {
                                        const int rs = decode_symbol(br, ac[c.ta], ok);
                                        if (!ok) { why = "corrupt entropy-coded data"; return false; }
                                        const int r = rs >> 4, sz = rs & 15;
                                        if (sz) {
                                            kk += r;
                                            if (kk > Se) { why = "corrupt entropy-coded data"; return false; }
                                            blk[kZigzag[kk]] = int16_t(extend(br.bits(sz), sz) * (1 << Al));
                                        } else if (r == 15) kk += 15;
                                        else { eobrun = (1 << r) + (r ? br.bits(r) : 0) - 1; break; }
                                    }
                                } else {                                     // progressive AC, refinement: G.1.2.3
                                    const int p1 = 1 << Al, m1 = -(1 << Al);
                                    auto refine = [&](int16_t& coef) {       // one correction bit for a coefficient that is already non-zero
                                        if (br.bit() && (coef & p1) == 0) coef = int16_t(coef >= 0 ? coef + p1 : coef + m1);
                                    };
                                    int kk = Ss;
                                    if (eobrun == 0) {
                                        for (; kk <= Se; kk++) {
                                            const int rs = decode_symbol(br, ac[c.ta], ok);
                                            if (!ok) { why = "corrupt entropy-coded data"; return false; }
                                            int r = rs >> 4, value = 0;
                                            if (rs & 15) value = br.bit() ? p1 : m1;          // the size must be 1
                                            else if (r != 15) { eobrun = (1 << r) + (r ? br.bits(r) : 0); break; }
                                            for (; kk <= Se; kk++) {          // pass the non-zero history, stop at the r-th zero
                                                int16_t& coef = blk[kZigzag[kk]];
                                                if (coef != 0) refine(coef);
                                                else if (--r < 0) break;
                                            }
                                            if (value && kk <= Se) blk[kZigzag[kk]] = int16_t(value);
                                        }
                                    }
                                    if (eobrun > 0) {
                                        for (; kk <= Se; kk++) { int16_t& coef = blk[kZigzag[kk]]; if (coef != 0) refine(coef); }
                                        eobrun--;
                                    }
                                }
                            }
                    }
                    restarts_left--;
                }
            saw_scan = true;
            pos = size_t(br.p - file.data());                                // the reader stopped at the next marker (or the end)
        }
    }
    if (!saw_scan) { why = "no scan found"; return false; }

    // ---- dequantisation + IDCT into the component planes ----
    for (auto& c : comps) {
        c.plane.assign(size_t(c.stride) * c.rows, 0);
        for (int by = 0; by < c.blocks_h; by++)
            for (int bx = 0; bx < c.blocks_w; bx++) {
                const int16_t* blk = c.coef.data() + (size_t(by) * c.blocks_w + bx) * 64;
                int coef[64];
                for (int k = 0; k < 64; k++) coef[k] = int(blk[k]) * qt[c.tq][k];
                idct_islow(coef, c.plane.data() + size_t(by) * 8 * c.stride + bx * 8, c.stride);
            }
        std::vector<int16_t>().swap(c.coef);
    }

    // ---- upsampling (jdsample.c, fancy) + colour conversion (jdcolor.c) ----
    width = w; height = h;
    pixels.assign(size_t(w) * h, 0);
    uint32_t gamma_lut[256];
    for (int v = 0; v < 256; v++) gamma_lut[v] = uint32_t(uint8_t(std::pow(float(v) * (1.0f / 255.0f), 2.2f) * 255.0f));   // image.cpp:10-18
    if (comps.size() == 1) {
        for (int y = 0; y < h; y++) {
            uint32_t* dst = pixels.data() + size_t(h - 1 - y) * w;           // bottom row first, image.cpp:223
            for (int x = 0; x < w; x++) dst[x] = gamma_lut[comps[0].plane[size_t(y) * comps[0].stride + x]];   // (grey, 0, 0, 0): :226-227
        }
        return true;
    }
    const int H = comps[0].h, V = comps[0].v;                                // chroma is (1, 1): subsampled by H x V
    const int cw = (w + H - 1) / H, chh = (h + V - 1) / V;                   // downsampled_width / height of the chroma planes
    std::vector<uint8_t> up[2];
    for (int k = 0; k < 2; k++) {
        const Component& c = comps[1 + k];
        std::vector<uint8_t>& o = up[k];
        const int ow = cw * H;
        o.assign(size_t(ow) * h, 0);
        auto row = [&](int r) { return c.plane.data() + size_t(std::min(std::max(r, 0), chh - 1)) * c.stride; };
        if (H == 1 && V == 1) {
            for (int y = 0; y < h; y++) std::memcpy(&o[size_t(y) * ow], row(y), cw);
        } else if (cw <= 2) {                                                // jinit_upsampler: fancy only if downsampled_width > 2, else replication
            for (int y = 0; y < h; y++)
                for (int x = 0; x < ow; x++) o[size_t(y) * ow + x] = row(y / V)[x / H];
        } else if (H == 2 && V == 1) {                                       // h2v1_fancy_upsample
            for (int y = 0; y < h; y++) {
                const uint8_t* in_row = row(y); uint8_t* out = &o[size_t(y) * ow];
                out[0] = in_row[0]; out[1] = uint8_t((in_row[0] * 3 + in_row[1] + 2) >> 2);
                for (int x = 1; x < cw - 1; x++) {
                    const int v3 = in_row[x] * 3;
                    out[2 * x] = uint8_t((v3 + in_row[x - 1] + 1) >> 2); out[2 * x + 1] = uint8_t((v3 + in_row[x + 1] + 2) >> 2);
                }
                out[2 * (cw - 1)] = uint8_t((in_row[cw - 1] * 3 + in_row[cw - 2] + 1) >> 2); out[2 * (cw - 1) + 1] = in_row[cw - 1];
            }
        } else {                                                             // h2v2_fancy_upsample
            for (int y = 0; y < h; y++) {
                const int r = y >> 1;
                const uint8_t* near_row = row(r);
                const uint8_t* far_row = row((y & 1) ? r + 1 : r - 1);        // the context rows replicate at the image edges (jdmainct.c)
                uint8_t* out = &o[size_t(y) * ow];
                auto colsum = [&](int x) { return near_row[x] * 3 + far_row[x]; };
                int last = colsum(0), cur = last, next = colsum(1);
                out[0] = uint8_t((cur * 4 + 8) >> 4); out[1] = uint8_t((cur * 3 + next + 7) >> 4);
                for (int x = 1; x < cw - 1; x++) {
                    last = cur; cur = next; next = colsum(x + 1);
                    out[2 * x] = uint8_t((cur * 3 + last + 8) >> 4); out[2 * x + 1] = uint8_t((cur * 3 + next + 7) >> 4);
                }
                last = cur; cur = next;
                out[2 * (cw - 1)] = uint8_t((cur * 3 + last + 8) >> 4); out[2 * (cw - 1) + 1] = uint8_t((cur * 4 + 7) >> 4);
            }
        }
    }
    // build_ycc_rgb_table / ycc_rgb_convert
    int cr_r[256], cb_b[256], cr_g[256], cb_g[256];
    for (int i = 0; i < 256; i++) {
        const int x = i - 128;
        cr_r[i] = (91881 * x + 32768) >> 16;            // FIX(1.40200)
        cb_b[i] = (116130 * x + 32768) >> 16;           // FIX(1.77200)
        cr_g[i] = -46802 * x;                           // FIX(0.71414)
        cb_g[i] = -22554 * x + 32768;                   // FIX(0.34414)
    }
    auto clamp8 = [](int v) { return v < 0 ? 0 : v > 255 ? 255 : v; };
    const int ow = cw * H;
    for (int y = 0; y < h; y++) {
        uint32_t* dst = pixels.data() + size_t(h - 1 - y) * w;
        const uint8_t* yr = comps[0].plane.data() + size_t(y) * comps[0].stride;
        const uint8_t* cb = &up[0][size_t(y) * ow];
        const uint8_t* cr = &up[1][size_t(y) * ow];
        for (int x = 0; x < w; x++) {
            int r, g, b;
            if (adobe_rgb) { r = yr[x]; g = cb[x]; b = cr[x]; }
            else {
                r = clamp8(yr[x] + cr_r[cr[x]]);
                g = clamp8(yr[x] + ((cb_g[cb[x]] + cr_g[cr[x]]) >> 16));
                b = clamp8(yr[x] + cb_b[cb[x]]);
            }
            dst[x] = gamma_lut[r] | gamma_lut[g] << 8 | gamma_lut[b] << 16;  // alpha stays 0 (:213, :226-227)
        }
    }
    return true;
}

}  // namespace rb200
