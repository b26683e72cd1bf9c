// Device-side shading math of the path tracer: surface element, BSDFs, light sampling,
// RNG -- the table-driven counterpart of the shaders the reference generates per scene
// (src/driver/converter.cpp:857-919) from src/render/{material,light,geometry}.impala and
// src/core/{random,common,matrix,color}.impala.  Plain fp32, operation order of the
// reference; the library is compiled without FMA contraction, so apart from sinf/cosf
// (CUDA's vs libm's) every sample follows the CPU oracle's arithmetic.
#pragma once

#include "common.cuh"
#include "poly_trig.h"

namespace rb200 {
namespace shade {

struct V3 { float x, y, z; };
struct Col { float r, g, b; };
struct M3 { V3 c0, c1, c2; };

__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float length(V3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ V3 normalize(V3 a) { return a * (1.0f / length(a)); }                      // vector.impala:82
__device__ __forceinline__ V3 reflect(V3 v, V3 n) { return n * (2.0f * dot(n, v)) - v; }               // vector.impala:74
__device__ __forceinline__ float lerp1(float a, float b, float k) { return (1.0f - k) * a + k * b; }     // common.impala:118-120
__device__ __forceinline__ float lerp2(float a, float b, float c, float k1, float k2) { return (1.0f - k1 - k2) * a + k1 * b + k2 * c; }
__device__ __forceinline__ Col col(float r, float g, float b) { return Col{r, g, b}; }
__device__ __forceinline__ Col operator*(Col a, Col b) { return col(a.r * b.r, a.g * b.g, a.b * b.b); }
__device__ __forceinline__ Col operator*(Col a, float f) { return col(a.r * f, a.g * f, a.b * f); }
__device__ __forceinline__ Col lerp(Col a, Col b, float t) { return col(lerp1(a.r, b.r, t), lerp1(a.g, b.g, t), lerp1(a.b, b.b, t)); }
__device__ __forceinline__ float luminance(Col c) { return c.r * 0.2126f + c.g * 0.7152f + c.b * 0.0722f; }   // color.impala:33-35

constexpr float kPi = 3.14159265359f;      // common.impala:7
constexpr float kOffset = 0.001f;          // renderer.impala:64

// matrix.impala:29-39, 107-111
__device__ __forceinline__ M3 orthonormal(V3 n) {
    const float sign = n.z >= 0.0f ? 1.0f : -1.0f;
    const float a = -1.0f / (sign + n.z);
    const float b = n.x * n.y * a;
    return M3{v3(1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x), v3(b, sign + n.y * n.y * a, -n.y), n};
}
__device__ __forceinline__ V3 mul(const M3& m, V3 v) {
    return v3(dot(v3(m.c0.x, m.c1.x, m.c2.x), v), dot(v3(m.c0.y, m.c1.y, m.c2.y), v), dot(v3(m.c0.z, m.c1.z, m.c2.z), v));
}

// random.impala:7-31, 116-126
__device__ __forceinline__ int xorshift(unsigned& seed) {
    unsigned x = seed;
    x = x == 0u ? 1u : x;
    x ^= x << 13; x ^= x >> 17; x ^= x << 5;
    seed = x;
    return int(x);
}
__device__ __forceinline__ float randf(unsigned& rnd) {
    const unsigned x = unsigned(xorshift(rnd));
    return __uint_as_float((127u << 23) | (x & 0x7FFFFFu)) - 1.0f;
}
__device__ __forceinline__ unsigned fnv_hash(unsigned h, unsigned d) {
    h = (h * 16777619u) ^ (d & 0xFFu);
    h = (h * 16777619u) ^ ((d >> 8) & 0xFFu);
    h = (h * 16777619u) ^ ((d >> 16) & 0xFFu);
    h = (h * 16777619u) ^ ((d >> 24) & 0xFFu);
    return h;
}

// common.impala:42-61
__device__ __forceinline__ float fastlog2(float x) {
    const unsigned vx = __float_as_uint(x);
    const unsigned mx = (vx & 0x007FFFFFu) | 0x3f000000u;
    const float y = float(vx) * 1.1920928955078125e-7f;
    const float z = __uint_as_float(mx);
    return y - 124.22551499f - 1.498030302f * z - 1.72587999f / (0.3520887068f + z);
}
__device__ __forceinline__ float fastpow2(float p) {
    const float offset = p < 0.0f ? 1.0f : 0.0f;
    const float clipp = p < -126.0f ? -126.0f : p;
    const int w = int(clipp);
    const float z = clipp - float(w) + offset;
    const int v = int(float(1u << 23) * (clipp + 121.2740575f + 27.7280233f / (4.84252568f - z) - 1.49012907f * z));
    return __int_as_float(v);
}
__device__ __forceinline__ float fastpow(float x, float p) { return fastpow2(p * fastlog2(x)); }

__device__ __forceinline__ float positive_cos(V3 a, V3 b) { const float c = dot(a, b); return c >= 0.0f ? c : 0.0f; }
__device__ __forceinline__ float cosine_hemisphere_pdf(float c) { return c * (1.0f / kPi); }
__device__ __forceinline__ float cosine_power_hemisphere_pdf(float c, float k) { return fastpow(c, k) * (k + 1.0f) * (1.0f / (2.0f * kPi)); }

struct DirSample { V3 dir; float pdf; };
// Test switch (rodent_b200_tune "render_poly_trig"): sin / cos from poly_trig.h instead of CUDA's sinf / cosf.
static __device__ int g_poly_trig = 0;                  // one per translation unit; render.cu sets its own
__device__ __forceinline__ DirSample make_dir_sample(float c, float s, float phi, float pdf) {          // random.impala:39-48
    float sn, cs;
    if (g_poly_trig) rb_poly_sincos(phi, &sn, &cs);
    else { sn = sinf(phi); cs = cosf(phi); }
    return DirSample{v3(s * cs, s * sn, c), pdf};
}
__device__ __forceinline__ DirSample sample_cosine_hemisphere(float u, float v) {                       // random.impala:72-77
    const float c = sqrtf(1.0f - v), s = sqrtf(v);
    return make_dir_sample(c, s, 2.0f * kPi * u, cosine_hemisphere_pdf(c));
}
__device__ __forceinline__ DirSample sample_cosine_power_hemisphere(float k, float u, float v) {        // random.impala:85-98
    const float c = fminf(fastpow(v, 1.0f / (k + 1.0f)), 1.0f);
    const float s = sqrtf(1.0f - c * c);
    const float pow_c_k = c != 0.0f ? v / c : 0.0f;
    return make_dir_sample(c, s, 2.0f * kPi * u, pow_c_k * (k + 1.0f) * (1.0f / (2.0f * kPi)));
}

// geometry.impala:21-53
struct Surf {
    bool is_entering;
    V3 point, face_normal;
    M3 local;
};
__device__ __forceinline__ V3 load3(const float4* base, int i) { const float4 v = __ldg(base + i); return v3(v.x, v.y, v.z); }

__device__ __forceinline__ Surf surface_element(const float4* __restrict__ normals, const float4* __restrict__ face_normals,
                                                const int4* __restrict__ indices, V3 org, V3 dir, int prim, float t, float u, float v) {
    const int4 idx = __ldg(indices + prim);
    const V3 fn = load3(face_normals, prim);
    const V3 n0 = load3(normals, idx.x), n1 = load3(normals, idx.y), n2 = load3(normals, idx.z);
    const V3 normal = normalize(v3(lerp2(n0.x, n1.x, n2.x, u, v), lerp2(n0.y, n1.y, n2.y, u, v), lerp2(n0.z, n1.z, n2.z, u, v)));
    Surf s;
    s.is_entering = dot(dir, fn) <= 0.0f;
    s.point = org + dir * t;
    s.face_normal = s.is_entering ? fn : -fn;
    s.local = orthonormal(dot(dir, normal) <= 0.0f ? normal : -normal);
    return s;
}

// make_texture(repeat border, bilinear filter, make_image_rgba32), src/render/image.impala:24-92
__device__ __forceinline__ Col rgba32_texture(const unsigned* __restrict__ pixels, int width, int height, float u, float v) {
    u = u - floorf(u); v = v - floorf(v);
    const float fu = u * float(width), fv = v * float(height);
    const int x0 = min(int(fu), width - 1), y0 = min(int(fv), height - 1);
    const int x1 = min(x0 + 1, width - 1), y1 = min(y0 + 1, height - 1);
    const float kx = fu - float(int(fu)), ky = fv - float(int(fv));
    auto px = [&](int x, int y) {
        const unsigned p = __ldg(pixels + y * width + x);
        return col(float(p & 0xFFu) * (1.0f / 255.0f), float((p >> 8) & 0xFFu) * (1.0f / 255.0f), float((p >> 16) & 0xFFu) * (1.0f / 255.0f));
    };
    const Col p00 = px(x0, y0), p10 = px(x1, y0), p01 = px(x0, y1), p11 = px(x1, y1);
    return col(lerp1(lerp1(p00.r, p10.r, kx), lerp1(p01.r, p11.r, kx), ky),
               lerp1(lerp1(p00.g, p10.g, kx), lerp1(p01.g, p11.g, kx), ky),
               lerp1(lerp1(p00.b, p10.b, kx), lerp1(p01.b, p11.b, kx), ky));
}

// The textured part of a material's shader (converter.cpp:876-903): kd / ks looked up at the hit's texture coordinates
// (surf.attr(0), the lerp of the three vertices' uv: geometry.impala:44-47), the mix weight recomputed from them.
__device__ __forceinline__ void apply_textures(RodentMaterial& m, const float4* __restrict__ texcoords, const int4* __restrict__ indices,
                                               const RodentTexture* __restrict__ textures, const unsigned* __restrict__ texture_pixels,
                                               int prim, float u, float v) {
    if ((m.map_kd | m.map_ks | m.map_ke) == 0) return;
    const int4 idx = __ldg(indices + prim);
    const float4 t0 = __ldg(texcoords + idx.x), t1 = __ldg(texcoords + idx.y), t2 = __ldg(texcoords + idx.z);
    const float tu = lerp2(t0.x, t1.x, t2.x, u, v), tv = lerp2(t0.y, t1.y, t2.y, u, v);
    if (m.map_kd) {
        const RodentTexture t = textures[m.map_kd - 1];
        const Col c = rgba32_texture(texture_pixels + t.offset, t.width, t.height, tu, tv);
        m.kd[0] = c.r; m.kd[1] = c.g; m.kd[2] = c.b;
    }
    if (m.map_ks) {
        const RodentTexture t = textures[m.map_ks - 1];
        const Col c = rgba32_texture(texture_pixels + t.offset, t.width, t.height, tu, tv);
        m.ks[0] = c.r; m.ks[1] = c.g; m.ks[2] = c.b;
    }
    if (m.map_ke) {                                              // the light's colour at this point (converter.cpp:794-801)
        const RodentTexture t = textures[m.map_ke - 1];
        const Col c = rgba32_texture(texture_pixels + t.offset, t.width, t.height, tu, tv);
        m.ke[0] = c.r; m.ke[1] = c.g; m.ke[2] = c.b;
    }
    if (m.bsdf == RODENT_BSDF_MIX) {
        const float lum_ks = luminance(col(m.ks[0], m.ks[1], m.ks[2])), lum_kd = luminance(col(m.kd[0], m.kd[1], m.kd[2]));
        m.mix_k = (lum_ks + lum_kd == 0.0f) ? 0.0f : lum_ks / (lum_ks + lum_kd);
    }
}

// Radiance a light emits from the point with barycentrics (u, v): its constant colour, or its material's map_Ke there.
__device__ __forceinline__ Col light_color(const RodentLight& l, const float4* __restrict__ texcoords, const int4* __restrict__ indices,
                                           const RodentTexture* __restrict__ textures, const unsigned* __restrict__ texture_pixels, float u, float v) {
    if (l.map_ke == 0) return col(l.color[0], l.color[1], l.color[2]);
    const int4 idx = __ldg(indices + l.prim);
    const float4 t0 = __ldg(texcoords + idx.x), t1 = __ldg(texcoords + idx.y), t2 = __ldg(texcoords + idx.z);
    const RodentTexture t = textures[l.map_ke - 1];
    return rgba32_texture(texture_pixels + t.offset, t.width, t.height, lerp2(t0.x, t1.x, t2.x, u, v), lerp2(t0.y, t1.y, t2.y, u, v));
}

// material.impala:54-192, interpreted from the RodentMaterial table
struct BsdfSample { V3 in_dir; float pdf, cos; Col color; };

__device__ __forceinline__ BsdfSample make_bsdf_sample(const Surf& s, V3 in_dir, float pdf, float cosv, Col color, bool inverted) {
    const bool valid = (pdf > 0.0f) && (inverted != (dot(in_dir, s.face_normal) > 0.0f));
    return BsdfSample{in_dir, valid ? pdf : 1.0f, cosv, valid ? color : col(0, 0, 0)};
}
__device__ __forceinline__ Col kd_of(const RodentMaterial& m) { return col(m.kd[0], m.kd[1], m.kd[2]); }
__device__ __forceinline__ Col ks_of(const RodentMaterial& m) { return col(m.ks[0], m.ks[1], m.ks[2]); }
__device__ __forceinline__ bool is_specular(const RodentMaterial& m) { return m.bsdf == RODENT_BSDF_MIRROR || m.bsdf == RODENT_BSDF_GLASS; }

__device__ __forceinline__ Col diffuse_eval(const RodentMaterial& m) { return kd_of(m) * (1.0f / kPi); }
__device__ __forceinline__ float diffuse_pdf(const Surf& s, V3 in_dir) { return cosine_hemisphere_pdf(positive_cos(in_dir, s.local.c2)); }
__device__ __forceinline__ BsdfSample diffuse_sample(const RodentMaterial& m, const Surf& s, unsigned& rnd) {
    const float u = randf(rnd), v = randf(rnd);
    const DirSample d = sample_cosine_hemisphere(u, v);
    return make_bsdf_sample(s, mul(s.local, d.dir), d.pdf, d.dir.z, kd_of(m) * (1.0f / kPi), false);
}
__device__ __forceinline__ Col phong_eval(const RodentMaterial& m, const Surf& s, V3 in_dir, V3 out_dir) {
    const float c = positive_cos(in_dir, reflect(out_dir, s.local.c2));
    return ks_of(m) * (fastpow(c, m.ns) * (m.ns + 2.0f) * (1.0f / (2.0f * kPi)));
}
__device__ __forceinline__ float phong_pdf(const RodentMaterial& m, const Surf& s, V3 in_dir, V3 out_dir) {
    return cosine_power_hemisphere_pdf(positive_cos(in_dir, reflect(out_dir, s.local.c2)), m.ns);
}
__device__ __forceinline__ BsdfSample phong_sample(const RodentMaterial& m, const Surf& s, unsigned& rnd, V3 out_dir) {
    const V3 reflect_out = reflect(out_dir, s.local.c2);
    const float u = randf(rnd), v = randf(rnd);
    const DirSample d = sample_cosine_power_hemisphere(m.ns, u, v);
    const V3 in_dir = mul(orthonormal(reflect_out), d.dir);
    return make_bsdf_sample(s, in_dir, d.pdf, positive_cos(in_dir, s.local.c2), ks_of(m) * (d.pdf * (m.ns + 2.0f) / (m.ns + 1.0f)), false);
}
__device__ __forceinline__ float fresnel_factor(float k, float cos_i, float cos_t) {
    const float rs = (k * cos_i - cos_t) / (k * cos_i + cos_t), rp = (cos_i - k * cos_t) / (cos_i + k * cos_t);
    return (rs * rs + rp * rp) * 0.5f;
}

__device__ __forceinline__ Col bsdf_eval(const RodentMaterial& m, const Surf& s, V3 in_dir, V3 out_dir) {
    switch (m.bsdf) {
        case RODENT_BSDF_DIFFUSE: return diffuse_eval(m);
        case RODENT_BSDF_PHONG:   return phong_eval(m, s, in_dir, out_dir);
        case RODENT_BSDF_MIX:     return lerp(diffuse_eval(m), phong_eval(m, s, in_dir, out_dir), m.mix_k);
        default:                  return col(0, 0, 0);
    }
}
__device__ __forceinline__ float bsdf_pdf(const RodentMaterial& m, const Surf& s, V3 in_dir, V3 out_dir) {
    switch (m.bsdf) {
        case RODENT_BSDF_DIFFUSE: return diffuse_pdf(s, in_dir);
        case RODENT_BSDF_PHONG:   return phong_pdf(m, s, in_dir, out_dir);
        case RODENT_BSDF_MIX:     return lerp1(diffuse_pdf(s, in_dir), phong_pdf(m, s, in_dir, out_dir), m.mix_k);
        default:                  return 0.0f;
    }
}
__device__ __forceinline__ BsdfSample bsdf_sample(const RodentMaterial& m, const Surf& s, unsigned& rnd, V3 out_dir) {
    switch (m.bsdf) {
        case RODENT_BSDF_DIFFUSE: return diffuse_sample(m, s, rnd);
        case RODENT_BSDF_PHONG:   return phong_sample(m, s, rnd, out_dir);
        case RODENT_BSDF_MIX: {                                                       // material.impala:178-190
            BsdfSample b;
            if (randf(rnd) >= m.mix_k) {
                b = diffuse_sample(m, s, rnd);
                const float p = lerp1(b.pdf, phong_pdf(m, s, b.in_dir, out_dir), m.mix_k);
                b.color = lerp(b.color, phong_eval(m, s, b.in_dir, out_dir), m.mix_k);
                b.pdf = p;
            } else {
                b = phong_sample(m, s, rnd, out_dir);
                const float p = lerp1(diffuse_pdf(s, b.in_dir), b.pdf, m.mix_k);
                b.color = lerp(diffuse_eval(m), b.color, m.mix_k);
                b.pdf = p;
            }
            return b;
        }
        case RODENT_BSDF_MIRROR:
            return make_bsdf_sample(s, reflect(out_dir, s.local.c2), 1.0f, 1.0f, ks_of(m), false);
        case RODENT_BSDF_GLASS: {                                                     // material.impala:131-164 with n1 = 1, n2 = ni
            const float k = s.is_entering ? 1.0f / m.ni : m.ni / 1.0f;
            const V3 n = s.local.c2;
            const float cos_i = dot(out_dir, n);
            const float cos2_t = 1.0f - k * k * (1.0f - cos_i * cos_i);
            if (cos2_t > 0.0f) {
                const float cos_t = sqrtf(cos2_t);
                if (randf(rnd) > fresnel_factor(k, cos_i, cos_t))
                    return make_bsdf_sample(s, n * (k * cos_i - cos_t) - out_dir * k, 1.0f, 1.0f, col(m.tf[0], m.tf[1], m.tf[2]), true);
            }
            return make_bsdf_sample(s, reflect(out_dir, n), 1.0f, 1.0f, ks_of(m), false);
        }
        default:
            return BsdfSample{out_dir, 1.0f, 1.0f, col(0, 0, 0)};                      // make_black_bsdf
    }
}

}  // namespace shade
}  // namespace rb200
