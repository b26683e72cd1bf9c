/*
 * render_oracle.c -- CPU restatement of the reference's path tracer.  TEST
 * INFRASTRUCTURE ONLY (see traversal_oracle.c): it checks the CUDA wavefront renderer
 * and is timed as the CPU baseline of the path-tracing metric.
 *
 * What is restated: make_path_tracing_renderer (src/render/renderer.impala:62-162) with
 * the camera emitter (:26-40), surface element (src/render/geometry.impala:21-53), BSDFs
 * (src/render/material.impala:65-192), triangle lights (src/render/light.impala:122-154),
 * RNG / sampling (src/core/random.impala), fastpow (src/core/common.impala:42-61),
 * pinhole camera (src/render/camera.impala:29-44), film accumulation (+= colour / spp,
 * src/render/mapping_gpu.impala:32-45) -- driven per camera sample instead of per
 * wavefront.  That is equivalent: every sample's random stream is a pure function of
 * (sample, iter, x, y) (renderer.impala:28-33), so the schedule (CPU tiles of
 * src/render/mapping_cpu.impala:352-473, GPU wavefronts of mapping_gpu.impala:308-369, or
 * this depth-first loop) does not change any sample's value.
 *
 * Rays are traced with the single-ray BVH8 kernel of traversal_oracle.c (closest hit for
 * path vertices, any hit for shadow rays, as mapping_gpu.impala:18-30,47-80 do).
 *
 * Parity status: PINNED statistically by testing/ref-cornell.png
 * (tests/test_render_oracle.py).  Bit parity with the reference is not defined for this
 * path: it calls libm sinf/cosf/sqrtf and is compiled -ffast-math.
 */
#ifdef ORACLE_DEBUG_PIXEL   /* developer aid: -DORACLE_DEBUG_PIXEL, then ORACLE_DEBUG_X / _Y print that pixel's closest-hit rays */
#include <stdio.h>
#endif
#include <math.h>

#include "traversal_oracle.c"

typedef struct { float x, y, z; } V3;
typedef struct { float r, g, b; } Col;
typedef struct { V3 c0, c1, c2; } M3;

static const float kPi = 3.14159265359f;          /* src/core/common.impala:7 */
static const float kOffset = 0.001f;              /* renderer.impala:64 */

static inline V3 v3(float x, float y, float z) { V3 v = {x, y, z}; return v; }
static inline V3 vadd(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 vsub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 vmulf(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
static inline V3 vneg(V3 a) { return v3(-a.x, -a.y, -a.z); }
static inline float vdot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float vlen(V3 a) { return sqrtf(vdot(a, a)); }
static inline V3 vnormalize(V3 a) { return vmulf(a, 1.0f / vlen(a)); }                         /* vector.impala:82 */
static inline V3 vreflect(V3 v, V3 n) { return vsub(vmulf(n, 2.0f * vdot(n, v)), v); }         /* vector.impala:74 */
static inline float lerp1(float a, float b, float k) { return (1.0f - k) * a + k * b; }         /* common.impala:118-120 */
static inline float lerp2(float a, float b, float c, float k1, float k2) { return (1.0f - k1 - k2) * a + k1 * b + k2 * c; }
static inline Col col(float r, float g, float b) { Col c = {r, g, b}; return c; }
static inline Col cmul(Col a, Col b) { return col(a.r * b.r, a.g * b.g, a.b * b.b); }
static inline Col cmulf(Col a, float f) { return col(a.r * f, a.g * f, a.b * f); }
static inline Col clerp(Col a, Col b, float t) { return col(lerp1(a.r, b.r, t), lerp1(a.g, b.g, t), lerp1(a.b, b.b, t)); }
static inline float luminance(Col c) { return c.r * 0.2126f + c.g * 0.7152f + c.b * 0.0722f; } /* color.impala:33-35 */
static const Col kBlack = {0.0f, 0.0f, 0.0f};

/* matrix.impala:29-39, 107-111 */
static inline M3 orthonormal(V3 n) {
    const float sign = n.z >= 0.0f ? 1.0f : -1.0f;
    const float a = -1.0f / (sign + n.z);
    const float b = n.x * n.y * a;
    M3 m;
    m.c0 = v3(1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x);
    m.c1 = v3(b, sign + n.y * n.y * a, -n.y);
    m.c2 = n;
    return m;
}
static inline V3 m3mul(M3 m, V3 v) {
    return v3(vdot(v3(m.c0.x, m.c1.x, m.c2.x), v), vdot(v3(m.c0.y, m.c1.y, m.c2.y), v), vdot(v3(m.c0.z, m.c1.z, m.c2.z), v));
}

/* random.impala:7-31, 116-126 */
static inline int32_t xorshift(uint32_t* seed) {
    uint32_t x = *seed;
    x = x == 0u ? 1u : x;
    x ^= x << 13; x ^= x >> 17; x ^= x << 5;
    *seed = x;
    return (int32_t)x;
}
static inline float randf(uint32_t* rnd) {
    const uint32_t x = (uint32_t)xorshift(rnd);
    return i2f((int32_t)((127u << 23) | (x & 0x7FFFFFu))) - 1.0f;
}
static inline uint32_t fnv_hash(uint32_t h, uint32_t d) {
    h = (h * 16777619u) ^ (d & 0xFFu);
    h = (h * 16777619u) ^ ((d >> 8) & 0xFFu);
    h = (h * 16777619u) ^ ((d >> 16) & 0xFFu);
    h = (h * 16777619u) ^ ((d >> 24) & 0xFFu);
    return h;
}

/* common.impala:42-61 */
static inline float fastlog2(float x) {
    const uint32_t vx = (uint32_t)f2i(x);
    const uint32_t mx = (vx & 0x007FFFFFu) | 0x3f000000u;
    const float y = (float)vx * 1.1920928955078125e-7f;
    const float z = i2f((int32_t)mx);
    return y - 124.22551499f - 1.498030302f * z - 1.72587999f / (0.3520887068f + z);
}
static inline float fastpow2(float p) {
    const float offset = p < 0.0f ? 1.0f : 0.0f;
    const float clipp = p < -126.0f ? -126.0f : p;
    const int32_t w = (int32_t)clipp;
    const float z = clipp - (float)w + offset;
    const int32_t v = (int32_t)((float)(1u << 23) * (clipp + 121.2740575f + 27.7280233f / (4.84252568f - z) - 1.49012907f * z));
    return i2f(v);
}
static inline float fastpow(float x, float p) { return fastpow2(p * fastlog2(x)); }

static inline float positive_cos(V3 a, V3 b) { const float c = vdot(a, b); return c >= 0.0f ? c : 0.0f; }   /* common.impala:133-136 */
static inline float cosine_hemisphere_pdf(float c) { return c * (1.0f / kPi); }                             /* random.impala:69 */
static inline float cosine_power_hemisphere_pdf(float c, float k) {                                       /* random.impala:80-82 */
    return fastpow(c, k) * (k + 1.0f) * (1.0f / (2.0f * kPi));
}

typedef struct { V3 dir; float pdf; } DirSample;
/* Test switch: sin / cos from the polynomial both sides share (rodent_b200/csrc/poly_trig.h) instead of libm's; with it
 * the device's films agree with this oracle's up to the order of the atomic adds (tests/test_gpu_render.py). */
#include "../rodent_b200/csrc/poly_trig.h"
static int g_poly_trig = 0;
void oracle_set_poly_trig(int on) { g_poly_trig = on; }
static inline DirSample make_dir_sample(float c, float s, float phi, float pdf) {                          /* random.impala:39-48 */
    float sn, cs;
    if (g_poly_trig) rb_poly_sincos(phi, &sn, &cs);
    else { sn = sinf(phi); cs = cosf(phi); }
    DirSample d; d.dir = v3(s * cs, s * sn, c); d.pdf = pdf; return d;
}
static inline DirSample sample_cosine_hemisphere(float u, float v) {                                       /* random.impala:72-77 */
    const float c = sqrtf(1.0f - v), s = sqrtf(v), phi = 2.0f * kPi * u;
    return make_dir_sample(c, s, phi, cosine_hemisphere_pdf(c));
}
static inline DirSample sample_cosine_power_hemisphere(float k, float u, float v) {                        /* random.impala:85-98 */
    const float c = fminf(fastpow(v, 1.0f / (k + 1.0f)), 1.0f);
    const float s = sqrtf(1.0f - c * c), phi = 2.0f * kPi * u;
    const float pow_c_k = c != 0.0f ? v / c : 0.0f;
    return make_dir_sample(c, s, phi, pow_c_k * (k + 1.0f) * (1.0f / (2.0f * kPi)));
}

/* ---- surface element, geometry.impala:21-53 ------------------------------------------ */
typedef struct {
    int is_entering;
    V3 point, face_normal;
    M3 local;
} Surf;

static inline V3 load3(const float* base, int i) { return v3(base[4 * i], base[4 * i + 1], base[4 * i + 2]); }

static Surf surface_element(const RodentSceneView* sc, V3 org, V3 dir, int prim, float t, float u, float v) {
    const int32_t* idx = sc->indices + 4 * prim;
    const V3 fn = load3(sc->face_normals, prim);
    const V3 n0 = load3(sc->normals, idx[0]), n1 = load3(sc->normals, idx[1]), n2 = load3(sc->normals, idx[2]);
    const V3 normal = vnormalize(v3(lerp2(n0.x, n1.x, n2.x, u, v), lerp2(n0.y, n1.y, n2.y, u, v), lerp2(n0.z, n1.z, n2.z, u, v)));
    Surf s;
    s.is_entering = vdot(dir, fn) <= 0.0f;
    s.point = vadd(org, vmulf(dir, t));
    s.face_normal = s.is_entering ? fn : vneg(fn);
    s.local = orthonormal(vdot(dir, normal) <= 0.0f ? normal : vneg(normal));
    return s;
}

/* ---- BSDFs, material.impala:54-192 ------------------------------------------------------ */
typedef struct { V3 in_dir; float pdf, cos; Col color; } BsdfSample;

static inline BsdfSample make_bsdf_sample(const Surf* s, V3 in_dir, float pdf, float cosv, Col color, int inverted) {
    const int valid = (pdf > 0.0f) && (inverted ^ (vdot(in_dir, s->face_normal) > 0.0f));
    BsdfSample b; b.in_dir = in_dir; b.pdf = valid ? pdf : 1.0f; b.cos = cosv; b.color = valid ? color : kBlack;
    return b;
}
static inline Col mkd(const RodentMaterial* m) { return col(m->kd[0], m->kd[1], m->kd[2]); }
static inline Col mks(const RodentMaterial* m) { return col(m->ks[0], m->ks[1], m->ks[2]); }

static Col diffuse_eval(const RodentMaterial* m) { return cmulf(mkd(m), 1.0f / kPi); }
static float diffuse_pdf(const Surf* s, V3 in_dir) { return cosine_hemisphere_pdf(positive_cos(in_dir, s->local.c2)); }
static BsdfSample diffuse_sample(const RodentMaterial* m, const Surf* s, uint32_t* rnd) {
    const float u = randf(rnd), v = randf(rnd);
    const DirSample d = sample_cosine_hemisphere(u, v);
    return make_bsdf_sample(s, m3mul(s->local, d.dir), d.pdf, d.dir.z, cmulf(mkd(m), 1.0f / kPi), 0);
}
static Col phong_eval(const RodentMaterial* m, const Surf* s, V3 in_dir, V3 out_dir) {
    const float c = positive_cos(in_dir, vreflect(out_dir, s->local.c2));
    return cmulf(mks(m), fastpow(c, m->ns) * (m->ns + 2.0f) * (1.0f / (2.0f * kPi)));
}
static float phong_pdf(const RodentMaterial* m, const Surf* s, V3 in_dir, V3 out_dir) {
    return cosine_power_hemisphere_pdf(positive_cos(in_dir, vreflect(out_dir, s->local.c2)), m->ns);
}
static BsdfSample phong_sample(const RodentMaterial* m, const Surf* s, uint32_t* rnd, V3 out_dir) {
    const V3 reflect_out = vreflect(out_dir, s->local.c2);
    const float u = randf(rnd), v = randf(rnd);
    const DirSample d = sample_cosine_power_hemisphere(m->ns, u, v);
    const V3 in_dir = m3mul(orthonormal(reflect_out), d.dir);
    const float c = positive_cos(in_dir, s->local.c2);
    return make_bsdf_sample(s, in_dir, d.pdf, c, cmulf(mks(m), d.pdf * (m->ns + 2.0f) / (m->ns + 1.0f)), 0);
}
static inline float fresnel_factor(float k, float cos_i, float cos_t) {
    const float rs = (k * cos_i - cos_t) / (k * cos_i + cos_t), rp = (cos_i - k * cos_t) / (cos_i + k * cos_t);
    return (rs * rs + rp * rp) * 0.5f;
}

static int bsdf_is_specular(const RodentMaterial* m) { return m->bsdf == RODENT_BSDF_MIRROR || m->bsdf == RODENT_BSDF_GLASS; }

static Col bsdf_eval(const RodentMaterial* m, const Surf* s, V3 in_dir, V3 out_dir) {
    switch (m->bsdf) {
        case RODENT_BSDF_DIFFUSE: return diffuse_eval(m);
        case RODENT_BSDF_PHONG:   return phong_eval(m, s, in_dir, out_dir);
        case RODENT_BSDF_MIX:     return clerp(diffuse_eval(m), phong_eval(m, s, in_dir, out_dir), m->mix_k);
        default:                  return kBlack;
    }
}
static float bsdf_pdf(const RodentMaterial* m, const Surf* s, V3 in_dir, V3 out_dir) {
    switch (m->bsdf) {
        case RODENT_BSDF_DIFFUSE: return diffuse_pdf(s, in_dir);
        case RODENT_BSDF_PHONG:   return phong_pdf(m, s, in_dir, out_dir);
        case RODENT_BSDF_MIX:     return lerp1(diffuse_pdf(s, in_dir), phong_pdf(m, s, in_dir, out_dir), m->mix_k);
        default:                  return 0.0f;
    }
}
static BsdfSample bsdf_sample(const RodentMaterial* m, const Surf* s, uint32_t* rnd, V3 out_dir) {
    switch (m->bsdf) {
        case RODENT_BSDF_DIFFUSE: return diffuse_sample(m, s, rnd);
        case RODENT_BSDF_PHONG:   return phong_sample(m, s, rnd, out_dir);
        case RODENT_BSDF_MIX: {                                                         /* material.impala:178-190 */
            BsdfSample b;
            if (randf(rnd) >= m->mix_k) {
                b = diffuse_sample(m, s, rnd);
                const float p = lerp1(b.pdf, phong_pdf(m, s, b.in_dir, out_dir), m->mix_k);
                b.color = clerp(b.color, phong_eval(m, s, b.in_dir, out_dir), m->mix_k);
                b.pdf = p;
            } else {
                b = phong_sample(m, s, rnd, out_dir);
                const float p = lerp1(diffuse_pdf(s, b.in_dir), b.pdf, m->mix_k);
                b.color = clerp(diffuse_eval(m), b.color, m->mix_k);
                b.pdf = p;
            }
            return b;
        }
        case RODENT_BSDF_MIRROR:
            return make_bsdf_sample(s, vreflect(out_dir, s->local.c2), 1.0f, 1.0f, mks(m), 0);
        case RODENT_BSDF_GLASS: {                                                       /* material.impala:131-164, n1 = 1, n2 = ni */
            const float k = s->is_entering ? 1.0f / m->ni : m->ni / 1.0f;
            const V3 n = s->local.c2;
            const float cos_i = vdot(out_dir, n);
            const float cos2_t = 1.0f - k * k * (1.0f - cos_i * cos_i);
            if (cos2_t > 0.0f) {
                const float cos_t = sqrtf(cos2_t);
                const float F = fresnel_factor(k, cos_i, cos_t);
                if (randf(rnd) > F) {
                    const V3 t = vsub(vmulf(n, k * cos_i - cos_t), vmulf(out_dir, k));
                    return make_bsdf_sample(s, t, 1.0f, 1.0f, col(m->tf[0], m->tf[1], m->tf[2]), 1);
                }
            }
            return make_bsdf_sample(s, vreflect(out_dir, n), 1.0f, 1.0f, mks(m), 0);
        }
        default: {                                                                      /* make_black_bsdf */
            BsdfSample b; b.in_dir = out_dir; b.pdf = 1.0f; b.cos = 1.0f; b.color = kBlack; return b;
        }
    }
}

/* ---- the renderer ------------------------------------------------------------------------ */
typedef struct OracleRenderStats { uint64_t samples, primary_rays, shadow_rays; OracleStats trav; } OracleRenderStats;

typedef struct {
    const RodentSceneView* sc; const Settings* cam;
    int width, height, spp, max_path_len, iter;
    float* film;
    int y0, y1;
    OracleRenderStats stats;
} RenderJob;

void oracle_bvh2_trace_one(int any_hit, const Node2* nodes, const Tri1* tris, const Ray1* ray, Hit1* hit, int32_t* geom, uint64_t* counters);   /* traversal_bvh2_oracle.c */

/* Through the scene's BVH2 / Tri1 when it carries one -- the reference GPU device's layout and traversal
 * (gpu_traverse_primary / gpu_traverse_secondary, mapping_gpu.impala:18-80) -- otherwise the BVH8 single-ray kernel: the
 * same choice as the CUDA render loop (rodent_b200/csrc/render.cu). */
static inline void trace(const RodentSceneView* sc, int any, V3 org, V3 dir, float tmin, float tmax, Hit1* hit, int32_t* geom, OracleStats* st) {
    Ray1 r = {{org.x, org.y, org.z}, tmin, {dir.x, dir.y, dir.z}, tmax};
    if (sc->nodes2) {                     /* the work counters then count Node2 visits and Tri1 tests */
        uint64_t c[2] = {0, 0};
        oracle_bvh2_trace_one(any, sc->nodes2, sc->tris1, &r, hit, geom, st ? c : NULL);
        if (st) { st->nodes += c[0]; st->tri4 += c[1]; }
        return;
    }
    if (any) traverse_single(8, 1, sc->nodes, sc->tris, &r, hit, st, geom);
    else     traverse_single(8, 0, sc->nodes, sc->tris, &r, hit, st, geom);
}

/* make_texture(repeat border, bilinear filter, make_image_rgba32), src/render/image.impala:24-92 */
static Col rgba32_texture(const uint32_t* pixels, int width, int height, float u, float v) {   /* image.impala:24-92: repeat border, bilinear */
    u = u - floorf(u); v = v - floorf(v);
    const float fu = u * (float)width, fv = v * (float)height;
    int x0 = (int)fu; if (x0 > width - 1) x0 = width - 1;
    int y0 = (int)fv; if (y0 > height - 1) y0 = height - 1;
    const int x1 = x0 + 1 < width - 1 ? x0 + 1 : width - 1, y1 = y0 + 1 < height - 1 ? y0 + 1 : height - 1;
    const float kx = fu - (float)(int)fu, ky = fv - (float)(int)fv;
    Col p[4];
    const int xs[4] = {x0, x1, x0, x1}, ys[4] = {y0, y0, y1, y1};
    for (int k = 0; k < 4; k++) {
        const uint32_t q = pixels[ys[k] * width + xs[k]];
        p[k] = col((float)(q & 0xFFu) * (1.0f / 255.0f), (float)((q >> 8) & 0xFFu) * (1.0f / 255.0f), (float)((q >> 16) & 0xFFu) * (1.0f / 255.0f));
    }
    return col(lerp1(lerp1(p[0].r, p[1].r, kx), lerp1(p[2].r, p[3].r, kx), ky),
               lerp1(lerp1(p[0].g, p[1].g, kx), lerp1(p[2].g, p[3].g, kx), ky),
               lerp1(lerp1(p[0].b, p[1].b, kx), lerp1(p[2].b, p[3].b, kx), ky));
}

/* The textured part of a material's shader, converter.cpp:876-903: kd / ks sampled at the hit's texture coordinates
 * (surf.attr(0), geometry.impala:44-47), the mix weight recomputed from the sampled colours. */
static RodentMaterial textured_material(const RodentSceneView* sc, int geom, int prim, float u, float v) {
    RodentMaterial m = sc->materials[geom];
    if ((m.map_kd | m.map_ks | m.map_ke) == 0) return m;
    const int i0 = sc->indices[prim * 4], i1 = sc->indices[prim * 4 + 1], i2 = sc->indices[prim * 4 + 2];
    const float* tc = sc->texcoords;
    const float tu = lerp2(tc[i0 * 4], tc[i1 * 4], tc[i2 * 4], u, v), tv = lerp2(tc[i0 * 4 + 1], tc[i1 * 4 + 1], tc[i2 * 4 + 1], u, v);
    if (m.map_kd) {
        const RodentTexture* t = &sc->textures[m.map_kd - 1];
        const Col c = rgba32_texture(sc->texture_pixels + t->offset, t->width, t->height, tu, tv);
        m.kd[0] = c.r; m.kd[1] = c.g; m.kd[2] = c.b;
    }
    if (m.map_ks) {
        const RodentTexture* t = &sc->textures[m.map_ks - 1];
        const Col c = rgba32_texture(sc->texture_pixels + t->offset, t->width, t->height, tu, tv);
        m.ks[0] = c.r; m.ks[1] = c.g; m.ks[2] = c.b;
    }
    if (m.map_ke) {                                              /* the light's colour at this point, converter.cpp:794-801 */
        const RodentTexture* t = &sc->textures[m.map_ke - 1];
        const Col c = rgba32_texture(sc->texture_pixels + t->offset, t->width, t->height, tu, tv);
        m.ke[0] = c.r; m.ke[1] = c.g; m.ke[2] = c.b;
    }
    if (m.bsdf == RODENT_BSDF_MIX) {
        const float lum_ks = luminance(col(m.ks[0], m.ks[1], m.ks[2])), lum_kd = luminance(col(m.kd[0], m.kd[1], m.kd[2]));
        m.mix_k = (lum_ks + lum_kd == 0.0f) ? 0.0f : lum_ks / (lum_ks + lum_kd);
    }
    return m;
}

/* Radiance a light emits from the point with barycentrics (u, v): its constant colour, or its material's map_Ke there. */
static Col light_color(const RodentSceneView* sc, const RodentLight* l, float u, float v) {
    if (l->map_ke == 0) return col(l->color[0], l->color[1], l->color[2]);
    const int i0 = sc->indices[l->prim * 4], i1 = sc->indices[l->prim * 4 + 1], i2 = sc->indices[l->prim * 4 + 2];
    const float* tc = sc->texcoords;
    const RodentTexture* t = &sc->textures[l->map_ke - 1];
    return rgba32_texture(sc->texture_pixels + t->offset, t->width, t->height,
                          lerp2(tc[i0 * 4], tc[i1 * 4], tc[i2 * 4], u, v), lerp2(tc[i0 * 4 + 1], tc[i1 * 4 + 1], tc[i2 * 4 + 1], u, v));
}

static void render_rows(RenderJob* job) {
    const RodentSceneView* sc = job->sc;
    const Settings* cam = job->cam;
    const V3 eye = v3(cam->eye.x, cam->eye.y, cam->eye.z), cdir = v3(cam->dir.x, cam->dir.y, cam->dir.z);
    const V3 up = v3(cam->up.x, cam->up.y, cam->up.z), right = v3(cam->right.x, cam->right.y, cam->right.z);
    const float pdf_lightpick = 1.0f / (float)sc->num_lights;                       /* renderer.impala:65 */
    const float inv_spp = 1.0f / (float)job->spp;                                   /* mapping_gpu.impala:40 */
    for (int y = job->y0; y < job->y1; y++)
        for (int x = 0; x < job->width; x++)
            for (int sample = 0; sample < job->spp; sample++) {
                /* make_camera_emitter, renderer.impala:26-40 ; camera.impala:35-44 */
                uint32_t rnd = fnv_hash(fnv_hash(fnv_hash(fnv_hash(0x811C9DC5u, (uint32_t)sample), (uint32_t)job->iter), (uint32_t)x), (uint32_t)y);
                const float kx = 2.0f * ((float)x + randf(&rnd)) / (float)job->width - 1.0f;
                const float ky = 1.0f - 2.0f * ((float)y + randf(&rnd)) / (float)job->height;
                V3 org = eye;
                V3 dir = vnormalize(vadd(vadd(vmulf(right, cam->width * kx), vmulf(up, cam->height * ky)), cdir));
                float tmin = 0.0f;
                Col contrib = col(1.0f, 1.0f, 1.0f), pixel = kBlack;
                float mis = 0.0f;
                int depth = 0;
                job->stats.samples++;
                for (;;) {
                    Hit1 hit; int32_t geom;
                    trace(sc, 0, org, dir, tmin, FLT_MAX_, &hit, &geom, &job->stats.trav);
                    job->stats.primary_rays++;
#ifdef ORACLE_DEBUG_PIXEL
                    if (getenv("ORACLE_DEBUG_X") && x == atoi(getenv("ORACLE_DEBUG_X")) && y == atoi(getenv("ORACLE_DEBUG_Y")))
                        fprintf(stderr, "RAY %d %d %d %08x %08x %08x %08x %08x %08x %08x -> %d %g geom %d\n", sample, depth,
                                0, f2i(org.x), f2i(org.y), f2i(org.z), f2i(tmin), f2i(dir.x), f2i(dir.y), f2i(dir.z), hit.tri_id, hit.t, geom);
#endif
                    if (hit.tri_id < 0) break;                                       /* misses leave the stream, mapping_gpu.impala:357 */
                    const RodentMaterial textured = textured_material(sc, geom, hit.tri_id, hit.u, hit.v);
                    const RodentMaterial* mat = &textured;
                    const Surf surf = surface_element(sc, org, dir, hit.tri_id, hit.t, hit.u, hit.v);
                    const V3 out_dir = vneg(dir);

                    /* on_hit, renderer.impala:111-127 */
                    if (mat->is_emissive && surf.is_entering) {
                        const RodentLight* l = &sc->lights[sc->light_ids[hit.tri_id]];
                        const float pdf_dir = cosine_hemisphere_pdf(vdot(v3(l->n[0], l->n[1], l->n[2]), out_dir));
                        Col intensity = kBlack; float pdf_area = 1.0f;                /* make_emission_value, light.impala:94-108 */
                        if (pdf_dir > 0.0f) { intensity = col(mat->ke[0], mat->ke[1], mat->ke[2]); pdf_area = l->inv_area; }   /* = the light's colour */
                        const float next_mis = mis * hit.t * hit.t / vdot(out_dir, surf.local.c2);
                        const float w = 1.0f / (1.0f + next_mis * pdf_lightpick * pdf_area);
                        const Col c = cmulf(cmul(contrib, intensity), w);
                        pixel.r += c.r * inv_spp; pixel.g += c.g * inv_spp; pixel.b += c.b * inv_spp;
                    }

                    /* on_shadow, renderer.impala:69-109 */
                    if (!bsdf_is_specular(mat) && sc->num_lights > 0) {
                        const int light_id = (xorshift(&rnd) & 0x7FFFFFFF) % sc->num_lights;
                        const RodentLight* l = &sc->lights[light_id];
                        /* make_area_light.sample_direct + sample_triangle, light.impala:122-128, random.impala:51-61 */
                        float u = randf(&rnd), v = randf(&rnd);
                        if (u + v > 1.0f) { u = 1.0f - u; v = 1.0f - v; }
                        const V3 lv0 = v3(l->v0[0], l->v0[1], l->v0[2]), lv1 = v3(l->v1[0], l->v1[1], l->v1[2]), lv2 = v3(l->v2[0], l->v2[1], l->v2[2]);
                        const V3 pos = vadd(vadd(vmulf(lv0, 1.0f - v - u), vmulf(lv1, u)), vmulf(lv2, v));
                        const V3 ln = v3(l->n[0], l->n[1], l->n[2]);
                        const V3 from_dir = vsub(surf.point, pos);
                        float cos_l = vdot(from_dir, ln) / vlen(from_dir);
                        Col intensity = light_color(sc, l, u, v);
                        float pdf_area = l->inv_area;
                        if (!(pdf_area > 0.0f && cosine_hemisphere_pdf(cos_l) > 0.0f && cos_l > 0.0f)) {   /* make_direct_sample */
                            intensity = kBlack; pdf_area = 1.0f; cos_l = 0.0f;
                        }
                        const V3 light_dir = vsub(pos, surf.point);
                        const float vis = vdot(light_dir, surf.local.c2);
                        if (vis > 0.0f && cos_l > 0.0f) {
                            const float inv_d = 1.0f / vlen(light_dir), inv_d2 = inv_d * inv_d;
                            const V3 in_dir = vmulf(light_dir, inv_d);
                            const float pdf_e = bsdf_pdf(mat, &surf, in_dir, out_dir);          /* has_area */
                            const float pdf_l = pdf_area * pdf_lightpick, inv_pdf_l = 1.0f / pdf_l;
                            const float cos_e = vis * inv_d;
                            const float w = 1.0f / (1.0f + pdf_e * cos_l * inv_d2 * inv_pdf_l);
                            const float geom_factor = cos_e * cos_l * inv_d2 * inv_pdf_l;
                            const Col c = cmulf(cmul(intensity, cmul(contrib, bsdf_eval(mat, &surf, in_dir, out_dir))), geom_factor * w);
                            Hit1 sh; int32_t sg;
                            trace(sc, 1, surf.point, light_dir, kOffset, 1.0f - kOffset, &sh, &sg, &job->stats.trav);
                            job->stats.shadow_rays++;
                            if (sh.tri_id < 0) { pixel.r += c.r * inv_spp; pixel.g += c.g * inv_spp; pixel.b += c.b * inv_spp; }
                        }
                    }

                    /* on_bounce, renderer.impala:129-152 */
                    float rr = 2.0f * luminance(contrib);                              /* russian_roulette, random.impala:128-131 */
                    if (rr > 0.75f) rr = 0.75f;
                    if (depth >= job->max_path_len || randf(&rnd) >= rr) break;
                    const BsdfSample bs = bsdf_sample(mat, &surf, &rnd, out_dir);
                    contrib = cmulf(cmul(contrib, bs.color), bs.cos / (bs.pdf * rr));
                    mis = bsdf_is_specular(mat) ? 0.0f : 1.0f / bs.pdf;
                    org = surf.point; dir = bs.in_dir; tmin = kOffset;
                    depth++;
                }
                float* px = job->film + 3 * ((size_t)y * job->width + x);
                px[0] += pixel.r; px[1] += pixel.g; px[2] += pixel.b;
            }
}
typedef struct { RenderJob* jobs; int first, step, n; } RenderWorker;
static void* render_worker(void* p) {
    RenderWorker* w = (RenderWorker*)p;
    for (int k = w->first; k < w->n; k += w->step) render_rows(&w->jobs[k]);
    return NULL;
}

/* One render(settings, iter) call: film += (sum over spp samples) / spp.  One job per image
 * row, rows dealt round-robin to `threads` host threads (balanced; rows never share pixels). */
void oracle_render(const RodentSceneView* sc, const Settings* cam, int width, int height, int spp, int max_path_len,
                   int iter, float* film, int threads, OracleRenderStats* stats) {
    pthread_once(&g_net_once, init_networks);
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    RenderJob* jobs = (RenderJob*)calloc((size_t)height, sizeof(RenderJob));
    for (int y = 0; y < height; y++) {
        RenderJob* j = &jobs[y];
        j->sc = sc; j->cam = cam; j->width = width; j->height = height; j->spp = spp; j->max_path_len = max_path_len;
        j->iter = iter; j->film = film; j->y0 = y; j->y1 = y + 1;
    }
    RenderWorker workers[256]; pthread_t th[256];
    for (int t = 0; t < threads; t++) { workers[t].jobs = jobs; workers[t].first = t; workers[t].step = threads; workers[t].n = height; }
    if (threads == 1) render_worker(&workers[0]);
    else {
        for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, render_worker, &workers[t]);
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    }
    if (stats) {
        memset(stats, 0, sizeof *stats);
        for (int y = 0; y < height; y++) {
            stats->samples += jobs[y].stats.samples; stats->primary_rays += jobs[y].stats.primary_rays;
            stats->shadow_rays += jobs[y].stats.shadow_rays;
            stats->trav.nodes += jobs[y].stats.trav.nodes; stats->trav.tri4 += jobs[y].stats.trav.tri4;
        }
    }
    free(jobs);
}

/* ---- cpu_bench_shading, tools/bench_shading/bench_shading.impala:22-104 ------------------------------------
 * CPU restatement (same signature, host pointers).  Parity status: UNPINNED by the reference -- the tool prints a
 * throughput and checks no value.  Known answers that follow from the source alone are asserted in
 * tests/test_shading_bench.py (constant-colour geometry 0: a pure function of the BSDF sample; depth + 1; tmin = offset).
 * The reference runs it with FTZ/DAZ set (bench_shading.cpp:57-59); no denormal occurs on its inputs.
 * Its texture is rgba32_texture above. */
void oracle_bench_shading(const PrimaryStream* in, PrimaryStream* out, const Vec3* vertices, const Vec3* normals,
                          const Vec3* face_normals, const Vec2* texcoords, const int32_t* indices, const uint32_t* pixels,
                          int32_t width, int32_t height, const int32_t* begins, const int32_t* ends, int32_t num_tris, int32_t num_iters) {
    (void)vertices; (void)num_tris;
    for (int iter = 0; iter < num_iters; iter++)
        for (int geom_id = 0; geom_id < 4; geom_id++)                                   /* iterate_rays, :7-20 */
            for (int i = begins[geom_id]; i < ends[geom_id]; i++) {
                const V3 org = v3(in->rays.org_x[i], in->rays.org_y[i], in->rays.org_z[i]);
                const V3 dir = v3(in->rays.dir_x[i], in->rays.dir_y[i], in->rays.dir_z[i]);
                const int prim = in->prim_id[i];
                const float t = in->t[i], hu = in->u[i], hv = in->v[i];
                uint32_t rnd = in->rnd[i];
                /* surface_element, geometry.impala:21-53 */
                const int i0 = indices[prim * 4], i1 = indices[prim * 4 + 1], i2 = indices[prim * 4 + 2];
                const V3 fn = v3(face_normals[prim].x, face_normals[prim].y, face_normals[prim].z);
                const V3 normal = vnormalize(v3(lerp2(normals[i0].x, normals[i1].x, normals[i2].x, hu, hv),
                                                lerp2(normals[i0].y, normals[i1].y, normals[i2].y, hu, hv),
                                                lerp2(normals[i0].z, normals[i1].z, normals[i2].z, hu, hv)));
                Surf surf;
                surf.is_entering = vdot(dir, fn) <= 0.0f;
                surf.point = vadd(org, vmulf(dir, t));
                surf.face_normal = surf.is_entering ? fn : vneg(fn);
                surf.local = orthonormal(vdot(dir, normal) <= 0.0f ? normal : vneg(normal));
                const float tu = lerp2(texcoords[i0].x, texcoords[i1].x, texcoords[i2].x, hu, hv);
                const float tv = lerp2(texcoords[i0].y, texcoords[i1].y, texcoords[i2].y, hu, hv);
                /* shader, :45-62 */
                const Col tex = rgba32_texture(pixels, width, height, tu, tv);
                const Col kd = (geom_id & 1) == 0 ? col(0.0f, 1.0f, 0.0f) : tex;
                const Col ks = (geom_id & 2) == 0 ? col(0.0f, 1.0f, 0.0f) : tex;
                RodentMaterial mat;
                memset(&mat, 0, sizeof mat);
                mat.bsdf = RODENT_BSDF_MIX; mat.ns = (geom_id & 2) == 0 ? 96.0f : 12.0f; mat.ni = 1.0f;
                mat.kd[0] = kd.r; mat.kd[1] = kd.g; mat.kd[2] = kd.b; mat.ks[0] = ks.r; mat.ks[1] = ks.g; mat.ks[2] = ks.b;
                const float lum_ks = luminance(ks), lum_kd = luminance(kd);
                mat.mix_k = (lum_ks + lum_kd == 0.0f) ? 0.0f : lum_ks / (lum_ks + lum_kd);
                /* :84-98 */
                const V3 out_dir = vneg(dir);
                const BsdfSample smp = bsdf_sample(&mat, &surf, &rnd, out_dir);
                const Col contrib = cmulf(cmul(col(in->contrib_r[i], in->contrib_g[i], in->contrib_b[i]), smp.color), smp.cos / smp.pdf);
                out->rays.org_x[i] = surf.point.x; out->rays.org_y[i] = surf.point.y; out->rays.org_z[i] = surf.point.z;
                out->rays.dir_x[i] = smp.in_dir.x; out->rays.dir_y[i] = smp.in_dir.y; out->rays.dir_z[i] = smp.in_dir.z;
                out->rays.tmin[i] = 0.0001f; out->rays.tmax[i] = 3.4028234664e+38f;
                out->rnd[i] = rnd; out->contrib_r[i] = contrib.r; out->contrib_g[i] = contrib.g; out->contrib_b[i] = contrib.b;
                out->mis[i] = 1.0f / smp.pdf; out->depth[i] = in->depth[i] + 1;
            }
}
