/*
 * shading_bench_oracle.c -- CPU restatement of the reference's shading-interface micro-benchmark,
 * tools/bench_interface/bench_interface.impala:67-143.  TEST INFRASTRUCTURE ONLY (see traversal_oracle.c).
 *
 * Parity status: UNPINNED by the reference -- tools/bench_interface prints a throughput and checks no
 * value, there is no golden vector for it anywhere in the tree.  This file is therefore the definition
 * (IEEE binary32, no contraction, source operation order), and two known answers follow from the source
 * alone and are asserted in tests/test_shading_bench.py: a constant kd texture c gives c / pi for every
 * hit whatever the border / filter mode, and a checkerboard sampled at texel centres returns the texels.
 * `mesh` holds HOST pointers here.
 */
#include <math.h>
#include <stdint.h>

#include "../include/rodent_b200.h"

static inline float lerp1(float a, float b, float k) { return (1.0f - k) * a + k * b; }                 /* common.impala:118-120 */
static inline float lerp2(float a, float b, float c, float k1, float k2) { return (1.0f - k1 - k2) * a + k1 * b + k2 * c; }   /* :122-124 */

/* make_clamp_border / make_repeat_border, src/render/image.impala:41-55 */
static inline float border(uint32_t mode, float x) { return mode == 0u ? fminf(1.0f, fmaxf(0.0f, x)) : x - floorf(x); }
static inline int imin(int a, int b) { return a < b ? a : b; }

/* lookup_tex, bench_interface.impala:67-89; filters: src/render/image.impala:57-92 */
static Color lookup_tex(const Tex* tex, float u, float v) {
    if (tex->border == 0u || tex->border == 1u) {
        u = border(tex->border, u);
        v = border(tex->border, v);
    } else if (u < 0.0f || u > 1.0f || v < 0.0f || v > 1.0f) {
        return tex->border_color;
    }
    const float fu = u * (float)tex->width, fv = v * (float)tex->height;
    const int x0 = imin((int)fu, tex->width - 1), y0 = imin((int)fv, tex->height - 1);
    if (tex->sampler == 0u) return tex->pixels[x0 + y0 * tex->width];
    const int x1 = imin(x0 + 1, tex->width - 1), y1 = imin(y0 + 1, tex->height - 1);
    const float kx = fu - (float)(int)fu, ky = fv - (float)(int)fv;
    const Color p00 = tex->pixels[x0 + y0 * tex->width], p10 = tex->pixels[x1 + y0 * tex->width];
    const Color p01 = tex->pixels[x0 + y1 * tex->width], p11 = tex->pixels[x1 + y1 * tex->width];
    Color c;
    c.r = lerp1(lerp1(p00.r, p10.r, kx), lerp1(p01.r, p11.r, kx), ky);
    c.g = lerp1(lerp1(p00.g, p10.g, kx), lerp1(p01.g, p11.g, kx), ky);
    c.b = lerp1(lerp1(p00.b, p10.b, kx), lerp1(p01.b, p11.b, kx), ky);
    return c;
}

/* compute_shader_input (:91-123) + shade (:125-135).  The diffuse evaluation, kd * (1 / pi)
 * (src/render/material.impala:75-79), reads nothing of the input but kd; the rest is computed by the
 * reference and discarded, so only what reaches `colors` is restated. */
void oracle_bench_interface(const ShadedMesh* mesh, const TriHit* tri_hits, const Vec3* in_dirs, const Vec3* out_dirs,
                            Color* colors, int32_t n) {
    (void)in_dirs; (void)out_dirs;
    const float inv_pi = 1.0f / 3.14159265359f;                                                          /* common.impala:7 */
    for (int32_t i = 0; i < n; i++) {
        const int32_t id = tri_hits[i].id;
        const float u = tri_hits[i].uv.x, v = tri_hits[i].uv.y;
        const uint32_t i0 = mesh->indices[id * 4 + 0], i1 = mesh->indices[id * 4 + 1], i2 = mesh->indices[id * 4 + 2];
        const Vec2 t0 = mesh->texcoords[i0], t1 = mesh->texcoords[i1], t2 = mesh->texcoords[i2];
        const float tu = lerp2(t0.x, t1.x, t2.x, u, v), tv = lerp2(t0.y, t1.y, t2.y, u, v);              /* vec2_lerp2, vector.impala:89-94 */
        const Color kd = lookup_tex(&mesh->tex_kd, tu, tv);
        colors[i].r = kd.r * inv_pi; colors[i].g = kd.g * inv_pi; colors[i].b = kd.b * inv_pi;           /* color_mulf */
    }
}
