/* sin and cos of an angle in [0, 2 pi] as fixed polynomials: plain C, the same operations in the same order wherever
 * it is compiled without FMA contraction (nvcc -fmad=false, gcc -ffp-contract=off).
 *
 * Why it exists: the path tracer's only libm calls are sinf / cosf in make_dir_sample (src/core/random.impala:39-48), and
 * CUDA's and glibc's differ in the last bits, which is the one source of per-sample differences between the device's
 * films and the CPU oracle's.  With `render_poly_trig` switched on (rodent_b200_tune; the oracle has the same switch)
 * both sides call this instead, and the films must then agree up to the order of the atomic adds
 * (tests/test_gpu_render.py::test_film_matches_oracle_with_shared_trig).  Off by default: the reference calls libm.
 * Accuracy: 2e-7 absolute over [0, 2 pi].
 */
#ifndef RODENT_B200_POLY_TRIG_H
#define RODENT_B200_POLY_TRIG_H

#ifdef __CUDACC__
#define RB_TRIG_FN __host__ __device__ __forceinline__
#else
#define RB_TRIG_FN static inline
#endif

RB_TRIG_FN void rb_poly_sincos(float phi, float* s_out, float* c_out) {
    const int q = (int)(phi * 0.63661977236f + 0.5f);                 /* nearest multiple of pi/2 */
    const float fq = (float)q;
    const float r = (phi - fq * 1.5707963705062866f) - fq * -4.371139000186243e-8f;    /* Cody-Waite, pi/2 in two parts */
    const float r2 = r * r;
    const float sp = r + r * r2 * (-1.6666667163e-1f + r2 * (8.3333337680e-3f + r2 * (-1.9841270114e-4f + r2 * 2.7557314297e-6f)));
    const float cp = 1.0f + r2 * (-0.5f + r2 * (4.1666667908e-2f + r2 * (-1.3888889225e-3f + r2 * (2.4801587642e-5f + r2 * -2.7557314297e-7f))));
    switch (q & 3) {
        case 0:  *s_out = sp;  *c_out = cp;  break;
        case 1:  *s_out = cp;  *c_out = -sp; break;
        case 2:  *s_out = -sp; *c_out = -cp; break;
        default: *s_out = -cp; *c_out = sp;  break;
    }
}

#endif
