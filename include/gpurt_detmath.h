/*
 * gpurt_detmath.h — deterministic fp32 sin / cos / pow / exp.
 *
 * The reference shaders call GLSL sin, cos and pow (src/shaders/rt/rtcommon.glsl:137-169, :255-343);
 * Vulkan only bounds their error (sin/cos: 2^-11 absolute on [-pi,pi]; pow: derived from
 * exp2(y*log2(x))), so any implementation inside those bounds is a valid restatement.  libm's and
 * CUDA's versions differ in the last bits, which would make CPU-vs-GPU image comparison fuzzy and
 * let stochastic decisions (Russian roulette, reservoir updates) diverge.  These versions use only
 * IEEE add/mul/div, explicit fmaf and exact integer/bit operations, so they return the same bits on
 * the host and on the device.  Accuracy: about 1e-7 absolute for sin/cos, about 2e-7 relative for
 * exp2/log2 — tighter than the Vulkan bounds.  This is a numeric library (the role libm plays), not
 * part of the algorithm: both the CUDA kernels and the CPU oracle include it.
 */
#ifndef GPURT_DETMATH_H
#define GPURT_DETMATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define GPURT_DM static __host__ __device__ __forceinline__
#else
#define GPURT_DM static inline
#endif

GPURT_DM uint32_t dm_f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
GPURT_DM float dm_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

/* quadrant reduction: x = q*(pi/2) + r, |r| <= pi/4 (Cody-Waite, three fp32 constants) */
GPURT_DM float dm_reduce(float x, int* quadrant) {
    float q = rintf(x * 0.636619772367581343f);
    float r = fmaf(-q, 1.5707962513e+0f, x);
    r = fmaf(-q, 7.5497894159e-08f, r);
    r = fmaf(-q, 5.3903029534e-15f, r);
    *quadrant = (int)q;
    return r;
}
GPURT_DM float dm_sin_poly(float r) {
    float s = r * r;
    float p = fmaf(s, -1.9515295891e-4f, 8.3321608736e-3f);
    p = fmaf(s, p, -1.6666654611e-1f);
    return fmaf(r * s, p, r);
}
GPURT_DM float dm_cos_poly(float r) {
    float s = r * r;
    float p = fmaf(s, 2.443315711809948e-5f, -1.388731625493765e-3f);
    p = fmaf(s, p, 4.166664568298827e-2f);
    return fmaf(s * s, p, fmaf(-0.5f, s, 1.0f));
}
GPURT_DM float dm_sin(float x) {
    int q;
    float r = dm_reduce(x, &q);
    float v = (q & 1) ? dm_cos_poly(r) : dm_sin_poly(r);
    return (q & 2) ? -v : v;
}
GPURT_DM float dm_cos(float x) {
    int q;
    float r = dm_reduce(x, &q);
    float v = (q & 1) ? dm_sin_poly(r) : dm_cos_poly(r);
    return ((q + 1) & 2) ? -v : v;
}

/* log2 of a positive, finite, normal-or-denormal float */
GPURT_DM float dm_log2(float x) {
    uint32_t u = dm_f2u(x);
    int e = 0;
    if(u < 0x00800000u) { /* denormal: scale up by 2^23 */
        x = x * 8388608.0f;
        u = dm_f2u(x);
        e = -23;
    }
    e += (int)(u >> 23) - 127;
    float m = dm_u2f((u & 0x007fffffu) | 0x3f800000u); /* [1,2) */
    if(m > 1.41421356f) {
        m = m * 0.5f;
        e += 1;
    }
    float z = (m - 1.0f) / (m + 1.0f); /* |z| <= 0.1716 */
    float s = z * z;
    float p = fmaf(s, 0.2222222222f, 0.2857142857f); /* 2/9, 2/7 */
    p = fmaf(s, p, 0.4f);
    p = fmaf(s, p, 0.6666666667f);
    p = fmaf(s, p, 2.0f);
    float ln_m = z * p;
    return fmaf(ln_m, 1.44269504088896341f, (float)e);
}
GPURT_DM float dm_exp2(float t) {
    if(!(t > -150.0f)) return t != t ? t : 0.0f;
    if(t > 128.0f) return dm_u2f(0x7f800000u);
    float n = rintf(t);
    float f = (t - n) * 0.693147180559945309f; /* |f| <= 0.3466 */
    float p = fmaf(f, 1.9841269841e-4f, 1.3888888889e-3f); /* 1/7!, 1/6! */
    p = fmaf(f, p, 8.3333333333e-3f);
    p = fmaf(f, p, 4.1666666667e-2f);
    p = fmaf(f, p, 1.6666666667e-1f);
    p = fmaf(f, p, 0.5f);
    p = fmaf(f, p, 1.0f);
    p = fmaf(f, p, 1.0f);
    int k = (int)n;
    /* scale by 2^k in two exact steps so that k in [-150,128] stays representable */
    int k1 = k / 2, k2 = k - k1;
    return (p * dm_u2f((uint32_t)(k1 + 127) << 23)) * dm_u2f((uint32_t)(k2 + 127) << 23);
}
/* GLSL pow(x,y): undefined for x<0 (NaN here), pow(0,y>0)=0 */
GPURT_DM float dm_pow(float x, float y) {
    if(y == 0.0f) return 1.0f;
    if(x == 0.0f) return y > 0.0f ? 0.0f : dm_u2f(0x7f800000u);
    if(!(x > 0.0f)) return dm_u2f(0x7fc00000u);
    if(x == dm_u2f(0x7f800000u)) return y > 0.0f ? x : 0.0f;
    return dm_exp2(y * dm_log2(x));
}
GPURT_DM float dm_exp(float x) { return dm_exp2(x * 1.44269504088896341f); }

#endif /* GPURT_DETMATH_H */
