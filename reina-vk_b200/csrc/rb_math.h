/*
 * rb_math.h — elementary fp32 functions with ONE definition for host and device.
 *
 * GLSL leaves the precision of sin/cos/log/exp/acos/inversesqrt implementation-defined and the reference
 * (the GLSL under shaders/raytrace/) only ever ran on one vendor's driver, so there is no "reference bit pattern" for
 * them. What this path needs instead is that the sm_100a kernels and the CPU oracle agree BIT FOR BIT, because
 * a 1-ulp difference in a bounce direction sends the two paths to different triangles and every later bounce
 * diverges chaotically. These routines therefore use only operations that IEEE-754 defines exactly
 * (+ - * / sqrt, fma, rint/floor, integer bit manipulation); compiled with `nvcc -fmad=false` and
 * `g++ -ffp-contract=off -mfma` they return identical bits on a B200 and on the host. Accuracy is ~1-2 ulp
 * (checked against float64 libm in tests/test_rb_math.py), i.e. at least as good as a GPU driver's GLSL
 * built-ins. Polynomial coefficients are the classic single-precision Cephes minimax sets.
 */
#ifndef RB_MATH_H
#define RB_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#  define RB_HD __host__ __device__ __forceinline__
#else
#  define RB_HD static inline __attribute__((always_inline))
#endif

#define RB_PI      3.14159265f      /* k_pi      shaders/raytrace/shaderCommon.h.glsl:47 */
#define RB_INV_PI  0.31830989f      /* k_inv_pi  shaders/raytrace/shaderCommon.h.glsl:48 */
/* "x / k_pi" in the reference's shaders: its compiler (spirv-opt) turns a division by a constant into a multiplication
 * by the constant's fp32 reciprocal, 1.0f / 3.14159265f = 0x3EA2F983 — one ulp below k_inv_pi. Seen in the compiled
 * lambertian.rchit.spv (pdfLambertian) and disney.rchit.spv (fBaseDiffuse); tests/test_spirv_golden.py. */
#define RB_RCP_PI  0.31830987334251404f

RB_HD uint32_t rb_f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
RB_HD float rb_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
RB_HD int32_t rb_f2i(float f) { return (int32_t)rb_f2u(f); }
RB_HD float   rb_i2f(int32_t i) { return rb_u2f((uint32_t)i); }

RB_HD float rb_nanf(void) { return rb_u2f(0x7fc00000u); }
RB_HD float rb_inff(void) { return rb_u2f(0x7f800000u); }
RB_HD bool  rb_isnan(float x) { return x != x; }

RB_HD float rb_min(float a, float b) { return fminf(a, b); }
RB_HD float rb_max(float a, float b) { return fmaxf(a, b); }
RB_HD float rb_clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
RB_HD float rb_mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
RB_HD float rb_sqrt(float x) { return sqrtf(x); }          /* IEEE: -prec-sqrt=true is nvcc's default */
RB_HD float rb_rsqrt(float x) { return 1.0f / sqrtf(x); }  /* inversesqrt(): two exactly-rounded ops */
RB_HD float rb_fract_mod1(float x) { return x - floorf(x); } /* GLSL mod(x, 1.0) = x - 1.0*floor(x/1.0) */

/* 2^k for integer k in [-126, 127] */
RB_HD float rb_pow2i(int k) { return rb_u2f((uint32_t)(k + 127) << 23); }

/* x^2 and x^5 by repeated multiplication. GLSL pow(x, y) is undefined for x < 0; the shaders call
 * pow(v, 2) on signed tangent-space components (brdfDisney.h.glsl:224-229) and pow(1 - c, 5) for Schlick
 * terms; the evident intent (and what shader compilers emit for small constant exponents) is a product. */
RB_HD float rb_sq(float x) { return x * x; }
RB_HD float rb_pow5(float x) { float x2 = x * x; return x2 * x2 * x; }

/* sin and cos together. Cody-Waite reduction by pi/2 in three parts, then Cephes sinf/cosf kernels. */
RB_HD void rb_sincos(float x, float* s_out, float* c_out) {
    if (!(fabsf(x) < 1.0e6f)) { *s_out = rb_nanf(); *c_out = rb_nanf(); return; }
    float k = rintf(x * 0.636619772f);
    int q = (int)k;
    float r = fmaf(k, -1.5703125f, x);
    r = fmaf(k, -4.837512969970703125e-4f, r);
    r = fmaf(k, -7.549789954891882e-8f, r);
    float z = r * r;
    float ps = -1.9515295891e-4f;
    ps = fmaf(ps, z, 8.3321608736e-3f);
    ps = fmaf(ps, z, -1.6666654611e-1f);
    float s = fmaf(ps * z, r, r);
    float pc = 2.443315711809948e-5f;
    pc = fmaf(pc, z, -1.388731625493765e-3f);
    pc = fmaf(pc, z, 4.166664568298827e-2f);
    float c = fmaf(pc * z, z, fmaf(-0.5f, z, 1.0f));
    switch (q & 3) {
        case 0:  *s_out = s;  *c_out = c;  break;
        case 1:  *s_out = c;  *c_out = -s; break;
        case 2:  *s_out = -s; *c_out = -c; break;
        default: *s_out = -c; *c_out = s;  break;
    }
}
RB_HD float rb_sin(float x) { float s, c; rb_sincos(x, &s, &c); return s; }
RB_HD float rb_cos(float x) { float s, c; rb_sincos(x, &s, &c); return c; }

/* natural logarithm (Cephes logf). x == 0 -> -inf, x < 0 -> NaN. */
RB_HD float rb_log(float x) {
    if (x != x || x < 0.0f) return rb_nanf();
    if (x == 0.0f) return -rb_inff();
    if (x == rb_inff()) return x;
    int e = 0;
    if (x < 1.17549435e-38f) { x *= 16777216.0f; e = -24; }
    uint32_t u = rb_f2u(x);
    e += (int)(u >> 23) - 126;
    float m = rb_u2f((u & 0x007fffffu) | 0x3f000000u);   /* [0.5, 1) */
    if (m < 0.707106781f) { e -= 1; m = m + m - 1.0f; } else { m = m - 1.0f; }
    float z = m * m;
    float y = 7.0376836292e-2f;
    y = fmaf(y, m, -1.1514610310e-1f);
    y = fmaf(y, m, 1.1676998740e-1f);
    y = fmaf(y, m, -1.2420140846e-1f);
    y = fmaf(y, m, 1.4249322787e-1f);
    y = fmaf(y, m, -1.6668057665e-1f);
    y = fmaf(y, m, 2.0000714765e-1f);
    y = fmaf(y, m, -2.4999993993e-1f);
    y = fmaf(y, m, 3.3333331174e-1f);
    y = y * m * z;
    float fe = (float)e;
    y = fmaf(-2.12194440e-4f, fe, y);
    y = fmaf(-0.5f, z, y);
    float r = m + y;
    return fmaf(0.693359375f, fe, r);
}

/* e^x (Cephes expf). Results that would be subnormal are flushed to 0 so that no platform FTZ setting matters. */
RB_HD float rb_exp(float x) {
    if (x != x) return x;
    if (x > 88.72f) return rb_inff();
    if (x < -87.3f) return 0.0f;
    float k = floorf(fmaf(x, 1.44269504088896341f, 0.5f));
    float r = fmaf(k, -0.693359375f, x);
    r = fmaf(k, 2.12194440e-4f, r);
    float z = r * r;
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    p = fmaf(p, z, r) + 1.0f;
    int ki = (int)k;
    int k1 = ki / 2, k2 = ki - k1;
    return p * rb_pow2i(k1) * rb_pow2i(k2);
}

/* 2^x; exact for integral x (exposure = 1 -> exactly 2, tonemapping.comp.glsl:62-64). */
RB_HD float rb_exp2(float x) {
    if (x != x) return x;
    if (x > 127.9f) return rb_inff();
    if (x < -125.9f) return 0.0f;
    float k = rintf(x);
    float f = x - k;
    float r = f * 0.693147180559945f;
    float z = r * r;
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    p = fmaf(p, z, r) + 1.0f;
    int ki = (int)k;
    int k1 = ki / 2, k2 = ki - k1;
    return p * rb_pow2i(k1) * rb_pow2i(k2);
}

/* arcsine kernel for |x| <= 0.5 (Cephes asinf) */
RB_HD float rb_asin_small(float x) {
    float z = x * x;
    float p = 4.2163199048e-2f;
    p = fmaf(p, z, 2.4181311049e-2f);
    p = fmaf(p, z, 4.5470025998e-2f);
    p = fmaf(p, z, 7.4953002686e-2f);
    p = fmaf(p, z, 1.6666752422e-1f);
    return fmaf(p * z, x, x);
}

/* arccosine on [-1, 1]; NaN outside (GLSL: undefined). */
RB_HD float rb_acos(float x) {
    if (!(x >= -1.0f && x <= 1.0f)) return rb_nanf();
    if (x > 0.5f)  return 2.0f * rb_asin_small(sqrtf(0.5f * (1.0f - x)));
    if (x < -0.5f) return 3.14159265358979f - 2.0f * rb_asin_small(sqrtf(0.5f * (1.0f + x)));
    return 1.5707963267948966f - rb_asin_small(x);
}

#endif /* RB_MATH_H */
