/*
 * rb_vec.h — 3-vector / 3x3 helpers with a fixed evaluation order, shared by the kernels and the CPU oracle.
 *
 * These restate the GLSL built-ins the reference shaders lean on (dot, cross, normalize, reflect, refract,
 * faceforward, mix, inverse(mat3)). Like rb_math.h they are part of the "elementary layer": the order of the
 * fp32 operations is spelled out once so device and host round identically. No FMA contraction may be applied
 * to this file (nvcc -fmad=false, g++ -ffp-contract=off).
 */
#ifndef RB_VEC_H
#define RB_VEC_H

#include "rb_math.h"

struct rb_v3 { float x, y, z; };
struct rb_v2 { float x, y; };
/* column-major 3x3: c0, c1, c2 are the columns (GLSL mat3(c0, c1, c2)) */
struct rb_m3 { rb_v3 c0, c1, c2; };

RB_HD rb_v3 rb_mk3(float x, float y, float z) { rb_v3 r; r.x = x; r.y = y; r.z = z; return r; }
RB_HD rb_v3 rb_splat3(float s) { return rb_mk3(s, s, s); }
RB_HD rb_v2 rb_mk2(float x, float y) { rb_v2 r; r.x = x; r.y = y; return r; }

RB_HD rb_v3 operator+(rb_v3 a, rb_v3 b) { return rb_mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
RB_HD rb_v3 operator-(rb_v3 a, rb_v3 b) { return rb_mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
RB_HD rb_v3 operator*(rb_v3 a, rb_v3 b) { return rb_mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
RB_HD rb_v3 operator/(rb_v3 a, rb_v3 b) { return rb_mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
RB_HD rb_v3 operator*(rb_v3 a, float s) { return rb_mk3(a.x * s, a.y * s, a.z * s); }
RB_HD rb_v3 operator*(float s, rb_v3 a) { return rb_mk3(s * a.x, s * a.y, s * a.z); }
RB_HD rb_v3 operator/(rb_v3 a, float s) { return rb_mk3(a.x / s, a.y / s, a.z / s); }
RB_HD rb_v3 operator-(rb_v3 a) { return rb_mk3(-a.x, -a.y, -a.z); }
RB_HD rb_v2 operator+(rb_v2 a, rb_v2 b) { return rb_mk2(a.x + b.x, a.y + b.y); }
RB_HD rb_v2 operator-(rb_v2 a, rb_v2 b) { return rb_mk2(a.x - b.x, a.y - b.y); }
RB_HD rb_v2 operator*(rb_v2 a, float s) { return rb_mk2(a.x * s, a.y * s); }

RB_HD float rb_dot(rb_v3 a, rb_v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
RB_HD float rb_dot2(rb_v2 a, rb_v2 b) { return a.x * b.x + a.y * b.y; }
RB_HD rb_v3 rb_cross(rb_v3 a, rb_v3 b) {
    return rb_mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
RB_HD float rb_length(rb_v3 a) { return sqrtf(rb_dot(a, a)); }
RB_HD rb_v3 rb_normalize(rb_v3 a) { float inv = 1.0f / sqrtf(rb_dot(a, a)); return a * inv; }
RB_HD rb_v3 rb_reflect(rb_v3 I, rb_v3 N) { float d2 = 2.0f * rb_dot(N, I); return I - N * d2; }
/* GLSL refract: zero vector on total internal reflection */
RB_HD rb_v3 rb_refract(rb_v3 I, rb_v3 N, float eta) {
    float ni = rb_dot(N, I);
    float k = 1.0f - eta * eta * (1.0f - ni * ni);
    if (k < 0.0f) return rb_splat3(0.0f);
    float f = eta * ni + sqrtf(k);
    return I * eta - N * f;
}
RB_HD rb_v3 rb_faceforward(rb_v3 N, rb_v3 I, rb_v3 Nref) { return rb_dot(Nref, I) < 0.0f ? N : -N; }
RB_HD rb_v3 rb_mix3(rb_v3 a, rb_v3 b, float t) { return a * (1.0f - t) + b * t; }
RB_HD rb_v3 rb_mix3v(rb_v3 a, rb_v3 b, rb_v3 t) {
    return rb_mk3(rb_mix(a.x, b.x, t.x), rb_mix(a.y, b.y, t.y), rb_mix(a.z, b.z, t.z));
}
RB_HD rb_v3 rb_clamp3(rb_v3 a, float lo, float hi) {
    return rb_mk3(rb_clamp(a.x, lo, hi), rb_clamp(a.y, lo, hi), rb_clamp(a.z, lo, hi));
}
/* Per-sample clamp of raytrace.rgen.glsl:266. GLSL leaves clamp(NaN) undefined; the reference relies on NaN
 * samples surviving it so that the isnan() test at :268-271 can drop them, so NaN is passed through here. */
RB_HD float rb_clamp_keepnan(float x, float lo, float hi) { return x != x ? x : fminf(fmaxf(x, lo), hi); }
RB_HD rb_v3 rb_clamp3_keepnan(rb_v3 a, float lo, float hi) {
    return rb_mk3(rb_clamp_keepnan(a.x, lo, hi), rb_clamp_keepnan(a.y, lo, hi), rb_clamp_keepnan(a.z, lo, hi));
}
RB_HD bool rb_anynan3(rb_v3 a) { return a.x != a.x || a.y != a.y || a.z != a.z; }

/* M * v (columns c0..c2) */
RB_HD rb_v3 rb_m3_mul(const rb_m3& m, rb_v3 v) {
    return rb_mk3(m.c0.x * v.x + m.c1.x * v.y + m.c2.x * v.z,
                  m.c0.y * v.x + m.c1.y * v.y + m.c2.y * v.z,
                  m.c0.z * v.x + m.c1.z * v.y + m.c2.z * v.z);
}
/* transpose(M) * v */
RB_HD rb_v3 rb_m3_tmul(const rb_m3& m, rb_v3 v) { return rb_mk3(rb_dot(m.c0, v), rb_dot(m.c1, v), rb_dot(m.c2, v)); }

/* transpose(inverse(M)) by cofactors: column j of the result is cross(c_{j+1}, c_{j+2}) / det */
RB_HD rb_m3 rb_m3_inverse_transpose(const rb_m3& m) {
    rb_v3 r0 = rb_cross(m.c1, m.c2);
    rb_v3 r1 = rb_cross(m.c2, m.c0);
    rb_v3 r2 = rb_cross(m.c0, m.c1);
    float inv = 1.0f / rb_dot(m.c0, r0);
    rb_m3 r; r.c0 = r0 * inv; r.c1 = r1 * inv; r.c2 = r2 * inv;
    return r;
}

/* column-major 4x4 (glm) applied to a point / its upper 3x3 */
RB_HD rb_v3 rb_m4_point(const float* m, rb_v3 p) {
    return rb_mk3(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
                  m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                  m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]);
}
RB_HD rb_m3 rb_m4_upper3(const float* m) {
    rb_m3 r; r.c0 = rb_mk3(m[0], m[1], m[2]); r.c1 = rb_mk3(m[4], m[5], m[6]); r.c2 = rb_mk3(m[8], m[9], m[10]);
    return r;
}

/* Two-level instancing (src/scene/Scene.cpp:93-111: one BLAS per object, a transform per instance; the Vulkan driver
 * moves the RAY into object space, src/tools/vktools.cpp:466-468 hands it the 3x4 instance matrix). The object-space ray
 * is inv * (o, 1), inv * (d, 0) with inv = the inverse of the instance's affine transform, three rows of four floats;
 * t is the same parameter in both spaces because d is not re-normalised. One definition for the kernels and the oracle.
 * rb_affine_inverse runs on the host only (fp64 cofactors, rounded once): both sides get the same twelve floats. */
RB_HD rb_v3 rb_inv_point(const float* inv, rb_v3 p) {
    return rb_mk3(inv[0] * p.x + inv[1] * p.y + inv[2] * p.z + inv[3],
                  inv[4] * p.x + inv[5] * p.y + inv[6] * p.z + inv[7],
                  inv[8] * p.x + inv[9] * p.y + inv[10] * p.z + inv[11]);
}
RB_HD rb_v3 rb_inv_vector(const float* inv, rb_v3 d) {
    return rb_mk3(inv[0] * d.x + inv[1] * d.y + inv[2] * d.z,
                  inv[4] * d.x + inv[5] * d.y + inv[6] * d.z,
                  inv[8] * d.x + inv[9] * d.y + inv[10] * d.z);
}
/* m: column-major 4x4 whose last row is (0, 0, 0, 1). Returns false for a singular (or non-finite) linear part. */
static inline bool rb_affine_inverse(const float* m, float* inv) {
    const double a00 = m[0], a01 = m[4], a02 = m[8], a10 = m[1], a11 = m[5], a12 = m[9], a20 = m[2], a21 = m[6], a22 = m[10];
    const double c00 = a11 * a22 - a12 * a21, c01 = a12 * a20 - a10 * a22, c02 = a10 * a21 - a11 * a20;
    const double det = a00 * c00 + a01 * c01 + a02 * c02;
    if (!(det != 0.0) || !(det - det == 0.0)) return false;
    const double r[3][3] = {{c00 / det, (a02 * a21 - a01 * a22) / det, (a01 * a12 - a02 * a11) / det},
                            {c01 / det, (a00 * a22 - a02 * a20) / det, (a02 * a10 - a00 * a12) / det},
                            {c02 / det, (a01 * a20 - a00 * a21) / det, (a00 * a11 - a01 * a10) / det}};
    const double tx = m[12], ty = m[13], tz = m[14];
    for (int i = 0; i < 3; i++) {
        inv[4 * i + 0] = (float)r[i][0]; inv[4 * i + 1] = (float)r[i][1]; inv[4 * i + 2] = (float)r[i][2];
        inv[4 * i + 3] = (float)(-(r[i][0] * tx + r[i][1] * ty + r[i][2] * tz));
    }
    for (int i = 0; i < 12; i++) if (!(inv[i] - inv[i] == 0.0f)) return false;
    return true;
}

/* Wächter & Binder self-intersection offset, shaders/raytrace/closestHitCommon.h.glsl:156-177.
 * Integer arithmetic on the fp32 bit patterns: must be (and is) bit-exact everywhere. */
RB_HD float rb_offset_component(float p, float n) {
    int of_i = (int)(256.0f * n);
    float p_i = rb_i2f(rb_f2i(p) + ((p < 0.0f) ? -of_i : of_i));
    return fabsf(p) < (1.0f / 32.0f) ? p + (1.0f / 65536.0f) * n : p_i;
}
RB_HD rb_v3 rb_offset_along_normal(rb_v3 p, rb_v3 n) {
    return rb_mk3(rb_offset_component(p.x, n.x), rb_offset_component(p.y, n.y), rb_offset_component(p.z, n.z));
}

/* PCG step, shaders/raytrace/shaderCommon.h.glsl:39-45. Returns [0, 1] inclusive:
 * float(word) rounds to nearest, 4294967295.0f is 2^32 in fp32. */
RB_HD float rb_random(uint32_t* state) {
    uint32_t s = *state * 747796405u + 1u;
    *state = s;
    uint32_t w = ((s >> ((s >> 28) + 4u)) ^ s) * 277803737u;
    w = (w >> 22) ^ w;
    return (float)w * 2.3283064365386963e-10f;
}

#endif /* RB_VEC_H */
