// ORACLE (test infrastructure) — the reference's Disney BSDF library (shaders/raytrace/brdfDisney.h.glsl),
// restated as plain scalar C++. Every lobe and every pdf is evaluated on every call, exactly as the shader does
// (brdfDisney.h.glsl:685-700): a zero-weight lobe that evaluates to inf/NaN still poisons the sum, and the
// caller drops the NaN sample (raytrace.rgen.glsl:268-271).
#ifndef ORACLE_DISNEY_H
#define ORACLE_DISNEY_H

#include "oracle_common.h"

namespace oracle {

struct DisneyParams {
    vec3 baseColor, specularTint, sheenTint;
    float anisotropic, roughness, subsurface, clearcoatGloss, eta, metallic, clearcoat, specularTransmission, sheen;
};

struct Alpha { float x, y; };

// the alpha computation repeated at brdfDisney.h.glsl:250-252, 273-275, 289-291, 406-408, 443-445
static inline Alpha ggx_alpha(float anisotropic, float roughness) {
    const float aspect = sqrtf(1.0f - 0.9f * anisotropic);
    Alpha a;
    a.x = rb_max(0.0001f, roughness * roughness / aspect);
    a.y = rb_max(0.0001f, roughness * roughness * aspect);
    return a;
}

// :9-40
static inline vec3 sample_ggx_vndf(vec3 V, float ax, float ay, uint32_t& rng) {
    const bool flip = V.z < 0.0f;
    if (flip) V.z = V.z * -1.0f;
    const float r1 = rnd(rng);
    const float r2 = rnd(rng);
    vec3 Vh = rb_normalize(rb_mk3(ax * V.x, ay * V.y, V.z));
    float lensq = Vh.x * Vh.x + Vh.y * Vh.y;
    vec3 T1 = lensq > 0.0f ? rb_mk3(-Vh.y, Vh.x, 0.0f) * rb_rsqrt(lensq) : rb_mk3(1.0f, 0.0f, 0.0f);
    vec3 T2 = rb_cross(Vh, T1);
    float r = sqrtf(r1);
    float phi = (2.0f * RB_PI) * r2;
    float sp, cp; rb_sincos(phi, &sp, &cp);
    float t1 = r * cp;
    float t2 = r * sp;
    float s = 0.5f * (1.0f + Vh.z);
    t2 = (1.0f - s) * sqrtf(1.0f - t1 * t1) + s * t2;
    vec3 Nh = T1 * t1 + T2 * t2 + Vh * sqrtf(rb_max(0.0f, 1.0f - t1 * t1 - t2 * t2));
    if (flip) Nh.z = Nh.z * -1.0f;
    return rb_normalize(rb_mk3(ax * Nh.x, ay * Nh.y, rb_max(0.0f, Nh.z)));
}

// :42-48
static inline float d_ggx_aniso(vec3 m, float ax, float ay) {
    float NoM = rb_max(m.z, 0.0f);
    float tx = m.x / ax, ty = m.y / ay;
    float inv = 1.0f / (tx * tx + ty * ty + NoM * NoM);
    return inv * inv / ((RB_PI * ax) * ay);
}

// :55-73 (bounded VNDF reflection pdf)
static inline float pdf_ggx_reflection(vec3 i, vec3 o, Alpha alpha) {
    vec3 m = rb_normalize(i + o);
    float ndf = d_ggx_aniso(m, alpha.x, alpha.y);
    float aix = alpha.x * i.x, aiy = alpha.y * i.y;
    float len2 = aix * aix + aiy * aiy;
    float t = sqrtf(len2 + i.z * i.z);
    if (i.z >= 0.0f) {
        float a = rb_clamp(rb_min(alpha.x, alpha.y), 0.0f, 1.0f);
        float s = 1.0f + sqrtf(i.x * i.x + i.y * i.y);
        float a2 = a * a, s2 = s * s;
        float k = (1.0f - a2) * s2 / (s2 + a2 * i.z * i.z);
        return ndf / (2.0f * (k * i.z + t));
    }
    return ndf * (t - i.z) / (2.0f * len2);
}

// :96-111
static inline vec3 sample_gtr1(float alpha, uint32_t& rng) {
    const float r1 = rnd(rng);
    const float r2 = rnd(rng);
    const float a = rb_max(0.001f, alpha);
    const float a2 = a * a;
    float cosEl = sqrtf((1.0f - rb_exp(r1 * rb_log(a2))) / (1.0f - a2));
    float el = rb_acos(cosEl);
    float az = (2.0f * RB_PI) * r2;
    float se, ce; rb_sincos(el, &se, &ce);
    float sa, ca; rb_sincos(az, &sa, &ca);
    return rb_mk3(se * ca, se * sa, cosEl);
}

// :124-168
static inline vec3 eval_diffuse(const DisneyParams& p, vec3 n, vec3 wi, vec3 wo, vec3 h) {
    const float hdwo = rb_dot(h, wo);
    const float ndwi = rb_max(rb_dot(n, wi), 0.0f), ndwo = rb_max(rb_dot(n, wo), 0.0f);
    // fBaseDiffuse :154-161
    // FD90 = 0.5 + x, and FD uses (FD90 - 1): the reference's compiler (spirv-opt) merges the two constants into
    // x + (-0.5), which skips the rounding of 0.5 + x (seen in the compiled disney.rchit.spv)
    const float fd90m1 = 2.0f * p.roughness * rb_max(hdwo, 0.0f) * rb_max(hdwo, 0.0f) + -0.5f;
    float fdIn = 1.0f + fd90m1 * rb_pow5(1.0f - ndwi);
    float fdOut = 1.0f + fd90m1 * rb_pow5(1.0f - ndwo);
    vec3 baseDiffuse = ((p.baseColor * RB_RCP_PI) * fdIn) * fdOut;     // "/ k_pi" is compiled as "* (1 / k_pi)"
    // fSubsurface :141-152
    vec3 k = (p.baseColor * 1.25f) * RB_INV_PI;
    float fss90 = p.roughness * rb_max(hdwo, 0.0f) * rb_max(hdwo, 0.0f);
    float fssIn = 1.0f + (fss90 - 1.0f) * rb_pow5(1.0f - ndwi);
    float fssOut = 1.0f + (fss90 - 1.0f) * rb_pow5(1.0f - ndwo);
    float third = 1.0f / (ndwi + ndwo) - 0.5f;
    vec3 fss = k * (fssIn * fssOut * third + 0.5f);
    return rb_mix3(baseDiffuse, fss, p.subsurface);
}

// :170-179
static inline float pdf_diffuse(vec3 wo, vec3 n) {
    float c = rb_dot(n, wo);
    if (c <= 0.0f) return 0.0f;
    return c * RB_INV_PI;
}

// :181-201
static inline vec3 sample_diffuse(vec3 n, uint32_t& rng) {
    float xi1 = rnd(rng);
    float xi2 = rnd(rng);
    float r = sqrtf(xi1);
    float phi = (2.0f * RB_PI) * xi2;
    float sp, cp; rb_sincos(phi, &sp, &cp);
    float x = r * cp, y = r * sp;
    float z = sqrtf(rb_max(0.0f, 1.0f - xi1));
    vec3 t = fabsf(n.x) < 0.5f ? rb_normalize(rb_cross(n, rb_mk3(1.0f, 0.0f, 0.0f)))
                               : rb_normalize(rb_cross(n, rb_mk3(0.0f, 1.0f, 0.0f)));
    vec3 b = rb_cross(n, t);
    return rb_normalize(t * x + b * y + n * z);
}

static inline float eval_r0(float ior) { return (ior - 1.0f) * (ior - 1.0f) / ((ior + 1.0f) * (ior + 1.0f)); }  // :207-209
static inline float luminance(vec3 c) { return rb_dot(c, rb_mk3(0.2126f, 0.7152f, 0.0722f)); }                   // :211-213

// :215-222
static inline vec3 eval_fm(vec3 baseColor, vec3 h, vec3 wo, float specular, vec3 specularTint, float metallic, float eta) {
    float lum = luminance(baseColor);
    vec3 ctint = lum > 0.0f ? baseColor / lum : rb_splat3(1.0f);
    vec3 ks = (rb_splat3(1.0f) - ctint) + specularTint * ctint;
    vec3 c0 = ks * (specular * eval_r0(eta) * (1.0f - metallic)) + baseColor * metallic;
    return c0 + (rb_splat3(1.0f) - c0) * rb_pow5(1.0f - fabsf(rb_dot(h, wo)));
}

// :224-229
static inline float eval_dm(vec3 hl, float ax, float ay) {
    float constant = (RB_PI * ax) * ay;
    float inner = rb_sq(hl.x) / rb_sq(ax) + rb_sq(hl.y) / rb_sq(ay) + rb_sq(hl.z);
    return 1.0f / (constant * rb_sq(inner));
}

// :231-245
static inline float smith_lambda(vec3 wl, float ax, float ay) {
    float sq = sqrtf(1.0f + (rb_sq(wl.x * ax) + rb_sq(wl.y * ay)) / rb_sq(wl.z));
    return (sq - 1.0f) / 2.0f;
}
static inline float smith_g(vec3 wl, float ax, float ay) { return 1.0f / (1.0f + smith_lambda(wl, ax, ay)); }
static inline float eval_gm(vec3 wi, vec3 wo, float ax, float ay) { return smith_g(wi, ax, ay) * smith_g(wo, ax, ay); }

// :247-267
static inline vec3 eval_metal(const mat3& tbn, vec3 baseColor, float anisotropic, float roughness, vec3 n, vec3 wi,
                              vec3 wo, vec3 h, float specular, vec3 specularTint, float metallic, float eta) {
    Alpha al = ggx_alpha(anisotropic, roughness);
    vec3 fm = eval_fm(baseColor, h, wo, specular, specularTint, metallic, eta);
    vec3 wiT = rb_normalize(rb_m3_tmul(tbn, wi));
    vec3 woT = rb_normalize(rb_m3_tmul(tbn, wo));
    vec3 hT = rb_normalize(rb_m3_tmul(tbn, h));
    float dm = eval_dm(hT, al.x, al.y);
    float gm = eval_gm(wiT, woT, al.x, al.y);
    float ndwi = fabsf(rb_dot(n, wi)), ndwo = fabsf(rb_dot(n, wo));
    return ((fm * dm) * gm) / (4.0f * ndwi * ndwo);
}

// :269-285
static inline vec3 sample_metal(const mat3& tbn, float anisotropic, float roughness, vec3 wi, uint32_t& rng) {
    Alpha al = ggx_alpha(anisotropic, roughness);
    vec3 wiT = rb_m3_tmul(tbn, wi);
    vec3 h = sample_ggx_vndf(wiT, al.x, al.y, rng);
    h = rb_normalize(rb_m3_mul(tbn, h));
    return rb_reflect(-wi, h);
}

// :287-299
static inline float pdf_metal(const mat3& tbn, vec3 wi, vec3 wo, float anisotropic, float roughness) {
    Alpha al = ggx_alpha(anisotropic, roughness);
    return pdf_ggx_reflection(rb_m3_tmul(tbn, wi), rb_m3_tmul(tbn, wo), al);
}

// :313-319
static inline float separable_smith_g1(vec3 w, float a) {
    float a2 = a * a;
    float c = w.z;
    return 2.0f / (1.0f + sqrtf(a2 + (1.0f - a2) * c * c));
}
static inline float clearcoat_alpha(float gloss) { return (1.0f - gloss) * 0.1f + gloss * 0.001f; }
// :328-333
static inline float eval_dc(float ag, vec3 hl) {
    float num = ag * ag - 1.0f;
    float den = (RB_PI * rb_log(ag * ag)) * (1.0f + (ag * ag - 1.0f) * (hl.z * hl.z));
    return num / den;
}
// :341-354
static inline vec3 eval_clearcoat(const mat3& tbn, vec3 wi, vec3 wo, float gloss, vec3 h) {
    float ag = clearcoat_alpha(gloss);
    vec3 hT = rb_m3_tmul(tbn, h), wiT = rb_m3_tmul(tbn, wi), woT = rb_m3_tmul(tbn, wo);
    float r0 = eval_r0(1.5f);
    float fc = r0 + (1.0f - r0) * rb_pow5(1.0f - rb_dot(h, wo));
    float gc = separable_smith_g1(wiT, 0.25f) * separable_smith_g1(woT, 0.25f);
    float dc = eval_dc(ag, hT);
    return rb_splat3(0.25f * fc * gc * dc);
}
// :356-366
static inline vec3 sample_clearcoat(const mat3& tbn, float gloss, vec3 wi, uint32_t& rng) {
    float ag = clearcoat_alpha(gloss);
    vec3 h = rb_normalize(sample_gtr1(ag, rng));
    h = rb_normalize(rb_m3_mul(tbn, h));
    return rb_normalize(rb_reflect(-wi, h));
}
// :368-383
static inline float pdf_clearcoat(const mat3& tbn, vec3 wi, vec3 wo, vec3 h, float gloss) {
    float ag = clearcoat_alpha(gloss);
    vec3 wiT = rb_m3_tmul(tbn, wi), woT = rb_m3_tmul(tbn, wo), hT = rb_m3_tmul(tbn, h);
    if (wiT.z <= 0.0f || woT.z <= 0.0f) return 0.0f;
    float dc = eval_dc(ag, hT) * hT.z;
    return dc / (4.0f * fabsf(rb_dot(woT, hT)));
}

// :389-394
static inline float schlick_reflectance(float cosine, float ri) {
    float r0 = (1.0f - ri) / (1.0f + ri);
    r0 = r0 * r0;
    return r0 + (1.0f - r0) * rb_pow5(1.0f - cosine);
}

// :396-431
static inline vec3 sample_glass(const mat3& tbn, vec3 wi, float roughness, float anisotropic, float eta,
                                uint32_t& rng, bool* refracted) {
    Alpha al = ggx_alpha(anisotropic, roughness);
    vec3 wiT = rb_m3_tmul(tbn, wi);
    vec3 hT = sample_ggx_vndf(wiT, al.x, al.y, rng);
    float cosTheta = rb_dot(wiT, hT);
    vec3 hW = rb_normalize(rb_m3_mul(tbn, hT));
    float reflectivity = schlick_reflectance(cosTheta, eta);
    float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
    bool cannotRefract = eta * sinTheta > 1.0f;
    if (cannotRefract || reflectivity > rnd(rng)) { *refracted = false; return rb_reflect(-wi, hW); }
    *refracted = true;
    return rb_refract(-wi, hW, eta);
}

// :433-439
static inline float smith_g_aniso(float ndv, float vdx, float vdy, float ax, float ay) {
    float a = vdx * ax, b = vdy * ay, c = ndv;
    return (2.0f * ndv) / (ndv + sqrtf(a * a + b * b + c * c));
}

// :441-467
static inline vec3 eval_microfacet_refraction(vec3 baseColor, float anisotropic, float roughness, float eta,
                                              vec3 V, vec3 L, vec3 H, float* pdf) {
    Alpha al = ggx_alpha(anisotropic, roughness);
    *pdf = 0.0f;
    if (L.z >= 0.0f) return rb_splat3(0.0f);
    float ldh = rb_dot(L, H), vdh = rb_dot(V, H);
    float D = eval_dm(H, al.x, al.y);
    float G1 = smith_g_aniso(fabsf(V.z), V.x, V.y, al.x, al.y);
    float G2 = G1 * smith_g_aniso(fabsf(L.z), L.x, L.y, al.x, al.y);
    float denom = ldh + vdh * eta;
    denom = denom * denom;
    float eta2 = eta * eta;
    float jac = fabsf(ldh) * eta2 / denom;
    *pdf = G1 * rb_max(0.0f, vdh) * D * jac / V.z;
    float F = schlick_reflectance(rb_dot(V, H), eta);
    vec3 sq = rb_mk3(sqrtf(baseColor.x), sqrtf(baseColor.y), sqrtf(baseColor.z));
    return (((((sq * (1.0f - F)) * D) * G2) * fabsf(vdh)) * jac) / fabsf(L.z * V.z);
}

// :539-581
static inline vec3 eval_glass(const mat3& tbn, vec3 baseColor, float anisotropic, float roughness, float eta, vec3 n,
                              vec3 wi, vec3 wo, bool didRefract, float* pdf) {
    vec3 h = didRefract ? rb_normalize(wo + wi * eta) : rb_normalize(wo + wi);
    if (rb_dot(h, n) < 0.0f) h = -h;
    float tpdf;
    vec3 glassf = eval_microfacet_refraction(baseColor, anisotropic, roughness, eta, rb_m3_tmul(tbn, wi),
                                             rb_m3_tmul(tbn, wo), rb_m3_tmul(tbn, h), &tpdf);
    float mpdf = pdf_metal(tbn, wi, wo, anisotropic, roughness);
    vec3 metalf = eval_metal(tbn, baseColor, anisotropic, roughness, n, wi, wo, h, 0.0f, rb_splat3(1.0f), 1.0f, eta);
    if (didRefract) { *pdf = tpdf; return glassf; }
    *pdf = mpdf;
    return metalf;
}

// :587-594
static inline vec3 eval_sheen(vec3 baseColor, vec3 wo, vec3 h, vec3 n, vec3 sheenTint) {
    float lum = luminance(baseColor);
    vec3 ctint = lum > 0.0f ? baseColor / lum : rb_splat3(1.0f);
    vec3 csheen = rb_mix3v(rb_splat3(1.0f), ctint, sheenTint);
    return (csheen * rb_pow5(1.0f - rb_max(rb_dot(h, wo), 0.0f))) * rb_max(rb_dot(n, wo), 0.0f);
}

// :609-654
static inline vec3 disney_sample(const mat3& tbn, const DisneyParams& p, vec3 n, vec3 wi, bool* didRefract,
                                 bool* choseGlass, uint32_t& rng) {
    float diffuseWt = (1.0f - p.specularTransmission) * (1.0f - p.metallic);
    float metalWt = p.metallic;
    float clearcoatWt = 0.25f * p.clearcoat;
    float glassWt = (1.0f - p.metallic) * p.specularTransmission;
    float c0 = diffuseWt, c1 = c0 + metalWt, c2 = c1 + clearcoatWt, c3 = c2 + glassWt;
    *didRefract = false; *choseGlass = false;
    float r = rnd(rng) * c3;
    if (r < c0) return sample_diffuse(n, rng);
    if (r < c1) return sample_metal(tbn, p.anisotropic, p.roughness, wi, rng);
    if (r < c2) return sample_clearcoat(tbn, p.clearcoatGloss, wi, rng);
    *choseGlass = true;
    return sample_glass(tbn, wi, p.roughness, p.anisotropic, p.eta, rng, didRefract);
}

// :656-711
static inline vec3 disney_eval(const mat3& tbn, const DisneyParams& p, bool didRefract, vec3 n, vec3 wi, vec3 wo,
                               vec3 h, float* pdf) {
    float diffuseWt = (1.0f - p.specularTransmission) * (1.0f - p.metallic);
    float metalWt = p.metallic;
    float clearcoatWt = 0.25f * p.clearcoat;
    float glassWt = (1.0f - p.metallic) * p.specularTransmission;
    float wtSum = diffuseWt + metalWt + glassWt;

    vec3 fdiffuse = eval_diffuse(p, n, wi, wo, h);
    float diffusePdf = pdf_diffuse(wo, n);
    vec3 fsheen = eval_sheen(p.baseColor, wo, h, n, p.sheenTint);
    vec3 fmetal = eval_metal(tbn, p.baseColor, p.anisotropic, p.roughness, n, wi, wo, h, p.specularTransmission,
                             p.specularTint, p.metallic, p.eta);
    float metalPdf = pdf_metal(tbn, wi, wo, p.anisotropic, p.roughness);
    vec3 fclear = eval_clearcoat(tbn, wi, wo, p.clearcoatGloss, h);
    float clearPdf = pdf_clearcoat(tbn, wi, wo, h, p.clearcoatGloss);
    float glassPdf;
    vec3 fglass = eval_glass(tbn, p.baseColor, p.anisotropic, p.roughness, p.eta, n, wi, wo, didRefract, &glassPdf);

    *pdf = diffusePdf * diffuseWt / wtSum + metalPdf * metalWt / wtSum + clearPdf * clearcoatWt +
           glassPdf * glassWt / wtSum;
    return (fdiffuse + fsheen * p.sheen) * (diffuseWt / wtSum) + fmetal * (metalWt / wtSum) + fclear * clearcoatWt +
           fglass * (glassWt / wtSum);
}

} // namespace oracle
#endif
