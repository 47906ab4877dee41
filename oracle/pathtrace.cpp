// ORACLE (test infrastructure) — scalar restatement of the reference's ray-generation shader, its four
// closest-hit material shaders, the miss shaders and the light sampler. One pixel at a time, one RNG stream per
// (pixel, sampleBatch), exactly the draw order of SURVEY.md Appendix A.
#include "oracle_common.h"
#include "disney.h"
#include <cmath>

namespace oracle {

float kat_random(uint32_t* state) { return rnd(*state); }

// ---------------------------------------------------------------------------------------------------
// textures: RGBA8 UNORM, bilinear, REPEAT, LOD 0 (src/tools/vktools.cpp:765-788, src/graphics/Image.cpp:43-86)
// ---------------------------------------------------------------------------------------------------
struct vec4 { float x, y, z, w; };

static inline int wrapi(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }

static vec4 sample_texture(const Scene& s, int id, vec2 uv) {
    const Scene::Tex& tx = s.textures[(size_t)id];
    float x = uv.x * (float)tx.w - 0.5f;
    float y = uv.y * (float)tx.h - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float ax = x - fx, ay = y - fy;
    // REPEAT addressing of a coordinate no int32 can hold (parallax steps at grazing angles, NaN) is not defined by
    // the float->int conversion: such a coordinate addresses texel 0 with weight 0 (same rule in the kernels)
    if (!(fabsf(fx) < 1073741824.0f)) { fx = 0.0f; ax = 0.0f; }
    if (!(fabsf(fy) < 1073741824.0f)) { fy = 0.0f; ay = 0.0f; }
    int x0 = wrapi((int)fx, (int)tx.w), y0 = wrapi((int)fy, (int)tx.h);
    int x1 = wrapi(x0 + 1, (int)tx.w), y1 = wrapi(y0 + 1, (int)tx.h);
    const uint8_t* p00 = &tx.rgba[4 * ((size_t)y0 * tx.w + x0)];
    const uint8_t* p10 = &tx.rgba[4 * ((size_t)y0 * tx.w + x1)];
    const uint8_t* p01 = &tx.rgba[4 * ((size_t)y1 * tx.w + x0)];
    const uint8_t* p11 = &tx.rgba[4 * ((size_t)y1 * tx.w + x1)];
    float out[4];
    for (int c = 0; c < 4; c++) {
        float c00 = (float)p00[c] / 255.0f, c10 = (float)p10[c] / 255.0f;
        float c01 = (float)p01[c] / 255.0f, c11 = (float)p11[c] / 255.0f;
        float top = c00 * (1.0f - ax) + c10 * ax;
        float bot = c01 * (1.0f - ax) + c11 * ax;
        out[c] = top * (1.0f - ay) + bot * ay;
    }
    vec4 r = {out[0], out[1], out[2], out[3]};
    return r;
}

// ---------------------------------------------------------------------------------------------------
// payload (shaders/raytrace/shaderCommon.h.glsl:18-36)
// ---------------------------------------------------------------------------------------------------
struct Payload {
    vec3 albedo, color, rayOrigin, rayDirection;
    uint32_t rngState;
    bool rayHitSky;
    vec3 emission, surfaceNormal;
    uint32_t materialID;
    bool skip;
    float pdf;
    float accumulatedDistance;
    bool insideDielectric;
    mat3 tbn;
    float eta;
    bool didRefract;
    const RB200InstanceProperties* props;
};

struct HitInfo {   // shaders/raytrace/closestHitCommon.h.glsl:41-50
    vec3 objectPosition, worldPosition, worldNormal, worldNormalGeometry;
    vec2 uv;
    bool frontFace;
    mat3 tbn;
};

static inline vec3 vtx(const Scene& s, uint32_t i) {
    return rb_mk3(s.vertices[4 * (size_t)i], s.vertices[4 * (size_t)i + 1], s.vertices[4 * (size_t)i + 2]);
}
static inline vec3 tbncol(const Scene& s, uint32_t i, int col) {
    const float* m = &s.tbns[9 * (size_t)i + 3 * col];
    return rb_mk3(m[0], m[1], m[2]);
}

// shaders/raytrace/closestHitCommon.h.glsl:52-148
static HitInfo hit_info(const Scene& s, const RB200Instance& inst, const RB200InstanceProperties& props,
                        uint32_t prim, float a1, float a2, vec3 rayDir) {
    HitInfo r;
    const uint32_t base = 3 * prim + props.indicesOffset;
    const uint32_t i0 = s.indices[base], i1 = s.indices[base + 1], i2 = s.indices[base + 2];
    const vec3 v0 = vtx(s, i0), v1 = vtx(s, i1), v2 = vtx(s, i2);
    const float bx = 1.0f - a1 - a2, by = a1, bz = a2;

    r.objectPosition = v0 * bx + v1 * by + v2 * bz;
    r.worldPosition = rb_m4_point(inst.transform, r.objectPosition);

    const vec3 ngObj = rb_normalize(rb_cross(v1 - v0, v2 - v0));
    const uint32_t tb = 3 * prim + props.tbnsIndicesOffset;
    const uint32_t t0 = s.tbnIndices[tb], t1 = s.tbnIndices[tb + 1], t2 = s.tbnIndices[tb + 2];
    vec3 nObj;
    if (!props.interpNormals) {
        nObj = ngObj;
    } else {
        const vec3 n0 = rb_normalize(tbncol(s, t0, 2));
        const vec3 n1 = rb_normalize(tbncol(s, t1, 2));
        const vec3 n2 = rb_normalize(tbncol(s, t2, 2));
        nObj = rb_normalize(n0 * bx + n1 * by + n2 * bz);
    }

    if (props.texIndicesOffset == 0xFFFFFFFFu) {
        r.uv = rb_mk2(0.0f, 0.0f);
    } else {
        const uint32_t xb = 3 * prim + props.texIndicesOffset;
        const uint32_t x0 = s.texIndices[xb], x1 = s.texIndices[xb + 1], x2 = s.texIndices[xb + 2];
        const vec2 c0 = rb_mk2(s.texCoords[2 * (size_t)x0], s.texCoords[2 * (size_t)x0 + 1]);
        const vec2 c1 = rb_mk2(s.texCoords[2 * (size_t)x1], s.texCoords[2 * (size_t)x1 + 1]);
        const vec2 c2 = rb_mk2(s.texCoords[2 * (size_t)x2], s.texCoords[2 * (size_t)x2 + 1]);
        r.uv = c0 * bx + c1 * by + c2 * bz;
    }

    const mat3 M = rb_m4_upper3(inst.transform);
    r.worldNormal = rb_normalize(rb_m3_mul(M, nObj));
    r.worldNormalGeometry = rb_normalize(rb_m3_mul(M, ngObj));

    r.frontFace = rb_dot(rayDir, r.worldNormalGeometry) < 0.0f;
    r.worldNormal = rb_faceforward(r.worldNormal, rayDir, r.worldNormalGeometry);
    r.worldNormalGeometry = rb_faceforward(r.worldNormalGeometry, rayDir, r.worldNormalGeometry);

    // tangent frame (:121-145)
    vec3 tangent = rb_normalize(tbncol(s, t0, 0) * bx + tbncol(s, t1, 0) * by + tbncol(s, t2, 0) * bz);
    vec3 bitangent = rb_normalize(tbncol(s, t0, 1) * bx + tbncol(s, t1, 1) * by + tbncol(s, t2, 1) * bz);
    vec3 normal = rb_normalize(tbncol(s, t0, 2) * bx + tbncol(s, t1, 2) * by + tbncol(s, t2, 2) * bz);
    const mat3 Nm = rb_m3_inverse_transpose(M);
    vec3 worldT = rb_normalize(rb_m3_mul(M, tangent));
    vec3 worldB = rb_normalize(rb_m3_mul(M, bitangent));
    vec3 worldN = rb_normalize(rb_m3_mul(Nm, normal));
    worldT = rb_normalize(worldT - worldN * rb_dot(worldN, worldT));
    worldB = rb_normalize(worldB - worldN * rb_dot(worldN, worldB));
    worldB = worldB * -1.0f;
    worldN = worldN * (r.frontFace ? 1.0f : -1.0f);
    r.tbn.c0 = worldT; r.tbn.c1 = worldB; r.tbn.c2 = worldN;
    return r;
}

// shaders/raytrace/closestHitCommon.h.glsl:156-177 — restated independently of rb_offset_along_normal
static vec3 offset_along_normal(vec3 p, vec3 n) {
    const float in[3] = {p.x, p.y, p.z}, nn[3] = {n.x, n.y, n.z};
    float out[3];
    for (int k = 0; k < 3; k++) {
        int32_t step = (int32_t)(256.0f * nn[k]);
        int32_t bits; memcpy(&bits, &in[k], 4);
        bits += (in[k] < 0.0f) ? -step : step;
        float moved; memcpy(&moved, &bits, 4);
        out[k] = (fabsf(in[k]) < 0.03125f) ? in[k] + 1.52587890625e-05f * nn[k] : moved;
    }
    return rb_mk3(out[0], out[1], out[2]);
}
vec3 kat_offset(vec3 p, vec3 n) { return offset_along_normal(p, n); }

// dielectric.rchit.glsl:14-38 == disney.rchit.glsl:10-34
static vec3 offset_for_dielectric(vec3 p, vec3 n, vec3 rayDir) {
    vec3 on = (rb_dot(n, rayDir) < 0.0f) ? -n : n;
    return offset_along_normal(p, on);
}

// closestHitCommon.h.glsl:185-193
static void do_skip(Payload& pld, const HitInfo& h, vec3 rayDir) {
    pld.rayOrigin = offset_along_normal(h.worldPosition, -h.worldNormal);
    pld.rayDirection = rayDir;
    pld.rayHitSky = false;
    pld.skip = true;
}

// texutils.h.glsl:4-41 — steep parallax search (64..512 layers, heightScale 0.2) + linear refinement between the
// last two layers. rayIn is the normalised world ray direction, T the hit's tangent frame. mix(a, b, t) is
// a*(1-t) + b*t, "V.xy / V.z * heightScale" associates left to right.
static vec2 bump_mapping(const Scene& s, vec2 uv, vec3 rayIn, const mat3& T, int heightMap) {
    const float heightScale = 0.2f, minLayers = 64.0f, maxLayers = 512.0f;
    const vec3 V = rb_normalize(rb_m3_tmul(T, -rayIn));
    if (V.z <= 0.0f) return uv;
    const float a = rb_clamp(V.z, 0.0f, 1.0f);
    const float numLayers = maxLayers * (1.0f - a) + minLayers * a;
    const float layerDepth = 1.0f / numLayers;
    const float Px = (V.x / V.z) * heightScale, Py = (V.y / V.z) * heightScale;
    const float dU = Px / numLayers, dV = Py / numLayers;
    float cu = uv.x, cv = uv.y;
    float depthSum = 0.0f;
    float h = sample_texture(s, heightMap, rb_mk2(cu, cv)).x * heightScale;
    while (depthSum < h) {
        cu = cu - dU; cv = cv - dV;
        depthSum = depthSum + layerDepth;
        h = sample_texture(s, heightMap, rb_mk2(cu, cv)).x * heightScale;
    }
    const float pu = cu + dU, pv = cv + dV;
    const float hPrev = sample_texture(s, heightMap, rb_mk2(pu, pv)).x * heightScale;
    const float sumPrev = depthSum - layerDepth;
    const float after = h - depthSum;
    const float before = hPrev - sumPrev;
    const float weight = after / (after - before);
    return rb_mk2(pu * (1.0f - weight) + cu * weight, pv * (1.0f - weight) + cv * weight);
}

// known-answer hook: the parallax search on a caller-provided height map
void kat_bump(const uint8_t* rgba, uint32_t w, uint32_t h, const float uv[2], const float rayIn[3], const float tbn[9],
              float out[2]) {
    Scene s;
    Scene::Tex t; t.w = w; t.h = h; t.rgba.assign(rgba, rgba + 4 * (size_t)w * h);
    s.textures.push_back(std::move(t));
    mat3 T; T.c0 = rb_mk3(tbn[0], tbn[1], tbn[2]); T.c1 = rb_mk3(tbn[3], tbn[4], tbn[5]); T.c2 = rb_mk3(tbn[6], tbn[7], tbn[8]);
    vec2 r = bump_mapping(s, rb_mk2(uv[0], uv[1]), rb_mk3(rayIn[0], rayIn[1], rayIn[2]), T, 0);
    out[0] = r.x; out[1] = r.y;
}

// closestHitCommon.h.glsl:195-205
static vec3 random_unit_vec(uint32_t& rng) {
    for (;;) {
        float a = rnd(rng), b = rnd(rng), c = rnd(rng);
        vec3 v = rb_mk3(2.0f * a - 1.0f, 2.0f * b - 1.0f, 2.0f * c - 1.0f);
        float l2 = rb_dot(v, v);
        if (0.0001f < l2 && l2 < 1.0f) return rb_normalize(v);
    }
}
// closestHitCommon.h.glsl:211-213. The function takes `inout uint rngState` but draws from the GLOBAL pld.rngState;
// every caller passes pld.rngState itself, so GLSL's copy-in / copy-out writes the value the state had BEFORE the
// call back over the advanced one: the draws of randomUnitVec are not consumed and the next random() repeats them.
// (Confirmed by executing the reference's compiled metal.rchit.spv: tests/test_spirv_golden.py.)
static vec3 fuzzy_reflection(vec3 in, vec3 n, float fuzz, uint32_t& rng) {
    const uint32_t stateAtCall = rng;
    vec3 r = rb_reflect(rb_normalize(in), rb_normalize(n));
    vec3 out = r + random_unit_vec(rng) * fuzz;
    rng = stateAtCall;
    return out;
}
// closestHitCommon.h.glsl:215-222
static vec3 diffuse_reflection(vec3 n, uint32_t& rng) {
    const float theta = (2.0f * RB_PI) * rnd(rng);
    const float u = 2.0f * rnd(rng) - 1.0f;
    const float r = sqrtf(1.0f - u * u);
    float sn, cs; rb_sincos(theta, &sn, &cs);
    return rb_normalize(n + rb_mk3(r * cs, r * sn, u));
}

// Common prologue of lambertian/metal/disney: cull, UV wrap + range skip, normal map, albedo texture with
// stochastic alpha. Returns false when the hit was turned into a skip.
// (lambertian.rchit.glsl:15-56, metal.rchit.glsl:12-48, disney.rchit.glsl:40-84)
static bool surface_prologue(const Scene& s, Payload& pld, const HitInfo& h, const RB200InstanceProperties& props,
                             vec3 rayDir, bool uvRangeSkip, vec3* worldNormal, vec3* albedo) {
    if (props.cullBackface != 0u && !h.frontFace) { do_skip(pld, h, rayDir); return false; }
    vec2 uv = rb_mk2(rb_fract_mod1(h.uv.x), rb_fract_mod1(h.uv.y));
    if (props.bumpMapTexID >= 0) uv = bump_mapping(s, uv, rb_normalize(rayDir), h.tbn, props.bumpMapTexID);
    if (uvRangeSkip && (uv.x < 0.0f || uv.x > 1.0f || uv.y < 0.0f || uv.y > 1.0f)) { do_skip(pld, h, rayDir); return false; }
    vec3 wn = h.worldNormal;
    if (props.normalMapTexID >= 0) {
        vec4 t = sample_texture(s, props.normalMapTexID, uv);
        vec3 tn = rb_mk3(t.x * 2.0f - 1.0f, t.y * 2.0f - 1.0f, t.z * 2.0f - 1.0f);
        tn.y = tn.y * -1.0f;
        wn = rb_normalize(rb_m3_mul(h.tbn, tn));
    }
    vec3 col = rb_mk3(props.albedo[0], props.albedo[1], props.albedo[2]);
    if (props.textureID >= 0) {
        vec4 t = sample_texture(s, props.textureID, uv);
        if (t.w < 0.999f && rnd(pld.rngState) > t.w) { do_skip(pld, h, rayDir); return false; }
        col = col * rb_mk3(t.x, t.y, t.z);
    }
    *worldNormal = wn; *albedo = col;
    return true;
}

static inline void track_distance(Payload& pld, const HitInfo& h, vec3 rayOrigin) {
    // lambertian.rchit.glsl:72-76 (same tail in metal and disney)
    if (pld.insideDielectric) pld.accumulatedDistance += rb_length(h.worldPosition - rayOrigin);
    else pld.accumulatedDistance = 0.0f;
}

// lambertian.rchit.glsl:11-78
static void shade_lambertian(const Scene& s, Payload& pld, const HitInfo& h, const RB200InstanceProperties& props,
                             vec3 rayOrigin, vec3 rayDir) {
    vec3 wn, col;
    if (!surface_prologue(s, pld, h, props, rayDir, true, &wn, &col)) return;
    pld.color = col; pld.albedo = col;
    pld.emission = rb_mk3(props.emission[0], props.emission[1], props.emission[2]);
    pld.rayOrigin = offset_along_normal(h.worldPosition, h.worldNormalGeometry);
    pld.rayDirection = diffuse_reflection(wn, pld.rngState);
    pld.rayHitSky = false; pld.skip = false; pld.insideDielectric = false;
    pld.materialID = 0; pld.surfaceNormal = wn;
    pld.pdf = rb_max(rb_dot(pld.surfaceNormal, pld.rayDirection), 0.0f) * RB_RCP_PI;   // pdf.h.glsl:10-12
    pld.tbn = h.tbn; pld.props = &props; pld.didRefract = false; pld.eta = 0.0f;
    track_distance(pld, h, rayOrigin);
}

// metal.rchit.glsl:7-70 (no UV-range skip in this shader)
static void shade_metal(const Scene& s, Payload& pld, const HitInfo& h, const RB200InstanceProperties& props,
                        vec3 rayOrigin, vec3 rayDir) {
    vec3 wn, col;
    if (!surface_prologue(s, pld, h, props, rayDir, false, &wn, &col)) return;
    pld.color = col; pld.albedo = col;
    pld.emission = rb_mk3(props.emission[0], props.emission[1], props.emission[2]);
    pld.rayOrigin = offset_along_normal(h.worldPosition, h.worldNormalGeometry);
    pld.rayDirection = fuzzy_reflection(rayDir, wn, props.roughness, pld.rngState);
    pld.rayHitSky = false; pld.skip = false; pld.insideDielectric = false;
    pld.materialID = 1; pld.surfaceNormal = wn; pld.pdf = 0.0f;
    pld.tbn = h.tbn; pld.props = &props; pld.didRefract = false; pld.eta = 0.0f;
    track_distance(pld, h, rayOrigin);
}

// dielectric.rchit.glsl:7-12
static float schlick(float cosine, float refIdx) {
    float r0 = (1.0f - refIdx) / (1.0f + refIdx);
    r0 = r0 * r0;
    return r0 + (1.0f - r0) * rb_pow5(1.0f - cosine);
}

// dielectric.rchit.glsl:40-113
static void shade_dielectric(const Scene& s, Payload& pld, const HitInfo& h, const RB200InstanceProperties& props,
                             vec3 rayOrigin, vec3 rayDir) {
    const float ri = h.frontFace ? 1.0f / props.ior : props.ior;
    const vec3 unitDir = rb_normalize(rayDir);
    const float cosTheta = rb_min(rb_dot(-unitDir, h.worldNormal), 1.0f);
    const float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
    const bool cannotRefract = ri * sinTheta > 1.0f;
    const float reflectivity = schlick(cosTheta, ri);
    const bool previouslyInside = pld.insideDielectric;

    vec2 uv = rb_mk2(rb_fract_mod1(h.uv.x), rb_fract_mod1(h.uv.y));
    if (props.bumpMapTexID >= 0) uv = bump_mapping(s, uv, rb_normalize(rayDir), h.tbn, props.bumpMapTexID);   // :59-61
    vec3 albedo = rb_mk3(props.albedo[0], props.albedo[1], props.albedo[2]);
    if (props.textureID >= 0) {
        vec4 t = sample_texture(s, props.textureID, uv);
        albedo = albedo * rb_mk3(t.x, t.y, t.z);
    }
    vec3 wn = h.worldNormal;
    if (props.normalMapTexID >= 0) {
        vec4 t = sample_texture(s, props.normalMapTexID, uv);
        vec3 tn = rb_mk3(t.x * 2.0f - 1.0f, t.y * 2.0f - 1.0f, t.z * 2.0f - 1.0f);
        tn.y = tn.y * -1.0f;
        wn = rb_normalize(rb_m3_mul(h.tbn, tn));
    }

    if (cannotRefract || reflectivity > rnd(pld.rngState)) {     // short-circuit: no draw on TIR
        pld.rayDirection = fuzzy_reflection(unitDir, wn, props.roughness, pld.rngState);
        pld.color = rb_splat3(1.0f);
        pld.rayOrigin = offset_along_normal(h.worldPosition, h.worldNormalGeometry);
    } else {
        pld.insideDielectric = h.frontFace;
        vec3 refr = rb_refract(unitDir, wn, ri);
        pld.rayDirection = refr + random_unit_vec(pld.rngState) * props.roughness;   // addFuzz, :207-209
        pld.color = albedo;
        pld.rayOrigin = offset_for_dielectric(h.worldPosition, h.worldNormalGeometry, unitDir);
    }

    // mix(0, len, bool): selects len when previouslyInside
    if (previouslyInside) pld.accumulatedDistance += rb_length(h.worldPosition - rayOrigin);
    else pld.accumulatedDistance += 0.0f;

    const bool exiting = previouslyInside && !pld.insideDielectric;
    if (exiting) {
        // Beer's law in metres (the x100 "cm" value at :93-94 is computed but unused)
        float a = rb_exp(-props.absorption * pld.accumulatedDistance);
        pld.color = rb_splat3(1.0f) * a;
        pld.color = pld.color * albedo;
        pld.accumulatedDistance = 0.0f;
    }
    pld.albedo = pld.color;
    pld.emission = rb_mk3(props.emission[0], props.emission[1], props.emission[2]);
    pld.rayHitSky = false; pld.skip = false; pld.materialID = 2;
    pld.surfaceNormal = wn; pld.pdf = 0.0f; pld.tbn = h.tbn; pld.props = &props;
    pld.didRefract = false; pld.eta = 0.0f;
}

static DisneyParams disney_params(const RB200InstanceProperties& p, vec3 albedo, float eta) {
    DisneyParams d;
    d.baseColor = albedo;
    d.specularTint = rb_mk3(p.specularTint[0], p.specularTint[1], p.specularTint[2]);
    d.sheenTint = rb_mk3(p.sheenTint[0], p.sheenTint[1], p.sheenTint[2]);
    d.anisotropic = p.anisotropic; d.roughness = p.roughness; d.subsurface = p.subsurface;
    d.clearcoatGloss = p.clearcoatGloss; d.eta = eta; d.metallic = p.metallic; d.clearcoat = p.clearcoat;
    d.specularTransmission = p.specularTransmission; d.sheen = p.sheen;
    return d;
}

// disney.rchit.glsl:36-198
static void shade_disney(const Scene& s, Payload& pld, const HitInfo& h, const RB200InstanceProperties& props,
                         vec3 rayOrigin, vec3 rayDir) {
    vec3 wn, albedo;
    if (!surface_prologue(s, pld, h, props, rayDir, true, &wn, &albedo)) return;
    const float eta = h.frontFace ? 1.0f / props.ior : props.ior;
    DisneyParams dp = disney_params(props, albedo, eta);
    bool didRefract = false, choseGlass = false;
    const vec3 wi = -rayDir;                       // NOT normalised in the reference
    vec3 wo = disney_sample(h.tbn, dp, wn, wi, &didRefract, &choseGlass, pld.rngState);
    vec3 hv = rb_normalize(wo + wi);
    float pdf;
    vec3 f = disney_eval(h.tbn, dp, didRefract, wn, wi, wo, hv, &pdf);
    float cosI = rb_max(rb_dot(wn, wo), 0.0f);
    pld.color = (f * cosI) / pdf;
    pld.albedo = albedo; pld.pdf = pdf;
    pld.emission = rb_mk3(props.emission[0], props.emission[1], props.emission[2]);
    pld.rayOrigin = offset_for_dielectric(h.worldPosition, h.worldNormalGeometry, wo);
    pld.rayDirection = wo;
    pld.rayHitSky = false; pld.skip = false; pld.materialID = 3;
    pld.surfaceNormal = wn; pld.tbn = h.tbn; pld.props = &props;
    pld.didRefract = false; pld.eta = 0.0f;        // :187-189: eta overwritten with 0
    pld.insideDielectric = choseGlass;
    track_distance(pld, h, rayOrigin);
}

// raytrace.rmiss.glsl:10-19
static void shade_miss(Payload& pld, vec3 rayDir) {
    const float y = rb_normalize(rayDir).y;
    const float t = 0.5f * (y + 1.0f);
    pld.color = rb_mix3(rb_splat3(0.1f), rb_mk3(0.4f, 1.7f, 2.0f), t) * 0.07f;
    pld.albedo = pld.color;
    pld.rayHitSky = true; pld.skip = false;
}
vec3 kat_sky(vec3 dir) { Payload p; shade_miss(p, dir); return p.color; }

// traceRayEXT for the main payload (raytrace.rgen.glsl:110-122): closest hit, then the hit group selected by the
// instance's SBT offset = materialIdx (src/tools/vktools.cpp:483), or the miss shader.
static void trace_main(const Scene& s, Payload& pld, vec3 org, vec3 dir, Counters* cnt) {
    if (cnt) cnt->extendRays++;
    Hit hit = closest_hit(s, org, dir, 10000.0f, false);
    if (!hit.valid) { shade_miss(pld, dir); return; }
    const WorldTri& wt = s.tris[hit.gid];
    const RB200Instance& inst = s.instances[wt.instance];
    const RB200InstanceProperties& props = s.props[inst.instancePropertiesID];
    HitInfo h = hit_info(s, inst, props, wt.primitive, hit.b1, hit.b2, dir);
    switch (inst.materialIdx) {
        case 0: shade_lambertian(s, pld, h, props, org, dir); break;
        case 1: shade_metal(s, pld, h, props, org, dir); break;
        case 2: shade_dielectric(s, pld, h, props, org, dir); break;
        default: shade_disney(s, pld, h, props, org, dir); break;
    }
}

// ---------------------------------------------------------------------------------------------------
// light sampling (shaders/raytrace/nee.h.glsl)
// ---------------------------------------------------------------------------------------------------
struct LightSample { vec3 point, normal, emission; float pdf; bool cullBackface; };

static LightSample random_emissive_point(const Scene& s, const RB200RtPushConsts& pc, uint32_t& rng) {
    // pickEmissiveInstance :52-67
    float u = rnd(rng);
    int lo = 0, hi = (int)s.cdfInstances.size() - 1;
    while (lo < hi) { int mid = (lo + hi) / 2; if (u <= s.cdfInstances[(size_t)mid]) hi = mid; else lo = mid + 1; }
    const uint32_t instanceIdx = (uint32_t)lo;
    const RB200InstanceData& md = s.emissive[instanceIdx];
    // pickEmissiveTriangle :69-84
    u = rnd(rng);
    lo = (int)md.cdfRangeStart; hi = (int)md.cdfRangeEnd;
    while (lo < hi) { int mid = (lo + hi) / 2; if (u <= s.cdfTriangles[(size_t)mid]) hi = mid; else lo = mid + 1; }
    const uint32_t tri = (uint32_t)lo;   // NB: index into cdfTriangles, used directly as the triangle index (:97-105)
    const uint32_t base = 3 * tri + md.indexOffset;
    vec3 v0 = rb_m4_point(md.transform, vtx(s, s.indices[base]));
    vec3 v1 = rb_m4_point(md.transform, vtx(s, s.indices[base + 1]));
    vec3 v2 = rb_m4_point(md.transform, vtx(s, s.indices[base + 2]));
    // randomPointOnTriangle :86-93
    float beta = 1.0f - sqrtf(rnd(rng));
    float gamma = (1.0f - beta) * rnd(rng);
    float alpha = 1.0f - beta - gamma;
    LightSample r;
    r.point = v0 * alpha + v1 * beta + v2 * gamma;
    r.normal = rb_normalize(rb_cross(v1 - v0, v2 - v0));
    r.emission = rb_mk3(md.emission[0], md.emission[1], md.emission[2]);
    r.pdf = (md.weight / pc.totalEmissiveWeight) * (1.0f / md.area);
    r.cullBackface = md.cullBackface != 0u;
    return r;
}

struct vec4r { vec3 rgb; float w; };

// raytrace.rgen.glsl:43-95
static vec4r direct_light(const Scene& s, const RB200RtPushConsts& pc, const Payload& pldIn, vec3 rayIn,
                          uint32_t& rng, Counters* cnt) {
    LightSample target = random_emissive_point(s, pc, rng);
    vec3 toLight = target.point - pldIn.rayOrigin;
    vec3 direction = rb_normalize(toLight);
    float dist = rb_length(toLight);
    float pdf = target.pdf * dist * dist / rb_max(rb_dot(target.normal, -direction), 0.0001f);
    vec4r out; out.w = pdf;
    if (cnt) cnt->shadowRays++;
    if (any_hit(s, pldIn.rayOrigin, direction, dist - 0.001f, false)) { out.rgb = rb_splat3(0.0f); return out; }

    vec3 brdf = rb_splat3(0.0f);
    if (pldIn.materialID == 0) {
        brdf = pldIn.albedo * RB_RCP_PI;
    } else if (pldIn.materialID == 3) {
        vec3 wi = -rayIn;
        vec3 hv = rb_normalize(direction + wi);
        DisneyParams dp = disney_params(*pldIn.props, pldIn.albedo, pldIn.eta);
        float ignorePdf;
        brdf = disney_eval(pldIn.tbn, dp, pldIn.didRefract, pldIn.surfaceNormal, wi, direction, hv, &ignorePdf);
    }
    float cosThetai = rb_dot(pldIn.surfaceNormal, direction);
    cosThetai = target.cullBackface ? rb_max(cosThetai, 0.0f) : fabsf(cosThetai);
    float geomNum = rb_dot(target.normal, -direction);
    geomNum = target.cullBackface ? rb_max(geomNum, 0.0f) : fabsf(geomNum);
    float geom = geomNum / (dist * dist);
    out.rgb = (((target.emission * brdf) * cosThetai) * geom) / target.pdf;
    return out;
}

// known-answer hooks: tests/test_oracle_kat.py restates nee.h.glsl and directLight's Lambertian branch in numpy
// out: point[3], normal[3], emission[3], pdf, cullBackface
void kat_light_sample(const Scene& s, const RB200RtPushConsts& pc, uint32_t* rng, float out[11]) {
    const LightSample t = random_emissive_point(s, pc, *rng);
    out[0] = t.point.x; out[1] = t.point.y; out[2] = t.point.z;
    out[3] = t.normal.x; out[4] = t.normal.y; out[5] = t.normal.z;
    out[6] = t.emission.x; out[7] = t.emission.y; out[8] = t.emission.z;
    out[9] = t.pdf; out[10] = t.cullBackface ? 1.0f : 0.0f;
}
// out: rgb[3], pdf (directLight's vec4) for a Lambertian surface point
void kat_direct_light_lambertian(const Scene& s, const RB200RtPushConsts& pc, const float o[3], const float n[3],
                                 const float albedo[3], uint32_t* rng, float out[4]) {
    Payload pld{};
    pld.rayOrigin = rb_mk3(o[0], o[1], o[2]);
    pld.surfaceNormal = rb_mk3(n[0], n[1], n[2]);
    pld.albedo = rb_mk3(albedo[0], albedo[1], albedo[2]);
    pld.materialID = 0;
    pld.props = nullptr;
    const vec4r r = direct_light(s, pc, pld, rb_mk3(0.0f, 0.0f, -1.0f), *rng, nullptr);
    out[0] = r.rgb.x; out[1] = r.rgb.y; out[2] = r.rgb.z; out[3] = r.w;
}

static inline float power_heuristic(float p1, float p2) { return p1 * p1 / (p1 * p1 + p2 * p2); }  // pdf.h.glsl:4-7
float kat_power_heuristic(float a, float b) { return power_heuristic(a, b); }

// raytrace.rgen.glsl:97-184
static vec3 trace_segments(const Scene& s, const RB200RtPushConsts& pc, bool nee, Payload& pld, vec3 org, vec3 dir,
                           Counters* cnt) {
    pld.insideDielectric = false;
    pld.accumulatedDistance = 0.0f;   // documented deviation: the reference never initialises it per path
    vec3 throughput = rb_splat3(1.0f);
    vec3 radiance = rb_splat3(0.0f);
    bool firstBounce = true, prevSkip = false;

    for (uint32_t seg = 0; seg < pc.maxBounces; seg++) {
        const vec3 rayIn = dir;
        const bool prevInside = pld.insideDielectric;
        trace_main(s, pld, org, dir, cnt);
        const bool leftDielectric = !pld.insideDielectric && prevInside;
        org = pld.rayOrigin; dir = pld.rayDirection;
        if (pld.skip) continue;
        if (pld.rayHitSky) { radiance = radiance + pld.color * throughput; break; }
        if (!pld.insideDielectric) {
            const vec3 indirect = pld.emission;
            const bool skipNEE = !nee || (pld.materialID != 0 && pld.materialID != 3);
            vec4r direct; direct.rgb = rb_splat3(0.0f); direct.w = 0.0f;
            if (!skipNEE) direct = direct_light(s, pc, pld, rayIn, pld.rngState, cnt);
            const float pdfNEE = direct.w, pdfBRDF = pld.pdf;
            float wNEE = 0.0f, wBRDF = 1.0f;
            if (!skipNEE) {
                if (firstBounce || prevSkip || leftDielectric) { wNEE = 1.0f; wBRDF = 1.0f; }
                else if (seg + 1 == pc.maxBounces) { wNEE = 0.0f; wBRDF = power_heuristic(pdfBRDF, pdfNEE); }
                else { wNEE = power_heuristic(pdfNEE, pdfBRDF); wBRDF = power_heuristic(pdfBRDF, pdfNEE); }
            }
            prevSkip = skipNEE;
            vec3 combined = direct.rgb * wNEE + indirect * wBRDF;
            radiance = radiance + combined * throughput;
            throughput = throughput * pld.color;
        }
        firstBounce = false;
    }
    return radiance;
}

// known-answer hook: one traceRayEXT of the bounce loop (intersection + the closest-hit / miss shader) on a caller-
// provided payload state. out[0..2] color, [3..5] albedo, [6..8] rayOrigin, [9..11] rayDirection, [12..14] emission,
// [15..17] surfaceNormal, [18] pdf, [19] accumulatedDistance, [20] eta; flags: bit 0 rayHitSky, 1 skip, 2 insideDielectric,
// 3 didRefract, bits 8.. materialID.
void kat_trace_main(const Scene& s, const float o[3], const float d[3], uint32_t* rngState, int insideDielectric,
                    float accumulatedDistance, float out[21], uint32_t* flags) {
    Payload pld{};
    pld.rngState = *rngState;
    pld.insideDielectric = insideDielectric != 0;
    pld.accumulatedDistance = accumulatedDistance;
    Counters cnt;
    trace_main(s, pld, rb_mk3(o[0], o[1], o[2]), rb_mk3(d[0], d[1], d[2]), &cnt);
    const vec3 v[6] = {pld.color, pld.albedo, pld.rayOrigin, pld.rayDirection, pld.emission, pld.surfaceNormal};
    for (int k = 0; k < 6; k++) { out[3 * k] = v[k].x; out[3 * k + 1] = v[k].y; out[3 * k + 2] = v[k].z; }
    out[18] = pld.pdf; out[19] = pld.accumulatedDistance; out[20] = pld.eta;
    *flags = (pld.rayHitSky ? 1u : 0u) | (pld.skip ? 2u : 0u) | (pld.insideDielectric ? 4u : 0u) | (pld.didRefract ? 8u : 0u) | (pld.materialID << 8);
    *rngState = pld.rngState;
}

// known-answer hook: kat_trace_main plus, where the hit is a non-skipped Lambertian or Disney surface outside glass, the
// directLight call that would follow it (on a COPY of the RNG state: the caller's restatement of the bounce loop decides
// whether that call happens and which state it continues with). direct = rgb + pdf.
void kat_bounce(const Scene& s, const RB200RtPushConsts& pc, const float o[3], const float d[3], uint32_t* rngState,
                int insideDielectric, float accumulatedDistance, float out[21], uint32_t* flags, float direct[4],
                uint32_t* rngAfterDirect) {
    Payload pld{};
    pld.rngState = *rngState;
    pld.insideDielectric = insideDielectric != 0;
    pld.accumulatedDistance = accumulatedDistance;
    Counters cnt;
    const vec3 dir = rb_mk3(d[0], d[1], d[2]);
    trace_main(s, pld, rb_mk3(o[0], o[1], o[2]), dir, &cnt);
    const vec3 v[6] = {pld.color, pld.albedo, pld.rayOrigin, pld.rayDirection, pld.emission, pld.surfaceNormal};
    for (int k = 0; k < 6; k++) { out[3 * k] = v[k].x; out[3 * k + 1] = v[k].y; out[3 * k + 2] = v[k].z; }
    out[18] = pld.pdf; out[19] = pld.accumulatedDistance; out[20] = pld.eta;
    *flags = (pld.rayHitSky ? 1u : 0u) | (pld.skip ? 2u : 0u) | (pld.insideDielectric ? 4u : 0u) | (pld.didRefract ? 8u : 0u) | (pld.materialID << 8);
    *rngState = pld.rngState;
    direct[0] = direct[1] = direct[2] = direct[3] = 0.0f;
    *rngAfterDirect = pld.rngState;
    if (!pld.rayHitSky && !pld.skip && !pld.insideDielectric && (pld.materialID == 0 || pld.materialID == 3)) {
        uint32_t r = pld.rngState;
        const vec4r dl = direct_light(s, pc, pld, dir, r, nullptr);
        direct[0] = dl.rgb.x; direct[1] = dl.rgb.y; direct[2] = dl.rgb.z; direct[3] = dl.w;
        *rngAfterDirect = r;
    }
}

// raytrace.rgen.glsl:34-41
static vec2 random_gaussian(uint32_t& rng) {
    const float u1 = rb_max(1e-5f, rnd(rng));
    const float u2 = rnd(rng);
    const float r = sqrtf(-2.0f * rb_log(u1));
    const float theta = (2.0f * RB_PI) * u2;
    float sn, cs; rb_sincos(theta, &sn, &cs);
    return rb_mk2(r * cs, r * sn);
}

// raytrace.rgen.glsl:194-204 — takes the state BY VALUE: the draws are not consumed by the caller
static vec2 random_in_unit_hexagon(uint32_t rng) {
    const float sqrt3 = 1.73205080757f;
    vec2 p;
    do {
        p.x = 2.0f * rnd(rng) - 1.0f;
        p.y = (rnd(rng) - 0.5f) * sqrt3;
    } while (fabsf(p.y) > (sqrt3 * 0.5f) || (sqrt3 * fabsf(p.x) + fabsf(p.y)) > sqrt3);
    return p;
}

static inline void m4_mul_v4(const float* m, const float v[4], float out[4]) {
    for (int r = 0; r < 4; r++) out[r] = m[r] * v[0] + m[4 + r] * v[1] + m[8 + r] * v[2] + m[12 + r] * v[3];
}

// raytrace.rgen.glsl:206-247
void starting_ray(const RB200RtPushConsts& pc, float px, float py, float resx, float resy, uint32_t& rng,
                  vec3* origin_out, vec3* dir_out) {
    vec2 g = random_gaussian(rng);
    float cx = px + 0.5f + 0.375f * g.x;
    float cy = py + 0.5f + 0.375f * g.y;
    float ndcx = (cx / resx) * 2.0f - 1.0f;
    float ndcy = -((cy / resy) * 2.0f - 1.0f);
    const float clip[4] = {ndcx, ndcy, -1.0f, 1.0f};
    float view[4]; m4_mul_v4(pc.invProjection, clip, view);
    vec3 viewDir = rb_normalize(rb_mk3(view[0] / view[3], view[1] / view[3], view[2] / view[3]));
    const float vd[4] = {viewDir.x, viewDir.y, viewDir.z, 0.0f};
    float wd[4]; m4_mul_v4(pc.invView, vd, wd);
    vec3 rayDirection = rb_normalize(rb_mk3(wd[0], wd[1], wd[2]));
    vec3 origin = rb_mk3(pc.invView[12], pc.invView[13], pc.invView[14]);
    vec3 focalPoint = origin + rayDirection * pc.focusDist;
    vec2 hex = random_in_unit_hexagon(rng);
    vec2 lens = hex * pc.defocusMultiplier;
    vec3 right = rb_normalize(rb_mk3(pc.invView[0], pc.invView[1], pc.invView[2]));
    vec3 up = rb_normalize(rb_mk3(pc.invView[4], pc.invView[5], pc.invView[6]));
    vec3 offset = right * lens.x + up * lens.y;
    vec3 newOrigin = origin + offset;
    *origin_out = newOrigin;
    *dir_out = rb_normalize(focalPoint - newOrigin);
}

// raytrace.rgen.glsl:250-285 for rows y0, y0+ystep, ... < y1
void render_rows(const Scene& s, uint32_t W, uint32_t H, uint32_t flags, const RB200RtPushConsts& pc,
                 float* hdr, uint32_t y0, uint32_t y1, uint32_t ystep, Counters* counters, const TilePartition& tiles) {
    const bool nee = (flags & RB200_FLAG_NEE) != 0;
    const bool sumMode = (flags & RB200_FLAG_ACCUM_SUM) != 0;
    const uint32_t tilesX = tiles.size ? (W + tiles.size - 1) / tiles.size : 0;
    for (uint32_t y = y0; y < y1; y += ystep) {
        for (uint32_t x = 0; x < W; x++) {
            // interleaved-tile partition: a rank renders the tiles whose row-major index is congruent
            // to its rank and leaves every other pixel untouched
            if (tiles.count > 1 && ((y / tiles.size) * tilesX + x / tiles.size) % tiles.count != tiles.rank) continue;
            Payload pld;
            pld.rngState = (pc.sampleBatch * H + y) * W + x;
            int actual = 0;
            vec3 sum = rb_splat3(0.0f);
            for (uint32_t sidx = 0; sidx < pc.samplesPerPixel; sidx++) {
                vec3 o, d;
                starting_ray(pc, (float)x, (float)y, (float)W, (float)H, pld.rngState, &o, &d);
                if (counters) counters->paths++;
                vec3 c = trace_segments(s, pc, nee, pld, o, d, counters);
                c = rb_clamp3_keepnan(c, 0.0f, pc.directClamp);
                if (rb_anynan3(c)) continue;
                actual++;
                sum = sum + c;
            }
            float* px = &hdr[4 * ((size_t)y * W + x)];
            if (actual == 0) {
                // documented deviation: the reference writes 0/0 here and poisons the pixel for ever.
                // Here: batch 0 writes black, later batches keep the previous value; sum mode adds nothing.
                if (!sumMode && pc.sampleBatch == 0) { px[0] = px[1] = px[2] = 0.0f; px[3] = 1.0f; }
                continue;
            }
            vec3 fin = sum / (float)actual;
            if (sumMode) {
                px[0] += fin.x; px[1] += fin.y; px[2] += fin.z; px[3] = 1.0f;
            } else {
                if (pc.sampleBatch > 0) {
                    vec3 prev = rb_mk3(px[0], px[1], px[2]);
                    fin = (prev * (float)pc.sampleBatch + fin) / (float)(pc.sampleBatch + 1u);
                }
                px[0] = fin.x; px[1] = fin.y; px[2] = fin.z; px[3] = 1.0f;
            }
        }
    }
}

} // namespace oracle
