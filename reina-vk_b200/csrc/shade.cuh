// shade.cuh — device-side shading: hit reconstruction, the four material models, sky, light sampling.
//
// CUDA counterparts of the reference's closest-hit / miss shaders and light sampler:
//   hit reconstruction       shaders/raytrace/closestHitCommon.h.glsl:52-148
//   skip / offsets           closestHitCommon.h.glsl:156-193, dielectric.rchit.glsl:14-38
//   lambertian               shaders/raytrace/lambertian.rchit.glsl:11-78
//   metal                    shaders/raytrace/metal.rchit.glsl:7-70
//   dielectric               shaders/raytrace/dielectric.rchit.glsl:40-113
//   disney                   shaders/raytrace/disney.rchit.glsl:36-198 + brdfDisney.h.glsl
//   sky                      shaders/raytrace/raytrace.rmiss.glsl:10-19
//   light sampling           shaders/raytrace/nee.h.glsl:52-124, raytrace.rgen.glsl:43-95
// The arithmetic is written in the evaluation order the CPU oracle uses so that both produce the same bits
// (compiled with -fmad=false; explicit fmaf only inside the shared elementary layer).
#pragma once
#include "common.cuh"

namespace rb200 {

struct Surf {
    rb_v3 worldPosition, worldNormal, worldNormalGeometry;
    rb_v2 uv;
    bool frontFace;
    rb_m3 tbn;
};

__device__ __forceinline__ rb_v3 ld_vertex(const DeviceScene& S, uint32_t i) {
    float4 v = __ldg(&S.vertices[i]);
    return rb_mk3(v.x, v.y, v.z);
}
__device__ __forceinline__ rb_v3 ld_tbn_col(const DeviceScene& S, uint32_t i, int col) {
    const float* m = S.tbns + 9 * (size_t)i + 3 * col;
    return rb_mk3(__ldg(m), __ldg(m + 1), __ldg(m + 2));
}
__device__ __forceinline__ rb_v3 ld3(const float* p) { return rb_mk3(__ldg(p), __ldg(p + 1), __ldg(p + 2)); }
// two consecutive float4 of a 32-byte aligned read-only record: one 256-bit load on the device (RB_WIDE_RECORDS, common.cuh)
__device__ __forceinline__ void ld_ro2(const float4* p, float4& a, float4& b) {
#if defined(__CUDA_ARCH__) && RB_WIDE_RECORDS
    asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
#else
    a = __ldg(p); b = __ldg(p + 1);
#endif
}

// RGBA8 UNORM texel fetch through the texture unit (point sampled); the bilinear REPEAT filter of the reference's
// sampler (src/tools/vktools.cpp:765-788) is applied in fp32 so that it is reproducible bit for bit.
__device__ __forceinline__ int wrapi(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }

__device__ __forceinline__ float4 sample_texture(const DeviceScene& S, int id, rb_v2 uv) {
    const cudaTextureObject_t tex = S.textures[id];
    const uint2 sz = S.texSizes[id];
    const float x = uv.x * (float)sz.x - 0.5f;
    const float y = uv.y * (float)sz.y - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float ax = x - fx, ay = y - fy;
    // a coordinate no int32 can hold (parallax steps at grazing angles, NaN) addresses texel 0 with weight 0: cvt
    // saturates on the GPU and does not on the host, so the rule is spelled out (same rule in the oracle)
    if (!(fabsf(fx) < 1073741824.0f)) { fx = 0.0f; ax = 0.0f; }
    if (!(fabsf(fy) < 1073741824.0f)) { fy = 0.0f; ay = 0.0f; }
    const int x0 = wrapi((int)fx, (int)sz.x), y0 = wrapi((int)fy, (int)sz.y);
    const int x1 = wrapi(x0 + 1, (int)sz.x), y1 = wrapi(y0 + 1, (int)sz.y);
    const uchar4 p00 = tex2D<uchar4>(tex, x0 + 0.5f, y0 + 0.5f);
    const uchar4 p10 = tex2D<uchar4>(tex, x1 + 0.5f, y0 + 0.5f);
    const uchar4 p01 = tex2D<uchar4>(tex, x0 + 0.5f, y1 + 0.5f);
    const uchar4 p11 = tex2D<uchar4>(tex, x1 + 0.5f, y1 + 0.5f);
    float4 r;
#define RB_BILERP(ch)                                                                     \
    {                                                                                     \
        float c00 = (float)p00.ch / 255.0f, c10 = (float)p10.ch / 255.0f;                 \
        float c01 = (float)p01.ch / 255.0f, c11 = (float)p11.ch / 255.0f;                 \
        float top = c00 * (1.0f - ax) + c10 * ax;                                         \
        float bot = c01 * (1.0f - ax) + c11 * ax;                                         \
        r.ch = top * (1.0f - ay) + bot * ay;                                              \
    }
    RB_BILERP(x) RB_BILERP(y) RB_BILERP(z) RB_BILERP(w)
#undef RB_BILERP
    return r;
}

// red channel only (height maps), same arithmetic as sample_texture's .x
__device__ __forceinline__ float sample_texture_r(const cudaTextureObject_t tex, const uint2 sz, float u, float v) {
    const float x = u * (float)sz.x - 0.5f;
    const float y = v * (float)sz.y - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float ax = x - fx, ay = y - fy;
    if (!(fabsf(fx) < 1073741824.0f)) { fx = 0.0f; ax = 0.0f; }
    if (!(fabsf(fy) < 1073741824.0f)) { fy = 0.0f; ay = 0.0f; }
    const int x0 = wrapi((int)fx, (int)sz.x), y0 = wrapi((int)fy, (int)sz.y);
    const int x1 = wrapi(x0 + 1, (int)sz.x), y1 = wrapi(y0 + 1, (int)sz.y);
    const float c00 = (float)tex2D<uchar4>(tex, x0 + 0.5f, y0 + 0.5f).x / 255.0f;
    const float c10 = (float)tex2D<uchar4>(tex, x1 + 0.5f, y0 + 0.5f).x / 255.0f;
    const float c01 = (float)tex2D<uchar4>(tex, x0 + 0.5f, y1 + 0.5f).x / 255.0f;
    const float c11 = (float)tex2D<uchar4>(tex, x1 + 0.5f, y1 + 0.5f).x / 255.0f;
    const float top = c00 * (1.0f - ax) + c10 * ax;
    const float bot = c01 * (1.0f - ax) + c11 * ax;
    return top * (1.0f - ay) + bot * ay;
}

// texutils.h.glsl:4-41 — parallax mapping: march the view ray through 64..512 depth layers of the height map (more
// layers at grazing angles), stop at the first layer below the surface, interpolate between the last two layers.
// Every lane walks its own number of layers (at most 0.2 * 512 + 1 steps, the height is at most heightScale).
__device__ __noinline__ rb_v2 bump_mapping_search(const cudaTextureObject_t tex, const uint2 sz, rb_v2 uv, rb_v3 V) {
    const float heightScale = 0.2f;
    if (V.z <= 0.0f) return uv;
    const float numLayers = rb_mix(512.0f, 64.0f, rb_clamp(V.z, 0.0f, 1.0f));
    const float layerDepth = 1.0f / numLayers;
    const rb_v2 delta = rb_mk2(((V.x / V.z) * heightScale) / numLayers, ((V.y / V.z) * heightScale) / numLayers);
    rb_v2 curr = uv;
    float depthSum = 0.0f;
    float h = sample_texture_r(tex, sz, curr.x, curr.y) * heightScale;
    while (depthSum < h) {
        curr = curr - delta;
        depthSum += layerDepth;
        h = sample_texture_r(tex, sz, curr.x, curr.y) * heightScale;
    }
    const rb_v2 prev = curr + delta;
    const float hPrev = sample_texture_r(tex, sz, prev.x, prev.y) * heightScale;
    const float after = h - depthSum;
    const float before = hPrev - (depthSum - layerDepth);
    const float weight = after / (after - before);
    return rb_mk2(rb_mix(prev.x, curr.x, weight), rb_mix(prev.y, curr.y, weight));
}
__device__ __forceinline__ rb_v2 bump_mapping(const DeviceScene& S, rb_v2 uv, rb_v3 rayIn, const rb_m3& T, int heightMap) {
    return bump_mapping_search(S.textures[heightMap], S.texSizes[heightMap], uv, rb_normalize(rb_m3_tmul(T, -rayIn)));
}

// closestHitCommon.h.glsl:52-148
template <bool NEED_TBN>
__device__ __forceinline__ void hit_info(const DeviceScene& S, const RB200Instance* inst, const RB200InstanceProperties* props,
                                         uint32_t prim, float a1, float a2, rb_v3 rayDir, Surf& r) {
    const uint32_t base = 3u * prim + __ldg(&props->indicesOffset);
    const uint32_t i0 = __ldg(&S.indices[base]), i1 = __ldg(&S.indices[base + 1]), i2 = __ldg(&S.indices[base + 2]);
    const rb_v3 v0 = ld_vertex(S, i0), v1 = ld_vertex(S, i1), v2 = ld_vertex(S, i2);
    const float bx = 1.0f - a1 - a2, by = a1, bz = a2;
    float M16[16];
#pragma unroll
    for (int k = 0; k < 16; k++) M16[k] = __ldg(&inst->transform[k]);

    const rb_v3 objectPosition = v0 * bx + v1 * by + v2 * bz;
    r.worldPosition = rb_m4_point(M16, objectPosition);
    const rb_v3 ngObj = rb_normalize(rb_cross(v1 - v0, v2 - v0));
    const uint32_t tb = 3u * prim + __ldg(&props->tbnsIndicesOffset);
    const bool interp = __ldg(&props->interpNormals) != 0u;
    uint32_t t0 = 0, t1 = 0, t2 = 0;
    if (interp || NEED_TBN) { t0 = __ldg(&S.tbnIndices[tb]); t1 = __ldg(&S.tbnIndices[tb + 1]); t2 = __ldg(&S.tbnIndices[tb + 2]); }
    rb_v3 c2_0, c2_1, c2_2;
    if (interp || NEED_TBN) { c2_0 = ld_tbn_col(S, t0, 2); c2_1 = ld_tbn_col(S, t1, 2); c2_2 = ld_tbn_col(S, t2, 2); }
    rb_v3 nObj;
    if (!interp) nObj = ngObj;
    else nObj = rb_normalize(rb_normalize(c2_0) * bx + rb_normalize(c2_1) * by + rb_normalize(c2_2) * bz);

    const uint32_t texOff = __ldg(&props->texIndicesOffset);
    if (texOff == 0xFFFFFFFFu) {
        r.uv = rb_mk2(0.0f, 0.0f);
    } else {
        const uint32_t xb = 3u * prim + texOff;
        const float2 q0 = __ldg(&S.texCoords[__ldg(&S.texIndices[xb])]);
        const float2 q1 = __ldg(&S.texCoords[__ldg(&S.texIndices[xb + 1])]);
        const float2 q2 = __ldg(&S.texCoords[__ldg(&S.texIndices[xb + 2])]);
        r.uv = rb_mk2(q0.x, q0.y) * bx + rb_mk2(q1.x, q1.y) * by + rb_mk2(q2.x, q2.y) * bz;
    }

    const rb_m3 M = rb_m4_upper3(M16);
    r.worldNormal = rb_normalize(rb_m3_mul(M, nObj));
    r.worldNormalGeometry = rb_normalize(rb_m3_mul(M, ngObj));
    r.frontFace = rb_dot(rayDir, r.worldNormalGeometry) < 0.0f;
    r.worldNormal = rb_faceforward(r.worldNormal, rayDir, r.worldNormalGeometry);
    r.worldNormalGeometry = rb_faceforward(r.worldNormalGeometry, rayDir, r.worldNormalGeometry);

    if (NEED_TBN) {
        rb_v3 tangent = rb_normalize(ld_tbn_col(S, t0, 0) * bx + ld_tbn_col(S, t1, 0) * by + ld_tbn_col(S, t2, 0) * bz);
        rb_v3 bitangent = rb_normalize(ld_tbn_col(S, t0, 1) * bx + ld_tbn_col(S, t1, 1) * by + ld_tbn_col(S, t2, 1) * bz);
        rb_v3 normal = rb_normalize(c2_0 * bx + c2_1 * by + c2_2 * bz);
        const rb_m3 Nm = rb_m3_inverse_transpose(M);
        rb_v3 worldT = rb_normalize(rb_m3_mul(M, tangent));
        rb_v3 worldB = rb_normalize(rb_m3_mul(M, bitangent));
        rb_v3 worldN = rb_normalize(rb_m3_mul(Nm, normal));
        worldT = rb_normalize(worldT - worldN * rb_dot(worldN, worldT));
        worldB = rb_normalize(worldB - worldN * rb_dot(worldN, worldB));
        r.tbn.c0 = worldT;
        r.tbn.c1 = worldB * -1.0f;
        r.tbn.c2 = worldN * (r.frontFace ? 1.0f : -1.0f);
    }
}

#if RB_SHADE_RECORDS
// The same reconstruction from the per-triangle shading record (DeviceScene::shadeBase / shadeFrame): identical
// floats, identical arithmetic and order, one or two contiguous reads instead of a three-level gather.
// `tri` is the triangle's slot in leaf order (RayHit::tri), which k_extend stores in the hit record.
template <bool NEED_TBN>
__device__ __forceinline__ void hit_info_rec(const DeviceScene& S, const RB200Instance* inst, const RB200InstanceProperties* props,
                                             uint32_t tri, float a1, float a2, rb_v3 rayDir, Surf& r) {
    const float4* sb = S.shadeBase + 4 * (size_t)tri;
#if RB_WIDE_RECORDS
    float4 r0, r1, r2, r3;
    ld_ro2(sb, r0, r1); ld_ro2(sb + 2, r2, r3);
#else
    const float4 r0 = __ldg(sb), r1 = __ldg(sb + 1), r2 = __ldg(sb + 2), r3 = __ldg(sb + 3);
#endif
    const rb_v3 v0 = rb_mk3(r0.x, r0.y, r0.z), v1 = rb_mk3(r1.x, r1.y, r1.z), v2 = rb_mk3(r2.x, r2.y, r2.z);
    const float bx = 1.0f - a1 - a2, by = a1, bz = a2;
    float M16[16];
#pragma unroll
    for (int k = 0; k < 16; k++) M16[k] = __ldg(&inst->transform[k]);

    const rb_v3 objectPosition = v0 * bx + v1 * by + v2 * bz;
    r.worldPosition = rb_m4_point(M16, objectPosition);
    const rb_v3 ngObj = rb_normalize(rb_cross(v1 - v0, v2 - v0));
    const bool interp = __ldg(&props->interpNormals) != 0u;
    float4 f0, f1, f2, f3, f4, f5, f6;
    rb_v3 c2_0, c2_1, c2_2;
    if (interp || NEED_TBN) {
        const float4* sf = S.shadeFrame + 8 * (size_t)tri;
#if RB_WIDE_RECORDS
        if (NEED_TBN) { float4 f7; ld_ro2(sf, f0, f1); ld_ro2(sf + 2, f2, f3); ld_ro2(sf + 4, f4, f5); ld_ro2(sf + 6, f6, f7); }
        else { ld_ro2(sf, f0, f1); f2 = __ldg(sf + 2); }
#else
        f0 = __ldg(sf); f1 = __ldg(sf + 1); f2 = __ldg(sf + 2);
        if (NEED_TBN) { f3 = __ldg(sf + 3); f4 = __ldg(sf + 4); f5 = __ldg(sf + 5); f6 = __ldg(sf + 6); }
#endif
        c2_0 = rb_mk3(f0.x, f0.y, f0.z); c2_1 = rb_mk3(f1.x, f1.y, f1.z); c2_2 = rb_mk3(f2.x, f2.y, f2.z);
    }
    rb_v3 nObj;
    if (!interp) nObj = ngObj;
    else nObj = rb_normalize(rb_normalize(c2_0) * bx + rb_normalize(c2_1) * by + rb_normalize(c2_2) * bz);

    if (__ldg(&props->texIndicesOffset) == 0xFFFFFFFFu) r.uv = rb_mk2(0.0f, 0.0f);
    else r.uv = rb_mk2(r0.w, r1.w) * bx + rb_mk2(r2.w, r3.x) * by + rb_mk2(r3.y, r3.z) * bz;

    const rb_m3 M = rb_m4_upper3(M16);
    r.worldNormal = rb_normalize(rb_m3_mul(M, nObj));
    r.worldNormalGeometry = rb_normalize(rb_m3_mul(M, ngObj));
    r.frontFace = rb_dot(rayDir, r.worldNormalGeometry) < 0.0f;
    r.worldNormal = rb_faceforward(r.worldNormal, rayDir, r.worldNormalGeometry);
    r.worldNormalGeometry = rb_faceforward(r.worldNormalGeometry, rayDir, r.worldNormalGeometry);

    if (NEED_TBN) {
        const rb_v3 t0 = rb_mk3(f0.w, f1.w, f2.w), t1 = rb_mk3(f3.x, f3.y, f3.z), t2 = rb_mk3(f4.x, f4.y, f4.z);
        const rb_v3 b0 = rb_mk3(f3.w, f4.w, f5.w), b1 = rb_mk3(f5.x, f5.y, f5.z), b2 = rb_mk3(f6.x, f6.y, f6.z);
        rb_v3 tangent = rb_normalize(t0 * bx + t1 * by + t2 * bz);
        rb_v3 bitangent = rb_normalize(b0 * bx + b1 * by + b2 * bz);
        rb_v3 normal = rb_normalize(c2_0 * bx + c2_1 * by + c2_2 * bz);
        const rb_m3 Nm = rb_m3_inverse_transpose(M);
        rb_v3 worldT = rb_normalize(rb_m3_mul(M, tangent));
        rb_v3 worldB = rb_normalize(rb_m3_mul(M, bitangent));
        rb_v3 worldN = rb_normalize(rb_m3_mul(Nm, normal));
        worldT = rb_normalize(worldT - worldN * rb_dot(worldN, worldT));
        worldB = rb_normalize(worldB - worldN * rb_dot(worldN, worldB));
        r.tbn.c0 = worldT;
        r.tbn.c1 = worldB * -1.0f;
        r.tbn.c2 = worldN * (r.frontFace ? 1.0f : -1.0f);
    }
}
#endif

// dielectric.rchit.glsl:14-38 / disney.rchit.glsl:10-34
__device__ __forceinline__ rb_v3 offset_for_dielectric(rb_v3 p, rb_v3 n, rb_v3 rayDir) {
    return rb_offset_along_normal(p, (rb_dot(n, rayDir) < 0.0f) ? -n : n);
}

// closestHitCommon.h.glsl:195-205
__device__ __forceinline__ rb_v3 random_unit_vec(uint32_t& rng) {
    for (;;) {
        float a = rb_random(&rng), b = rb_random(&rng), c = rb_random(&rng);
        rb_v3 v = rb_mk3(2.0f * a - 1.0f, 2.0f * b - 1.0f, 2.0f * c - 1.0f);
        float l2 = rb_dot(v, v);
        if (0.0001f < l2 && l2 < 1.0f) return rb_normalize(v);
    }
}
// closestHitCommon.h.glsl:211-213. The GLSL function takes `inout uint rngState` but draws from the GLOBAL pld.rngState,
// and every caller passes pld.rngState itself: copy-out writes the value the state had BEFORE the call back over the
// advanced one, so the draws of randomUnitVec are not consumed (the next random() repeats them). Confirmed by executing
// the reference's compiled metal.rchit.spv / dielectric.rchit.spv (tests/test_spirv_golden.py).
__device__ __forceinline__ rb_v3 fuzzy_reflection(rb_v3 in, rb_v3 n, float fuzz, uint32_t& rng) {
    const uint32_t stateAtCall = rng;
    rb_v3 r = rb_reflect(rb_normalize(in), rb_normalize(n));
    const rb_v3 out = r + random_unit_vec(rng) * fuzz;
    rng = stateAtCall;
    return out;
}
__device__ __forceinline__ rb_v3 diffuse_reflection(rb_v3 n, uint32_t& rng) {
    const float theta = (2.0f * RB_PI) * rb_random(&rng);
    const float u = 2.0f * rb_random(&rng) - 1.0f;
    const float r = sqrtf(1.0f - u * u);
    float sn, cs; rb_sincos(theta, &sn, &cs);
    return rb_normalize(n + rb_mk3(r * cs, r * sn, u));
}
__device__ __forceinline__ float schlick(float cosine, float refIdx) {
    float r0 = (1.0f - refIdx) / (1.0f + refIdx);
    r0 = r0 * r0;
    return r0 + (1.0f - r0) * rb_pow5(1.0f - cosine);
}

// raytrace.rmiss.glsl:10-19
__device__ __forceinline__ rb_v3 sky_color(rb_v3 rayDir) {
    const float y = rb_normalize(rayDir).y;
    const float t = 0.5f * (y + 1.0f);
    return rb_mix3(rb_splat3(0.1f), rb_mk3(0.4f, 1.7f, 2.0f), t) * 0.07f;
}

// ---------------------------------------------------------------------------------------------------
// Disney BSDF (brdfDisney.h.glsl)
// ---------------------------------------------------------------------------------------------------
struct DisneyP {
    rb_v3 baseColor, specularTint, sheenTint;
    float anisotropic, roughness, subsurface, clearcoatGloss, eta, metallic, clearcoat, specularTransmission, sheen;
};

__device__ __forceinline__ DisneyP load_disney(const RB200InstanceProperties* p, rb_v3 albedo, float eta) {
    DisneyP d;
    d.baseColor = albedo;
    d.specularTint = ld3(p->specularTint);
    d.sheenTint = ld3(p->sheenTint);
    d.anisotropic = __ldg(&p->anisotropic); d.roughness = __ldg(&p->roughness); d.subsurface = __ldg(&p->subsurface);
    d.clearcoatGloss = __ldg(&p->clearcoatGloss); d.eta = eta; d.metallic = __ldg(&p->metallic);
    d.clearcoat = __ldg(&p->clearcoat); d.specularTransmission = __ldg(&p->specularTransmission); d.sheen = __ldg(&p->sheen);
    return d;
}

__device__ __forceinline__ void ggx_alpha(float anisotropic, float roughness, float& ax, float& ay) {
    const float aspect = sqrtf(1.0f - 0.9f * anisotropic);
    ax = rb_max(0.0001f, roughness * roughness / aspect);
    ay = rb_max(0.0001f, roughness * roughness * aspect);
}

__device__ rb_v3 sample_ggx_vndf(rb_v3 V, float ax, float ay, uint32_t& rng) {
    const bool flip = V.z < 0.0f;
    if (flip) V.z = V.z * -1.0f;
    const float r1 = rb_random(&rng);
    const float r2 = rb_random(&rng);
    rb_v3 Vh = rb_normalize(rb_mk3(ax * V.x, ay * V.y, V.z));
    float lensq = Vh.x * Vh.x + Vh.y * Vh.y;
    rb_v3 T1 = lensq > 0.0f ? rb_mk3(-Vh.y, Vh.x, 0.0f) * rb_rsqrt(lensq) : rb_mk3(1.0f, 0.0f, 0.0f);
    rb_v3 T2 = rb_cross(Vh, T1);
    float r = sqrtf(r1);
    float phi = (2.0f * RB_PI) * r2;
    float sp, cp; rb_sincos(phi, &sp, &cp);
    float t1 = r * cp;
    float t2 = r * sp;
    float s = 0.5f * (1.0f + Vh.z);
    t2 = (1.0f - s) * sqrtf(1.0f - t1 * t1) + s * t2;
    rb_v3 Nh = T1 * t1 + T2 * t2 + Vh * sqrtf(rb_max(0.0f, 1.0f - t1 * t1 - t2 * t2));
    if (flip) Nh.z = Nh.z * -1.0f;
    return rb_normalize(rb_mk3(ax * Nh.x, ay * Nh.y, rb_max(0.0f, Nh.z)));
}

__device__ __forceinline__ float d_ggx_aniso(rb_v3 m, float ax, float ay) {
    float NoM = rb_max(m.z, 0.0f);
    float tx = m.x / ax, ty = m.y / ay;
    float inv = 1.0f / (tx * tx + ty * ty + NoM * NoM);
    return inv * inv / ((RB_PI * ax) * ay);
}

__device__ float pdf_ggx_reflection(rb_v3 i, rb_v3 o, float ax, float ay) {
    rb_v3 m = rb_normalize(i + o);
    float ndf = d_ggx_aniso(m, ax, ay);
    float aix = ax * i.x, aiy = ay * i.y;
    float len2 = aix * aix + aiy * aiy;
    float t = sqrtf(len2 + i.z * i.z);
    if (i.z >= 0.0f) {
        float a = rb_clamp(rb_min(ax, ay), 0.0f, 1.0f);
        float s = 1.0f + sqrtf(i.x * i.x + i.y * i.y);
        float a2 = a * a, s2 = s * s;
        float k = (1.0f - a2) * s2 / (s2 + a2 * i.z * i.z);
        return ndf / (2.0f * (k * i.z + t));
    }
    return ndf * (t - i.z) / (2.0f * len2);
}

__device__ rb_v3 sample_gtr1(float alpha, uint32_t& rng) {
    const float r1 = rb_random(&rng);
    const float r2 = rb_random(&rng);
    const float a = rb_max(0.001f, alpha);
    const float a2 = a * a;
    float cosEl = sqrtf((1.0f - rb_exp(r1 * rb_log(a2))) / (1.0f - a2));
    float el = rb_acos(cosEl);
    float az = (2.0f * RB_PI) * r2;
    float se, ce; rb_sincos(el, &se, &ce);
    float sa, ca; rb_sincos(az, &sa, &ca);
    return rb_mk3(se * ca, se * sa, cosEl);
}

__device__ rb_v3 eval_diffuse(const DisneyP& p, rb_v3 n, rb_v3 wi, rb_v3 wo, rb_v3 h) {
    const float hdwo = rb_dot(h, wo);
    const float ndwi = rb_max(rb_dot(n, wi), 0.0f), ndwo = rb_max(rb_dot(n, wo), 0.0f);
    // FD90 = 0.5 + x, and FD uses (FD90 - 1): the reference's compiler (spirv-opt) merges the two constants into
    // x + (-0.5), which skips the rounding of 0.5 + x (seen in the compiled disney.rchit.spv)
    const float fd90m1 = 2.0f * p.roughness * rb_max(hdwo, 0.0f) * rb_max(hdwo, 0.0f) + -0.5f;
    float fdIn = 1.0f + fd90m1 * rb_pow5(1.0f - ndwi);
    float fdOut = 1.0f + fd90m1 * rb_pow5(1.0f - ndwo);
    rb_v3 baseDiffuse = ((p.baseColor * RB_RCP_PI) * fdIn) * fdOut;     // "/ k_pi" is compiled as "* (1 / k_pi)"
    rb_v3 k = (p.baseColor * 1.25f) * RB_INV_PI;
    float fss90 = p.roughness * rb_max(hdwo, 0.0f) * rb_max(hdwo, 0.0f);
    float fssIn = 1.0f + (fss90 - 1.0f) * rb_pow5(1.0f - ndwi);
    float fssOut = 1.0f + (fss90 - 1.0f) * rb_pow5(1.0f - ndwo);
    float third = 1.0f / (ndwi + ndwo) - 0.5f;
    rb_v3 fss = k * (fssIn * fssOut * third + 0.5f);
    return rb_mix3(baseDiffuse, fss, p.subsurface);
}

__device__ rb_v3 sample_diffuse_lobe(rb_v3 n, uint32_t& rng) {
    float xi1 = rb_random(&rng);
    float xi2 = rb_random(&rng);
    float r = sqrtf(xi1);
    float phi = (2.0f * RB_PI) * xi2;
    float sp, cp; rb_sincos(phi, &sp, &cp);
    float x = r * cp, y = r * sp;
    float z = sqrtf(rb_max(0.0f, 1.0f - xi1));
    rb_v3 t = fabsf(n.x) < 0.5f ? rb_normalize(rb_cross(n, rb_mk3(1.0f, 0.0f, 0.0f)))
                                : rb_normalize(rb_cross(n, rb_mk3(0.0f, 1.0f, 0.0f)));
    rb_v3 b = rb_cross(n, t);
    return rb_normalize(t * x + b * y + n * z);
}

__device__ __forceinline__ float eval_r0(float ior) { return (ior - 1.0f) * (ior - 1.0f) / ((ior + 1.0f) * (ior + 1.0f)); }
__device__ __forceinline__ float luminance(rb_v3 c) { return rb_dot(c, rb_mk3(0.2126f, 0.7152f, 0.0722f)); }

__device__ __forceinline__ rb_v3 eval_fm(rb_v3 baseColor, rb_v3 h, rb_v3 wo, float specular, rb_v3 specularTint,
                                         float metallic, float eta) {
    float lum = luminance(baseColor);
    rb_v3 ctint = lum > 0.0f ? baseColor / lum : rb_splat3(1.0f);
    rb_v3 ks = (rb_splat3(1.0f) - ctint) + specularTint * ctint;
    rb_v3 c0 = ks * (specular * eval_r0(eta) * (1.0f - metallic)) + baseColor * metallic;
    return c0 + (rb_splat3(1.0f) - c0) * rb_pow5(1.0f - fabsf(rb_dot(h, wo)));
}
__device__ __forceinline__ float eval_dm(rb_v3 hl, float ax, float ay) {
    float constant = (RB_PI * ax) * ay;
    float inner = rb_sq(hl.x) / rb_sq(ax) + rb_sq(hl.y) / rb_sq(ay) + rb_sq(hl.z);
    return 1.0f / (constant * rb_sq(inner));
}
__device__ __forceinline__ float smith_g(rb_v3 wl, float ax, float ay) {
    float sq = sqrtf(1.0f + (rb_sq(wl.x * ax) + rb_sq(wl.y * ay)) / rb_sq(wl.z));
    float lambda = (sq - 1.0f) / 2.0f;
    return 1.0f / (1.0f + lambda);
}

__device__ rb_v3 eval_metal(const rb_m3& tbn, rb_v3 baseColor, float ax, float ay, rb_v3 n, rb_v3 wi, rb_v3 wo, rb_v3 h,
                            float specular, rb_v3 specularTint, float metallic, float eta) {
    rb_v3 fm = eval_fm(baseColor, h, wo, specular, specularTint, metallic, eta);
    rb_v3 wiT = rb_normalize(rb_m3_tmul(tbn, wi));
    rb_v3 woT = rb_normalize(rb_m3_tmul(tbn, wo));
    rb_v3 hT = rb_normalize(rb_m3_tmul(tbn, h));
    float dm = eval_dm(hT, ax, ay);
    float gm = smith_g(wiT, ax, ay) * smith_g(woT, ax, ay);
    float ndwi = fabsf(rb_dot(n, wi)), ndwo = fabsf(rb_dot(n, wo));
    return ((fm * dm) * gm) / (4.0f * ndwi * ndwo);
}

__device__ __forceinline__ float separable_smith_g1(rb_v3 w, float a) {
    float a2 = a * a;
    float c = w.z;
    return 2.0f / (1.0f + sqrtf(a2 + (1.0f - a2) * c * c));
}
__device__ __forceinline__ float clearcoat_alpha(float gloss) { return (1.0f - gloss) * 0.1f + gloss * 0.001f; }
__device__ __forceinline__ float eval_dc(float ag, rb_v3 hl) {
    float num = ag * ag - 1.0f;
    float den = (RB_PI * rb_log(ag * ag)) * (1.0f + (ag * ag - 1.0f) * (hl.z * hl.z));
    return num / den;
}

__device__ __forceinline__ float smith_g_aniso(float ndv, float vdx, float vdy, float ax, float ay) {
    float a = vdx * ax, b = vdy * ay, c = ndv;
    return (2.0f * ndv) / (ndv + sqrtf(a * a + b * b + c * c));
}

__device__ rb_v3 eval_microfacet_refraction(rb_v3 baseColor, float ax, float ay, float eta, rb_v3 V, rb_v3 L, rb_v3 H, float* pdf) {
    *pdf = 0.0f;
    if (L.z >= 0.0f) return rb_splat3(0.0f);
    float ldh = rb_dot(L, H), vdh = rb_dot(V, H);
    float D = eval_dm(H, ax, ay);
    float G1 = smith_g_aniso(fabsf(V.z), V.x, V.y, ax, ay);
    float G2 = G1 * smith_g_aniso(fabsf(L.z), L.x, L.y, ax, ay);
    float denom = ldh + vdh * eta;
    denom = denom * denom;
    float eta2 = eta * eta;
    float jac = fabsf(ldh) * eta2 / denom;
    *pdf = G1 * rb_max(0.0f, vdh) * D * jac / V.z;
    float F = schlick(rb_dot(V, H), eta);
    rb_v3 sq = rb_mk3(sqrtf(baseColor.x), sqrtf(baseColor.y), sqrtf(baseColor.z));
    return (((((sq * (1.0f - F)) * D) * G2) * fabsf(vdh)) * jac) / fabsf(L.z * V.z);
}

// sampleDisney, brdfDisney.h.glsl:609-654
__device__ rb_v3 disney_sample(const rb_m3& tbn, const DisneyP& p, rb_v3 n, rb_v3 wi, bool* didRefract, bool* choseGlass,
                               uint32_t& rng) {
    const float diffuseWt = (1.0f - p.specularTransmission) * (1.0f - p.metallic);
    const float metalWt = p.metallic;
    const float clearcoatWt = 0.25f * p.clearcoat;
    const float glassWt = (1.0f - p.metallic) * p.specularTransmission;
    const float c0 = diffuseWt, c1 = c0 + metalWt, c2 = c1 + clearcoatWt, c3 = c2 + glassWt;
    *didRefract = false; *choseGlass = false;
    const float r = rb_random(&rng) * c3;
    if (r < c0) return sample_diffuse_lobe(n, rng);
    float ax, ay; ggx_alpha(p.anisotropic, p.roughness, ax, ay);
    if (r < c1) {
        rb_v3 h = sample_ggx_vndf(rb_m3_tmul(tbn, wi), ax, ay, rng);
        h = rb_normalize(rb_m3_mul(tbn, h));
        return rb_reflect(-wi, h);
    }
    if (r < c2) {
        rb_v3 h = rb_normalize(sample_gtr1(clearcoat_alpha(p.clearcoatGloss), rng));
        h = rb_normalize(rb_m3_mul(tbn, h));
        return rb_normalize(rb_reflect(-wi, h));
    }
    *choseGlass = true;
    rb_v3 wiT = rb_m3_tmul(tbn, wi);
    rb_v3 hT = sample_ggx_vndf(wiT, ax, ay, rng);
    float cosTheta = rb_dot(wiT, hT);
    rb_v3 hW = rb_normalize(rb_m3_mul(tbn, hT));
    float reflectivity = schlick(cosTheta, p.eta);
    float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
    bool cannotRefract = p.eta * sinTheta > 1.0f;
    if (cannotRefract || reflectivity > rb_random(&rng)) { *didRefract = false; return rb_reflect(-wi, hW); }
    *didRefract = true;
    return rb_refract(-wi, hW, p.eta);
}

// evalDisney, brdfDisney.h.glsl:656-711 — all lobes, all pdfs, every time
__device__ rb_v3 disney_eval(const rb_m3& tbn, const DisneyP& p, bool didRefract, rb_v3 n, rb_v3 wi, rb_v3 wo, rb_v3 h, float* pdf) {
    const float diffuseWt = (1.0f - p.specularTransmission) * (1.0f - p.metallic);
    const float metalWt = p.metallic;
    const float clearcoatWt = 0.25f * p.clearcoat;
    const float glassWt = (1.0f - p.metallic) * p.specularTransmission;
    const float wtSum = diffuseWt + metalWt + glassWt;
    float ax, ay; ggx_alpha(p.anisotropic, p.roughness, ax, ay);
    const rb_v3 wiT = rb_m3_tmul(tbn, wi), woT = rb_m3_tmul(tbn, wo), hT = rb_m3_tmul(tbn, h);

    const rb_v3 fdiffuse = eval_diffuse(p, n, wi, wo, h);
    const float ndwo = rb_dot(n, wo);
    const float diffusePdf = ndwo <= 0.0f ? 0.0f : ndwo * RB_INV_PI;
    // sheen :587-594
    rb_v3 fsheen;
    {
        float lum = luminance(p.baseColor);
        rb_v3 ctint = lum > 0.0f ? p.baseColor / lum : rb_splat3(1.0f);
        rb_v3 csheen = rb_mix3v(rb_splat3(1.0f), ctint, p.sheenTint);
        fsheen = (csheen * rb_pow5(1.0f - rb_max(rb_dot(h, wo), 0.0f))) * rb_max(ndwo, 0.0f);
    }
    const rb_v3 fmetal = eval_metal(tbn, p.baseColor, ax, ay, n, wi, wo, h, p.specularTransmission, p.specularTint, p.metallic, p.eta);
    const float metalPdf = pdf_ggx_reflection(wiT, woT, ax, ay);
    // clearcoat :341-383
    rb_v3 fclear; float clearPdf;
    {
        float ag = clearcoat_alpha(p.clearcoatGloss);
        float r0 = eval_r0(1.5f);
        float fc = r0 + (1.0f - r0) * rb_pow5(1.0f - rb_dot(h, wo));
        float gc = separable_smith_g1(wiT, 0.25f) * separable_smith_g1(woT, 0.25f);
        float dc = eval_dc(ag, hT);
        fclear = rb_splat3(0.25f * fc * gc * dc);
        if (wiT.z <= 0.0f || woT.z <= 0.0f) clearPdf = 0.0f;
        else clearPdf = (dc * hT.z) / (4.0f * fabsf(rb_dot(woT, hT)));
    }
    // glass :539-581
    rb_v3 fglass; float glassPdf;
    {
        rb_v3 hg = didRefract ? rb_normalize(wo + wi * p.eta) : rb_normalize(wo + wi);
        if (rb_dot(hg, n) < 0.0f) hg = -hg;
        if (didRefract) {
            fglass = eval_microfacet_refraction(p.baseColor, ax, ay, p.eta, wiT, woT, rb_m3_tmul(tbn, hg), &glassPdf);
        } else {
            glassPdf = metalPdf;
            fglass = eval_metal(tbn, p.baseColor, ax, ay, n, wi, wo, hg, 0.0f, rb_splat3(1.0f), 1.0f, p.eta);
        }
    }
    *pdf = diffusePdf * diffuseWt / wtSum + metalPdf * metalWt / wtSum + clearPdf * clearcoatWt + glassPdf * glassWt / wtSum;
    return (fdiffuse + fsheen * p.sheen) * (diffuseWt / wtSum) + fmetal * (metalWt / wtSum) + fclear * clearcoatWt +
           fglass * (glassWt / wtSum);
}

// ---------------------------------------------------------------------------------------------------
// light sampling (nee.h.glsl:52-124)
// ---------------------------------------------------------------------------------------------------
struct LightSample { rb_v3 point, normal, emission; float pdf; bool cullBackface; };

__device__ LightSample random_emissive_point(const DeviceScene& S, float totalEmissiveWeight, uint32_t& rng) {
    float u = rb_random(&rng);
    int lo = 0, hi = (int)S.numCdfInstances - 1;
    while (lo < hi) { int mid = (lo + hi) / 2; if (u <= __ldg(&S.cdfInstances[mid])) hi = mid; else lo = mid + 1; }
    const RB200InstanceData* md = &S.emissive[lo];
    u = rb_random(&rng);
    lo = (int)__ldg(&md->cdfRangeStart); hi = (int)__ldg(&md->cdfRangeEnd);
    while (lo < hi) { int mid = (lo + hi) / 2; if (u <= __ldg(&S.cdfTriangles[mid])) hi = mid; else lo = mid + 1; }
    const uint32_t base = 3u * (uint32_t)lo + __ldg(&md->indexOffset);
    float M16[16];
#pragma unroll
    for (int k = 0; k < 16; k++) M16[k] = __ldg(&md->transform[k]);
    rb_v3 v0 = rb_m4_point(M16, ld_vertex(S, __ldg(&S.indices[base])));
    rb_v3 v1 = rb_m4_point(M16, ld_vertex(S, __ldg(&S.indices[base + 1])));
    rb_v3 v2 = rb_m4_point(M16, ld_vertex(S, __ldg(&S.indices[base + 2])));
    float beta = 1.0f - sqrtf(rb_random(&rng));
    float gamma = (1.0f - beta) * rb_random(&rng);
    float alpha = 1.0f - beta - gamma;
    LightSample r;
    r.point = v0 * alpha + v1 * beta + v2 * gamma;
    r.normal = rb_normalize(rb_cross(v1 - v0, v2 - v0));
    r.emission = ld3(md->emission);
    r.pdf = (__ldg(&md->weight) / totalEmissiveWeight) * (1.0f / __ldg(&md->area));
    r.cullBackface = __ldg(&md->cullBackface) != 0u;
    return r;
}

} // namespace rb200
