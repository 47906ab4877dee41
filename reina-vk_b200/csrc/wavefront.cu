// wavefront.cu — the wavefront path tracer: generate -> { extend -> shade_<material> / miss -> shadow -> finish }*.
//
// Replaces the single vkCmdTraceRaysKHR(W,H,1) of Reina::traceRays (src/Reina.cpp:425-470) whose raygen shader
// runs the whole bounce loop per pixel (shaders/raytrace/raytrace.rgen.glsl:97-184, 250-285). Here a pixel owns a
// path-state slot; every wave advances all live paths by one segment:
//   extend        closest hit for every queued ray, then the slot is binned by the hit instance's material
//   shade_<m>     the closest-hit shader of material m + the raygen bookkeeping that follows traceRayEXT
//   miss          sky radiance, path ends
//   shadow        any-hit visibility for this wave's light samples, folds the NEE term into the path radiance
//   finish        end of path: clamp, drop NaN, add to the pixel's batch sum; next sample of the pixel re-uses the
//                 slot (the RNG stream is one per (pixel, batch) and continues across samples, rgen.glsl:259,264),
//                 the last sample writes the running average into the HDR image (rgen.glsl:277-284)
// Queues are compacted with warp-aggregated atomics; their order is irrelevant to the result because every slot is
// private to its pixel. Draw order of the RNG follows SURVEY.md Appendix A exactly.
#include "context.cuh"
#include "traverse.cuh"
#include "shade.cuh"

namespace rb200 {

static constexpr int BLOCK = 256;
#ifndef RB_TRAV_BLOCK
#define RB_TRAV_BLOCK 256        // threads per block of k_extend / k_shadow (with RB_TRAV_MINBLOCKS: 1024 threads per SM at 64 registers)
#endif
#ifndef RB_TRAV_MINBLOCKS
#define RB_TRAV_MINBLOCKS (1024 / RB_TRAV_BLOCK)      // resident blocks per SM the traversal kernels are compiled for (register cap)
#endif

// ---------------------------------------------------------------------------------------------------
// queue helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void queue_push(uint32_t* __restrict__ q, uint32_t* counter, uint32_t value) {
    const uint32_t active = __activemask();
    const uint32_t lane = threadIdx.x & 31u;
    const int leader = __ffs(active) - 1;
    const uint32_t rank = __popc(active & ((1u << lane) - 1u));
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(counter, (uint32_t)__popc(active));
    base = __shfl_sync(active, base, leader);
    q[base + rank] = value;
}

__device__ __forceinline__ uint32_t queue_reserve(uint32_t* counter) {
    const uint32_t active = __activemask();
    const uint32_t lane = threadIdx.x & 31u;
    const int leader = __ffs(active) - 1;
    const uint32_t rank = __popc(active & ((1u << lane) - 1u));
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(counter, (uint32_t)__popc(active));
    base = __shfl_sync(active, base, leader);
    return base + rank;
}

// ---------------------------------------------------------------------------------------------------
// camera (raytrace.rgen.glsl:34-41, 194-247)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void m4_mul_v4(const float* m, float v0, float v1, float v2, float v3, float out[4]) {
#pragma unroll
    for (int r = 0; r < 4; r++) out[r] = m[r] * v0 + m[4 + r] * v1 + m[8 + r] * v2 + m[12 + r] * v3;
}

__device__ void starting_ray(const RB200RtPushConsts& pc, float px, float py, float resx, float resy, uint32_t& rng,
                             rb_v3& origin, rb_v3& dir) {
    // randomGaussian (:34-41)
    const float u1 = rb_max(1e-5f, rb_random(&rng));
    const float u2 = rb_random(&rng);
    const float r = sqrtf(-2.0f * rb_log(u1));
    const float theta = (2.0f * RB_PI) * u2;
    float sn, cs; rb_sincos(theta, &sn, &cs);
    const float cx = px + 0.5f + 0.375f * (r * cs);
    const float cy = py + 0.5f + 0.375f * (r * sn);
    const float ndcx = (cx / resx) * 2.0f - 1.0f;
    const float ndcy = -((cy / resy) * 2.0f - 1.0f);
    float view[4]; m4_mul_v4(pc.invProjection, ndcx, ndcy, -1.0f, 1.0f, view);
    const rb_v3 viewDir = rb_normalize(rb_mk3(view[0] / view[3], view[1] / view[3], view[2] / view[3]));
    float wd[4]; m4_mul_v4(pc.invView, viewDir.x, viewDir.y, viewDir.z, 0.0f, wd);
    const rb_v3 rayDirection = rb_normalize(rb_mk3(wd[0], wd[1], wd[2]));
    const rb_v3 camPos = rb_mk3(pc.invView[12], pc.invView[13], pc.invView[14]);
    const rb_v3 focalPoint = camPos + rayDirection * pc.focusDist;
    // randomInUnitHexagon works on a COPY of the state (:194): the draws are not consumed
    uint32_t tmp = rng;
    const float sqrt3 = 1.73205080757f;
    float hx, hy;
    do {
        hx = 2.0f * rb_random(&tmp) - 1.0f;
        hy = (rb_random(&tmp) - 0.5f) * sqrt3;
    } while (fabsf(hy) > (sqrt3 * 0.5f) || (sqrt3 * fabsf(hx) + fabsf(hy)) > sqrt3);
    const float lx = hx * pc.defocusMultiplier, ly = hy * pc.defocusMultiplier;
    const rb_v3 right = rb_normalize(rb_mk3(pc.invView[0], pc.invView[1], pc.invView[2]));
    const rb_v3 up = rb_normalize(rb_mk3(pc.invView[4], pc.invView[5], pc.invView[6]));
    const rb_v3 offset = right * lx + up * ly;
    origin = camPos + offset;
    dir = rb_normalize(focalPoint - origin);
}

// start (or restart) the path of a slot: traceSegments' locals (rgen.glsl:98-104)
__device__ __forceinline__ void begin_path(const WaveParams& P, uint32_t slot, uint32_t& rng, uint32_t sampleIdx) {
    const uint32_t x = slot % P.W, y = slot / P.W;
    rb_v3 o, d;
    starting_ray(P.pc, (float)x, (float)y, (float)P.W, (float)P.H, rng, o, d);
    P.rayO[slot] = make_float4(o.x, o.y, o.z, 0.f);
    P.rayD[slot] = make_float4(d.x, d.y, d.z, 0.f);
    P.thr[slot] = make_float4(1.f, 1.f, 1.f, 0.f);     // throughput 1, accumulatedDistance 0 (documented deviation)
    P.rad[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    P.st[slot] = make_uint4(rng, F_FIRST, sampleIdx, 0u);
}

__global__ void __launch_bounds__(BLOCK) k_generate(WaveParams P) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (P.tileCount <= 1u) {
        if (slot == 0) atomicAdd(&P.stats[ST_PATHS], (unsigned long long)P.N);
        if (slot >= P.N) return;
        const uint32_t x = slot % P.W, y = slot / P.W;
        uint32_t rng = (P.pc.sampleBatch * P.H + y) * P.W + x;       // rgen.glsl:259
        P.sum[slot] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0u));
        begin_path(P, slot, rng, 0u);
        P.rayQ[0][slot] = slot;
        if (slot == 0) P.counters[CNT_RAYS] = P.N;
        return;
    }
    // interleaved-tile partition (latency mode): this context traces the pixels of the tiles whose
    // row-major index is congruent to tileRank; the other pixels are not touched (mean.w < 0 makes k_accumulate
    // skip them), so the partial images of all ranks add up to the single-GPU image bit for bit
    if (slot >= P.N) return;
    const uint32_t x = slot % P.W, y = slot / P.W;
    const bool mine = ((y / P.tileSize) * P.tilesX + x / P.tileSize) % P.tileCount == P.tileRank;
    if (!mine) { P.mean[slot] = make_float4(0.f, 0.f, 0.f, -1.f); return; }     // w < 0: not this rank's pixel
    uint32_t rng = (P.pc.sampleBatch * P.H + y) * P.W + x;
    P.sum[slot] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0u));
    begin_path(P, slot, rng, 0u);
    const uint32_t active = __activemask();
    queue_push(P.rayQ[0], &P.counters[CNT_RAYS], slot);
    if ((threadIdx.x & 31u) == (uint32_t)(__ffs(active) - 1)) atomicAdd(&P.stats[ST_PATHS], (unsigned long long)__popc(active));
}

// ---------------------------------------------------------------------------------------------------
// extend
// ---------------------------------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(RB_TRAV_BLOCK, RB_TRAV_MINBLOCKS) k_extend(WaveParams P, int parity) {
    uint32_t* cnt = P.counters + parity * CNT_SET;
    const uint32_t n = cnt[CNT_RAYS];
    const uint32_t* __restrict__ q = P.rayQ[parity];
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&P.stats[ST_EXTEND], (unsigned long long)n);
    uint32_t nodeVisits = 0, triTests = 0;
    extern __shared__ __align__(16) unsigned char rb_dyn_smem[];      // WarpShared<false>[RB_TRAV_BLOCK / 32]: may exceed the 48 KB static limit
    WarpShared<false>* ws = reinterpret_cast<WarpShared<false>*>(rb_dyn_smem);
    trace_queue<false, COUNT>(
        P.S.nodes, P.S.tris, n, &cnt[CNT_CURSOR_EXTEND],
        [&](uint32_t i, rb_v3& o, rb_v3& d, float& tmax) {
            const uint32_t slot = q[i];
            const float4 o4 = P.rayO[slot], d4 = P.rayD[slot];
            o = rb_mk3(o4.x, o4.y, o4.z); d = rb_mk3(d4.x, d4.y, d4.z); tmax = 10000.0f;
        },
        [&](uint32_t i, const RayHit& h) {
            const uint32_t slot = q[i];
            uint32_t bin = 4;
            if (h.tri != 0xFFFFFFFFu) {
                const float4* tp = reinterpret_cast<const float4*>(P.S.tris + h.tri);
#if RB_SHADE_RECORDS
                const uint32_t prim = h.tri;       // the shading kernels read the per-triangle record of this slot
#else
                const uint32_t prim = __float_as_uint(__ldg(tp).w);
#endif
                const uint32_t inst = __float_as_uint(__ldg(tp + 1).w);
                P.hit[slot] = make_uint4(__float_as_uint(h.b1), __float_as_uint(h.b2), prim, inst);
                const uint32_t m = __ldg(&P.S.instances[inst].materialIdx);
                bin = m > 3u ? 3u : m;
            }
            // bin the slot by material (warp-aggregated per bin)
#pragma unroll
            for (uint32_t b = 0; b < 5; b++)
                if (bin == b) queue_push(P.matQ[b], &cnt[CNT_MAT0 + b], slot);
        },
        nodeVisits, triTests, ws[threadIdx.x >> 5]);
    if (COUNT) {
        atomicAdd(&P.stats[ST_NODES], (unsigned long long)nodeVisits);
        atomicAdd(&P.stats[ST_TRIS], (unsigned long long)triTests);
    }
}

// ---------------------------------------------------------------------------------------------------
// shade
// ---------------------------------------------------------------------------------------------------
struct ShadeOut {
    rb_v3 color, albedo, emission, newO, newD, normal;
    float pdf;
    bool skip, inside;
};

// closestHitCommon.h.glsl:185-193
__device__ __forceinline__ void do_skip(ShadeOut& o, const Surf& s, rb_v3 rayDir) {
    o.newO = rb_offset_along_normal(s.worldPosition, -s.worldNormal);
    o.newD = rayDir;
    o.skip = true;
}

// cull / UV wrap + range skip / normal map / albedo texture with stochastic alpha
// (lambertian.rchit.glsl:15-56, metal.rchit.glsl:12-48, disney.rchit.glsl:40-84)
template <bool UV_RANGE_SKIP>
__device__ __forceinline__ bool surface_prologue(const WaveParams& P, const RB200InstanceProperties* props, const Surf& s,
                                                 const rb_m3* tbn, rb_v3 rayDir, uint32_t& rng, ShadeOut& o, rb_v3& wn, rb_v3& col) {
    if (__ldg(&props->cullBackface) != 0u && !s.frontFace) { do_skip(o, s, rayDir); return false; }
    rb_v2 uv = rb_mk2(rb_fract_mod1(s.uv.x), rb_fract_mod1(s.uv.y));
    const int bm = __ldg(&props->bumpMapTexID);
    if (bm >= 0) uv = bump_mapping(P.S, uv, rb_normalize(rayDir), *tbn, bm);
    if (UV_RANGE_SKIP && (uv.x < 0.0f || uv.x > 1.0f || uv.y < 0.0f || uv.y > 1.0f)) { do_skip(o, s, rayDir); return false; }
    wn = s.worldNormal;
    const int nm = __ldg(&props->normalMapTexID);
    if (nm >= 0) {
        const float4 t = sample_texture(P.S, nm, uv);
        rb_v3 tn = rb_mk3(t.x * 2.0f - 1.0f, t.y * 2.0f - 1.0f, t.z * 2.0f - 1.0f);
        tn.y = tn.y * -1.0f;
        wn = rb_normalize(rb_m3_mul(*tbn, tn));
    }
    col = ld3(props->albedo);
    const int tid = __ldg(&props->textureID);
    if (tid >= 0) {
        const float4 t = sample_texture(P.S, tid, uv);
        if (t.w < 0.999f && rb_random(&rng) > t.w) { do_skip(o, s, rayDir); return false; }
        col = col * rb_mk3(t.x, t.y, t.z);
    }
    return true;
}

// End of a path (rgen.glsl:264-284): clamp, drop NaN samples, add to the pixel's batch sum; then either restart the
// slot with the pixel's next sample (returns 1) or store the batch mean of the pixel (returns 0). L is the path's
// final radiance, st the slot's state record.
#ifndef RB_FINISH_IN_MISS
#define RB_FINISH_IN_MISS 1      // sky misses finish their path inside the miss shader instead of going through endQ + k_finish
#endif
__device__ __forceinline__ uint32_t finish_slot(const WaveParams& P, const uint32_t slot, const rb_v3 L, const uint4 st,
                                                uint32_t* cntNext, int parity) {
    const rb_v3 c = rb_clamp3_keepnan(L, 0.0f, P.pc.directClamp);
    float4 s4 = P.sum[slot];
    uint32_t actual = __float_as_uint(s4.w);
    if (!rb_anynan3(c)) { actual += 1u; s4.x += c.x; s4.y += c.y; s4.z += c.z; }
    const uint32_t sampleIdx = st.z + 1u;
    if (sampleIdx < P.pc.samplesPerPixel) {
        P.sum[slot] = make_float4(s4.x, s4.y, s4.z, __uint_as_float(actual));
        uint32_t rng = st.x;
        begin_path(P, slot, rng, sampleIdx);
        queue_push(P.rayQ[parity ^ 1], &cntNext[CNT_RAYS], slot);
        return 1u;
    }
    // last sample of the pixel: this batch's mean over the valid samples (rgen.glsl:275); folded into the image by
    // k_accumulate once the whole batch is done
    if (actual == 0u) P.mean[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    else {
        const rb_v3 fin = rb_mk3(s4.x, s4.y, s4.z) / (float)actual;
        P.mean[slot] = make_float4(fin.x, fin.y, fin.z, 1.f);
    }
    return 0u;
}

template <int MAT>
__device__ __forceinline__ uint32_t shade_slot(const WaveParams& P, const uint32_t slot, uint32_t* cnt, uint32_t* cntNext, int parity) {
    const float4 ro4 = P.rayO[slot], rd4 = P.rayD[slot];
    const rb_v3 rayOrigin = rb_mk3(ro4.x, ro4.y, ro4.z), rayDir = rb_mk3(rd4.x, rd4.y, rd4.z);
    uint4 st = P.st[slot];
    float4 T4 = P.thr[slot];
    rb_v3 T = rb_mk3(T4.x, T4.y, T4.z);
    float accDist = T4.w;
    uint32_t rng = st.x;
    uint32_t flags = st.y & 0xFFu;
    uint32_t segments = st.y >> 8;
    const bool nee = (P.flags & RB200_FLAG_NEE) != 0u;

    if (MAT == 4) {
        // miss shader + raygen's sky branch (rgen.glsl:138-141): radiance += sky * throughput, path ends
        const float4 L4 = P.rad[slot];
        const rb_v3 L = rb_mk3(L4.x, L4.y, L4.z) + sky_color(rayDir) * T;
#if RB_FINISH_IN_MISS
        // nothing else can add to this path (a miss casts no shadow ray, and the previous hit's shadow ray was resolved
        // in the previous wave), so the path ends here: about 60 % of all path ends skip endQ and k_finish
        return finish_slot(P, slot, L, st, cntNext, parity);
#else
        P.rad[slot] = make_float4(L.x, L.y, L.z, 0.f);
        queue_push(P.endQ, &cnt[CNT_END], slot);
        return 0u;
#endif
    }

    const uint4 h = P.hit[slot];
    const float a1 = __uint_as_float(h.x), a2 = __uint_as_float(h.y);
    const RB200Instance* inst = &P.S.instances[h.w];
#if RB_INST_RECORDS
    const RB200InstanceProperties* props = &P.S.instProps[h.w];
#else
    const RB200InstanceProperties* props = &P.S.props[__ldg(&inst->instancePropertiesID)];
#endif
    const bool prevInside = (flags & F_INSIDE) != 0u;

    Surf s;
    const bool hasNormalMap = __ldg(&props->normalMapTexID) >= 0;
    const bool needTbn = hasNormalMap || __ldg(&props->bumpMapTexID) >= 0;     // the parallax search runs in tangent space
#if RB_SHADE_RECORDS
    if (MAT == 3) hit_info_rec<true>(P.S, inst, props, h.z, a1, a2, rayDir, s);
    else if (needTbn) hit_info_rec<true>(P.S, inst, props, h.z, a1, a2, rayDir, s);
    else hit_info_rec<false>(P.S, inst, props, h.z, a1, a2, rayDir, s);
#else
    if (MAT == 3) hit_info<true>(P.S, inst, props, h.z, a1, a2, rayDir, s);
    else if (needTbn) hit_info<true>(P.S, inst, props, h.z, a1, a2, rayDir, s);
    else hit_info<false>(P.S, inst, props, h.z, a1, a2, rayDir, s);
#endif

    ShadeOut o;
    o.skip = false; o.pdf = 0.0f; o.inside = prevInside;
    o.color = o.albedo = o.emission = o.normal = rb_splat3(0.0f);
    bool didRefract = false;   // always false when NEE sees it (disney.rchit.glsl:187)
    rb_v3 wn, col;

    if (MAT == 0) {           // lambertian.rchit.glsl:11-78
        if (surface_prologue<true>(P, props, s, &s.tbn, rayDir, rng, o, wn, col)) {
            o.color = col; o.albedo = col; o.emission = ld3(props->emission);
            o.newO = rb_offset_along_normal(s.worldPosition, s.worldNormalGeometry);
            o.newD = diffuse_reflection(wn, rng);
            o.inside = false; o.normal = wn;
            o.pdf = rb_max(rb_dot(wn, o.newD), 0.0f) * RB_RCP_PI;
            accDist = 0.0f;
        }
    } else if (MAT == 1) {    // metal.rchit.glsl:7-70
        if (surface_prologue<false>(P, props, s, &s.tbn, rayDir, rng, o, wn, col)) {
            o.color = col; o.albedo = col; o.emission = ld3(props->emission);
            o.newO = rb_offset_along_normal(s.worldPosition, s.worldNormalGeometry);
            o.newD = fuzzy_reflection(rayDir, wn, __ldg(&props->roughness), rng);
            o.inside = false; o.normal = wn; o.pdf = 0.0f;
            accDist = 0.0f;
        }
    } else if (MAT == 2) {    // dielectric.rchit.glsl:40-113
        const float ior = __ldg(&props->ior), rough = __ldg(&props->roughness);
        const float ri = s.frontFace ? 1.0f / ior : ior;
        const rb_v3 unitDir = rb_normalize(rayDir);
        const float cosTheta = rb_min(rb_dot(-unitDir, s.worldNormal), 1.0f);
        const float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
        const bool cannotRefract = ri * sinTheta > 1.0f;
        const float reflectivity = schlick(cosTheta, ri);
        rb_v2 uv = rb_mk2(rb_fract_mod1(s.uv.x), rb_fract_mod1(s.uv.y));
        const int bm = __ldg(&props->bumpMapTexID);
        if (bm >= 0) uv = bump_mapping(P.S, uv, unitDir, s.tbn, bm);     // dielectric.rchit.glsl:59-61
        rb_v3 albedo = ld3(props->albedo);
        const int tid = __ldg(&props->textureID);
        if (tid >= 0) { const float4 t = sample_texture(P.S, tid, uv); albedo = albedo * rb_mk3(t.x, t.y, t.z); }
        wn = s.worldNormal;
        if (hasNormalMap) {
            const float4 t = sample_texture(P.S, __ldg(&props->normalMapTexID), uv);
            rb_v3 tn = rb_mk3(t.x * 2.0f - 1.0f, t.y * 2.0f - 1.0f, t.z * 2.0f - 1.0f);
            tn.y = tn.y * -1.0f;
            wn = rb_normalize(rb_m3_mul(s.tbn, tn));
        }
        if (cannotRefract || reflectivity > rb_random(&rng)) {
            o.newD = fuzzy_reflection(unitDir, wn, rough, rng);
            o.color = rb_splat3(1.0f);
            o.newO = rb_offset_along_normal(s.worldPosition, s.worldNormalGeometry);
        } else {
            o.inside = s.frontFace;
            const rb_v3 refr = rb_refract(unitDir, wn, ri);
            o.newD = refr + random_unit_vec(rng) * rough;
            o.color = albedo;
            o.newO = offset_for_dielectric(s.worldPosition, s.worldNormalGeometry, unitDir);
        }
        if (prevInside) accDist += rb_length(s.worldPosition - rayOrigin);
        else accDist += 0.0f;
        if (prevInside && !o.inside) {     // leaving the medium: Beer's law in metres (:89-100)
            const float a = rb_exp(-__ldg(&props->absorption) * accDist);
            o.color = rb_splat3(1.0f) * a;
            o.color = o.color * albedo;
            accDist = 0.0f;
        }
        o.albedo = o.color; o.emission = ld3(props->emission); o.normal = wn; o.pdf = 0.0f;
    } else {                  // disney.rchit.glsl:36-198
        if (surface_prologue<true>(P, props, s, &s.tbn, rayDir, rng, o, wn, col)) {
            const float ior = __ldg(&props->ior);
            const float eta = s.frontFace ? 1.0f / ior : ior;
            const DisneyP dp = load_disney(props, col, eta);
            bool choseGlass = false;
            const rb_v3 wi = -rayDir;     // not normalised, as in the reference
            const rb_v3 wo = disney_sample(s.tbn, dp, wn, wi, &didRefract, &choseGlass, rng);
            const rb_v3 hv = rb_normalize(wo + wi);
            float pdf;
            const rb_v3 f = disney_eval(s.tbn, dp, didRefract, wn, wi, wo, hv, &pdf);
            const float cosI = rb_max(rb_dot(wn, wo), 0.0f);
            o.color = (f * cosI) / pdf;
            o.albedo = col; o.pdf = pdf; o.emission = ld3(props->emission);
            o.newO = offset_for_dielectric(s.worldPosition, s.worldNormalGeometry, wo);
            o.newD = wo; o.normal = wn; o.inside = choseGlass;
            if (o.inside) accDist += rb_length(s.worldPosition - rayOrigin);
            else accDist = 0.0f;
        }
    }

    // ---- what raygen does after traceRayEXT returns (rgen.glsl:124-181) ----
    P.rayO[slot] = make_float4(o.newO.x, o.newO.y, o.newO.z, 0.f);
    P.rayD[slot] = make_float4(o.newD.x, o.newD.y, o.newD.z, 0.f);
    const bool leftDielectric = !o.inside && prevInside;
    segments += 1u;
    uint32_t newFlags = (flags & (F_FIRST | F_PREVSKIP)) | (o.inside ? F_INSIDE : 0u);
    if (!o.skip) {
        if (!o.inside) {
            const rb_v3 indirect = o.emission;
            const bool skipNEE = !nee || (MAT != 0 && MAT != 3);
            if (skipNEE) {
                const float4 L4 = P.rad[slot];
                const rb_v3 combined = rb_splat3(0.0f) * 0.0f + indirect * 1.0f;
                const rb_v3 L = rb_mk3(L4.x, L4.y, L4.z) + combined * T;
                P.rad[slot] = make_float4(L.x, L.y, L.z, 0.f);
            } else {
                // directLight (rgen.glsl:43-95) up to the visibility test, which the shadow stage resolves
                const LightSample target = random_emissive_point(P.S, P.pc.totalEmissiveWeight, rng);
                const rb_v3 toLight = target.point - o.newO;
                const rb_v3 direction = rb_normalize(toLight);
                const float dist = rb_length(toLight);
                const float pdfNEE = target.pdf * dist * dist / rb_max(rb_dot(target.normal, -direction), 0.0001f);
                rb_v3 brdf;
                if (MAT == 0) {
                    brdf = o.albedo * RB_RCP_PI;
                } else {
                    const rb_v3 wi = -rayDir;
                    const rb_v3 hv = rb_normalize(direction + wi);
                    const DisneyP dp = load_disney(props, o.albedo, 0.0f);     // pld.eta was overwritten with 0
                    float ignorePdf;
                    brdf = disney_eval(s.tbn, dp, false, o.normal, wi, direction, hv, &ignorePdf);
                }
                float cosThetai = rb_dot(o.normal, direction);
                cosThetai = target.cullBackface ? rb_max(cosThetai, 0.0f) : fabsf(cosThetai);
                float geomNum = rb_dot(target.normal, -direction);
                geomNum = target.cullBackface ? rb_max(geomNum, 0.0f) : fabsf(geomNum);
                const float geom = geomNum / (dist * dist);
                const rb_v3 D = (((target.emission * brdf) * cosThetai) * geom) / target.pdf;
                const float pdfBRDF = o.pdf;
                float wNEE, wBRDF;
                if ((flags & F_FIRST) || (flags & F_PREVSKIP) || leftDielectric) { wNEE = 1.0f; wBRDF = 1.0f; }
                else if (segments == P.pc.maxBounces) { wNEE = 0.0f; wBRDF = pdfBRDF * pdfBRDF / (pdfBRDF * pdfBRDF + pdfNEE * pdfNEE); }
                else {
                    wNEE = pdfNEE * pdfNEE / (pdfNEE * pdfNEE + pdfBRDF * pdfBRDF);
                    wBRDF = pdfBRDF * pdfBRDF / (pdfBRDF * pdfBRDF + pdfNEE * pdfNEE);
                }
                const rb_v3 Bv = indirect * wBRDF;
                const uint32_t k = queue_reserve(&cnt[CNT_SHADOW]);
                P.shO[k] = make_float4(o.newO.x, o.newO.y, o.newO.z, dist - 0.001f);
                P.shD[k] = make_float4(direction.x, direction.y, direction.z, 0.f);
                P.shA[k] = make_float4(D.x, D.y, D.z, wNEE);
                P.shB[k] = make_float4(Bv.x, Bv.y, Bv.z, __uint_as_float(slot));
                P.shT[k] = make_float4(T.x, T.y, T.z, 0.f);
            }
            if (skipNEE) newFlags |= F_PREVSKIP; else newFlags &= ~F_PREVSKIP;
            T = T * o.color;
        }
        newFlags &= ~F_FIRST;
    }
    P.thr[slot] = make_float4(T.x, T.y, T.z, accDist);
    P.st[slot] = make_uint4(rng, newFlags | (segments << 8), st.z, st.w);
    if (segments < P.pc.maxBounces) queue_push(P.rayQ[parity ^ 1], &cntNext[CNT_RAYS], slot);
    else queue_push(P.endQ, &cnt[CNT_END], slot);
    return 0u;
}

#ifndef RB_SHADE_MINBLOCKS
#define RB_SHADE_MINBLOCKS 1
#endif
#ifndef RB_SHADE_BLOCK
#define RB_SHADE_BLOCK 128       // threads per block of the shading kernels (85-119 registers: smaller blocks pack an SM's register file better)
#endif
#ifndef RB_DISNEY_MINBLOCKS
#define RB_DISNEY_MINBLOCKS 1
#endif
template <int MAT>
__global__ void __launch_bounds__(RB_SHADE_BLOCK, MAT == 3 ? RB_DISNEY_MINBLOCKS : RB_SHADE_MINBLOCKS) k_shade(WaveParams P, int parity) {
    uint32_t* cnt = P.counters + parity * CNT_SET;
    uint32_t* cntNext = P.counters + (parity ^ 1) * CNT_SET;
    const uint32_t n = cnt[CNT_MAT0 + MAT];
    const uint32_t* __restrict__ q = P.matQ[MAT];
    uint32_t started = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        started += shade_slot<MAT>(P, q[i], cnt, cntNext, parity);
    if (MAT == 4) {      // paths restarted by the miss shader (one counter update per warp)
        const uint32_t warpStarted = __reduce_add_sync(0xffffffffu, started);
        if ((threadIdx.x & 31u) == 0u && warpStarted) atomicAdd(&P.stats[ST_PATHS], (unsigned long long)warpStarted);
    }
}

// ---------------------------------------------------------------------------------------------------
// shadow: shadowRayOccluded (nee.h.glsl:126-144) + the radiance update of rgen.glsl:174-177
// ---------------------------------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(RB_TRAV_BLOCK, RB_TRAV_MINBLOCKS) k_shadow(WaveParams P, int parity) {
    uint32_t* cnt = P.counters + parity * CNT_SET;
    const uint32_t n = cnt[CNT_SHADOW];
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&P.stats[ST_SHADOW], (unsigned long long)n);
    uint32_t nodeVisits = 0, triTests = 0;
    extern __shared__ __align__(16) unsigned char rb_dyn_smem[];      // WarpShared<true>[RB_TRAV_BLOCK / 32]
    WarpShared<true>* ws = reinterpret_cast<WarpShared<true>*>(rb_dyn_smem);
    trace_queue<true, COUNT>(
        P.S.nodes, P.S.tris, n, &cnt[CNT_CURSOR_SHADOW],
        [&](uint32_t i, rb_v3& o, rb_v3& d, float& tmax) {
            const float4 o4 = P.shO[i], d4 = P.shD[i];
            o = rb_mk3(o4.x, o4.y, o4.z); d = rb_mk3(d4.x, d4.y, d4.z); tmax = o4.w;
        },
        [&](uint32_t i, const RayHit& h) {
            const bool occluded = h.tri != 0xFFFFFFFFu;
            const float4 A = P.shA[i], B = P.shB[i], T = P.shT[i];
            const uint32_t slot = __float_as_uint(B.w);
            const rb_v3 direct = occluded ? rb_splat3(0.0f) : rb_mk3(A.x, A.y, A.z);
            const rb_v3 combined = direct * A.w + rb_mk3(B.x, B.y, B.z);
            const float4 L4 = P.rad[slot];
            const rb_v3 L = rb_mk3(L4.x, L4.y, L4.z) + combined * rb_mk3(T.x, T.y, T.z);
            P.rad[slot] = make_float4(L.x, L.y, L.z, 0.f);
        },
        nodeVisits, triTests, ws[threadIdx.x >> 5]);
    if (COUNT) {
        atomicAdd(&P.stats[ST_NODES], (unsigned long long)nodeVisits);
        atomicAdd(&P.stats[ST_TRIS], (unsigned long long)triTests);
    }
}

// ---------------------------------------------------------------------------------------------------
// finish: end of a path (rgen.glsl:264-284)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK) k_finish(WaveParams P, int parity) {
    uint32_t* cnt = P.counters + parity * CNT_SET;
    uint32_t* cntNext = P.counters + (parity ^ 1) * CNT_SET;
    const uint32_t n = cnt[CNT_END];
    uint32_t started = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t slot = P.endQ[i];
        const float4 L4 = P.rad[slot];
        started += finish_slot(P, slot, rb_mk3(L4.x, L4.y, L4.z), P.st[slot], cntNext, parity);
    }
    // one counter update per warp, not per thread: a same-address 64-bit atomic from every thread serialises in L2
    const uint32_t warpStarted = __reduce_add_sync(0xffffffffu, started);
    if ((threadIdx.x & 31u) == 0u && warpStarted) atomicAdd(&P.stats[ST_PATHS], (unsigned long long)warpStarted);
}

// rgen.glsl:277-284: running average over batches (or the plain sum with RB200_FLAG_ACCUM_SUM). Runs once per batch,
// in batch order, after every pixel of the batch has delivered its mean.
__global__ void __launch_bounds__(BLOCK) k_accumulate(float4* __restrict__ image, const float4* __restrict__ mean, uint32_t n,
                                                      uint32_t sampleBatch, uint32_t flags,
                                                      const unsigned long long* __restrict__ laneStats,
                                                      unsigned long long* __restrict__ cumStats) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ST_COUNT) cumStats[i] += laneStats[i];     // batches are folded one at a time, in order: no race
    if (i >= n) return;
    const float4 m = mean[i];
    if (m.w < 0.f) return;       // tile partition: another rank's pixel
    if (m.w == 0.f) {
        // documented deviation: the reference writes 0/0 and poisons the pixel; batch 0 writes black, later batches keep
        // the previous value, sum mode adds nothing
        if (!(flags & RB200_FLAG_ACCUM_SUM) && sampleBatch == 0u) image[i] = make_float4(0.f, 0.f, 0.f, 1.f);
        return;
    }
    rb_v3 fin = rb_mk3(m.x, m.y, m.z);
    if (flags & RB200_FLAG_ACCUM_SUM) {
        const float4 prev = image[i];
        image[i] = make_float4(prev.x + fin.x, prev.y + fin.y, prev.z + fin.z, 1.f);
    } else {
        if (sampleBatch > 0u) {
            const float4 prev = image[i];
            fin = (rb_mk3(prev.x, prev.y, prev.z) * (float)sampleBatch + fin) / (float)(sampleBatch + 1u);
        }
        image[i] = make_float4(fin.x, fin.y, fin.z, 1.f);
    }
}

__global__ void k_resolve_sum(float4* image, uint32_t n, float inv) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = image[i];
    image[i] = make_float4(v.x * inv, v.y * inv, v.z * inv, 1.f);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
template <class K> static int persistent_grid(K kernel, int numSMs, int block = BLOCK, size_t dynSmem = 0) {
    int perSM = 0;
    if (dynSmem) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dynSmem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, block, dynSmem) != cudaSuccess || perSM < 1) perSM = 1;
    return numSMs * perSM;
}
static constexpr size_t SMEM_EXTEND = sizeof(WarpShared<false>) * (RB_TRAV_BLOCK / 32);
static constexpr size_t SMEM_SHADOW = sizeof(WarpShared<true>) * (RB_TRAV_BLOCK / 32);

// Per-triangle shading records (DeviceScene::shadeBase / shadeFrame): one thread per triangle slot copies exactly the
// floats hit_info() would gather for that triangle.
__global__ void k_build_shade_records(DeviceScene S, float4* __restrict__ base, float4* __restrict__ frame) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S.numTris) return;
    const float4* tp = reinterpret_cast<const float4*>(S.tris + t);
    const uint32_t prim = __float_as_uint(tp[0].w), inst = __float_as_uint(tp[1].w);
    const RB200InstanceProperties* props = &S.props[S.instances[inst].instancePropertiesID];
    const uint32_t ib = 3u * prim + props->indicesOffset;
    const float4 v0 = S.vertices[S.indices[ib]], v1 = S.vertices[S.indices[ib + 1]], v2 = S.vertices[S.indices[ib + 2]];
    float2 q0 = make_float2(0.f, 0.f), q1 = q0, q2 = q0;
    if (props->texIndicesOffset != 0xFFFFFFFFu) {
        const uint32_t xb = 3u * prim + props->texIndicesOffset;
        q0 = S.texCoords[S.texIndices[xb]]; q1 = S.texCoords[S.texIndices[xb + 1]]; q2 = S.texCoords[S.texIndices[xb + 2]];
    }
    float4* b = base + 4 * (size_t)t;
    b[0] = make_float4(v0.x, v0.y, v0.z, q0.x);
    b[1] = make_float4(v1.x, v1.y, v1.z, q0.y);
    b[2] = make_float4(v2.x, v2.y, v2.z, q1.x);
    b[3] = make_float4(q1.y, q2.x, q2.y, 0.f);
    const uint32_t tb = 3u * prim + props->tbnsIndicesOffset;
    const float* m0 = S.tbns + 9 * (size_t)S.tbnIndices[tb];
    const float* m1 = S.tbns + 9 * (size_t)S.tbnIndices[tb + 1];
    const float* m2 = S.tbns + 9 * (size_t)S.tbnIndices[tb + 2];
    // TBN entry: columns T (0..2), B (3..5), N (6..8)
    float4* f = frame + 8 * (size_t)t;
    f[0] = make_float4(m0[6], m0[7], m0[8], m0[0]);
    f[1] = make_float4(m1[6], m1[7], m1[8], m0[1]);
    f[2] = make_float4(m2[6], m2[7], m2[8], m0[2]);
    f[3] = make_float4(m1[0], m1[1], m1[2], m0[3]);
    f[4] = make_float4(m2[0], m2[1], m2[2], m0[4]);
    f[5] = make_float4(m1[3], m1[4], m1[5], m0[5]);
    f[6] = make_float4(m2[3], m2[4], m2[5], 0.f);
    f[7] = make_float4(0.f, 0.f, 0.f, 0.f);
}

int build_shade_records(const DeviceScene& S, float4* base, float4* frame, cudaStream_t stream) {
    if (S.numTris == 0) return RB200_OK;
    k_build_shade_records<<<(S.numTris + 255) / 256, 256, 0, stream>>>(S, base, frame);
    RB_CUDA(cudaGetLastError());
    return RB200_OK;
}

// see preload_post_kernels (post.cu)
void preload_wave_kernels() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, k_generate);
    cudaFuncGetAttributes(&a, k_extend<false>); cudaFuncGetAttributes(&a, k_extend<true>);
    cudaFuncGetAttributes(&a, k_shadow<false>); cudaFuncGetAttributes(&a, k_shadow<true>);
    cudaFuncGetAttributes(&a, k_shade<0>); cudaFuncGetAttributes(&a, k_shade<1>); cudaFuncGetAttributes(&a, k_shade<2>);
    cudaFuncGetAttributes(&a, k_shade<3>); cudaFuncGetAttributes(&a, k_shade<4>);
    cudaFuncGetAttributes(&a, k_finish);
    cudaFuncGetAttributes(&a, k_accumulate);
    cudaFuncGetAttributes(&a, k_resolve_sum);
    cudaGetLastError();
}

int render_batch(RB200Context* ctx, const RB200Scene* scene, const RB200RtPushConsts* pc) {
    if (pc->samplesPerPixel == 0 || pc->maxBounces == 0) { set_error("samplesPerPixel and maxBounces must be > 0"); return RB200_ERR_INVALID_ARGUMENT; }
    if ((ctx->flags & RB200_FLAG_NEE) && scene->numEmissive == 0) {
        set_error("Scene must have at least one emissive object");   // src/scene/Instances.cpp:125-127
        return RB200_ERR_NO_EMITTER;
    }
    // Lane selection: consecutive calls rotate through RB_LANES path-state sets on as many internal streams, so the
    // thin tail of batch b (few live paths, latency-bound launches) overlaps the heads of the following batches.
    const int lane = (int)(ctx->batchCalls % RB_LANES);
    const int other = (lane + RB_LANES - 1) % RB_LANES;        // the lane of the previous batch
    const bool hadPrevious = ctx->batchCalls > 0;
    ctx->batchCalls++;
    WaveParams& P = ctx->lanes[lane];
    P.S = scene->dev;
    P.pc = *pc;
    cudaStream_t s = ctx->laneStream[lane];
    // everything the caller enqueued on the front-end stream before this call (write_hdr, postprocess of the previous
    // frame, an external reduce of the image ...) must precede this batch's accumulation — not its tracing
    RB_CUDA(cudaEventRecord(ctx->frontMark, ctx->stream));
    const bool count = (ctx->flags & RB200_FLAG_COUNT_BVH) != 0;
    if (ctx->flags & RB200_FLAG_TIME_KERNELS)      // timing pass: no overlap
        for (int l = 0; l < RB_LANES; l++) if (l != lane) RB_CUDA(cudaStreamSynchronize(ctx->laneStream[l]));

    static int gExtend = 0, gExtendC = 0, gShadow = 0, gShadowC = 0, gShade[5] = {0, 0, 0, 0, 0}, gFinish = 0;
    if (!gExtend) {
        gExtend = persistent_grid(k_extend<false>, ctx->numSMs, RB_TRAV_BLOCK, SMEM_EXTEND);
        gExtendC = persistent_grid(k_extend<true>, ctx->numSMs, RB_TRAV_BLOCK, SMEM_EXTEND);
        gShadow = persistent_grid(k_shadow<false>, ctx->numSMs, RB_TRAV_BLOCK, SMEM_SHADOW);
        gShadowC = persistent_grid(k_shadow<true>, ctx->numSMs, RB_TRAV_BLOCK, SMEM_SHADOW);
        gShade[0] = persistent_grid(k_shade<0>, ctx->numSMs, RB_SHADE_BLOCK);
        gShade[1] = persistent_grid(k_shade<1>, ctx->numSMs, RB_SHADE_BLOCK);
        gShade[2] = persistent_grid(k_shade<2>, ctx->numSMs, RB_SHADE_BLOCK);
        gShade[3] = persistent_grid(k_shade<3>, ctx->numSMs, RB_SHADE_BLOCK);
        gShade[4] = persistent_grid(k_shade<4>, ctx->numSMs, RB_SHADE_BLOCK);
        gFinish = persistent_grid(k_finish, ctx->numSMs);
    }

    // lane stagger (see RB200Context::staggerWave): start behind wave `staggerWave` of the previous batch
    if (ctx->staggerWave >= 0 && hadPrevious && other != lane && !(ctx->flags & RB200_FLAG_TIME_KERNELS))
        RB_CUDA(cudaStreamWaitEvent(s, ctx->staggerEv[other], 0));
    // snapshot of the cumulative device counters at batch start (device-to-device: no host synchronisation here;
    // rb200_get_stats resolves "last batch" = cumulative - snapshot after synchronising)
    RB_CUDA(cudaMemsetAsync(P.stats, 0, ST_COUNT * sizeof(unsigned long long), s));   // this lane's per-batch counters
    RB_CUDA(cudaMemsetAsync(P.counters, 0, 2 * CNT_SET * sizeof(uint32_t), s));
    uint64_t nl = 0;
    const bool timed = (ctx->flags & RB200_FLAG_TIME_KERNELS) != 0;
    ctx->evUsed = 0; ctx->evClass.clear();
    auto tic = [&](int cls) {
        if (!timed) return;
        while (ctx->evPool.size() < ctx->evUsed + 2) { cudaEvent_t e; cudaEventCreate(&e); ctx->evPool.push_back(e); }
        ctx->evClass.push_back(cls);
        cudaEventRecord(ctx->evPool[ctx->evUsed], s);
    };
    auto toc = [&]() {
        if (!timed) return;
        cudaEventRecord(ctx->evPool[ctx->evUsed + 1], s);
        ctx->evUsed += 2;
    };
    tic(0); k_generate<<<(P.N + BLOCK - 1) / BLOCK, BLOCK, 0, s>>>(P); toc(); nl++;
    const uint32_t maxWaves = pc->samplesPerPixel * pc->maxBounces;
    // developer aid: RB200_WAVE_LOG=<file> (with RB200_FLAG_TIME_KERNELS) writes one CSV row per wave — queue counters
    // and the device time of every kernel — for the last batch rendered; it synchronises after every wave
    const char* waveLogPath = timed ? getenv("RB200_WAVE_LOG") : nullptr;
    std::vector<uint32_t> waveCounters;
    const bool nee = (ctx->flags & RB200_FLAG_NEE) != 0;
    if (timed) {
        if (ctx->waveCountsCap < maxWaves) {
            if (ctx->waveCountsDev) cudaFree(ctx->waveCountsDev);
            ctx->waveCountsDev = nullptr; ctx->waveCountsCap = 0;
            RB_CUDA(cudaMalloc(&ctx->waveCountsDev, (size_t)maxWaves * 2 * sizeof(uint32_t)));
            ctx->waveCountsCap = maxWaves;
        }
        ctx->waveCountsWaves = maxWaves;
    }
    const uint32_t staggerWant = ctx->staggerWave < 0 ? 0u : ctx->staggerWave > 0 ? (uint32_t)ctx->staggerWave : std::max(1u, maxWaves * 5u / 32u);
    const uint32_t staggerAt = timed ? 0u : std::min(staggerWant, maxWaves);
    bool capturing = false;
    const uint32_t mats = P.S.materialMask & 15u;
    auto launch_waves_on = [&](WaveParams& P, cudaStream_t s, int lane) -> int {
        for (uint32_t w = 0; w < maxWaves; w++) {
            const int p = (int)(w & 1u);
            RB_CUDA(cudaMemsetAsync(P.counters + (p ^ 1) * CNT_SET, 0, CNT_SET * sizeof(uint32_t), s));
            tic(1);
            if (count) k_extend<true><<<gExtendC, RB_TRAV_BLOCK, SMEM_EXTEND, s>>>(P, p); else k_extend<false><<<gExtend, RB_TRAV_BLOCK, SMEM_EXTEND, s>>>(P, p);
            toc();
            tic(6); k_shade<4><<<gShade[4], RB_SHADE_BLOCK, 0, s>>>(P, p); toc();
            // a material no instance uses has an empty queue in every wave: its kernel is not launched
            if (mats & 1u) { tic(2); k_shade<0><<<gShade[0], RB_SHADE_BLOCK, 0, s>>>(P, p); toc(); }
            if (mats & 2u) { tic(3); k_shade<1><<<gShade[1], RB_SHADE_BLOCK, 0, s>>>(P, p); toc(); }
            if (mats & 4u) { tic(4); k_shade<2><<<gShade[2], RB_SHADE_BLOCK, 0, s>>>(P, p); toc(); }
            if (mats & 8u) { tic(5); k_shade<3><<<gShade[3], RB_SHADE_BLOCK, 0, s>>>(P, p); toc(); }
            if (nee) {
                tic(7);
                if (count) k_shadow<true><<<gShadowC, RB_TRAV_BLOCK, SMEM_SHADOW, s>>>(P, p); else k_shadow<false><<<gShadow, RB_TRAV_BLOCK, SMEM_SHADOW, s>>>(P, p);
                toc();
            }
            tic(8); k_finish<<<gFinish, BLOCK, 0, s>>>(P, p); toc();
            if (timed && ctx->waveCountsDev) {      // CNT_RAYS / CNT_SHADOW of this wave, device to device: no host wait
                RB_CUDA(cudaMemcpyAsync(ctx->waveCountsDev + 2 * w, P.counters + p * CNT_SET + CNT_RAYS, sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
                RB_CUDA(cudaMemcpyAsync(ctx->waveCountsDev + 2 * w + 1, P.counters + p * CNT_SET + CNT_SHADOW, sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
            }
            if (staggerAt && w + 1 == staggerAt)
                RB_CUDA(cudaEventRecordWithFlags(ctx->staggerEv[lane], s, capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
            if (waveLogPath) {
                waveCounters.resize((size_t)(w + 1) * CNT_SET);
                RB_CUDA(cudaMemcpyAsync(&waveCounters[(size_t)w * CNT_SET], P.counters + p * CNT_SET, CNT_SET * sizeof(uint32_t),
                                        cudaMemcpyDeviceToHost, s));
                RB_CUDA(cudaStreamSynchronize(s));
            }
        }
        return RB200_OK;
    };
    auto launch_waves = [&]() -> int { return launch_waves_on(P, s, lane); };
    nl += (uint64_t)maxWaves * ((nee ? 4u : 3u) + (uint32_t)__builtin_popcount(mats));
    static const bool graphsOff = getenv("RB200_NO_GRAPH") != nullptr;
    if (timed || count || graphsOff) {
        const int rc = launch_waves();
        if (rc != RB200_OK) return rc;
    } else {
        // the wave loop as one graph launch: ~1150 stream operations per batch become one, and the device-side gap
        // between the dependent kernels of a thin wave shrinks
        // (Re)capture when the arguments changed. Every lane is captured at once — the other lanes' graphs differ only in
        // their buffers — so the cost (a few ms per lane) falls on one call instead of on the first call of each lane.
        const uint32_t wavesKey = maxWaves | (staggerAt << 16);
        {
            WaveParams key = P;
            key.pc.sampleBatch = 0u;
            if (!ctx->waveGraph[lane] || ctx->waveGraphWaves[lane] != wavesKey || memcmp(&key, &ctx->waveGraphKey[lane], sizeof(WaveParams)) != 0) {
                const int laneNow = lane;
                for (int l = 0; l < RB_LANES; l++) {
                    WaveParams& PL = ctx->lanes[l];
                    PL.S = P.S; PL.pc = P.pc;
                    WaveParams k2 = PL;
                    k2.pc.sampleBatch = 0u;
                    if (ctx->waveGraph[l] && ctx->waveGraphWaves[l] == wavesKey && memcmp(&k2, &ctx->waveGraphKey[l], sizeof(WaveParams)) == 0) continue;
                    if (ctx->waveGraph[l]) { cudaGraphExecDestroy(ctx->waveGraph[l]); ctx->waveGraph[l] = nullptr; }
                    cudaGraph_t g = nullptr;
                    cudaStream_t sl = ctx->laneStream[l];
                    RB_CUDA(cudaStreamBeginCapture(sl, cudaStreamCaptureModeThreadLocal));
                    capturing = true;
                    const int rc = launch_waves_on(PL, sl, l);
                    capturing = false;
                    const cudaError_t ce = cudaStreamEndCapture(sl, &g);
                    if (rc != RB200_OK || ce != cudaSuccess || !g) {
                        cudaGetLastError();
                        if (g) cudaGraphDestroy(g);
                        set_error("CUDA graph capture of the wave loop failed: %s", cudaGetErrorString(ce));
                        return RB200_ERR_CUDA;
                    }
                    const cudaError_t ie = cudaGraphInstantiate(&ctx->waveGraph[l], g, 0);
                    cudaGraphDestroy(g);
                    if (ie != cudaSuccess) { cudaGetLastError(); ctx->waveGraph[l] = nullptr; set_error("cudaGraphInstantiate: %s", cudaGetErrorString(ie)); return RB200_ERR_CUDA; }
                    memcpy(&ctx->waveGraphKey[l], &k2, sizeof(WaveParams));
                    ctx->waveGraphWaves[l] = wavesKey;
                    ctx->graphCaptures++;
                }
                (void)laneNow;
            }
        }
        RB_CUDA(cudaGraphLaunch(ctx->waveGraph[lane], s));
    }
    if (waveLogPath) {
        if (FILE* f = fopen(waveLogPath, "w")) {
            fprintf(f, "wave,rays,lambertian,metal,dielectric,disney,miss,shadow,end,extend_us,miss_us,lambertian_us,metal_us,dielectric_us,disney_us,shadow_us,finish_us\n");
            size_t e = 2;      // event pair 0 is k_generate
            for (uint32_t w = 0; w < maxWaves; w++) {
                const uint32_t* c = &waveCounters[(size_t)w * CNT_SET];
                fprintf(f, "%u,%u,%u,%u,%u,%u,%u,%u,%u", w, c[CNT_RAYS], c[CNT_MAT0], c[CNT_MAT0 + 1], c[CNT_MAT0 + 2], c[CNT_MAT0 + 3],
                        c[CNT_MISS], c[CNT_SHADOW], c[CNT_END]);
                // event classes: 1 extend, 6 miss, 2..5 materials, 7 shadow, 8 finish; columns in that order
                float us[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                for (;;) {
                    const int cls = ctx->evClass[e / 2];
                    float ms = 0.f;
                    cudaEventElapsedTime(&ms, ctx->evPool[e], ctx->evPool[e + 1]);
                    us[cls] = ms * 1000.f;
                    e += 2;
                    if (cls == 8) break;
                }
                fprintf(f, ",%.1f,%.1f,%.1f,%.1f,%.1f,%.1f,%.1f,%.1f\n", us[1], us[6], us[2], us[3], us[4], us[5], us[7], us[8]);
            }
            fclose(f);
        }
    }
    // fold this batch's pixel means into the shared image: after the previous batch's fold and after whatever the
    // caller had queued on the front-end stream; then let the front-end stream see the result
    if (hadPrevious && other != lane) RB_CUDA(cudaStreamWaitEvent(s, ctx->accumDone[other], 0));
    RB_CUDA(cudaStreamWaitEvent(s, ctx->frontMark, 0));
    k_accumulate<<<(P.N + BLOCK - 1) / BLOCK, BLOCK, 0, s>>>(P.image, P.mean.p, P.N, pc->sampleBatch, ctx->flags, P.stats,
                                                             ctx->statsSnap); nl++;
    RB_CUDA(cudaEventRecord(ctx->accumDone[lane], s));
    RB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->accumDone[lane], 0));
    RB_CUDA(cudaGetLastError());
    ctx->last.waves = maxWaves;
    ctx->last.kernelLaunches = nl;
    ctx->cumulative.waves += maxWaves;
    ctx->launches += nl;
    return RB200_OK;
}

int resolve_sum(RB200Context* ctx, uint32_t numBatches) {
    if (numBatches == 0) { set_error("numBatches must be > 0"); return RB200_ERR_INVALID_ARGUMENT; }
    WaveParams& P = ctx->wp;
    k_resolve_sum<<<(P.N + BLOCK - 1) / BLOCK, BLOCK, 0, ctx->stream>>>(P.image, P.N, 1.0f / (float)numBatches);
    ctx->launches++;
    RB_CUDA(cudaGetLastError());
    return RB200_OK;
}

// ---------------------------------------------------------------------------------------------------
// parity / measurement entry points: plain closest-hit and any-hit queries
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK) k_primary_rays(WaveParams P, float4* o, float4* d) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= P.N) return;
    const uint32_t x = slot % P.W, y = slot / P.W;
    uint32_t rng = (P.pc.sampleBatch * P.H + y) * P.W + x;
    rb_v3 ro, rd;
    starting_ray(P.pc, (float)x, (float)y, (float)P.W, (float)P.H, rng, ro, rd);
    o[slot] = make_float4(ro.x, ro.y, ro.z, 10000.0f);
    d[slot] = make_float4(rd.x, rd.y, rd.z, 0.f);
}

template <bool ANY>
__global__ void __launch_bounds__(BLOCK) k_trace_query(const WideNode* nodes, const TriRecord* tris, uint32_t n,
                                                       const float4* __restrict__ o, const float4* __restrict__ d,
                                                       RB200PrimaryHit* __restrict__ out, uint32_t* cursor) {
    uint32_t nv = 0, tt = 0;
    extern __shared__ __align__(16) unsigned char rb_dyn_smem[];      // WarpShared<ANY>[BLOCK / 32]
    WarpShared<ANY>* ws = reinterpret_cast<WarpShared<ANY>*>(rb_dyn_smem);
    trace_queue<ANY, false>(
        nodes, tris, n, cursor,
        [&](uint32_t i, rb_v3& ro, rb_v3& rd, float& tmax) {
            const float4 o4 = o[i], d4 = d[i];
            ro = rb_mk3(o4.x, o4.y, o4.z); rd = rb_mk3(d4.x, d4.y, d4.z); tmax = o4.w;
        },
        [&](uint32_t i, const RayHit& h) {
            RB200PrimaryHit r;
            r.u = h.b1; r.v = h.b2;
            if (h.tri != 0xFFFFFFFFu) {
                const float4* tp = reinterpret_cast<const float4*>(tris + h.tri);
                r.t = h.t;
                r.primitive = __float_as_uint(__ldg(tp).w);
                r.instance = __float_as_uint(__ldg(tp + 1).w);
            } else { r.t = -1.0f; r.u = r.v = 0.f; r.primitive = r.instance = 0xFFFFFFFFu; }
            out[i] = r;
        },
        nv, tt, ws[threadIdx.x >> 5]);
}

static int run_query(RB200Context* ctx, const RB200Scene* scene, uint32_t n, const float4* dO, const float4* dD, int any,
                     RB200PrimaryHit* out) {
    RB200PrimaryHit* dOut;
    for (int lane = 0; lane < RB_LANES; lane++) RB_CUDA(cudaStreamSynchronize(ctx->laneStream[lane]));   // lane 0's arrays are the scratch
    uint32_t* dCursor;
    RB_CUDA(cudaMalloc(&dOut, (size_t)n * sizeof(RB200PrimaryHit)));
    RB_CUDA(cudaMalloc(&dCursor, sizeof(uint32_t)));
    RB_CUDA(cudaMemsetAsync(dCursor, 0, sizeof(uint32_t), ctx->stream));
    const int grid = (int)std::min<uint64_t>((n + BLOCK - 1) / BLOCK, (uint64_t)ctx->numSMs * 8);
    constexpr size_t smAny = sizeof(WarpShared<true>) * (BLOCK / 32), smClosest = sizeof(WarpShared<false>) * (BLOCK / 32);
    RB_CUDA(cudaFuncSetAttribute(k_trace_query<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smAny));
    RB_CUDA(cudaFuncSetAttribute(k_trace_query<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smClosest));
    if (any) k_trace_query<true><<<grid, BLOCK, smAny, ctx->stream>>>(scene->dev.nodes, scene->dev.tris, n, dO, dD, dOut, dCursor);
    else k_trace_query<false><<<grid, BLOCK, smClosest, ctx->stream>>>(scene->dev.nodes, scene->dev.tris, n, dO, dD, dOut, dCursor);
    ctx->launches++;
    RB_CUDA(cudaGetLastError());
    RB_CUDA(cudaMemcpyAsync(out, dOut, (size_t)n * sizeof(RB200PrimaryHit), cudaMemcpyDeviceToHost, ctx->stream));
    RB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(dOut); cudaFree(dCursor);
    return RB200_OK;
}

int trace_primary(RB200Context* ctx, const RB200Scene* scene, const RB200RtPushConsts* pc, RB200PrimaryHit* out) {
    WaveParams& P = ctx->wp;
    for (int lane = 0; lane < RB_LANES; lane++) RB_CUDA(cudaStreamSynchronize(ctx->laneStream[lane]));
    P.S = scene->dev; P.pc = *pc;
    k_primary_rays<<<(P.N + BLOCK - 1) / BLOCK, BLOCK, 0, ctx->stream>>>(P, P.shO.p, P.shD.p);
    ctx->launches++;
    return run_query(ctx, scene, P.N, P.shO.p, P.shD.p, 0, out);
}

int trace_rays(RB200Context* ctx, const RB200Scene* scene, uint32_t n, const float* o, const float* d, const float* tmax,
               int any, RB200PrimaryHit* out) {
    if (n == 0) return RB200_OK;
    std::vector<float4> ho(n), hd(n);
    for (uint32_t i = 0; i < n; i++) {
        ho[i] = make_float4(o[3 * i], o[3 * i + 1], o[3 * i + 2], tmax[i]);
        hd[i] = make_float4(d[3 * i], d[3 * i + 1], d[3 * i + 2], 0.f);
    }
    float4 *dO, *dD;
    RB_CUDA(cudaMalloc(&dO, (size_t)n * sizeof(float4)));
    RB_CUDA(cudaMalloc(&dD, (size_t)n * sizeof(float4)));
    RB_CUDA(cudaMemcpyAsync(dO, ho.data(), (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    RB_CUDA(cudaMemcpyAsync(dD, hd.data(), (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    int rc = run_query(ctx, scene, n, dO, dD, any, out);
    cudaFree(dO); cudaFree(dD);
    return rc;
}

} // namespace rb200
