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
#include "two_level.cuh"
#include "shade.cuh"

namespace rb200 {

static constexpr int BLOCK = 256;
#ifndef RB_TRAV_BLOCK
#define RB_TRAV_BLOCK 256        // threads per block of k_extend / k_shadow (with RB_TRAV_MINBLOCKS: 1024 threads per SM at 64 registers)
#endif
// RB_SHADOW_LDCS=1: the shadow-ray records are read once, by k_shadow, and are dead afterwards: ld.global.cs (evict
// first) keeps their 5 x 33 MB per wave from displacing the hierarchy and the path state in L2.
#ifndef RB_SHADOW_LDCS
#define RB_SHADOW_LDCS 0
#endif
// RB_BIN_MATCH=1: k_extend bins the finished rays of a warp by material with one __match_any_sync instead of five
// predicated passes.
#ifndef RB_BIN_MATCH
#define RB_BIN_MATCH 1
#endif
#ifndef RB_TRAV_MINBLOCKS
#define RB_TRAV_MINBLOCKS (1024 / RB_TRAV_BLOCK)      // resident blocks per SM the traversal kernels are compiled for (register cap)
#endif
// The two traversal kernels want different shapes (B200, headline step, ms per step, profiles/r02l_variant_sweep.txt):
//   256 threads x 4 blocks (1024 threads per SM, 64 registers): k_extend 19.13, k_shadow 9.49
//   128 threads x 7 blocks ( 896 threads per SM, 72 registers, no spill): k_extend 18.37, k_shadow 9.82
//   256 threads x 3 blocks ( 768 threads per SM, 80 registers): k_extend 19.36, k_shadow 10.52
// The closest-hit kernel carries more live state (best hit, barycentrics, triangle id) and gains from eight more registers
// what it loses with an eighth of its warps; the any-hit kernel does not.
#ifndef RB_EXTEND_BLOCK
#define RB_EXTEND_BLOCK 128
#endif
#ifndef RB_EXTEND_MINBLOCKS
#define RB_EXTEND_MINBLOCKS 7
#endif
#ifndef RB_SHADOW_BLOCK
#define RB_SHADOW_BLOCK RB_TRAV_BLOCK
#endif
#ifndef RB_SHADOW_MINBLOCKS
#define RB_SHADOW_MINBLOCKS RB_TRAV_MINBLOCKS
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

// lane of an engine slot (slot = lane * N + pixel); an engine has at most RB_MAX_LANES lanes
__device__ __forceinline__ uint32_t lane_of(const WaveParams& P, uint32_t slot) {
    uint32_t l = 0;
    for (uint32_t k = 1; k < P.numLanes; k++) l += slot >= k * P.N ? 1u : 0u;
    return l;
}

// Per-block accumulators of the per-lane batch counters (shared memory; flushed to WaveParams::stats once per block).
// word 0: extend rays (low 32 bits) | shadow rays (high 32 bits) of the paths that ended; word 1: paths started.
struct LaneAcc { unsigned long long rays[RB_MAX_LANES]; unsigned int started[RB_MAX_LANES]; };
__device__ __forceinline__ void lane_acc_init(LaneAcc& a) {
    if (threadIdx.x < RB_MAX_LANES) { a.rays[threadIdx.x] = 0ull; a.started[threadIdx.x] = 0u; }
    __syncthreads();
}
__device__ __forceinline__ void lane_acc_flush(const WaveParams& P, LaneAcc& a) {
    __syncthreads();
    if (threadIdx.x < P.numLanes) {
        const unsigned long long r = a.rays[threadIdx.x];
        const unsigned int st = a.started[threadIdx.x];
        unsigned long long* dst = P.stats + (size_t)threadIdx.x * ST_COUNT;
        if (r & 0xFFFFFFFFull) atomicAdd(&dst[ST_EXTEND], r & 0xFFFFFFFFull);
        if (r >> 32) atomicAdd(&dst[ST_SHADOW], r >> 32);
        if (st) atomicAdd(&dst[ST_PATHS], (unsigned long long)st);
    }
}

// start (or restart) the path of a slot: traceSegments' locals (rgen.glsl:98-104)
__device__ __forceinline__ void begin_path(const WaveParams& P, uint32_t slot, uint32_t pixel, uint32_t& rng, uint32_t sampleIdx) {
    const uint32_t x = pixel % P.W, y = pixel / P.W;
    rb_v3 o, d;
    starting_ray(P.pc, (float)x, (float)y, (float)P.W, (float)P.H, rng, o, d);
    store_pair(P.rayO, P.rayD, slot, make_float4(o.x, o.y, o.z, 0.f), make_float4(d.x, d.y, d.z, 0.f));
    // throughput 1, accumulatedDistance 0 (documented deviation)
    store_pair(P.thr, P.st, slot, make_float4(1.f, 1.f, 1.f, 0.f), make_uint4(rng, F_FIRST, sampleIdx, 0u));
    P.rad[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// Starts batch `sampleBatch` in lane `lane` of the engine: one camera path per pixel, appended to the ray queue that the
// next wave (queue set `parity`) reads. Runs between two waves on the engine's stream.
__global__ void __launch_bounds__(BLOCK) k_generate(WaveParams P, uint32_t lane, uint32_t sampleBatch, int parity) {
    const uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t* cnt = P.counters + parity * CNT_SET;
    unsigned long long* laneStats = P.stats + (size_t)lane * ST_COUNT;
    bool mine = pixel < P.N;
    const uint32_t x = pixel % P.W, y = pixel / P.W;
    // interleaved-tile partition (latency mode): this context traces the pixels of the tiles whose
    // row-major index is congruent to tileRank; the other pixels are not touched (mean.w < 0 makes k_accumulate
    // skip them), so the partial images of all ranks add up to the single-GPU image bit for bit
    if (mine && P.tileCount > 1u) {
        mine = ((y / P.tileSize) * P.tilesX + x / P.tileSize) % P.tileCount == P.tileRank;
        if (!mine) P.mean[lane * P.N + pixel] = make_float4(0.f, 0.f, 0.f, -1.f);     // w < 0: not this rank's pixel
    }
    // one queue reservation per block
    __shared__ uint32_t s_base, s_count;
    if (threadIdx.x == 0) s_count = 0u;
    __syncthreads();
    uint32_t rank = 0;
    if (mine) rank = atomicAdd(&s_count, 1u);
    __syncthreads();
    if (threadIdx.x == 0 && s_count) {
        s_base = atomicAdd(&cnt[CNT_RAYS], s_count);
        atomicAdd(&laneStats[ST_PATHS], (unsigned long long)s_count);
    }
    __syncthreads();
    if (!mine) return;
    const uint32_t slot = lane * P.N + pixel;
    uint32_t rng = (sampleBatch * P.H + y) * P.W + x;       // rgen.glsl:259
    P.sum[slot] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0u));
    begin_path(P, slot, pixel, rng, 0u);
    P.rayQ[parity][s_base + rank] = slot;
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// ---------------------------------------------------------------------------------------------------
// extend
// ---------------------------------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(RB_EXTEND_BLOCK, RB_EXTEND_MINBLOCKS) k_extend(WaveParams P, int parity) {
    uint32_t* cnt = P.counters + parity * CNT_SET;
    const uint32_t n = cnt[CNT_RAYS];
    const uint32_t* __restrict__ q = P.rayQ[parity];
    // (rays are counted per batch when their path ends: finish_slot)
    uint32_t nodeVisits = 0, triTests = 0;
    extern __shared__ __align__(16) unsigned char rb_dyn_smem[];      // WarpShared<false>[RB_EXTEND_BLOCK / 32]: may exceed the 48 KB static limit
    WarpShared<false>* ws = reinterpret_cast<WarpShared<false>*>(rb_dyn_smem);
    trace_queue<false, COUNT>(
        P.S.nodes, P.S.tris, n, &cnt[CNT_CURSOR_EXTEND],
        [&](uint32_t i, rb_v3& o, rb_v3& d, float& tmax) {
            const uint32_t slot = q[i];
            float4 o4, d4;
            load_pair(P.rayO, P.rayD, slot, o4, d4);
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
                const uint32_t iw = __float_as_uint(__ldg(tp + 1).w);      // instance | material kernel << 30: no instance-table gather
                P.hit[slot] = make_uint4(__float_as_uint(h.b1), __float_as_uint(h.b2), prim, iw & TRI_INST_MASK);
                bin = iw >> 30;
            }
            // bin the slot by material (warp-aggregated per bin)
#if RB_BIN_MATCH
            const uint32_t active = __activemask();
            const uint32_t peers = __match_any_sync(active, bin);
            const uint32_t lane = threadIdx.x & 31u;
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(&cnt[CNT_MAT0 + bin], (uint32_t)__popc(peers));
            base = __shfl_sync(peers, base, leader);
            uint32_t* mq = bin == 0u ? P.matQ[0] : bin == 1u ? P.matQ[1] : bin == 2u ? P.matQ[2] : bin == 3u ? P.matQ[3] : P.matQ[4];
            mq[base + __popc(peers & ((1u << lane) - 1u))] = slot;
#else
#pragma unroll
            for (uint32_t b = 0; b < 5; b++)
                if (bin == b) queue_push(P.matQ[b], &cnt[CNT_MAT0 + b], slot);
#endif
        },
        nodeVisits, triTests, ws[threadIdx.x >> 5],
        [&](uint32_t i) { return q[i]; },
        [&](uint32_t slot) { prefetch_l1(reinterpret_cast<const char*>(P.rayO.p + (size_t)slot * STATE_STRIDE)); });
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
// slot with the pixel's next sample or store the batch mean of the pixel. L is the path's final radiance, st the slot's
// state record, extendRays the rays this path traced. The batch counters of the slot's lane are updated through the
// block's accumulators.
__device__ __forceinline__ void finish_slot(const WaveParams& P, const uint32_t slot, const rb_v3 L, const uint4 st,
                                            const uint32_t extendRays, uint32_t* cntNext, int parity, LaneAcc& acc) {
    const uint32_t lane = lane_of(P, slot);
    const uint32_t sampleIdx = st.z + 1u;
    const bool restart = sampleIdx < P.pc.samplesPerPixel;
    {
        // One shared-memory atomic per engine lane and warp instead of one per thread: a 64-bit shared atomicAdd is a
        // compare-and-swap loop, and the threads of a warp mostly end paths of the same lane (same address): the miss
        // shader spent 24 of its 45 stall cycles per issue there (short scoreboard, profiles/r02b).
        const uint32_t peers = __match_any_sync(__activemask(), lane);
        const uint32_t ext = __reduce_add_sync(peers, extendRays);
        const uint32_t shd = __reduce_add_sync(peers, st.w);
        const uint32_t started = __reduce_add_sync(peers, restart ? 1u : 0u);
        if ((int)(threadIdx.x & 31u) == __ffs(peers) - 1) {
            atomicAdd(&acc.rays[lane], (unsigned long long)ext | ((unsigned long long)shd << 32));
            if (started) atomicAdd(&acc.started[lane], started);
        }
    }
    const rb_v3 c = rb_clamp3_keepnan(L, 0.0f, P.pc.directClamp);
    float4 s4 = P.sum[slot];
    uint32_t actual = __float_as_uint(s4.w);
    if (!rb_anynan3(c)) { actual += 1u; s4.x += c.x; s4.y += c.y; s4.z += c.z; }
    if (restart) {
        P.sum[slot] = make_float4(s4.x, s4.y, s4.z, __uint_as_float(actual));
        uint32_t rng = st.x;
        begin_path(P, slot, slot - lane * P.N, rng, sampleIdx);
        queue_push(P.rayQ[parity ^ 1], &cntNext[CNT_RAYS], slot);
        return;
    }
    // last sample of the pixel: this batch's mean over the valid samples (rgen.glsl:275); folded into the image by
    // k_accumulate once the whole batch is done
    if (actual == 0u) P.mean[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    else {
        const rb_v3 fin = rb_mk3(s4.x, s4.y, s4.z) / (float)actual;
        P.mean[slot] = make_float4(fin.x, fin.y, fin.z, 1.f);
    }
}

// The closest-hit shader of material MAT on one hit (lambertian / metal / dielectric / disney .rchit.glsl): everything
// the shader writes into the payload — new ray, colour, albedo, emission, normal, pdf, the skip and insideDielectric
// flags, the RNG state and the accumulated distance. Used by the wave loop (shade_slot) and by rb200_shade_hits.
template <int MAT>
__device__ __forceinline__ void eval_hit(const WaveParams& P, const uint4 h, const rb_v3 rayOrigin, const rb_v3 rayDir, const bool prevInside,
                                         uint32_t& rng, float& accDist, ShadeOut& o, Surf& s, const RB200InstanceProperties*& propsOut,
                                         bool& didRefract) {
    const float a1 = __uint_as_float(h.x), a2 = __uint_as_float(h.y);
    const RB200Instance* inst = &P.S.instances[h.w];
#if RB_INST_RECORDS
    const RB200InstanceProperties* props = &P.S.instProps[h.w];
#else
    const RB200InstanceProperties* props = &P.S.props[__ldg(&inst->instancePropertiesID)];
#endif

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

    o.skip = false; o.pdf = 0.0f; o.inside = prevInside;
    o.color = o.albedo = o.emission = o.normal = rb_splat3(0.0f);
    didRefract = false;        // always false when NEE sees it (disney.rchit.glsl:187)
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

    propsOut = props;
}

template <int MAT>
__device__ __forceinline__ void shade_slot(const WaveParams& P, const uint32_t slot, uint32_t* cnt, uint32_t* cntNext, int parity, LaneAcc& acc) {
    float4 ro4, rd4, T4;
    uint4 st;
    load_pair(P.rayO, P.rayD, slot, ro4, rd4);
    load_pair(P.thr, P.st, slot, T4, st);
    const rb_v3 rayOrigin = rb_mk3(ro4.x, ro4.y, ro4.z), rayDir = rb_mk3(rd4.x, rd4.y, rd4.z);
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
        // nothing else can add to this path (a miss casts no shadow ray, and the previous hit's shadow ray was resolved
        // in the previous wave), so the path ends here: about 60 % of all path ends skip endQ and k_finish.
        // Rays traced by this path: one per shaded segment plus the one that missed.
        finish_slot(P, slot, L, st, segments + 1u, cntNext, parity, acc);
        return;
    }

    uint4 h;
    float4 radIn;          // the path's radiance so far: the other half of the hit record's sector
    load_pair(P.hit, P.rad, slot, h, radIn);
    const bool prevInside = (flags & F_INSIDE) != 0u;
    ShadeOut o;
    Surf s;
    const RB200InstanceProperties* props;
    bool didRefract;
    eval_hit<MAT>(P, h, rayOrigin, rayDir, prevInside, rng, accDist, o, s, props, didRefract);

    // ---- what raygen does after traceRayEXT returns (rgen.glsl:124-181) ----
    store_pair(P.rayO, P.rayD, slot, make_float4(o.newO.x, o.newO.y, o.newO.z, 0.f), make_float4(o.newD.x, o.newD.y, o.newD.z, 0.f));
    const bool leftDielectric = !o.inside && prevInside;
    segments += 1u;
    uint32_t newFlags = (flags & (F_FIRST | F_PREVSKIP)) | (o.inside ? F_INSIDE : 0u);
    if (!o.skip) {
        if (!o.inside) {
            const rb_v3 indirect = o.emission;
            const bool skipNEE = !nee || (MAT != 0 && MAT != 3);
            if (skipNEE) {
                const float4 L4 = radIn;
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
                st.w += 1u;       // shadow rays of this path (counted into the batch when the path ends)
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
    store_pair(P.thr, P.st, slot, make_float4(T.x, T.y, T.z, accDist), make_uint4(rng, newFlags | (segments << 8), st.z, st.w));
    if (segments < P.pc.maxBounces) queue_push(P.rayQ[parity ^ 1], &cntNext[CNT_RAYS], slot);
    else queue_push(P.endQ, &cnt[CNT_END], slot);
}

// 5 resident blocks of 128 threads = 96 registers: with the 256-bit state accesses ptxas otherwise takes 102-108 and the
// fifth block is lost (Lambertian 3.48 -> 3.28 ms per step with the cap, no spill).
#ifndef RB_SHADE_MINBLOCKS
#define RB_SHADE_MINBLOCKS 5
#endif
#ifndef RB_SHADE_BLOCK
#define RB_SHADE_BLOCK 128       // threads per block of the shading kernels (85-119 registers: smaller blocks pack an SM's register file better)
#endif
// Disney: 128 registers without a cap (4 resident blocks). B200, headline step, ms per step with __launch_bounds__(128, n):
// n = 3 / 4 (no cap) 3.94, 5 (96 registers, 76 B spilled) 3.84, 6 (80 registers, 180 B spilled) 3.77: the kernel issues on
// 45 % of the cycles with 4 warps per scheduler, so two more resident blocks are worth more than the spills cost.
#ifndef RB_DISNEY_MINBLOCKS
#define RB_DISNEY_MINBLOCKS 6
#endif
// RB_SHADE_PREFETCH=1: software pipeline of the shading kernels' gathers — while item k is shaded, the path-state line
// of item k+2 and the shading records of item k+1 (whose hit record arrived with the line requested one iteration
// earlier) are requested into L2 with prefetch.global.L2. The idea: a queue item costs a chain of dependent fetches
// (queue entry -> 128-byte state line -> shading records) and the kernels wait on it (12 warps per scheduler on a long
// scoreboard per issue, 26 % of the issue slots used, DRAM at 1.7 of 6.5 TB/s, profiles/r02b). MEASURED ON B200: WORSE —
// Lambertian 3.84 -> 5.48 ms, Disney 4.05 -> 4.82 ms, miss 2.30 -> 3.22 ms per step, with or without a register cap
// that keeps the fifth resident block (profiles/r02d_variant_sweep.txt). Every prefetch of a scattered address is one
// more L1 -> L2 request per lane, and the request rate of that path, not the latency behind it, is what these kernels
// run into; the pipeline doubles the requests. Kept as a switch, off.
#ifndef RB_SHADE_PREFETCH
#define RB_SHADE_PREFETCH 0
#endif
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_state_line(const WaveParams& P, uint32_t slot, bool withSum) {
    const char* line = reinterpret_cast<const char*>(P.rayO.p + (size_t)slot * STATE_STRIDE);
    if (RB_PAIR_STATE == 2) {
        prefetch_l2(line); prefetch_l2(line + 32); prefetch_l2(line + 64);
        if (withSum) prefetch_l2(line + 96);
    }
}
template <int MAT>
__device__ __forceinline__ void prefetch_hit_records(const DeviceScene& S, uint32_t tri, uint32_t instance) {
#if RB_SHADE_RECORDS && RB_INST_RECORDS
    const char* sb = reinterpret_cast<const char*>(S.shadeBase + 4 * (size_t)tri);
    prefetch_l2(sb); prefetch_l2(sb + 32);
    const RB200InstanceProperties* props = &S.instProps[instance];
    const bool tbn = MAT == 3 || __ldg(&props->normalMapTexID) >= 0 || __ldg(&props->bumpMapTexID) >= 0;
    if (tbn || __ldg(&props->interpNormals) != 0u) {
        const char* sf = reinterpret_cast<const char*>(S.shadeFrame + 8 * (size_t)tri);
        prefetch_l2(sf); prefetch_l2(sf + 32);
        if (tbn) { prefetch_l2(sf + 64); prefetch_l2(sf + 96); }
    }
#endif
}

template <int MAT>
__global__ void __launch_bounds__(RB_SHADE_BLOCK, MAT == 3 ? RB_DISNEY_MINBLOCKS : RB_SHADE_MINBLOCKS) k_shade(WaveParams P, int parity) {
    uint32_t* cnt = P.counters + parity * CNT_SET;
    uint32_t* cntNext = P.counters + (parity ^ 1) * CNT_SET;
    const uint32_t n = cnt[CNT_MAT0 + MAT];
    const uint32_t* __restrict__ q = P.matQ[MAT];
    __shared__ LaneAcc acc;           // only the miss shader ends paths
    if (MAT == 4) lane_acc_init(acc);
#if RB_SHADE_PREFETCH
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t NONE = 0xFFFFFFFFu;
    uint32_t slot = i < n ? q[i] : NONE;
    uint32_t slotNext = i + stride < n && i + stride >= i ? q[i + stride] : NONE;
    if (slotNext != NONE) prefetch_state_line(P, slotNext, MAT == 4);
    for (; slot != NONE; i += stride) {
        const uint32_t i2 = i + 2u * stride;
        uint32_t slotNext2 = NONE;
        if (i2 < n && i2 >= i) { slotNext2 = q[i2]; prefetch_state_line(P, slotNext2, MAT == 4); }
        uint4 hitNext = make_uint4(0u, 0u, 0u, 0u);
        if (MAT != 4 && slotNext != NONE) hitNext = P.hit[slotNext];      // requested now, needed after this item is shaded
        shade_slot<MAT>(P, slot, cnt, cntNext, parity, acc);
        if (MAT != 4 && slotNext != NONE) prefetch_hit_records<MAT>(P.S, hitNext.z, hitNext.w);
        slot = slotNext; slotNext = slotNext2;
    }
#else
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        shade_slot<MAT>(P, q[i], cnt, cntNext, parity, acc);
#endif
    if (MAT == 4) lane_acc_flush(P, acc);
}

// ---------------------------------------------------------------------------------------------------
// shadow: shadowRayOccluded (nee.h.glsl:126-144) + the radiance update of rgen.glsl:174-177
// ---------------------------------------------------------------------------------------------------
template <bool COUNT, bool SKIPNULL = false>
__global__ void __launch_bounds__(RB_SHADOW_BLOCK, RB_SHADOW_MINBLOCKS) k_shadow(WaveParams P, int parity) {
    uint32_t* cnt = P.counters + parity * CNT_SET;
    const uint32_t n = cnt[CNT_SHADOW];
    uint32_t nodeVisits = 0, triTests = 0;
    // RB200_FLAG_SKIP_NULL_SHADOW_RAYS: a record whose `direct` term is +0 in every channel adds the same radiance whether
    // the light is visible or not (commit below: direct * wNEE with direct = occluded ? 0 : +0). Such a ray is handed to the
    // traversal with tmax = 0, for which Traversal::init queues no node work: it is committed as "not occluded" at the
    // end of the iteration, through the same code as every other ray.
    // (a separate instantiation, so that the kernel of the default path is exactly the one without this code)
    constexpr bool skipNull = SKIPNULL;
    uint32_t skipped = 0;
    extern __shared__ __align__(16) unsigned char rb_dyn_smem[];      // WarpShared<true>[RB_SHADOW_BLOCK / 32]
    WarpShared<true>* ws = reinterpret_cast<WarpShared<true>*>(rb_dyn_smem);
    trace_queue<true, COUNT>(
        P.S.nodes, P.S.tris, n, &cnt[CNT_CURSOR_SHADOW],
        [&](uint32_t i, rb_v3& o, rb_v3& d, float& tmax) {
#if RB_SHADOW_LDCS
            const float4 o4 = __ldcs(P.shO.p + i), d4 = __ldcs(P.shD.p + i);
#else
            const float4 o4 = P.shO[i], d4 = P.shD[i];
#endif
            o = rb_mk3(o4.x, o4.y, o4.z); d = rb_mk3(d4.x, d4.y, d4.z); tmax = o4.w;
            if (skipNull) {
                const float4 A = P.shA[i];
                if ((__float_as_uint(A.x) | __float_as_uint(A.y) | __float_as_uint(A.z)) == 0u) { tmax = 0.0f; skipped++; }
            }
        },
        [&](uint32_t i, const RayHit& h) {
            const bool occluded = h.tri != 0xFFFFFFFFu;
#if RB_SHADOW_LDCS
            const float4 A = __ldcs(P.shA.p + i), B = __ldcs(P.shB.p + i), T = __ldcs(P.shT.p + i);
#else
            const float4 A = P.shA[i], B = P.shB[i], T = P.shT[i];
#endif
            const uint32_t slot = __float_as_uint(B.w);
            const rb_v3 direct = occluded ? rb_splat3(0.0f) : rb_mk3(A.x, A.y, A.z);
            const rb_v3 combined = direct * A.w + rb_mk3(B.x, B.y, B.z);
            const float4 L4 = P.rad[slot];
            const rb_v3 L = rb_mk3(L4.x, L4.y, L4.z) + combined * rb_mk3(T.x, T.y, T.z);
            P.rad[slot] = make_float4(L.x, L.y, L.z, 0.f);
        },
        nodeVisits, triTests, ws[threadIdx.x >> 5],
        [&](uint32_t i) { return i; },
        [&](uint32_t i) { prefetch_l1(reinterpret_cast<const char*>(P.shO.p + i)); prefetch_l1(reinterpret_cast<const char*>(P.shD.p + i)); });
    if (skipNull) {
        skipped = __reduce_add_sync(0xffffffffu, skipped);
        if ((threadIdx.x & 31u) == 0u && skipped) atomicAdd(P.nullShadow, (unsigned long long)skipped);
    }
    if (COUNT) {      // the counting pass runs one lane (context_create), so lane 0's counters are the batch's
        atomicAdd(&P.stats[ST_NODES_SHADOW], (unsigned long long)nodeVisits);
        atomicAdd(&P.stats[ST_TRIS_SHADOW], (unsigned long long)triTests);
    }
}

// ---------------------------------------------------------------------------------------------------
// two-level mode (RB200_FLAG_TWO_LEVEL, two_level.cuh): the same two stages, one ray per thread over the same queues
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK) k_extend_two_level(WaveParams P, int parity) {
    uint32_t* cnt = P.counters + parity * CNT_SET;
    const uint32_t n = cnt[CNT_RAYS];
    const uint32_t* __restrict__ q = P.rayQ[parity];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t slot = q[i];
        float4 o4, d4;
        load_pair(P.rayO, P.rayD, slot, o4, d4);
        const TLHit h = trace_two_level<false>(P.S, rb_mk3(o4.x, o4.y, o4.z), rb_mk3(d4.x, d4.y, d4.z), 10000.0f);
        uint32_t bin = 4;
        if (h.tri != 0xFFFFFFFFu) {
            P.hit[slot] = make_uint4(__float_as_uint(h.b1), __float_as_uint(h.b2), h.tri, h.inst);
            bin = __ldg(&P.S.tlInstances[h.inst].material);
        }
        const uint32_t peers = __match_any_sync(__activemask(), bin);
        const uint32_t lane = threadIdx.x & 31u;
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if ((int)lane == leader) base = atomicAdd(&cnt[CNT_MAT0 + bin], (uint32_t)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        uint32_t* mq = bin == 0u ? P.matQ[0] : bin == 1u ? P.matQ[1] : bin == 2u ? P.matQ[2] : bin == 3u ? P.matQ[3] : P.matQ[4];
        mq[base + __popc(peers & ((1u << lane) - 1u))] = slot;
    }
}

__global__ void __launch_bounds__(BLOCK) k_shadow_two_level(WaveParams P, int parity) {
    uint32_t* cnt = P.counters + parity * CNT_SET;
    const uint32_t n = cnt[CNT_SHADOW];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 o4 = P.shO[i], d4 = P.shD[i];
        const TLHit h = trace_two_level<true>(P.S, rb_mk3(o4.x, o4.y, o4.z), rb_mk3(d4.x, d4.y, d4.z), o4.w);
        const bool occluded = h.tri != 0xFFFFFFFFu;
        const float4 A = P.shA[i], B = P.shB[i], T = P.shT[i];
        const uint32_t slot = __float_as_uint(B.w);
        const rb_v3 direct = occluded ? rb_splat3(0.0f) : rb_mk3(A.x, A.y, A.z);
        const rb_v3 combined = direct * A.w + rb_mk3(B.x, B.y, B.z);
        const float4 L4 = P.rad[slot];
        const rb_v3 L = rb_mk3(L4.x, L4.y, L4.z) + combined * rb_mk3(T.x, T.y, T.z);
        P.rad[slot] = make_float4(L.x, L.y, L.z, 0.f);
    }
}

// ---------------------------------------------------------------------------------------------------
// finish: end of a path (rgen.glsl:264-284)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK) k_finish(WaveParams P, int parity) {
    uint32_t* cnt = P.counters + parity * CNT_SET;
    uint32_t* cntNext = P.counters + (parity ^ 1) * CNT_SET;
    const uint32_t n = cnt[CNT_END];
    __shared__ LaneAcc acc;
    lane_acc_init(acc);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t slot = P.endQ[i];
        const float4 L4 = P.rad[slot];
        const uint4 st = P.st[slot];
        finish_slot(P, slot, rb_mk3(L4.x, L4.y, L4.z), st, st.y >> 8, cntNext, parity, acc);
    }
    lane_acc_flush(P, acc);
}

// rgen.glsl:277-284: running average over batches (or the plain sum with RB200_FLAG_ACCUM_SUM). Runs once per batch,
// in batch order, after every pixel of the batch has delivered its mean.
__global__ void __launch_bounds__(BLOCK) k_accumulate(float4* __restrict__ image, const float4* __restrict__ mean, uint32_t n,
                                                      uint32_t sampleBatch, uint32_t flags,
                                                      const unsigned long long* __restrict__ laneStats,
                                                      unsigned long long* __restrict__ cumStats,
                                                      unsigned long long* __restrict__ lastStats, float4* __restrict__ peer) {
    // peer (latency mode of a one-process group, group.cu): device 0's image. Every pixel has exactly one owner, so the
    // owner stores its new value there as well — plain stores over NVLink from the kernel that produces the value, instead
    // of a separate reduce of images that are zero outside each device's tiles.
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    // batches are folded one at a time, in order: no race (the skipped-shadow-ray slot of the cumulative array is k_shadow's)
    if (i < ST_COUNT && i != ST_SHADOW_SKIPPED) { const unsigned long long v = laneStats[i]; cumStats[i] += v; lastStats[i] = v; }
    if (i >= n) return;
    const float4 m = mean[i];
    if (m.w < 0.f) return;       // tile partition: another rank's pixel
    if (m.w == 0.f) {
        // documented deviation: the reference writes 0/0 and poisons the pixel; batch 0 writes black, later batches keep
        // the previous value, sum mode adds nothing
        if (!(flags & RB200_FLAG_ACCUM_SUM) && sampleBatch == 0u) {
            image[i] = make_float4(0.f, 0.f, 0.f, 1.f);
            if (peer) peer[i] = make_float4(0.f, 0.f, 0.f, 1.f);
        }
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
        if (peer) peer[i] = make_float4(fin.x, fin.y, fin.z, 1.f);
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
static constexpr size_t SMEM_EXTEND = sizeof(WarpShared<false>) * (RB_EXTEND_BLOCK / 32);
static constexpr size_t SMEM_SHADOW = sizeof(WarpShared<true>) * (RB_SHADOW_BLOCK / 32);

// Per-triangle shading records (DeviceScene::shadeBase / shadeFrame): one thread per triangle slot copies exactly the
// floats hit_info() would gather for that triangle.
__global__ void k_build_shade_records(DeviceScene S, float4* __restrict__ base, float4* __restrict__ frame) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S.numTris) return;
    const float4* tp = reinterpret_cast<const float4*>(S.tris + t);
    const uint32_t prim = __float_as_uint(tp[0].w), inst = __float_as_uint(tp[1].w) & TRI_INST_MASK;
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

// ---------------------------------------------------------------------------------------------------
// the wave loop of an engine (see context.cuh: engines, lanes and speculation)
// ---------------------------------------------------------------------------------------------------
struct CallTimer {          // RB200_FLAG_TIME_KERNELS: one event pair per kernel of the call
    RB200Context* ctx; cudaStream_t s; bool on;
    void tic(int cls) {
        if (!on) return;
        while (ctx->evPool.size() < ctx->evUsed + 2) { cudaEvent_t e; cudaEventCreate(&e); ctx->evPool.push_back(e); }
        ctx->evClass.push_back(cls);
        cudaEventRecord(ctx->evPool[ctx->evUsed], s);
    }
    void toc() {
        if (!on) return;
        cudaEventRecord(ctx->evPool[ctx->evUsed + 1], s);
        ctx->evUsed += 2;
    }
};

// Issues `count` waves of engine E on its stream, starting with queue set `parity`. staggerAt (1-based wave within this
// span, 0 = none) records E.staggerEv behind that wave. `waveBase`: index of the first wave in the call (timing pass).
static int issue_waves(RB200Context* ctx, Engine& E, uint32_t count, uint32_t parity, uint32_t staggerAt, bool capturing,
                       CallTimer& tm, uint32_t waveBase, std::vector<uint32_t>* waveLog) {
    WaveParams& P = E.P;
    cudaStream_t s = E.stream;
    const bool countBvh = (ctx->flags & RB200_FLAG_COUNT_BVH) != 0;
    const bool nee = (ctx->flags & RB200_FLAG_NEE) != 0;
    const uint32_t mats = P.S.materialMask & 15u;
    const bool twoLevel = P.S.tlasNodes != nullptr;
    for (uint32_t w = 0; w < count; w++) {
        const int p = (int)((parity + w) & 1u);
        RB_CUDA(cudaMemsetAsync(P.counters + (p ^ 1) * CNT_SET, 0, CNT_SET * sizeof(uint32_t), s));
        tm.tic(1);
        if (twoLevel) k_extend_two_level<<<ctx->gTwoLevel[0], BLOCK, 0, s>>>(P, p);
        else if (countBvh) k_extend<true><<<ctx->gExtendC, RB_EXTEND_BLOCK, SMEM_EXTEND, s>>>(P, p);
        else k_extend<false><<<ctx->gExtend, RB_EXTEND_BLOCK, SMEM_EXTEND, s>>>(P, p);
        tm.toc();
        tm.tic(6); k_shade<4><<<ctx->gShade[4], RB_SHADE_BLOCK, 0, s>>>(P, p); tm.toc();
        // a material no instance uses has an empty queue in every wave: its kernel is not launched
        if (mats & 1u) { tm.tic(2); k_shade<0><<<ctx->gShade[0], RB_SHADE_BLOCK, 0, s>>>(P, p); tm.toc(); }
        if (mats & 2u) { tm.tic(3); k_shade<1><<<ctx->gShade[1], RB_SHADE_BLOCK, 0, s>>>(P, p); tm.toc(); }
        if (mats & 4u) { tm.tic(4); k_shade<2><<<ctx->gShade[2], RB_SHADE_BLOCK, 0, s>>>(P, p); tm.toc(); }
        if (mats & 8u) { tm.tic(5); k_shade<3><<<ctx->gShade[3], RB_SHADE_BLOCK, 0, s>>>(P, p); tm.toc(); }
        if (nee) {
            tm.tic(7);
            if (twoLevel) k_shadow_two_level<<<ctx->gTwoLevel[1], BLOCK, 0, s>>>(P, p);
            else if (ctx->flags & RB200_FLAG_SKIP_NULL_SHADOW_RAYS) {
                if (countBvh) k_shadow<true, true><<<ctx->gShadowC, RB_SHADOW_BLOCK, SMEM_SHADOW, s>>>(P, p);
                else k_shadow<false, true><<<ctx->gShadow, RB_SHADOW_BLOCK, SMEM_SHADOW, s>>>(P, p);
            }
            else if (countBvh) k_shadow<true><<<ctx->gShadowC, RB_SHADOW_BLOCK, SMEM_SHADOW, s>>>(P, p);
            else k_shadow<false><<<ctx->gShadow, RB_SHADOW_BLOCK, SMEM_SHADOW, s>>>(P, p);
            tm.toc();
        }
        tm.tic(8); k_finish<<<ctx->gFinish, BLOCK, 0, s>>>(P, p); tm.toc();
        if (tm.on && ctx->waveCountsDev && waveBase + w < ctx->waveCountsCap)      // the queue counters of this wave, device to device: no host wait
            RB_CUDA(cudaMemcpyAsync(ctx->waveCountsDev + (size_t)CNT_SET * (waveBase + w), P.counters + p * CNT_SET, CNT_SET * sizeof(uint32_t),
                                    cudaMemcpyDeviceToDevice, s));
        if (staggerAt && w + 1 == staggerAt)
            RB_CUDA(cudaEventRecordWithFlags(E.staggerEv, s, capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
        if (waveLog) {
            waveLog->resize((size_t)(waveBase + w + 1) * CNT_SET);
            RB_CUDA(cudaMemcpyAsync(&(*waveLog)[(size_t)(waveBase + w) * CNT_SET], P.counters + p * CNT_SET, CNT_SET * sizeof(uint32_t),
                                    cudaMemcpyDeviceToHost, s));
            RB_CUDA(cudaStreamSynchronize(s));
        }
    }
    RB_CUDA(cudaGetLastError());
    return RB200_OK;
}

static uint32_t kernels_per_wave(const RB200Context* ctx, const Engine& E) {
    return ((ctx->flags & RB200_FLAG_NEE) ? 4u : 3u) + (uint32_t)__builtin_popcount(E.P.S.materialMask & 15u);
}

static void drop_graphs(Engine& E) {
    for (WaveGraph& g : E.graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    E.graphs.clear();
}

// `count` waves as graph launches: chunks of RB_GRAPH_CHUNK waves (a few hundred stream operations become one launch, and
// the device-side gap between dependent kernels shrinks); a chunk is captured once per (length, parity, stagger position)
// and reused until the engine's key (scene, camera, sample counts) changes.
static int run_waves(RB200Context* ctx, Engine& E, uint32_t count, uint32_t staggerAt, CallTimer& tm, uint32_t waveBase,
                     std::vector<uint32_t>* waveLog) {
    static const bool graphsOff = getenv("RB200_NO_GRAPH") != nullptr;
    const bool direct = tm.on || waveLog || graphsOff || (ctx->flags & RB200_FLAG_COUNT_BVH);
    uint32_t done = 0;
    while (done < count) {
        const uint32_t c = std::min<uint32_t>(count - done, RB_GRAPH_CHUNK);
        const uint32_t parity = E.globalWave & 1u;
        const uint32_t st = (staggerAt > done && staggerAt <= done + c) ? staggerAt - done : 0u;
        if (direct) {
            const int rc = issue_waves(ctx, E, c, parity, st, false, tm, waveBase + done, waveLog);
            if (rc != RB200_OK) return rc;
        } else {
            WaveGraph* g = nullptr;
            for (WaveGraph& k : E.graphs) if (k.waves == c && k.parity == parity && k.staggerAt == st) { g = &k; break; }
            if (!g) {
                cudaGraph_t graph = nullptr;
                RB_CUDA(cudaStreamBeginCapture(E.stream, cudaStreamCaptureModeThreadLocal));
                const int rc = issue_waves(ctx, E, c, parity, st, true, tm, 0, nullptr);
                const cudaError_t ce = cudaStreamEndCapture(E.stream, &graph);
                if (rc != RB200_OK || ce != cudaSuccess || !graph) {
                    cudaGetLastError();
                    if (graph) cudaGraphDestroy(graph);
                    set_error("CUDA graph capture of the wave loop failed: %s", cudaGetErrorString(ce));
                    return RB200_ERR_CUDA;
                }
                WaveGraph ng; ng.waves = c; ng.parity = parity; ng.staggerAt = st;
                const cudaError_t ie = cudaGraphInstantiate(&ng.exec, graph, 0);
                cudaGraphDestroy(graph);
                if (ie != cudaSuccess) { cudaGetLastError(); set_error("cudaGraphInstantiate: %s", cudaGetErrorString(ie)); return RB200_ERR_CUDA; }
                E.graphs.push_back(ng);
                g = &E.graphs.back();
                ctx->graphCaptures++;
            }
            RB_CUDA(cudaGraphLaunch(g->exec, E.stream));
        }
        E.globalWave += c;
        done += c;
    }
    return RB200_OK;
}

// Discards every batch in flight on the engine (their rays are dropped with the queue counters).
static int reset_engine(Engine& E) {
    for (int l = 0; l < E.numLanes; l++) {
        if (E.lane[l].active) E.wastedBatches++;
        E.lane[l] = LaneState();
    }
    E.head = 0; E.active = 0; E.globalWave = 0;
    RB_CUDA(cudaMemsetAsync(E.P.counters, 0, 2 * CNT_SET * sizeof(uint32_t), E.stream));
    return RB200_OK;
}

void invalidate_speculation(RB200Context* ctx) {
    for (int e = 0; e < ctx->numEngines; e++) {
        Engine& E = ctx->eng[e];
        cudaStreamSynchronize(E.stream);
        reset_engine(E);
        E.keyValid = false; E.havePrev = false; E.streak = 0; E.stride = 1;
        drop_graphs(E);
    }
}

// Starts batch `sampleBatch` in the next free lane of the engine (behind the batches already in flight).
static int inject_batch(RB200Context* ctx, Engine& E, uint32_t sampleBatch, CallTimer& tm, uint64_t& nl) {
    const int l = (E.head + E.active) % E.numLanes;
    E.lane[l].active = true; E.lane[l].sampleBatch = sampleBatch; E.lane[l].wavesDone = 0;
    E.active++;
    WaveParams& P = E.P;
    RB_CUDA(cudaMemsetAsync(P.stats + (size_t)l * ST_COUNT, 0, ST_COUNT * sizeof(unsigned long long), E.stream));
    tm.tic(0);
    k_generate<<<(P.N + BLOCK - 1) / BLOCK, BLOCK, 0, E.stream>>>(P, (uint32_t)l, sampleBatch, (int)(E.globalWave & 1u));
    tm.toc();
    nl++;
    RB_CUDA(cudaGetLastError());
    return RB200_OK;
}

int render_batch(RB200Context* ctx, const RB200Scene* scene, const RB200RtPushConsts* pc) {
    if (pc->samplesPerPixel == 0 || pc->maxBounces == 0) { set_error("samplesPerPixel and maxBounces must be > 0"); return RB200_ERR_INVALID_ARGUMENT; }
    const uint64_t waves64 = (uint64_t)pc->samplesPerPixel * pc->maxBounces;
    if (waves64 > (1ull << 22)) {
        set_error("samplesPerPixel * maxBounces = %llu waves per batch: more than 2^22 (split the samples over several batches)",
                  (unsigned long long)waves64);
        return RB200_ERR_INVALID_ARGUMENT;
    }
    if ((ctx->flags & RB200_FLAG_NEE) && scene->numEmissive == 0) {
        set_error("Scene must have at least one emissive object");   // src/scene/Instances.cpp:125-127
        return RB200_ERR_NO_EMITTER;
    }
    const uint32_t maxWaves = (uint32_t)waves64;
    const int e = (int)(ctx->batchCalls % (uint64_t)ctx->numEngines);
    const int prevEngine = ctx->lastEngine;
    ctx->batchCalls++;
    ctx->lastEngine = e;
    Engine& E = ctx->eng[e];
    E.calls++;
    WaveParams& P = E.P;
    P.S = scene->dev;
    P.pc = *pc;
    cudaStream_t s = E.stream;
    // everything the caller enqueued on the front-end stream before this call (write_hdr, postprocess of the previous
    // frame, an external reduce of the image ...) must precede this batch's accumulation — not its tracing
    RB_CUDA(cudaEventRecord(ctx->frontMark, ctx->stream));
    const bool timed = (ctx->flags & RB200_FLAG_TIME_KERNELS) != 0;
    if (timed)      // timing pass: no overlap between engines
        for (int o = 0; o < ctx->numEngines; o++) if (o != e) RB_CUDA(cudaStreamSynchronize(ctx->eng[o].stream));
    CallTimer tm{ctx, s, timed};
    ctx->evUsed = 0; ctx->evClass.clear();

    // ---- does this call continue the sequence the engine predicted? ----
    WaveParams key = P;
    key.pc.sampleBatch = 0u;
    const bool sameKey = E.keyValid && E.scene == scene && memcmp(&key, &E.key, sizeof(WaveParams)) == 0;
    if (!sameKey) drop_graphs(E);
    const uint32_t delta = pc->sampleBatch - E.prevBatch;        // modulo 2^32, like the seed
    if (sameKey && E.havePrev && delta == E.stride) E.streak++;
    else {
        E.streak = 0;
        E.stride = (sameKey && E.havePrev && delta != 0u && delta <= 65536u) ? delta : 1u;     // learn the stride (ranks of a sample split)
    }
    bool inFlightOk = sameKey && E.active > 0 && E.lane[E.head].sampleBatch == pc->sampleBatch;
    for (int k = 1; inFlightOk && k < E.active; k++)
        inFlightOk = E.lane[(E.head + k) % E.numLanes].sampleBatch == pc->sampleBatch + (uint32_t)k * E.stride;
    uint64_t nl = 0;
    if (!inFlightOk) {
        int rc = reset_engine(E);
        if (rc != RB200_OK) return rc;
        if ((rc = inject_batch(ctx, E, pc->sampleBatch, tm, nl)) != RB200_OK) return rc;
    }
    memcpy(&E.key, &key, sizeof(WaveParams));
    E.keyValid = true; E.scene = scene;

    // Speculate from the second call of a regular sequence on; never in the counting pass (its counters are per call).
    const bool speculate = E.numLanes > 1 && E.streak >= 1u && !(ctx->flags & RB200_FLAG_COUNT_BVH) && !getenv("RB200_NO_SPECULATION");
    const uint32_t span = speculate ? (maxWaves + (uint32_t)E.numLanes - 1u) / (uint32_t)E.numLanes : maxWaves;

    // engine stagger (one lane per engine only): start behind wave `staggerWave` of the previous batch
    const bool stagger = ctx->numEngines > 1 && E.numLanes == 1 && ctx->staggerWave >= 0 && !timed;
    if (stagger && prevEngine >= 0 && prevEngine != e) RB_CUDA(cudaStreamWaitEvent(s, ctx->eng[prevEngine].staggerEv, 0));
    const uint32_t staggerWant = ctx->staggerWave > 0 ? (uint32_t)ctx->staggerWave : std::max(1u, maxWaves * 5u / 32u);
    uint32_t staggerAt = stagger ? std::min(staggerWant, maxWaves) : 0u;

    // developer aid: RB200_WAVE_LOG=<file> (with RB200_FLAG_TIME_KERNELS) writes one CSV row per wave of the last call —
    // queue counters and the device time of every kernel; it synchronises after every wave
    const char* waveLogPath = timed ? getenv("RB200_WAVE_LOG") : nullptr;
    std::vector<uint32_t> waveCounters;
    if (timed) {
        if (ctx->waveCountsCap < maxWaves) {
            if (ctx->waveCountsDev) cudaFree(ctx->waveCountsDev);
            ctx->waveCountsDev = nullptr; ctx->waveCountsCap = 0;
            RB_CUDA(cudaMalloc(&ctx->waveCountsDev, (size_t)maxWaves * CNT_SET * sizeof(uint32_t)));
            ctx->waveCountsCap = maxWaves;
        }
    }

    uint32_t issued = 0;
    while (E.lane[E.head].wavesDone < maxWaves) {
        if (speculate && E.active < E.numLanes) {
            const LaneState& newest = E.lane[(E.head + E.active - 1) % E.numLanes];
            if (newest.wavesDone >= span) {
                const int rc = inject_batch(ctx, E, newest.sampleBatch + E.stride, tm, nl);
                if (rc != RB200_OK) return rc;
            }
        }
        const uint32_t n = std::min(span, maxWaves - E.lane[E.head].wavesDone);
        const uint32_t st = (staggerAt > issued && staggerAt <= issued + n) ? staggerAt - issued : 0u;
        const int rc = run_waves(ctx, E, n, st, tm, issued, waveLogPath ? &waveCounters : nullptr);
        if (rc != RB200_OK) return rc;
        issued += n;
        for (int k = 0; k < E.active; k++) E.lane[(E.head + k) % E.numLanes].wavesDone += n;
    }
    nl += (uint64_t)issued * kernels_per_wave(ctx, E);
    if (timed) ctx->waveCountsWaves = std::min(issued, ctx->waveCountsCap);

    if (waveLogPath) {
        if (FILE* f = fopen(waveLogPath, "w")) {
            fprintf(f, "wave,rays,lambertian,metal,dielectric,disney,miss,shadow,end,extend_us,miss_us,lambertian_us,metal_us,dielectric_us,disney_us,shadow_us,finish_us\n");
            size_t ev = 0;
            while (ev < ctx->evClass.size() && ctx->evClass[ev] == 0) ev++;      // k_generate pairs come first
            for (uint32_t w = 0; w < issued && (size_t)(w + 1) * CNT_SET <= waveCounters.size(); w++) {
                const uint32_t* c = &waveCounters[(size_t)w * CNT_SET];
                fprintf(f, "%u,%u,%u,%u,%u,%u,%u,%u,%u", w, c[CNT_RAYS], c[CNT_MAT0], c[CNT_MAT0 + 1], c[CNT_MAT0 + 2], c[CNT_MAT0 + 3],
                        c[CNT_MISS], c[CNT_SHADOW], c[CNT_END]);
                // event classes: 1 extend, 6 miss, 2..5 materials, 7 shadow, 8 finish; columns in that order
                float us[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                while (ev < ctx->evClass.size()) {
                    const int cls = ctx->evClass[ev];
                    float ms = 0.f;
                    cudaEventElapsedTime(&ms, ctx->evPool[2 * ev], ctx->evPool[2 * ev + 1]);
                    ev++;
                    if (cls == 0) continue;       // a speculative batch was started between two waves
                    us[cls] = ms * 1000.f;
                    if (cls == 8) break;
                }
                fprintf(f, ",%.1f,%.1f,%.1f,%.1f,%.1f,%.1f,%.1f,%.1f\n", us[1], us[6], us[2], us[3], us[4], us[5], us[7], us[8]);
            }
            fclose(f);
        }
    }

    // fold the head batch's pixel means into the shared image: after the previous call's fold (another engine's stream)
    // and after whatever the caller had queued on the front-end stream; then let the front-end stream see the result
    const int hl = E.head;
    if (prevEngine >= 0 && prevEngine != e) RB_CUDA(cudaStreamWaitEvent(s, ctx->eng[prevEngine].accumDone, 0));
    RB_CUDA(cudaStreamWaitEvent(s, ctx->frontMark, 0));
    k_accumulate<<<(P.N + BLOCK - 1) / BLOCK, BLOCK, 0, s>>>(P.image, P.mean.p + (size_t)hl * P.N, P.N, pc->sampleBatch, ctx->flags,
                                                             P.stats + (size_t)hl * ST_COUNT, ctx->statsSnap, ctx->statsLast,
                                                             (ctx->flags & RB200_FLAG_ACCUM_SUM) ? nullptr : ctx->peerImage); nl++;
    RB_CUDA(cudaEventRecord(E.accumDone, s));
    RB_CUDA(cudaStreamWaitEvent(ctx->stream, E.accumDone, 0));
    RB_CUDA(cudaGetLastError());
    E.lane[hl] = LaneState();
    E.head = (E.head + 1) % E.numLanes;
    E.active--;
    E.havePrev = true; E.prevBatch = pc->sampleBatch;
    ctx->last.waves = issued;
    ctx->last.kernelLaunches = nl;
    ctx->cumulative.waves += issued;
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
                r.instance = __float_as_uint(__ldg(tp + 1).w) & TRI_INST_MASK;
            } else { r.t = -1.0f; r.u = r.v = 0.f; r.primitive = r.instance = 0xFFFFFFFFu; }
            out[i] = r;
        },
        nv, tt, ws[threadIdx.x >> 5]);
}

template <bool ANY>
__global__ void __launch_bounds__(BLOCK) k_trace_query_two_level(DeviceScene S, uint32_t n, const float4* __restrict__ o,
                                                                 const float4* __restrict__ d, RB200PrimaryHit* __restrict__ out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 o4 = o[i], d4 = d[i];
        const TLHit h = trace_two_level<ANY>(S, rb_mk3(o4.x, o4.y, o4.z), rb_mk3(d4.x, d4.y, d4.z), o4.w);
        RB200PrimaryHit r;
        if (h.tri != 0xFFFFFFFFu) {
            r.t = h.t; r.u = h.b1; r.v = h.b2; r.instance = h.inst;
            r.primitive = __float_as_uint(__ldg(reinterpret_cast<const float4*>(S.tris + h.tri)).w);
        } else { r.t = -1.0f; r.u = r.v = 0.f; r.primitive = r.instance = 0xFFFFFFFFu; }
        out[i] = r;
    }
}

static int sync_engines(RB200Context* ctx) {
    for (int e = 0; e < ctx->numEngines; e++) RB_CUDA(cudaStreamSynchronize(ctx->eng[e].stream));
    return RB200_OK;
}

struct DeviceBuf {          // frees on every return path
    void* p = nullptr;
    ~DeviceBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
};

static int launch_query(RB200Context* ctx, const RB200Scene* scene, uint32_t n, const float4* dO, const float4* dD, int any,
                        RB200PrimaryHit* dOut) {
    RB_CUDA(cudaMemsetAsync(ctx->queryCursor, 0, sizeof(uint32_t), ctx->stream));
    if (scene->dev.tlasNodes) {
        const int g2 = (int)std::min<uint64_t>((n + BLOCK - 1) / BLOCK, (uint64_t)ctx->gTwoLevel[any ? 1 : 0]);
        if (any) k_trace_query_two_level<true><<<g2, BLOCK, 0, ctx->stream>>>(scene->dev, n, dO, dD, dOut);
        else k_trace_query_two_level<false><<<g2, BLOCK, 0, ctx->stream>>>(scene->dev, n, dO, dD, dOut);
        ctx->launches++;
        RB_CUDA(cudaGetLastError());
        return RB200_OK;
    }
    const int grid = (int)std::min<uint64_t>((n + BLOCK - 1) / BLOCK, (uint64_t)ctx->gQuery[any ? 1 : 0]);
    constexpr size_t smAny = sizeof(WarpShared<true>) * (BLOCK / 32), smClosest = sizeof(WarpShared<false>) * (BLOCK / 32);
    if (any) k_trace_query<true><<<grid, BLOCK, smAny, ctx->stream>>>(scene->dev.nodes, scene->dev.tris, n, dO, dD, dOut, ctx->queryCursor);
    else k_trace_query<false><<<grid, BLOCK, smClosest, ctx->stream>>>(scene->dev.nodes, scene->dev.tris, n, dO, dD, dOut, ctx->queryCursor);
    ctx->launches++;
    RB_CUDA(cudaGetLastError());
    return RB200_OK;
}

static int run_query(RB200Context* ctx, const RB200Scene* scene, uint32_t n, const float4* dO, const float4* dD, int any,
                     RB200PrimaryHit* out) {
    DeviceBuf dOut;
    RB_CUDA(dOut.alloc((size_t)n * sizeof(RB200PrimaryHit)));
    const int rc = launch_query(ctx, scene, n, dO, dD, any, static_cast<RB200PrimaryHit*>(dOut.p));
    if (rc != RB200_OK) return rc;
    RB_CUDA(cudaMemcpyAsync(out, dOut.p, (size_t)n * sizeof(RB200PrimaryHit), cudaMemcpyDeviceToHost, ctx->stream));
    RB_CUDA(cudaStreamSynchronize(ctx->stream));
    return RB200_OK;
}

int trace_primary(RB200Context* ctx, const RB200Scene* scene, const RB200RtPushConsts* pc, RB200PrimaryHit* out) {
    const int src = sync_engines(ctx);
    if (src != RB200_OK) return src;
    // engine 0's shadow-record arrays are the scratch for the camera rays: between two calls no wave is running and the
    // records of a wave do not outlive it
    WaveParams P = ctx->wp;
    P.S = scene->dev; P.pc = *pc;
    k_primary_rays<<<(P.N + BLOCK - 1) / BLOCK, BLOCK, 0, ctx->stream>>>(P, P.shO.p, P.shD.p);
    ctx->launches++;
    return run_query(ctx, scene, P.N, P.shO.p, P.shD.p, 0, out);
}

static int upload_rays(RB200Context* ctx, uint32_t n, const float* o, const float* d, const float* tmax, DeviceBuf& dO, DeviceBuf& dD) {
    std::vector<float4> ho(n), hd(n);
    for (uint32_t i = 0; i < n; i++) {
        ho[i] = make_float4(o[3 * i], o[3 * i + 1], o[3 * i + 2], tmax[i]);
        hd[i] = make_float4(d[3 * i], d[3 * i + 1], d[3 * i + 2], 0.f);
    }
    RB_CUDA(dO.alloc((size_t)n * sizeof(float4)));
    RB_CUDA(dD.alloc((size_t)n * sizeof(float4)));
    RB_CUDA(cudaMemcpyAsync(dO.p, ho.data(), (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    RB_CUDA(cudaMemcpyAsync(dD.p, hd.data(), (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    RB_CUDA(cudaStreamSynchronize(ctx->stream));       // the staging vectors go out of scope
    return RB200_OK;
}

int trace_rays(RB200Context* ctx, const RB200Scene* scene, uint32_t n, const float* o, const float* d, const float* tmax,
               int any, RB200PrimaryHit* out) {
    if (n == 0) return RB200_OK;
    DeviceBuf dO, dD;
    const int rc = upload_rays(ctx, n, o, d, tmax, dO, dD);
    if (rc != RB200_OK) return rc;
    return run_query(ctx, scene, n, static_cast<float4*>(dO.p), static_cast<float4*>(dD.p), any, out);
}

// Measurement entry point (tools/trav_bench.py): the traversal kernel alone on a caller-supplied ray set, `reps` timed
// launches (CUDA events on the launching stream) after one warm-up; the hits of the last launch are reduced to a checksum
// so that kernel variants can be compared for equal results.
__global__ void k_hit_checksum(const RB200PrimaryHit* __restrict__ hits, uint32_t n, unsigned long long* out) {
    unsigned long long h = 0ull;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const RB200PrimaryHit r = hits[i];
        unsigned long long v = ((unsigned long long)__float_as_uint(r.t) << 32) ^ ((unsigned long long)r.primitive * 0x9E3779B97F4A7C15ull)
                               ^ ((unsigned long long)r.instance << 17) ^ __float_as_uint(r.u) ^ ((unsigned long long)__float_as_uint(r.v) << 13);
        v ^= (unsigned long long)i * 0xD6E8FEB86659FD93ull;
        v ^= v >> 29; v *= 0xBF58476D1CE4E5B9ull; v ^= v >> 32;
        h += v;
    }
    atomicAdd(out, h);
}

int bench_trace(RB200Context* ctx, const RB200Scene* scene, uint32_t n, const float* o, const float* d, const float* tmax, int any,
                uint32_t reps, float* outMs, uint64_t* outChecksum) {
    if (n == 0 || reps == 0) { set_error("bench_trace: n and reps must be > 0"); return RB200_ERR_INVALID_ARGUMENT; }
    int rc = sync_engines(ctx);
    if (rc != RB200_OK) return rc;
    DeviceBuf dO, dD, dOut, dSum;
    if ((rc = upload_rays(ctx, n, o, d, tmax, dO, dD)) != RB200_OK) return rc;
    RB_CUDA(dOut.alloc((size_t)n * sizeof(RB200PrimaryHit)));
    RB_CUDA(dSum.alloc(sizeof(unsigned long long)));
    RB_CUDA(cudaMemsetAsync(dSum.p, 0, sizeof(unsigned long long), ctx->stream));
    cudaEvent_t e0, e1;
    RB_CUDA(cudaEventCreate(&e0)); RB_CUDA(cudaEventCreate(&e1));
    rc = launch_query(ctx, scene, n, static_cast<float4*>(dO.p), static_cast<float4*>(dD.p), any, static_cast<RB200PrimaryHit*>(dOut.p));
    cudaEventRecord(e0, ctx->stream);
    for (uint32_t r = 0; r < reps && rc == RB200_OK; r++)
        rc = launch_query(ctx, scene, n, static_cast<float4*>(dO.p), static_cast<float4*>(dD.p), any, static_cast<RB200PrimaryHit*>(dOut.p));
    cudaEventRecord(e1, ctx->stream);
    if (rc == RB200_OK) {
        k_hit_checksum<<<256, 256, 0, ctx->stream>>>(static_cast<const RB200PrimaryHit*>(dOut.p), n, static_cast<unsigned long long*>(dSum.p));
        ctx->launches++;
    }
    unsigned long long sum = 0ull;
    cudaError_t ce = cudaMemcpyAsync(&sum, dSum.p, sizeof(sum), cudaMemcpyDeviceToHost, ctx->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
    float ms = 0.f;
    if (ce == cudaSuccess) ce = cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (rc != RB200_OK) return rc;
    if (ce != cudaSuccess) { set_error("bench_trace: %s", cudaGetErrorString(ce)); return RB200_ERR_CUDA; }
    if (outMs) *outMs = ms / (float)reps;
    if (outChecksum) *outChecksum = sum;
    return RB200_OK;
}

// ---------------------------------------------------------------------------------------------------
// rb200_shade_hits: ONE traceRayEXT + closest-hit shader invocation per caller-supplied ray — the payload the shader
// leaves behind (shaderCommon.h.glsl payload struct), for parity against the reference's compiled *.rchit.spv
// (tests/golden/spirv_hits.npz). The traversal kernel finds the hits, k_shade_hits runs the very eval_hit<MAT> the
// wave loop runs.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK) k_hits_for_shading(const WideNode* nodes, const TriRecord* tris, uint32_t n,
                                                            const float4* __restrict__ o, const float4* __restrict__ d,
                                                            uint4* __restrict__ hits, uint32_t* __restrict__ material, uint32_t* cursor) {
    uint32_t nv = 0, tt = 0;
    extern __shared__ __align__(16) unsigned char rb_dyn_smem[];
    WarpShared<false>* ws = reinterpret_cast<WarpShared<false>*>(rb_dyn_smem);
    trace_queue<false, false>(
        nodes, tris, n, cursor,
        [&](uint32_t i, rb_v3& ro, rb_v3& rd, float& tmax) {
            const float4 o4 = o[i], d4 = d[i];
            ro = rb_mk3(o4.x, o4.y, o4.z); rd = rb_mk3(d4.x, d4.y, d4.z); tmax = 10000.0f;
        },
        [&](uint32_t i, const RayHit& h) {
            if (h.tri == 0xFFFFFFFFu) { material[i] = 4u; hits[i] = make_uint4(0u, 0u, 0u, 0u); return; }
            const float4* tp = reinterpret_cast<const float4*>(tris + h.tri);
            const uint32_t iw = __float_as_uint(__ldg(tp + 1).w);
#if RB_SHADE_RECORDS
            const uint32_t prim = h.tri;
#else
            const uint32_t prim = __float_as_uint(__ldg(tp).w);
#endif
            hits[i] = make_uint4(__float_as_uint(h.b1), __float_as_uint(h.b2), prim, iw & TRI_INST_MASK);
            material[i] = iw >> 30;
        },
        nv, tt, ws[threadIdx.x >> 5]);
}

__global__ void __launch_bounds__(RB_SHADE_BLOCK) k_shade_hits(WaveParams P, uint32_t n, const float4* __restrict__ o4, const float4* __restrict__ d4,
                                                              const uint4* __restrict__ hits, const uint32_t* __restrict__ material,
                                                              const uint32_t* __restrict__ rngIn, const uint32_t* __restrict__ insideIn,
                                                              const float* __restrict__ accIn, RB200ShadeResult* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RB200ShadeResult r;
    memset(&r, 0, sizeof(r));
    const uint32_t m = material[i];
    r.material = m;
    r.rngState = rngIn[i];
    r.accumulatedDistance = accIn[i];
    if (m > 3u) { out[i] = r; return; }          // miss: the payload is the miss shader's business
    const rb_v3 ro = rb_mk3(o4[i].x, o4[i].y, o4[i].z), rd = rb_mk3(d4[i].x, d4[i].y, d4[i].z);
    const bool prevInside = insideIn[i] != 0u;
    uint32_t rng = rngIn[i];
    float acc = accIn[i];
    ShadeOut so;
    Surf s;
    const RB200InstanceProperties* props;
    bool didRefract;
    const uint4 h = hits[i];
    if (m == 0u) eval_hit<0>(P, h, ro, rd, prevInside, rng, acc, so, s, props, didRefract);
    else if (m == 1u) eval_hit<1>(P, h, ro, rd, prevInside, rng, acc, so, s, props, didRefract);
    else if (m == 2u) eval_hit<2>(P, h, ro, rd, prevInside, rng, acc, so, s, props, didRefract);
    else eval_hit<3>(P, h, ro, rd, prevInside, rng, acc, so, s, props, didRefract);
    r.color[0] = so.color.x; r.color[1] = so.color.y; r.color[2] = so.color.z;
    r.albedo[0] = so.albedo.x; r.albedo[1] = so.albedo.y; r.albedo[2] = so.albedo.z;
    r.origin[0] = so.newO.x; r.origin[1] = so.newO.y; r.origin[2] = so.newO.z;
    r.direction[0] = so.newD.x; r.direction[1] = so.newD.y; r.direction[2] = so.newD.z;
    r.emission[0] = so.emission.x; r.emission[1] = so.emission.y; r.emission[2] = so.emission.z;
    r.normal[0] = so.normal.x; r.normal[1] = so.normal.y; r.normal[2] = so.normal.z;
    r.pdf = so.pdf;
    r.accumulatedDistance = acc;
    r.rngState = rng;
    r.flags = 1u | (so.skip ? 2u : 0u) | (so.inside ? 4u : 0u);
    out[i] = r;
}

int shade_hits(RB200Context* ctx, const RB200Scene* scene, uint32_t n, const float* o, const float* d, const uint32_t* rng,
               const uint32_t* inside, const float* acc, RB200ShadeResult* out) {
    if (n == 0) return RB200_OK;
    if (scene->dev.tlasNodes) { set_error("rb200_shade_hits is a parity aid of the flattened path (context without RB200_FLAG_TWO_LEVEL)"); return RB200_ERR_INVALID_ARGUMENT; }
    int rc = sync_engines(ctx);
    if (rc != RB200_OK) return rc;
    std::vector<float> tmax(n, 10000.0f);
    DeviceBuf dO, dD, dHits, dMat, dRng, dIn, dAcc, dOut;
    if ((rc = upload_rays(ctx, n, o, d, tmax.data(), dO, dD)) != RB200_OK) return rc;
    RB_CUDA(dHits.alloc((size_t)n * sizeof(uint4))); RB_CUDA(dMat.alloc((size_t)n * 4)); RB_CUDA(dRng.alloc((size_t)n * 4));
    RB_CUDA(dIn.alloc((size_t)n * 4)); RB_CUDA(dAcc.alloc((size_t)n * 4)); RB_CUDA(dOut.alloc((size_t)n * sizeof(RB200ShadeResult)));
    cudaStream_t s = ctx->stream;
    RB_CUDA(cudaMemcpyAsync(dRng.p, rng, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    RB_CUDA(cudaMemcpyAsync(dIn.p, inside, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    RB_CUDA(cudaMemcpyAsync(dAcc.p, acc, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    RB_CUDA(cudaMemsetAsync(ctx->queryCursor, 0, sizeof(uint32_t), s));
    constexpr size_t smClosest = sizeof(WarpShared<false>) * (BLOCK / 32);
    RB_CUDA(cudaFuncSetAttribute(k_hits_for_shading, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smClosest));
    const int grid = (int)std::min<uint64_t>((n + BLOCK - 1) / BLOCK, (uint64_t)ctx->gQuery[0]);
    k_hits_for_shading<<<grid, BLOCK, smClosest, s>>>(scene->dev.nodes, scene->dev.tris, n, static_cast<float4*>(dO.p), static_cast<float4*>(dD.p),
                                                      static_cast<uint4*>(dHits.p), static_cast<uint32_t*>(dMat.p), ctx->queryCursor);
    WaveParams P = ctx->wp;
    P.S = scene->dev;
    k_shade_hits<<<(n + RB_SHADE_BLOCK - 1) / RB_SHADE_BLOCK, RB_SHADE_BLOCK, 0, s>>>(
        P, n, static_cast<float4*>(dO.p), static_cast<float4*>(dD.p), static_cast<uint4*>(dHits.p), static_cast<uint32_t*>(dMat.p),
        static_cast<uint32_t*>(dRng.p), static_cast<uint32_t*>(dIn.p), static_cast<float*>(dAcc.p), static_cast<RB200ShadeResult*>(dOut.p));
    ctx->launches += 2;
    RB_CUDA(cudaGetLastError());
    RB_CUDA(cudaMemcpyAsync(out, dOut.p, (size_t)n * sizeof(RB200ShadeResult), cudaMemcpyDeviceToHost, s));
    RB_CUDA(cudaStreamSynchronize(s));
    return RB200_OK;
}

// Per context, i.e. per DEVICE (function attributes and occupancy are per device): shared-memory limits of the traversal
// kernels, persistent-grid sizes, and the kernels' code loaded before the first frame (a first launch loads its code only
// once the device is idle, which with several batches in flight stalled the first presented frame by 170 ms).
template <class K> static int persistent_grid(K kernel, int numSMs, int block = BLOCK, size_t dynSmem = 0) {
    int perSM = 0;
    if (dynSmem) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dynSmem);
    // RB200_SMEM_CARVEOUT=<percent>: preferred shared-memory carve-out of the traversal kernels (developer aid; the
    // driver's default picks the smallest configuration that holds the resident blocks)
    static const char* carve = getenv("RB200_SMEM_CARVEOUT");
    if (dynSmem && carve) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(carve));
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, block, dynSmem) != cudaSuccess || perSM < 1) perSM = 1;
    // RB200_TRAV_BLOCKS_PER_SM / RB200_SHADE_BLOCKS_PER_SM=<n>: cap the resident blocks of the traversal (dynamic shared
    // memory) / the other persistent kernels below what fits, leaving room for the other engine's kernels to co-reside
    // (developer aid for the co-scheduling experiment of profiles/r02_summary.md)
    static const char* capTrav = getenv("RB200_TRAV_BLOCKS_PER_SM");
    static const char* capShade = getenv("RB200_SHADE_BLOCKS_PER_SM");
    const char* cap = dynSmem ? capTrav : capShade;
    if (cap && atoi(cap) >= 1) perSM = std::min(perSM, atoi(cap));
    return numSMs * perSM;
}

int configure_wave_kernels(RB200Context* ctx) {
    constexpr size_t smAny = sizeof(WarpShared<true>) * (BLOCK / 32), smClosest = sizeof(WarpShared<false>) * (BLOCK / 32);
    ctx->gExtend = persistent_grid(k_extend<false>, ctx->numSMs, RB_EXTEND_BLOCK, SMEM_EXTEND);
    ctx->gExtendC = persistent_grid(k_extend<true>, ctx->numSMs, RB_EXTEND_BLOCK, SMEM_EXTEND);
    ctx->gShadow = persistent_grid(k_shadow<false>, ctx->numSMs, RB_SHADOW_BLOCK, SMEM_SHADOW);
    ctx->gShadowC = persistent_grid(k_shadow<true>, ctx->numSMs, RB_SHADOW_BLOCK, SMEM_SHADOW);
    if (ctx->flags & RB200_FLAG_SKIP_NULL_SHADOW_RAYS) {       // same shape, register cap and shared memory: the grids above fit
        ctx->gShadow = std::min(ctx->gShadow, persistent_grid(k_shadow<false, true>, ctx->numSMs, RB_SHADOW_BLOCK, SMEM_SHADOW));
        ctx->gShadowC = std::min(ctx->gShadowC, persistent_grid(k_shadow<true, true>, ctx->numSMs, RB_SHADOW_BLOCK, SMEM_SHADOW));
    }
    ctx->gShade[0] = persistent_grid(k_shade<0>, ctx->numSMs, RB_SHADE_BLOCK);
    ctx->gShade[1] = persistent_grid(k_shade<1>, ctx->numSMs, RB_SHADE_BLOCK);
    ctx->gShade[2] = persistent_grid(k_shade<2>, ctx->numSMs, RB_SHADE_BLOCK);
    ctx->gShade[3] = persistent_grid(k_shade<3>, ctx->numSMs, RB_SHADE_BLOCK);
    ctx->gShade[4] = persistent_grid(k_shade<4>, ctx->numSMs, RB_SHADE_BLOCK);
    ctx->gFinish = persistent_grid(k_finish, ctx->numSMs);
    ctx->gQuery[0] = persistent_grid(k_trace_query<false>, ctx->numSMs, BLOCK, smClosest);
    ctx->gQuery[1] = persistent_grid(k_trace_query<true>, ctx->numSMs, BLOCK, smAny);
    ctx->gTwoLevel[0] = persistent_grid(k_extend_two_level, ctx->numSMs, BLOCK);
    ctx->gTwoLevel[1] = persistent_grid(k_shadow_two_level, ctx->numSMs, BLOCK);
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, k_generate);
    cudaFuncGetAttributes(&a, k_accumulate);
    cudaFuncGetAttributes(&a, k_resolve_sum);
    cudaFuncGetAttributes(&a, k_primary_rays);
    RB_CUDA(cudaGetLastError());
    return RB200_OK;
}

} // namespace rb200
