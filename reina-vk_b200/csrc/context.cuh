// context.cuh — host-side objects behind the opaque RB200Context / RB200Scene handles.
#pragma once
#include "common.cuh"
#include <deque>

struct RB200Scene {
    RB200Context* ctx = nullptr;
    rb200::DeviceScene dev{};
    rb200::Bvh bvh{};
    rb200::Bvh tlas{};                        // RB200_FLAG_TWO_LEVEL: the hierarchy over the instances (bvh = the objects')
    uint32_t numModels = 0;                   // distinct objects of a two-level scene
    std::vector<void*> allocations;           // device buffers owned by the scene
    std::vector<cudaArray_t> texArrays;
    std::vector<cudaTextureObject_t> texObjects;
    std::vector<RB200Instance> hostInstances;
    uint32_t numEmissive = 0;
    size_t l2PersistBytes = 0;                // L2 set aside as persisting for the BVH (0 = hint not applied)
    uint64_t hash = 0;
    bool hashValid = false;
};

namespace rb200 {

// counters of one wave parity set
enum { CNT_RAYS = 0, CNT_MAT0 = 1, CNT_MISS = 5, CNT_SHADOW = 6, CNT_END = 7, CNT_CURSOR_EXTEND = 8, CNT_CURSOR_SHADOW = 9, CNT_SET = 12 };
// device statistics (unsigned long long each)
enum { ST_EXTEND = 0, ST_SHADOW = 1, ST_PATHS = 2, ST_NODES = 3, ST_TRIS = 4, ST_NODES_SHADOW = 5, ST_TRIS_SHADOW = 6, ST_SHADOW_SKIPPED = 7 /* cumulative array only */, ST_COUNT = 8 };

// path flags kept in PathState.st.y (low byte); the traced-segment counter lives in bits 8..31
enum { F_INSIDE = 1u, F_FIRST = 2u, F_PREVSKIP = 4u };

// Path-state arrays go through this proxy so that their cache policy is one switch. RB_STREAM_STATE=1 sends them
// through ld/st.global.cs (evict-first) to keep the per-wave sweep (~0.6 GB at 1080p) from pushing the BVH out of L2.
// Measured on B200 (dragon, 1080p): traversal +0.8 %, but the shading / finish kernels lose more (lambertian 6.2 ->
// 7.3 ms, finish 3.6 -> 4.5 ms per batch) because in thin waves the state written by one kernel is still in L2 when
// the next kernel reads it; net -4.5 %, so the default is plain loads and stores.
#ifndef RB_STREAM_STATE
#define RB_STREAM_STATE 0
#endif
template <class T> struct StateRef {
    T* p;
    __device__ __forceinline__ operator T() const {
#if RB_STREAM_STATE
        return __ldcs(p);
#else
        return *p;
#endif
    }
    __device__ __forceinline__ void operator=(const T& v) const {
#if RB_STREAM_STATE
        __stcs(p, v);
#else
        *p = v;
#endif
    }
};
template <class T, int STRIDE = 1> struct StateArr {
    T* p;
    __device__ __forceinline__ StateRef<T> operator[](uint32_t i) const { return StateRef<T>{p + (size_t)i * STRIDE}; }
};
// Path-state layout (RB_PAIR_STATE). A material queue holds a scattered subset of the slots and the shading kernels
// are bound by DRAM sector throughput (~2.5 TB/s), so what a kernel touches for one slot should sit together:
//   0  one array per record: every 16-byte access moved a half-used 32-byte sector;
//   1  interleaved pairs, one sector per slot: (rayO, rayD), (thr, st), (hit, rad): +4 % rays/s over 0;
//   2  one 128-byte line per slot: rayO, rayD | thr, st | hit, rad | sum, spare — a shading kernel's three sectors
//      are one DRAM burst: Disney 8.0 -> 7.5 ms, Lambertian 5.6 -> 5.3 ms per batch, shadow 18.0 -> 18.2 ms,
//      +2.2 % rays/s over 1 on B200 (headline scene). Default.
#ifndef RB_PAIR_STATE
#define RB_PAIR_STATE 2
#endif
// RB_WIDE_STATE=1 (needs the one-line-per-slot layout): the two records of a 32-byte sector — (rayO, rayD), (thr, st),
// (hit, rad) — move with ONE 256-bit access (LDG / STG.E.256, sm_100) instead of two 128-bit ones. The shading kernels
// run into the rate at which an SM can send scattered requests to L2, not into latency or DRAM bandwidth (the software
// prefetch experiment in wavefront.cu made them slower by adding requests), so fewer, wider requests for the same
// sectors are what helps: a shaded hit reads its state with 3 requests instead of 6 and writes it back with 3 or 4
// instead of 5.
#ifndef RB_WIDE_STATE
#define RB_WIDE_STATE (RB_PAIR_STATE == 2)
#endif
#if RB_WIDE_STATE && RB_PAIR_STATE != 2
#error "RB_WIDE_STATE needs RB_PAIR_STATE == 2"
#endif
static constexpr int STATE_STRIDE = RB_PAIR_STATE == 2 ? 8 : (RB_PAIR_STATE ? 2 : 1);
static constexpr int SUM_STRIDE = RB_PAIR_STATE == 2 ? 8 : 1;

// first[slot] and second[slot], which with RB_WIDE_STATE are the two halves of one 32-byte sector
template <class TA, class TB, int S>
__device__ __forceinline__ void load_pair(const StateArr<TA, S>& first, const StateArr<TB, S>& second, uint32_t slot, TA& a, TB& b) {
    static_assert(sizeof(TA) == 16 && sizeof(TB) == 16, "16-byte records");
#if RB_WIDE_STATE
    uint32_t r[8];
    asm volatile("ld.global.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(first.p + (size_t)slot * S) : "memory");
    const uint4 ua = make_uint4(r[0], r[1], r[2], r[3]), ub = make_uint4(r[4], r[5], r[6], r[7]);
    a = *reinterpret_cast<const TA*>(&ua); b = *reinterpret_cast<const TB*>(&ub);
#else
    a = first[slot]; b = second[slot];
#endif
}
template <class TA, class TB, int S>
__device__ __forceinline__ void store_pair(const StateArr<TA, S>& first, const StateArr<TB, S>& second, uint32_t slot, const TA& a, const TB& b) {
    static_assert(sizeof(TA) == 16 && sizeof(TB) == 16, "16-byte records");
#if RB_WIDE_STATE
    const uint4 ua = *reinterpret_cast<const uint4*>(&a), ub = *reinterpret_cast<const uint4*>(&b);
    asm volatile("st.global.v8.u32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "l"(first.p + (size_t)slot * S), "r"(ua.x), "r"(ua.y), "r"(ua.z), "r"(ua.w), "r"(ub.x), "r"(ub.y), "r"(ub.z), "r"(ub.w) : "memory");
#else
    first[slot] = a; second[slot] = b;
#endif
}

struct WaveParams {
    DeviceScene S;
    RB200RtPushConsts pc;
    uint32_t W, H, N, flags;                  // N = W * H = pixels = slots of ONE lane
    uint32_t tileRank, tileCount, tileSize, tilesX;   // interleaved-tile partition (rb200_context_set_tiles); tileCount 1 = whole image
    uint32_t numLanes, NT;                    // lanes of this engine; NT = numLanes * N = slots of the engine (slot = lane * N + pixel)
    StateArr<float4, STATE_STRIDE> rayO;      // xyz origin
    StateArr<float4, STATE_STRIDE> rayD;      // xyz direction (not necessarily unit)
    StateArr<uint4, STATE_STRIDE> hit;        // x = bits(b1), y = bits(b2), z = primitive, w = instance
    StateArr<float4, STATE_STRIDE> thr;       // xyz throughput, w = accumulatedDistance
    StateArr<float4, STATE_STRIDE> rad;       // xyz radiance of the current path
    StateArr<float4, SUM_STRIDE> sum;       // xyz summed sample colours of this batch, w = bits(actualSamples)
    StateArr<uint4, STATE_STRIDE> st;         // x = rng state, y = flags | segments << 8, z = sample index, w = shadow rays of this path
    StateArr<float4> shO, shD, shA, shB, shT;   // shadow-ray records (compacted): origin/tmax, dir, D/wNEE, E*wBRDF/slot, throughput
    uint32_t* rayQ[2];
    uint32_t* matQ[5]; // 0..3 materials, 4 = miss
    uint32_t* endQ;
    uint32_t* counters;            // [2][CNT_SET]
    unsigned long long* stats;     // [RB_MAX_LANES][ST_COUNT]: counters of the batch each lane is rendering
    StateArr<float4> mean;         // per slot: mean of this batch's valid samples (xyz), w = 1 if any sample was valid
    float4* image;                 // HDR accumulation image (shared by all engines and lanes)
    unsigned long long* nullShadow; // RB200_FLAG_SKIP_NULL_SHADOW_RAYS: cumulative count of shadow rays answered without a traversal
};

} // namespace rb200

// ---------------------------------------------------------------------------------------------------
// Engines and lanes: how several batches share the kernels of a wave.
//
// A batch's waves thin out — at 1080p, 8 spp x 16 bounces, wave 40 of 128 carries 8 % of the paths, wave 64 0.2 % — but
// a thin wave still costs the dependent-fetch latency of its longest ray in every traversal launch, and the samples of a
// pixel are sequential by construction (one RNG stream per pixel and batch), so a batch cannot be made shorter.
// Round 1 overlapped the thin tails by running RB_LANES independent wave loops on as many streams; every loop still
// launched its own thin waves (102 of a batch's 128 extend launches took 38 % of the extend time for 15 % of the rays).
//
// Now an ENGINE owns L lanes (path-state sets) that SHARE one set of queues and one wave loop on one stream: the
// kernels of a wave process the rays of every lane's batch together (slot = lane * N + pixel; a slot is still private
// to its pixel and batch, so the image does not depend on how batches are mixed). The L batches in flight are L
// consecutive batches of the caller's sequence, staggered by ceil(maxWaves / L) waves: while batch b runs waves
// 96..127, b+1 runs 64..95, b+2 32..63 and b+3 0..31, so every launch of the engine carries about one batch's worth of
// rays in total and no launch is thin. Batches b+1.. are started SPECULATIVELY: rb200_render_batch(b) predicts that the
// next calls will differ from this one only by sampleBatch += stride (progressive rendering, the reference's frame loop
// src/Reina.cpp:425-470; stride = number of ranks under a sample split). A batch is only ever folded into the image by
// the call that asks for it, in call order, after all of its waves — the contract of rb200_render_batch is unchanged.
// A call that does not match the prediction (camera moved, other scene, other sample counts) discards the speculative
// lanes and starts over; speculation begins with the second call of a regular sequence, so isolated calls pay nothing.
// A context owns E engines on E streams (calls rotate through them) so that kernels of different classes — traversal:
// issue / latency-bound, shading: DRAM-bound — still overlap. E = 4, L = 1 is the round-1 organisation.
// Headline scene on B200 (Mrays/s of the device loop / end to end / fraction of the HBM roofline reached by k_extend over
// ALL launches of a step), E x L: 4 x 1 2277 / 2253 / 0.385; 1 x 4 2286 / 2272 / 0.507; 1 x 8 2438 / 2420 / 0.555;
// 2 x 4 2519 / 2496 / 0.515; 2 x 6 2571 / 2552 / 0.545; 2 x 8 2643 / 2605 / 0.563 (default); 3 x 4 2583 / 2540; 3 x 6
// 2690 / 2462; 4 x 4 2593 / 2452 (profiles/r02_summary.md). Path state: 0.53 GB per lane at 1080p.
// ---------------------------------------------------------------------------------------------------
#ifndef RB_MAX_ENGINES
#define RB_MAX_ENGINES 4
#endif
#ifndef RB_MAX_LANES
#define RB_MAX_LANES 16
#endif
#ifndef RB_ENGINES
#define RB_ENGINES 2           // default engines per context (RB200_ENGINES overrides at context creation)
#endif
#ifndef RB_LANES
#define RB_LANES 8             // default lanes per engine (RB200_LANES overrides)
#endif
#ifndef RB_GRAPH_CHUNK
#define RB_GRAPH_CHUNK 32      // waves per captured graph (even: a chunk preserves the queue parity)
#endif

#ifndef RB_STAGGER_WAVE
#define RB_STAGGER_WAVE 0      // 0: 5/32 of the batch's waves (wave 20 of 128); see RB200Context::staggerWave
#endif

namespace rb200 {

struct LaneState {
    bool active = false;
    uint32_t sampleBatch = 0;      // the batch this lane renders
    uint32_t wavesDone = 0;
};

struct WaveGraph {
    cudaGraphExec_t exec = nullptr;
    uint32_t waves = 0, parity = 0, staggerAt = 0;
};

struct Engine {
    WaveParams P{};                         // arrays sized numLanes * N
    cudaStream_t stream = nullptr;
    cudaEvent_t accumDone = nullptr, staggerEv = nullptr;
    int numLanes = 1;
    LaneState lane[RB_MAX_LANES];
    int head = 0;                            // lane of the oldest batch in flight
    int active = 0;                          // lanes in flight: head, head+1, ... (mod numLanes), in batch order
    uint32_t globalWave = 0;                 // waves issued since the last reset: its parity selects the queue set
    // prediction state
    bool keyValid = false;
    WaveParams key{};                        // P with pc.sampleBatch = 0 at the last call
    const RB200Scene* scene = nullptr;
    bool havePrev = false;
    uint32_t prevBatch = 0;
    uint32_t stride = 1;                     // predicted sampleBatch increment between this engine's calls
    uint32_t streak = 0;                     // consecutive calls that continued the sequence
    std::vector<WaveGraph> graphs;           // captured chunks of the wave loop for the current key
    uint64_t calls = 0;
    uint64_t wastedBatches = 0;              // speculative batches discarded by a mismatch
};

} // namespace rb200

struct RB200Context {
    uint32_t width = 0, height = 0, flags = 0;
    int device = 0;
    int numSMs = 148;
    cudaStream_t stream = nullptr;             // front-end stream: API calls are ordered on it (may be the caller's)
    bool ownStream = false;
    int numEngines = RB_ENGINES, numLanes = RB_LANES;
    rb200::Engine eng[RB_MAX_ENGINES];
    rb200::WaveParams& wp = eng[0].P;          // eng[0]'s arrays are also the scratch of the query entry points
    // persistent-grid sizes of the wave kernels on THIS device (function attributes are per device)
    int gExtend = 0, gExtendC = 0, gShadow = 0, gShadowC = 0, gShade[5] = {0, 0, 0, 0, 0}, gFinish = 0, gQuery[2] = {0, 0}, gTwoLevel[2] = {0, 0};
    uint64_t graphCaptures = 0;
    // Engine stagger (only with one lane per engine, i.e. without speculation): a batch may start once the previous
    // batch (on the previous engine) has finished wave staggerWave. Without it batches submitted faster than they render
    // run in lockstep: all engines are in their full waves together and in their thin tails together.
    // staggerWave: 0 = 5/32 of the batch's waves, -1 = off, else the wave.
    int staggerWave = 0;
    cudaEvent_t frontMark = nullptr;
    float4* peerImage = nullptr;               // latency mode of a one-process group: device 0's image, written by k_accumulate over NVLink (group.cu)
    std::vector<cudaEvent_t> ldrPendingEvents; // completion events of the outstanding rb200_read_ldr_async copies, oldest first
    std::vector<cudaEvent_t> ldrEventPool;
    std::deque<cudaEvent_t> batchEvents;       // one per rb200_render_batch still in flight (rb200_wait_batches_pending)
    std::vector<cudaEvent_t> batchEventPool;
    uint64_t batchCalls = 0;
    int lastEngine = -1;                       // engine of the most recent rb200_render_batch call
    std::vector<void*> allocations;
    float4 *ping = nullptr, *pong = nullptr;   // bloom work images
    float4* resolved = nullptr;                // rb200_present_sum: mean image of a SUM image (allocated on first use)
    uchar4* ldr = nullptr;
    bool hdrMayBeNonFinite = false;            // the HDR image may hold NaN / infinities (rb200_write_hdr of such an image, or a
                                               // non-finite directClamp): post-processing then runs the shader's loops as written
    uint32_t* queryCursor = nullptr;           // cursor of the query kernels (rb200_trace_*)
    RB200Stats last{}, cumulative{};
    unsigned long long* statsSnap = nullptr;   // cumulative device counters (a lane's per-batch counters are added by
                                               // k_accumulate when its batch is folded into the image)
    unsigned long long* statsLast = nullptr;   // counters of the batch folded last
    uint64_t launches = 0;                     // kernels launched by this context (all entry points)
    // RB200_FLAG_TIME_KERNELS: event pairs recorded around the kernels of the last call, tagged by class
    std::vector<cudaEvent_t> evPool;
    std::vector<int> evClass;                  // class of pair i (events 2i, 2i+1): 0 generate, 1 extend, 2..6 shade, 7 shadow, 8 finish
    size_t evUsed = 0;
    uint32_t* waveCountsDev = nullptr;         // timing pass: per wave (rays extended, shadow rays), copied from the queue counters
    uint32_t waveCountsWaves = 0;              // waves of the last timed call
    uint32_t waveCountsCap = 0;
};

namespace rb200 {
// wavefront.cu
int render_batch(RB200Context* ctx, const RB200Scene* scene, const RB200RtPushConsts* pc);
int trace_primary(RB200Context* ctx, const RB200Scene* scene, const RB200RtPushConsts* pc, RB200PrimaryHit* out);
int trace_rays(RB200Context* ctx, const RB200Scene* scene, uint32_t n, const float* o, const float* d, const float* tmax,
               int any, RB200PrimaryHit* out);
int resolve_sum(RB200Context* ctx, uint32_t numBatches);
int bench_trace(RB200Context* ctx, const RB200Scene* scene, uint32_t n, const float* o, const float* d, const float* tmax, int any,
                uint32_t reps, float* outMs, uint64_t* outChecksum);
int shade_hits(RB200Context* ctx, const RB200Scene* scene, uint32_t n, const float* o, const float* d, const uint32_t* rng,
               const uint32_t* inside, const float* acc, RB200ShadeResult* out);
int context_create(uint32_t width, uint32_t height, int device, uint32_t flags, int lanesHint, RB200Context** out);   // api.cu
int configure_wave_kernels(RB200Context* ctx);     // per device: shared-memory limits, persistent grids, code preload
void invalidate_speculation(RB200Context* ctx);    // drains the engines and discards speculative batches
// post.cu
int postprocess(RB200Context* ctx, const RB200BloomPushConsts* bloom, const RB200TonemappingPushConsts* tm,
                const float4* source = nullptr);      // source: HDR image to post-process (default: the context's)
int build_shade_records(const DeviceScene& S, float4* base, float4* frame, cudaStream_t stream);
void preload_post_kernels();
int present_sum(RB200Context* ctx, const float4* deviceSum, uint32_t numBatches, const RB200BloomPushConsts* bloom,
                const RB200TonemappingPushConsts* tm);
} // namespace rb200
