// context.cuh — host-side objects behind the opaque RB200Context / RB200Scene handles.
#pragma once
#include "common.cuh"

struct RB200Scene {
    RB200Context* ctx = nullptr;
    rb200::DeviceScene dev{};
    rb200::Bvh bvh{};
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
enum { ST_EXTEND = 0, ST_SHADOW = 1, ST_PATHS = 2, ST_NODES = 3, ST_TRIS = 4, ST_COUNT = 8 };

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
static constexpr int STATE_STRIDE = RB_PAIR_STATE == 2 ? 8 : (RB_PAIR_STATE ? 2 : 1);
static constexpr int SUM_STRIDE = RB_PAIR_STATE == 2 ? 8 : 1;

struct WaveParams {
    DeviceScene S;
    RB200RtPushConsts pc;
    uint32_t W, H, N, flags;
    uint32_t tileRank, tileCount, tileSize, tilesX;   // interleaved-tile partition (rb200_context_set_tiles); tileCount 1 = whole image
    StateArr<float4, STATE_STRIDE> rayO;      // xyz origin
    StateArr<float4, STATE_STRIDE> rayD;      // xyz direction (not necessarily unit)
    StateArr<uint4, STATE_STRIDE> hit;        // x = bits(b1), y = bits(b2), z = primitive, w = instance
    StateArr<float4, STATE_STRIDE> thr;       // xyz throughput, w = accumulatedDistance
    StateArr<float4, STATE_STRIDE> rad;       // xyz radiance of the current path
    StateArr<float4, SUM_STRIDE> sum;       // xyz summed sample colours of this batch, w = bits(actualSamples)
    StateArr<uint4, STATE_STRIDE> st;         // x = rng state, y = flags | segments << 8, z = sample index
    StateArr<float4> shO, shD, shA, shB, shT;   // shadow-ray records (compacted): origin/tmax, dir, D/wNEE, E*wBRDF/slot, throughput
    uint32_t* rayQ[2];
    uint32_t* matQ[5]; // 0..3 materials, 4 = miss
    uint32_t* endQ;
    uint32_t* counters;            // [2][CNT_SET]
    unsigned long long* stats;     // [ST_COUNT]
    StateArr<float4> mean;         // per pixel: mean of this batch's valid samples (xyz), w = 1 if any sample was valid
    float4* image;                 // HDR accumulation image (shared by both lanes)
};

} // namespace rb200

// Number of batches in flight. A batch's waves thin out (at 1080p, 8 spp x 16 bounces: wave 40 of 128 carries 8 % of
// the paths, wave 64 0.2 %) but a thin wave still costs 50-170 us of dependent-fetch latency per kernel; with more
// lanes more of those tails run beside another batch's full waves. 0.5 GB of path state per lane at 1080p.
// Headline scene on B200, device-resident loop: 1 / 2 / 3 / 4 / 6 lanes = 1478 / 1798 / 1891 / 1955 / 1985 Mrays/s.
#ifndef RB_LANES
#define RB_LANES 4
#endif

#ifndef RB_STAGGER_WAVE
#define RB_STAGGER_WAVE 0      // 0: 5/32 of the batch's waves (wave 20 of 128); see RB200Context::staggerWave
#endif

struct RB200Context {
    uint32_t width = 0, height = 0, flags = 0;
    int device = 0;
    int numSMs = 148;
    cudaStream_t stream = nullptr;             // front-end stream: API calls are ordered on it (may be the caller's)
    bool ownStream = false;
    // RB_LANES lanes (path-state sets + internal streams). Consecutive rb200_render_batch calls rotate through the
    // lanes so the long, thinly populated tail of one batch overlaps the heads of the next ones; the per-pixel
    // accumulation into the shared HDR image stays in batch order (k_accumulate of batch b waits for that of b-1).
    rb200::WaveParams lanes[RB_LANES]{};       // lanes[0] is also the scratch of the query entry points
    rb200::WaveParams& wp = lanes[0];
    cudaStream_t laneStream[RB_LANES] = {};
    cudaEvent_t accumDone[RB_LANES] = {};
    // The wave loop of a lane (maxWaves x (counter reset + extend + 5 shade + shadow + finish)) as a CUDA graph: its
    // kernel arguments depend on the lane, the scene and the push constants but not on sampleBatch (only k_generate
    // and k_accumulate read it), so one capture serves every batch until the camera / scene / sample counts change.
    cudaGraphExec_t waveGraph[RB_LANES] = {};
    rb200::WaveParams waveGraphKey[RB_LANES]{};
    uint32_t waveGraphWaves[RB_LANES] = {};
    uint64_t graphCaptures = 0;
    // Lane stagger: a batch may start once the previous batch (on the previous lane) has finished wave staggerWave.
    // Without it batches submitted faster than they render (graph launches cost the host ~0.3 ms) run in lockstep:
    // all lanes are in their full waves together and in their thin tails together, which is what the lanes are there
    // to avoid. Headline scene, 4 lanes, graph launches, end-to-end ms per step with the gate at wave 0 (off) / 6 / 12 /
    // 20 / 32 of 128: 54.9 / 50.9 / 50.6 / 50.1 / 53.0. staggerWave: 0 = 5/32 of the batch's waves, -1 = off, else the wave.
    cudaEvent_t staggerEv[RB_LANES] = {};
    int staggerWave = 0;
    cudaEvent_t frontMark = nullptr;
    std::vector<cudaEvent_t> ldrPendingEvents; // completion events of the outstanding rb200_read_ldr_async copies, oldest first
    std::vector<cudaEvent_t> ldrEventPool;
    uint64_t batchCalls = 0;
    std::vector<void*> allocations;
    float4 *ping = nullptr, *pong = nullptr;   // bloom work images
    float4* resolved = nullptr;                // rb200_present_sum: mean image of a SUM image (allocated on first use)
    uchar4* ldr = nullptr;
    RB200Stats last{}, cumulative{};
    unsigned long long* statsSnap = nullptr;   // cumulative device counters (each lane's per-batch counters are added
                                               // by k_accumulate when its batch ends)
    uint64_t launches = 0;                     // kernels launched by this context (all entry points)
    // RB200_FLAG_TIME_KERNELS: event pairs recorded around the kernels of the last batch, tagged by class
    std::vector<cudaEvent_t> evPool;
    std::vector<int> evClass;                  // class of pair i (events 2i, 2i+1): 0 generate, 1 extend, 2..6 shade, 7 shadow, 8 finish
    size_t evUsed = 0;
    uint32_t* waveCountsDev = nullptr;         // timing pass: per wave (rays extended, shadow rays), copied from the lane counters
    uint32_t waveCountsWaves = 0;              // waves of the last timed batch
    uint32_t waveCountsCap = 0;
};

namespace rb200 {
// wavefront.cu
int render_batch(RB200Context* ctx, const RB200Scene* scene, const RB200RtPushConsts* pc);
int trace_primary(RB200Context* ctx, const RB200Scene* scene, const RB200RtPushConsts* pc, RB200PrimaryHit* out);
int trace_rays(RB200Context* ctx, const RB200Scene* scene, uint32_t n, const float* o, const float* d, const float* tmax,
               int any, RB200PrimaryHit* out);
int resolve_sum(RB200Context* ctx, uint32_t numBatches);
// post.cu
int postprocess(RB200Context* ctx, const RB200BloomPushConsts* bloom, const RB200TonemappingPushConsts* tm,
                const float4* source = nullptr);      // source: HDR image to post-process (default: the context's)
int build_shade_records(const DeviceScene& S, float4* base, float4* frame, cudaStream_t stream);
void preload_post_kernels();
void preload_wave_kernels();
int present_sum(RB200Context* ctx, const float4* deviceSum, uint32_t numBatches, const RB200BloomPushConsts* bloom,
                const RB200TonemappingPushConsts* tm);
} // namespace rb200
