/*
 * reina_b200.h — C ABI of the B200-native path-tracing core for Reina.
 *
 * This is the drop-in boundary for ONE hot path of AlexanderJCS/reina-vk:
 *     scene hand-off -> trace + shade + accumulate one sample batch -> bloom + tonemap -> read back.
 * Everything the reference binds to its ray-tracing descriptor set is passed here as plain host
 * pointers and counts; the library copies them to the GPU, builds its own 8-wide BVH (the reference lets the
 * Vulkan driver build BLAS/TLAS) and runs hand-written sm_100a kernels.
 *
 * Reference interfaces replaced (paths relative to the reference repo):
 *   RB200InstanceProperties      <- polyglot/raytrace.h:18-41     (byte-for-byte, 120 B)
 *   RB200RtPushConsts            <- polyglot/raytrace.h:43-54     (byte-for-byte, 160 B)
 *   RB200BloomPushConsts         <- polyglot/bloom.h:4-8
 *   RB200TonemappingPushConsts   <- polyglot/tonemapping.h:4-6
 *   RB200InstanceData            <- src/scene/Instances.h:15-27 == shaders/raytrace/nee.h.glsl:11-22 (112 B)
 *   RB200Instance                <- the TLAS record, src/tools/vktools.cpp:464-488 + src/graphics/Blas.cpp:34-39
 *   RB200SceneDesc               <- descriptor bindings 2..11,13 written in src/Reina.cpp:394-407
 *   rb200_scene_create           <- Scene::build, src/scene/Scene.cpp:56-125 (BLAS/TLAS build replaced by LBVH->BVH8)
 *   rb200_render_batch           <- Reina::traceRays, src/Reina.cpp:425-470 (vkCmdTraceRaysKHR(W,H,1))
 *   rb200_postprocess            <- Reina::applyBloom + applyTonemapping, src/Reina.cpp:472-577
 *   rb200_read_ldr               <- save(), src/Reina.cpp:19-47 (RGBA8, row 0 = top, stride 4*W)
 *
 * Error convention: the reference throws std::runtime_error and main() prints what() (src/main.cpp:6-14).
 * Nothing is thrown across this boundary: every call returns RB200_OK (0) or a negative status and
 * rb200_last_error() returns the message of the last failure on the calling thread.
 *
 * Threading: a context is not re-entrant (the reference has one queue and one command buffer in flight,
 * src/Reina.cpp:327-328). Work is enqueued on the context's CUDA stream; rb200_read_* synchronise.
 */
#ifndef REINA_B200_H
#define REINA_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#  define RB200_API __attribute__((visibility("default")))
#else
#  define RB200_API
#endif

/* ------------------------------------------------------------------------------------------------ */
/* status codes                                                                                      */
/* ------------------------------------------------------------------------------------------------ */
enum {
    RB200_OK = 0,
    RB200_ERR_INVALID_ARGUMENT = -1,
    RB200_ERR_CUDA = -2,
    RB200_ERR_NO_EMITTER = -3,   /* "Scene must have at least one emissive object" (src/scene/Instances.cpp:125-127); only with NEE */
    RB200_ERR_OUT_OF_MEMORY = -4,
    RB200_ERR_NO_DEVICE = -5
};

/* context flags */
enum {
    RB200_FLAG_NEE = 1u << 0,        /* enable next-event estimation + MIS (compiled out in the shipped reference:
                                        shaders/raytrace/raytrace.rgen.glsl:145-146). Off = as-shipped behaviour. */
    RB200_FLAG_ACCUM_SUM = 1u << 1,  /* keep the SUM of per-batch pixel means instead of the running mean of
                                        raytrace.rgen.glsl:277-284; used when the sample batches are split over
                                        several GPUs and reduced once (see rb200_resolve_sum). */
    RB200_FLAG_COUNT_BVH = 1u << 2,  /* counting build: also count wide-node visits and triangle tests per ray */
    RB200_FLAG_GROUP_TILES = 1u << 4, /* rb200_group_create only: latency mode — every device traces its interleaved 32 x 32
                                       * tiles of every batch instead of every n-th batch of all pixels */
    RB200_FLAG_TIME_KERNELS = 1u << 3, /* bracket every kernel of rb200_render_batch with CUDA events (per-class device
                                        times for the roofline report; serialises nothing but adds event overhead) */
    RB200_FLAG_SKIP_NULL_SHADOW_RAYS = 1u << 6, /* next-event estimation: a shadow ray whose `direct` term is exactly +0 in all
                                        three channels BEFORE the visibility test (the light faces away from the hit point, or lies
                                        below its horizon: cosThetai or the geometry term is max(., 0) = 0) is answered "not
                                        occluded" without being traversed — raygen adds (occluded ? 0 : direct) * wNEE + ..., and
                                        direct is +0 either way (shaders/raytrace/raytrace.rgen.glsl:43-95, 174-177). The wavefront
                                        evaluates `direct` before the shadow stage, so it can tell; the reference traces first.
                                        Images are bit-identical with and without the flag. Such rays stay in RB200Stats::shadowRays
                                        (they are rays of the estimator) and are also counted in shadowRaysSkipped. Off by default,
                                        so that the traced ray set is the reference's. Flattened scenes only. */
    RB200_FLAG_TWO_LEVEL = 1u << 5   /* scenes of this context keep the reference's two-level structure (src/scene/Scene.cpp:
                                        93-111: ONE hierarchy per distinct object, a top-level hierarchy over the instances;
                                        rays are moved into object space with the inverse instance transform, as the Vulkan
                                        driver does) instead of flattening every instance into world-space triangles. Triangle
                                        and shading records are then stored once per object, not once per instance. Slower per
                                        ray (a ray set-up per visited instance); results differ from the flattened path in the
                                        last bits of t (object-space arithmetic). Singular instance transforms are refused. */
};

/* ------------------------------------------------------------------------------------------------ */
/* PODs shared with the reference (GLSL "scalar" layout == C packed-by-4)                            */
/* ------------------------------------------------------------------------------------------------ */
typedef struct RB200InstanceProperties {   /* polyglot/raytrace.h:18-41 */
    uint32_t indicesOffset;          /*   0 */
    float    albedo[3];              /*   4 */
    float    emission[3];            /*  16 */
    uint32_t tbnsIndicesOffset;      /*  28 */
    uint32_t texIndicesOffset;       /*  32  0xFFFFFFFF = model has no UVs */
    float    roughness;              /*  36 */
    float    ior;                    /*  40 */
    uint32_t interpNormals;          /*  44  C++ bool + 3 pad == GLSL 4-byte bool; only != 0 is tested */
    float    absorption;             /*  48 */
    int32_t  textureID;              /*  52 */
    int32_t  normalMapTexID;         /*  56 */
    int32_t  bumpMapTexID;           /*  60 */
    uint32_t cullBackface;           /*  64 */
    float    anisotropic;            /*  68 */
    float    subsurface;             /*  72 */
    float    clearcoatGloss;         /*  76 */
    float    sheenTint[3];           /*  80 */
    float    specularTint[3];        /*  92 */
    float    metallic;               /* 104 */
    float    clearcoat;              /* 108 */
    float    specularTransmission;   /* 112 */
    float    sheen;                  /* 116 */
} RB200InstanceProperties;           /* 120 bytes */

typedef struct RB200RtPushConsts {   /* polyglot/raytrace.h:43-54 */
    float    invView[16];            /*   0  column-major (glm) */
    float    invProjection[16];      /*  64  column-major (glm) */
    uint32_t sampleBatch;            /* 128  caller owns the sequencing (src/Reina.cpp:431-432) */
    float    totalEmissiveWeight;    /* 132 */
    float    focusDist;              /* 136 */
    float    defocusMultiplier;      /* 140  already divided by 100 (src/Reina.cpp:150) */
    float    directClamp;            /* 144 */
    float    indirectClamp;          /* 148  pushed but never read by any shader; kept for layout */
    uint32_t samplesPerPixel;        /* 152 */
    uint32_t maxBounces;             /* 156 */
} RB200RtPushConsts;                 /* 160 bytes */

typedef struct RB200BloomPushConsts {      /* polyglot/bloom.h:4-8 */
    float radius;      /* percent of the image dimension AND the Gaussian sigma (blurCommon.h.glsl:23-31) */
    float threshold;
    float intensity;
} RB200BloomPushConsts;

typedef struct RB200TonemappingPushConsts { /* polyglot/tonemapping.h:4-6 */
    float exposure;
} RB200TonemappingPushConsts;

typedef struct RB200InstanceData {   /* src/scene/Instances.h:15-27, alignas(16) */
    float    transform[16];          /*   0  column-major object->world */
    uint32_t materialOffset;         /*  64 */
    uint32_t cdfRangeStart;          /*  68 */
    uint32_t cdfRangeEnd;            /*  72  inclusive */
    uint32_t indexOffset;            /*  76 */
    float    emission[3];            /*  80 */
    float    weight;                 /*  92 */
    float    area;                   /*  96 */
    uint32_t cullBackface;           /* 100  bool + pad */
    float    padding[2];             /* 104 */
} RB200InstanceData;                 /* 112 bytes */

/* One TLAS instance record: what vktools::createTlas hands the driver per instance. */
typedef struct RB200Instance {
    float    transform[16];          /* column-major glm::mat4 object->world (Instance::getTransform) */
    uint32_t instancePropertiesID;   /* gl_InstanceCustomIndexEXT */
    uint32_t materialIdx;            /* SBT hit-group offset: 0 lambertian, 1 metal, 2 dielectric, 3 disney */
    uint32_t indexOffset;            /* ModelRange.indexOffset: first entry of this model in `indices` */
    uint32_t triangleCount;          /* ModelRange.indexCount (already /3, src/scene/Models.cpp:52) */
} RB200Instance;                     /* 80 bytes */

typedef struct RB200Texture {        /* src/graphics/Image.cpp:43-86: RGBA8 UNORM, no sRGB decode, one mip */
    const uint8_t* rgba8;            /* row 0 first, 4*width bytes per row, already flipped as the caller wants */
    uint32_t width;
    uint32_t height;
} RB200Texture;

/* The tables the reference binds in Reina::writeDescriptorSets (src/Reina.cpp:394-407). Host pointers.
 * The library copies everything during rb200_scene_create; the caller may free them afterwards. */
typedef struct RB200SceneDesc {
    const float*    vertices;        uint32_t numVertices;       /* binding 2: float4 per vertex (w = 1) */
    const uint32_t* indices;         uint32_t numIndices;        /* binding 3: global vertex ids */
    const RB200InstanceProperties* instanceProperties; uint32_t numInstanceProperties; /* binding 4 */
    const float*    tbns;            uint32_t numTbns;           /* binding 5: 9 floats, column-major [T,B,N] */
    const uint32_t* tbnIndices;      uint32_t numTbnIndices;     /* binding 6 */
    const RB200InstanceData* emissiveMetadata; uint32_t numEmissive;  /* binding 7 (may be 0 without NEE) */
    const float*    cdfTriangles;    uint32_t numCdfTriangles;   /* binding 8 */
    const float*    cdfInstances;    uint32_t numCdfInstances;   /* binding 9 payload (the {count,total} header
                                                                    is numCdfInstances / RtPushConsts.totalEmissiveWeight) */
    const float*    texCoords;       uint32_t numTexCoords;      /* binding 10: float2 per entry */
    const uint32_t* texIndices;      uint32_t numTexIndices;     /* binding 11 */
    const RB200Texture* textures;    uint32_t numTextures;       /* binding 13 */
    const RB200Instance* instances;  uint32_t numInstances;      /* replaces binding 1 (the TLAS) */
} RB200SceneDesc;

typedef struct RB200Context RB200Context;
typedef struct RB200Scene   RB200Scene;

/* Counters of the last rb200_render_batch (and cumulative since context creation). */
typedef struct RB200Stats {
    uint64_t extendRays;     /* closest-hit rays traced */
    uint64_t shadowRays;     /* any-hit rays traced */
    uint64_t paths;          /* camera paths started */
    uint64_t nodeVisits;     /* wide-node visits   (only with RB200_FLAG_COUNT_BVH) */
    uint64_t triTests;       /* triangle tests     (only with RB200_FLAG_COUNT_BVH) */
    uint64_t waves;          /* wavefront iterations executed */
    uint64_t kernelLaunches; /* kernels launched by the library */
    uint64_t shadowNodeVisits; /* the part of nodeVisits / triTests spent on any-hit (shadow) rays */
    uint64_t shadowTriTests;
    uint64_t shadowRaysSkipped; /* RB200_FLAG_SKIP_NULL_SHADOW_RAYS: shadow rays (included in shadowRays) answered without a traversal.
                                 * Cumulative only, and counted when the shadow stage runs — with batches in flight it can be ahead
                                 * of shadowRays, which is counted when a path ends; differences over a steady-state window agree */
} RB200Stats;

/* Device time per kernel class of the launches issued by the last rb200_render_batch (needs RB200_FLAG_TIME_KERNELS).
 * With several lanes a launch carries the rays of every batch in flight, so in steady state the launches of one call
 * trace one batch's worth of rays in total. */
typedef struct RB200KernelTimes {
    float    generateMs, extendMs, shadeMs[5] /* lambertian, metal, dielectric, disney, miss */, shadowMs, finishMs;
    uint32_t extendLaunches, shadeLaunches, shadowLaunches, finishLaunches;
    /* the same split by occupancy: launches whose ray queue held at least half the image's pixels ("full waves", where
     * the kernel is throughput-bound) — the rest of a batch's launches are thin waves bound by single-ray latency */
    float    extendFullMs, shadowFullMs;
    uint32_t extendFullLaunches, shadowFullLaunches;
    uint64_t extendFullRays, shadowFullRays;
    /* what the launches of the call processed: rays traced by k_extend / k_shadow, slots shaded per material kernel
     * (lambertian, metal, dielectric, disney, miss) and paths ended by k_finish */
    uint64_t extendRays, shadowRays, shadeItems[5], finishItems;
} RB200KernelTimes;

typedef struct RB200BvhInfo {
    uint32_t numTriangles;
    uint32_t numWideNodes;
    uint32_t maxDepth;
    uint32_t reserved;       /* KiB of L2 set aside as persisting for the node+triangle arrays (0 = hint not applied) */
    uint64_t nodeBytes;
    uint64_t triangleBytes;
    uint64_t hash;           /* FNV-1a over the node and triangle arrays: equal across runs/GPUs (deterministic build) */
    float    buildMs;        /* device time of the build */
    float    sceneMin[3];
    float    sceneMax[3];
} RB200BvhInfo;

/* One primary-ray hit, for parity checks against the oracle (north_star: prim ids equal, t within 1e-5). */
typedef struct RB200PrimaryHit {
    float    t;              /* < 0 : miss */
    float    u, v;           /* barycentric weights of vertex 1 and 2 */
    uint32_t instance;       /* index into RB200SceneDesc.instances */
    uint32_t primitive;      /* triangle index within the model (gl_PrimitiveID) */
} RB200PrimaryHit;

/* What one closest-hit shader invocation leaves in the ray payload (shaders/raytrace/shaderCommon.h.glsl:17-35), for
 * rb200_shade_hits. */
typedef struct RB200ShadeResult {
    float    color[3], albedo[3], origin[3], direction[3], emission[3], normal[3];   /* pld.color .. pld.surfaceNormal */
    float    pdf;                    /* pld.pdf */
    float    accumulatedDistance;    /* pld.accumulatedDistance after the shader */
    uint32_t rngState;               /* pld.rngState after the shader */
    uint32_t flags;                  /* bit 0: the ray hit something; bit 1: pld.skip; bit 2: pld.insideDielectric */
    uint32_t material;               /* hit group that ran: 0 lambertian, 1 metal, 2 dielectric, 3 disney; 4 = miss */
    uint32_t reserved;
} RB200ShadeResult;

/* ------------------------------------------------------------------------------------------------ */
/* entry points                                                                                      */
/* ------------------------------------------------------------------------------------------------ */

/* Library/ABI version (major<<16 | minor). */
RB200_API uint32_t rb200_version(void);

/* Message of the last error on this thread ("" if none). Never NULL. */
RB200_API const char* rb200_last_error(void);

/* Create a render context of fixed size on CUDA device `device` (src/Reina.cpp:54-55,76-80: the reference
 * fixes W x H at construction). `flags` = RB200_FLAG_*. */
RB200_API int rb200_context_create(uint32_t width, uint32_t height, int device, uint32_t flags, RB200Context** out);
RB200_API int rb200_context_destroy(RB200Context* ctx);

/* Use an externally owned cudaStream_t (e.g. the harness's stream) instead of the context's own. */
RB200_API int rb200_context_set_stream(RB200Context* ctx, void* cuda_stream);

/* Interleaved-tile partition for several GPUs in latency mode: this context traces only the pixels of
 * the tileSize x tileSize tiles whose row-major index is congruent to tileRank modulo tileCount, every batch of them,
 * and leaves all other pixels of its image untouched (zero after a rb200_write_hdr of zeros / in a fresh context).
 * Because every pixel is still accumulated by one context in batch order, the sum (ncclReduce) of the ranks' images is
 * bit-identical to the single-GPU image; bloom needs the whole image and runs on the root after the reduce.
 * tileCount = 1 (the default) renders the whole image. Takes effect with the next rb200_render_batch. */
RB200_API int rb200_context_set_tiles(RB200Context* ctx, uint32_t tileRank, uint32_t tileCount, uint32_t tileSize);

/* Copy the scene tables to the device and build the acceleration structure (LBVH -> 8-wide compressed BVH).
 * The scene is immutable afterwards (the reference never updates an acceleration structure).
 * With RB200_FLAG_NEE: RB200_ERR_NO_EMITTER without an emissive instance (src/scene/Instances.cpp:125-127), and
 * RB200_ERR_INVALID_ARGUMENT if light sampling would read past the index buffer — nee.h.glsl:97-105 addresses an
 * emitter's triangles as indices[3 * cdfSlot + indexOffset] with cdfSlot counted over the CONCATENATED triangle CDF,
 * which is only the emitter's own triangle for an emitter whose CDF starts at slot 0 (reproduced as is otherwise). */
RB200_API int rb200_scene_create(RB200Context* ctx, const RB200SceneDesc* desc, RB200Scene** out);
RB200_API int rb200_scene_destroy(RB200Scene* scene);
RB200_API int rb200_scene_bvh_info(const RB200Scene* scene, RB200BvhInfo* out);

/* Trace + shade + accumulate one sample batch (pc->samplesPerPixel samples per pixel, <= pc->maxBounces
 * segments each). pc->sampleBatch == 0 overwrites the HDR image, > 0 folds into the running average
 * (raytrace.rgen.glsl:277-284). Asynchronous on the context stream.
 * Replaces Reina::traceRays (src/Reina.cpp:425-470: push RtPushConsts, sampleBatch++, vkCmdTraceRaysKHR(W,H,1)).
 * When the stream-ordered work of this call has run, the image holds exactly the batches asked for so far, in call
 * order. Inside the library the NEXT batches of a regular sequence (same push constants, sampleBatch advancing by a
 * constant stride — the reference's frame loop) are traced speculatively alongside this one so that every kernel
 * launch stays full; they are folded into the image only by the calls that ask for them and are discarded if the
 * sequence changes (camera move, other scene, other sample counts). RB200_NO_SPECULATION=1 turns this off.
 * samplesPerPixel * maxBounces must not exceed 2^22. */
RB200_API int rb200_render_batch(RB200Context* ctx, const RB200Scene* scene, const RB200RtPushConsts* pc);

/* Back-pressure for a frame loop (the reference waits for its command buffer every frame, src/Reina.cpp:327-328): block
 * until at most `max_pending` of the rb200_render_batch calls issued so far are still unfinished. */
RB200_API int rb200_wait_batches_pending(RB200Context* ctx, uint32_t max_pending);

/* With RB200_FLAG_ACCUM_SUM: turn the accumulated sum of `numBatches` batch means into the mean image. */
RB200_API int rb200_resolve_sum(RB200Context* ctx, uint32_t numBatches);

/* Bloom (blurX with threshold -> blurY -> combine) and ACES tonemap into the RGBA8 image. */
RB200_API int rb200_postprocess(RB200Context* ctx, const RB200BloomPushConsts* bloom,
                                const RB200TonemappingPushConsts* tonemap);

/* Present a frame from a SUM image without touching the context's accumulation image: resolve `device_sum_rgba32f`
 * (W*H*4 floats in DEVICE memory: the sum of `numBatches` batch means, e.g. the NCCL reduce of every rank's
 * rb200_hdr_device_ptr image into a separate buffer; NULL = this context's own image) into a staging image with the
 * arithmetic of rb200_resolve_sum, then bloom + tonemap into the RGBA8 frame. Asynchronous on the context stream, so a
 * multi-GPU frame loop needs no host synchronisation between the reduce and the read-back. */
RB200_API int rb200_present_sum(RB200Context* ctx, const void* device_sum_rgba32f, uint32_t numBatches,
                                const RB200BloomPushConsts* bloom, const RB200TonemappingPushConsts* tonemap);

/* Read-back (synchronises the stream). `rgba8`: W*H*4 bytes, row 0 = top. `rgba32f`: W*H*4 floats. */
RB200_API int rb200_read_ldr(RB200Context* ctx, uint8_t* rgba8);
RB200_API int rb200_read_hdr(RB200Context* ctx, float* rgba32f);
/* Pipelined read-back: enqueue the copy of the RGBA8 frame behind the post-processing queued so far and return at
 * once; rb200_wait_ldr blocks until the most recent such copy has landed. With `rgba8` from rb200_host_alloc (pinned)
 * the caller can issue rb200_render_batch(i+1) before waiting for frame i, so post-processing and read-back of one
 * frame overlap the next batch (what a swap chain gives the reference's loop between endSubmit and present,
 * src/Reina.cpp:380-385). Several copies may be outstanding, each into its own host frame: rb200_wait_ldr_pending
 * blocks until at most `max_pending` of them (the most recent ones) are still in flight, rb200_wait_ldr until none is.
 * rb200_pipeline_depth: how many frames a display loop should keep outstanding so that the host never drains the
 * device. */
RB200_API int rb200_read_ldr_async(RB200Context* ctx, uint8_t* rgba8);
RB200_API int rb200_wait_ldr(RB200Context* ctx);
RB200_API int rb200_wait_ldr_pending(RB200Context* ctx, uint32_t max_pending);
RB200_API uint32_t rb200_pipeline_depth(void);
/* Page-locked host memory for the asynchronous read-back (callers without a CUDA runtime of their own). */
RB200_API int rb200_host_alloc(size_t bytes, void** out);
RB200_API int rb200_host_free(void* p);
/* Replace the HDR accumulation image (resume / post-processing parity on identical input). */
RB200_API int rb200_write_hdr(RB200Context* ctx, const float* rgba32f);
/* Device pointer of the float4 HDR accumulation image (for a collective over NVLink issued by the harness). */
RB200_API int rb200_hdr_device_ptr(RB200Context* ctx, void** out_device_ptr);

/* Primary-ray closest hits for batch `pc->sampleBatch`, first sample of each pixel (W*H records). */
RB200_API int rb200_trace_primary(RB200Context* ctx, const RB200Scene* scene, const RB200RtPushConsts* pc,
                                  RB200PrimaryHit* out_hits);

/* Closest-hit / any-hit queries on caller-supplied rays (n rays; origins/directions as xyz triples;
 * tmax per ray). `any_hit` != 0 -> only hit.t >= 0 / < 0 is meaningful. Used by traversal parity tests
 * and the traversal micro-benchmark. Host pointers. */
RB200_API int rb200_trace_rays(RB200Context* ctx, const RB200Scene* scene, uint32_t n, const float* origins,
                               const float* directions, const float* tmax, int any_hit, RB200PrimaryHit* out_hits);

/* Parity aid: for each of n caller-supplied rays (xyz triples, not necessarily unit length) ONE closest-hit traversal and
 * ONE invocation of the hit instance's material shader — the device code the wave loop runs — with the payload's incoming
 * state set from rngStates / insideDielectric (0 or 1) / accumulatedDistance; *out = the payload afterwards. This is
 * traceRayEXT + <material>.rchit of the reference for one ray (raytrace.rgen.glsl:110-122). Host pointers, blocking. */
RB200_API int rb200_shade_hits(RB200Context* ctx, const RB200Scene* scene, uint32_t n, const float* origins,
                               const float* directions, const uint32_t* rngStates, const uint32_t* insideDielectric,
                               const float* accumulatedDistance, RB200ShadeResult* out);

/* Measurement aid: the traversal kernel alone on the caller's rays (as rb200_trace_rays), one warm-up launch and `reps`
 * timed ones (CUDA events on the launching stream); *out_ms_per_launch = mean device time of a launch, *out_checksum =
 * a hash of all hit records of the last launch (equal results <=> equal checksums, whatever the kernel variant). */
RB200_API int rb200_bench_trace(RB200Context* ctx, const RB200Scene* scene, uint32_t n, const float* origins,
                                const float* directions, const float* tmax, int any_hit, uint32_t reps,
                                float* out_ms_per_launch, uint64_t* out_checksum);

/* How the context schedules batches: `engines` wave loops on as many streams, each tracing `lanes` consecutive batches
 * of the caller's sequence together (RB200_ENGINES / RB200_LANES at context creation); *discarded_batches = speculative
 * batches thrown away because the caller's sequence changed. Any pointer may be NULL. */
RB200_API int rb200_engine_config(RB200Context* ctx, uint32_t* engines, uint32_t* lanes, uint64_t* discarded_batches);

/* last_batch = the batch folded into the image by the most recent rb200_render_batch (its rays are counted per path,
 * whichever launches traced them); cumulative = all batches folded so far. */
RB200_API int rb200_get_stats(RB200Context* ctx, RB200Stats* last_batch, RB200Stats* cumulative);
RB200_API int rb200_get_kernel_times(RB200Context* ctx, RB200KernelTimes* out);
RB200_API int rb200_synchronize(RB200Context* ctx);

/* Measurement aid (the traversal roofline is the L2 gather bandwidth, which MEASURED_PEAKS.json does not
 * hold): GB/s of independent random reads of whole `recordBytes`-sized records (80 = wide node, 48 = triangle, 16)
 * from a table of `tableBytes` (make it the size of the BVH so that it is L2-resident the way the BVH is). */
RB200_API int rb200_measure_gather(RB200Context* ctx, size_t tableBytes, uint32_t recordBytes, float* out_gbps);

/* ------------------------------------------------------------------------------------------------ */
/* several GPUs (SURVEY.md 8b "Context", 8e): BVH replicated per device, one ncclReduce of the float4   */
/* accumulation images per presented frame. NCCL is loaded at run time (libnccl.so.2).                 */
/* ------------------------------------------------------------------------------------------------ */
typedef struct RB200Group      RB200Group;
typedef struct RB200GroupScene RB200GroupScene;

/* One process, n devices (what `Reina` with n GPUs would create once, src/Reina.cpp:54-80): a context per device, each
 * driven by its own host thread, ncclCommInitAll over `devices`. Sample split by default (device i renders batches
 * i, i + n, ... into a local SUM image; RB200_FLAG_ACCUM_SUM is implied); RB200_FLAG_GROUP_TILES selects latency mode. */
RB200_API int rb200_group_create(uint32_t width, uint32_t height, const int* devices, int numDevices, uint32_t flags,
                                 RB200Group** out);
RB200_API int rb200_group_destroy(RB200Group* group);
RB200_API int rb200_group_size(const RB200Group* group);
/* 1 when the group runs latency mode (RB200_FLAG_GROUP_TILES) with peer stores: every device can address device 0's memory,
 * so the kernel that folds a batch into the image (k_accumulate) stores each pixel its device owns into device 0's image
 * over NVLink as well, and a presented frame needs no collective — device 0 waits for the others' "tiles done" events,
 * snapshots its image and post-processes it. 0: one ncclReduce per presented frame (sample split; latency mode without
 * peer access, or with RB200_GROUP_TILES_REDUCE=1 in the environment). Frames are bit-identical either way. */
RB200_API int rb200_group_uses_peer_stores(const RB200Group* group);
RB200_API int rb200_group_context(RB200Group* group, int index, RB200Context** out);       /* member context (not owned by the caller) */
RB200_API int rb200_group_set_tile_size(RB200Group* group, uint32_t tileSize);             /* latency mode, before the first batch */
/* Upload the tables and build the hierarchy on every device concurrently; fails if the replicas' hashes differ (the
 * build is deterministic). */
RB200_API int rb200_group_scene_create(RB200Group* group, const RB200SceneDesc* desc, RB200GroupScene** out);
RB200_API int rb200_group_scene_destroy(RB200GroupScene* scene);
RB200_API int rb200_group_scene_bvh_info(const RB200GroupScene* scene, int index, RB200BvhInfo* out);
/* Sample split: device i renders batches firstBatch + i + k * n, k < batchesPerDevice, with *pc (sampleBatch replaced).
 * Latency mode: every device renders its tiles of batches firstBatch .. firstBatch + batchesPerDevice - 1. Asynchronous. */
RB200_API int rb200_group_render_batches(RB200Group* group, const RB200GroupScene* scene, const RB200RtPushConsts* pc,
                                         uint32_t firstBatch, uint32_t batchesPerDevice);
/* One frame: snapshot of every device's image -> ONE ncclReduce(float, 4 * W * H) to device 0 -> resolve (sum / batches
 * rendered so far; nothing to resolve in latency mode) + bloom + tonemap on device 0. Asynchronous on the devices' streams;
 * rendering continues underneath. Replaces applyBloom + applyTonemapping of src/Reina.cpp:472-577 for n devices. */
RB200_API int rb200_group_present(RB200Group* group, const RB200BloomPushConsts* bloom, const RB200TonemappingPushConsts* tonemap);
RB200_API int rb200_group_read_ldr(RB200Group* group, uint8_t* rgba8);       /* the frame of the last rb200_group_present (blocking) */
RB200_API int rb200_group_read_hdr(RB200Group* group, float* rgba32f);       /* reduce + resolve, blocking: the mean image so far */
RB200_API int rb200_group_synchronize(RB200Group* group);
RB200_API int rb200_group_get_stats(RB200Group* group, RB200Stats* cumulative);   /* summed over the devices */

/* One process per GPU (torchrun, MPI): rank 0 calls rb200_comm_unique_id and hands the 128 bytes to the other ranks by
 * its own means; every rank then calls rb200_context_comm_init (ncclCommInitRank on the context's device) and, per
 * frame, rb200_context_reduce_present: snapshot of the rank's image, ONE ncclReduce to rank 0, and on rank 0 resolve
 * (with RB200_FLAG_ACCUM_SUM: sum / totalBatches) + bloom + tonemap of the reduced copy. bloom / tonemap are only read
 * on rank 0. rb200_context_reduced_device_ptr: rank 0's reduced image (device memory, W*H*4 floats). */
RB200_API int rb200_comm_unique_id(void* out_128_bytes);
RB200_API int rb200_context_comm_init(RB200Context* ctx, const void* unique_id_128_bytes, int rank, int nranks);
RB200_API int rb200_context_comm_destroy(RB200Context* ctx);
RB200_API int rb200_context_reduce_present(RB200Context* ctx, uint32_t totalBatches, const RB200BloomPushConsts* bloom,
                                           const RB200TonemappingPushConsts* tonemap);
RB200_API int rb200_context_reduced_device_ptr(RB200Context* ctx, void** out_device_ptr);

#ifdef __cplusplus
}
#endif
#endif /* REINA_B200_H */
