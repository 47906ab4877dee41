// api.cu — the extern "C" boundary declared in include/reina_b200.h.
#include "context.cuh"
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <cmath>

namespace rb200 {

static thread_local std::string g_error;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
}

template <class T> static int upload(RB200Scene* sc, const T* host, size_t count, const T** dev, cudaStream_t s) {
    T* d = nullptr;
    const size_t bytes = (count ? count : 1) * sizeof(T);
    RB_CUDA(cudaMalloc((void**)&d, bytes));
    sc->allocations.push_back(d);
    if (count) RB_CUDA(cudaMemcpyAsync(d, host, count * sizeof(T), cudaMemcpyHostToDevice, s));
    else RB_CUDA(cudaMemsetAsync(d, 0, bytes, s));
    *dev = d;
    return RB200_OK;
}

template <class T> static int ctx_alloc(RB200Context* c, T** p, size_t count) {
    RB_CUDA(cudaMalloc((void**)p, (count ? count : 1) * sizeof(T)));
    c->allocations.push_back(*p);
    return RB200_OK;
}

static int validate_desc(const RB200SceneDesc* d) {
    if (!d->vertices || !d->indices || !d->instanceProperties || !d->instances || d->numInstances == 0) {
        set_error("scene description lacks vertices / indices / instanceProperties / instances");
        return RB200_ERR_INVALID_ARGUMENT;
    }
    if (!d->tbns || !d->tbnIndices) { set_error("scene description lacks the TBN tables (bindings 5, 6)"); return RB200_ERR_INVALID_ARGUMENT; }
    // every table whose count is non-zero must be there
    if ((d->numEmissive && !d->emissiveMetadata) || (d->numCdfTriangles && !d->cdfTriangles) || (d->numCdfInstances && !d->cdfInstances) ||
        (d->numTexCoords && !d->texCoords) || (d->numTexIndices && !d->texIndices) || (d->numTextures && !d->textures)) {
        set_error("scene description: a table with a non-zero count is NULL"); return RB200_ERR_INVALID_ARGUMENT;
    }
    bool usesTexCoords = false;
    for (uint32_t i = 0; i < d->numInstances; i++) {
        const RB200Instance& in = d->instances[i];
        if (in.instancePropertiesID >= d->numInstanceProperties) { set_error("instance %u: instancePropertiesID out of range", i); return RB200_ERR_INVALID_ARGUMENT; }
        if ((uint64_t)in.indexOffset + 3ull * in.triangleCount > d->numIndices) { set_error("instance %u: index range out of bounds", i); return RB200_ERR_INVALID_ARGUMENT; }
        const RB200InstanceProperties& p = d->instanceProperties[in.instancePropertiesID];
        if (p.textureID >= (int)d->numTextures || p.normalMapTexID >= (int)d->numTextures || p.bumpMapTexID >= (int)d->numTextures) {
            set_error("instance %u: texture id out of range", i); return RB200_ERR_INVALID_ARGUMENT;
        }
        // the shading kernels address the index buffer through the PROPERTIES' offset (closestHitCommon.h.glsl:61-66)
        if ((uint64_t)p.indicesOffset + 3ull * in.triangleCount > d->numIndices) { set_error("instance %u: InstanceProperties.indicesOffset range out of bounds", i); return RB200_ERR_INVALID_ARGUMENT; }
        if (p.texIndicesOffset != 0xFFFFFFFFu) usesTexCoords = true;
        if ((uint64_t)p.tbnsIndicesOffset + 3ull * in.triangleCount > d->numTbnIndices) { set_error("instance %u: TBN index range out of bounds", i); return RB200_ERR_INVALID_ARGUMENT; }
        if (p.texIndicesOffset != 0xFFFFFFFFu && (uint64_t)p.texIndicesOffset + 3ull * in.triangleCount > d->numTexIndices) {
            set_error("instance %u: texcoord index range out of bounds", i); return RB200_ERR_INVALID_ARGUMENT;
        }
    }
    for (uint32_t i = 0; i < d->numIndices; i++) if (d->indices[i] >= d->numVertices) { set_error("vertex index %u out of range", i); return RB200_ERR_INVALID_ARGUMENT; }
    for (uint32_t i = 0; i < d->numTbnIndices; i++) if (d->tbnIndices[i] >= d->numTbns) { set_error("TBN index %u out of range", i); return RB200_ERR_INVALID_ARGUMENT; }
    if (usesTexCoords && d->numTexCoords == 0) { set_error("an instance refers to texture coordinates but the texcoord table is empty"); return RB200_ERR_INVALID_ARGUMENT; }
    if (d->numTexCoords || usesTexCoords) for (uint32_t i = 0; i < d->numTexIndices; i++)
        if (d->texIndices[i] != 0xFFFFFFFFu && d->texIndices[i] >= d->numTexCoords) { set_error("texcoord index %u out of range", i); return RB200_ERR_INVALID_ARGUMENT; }
    for (uint32_t i = 0; i < d->numEmissive; i++) {
        const RB200InstanceData& e = d->emissiveMetadata[i];
        if (e.cdfRangeEnd >= d->numCdfTriangles || e.cdfRangeStart > e.cdfRangeEnd) { set_error("emissive %u: CDF range out of bounds", i); return RB200_ERR_INVALID_ARGUMENT; }
    }
    if (d->numEmissive != d->numCdfInstances) { set_error("numEmissive != numCdfInstances"); return RB200_ERR_INVALID_ARGUMENT; }
    return RB200_OK;
}

} // namespace rb200

using namespace rb200;

// lanesHint > 0: lanes per engine instead of the compiled default (the environment still overrides) — group.cu asks for more
// lanes in latency mode, where a device traces 1 / n of the pixels and its launches would otherwise be thin
int rb200::context_create(uint32_t width, uint32_t height, int device, uint32_t flags, int lanesHint, RB200Context** out) {
    if (!out || width == 0 || height == 0 || (uint64_t)width * height > 0x7FFFFFFFull) { set_error("invalid context size"); return RB200_ERR_INVALID_ARGUMENT; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device available (this library has no CPU path)"); return RB200_ERR_NO_DEVICE; }
    if (device < 0 || device >= ndev) { set_error("device %d out of range (%d devices)", device, ndev); return RB200_ERR_INVALID_ARGUMENT; }
    RB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RB_CUDA(cudaGetDeviceProperties(&prop, device));
    cudaStream_t front = nullptr;
    RB_CUDA(cudaStreamCreateWithFlags(&front, cudaStreamNonBlocking));
    RB200Context* c = new RB200Context();
    c->width = width; c->height = height; c->flags = flags; c->device = device;
    c->numSMs = prop.multiProcessorCount;
    c->stream = front;
    c->ownStream = true;
    const size_t N = (size_t)width * height;
    // engines x lanes (context.cuh): RB200_ENGINES / RB200_LANES override the compiled defaults; the counting pass renders
    // one batch at a time
    c->numEngines = RB_ENGINES; c->numLanes = lanesHint > 0 ? std::min(lanesHint, (int)RB_MAX_LANES) : RB_LANES;
    if (const char* e = getenv("RB200_ENGINES")) c->numEngines = atoi(e);
    if (const char* e = getenv("RB200_LANES")) c->numLanes = atoi(e);
    if (flags & RB200_FLAG_COUNT_BVH) { c->numEngines = 1; c->numLanes = 1; }
    if (c->numEngines < 1 || c->numEngines > RB_MAX_ENGINES || c->numLanes < 1 || c->numLanes > RB_MAX_LANES ||
        (uint64_t)N * (uint64_t)c->numLanes > 0x7FFFFFFFull) {
        set_error("engines x lanes = %d x %d out of range (at most %d x %d, lanes * pixels < 2^31)", c->numEngines, c->numLanes,
                  RB_MAX_ENGINES, RB_MAX_LANES);
        rb200_context_destroy(c);
        return RB200_ERR_INVALID_ARGUMENT;
    }
    int rc;
#define A(ptr, n) if ((rc = ctx_alloc(c, &(ptr), (n))) != RB200_OK) { rb200_context_destroy(c); return rc; }
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); rb200_context_destroy(c); return RB200_ERR_CUDA; } } while (0)
    float4* image = nullptr;
    A(image, N);
    for (int en = 0; en < c->numEngines; en++) {
        Engine& E = c->eng[en];
        E.numLanes = c->numLanes;
        WaveParams& L = E.P;
        const size_t NT = N * (size_t)c->numLanes;
        L.W = width; L.H = height; L.N = (uint32_t)N; L.flags = flags;
        L.numLanes = (uint32_t)c->numLanes; L.NT = (uint32_t)NT;
        L.tileRank = 0; L.tileCount = 1; L.tileSize = 32; L.tilesX = (width + 31u) / 32u;
        L.image = image;
#if RB_PAIR_STATE == 2
        // one 128-byte line per slot: float4 records 0..7 = rayO, rayD, thr, st, hit, rad, sum, (spare)
        A(L.rayO.p, 8 * NT); L.rayD.p = L.rayO.p + 1; L.thr.p = L.rayO.p + 2; L.st.p = reinterpret_cast<uint4*>(L.rayO.p + 3);
        L.hit.p = reinterpret_cast<uint4*>(L.rayO.p + 4); L.rad.p = L.rayO.p + 5; L.sum.p = L.rayO.p + 6;
#elif RB_PAIR_STATE
        // interleaved pairs (see context.cuh): record 2*slot is the first member, 2*slot+1 the second
        A(L.rayO.p, 2 * NT); L.rayD.p = L.rayO.p + 1;
        A(L.thr.p, 2 * NT);  L.st.p = reinterpret_cast<uint4*>(L.thr.p + 1);
        A(L.hit.p, 2 * NT);  L.rad.p = reinterpret_cast<float4*>(L.hit.p + 1);
        A(L.sum.p, NT);
#else
        A(L.rayO.p, NT); A(L.rayD.p, NT); A(L.hit.p, NT); A(L.thr.p, NT); A(L.rad.p, NT); A(L.sum.p, NT); A(L.st.p, NT);
#endif
        A(L.shO.p, NT); A(L.shD.p, NT); A(L.shA.p, NT); A(L.shB.p, NT); A(L.shT.p, NT);
        A(L.rayQ[0], NT); A(L.rayQ[1], NT);
        for (int m = 0; m < 5; m++) A(L.matQ[m], NT);
        A(L.endQ, NT); A(L.counters, 2 * CNT_SET); A(L.mean.p, NT); A(L.stats, (size_t)RB_MAX_LANES * ST_COUNT);
        CK(cudaMemsetAsync(L.stats, 0, (size_t)RB_MAX_LANES * ST_COUNT * sizeof(unsigned long long), c->stream));
        CK(cudaMemsetAsync(L.counters, 0, 2 * CNT_SET * sizeof(uint32_t), c->stream));
        CK(cudaStreamCreateWithFlags(&E.stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&E.accumDone, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&E.staggerEv, cudaEventDisableTiming));
    }
    c->staggerWave = RB_STAGGER_WAVE;
    if (const char* e = getenv("RB200_STAGGER_WAVE")) c->staggerWave = atoi(e);
    if ((rc = configure_wave_kernels(c)) != RB200_OK) { rb200_context_destroy(c); return rc; }
    preload_post_kernels();
    CK(cudaEventCreateWithFlags(&c->frontMark, cudaEventDisableTiming));
    A(c->statsSnap, ST_COUNT); A(c->statsLast, ST_COUNT); A(c->queryCursor, 1);
    A(c->ping, N); A(c->pong, N); A(c->ldr, N);
    // rb200_present_sum's staging image: allocated here in sum mode (cudaMalloc synchronises the device, which would
    // drain the engines if it happened on the first presented frame), on first use otherwise
    if (flags & RB200_FLAG_ACCUM_SUM) A(c->resolved, N);
    for (int en = 0; en < c->numEngines; en++) c->eng[en].P.nullShadow = c->statsSnap + ST_SHADOW_SKIPPED;
    WaveParams& P = c->wp;
#undef A
    CK(cudaMemsetAsync(c->statsSnap, 0, ST_COUNT * sizeof(unsigned long long), c->stream));
    CK(cudaMemsetAsync(c->statsLast, 0, ST_COUNT * sizeof(unsigned long long), c->stream));
    CK(cudaMemsetAsync(P.image, 0, N * sizeof(float4), c->stream));
    CK(cudaMemsetAsync(c->ldr, 0, N * sizeof(uchar4), c->stream));
    CK(cudaStreamSynchronize(c->stream));
#undef CK
    *out = c;
    return RB200_OK;
}

extern "C" {

RB200_API uint32_t rb200_version(void) { return (1u << 16) | 0u; }

RB200_API const char* rb200_last_error(void) { return g_error.c_str(); }

RB200_API int rb200_context_create(uint32_t width, uint32_t height, int device, uint32_t flags, RB200Context** out) {
    return rb200::context_create(width, height, device, flags, 0, out);
}


RB200_API int rb200_context_destroy(RB200Context* ctx) {
    if (!ctx) return RB200_OK;
    cudaSetDevice(ctx->device);
    for (int e = 0; e < RB_MAX_ENGINES; e++) if (ctx->eng[e].stream) cudaStreamSynchronize(ctx->eng[e].stream);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (void* p : ctx->allocations) cudaFree(p);
    for (cudaEvent_t e : ctx->evPool) cudaEventDestroy(e);
    for (int e = 0; e < RB_MAX_ENGINES; e++) {
        Engine& E = ctx->eng[e];
        for (WaveGraph& g : E.graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
        if (E.staggerEv) cudaEventDestroy(E.staggerEv);
        if (E.accumDone) cudaEventDestroy(E.accumDone);
        if (E.stream) cudaStreamDestroy(E.stream);
    }
    if (ctx->frontMark) cudaEventDestroy(ctx->frontMark);
    if (ctx->waveCountsDev) cudaFree(ctx->waveCountsDev);
    for (cudaEvent_t e : ctx->ldrPendingEvents) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ldrEventPool) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->batchEvents) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->batchEventPool) cudaEventDestroy(e);
    if (ctx->ownStream && ctx->stream) cudaStreamDestroy(ctx->stream);
    cudaGetLastError();
    delete ctx;
    return RB200_OK;
}

RB200_API int rb200_context_set_stream(RB200Context* ctx, void* cuda_stream) {
    if (!ctx) { set_error("null context"); return RB200_ERR_INVALID_ARGUMENT; }
    RB_CUDA(cudaSetDevice(ctx->device));
    for (int e = 0; e < ctx->numEngines; e++) RB_CUDA(cudaStreamSynchronize(ctx->eng[e].stream));
    RB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->ownStream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->ownStream = false;
    return RB200_OK;
}

RB200_API int rb200_scene_create(RB200Context* ctx, const RB200SceneDesc* d, RB200Scene** out) {
    if (!ctx || !d || !out) { set_error("null argument"); return RB200_ERR_INVALID_ARGUMENT; }
    int rc = validate_desc(d);
    if (rc != RB200_OK) return rc;
    if ((ctx->flags & RB200_FLAG_NEE) && d->numEmissive == 0) {
        set_error("Scene must have at least one emissive object");   // src/scene/Instances.cpp:125-127
        return RB200_ERR_NO_EMITTER;
    }
    if (ctx->flags & RB200_FLAG_NEE)
        for (uint32_t i = 0; i < d->numEmissive; i++) {
            // nee.h.glsl:97-105 addresses an emitter's triangles by their position in the CONCATENATED triangle CDF
            // (indices[3 * cdfIndex + indexOffset]). For an emitter whose CDF does not start at 0 that reaches past its own
            // triangles (reproduced as is), and past the end of the index buffer it is an out-of-bounds read: refused.
            const RB200InstanceData& e = d->emissiveMetadata[i];
            if (3ull * e.cdfRangeEnd + e.indexOffset + 2ull >= (unsigned long long)d->numIndices) {
                set_error("emissive %u: light sampling would read past the index buffer (triangle CDF slot %u + indexOffset %u)",
                          i, e.cdfRangeEnd, e.indexOffset);
                return RB200_ERR_INVALID_ARGUMENT;
            }
        }
    RB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    RB200Scene* sc = new RB200Scene();
    // from here on every failure path releases the half-built scene
#define SC_CUDA(call)                                                                                  \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            rb200::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));    \
            cudaGetLastError();                                                                        \
            rb200_scene_destroy(sc);                                                                   \
            return RB200_ERR_CUDA;                                                                     \
        }                                                                                              \
    } while (0)
    sc->ctx = ctx;
    DeviceScene& D = sc->dev;
#define U(field, host, count) if ((rc = upload(sc, host, (size_t)(count), &D.field, s)) != RB200_OK) { rb200_scene_destroy(sc); return rc; }
    U(vertices, reinterpret_cast<const float4*>(d->vertices), d->numVertices);
    U(indices, d->indices, d->numIndices);
    U(props, d->instanceProperties, d->numInstanceProperties);
    D.instProps = nullptr;
#if RB_INST_RECORDS
    {   // the properties of every instance, by instance index (validate_desc checked instancePropertiesID)
        std::vector<RB200InstanceProperties> perInstance(d->numInstances);
        for (uint32_t i = 0; i < d->numInstances; i++) perInstance[i] = d->instanceProperties[d->instances[i].instancePropertiesID];
        U(instProps, perInstance.data(), perInstance.size());
        SC_CUDA(cudaStreamSynchronize(s));      // perInstance goes out of scope
    }
#endif
    U(tbns, d->tbns, 9 * (size_t)d->numTbns);
    U(tbnIndices, d->tbnIndices, d->numTbnIndices);
    U(emissive, d->emissiveMetadata, d->numEmissive);
    U(cdfTriangles, d->cdfTriangles, d->numCdfTriangles);
    U(cdfInstances, d->cdfInstances, d->numCdfInstances);
    U(texCoords, reinterpret_cast<const float2*>(d->texCoords), d->numTexCoords);
    U(texIndices, d->texIndices, d->numTexIndices);
    U(instances, d->instances, d->numInstances);
#undef U
    D.numCdfInstances = d->numCdfInstances;
    D.numInstances = d->numInstances;
    sc->numEmissive = d->numEmissive;
    sc->hostInstances.assign(d->instances, d->instances + d->numInstances);
    sc->dev.materialMask = 0u;
    for (const RB200Instance& in : sc->hostInstances) sc->dev.materialMask |= 1u << (in.materialIdx > 3u ? 3u : in.materialIdx);   // as k_extend bins
    sc->dev.pad_ = 0u;

    // textures: RGBA8 UNORM cudaArrays behind texture objects (point sampled; the bilinear REPEAT filter is applied in fp32)
    std::vector<uint2> sizes;
    for (uint32_t i = 0; i < d->numTextures; i++) {
        const RB200Texture& t = d->textures[i];
        if (!t.rgba8 || t.width == 0 || t.height == 0) { set_error("texture %u is empty", i); rb200_scene_destroy(sc); return RB200_ERR_INVALID_ARGUMENT; }
        cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uchar4>();
        cudaArray_t arr;
        SC_CUDA(cudaMallocArray(&arr, &fmt, t.width, t.height));
        sc->texArrays.push_back(arr);
        SC_CUDA(cudaMemcpy2DToArrayAsync(arr, 0, 0, t.rgba8, (size_t)t.width * 4, (size_t)t.width * 4, t.height, cudaMemcpyHostToDevice, s));
        cudaResourceDesc rd; memset(&rd, 0, sizeof(rd));
        rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
        cudaTextureDesc td; memset(&td, 0, sizeof(td));
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
        cudaTextureObject_t obj;
        SC_CUDA(cudaCreateTextureObject(&obj, &rd, &td, nullptr));
        sc->texObjects.push_back(obj);
        sizes.push_back(make_uint2(t.width, t.height));
    }
    if ((rc = upload(sc, sc->texObjects.data(), sc->texObjects.size(), &D.textures, s)) != RB200_OK) { rb200_scene_destroy(sc); return rc; }
    if ((rc = upload(sc, sizes.data(), sizes.size(), &D.texSizes, s)) != RB200_OK) { rb200_scene_destroy(sc); return rc; }
    SC_CUDA(cudaStreamSynchronize(s));   // host staging vectors go out of scope below

    // RB200_BVH_BUILDER=lbvh | ploc selects the binary hierarchy under the collapse (see common.cuh); default ploc
    int builder = BUILDER_PLOC;
    if (const char* e = getenv("RB200_BVH_BUILDER")) {
        if (!strcmp(e, "lbvh")) builder = BUILDER_LBVH;
        else if (!strcmp(e, "ploc")) builder = BUILDER_PLOC;
        else { set_error("RB200_BVH_BUILDER must be lbvh or ploc"); rb200_scene_destroy(sc); return RB200_ERR_INVALID_ARGUMENT; }
    }
    BuildInput bi{D.vertices, D.indices, D.instances, &sc->hostInstances, d->numInstances, builder};
    D.tlasNodes = nullptr; D.tlasLeaves = nullptr; D.tlInstances = nullptr;
    if (ctx->flags & RB200_FLAG_TWO_LEVEL) {
        // the reference's structure (src/scene/Scene.cpp:93-111): a hierarchy per distinct object in object space + one over
        // the instances; sc->bvh holds the former (its slots index the shading records), sc->tlas the latter
        std::vector<TwoLevelInstance> entries;
        rc = build_two_level(bi, s, &sc->bvh, &sc->tlas, &entries, &sc->numModels, &ctx->launches);
        if (rc != RB200_OK) { rb200_scene_destroy(sc); return rc; }
        if (2u * sc->tlas.maxDepth + sc->bvh.maxDepth + 2u > (uint32_t)TL_STACK) {
            set_error("two-level hierarchy too deep for the traversal stack (top level %u, objects %u)", sc->tlas.maxDepth, sc->bvh.maxDepth);
            rb200_scene_destroy(sc); return RB200_ERR_INVALID_ARGUMENT;
        }
        if ((rc = upload(sc, entries.data(), entries.size(), &D.tlInstances, s)) != RB200_OK) { rb200_scene_destroy(sc); return rc; }
        SC_CUDA(cudaStreamSynchronize(s));
        D.tlasNodes = sc->tlas.nodes; D.tlasLeaves = sc->tlas.tris;
    } else {
        rc = build_bvh(bi, s, &sc->bvh, &ctx->launches);
        if (rc != RB200_OK) { rb200_scene_destroy(sc); return rc; }
    }
    /* TRAV_MAX_DEPTH = 22: the per-lane traversal stack holds at most 2 groups per tree level */
    if (sc->bvh.maxDepth > (uint32_t)22) { set_error("BVH depth %u exceeds the traversal stack", sc->bvh.maxDepth); rb200_scene_destroy(sc); return RB200_ERR_INVALID_ARGUMENT; }
    D.nodes = sc->bvh.nodes; D.tris = sc->bvh.tris; D.numTris = sc->bvh.numTris;
    D.shadeBase = nullptr; D.shadeFrame = nullptr;
#if RB_SHADE_RECORDS
    {   // per-triangle shading records in leaf order (common.cuh): 192 bytes per triangle
        float4 *base = nullptr, *frame = nullptr;
        const size_t n = std::max<size_t>(D.numTris, 1);
        if (cudaMalloc(&base, n * 4 * sizeof(float4)) != cudaSuccess || cudaMalloc(&frame, n * 8 * sizeof(float4)) != cudaSuccess) {
            cudaGetLastError();
            if (base) cudaFree(base);
            set_error("out of device memory (shading records)");
            rb200_scene_destroy(sc);
            return RB200_ERR_OUT_OF_MEMORY;
        }
        sc->allocations.push_back(base); sc->allocations.push_back(frame);
        if ((rc = build_shade_records(D, base, frame, s)) != RB200_OK) { rb200_scene_destroy(sc); return rc; }
        ctx->launches++;
        D.shadeBase = base; D.shadeFrame = frame;
        SC_CUDA(cudaStreamSynchronize(s));
    }
#endif

    // Optional (RB200_L2_PERSIST=1): pin the hierarchy in L2 with a persisting access-policy window over the
    // node+triangle allocation. Every wave streams ~0.6 GB of path state through the 126 MB L2 and evicts about a
    // third of the 52 MB BVH between visits, but measured on B200 the traversal kernels are issue-bound, not
    // latency-bound (extend 1285 vs 1280 Mrays/s), while the set-aside slows the shading kernels (step 81 vs
    // 75 ms) — so the hint is off by default.
    if (getenv("RB200_L2_PERSIST")) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, ctx->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0) {
            const size_t setAside = std::min((size_t)prop.persistingL2CacheMaxSize, sc->bvh.blobBytes);
            const size_t window = std::min(sc->bvh.blobBytes, (size_t)prop.accessPolicyMaxWindowSize);
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, setAside) == cudaSuccess && window > 0) {
                cudaStreamAttrValue attr;
                memset(&attr, 0, sizeof(attr));
                attr.accessPolicyWindow.base_ptr = sc->bvh.blob;
                attr.accessPolicyWindow.num_bytes = window;
                attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)setAside / (double)window);
                attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                bool ok = true;
                for (int e = 0; e < ctx->numEngines; e++)
                    ok &= cudaStreamSetAttribute(ctx->eng[e].stream, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess;
                ok &= cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess;
                if (ok) sc->l2PersistBytes = setAside;
            }
        }
        cudaGetLastError();   // best effort: a refused hint is not an error
    }
    *out = sc;
    return RB200_OK;
#undef SC_CUDA
}

RB200_API int rb200_scene_destroy(RB200Scene* sc) {
    if (!sc) return RB200_OK;
    if (sc->ctx) { cudaSetDevice(sc->ctx->device); cudaStreamSynchronize(sc->ctx->stream); }
    for (cudaTextureObject_t t : sc->texObjects) cudaDestroyTextureObject(t);
    for (cudaArray_t a : sc->texArrays) cudaFreeArray(a);
    for (void* p : sc->allocations) cudaFree(p);
    free_bvh(&sc->bvh);
    free_bvh(&sc->tlas);
    delete sc;
    return RB200_OK;
}

RB200_API int rb200_scene_bvh_info(const RB200Scene* scene, RB200BvhInfo* out) {
    if (!scene || !out) { set_error("null argument"); return RB200_ERR_INVALID_ARGUMENT; }
    RB200Scene* sc = const_cast<RB200Scene*>(scene);
    if (!sc->hashValid) {
        int rc = hash_bvh(sc->bvh, sc->ctx->stream, &sc->hash);
        if (rc != RB200_OK) return rc;
        if (sc->tlas.nodes) {
            uint64_t top = 0;
            if ((rc = hash_bvh(sc->tlas, sc->ctx->stream, &top)) != RB200_OK) return rc;
            sc->hash = (sc->hash ^ top) * 1099511628211ull;
        }
        sc->hashValid = true;
    }
    memset(out, 0, sizeof(*out));
    out->numTriangles = sc->bvh.numTris; out->numWideNodes = sc->bvh.numNodes; out->maxDepth = sc->bvh.maxDepth;
    out->nodeBytes = (uint64_t)sc->bvh.numNodes * sizeof(WideNode);
    out->triangleBytes = (uint64_t)sc->bvh.numTris * sizeof(TriRecord);
    out->hash = sc->hash; out->buildMs = sc->bvh.buildMs;
    out->reserved = (uint32_t)(sc->l2PersistBytes >> 10);
    for (int a = 0; a < 3; a++) { out->sceneMin[a] = sc->bvh.sceneMin[a]; out->sceneMax[a] = sc->bvh.sceneMax[a]; }
    if (sc->tlas.nodes) {
        // two-level scene: the objects' hierarchies (triangles counted once per object) + the one over the instances
        out->numWideNodes += sc->tlas.numNodes; out->maxDepth += sc->tlas.maxDepth;
        out->nodeBytes += (uint64_t)sc->tlas.numNodes * sizeof(WideNode);
        out->triangleBytes += (uint64_t)sc->tlas.numTris * (sizeof(TriRecord) + sizeof(TwoLevelInstance));
        out->buildMs += sc->tlas.buildMs;
        for (int a = 0; a < 3; a++) { out->sceneMin[a] = sc->tlas.sceneMin[a]; out->sceneMax[a] = sc->tlas.sceneMax[a]; }
    }
    return RB200_OK;
}

RB200_API int rb200_render_batch(RB200Context* ctx, const RB200Scene* scene, const RB200RtPushConsts* pc) {
    if (!ctx || !scene || !pc) { set_error("null argument"); return RB200_ERR_INVALID_ARGUMENT; }
    if (scene->ctx != ctx) { set_error("scene belongs to another context"); return RB200_ERR_INVALID_ARGUMENT; }
    RB_CUDA(cudaSetDevice(ctx->device));
    const int rc = render_batch(ctx, scene, pc);
    if (rc != RB200_OK) return rc;
    // samples are clamped to [0, directClamp] and NaN samples are dropped: the image stays finite unless the clamp is not
    if (!std::isfinite(pc->directClamp)) ctx->hdrMayBeNonFinite = true;
    else if (pc->sampleBatch == 0u && !(ctx->flags & RB200_FLAG_ACCUM_SUM)) ctx->hdrMayBeNonFinite = false;    // batch 0 overwrites the image
    // completion marker of this batch on the front-end stream (which waits for the batch's fold)
    cudaEvent_t e;
    if (!ctx->batchEventPool.empty()) { e = ctx->batchEventPool.back(); ctx->batchEventPool.pop_back(); }
    else RB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync));
    RB_CUDA(cudaEventRecord(e, ctx->stream));
    ctx->batchEvents.push_back(e);
    while (ctx->batchEvents.size() > 64) {      // bound the list for callers that never wait
        if (cudaEventQuery(ctx->batchEvents.front()) != cudaSuccess) break;
        ctx->batchEventPool.push_back(ctx->batchEvents.front()); ctx->batchEvents.pop_front();
    }
    cudaGetLastError();
    return RB200_OK;
}

RB200_API int rb200_wait_batches_pending(RB200Context* ctx, uint32_t max_pending) {
    if (!ctx) { set_error("null context"); return RB200_ERR_INVALID_ARGUMENT; }
    RB_CUDA(cudaSetDevice(ctx->device));
    while (ctx->batchEvents.size() > max_pending) {
        cudaEvent_t e = ctx->batchEvents.front();
        ctx->batchEvents.pop_front();
        ctx->batchEventPool.push_back(e);
        RB_CUDA(cudaEventSynchronize(e));
    }
    return RB200_OK;
}

RB200_API int rb200_resolve_sum(RB200Context* ctx, uint32_t numBatches) {
    if (!ctx) { set_error("null context"); return RB200_ERR_INVALID_ARGUMENT; }
    RB_CUDA(cudaSetDevice(ctx->device));
    return resolve_sum(ctx, numBatches);
}

RB200_API int rb200_postprocess(RB200Context* ctx, const RB200BloomPushConsts* bloom, const RB200TonemappingPushConsts* tm) {
    if (!ctx || !bloom || !tm) { set_error("null argument"); return RB200_ERR_INVALID_ARGUMENT; }
    if (!(bloom->radius > 0.0f)) { set_error("bloom radius must be > 0"); return RB200_ERR_INVALID_ARGUMENT; }
    RB_CUDA(cudaSetDevice(ctx->device));
    return postprocess(ctx, bloom, tm);
}

RB200_API int rb200_context_set_tiles(RB200Context* ctx, uint32_t tileRank, uint32_t tileCount, uint32_t tileSize) {
    if (!ctx) { set_error("null context"); return RB200_ERR_INVALID_ARGUMENT; }
    if (tileCount == 0 || tileRank >= tileCount || tileSize == 0) { set_error("invalid tile partition: need tileRank < tileCount and tileSize > 0"); return RB200_ERR_INVALID_ARGUMENT; }
    RB_CUDA(cudaSetDevice(ctx->device));
    invalidate_speculation(ctx);      // batches in flight were generated with the old partition
    for (int e = 0; e < ctx->numEngines; e++) {
        WaveParams& L = ctx->eng[e].P;
        L.tileRank = tileRank; L.tileCount = tileCount; L.tileSize = tileSize; L.tilesX = (ctx->width + tileSize - 1u) / tileSize;
    }
    return RB200_OK;
}

RB200_API int rb200_present_sum(RB200Context* ctx, const void* device_sum_rgba32f, uint32_t numBatches,
                                const RB200BloomPushConsts* bloom, const RB200TonemappingPushConsts* tm) {
    if (!ctx || !bloom || !tm) { set_error("null argument"); return RB200_ERR_INVALID_ARGUMENT; }
    if (numBatches == 0) { set_error("numBatches must be > 0"); return RB200_ERR_INVALID_ARGUMENT; }
    if (!(bloom->radius > 0.0f)) { set_error("bloom radius must be > 0"); return RB200_ERR_INVALID_ARGUMENT; }
    RB_CUDA(cudaSetDevice(ctx->device));
    return present_sum(ctx, static_cast<const float4*>(device_sum_rgba32f), numBatches, bloom, tm);
}

RB200_API int rb200_read_ldr(RB200Context* ctx, uint8_t* rgba8) {
    if (!ctx || !rgba8) { set_error("null argument"); return RB200_ERR_INVALID_ARGUMENT; }
    RB_CUDA(cudaSetDevice(ctx->device));
    RB_CUDA(cudaMemcpyAsync(rgba8, ctx->ldr, (size_t)ctx->width * ctx->height * 4, cudaMemcpyDeviceToHost, ctx->stream));
    RB_CUDA(cudaStreamSynchronize(ctx->stream));
    return RB200_OK;
}

RB200_API int rb200_read_ldr_async(RB200Context* ctx, uint8_t* rgba8) {
    if (!ctx || !rgba8) { set_error("null argument"); return RB200_ERR_INVALID_ARGUMENT; }
    RB_CUDA(cudaSetDevice(ctx->device));
    RB_CUDA(cudaMemcpyAsync(rgba8, ctx->ldr, (size_t)ctx->width * ctx->height * 4, cudaMemcpyDeviceToHost, ctx->stream));
    cudaEvent_t e;
    if (!ctx->ldrEventPool.empty()) { e = ctx->ldrEventPool.back(); ctx->ldrEventPool.pop_back(); }
    else RB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync));
    ctx->ldrPendingEvents.push_back(e);
    RB_CUDA(cudaEventRecord(e, ctx->stream));
    return RB200_OK;
}

RB200_API int rb200_wait_ldr_pending(RB200Context* ctx, uint32_t max_pending) {
    if (!ctx) { set_error("null context"); return RB200_ERR_INVALID_ARGUMENT; }
    RB_CUDA(cudaSetDevice(ctx->device));
    while (ctx->ldrPendingEvents.size() > max_pending) {
        cudaEvent_t e = ctx->ldrPendingEvents.front();
        ctx->ldrPendingEvents.erase(ctx->ldrPendingEvents.begin());
        ctx->ldrEventPool.push_back(e);
        RB_CUDA(cudaEventSynchronize(e));
    }
    return RB200_OK;
}

RB200_API int rb200_wait_ldr(RB200Context* ctx) { return rb200_wait_ldr_pending(ctx, 0); }

// Frames a display loop should keep in flight so that the host never drains the device: the batches behind the one being
// presented are traced speculatively inside the library, so this no longer depends on the lane count.
RB200_API uint32_t rb200_pipeline_depth(void) { return 4u; }

RB200_API int rb200_engine_config(RB200Context* ctx, uint32_t* engines, uint32_t* lanes, uint64_t* discarded_batches) {
    if (!ctx) { set_error("null context"); return RB200_ERR_INVALID_ARGUMENT; }
    if (engines) *engines = (uint32_t)ctx->numEngines;
    if (lanes) *lanes = (uint32_t)ctx->numLanes;
    if (discarded_batches) {
        uint64_t w = 0;
        for (int e = 0; e < ctx->numEngines; e++) w += ctx->eng[e].wastedBatches;
        *discarded_batches = w;
    }
    return RB200_OK;
}

RB200_API int rb200_host_alloc(size_t bytes, void** out) {
    if (!out || bytes == 0) { set_error("null argument"); return RB200_ERR_INVALID_ARGUMENT; }
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) { cudaGetLastError(); set_error(cudaGetErrorString(e)); return e == cudaErrorMemoryAllocation ? RB200_ERR_OUT_OF_MEMORY : RB200_ERR_CUDA; }
    return RB200_OK;
}

RB200_API int rb200_host_free(void* p) {
    if (!p) return RB200_OK;
    RB_CUDA(cudaFreeHost(p));
    return RB200_OK;
}

RB200_API int rb200_read_hdr(RB200Context* ctx, float* rgba32f) {
    if (!ctx || !rgba32f) { set_error("null argument"); return RB200_ERR_INVALID_ARGUMENT; }
    RB_CUDA(cudaSetDevice(ctx->device));
    RB_CUDA(cudaMemcpyAsync(rgba32f, ctx->wp.image, (size_t)ctx->width * ctx->height * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
    RB_CUDA(cudaStreamSynchronize(ctx->stream));
    return RB200_OK;
}

RB200_API int rb200_write_hdr(RB200Context* ctx, const float* rgba32f) {
    if (!ctx || !rgba32f) { set_error("null argument"); return RB200_ERR_INVALID_ARGUMENT; }
    RB_CUDA(cudaSetDevice(ctx->device));
    RB_CUDA(cudaMemcpyAsync(ctx->wp.image, rgba32f, (size_t)ctx->width * ctx->height * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    // NaN / infinite pixels change how far a pixel reaches in the bloom passes (post.cu, k_blur_exact): note whether there are any
    bool nonFinite = false;
    const size_t n = (size_t)ctx->width * ctx->height;
    for (size_t i = 0; i < n && !nonFinite; i++)
        nonFinite = !std::isfinite(rgba32f[4 * i]) || !std::isfinite(rgba32f[4 * i + 1]) || !std::isfinite(rgba32f[4 * i + 2]);
    ctx->hdrMayBeNonFinite = nonFinite;
    RB_CUDA(cudaStreamSynchronize(ctx->stream));
    return RB200_OK;
}

RB200_API int rb200_hdr_device_ptr(RB200Context* ctx, void** out_device_ptr) {
    if (!ctx || !out_device_ptr) { set_error("null argument"); return RB200_ERR_INVALID_ARGUMENT; }
    *out_device_ptr = ctx->wp.image;
    return RB200_OK;
}

RB200_API int rb200_trace_primary(RB200Context* ctx, const RB200Scene* scene, const RB200RtPushConsts* pc, RB200PrimaryHit* out_hits) {
    if (!ctx || !scene || !pc || !out_hits) { set_error("null argument"); return RB200_ERR_INVALID_ARGUMENT; }
    RB_CUDA(cudaSetDevice(ctx->device));
    return trace_primary(ctx, scene, pc, out_hits);
}

RB200_API int rb200_trace_rays(RB200Context* ctx, const RB200Scene* scene, uint32_t n, const float* origins, const float* directions,
                               const float* tmax, int any_hit, RB200PrimaryHit* out_hits) {
    if (!ctx || !scene || (n && (!origins || !directions || !tmax || !out_hits))) { set_error("null argument"); return RB200_ERR_INVALID_ARGUMENT; }
    RB_CUDA(cudaSetDevice(ctx->device));
    return trace_rays(ctx, scene, n, origins, directions, tmax, any_hit, out_hits);
}

RB200_API int rb200_shade_hits(RB200Context* ctx, const RB200Scene* scene, uint32_t n, const float* origins, const float* directions,
                               const uint32_t* rngStates, const uint32_t* insideDielectric, const float* accumulatedDistance,
                               RB200ShadeResult* out) {
    if (!ctx || !scene || (n && (!origins || !directions || !rngStates || !insideDielectric || !accumulatedDistance || !out))) {
        set_error("null argument"); return RB200_ERR_INVALID_ARGUMENT;
    }
    if (scene->ctx != ctx) { set_error("scene belongs to another context"); return RB200_ERR_INVALID_ARGUMENT; }
    RB_CUDA(cudaSetDevice(ctx->device));
    return shade_hits(ctx, scene, n, origins, directions, rngStates, insideDielectric, accumulatedDistance, out);
}

RB200_API int rb200_bench_trace(RB200Context* ctx, const RB200Scene* scene, uint32_t n, const float* origins, const float* directions,
                                const float* tmax, int any_hit, uint32_t reps, float* out_ms_per_launch, uint64_t* out_checksum) {
    if (!ctx || !scene || !origins || !directions || !tmax) { set_error("null argument"); return RB200_ERR_INVALID_ARGUMENT; }
    RB_CUDA(cudaSetDevice(ctx->device));
    return bench_trace(ctx, scene, n, origins, directions, tmax, any_hit, reps, out_ms_per_launch, out_checksum);
}

RB200_API int rb200_get_stats(RB200Context* ctx, RB200Stats* last_batch, RB200Stats* cumulative) {
    if (!ctx) { set_error("null context"); return RB200_ERR_INVALID_ARGUMENT; }
    // a lane's counters hold its batch; k_accumulate copies them to statsLast and adds them to the cumulative array when
    // the batch is folded. The front-end stream waits for the last fold, which is ordered after every earlier one.
    unsigned long long lastc[ST_COUNT] = {0}, cum[ST_COUNT];
    RB_CUDA(cudaSetDevice(ctx->device));
    RB_CUDA(cudaMemcpyAsync(lastc, ctx->statsLast, sizeof(lastc), cudaMemcpyDeviceToHost, ctx->stream));
    RB_CUDA(cudaMemcpyAsync(cum, ctx->statsSnap, sizeof(cum), cudaMemcpyDeviceToHost, ctx->stream));
    RB_CUDA(cudaStreamSynchronize(ctx->stream));
    RB200Stats& L = ctx->last; RB200Stats& Cm = ctx->cumulative;
    L.extendRays = lastc[ST_EXTEND]; L.shadowRays = lastc[ST_SHADOW];
    L.paths = lastc[ST_PATHS];
    L.nodeVisits = lastc[ST_NODES] + lastc[ST_NODES_SHADOW]; L.triTests = lastc[ST_TRIS] + lastc[ST_TRIS_SHADOW];
    L.shadowNodeVisits = lastc[ST_NODES_SHADOW]; L.shadowTriTests = lastc[ST_TRIS_SHADOW];
    Cm.extendRays = cum[ST_EXTEND]; Cm.shadowRays = cum[ST_SHADOW]; Cm.paths = cum[ST_PATHS];
    Cm.nodeVisits = cum[ST_NODES] + cum[ST_NODES_SHADOW]; Cm.triTests = cum[ST_TRIS] + cum[ST_TRIS_SHADOW];
    Cm.shadowNodeVisits = cum[ST_NODES_SHADOW]; Cm.shadowTriTests = cum[ST_TRIS_SHADOW]; Cm.shadowRaysSkipped = cum[ST_SHADOW_SKIPPED]; Cm.kernelLaunches = ctx->launches;
    if (last_batch) *last_batch = L;
    if (cumulative) *cumulative = Cm;
    return RB200_OK;
}

RB200_API int rb200_get_kernel_times(RB200Context* ctx, RB200KernelTimes* out) {
    if (!ctx || !out) { set_error("null argument"); return RB200_ERR_INVALID_ARGUMENT; }
    if (!(ctx->flags & RB200_FLAG_TIME_KERNELS)) { set_error("context was not created with RB200_FLAG_TIME_KERNELS"); return RB200_ERR_INVALID_ARGUMENT; }
    RB_CUDA(cudaSetDevice(ctx->device));
    RB_CUDA(cudaStreamSynchronize(ctx->stream));
    memset(out, 0, sizeof(*out));
    std::vector<uint32_t> waveCounts((size_t)ctx->waveCountsWaves * CNT_SET, 0u);
    if (ctx->waveCountsDev && ctx->waveCountsWaves)
        RB_CUDA(cudaMemcpy(waveCounts.data(), ctx->waveCountsDev, waveCounts.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (uint32_t w = 0; w < ctx->waveCountsWaves; w++) {
        const uint32_t* c = &waveCounts[(size_t)w * CNT_SET];
        out->extendRays += c[CNT_RAYS]; out->shadowRays += c[CNT_SHADOW]; out->finishItems += c[CNT_END];
        for (int m = 0; m < 5; m++) out->shadeItems[m] += c[CNT_MAT0 + m];
    }
    const uint32_t half = ctx->width * ctx->height / 2u;
    uint32_t extendWave = 0, shadowWave = 0;
    for (size_t i = 0; i < ctx->evClass.size(); i++) {
        float ms = 0.f;
        RB_CUDA(cudaEventElapsedTime(&ms, ctx->evPool[2 * i], ctx->evPool[2 * i + 1]));
        const int c = ctx->evClass[i];
        if (c == 0) out->generateMs += ms;
        else if (c == 1) {
            out->extendMs += ms; out->extendLaunches++;
            const uint32_t n = extendWave < ctx->waveCountsWaves ? waveCounts[(size_t)CNT_SET * extendWave + CNT_RAYS] : 0u;
            if (n >= half) { out->extendFullMs += ms; out->extendFullLaunches++; out->extendFullRays += n; }
            extendWave++;
        }
        else if (c >= 2 && c <= 6) { out->shadeMs[c - 2] += ms; out->shadeLaunches++; }
        else if (c == 7) {
            out->shadowMs += ms; out->shadowLaunches++;
            const uint32_t n = shadowWave < ctx->waveCountsWaves ? waveCounts[(size_t)CNT_SET * shadowWave + CNT_SHADOW] : 0u;
            if (n >= half) { out->shadowFullMs += ms; out->shadowFullLaunches++; out->shadowFullRays += n; }
            shadowWave++;
        }
        else { out->finishMs += ms; out->finishLaunches++; }
    }
    return RB200_OK;
}

// ---------------------------------------------------------------------------------------------------
// Roofline denominator for the traversal kernels: how fast can this GPU gather node-sized records from an
// L2-resident table? Every thread draws independent pseudo-random record indices and reads the whole record with
// 16-byte loads; nothing depends on the loaded values, so the figure is a bandwidth (an upper bound for a
// traversal, whose next address depends on the node it just read).
// ---------------------------------------------------------------------------------------------------
} // extern "C" (templates need C++ linkage)
template <int VEC>
__global__ void __launch_bounds__(256) k_gather_bench(const uint4* __restrict__ table, uint32_t numRecords, uint32_t perThread,
                                                        uint32_t seed, uint32_t* __restrict__ sink) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + seed;
    uint4 acc = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 4
    for (uint32_t k = 0; k < perThread; k++) {
        s = s * 747796405u + 2891336453u;
        const uint32_t w = ((s >> ((s >> 28) + 4u)) ^ s) * 277803737u;
        const uint32_t r = (uint32_t)(((unsigned long long)(w ^ (w >> 22)) * numRecords) >> 32);
        const uint4* p = table + (size_t)r * VEC;
#pragma unroll
        for (int j = 0; j < VEC; j++) { const uint4 v = __ldg(p + j); acc.x ^= v.x; acc.y += v.y; acc.z ^= v.z; acc.w += v.w; }
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x9E3779B9u) sink[0] = acc.x;     // keeps the loads alive
}
extern "C" {

RB200_API int rb200_measure_gather(RB200Context* ctx, size_t tableBytes, uint32_t recordBytes, float* out_gbps) {
    if (!ctx || !out_gbps) { set_error("null argument"); return RB200_ERR_INVALID_ARGUMENT; }
    if (recordBytes != 80 && recordBytes != 48 && recordBytes != 16) { set_error("recordBytes must be 16, 48 or 80"); return RB200_ERR_INVALID_ARGUMENT; }
    if (tableBytes < (size_t)recordBytes * 1024 || tableBytes > ((size_t)8 << 30)) { set_error("tableBytes out of range"); return RB200_ERR_INVALID_ARGUMENT; }
    RB_CUDA(cudaSetDevice(ctx->device));
    const uint32_t numRecords = (uint32_t)(tableBytes / recordBytes);
    uint4* table = nullptr; uint32_t* sink = nullptr;
    if (cudaMalloc(&table, (size_t)numRecords * recordBytes) != cudaSuccess || cudaMalloc(&sink, 4) != cudaSuccess) {
        cudaGetLastError(); if (table) cudaFree(table);
        set_error("out of device memory"); return RB200_ERR_OUT_OF_MEMORY;
    }
    cudaStream_t s = ctx->stream;
    cudaMemsetAsync(table, 0x5A, (size_t)numRecords * recordBytes, s);
    const uint32_t blocks = (uint32_t)ctx->numSMs * 8u, perThread = 256u;
    auto launch = [&](uint32_t seed) {
        if (recordBytes == 80) k_gather_bench<5><<<blocks, 256, 0, s>>>(table, numRecords, perThread, seed, sink);
        else if (recordBytes == 48) k_gather_bench<3><<<blocks, 256, 0, s>>>(table, numRecords, perThread, seed, sink);
        else k_gather_bench<1><<<blocks, 256, 0, s>>>(table, numRecords, perThread, seed, sink);
    };
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (uint32_t i = 0; i < 3; i++) launch(i);          // warm-up: pulls the table into L2
    const int reps = 10;
    cudaEventRecord(e0, s);
    for (int i = 0; i < reps; i++) launch(100u + (uint32_t)i);
    cudaEventRecord(e1, s);
    cudaError_t err = cudaStreamSynchronize(s);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(table); cudaFree(sink);
    ctx->launches += 3 + reps;
    if (err != cudaSuccess || cudaGetLastError() != cudaSuccess || !(ms > 0.f)) { set_error("gather micro-benchmark failed"); return RB200_ERR_CUDA; }
    const double bytes = (double)reps * blocks * 256.0 * perThread * recordBytes;
    *out_gbps = (float)(bytes / (ms * 1e-3) / 1e9);
    return RB200_OK;
}

RB200_API int rb200_synchronize(RB200Context* ctx) {
    if (!ctx) { set_error("null context"); return RB200_ERR_INVALID_ARGUMENT; }
    RB_CUDA(cudaSetDevice(ctx->device));
    RB_CUDA(cudaStreamSynchronize(ctx->stream));
    return RB200_OK;
}

} // extern "C"
