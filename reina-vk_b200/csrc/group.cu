// group.cu — several B200s behind the C ABI (SURVEY.md 8b "Context", 8e).
//
// The path shards by samples: device g of n renders batches first + g, first + g + n, ... with the unmodified seed
// formula into a local SUM image (RB200_FLAG_ACCUM_SUM), the scene and its BVH are replicated (the build is
// deterministic, so the hashes must agree — checked at scene creation), and ONE ncclReduce(float, 4 * W * H) of
// stream-ordered snapshots of the images to the root closes a frame; the root resolves (sum / batches), blooms,
// tonemaps and reads back. In latency mode (RB200_FLAG_GROUP_TILES) every device traces its interleaved tiles of every
// batch with the reference's running average instead, and the same reduce yields the single-GPU image bit for bit.
//
// Two ways in:
//   * one process, n devices: rb200_group_* — ncclCommInitAll, one host thread per GPU (each device's calls, graph
//     launches and NCCL calls are issued by its own thread, so no device waits for another device's host work);
//   * one process per GPU (torchrun, MPI): rb200_comm_unique_id + rb200_context_comm_init (ncclCommInitRank) and
//     rb200_context_reduce_present on every rank.
// NCCL is loaded at run time (dlopen "libnccl.so.2"): a process that already holds an NCCL (torch's) shares it, and
// librb200.so itself does not depend on NCCL being installed for single-GPU use.
#include "context.cuh"
#include <dlfcn.h>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>

namespace rb200 {

// ---- the few NCCL entry points used, resolved at run time (types as in nccl.h 2.27) ----
typedef struct { char internal[128]; } NcclUniqueId;
typedef void* NcclComm;
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommInitAll)(NcclComm*, int, const int*) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
};
static constexpr int NCCL_FLOAT32 = 7, NCCL_SUM = 0;

static NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.handle) break; }
        if (!api.handle) return;
#define SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name))
        SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommInitAll, "ncclCommInitAll");
        SYM(CommDestroy, "ncclCommDestroy"); SYM(Reduce, "ncclReduce"); SYM(GroupStart, "ncclGroupStart");
        SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString"); SYM(GetVersion, "ncclGetVersion");
#undef SYM
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommInitAll || !api.CommDestroy || !api.Reduce || !api.GroupStart ||
            !api.GroupEnd || !api.GetErrorString) { dlclose(api.handle); api.handle = nullptr; }
    });
    return api.handle ? &api : nullptr;
}

#define RB_NCCL(api, call)                                                                                     \
    do {                                                                                                       \
        int r_ = (call);                                                                                       \
        if (r_ != 0) { rb200::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, (api)->GetErrorString(r_)); return RB200_ERR_CUDA; } \
    } while (0)

// per-context communicator state (multi-process flavour and the members of a group)
struct CommState {
    NcclComm comm = nullptr;
    int rank = 0, nranks = 1;
    float4* snapshot = nullptr;     // stream-ordered copy of this rank's image: what the reduce reads
    float4* reduced = nullptr;      // root only: the reduce's result
};

static std::mutex g_commMutex;
static std::vector<std::pair<RB200Context*, CommState*>> g_comms;
static CommState* comm_of(RB200Context* ctx) {
    std::lock_guard<std::mutex> lk(g_commMutex);
    for (auto& p : g_comms) if (p.first == ctx) return p.second;
    return nullptr;
}

static int comm_attach(RB200Context* ctx, NcclComm comm, int rank, int nranks, int root) {
    CommState* cs = new CommState();
    cs->comm = comm; cs->rank = rank; cs->nranks = nranks;
    const size_t n = (size_t)ctx->width * ctx->height;
    RB_CUDA(cudaSetDevice(ctx->device));
    RB_CUDA(cudaMalloc(&cs->snapshot, n * sizeof(float4)));
    if (rank == root) RB_CUDA(cudaMalloc(&cs->reduced, n * sizeof(float4)));
    std::lock_guard<std::mutex> lk(g_commMutex);
    g_comms.push_back({ctx, cs});
    return RB200_OK;
}

static void comm_detach(RB200Context* ctx) {
    CommState* cs = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_commMutex);
        for (size_t i = 0; i < g_comms.size(); i++) if (g_comms[i].first == ctx) { cs = g_comms[i].second; g_comms.erase(g_comms.begin() + i); break; }
    }
    if (!cs) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (NcclApi* api = nccl_api()) if (cs->comm) api->CommDestroy(cs->comm);
    if (cs->snapshot) cudaFree(cs->snapshot);
    if (cs->reduced) cudaFree(cs->reduced);
    delete cs;
}

// snapshot of the context's image behind everything queued so far, reduced to `root` (one ncclReduce on the context's
// front-end stream: the engines keep tracing the next batches underneath)
static int reduce_image(RB200Context* ctx, CommState* cs, int root) {
    NcclApi* api = nccl_api();
    const size_t n = (size_t)ctx->width * ctx->height;
    RB_CUDA(cudaSetDevice(ctx->device));
    RB_CUDA(cudaMemcpyAsync(cs->snapshot, ctx->wp.image, n * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
    RB_NCCL(api, api->Reduce(cs->snapshot, cs->rank == root ? cs->reduced : nullptr, n * 4, NCCL_FLOAT32, NCCL_SUM, root, cs->comm, ctx->stream));
    return RB200_OK;
}

// ---- worker threads of a group: one per device ----
struct Worker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv, done;
    std::deque<std::function<int()>> jobs;
    int pending = 0;
    int firstError = RB200_OK;
    std::string errorText;
    bool quit = false;
    void start() {
        th = std::thread([this] {
            for (;;) {
                std::function<int()> job;
                {
                    std::unique_lock<std::mutex> lk(m);
                    cv.wait(lk, [this] { return quit || !jobs.empty(); });
                    if (jobs.empty()) return;
                    job = std::move(jobs.front()); jobs.pop_front();
                }
                const int rc = job();
                {
                    std::lock_guard<std::mutex> lk(m);
                    if (rc != RB200_OK && firstError == RB200_OK) { firstError = rc; errorText = rb200_last_error(); }
                    pending--;
                }
                done.notify_all();
            }
        });
    }
    void post(std::function<int()> f) {
        { std::lock_guard<std::mutex> lk(m); jobs.push_back(std::move(f)); pending++; }
        cv.notify_one();
    }
    int wait() {      // until every posted job has run; returns (and clears) the first error
        std::unique_lock<std::mutex> lk(m);
        done.wait(lk, [this] { return pending == 0; });
        const int rc = firstError;
        if (rc != RB200_OK) set_error("%s", errorText.c_str());
        firstError = RB200_OK;
        return rc;
    }
    void stop() {
        { std::lock_guard<std::mutex> lk(m); quit = true; }
        cv.notify_one();
        if (th.joinable()) th.join();
    }
};

} // namespace rb200

using namespace rb200;

// Host-side hand-shake of one presented frame in peer-store latency mode (see rb200_group_present): the members tell the
// root that their "tiles done" events are recorded, the root tells the members that its snapshot event is recorded.
struct PresentSync {
    std::mutex m;
    std::condition_variable cv;
    uint64_t arrived = 0, released = 0;
    void arrive() { { std::lock_guard<std::mutex> lk(m); arrived++; } cv.notify_all(); }
    void wait_arrivals(uint64_t target) { std::unique_lock<std::mutex> lk(m); cv.wait(lk, [&] { return arrived >= target; }); }
    void release(uint64_t gen) { { std::lock_guard<std::mutex> lk(m); released = gen; } cv.notify_all(); }
    void wait_release(uint64_t gen) { std::unique_lock<std::mutex> lk(m); cv.wait(lk, [&] { return released >= gen; }); }
};

struct RB200Group {
    std::vector<int> devices;
    std::vector<RB200Context*> ctx;
    std::vector<Worker*> workers;
    uint32_t width = 0, height = 0, flags = 0;
    bool tiles = false;
    // Latency mode with peer access between the devices: every member's k_accumulate stores the pixels it owns straight
    // into device 0's image over NVLink (RB200Context::peerImage), so a presented frame needs no collective at all — the
    // root waits for the members' "tiles done" events, snapshots its own image and post-processes the snapshot.
    bool peerTiles = false;
    std::vector<cudaEvent_t> tileDone;      // per member, recorded on its front-end stream
    cudaEvent_t rootSnap = nullptr;         // root: its snapshot of the image has been taken
    float4* rootSnapshot = nullptr;
    uint64_t presentGen = 0;
    PresentSync sync;
    uint64_t batchesRendered = 0;       // over all devices (sum mode: the divisor of the resolve)
    int n() const { return (int)devices.size(); }
    int wait_all() {
        int rc = RB200_OK;
        for (Worker* w : workers) { const int r = w->wait(); if (rc == RB200_OK) rc = r; }
        return rc;
    }
};

struct RB200GroupScene {
    RB200Group* group = nullptr;
    std::vector<RB200Scene*> scenes;
};

extern "C" {

RB200_API int rb200_comm_unique_id(void* out_128_bytes) {
    if (!out_128_bytes) { set_error("null argument"); return RB200_ERR_INVALID_ARGUMENT; }
    NcclApi* api = nccl_api();
    if (!api) { set_error("NCCL is not available (libnccl.so.2 could not be loaded)"); return RB200_ERR_NO_DEVICE; }
    NcclUniqueId id;
    RB_NCCL(api, api->GetUniqueId(&id));
    memcpy(out_128_bytes, &id, sizeof(id));
    return RB200_OK;
}

RB200_API int rb200_context_comm_init(RB200Context* ctx, const void* unique_id_128_bytes, int rank, int nranks) {
    if (!ctx || !unique_id_128_bytes || nranks < 1 || rank < 0 || rank >= nranks) { set_error("invalid communicator arguments"); return RB200_ERR_INVALID_ARGUMENT; }
    if (comm_of(ctx)) { set_error("context already has a communicator"); return RB200_ERR_INVALID_ARGUMENT; }
    NcclApi* api = nccl_api();
    if (!api) { set_error("NCCL is not available (libnccl.so.2 could not be loaded)"); return RB200_ERR_NO_DEVICE; }
    RB_CUDA(cudaSetDevice(ctx->device));
    NcclUniqueId id;
    memcpy(&id, unique_id_128_bytes, sizeof(id));
    NcclComm comm = nullptr;
    RB_NCCL(api, api->CommInitRank(&comm, nranks, id, rank));
    return comm_attach(ctx, comm, rank, nranks, 0);
}

RB200_API int rb200_context_comm_destroy(RB200Context* ctx) {
    if (!ctx) return RB200_OK;
    comm_detach(ctx);
    return RB200_OK;
}

RB200_API int rb200_context_reduce_present(RB200Context* ctx, uint32_t totalBatches, const RB200BloomPushConsts* bloom,
                                           const RB200TonemappingPushConsts* tm) {
    if (!ctx) { set_error("null context"); return RB200_ERR_INVALID_ARGUMENT; }
    CommState* cs = comm_of(ctx);
    if (!cs) { set_error("context has no communicator (rb200_context_comm_init)"); return RB200_ERR_INVALID_ARGUMENT; }
    int rc = reduce_image(ctx, cs, 0);
    if (rc != RB200_OK || cs->rank != 0) return rc;
    if (!bloom || !tm) { set_error("null argument"); return RB200_ERR_INVALID_ARGUMENT; }
    if (ctx->flags & RB200_FLAG_ACCUM_SUM) return rb200_present_sum(ctx, cs->reduced, totalBatches, bloom, tm);
    return postprocess(ctx, bloom, tm, cs->reduced);       // tile mode: the reduced image is the running average itself
}

RB200_API int rb200_context_reduced_device_ptr(RB200Context* ctx, void** out_device_ptr) {
    if (!ctx || !out_device_ptr) { set_error("null argument"); return RB200_ERR_INVALID_ARGUMENT; }
    CommState* cs = comm_of(ctx);
    if (!cs || !cs->reduced) { set_error("not the root of a communicator"); return RB200_ERR_INVALID_ARGUMENT; }
    *out_device_ptr = cs->reduced;
    return RB200_OK;
}

// ---------------------------------------------------------------------------------------------------
// one process, n devices
// ---------------------------------------------------------------------------------------------------
RB200_API int rb200_group_destroy(RB200Group* g) {
    if (!g) return RB200_OK;
    for (Worker* w : g->workers) { w->wait(); w->stop(); delete w; }
    // peer-store latency mode (also after a failed set-up: whatever was created is released)
    for (size_t i = 0; i < g->tileDone.size(); i++)
        if (g->tileDone[i]) {
            cudaSetDevice(g->devices[i]);
            if (g->ctx[i]) cudaStreamSynchronize(g->ctx[i]->stream);
            cudaEventDestroy(g->tileDone[i]);
        }
    if (g->rootSnap || g->rootSnapshot) {
        cudaSetDevice(g->devices[0]);
        if (g->ctx[0]) cudaStreamSynchronize(g->ctx[0]->stream);
        if (g->rootSnap) cudaEventDestroy(g->rootSnap);
        if (g->rootSnapshot) cudaFree(g->rootSnapshot);
    }
    for (RB200Context* c : g->ctx) if (c) { comm_detach(c); rb200_context_destroy(c); }
    delete g;
    return RB200_OK;
}

RB200_API int rb200_group_create(uint32_t width, uint32_t height, const int* devices, int numDevices, uint32_t flags, RB200Group** out) {
    if (!out || !devices || numDevices < 1 || numDevices > 64) { set_error("invalid device list"); return RB200_ERR_INVALID_ARGUMENT; }
    for (int i = 0; i < numDevices; i++) for (int j = 0; j < i; j++)
        if (devices[i] == devices[j]) { set_error("device %d listed twice", devices[i]); return RB200_ERR_INVALID_ARGUMENT; }
    NcclApi* api = numDevices > 1 ? nccl_api() : nullptr;
    if (numDevices > 1 && !api) { set_error("NCCL is not available (libnccl.so.2 could not be loaded)"); return RB200_ERR_NO_DEVICE; }
    RB200Group* g = new RB200Group();
    g->devices.assign(devices, devices + numDevices);
    g->width = width; g->height = height;
    g->tiles = (flags & RB200_FLAG_GROUP_TILES) != 0;
    // sample split: local sums, one reduce, resolve on the root; latency mode: running averages of disjoint tiles
    g->flags = g->tiles ? (flags & ~(uint32_t)(RB200_FLAG_GROUP_TILES | RB200_FLAG_ACCUM_SUM)) : (flags | RB200_FLAG_ACCUM_SUM);
    g->ctx.assign(numDevices, nullptr);
    for (int i = 0; i < numDevices; i++) { Worker* w = new Worker(); w->start(); g->workers.push_back(w); }
    for (int i = 0; i < numDevices; i++)
        // latency mode: a member traces 1 / n of the pixels, so twice the lanes keep its launches from getting thin (one device's
        // share of an 8-device frame, measured alone: 7.5 ms per batch with 8 lanes per engine, 6.05 ms with 16)
        g->workers[i]->post([g, i, numDevices] {
            // (path state is allocated per lane for the whole image: ~0.5 KB per pixel and lane over both engines, 16 GB per
            // device at 1080p with 16 lanes; larger images keep the default)
            const bool roomy = (uint64_t)g->width * g->height <= (4ull << 20);
            return rb200::context_create(g->width, g->height, g->devices[i], g->flags, g->tiles && numDevices > 1 && roomy ? 16 : 0, &g->ctx[i]);
        });
    int rc = g->wait_all();
    if (rc != RB200_OK) { rb200_group_destroy(g); return rc; }
    if (g->tiles)
        for (int i = 0; i < numDevices; i++)
            if ((rc = rb200_context_set_tiles(g->ctx[i], (uint32_t)i, (uint32_t)numDevices, 32u)) != RB200_OK) { rb200_group_destroy(g); return rc; }
    if (numDevices > 1) {
        std::vector<NcclComm> comms(numDevices, nullptr);
        const int r = api->CommInitAll(comms.data(), numDevices, devices);
        if (r != 0) { set_error("ncclCommInitAll -> %s", api->GetErrorString(r)); rb200_group_destroy(g); return RB200_ERR_CUDA; }
        for (int i = 0; i < numDevices; i++)
            if ((rc = comm_attach(g->ctx[i], comms[i], i, numDevices, 0)) != RB200_OK) { rb200_group_destroy(g); return rc; }
    }
    // latency mode: peer stores instead of the reduce when every member can reach device 0's memory
    // (RB200_GROUP_TILES_REDUCE=1 keeps the reduce, for comparison)
    const char* forceReduce = getenv("RB200_GROUP_TILES_REDUCE");
    if (g->tiles && numDevices > 1 && !(forceReduce && atoi(forceReduce) != 0)) {
        bool all = true;
        for (int i = 1; i < numDevices && all; i++) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[i], devices[0]) != cudaSuccess || !can) all = false;
        }
        if (all) {
            g->tileDone.assign(numDevices, nullptr);
            for (int i = 1; i < numDevices; i++)
                g->workers[i]->post([g, i]() -> int {
                    RB_CUDA(cudaSetDevice(g->devices[i]));
                    const cudaError_t e = cudaDeviceEnablePeerAccess(g->devices[0], 0);
                    if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                    else if (e != cudaSuccess) { set_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", g->devices[i], g->devices[0], cudaGetErrorString(e)); return RB200_ERR_CUDA; }
                    RB_CUDA(cudaEventCreateWithFlags(&g->tileDone[i], cudaEventDisableTiming));
                    return RB200_OK;
                });
            g->workers[0]->post([g]() -> int {
                RB_CUDA(cudaSetDevice(g->devices[0]));
                RB_CUDA(cudaEventCreateWithFlags(&g->rootSnap, cudaEventDisableTiming));
                RB_CUDA(cudaMalloc(&g->rootSnapshot, (size_t)g->width * g->height * sizeof(float4)));
                return RB200_OK;
            });
            if ((rc = g->wait_all()) != RB200_OK) { rb200_group_destroy(g); return rc; }
            for (int i = 1; i < numDevices; i++) g->ctx[i]->peerImage = g->ctx[0]->wp.image;
            g->peerTiles = true;
        }
    }
    *out = g;
    return RB200_OK;
}

RB200_API int rb200_group_size(const RB200Group* g) { return g ? g->n() : 0; }

RB200_API int rb200_group_uses_peer_stores(const RB200Group* g) { return g && g->peerTiles ? 1 : 0; }

RB200_API int rb200_group_context(RB200Group* g, int index, RB200Context** out) {
    if (!g || !out || index < 0 || index >= g->n()) { set_error("invalid group member"); return RB200_ERR_INVALID_ARGUMENT; }
    *out = g->ctx[index];
    return RB200_OK;
}

RB200_API int rb200_group_set_tile_size(RB200Group* g, uint32_t tileSize) {
    if (!g) { set_error("null group"); return RB200_ERR_INVALID_ARGUMENT; }
    if (!g->tiles) { set_error("the group was not created with RB200_FLAG_GROUP_TILES"); return RB200_ERR_INVALID_ARGUMENT; }
    if (tileSize == 0) { set_error("tileSize must be > 0"); return RB200_ERR_INVALID_ARGUMENT; }
    if (g->batchesRendered) { set_error("the tile size must be chosen before the first batch"); return RB200_ERR_INVALID_ARGUMENT; }
    int rc = g->wait_all();
    for (int i = 0; i < g->n() && rc == RB200_OK; i++) rc = rb200_context_set_tiles(g->ctx[i], (uint32_t)i, (uint32_t)g->n(), tileSize);
    return rc;
}

RB200_API int rb200_group_scene_destroy(RB200GroupScene* s) {
    if (!s) return RB200_OK;
    RB200Group* g = s->group;
    for (int i = 0; i < g->n(); i++)
        if (s->scenes[i]) { RB200Scene* sc = s->scenes[i]; g->workers[i]->post([sc] { return rb200_scene_destroy(sc); }); }
    g->wait_all();
    delete s;
    return RB200_OK;
}

RB200_API int rb200_group_scene_create(RB200Group* g, const RB200SceneDesc* desc, RB200GroupScene** out) {
    if (!g || !desc || !out) { set_error("null argument"); return RB200_ERR_INVALID_ARGUMENT; }
    RB200GroupScene* s = new RB200GroupScene();
    s->group = g;
    s->scenes.assign(g->n(), nullptr);
    // every device uploads the tables and builds its own copy of the hierarchy, concurrently
    for (int i = 0; i < g->n(); i++)
        g->workers[i]->post([g, s, desc, i] { return rb200_scene_create(g->ctx[i], desc, &s->scenes[i]); });
    int rc = g->wait_all();
    if (rc != RB200_OK) { rb200_group_scene_destroy(s); return rc; }
    // the build is deterministic: every replica must be the same structure, bit for bit
    std::vector<RB200BvhInfo> info(g->n());
    for (int i = 0; i < g->n(); i++)
        g->workers[i]->post([s, &info, i] { return rb200_scene_bvh_info(s->scenes[i], &info[i]); });
    rc = g->wait_all();
    if (rc != RB200_OK) { rb200_group_scene_destroy(s); return rc; }
    for (int i = 1; i < g->n(); i++)
        if (info[i].hash != info[0].hash) {
            set_error("BVH replicas differ: device %d hash %016llx, device %d hash %016llx", g->devices[0], (unsigned long long)info[0].hash,
                      g->devices[i], (unsigned long long)info[i].hash);
            rb200_group_scene_destroy(s);
            return RB200_ERR_CUDA;
        }
    *out = s;
    return RB200_OK;
}

RB200_API int rb200_group_scene_bvh_info(const RB200GroupScene* s, int index, RB200BvhInfo* out) {
    if (!s || !out || index < 0 || index >= s->group->n()) { set_error("invalid argument"); return RB200_ERR_INVALID_ARGUMENT; }
    RB200Group* g = s->group;
    g->workers[index]->post([s, out, index] { return rb200_scene_bvh_info(s->scenes[index], out); });
    return g->workers[index]->wait();
}

// Sample split: device i renders batches firstBatch + i + k * n, k = 0 .. batchesPerDevice - 1, of the sequence whose push
// constants are *pc with sampleBatch replaced (n * batchesPerDevice batches in all). Latency mode (RB200_FLAG_GROUP_TILES):
// every device renders its tiles of batches firstBatch .. firstBatch + batchesPerDevice - 1. Asynchronous: returns when
// the jobs are queued on the devices' threads.
RB200_API int rb200_group_render_batches(RB200Group* g, const RB200GroupScene* s, const RB200RtPushConsts* pc, uint32_t firstBatch,
                                         uint32_t batchesPerDevice) {
    if (!g || !s || !pc || s->group != g) { set_error("invalid argument"); return RB200_ERR_INVALID_ARGUMENT; }
    const RB200RtPushConsts base = *pc;
    const int n = g->n();
    for (int i = 0; i < n; i++)
        g->workers[i]->post([g, s, base, firstBatch, batchesPerDevice, i, n] {
            RB200RtPushConsts p = base;
            for (uint32_t k = 0; k < batchesPerDevice; k++) {
                p.sampleBatch = g->tiles ? firstBatch + k : firstBatch + (uint32_t)i + k * (uint32_t)n;
                const int rc = rb200_render_batch(g->ctx[i], s->scenes[i], &p);
                if (rc != RB200_OK) return rc;
            }
            return (int)RB200_OK;
        });
    g->batchesRendered += g->tiles ? (uint64_t)batchesPerDevice : (uint64_t)batchesPerDevice * (uint64_t)n;
    return RB200_OK;
}

// One presented frame: every device snapshots its SUM image behind the batches queued so far and joins ONE ncclReduce to
// device 0, which resolves (sum / batches rendered so far), blooms and tonemaps the reduced copy into its RGBA8 frame.
// Asynchronous on every device's stream; the accumulation images are not touched, so rendering continues underneath.
RB200_API int rb200_group_present(RB200Group* g, const RB200BloomPushConsts* bloom, const RB200TonemappingPushConsts* tm) {
    if (!g || !bloom || !tm) { set_error("null argument"); return RB200_ERR_INVALID_ARGUMENT; }
    if (g->batchesRendered == 0) { set_error("nothing rendered yet"); return RB200_ERR_INVALID_ARGUMENT; }
    const uint32_t total = (uint32_t)g->batchesRendered;
    const RB200BloomPushConsts b = *bloom;
    const RB200TonemappingPushConsts t = *tm;
    if (g->n() == 1) {
        if (g->tiles) g->workers[0]->post([g, b, t] { return rb200_postprocess(g->ctx[0], &b, &t); });
        else g->workers[0]->post([g, total, b, t] { return rb200_present_sum(g->ctx[0], nullptr, total, &b, &t); });
        return RB200_OK;
    }
    if (g->peerTiles) {
        // Every pixel of device 0's image was stored by its owner (k_accumulate, over NVLink for the other devices). A frame:
        // members record "my tiles of everything queued so far are in" on their streams; the root waits for those events,
        // snapshots its image and post-processes the snapshot; the members' streams wait for the snapshot, so the stores of
        // later batches cannot tear the frame. The condition variables only order the host-side record / wait calls.
        const uint64_t gen = ++g->presentGen;
        const int n = g->n();
        for (int i = 1; i < n; i++)
            g->workers[i]->post([g, i, gen]() -> int {
                cudaError_t e = cudaSetDevice(g->devices[i]);
                if (e == cudaSuccess) e = cudaEventRecord(g->tileDone[i], g->ctx[i]->stream);
                g->sync.arrive();
                g->sync.wait_release(gen);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(g->ctx[i]->stream, g->rootSnap, 0);
                if (e != cudaSuccess) { set_error("latency-mode present on device %d: %s", g->devices[i], cudaGetErrorString(e)); return RB200_ERR_CUDA; }
                return RB200_OK;
            });
        g->workers[0]->post([g, b, t, gen, n]() -> int {
            g->sync.wait_arrivals(gen * (uint64_t)(n - 1));
            RB200Context* c = g->ctx[0];
            cudaError_t e = cudaSetDevice(c->device);
            for (int i = 1; i < n && e == cudaSuccess; i++) e = cudaStreamWaitEvent(c->stream, g->tileDone[i], 0);
            if (e == cudaSuccess) e = cudaMemcpyAsync(g->rootSnapshot, c->wp.image, (size_t)g->width * g->height * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream);
            if (e == cudaSuccess) e = cudaEventRecord(g->rootSnap, c->stream);
            g->sync.release(gen);
            if (e != cudaSuccess) { set_error("latency-mode present on device %d: %s", c->device, cudaGetErrorString(e)); return RB200_ERR_CUDA; }
            return postprocess(c, &b, &t, g->rootSnapshot);
        });
        return RB200_OK;
    }
    for (int i = 0; i < g->n(); i++)
        g->workers[i]->post([g, total, b, t, i] { return rb200_context_reduce_present(g->ctx[i], total, &b, &t); });
    return RB200_OK;
}

RB200_API int rb200_group_synchronize(RB200Group* g) {
    if (!g) { set_error("null group"); return RB200_ERR_INVALID_ARGUMENT; }
    int rc = g->wait_all();
    if (rc != RB200_OK) return rc;
    for (int i = 0; i < g->n(); i++) g->workers[i]->post([g, i] { return rb200_synchronize(g->ctx[i]); });
    return g->wait_all();
}

RB200_API int rb200_group_read_ldr(RB200Group* g, uint8_t* rgba8) {
    if (!g || !rgba8) { set_error("null argument"); return RB200_ERR_INVALID_ARGUMENT; }
    int rc = g->wait_all();
    if (rc != RB200_OK) return rc;
    g->workers[0]->post([g, rgba8] { return rb200_read_ldr(g->ctx[0], rgba8); });
    return g->workers[0]->wait();
}

// The mean image of everything rendered so far (W*H*4 floats): reduce + resolve on device 0, blocking.
RB200_API int rb200_group_read_hdr(RB200Group* g, float* rgba32f) {
    if (!g || !rgba32f) { set_error("null argument"); return RB200_ERR_INVALID_ARGUMENT; }
    if (g->batchesRendered == 0) { set_error("nothing rendered yet"); return RB200_ERR_INVALID_ARGUMENT; }
    int rc = g->wait_all();
    if (rc != RB200_OK) return rc;
    const uint32_t total = (uint32_t)g->batchesRendered;
    const size_t n = (size_t)g->width * g->height;
    if (g->peerTiles) {
        // device 0's image is complete once every device has folded its batches: drain them all, then read it
        for (int i = 0; i < g->n(); i++) g->workers[i]->post([g, i] { return rb200_synchronize(g->ctx[i]); });
        if ((rc = g->wait_all()) != RB200_OK) return rc;
        g->workers[0]->post([g, n, rgba32f]() -> int {
            RB200Context* c = g->ctx[0];
            RB_CUDA(cudaSetDevice(c->device));
            RB_CUDA(cudaMemcpyAsync(rgba32f, c->wp.image, n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
            RB_CUDA(cudaStreamSynchronize(c->stream));
            return RB200_OK;
        });
        return g->workers[0]->wait();
    }
    for (int i = 0; i < g->n(); i++)
        g->workers[i]->post([g, i, total, n, rgba32f]() -> int {
            RB200Context* c = g->ctx[i];
            const float4* src = c->wp.image;
            if (g->n() > 1) {
                CommState* cs = comm_of(c);
                const int r = reduce_image(c, cs, 0);
                if (r != RB200_OK || i != 0) return r;
                src = cs->reduced;
            }
            RB_CUDA(cudaSetDevice(c->device));
            RB_CUDA(cudaMemcpyAsync(rgba32f, src, n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
            RB_CUDA(cudaStreamSynchronize(c->stream));
            if (g->tiles) return RB200_OK;                  // running averages of disjoint tiles: nothing to resolve
            const float inv = 1.0f / (float)total;          // the arithmetic of rb200_resolve_sum
            for (size_t k = 0; k < n; k++) { rgba32f[4 * k] *= inv; rgba32f[4 * k + 1] *= inv; rgba32f[4 * k + 2] *= inv; rgba32f[4 * k + 3] = 1.0f; }
            return RB200_OK;
        });
    return g->wait_all();
}

RB200_API int rb200_group_get_stats(RB200Group* g, RB200Stats* cumulative) {
    if (!g || !cumulative) { set_error("null argument"); return RB200_ERR_INVALID_ARGUMENT; }
    int rc = g->wait_all();
    if (rc != RB200_OK) return rc;
    std::vector<RB200Stats> st(g->n());
    for (int i = 0; i < g->n(); i++) g->workers[i]->post([g, &st, i] { RB200Stats last; return rb200_get_stats(g->ctx[i], &last, &st[i]); });
    rc = g->wait_all();
    if (rc != RB200_OK) return rc;
    memset(cumulative, 0, sizeof(*cumulative));
    for (const RB200Stats& s : st) {
        cumulative->extendRays += s.extendRays; cumulative->shadowRays += s.shadowRays; cumulative->paths += s.paths;
        cumulative->nodeVisits += s.nodeVisits; cumulative->triTests += s.triTests; cumulative->waves += s.waves;
        cumulative->kernelLaunches += s.kernelLaunches; cumulative->shadowNodeVisits += s.shadowNodeVisits; cumulative->shadowTriTests += s.shadowTriTests;
        cumulative->shadowRaysSkipped += s.shadowRaysSkipped;
    }
    return RB200_OK;
}

} // extern "C"
