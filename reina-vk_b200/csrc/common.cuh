// common.cuh — shared declarations of the CUDA library (librb200.so).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/reina_b200.h"
#include "rb_math.h"
#include "rb_vec.h"
#include "rb_tri.h"

namespace rb200 {

void set_error(const char* fmt, ...);

#define RB_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            rb200::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return RB200_ERR_CUDA;                                                                 \
        }                                                                                          \
    } while (0)

// ---------------------------------------------------------------------------------------------------
// 8-wide compressed BVH (Ylitie, Karras, Laine 2017 layout): 80-byte nodes, 48-byte triangles
//
// RB_WIDE_LOADS=1: records padded to a multiple of 32 bytes (node 96 B, triangle 64 B, 32-byte aligned) and read with
// 256-bit loads (LDG.E.256, sm_100): a wide node is 3 load instructions instead of 5, a triangle 2 instead of 3. Every
// lane reads its own node, so each load instruction of a warp is up to 32 separate L1 requests, and the L1 data pipe was
// the traversal kernels' busiest unit next to the issue slots (76 % of its peak, profiles/r02b). MEASURED ON B200: no
// gain — k_extend 19.63 against 19.65 ms per step, k_shadow 10.50 against 10.37 (profiles/r02e_variant_sweep.txt): the
// five 128-bit loads of a node already hit the one or two lines the first of them brought in. Off: the records keep
// their 80 / 48 bytes (the padding would also cost 15.5 MB of L2 on the headline scene).
// (RB_WIDE_RECORDS below — the same idea for the shading kernels, which ARE bound by the request rate — pays.)
// ---------------------------------------------------------------------------------------------------
#ifndef RB_WIDE_LOADS
#define RB_WIDE_LOADS 0
#endif
// RB_WIDE_RECORDS=1: the per-triangle shading records (64 and 128 bytes, 64-byte aligned) are read with 256-bit loads.
#ifndef RB_WIDE_RECORDS
#define RB_WIDE_RECORDS 1
#endif
#if RB_WIDE_LOADS
#define RB_BVH_ALIGN 32
#else
#define RB_BVH_ALIGN 16
#endif
static constexpr uint32_t WIDE_NODE_INFO_BYTES = 80, TRI_RECORD_INFO_BYTES = 48;     // what a record carries, without padding
struct alignas(RB_BVH_ALIGN) WideNode {
    float px, py, pz;        // origin of the quantisation grid = node AABB min
    uint8_t ex, ey, ez;      // biased fp32 exponents of the per-axis grid step
    uint8_t imask;           // bit s set: slot s holds an internal child
    uint32_t childBase;      // index of the first internal child (children are contiguous, in slot order)
    uint32_t triBase;        // index of the first triangle referenced by this node's leaf slots
    uint8_t meta[8];         // per slot: internal -> (1<<5) | (24+slot); leaf -> (unary count << 5) | tri offset; 0 = empty
    uint8_t qlox[8], qloy[8];
    uint8_t qloz[8], qhix[8];
    uint8_t qhiy[8], qhiz[8];
#if RB_WIDE_LOADS
    uint32_t pad[4];         // zero
#endif
};
static_assert(sizeof(WideNode) == (RB_WIDE_LOADS ? 96 : 80), "WideNode must be 80 bytes (+ 16 of padding with RB_WIDE_LOADS)");

// triangle record in leaf order: three float4, w lanes carry ids
//   v0.w = primitive id within the model, v1.w = instance index | material kernel (0..3) << 30, v2.w = global primitive
//   id (tie-break key). The material bits let k_extend bin a finished ray without touching the instance table.
static constexpr uint32_t TRI_INST_MASK = 0x3FFFFFFFu;
struct alignas(RB_BVH_ALIGN) TriRecord {
    float4 v0, v1, v2;
#if RB_WIDE_LOADS
    float4 pad;              // zero
#endif
};
static_assert(sizeof(TriRecord) == (RB_WIDE_LOADS ? 64 : 48), "TriRecord must be 48 bytes (+ 16 of padding with RB_WIDE_LOADS)");

struct Bvh {
    void* blob = nullptr;            // one allocation: nodes, then (256-byte aligned) triangles
    size_t blobBytes = 0;
    WideNode* nodes = nullptr;
    TriRecord* tris = nullptr;
    uint32_t numNodes = 0, numTris = 0, maxDepth = 0;
    float buildMs = 0.f;
    float sceneMin[3] = {0, 0, 0}, sceneMax[3] = {0, 0, 0};
};

// Scene tables resident in HBM (bindings 2..11, 13 of the reference + the BVH replacing binding 1)
struct DeviceScene {
    const float4* vertices;
    const uint32_t* indices;
    const RB200InstanceProperties* props;
    const float* tbns;               // 9 floats per entry
    const uint32_t* tbnIndices;
    const RB200InstanceData* emissive;
    const float* cdfTriangles;
    const float* cdfInstances;
    uint32_t numCdfInstances;
    const float2* texCoords;
    const uint32_t* texIndices;
    const RB200Instance* instances;
    const cudaTextureObject_t* textures;
    const uint2* texSizes;
    const WideNode* nodes;
    const TriRecord* tris;
    uint32_t numInstances, numTris;
    uint32_t materialMask;           // bit m: some instance is shaded by material kernel m (0..3); host-side launch filter
    uint32_t pad_;
    // Per-triangle shading records in leaf order (index = TriRecord index), built once at scene creation: the floats
    // getObjectHitInfo gathers through three levels of indirection (instance -> properties -> index tables -> vertex /
    // frame / uv tables), laid out contiguously. shadeBase: 4 float4 per triangle = object-space v0.xyz|uv0.x,
    // v1.xyz|uv0.y, v2.xyz|uv1.x, (uv1.y, uv2.x, uv2.y, 0). shadeFrame: 8 float4 = n0.xyz|t0.x, n1.xyz|t0.y,
    // n2.xyz|t0.z, t1.xyz|b0.x, t2.xyz|b0.y, b1.xyz|b0.z, b2.xyz|0, pad (n / t / b = columns 2 / 0 / 1 of the vertex
    // TBNs). Null when RB_SHADE_RECORDS is 0.
    const float4* shadeBase;
    const float4* shadeFrame;
    // InstanceProperties by INSTANCE index (a copy of props[instances[i].instancePropertiesID]): the material kernels
    // fetch the instance record and its properties side by side instead of one after the other. Null when
    // RB_INST_RECORDS is 0.
    const RB200InstanceProperties* instProps;
    // Two-level mode (RB200_FLAG_TWO_LEVEL; null / 0 otherwise). `nodes` / `tris` then hold the object-space hierarchies of
    // the distinct model ranges, one after the other (child and triangle indices already absolute), so a triangle slot
    // still indexes shadeBase / shadeFrame; tlasNodes / tlasLeaves are a hierarchy over the instances' world-space boxes
    // whose "triangles" are box records (v0 = min, v1 = max, v0.w = instance index); tlInstances[i] is what a ray needs
    // to enter instance i.
    const WideNode* tlasNodes;
    const TriRecord* tlasLeaves;
    const struct TwoLevelInstance* tlInstances;
};

// Entry record of an instance in two-level mode (64 bytes)
struct alignas(16) TwoLevelInstance {
    float inv[12];           // inverse of the instance transform, three rows of four (rb_inv_point / rb_inv_vector)
    uint32_t rootNode;       // index of the root wide node of the instance's model in DeviceScene::nodes
    uint32_t gidBase;        // global primitive id of the instance's first triangle (instance-major order, as flattening numbers them)
    uint32_t material;       // material kernel 0..3 (as k_extend bins)
    uint32_t pad;
};
static_assert(sizeof(TwoLevelInstance) == 64, "TwoLevelInstance must be 64 bytes");
// per-lane stack of the two-level traversal (two_level.cuh): node groups and instance groups of both levels + one sentinel;
// rb200_scene_create refuses hierarchies with 2 * depth(top level) + depth(objects) + 2 beyond it
static constexpr int TL_STACK = 64;

#ifndef RB_SHADE_RECORDS
#define RB_SHADE_RECORDS 1
#endif
#ifndef RB_INST_RECORDS
#define RB_INST_RECORDS 1
#endif

struct BuildInput {
    const float4* vertices;
    const uint32_t* indices;
    const RB200Instance* d_instances;
    const std::vector<RB200Instance>* h_instances;
    uint32_t numInstances;
    int builder;       // BUILDER_*
    // non-null: build over these `numPremade` records (device memory) instead of flattening the instances — the top level of
    // the two-level mode, whose leaves are instance boxes
    const TriRecord* premade = nullptr;
    uint32_t numPremade = 0;
};
// Binary hierarchy under the wide-node collapse: Karras 2012 over the sorted Morton keys (the north-star pipeline) or
// PLOC clustering of the same sorted leaves (better surface-area cost, a few more milliseconds of build).
enum { BUILDER_LBVH = 0, BUILDER_PLOC = 1 };

// bvh_build.cu
int build_bvh(const BuildInput& in, cudaStream_t stream, Bvh* out, uint64_t* launches);
int build_two_level(const BuildInput& in, cudaStream_t stream, Bvh* blas, Bvh* tlas, std::vector<TwoLevelInstance>* entries,
                    uint32_t* numModels, uint64_t* launches);
int hash_bvh(const Bvh& bvh, cudaStream_t stream, uint64_t* hash);
void free_bvh(Bvh* b);

} // namespace rb200
