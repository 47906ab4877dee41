/*
 * ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product.
 *
 * CPU restatement of the path AlexanderJCS/reina-vk runs on the GPU (trace + shade + accumulate + bloom +
 * tonemap). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the CUDA library never links or calls it.
 *
 * PARITY PINNING: the reference ships no tests, golden vectors or fixtures and cannot be
 * built here (Vulkan RT + glslc + Assimp absent), and ray/triangle intersection lives in the Vulkan driver
 * (traversal: UNPINNED). What the reference does ship is the compiled SPIR-V of every shader; tests/spirv_interp.py
 * executes those binaries on the CPU and tests/golden/spirv_*.npz holds their outputs: the parts of this
 * restatement listed in tests/test_spirv_golden.py (post-processing; camera, RNG, bounce loop, the four materials,
 * parallax mapping, sky and accumulation of the as-shipped estimator) are pinned to the reference's compiled code bit
 * for bit. Next-event estimation is compiled out of those binaries: nee.h.glsl, directLight and the bounce loop with
 * NEE on are pinned by independent numpy / Python restatements (tests/test_oracle_kat.py, bit-exact). The rest is
 * defended line-by-line against the GLSL (every function cites the lines it follows) and pinned by the known-answer
 * values in SURVEY.md Appendix A3 (tests/test_oracle_kat.py).
 *
 * Elementary layer: rb_math.h / rb_vec.h / rb_tri.h (sin, cos, log, exp, acos, vector built-ins, the
 * watertight triangle test) are shared with the kernels on purpose, so that both sides round identically and
 * the parity tests can demand bit-equal images. Everything above that layer (camera, bounce loop, materials,
 * Disney BSDF, light sampling, accumulation, bloom, tonemap, BVH) is written independently here.
 */
#ifndef ORACLE_COMMON_H
#define ORACLE_COMMON_H

#include <cstdint>
#include <cstddef>
#include <vector>
#include "../include/reina_b200.h"
#include "../reina-vk_b200/csrc/rb_math.h"
#include "../reina-vk_b200/csrc/rb_vec.h"
#include "../reina-vk_b200/csrc/rb_tri.h"

namespace oracle {

typedef rb_v3 vec3;
typedef rb_v2 vec2;
typedef rb_m3 mat3;

struct WorldTri {
    vec3 v0, v1, v2;
    uint32_t instance;   // index into instances[]
    uint32_t primitive;  // triangle index within the model (gl_PrimitiveID)
};

struct BvhNode {          // plain binary BVH used only by the oracle
    float lo[3], hi[3];
    uint32_t left;        // internal: left child, right = left + 1 ; leaf: first triangle slot
    uint32_t count;       // 0 = internal, else number of triangles
};

struct Hit {
    float t, b1, b2;
    uint32_t gid;         // global primitive id = position in the flattened (instance-major) triangle list
    bool valid;
};

struct Scene {
    std::vector<float> vertices;            // float4
    std::vector<uint32_t> indices;
    std::vector<RB200InstanceProperties> props;
    std::vector<float> tbns;                // 9 floats each
    std::vector<uint32_t> tbnIndices;
    std::vector<RB200InstanceData> emissive;
    std::vector<float> cdfTriangles;
    std::vector<float> cdfInstances;
    std::vector<float> texCoords;           // float2
    std::vector<uint32_t> texIndices;
    struct Tex { std::vector<uint8_t> rgba; uint32_t w, h; };
    std::vector<Tex> textures;
    std::vector<RB200Instance> instances;

    std::vector<WorldTri> tris;             // flattened, instance-major: gid = index
    std::vector<BvhNode> nodes;
    std::vector<uint32_t> bvhTriOrder;      // leaf slots -> gid
    bool useBvh = false;

    // Two-level instancing (src/scene/Scene.cpp:93-111: one BLAS per object, a transform per instance): rays are moved into
    // object space with the inverse instance transform and intersected with the object's triangles there, as the Vulkan
    // driver does with its TLAS / BLAS. `blas[b]` holds the object-space triangles (and hierarchy) of one distinct model
    // range, `tl[i]` what instance i needs. `tris` above stays filled (gid -> instance, primitive).
    struct TwoLevelInstance { float inv[12]; uint32_t blas; uint32_t gidBase; };
    bool twoLevel = false;
    std::vector<Scene> blas;
    std::vector<TwoLevelInstance> tl;
};

// intersect.cpp
void flatten_scene(Scene& s);
void build_bvh(Scene& s);
bool build_two_level(Scene& s, bool withBvh);      // false: an instance transform is singular
Hit closest_hit(const Scene& s, vec3 org, vec3 dir, float tmax, bool brute, uint64_t* tri_tests = nullptr);
bool any_hit(const Scene& s, vec3 org, vec3 dir, float tmax, bool brute);

// PCG step of shaders/raytrace/shaderCommon.h.glsl:39-45, written out again for the oracle (the kernels use
// rb_random from rb_vec.h; tests/test_oracle_kat.py checks both against SURVEY.md A3's known answers).
// float(word) rounds to nearest; 4294967295.0f is 2^32 in fp32, so the result lies in [0, 1] inclusive.
static inline float rnd(uint32_t& state) {
    state = state * 747796405u + 1u;
    uint32_t word = ((state >> ((state >> 28) + 4u)) ^ state) * 277803737u;
    word = (word >> 22) ^ word;
    return (float)word / 4294967295.0f;
}

struct Counters { uint64_t extendRays = 0, shadowRays = 0, paths = 0; };

// pathtrace.cpp
struct TilePartition { uint32_t rank = 0, count = 1, size = 32; };
void render_rows(const Scene& s, uint32_t W, uint32_t H, uint32_t flags, const RB200RtPushConsts& pc,
                 float* hdr, uint32_t y0, uint32_t y1, uint32_t ystep, Counters* counters,
                 const TilePartition& tiles = TilePartition());

} // namespace oracle
#endif
