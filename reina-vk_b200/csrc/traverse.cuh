// traverse.cuh — stack-based traversal of the 8-wide compressed BVH (closest hit and any hit).
//
// Software replacement for the fixed-function traversal behind traceRayEXT in the reference
// (shaders/raytrace/raytrace.rgen.glsl:110-122: opaque, tmin 0, tmax 1e4, closest hit;
//  shaders/raytrace/nee.h.glsl:126-144: opaque | terminateOnFirstHit | skipClosestHit, tmax dist - 0.001).
//
// One ray per lane. A stack entry is a "group": (base index, bit mask). Node groups carry the hit mask of the
// internal children of one wide node in bits 24..31 (already permuted by the ray octant so that the highest set bit
// is the nearest child) and the node's imask in bits 0..7; triangle groups carry up to 24 triangle bits.
// Child boxes are tested directly in the quantised grid: t = q * (2^e / d) + (p - o) / d, one fma per plane.
// Every plane is pushed outwards by a slack that bounds the fp32 rounding of that expression and the few-ulp
// acceptance band of the watertight triangle test, so a triangle the test accepts is never culled by a box.
// Closest-hit rule: smallest t, ties -> smallest global primitive id; boxes are culled with <= so ties are visited
// and the result does not depend on the order in which nodes and triangles are processed.
//
// Warp organisation (trace_queue): all 32 lanes run one convergent loop.
//   1. refill   idle lanes pull the next rays from the queue with one aggregated atomic, so lanes whose rays finish
//               early do not wait for the slowest ray of the warp;
//   2. nodes    every lane with a ray performs RB_CHUNK wide-node steps (~23 of 32 lanes active) and queues the
//               triangle groups they produce;
//   3. triangles the warp pools ALL queued triangles of its 32 rays in shared memory and tests them 32 at a time,
//               any lane testing any ray's triangle (ray data is staged in shared memory). A wide node yields ~0.5
//               triangles, so per-lane triangle loops ran with 3-6 active lanes and were > 1/3 of all issued
//               instructions; pooled, a chunk's ~90 triangles take 3 rounds instead of ~16 iterations.
//               Closest hits are merged per ray with a 64-bit shared-memory atomicMin on (t bits << 32 | primitive id),
//               which is exactly the closest-hit rule above.
#pragma once
#include "common.cuh"

namespace rb200 {

struct RayHit {
    float t, b1, b2;
    uint32_t tri;      // index into Bvh::tris, 0xFFFFFFFF = miss
    uint32_t gid;
};

static constexpr int TRAV_STACK = 24;          // pending node groups: at most one per tree level
static constexpr uint32_t TRAV_MAX_DEPTH = 22;

#ifndef RB_REFILL
#define RB_REFILL 32      // refill when fewer than this many lanes own a ray
#endif
#ifndef RB_CHUNK
#define RB_CHUNK 6        // wide-node steps a lane performs between two triangle phases (swept 2..24 on B200)
#endif
#ifndef RB_PREFETCH
#define RB_PREFETCH 0      // L1 prefetch of the next wide node / first queued triangle at the end of a node step:
                           // measured -5 % on B200 (extend 1426 vs 1504 Mrays/s), kept only as a switch
#endif
#ifndef RB_WORK_CAP
#define RB_WORK_CAP 256   // pooled triangles per pass (a chunk produces ~90 per warp; more are handled by extra passes).
                          // 64 / 96 / 128 / 256 entries: -4 / -1 / 0 / +1 % on B200 with 4-byte entries
#endif

// RB_COOP_FILL=1: warp-cooperative fill of the pooled work list (see trace_queue): position p of the list is computed by
// lane p % 32 from the lanes' group lists in shared memory, total / 32 rounds with all lanes instead of the per-lane loop's
// ~17 iterations with 5.6 active lanes (11 % of the kernel's issued instructions, profiles/r02f SASS page). MEASURED ON
// B200: SLOWER — k_extend 19.50 -> 20.13 ms, k_shadow 10.31 -> 10.75 ms per step (profiles/r02j_variant_sweep.txt), results
// identical: the owner search, the walk over the owner's groups and the n-th-set-bit loop are chains of dependent
// shared-memory loads, and the few issue slots the narrow loop wastes are cheaper than those stalls. Off.
#ifndef RB_LEAN_FILL
#define RB_LEAN_FILL 1       // nested-loop fill of the pooled work list when a chunk's triangles fit (see trace_queue)
#endif

// RB_DEFER_BARY=1 (closest hit): a candidate's barycentric divisions V / det, W / det are postponed to the commit of the
// ray. The warp's payload row of a ray keeps (V, W, triangle, det) of its current best — it is only ever rewritten by a
// better candidate of the same ray, so it survives from pass to pass — and the lane keeps just (t, primitive id) in
// registers. Two IEEE divisions (~30 instructions) leave every round of the pooled triangle phase, three registers
// (b1, b2, tri) leave the lane's persistent state (k_extend then fits 64 registers without spilling); the quotients are
// the same operands divided once, later. MEASURED ON B200 (profiles/r02m_variant_sweep.txt), bit-identical (46 parity
// tests): SLOWER — k_extend 18.17 -> 18.86 ms per step at 7 x 128 threads, 18.18 at 8 x 128 / 4 x 256 with 64 registers:
// the divisions of a round are already skipped when no lane's line crosses its triangle, and the commit — which runs
// with ~6 lanes — now carries two divisions and a shared-memory read. Off.
#ifndef RB_DEFER_BARY
#define RB_DEFER_BARY 0
#endif
// RB_FAST_RCP=1: 1 / d for the BOX tests from MUFU.RCP (1 ulp) instead of the IEEE sequence (~12 instructions each): the
// slack of node_step (2^-20 relative on every plane) covers a 2^-23 relative change of all planes of an axis; the
// triangle test's shear constants stay IEEE. B200: k_extend 18.20 -> 18.06 ms, k_shadow 9.61 -> 9.53 ms per step,
// results bit-identical (46 parity tests). On.
#ifndef RB_FAST_RCP
#define RB_FAST_RCP 1
#endif
// RB_DOM_AXIS=1: the ray's dominant axis for the culling margin (node_step) is found once per ray and kept in bits 8..9
// of oct_inv instead of three compares per node step. B200: SLOWER, k_extend 18.86 -> 19.24 ms (the extra live bits cost
// more in the node step's register allocation than the compares). Off.
#ifndef RB_DOM_AXIS
#define RB_DOM_AXIS 0
#endif

// RB_PRECLAIM=1: a lane whose ray has no node work left will be idle after the coming triangle phase, whatever the tests
// say; it claims its NEXT ray before that phase (the aggregated atomic), reads the queue entry while the work list is
// filled and requests the ray's record into L2 / L1 while the triangles are tested — so the refill's chain of three
// dependent long-latency accesses (atomic -> queue entry -> ray record: a tenth of the kernel's stall samples,
// profiles/r02a_k_extend_source_lines.csv, with the whole warp waiting) is off the critical path. Which lane traces which
// ray changes; results do not depend on that. MEASURED ON B200 (profiles/r02p_variant_sweep.txt), bit-identical (46 parity
// tests): SLOWER — k_extend 17.68 -> 17.91 ms, k_shadow 9.44 -> 9.68 ms per step: the second ballot / atomic / shuffle
// sequence and the prefetches are issued by the same few lanes, and the other warps of the SM were already covering that
// latency. Off.
#ifndef RB_PRECLAIM
#define RB_PRECLAIM 0
#endif

#ifndef RB_FILL_UNROLL2
#define RB_FILL_UNROLL2 0    // the nested-loop fill writes two triangles per trip (see trace_queue). B200: k_extend 17.69 ->
                             // 17.86 ms (profiles/r02q_variant_sweep.txt), same checksums. Off.
#endif
// RB_BARY_LATE=1 (closest hit): the barycentric divisions V / det, W / det of a crossing are made only after its t passed
// the (0, tmax) test — same operands, same quotients; a round in which no crossing is in range skips them. B200: k_extend
// 17.69 -> 18.58 ms, k_shadow 9.47 -> 9.91 ms, same checksums (the split form of the test schedules worse than the
// divisions cost, as with RB_DEFER_BARY). Off.
#ifndef RB_BARY_LATE
#define RB_BARY_LATE 0
#endif

#ifndef RB_STACK_IN_STRUCT
#define RB_STACK_IN_STRUCT 0    // 1: the round-1 layout (deep stack array as a member of Traversal), kept for comparison
#endif

#ifndef RB_COOP_FILL
#define RB_COOP_FILL 0
#endif

#ifndef RB_TRI_LDCG
#define RB_TRI_LDCG 0     // 1: triangle records bypass L1 (ld.global.cg): measured -1 % closest-hit, -5 % any-hit on B200
                          // (neighbouring rays do re-use each other's triangles); ray records through L2 only: no change
#endif

// Where a lane keeps (a) the triangle groups its node steps produce during a chunk (at most RB_CHUNK) and (b) its stack
// of pending node groups. In local memory, lanes at different fill levels touch different rows — one L1 wavefront per
// lane and access; local memory was half of the traversal kernels' L1 data-pipe wavefronts (the busiest unit, 67 % of its
// peak) and its lines compete with nodes and triangles for L1. In shared memory as [entry][lane] an access is
// conflict-free (2 wavefronts per warp), at 256 B per entry and warp taken from L1.
// Measured on B200 (headline scene, ms per batch, 4 lanes; Mrays/s of the whole step):
//   everything local                                        extend 33.5  shadow 18.2   1962
//   (a) shared, any-hit only                                       33.4         17.0   1972
//   (a) shared any-hit + first 4 of (b) shared any-hit             33.5         15.7   1986
//   ... + first 4 / 6 of (b) shared closest-hit, (a) local         33.8 / 33.3  15.7   2003 / 2025
//   (a) and first 4 of (b) shared in both kernels (default)        29.4         15.7   2232
//   ... first 6 of (b) in closest-hit (50.6 KB per block)          30.1         15.6   2090
//   first 6 / 8 of (b) any-hit (> 48 KB per block)                 33.3         16.7   1835
// Either half alone does nothing for the closest-hit kernel (the other half keeps L1 busy); blocks beyond ~47 KB leave
// no shared memory for the other lanes' kernels and the step slows down although the kernel itself does not.
// RB_TSTACK_SHARED: bit 0 any-hit kernels, bit 1 closest-hit kernels.
#ifndef RB_TSTACK_SHARED
#define RB_TSTACK_SHARED 3
#endif
template <bool ANY> struct TStackShared { static constexpr bool value = ((RB_TSTACK_SHARED >> (ANY ? 0 : 1)) & 1) != 0; };
// The first RB_STACK_SHARED_* entries of the node-group stack live in shared memory, deeper ones in local memory.
// With 128-thread blocks of k_extend (7 per SM) and no payload rows in the any-hit stage, the blocks of either kernel need
// 160-172 KB per SM, and the carve-out the driver picks for that (196 KB) has room for deeper shared stacks at no cost in
// L1 (B200, headline step, profiles/r02n_variant_sweep.txt): closest hit 4 / 5 / 6 / 7 entries: k_extend 18.07 / 17.95 /
// 17.68 / 17.63 ms; any hit 4 / 5 / 6: k_shadow 9.49 / 9.42 / 9.54 ms. (Forcing the 228 KB carve-out instead — 28 KB of L1 —
// costs k_extend 3 % and k_shadow 9 %; a 5-entry triangle-group list that would fit the 164 KB carve-out costs 10 %.)
#ifndef RB_STACK_SHARED_ANY
#define RB_STACK_SHARED_ANY 5
#endif
#ifndef RB_STACK_SHARED_CLOSEST
#define RB_STACK_SHARED_CLOSEST 6
#endif
template <bool ANY> struct StackShared { static constexpr int value = ANY ? RB_STACK_SHARED_ANY : RB_STACK_SHARED_CLOSEST; };

// RB_BVH_EVICT_LAST=1: node and triangle loads carry an L2 evict_last policy so that the per-wave sweep of path state
// (~0.6 GB at 1080p, against 48 MB of hierarchy) evicts path state first (r01i: L2 hit rate 74 %, DRAM traffic of a
// launch 4.9x its algorithmic HBM bytes).
#ifndef RB_BVH_EVICT_LAST
#define RB_BVH_EVICT_LAST 0
#endif
#if RB_BVH_EVICT_LAST
__device__ __forceinline__ unsigned long long bvh_policy() {
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ld_bvh(const float4* p, unsigned long long pol) {
    float4 v;
    asm("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
#else
__device__ __forceinline__ unsigned long long bvh_policy() { return 0ull; }
__device__ __forceinline__ float4 ld_bvh(const float4* p, unsigned long long) { return __ldg(p); }
#endif
#if RB_WIDE_LOADS
// two consecutive float4 of a 32-byte aligned record with one 256-bit read-only load (LDG.E.256.CONSTANT)
__device__ __forceinline__ void ld_bvh2(const float4* p, float4& a, float4& b, unsigned long long) {
    asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
#endif

// The first RB_TSTACK_N triangle groups a lane queues during a chunk live in shared memory, the rest (a lane queues
// more than 4 groups in 2 % of its chunks) in local memory: with RB_WORK_CAP 128 a block then needs 38 KB instead of
// 46 KB, 4 blocks fit under the 164 KB carve-out and the SM keeps 64 KB of L1 instead of 32 KB.
#ifndef RB_TSTACK_N
#define RB_TSTACK_N RB_CHUNK
#endif

// Chunk length of the any-hit kernels (<= RB_CHUNK, which sizes the triangle-group lists): a shadow ray ends with its
// first hit, so testing the queued triangles sooner saves node steps that a found hit makes useless. B200, headline
// step, k_shadow ms: 3 / 4 / 5 / 6 steps 10.55 / 10.35 / 10.36 / 10.55.
#ifndef RB_CHUNK_ANY
#define RB_CHUNK_ANY 4
#endif
static_assert(RB_CHUNK_ANY <= RB_CHUNK, "the triangle-group lists hold RB_CHUNK entries");

// per-warp staging area of the pooled triangle phase
template <bool TSTACK, int N> struct WarpTStack { };
template <int N> struct WarpTStack<true, N> { uint2 tstack[N][32]; };    // per lane: triangle groups of the current chunk
// a chunk of the any-hit kernels is RB_CHUNK_ANY node steps, so their lists need no more entries than that
template <bool ANY> struct TStackEntries { static constexpr int value = (ANY && RB_TSTACK_N == RB_CHUNK) ? RB_CHUNK_ANY : RB_TSTACK_N; };
template <int K> struct WarpStack { uint2 nstack[K][32]; };
template <> struct WarpStack<0> { };
// per owner: b1, b2, bits(triangle index) of the current best — closest hit only (an any-hit ray has no use for it, and
// without these 512 bytes four 256-thread blocks of k_shadow need 160 KB instead of 176 KB of the SM's shared memory)
template <bool ANY> struct WarpPayload { float4 payload[32]; };
template <> struct WarpPayload<true> { };
template <bool ANY>
struct WarpShared : WarpTStack<TStackShared<ANY>::value, TStackEntries<ANY>::value>, WarpStack<StackShared<ANY>::value>, WarpPayload<ANY> {
    float4 ray[32][3];                    // per lane: (o, tmax), (mx, Sz), (my, bits(kz)) — shear rows of rb_tri.h
    uint32_t work[RB_WORK_CAP];           // triangle index << 5 | owner lane (the build refuses >= 2^27 triangles)
    unsigned long long bestKey[32];       // per owner: min over candidates of (t bits << 32 | global primitive id)
};

// per byte: 0xFF if bit 7 is set, else 0x00 (prmt's sign-replicate mode; __byte_perm only honours 3 selector bits)
__device__ __forceinline__ uint32_t sign_extend_s8x4(uint32_t x) {
    uint32_t r;
    asm("prmt.b32 %0, %1, 0, 0x0000ba98;" : "=r"(r) : "r"(x));
    return r;
}
// byte j of w as the float 2^15 + byte, built by ONE PRMT (0x4700bb00): no conversion-pipe (XU) instruction and no
// subtraction — the bias is folded into the plane offsets (c' = c - 2^15 * s), whose extra rounding (<= 2^-9 of a
// grid step) is covered by the slack below.
static constexpr float BYTE_BIAS = 32768.0f;
// The 0x47000000 word comes from constant memory on purpose: with both PRMT inputs known at compile time ptxas keeps
// the WORD as the immediate and the four selectors in registers, and under the 64-register cap it re-materialises
// those selectors with an extra move in front of most of the 48 PRMTs of a node step. A run-time word leaves the
// selector as the immediate: one register, no moves.
__constant__ uint32_t c_byteBiasWord = 0x47000000u;
__device__ __forceinline__ float byte_f(uint32_t w, int j, uint32_t biasWord) {
    return __uint_as_float(__byte_perm(w, biasWord, 0x7404u + ((uint32_t)j << 4)));
}

// Culling margin as a fraction of the node's extent along the ray's dominant axis (see node_step): 2^-14.
#ifndef RB_CULL_MARGIN
#define RB_CULL_MARGIN 6.103515625e-05f
#endif

// RB_ORIGIN_FROM_STAGE=1: the node step reads the ray origin from the warp's shared-memory ray stage (one LDS.128, the
// stage holds it for the triangle phase anyway) instead of keeping it in three registers: under the 64-register cap ptxas
// spilled exactly those three and re-loaded them from local memory in every node step (3 LDL per step, r02a).
#ifndef RB_ORIGIN_FROM_STAGE
#define RB_ORIGIN_FROM_STAGE 1
#endif

__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

template <bool ANY, bool COUNT>
struct Traversal {
#if !RB_ORIGIN_FROM_STAGE
    rb_v3 o;
#endif
    float idx, idy, idz;
    uint32_t oct_inv;
    __device__ __forceinline__ uint32_t oct3() const {
#if RB_DOM_AXIS
        return oct_inv & 7u;
#else
        return oct_inv;
#endif
    }
    uint2 ngroup, tgroup;
    int sp, tsp;
    uint32_t tcount;              // triangles queued in tgroup + tstack
    RayHit best;
#if RB_STACK_IN_STRUCT
    uint2 stack[TRAV_STACK];
#else
    uint2* stack;                 // pending node groups below the shared-memory entries: an array in local memory that
                                  // lives OUTSIDE this struct, so that the scalar members stay in registers (with the array
                                  // inside, ptxas kept ngroup / sp / tsp / tcount in the local frame: ~10 STL / LDL per node step)
#endif
    // triangle groups produced by the node steps of the current chunk (see RB_TSTACK_SHARED)
    uint2 tstackLocal[TStackShared<ANY>::value ? (RB_TSTACK_N < RB_CHUNK ? RB_CHUNK - RB_TSTACK_N : 1) : RB_CHUNK];
    __device__ __forceinline__ uint2& tst(WarpShared<ANY>& ws, int k) {
        if constexpr (TStackShared<ANY>::value) {
            if constexpr (RB_TSTACK_N < RB_CHUNK) { if (k >= RB_TSTACK_N) return tstackLocal[k - RB_TSTACK_N]; }
            return ws.tstack[k][threadIdx.x & 31u];
        }
        else return tstackLocal[k];
    }

    // History: with the triangle-group list still in local memory, putting the first 6 / 8 / 10 entries of this stack in
    // shared memory measured -3 / -8 / -8 % (closest hit) and a register-cached top entry -9 %; together with a shared
    // triangle-group list the first 4 entries in shared memory give +17 % (see RB_TSTACK_SHARED above).
    __device__ __forceinline__ void push(const uint2 v, WarpShared<ANY>& ws) {
        constexpr int K = StackShared<ANY>::value;
        if constexpr (K > 0) {
            if (sp < K) ws.nstack[sp][threadIdx.x & 31u] = v; else stack[sp - K] = v;
            sp++;
        } else stack[sp++] = v;
    }

    __device__ __forceinline__ void init(const rb_v3 org, const rb_v3 d, const float tmax_, float4* rayStage) {
#if !RB_ORIGIN_FROM_STAGE
        o = org;
#endif
        best.t = tmax_; best.gid = 0xFFFFFFFFu;
        if (ANY || !RB_DEFER_BARY) { best.b1 = 0.f; best.b2 = 0.f; best.tri = 0xFFFFFFFFu; }
        const float ooeps = 8.2718061e-25f;   // 2^-80: keeps 1/d finite for axis-parallel rays (box tests only)
#if RB_FAST_RCP
        idx = fast_rcp(fabsf(d.x) > ooeps ? d.x : copysignf(ooeps, d.x));
        idy = fast_rcp(fabsf(d.y) > ooeps ? d.y : copysignf(ooeps, d.y));
        idz = fast_rcp(fabsf(d.z) > ooeps ? d.z : copysignf(ooeps, d.z));
#else
        idx = 1.0f / (fabsf(d.x) > ooeps ? d.x : copysignf(ooeps, d.x));
        idy = 1.0f / (fabsf(d.y) > ooeps ? d.y : copysignf(ooeps, d.y));
        idz = 1.0f / (fabsf(d.z) > ooeps ? d.z : copysignf(ooeps, d.z));
#endif
        oct_inv = (idx < 0.f ? 0u : 4u) | (idy < 0.f ? 0u : 2u) | (idz < 0.f ? 0u : 1u);
#if RB_DOM_AXIS
        {
            const float ax = fabsf(idx), ay = fabsf(idy), az = fabsf(idz);
            oct_inv |= ((ax <= ay && ax <= az) ? 0u : (ay <= az ? 1u : 2u)) << 8;
        }
#endif
        const rb_ray_shear sh = rb_ray_prepare(d);
        rayStage[0] = make_float4(org.x, org.y, org.z, tmax_);
        rayStage[1] = make_float4(sh.mx.x, sh.mx.y, sh.mx.z, sh.Sz);
        rayStage[2] = make_float4(sh.my.x, sh.my.y, sh.my.z, __int_as_float(sh.kz));
        sp = 0; tsp = 0; tcount = 0;
        // any hit: nothing can lie in (0, tmax) when tmax <= 0 — no node work is queued and the ray finishes, unoccluded,
        // with the iteration it was fetched in (the light sample closer than the 0.001 the reference subtracts; k_shadow's
        // null-contribution rays). The closest-hit kernels pass a constant tmax > 0.
        ngroup = make_uint2(0u, (ANY && !(tmax_ > 0.0f)) ? 0u : 0x80000000u);
        tgroup = make_uint2(0u, 0u);
    }

    __device__ __forceinline__ bool want_node() const { return ngroup.y > 0x00FFFFFFu; }
    __device__ __forceinline__ bool stack_empty() const { return sp == 0; }

    // No node group current: take the next one from the stack. Returns false when no node work is left.
    __device__ __forceinline__ bool pop(WarpShared<ANY>& ws) {
        if (sp == 0) return false;
        constexpr int K = StackShared<ANY>::value;
        --sp;
        if constexpr (K > 0) { if (sp < K) ngroup = ws.nstack[sp][threadIdx.x & 31u]; else ngroup = stack[sp - K]; }
        else ngroup = stack[sp];
        return true;
    }

    // next queued triangle index; requires tcount > 0
    __device__ __forceinline__ uint32_t take_tri(WarpShared<ANY>& ws) {
        if (tgroup.y == 0u) tgroup = tst(ws, --tsp);
        const uint32_t ti = 31u - (uint32_t)__clz(tgroup.y);
        tgroup.y &= ~(1u << ti);
        tcount--;
        return tgroup.x + ti;
    }

    // Pop the nearest pending child of the current node group and test its 8 children; the triangles it yields are
    // queued on tstack for the warp's pooled triangle phase. Requires want_node().
    __device__ __forceinline__ void node_step(const WideNode* __restrict__ nodes, const TriRecord* __restrict__ tris_for_prefetch,
                                              uint32_t& nodeVisits, WarpShared<ANY>& ws, const unsigned long long pol) {
        const uint32_t hits = ngroup.y;
        const uint32_t bitIndex = 31u - (uint32_t)__clz(hits);
        const uint32_t base = ngroup.x;
        ngroup.y &= ~(1u << bitIndex);
        if (ngroup.y > 0x00FFFFFFu) push(ngroup, ws);
        const uint32_t slot = (bitIndex - 24u) ^ oct3();
        const uint32_t rel = __popc(hits & ~(0xFFFFFFFFu << slot) & 0xFFu);
        const float4* np = reinterpret_cast<const float4*>(nodes + (base + rel));
#if RB_WIDE_LOADS
        float4 n0, n1, n2, n3, n4, npad;
        ld_bvh2(np + 0, n0, n1, pol); ld_bvh2(np + 2, n2, n3, pol); ld_bvh2(np + 4, n4, npad, pol);
#else
        const float4 n0 = ld_bvh(np + 0, pol), n1 = ld_bvh(np + 1, pol), n2 = ld_bvh(np + 2, pol), n3 = ld_bvh(np + 3, pol), n4 = ld_bvh(np + 4, pol);
#endif
        if (COUNT) nodeVisits++;

        const uint32_t eim = __float_as_uint(n0.w);
        const float sx = __uint_as_float((eim & 0xFFu) << 23) * idx;
        const float sy = __uint_as_float(((eim >> 8) & 0xFFu) << 23) * idy;
        const float sz = __uint_as_float(((eim >> 16) & 0xFFu) << 23) * idz;
#if RB_ORIGIN_FROM_STAGE
        const float4 o = ws.ray[threadIdx.x & 31u][0];
#endif
        const float cx = (n0.x - o.x) * idx, cy = (n0.y - o.y) * idy, cz = (n0.z - o.z) * idz;
        const float eps = 9.5367431640625e-07f;   // 2^-20
        // slack: 2^-20 * (255 |s| + |c|) bounds the rounding of q*s + c and the acceptance band of the triangle test;
        // + 2^-9 |s| (= 2^-20 * 2048, here 2305 with margin) bounds the rounding of the biased offsets c - 2^15 s
        const float jx = eps * fmaf(2560.0f, fabsf(sx), fabsf(cx));
        const float jy = eps * fmaf(2560.0f, fabsf(sy), fabsf(cy));
        const float jz = eps * fmaf(2560.0f, fabsf(sz), fabsf(cz));
        // + a margin in ray parameters, the same on every plane and on the best-t limit: the watertight test computes t as
        // the barycentric mean of the vertices' ray parameters, whose fp32 error is proportional to the EXTENT of the triangle
        // in ray parameters, not to t. A ray leaving a large floor triangle finds the neighbouring triangle at
        // t = 2.65e-4 +- 3e-8, or — starting 1e-7 BEHIND a two-metre triangle and moving away from it — at t = +4e-8.
        // What the triangle test accepts must never be culled, or the closest hit depends on the visiting order (found by
        // the full-size C2 parity test: 4 of 518,400 pixels; tests/test_gpu_parity2.py has the ray set). 2^-14 of the node's
        // extent in ray parameters bounds that error for triangles with an aspect ratio up to ~64 and weakens the cull by
        // less than 1e-4 of a node's size.
        // (the vertices' ray parameters are measured along the ray's dominant axis — the one with the smallest |1/d| —, so the
        // node's slab along THAT axis bounds them; the other axes' slabs can be arbitrarily wide for a nearly parallel ray)
#if RB_DOM_AXIS
        const uint32_t dom = oct_inv >> 8;
        const float es = dom == 0u ? sx : (dom == 1u ? sy : sz), ec = dom == 0u ? cx : (dom == 1u ? cy : cz);
        const float margin = RB_CULL_MARGIN * fmaf(256.0f, fabsf(es), fabsf(ec));
#else
        const float ax = fabsf(idx), ay = fabsf(idy), az = fabsf(idz);
        const float ex = fmaf(256.0f, fabsf(sx), fabsf(cx)), ey = fmaf(256.0f, fabsf(sy), fabsf(cy)), ez = fmaf(256.0f, fabsf(sz), fabsf(cz));
        const float margin = RB_CULL_MARGIN * ((ax <= ay && ax <= az) ? ex : (ay <= az ? ey : ez));
#endif
        const float kx = jx + margin, ky = jy + margin, kz = jz + margin;
        const float cnx = fmaf(-BYTE_BIAS, sx, cx - kx), cfx = fmaf(-BYTE_BIAS, sx, cx + kx);
        const float cny = fmaf(-BYTE_BIAS, sy, cy - ky), cfy = fmaf(-BYTE_BIAS, sy, cy + ky);
        const float cnz = fmaf(-BYTE_BIAS, sz, cz - kz), cfz = fmaf(-BYTE_BIAS, sz, cz + kz);
        const uint32_t oct_inv4 = oct3() * 0x01010101u;
        const uint32_t kb = c_byteBiasWord;

        ngroup.x = __float_as_uint(n1.x);
        uint32_t hitmask = 0u;
        const float tcur = best.t + margin;
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const uint32_t meta4 = __float_as_uint(half == 0 ? n1.z : n1.w);
            const uint32_t isInner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
            const uint32_t innerMask4 = sign_extend_s8x4(isInner4 << 3);
            const uint32_t bitIndex4 = (meta4 ^ (oct_inv4 & innerMask4)) & 0x1F1F1F1Fu;
            const uint32_t childBits4 = (meta4 >> 5) & 0x07070707u;
            const uint32_t qlox = __float_as_uint(half == 0 ? n2.x : n2.y);
            const uint32_t qloy = __float_as_uint(half == 0 ? n2.z : n2.w);
            const uint32_t qloz = __float_as_uint(half == 0 ? n3.x : n3.y);
            const uint32_t qhix = __float_as_uint(half == 0 ? n3.z : n3.w);
            const uint32_t qhiy = __float_as_uint(half == 0 ? n4.x : n4.y);
            const uint32_t qhiz = __float_as_uint(half == 0 ? n4.z : n4.w);
            const uint32_t nx = idx < 0.f ? qhix : qlox, fx = idx < 0.f ? qlox : qhix;
            const uint32_t ny = idy < 0.f ? qhiy : qloy, fy = idy < 0.f ? qloy : qhiy;
            const uint32_t nz = idz < 0.f ? qhiz : qloz, fz = idz < 0.f ? qloz : qhiz;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float t0x = fmaf(byte_f(nx, j, kb), sx, cnx), t1x = fmaf(byte_f(fx, j, kb), sx, cfx);
                const float t0y = fmaf(byte_f(ny, j, kb), sy, cny), t1y = fmaf(byte_f(fy, j, kb), sy, cfy);
                const float t0z = fmaf(byte_f(nz, j, kb), sz, cnz), t1z = fmaf(byte_f(fz, j, kb), sz, cfz);
                const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, 0.0f));
                const float tf = fminf(fminf(t1x, t1y), fminf(t1z, tcur));
                if (tn <= tf) {
                    const uint32_t cb = (childBits4 >> (8 * j)) & 0xFFu;
                    const uint32_t bi = (bitIndex4 >> (8 * j)) & 0xFFu;
                    hitmask |= cb << bi;
                }
            }
        }
        ngroup.y = (hitmask & 0xFF000000u) | (eim >> 24);
        const uint32_t tmask = hitmask & 0x00FFFFFFu;
        if (tmask) { tst(ws, tsp++) = make_uint2(__float_as_uint(n1.y), tmask); tcount += (uint32_t)__popc(tmask); }
#if RB_PREFETCH
        // pull what this ray touches next towards L1 while other warps run: the nearest hit child node and the first
        // queued triangle (the traversal kernels stall mostly on these dependent fetches)
        if (ngroup.y > 0x00FFFFFFu) {
            const uint32_t nb = 31u - (uint32_t)__clz(ngroup.y);
            const uint32_t nslot = (nb - 24u) ^ oct3();
            const uint32_t nrel = __popc(ngroup.y & ~(0xFFFFFFFFu << nslot) & 0xFFu);
            const char* pn = reinterpret_cast<const char*>(nodes + (ngroup.x + nrel));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pn));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pn + 64));
        }
        if (tmask) {
            const char* pt = reinterpret_cast<const char*>(tris_for_prefetch + (__float_as_uint(n1.y) + (31u - (uint32_t)__clz(tmask))));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pt));
        }
#endif
    }
};

// t > 0 always, so its bit pattern orders like the value; ties fall to the smaller global primitive id
__device__ __forceinline__ unsigned long long hit_key(float t, uint32_t gid) {
    return ((unsigned long long)__float_as_uint(t) << 32) | (unsigned long long)gid;
}

// Warp-cooperative persistent trace loop (see the header comment).
//   fetch(i)      -> load ray i into (o, d, tmax); called for i < n
//   commit(i, h)  -> consume the finished ray i
//   look(i)       -> (RB_PRECLAIM) first half of a prefetch of ray i: returns what is needed to address its record
//                    (the queue entry), request(v) -> second half: asks for the record itself (prefetch instructions)
struct NoLook { __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return i; } };
struct NoRequest { __device__ __forceinline__ void operator()(uint32_t) const { } };
template <bool ANY, bool COUNT, class Fetch, class Commit, class Look = NoLook, class Request = NoRequest>
__device__ __forceinline__ void trace_queue(const WideNode* __restrict__ nodes, const TriRecord* __restrict__ tris,
                                            uint32_t n, uint32_t* cursor, Fetch fetch, Commit commit,
                                            uint32_t& nodeVisits, uint32_t& triTests, WarpShared<ANY>& ws,
                                            Look look = Look(), Request request = Request()) {
    const uint32_t lane = threadIdx.x & 31u;
    Traversal<ANY, COUNT> tr;
#if !RB_STACK_IN_STRUCT
    uint2 deepStack[TRAV_STACK];
    tr.stack = deepStack;
#endif
    tr.tcount = 0;
    const unsigned long long pol = bvh_policy();

    bool has = false;
    bool exhausted = false;
    uint32_t rayIdx = 0;
#if RB_PRECLAIM
    uint32_t nextI = 0xFFFFFFFFu;        // index of the ray this lane claimed ahead of time (0xFFFFFFFF: none)
#endif
    for (;;) {
        // ---- 1. refill ----
        uint32_t busy = __ballot_sync(0xffffffffu, has);
#if RB_PRECLAIM
        {
            // idle lanes that claimed their next ray before the last triangle phase already hold its index
            uint32_t idx = has ? 0xFFFFFFFFu : nextI;
            nextI = 0xFFFFFFFFu;
            const uint32_t need = __ballot_sync(0xffffffffu, !has && idx == 0xFFFFFFFFu);
            if (!exhausted && need != 0u && __popc(busy) < RB_REFILL) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(cursor, (uint32_t)__popc(need));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (!has && idx == 0xFFFFFFFFu) idx = base + __popc(need & ((1u << lane) - 1u));
                if (base + (uint32_t)__popc(need) >= n) exhausted = true;
            }
            if (!has && idx < n) {
                rb_v3 o, d; float tmax;
                fetch(idx, o, d, tmax);
                tr.init(o, d, tmax, ws.ray[lane]);
                rayIdx = idx; has = true;
            }
            busy = __ballot_sync(0xffffffffu, has);
        }
#else
        if (!exhausted && __popc(busy) < RB_REFILL) {
            const uint32_t need = ~busy;
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(cursor, (uint32_t)__popc(need));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (!has) {
                const uint32_t i = base + __popc(need & ((1u << lane) - 1u));
                if (i < n) {
                    rb_v3 o, d; float tmax;
                    fetch(i, o, d, tmax);
                    tr.init(o, d, tmax, ws.ray[lane]);
                    rayIdx = i; has = true;
                }
            }
            if (base + (uint32_t)__popc(need) >= n) exhausted = true;
            busy = __ballot_sync(0xffffffffu, has);
        }
#endif
        if (busy == 0u) break;

        // ---- 2. a chunk of wide-node steps per lane ----
        // (ending the chunk early once fewer than 8 / 16 / 24 lanes still have node work — a warp-uniform loop with a
        // ballot per step — was measured at -12 / -14 / -18 % closest-hit rays/s: the fixed chunk stays)
        if (has) {
#pragma unroll 1
            for (int it = 0; it < (ANY ? RB_CHUNK_ANY : RB_CHUNK); it++) {
                if (!tr.want_node() && !tr.pop(ws)) break;
                tr.node_step(nodes, tris, nodeVisits, ws, pol);
            }
        }

#if RB_PRECLAIM
        // ---- 2b. lanes that will be idle after the triangle phase claim their next ray now ----
        uint32_t looked = 0u;
        {
            const bool leaving = has && !tr.want_node() && tr.stack_empty();
            const uint32_t lv = __ballot_sync(0xffffffffu, leaving);
            if (!exhausted && lv != 0u) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(cursor, (uint32_t)__popc(lv));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (leaving) {
                    nextI = base + __popc(lv & ((1u << lane) - 1u));
                    if (nextI < n) looked = look(nextI);                 // the load is in flight while the work list is filled
                    else nextI = 0xFFFFFFFFu;
                }
                if (base + (uint32_t)__popc(lv) >= n) exhausted = true;
            }
        }
        bool requested = false;
#endif

        // ---- 3. pooled triangle phase ----
        bool anyHitFound = false;
        for (;;) {
            const uint32_t c = has ? tr.tcount : 0u;
            uint32_t incl = c;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
                if ((int)lane >= off) incl += v;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            if (total == 0u) break;
            uint32_t pos = incl - c;
            const unsigned long long seed = has ? hit_key(tr.best.t, tr.best.gid) : ~0ull;
            ws.bestKey[lane] = seed;
#if RB_COOP_FILL
            static_assert(!RB_DEFER_BARY, "the cooperative fill uses the payload rows as scratch; RB_DEFER_BARY keeps the best candidate there");
            bool filled = false;
            if constexpr (TStackShared<ANY>::value && RB_TSTACK_N >= RB_CHUNK) if (total <= (uint32_t)RB_WORK_CAP) {
                filled = true;
                // Cooperative fill of the work list: position p of the list is computed by lane p % 32, whoever owns the
                // triangle. The per-lane loop below ran max-over-lanes iterations (~17) with 5.6 of 32 lanes active — a few
                // rays own most of a chunk's triangles — and was 11 % of the kernel's issued instructions (profiles/r02f,
                // SASS page); this takes total / 32 rounds (~3) with all lanes. The order of the list is irrelevant.
                uint32_t* sc = reinterpret_cast<uint32_t*>(ws.payload);      // 128 words, free until the tests start:
                const uint32_t excl = pos;                                  // [0..31] first position of each lane's triangles
                sc[lane] = excl;
                sc[32u + lane] = tr.tgroup.x;                                // [32..63], [64..95] the group a lane holds in registers
                sc[64u + lane] = c ? tr.tgroup.y : 0u;
                const uint32_t owners = __ballot_sync(0xffffffffu, c > 0u);
                if (c) sc[96u + __popc(owners & ((1u << lane) - 1u))] = lane | ((uint32_t)tr.tsp << 8);   // [96..127] owners in order + their list length
                __syncwarp();
                for (uint32_t b = 0; b < total; b += 32u) {
                    const uint32_t rel = excl - b;
                    const uint32_t starts = __reduce_or_sync(0xffffffffu, (c > 0u && rel < 32u) ? (1u << rel) : 0u);
                    const uint32_t before = (uint32_t)__popc(__ballot_sync(0xffffffffu, c > 0u && excl < b));
                    const uint32_t p = b + lane;
                    if (p < total) {
                        const uint32_t ord = before + (uint32_t)__popc(starts & (0xFFFFFFFFu >> (31u - lane))) - 1u;
                        const uint32_t info = sc[96u + ord];
                        const uint32_t L = info & 31u;
                        int k = (int)(info >> 8);
                        uint32_t off = p - sc[L];
                        uint32_t mask = sc[64u + L], base = sc[32u + L];
                        uint32_t cnt = (uint32_t)__popc(mask);
                        while (off >= cnt) {             // the register group first, then the lane's list from the top
                            off -= cnt;
                            const uint2 g = ws.tstack[--k][L];
                            base = g.x; mask = g.y; cnt = (uint32_t)__popc(mask);
                        }
                        for (; off > 0u; off--) mask &= mask - 1u;
                        ws.work[p] = ((base + (uint32_t)__ffs(mask) - 1u) << 5) | L;
                    }
                }
                if (has) { tr.tcount = 0u; tr.tsp = 0; tr.tgroup.y = 0u; }
            }
            if (!filled)
#endif
#if RB_LEAN_FILL
            if (total <= (uint32_t)RB_WORK_CAP) {
                // Everything fits (the common case): two nested loops — groups, then the bits of a group — without the
                // per-triangle bookkeeping of take_tri (group reload test, tcount, list bound): 8 instructions per
                // triangle instead of 21. The loop runs max-over-lanes iterations with ~6 of 32 lanes active (a few rays
                // own most of a chunk's triangles) and was 11 % of the kernel's issued instructions (profiles/r02f).
                if (c) {
                    uint32_t* w = &ws.work[pos];
                    uint2 g = tr.tgroup;
                    int k = tr.tsp;
                    for (;;) {
                        const uint32_t tag = (g.x << 5) | lane;
#if RB_FILL_UNROLL2
                        // two triangles per trip: the second store is predicated, the loop overhead is paid once per pair
                        while (g.y) {
                            const uint32_t ti = 31u - (uint32_t)__clz(g.y);
                            g.y ^= 1u << ti;
                            w[0] = tag + (ti << 5);
                            const bool more = g.y != 0u;
                            const uint32_t tj = 31u - (uint32_t)__clz(g.y | 1u);
                            if (more) { g.y ^= 1u << tj; w[1] = tag + (tj << 5); }
                            w += more ? 2 : 1;
                        }
#else
                        while (g.y) {
                            const uint32_t ti = 31u - (uint32_t)__clz(g.y);
                            g.y ^= 1u << ti;
                            *w++ = tag + (ti << 5);
                        }
#endif
                        if (k == 0) break;
                        g = tr.tst(ws, --k);
                    }
                    tr.tcount = 0u; tr.tsp = 0; tr.tgroup.y = 0u;
                }
            } else
#endif
            while (has && tr.tcount > 0u && pos < (uint32_t)RB_WORK_CAP) ws.work[pos++] = (tr.take_tri(ws) << 5) | lane;
            __syncwarp();
#if RB_PRECLAIM
            if (!requested) { requested = true; if (nextI != 0xFFFFFFFFu) request(looked); }      // ... and the record while the triangles are tested
#endif
            const uint32_t count = min(total, (uint32_t)RB_WORK_CAP);
            for (uint32_t b = 0; b < count; b += 32u) {
                bool cand = false;
                unsigned long long mykey = 0ull;
                uint32_t owner = 0, triIdx = 0;
                float b1 = 0.f, b2 = 0.f, det = 0.f;
                if (b + lane < count) {
                    const uint32_t item = ws.work[b + lane];
                    triIdx = item >> 5; owner = item & 31u;
                    const float4 r0 = ws.ray[owner][0], r1 = ws.ray[owner][1], r2 = ws.ray[owner][2];
                    const float4* tp = reinterpret_cast<const float4*>(tris + triIdx);
#if RB_WIDE_LOADS
                    float4 va, vb, vc, vpad;
                    ld_bvh2(tp + 0, va, vb, pol); ld_bvh2(tp + 2, vc, vpad, pol);
#elif RB_TRI_LDCG
                    // triangles are read about once per ray: keep them out of L1 so that it holds wide nodes
                    const float4 va = __ldcg(tp + 0), vb = __ldcg(tp + 1), vc = __ldcg(tp + 2);
#else
                    const float4 va = ld_bvh(tp + 0, pol), vb = ld_bvh(tp + 1, pol), vc = ld_bvh(tp + 2, pol);
#endif
                    if (COUNT) triTests++;
                    rb_ray_shear sh;
                    sh.mx = rb_mk3(r1.x, r1.y, r1.z); sh.my = rb_mk3(r2.x, r2.y, r2.z); sh.mz = rb_axis3(__float_as_int(r2.w), r1.w);
                    float t;
#if RB_DEFER_BARY
                    float T;
                    const bool crosses = rb_tri_edges(rb_mk3(r0.x, r0.y, r0.z), sh, rb_mk3(va.x, va.y, va.z), rb_mk3(vb.x, vb.y, vb.z),
                                                      rb_mk3(vc.x, vc.y, vc.z), &det, &T, &b1, &b2);      // b1, b2 hold V, W until the commit
                    if (crosses) t = T / det;
#elif RB_BARY_LATE
                    float T;
                    const bool crosses = rb_tri_edges(rb_mk3(r0.x, r0.y, r0.z), sh, rb_mk3(va.x, va.y, va.z), rb_mk3(vb.x, vb.y, vb.z),
                                                      rb_mk3(vc.x, vc.y, vc.z), &det, &T, &b1, &b2);
                    if (crosses) t = T / det;
#else
                    const bool crosses = rb_tri_intersect(rb_mk3(r0.x, r0.y, r0.z), sh, rb_mk3(va.x, va.y, va.z), rb_mk3(vb.x, vb.y, vb.z),
                                                          rb_mk3(vc.x, vc.y, vc.z), &t, &b1, &b2);
#endif
                    if (crosses) {
                        if (t > 0.0f && t < r0.w) {
#if RB_BARY_LATE && !RB_DEFER_BARY
                            if constexpr (!ANY) { b1 = b1 / det; b2 = b2 / det; }
#endif
                            mykey = hit_key(t, __float_as_uint(vc.w));
                            atomicMin(&ws.bestKey[owner], mykey);
                            cand = true;
                        }
                    }
                }
                if constexpr (!ANY) {
                    __syncwarp();
                    if (cand && ws.bestKey[owner] == mykey) ws.payload[owner] = make_float4(b1, b2, __uint_as_float(triIdx), det);
                }
            }
            __syncwarp();
            if (has) {
                const unsigned long long k = ws.bestKey[lane];
                if (k != seed) {
                    if constexpr (ANY) anyHitFound = true;
                    else {
                        tr.best.t = __uint_as_float((uint32_t)(k >> 32)); tr.best.gid = (uint32_t)k;
#if !RB_DEFER_BARY
                        const float4 p = ws.payload[lane];
                        tr.best.b1 = p.x; tr.best.b2 = p.y; tr.best.tri = __float_as_uint(p.z);
#endif
                    }
                }
            }
            __syncwarp();     // bestKey / work are rewritten by the next pass
            if (total <= (uint32_t)RB_WORK_CAP) break;
        }

#if RB_PRECLAIM
        if (!requested && nextI != 0xFFFFFFFFu) request(looked);      // no triangles were queued in this chunk
#endif
        // ---- finished rays ----
        if (has) {
            if (ANY && anyHitFound) {
                tr.best.tri = 0u;     // any-hit: only "occluded or not" is meaningful
                commit(rayIdx, tr.best); has = false; tr.tcount = 0u;
            } else if (!tr.want_node() && tr.stack_empty() && tr.tcount == 0u) {
#if RB_DEFER_BARY
                if constexpr (!ANY) {
                    // the payload row still holds (V, W, triangle, det) of the candidate that set the final key
                    tr.best.b1 = 0.f; tr.best.b2 = 0.f; tr.best.tri = 0xFFFFFFFFu;
                    if (tr.best.gid != 0xFFFFFFFFu) {
                        const float4 p = ws.payload[lane];
                        tr.best.b1 = p.x / p.w; tr.best.b2 = p.y / p.w; tr.best.tri = __float_as_uint(p.z);
                    }
                }
#endif
                commit(rayIdx, tr.best); has = false;
            }
        }
        __syncwarp();
    }
}

} // namespace rb200
