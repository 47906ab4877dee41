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
// Warp organisation (trace_queue): all 32 lanes run one convergent loop; a lane that owns a ray performs RB_CHUNK
// wide-node steps, then tests the triangles those steps produced in ONE loop (a single node yields ~0.5 triangles,
// so testing per node leaves ~3 of 32 lanes active; per chunk several times more lanes have work), then the warp
// checks how many lanes are idle and the idle lanes refill from the ray queue with one aggregated atomic, so lanes
// whose rays finish early do not wait for the slowest ray of the warp. (Measured on B200: warp-level triangle
// phases with waiting lanes, and unbounded postponing, lost more than they gained.)
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
#define RB_CHUNK 8        // wide-node steps a lane performs between two warp-level refill checks (swept 2..24 on B200)
#endif

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
__device__ __forceinline__ float byte_f(uint32_t w, int j) {
    return __uint_as_float(__byte_perm(w, 0x47000000u, 0x7404u + ((uint32_t)j << 4)));
}

template <bool ANY, bool COUNT>
struct Traversal {
    rb_v3 o;
    float idx, idy, idz, tmax;
    rb_ray_shear shear;
    uint32_t oct_inv;
    uint2 ngroup, tgroup;
    int sp;
    RayHit best;
    uint2 stack[TRAV_STACK];
    uint2 tstack[RB_CHUNK];   // triangle groups found during the current chunk of node steps, tested together
    int tsp;

    __device__ __forceinline__ void init(const rb_v3 org, const rb_v3 d, const float tmax_) {
        o = org; tmax = tmax_;
        best.t = tmax_; best.b1 = 0.f; best.b2 = 0.f; best.tri = 0xFFFFFFFFu; best.gid = 0xFFFFFFFFu;
        const float ooeps = 8.2718061e-25f;   // 2^-80: keeps 1/d finite for axis-parallel rays (box tests only)
        idx = 1.0f / (fabsf(d.x) > ooeps ? d.x : copysignf(ooeps, d.x));
        idy = 1.0f / (fabsf(d.y) > ooeps ? d.y : copysignf(ooeps, d.y));
        idz = 1.0f / (fabsf(d.z) > ooeps ? d.z : copysignf(ooeps, d.z));
        oct_inv = (idx < 0.f ? 0u : 4u) | (idy < 0.f ? 0u : 2u) | (idz < 0.f ? 0u : 1u);
        shear = rb_ray_prepare(d);
        sp = 0; tsp = 0;
        ngroup = make_uint2(0u, 0x80000000u);
        tgroup = make_uint2(0u, 0u);
    }

    __device__ __forceinline__ bool want_node() const { return ngroup.y > 0x00FFFFFFu; }
    __device__ __forceinline__ bool want_tri() const { return tgroup.y != 0u || tsp > 0; }

    // No node group current: take the next one from the stack. Returns false when no node work is left.
    __device__ __forceinline__ bool pop() {
        if (sp == 0) return false;
        ngroup = stack[--sp];
        return true;
    }

    // Pop the nearest pending child of the current node group and test its 8 children; the triangles it yields are
    // queued on tstack (tested later, together with those of the other node steps of the chunk). Requires want_node().
    __device__ __forceinline__ void node_step(const WideNode* __restrict__ nodes, uint32_t& nodeVisits) {
        const uint32_t hits = ngroup.y;
        const uint32_t bitIndex = 31u - (uint32_t)__clz(hits);
        const uint32_t base = ngroup.x;
        ngroup.y &= ~(1u << bitIndex);
        if (ngroup.y > 0x00FFFFFFu) { stack[sp++] = ngroup; }
        const uint32_t slot = (bitIndex - 24u) ^ oct_inv;
        const uint32_t rel = __popc(hits & ~(0xFFFFFFFFu << slot) & 0xFFu);
        const float4* np = reinterpret_cast<const float4*>(nodes + (base + rel));
        const float4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
        if (COUNT) nodeVisits++;

        const uint32_t eim = __float_as_uint(n0.w);
        const float sx = __uint_as_float((eim & 0xFFu) << 23) * idx;
        const float sy = __uint_as_float(((eim >> 8) & 0xFFu) << 23) * idy;
        const float sz = __uint_as_float(((eim >> 16) & 0xFFu) << 23) * idz;
        const float cx = (n0.x - o.x) * idx, cy = (n0.y - o.y) * idy, cz = (n0.z - o.z) * idz;
        const float eps = 9.5367431640625e-07f;   // 2^-20
        // slack: 2^-20 * (255 |s| + |c|) bounds the rounding of q*s + c and the acceptance band of the triangle test;
        // + 2^-9 |s| (= 2^-20 * 2048, here 2305 with margin) bounds the rounding of the biased offsets c - 2^15 s
        const float kx = eps * fmaf(2560.0f, fabsf(sx), fabsf(cx));
        const float ky = eps * fmaf(2560.0f, fabsf(sy), fabsf(cy));
        const float kz = eps * fmaf(2560.0f, fabsf(sz), fabsf(cz));
        const float cnx = fmaf(-BYTE_BIAS, sx, cx - kx), cfx = fmaf(-BYTE_BIAS, sx, cx + kx);
        const float cny = fmaf(-BYTE_BIAS, sy, cy - ky), cfy = fmaf(-BYTE_BIAS, sy, cy + ky);
        const float cnz = fmaf(-BYTE_BIAS, sz, cz - kz), cfz = fmaf(-BYTE_BIAS, sz, cz + kz);
        const uint32_t oct_inv4 = oct_inv * 0x01010101u;

        ngroup.x = __float_as_uint(n1.x);
        uint32_t hitmask = 0u;
        const float tcur = best.t;
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
                const float t0x = fmaf(byte_f(nx, j), sx, cnx), t1x = fmaf(byte_f(fx, j), sx, cfx);
                const float t0y = fmaf(byte_f(ny, j), sy, cny), t1y = fmaf(byte_f(fy, j), sy, cfy);
                const float t0z = fmaf(byte_f(nz, j), sz, cnz), t1z = fmaf(byte_f(fz, j), sz, cfz);
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
        if (hitmask & 0x00FFFFFFu) tstack[tsp++] = make_uint2(__float_as_uint(n1.y), hitmask & 0x00FFFFFFu);
    }

    // Test one pending triangle. Returns true when an any-hit query is decided. Requires want_tri().
    __device__ __forceinline__ bool tri_step(const TriRecord* __restrict__ tris, uint32_t& triTests) {
        if (tgroup.y == 0u) tgroup = tstack[--tsp];
        const uint32_t ti = 31u - (uint32_t)__clz(tgroup.y);
        tgroup.y &= ~(1u << ti);
        const uint32_t triIdx = tgroup.x + ti;
        const float4* tp = reinterpret_cast<const float4*>(tris + triIdx);
        const float4 a = __ldg(tp + 0), b = __ldg(tp + 1), c = __ldg(tp + 2);
        if (COUNT) triTests++;
        float t, b1, b2;
        if (rb_tri_intersect(o, shear, rb_mk3(a.x, a.y, a.z), rb_mk3(b.x, b.y, b.z), rb_mk3(c.x, c.y, c.z), &t, &b1, &b2)) {
            if (t > 0.0f && t < tmax) {
                const uint32_t gid = __float_as_uint(c.w);
                if (ANY) { best.t = t; best.tri = triIdx; best.gid = gid; return true; }
                if (t < best.t || (t == best.t && gid < best.gid)) {
                    best.t = t; best.b1 = b1; best.b2 = b2; best.tri = triIdx; best.gid = gid;
                }
            }
        }
        return false;
    }
};

// Warp-cooperative persistent trace loop (see the header comment).
//   fetch(i)      -> load ray i into (o, d, tmax); called for i < n
//   commit(i, h)  -> consume the finished ray i
template <bool ANY, bool COUNT, class Fetch, class Commit>
__device__ __forceinline__ void trace_queue(const WideNode* __restrict__ nodes, const TriRecord* __restrict__ tris,
                                            uint32_t n, uint32_t* cursor, Fetch fetch, Commit commit,
                                            uint32_t& nodeVisits, uint32_t& triTests) {
    const uint32_t lane = threadIdx.x & 31u;
    Traversal<ANY, COUNT> tr;
    bool has = false;
    bool exhausted = false;
    uint32_t rayIdx = 0;
    for (;;) {
        uint32_t busy = __ballot_sync(0xffffffffu, has);
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
                    tr.init(o, d, tmax);
                    rayIdx = i; has = true;
                }
            }
            if (base + (uint32_t)__popc(need) >= n) exhausted = true;
            busy = __ballot_sync(0xffffffffu, has);
        }
        if (busy == 0u) break;

        if (has) {
            bool done = false;
#pragma unroll 1
            for (int it = 0; it < RB_CHUNK; it++) {
                if (!tr.want_node() && !tr.pop()) break;
                tr.node_step(nodes, nodeVisits);
            }
            // the triangles of the whole chunk in one loop: more lanes have work at the same time
            while (tr.want_tri()) {
                if (tr.tri_step(tris, triTests)) { done = true; break; }
            }
            if (!done && !tr.want_node() && tr.sp == 0) done = true;
            if (done) { commit(rayIdx, tr.best); has = false; }
        }
        __syncwarp();
    }
}

// plain one-ray traversal (kept for callers that own exactly one ray per thread)
template <bool ANY, bool COUNT>
__device__ __forceinline__ void traverse(const WideNode* __restrict__ nodes, const TriRecord* __restrict__ tris,
                                         const rb_v3 o, const rb_v3 d, const float tmax, RayHit& best,
                                         uint32_t& nodeVisits, uint32_t& triTests) {
    Traversal<ANY, COUNT> tr;
    tr.init(o, d, tmax);
    for (;;) {
        if (tr.want_tri()) { if (tr.tri_step(tris, triTests)) break; }
        else if (tr.want_node()) tr.node_step(nodes, nodeVisits);
        else if (!tr.pop()) break;
    }
    best = tr.best;
}

} // namespace rb200
