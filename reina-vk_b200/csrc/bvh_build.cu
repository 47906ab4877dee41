// bvh_build.cu — deterministic GPU builder: triangles -> 63-bit Morton codes -> stable LSD radix sort ->
// Karras-2012 binary hierarchy -> bottom-up refit -> level-by-level (scan-allocated) collapse to an 8-wide
// compressed BVH with quantised child boxes -> triangles copied into leaf order.
//
// Replaces the driver-side acceleration-structure build of the reference:
//   BLAS  vkCmdBuildAccelerationStructuresKHR  src/graphics/Blas.cpp:8-124 (PREFER_FAST_TRACE, opaque triangles)
//   TLAS  vktools::createTlas                  src/tools/vktools.cpp:460-596
// Instances are flattened into world space (the reference never updates an acceleration structure, so a single
// static hierarchy over all instances is equivalent and avoids the two-level transform on every ray).
//
// Determinism: every stage is a pure function of its input (sort is stable, allocation is by prefix sum, refit is
// min/max), so node and triangle arrays are bit-identical across runs and GPUs (hash exposed in RB200BvhInfo).
#include "common.cuh"
#include <algorithm>

namespace rb200 {

static constexpr uint32_t LEAF_FLAG = 0x80000000u;
#ifndef RB_MAX_LEAF_TRIS
#define RB_MAX_LEAF_TRIS 3
#endif
static constexpr int MAX_LEAF_TRIS = RB_MAX_LEAF_TRIS;   // triangles per leaf slot (<= 3: unary count in 3 meta bits)

// ---------------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

// ---------------------------------------------------------------------------------------------------
// 1. flatten instances into world-space triangles, scene bounds
// ---------------------------------------------------------------------------------------------------
__global__ void k_flatten(const float4* __restrict__ verts, const uint32_t* __restrict__ indices,
                          const RB200Instance* __restrict__ inst, const uint32_t* __restrict__ instPrefix,
                          uint32_t numInst, uint32_t N, TriRecord* __restrict__ out, uint32_t* __restrict__ bounds) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    if (g < N) {
        uint32_t a = 0, b = numInst;           // instance i with instPrefix[i] <= g < instPrefix[i+1]
        while (b - a > 1) { uint32_t m = (a + b) >> 1; if (instPrefix[m] <= g) a = m; else b = m; }
        const RB200Instance& in = inst[a];
        uint32_t p = g - instPrefix[a];
        rb_v3 w[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            uint32_t vi = indices[3 * p + in.indexOffset + k];
            float4 v = verts[vi];
            w[k] = rb_m4_point(in.transform, rb_mk3(v.x, v.y, v.z));
            lo[0] = fminf(lo[0], w[k].x); lo[1] = fminf(lo[1], w[k].y); lo[2] = fminf(lo[2], w[k].z);
            hi[0] = fmaxf(hi[0], w[k].x); hi[1] = fmaxf(hi[1], w[k].y); hi[2] = fmaxf(hi[2], w[k].z);
        }
        TriRecord t;
        t.v0 = make_float4(w[0].x, w[0].y, w[0].z, __uint_as_float(p));
        t.v1 = make_float4(w[1].x, w[1].y, w[1].z, __uint_as_float(a));
        t.v2 = make_float4(w[2].x, w[2].y, w[2].z, __uint_as_float(g));
        out[g] = t;
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
        float l = lo[a], h = hi[a];
        for (int o = 16; o > 0; o >>= 1) {
            l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
            h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o));
        }
        if ((threadIdx.x & 31) == 0 && l <= h) {
            atomicMin(&bounds[a], float_to_ordered(l));
            atomicMax(&bounds[3 + a], float_to_ordered(h));
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// 2. Morton codes (21 bits per axis of the triangle-AABB centre, normalised to the scene bounds)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t expand21(uint64_t v) {
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

__global__ void k_morton(const TriRecord* __restrict__ tris, uint32_t N, const uint32_t* __restrict__ bounds,
                         uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= N) return;
    TriRecord t = tris[g];
    float c[3] = {0.5f * (fminf(fminf(t.v0.x, t.v1.x), t.v2.x) + fmaxf(fmaxf(t.v0.x, t.v1.x), t.v2.x)),
                  0.5f * (fminf(fminf(t.v0.y, t.v1.y), t.v2.y) + fmaxf(fmaxf(t.v0.y, t.v1.y), t.v2.y)),
                  0.5f * (fminf(fminf(t.v0.z, t.v1.z), t.v2.z) + fmaxf(fmaxf(t.v0.z, t.v1.z), t.v2.z))};
    uint64_t q[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        float lo = ordered_to_float(bounds[a]), hi = ordered_to_float(bounds[3 + a]);
        float ext = hi - lo;
        float n = ext > 0.0f ? (c[a] - lo) / ext : 0.0f;
        float s = fminf(fmaxf(n * 2097152.0f, 0.0f), 2097151.0f);
        q[a] = (uint64_t)(uint32_t)s;
    }
    keys[g] = (expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]);
    vals[g] = g;
}

// ---------------------------------------------------------------------------------------------------
// 3. stable LSD radix sort, 8 bits per pass, 256 keys per block (one per thread)
// ---------------------------------------------------------------------------------------------------
static constexpr int SORT_BLOCK = 256;

__device__ __forceinline__ void block_digit_ranks(uint32_t digit, bool valid, uint32_t (*wcnt)[256], uint32_t* rankInWarp) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 8 * 256; i += SORT_BLOCK) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    uint32_t key = valid ? digit : 0xFFFFFFFFu;
    uint32_t peers = __match_any_sync(0xffffffffu, key);
    uint32_t r = __popc(peers & ((1u << lane) - 1u));
    if (valid && r == 0) wcnt[warp][digit] = __popc(peers);
    *rankInWarp = r;
    __syncthreads();
    // thread t owns digit t: exclusive prefix over the 8 warps, total left in wcnt[8-1]... stored separately by caller
}

__global__ void k_radix_hist(const uint64_t* __restrict__ keys, uint32_t N, int shift, uint32_t* __restrict__ hist,
                             uint32_t numBlocks) {
    __shared__ uint32_t wcnt[8][256];
    uint32_t i = blockIdx.x * SORT_BLOCK + threadIdx.x;
    bool valid = i < N;
    uint32_t digit = valid ? (uint32_t)((keys[i] >> shift) & 0xFF) : 0;
    uint32_t r;
    block_digit_ranks(digit, valid, wcnt, &r);
    uint32_t t = threadIdx.x, sum = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) sum += wcnt[w][t];
    hist[t * numBlocks + blockIdx.x] = sum;
}

__global__ void k_radix_scatter(const uint64_t* __restrict__ keysIn, const uint32_t* __restrict__ valsIn, uint32_t N,
                                int shift, const uint32_t* __restrict__ offsets, uint32_t numBlocks,
                                uint64_t* __restrict__ keysOut, uint32_t* __restrict__ valsOut) {
    __shared__ uint32_t wcnt[8][256];
    uint32_t i = blockIdx.x * SORT_BLOCK + threadIdx.x;
    bool valid = i < N;
    uint64_t key = valid ? keysIn[i] : 0;
    uint32_t digit = (uint32_t)((key >> shift) & 0xFF);
    uint32_t r;
    block_digit_ranks(digit, valid, wcnt, &r);
    {   // exclusive prefix over warps for digit = threadIdx.x
        uint32_t t = threadIdx.x, run = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) { uint32_t c = wcnt[w][t]; wcnt[w][t] = run; run += c; }
    }
    __syncthreads();
    if (valid) {
        uint32_t pos = offsets[digit * numBlocks + blockIdx.x] + wcnt[threadIdx.x >> 5][digit] + r;
        keysOut[pos] = key;
        valsOut[pos] = valsIn[i];
    }
}

// single-block exclusive scan of n uint32 (n up to a few million); total written to *total if not null
__global__ void k_exclusive_scan(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n, uint32_t* total) {
    __shared__ uint32_t part[1024];
    const uint32_t T = blockDim.x, t = threadIdx.x;
    const uint32_t chunk = (n + T - 1) / T;
    const uint32_t b = min(n, t * chunk), e = min(n, b + chunk);
    uint32_t s = 0;
    for (uint32_t i = b; i < e; i++) s += in[i];
    part[t] = s;
    __syncthreads();
    for (uint32_t o = 1; o < T; o <<= 1) {
        uint32_t v = (t >= o) ? part[t - o] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    uint32_t run = part[t] - s;
    if (t == T - 1 && total) *total = part[t];
    for (uint32_t i = b; i < e; i++) { uint32_t v = in[i]; out[i] = run; run += v; }
}

// ---------------------------------------------------------------------------------------------------
// 4. Karras 2012 hierarchy over the sorted keys (ties broken by position)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, int N, int i, int j) {
    if (j < 0 || j >= N) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

__global__ void k_karras(const uint64_t* __restrict__ keys, int N, uint32_t* __restrict__ childL,
                         uint32_t* __restrict__ childR, uint32_t* __restrict__ parentInt, uint32_t* __restrict__ parentLeaf,
                         uint32_t* __restrict__ rangeFirst, uint32_t* __restrict__ rangeLast) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N - 1) return;
    int d = (delta(keys, N, i, i + 1) - delta(keys, N, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, N, i, i - d);
    int lmax = 2;
    while (delta(keys, N, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, N, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, N, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, N, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    uint32_t L = (lo == gamma) ? (LEAF_FLAG | (uint32_t)gamma) : (uint32_t)gamma;
    uint32_t R = (hi == gamma + 1) ? (LEAF_FLAG | (uint32_t)(gamma + 1)) : (uint32_t)(gamma + 1);
    childL[i] = L; childR[i] = R;
    rangeFirst[i] = (uint32_t)lo; rangeLast[i] = (uint32_t)hi;
    if (L & LEAF_FLAG) parentLeaf[gamma] = (uint32_t)i; else parentInt[gamma] = (uint32_t)i;
    if (R & LEAF_FLAG) parentLeaf[gamma + 1] = (uint32_t)i; else parentInt[gamma + 1] = (uint32_t)i;
    if (i == 0) parentInt[0] = 0xFFFFFFFFu;
}

// gather triangles into sorted order and compute leaf boxes
__global__ void k_gather_sorted(const TriRecord* __restrict__ in, const uint32_t* __restrict__ order, uint32_t N,
                                TriRecord* __restrict__ out, float4* __restrict__ leafLo, float4* __restrict__ leafHi) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    TriRecord t = in[order[i]];
    out[i] = t;
    leafLo[i] = make_float4(fminf(fminf(t.v0.x, t.v1.x), t.v2.x), fminf(fminf(t.v0.y, t.v1.y), t.v2.y),
                            fminf(fminf(t.v0.z, t.v1.z), t.v2.z), 0.f);
    leafHi[i] = make_float4(fmaxf(fmaxf(t.v0.x, t.v1.x), t.v2.x), fmaxf(fmaxf(t.v0.y, t.v1.y), t.v2.y),
                            fmaxf(fmaxf(t.v0.z, t.v1.z), t.v2.z), 0.f);
}

// 5. bottom-up refit; the second thread to arrive at a node continues upwards
__global__ void k_refit(uint32_t N, const uint32_t* __restrict__ childL, const uint32_t* __restrict__ childR,
                        const uint32_t* __restrict__ parentInt, const uint32_t* __restrict__ parentLeaf,
                        const float4* __restrict__ leafLo, const float4* __restrict__ leafHi, float4* nodeLo,
                        float4* nodeHi, uint32_t* __restrict__ flags) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || N < 2) return;
    uint32_t p = parentLeaf[i];
    while (p != 0xFFFFFFFFu) {
        __threadfence();
        if (atomicAdd(&flags[p], 1u) == 0u) return;
        __threadfence();
        uint32_t L = childL[p], R = childR[p];
        // child boxes were written by other SMs: read them through L2 (ld.global.cg), never from a stale L1 line
        float4 a = (L & LEAF_FLAG) ? leafLo[L & ~LEAF_FLAG] : __ldcg(&nodeLo[L]);
        float4 ah = (L & LEAF_FLAG) ? leafHi[L & ~LEAF_FLAG] : __ldcg(&nodeHi[L]);
        float4 b = (R & LEAF_FLAG) ? leafLo[R & ~LEAF_FLAG] : __ldcg(&nodeLo[R]);
        float4 bh = (R & LEAF_FLAG) ? leafHi[R & ~LEAF_FLAG] : __ldcg(&nodeHi[R]);
        nodeLo[p] = make_float4(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z), 0.f);
        nodeHi[p] = make_float4(fmaxf(ah.x, bh.x), fmaxf(ah.y, bh.y), fmaxf(ah.z, bh.z), 0.f);
        p = parentInt[p];
    }
}

// ---------------------------------------------------------------------------------------------------
// 6. collapse: one thread per wide node of the current level
// ---------------------------------------------------------------------------------------------------
struct CollapseArrays {
    const uint32_t *childL, *childR, *rangeFirst, *rangeLast;
    const float4 *leafLo, *leafHi, *nodeLo, *nodeHi;
    WideNode* nodes;
    uint32_t* nodeInternalCount;   // per wide node
    uint32_t* nodeTriCount;        // per wide node
    uint32_t* nodeChildRefs;       // 8 per wide node: internal children (binary refs) in slot order
    uint32_t* slotTriFirst;        // 8 per wide node: first sorted-triangle index of a leaf slot
};

__device__ __forceinline__ void ref_box(const CollapseArrays& A, uint32_t ref, float lo[3], float hi[3]) {
    float4 l = (ref & LEAF_FLAG) ? A.leafLo[ref & ~LEAF_FLAG] : A.nodeLo[ref];
    float4 h = (ref & LEAF_FLAG) ? A.leafHi[ref & ~LEAF_FLAG] : A.nodeHi[ref];
    lo[0] = l.x; lo[1] = l.y; lo[2] = l.z; hi[0] = h.x; hi[1] = h.y; hi[2] = h.z;
}
__device__ __forceinline__ float box_area(const float lo[3], const float hi[3]) {
    float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return dx * dy + dy * dz + dz * dx;
}
__device__ __forceinline__ uint32_t ref_count(const CollapseArrays& A, uint32_t ref) {
    return (ref & LEAF_FLAG) ? 1u : (A.rangeLast[ref] - A.rangeFirst[ref] + 1u);
}
__device__ __forceinline__ uint32_t ref_first(const CollapseArrays& A, uint32_t ref) {
    return (ref & LEAF_FLAG) ? (ref & ~LEAF_FLAG) : A.rangeFirst[ref];
}

__global__ void k_collapse_level(CollapseArrays A, const uint32_t* __restrict__ work, uint32_t count, uint32_t levelBase) {
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= count) return;
    const uint32_t wideIdx = levelBase + w;
    uint32_t refs[8];
    float area[8];
    int n = 1;
    refs[0] = work[w];
    {
        float lo[3], hi[3]; ref_box(A, refs[0], lo, hi);
        area[0] = (refs[0] & LEAF_FLAG) ? -1.0f : box_area(lo, hi);
    }
    // SAH-guided greedy expansion: always open the internal child with the largest surface area
    while (n < 8) {
        int best = -1; float bestA = -1.0f;
        for (int j = 0; j < n; j++) if (area[j] > bestA) { bestA = area[j]; best = j; }
        if (best < 0) break;
        uint32_t b = refs[best];
        uint32_t L = A.childL[b], R = A.childR[b];
        float lo[3], hi[3];
        refs[best] = L; ref_box(A, L, lo, hi); area[best] = (L & LEAF_FLAG) ? -1.0f : box_area(lo, hi);
        refs[n] = R;    ref_box(A, R, lo, hi); area[n] = (R & LEAF_FLAG) ? -1.0f : box_area(lo, hi);
        n++;
    }
    // node bounds = union of children
    float clo[8][3], chi[8][3];
    float nlo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, nhi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int j = 0; j < n; j++) {
        ref_box(A, refs[j], clo[j], chi[j]);
        for (int a = 0; a < 3; a++) { nlo[a] = fminf(nlo[a], clo[j][a]); nhi[a] = fmaxf(nhi[a], chi[j][a]); }
    }
    // octant-ordered slot assignment (greedy maximum of the 8x8 affinity table)
    int slotOf[8]; bool slotUsed[8], childDone[8];
    for (int j = 0; j < 8; j++) { slotOf[j] = -1; slotUsed[j] = false; childDone[j] = false; }
    float cx[8], cy[8], cz[8];
    for (int j = 0; j < n; j++) {
        cx[j] = 0.5f * (clo[j][0] + chi[j][0]) - 0.5f * (nlo[0] + nhi[0]);
        cy[j] = 0.5f * (clo[j][1] + chi[j][1]) - 0.5f * (nlo[1] + nhi[1]);
        cz[j] = 0.5f * (clo[j][2] + chi[j][2]) - 0.5f * (nlo[2] + nhi[2]);
    }
    for (int it = 0; it < n; it++) {
        float bestC = -3.0e38f; int bj = -1, bs = -1;
        for (int j = 0; j < n; j++) {
            if (childDone[j]) continue;
            for (int s = 0; s < 8; s++) {
                if (slotUsed[s]) continue;
                float c = ((s & 4) ? cx[j] : -cx[j]) + ((s & 2) ? cy[j] : -cy[j]) + ((s & 1) ? cz[j] : -cz[j]);
                if (c > bestC) { bestC = c; bj = j; bs = s; }
            }
        }
        slotOf[bj] = bs; slotUsed[bs] = true; childDone[bj] = true;
    }
    int childAt[8];
    for (int s = 0; s < 8; s++) childAt[s] = -1;
    for (int j = 0; j < n; j++) childAt[slotOf[j]] = j;

    // quantisation grid: smallest power-of-two step e with lo + 255 * 2^e >= hi (checked in fp64)
    WideNode node;
    memset(&node, 0, sizeof(node));
    node.px = nlo[0]; node.py = nlo[1]; node.pz = nlo[2];
    uint8_t eb[3];
    double step[3];
    for (int a = 0; a < 3; a++) {
        double ext = (double)nhi[a] - (double)nlo[a];
        int e = 1;   // biased exponent; 2^(e-127)
        if (ext > 0.0) {
            float q = (float)(ext / 255.0);
            e = (int)((__float_as_uint(q) >> 23) & 0xFF);
            if (e < 1) e = 1;
            while (255.0 * exp2((double)(e - 127)) < ext) e++;
        }
        eb[a] = (uint8_t)e;
        step[a] = exp2((double)(e - 127));
    }
    node.ex = eb[0]; node.ey = eb[1]; node.ez = eb[2];

    uint32_t nInternal = 0, triOffset = 0;
    for (int s = 0; s < 8; s++) {
        int j = childAt[s];
        A.slotTriFirst[wideIdx * 8 + s] = 0;
        if (j < 0) continue;
        uint8_t q[6];
        for (int a = 0; a < 3; a++) {
            double l = floor(((double)clo[j][a] - (double)nlo[a]) / step[a]);
            double h = ceil(((double)chi[j][a] - (double)nlo[a]) / step[a]);
            q[a] = (uint8_t)fmin(fmax(l, 0.0), 255.0);
            q[3 + a] = (uint8_t)fmin(fmax(h, 0.0), 255.0);
        }
        node.qlox[s] = q[0]; node.qloy[s] = q[1]; node.qloz[s] = q[2];
        node.qhix[s] = q[3]; node.qhiy[s] = q[4]; node.qhiz[s] = q[5];
        uint32_t cnt = ref_count(A, refs[j]);
        if ((refs[j] & LEAF_FLAG) || cnt <= (uint32_t)MAX_LEAF_TRIS) {
            uint32_t unary = (1u << cnt) - 1u;      // 1 -> 001, 2 -> 011, 3 -> 111
            node.meta[s] = (uint8_t)((unary << 5) | triOffset);
            A.slotTriFirst[wideIdx * 8 + s] = ref_first(A, refs[j]);
            triOffset += cnt;
        } else {
            node.imask |= (uint8_t)(1u << s);
            node.meta[s] = (uint8_t)((1u << 5) | (24u + (uint32_t)s));
            A.nodeChildRefs[wideIdx * 8 + nInternal] = refs[j];
            nInternal++;
        }
    }
    A.nodes[wideIdx] = node;
    A.nodeInternalCount[wideIdx] = nInternal;
    A.nodeTriCount[wideIdx] = triOffset;
}

__global__ void k_link_children(WideNode* nodes, const uint32_t* __restrict__ prefix, const uint32_t* __restrict__ internalCount,
                                const uint32_t* __restrict__ childRefs, uint32_t count, uint32_t levelBase, uint32_t nextBase,
                                uint32_t* __restrict__ nextWork) {
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= count) return;
    uint32_t wideIdx = levelBase + w;
    uint32_t p = prefix[w];
    nodes[wideIdx].childBase = nextBase + p;
    uint32_t c = internalCount[wideIdx];
    for (uint32_t k = 0; k < c; k++) nextWork[p + k] = childRefs[wideIdx * 8 + k];
}

// 7. final triangle placement: node.triBase from a scan over all wide nodes, triangles copied in slot order
__global__ void k_place_triangles(WideNode* nodes, uint32_t numNodes, const uint32_t* __restrict__ triPrefix,
                                  const uint32_t* __restrict__ slotTriFirst, const TriRecord* __restrict__ sorted,
                                  TriRecord* __restrict__ out) {
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= numNodes) return;
    uint32_t base = triPrefix[w];
    nodes[w].triBase = base;
    for (int s = 0; s < 8; s++) {
        uint8_t m = nodes[w].meta[s];
        if (m == 0 || (nodes[w].imask >> s) & 1) continue;
        uint32_t cnt = __popc((uint32_t)(m >> 5));
        uint32_t off = m & 31u;
        uint32_t first = slotTriFirst[w * 8 + s];
        for (uint32_t k = 0; k < cnt; k++) out[base + off + k] = sorted[first + k];
    }
}

// ---------------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------------
template <class T> static cudaError_t dalloc(T** p, size_t n) { return cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)); }

int build_bvh(const BuildInput& in, cudaStream_t stream, Bvh* out, uint64_t* launches) {
    const std::vector<RB200Instance>& hi = *in.h_instances;
    std::vector<uint32_t> prefix(hi.size() + 1, 0);
    for (size_t i = 0; i < hi.size(); i++) prefix[i + 1] = prefix[i] + hi[i].triangleCount;
    const uint32_t N = prefix.back();
    if (N == 0) { set_error("scene has no triangles"); return RB200_ERR_INVALID_ARGUMENT; }
    if (N >= 0x40000000u) { set_error("too many triangles"); return RB200_ERR_INVALID_ARGUMENT; }

    cudaEvent_t e0, e1;
    RB_CUDA(cudaEventCreate(&e0)); RB_CUDA(cudaEventCreate(&e1));
    RB_CUDA(cudaEventRecord(e0, stream));
    uint64_t nl = 0;

    uint32_t *dPrefix, *dBounds, *vals[2], *hist, *histScan, *childL, *childR, *parentInt, *parentLeaf, *rFirst, *rLast, *flags;
    uint64_t* keys[2];
    TriRecord *unsorted, *sorted;
    float4 *leafLo, *leafHi, *nodeLo, *nodeHi;
    RB_CUDA(dalloc(&dPrefix, prefix.size()));
    RB_CUDA(cudaMemcpyAsync(dPrefix, prefix.data(), prefix.size() * 4, cudaMemcpyHostToDevice, stream));
    RB_CUDA(dalloc(&dBounds, 6));
    uint32_t initB[6] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0, 0, 0};
    RB_CUDA(cudaMemcpyAsync(dBounds, initB, sizeof(initB), cudaMemcpyHostToDevice, stream));
    RB_CUDA(dalloc(&unsorted, N)); RB_CUDA(dalloc(&sorted, N));
    RB_CUDA(dalloc(&keys[0], N)); RB_CUDA(dalloc(&keys[1], N)); RB_CUDA(dalloc(&vals[0], N)); RB_CUDA(dalloc(&vals[1], N));
    const uint32_t sortBlocks = (N + SORT_BLOCK - 1) / SORT_BLOCK;
    RB_CUDA(dalloc(&hist, (size_t)256 * sortBlocks)); RB_CUDA(dalloc(&histScan, (size_t)256 * sortBlocks));
    RB_CUDA(dalloc(&childL, N)); RB_CUDA(dalloc(&childR, N)); RB_CUDA(dalloc(&parentInt, N)); RB_CUDA(dalloc(&parentLeaf, N));
    RB_CUDA(dalloc(&rFirst, N)); RB_CUDA(dalloc(&rLast, N)); RB_CUDA(dalloc(&flags, N));
    RB_CUDA(dalloc(&leafLo, N)); RB_CUDA(dalloc(&leafHi, N)); RB_CUDA(dalloc(&nodeLo, N)); RB_CUDA(dalloc(&nodeHi, N));
    RB_CUDA(cudaMemsetAsync(flags, 0, (size_t)N * 4, stream));
    RB_CUDA(cudaMemsetAsync(parentLeaf, 0xFF, (size_t)N * 4, stream));

    const uint32_t B = 256, G = (N + B - 1) / B;
    k_flatten<<<G, B, 0, stream>>>(in.vertices, in.indices, in.d_instances, dPrefix, in.numInstances, N, unsorted, dBounds); nl++;
    k_morton<<<G, B, 0, stream>>>(unsorted, N, dBounds, keys[0], vals[0]); nl++;
    int cur = 0;
    for (int pass = 0; pass < 8; pass++) {
        k_radix_hist<<<sortBlocks, SORT_BLOCK, 0, stream>>>(keys[cur], N, pass * 8, hist, sortBlocks); nl++;
        k_exclusive_scan<<<1, 1024, 0, stream>>>(hist, histScan, 256u * sortBlocks, nullptr); nl++;
        k_radix_scatter<<<sortBlocks, SORT_BLOCK, 0, stream>>>(keys[cur], vals[cur], N, pass * 8, histScan, sortBlocks,
                                                               keys[cur ^ 1], vals[cur ^ 1]); nl++;
        cur ^= 1;
    }
    k_gather_sorted<<<G, B, 0, stream>>>(unsorted, vals[cur], N, sorted, leafLo, leafHi); nl++;
    if (N > 1) {
        k_karras<<<G, B, 0, stream>>>(keys[cur], (int)N, childL, childR, parentInt, parentLeaf, rFirst, rLast); nl++;
        k_refit<<<G, B, 0, stream>>>(N, childL, childR, parentInt, parentLeaf, leafLo, leafHi, nodeLo, nodeHi, flags); nl++;
    }
    RB_CUDA(cudaGetLastError());

    // collapse, level by level
    WideNode* nodesTmp; uint32_t *nodeInternalCount, *nodeTriCount, *nodeChildRefs, *slotTriFirst, *work[2], *levelPrefix, *dTotal;
    const size_t maxNodes = N;   // every wide node has >= 1 leaf slot or >= 2 children: never more nodes than triangles
    RB_CUDA(dalloc(&nodesTmp, maxNodes)); RB_CUDA(dalloc(&nodeInternalCount, maxNodes)); RB_CUDA(dalloc(&nodeTriCount, maxNodes));
    RB_CUDA(dalloc(&nodeChildRefs, maxNodes * 8)); RB_CUDA(dalloc(&slotTriFirst, maxNodes * 8));
    RB_CUDA(dalloc(&work[0], maxNodes)); RB_CUDA(dalloc(&work[1], maxNodes)); RB_CUDA(dalloc(&levelPrefix, maxNodes));
    RB_CUDA(dalloc(&dTotal, 1));
    CollapseArrays A{childL, childR, rFirst, rLast, leafLo, leafHi, nodeLo, nodeHi, nodesTmp, nodeInternalCount, nodeTriCount,
                     nodeChildRefs, slotTriFirst};
    uint32_t rootRef = (N > 1) ? 0u : (LEAF_FLAG | 0u);
    RB_CUDA(cudaMemcpyAsync(work[0], &rootRef, 4, cudaMemcpyHostToDevice, stream));
    uint32_t levelBase = 0, levelCount = 1, depth = 0;
    int wcur = 0;
    while (levelCount > 0) {
        if ((size_t)levelBase + levelCount > maxNodes) { set_error("internal: wide node overflow"); return RB200_ERR_CUDA; }
        uint32_t g = (levelCount + 127) / 128;
        k_collapse_level<<<g, 128, 0, stream>>>(A, work[wcur], levelCount, levelBase); nl++;
        k_exclusive_scan<<<1, 1024, 0, stream>>>(nodeInternalCount + levelBase, levelPrefix, levelCount, dTotal); nl++;
        uint32_t nextBase = levelBase + levelCount;
        k_link_children<<<g, 128, 0, stream>>>(nodesTmp, levelPrefix, nodeInternalCount, nodeChildRefs, levelCount, levelBase,
                                               nextBase, work[wcur ^ 1]); nl++;
        uint32_t total = 0;
        RB_CUDA(cudaMemcpyAsync(&total, dTotal, 4, cudaMemcpyDeviceToHost, stream));
        RB_CUDA(cudaStreamSynchronize(stream));
        levelBase = nextBase; levelCount = total; wcur ^= 1; depth++;
    }
    const uint32_t numNodes = levelBase;

    // triangles into leaf order
    uint32_t* triPrefix;
    RB_CUDA(dalloc(&triPrefix, numNodes));
    k_exclusive_scan<<<1, 1024, 0, stream>>>(nodeTriCount, triPrefix, numNodes, dTotal); nl++;
    // nodes and triangles live in ONE allocation so that a single L2 access-policy window can keep the whole
    // hierarchy resident (see rb200_scene_create)
    const size_t nodeBytes = ((size_t)numNodes * sizeof(WideNode) + 255) & ~(size_t)255;
    out->blobBytes = nodeBytes + (size_t)N * sizeof(TriRecord);
    RB_CUDA(cudaMalloc(&out->blob, out->blobBytes));
    out->nodes = reinterpret_cast<WideNode*>(out->blob);
    out->tris = reinterpret_cast<TriRecord*>(reinterpret_cast<char*>(out->blob) + nodeBytes);
    k_place_triangles<<<(numNodes + 127) / 128, 128, 0, stream>>>(nodesTmp, numNodes, triPrefix, slotTriFirst, sorted, out->tris); nl++;
    RB_CUDA(cudaMemcpyAsync(out->nodes, nodesTmp, (size_t)numNodes * sizeof(WideNode), cudaMemcpyDeviceToDevice, stream));
    uint32_t totalTris = 0, hb[6];
    RB_CUDA(cudaMemcpyAsync(&totalTris, dTotal, 4, cudaMemcpyDeviceToHost, stream));
    RB_CUDA(cudaMemcpyAsync(hb, dBounds, sizeof(hb), cudaMemcpyDeviceToHost, stream));
    RB_CUDA(cudaEventRecord(e1, stream));
    RB_CUDA(cudaStreamSynchronize(stream));
    RB_CUDA(cudaGetLastError());
    if (totalTris != N) { set_error("internal: BVH references %u of %u triangles", totalTris, N); return RB200_ERR_CUDA; }
    RB_CUDA(cudaEventElapsedTime(&out->buildMs, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    out->numNodes = numNodes; out->numTris = N; out->maxDepth = depth;
    for (int a = 0; a < 3; a++) { out->sceneMin[a] = ordered_to_float(hb[a]); out->sceneMax[a] = ordered_to_float(hb[3 + a]); }

    void* frees[] = {dPrefix, dBounds, unsorted, sorted, keys[0], keys[1], vals[0], vals[1], hist, histScan, childL, childR,
                     parentInt, parentLeaf, rFirst, rLast, flags, leafLo, leafHi, nodeLo, nodeHi, nodesTmp, nodeInternalCount,
                     nodeTriCount, nodeChildRefs, slotTriFirst, work[0], work[1], levelPrefix, dTotal, triPrefix};
    for (void* p : frees) cudaFree(p);
    if (launches) *launches += nl;
    return RB200_OK;
}

int hash_bvh(const Bvh& bvh, cudaStream_t stream, uint64_t* hash) {
    std::vector<uint8_t> buf((size_t)bvh.numNodes * sizeof(WideNode) + (size_t)bvh.numTris * sizeof(TriRecord));
    RB_CUDA(cudaMemcpyAsync(buf.data(), bvh.nodes, (size_t)bvh.numNodes * sizeof(WideNode), cudaMemcpyDeviceToHost, stream));
    RB_CUDA(cudaMemcpyAsync(buf.data() + (size_t)bvh.numNodes * sizeof(WideNode), bvh.tris,
                            (size_t)bvh.numTris * sizeof(TriRecord), cudaMemcpyDeviceToHost, stream));
    RB_CUDA(cudaStreamSynchronize(stream));
    uint64_t h = 1469598103934665603ull;
    for (uint8_t b : buf) { h ^= b; h *= 1099511628211ull; }
    *hash = h;
    return RB200_OK;
}

void free_bvh(Bvh* b) {
    if (b->blob) cudaFree(b->blob);
    b->blob = nullptr; b->nodes = nullptr; b->tris = nullptr;
}

} // namespace rb200
