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
        t.v1 = make_float4(w[1].x, w[1].y, w[1].z, __uint_as_float(a | ((in.materialIdx > 3u ? 3u : in.materialIdx) << 30)));
        t.v2 = make_float4(w[2].x, w[2].y, w[2].z, __uint_as_float(g));
#if RB_WIDE_LOADS
        t.pad = make_float4(0.f, 0.f, 0.f, 0.f);
#endif
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

// Premade records (BuildInput::premade: the instance boxes of the two-level mode's top level) instead of flattening:
// copy, zero the padding, scene bounds as above.
__global__ void k_premade(const TriRecord* __restrict__ in, uint32_t N, TriRecord* __restrict__ out, uint32_t* __restrict__ bounds) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    if (g < N) {
        TriRecord t = in[g];
#if RB_WIDE_LOADS
        t.pad = make_float4(0.f, 0.f, 0.f, 0.f);
#endif
        const float4 v[3] = {t.v0, t.v1, t.v2};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lo[0] = fminf(lo[0], v[k].x); lo[1] = fminf(lo[1], v[k].y); lo[2] = fminf(lo[2], v[k].z);
            hi[0] = fmaxf(hi[0], v[k].x); hi[1] = fmaxf(hi[1], v[k].y); hi[2] = fmaxf(hi[2], v[k].z);
        }
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

// Exclusive scans (integer sums: the result does not depend on the order of evaluation).
// k_scan_small: one block, any n (each thread takes a contiguous chunk); safe in place; total to *total if not null.
template <class T>
__global__ void k_scan_small(const T* in, T* out, uint32_t n, T* total) {
    __shared__ T part[1024];
    const uint32_t nt = blockDim.x, t = threadIdx.x;
    const uint32_t chunk = (n + nt - 1) / nt;
    const uint32_t b = min(n, t * chunk), e = min(n, b + chunk);
    T s = 0;
    for (uint32_t i = b; i < e; i++) s += in[i];
    part[t] = s;
    __syncthreads();
    for (uint32_t o = 1; o < nt; o <<= 1) {
        T v = (t >= o) ? part[t - o] : T(0);
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    T run = part[t] - s;
    if (t == nt - 1 && total) *total = part[t];
    for (uint32_t i = b; i < e; i++) { T v = in[i]; out[i] = run; run += v; }
}

// large n: scan tiles of SCAN_TILE items independently, scan the tile sums with k_scan_small, add the offsets back
static constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <class T>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const T* __restrict__ in, T* __restrict__ out, uint32_t n,
                                                              T* __restrict__ tileSums) {
    __shared__ T warpSums[SCAN_THREADS / 32];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    T v[SCAN_ITEMS];
    T sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < n) ? in[base + k] : T(0); sum += v[k]; }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    T inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { T up = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += up; }
    if (lane == 31) warpSums[w] = inc;
    __syncthreads();
    T offset = 0;
    for (int k = 0; k < w; k++) offset += warpSums[k];
    T run = offset + inc - sum;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) out[base + k] = run; run += v[k]; }
    if (threadIdx.x == SCAN_THREADS - 1) tileSums[blockIdx.x] = run;
}

template <class T>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(T* __restrict__ out, uint32_t n, const T* __restrict__ tileOffsets) {
    const T off = tileOffsets[blockIdx.x];
    const uint32_t base = blockIdx.x * SCAN_TILE;
    for (uint32_t k = threadIdx.x; k < (uint32_t)SCAN_TILE; k += SCAN_THREADS)
        if (base + k < n) out[base + k] += off;
}

// tileScratch: at least (n + SCAN_TILE - 1) / SCAN_TILE entries
template <class T>
static void exclusive_scan(const T* in, T* out, uint32_t n, T* total, T* tileScratch, cudaStream_t stream, uint64_t& nl) {
    if (n <= 2u * SCAN_TILE) { k_scan_small<T><<<1, 1024, 0, stream>>>(in, out, n, total); nl++; return; }
    const uint32_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    k_scan_tiles<T><<<tiles, SCAN_THREADS, 0, stream>>>(in, out, n, tileScratch);
    k_scan_small<T><<<1, 1024, 0, stream>>>(tileScratch, tileScratch, tiles, total);
    k_scan_add<T><<<tiles, SCAN_THREADS, 0, stream>>>(out, n, tileScratch);
    nl += 3;
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
// 4b. PLOC (Meister & Bittner 2018, parallel locally-ordered clustering) as the alternative to Karras + refit:
// clusters start as the Morton-sorted leaves; every round each cluster looks RB_PLOC_RADIUS positions to either side
// for the neighbour whose union with it has the smallest surface area, mutually-nearest pairs merge into a new
// binary node, the array is compacted (order kept) and the round repeats until one cluster is left. Node ids and
// positions come from prefix sums and ties are broken by index, so the tree is the same on every run.
// ---------------------------------------------------------------------------------------------------
#ifndef RB_PLOC_RADIUS
#define RB_PLOC_RADIUS 16
#endif
static constexpr int PLOC_BLOCK = 256;
static constexpr uint32_t NONE = 0xFFFFFFFFu;

__global__ void __launch_bounds__(PLOC_BLOCK) k_ploc_nearest(uint32_t c, const float4* __restrict__ cLo,
                                                              const float4* __restrict__ cHi, uint32_t* __restrict__ nn) {
    constexpr int R = RB_PLOC_RADIUS;
    __shared__ float4 sLo[PLOC_BLOCK + 2 * R], sHi[PLOC_BLOCK + 2 * R];
    const int base = (int)(blockIdx.x * PLOC_BLOCK) - R;
    for (int t = threadIdx.x; t < PLOC_BLOCK + 2 * R; t += PLOC_BLOCK) {
        const int g = base + t;
        if (g >= 0 && g < (int)c) { sLo[t] = cLo[g]; sHi[t] = cHi[g]; }
    }
    __syncthreads();
    const int i = blockIdx.x * PLOC_BLOCK + threadIdx.x;
    if (i >= (int)c) return;
    const float4 lo = sLo[threadIdx.x + R], hi = sHi[threadIdx.x + R];
    float best = 3.0e38f;
    int bj = -1;
    uint32_t bx = 0xFFFFFFFFu;
    for (int d = -R; d <= R; d++) {
        const int j = i + d;
        if (d == 0 || j < 0 || j >= (int)c) continue;
        const float4 l2 = sLo[threadIdx.x + R + d], h2 = sHi[threadIdx.x + R + d];
        const float dx = fmaxf(hi.x, h2.x) - fminf(lo.x, l2.x), dy = fmaxf(hi.y, h2.y) - fminf(lo.y, l2.y),
                    dz = fmaxf(hi.z, h2.z) - fminf(lo.z, l2.z);
        const float a = dx * dy + dy * dz + dz * dx;
        // equal areas (regular grids) go to the partner with the smaller i ^ j: neighbours then pair up (0,1), (2,3), ...
        // instead of forming one long chain with a single mutual pair per round; the rule is symmetric in i and j, so
        // the pair that minimises (area, i ^ j) overall is always mutual and every round makes progress
        const uint32_t x = (uint32_t)i ^ (uint32_t)j;
        if (a < best || (a == best && x < bx)) { best = a; bj = j; bx = x; }
    }
    nn[i] = (uint32_t)bj;
}

// Safety net for inputs on which nearest-neighbour chains leave almost nothing mutual (e.g. a row of boxes of steadily
// growing size: every cluster prefers its smaller neighbour and only the first pair is mutual): pair by position.
__global__ void k_ploc_pair_adjacent(uint32_t c, uint32_t* __restrict__ nn) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c) nn[i] = (i ^ 1u) < c ? (i ^ 1u) : NONE;
}

// per cluster: high word 1 = survives this round (not the right half of a merging pair), low word 1 = starts a merge
__global__ void k_ploc_flags(uint32_t c, const uint32_t* __restrict__ nn, unsigned long long* __restrict__ flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    const uint32_t j = nn[i];
    const bool mutual = j < c && nn[j] == i;
    const unsigned long long valid = (mutual && i > j) ? 0ull : 1ull, merge = (mutual && i < j) ? 1ull : 0ull;
    flags[i] = (valid << 32) | merge;
}

__global__ void k_ploc_apply(uint32_t c, const uint32_t* __restrict__ nn, const unsigned long long* __restrict__ scan,
                             const uint32_t* __restrict__ clIn, const float4* __restrict__ loIn, const float4* __restrict__ hiIn,
                             uint32_t* __restrict__ clOut, float4* __restrict__ loOut, float4* __restrict__ hiOut,
                             uint32_t nodeBase, uint32_t* __restrict__ childL, uint32_t* __restrict__ childR,
                             float4* __restrict__ nodeLo, float4* __restrict__ nodeHi, uint32_t* __restrict__ count,
                             uint32_t* __restrict__ parentInt, uint32_t* __restrict__ parentLeaf) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    const uint32_t j = nn[i];
    const bool mutual = j < c && nn[j] == i;
    if (mutual && i > j) return;
    const unsigned long long sc = scan[i];
    const uint32_t pos = (uint32_t)(sc >> 32);
    if (!mutual) { clOut[pos] = clIn[i]; loOut[pos] = loIn[i]; hiOut[pos] = hiIn[i]; return; }
    const uint32_t id = nodeBase + (uint32_t)(sc & 0xFFFFFFFFull);
    const uint32_t L = clIn[i], R = clIn[j];
    const float4 a = loIn[i], b = loIn[j], ah = hiIn[i], bh = hiIn[j];
    const float4 lo = make_float4(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z), 0.f);
    const float4 hi = make_float4(fmaxf(ah.x, bh.x), fmaxf(ah.y, bh.y), fmaxf(ah.z, bh.z), 0.f);
    childL[id] = L; childR[id] = R;
    nodeLo[id] = lo; nodeHi[id] = hi;
    count[id] = ((L & LEAF_FLAG) ? 1u : count[L]) + ((R & LEAF_FLAG) ? 1u : count[R]);
    if (L & LEAF_FLAG) parentLeaf[L & ~LEAF_FLAG] = id; else parentInt[L] = id;
    if (R & LEAF_FLAG) parentLeaf[R & ~LEAF_FLAG] = id; else parentInt[R] = id;
    parentInt[id] = NONE;
    clOut[pos] = id; loOut[pos] = lo; hiOut[pos] = hi;
}

__global__ void k_ploc_init(uint32_t N, uint32_t* __restrict__ cl) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) cl[i] = LEAF_FLAG | i;
}

// depth-first position of every leaf and first/last leaf of every internal node: the number of leaves in the left
// siblings met on the way up to the root
__global__ void k_ploc_ranges(uint32_t N, const uint32_t* __restrict__ childL, const uint32_t* __restrict__ parentInt,
                              const uint32_t* __restrict__ parentLeaf, const uint32_t* __restrict__ count,
                              uint32_t* __restrict__ rangeFirst, uint32_t* __restrict__ rangeLast, uint32_t* __restrict__ leafPos) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2u * N - 1u) return;
    const bool leaf = t < N;
    uint32_t cur = leaf ? (LEAF_FLAG | t) : (t - N);
    uint32_t p = leaf ? parentLeaf[t] : parentInt[t - N];
    uint32_t first = 0;
    while (p != NONE) {
        const uint32_t L = childL[p];
        if (L != cur) first += (L & LEAF_FLAG) ? 1u : count[L];
        cur = p;
        p = parentInt[p];
    }
    if (leaf) leafPos[t] = first;
    else { rangeFirst[t - N] = first; rangeLast[t - N] = first + count[t - N] - 1u; }
}

__global__ void k_ploc_permute(uint32_t N, const uint32_t* __restrict__ leafPos, const TriRecord* __restrict__ triIn,
                               const float4* __restrict__ loIn, const float4* __restrict__ hiIn, TriRecord* __restrict__ triOut,
                               float4* __restrict__ loOut, float4* __restrict__ hiOut, uint32_t* childL, uint32_t* childR) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) {
        const uint32_t p = leafPos[i];
        triOut[p] = triIn[i]; loOut[p] = loIn[i]; hiOut[p] = hiIn[i];
    }
    if (i + 1 < N) {      // internal node i: leaf references follow their leaves
        const uint32_t L = childL[i], R = childR[i];
        if (L & LEAF_FLAG) childL[i] = LEAF_FLAG | leafPos[L & ~LEAF_FLAG];
        if (R & LEAF_FLAG) childR[i] = LEAF_FLAG | leafPos[R & ~LEAF_FLAG];
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

// 6a. SAH-optimal collapse (Ylitie, Karras, Laine 2017, section 3.2): for every binary node n and every i in 1..7,
// cost[n][i] is the cheapest way to turn n's subtree into at most i children of a wide node:
//     cost[n][1] = min( area(n) * P(n) * c_prim            (one leaf slot, only if P(n) <= MAX_LEAF_TRIS),
//                       dist(n, 8) + area(n) * c_node )    (one wide node with up to 8 children)
//     cost[n][i] = min( dist(n, i), cost[n][i-1] ),   dist(n, j) = min over 0 < k < j of cost[left][k] + cost[right][j-k]
// computed bottom-up (the second thread to reach a node evaluates it, as in the refit), with the arg-min of every
// choice kept in dec[n] so that the level-by-level pass below can unfold the decisions. dec[n].x byte i-1 (i = 2..7):
// 0 = "same as i-1", else k; byte 0: 1 = wide node, 0 = leaf slot; byte 7: the k of dist(n, 8).
#ifndef RB_COLLAPSE_DP
#define RB_COLLAPSE_DP 1
#endif
#ifndef RB_COST_NODE
#define RB_COST_NODE 1.0f
#endif
#ifndef RB_COST_PRIM
#define RB_COST_PRIM 0.3f
#endif

__device__ __forceinline__ void ref_costs(const CollapseArrays& A, const float* __restrict__ cost, uint32_t ref, float c[8]) {
    if (ref & LEAF_FLAG) {
        float lo[3], hi[3]; ref_box(A, ref, lo, hi);
        const float v = box_area(lo, hi) * RB_COST_PRIM;
#pragma unroll
        for (int i = 1; i <= 7; i++) c[i] = v;
    } else {
        const float4 a = __ldcg(reinterpret_cast<const float4*>(cost + (size_t)ref * 8));
        const float4 b = __ldcg(reinterpret_cast<const float4*>(cost + (size_t)ref * 8) + 1);
        c[1] = a.x; c[2] = a.y; c[3] = a.z; c[4] = a.w; c[5] = b.x; c[6] = b.y; c[7] = b.z;
    }
}

__global__ void k_collapse_cost(uint32_t N, CollapseArrays A, const uint32_t* __restrict__ parentInt,
                                const uint32_t* __restrict__ parentLeaf, uint32_t* __restrict__ flags, float* cost, uint2* dec) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || N < 2) return;
    uint32_t p = parentLeaf[i];
    while (p != 0xFFFFFFFFu) {
        __threadfence();
        if (atomicAdd(&flags[p], 1u) == 0u) return;
        __threadfence();
        float cl[8], cr[8];
        ref_costs(A, cost, A.childL[p], cl);
        ref_costs(A, cost, A.childR[p], cr);
        float lo[3], hi[3]; ref_box(A, p, lo, hi);
        const float area = box_area(lo, hi);
        const uint32_t P = A.rangeLast[p] - A.rangeFirst[p] + 1u;
        float dist[9]; uint32_t kk[9];
#pragma unroll
        for (int j = 2; j <= 8; j++) {
            float best = 3.0e38f; uint32_t bk = 1;
#pragma unroll
            for (int k = 1; k < j; k++) {
                if (k > 7 || j - k > 7) continue;
                const float v = cl[k] + cr[j - k];
                if (v < best) { best = v; bk = (uint32_t)k; }
            }
            dist[j] = best; kk[j] = bk;
        }
        const float cLeaf = P <= (uint32_t)MAX_LEAF_TRIS ? area * (float)P * RB_COST_PRIM : 3.0e38f;
        const float cInt = dist[8] + area * RB_COST_NODE;
        float c[8];
        uint32_t d0 = 0u, d1 = 0u;
        c[1] = fminf(cLeaf, cInt);
        if (cInt < cLeaf) d0 |= 1u;
#pragma unroll
        for (int j = 2; j <= 7; j++) {
            if (dist[j] < c[j - 1]) { c[j] = dist[j]; if (j <= 4) d0 |= kk[j] << (8 * (j - 1)); else d1 |= kk[j] << (8 * (j - 5)); }
            else c[j] = c[j - 1];
        }
        d1 |= kk[8] << 24;
        float4* out = reinterpret_cast<float4*>(cost + (size_t)p * 8);
        out[0] = make_float4(c[1], c[2], c[3], c[4]);
        out[1] = make_float4(c[5], c[6], c[7], 0.f);
        dec[p] = make_uint2(d0, d1);
        p = parentInt[p];
    }
}

__device__ __forceinline__ uint32_t dec_byte(uint2 d, int i) { return ((i < 4 ? d.x : d.y) >> (8 * (i & 3))) & 0xFFu; }

__global__ void k_collapse_level(CollapseArrays A, const uint2* __restrict__ dec, const uint32_t* __restrict__ work, uint32_t count,
                                 uint32_t levelBase) {
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= count) return;
    const uint32_t wideIdx = levelBase + w;
    uint32_t refs[8];
    bool leafSlot[8];
    int n = 0;
#if RB_COLLAPSE_DP
    {
        // unfold the decisions of k_collapse_cost for the wide node rooted at work[w]: children come out left to right
        const uint32_t root = work[w];
        uint32_t sref[16]; int sbud[16]; int top = 0;
        if (root & LEAF_FLAG) { refs[0] = root; leafSlot[0] = true; n = 1; }
        else {
            const uint32_t k8 = dec_byte(dec[root], 7);
            sref[top] = A.childR[root]; sbud[top++] = 8 - (int)k8;
            sref[top] = A.childL[root]; sbud[top++] = (int)k8;
        }
        while (top > 0) {
            const uint32_t ref = sref[--top];
            int j = sbud[top];
            if (ref & LEAF_FLAG) { refs[n] = ref; leafSlot[n++] = true; continue; }
            const uint2 d = dec[ref];
            while (j > 1 && dec_byte(d, j - 1) == 0u) j--;
            if (j == 1) { refs[n] = ref; leafSlot[n++] = (d.x & 1u) == 0u; continue; }
            const int k = (int)dec_byte(d, j - 1);
            sref[top] = A.childR[ref]; sbud[top++] = j - k;
            sref[top] = A.childL[ref]; sbud[top++] = k;
        }
    }
#else
    float area[8];
    n = 1;
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
    for (int j = 0; j < n; j++) leafSlot[j] = (refs[j] & LEAF_FLAG) || ref_count(A, refs[j]) <= (uint32_t)MAX_LEAF_TRIS;
#endif
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
        if (leafSlot[j]) {
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
// All build temporaries come out of one allocation (about 600 bytes per triangle): ~50 cudaMalloc / cudaFree calls
// cost several milliseconds of host time, more than the kernels of the build.
struct Arena {
    char* base = nullptr;
    size_t cap = 0, used = 0;
    ~Arena() { if (base) cudaFree(base); }
};
static thread_local Arena* g_arena = nullptr;
template <class T> static cudaError_t dalloc(T** p, size_t n) {
    const size_t bytes = (std::max<size_t>(n, 1) * sizeof(T) + 255) & ~(size_t)255;
    if (!g_arena || g_arena->used + bytes > g_arena->cap) return cudaErrorMemoryAllocation;
    *p = reinterpret_cast<T*>(g_arena->base + g_arena->used);
    g_arena->used += bytes;
    return cudaSuccess;
}

int build_bvh(const BuildInput& in, cudaStream_t stream, Bvh* out, uint64_t* launches) {
    const std::vector<RB200Instance>& hi = *in.h_instances;
    std::vector<uint32_t> prefix(hi.size() + 1, 0);
    for (size_t i = 0; i < hi.size(); i++) prefix[i + 1] = prefix[i] + hi[i].triangleCount;
    const uint32_t N = in.premade ? in.numPremade : prefix.back();
    if (N == 0) { set_error("scene has no triangles"); return RB200_ERR_INVALID_ARGUMENT; }
    if (N >= 0x08000000u) { set_error("too many triangles (limit 2^27: the traversal packs triangle index and lane into 32 bits)"); return RB200_ERR_INVALID_ARGUMENT; }
    if (hi.size() > (size_t)TRI_INST_MASK) { set_error("too many instances (limit 2^30)"); return RB200_ERR_INVALID_ARGUMENT; }

    Arena arena;
    arena.cap = (size_t)N * 704 + prefix.size() * 4 + (1u << 20);
    if (cudaMalloc((void**)&arena.base, arena.cap) != cudaSuccess) {
        cudaGetLastError(); arena.base = nullptr;
        set_error("out of device memory for the BVH build (%zu MiB of temporaries)", arena.cap >> 20);
        return RB200_ERR_OUT_OF_MEMORY;
    }
    g_arena = &arena;
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
    uint32_t* scanScratch;
    RB_CUDA(dalloc(&scanScratch, ((size_t)256 * sortBlocks + N) / SCAN_TILE + 2));
    RB_CUDA(dalloc(&childL, N)); RB_CUDA(dalloc(&childR, N)); RB_CUDA(dalloc(&parentInt, N)); RB_CUDA(dalloc(&parentLeaf, N));
    RB_CUDA(dalloc(&rFirst, N)); RB_CUDA(dalloc(&rLast, N)); RB_CUDA(dalloc(&flags, N));
    RB_CUDA(dalloc(&leafLo, N)); RB_CUDA(dalloc(&leafHi, N)); RB_CUDA(dalloc(&nodeLo, N)); RB_CUDA(dalloc(&nodeHi, N));
    RB_CUDA(cudaMemsetAsync(flags, 0, (size_t)N * 4, stream));
    RB_CUDA(cudaMemsetAsync(parentLeaf, 0xFF, (size_t)N * 4, stream));

    const uint32_t B = 256, G = (N + B - 1) / B;
    if (in.premade) k_premade<<<G, B, 0, stream>>>(in.premade, N, unsorted, dBounds);
    else k_flatten<<<G, B, 0, stream>>>(in.vertices, in.indices, in.d_instances, dPrefix, in.numInstances, N, unsorted, dBounds);
    nl++;
    k_morton<<<G, B, 0, stream>>>(unsorted, N, dBounds, keys[0], vals[0]); nl++;
    int cur = 0;
    for (int pass = 0; pass < 8; pass++) {
        k_radix_hist<<<sortBlocks, SORT_BLOCK, 0, stream>>>(keys[cur], N, pass * 8, hist, sortBlocks); nl++;
        exclusive_scan<uint32_t>(hist, histScan, 256u * sortBlocks, nullptr, scanScratch, stream, nl);
        k_radix_scatter<<<sortBlocks, SORT_BLOCK, 0, stream>>>(keys[cur], vals[cur], N, pass * 8, histScan, sortBlocks,
                                                               keys[cur ^ 1], vals[cur ^ 1]); nl++;
        cur ^= 1;
    }
    k_gather_sorted<<<G, B, 0, stream>>>(unsorted, vals[cur], N, sorted, leafLo, leafHi); nl++;
    uint32_t rootRef = LEAF_FLAG | 0u;
    const TriRecord* leafTris = sorted;          // triangles in the order the binary tree's leaf ranges refer to
    const float4 *leafLoFinal = leafLo, *leafHiFinal = leafHi;
    if (N > 1 && in.builder == BUILDER_LBVH) {
        k_karras<<<G, B, 0, stream>>>(keys[cur], (int)N, childL, childR, parentInt, parentLeaf, rFirst, rLast); nl++;
        k_refit<<<G, B, 0, stream>>>(N, childL, childR, parentInt, parentLeaf, leafLo, leafHi, nodeLo, nodeHi, flags); nl++;
        rootRef = 0u;
    } else if (N > 1) {
        uint32_t *cl[2], *nn, *count, *leafPos;
        float4 *cLo[2], *cHi[2], *permLo, *permHi;
        unsigned long long *mflags, *mscan, *mtotal, *mscratch;
        RB_CUDA(dalloc(&cl[0], N)); RB_CUDA(dalloc(&cl[1], N)); RB_CUDA(dalloc(&nn, N)); RB_CUDA(dalloc(&count, N));
        RB_CUDA(dalloc(&leafPos, N));
        RB_CUDA(dalloc(&cLo[0], N)); RB_CUDA(dalloc(&cLo[1], N)); RB_CUDA(dalloc(&cHi[0], N)); RB_CUDA(dalloc(&cHi[1], N));
        RB_CUDA(dalloc(&permLo, N)); RB_CUDA(dalloc(&permHi, N));
        RB_CUDA(dalloc(&mflags, N)); RB_CUDA(dalloc(&mscan, N)); RB_CUDA(dalloc(&mtotal, 1));
        RB_CUDA(dalloc(&mscratch, (size_t)N / SCAN_TILE + 2));
        k_ploc_init<<<G, B, 0, stream>>>(N, cl[0]); nl++;
        RB_CUDA(cudaMemcpyAsync(cLo[0], leafLo, (size_t)N * sizeof(float4), cudaMemcpyDeviceToDevice, stream));
        RB_CUDA(cudaMemcpyAsync(cHi[0], leafHi, (size_t)N * sizeof(float4), cudaMemcpyDeviceToDevice, stream));
        uint32_t c = N, nodeBase = 0, rounds = 0, forced = 0;
        bool forcePairing = false;
        int pc = 0;
        while (c > 1) {
            const uint32_t g = (c + PLOC_BLOCK - 1) / PLOC_BLOCK;
            // a round that merged fewer than 1/8 of the clusters is followed by one that pairs neighbours by position
            // (the LBVH choice), which halves the array: the number of rounds stays O(log N) on any input. Only while
            // more than 1024 clusters are left: the last rounds build the top of the tree, where a poor pairing is
            // expensive (nodes per ray 9.6 -> 11.3 on the headline scene when the rule also fired there) and a slow
            // round is cheap
            if (forcePairing) k_ploc_pair_adjacent<<<g, PLOC_BLOCK, 0, stream>>>(c, nn);
            else k_ploc_nearest<<<g, PLOC_BLOCK, 0, stream>>>(c, cLo[pc], cHi[pc], nn);
            nl++;
            k_ploc_flags<<<g, PLOC_BLOCK, 0, stream>>>(c, nn, mflags); nl++;
            exclusive_scan<unsigned long long>(mflags, mscan, c, mtotal, mscratch, stream, nl);
            k_ploc_apply<<<g, PLOC_BLOCK, 0, stream>>>(c, nn, mscan, cl[pc], cLo[pc], cHi[pc], cl[pc ^ 1], cLo[pc ^ 1], cHi[pc ^ 1],
                                                       nodeBase, childL, childR, nodeLo, nodeHi, count, parentInt, parentLeaf); nl++;
            unsigned long long tot = 0;
            RB_CUDA(cudaMemcpyAsync(&tot, mtotal, sizeof(tot), cudaMemcpyDeviceToHost, stream));
            RB_CUDA(cudaStreamSynchronize(stream));
            const uint32_t merges = (uint32_t)(tot & 0xFFFFFFFFull), survivors = (uint32_t)(tot >> 32);
            if (merges == 0 || survivors != c - merges) { set_error("internal: PLOC round made no progress"); return RB200_ERR_CUDA; }
            forcePairing = !forcePairing && c > 1024u && (unsigned long long)merges * 8ull < c;
            forced += forcePairing ? 1u : 0u;
            nodeBase += merges; c = survivors; pc ^= 1; rounds++;
        }
        if (getenv("RB200_DEBUG_BUILD")) fprintf(stderr, "[rb200] PLOC: %u leaves, %u rounds (%u paired by position)\n", N, rounds, forced);
        if (nodeBase != N - 1) { set_error("internal: PLOC built %u of %u nodes", nodeBase, N - 1); return RB200_ERR_CUDA; }
        rootRef = N - 2;       // the last node created
        k_ploc_ranges<<<(2 * N - 1 + B - 1) / B, B, 0, stream>>>(N, childL, parentInt, parentLeaf, count, rFirst, rLast, leafPos); nl++;
        k_ploc_permute<<<G, B, 0, stream>>>(N, leafPos, sorted, leafLo, leafHi, unsorted, permLo, permHi, childL, childR); nl++;
        leafTris = unsorted; leafLoFinal = permLo; leafHiFinal = permHi;
    }
    RB_CUDA(cudaGetLastError());

    // collapse, level by level
    WideNode* nodesTmp; uint32_t *nodeInternalCount, *nodeTriCount, *nodeChildRefs, *slotTriFirst, *work[2], *levelPrefix, *dTotal;
    const size_t maxNodes = N;   // every wide node has >= 1 leaf slot or >= 2 children: never more nodes than triangles
    RB_CUDA(dalloc(&nodesTmp, maxNodes)); RB_CUDA(dalloc(&nodeInternalCount, maxNodes)); RB_CUDA(dalloc(&nodeTriCount, maxNodes));
    RB_CUDA(dalloc(&nodeChildRefs, maxNodes * 8)); RB_CUDA(dalloc(&slotTriFirst, maxNodes * 8));
    RB_CUDA(dalloc(&work[0], maxNodes)); RB_CUDA(dalloc(&work[1], maxNodes)); RB_CUDA(dalloc(&levelPrefix, maxNodes));
    RB_CUDA(dalloc(&dTotal, 1));
    CollapseArrays A{childL, childR, rFirst, rLast, leafLoFinal, leafHiFinal, nodeLo, nodeHi, nodesTmp, nodeInternalCount, nodeTriCount,
                     nodeChildRefs, slotTriFirst};
    RB_CUDA(cudaMemcpyAsync(work[0], &rootRef, 4, cudaMemcpyHostToDevice, stream));
    float* dpCost = nullptr; uint2* dpDec = nullptr;
#if RB_COLLAPSE_DP
    if (N > 1) {
        RB_CUDA(dalloc(&dpCost, (size_t)N * 8)); RB_CUDA(dalloc(&dpDec, N));
        RB_CUDA(cudaMemsetAsync(flags, 0, (size_t)N * 4, stream));
        k_collapse_cost<<<G, B, 0, stream>>>(N, A, parentInt, parentLeaf, flags, dpCost, dpDec); nl++;
    }
#endif
    uint32_t levelBase = 0, levelCount = 1, depth = 0;
    int wcur = 0;
    while (levelCount > 0) {
        if ((size_t)levelBase + levelCount > maxNodes) { set_error("internal: wide node overflow"); return RB200_ERR_CUDA; }
        uint32_t g = (levelCount + 127) / 128;
        k_collapse_level<<<g, 128, 0, stream>>>(A, dpDec, work[wcur], levelCount, levelBase); nl++;
        exclusive_scan<uint32_t>(nodeInternalCount + levelBase, levelPrefix, levelCount, dTotal, scanScratch, stream, nl);
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
    exclusive_scan<uint32_t>(nodeTriCount, triPrefix, numNodes, dTotal, scanScratch, stream, nl);
    // nodes and triangles live in ONE allocation so that a single L2 access-policy window can keep the whole
    // hierarchy resident (see rb200_scene_create)
    const size_t nodeBytes = ((size_t)numNodes * sizeof(WideNode) + 255) & ~(size_t)255;
    out->blobBytes = nodeBytes + (size_t)N * sizeof(TriRecord);
    RB_CUDA(cudaMalloc(&out->blob, out->blobBytes));
    out->nodes = reinterpret_cast<WideNode*>(out->blob);
    out->tris = reinterpret_cast<TriRecord*>(reinterpret_cast<char*>(out->blob) + nodeBytes);
    k_place_triangles<<<(numNodes + 127) / 128, 128, 0, stream>>>(nodesTmp, numNodes, triPrefix, slotTriFirst, leafTris, out->tris); nl++;
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

    g_arena = nullptr;      // the arena itself is released when `arena` goes out of scope
    if (launches) *launches += nl;
    return RB200_OK;
}

// ---------------------------------------------------------------------------------------------------
// Two-level mode (RB200_FLAG_TWO_LEVEL; src/scene/Scene.cpp:93-111: a BLAS per object, a TLAS entry per instance)
// ---------------------------------------------------------------------------------------------------
__global__ void k_merge_nodes(const WideNode* __restrict__ in, uint32_t n, uint32_t nodeOff, uint32_t triOff, WideNode* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    WideNode w = in[i];
    w.childBase += nodeOff; w.triBase += triOff;
    out[i] = w;
}
__global__ void k_merge_tris(const TriRecord* __restrict__ in, uint32_t n, uint32_t representative, TriRecord* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    TriRecord t = in[i];
    t.v1.w = __uint_as_float(representative);      // an instance of this model: the shading records are built through it
    out[i] = t;
}

// One hierarchy per distinct model range in OBJECT space (the existing builder on a single identity-transform instance, so
// the vertex values are exactly the flattening path's rb_m4_point(identity, v)), merged into one node and one triangle
// array with absolute indices; a hierarchy over the instances' padded world-space boxes; the instances' entry records.
int build_two_level(const BuildInput& in, cudaStream_t stream, Bvh* blas, Bvh* tlas, std::vector<TwoLevelInstance>* entries,
                    uint32_t* numModels, uint64_t* launches) {
    const std::vector<RB200Instance>& hi = *in.h_instances;
    struct Model { uint32_t indexOffset, triangleCount, representative; Bvh bvh; uint32_t nodeOff, triOff; };
    std::vector<Model> models;
    std::vector<uint32_t> modelOf(hi.size());
    for (size_t i = 0; i < hi.size(); i++) {
        size_t m = 0;
        while (m < models.size() && !(models[m].indexOffset == hi[i].indexOffset && models[m].triangleCount == hi[i].triangleCount)) m++;
        if (m == models.size()) { Model mm{}; mm.indexOffset = hi[i].indexOffset; mm.triangleCount = hi[i].triangleCount; mm.representative = (uint32_t)i; models.push_back(mm); }
        modelOf[i] = (uint32_t)m;
    }
    auto release = [&]() { for (Model& m : models) free_bvh(&m.bvh); };
    RB200Instance* dOne = nullptr;
    if (cudaMalloc(&dOne, sizeof(RB200Instance)) != cudaSuccess) { cudaGetLastError(); set_error("out of device memory"); return RB200_ERR_OUT_OF_MEMORY; }
    uint32_t totalNodes = 0, totalTris = 0, maxDepth = 0;
    float buildMs = 0.f;
    for (Model& m : models) {
        RB200Instance one = hi[m.representative];
        static const float identity[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        memcpy(one.transform, identity, sizeof identity);
        std::vector<RB200Instance> hone(1, one);
        cudaError_t e = cudaMemcpyAsync(dOne, &one, sizeof one, cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) { set_error("two-level build: %s", cudaGetErrorString(e)); cudaFree(dOne); release(); return RB200_ERR_CUDA; }
        BuildInput bi{in.vertices, in.indices, dOne, &hone, 1u, in.builder};
        const int rc = build_bvh(bi, stream, &m.bvh, launches);
        if (rc != RB200_OK) { cudaFree(dOne); release(); return rc; }
        m.nodeOff = totalNodes; m.triOff = totalTris;
        totalNodes += m.bvh.numNodes; totalTris += m.bvh.numTris;
        maxDepth = std::max(maxDepth, m.bvh.maxDepth);
        buildMs += m.bvh.buildMs;
    }
    cudaFree(dOne);
    if (totalTris >= 0x08000000u) { set_error("too many triangles (limit 2^27)"); release(); return RB200_ERR_INVALID_ARGUMENT; }

    const size_t nodeBytes = ((size_t)totalNodes * sizeof(WideNode) + 255) & ~(size_t)255;
    blas->blobBytes = nodeBytes + (size_t)totalTris * sizeof(TriRecord);
    if (cudaMalloc(&blas->blob, blas->blobBytes) != cudaSuccess) { cudaGetLastError(); blas->blob = nullptr; set_error("out of device memory (two-level hierarchy)"); release(); return RB200_ERR_OUT_OF_MEMORY; }
    blas->nodes = reinterpret_cast<WideNode*>(blas->blob);
    blas->tris = reinterpret_cast<TriRecord*>(reinterpret_cast<char*>(blas->blob) + nodeBytes);
    for (Model& m : models) {
        k_merge_nodes<<<(m.bvh.numNodes + 127) / 128, 128, 0, stream>>>(m.bvh.nodes, m.bvh.numNodes, m.nodeOff, m.triOff, blas->nodes + m.nodeOff);
        k_merge_tris<<<(m.bvh.numTris + 127) / 128, 128, 0, stream>>>(m.bvh.tris, m.bvh.numTris, m.representative, blas->tris + m.triOff);
        if (launches) *launches += 2;
    }
    {
        cudaError_t e = cudaStreamSynchronize(stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) { set_error("two-level build: %s", cudaGetErrorString(e)); release(); free_bvh(blas); return RB200_ERR_CUDA; }
    }
    blas->numNodes = totalNodes; blas->numTris = totalTris; blas->maxDepth = maxDepth; blas->buildMs = buildMs;

    // entry records and world-space boxes of the instances
    entries->assign(hi.size(), TwoLevelInstance{});
    std::vector<TriRecord> boxes(hi.size());
    uint32_t gid = 0;
    for (size_t i = 0; i < hi.size(); i++) {
        const Model& m = models[modelOf[i]];
        TwoLevelInstance& e = (*entries)[i];
        if (!rb_affine_inverse(hi[i].transform, e.inv)) {
            set_error("instance %zu: the transform is singular (or not finite); two-level mode needs its inverse", i);
            release(); free_bvh(blas); return RB200_ERR_INVALID_ARGUMENT;
        }
        e.rootNode = m.nodeOff; e.gidBase = gid; e.material = hi[i].materialIdx > 3u ? 3u : hi[i].materialIdx; e.pad = 0u;
        gid += hi[i].triangleCount;
        // world box = box of the 8 transformed corners of the model's bounds (fp64), padded: the ray is intersected with the
        // triangles in object space after an fp32 inverse transform, so "the world ray meets the world box" and "the object ray
        // meets a triangle" can disagree by a few 1e-7 of the coordinates involved; 2^-15 of them (and of the extent) is ample
        const float* M = hi[i].transform;
        double lo[3] = {1e300, 1e300, 1e300}, hb[3] = {-1e300, -1e300, -1e300}, mag = 0.0;
        for (int c = 0; c < 8; c++) {
            const double x = (c & 1) ? m.bvh.sceneMax[0] : m.bvh.sceneMin[0], y = (c & 2) ? m.bvh.sceneMax[1] : m.bvh.sceneMin[1],
                         z = (c & 4) ? m.bvh.sceneMax[2] : m.bvh.sceneMin[2];
            for (int a = 0; a < 3; a++) {
                const double w = (double)M[a] * x + (double)M[4 + a] * y + (double)M[8 + a] * z + (double)M[12 + a];
                lo[a] = std::min(lo[a], w); hb[a] = std::max(hb[a], w); mag = std::max(mag, std::fabs(w));
            }
        }
        const double ext = std::max(std::max(hb[0] - lo[0], hb[1] - lo[1]), hb[2] - lo[2]);
        // ... times the condition of the transform (1 for a rotation with uniform scale): the rounding of the inverse
        // transform, mapped back to world space, grows with it
        double fm = 0.0, fi = 0.0;
        for (int c = 0; c < 3; c++)
            for (int a = 0; a < 3; a++) { fm += (double)M[4 * c + a] * M[4 * c + a]; fi += (double)e.inv[4 * a + c] * e.inv[4 * a + c]; }
        const double cond = std::max(1.0, std::sqrt(fm * fi) / 3.0);
        const double pad = 3.0517578125e-05 * cond * (mag + ext) + 1e-30;
        TriRecord& b = boxes[i];
        memset(&b, 0, sizeof b);
        float idAsFloat;
        const uint32_t id = (uint32_t)i;
        memcpy(&idAsFloat, &id, 4);
        b.v0 = make_float4((float)(lo[0] - pad), (float)(lo[1] - pad), (float)(lo[2] - pad), idAsFloat);
        b.v1 = make_float4((float)(hb[0] + pad), (float)(hb[1] + pad), (float)(hb[2] + pad), 0.f);
        b.v2 = b.v0;
    }
    release();
    TriRecord* dBoxes = nullptr;
    if (cudaMalloc(&dBoxes, boxes.size() * sizeof(TriRecord)) != cudaSuccess) { cudaGetLastError(); set_error("out of device memory"); free_bvh(blas); return RB200_ERR_OUT_OF_MEMORY; }
    cudaError_t e = cudaMemcpyAsync(dBoxes, boxes.data(), boxes.size() * sizeof(TriRecord), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) { set_error("two-level build: %s", cudaGetErrorString(e)); cudaFree(dBoxes); free_bvh(blas); return RB200_ERR_CUDA; }
    BuildInput bt = in;
    bt.premade = dBoxes; bt.numPremade = (uint32_t)boxes.size();
    const int rc = build_bvh(bt, stream, tlas, launches);
    cudaFree(dBoxes);
    if (rc != RB200_OK) { free_bvh(blas); return rc; }
    *numModels = (uint32_t)models.size();
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
