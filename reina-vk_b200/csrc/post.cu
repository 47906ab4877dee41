// post.cu — bloom (thresholded blur X, blur Y) and fused combine + ACES-fitted tonemap.
//
// Replaces Reina::applyBloom / applyTonemapping (src/Reina.cpp:472-577), i.e. the compute shaders
//   shaders/postprocessing/bloom/blurCommon.h.glsl:15-56 (blurX.comp.glsl:7-9 threshold on, blurY.comp.glsl:7-9 off)
//   shaders/postprocessing/bloom/combine.comp.glsl:15-27
//   shaders/postprocessing/tonemap/tonemapping.comp.glsl:17-84
// Reference quirk kept: the Gaussian sigma is the radius *percent* value while the loop spans
// +-3 * radius% * dimension taps and weightSum accumulates every tap, in or out of the image. Weights underflow
// to exactly 0 beyond |i| ~ 13.2 sigma (rb_exp flushes below 2^-126), and adding 0-weighted taps changes neither
// the colour sum nor weightSum, so only the +-R taps with non-zero weight are visited — in the same ascending
// order as the shader's loop, which keeps the fp32 sums bit-identical to the oracle.
// Tiles are staged in shared memory with 128-bit loads; thresholded / out-of-image taps are staged as zeros.
// Every thread visits only the span of NON-ZERO pixels of its tap window (an occupancy bit mask of the staged tile is built
// while staging): most of a frame is zero after the threshold, a zero pixel contributes +-0 to sums that start at +0, so
// the result is what the full loops would produce, bit for bit (NaN and infinities count as non-zero). With it the two
// blur passes read and write each image once and approach the HBM bound like combine + tonemap.
#include "context.cuh"

namespace rb200 {

#ifndef RB_BLUR_OUT
#define RB_BLUR_OUT 4
#endif
static constexpr int BLUR_OUT = RB_BLUR_OUT;                   // outputs per thread along the blur axis (register blocking)
static constexpr int BX_THREADS = 512 / BLUR_OUT;
static constexpr int BX_TILE = BX_THREADS * BLUR_OUT; // blur X: pixels of one row per block
static constexpr int BY_W = 16;                       // blur Y: tile width
static constexpr int BY_H = 128;                      // blur Y: tile height (rows per block)
static constexpr int BY_THREADS = BY_W * (BY_H / BLUR_OUT);

__device__ __forceinline__ float gauss(float x, float sigma) { return rb_exp(-x * x / (2.0f * sigma * sigma)); }

// Each thread produces BLUR_OUT adjacent outputs along the blur axis: a staged pixel is loaded from shared memory once
// and used by up to BLUR_OUT outputs (tap i of output o is pixel t = i + o), with the weights sliding through
// registers. Every output still adds its taps in ascending order, one product and one add per channel, so the sums
// are the ones the shader's loop produces. Only the pixels t0 .. t1 of the thread's window are visited — the span of its
// non-zero pixels: a zero pixel contributes +-0 to sums that start at +0, which never changes them (x + 0 = x, and
// +0 + -0 = +0), so skipping it is exact. weightSum does not depend on the pixel and comes from the host.
// `tile` points at tap 0 of output 0, consecutive pixels along the axis are `stride` entries apart.
__device__ __forceinline__ void blur_outputs(const float* __restrict__ wts, const float4* __restrict__ tile, int stride, int R, int t0, int t1,
                                             float (&cr)[BLUR_OUT], float (&cg)[BLUR_OUT], float (&cb)[BLUR_OUT]) {
    const int n = 2 * R + 1;
    float w[BLUR_OUT];
#pragma unroll
    for (int o = 0; o < BLUR_OUT; o++) { cr[o] = 0.f; cg[o] = 0.f; cb[o] = 0.f; }
    // state of the sliding window as if tap t0 - 1 had just been processed: w[o] = wts[t0 - 1 - o]
#pragma unroll
    for (int o = 0; o < BLUR_OUT; o++) { const int i = t0 - 1 - o; w[o] = (i >= 0 && i < n) ? wts[i] : 0.f; }
    for (int t = t0; t <= t1; t++) {
#pragma unroll
        for (int o = BLUR_OUT - 1; o > 0; o--) w[o] = w[o - 1];
        w[0] = t < n ? wts[t] : 0.f;
        const float4 p = tile[t * stride];
        if (t >= BLUR_OUT - 1 && t < n) {          // the tap is valid for every output of this thread
#pragma unroll
            for (int o = 0; o < BLUR_OUT; o++) { cr[o] += p.x * w[o]; cg[o] += p.y * w[o]; cb[o] += p.z * w[o]; }
        } else {
#pragma unroll
            for (int o = 0; o < BLUR_OUT; o++) {
                const int i = t - o;
                if (i >= 0 && i < n) { cr[o] += p.x * w[o]; cg[o] += p.y * w[o]; cb[o] += p.z * w[o]; }
            }
        }
    }
}

// first and last set bit of mask[] within [lo, hi) (bit i of word i >> 5); first > last when none is set
__device__ __forceinline__ void nonzero_span(const uint32_t* __restrict__ mask, int lo, int hi, int& first, int& last) {
    first = hi; last = lo - 1;
    for (int w = lo >> 5; w <= (hi - 1) >> 5; w++) {
        uint32_t m = mask[w];
        if (w == (lo >> 5)) m &= 0xFFFFFFFFu << (lo & 31);
        if (w == ((hi - 1) >> 5) && (hi & 31)) m &= 0xFFFFFFFFu >> (32 - (hi & 31));
        if (m) {
            first = min(first, w * 32 + (__ffs(m) - 1));
            last = max(last, w * 32 + 31 - __clz(m));
        }
    }
}

__device__ __forceinline__ float4 blur_resolve(float cr, float cg, float cb, float ws) {
    if (ws < 0.0001f) { cr = cg = cb = 0.f; } else { cr /= ws; cg /= ws; cb /= ws; }
    return make_float4(cr, cg, cb, 1.f);
}

// Blur along x with the threshold. Shared memory: weights | occupancy mask of the staged row segment | the segment.
__global__ void __launch_bounds__(BX_THREADS) k_blur_x(const float4* __restrict__ in, float4* __restrict__ out, int W, int H,
                                                       int R, float sigma, float threshold, float ws) {
    extern __shared__ float4 smem[];
    const int TW = BX_TILE + 2 * R, MW = (TW + 31) / 32;
    float* wts = reinterpret_cast<float*>(smem);              // 2R+1 weights (padded to a multiple of 4 floats)
    uint32_t* mask = reinterpret_cast<uint32_t*>(smem + ((2 * R + 1 + 3) / 4));     // bit i: staged pixel i is non-zero
    float4* tile = smem + ((2 * R + 1 + 3) / 4) + ((MW + 3) / 4);                   // BX_TILE + 2R pixels
    const int y = blockIdx.y;
    const int x0 = blockIdx.x * BX_TILE;
    for (int i = threadIdx.x; i <= 2 * R; i += BX_THREADS) wts[i] = gauss((float)(i - R), sigma);
    const int lane = threadIdx.x & 31;
    for (int base = (int)(threadIdx.x & ~31u); base < TW; base += BX_THREADS) {      // a warp stages 32 consecutive pixels
        const int i = base + lane, x = x0 - R + i;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < TW && x >= 0 && x < W) {
            p = in[(size_t)y * W + x];
            const float lum = p.x * 0.299f + p.y * 0.587f + p.z * 0.114f;
            if (lum < threshold) p = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (i < TW) tile[i] = p;
        const uint32_t b = __ballot_sync(0xffffffffu, (p.x != 0.f) | (p.y != 0.f) | (p.z != 0.f));
        if (lane == 0) mask[base >> 5] = b;
    }
    __syncthreads();
    const int xo = x0 + (int)threadIdx.x * BLUR_OUT;
    if (xo >= W) return;
    const int off = (int)threadIdx.x * BLUR_OUT, n = 2 * R + 1;
    int first, last;
    nonzero_span(mask, off, off + n + BLUR_OUT - 1, first, last);
    float cr[BLUR_OUT], cg[BLUR_OUT], cb[BLUR_OUT];
    if (first > last) {
#pragma unroll
        for (int o = 0; o < BLUR_OUT; o++) { cr[o] = 0.f; cg[o] = 0.f; cb[o] = 0.f; }
    } else blur_outputs(wts, tile + off, 1, R, first - off, last - off, cr, cg, cb);
#pragma unroll
    for (int o = 0; o < BLUR_OUT; o++)
        if (xo + o < W) out[(size_t)y * W + xo + o] = blur_resolve(cr[o], cg[o], cb[o], ws);
}

// Blur along y. Shared memory: weights | per-column occupancy masks of the staged rows | (BY_H + 2R) rows x BY_W.
__global__ void __launch_bounds__(BY_THREADS) k_blur_y(const float4* __restrict__ in, float4* __restrict__ out, int W, int H,
                                                       int R, float sigma, float ws) {
    extern __shared__ float4 smem[];
    const int rows = BY_H + 2 * R, MW = (rows + 31) / 32;
    float* wts = reinterpret_cast<float*>(smem);
    uint32_t* mask = reinterpret_cast<uint32_t*>(smem + ((2 * R + 1 + 3) / 4));     // [BY_W][MW]: bit r of column c
    float4* tile = smem + ((2 * R + 1 + 3) / 4) + ((BY_W * MW + 3) / 4);
    const int x0 = blockIdx.x * BY_W, y0 = blockIdx.y * BY_H;
    const int tx = threadIdx.x % BY_W, tg = threadIdx.x / BY_W;     // column, group of BLUR_OUT rows
    for (int i = threadIdx.x; i <= 2 * R; i += BY_THREADS) wts[i] = gauss((float)(i - R), sigma);
    for (int i = threadIdx.x; i < BY_W * MW; i += BY_THREADS) mask[i] = 0u;
    __syncthreads();
    for (int r = tg; r < rows; r += BY_THREADS / BY_W) {
        const int y = y0 - R + r, x = x0 + tx;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (y >= 0 && y < H && x < W) p = in[(size_t)y * W + x];
        if ((p.x != 0.f) | (p.y != 0.f) | (p.z != 0.f)) atomicOr(&mask[tx * MW + (r >> 5)], 1u << (r & 31));
        tile[r * BY_W + tx] = p;
    }
    __syncthreads();
    const int x = x0 + tx, yo = y0 + tg * BLUR_OUT;
    if (x >= W || yo >= H) return;
    const int off = tg * BLUR_OUT, n = 2 * R + 1;
    int first, last;
    nonzero_span(mask + tx * MW, off, off + n + BLUR_OUT - 1, first, last);
    float cr[BLUR_OUT], cg[BLUR_OUT], cb[BLUR_OUT];
    if (first > last) {
#pragma unroll
        for (int o = 0; o < BLUR_OUT; o++) { cr[o] = 0.f; cg[o] = 0.f; cb[o] = 0.f; }
    } else blur_outputs(wts, tile + off * BY_W + tx, BY_W, R, first - off, last - off, cr, cg, cb);
#pragma unroll
    for (int o = 0; o < BLUR_OUT; o++)
        if (yo + o < H) out[(size_t)(yo + o) * W + x] = blur_resolve(cr[o], cg[o], cb[o], ws);
}

// blurCommon.h.glsl:15-56 as written: one thread per pixel, every one of the 2k+1 taps, straight from global memory. Only
// used when the image may hold NaN or infinite values (RB200Context::hdrMayBeNonFinite): a tap whose weight underflowed
// to 0 still turns a NaN or an infinity into NaN in the shader's loop, so the window of non-zero weights the tiled kernels
// visit is not enough there. ax = 1: blur along x with the threshold; ax = 0: along y without.
__global__ void __launch_bounds__(256) k_blur_exact(const float4* __restrict__ in, float4* __restrict__ out, int W, int H, int k, float sigma,
                                                    float threshold, int ax) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    float cr = 0.f, cg = 0.f, cb = 0.f, ws = 0.f;
    for (int i = -k; i <= k; i++) {
        const float w = gauss((float)i, sigma);
        ws += w;
        const int cx = ax ? x + i : x, cy = ax ? y : y + i;
        if (cx < 0 || cx >= W || cy < 0 || cy >= H) continue;
        const float4 p = in[(size_t)cy * W + cx];
        if (ax) {
            const float lum = p.x * 0.299f + p.y * 0.587f + p.z * 0.114f;
            if (lum < threshold) continue;
        }
        cr += p.x * w; cg += p.y * w; cb += p.z * w;
    }
    out[(size_t)y * W + x] = blur_resolve(cr, cg, cb, ws);
}

// tonemapping.comp.glsl:34-39
__device__ __forceinline__ float rrt_odt_fit(float v) {
    const float a = v * (v + 0.0245786f) - 0.000090537f;
    const float b = v * (0.983729f * v + 0.4329510f) + 0.238081f;
    return a / b;
}

// combine (rt + bloom * intensity) followed by exposure, ACES input matrix, RRT/ODT fit, ACES output matrix, clamp,
// UNORM8 store. `c * M` in the shader is row-vector times matrix: component j = dot(c, column j).
__device__ __forceinline__ uint32_t combine_tonemap(const float4 rt, const float4 bl, float intensity, float e) {
    const rb_v3 hdr = rb_mk3(rt.x + bl.x * intensity, rt.y + bl.y * intensity, rt.z + bl.z * intensity);
    const rb_v3 c = rb_mk3(hdr.x * e, hdr.y * e, hdr.z * e);
    rb_v3 a = rb_mk3(rb_dot(c, rb_mk3(0.59719f, 0.35458f, 0.04823f)), rb_dot(c, rb_mk3(0.07600f, 0.90834f, 0.01566f)),
                     rb_dot(c, rb_mk3(0.02840f, 0.13383f, 0.83777f)));
    a = rb_mk3(rrt_odt_fit(a.x), rrt_odt_fit(a.y), rrt_odt_fit(a.z));
    rb_v3 o = rb_mk3(rb_dot(a, rb_mk3(1.60475f, -0.53108f, -0.07367f)), rb_dot(a, rb_mk3(-0.10208f, 1.10813f, -0.00605f)),
                     rb_dot(a, rb_mk3(-0.00327f, -0.07276f, 1.07602f)));
    o = rb_clamp3(o, 0.0f, 1.0f);
    const uint32_t r = (uint32_t)(int)rintf(o.x * 255.0f), g = (uint32_t)(int)rintf(o.y * 255.0f), b = (uint32_t)(int)rintf(o.z * 255.0f);
    return r | (g << 8) | (b << 16) | (255u << 24);
}

__global__ void __launch_bounds__(256) k_combine_tonemap(const float4* __restrict__ rt, const float4* __restrict__ bloom,
                                                         uint32_t* __restrict__ ldr, uint32_t n, float intensity, float exposure) {
    const float e = rb_exp2(exposure);
    const uint32_t i4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
    if (i4 >= n) return;
    if (i4 + 4u <= n) {
        uint4 o;
        o.x = combine_tonemap(rt[i4 + 0], bloom[i4 + 0], intensity, e);
        o.y = combine_tonemap(rt[i4 + 1], bloom[i4 + 1], intensity, e);
        o.z = combine_tonemap(rt[i4 + 2], bloom[i4 + 2], intensity, e);
        o.w = combine_tonemap(rt[i4 + 3], bloom[i4 + 3], intensity, e);
        *reinterpret_cast<uint4*>(ldr + i4) = o;
    } else {
        for (uint32_t i = i4; i < n; i++) ldr[i] = combine_tonemap(rt[i], bloom[i], intensity, e);
    }
}

// number of taps on each side with a non-zero weight: the same fp32 expression rb_exp sees, evaluated on the host
static int effective_radius(int k, float sigma) {
    int R = 0;
    for (int i = 0; i <= k; i++) {
        const float x = (float)i;
        const float arg = -x * x / (2.0f * sigma * sigma);
        if (arg < -87.3f) break;
        R = i;
    }
    return R;
}

// the mean image of a SUM image: v * (1 / numBatches), the arithmetic of k_resolve_sum (wavefront.cu), into `dst`
__global__ void k_resolve_into(const float4* __restrict__ src, float4* __restrict__ dst, uint32_t n, float inv) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = src[i];
    dst[i] = make_float4(v.x * inv, v.y * inv, v.z * inv, 1.f);
}

// CUDA loads a kernel's code on its first launch, and that load waits for the device to go idle: with several batches
// in flight the first presented frame would stall the host for a whole pipeline depth. Load everything up front.
void preload_post_kernels() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, k_blur_x);
    cudaFuncGetAttributes(&a, k_blur_y);
    cudaFuncGetAttributes(&a, k_combine_tonemap);
    cudaFuncGetAttributes(&a, k_resolve_into);
    cudaGetLastError();
}

int present_sum(RB200Context* ctx, const float4* deviceSum, uint32_t numBatches, const RB200BloomPushConsts* bloom,
                const RB200TonemappingPushConsts* tm) {
    const uint32_t n = ctx->width * ctx->height;
    if (!ctx->resolved) {
        cudaError_t e = cudaMalloc(&ctx->resolved, (size_t)n * sizeof(float4));
        if (e != cudaSuccess) { cudaGetLastError(); set_error("out of device memory (resolved image)"); return RB200_ERR_OUT_OF_MEMORY; }
        ctx->allocations.push_back(ctx->resolved);     // (contexts created with RB200_FLAG_ACCUM_SUM own it from the start)
    }
    k_resolve_into<<<(n + 255) / 256, 256, 0, ctx->stream>>>(deviceSum ? deviceSum : ctx->wp.image, ctx->resolved, n,
                                                              1.0f / (float)numBatches);
    ctx->launches++;
    RB_CUDA(cudaGetLastError());
    return postprocess(ctx, bloom, tm, ctx->resolved);
}

int postprocess(RB200Context* ctx, const RB200BloomPushConsts* bloom, const RB200TonemappingPushConsts* tm, const float4* source) {
    const int W = (int)ctx->width, H = (int)ctx->height;
    cudaStream_t s = ctx->stream;
    const float4* rt = source ? source : ctx->wp.image;
    // blurCommon.h.glsl:23-24
    const float radiusPxX = (float)W * bloom->radius / 100.0f, radiusPxY = (float)H * bloom->radius / 100.0f;
    const int kX = (int)(radiusPxX * 3.0f + 0.5f), kY = (int)(radiusPxY * 3.0f + 0.5f);
    const int RX = effective_radius(kX, bloom->radius), RY = effective_radius(kY, bloom->radius);

    const size_t smX = (size_t)(((2 * RX + 1 + 3) / 4) + (((BX_TILE + 2 * RX + 31) / 32 + 3) / 4) + BX_TILE + 2 * RX) * sizeof(float4);
    const size_t smY = (size_t)(((2 * RY + 1 + 3) / 4) + ((BY_W * ((BY_H + 2 * RY + 31) / 32) + 3) / 4) + (BY_H + 2 * RY) * BY_W) * sizeof(float4);
    // weightSum of blurCommon.h.glsl:49-53: every tap's weight added in ascending order (the zero-weight taps beyond +-R do
    // not change it); the same fp32 sequence the kernels used to form per thread, from the shared elementary layer
    float wsX = 0.f, wsY = 0.f;
    for (int i = 0; i <= 2 * RX; i++) { const float x = (float)(i - RX); wsX += rb_exp(-x * x / (2.0f * bloom->radius * bloom->radius)); }
    for (int i = 0; i <= 2 * RY; i++) { const float x = (float)(i - RY); wsY += rb_exp(-x * x / (2.0f * bloom->radius * bloom->radius)); }
    if (smX > 200 * 1024 || smY > 200 * 1024) { set_error("bloom radius too large for the shared-memory tiles"); return RB200_ERR_INVALID_ARGUMENT; }
    RB_CUDA(cudaFuncSetAttribute(k_blur_x, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smX));
    RB_CUDA(cudaFuncSetAttribute(k_blur_y, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smY));
    dim3 gx((W + BX_TILE - 1) / BX_TILE, H), gy((W + BY_W - 1) / BY_W, (H + BY_H - 1) / BY_H);
    if (ctx->hdrMayBeNonFinite && !source) {
        // NaN / infinite pixels (only rb200_write_hdr or a non-finite directClamp can bring them in): the shader's loop as written
        dim3 ge((W + 255) / 256, H);
        k_blur_exact<<<ge, 256, 0, s>>>(rt, ctx->ping, W, H, kX, bloom->radius, bloom->threshold, 1);
        k_blur_exact<<<ge, 256, 0, s>>>(ctx->ping, ctx->pong, W, H, kY, bloom->radius, 0.f, 0);
    } else {
        k_blur_x<<<gx, BX_THREADS, smX, s>>>(rt, ctx->ping, W, H, RX, bloom->radius, bloom->threshold, wsX);
        k_blur_y<<<gy, BY_THREADS, smY, s>>>(ctx->ping, ctx->pong, W, H, RY, bloom->radius, wsY);
    }
    const uint32_t n = (uint32_t)W * (uint32_t)H;
    k_combine_tonemap<<<(n / 4 + 256) / 256, 256, 0, s>>>(rt, ctx->pong, reinterpret_cast<uint32_t*>(ctx->ldr), n,
                                                          bloom->intensity, tm->exposure);
    ctx->launches += 3;
    RB_CUDA(cudaGetLastError());
    return RB200_OK;
}

} // namespace rb200
