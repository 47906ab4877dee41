// ORACLE (test infrastructure) — bloom (blurX with threshold, blurY, combine) and ACES-fitted tonemap,
// restating shaders/postprocessing/bloom/blurCommon.h.glsl:15-56, blurX/blurY.comp.glsl:7-9,
// combine.comp.glsl:15-27 and tonemap/tonemapping.comp.glsl:17-84. Pass chaining as in
// Reina::applyBloom / applyTonemapping (src/Reina.cpp:472-577): rt -> ping (X), ping -> pong (Y),
// rt + pong*intensity -> ping, ping -> RGBA8.
#include "oracle_common.h"
#include <thread>
#include <vector>
#include <cmath>

namespace oracle {

// blurCommon.h.glsl:10-13. sigma is pushConstants.radius, the PERCENT value (quirk F9 of SURVEY.md).
static inline float gauss(float x, float sigma) { return rb_exp(-x * x / (2.0f * sigma * sigma)); }

// blurCommon.h.glsl:15-56 for the rows [y0, y1)
static void blur_axis_rows(const float* in, float* out, int W, int H, int ax, int ay, bool applyThreshold,
                           const RB200BloomPushConsts& pc, int y0, int y1) {
    const float radiusPx = (float)(ax == 1 ? W : H) * pc.radius / 100.0f;
    const int k = (int)(radiusPx * 3.0f + 0.5f);
    std::vector<float> w((size_t)(2 * k + 1));
    for (int i = -k; i <= k; i++) w[(size_t)(i + k)] = gauss((float)i, pc.radius);
    for (int y = y0; y < y1; y++) {
        for (int x = 0; x < W; x++) {
            float cr = 0.0f, cg = 0.0f, cb = 0.0f, weightSum = 0.0f;
            for (int i = -k; i <= k; i++) {
                const float weight = w[(size_t)(i + k)];
                weightSum += weight;
                const int cx = x + i * ax, cy = y + i * ay;
                if (cx < 0 || cx >= W || cy < 0 || cy >= H) continue;
                const float* p = &in[4 * ((size_t)cy * W + cx)];
                const float lum = p[0] * 0.299f + p[1] * 0.587f + p[2] * 0.114f;
                if (applyThreshold && lum < pc.threshold) continue;
                cr += p[0] * weight; cg += p[1] * weight; cb += p[2] * weight;
            }
            if (weightSum < 0.0001f) { cr = cg = cb = 0.0f; }
            else { cr /= weightSum; cg /= weightSum; cb /= weightSum; }
            float* o = &out[4 * ((size_t)y * W + x)];
            o[0] = cr; o[1] = cg; o[2] = cb; o[3] = 1.0f;
        }
    }
}

// tonemapping.comp.glsl:34-39
static inline float rrt_odt_fit(float v) {
    float a = v * (v + 0.0245786f) - 0.000090537f;
    float b = v * (0.983729f * v + 0.4329510f) + 0.238081f;
    return a / b;
}

// tonemapping.comp.glsl:62-84 for one pixel. `color * M` is row-vector times matrix: component j = dot(color, column j).
void tonemap_pixel(const float rgb[3], float exposure, uint8_t out[4]) {
    const float e = rb_exp2(exposure);
    vec3 c = rb_mk3(rgb[0] * e, rgb[1] * e, rgb[2] * e);
    vec3 a = rb_mk3(rb_dot(c, rb_mk3(0.59719f, 0.35458f, 0.04823f)),
                    rb_dot(c, rb_mk3(0.07600f, 0.90834f, 0.01566f)),
                    rb_dot(c, rb_mk3(0.02840f, 0.13383f, 0.83777f)));
    a = rb_mk3(rrt_odt_fit(a.x), rrt_odt_fit(a.y), rrt_odt_fit(a.z));
    vec3 o = rb_mk3(rb_dot(a, rb_mk3(1.60475f, -0.53108f, -0.07367f)),
                    rb_dot(a, rb_mk3(-0.10208f, 1.10813f, -0.00605f)),
                    rb_dot(a, rb_mk3(-0.00327f, -0.07276f, 1.07602f)));
    o = rb_clamp3(o, 0.0f, 1.0f);
    // rgba8 image store: UNORM conversion, round to nearest
    out[0] = (uint8_t)(int)rintf(o.x * 255.0f);
    out[1] = (uint8_t)(int)rintf(o.y * 255.0f);
    out[2] = (uint8_t)(int)rintf(o.z * 255.0f);
    out[3] = 255;
}

template <class F> static void parallel_rows(int H, int threads, F f) {
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    int chunk = (H + threads - 1) / threads;
    for (int t = 0; t < threads; t++) {
        int y0 = t * chunk, y1 = std::min(H, y0 + chunk);
        if (y0 >= y1) break;
        pool.emplace_back([=]() { f(y0, y1); });
    }
    for (auto& th : pool) th.join();
}

void postprocess(int W, int H, const float* hdr, const RB200BloomPushConsts& bloom, const RB200TonemappingPushConsts& tm,
                 uint8_t* ldr, float* combined_out, int threads) {
    std::vector<float> ping((size_t)W * H * 4), pong((size_t)W * H * 4);
    parallel_rows(H, threads, [&](int y0, int y1) { blur_axis_rows(hdr, ping.data(), W, H, 1, 0, true, bloom, y0, y1); });
    parallel_rows(H, threads, [&](int y0, int y1) { blur_axis_rows(ping.data(), pong.data(), W, H, 0, 1, false, bloom, y0, y1); });
    // combine.comp.glsl:15-27: rt + bloom * intensity -> ping
    parallel_rows(H, threads, [&](int y0, int y1) {
        for (size_t i = (size_t)y0 * W; i < (size_t)y1 * W; i++) {
            ping[4 * i + 0] = hdr[4 * i + 0] + pong[4 * i + 0] * bloom.intensity;
            ping[4 * i + 1] = hdr[4 * i + 1] + pong[4 * i + 1] * bloom.intensity;
            ping[4 * i + 2] = hdr[4 * i + 2] + pong[4 * i + 2] * bloom.intensity;
            ping[4 * i + 3] = 1.0f;
        }
    });
    parallel_rows(H, threads, [&](int y0, int y1) {
        for (size_t i = (size_t)y0 * W; i < (size_t)y1 * W; i++) tonemap_pixel(&ping[4 * i], tm.exposure, &ldr[4 * i]);
    });
    if (combined_out) memcpy(combined_out, ping.data(), ping.size() * sizeof(float));
}

} // namespace oracle
