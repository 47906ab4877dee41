// ORACLE (test infrastructure) — C entry points loaded by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs through ctypes. Mirrors the shape of include/reina_b200.h so that the same
// RB200SceneDesc / RB200RtPushConsts blocks feed both sides.
#include "oracle_common.h"
#include <algorithm>
#include <thread>
#include <cstring>
#include <string>

namespace oracle {
float kat_random(uint32_t* state);
vec3 kat_offset(vec3 p, vec3 n);
vec3 kat_sky(vec3 dir);
float kat_power_heuristic(float a, float b);
void kat_light_sample(const Scene& s, const RB200RtPushConsts& pc, uint32_t* rng, float out[11]);
void kat_direct_light_lambertian(const Scene& s, const RB200RtPushConsts& pc, const float o[3], const float n[3],
                                 const float albedo[3], uint32_t* rng, float out[4]);
void kat_trace_main(const Scene& s, const float o[3], const float d[3], uint32_t* rngState, int insideDielectric,
                    float accumulatedDistance, float out[21], uint32_t* flags);
void kat_bounce(const Scene& s, const RB200RtPushConsts& pc, const float o[3], const float d[3], uint32_t* rngState,
                int insideDielectric, float accumulatedDistance, float out[21], uint32_t* flags, float direct[4],
                uint32_t* rngAfterDirect);
void kat_bump(const uint8_t* rgba, uint32_t w, uint32_t h, const float uv[2], const float rayIn[3], const float tbn[9],
              float out[2]);
void starting_ray(const RB200RtPushConsts& pc, float px, float py, float resx, float resy, uint32_t& rng,
                  vec3* origin_out, vec3* dir_out);
void tonemap_pixel(const float rgb[3], float exposure, uint8_t out[4]);
void postprocess(int W, int H, const float* hdr, const RB200BloomPushConsts& bloom, const RB200TonemappingPushConsts& tm,
                 uint8_t* ldr, float* combined_out, int threads);
}

using namespace oracle;

extern "C" {

#define ORACLE_API __attribute__((visibility("default")))

struct OracleCounters { uint64_t extendRays, shadowRays, paths; };

ORACLE_API int oracle_scene_create(const RB200SceneDesc* d, int bvh_threshold, void** out) {
    if (!d || !out) return -1;
    Scene* s = new Scene();
    s->vertices.assign(d->vertices, d->vertices + 4 * (size_t)d->numVertices);
    s->indices.assign(d->indices, d->indices + d->numIndices);
    s->props.assign(d->instanceProperties, d->instanceProperties + d->numInstanceProperties);
    s->tbns.assign(d->tbns, d->tbns + 9 * (size_t)d->numTbns);
    s->tbnIndices.assign(d->tbnIndices, d->tbnIndices + d->numTbnIndices);
    if (d->numEmissive) s->emissive.assign(d->emissiveMetadata, d->emissiveMetadata + d->numEmissive);
    if (d->numCdfTriangles) s->cdfTriangles.assign(d->cdfTriangles, d->cdfTriangles + d->numCdfTriangles);
    if (d->numCdfInstances) s->cdfInstances.assign(d->cdfInstances, d->cdfInstances + d->numCdfInstances);
    if (d->numTexCoords) s->texCoords.assign(d->texCoords, d->texCoords + 2 * (size_t)d->numTexCoords);
    if (d->numTexIndices) s->texIndices.assign(d->texIndices, d->texIndices + d->numTexIndices);
    for (uint32_t i = 0; i < d->numTextures; i++) {
        Scene::Tex t; t.w = d->textures[i].width; t.h = d->textures[i].height;
        t.rgba.assign(d->textures[i].rgba8, d->textures[i].rgba8 + 4 * (size_t)t.w * t.h);
        s->textures.push_back(std::move(t));
    }
    s->instances.assign(d->instances, d->instances + d->numInstances);
    flatten_scene(*s);
    if ((int)s->tris.size() > bvh_threshold) build_bvh(*s);
    *out = s;
    return 0;
}

// Switch the scene to two-level intersection (object-space BLAS per distinct model range + inverse instance transforms,
// src/scene/Scene.cpp:93-111). Returns -2 if an instance transform is singular.
ORACLE_API int oracle_scene_set_two_level(void* scene, int bvh_threshold) {
    Scene* s = (Scene*)scene;
    size_t largest = 0;
    for (const RB200Instance& in : s->instances) largest = std::max<size_t>(largest, in.triangleCount);
    return build_two_level(*s, (int)largest > bvh_threshold) ? 0 : -2;
}

ORACLE_API int oracle_scene_destroy(void* scene) { delete (Scene*)scene; return 0; }
ORACLE_API uint32_t oracle_scene_num_triangles(void* scene) { return (uint32_t)((Scene*)scene)->tris.size(); }

// One sample batch over the whole image (raytrace.rgen.glsl:250-285), rows interleaved over `threads` threads.
// tileCount > 1: only the pixels of the tiles (tileSize x tileSize, row-major index) congruent to tileRank.
ORACLE_API int oracle_render_batch_tiles(void* scene, uint32_t W, uint32_t H, uint32_t flags, const RB200RtPushConsts* pc,
                                         float* hdr, int threads, OracleCounters* counters, uint32_t tileRank,
                                         uint32_t tileCount, uint32_t tileSize);
ORACLE_API int oracle_render_batch(void* scene, uint32_t W, uint32_t H, uint32_t flags, const RB200RtPushConsts* pc,
                                   float* hdr, int threads, OracleCounters* counters) {
    return oracle_render_batch_tiles(scene, W, H, flags, pc, hdr, threads, counters, 0, 1, 32);
}
ORACLE_API int oracle_render_batch_tiles(void* scene, uint32_t W, uint32_t H, uint32_t flags, const RB200RtPushConsts* pc,
                                         float* hdr, int threads, OracleCounters* counters, uint32_t tileRank,
                                         uint32_t tileCount, uint32_t tileSize) {
    Scene* s = (Scene*)scene;
    if (!s || !pc || !hdr) return -1;
    if (tileCount == 0 || tileRank >= tileCount || tileSize == 0) return -1;
    TilePartition tiles; tiles.rank = tileRank; tiles.count = tileCount; tiles.size = tileSize;
    if ((flags & RB200_FLAG_NEE) && s->cdfTriangles.empty()) return RB200_ERR_NO_EMITTER;
    if (flags & RB200_FLAG_NEE)                 // same refusal as rb200_scene_create: nee.h.glsl:97-105 would read past the indices
        for (const RB200InstanceData& e : s->emissive)
            if (3ull * e.cdfRangeEnd + e.indexOffset + 2ull >= (unsigned long long)s->indices.size()) return RB200_ERR_INVALID_ARGUMENT;
    if (threads < 1) threads = 1;
    std::vector<Counters> cnt((size_t)threads);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&, t]() { render_rows(*s, W, H, flags, *pc, hdr, (uint32_t)t, H, (uint32_t)threads, &cnt[(size_t)t], tiles); });
    for (auto& th : pool) th.join();
    if (counters) {
        counters->extendRays = counters->shadowRays = counters->paths = 0;
        for (auto& c : cnt) { counters->extendRays += c.extendRays; counters->shadowRays += c.shadowRays; counters->paths += c.paths; }
    }
    return 0;
}

// Primary-ray closest hits of the first sample of batch pc->sampleBatch. brute != 0 forces the O(N) loop.
ORACLE_API int oracle_trace_primary(void* scene, uint32_t W, uint32_t H, const RB200RtPushConsts* pc,
                                    RB200PrimaryHit* out, int threads, int brute) {
    Scene* s = (Scene*)scene;
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) {
        pool.emplace_back([=]() {
            for (uint32_t y = (uint32_t)t; y < H; y += (uint32_t)threads)
                for (uint32_t x = 0; x < W; x++) {
                    uint32_t rng = (pc->sampleBatch * H + y) * W + x;
                    vec3 o, d;
                    starting_ray(*pc, (float)x, (float)y, (float)W, (float)H, rng, &o, &d);
                    Hit h = closest_hit(*s, o, d, 10000.0f, brute != 0);
                    RB200PrimaryHit& r = out[(size_t)y * W + x];
                    if (h.valid) {
                        r.t = h.t; r.u = h.b1; r.v = h.b2;
                        r.instance = s->tris[h.gid].instance; r.primitive = s->tris[h.gid].primitive;
                    } else { r.t = -1.0f; r.u = r.v = 0.0f; r.instance = r.primitive = 0xFFFFFFFFu; }
                }
        });
    }
    for (auto& th : pool) th.join();
    return 0;
}

ORACLE_API int oracle_trace_rays(void* scene, uint32_t n, const float* origins, const float* directions,
                                 const float* tmax, int any, RB200PrimaryHit* out, int threads, int brute) {
    Scene* s = (Scene*)scene;
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) {
        pool.emplace_back([=]() {
            for (uint32_t i = (uint32_t)t; i < n; i += (uint32_t)threads) {
                vec3 o = rb_mk3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
                vec3 d = rb_mk3(directions[3 * i], directions[3 * i + 1], directions[3 * i + 2]);
                RB200PrimaryHit& r = out[i];
                r.u = r.v = 0.0f; r.instance = r.primitive = 0xFFFFFFFFu;
                if (any) { r.t = any_hit(*s, o, d, tmax[i], brute != 0) ? 1.0f : -1.0f; continue; }
                Hit h = closest_hit(*s, o, d, tmax[i], brute != 0);
                if (h.valid) {
                    r.t = h.t; r.u = h.b1; r.v = h.b2;
                    r.instance = s->tris[h.gid].instance; r.primitive = s->tris[h.gid].primitive;
                } else r.t = -1.0f;
            }
        });
    }
    for (auto& th : pool) th.join();
    return 0;
}

ORACLE_API int oracle_postprocess(uint32_t W, uint32_t H, const float* hdr, const RB200BloomPushConsts* bloom,
                                  const RB200TonemappingPushConsts* tm, uint8_t* ldr, float* combined_out, int threads) {
    postprocess((int)W, (int)H, hdr, *bloom, *tm, ldr, combined_out, threads);
    return 0;
}

// ---- known-answer hooks ----
ORACLE_API float oracle_kat_random(uint32_t* state) { return kat_random(state); }
ORACLE_API void oracle_kat_offset(const float* p, const float* n, float* out) {
    vec3 r = kat_offset(rb_mk3(p[0], p[1], p[2]), rb_mk3(n[0], n[1], n[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
ORACLE_API void oracle_kat_sky(const float* d, float* out) {
    vec3 r = kat_sky(rb_mk3(d[0], d[1], d[2])); out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
ORACLE_API void oracle_kat_bump(const uint8_t* rgba, uint32_t w, uint32_t h, const float* uv, const float* rayIn,
                                const float* tbn, float* out) { kat_bump(rgba, w, h, uv, rayIn, tbn, out); }
ORACLE_API void oracle_kat_trace_main(void* scene, const float* o, const float* d, uint32_t* rngState, int insideDielectric,
                                      float accumulatedDistance, float* out, uint32_t* flags) {
    kat_trace_main(*(Scene*)scene, o, d, rngState, insideDielectric, accumulatedDistance, out, flags);
}
ORACLE_API void oracle_kat_bounce(void* scene, const RB200RtPushConsts* pc, const float* o, const float* d, uint32_t* rngState,
                                  int insideDielectric, float accumulatedDistance, float* out, uint32_t* flags, float* direct,
                                  uint32_t* rngAfterDirect) {
    kat_bounce(*(Scene*)scene, *pc, o, d, rngState, insideDielectric, accumulatedDistance, out, flags, direct, rngAfterDirect);
}
ORACLE_API float oracle_kat_power_heuristic(float a, float b) { return kat_power_heuristic(a, b); }
ORACLE_API void oracle_kat_light_sample(void* scene, const RB200RtPushConsts* pc, uint32_t* rngState, float* out) {
    kat_light_sample(*(Scene*)scene, *pc, rngState, out);
}
ORACLE_API void oracle_kat_direct_light_lambertian(void* scene, const RB200RtPushConsts* pc, const float* o, const float* n,
                                                   const float* albedo, uint32_t* rngState, float* out) {
    kat_direct_light_lambertian(*(Scene*)scene, *pc, o, n, albedo, rngState, out);
}
ORACLE_API void oracle_kat_tonemap(const float* rgb, float exposure, uint8_t* out) { tonemap_pixel(rgb, exposure, out); }
ORACLE_API void oracle_kat_starting_ray(const RB200RtPushConsts* pc, uint32_t x, uint32_t y, uint32_t W, uint32_t H,
                                        float* origin, float* dir, uint32_t* rng_after) {
    uint32_t rng = (pc->sampleBatch * H + y) * W + x;
    vec3 o, d;
    starting_ray(*pc, (float)x, (float)y, (float)W, (float)H, rng, &o, &d);
    origin[0] = o.x; origin[1] = o.y; origin[2] = o.z; dir[0] = d.x; dir[1] = d.y; dir[2] = d.z;
    *rng_after = rng;
}
// elementary layer, exposed so tests can compare rb_math.h against float64 libm
ORACLE_API void oracle_rb_math(int fn, uint32_t n, const float* x, float* y) {
    for (uint32_t i = 0; i < n; i++) {
        switch (fn) {
            case 0: y[i] = rb_sin(x[i]); break;
            case 1: y[i] = rb_cos(x[i]); break;
            case 2: y[i] = rb_log(x[i]); break;
            case 3: y[i] = rb_exp(x[i]); break;
            case 4: y[i] = rb_exp2(x[i]); break;
            case 5: y[i] = rb_acos(x[i]); break;
            default: y[i] = 0.0f;
        }
    }
}
ORACLE_API void oracle_rb_random(uint32_t* state, uint32_t n, float* out) {
    for (uint32_t i = 0; i < n; i++) out[i] = rb_random(state);
}

} // extern "C"
