// Host compilation of the kernels' pure shading functions (csrc/shade.cuh) next to the oracle's (oracle/disney.h,
// pathtrace.cpp is not needed): random inputs, bit-for-bit comparison.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T tex2D(cudaTextureObject_t, float, float) { return T(); }
#define __forceinline__ inline
#define __noinline__
#include "shade.cuh"
#include "oracle_common.h"
#include "disney.h"

namespace oracle { void kat_light_sample(const Scene& s, const RB200RtPushConsts& pc, uint32_t* rng, float out[11]); }

// csrc/shade.cuh random_emissive_point against the oracle's (oracle/pathtrace.cpp, linked in) on a scene with THREE emitters
// whose CDF ranges do not start at 0: the reference's concatenated-CDF addressing must come out the same on both sides.
static long light_sampling_mismatches(std::mt19937& gen) {
    std::uniform_real_distribution<float> U(0.f, 1.f);
    const uint32_t trisPer[4] = {4, 6, 3, 40};                 // three emitters, then a filler mesh the slots run into
    oracle::Scene os;
    uint32_t indexOffset[4], off = 0;
    for (int m = 0; m < 4; m++) {
        indexOffset[m] = off;
        for (uint32_t t = 0; t < trisPer[m]; t++)
            for (int k = 0; k < 3; k++) {
                os.indices.push_back(uint32_t(os.vertices.size() / 4));
                os.vertices.insert(os.vertices.end(), {U(gen) * 4 - 2, U(gen) * 4 - 2, U(gen) * 4 - 2, 1.0f});
            }
        off += 3 * trisPer[m];
    }
    float total = 0, cum = 0;
    for (int m = 0; m < 3; m++) {
        RB200InstanceData e{};
        for (int k = 0; k < 16; k++) e.transform[k] = (k % 5 == 0 ? 1.0f : 0.0f) + 0.3f * (U(gen) - 0.5f);
        e.transform[3] = e.transform[7] = e.transform[11] = 0.0f; e.transform[15] = 1.0f;
        e.cdfRangeStart = uint32_t(os.cdfTriangles.size());
        float c = 0;
        std::vector<float> w(trisPer[m]);
        for (float& x : w) { x = 0.1f + U(gen); c += x; }
        float run = 0;
        for (float x : w) { run += x; os.cdfTriangles.push_back(run / c); }
        os.cdfTriangles.back() = 1.0f;
        e.cdfRangeEnd = uint32_t(os.cdfTriangles.size()) - 1;
        e.indexOffset = indexOffset[m];
        e.emission[0] = U(gen) * 9; e.emission[1] = U(gen) * 9; e.emission[2] = U(gen) * 9;
        e.weight = 0.5f + U(gen); e.area = 0.2f + U(gen); e.cullBackface = m & 1;
        total += e.weight;
        os.emissive.push_back(e);
    }
    for (int m = 0; m < 3; m++) { cum += os.emissive[m].weight; os.cdfInstances.push_back(cum / total); }
    os.cdfInstances.back() = 1.0f;
    rb200::DeviceScene S{};
    S.vertices = reinterpret_cast<const float4*>(os.vertices.data());
    S.indices = os.indices.data();
    S.emissive = os.emissive.data();
    S.cdfTriangles = os.cdfTriangles.data();
    S.cdfInstances = os.cdfInstances.data();
    S.numCdfInstances = 3;
    RB200RtPushConsts pc{};
    pc.totalEmissiveWeight = total;
    long bad = 0;
    bool seen[3] = {false, false, false};
    for (int it = 0; it < 20000; it++) {
        uint32_t r1 = gen(), r2 = r1;
        const rb200::LightSample a = rb200::random_emissive_point(S, total, r1);
        float o[11];
        oracle::kat_light_sample(os, pc, &r2, o);
        const float mine[11] = {a.point.x, a.point.y, a.point.z, a.normal.x, a.normal.y, a.normal.z, a.emission.x, a.emission.y, a.emission.z,
                                a.pdf, a.cullBackface ? 1.0f : 0.0f};
        if (r1 != r2 || memcmp(mine, o, sizeof mine) != 0) bad++;
        for (int m = 0; m < 3; m++) if (a.emission.x == os.emissive[m].emission[0]) seen[m] = true;
    }
    return bad + ((seen[0] && seen[1] && seen[2]) ? 0 : 1000000);
}

static uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static bool same(float a, float b) { return bits(a) == bits(b) || (a != a && b != b); }

int main() {
    std::mt19937 gen(7);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    auto rv = [&]() { rb_v3 v = rb_mk3(U(gen) * 2 - 1, U(gen) * 2 - 1, U(gen) * 2 - 1); return rb_normalize(v); };
    long bad_eval = 0, bad_sample = 0, bad_fuzzy = 0, n = 200000;
    for (long it = 0; it < n; it++) {
        rb_m3 tbn; rb_v3 N = rv();
        rb_v3 t = rb_normalize(rb_cross(N, fabsf(N.x) < 0.5f ? rb_mk3(1, 0, 0) : rb_mk3(0, 1, 0)));
        tbn.c0 = t; tbn.c1 = rb_cross(N, t); tbn.c2 = N;
        rb200::DisneyP p;
        p.baseColor = rb_mk3(U(gen), U(gen), U(gen)); p.specularTint = rb_mk3(U(gen), U(gen), U(gen)); p.sheenTint = rb_mk3(U(gen), U(gen), U(gen));
        p.anisotropic = U(gen); p.roughness = 0.02f + U(gen); p.subsurface = U(gen); p.clearcoatGloss = U(gen); p.eta = 0.5f + 1.5f * U(gen);
        p.metallic = U(gen); p.clearcoat = U(gen); p.specularTransmission = U(gen); p.sheen = U(gen);
        if (it % 7 == 0) { p.metallic = 0; p.clearcoat = 0; p.specularTransmission = 0; }
        oracle::DisneyParams q;
        q.baseColor = p.baseColor; q.specularTint = p.specularTint; q.sheenTint = p.sheenTint; q.anisotropic = p.anisotropic; q.roughness = p.roughness;
        q.subsurface = p.subsurface; q.clearcoatGloss = p.clearcoatGloss; q.eta = p.eta; q.metallic = p.metallic; q.clearcoat = p.clearcoat;
        q.specularTransmission = p.specularTransmission; q.sheen = p.sheen;
        rb_v3 wi = rv(); if (rb_dot(wi, N) < 0 && it % 3) wi = -wi;
        wi = wi * (0.5f + 1.5f * U(gen));        // un-normalised, as the reference passes it
        uint32_t r1 = gen(), r2 = r1;
        bool dr1, cg1, dr2, cg2;
        rb_v3 wo1 = rb200::disney_sample(tbn, p, N, wi, &dr1, &cg1, r1);
        rb_v3 wo2 = oracle::disney_sample(tbn, q, N, wi, &dr2, &cg2, r2);
        if (!(same(wo1.x, wo2.x) && same(wo1.y, wo2.y) && same(wo1.z, wo2.z)) || r1 != r2 || dr1 != dr2 || cg1 != cg2) bad_sample++;
        rb_v3 h = rb_normalize(wo1 + wi);
        float pdf1, pdf2;
        rb_v3 f1 = rb200::disney_eval(tbn, p, dr1, N, wi, wo1, h, &pdf1);
        rb_v3 f2 = oracle::disney_eval(tbn, q, dr1, N, wi, wo1, h, &pdf2);
        if (!(same(f1.x, f2.x) && same(f1.y, f2.y) && same(f1.z, f2.z) && same(pdf1, pdf2))) bad_eval++;
        uint32_t s1 = gen(), s0 = s1;
        rb_v3 fr = rb200::fuzzy_reflection(wi, N, p.roughness, s1);
        if (s1 != s0 || !(fr.x == fr.x)) bad_fuzzy++;
    }
    printf("hits %ld  disney_sample mismatches %ld  disney_eval mismatches %ld  fuzzy_reflection state not restored %ld\n", n, bad_sample, bad_eval, bad_fuzzy);
    const long bad_light = light_sampling_mismatches(gen);
    printf("light samples 20000 (three emitters)  random_emissive_point mismatches %ld\n", bad_light);
    return (bad_eval || bad_sample || bad_fuzzy || bad_light) ? 1 : 0;
}
