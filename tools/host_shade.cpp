// Host compilation of the kernels' pure shading functions (csrc/shade.cuh) next to the oracle's (oracle/disney.h,
// pathtrace.cpp is not needed): random inputs, bit-for-bit comparison.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <random>
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T tex2D(cudaTextureObject_t, float, float) { return T(); }
#define __forceinline__ inline
#define __noinline__
#include "shade.cuh"
#include "oracle_common.h"
#include "disney.h"

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
    return (bad_eval || bad_sample || bad_fuzzy) ? 1 : 0;
}
