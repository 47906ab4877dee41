/*
 * rb_tri.h — watertight ray/triangle intersection (Woop, Benthin, Wald 2013), one definition for device + host.
 *
 * In the reference this arithmetic lives inside the Vulkan driver / RT cores (traceRayEXT at
 * shaders/raytrace/raytrace.rgen.glsl:110-122 and shaders/raytrace/nee.h.glsl:129-141); Vulkan only promises
 * watertightness, not bit patterns, so the rule is ours to fix:
 *   - the ray is sheared so its dominant axis becomes +z, edge functions U, V, W are evaluated in fp32 with
 *     plain (uncontracted) products so that a shared edge yields exactly opposite values in both triangles;
 *   - if any edge function is exactly 0 it is re-evaluated in fp64 (the paper's fallback);
 *   - no back-face culling (the reference culls in its closest-hit shaders, not in traversal);
 *   - t = T / det, barycentrics (b1, b2) = (V, W) / det are the Vulkan hit attributes (weights of v1, v2);
 *   - the caller applies tmin < t < tmax and the closest-hit tie rule (equal t -> smaller global primitive id),
 *     which makes the result independent of traversal order and therefore of the BVH.
 * Directions need not be unit length (metal/dielectric leave them unnormalised,
 * shaders/raytrace/metal.rchit.glsl:53); t is in units of |dir|.
 */
#ifndef RB_TRI_H
#define RB_TRI_H

#include "rb_vec.h"

struct rb_ray_shear {
    int kx, ky, kz;
    float Sx, Sy, Sz;
};

RB_HD float rb_sel3(rb_v3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }

RB_HD rb_ray_shear rb_ray_prepare(rb_v3 d) {
    rb_ray_shear s;
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
    int kx = kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    float dz = rb_sel3(d, kz);
    if (dz < 0.0f) { int t = kx; kx = ky; ky = t; }
    s.kx = kx; s.ky = ky; s.kz = kz;
    s.Sx = rb_sel3(d, kx) / dz;
    s.Sy = rb_sel3(d, ky) / dz;
    s.Sz = 1.0f / dz;
    return s;
}

/* Returns true when the (infinite) ray line crosses the triangle with det != 0; outputs t, b1, b2. */
RB_HD bool rb_tri_intersect(rb_v3 org, const rb_ray_shear& s, rb_v3 v0, rb_v3 v1, rb_v3 v2,
                            float* t_out, float* b1_out, float* b2_out) {
    rb_v3 A = v0 - org, B = v1 - org, C = v2 - org;
    float Akz = rb_sel3(A, s.kz), Bkz = rb_sel3(B, s.kz), Ckz = rb_sel3(C, s.kz);
    float Ax = fmaf(-s.Sx, Akz, rb_sel3(A, s.kx));
    float Ay = fmaf(-s.Sy, Akz, rb_sel3(A, s.ky));
    float Bx = fmaf(-s.Sx, Bkz, rb_sel3(B, s.kx));
    float By = fmaf(-s.Sy, Bkz, rb_sel3(B, s.ky));
    float Cx = fmaf(-s.Sx, Ckz, rb_sel3(C, s.kx));
    float Cy = fmaf(-s.Sy, Ckz, rb_sel3(C, s.ky));

    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;

    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        double CxBy = (double)Cx * (double)By, CyBx = (double)Cy * (double)Bx;
        U = (float)(CxBy - CyBx);
        double AxCy = (double)Ax * (double)Cy, AyCx = (double)Ay * (double)Cx;
        V = (float)(AxCy - AyCx);
        double BxAy = (double)Bx * (double)Ay, ByAx = (double)By * (double)Ax;
        W = (float)(BxAy - ByAx);
    }

    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    float det = U + V + W;
    if (det == 0.0f) return false;

    float Az = s.Sz * Akz, Bz = s.Sz * Bkz, Cz = s.Sz * Ckz;
    float T = U * Az + V * Bz + W * Cz;
    *t_out = T / det;
    *b1_out = V / det;
    *b2_out = W / det;
    return true;
}

#endif /* RB_TRI_H */
