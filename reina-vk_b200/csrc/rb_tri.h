/*
 * rb_tri.h — watertight ray/triangle intersection (Woop, Benthin, Wald 2013), one definition for device + host.
 *
 * In the reference this arithmetic lives inside the Vulkan driver / RT cores (traceRayEXT at
 * shaders/raytrace/raytrace.rgen.glsl:110-122 and shaders/raytrace/nee.h.glsl:129-141); Vulkan only promises
 * watertightness, not bit patterns, so the rule is ours to fix:
 *   - the ray is sheared so its dominant axis becomes +z (each vertex is mapped by the same three row products, so
 *     the mapping is a function of (vertex, ray) only), edge functions U, V, W are evaluated in fp32 with plain
 *     (uncontracted) products so that a shared edge yields exactly opposite values in both triangles;
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
    /* rows of the shear+permutation that maps the ray onto +z: x' = mx . p, y' = my . p, z' = mz . p.
     * mx = e_kx - Sx e_kz, my = e_ky - Sy e_kz, mz = Sz e_kz with kz the dominant axis of the direction
     * (Woop et al. 2013, eq. 2-3); stored as dense rows so a vertex needs 3 multiply-adds per coordinate and no
     * per-vertex axis selects. */
    rb_v3 mx, my, mz;
    float Sz; int kz;          /* mz = Sz e_kz, kept so the row can be stored in two words */
};

RB_HD float rb_sel3(rb_v3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
RB_HD rb_v3 rb_axis3(int k, float val) { return rb_mk3(k == 0 ? val : 0.0f, k == 1 ? val : 0.0f, k == 2 ? val : 0.0f); }

RB_HD rb_ray_shear rb_ray_prepare(rb_v3 d) {
    rb_ray_shear s;
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
    int kx = kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    float dz = rb_sel3(d, kz);
    if (dz < 0.0f) { int t = kx; kx = ky; ky = t; }     /* keep the winding */
    const float Sx = rb_sel3(d, kx) / dz;
    const float Sy = rb_sel3(d, ky) / dz;
    const float Sz = 1.0f / dz;
    s.mx = rb_axis3(kx, 1.0f) + rb_axis3(kz, -Sx);
    s.my = rb_axis3(ky, 1.0f) + rb_axis3(kz, -Sy);
    s.mz = rb_axis3(kz, Sz);
    s.Sz = Sz; s.kz = kz;
    return s;
}

/* m . p with a fixed evaluation order (one multiply, two fused multiply-adds) */
RB_HD float rb_row3(rb_v3 m, rb_v3 p) { return fmaf(m.z, p.z, fmaf(m.y, p.y, m.x * p.x)); }

/* The test up to (not including) its three divisions: returns true when the (infinite) ray line crosses the triangle
 * with det != 0 and yields det, T and the edge functions V, W: t = T / det, (b1, b2) = (V, W) / det. Split out so that
 * the traversal kernel can postpone the barycentric divisions to the one candidate that ends up closest; the
 * operands, and therefore the quotients, are the same either way. */
RB_HD bool rb_tri_edges(rb_v3 org, const rb_ray_shear& s, rb_v3 v0, rb_v3 v1, rb_v3 v2,
                        float* det_out, float* T_out, float* V_out, float* W_out) {
    const rb_v3 A = v0 - org, B = v1 - org, C = v2 - org;
    const float Ax = rb_row3(s.mx, A), Ay = rb_row3(s.my, A);
    const float Bx = rb_row3(s.mx, B), By = rb_row3(s.my, B);
    const float Cx = rb_row3(s.mx, C), Cy = rb_row3(s.my, C);

    /* edge functions: plain products and differences, so a shared edge gives exactly opposite values */
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
    const float det = U + V + W;
    if (det == 0.0f) return false;

    const float Az = rb_row3(s.mz, A), Bz = rb_row3(s.mz, B), Cz = rb_row3(s.mz, C);
    *T_out = U * Az + V * Bz + W * Cz;
    *det_out = det; *V_out = V; *W_out = W;
    return true;
}

/* Returns true when the (infinite) ray line crosses the triangle with det != 0; outputs t, b1, b2. */
RB_HD bool rb_tri_intersect(rb_v3 org, const rb_ray_shear& s, rb_v3 v0, rb_v3 v1, rb_v3 v2,
                            float* t_out, float* b1_out, float* b2_out) {
    float det, T, V, W;
    if (!rb_tri_edges(org, s, v0, v1, v2, &det, &T, &V, &W)) return false;
    *t_out = T / det;
    *b1_out = V / det;
    *b2_out = W / det;
    return true;
}

#endif /* RB_TRI_H */
