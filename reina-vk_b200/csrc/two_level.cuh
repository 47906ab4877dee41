// two_level.cuh — traversal of the two-level structure (RB200_FLAG_TWO_LEVEL).
//
// The reference keeps one BLAS per object and one TLAS entry per instance (src/scene/Scene.cpp:93-111; instance matrix handed
// to the driver at src/tools/vktools.cpp:466-468) and the Vulkan driver moves the RAY into object space. The default path of
// this library flattens every instance into world-space triangles instead (one hierarchy, no per-instance ray set-up: what
// the tuned kernels of traverse.cuh want). This file is the other choice, for scenes whose instancing makes flattening
// too large: triangle and shading records exist once per object; a ray walks the top-level hierarchy over the instances'
// world boxes, and for every instance box it meets it is transformed with the instance's inverse matrix (t is the same
// parameter in both spaces because the direction is not re-normalised), set up again (1/d, shear constants) and walks that
// object's hierarchy. Closest-hit rule as everywhere: smallest t, ties -> smallest global primitive id (instance-major
// numbering, as the flattening numbers them), so the result does not depend on either hierarchy — the oracle's two-level
// path (oracle/intersect.cpp) tests every instance without a TLAS and must agree bit for bit.
//
// One ray per lane, per-lane stack in local memory, triangles tested where they are found: none of the warp-cooperative
// machinery of traverse.cuh (pooled triangle phase, shared-memory stacks, refill) — the set-up per visited instance is the
// badly utilised part of a traversal kernel, so this path is slower per ray by design and exists for capacity.
#pragma once
#include "traverse.cuh"

namespace rb200 {

struct TLRay {
    rb_v3 o;
    float idx, idy, idz;
    uint32_t oct_inv;
    rb_ray_shear sh;
};

__device__ __forceinline__ void tl_setup(TLRay& r, const rb_v3 o, const rb_v3 d) {
    r.o = o;
    const float ooeps = 8.2718061e-25f;   // 2^-80, as Traversal::init
    r.idx = 1.0f / (fabsf(d.x) > ooeps ? d.x : copysignf(ooeps, d.x));
    r.idy = 1.0f / (fabsf(d.y) > ooeps ? d.y : copysignf(ooeps, d.y));
    r.idz = 1.0f / (fabsf(d.z) > ooeps ? d.z : copysignf(ooeps, d.z));
    r.oct_inv = (r.idx < 0.f ? 0u : 4u) | (r.idy < 0.f ? 0u : 2u) | (r.idz < 0.f ? 0u : 1u);
    r.sh = rb_ray_prepare(d);
}

// The eight child boxes of one wide node against a ray: the arithmetic of Traversal::node_step (quantised-grid slab test,
// rounding slack, culling margin of 2^-14 of the node's extent along the ray's dominant axis), returning the hit mask in
// node_step's layout (bits 24..31 internal children permuted by the ray octant, bits 0..23 triangle bits).
__device__ __forceinline__ uint32_t tl_node_hits(const float4 n0, const float4 n1, const float4 n2, const float4 n3, const float4 n4,
                                                 const TLRay& r, const float bestT) {
    const uint32_t eim = __float_as_uint(n0.w);
    const float idx = r.idx, idy = r.idy, idz = r.idz;
    const float sx = __uint_as_float((eim & 0xFFu) << 23) * idx;
    const float sy = __uint_as_float(((eim >> 8) & 0xFFu) << 23) * idy;
    const float sz = __uint_as_float(((eim >> 16) & 0xFFu) << 23) * idz;
    const float cx = (n0.x - r.o.x) * idx, cy = (n0.y - r.o.y) * idy, cz = (n0.z - r.o.z) * idz;
    const float eps = 9.5367431640625e-07f;   // 2^-20
    const float jx = eps * fmaf(2560.0f, fabsf(sx), fabsf(cx));
    const float jy = eps * fmaf(2560.0f, fabsf(sy), fabsf(cy));
    const float jz = eps * fmaf(2560.0f, fabsf(sz), fabsf(cz));
    const float ax = fabsf(idx), ay = fabsf(idy), az = fabsf(idz);
    const float ex = fmaf(256.0f, fabsf(sx), fabsf(cx)), ey = fmaf(256.0f, fabsf(sy), fabsf(cy)), ez = fmaf(256.0f, fabsf(sz), fabsf(cz));
    const float margin = RB_CULL_MARGIN * ((ax <= ay && ax <= az) ? ex : (ay <= az ? ey : ez));
    const float kx = jx + margin, ky = jy + margin, kz = jz + margin;
    const float cnx = fmaf(-BYTE_BIAS, sx, cx - kx), cfx = fmaf(-BYTE_BIAS, sx, cx + kx);
    const float cny = fmaf(-BYTE_BIAS, sy, cy - ky), cfy = fmaf(-BYTE_BIAS, sy, cy + ky);
    const float cnz = fmaf(-BYTE_BIAS, sz, cz - kz), cfz = fmaf(-BYTE_BIAS, sz, cz + kz);
    const uint32_t oct_inv4 = r.oct_inv * 0x01010101u;
    const uint32_t kb = c_byteBiasWord;
    uint32_t hitmask = 0u;
    const float tcur = bestT + margin;
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
            const float t0x = fmaf(byte_f(nx, j, kb), sx, cnx), t1x = fmaf(byte_f(fx, j, kb), sx, cfx);
            const float t0y = fmaf(byte_f(ny, j, kb), sy, cny), t1y = fmaf(byte_f(fy, j, kb), sy, cfy);
            const float t0z = fmaf(byte_f(nz, j, kb), sz, cnz), t1z = fmaf(byte_f(fz, j, kb), sz, cfz);
            const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, 0.0f));
            const float tf = fminf(fminf(t1x, t1y), fminf(t1z, tcur));
            if (tn <= tf) {
                const uint32_t cb = (childBits4 >> (8 * j)) & 0xFFu;
                const uint32_t bi = (bitIndex4 >> (8 * j)) & 0xFFu;
                hitmask |= cb << bi;
            }
        }
    }
    return hitmask;
}

struct TLHit {
    float t, b1, b2;
    uint32_t tri;        // slot in DeviceScene::tris (= index of the shading records), 0xFFFFFFFF = miss
    uint32_t inst;       // instance index
    uint32_t gid;        // global primitive id (tie-break key)
};


// Closest hit (ANY = false) or any hit (ANY = true) of one ray against the two-level structure.
template <bool ANY>
__device__ __noinline__ TLHit trace_two_level(const DeviceScene& S, const rb_v3 org, const rb_v3 dir, const float tmax) {
    TLHit best;
    best.t = tmax; best.b1 = 0.f; best.b2 = 0.f; best.tri = 0xFFFFFFFFu; best.inst = 0xFFFFFFFFu; best.gid = 0xFFFFFFFFu;
    unsigned long long bestKey = hit_key(tmax, 0xFFFFFFFFu);
    TLRay world, cur;
    tl_setup(world, org, dir);
    cur = world;
    bool inObject = false;
    uint32_t inst = 0u, gidBase = 0u;
    uint2 stack[TL_STACK];
    int sp = 0;
    uint2 ngroup = make_uint2(0u, 0x80000000u);       // root of the top level
    for (;;) {
        if (ngroup.y <= 0x00FFFFFFu) {                // no node group current: next entry of the stack
            if (sp == 0) break;
            uint2 e = stack[--sp];
            if (e.y > 0x00FFFFFFu) { ngroup = e; continue; }
            if (e.y == 0u) { cur = world; inObject = false; continue; }      // sentinel: the instance is done, back to world space
            // a group of instance boxes found in a top-level node: enter the next one
            const uint32_t ti = 31u - (uint32_t)__clz(e.y);
            e.y &= ~(1u << ti);
            if (e.y) stack[sp++] = e;
            inst = __float_as_uint(__ldg(reinterpret_cast<const float4*>(S.tlasLeaves + (e.x + ti))).w);
            const TwoLevelInstance* en = S.tlInstances + inst;
            float inv[12];
            const float4 i0 = __ldg(reinterpret_cast<const float4*>(en)), i1 = __ldg(reinterpret_cast<const float4*>(en) + 1),
                         i2 = __ldg(reinterpret_cast<const float4*>(en) + 2), i3 = __ldg(reinterpret_cast<const float4*>(en) + 3);
            inv[0] = i0.x; inv[1] = i0.y; inv[2] = i0.z; inv[3] = i0.w; inv[4] = i1.x; inv[5] = i1.y; inv[6] = i1.z; inv[7] = i1.w;
            inv[8] = i2.x; inv[9] = i2.y; inv[10] = i2.z; inv[11] = i2.w;
            gidBase = __float_as_uint(i3.y);
            tl_setup(cur, rb_inv_point(inv, org), rb_inv_vector(inv, dir));
            inObject = true;
            stack[sp++] = make_uint2(0u, 0u);         // sentinel
            ngroup = make_uint2(__float_as_uint(i3.x), 0x80000000u);
            continue;
        }
        // one wide-node step at the current level (Traversal::node_step without the shared-memory staging)
        const uint32_t hits = ngroup.y;
        const uint32_t bitIndex = 31u - (uint32_t)__clz(hits);
        const uint32_t base = ngroup.x;
        ngroup.y &= ~(1u << bitIndex);
        if (ngroup.y > 0x00FFFFFFu) stack[sp++] = ngroup;
        const uint32_t slot = (bitIndex - 24u) ^ cur.oct_inv;
        const uint32_t rel = __popc(hits & ~(0xFFFFFFFFu << slot) & 0xFFu);
        const float4* np = reinterpret_cast<const float4*>((inObject ? S.nodes : S.tlasNodes) + (base + rel));
        const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
        const uint32_t hitmask = tl_node_hits(n0, n1, n2, n3, n4, cur, best.t);
        ngroup = make_uint2(__float_as_uint(n1.x), (hitmask & 0xFF000000u) | (__float_as_uint(n0.w) >> 24));
        uint32_t tmask = hitmask & 0x00FFFFFFu;
        if (tmask == 0u) continue;
        const uint32_t triBase = __float_as_uint(n1.y);
        if (!inObject) { stack[sp++] = make_uint2(triBase, tmask); continue; }      // instance boxes: entered one at a time
        while (tmask) {
            const uint32_t ti = 31u - (uint32_t)__clz(tmask);
            tmask &= ~(1u << ti);
            const uint32_t triIdx = triBase + ti;
            const float4* tp = reinterpret_cast<const float4*>(S.tris + triIdx);
            const float4 va = __ldg(tp), vb = __ldg(tp + 1), vc = __ldg(tp + 2);
            float t, b1, b2;
            if (rb_tri_intersect(cur.o, cur.sh, rb_mk3(va.x, va.y, va.z), rb_mk3(vb.x, vb.y, vb.z), rb_mk3(vc.x, vc.y, vc.z), &t, &b1, &b2)) {
                if (t > 0.0f && t < tmax) {
                    if (ANY) { best.t = t; best.tri = triIdx; best.inst = inst; return best; }
                    const uint32_t gid = gidBase + __float_as_uint(va.w);
                    const unsigned long long key = hit_key(t, gid);
                    if (key < bestKey) { bestKey = key; best.t = t; best.b1 = b1; best.b2 = b2; best.tri = triIdx; best.inst = inst; best.gid = gid; }
                }
            }
        }
    }
    return best;
}

} // namespace rb200
