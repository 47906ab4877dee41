// ORACLE (test infrastructure) — ray/scene intersection on the CPU.
//
// Stands in for the Vulkan driver's TLAS/BLAS traversal behind traceRayEXT
// (shaders/raytrace/raytrace.rgen.glsl:110-122 closest hit, shaders/raytrace/nee.h.glsl:126-144 any hit;
// acceleration structures built at src/graphics/Blas.cpp:8-124 and src/tools/vktools.cpp:460-596).
// Rule (see rb_tri.h): world-space triangles, watertight test, hit iff tmin(0) < t < tmax, closest hit =
// smallest t, equal t -> smallest global primitive id. With that rule the answer does not depend on the
// acceleration structure, so a brute-force loop (small scenes) and a binned-SAH BVH (large scenes) agree.
#include "oracle_common.h"
#include <algorithm>
#include <cmath>
#include <cfloat>

namespace oracle {

// Instances are flattened into world space in instance order, triangles in model order; a triangle's global
// id is its position in that list. Vertex transform: column-major mat4 * (v, 1) (Instance::getTransform,
// src/scene/Instance.cpp; TLAS transform src/tools/vktools.cpp:466-468).
void flatten_scene(Scene& s) {
    s.tris.clear();
    for (uint32_t i = 0; i < s.instances.size(); i++) {
        const RB200Instance& in = s.instances[i];
        for (uint32_t p = 0; p < in.triangleCount; p++) {
            WorldTri t;
            vec3 v[3];
            for (int k = 0; k < 3; k++) {
                uint32_t vi = s.indices[3 * p + in.indexOffset + k];
                vec3 o = rb_mk3(s.vertices[4 * vi + 0], s.vertices[4 * vi + 1], s.vertices[4 * vi + 2]);
                v[k] = rb_m4_point(in.transform, o);
            }
            t.v0 = v[0]; t.v1 = v[1]; t.v2 = v[2];
            t.instance = i; t.primitive = p;
            s.tris.push_back(t);
        }
    }
}

// Two-level mode: the distinct model ranges in object space (flattened with the identity, so the vertex values go through
// the same rb_m4_point as the kernels' BLAS build) and the inverse transform of every instance.
bool build_two_level(Scene& s, bool withBvh) {
    static const float identity[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    s.blas.clear(); s.tl.clear();
    std::vector<std::pair<uint32_t, uint32_t>> ranges;
    uint32_t gidBase = 0;
    for (uint32_t i = 0; i < s.instances.size(); i++) {
        const RB200Instance& in = s.instances[i];
        const std::pair<uint32_t, uint32_t> key(in.indexOffset, in.triangleCount);
        uint32_t b = 0;
        while (b < ranges.size() && ranges[b] != key) b++;
        if (b == ranges.size()) {
            ranges.push_back(key);
            s.blas.emplace_back();
            Scene& sub = s.blas.back();
            for (uint32_t p = 0; p < in.triangleCount; p++) {
                WorldTri t;
                vec3 v[3];
                for (int k = 0; k < 3; k++) {
                    uint32_t vi = s.indices[3 * p + in.indexOffset + k];
                    v[k] = rb_m4_point(identity, rb_mk3(s.vertices[4 * vi + 0], s.vertices[4 * vi + 1], s.vertices[4 * vi + 2]));
                }
                t.v0 = v[0]; t.v1 = v[1]; t.v2 = v[2]; t.instance = i; t.primitive = p;
                sub.tris.push_back(t);
            }
            if (withBvh) build_bvh(sub);
        }
        Scene::TwoLevelInstance ti;
        if (!rb_affine_inverse(in.transform, ti.inv)) return false;
        ti.blas = b; ti.gidBase = gidBase;
        s.tl.push_back(ti);
        gidBase += in.triangleCount;
    }
    s.twoLevel = true;
    return true;
}

// ---------------------------------------------------------------------------------------------------
// binned-SAH binary BVH (16 bins), leaves of <= 4 triangles
// ---------------------------------------------------------------------------------------------------
namespace {
struct Box {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    void grow(const vec3& p) {
        lo[0] = std::min(lo[0], p.x); lo[1] = std::min(lo[1], p.y); lo[2] = std::min(lo[2], p.z);
        hi[0] = std::max(hi[0], p.x); hi[1] = std::max(hi[1], p.y); hi[2] = std::max(hi[2], p.z);
    }
    void grow(const Box& b) {
        for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); }
    }
    double area() const {
        double dx = (double)hi[0] - lo[0], dy = (double)hi[1] - lo[1], dz = (double)hi[2] - lo[2];
        if (dx < 0) return 0.0;
        return 2.0 * (dx * dy + dy * dz + dz * dx);
    }
};

struct Builder {
    Scene& s;
    std::vector<Box> tbox;
    std::vector<float> cen;   // 3 per tri
    std::vector<uint32_t>& order;
    Builder(Scene& sc) : s(sc), order(sc.bvhTriOrder) {}

    void subdivide(uint32_t nodeIdx, uint32_t first, uint32_t count) {
        Box nb, cb;
        for (uint32_t i = first; i < first + count; i++) {
            uint32_t t = order[i];
            nb.grow(tbox[t]);
            vec3 c = rb_mk3(cen[3 * t], cen[3 * t + 1], cen[3 * t + 2]);
            cb.grow(c);
        }
        BvhNode& n0 = s.nodes[nodeIdx];
        for (int a = 0; a < 3; a++) { n0.lo[a] = nb.lo[a]; n0.hi[a] = nb.hi[a]; }
        if (count <= 4) { n0.left = first; n0.count = count; return; }

        const int NB = 16;
        int bestAxis = -1, bestSplit = -1;
        double bestCost = 1e300;
        for (int a = 0; a < 3; a++) {
            float cmin = cb.lo[a], cmax = cb.hi[a];
            if (!(cmax > cmin)) continue;
            Box bins[NB]; uint32_t cnt[NB] = {0};
            float scale = NB / (cmax - cmin);
            for (uint32_t i = first; i < first + count; i++) {
                uint32_t t = order[i];
                int b = std::min(NB - 1, (int)((cen[3 * t + a] - cmin) * scale));
                bins[b].grow(tbox[t]); cnt[b]++;
            }
            double leftArea[NB - 1], rightArea[NB - 1]; uint32_t leftCnt[NB - 1], rightCnt[NB - 1];
            Box lb, rb; uint32_t lc = 0, rc = 0;
            for (int i = 0; i < NB - 1; i++) {
                lb.grow(bins[i]); lc += cnt[i]; leftArea[i] = lb.area(); leftCnt[i] = lc;
                rb.grow(bins[NB - 1 - i]); rc += cnt[NB - 1 - i]; rightArea[NB - 2 - i] = rb.area(); rightCnt[NB - 2 - i] = rc;
            }
            for (int i = 0; i < NB - 1; i++) {
                if (leftCnt[i] == 0 || rightCnt[i] == 0) continue;
                double cost = leftArea[i] * leftCnt[i] + rightArea[i] * rightCnt[i];
                if (cost < bestCost) { bestCost = cost; bestAxis = a; bestSplit = i; }
            }
        }
        uint32_t mid;
        if (bestAxis < 0) {
            mid = first + count / 2;   // all centroids coincide: split by position in the list
        } else {
            float cmin = cb.lo[bestAxis], cmax = cb.hi[bestAxis];
            float scale = NB / (cmax - cmin);
            auto it = std::partition(order.begin() + first, order.begin() + first + count, [&](uint32_t t) {
                int b = std::min(NB - 1, (int)((cen[3 * t + bestAxis] - cmin) * scale));
                return b <= bestSplit;
            });
            mid = (uint32_t)(it - order.begin());
            if (mid == first || mid == first + count) mid = first + count / 2;
        }
        uint32_t left = (uint32_t)s.nodes.size();
        s.nodes.push_back(BvhNode()); s.nodes.push_back(BvhNode());
        s.nodes[nodeIdx].left = left; s.nodes[nodeIdx].count = 0;
        subdivide(left, first, mid - first);
        subdivide(left + 1, mid, first + count - mid);
    }
};
} // namespace

void build_bvh(Scene& s) {
    uint32_t n = (uint32_t)s.tris.size();
    s.nodes.clear(); s.bvhTriOrder.resize(n);
    if (n == 0) return;
    Builder b(s);
    b.tbox.resize(n); b.cen.resize(3 * (size_t)n);
    for (uint32_t i = 0; i < n; i++) {
        s.bvhTriOrder[i] = i;
        const WorldTri& t = s.tris[i];
        b.tbox[i].grow(t.v0); b.tbox[i].grow(t.v1); b.tbox[i].grow(t.v2);
        for (int a = 0; a < 3; a++) b.cen[3 * i + a] = 0.5f * (b.tbox[i].lo[a] + b.tbox[i].hi[a]);
    }
    s.nodes.reserve(2 * (size_t)n);
    s.nodes.push_back(BvhNode());
    b.subdivide(0, 0, n);
    s.useBvh = true;
}

// Conservative fp64 slab test. The watertight triangle test may accept rays that graze a triangle by a few
// fp32 ulps of |vertex - origin|, so boxes are padded by a slack proportional to the coordinates involved.
static inline bool box_hit(const BvhNode& n, const double o[3], const double inv[3], double tmax, double slackScale,
                           double* tnear) {
    // Every box is dilated by a margin in ray parameters, and so is the best-t limit: the watertight test computes t as the
    // barycentric mean of the vertices' ray parameters, with an fp32 error proportional to the triangle's extent in ray
    // parameters (not to t). A triangle whose computed t beats the best so far can sit in a box that starts a few 1e-8 beyond
    // it (a ray leaving a large floor triangle), and a ray that starts 1e-7 behind a two-metre triangle and moves away from
    // it is reported at t = +4e-8. 2^-14 of the box's extent in ray parameters, as in csrc/traverse.cuh.
    double t0 = -1e300, t1 = 1e300, ext = 0.0;
    for (int a = 0; a < 3; a++) {
        double m = std::max(std::max(std::fabs((double)n.lo[a]), std::fabs((double)n.hi[a])), std::fabs(o[a]));
        double slack = slackScale * m + 1e-30;
        double ta = ((double)n.lo[a] - slack - o[a]) * inv[a];
        double tb = ((double)n.hi[a] + slack - o[a]) * inv[a];
        if (ta != ta || tb != tb) continue;            // 0 * inf: origin on the slab plane, direction parallel -> inside
        if (ta > tb) std::swap(ta, tb);
        // the vertices' ray parameters are measured along the dominant axis of the ray (smallest |1 / d|): its slab bounds them
        if (std::fabs(inv[a]) <= std::fabs(inv[(a + 1) % 3]) && std::fabs(inv[a]) <= std::fabs(inv[(a + 2) % 3]))
            ext = std::max(ext, std::max(std::fabs(ta), std::fabs(tb)));
        t0 = std::max(t0, ta); t1 = std::min(t1, tb);
    }
    const double margin = 6.103515625e-05 * ext;        // 2^-14
    t0 -= margin; t1 += margin;
    *tnear = std::max(t0, 0.0);
    return t0 <= t1 * (1.0 + 1e-9) + 1e-30 && t1 >= 0.0 && t0 <= (tmax + margin) * (1.0 + 1e-9) + 1e-30;
}

static inline void test_tri(const Scene& s, uint32_t gid, vec3 org, const rb_ray_shear& sh, float tmax, Hit& best) {
    const WorldTri& tr = s.tris[gid];
    float t, b1, b2;
    if (!rb_tri_intersect(org, sh, tr.v0, tr.v1, tr.v2, &t, &b1, &b2)) return;
    if (!(t > 0.0f && t < tmax)) return;
    if (!best.valid || t < best.t || (t == best.t && gid < best.gid)) {
        best.valid = true; best.t = t; best.b1 = b1; best.b2 = b2; best.gid = gid;
    }
}

// Two-level closest hit: every instance in order, the ray in that instance's object space, t < best so far (a later
// instance has larger global ids, so it loses a tie: strict is the closest-hit rule). No culling of instances: the answer
// must not depend on a TLAS.
static Hit closest_hit_two_level(const Scene& s, vec3 org, vec3 dir, float tmax, bool brute, uint64_t* tri_tests) {
    Hit best; best.valid = false; best.t = tmax; best.b1 = best.b2 = 0; best.gid = 0xFFFFFFFFu;
    for (size_t i = 0; i < s.tl.size(); i++) {
        const Scene::TwoLevelInstance& ti = s.tl[i];
        const Hit h = closest_hit(s.blas[ti.blas], rb_inv_point(ti.inv, org), rb_inv_vector(ti.inv, dir), best.t, brute, tri_tests);
        if (h.valid) { best = h; best.gid = ti.gidBase + h.gid; }
    }
    return best;
}

Hit closest_hit(const Scene& s, vec3 org, vec3 dir, float tmax, bool brute, uint64_t* tri_tests) {
    if (s.twoLevel) return closest_hit_two_level(s, org, dir, tmax, brute, tri_tests);
    Hit best; best.valid = false; best.t = tmax; best.b1 = best.b2 = 0; best.gid = 0xFFFFFFFFu;
    rb_ray_shear sh = rb_ray_prepare(dir);
    if (brute || !s.useBvh) {
        for (uint32_t g = 0; g < s.tris.size(); g++) test_tri(s, g, org, sh, tmax, best);
        if (tri_tests) *tri_tests += s.tris.size();
        return best;
    }
    const double o[3] = {org.x, org.y, org.z};
    const double inv[3] = {1.0 / (double)dir.x, 1.0 / (double)dir.y, 1.0 / (double)dir.z};
    const double slackScale = 4e-6;
    uint32_t stack[128]; int sp = 0;
    stack[sp++] = 0;
    while (sp) {
        uint32_t ni = stack[--sp];
        const BvhNode& n = s.nodes[ni];
        double tn;
        // <= best.t (not <): equal-t candidates with a smaller id must still be visited
        if (!box_hit(n, o, inv, best.valid ? (double)best.t : (double)tmax, slackScale, &tn)) continue;
        if (n.count) {
            for (uint32_t i = 0; i < n.count; i++) test_tri(s, s.bvhTriOrder[n.left + i], org, sh, tmax, best);
            if (tri_tests) *tri_tests += n.count;
        } else {
            double ta, tb;
            bool ha = box_hit(s.nodes[n.left], o, inv, 1e300, slackScale, &ta);
            bool hb = box_hit(s.nodes[n.left + 1], o, inv, 1e300, slackScale, &tb);
            (void)ha; (void)hb;
            if (ta < tb) { stack[sp++] = n.left + 1; stack[sp++] = n.left; }
            else         { stack[sp++] = n.left; stack[sp++] = n.left + 1; }
        }
    }
    return best;
}

bool any_hit(const Scene& s, vec3 org, vec3 dir, float tmax, bool brute) {
    if (s.twoLevel) {
        for (size_t i = 0; i < s.tl.size(); i++) {
            const Scene::TwoLevelInstance& ti = s.tl[i];
            if (any_hit(s.blas[ti.blas], rb_inv_point(ti.inv, org), rb_inv_vector(ti.inv, dir), tmax, brute)) return true;
        }
        return false;
    }
    rb_ray_shear sh = rb_ray_prepare(dir);
    if (brute || !s.useBvh) {
        for (uint32_t g = 0; g < s.tris.size(); g++) {
            const WorldTri& tr = s.tris[g];
            float t, b1, b2;
            if (rb_tri_intersect(org, sh, tr.v0, tr.v1, tr.v2, &t, &b1, &b2) && t > 0.0f && t < tmax) return true;
        }
        return false;
    }
    const double o[3] = {org.x, org.y, org.z};
    const double inv[3] = {1.0 / (double)dir.x, 1.0 / (double)dir.y, 1.0 / (double)dir.z};
    uint32_t stack[128]; int sp = 0;
    stack[sp++] = 0;
    while (sp) {
        uint32_t ni = stack[--sp];
        const BvhNode& n = s.nodes[ni];
        double tn;
        if (!box_hit(n, o, inv, (double)tmax, 4e-6, &tn)) continue;
        if (n.count) {
            for (uint32_t i = 0; i < n.count; i++) {
                const WorldTri& tr = s.tris[s.bvhTriOrder[n.left + i]];
                float t, b1, b2;
                if (rb_tri_intersect(org, sh, tr.v0, tr.v1, tr.v2, &t, &b1, &b2) && t > 0.0f && t < tmax) return true;
            }
        } else {
            stack[sp++] = n.left; stack[sp++] = n.left + 1;
        }
    }
    return false;
}

} // namespace oracle
