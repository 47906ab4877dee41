"""Known-answer tests pinning the oracle (and the shared elementary layer) to the values derived from the reference's
GLSL in SURVEY.md Appendix A3. The reference ships no tests or golden vectors (parity unpinned), so these KATs — plus
independent numpy restatements of the integer-exact pieces below — are what anchors the restatement."""
import ctypes as C

import numpy as np
import pytest


def np_random(state):
    """shaders/raytrace/shaderCommon.h.glsl:39-45 in numpy (third, independent statement)."""
    state = (np.uint64(state) * np.uint64(747796405) + np.uint64(1)) & np.uint64(0xFFFFFFFF)
    s = int(state)
    word = (((s >> ((s >> 28) + 4)) ^ s) * 277803737) & 0xFFFFFFFF
    word = ((word >> 22) ^ word) & 0xFFFFFFFF
    return s, word, np.float32(np.float32(np.uint32(word)) / np.float32(4294967295.0))


def oracle_seq(ol, state, n):
    st = C.c_uint32(state)
    return [np.float32(ol.lib().oracle_kat_random(C.byref(st))) for _ in range(n)], st.value


def test_rng_known_answers(ol):
    seq, _ = oracle_seq(ol, 0, 4)
    assert np.allclose(seq, [0.06468121, 4.9004331e-05, 0.86248976, 0.8270753], rtol=2e-7, atol=0)
    s, w, _ = np_random(0)
    assert (s, w) == (0x00000001, 0x108EF29B)
    s, w, _ = np_random(s)
    assert (s, w) == (0x2C9277B6, 0x00033628)
    for seed, expect in ((479999, [0.5611887, 0.60613185, 0.7933254, 0.37940344]),
                         (2073599, [0.6102007, 0.23335038, 0.50339574, 0.71383315]),
                         (0xFFFFFFFF, [0.064191855, 0.93543744, 0.32334736, 0.6997081])):
        seq, _ = oracle_seq(ol, seed, 4)
        assert np.allclose(seq, expect, rtol=2e-7, atol=0), (seed, seq)


def test_rng_matches_numpy_and_shared_layer(ol):
    rng = np.random.RandomState(0)
    for seed in [0, 1, 485605, 0xFFFFFFFF] + [int(x) for x in rng.randint(0, 2 ** 32, 50, dtype=np.uint64)]:
        seq, end = oracle_seq(ol, seed, 16)
        s = seed
        for v in seq:
            s, _, ref = np_random(s)
            assert v == ref
        assert end == s
        st = C.c_uint32(seed)
        out = np.empty(16, np.float32)
        ol.lib().oracle_rb_random(C.byref(st), 16, out.ctypes.data_as(C.c_void_p))    # rb_random (kernel side)
        assert (out == np.array(seq, np.float32)).all() and st.value == end


def test_random_can_return_one():
    # float(w) rounds w >= 2^32 - 128 up to 2^32 and 4294967295.0f is 2^32: random() lies in [0, 1] inclusive
    assert np.float32(np.uint32(0xFFFFFF80)) / np.float32(4294967295.0) == np.float32(1.0)
    assert np.float32(np.uint32(0xFFFFFF7F)) / np.float32(4294967295.0) == np.float32(0.99999994)


def test_seed_formula_and_lens_draws_not_consumed(ol, rb):
    wl = rb.configs.cornell(800, 600)
    pc = wl.push_constants(1)
    o, d = np.zeros(3, np.float32), np.zeros(3, np.float32)
    after = C.c_uint32(0)
    ol.lib().oracle_kat_starting_ray(C.byref(pc), 5, 7, 800, 600, o.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p),
                                     C.byref(after))
    s = (1 * 600 + 7) * 800 + 5
    assert s == 485605                                   # raytrace.rgen.glsl:259
    for _ in range(2):                                   # only the two Gaussian draws advance the state (:186,194,234)
        s, _, _ = np_random(s)
    assert after.value == s
    assert abs(np.linalg.norm(d) - 1) < 1e-6 and abs(o[2] - 3.9) < 0.05


def test_sky_and_power_heuristic(ol):
    out = np.zeros(3, np.float32)
    for y, expect in ((1.0, (0.028, 0.119, 0.14)), (0.0, (0.0175, 0.063, 0.0735)), (-1.0, (0.007, 0.007, 0.007))):
        d = np.array([np.sqrt(max(0.0, 1 - y * y)), y, 0], np.float32)
        ol.lib().oracle_kat_sky(d.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        assert np.allclose(out, expect, rtol=1e-5)
    assert abs(ol.lib().oracle_kat_power_heuristic(0.5, 0.25) - 0.8) < 1e-7        # pdf.h.glsl:4-7


def test_tonemap_known_answers(ol):
    def tm(rgb, exposure=1.0):
        a = np.array(rgb, np.float32)
        o = np.zeros(4, np.uint8)
        ol.lib().oracle_kat_tonemap(a.ctypes.data_as(C.c_void_p), exposure, o.ctypes.data_as(C.c_void_p))
        return tuple(int(x) for x in o)
    assert tm((0.18, 0.18, 0.18)) == (68, 68, 68, 255)
    assert tm((0.5, 0.5, 0.5)) == (158, 158, 158, 255)
    assert tm((1.0, 1.0, 1.0)) == (205, 205, 205, 255)
    assert tm((4.0, 0.5, 0.1)) == (255, 184, 94, 255)
    assert tm((0.0, 0.0, 0.0)) == (0, 0, 0, 255)
    assert tm((16.0, 16.0, 16.0)) == (255, 255, 255, 255)


def np_offset(p, n):
    """closestHitCommon.h.glsl:156-177 in numpy."""
    p = np.asarray(p, np.float32); n = np.asarray(n, np.float32)
    of_i = (np.float32(256.0) * n).astype(np.int32)          # truncation toward zero
    bits = p.view(np.int32) + np.where(p < 0, -of_i, of_i)
    p_i = bits.astype(np.int32).view(np.float32)
    return np.where(np.abs(p) < np.float32(1.0 / 32.0), p + np.float32(1.0 / 65536.0) * n, p_i).astype(np.float32)


def test_offset_along_normal_bit_exact(ol):
    rng = np.random.RandomState(3)
    pts = np.concatenate([rng.uniform(-3, 3, (200, 3)), rng.uniform(-0.05, 0.05, (200, 3)), rng.uniform(-1e4, 1e4, (50, 3))])
    for p in pts.astype(np.float32):
        n = rng.normal(size=3); n = (n / np.linalg.norm(n)).astype(np.float32)
        out = np.zeros(3, np.float32)
        ol.lib().oracle_kat_offset(p.ctypes.data_as(C.c_void_p), n.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        assert (out.view(np.uint32) == np_offset(p, n).view(np.uint32)).all(), (p, n)


@pytest.mark.parametrize("dim,k", [(800, 120), (600, 90), (1920, 288), (1080, 162), (3840, 576), (2160, 324)])
def test_bloom_half_width(dim, k):
    # blurCommon.h.glsl:23-24 with radius = 5 (percent)
    assert int(np.float32(np.float32(dim) * np.float32(5.0) / np.float32(100.0)) * np.float32(3) + np.float32(0.5)) == k


def test_bloom_impulse_response(ol, rb):
    """sigma is the percent value (5), weightSum covers every tap (= sigma*sqrt(2 pi) = 12.5331 once k >= ~30)."""
    W, H = 200, 160
    hdr = np.zeros((H, W, 4), np.float32); hdr[..., 3] = 1
    V = 50.0
    hdr[80, 100, :3] = V
    bloom = rb.abi.BloomPushConsts(5.0, 1.0, 0.05)
    _, comb = ol.postprocess(hdr, bloom=bloom, want_combined=True)
    S = 12.5331
    for dx, dy in ((0, 0), (10, 0), (0, 10), (7, -4), (20, 0)):
        expect = V * np.exp(-dx * dx / 50.0) / S * np.exp(-dy * dy / 50.0) / S * 0.05 + (V if (dx, dy) == (0, 0) else 0)
        assert abs(comb[80 + dy, 100 + dx, 0] - expect) < 2e-4 * max(1.0, expect), (dx, dy)
    # below-threshold pixels do not bloom but still count in weightSum
    hdr2 = hdr.copy(); hdr2[80, 100, :3] = 0.5
    _, comb2 = ol.postprocess(hdr2, bloom=bloom, want_combined=True)
    assert comb2[80, 110, 0] == 0.0


# ---- parallax bump mapping (texutils.h.glsl:4-41), numpy float32 restatement ------------------------------------
F = np.float32


def np_texture_r(img, u, v):
    """Bilinear, REPEAT, UNORM8, LOD 0: red channel."""
    h, w = img.shape[:2]
    x = F(F(u) * F(w)) - F(0.5); y = F(F(v) * F(h)) - F(0.5)
    fx, fy = np.floor(x), np.floor(y)
    ax, ay = F(x - fx), F(y - fy)
    x0, y0 = int(fx) % w, int(fy) % h
    x1, y1 = (x0 + 1) % w, (y0 + 1) % h
    c = lambda yy, xx: F(F(img[yy, xx, 0]) / F(255.0))
    top = F(F(c(y0, x0) * F(F(1) - ax)) + F(c(y0, x1) * ax))
    bot = F(F(c(y1, x0) * F(F(1) - ax)) + F(c(y1, x1) * ax))
    return F(F(top * F(F(1) - ay)) + F(bot * ay))


def np_bump(img, uv, ray_in, T):
    hs = F(0.2)
    m = -np.asarray(ray_in, np.float32)
    V = np.array([F(F(F(T[c][0] * m[0]) + F(T[c][1] * m[1])) + F(T[c][2] * m[2])) for c in range(3)], np.float32)
    inv = F(F(1) / np.sqrt(F(F(F(V[0] * V[0]) + F(V[1] * V[1])) + F(V[2] * V[2]))))
    V = (V * inv).astype(np.float32)
    if V[2] <= 0:
        return F(uv[0]), F(uv[1]), 0
    a = min(max(V[2], F(0)), F(1))
    layers = F(F(F(512) * F(F(1) - a)) + F(F(64) * a))
    depth = F(F(1) / layers)
    du = F(F(F(V[0] / V[2]) * hs) / layers); dv = F(F(F(V[1] / V[2]) * hs) / layers)
    cu, cv, s, steps = F(uv[0]), F(uv[1]), F(0), 0
    h = F(np_texture_r(img, cu, cv) * hs)
    while s < h:
        cu = F(cu - du); cv = F(cv - dv); s = F(s + depth); steps += 1
        h = F(np_texture_r(img, cu, cv) * hs)
    pu, pv = F(cu + du), F(cv + dv)
    hp = F(np_texture_r(img, pu, pv) * hs)
    after = F(h - s); before = F(hp - F(s - depth))
    wgt = F(after / F(after - before))
    mix = lambda p, c: F(F(p * F(F(1) - wgt)) + F(c * wgt))
    return mix(pu, cu), mix(pv, cv), steps


def test_parallax_bump_mapping_matches_numpy(ol, rb):
    img = rb.configs.brick_height_map(64)
    rng = np.random.RandomState(11)
    total_steps = 0
    for k in range(300):
        # an orthonormal tangent frame, columns T B N, and a ray arriving from the N side (or not: V.z <= 0 returns uv)
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        T = [q[:, 0].astype(np.float32), q[:, 1].astype(np.float32), q[:, 2].astype(np.float32)]
        graze = rng.uniform(0.02, 1.0)
        d = -(graze * q[:, 2] + rng.normal(size=3) * 0.7)
        if k % 10 == 0:
            d = -d
        d = (d / np.linalg.norm(d)).astype(np.float32)
        uv = rng.uniform(0, 1, 2).astype(np.float32)
        tbn = np.concatenate(T).astype(np.float32)
        out = np.zeros(2, np.float32)
        ol.lib().oracle_kat_bump(img.ctypes.data_as(C.c_void_p), 64, 64, uv.ctypes.data_as(C.c_void_p),
                                 d.ctypes.data_as(C.c_void_p), tbn.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        eu, ev, steps = np_bump(img, uv, d, T)
        total_steps += steps
        assert out[0].view(np.uint32) == np.float32(eu).view(np.uint32) and out[1].view(np.uint32) == np.float32(ev).view(np.uint32), (k, out, eu, ev)
    assert total_steps > 1000        # the search really walked layers


def test_parallax_flat_height_map_is_identity_at_zero_height(ol):
    # h = 0 everywhere: the loop never runs, weight = -0/(0 - layerDepth) = -0... and mix returns prevUV = uv + deltaUV
    img = np.zeros((8, 8, 4), np.uint8)
    tbn = np.eye(3, dtype=np.float32).reshape(-1)
    uv = np.array([0.25, 0.5], np.float32)
    d = np.array([0.0, 0.0, -1.0], np.float32)             # head-on: V = (0, 0, 1), deltaUV = 0
    out = np.zeros(2, np.float32)
    ol.lib().oracle_kat_bump(img.ctypes.data_as(C.c_void_p), 8, 8, uv.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p),
                             tbn.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert (out == uv).all()
    # white height map, 45 degrees in the tangent x-z plane: P = V.xy / V.z * 0.2 spans the full depth range [0, 1] and
    # the height is scaled by 0.2 as well, so the ray meets the surface after 0.2 of P: uv.x moves by -0.2 * 0.2
    img[...] = 255
    d = np.array([-1.0, 0.0, -1.0], np.float32) / np.float32(np.sqrt(2))
    ol.lib().oracle_kat_bump(img.ctypes.data_as(C.c_void_p), 8, 8, uv.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p),
                             tbn.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert abs(out[0] - (0.25 - 0.2 * 0.2)) < 2e-3 and out[1] == np.float32(0.5)


# ---------------------------------------------------------------------------------------------------------------------
# next-event estimation: compiled OUT of the binaries upstream ships, so it cannot be pinned by
# executing them; nee.h.glsl and directLight are restated here in numpy float32, independently of oracle/pathtrace.cpp
# ---------------------------------------------------------------------------------------------------------------------
F = np.float32


def _np_random(state):
    """shaderCommon.h.glsl random(): PCG step + output permutation, float(w) / 2^32 (rounds up to 1.0 for large w)."""
    s = (state * 747796405 + 1) & 0xFFFFFFFF
    w = (((s >> ((s >> 28) + 4)) ^ s) * 277803737) & 0xFFFFFFFF
    w = ((w >> 22) ^ w) & 0xFFFFFFFF
    return s, F(F(w) * F(2.3283064365386963e-10))


def _np_dot(a, b):
    return F(F(F(a[0] * b[0]) + F(a[1] * b[1])) + F(a[2] * b[2]))


def _np_normalize(v):
    inv = F(F(1) / F(np.sqrt(_np_dot(v, v))))
    return np.array([F(c * inv) for c in v], F)


def _np_random_emissive_point(t, rb, pc, state):
    """nee.h.glsl:52-124: instance by CDF, triangle by CDF (lower bound, u <= cdf[mid]), point by the square-root
    parametrisation, pdf = (weight / totalEmissiveWeight) * (1 / area)."""
    def lower_bound(cdf, lo, hi, u):
        while lo < hi:
            mid = (lo + hi) // 2
            if u <= cdf[mid]:
                hi = mid
            else:
                lo = mid + 1
        return lo
    md_all = np.frombuffer(t.emissive.tobytes()[: t.numEmissive * 112], np.uint8).reshape(-1, 112)
    state, u = _np_random(state)
    inst = lower_bound(t.cdfInstances, 0, t.numEmissive - 1, u)
    md = rb.abi.InstanceData.from_buffer_copy(md_all[inst].tobytes())
    state, u = _np_random(state)
    tri = lower_bound(t.cdfTriangles, md.cdfRangeStart, md.cdfRangeEnd, u)
    M = np.array(md.transform[:], F).reshape(4, 4)          # column-major: M[c][r]
    vs = []
    for k in range(3):
        v = t.vertices.reshape(-1, 4)[t.indices[3 * tri + md.indexOffset + k]][:3]
        # mat4 * vec4(v, 1): x*c0 + y*c1 + z*c2 + c3, summed left to right
        vs.append(np.array([F(F(F(F(M[0][r] * v[0]) + F(M[1][r] * v[1])) + F(M[2][r] * v[2])) + M[3][r]) for r in range(3)], F))
    state, r1 = _np_random(state)
    beta = F(F(1) - F(np.sqrt(r1)))
    state, r2 = _np_random(state)
    gamma = F(F(F(1) - beta) * r2)
    alpha = F(F(F(1) - beta) - gamma)
    point = np.array([F(F(F(alpha * vs[0][c]) + F(beta * vs[1][c])) + F(gamma * vs[2][c])) for c in range(3)], F)
    e1, e2 = (vs[1] - vs[0]).astype(F), (vs[2] - vs[0]).astype(F)
    cross = np.array([F(F(e1[1] * e2[2]) - F(e1[2] * e2[1])), F(F(e1[2] * e2[0]) - F(e1[0] * e2[2])), F(F(e1[0] * e2[1]) - F(e1[1] * e2[0]))], F)
    pdf = F(F(F(md.weight) / F(pc.totalEmissiveWeight)) * F(F(1) / F(md.area)))
    return state, point, _np_normalize(cross), np.array(md.emission[:], F), pdf, bool(md.cullBackface)


def _two_light_scene(rb, emitters_first=True):
    """Two emitters (rb.configs.two_lights): see there for how nee.h.glsl:97-105 addresses their triangles."""
    return rb.configs.two_lights(8, 8, emitters_first=emitters_first).tables


def test_light_sampling_past_the_index_buffer_is_refused(ol, rb):
    t = _two_light_scene(rb, emitters_first=False)
    pc = rb.camera.push_constants(8, 8, (0, 1, 3.9), (0, 1, 0), 40.0, total_emissive_weight=t.totalEmissiveWeight)
    sc = ol.OracleScene(t)
    hdr = np.zeros((8, 8, 4), np.float32)
    args = (sc._h, 8, 8, C.byref(pc), hdr.ctypes.data_as(C.c_void_p), 1, None, 0, 1, 32)
    assert ol.lib().oracle_render_batch_tiles(args[0], 8, 8, rb.RB200_FLAG_NEE, *args[3:]) == -1       # RB200_ERR_INVALID_ARGUMENT
    assert ol.lib().oracle_render_batch_tiles(args[0], 8, 8, 0, *args[3:]) == 0                          # without NEE nothing samples lights


def test_light_sampling_matches_a_numpy_restatement(ol, rb):
    t = _two_light_scene(rb)
    assert t.numEmissive == 2
    pc = rb.camera.push_constants(8, 8, (0, 1, 3.9), (0, 1, 0), 40.0, total_emissive_weight=t.totalEmissiveWeight)
    sc = ol.OracleScene(t)
    picked = set()
    for state0 in [0, 1, 7, 12345, 0xFFFFFFFF, 0x80000000] + [int(x) for x in np.random.RandomState(2).randint(0, 2 ** 32, 300, dtype=np.uint64)]:
        st = C.c_uint32(state0)
        out = np.zeros(11, np.float32)
        ol.lib().oracle_kat_light_sample(sc._h, C.byref(pc), C.byref(st), out.ctypes.data_as(C.c_void_p))
        state, point, normal, emission, pdf, cull = _np_random_emissive_point(t, rb, pc, state0)
        want = np.concatenate([point, normal, emission, [pdf, F(cull)]]).astype(np.float32)
        assert st.value == state and (out.view(np.uint32) == want.view(np.uint32)).all(), hex(state0)
        picked.add(tuple(emission))
    assert len(picked) == 2                                  # both emitters are drawn


def test_direct_light_lambertian_matches_a_numpy_restatement(ol, rb):
    """raytrace.rgen.glsl:43-95 for materialID 0: solid-angle pdf, BRDF = albedo / k_pi (a reciprocal multiply in the
    reference build), cosine and geometry terms with the emitter's cull rule, occlusion from the oracle's own
    any-hit query (the traversal is tested on its own)."""
    t = _two_light_scene(rb)
    pc = rb.camera.push_constants(8, 8, (0, 1, 3.9), (0, 1, 0), 40.0, total_emissive_weight=t.totalEmissiveWeight)
    sc = ol.OracleScene(t)
    rng = np.random.RandomState(9)
    lit = dark = 0
    for _ in range(300):
        o = np.array([rng.uniform(-0.9, 0.9), rng.uniform(0.05, 1.9), rng.uniform(-0.9, 0.9)], F)
        n = _np_normalize(rng.normal(size=3).astype(F))
        albedo = rng.uniform(0, 1, 3).astype(F)
        state0 = int(rng.randint(0, 2 ** 32, dtype=np.uint64))
        st = C.c_uint32(state0)
        out = np.zeros(4, np.float32)
        ol.lib().oracle_kat_direct_light_lambertian(sc._h, C.byref(pc), o.ctypes.data_as(C.c_void_p), n.ctypes.data_as(C.c_void_p),
                                                    albedo.ctypes.data_as(C.c_void_p), C.byref(st), out.ctypes.data_as(C.c_void_p))
        state, point, ln, emission, lpdf, cull = _np_random_emissive_point(t, rb, pc, state0)
        to = (point - o).astype(F)
        d = _np_normalize(to)
        dist = F(np.sqrt(_np_dot(to, to)))
        pdf = F(F(F(lpdf * dist) * dist) / max(_np_dot(ln, -d), F(0.0001)))
        occluded = sc.trace_rays(o[None], d[None], np.array([dist - F(0.001)], F), any_hit=True, threads=1)["t"][0] >= 0
        if occluded:
            want = np.array([0, 0, 0, pdf], F)
            dark += 1
        else:
            brdf = (albedo * F(0.31830987334251404)).astype(F)
            cos_i = _np_dot(n, d)
            cos_i = max(cos_i, F(0)) if cull else abs(cos_i)
            g = _np_dot(ln, -d)
            g = max(g, F(0)) if cull else abs(g)
            geom = F(g / F(dist * dist))
            rgb = [F(F(F(F(emission[c] * brdf[c]) * cos_i) * geom) / lpdf) for c in range(3)]
            want = np.array(rgb + [pdf], F)
            lit += 1
        assert st.value == state and (out.view(np.uint32) == want.view(np.uint32)).all(), (o, n, state0, out, want)
    assert lit > 50 and dark > 20


def test_bounce_loop_with_nee_matches_a_python_restatement(ol, rb):
    """raytrace.rgen.glsl:97-184 with next-event estimation switched ON (upstream ships it compiled out, so the binaries
    cannot pin this part): the skip / sky / inside-dielectric branches, skipNEE, the MIS weight cases (first bounce,
    previous skip, left a dielectric, last bounce, general), the order of the radiance and throughput updates, the clamp
    and the running average — restated here over three building blocks that are pinned elsewhere: one traceRayEXT (the
    compiled closest-hit / miss shaders, test_spirv_golden.py), directLight (numpy restatement above) and the starting
    ray. Whole images of the Lambertian Cornell box must agree bit for bit over three batches."""
    W, H, BOUNCES = 12, 9, 5
    wl = rb.configs.cornell(W, H, nee=True, samples_per_pixel=1, max_bounces=BOUNCES)
    sc = ol.OracleScene(wl.tables)
    L = ol.lib()
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)

    def heuristic(a, b):
        return F(F(a * a) / F(F(a * a) + F(b * b)))

    def trace_segments(pc, org, d, state):
        inside, acc = False, F(0)
        T, rad = np.ones(3, F), np.zeros(3, F)
        first, prev_skip = True, False
        for seg in range(BOUNCES):
            prev_inside = inside
            st, out, flags = C.c_uint32(state), np.zeros(21, F), C.c_uint32()
            L.oracle_kat_trace_main(sc._h, ptr(org), ptr(d), C.byref(st), int(inside), float(acc), ptr(out), C.byref(flags))
            state = st.value
            color, albedo, org, d, emission, normal = (out[3 * k:3 * k + 3].copy() for k in range(6))
            pdf_brdf, acc = out[18], out[19]
            hit_sky, skip, inside, material = bool(flags.value & 1), bool(flags.value & 2), bool(flags.value & 4), flags.value >> 8
            left = (not inside) and prev_inside
            if skip:
                continue
            if hit_sky:
                rad = (rad + color * T).astype(F)
                break
            if not inside:
                skip_nee = material not in (0, 3)
                direct = np.zeros(4, F)
                if not skip_nee:
                    assert material == 0                     # this scene is Lambertian throughout
                    st = C.c_uint32(state)
                    L.oracle_kat_direct_light_lambertian(sc._h, C.byref(pc), ptr(org), ptr(normal), ptr(albedo), C.byref(st), ptr(direct))
                    state = st.value
                w_nee, w_brdf = F(0), F(1)
                if not skip_nee:
                    if first or prev_skip or left:
                        w_nee, w_brdf = F(1), F(1)
                    elif seg + 1 == BOUNCES:
                        w_nee, w_brdf = F(0), heuristic(pdf_brdf, direct[3])
                    else:
                        w_nee, w_brdf = heuristic(direct[3], pdf_brdf), heuristic(pdf_brdf, direct[3])
                prev_skip = skip_nee
                combined = ((direct[:3] * w_nee).astype(F) + (emission * w_brdf).astype(F)).astype(F)
                rad = (rad + (combined * T).astype(F)).astype(F)
                T = (T * color).astype(F)
            first = False
        return rad

    img = np.zeros((H, W, 4), F)
    want = np.zeros((H, W, 4), F)
    seen = {"clamped": 0, "lit": 0}
    for batch in range(3):
        pc = wl.push_constants(batch, direct_clamp=2.0)
        sc.render_batch(W, H, rb.RB200_FLAG_NEE, pc, want, threads=1)
        for y in range(H):
            for x in range(W):
                o, d, st = np.zeros(3, F), np.zeros(3, F), C.c_uint32()
                L.oracle_kat_starting_ray(C.byref(pc), x, y, W, H, ptr(o), ptr(d), C.byref(st))
                c = trace_segments(pc, o, d, st.value)
                seen["clamped"] += int((c > 2.0).any())
                seen["lit"] += int((c > 0).any())
                c = np.minimum(np.maximum(c, F(0)), F(2.0)).astype(F)
                assert not np.isnan(c).any()
                if batch:
                    c = (((img[y, x, :3] * F(batch)).astype(F) + c).astype(F) / F(batch + 1)).astype(F)
                img[y, x] = (c[0], c[1], c[2], 1)
        assert (img.view(np.uint32) == want.view(np.uint32)).all(), batch
    assert seen["lit"] > 200 and seen["clamped"] > 0


def test_bounce_loop_with_nee_on_every_material(ol, rb):
    """The same restatement of raytrace.rgen.glsl:97-184 on a scene with all four materials, textures and skips, so that
    the branches the Lambertian box never takes are compared too: skipNEE for metal and glass (prevSkip makes the next
    diffuse hit count its light sample fully), paths inside glass (no light sample, no throughput update), leftDielectric,
    cull / alpha skips that do not consume a bounce's bookkeeping. One bounce = oracle_kat_bounce (the closest-hit shader
    and, for Lambertian / Disney hits, the directLight that would follow, on a copy of the RNG state); WHETHER it follows,
    its weight and everything after it are decided here."""
    W, H, BOUNCES = 16, 12, 8
    wl = rb.configs.small_mixed(W, H, nee=True, samples_per_pixel=1, max_bounces=BOUNCES)
    sc = ol.OracleScene(wl.tables)
    L = ol.lib()
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    heuristic = lambda a, b: F(F(a * a) / F(F(a * a) + F(b * b)))
    taken = {"skip": 0, "prev_skip_full_weight": 0, "left_dielectric": 0, "inside": 0, "last_bounce": 0, "general": 0, "disney_nee": 0}

    def trace_segments(pc, org, d, state):
        inside, acc = False, F(0)
        T, rad = np.ones(3, F), np.zeros(3, F)
        first, prev_skip = True, False
        for seg in range(BOUNCES):
            prev_inside = inside
            st, out, flags, direct, st_direct = C.c_uint32(state), np.zeros(21, F), C.c_uint32(), np.zeros(4, F), C.c_uint32()
            L.oracle_kat_bounce(sc._h, C.byref(pc), ptr(org), ptr(d), C.byref(st), int(inside), float(acc), ptr(out), C.byref(flags),
                                ptr(direct), C.byref(st_direct))
            state = st.value
            color, albedo, org, d, emission, normal = (out[3 * k:3 * k + 3].copy() for k in range(6))
            pdf_brdf, acc = out[18], out[19]
            hit_sky, skip, inside, material = bool(flags.value & 1), bool(flags.value & 2), bool(flags.value & 4), flags.value >> 8
            left = (not inside) and prev_inside
            if skip:
                taken["skip"] += 1
                continue
            if hit_sky:
                rad = (rad + color * T).astype(F)
                break
            if inside:
                taken["inside"] += 1
            else:
                skip_nee = material not in (0, 3)
                w_nee, w_brdf = F(0), F(1)
                if skip_nee:
                    direct = np.zeros(4, F)
                else:
                    state = st_direct.value                   # directLight ran: its four draws are consumed
                    taken["disney_nee"] += int(material == 3)
                    if first or prev_skip or left:
                        w_nee, w_brdf = F(1), F(1)
                        taken["prev_skip_full_weight"] += int(prev_skip and not first)
                        taken["left_dielectric"] += int(left)
                    elif seg + 1 == BOUNCES:
                        w_nee, w_brdf = F(0), heuristic(pdf_brdf, direct[3])
                        taken["last_bounce"] += 1
                    else:
                        w_nee, w_brdf = heuristic(direct[3], pdf_brdf), heuristic(pdf_brdf, direct[3])
                        taken["general"] += 1
                prev_skip = skip_nee
                combined = ((direct[:3] * w_nee).astype(F) + (emission * w_brdf).astype(F)).astype(F)
                rad = (rad + (combined * T).astype(F)).astype(F)
                T = (T * color).astype(F)
            first = False
        return rad

    img, want = np.zeros((H, W, 4), F), np.zeros((H, W, 4), F)
    for batch in range(2):
        pc = wl.push_constants(batch)
        clamp = F(pc.directClamp)
        sc.render_batch(W, H, rb.RB200_FLAG_NEE, pc, want, threads=1)
        for y in range(H):
            for x in range(W):
                o, d, st = np.zeros(3, F), np.zeros(3, F), C.c_uint32()
                L.oracle_kat_starting_ray(C.byref(pc), x, y, W, H, ptr(o), ptr(d), C.byref(st))
                c = trace_segments(pc, o, d, st.value)
                if np.isnan(c).any():                     # a NaN sample is dropped (rgen.glsl:268-272); with 1 spp the pixel keeps its value
                    if batch == 0:
                        img[y, x] = (0, 0, 0, 1)
                    continue
                c = np.minimum(np.maximum(c, F(0)), clamp).astype(F)
                if batch:
                    c = (((img[y, x, :3] * F(batch)).astype(F) + c).astype(F) / F(batch + 1)).astype(F)
                img[y, x] = (c[0], c[1], c[2], 1)
        assert (img.view(np.uint32) == want.view(np.uint32)).all(), batch
    assert all(v > 0 for v in taken.values()), taken
