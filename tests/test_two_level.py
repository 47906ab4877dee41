"""Two-level instancing (src/scene/Scene.cpp:93-111: one BLAS per object, one TLAS entry per instance,
the driver moves the ray into object space). CPU: the oracle's two-level intersection against its flattened one. GPU: the
library's RB200_FLAG_TWO_LEVEL path against the oracle's two-level path, bit for bit."""
import importlib

import numpy as np
import pytest

import oracle_lib as ol_mod


def _random_rays(n, seed, centre=(0.0, 0.8, 0.0), radius=2.5):
    rng = np.random.RandomState(seed)
    u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    o = (np.asarray(centre) + radius * u).astype(np.float32)
    tgt = np.asarray(centre) + rng.uniform(-0.9, 0.9, (n, 3))
    d = tgt - o
    d = (d / np.linalg.norm(d, axis=1, keepdims=True) * rng.uniform(0.5, 2.0, (n, 1))).astype(np.float32)
    return o, d


def _same_hits(a, b):
    return all(np.array_equal(a[k].view(np.uint32) if a[k].dtype == np.float32 else a[k],
                              b[k].view(np.uint32) if b[k].dtype == np.float32 else b[k]) for k in ("t", "u", "v", "instance", "primitive"))


def test_two_level_equals_flattening_when_every_transform_is_the_identity(ol, rb):
    """With identity transforms object space IS world space: the inverse is the identity, inv * o = o exactly, and both paths
    test the same vertices with the same ray — ids, t and barycentrics must be bit-identical (closest and any hit)."""
    s = rb.Scene()
    I = rb.configs.IDENT
    s.addObject(rb.meshes.cornell_box(), I, rb.Material(**rb.configs.CORNELL_WALL))
    s.addObject(rb.meshes.cornell_light(), I, rb.Material(**rb.configs.LIGHT))
    blob = s.defineObject(rb.meshes.subdivided_blob(levels=3, seed=5, displacement=0.2, radius=0.45))
    s.addInstance(blob, I, rb.Material(materialIdx=0, albedo=(0.5, 0.5, 0.5)))
    s.addInstance(blob, I, rb.Material(materialIdx=1, albedo=(0.5, 0.5, 0.5)))      # coincident instance: ties go to the smaller global id
    tables = s.build(require_emitter=True)
    o, d = _random_rays(20000, 1, centre=(0.0, 0.5, 0.0))
    flat = ol.OracleScene(tables, bvh_threshold=10 ** 9)
    two = ol.OracleScene(tables, bvh_threshold=10 ** 9)
    two.set_two_level(bvh_threshold=10 ** 9)
    for any_hit in (False, True):
        a = flat.trace_rays(o, d, 1e4, any_hit=any_hit, brute=True)
        b = two.trace_rays(o, d, 1e4, any_hit=any_hit, brute=True)
        assert (a["t"] > 0).mean() > 0.5
        if any_hit:
            assert np.array_equal(a["t"] > 0, b["t"] > 0)
        else:
            assert _same_hits(a, b)
    # and with the per-model hierarchies instead of brute force
    two_bvh = ol.OracleScene(tables, bvh_threshold=10 ** 9)
    two_bvh.set_two_level(bvh_threshold=100)
    assert _same_hits(flat.trace_rays(o, d, 1e4, brute=True), two_bvh.trace_rays(o, d, 1e4))
    flat.close(); two.close(); two_bvh.close()


def test_two_level_agrees_with_flattening_on_rotated_and_scaled_instances(ol, rb):
    """General instance transforms: the two paths round differently (world-space vertices against an object-space ray), so
    the comparison is a tolerance — same triangle for all but grazing rays, t to 1e-5 — and an image that is close."""
    wl = rb.configs.instanced(grid=3)
    o, d = _random_rays(40000, 2)
    flat = ol.OracleScene(wl.tables, bvh_threshold=100)
    two = ol.OracleScene(wl.tables, bvh_threshold=100)
    two.set_two_level(bvh_threshold=100)
    a, b = flat.trace_rays(o, d, 1e4), two.trace_rays(o, d, 1e4)
    hit = (a["t"] > 0) & (b["t"] > 0)
    assert ((a["t"] > 0) == (b["t"] > 0)).mean() > 0.9995
    same = hit & (a["instance"] == b["instance"]) & (a["primitive"] == b["primitive"])
    assert same.sum() / hit.sum() > 0.999
    rel = np.abs(a["t"][same] - b["t"][same]) / np.abs(a["t"][same])
    assert np.percentile(rel, 99.9) < 1e-5 and rel.max() < 1e-3, (np.percentile(rel, 99.9), rel.max())
    # the hierarchies of the two-level path change nothing
    two_brute = two.trace_rays(o[:4000], d[:4000], 1e4, brute=True)
    assert _same_hits(two_brute, b[:4000])
    # shadow rays
    sa, sb = flat.trace_rays(o, d, 3.0, any_hit=True), two.trace_rays(o, d, 3.0, any_hit=True)
    assert ((sa["t"] > 0) == (sb["t"] > 0)).mean() > 0.9995
    flat.close(); two.close()


def test_two_level_render_is_close_to_the_flattened_render(ol, rb):
    wl = rb.configs.instanced(width=96, height=72, grid=3, samples_per_pixel=4)
    flat = ol.OracleScene(wl.tables, bvh_threshold=100)
    two = ol.OracleScene(wl.tables, bvh_threshold=100)
    two.set_two_level(bvh_threshold=100)
    pc = wl.push_constants(0)
    a, ca = flat.render_batch(wl.width, wl.height, rb.RB200_FLAG_NEE, pc)
    b, cb = two.render_batch(wl.width, wl.height, rb.RB200_FLAG_NEE, pc)
    # a 1-ulp difference in a hit point sends a path elsewhere, so pixels differ — but most paths are identical and the image
    # means agree closely
    assert (np.abs(a - b).max(axis=2) < 1e-4).mean() > 0.7
    assert abs(float(a[..., :3].mean()) - float(b[..., :3].mean())) < 0.02 * float(a[..., :3].mean())
    assert abs(ca["extendRays"] - cb["extendRays"]) < 0.02 * ca["extendRays"]
    flat.close(); two.close()


def test_singular_instance_transform_is_refused(ol, rb):
    s = rb.Scene()
    s.addObject(rb.meshes.cornell_light(), rb.configs.IDENT, rb.Material(**rb.configs.LIGHT))
    flatM = rb.camera.scale((1.0, 0.0, 1.0))
    s.addObject(rb.meshes.uv_sphere(8, 4), flatM, rb.Material(materialIdx=0, albedo=(0.5, 0.5, 0.5)))
    tables = s.build(require_emitter=True)
    o = ol.OracleScene(tables)
    with pytest.raises(RuntimeError):
        o.set_two_level()
    o.close()


# ---------------------------------------------------------------------------------------------------
# GPU: the library's two-level path (RB200_FLAG_TWO_LEVEL) against the oracle's, through the C ABI
# ---------------------------------------------------------------------------------------------------
def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.gpu
def test_cuda_two_level_hits_equal_the_oracle_two_level_hits(ol, rb):
    """Closest and any hit on 200,000 random rays over 32 rotated, non-uniformly scaled instances of two models: instance,
    primitive, t and barycentrics bit-identical to the oracle's two-level path, which tests EVERY instance without a
    top-level hierarchy (brute force over the instances, hierarchy or brute force inside) — so neither the instance boxes
    nor either hierarchy of the library may change an answer."""
    wl = rb.configs.instanced(grid=4)
    o, d = _random_rays(200000, 3)
    r = rb.Renderer(wl.width, wl.height, wl.tables, flags=rb.RB200_FLAG_NEE | rb.RB200_FLAG_TWO_LEVEL)
    sc = ol.OracleScene(wl.tables, bvh_threshold=100)
    sc.set_two_level(bvh_threshold=100)
    g, c = r.trace_rays(o, d, 1e4), sc.trace_rays(o, d, 1e4)
    assert (g["t"] > 0).mean() > 0.5
    for k in ("instance", "primitive"):
        assert np.array_equal(g[k], c[k]), k
    for k in ("t", "u", "v"):
        assert np.array_equal(_bits(g[k]), _bits(c[k])), k
    brute = sc.trace_rays(o[:5000], d[:5000], 1e4, brute=True)
    assert np.array_equal(_bits(g["t"][:5000]), _bits(brute["t"])) and np.array_equal(g["primitive"][:5000], brute["primitive"])
    ga, ca = r.trace_rays(o, d, 2.5, any_hit=True), sc.trace_rays(o, d, 2.5, any_hit=True)
    assert np.array_equal(ga["t"] > 0, ca["t"] > 0) and 0.05 < (ga["t"] > 0).mean() < 0.95
    # primary rays of the camera too
    pc = wl.push_constants(0)
    gp, cp = r.trace_primary(pc), sc.trace_primary(wl.width, wl.height, pc)
    assert np.array_equal(gp["primitive"], cp["primitive"]) and np.array_equal(_bits(gp["t"]), _bits(cp["t"]))
    r.close(); sc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("nee", [True, False])
def test_cuda_two_level_render_is_bit_identical_to_the_oracle_two_level_render(ol, rb, nee):
    """Whole batches through the wave loop with the two-level traversal kernels: every material on instanced geometry, 3
    samples per pixel, 2 batches; HDR image and ray counters equal the oracle's two-level render bit for bit."""
    wl = rb.configs.instanced(width=128, height=96, grid=3, samples_per_pixel=3, max_bounces=7, nee=nee)
    flags = rb.RB200_FLAG_NEE if nee else 0
    r = rb.Renderer(wl.width, wl.height, wl.tables, flags=flags | rb.RB200_FLAG_TWO_LEVEL)
    sc = ol.OracleScene(wl.tables, bvh_threshold=100)
    sc.set_two_level(bvh_threshold=100)
    hdr_o = np.zeros((wl.height, wl.width, 4), np.float32)
    for b in range(2):
        pc = wl.push_constants(b)
        r.render_batch(pc)
        hdr_o, cnt = sc.render_batch(wl.width, wl.height, flags, pc, hdr_o)
        last, _ = r.stats()
        assert (last["extendRays"], last["shadowRays"], last["paths"]) == (cnt["extendRays"], cnt["shadowRays"], cnt["paths"])
    assert np.array_equal(_bits(r.read_hdr()), _bits(hdr_o))
    r.postprocess()
    assert np.array_equal(r.read_ldr(), ol.postprocess(hdr_o))
    r.close(); sc.close()


@pytest.mark.gpu
def test_two_level_stores_every_object_once_and_agrees_with_the_flattened_path(ol, rb):
    """What the mode is for: 128 instances of two models cost the triangles of two models (+ 64 B per instance), not of 128;
    and its answers are the flattened path's up to the last bits of t (the same triangle for all but grazing rays)."""
    wl = rb.configs.instanced(grid=8, segments=48, rings=24)
    flat = rb.Renderer(wl.width, wl.height, wl.tables, flags=rb.RB200_FLAG_NEE)
    two = rb.Renderer(wl.width, wl.height, wl.tables, flags=rb.RB200_FLAG_NEE | rb.RB200_FLAG_TWO_LEVEL)
    bf, bt = flat.bvh_info(), two.bvh_info()
    assert bf["numTriangles"] == wl.tables.num_triangles()
    assert bt["numTriangles"] < bf["numTriangles"] // 30
    assert bt["triangleBytes"] + bt["nodeBytes"] < (bf["triangleBytes"] + bf["nodeBytes"]) // 20
    o, d = _random_rays(100000, 4)
    a, b = flat.trace_rays(o, d, 1e4), two.trace_rays(o, d, 1e4)
    hit = (a["t"] > 0) & (b["t"] > 0)
    assert ((a["t"] > 0) == (b["t"] > 0)).mean() > 0.9995
    same = hit & (a["instance"] == b["instance"]) & (a["primitive"] == b["primitive"])
    assert same.sum() / hit.sum() > 0.999
    rel = np.abs(a["t"][same] - b["t"][same]) / np.abs(a["t"][same])
    assert np.percentile(rel, 99.9) < 1e-5
    # the builds are deterministic in this mode as well
    two2 = rb.Renderer(wl.width, wl.height, wl.tables, flags=rb.RB200_FLAG_NEE | rb.RB200_FLAG_TWO_LEVEL)
    assert two2.bvh_info()["hash"] == bt["hash"] != bf["hash"]
    flat.close(); two.close(); two2.close()


@pytest.mark.gpu
def test_two_level_with_identity_transforms_is_the_flattened_path_bit_for_bit(ol, rb):
    s = rb.Scene()
    I = rb.configs.IDENT
    s.addObject(rb.meshes.cornell_box(), I, rb.Material(**rb.configs.CORNELL_WALL))
    s.addObject(rb.meshes.cornell_light(), I, rb.Material(**rb.configs.LIGHT))
    blob = s.defineObject(rb.meshes.subdivided_blob(levels=3, seed=5, displacement=0.2, radius=0.45))
    s.addInstance(blob, I, rb.Material(materialIdx=0, albedo=(0.5, 0.5, 0.5)))
    s.addInstance(blob, I, rb.Material(materialIdx=1, albedo=(0.5, 0.5, 0.5)))      # coincident: ties go to the smaller global id
    tables = s.build(require_emitter=True)
    o, d = _random_rays(50000, 5, centre=(0.0, 0.5, 0.0))
    flat = rb.Renderer(64, 48, tables, flags=0)
    two = rb.Renderer(64, 48, tables, flags=rb.RB200_FLAG_TWO_LEVEL)
    a, b = flat.trace_rays(o, d, 1e4), two.trace_rays(o, d, 1e4)
    for k in ("instance", "primitive"):
        assert np.array_equal(a[k], b[k]), k
    for k in ("t", "u", "v"):
        assert np.array_equal(_bits(a[k]), _bits(b[k])), k
    flat.close(); two.close()


@pytest.mark.gpu
def test_two_level_refuses_a_singular_instance_transform(rb):
    s = rb.Scene()
    s.addObject(rb.meshes.cornell_light(), rb.configs.IDENT, rb.Material(**rb.configs.LIGHT))
    s.addObject(rb.meshes.uv_sphere(8, 4), rb.camera.scale((1.0, 0.0, 1.0)), rb.Material(materialIdx=0, albedo=(0.5, 0.5, 0.5)))
    tables = s.build(require_emitter=True)
    with pytest.raises(Exception, match="singular"):
        rb.Renderer(32, 32, tables, flags=rb.RB200_FLAG_TWO_LEVEL)
