"""Self-consistency of the CPU oracle: it is the parity checker for the CUDA path, so its own invariants are tested
here on the CPU (no GPU): BVH vs brute-force intersection, thread-count independence, accumulation modes, light
sampling, energy sanity and the committed golden fixtures."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_bvh_equals_brute_force(ol, rb):
    wl = rb.configs.dragon(64, 48, n_along=300, n_ring=10)
    fast = ol.OracleScene(wl.tables, bvh_threshold=0)       # always BVH
    rng = np.random.RandomState(5)
    n = 4000
    o = rng.uniform(-0.9, 0.9, (n, 3)).astype(np.float32) + np.array([0, 1, 0], np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    a = fast.trace_rays(o, d, 1e4)
    b = fast.trace_rays(o, d, 1e4, brute=True)
    assert (bits(a["t"]) == bits(b["t"])).all() and (a["primitive"] == b["primitive"]).all() and (a["instance"] == b["instance"]).all()
    tm = rng.uniform(0.05, 2.0, n).astype(np.float32)
    assert ((fast.trace_rays(o, d, tm, any_hit=True)["t"] >= 0) == (fast.trace_rays(o, d, tm, any_hit=True, brute=True)["t"] >= 0)).all()
    # axis-parallel rays and rays starting exactly on geometry
    o2 = np.array([[0, 1, 0], [0, 0, 0], [0.3, 1.0, -0.2], [0, 1.989, 0]], np.float32)
    d2 = np.array([[0, -1, 0], [0, 1, 0], [1, 0, 0], [0, 0, 1]], np.float32)
    a, b = fast.trace_rays(o2, d2, 1e4), fast.trace_rays(o2, d2, 1e4, brute=True)
    assert (bits(a["t"]) == bits(b["t"])).all() and (a["primitive"] == b["primitive"]).all()


def test_closest_hit_tie_rule_picks_smallest_primitive(ol, rb):
    """Two coincident quads: equal t, the smaller global primitive id wins regardless of traversal order."""
    s = rb.Scene()
    q = rb.meshes.quad((-1, 0, -1), (1, 0, -1), (1, 0, 1), (-1, 0, 1))
    s.addObject(q, np.eye(4, dtype=np.float32), rb.Material(emission=(1, 1, 1)))
    s.addObject(q, np.eye(4, dtype=np.float32), rb.Material())
    sc = ol.OracleScene(s.build())
    h = sc.trace_rays(np.array([[0.2, 1, 0.3]], np.float32), np.array([[0, -1, 0]], np.float32), 1e4)
    assert h["instance"][0] == 0 and abs(h["t"][0] - 1.0) < 1e-6


def test_thread_count_does_not_change_the_image(ol, rb):
    wl = rb.configs.small_mixed(48, 36, samples_per_pixel=2, max_bounces=5)
    sc = ol.OracleScene(wl.tables)
    pc = wl.push_constants(0)
    a, ca = sc.render_batch(48, 36, rb.RB200_FLAG_NEE, pc, threads=1)
    b, cb = sc.render_batch(48, 36, rb.RB200_FLAG_NEE, pc, threads=7)
    assert (bits(a) == bits(b)).all() and ca == cb


def test_running_mean_vs_sum_mode(ol, rb):
    wl = rb.configs.cornell(40, 30, samples_per_pixel=1, max_bounces=4)
    sc = ol.OracleScene(wl.tables)
    mean = np.zeros((30, 40, 4), np.float32)
    acc = np.zeros((30, 40, 4), np.float32)
    for b in range(4):
        pc = wl.push_constants(b)
        mean, _ = sc.render_batch(40, 30, rb.RB200_FLAG_NEE, pc, mean)
        acc, _ = sc.render_batch(40, 30, rb.RB200_FLAG_NEE | rb.RB200_FLAG_ACCUM_SUM, pc, acc)
    assert np.allclose(acc[..., :3] / 4, mean[..., :3], rtol=2e-6, atol=1e-7)


def test_closed_box_ray_count_is_deterministic(ol, rb):
    """No Russian roulette, no throughput cut-off (raytrace.rgen.glsl:106,138-141): in the closed Cornell box every path
    runs all maxBounces segments; rays whose camera ray misses the box silhouette escape on segment 1."""
    wl = rb.configs.cornell(64, 48, samples_per_pixel=2, max_bounces=6)
    sc = ol.OracleScene(wl.tables)
    pc = wl.push_constants(0)
    hits = sc.trace_primary(64, 48, pc)
    _, cnt = sc.render_batch(64, 48, 0, pc)
    assert cnt["paths"] == 64 * 48 * 2
    inside = int((hits["t"] >= 0).sum())
    assert inside * 6 * 2 * 0.9 < cnt["extendRays"] <= 64 * 48 * 2 * 6
    assert cnt["shadowRays"] == 0


def test_nee_and_brdf_sampling_agree_statistically(ol, rb):
    """Lambertian Cornell box: the estimator with NEE + the reference's MIS weights and the as-shipped one (emission
    only) estimate the same image; means agree to a few percent at 64 spp on a small frame."""
    wl = rb.configs.cornell(32, 24, samples_per_pixel=64, max_bounces=8, textured=False)
    sc = ol.OracleScene(wl.tables)
    pc = wl.push_constants(0)
    a, _ = sc.render_batch(32, 24, rb.RB200_FLAG_NEE, pc)
    b, _ = sc.render_batch(32, 24, 0, pc)
    ma, mb = a[..., :3].mean(), b[..., :3].mean()
    assert np.isfinite(a).all() and np.isfinite(b).all()
    assert abs(ma - mb) / mb < 0.25      # the reference's MIS pairs pdfs of the same hit (not textbook): loose bound


def test_light_sampling_hits_the_light(ol, rb):
    """Shadow rays towards sampled light points from the floor centre are unoccluded in the empty Cornell box."""
    wl = rb.configs.cornell(16, 12)
    sc = ol.OracleScene(wl.tables)
    rng = np.random.RandomState(1)
    pts = np.stack([rng.uniform(-0.24, 0.23, 500), np.full(500, 1.989), rng.uniform(-0.22, 0.16, 500)], 1).astype(np.float32)
    o = np.tile(np.array([[0.1, 0.01, 0.05]], np.float32), (500, 1))
    d = pts - o
    dist = np.linalg.norm(d, axis=1).astype(np.float32)
    d = (d / dist[:, None]).astype(np.float32)
    occ = sc.trace_rays(o, d, dist - np.float32(0.001), any_hit=True)
    assert (occ["t"] < 0).all()
    hit = sc.trace_rays(o, d, 1e4)
    assert (hit["instance"] == 1).all()


def test_parallax_height_map_changes_only_bumped_surfaces(ol, rb):
    """texutils.h.glsl:4-41: with bumpMapTexID the UV used for the albedo / normal fetch is shifted along the view
    ray; a height map of 0 (surface level everywhere) shifts by exactly one layer step, a real one moves the texture
    and pushes UVs near the border out of [0, 1] (lambertian / disney then skip the hit)."""
    wl = rb.configs.parallax(64, 48, samples_per_pixel=2, max_bounces=4)
    sc = ol.OracleScene(wl.tables)
    pc = wl.push_constants(0)
    a, ca = sc.render_batch(64, 48, rb.RB200_FLAG_NEE, pc, threads=1)
    b, cb = sc.render_batch(64, 48, rb.RB200_FLAG_NEE, pc, threads=5)
    assert (bits(a) == bits(b)).all() and ca == cb and np.isfinite(a).all()
    # same tables without the height maps
    t2 = rb.configs.parallax(64, 48, samples_per_pixel=2, max_bounces=4).tables
    t2.instanceProperties.view(np.int32).reshape(-1, 30)[:, 15] = -1        # bumpMapTexID, byte offset 60 of 120
    c, cc = ol.OracleScene(t2).render_batch(64, 48, rb.RB200_FLAG_NEE, pc)
    changed = (bits(a) != bits(c)).any(axis=2)
    assert 0.05 < changed.mean() < 0.95          # the bumped surfaces (and what reflects them) changed, the far walls' first hits did not


@pytest.mark.parametrize("name", ["small_mixed_nee", "small_mixed_shipped", "cornell_nee", "parallax_nee"])
def test_golden_fixtures(ol, rb, name):
    """tests/golden/*.npz were written by tests/golden/make_golden.py from the oracle; they pin it against drift
    (the reference has no fixtures of its own). The CUDA path is checked against the same files in test_gpu_parity."""
    path = os.path.join(GOLD, name + ".npz")
    g = np.load(path)
    import golden.make_golden as mg
    wl, flags, batches = mg.CASES[name](rb)
    sc = ol.OracleScene(wl.tables)
    hdr = np.zeros((wl.height, wl.width, 4), np.float32)
    for b in range(batches):
        hdr, _ = sc.render_batch(wl.width, wl.height, flags, wl.push_constants(b), hdr)
    assert (bits(hdr) == bits(g["hdr"])).all()
    assert (ol.postprocess(hdr) == g["ldr"]).all()
    hits = sc.trace_primary(wl.width, wl.height, wl.push_constants(0))
    assert (hits["primitive"] == g["prim"]).all() and (bits(hits["t"]) == bits(g["t"])).all()
