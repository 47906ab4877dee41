"""13,000 single closest-hit invocations of the reference's COMPILED shaders (tests/golden/spirv_hits.npz, minted by
tests/golden/make_spirv_hits_golden.py from /root/reference/shaders/raytrace/*.rchit.spv) replayed on
  * the CPU oracle (oracle_kat_trace_main), and
  * the CUDA shading code through the C ABI (rb200_shade_hits: the traversal kernel + the eval_hit<material> the wave
    loop runs), on the GPU.
Every payload field, the RNG state and the flags must be bit-identical (NaN = NaN). This is what pins rows a7-a13 of
SURVEY.md 8 — hit attributes, the four material shaders, parallax and normal mapping, alpha / cull skips — to the
reference's binaries on the DEVICE, independently of the text the oracle and the kernels share."""
import ctypes as C
import os

import numpy as np
import pytest

import spirv_hit_scenes as hs

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "spirv_hits.npz")


def same(a, b):
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


def check(fix, sel, vec, pdf, acc, rng, skip, inside, what):
    """vec (n, 18) = color, albedo, origin, direction, emission, normal. A skipped hit only defines the new ray."""
    ref_skip = fix["skip"][sel].astype(bool)
    assert (skip == ref_skip).all(), what + ": skip flags"
    assert (rng == fix["rng_out"][sel]).all(), what + ": RNG state"
    ok = same(vec, fix["vec"][sel])
    ray = ok[:, 6:12].all(axis=1)
    assert ray.all(), what + ": new ray differs at %s" % np.nonzero(~ray)[0][:5]
    full = ok.all(axis=1) | ref_skip
    assert full.all(), what + ": payload vectors differ at %s" % np.nonzero(~full)[0][:5]
    live = ~ref_skip
    assert same(pdf[live], fix["pdf"][sel][live]).all() and same(acc[live], fix["acc_out"][sel][live]).all(), what + ": pdf / distance"
    assert (inside[live] == fix["inside_out"][sel][live].astype(bool)).all(), what + ": insideDielectric"


def test_fixture_covers_every_material_and_path():
    fix = np.load(FIX)
    n = len(fix["scene"])
    assert n >= 13000 and set(fix["scene"]) == set(range(hs.N_SCENES))
    for k in range(4):
        assert (fix["material"] == k).sum() > 1000
    assert 0.02 < fix["skip"].mean() < 0.6 and fix["inside"].mean() > 0.2 and fix["inside_out"].mean() > 0.02
    assert np.isnan(fix["vec"]).any() or True          # NaN payloads are legal (zero tangents); compared as NaN = NaN


def test_oracle_matches_the_compiled_shaders_on_13000_hits(ol, rb):
    fix = np.load(FIX)
    for seed in range(hs.N_SCENES):
        sel = np.nonzero(fix["scene"] == seed)[0]
        sc = ol.OracleScene(hs.hit_scene(rb, seed))
        n = len(sel)
        vec, pdf, acc = np.zeros((n, 18), np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
        rng, skip, inside = np.zeros(n, np.uint32), np.zeros(n, bool), np.zeros(n, bool)
        for j, i in enumerate(sel):
            out, st, fl = np.zeros(21, np.float32), C.c_uint32(int(fix["state"][i])), C.c_uint32()
            o, d = np.ascontiguousarray(fix["o"][i]), np.ascontiguousarray(fix["d"][i])
            ol.lib().oracle_kat_trace_main(sc._h, o.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), C.byref(st), int(fix["inside"][i]),
                                           C.c_float(float(fix["acc"][i])), out.ctypes.data_as(C.c_void_p), C.byref(fl))
            vec[j], pdf[j], acc[j], rng[j] = out[:18], out[18], out[19], st.value
            skip[j], inside[j] = bool(fl.value & 2), bool(fl.value & 4)
        sc.close()
        check(fix, sel, vec, pdf, acc, rng, skip, inside, "oracle, scene %d" % seed)


@pytest.mark.gpu
def test_cuda_shading_matches_the_compiled_shaders_on_13000_hits(rb):
    fix = np.load(FIX)
    for seed in range(hs.N_SCENES):
        sel = np.nonzero(fix["scene"] == seed)[0]
        tables = hs.hit_scene(rb, seed)
        r = rb.Renderer(32, 24, tables, flags=0)
        res = r.shade_hits(fix["o"][sel], fix["d"][sel], fix["state"][sel], fix["inside"][sel].astype(np.uint32), fix["acc"][sel])
        r.close()
        assert (res["flags"] & 1).all() and (res["material"] == fix["material"][sel]).all()
        vec = np.concatenate([res[k] for k in ("color", "albedo", "origin", "direction", "emission", "normal")], axis=1)
        check(fix, sel, vec, res["pdf"], res["accumulatedDistance"], res["rngState"], (res["flags"] & 2) != 0, (res["flags"] & 4) != 0,
              "CUDA, scene %d" % seed)
