"""Developer check run on the GPU box: CUDA path vs oracle on small scenes, with verbose diagnostics.
Usage: python tools/gpu_check.py [--scene small_mixed|cornell] [--w 160 --h 120] [--spp 2] [--bounces 8]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402

rb = ol.rb


def compare_hits(a, b, label):
    same_prim = (a["instance"] == b["instance"]) & (a["primitive"] == b["primitive"])
    hit = a["t"] >= 0
    t_equal = a["t"].view(np.uint32) == b["t"].view(np.uint32)
    uv_equal = (a["u"].view(np.uint32) == b["u"].view(np.uint32)) & (a["v"].view(np.uint32) == b["v"].view(np.uint32))
    n = a.size
    print(f"[{label}] rays {n}  hits {int(hit.sum())}  prim match {same_prim.mean() * 100:.4f}%  "
          f"t bit-equal {t_equal.mean() * 100:.4f}%  uv bit-equal {uv_equal.mean() * 100:.4f}%")
    bad = np.nonzero(~(same_prim & t_equal))[0]
    for i in bad[:8]:
        print("   mismatch", i, "gpu", a[i], "oracle", b[i])
    return bool(same_prim.all() and t_equal.all() and uv_equal.all())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="small_mixed")
    ap.add_argument("--w", type=int, default=160)
    ap.add_argument("--h", type=int, default=120)
    ap.add_argument("--spp", type=int, default=2)
    ap.add_argument("--bounces", type=int, default=8)
    ap.add_argument("--batches", type=int, default=2)
    ap.add_argument("--nee", type=int, default=1)
    args = ap.parse_args()

    if args.scene == "cornell":
        wl = rb.configs.cornell(args.w, args.h, nee=bool(args.nee), samples_per_pixel=args.spp, max_bounces=args.bounces)
    elif args.scene == "dragon":
        wl = rb.configs.dragon(args.w, args.h, n_along=1500, n_ring=16, nee=bool(args.nee), samples_per_pixel=args.spp,
                               max_bounces=args.bounces)
    else:
        wl = rb.configs.small_mixed(args.w, args.h, nee=bool(args.nee), samples_per_pixel=args.spp, max_bounces=args.bounces)
    flags = rb.RB200_FLAG_NEE if args.nee else 0
    print("scene", wl.name, "triangles", wl.tables.num_triangles())

    t0 = time.time()
    r = rb.Renderer(wl.width, wl.height, wl.tables, flags=flags | rb.RB200_FLAG_COUNT_BVH)
    print("scene_create s", round(time.time() - t0, 3), json.dumps(r.bvh_info()))
    osc = ol.OracleScene(wl.tables)

    pc = wl.push_constants(0)
    ok = compare_hits(r.trace_primary(pc), osc.trace_primary(wl.width, wl.height, pc), "primary")

    rng = np.random.RandomState(1)
    n = 20000
    o = rng.uniform(-0.9, 0.9, (n, 3)).astype(np.float32) + np.array([0, 1, 0], np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    gh, oh = r.trace_rays(o, d, 1e4), osc.trace_rays(o, d, 1e4, brute=True)
    ok &= compare_hits(gh, oh, "random closest vs brute force")
    octant = (d[:, 0] > 0) * 4 + (d[:, 1] > 0) * 2 + (d[:, 2] > 0) * 1
    badm = ~((gh["instance"] == oh["instance"]) & (gh["primitive"] == oh["primitive"]))
    kz = np.argmax(np.abs(d), axis=1)
    print("   mismatch rate by direction octant (x+ y+ z+ bits):", [round(float(badm[octant == k].mean()), 3) for k in range(8)])
    print("   mismatch rate by dominant axis:", [round(float(badm[kz == k].mean()), 3) for k in range(3)])
    tm = rng.uniform(0.05, 2.0, n).astype(np.float32)
    ga, oa = r.trace_rays(o, d, tm, any_hit=True), osc.trace_rays(o, d, tm, any_hit=True, brute=True)
    occ_same = ((ga["t"] >= 0) == (oa["t"] >= 0))
    print(f"[any-hit] agree {occ_same.mean() * 100:.4f}%  occluded {(oa['t'] >= 0).mean() * 100:.1f}%")
    ok &= bool(occ_same.all())

    hdr_o = np.zeros((wl.height, wl.width, 4), np.float32)
    for b in range(args.batches):
        pc = wl.push_constants(b)
        t0 = time.time()
        r.render_batch(pc)
        r.synchronize()
        tg = time.time() - t0
        last, cum = r.stats()
        t0 = time.time()
        hdr_o, cnt = osc.render_batch(wl.width, wl.height, flags, pc, hdr_o)
        tc = time.time() - t0
        hdr_g = r.read_hdr()
        eq = (hdr_g.view(np.uint32) == hdr_o.view(np.uint32)).all(axis=2)
        diff = np.abs(hdr_g[..., :3] - hdr_o[..., :3])
        print(f"[batch {b}] gpu {tg * 1e3:.1f} ms  oracle {tc * 1e3:.1f} ms  pixels bit-equal {eq.mean() * 100:.4f}%  "
              f"max abs diff {np.nanmax(diff):.3e}  mean gpu {np.nanmean(hdr_g[..., :3]):.6f} oracle {np.nanmean(hdr_o[..., :3]):.6f}")
        print("   gpu stats", last, "\n   oracle   ", cnt)
        ok &= bool(eq.all())
        ok &= last["extendRays"] == cnt["extendRays"] and last["shadowRays"] == cnt["shadowRays"] and last["paths"] == cnt["paths"]
        if not eq.all():
            ys, xs = np.nonzero(~eq)
            for y, x in list(zip(ys, xs))[:6]:
                print("   pixel", x, y, "gpu", hdr_g[y, x], "oracle", hdr_o[y, x])

    r.postprocess()
    ldr_g = r.read_ldr()
    ldr_o = ol.postprocess(hdr_o)
    eq = (ldr_g == ldr_o).all(axis=2)
    print(f"[post] LDR pixels equal {eq.mean() * 100:.4f}%  max diff {np.abs(ldr_g.astype(int) - ldr_o.astype(int)).max()}")
    ok &= bool(eq.all())
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    try:
        from PIL import Image
        Image.fromarray(ldr_g[..., :3]).save(os.path.join(ROOT, "gpurun_out", f"gpu_{wl.name}.png"))
    except Exception as e:  # pragma: no cover
        print("png save failed", e)
    print("RESULT", "PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
