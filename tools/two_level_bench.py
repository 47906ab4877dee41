"""Flattened against two-level (RB200_FLAG_TWO_LEVEL) on a heavily instanced scene: hierarchy memory, build time and
Mrays/s of the wave loop (8 spp x 8 bounces at 1280x720, NEE on), one JSON line per mode.
usage: python tools/two_level_bench.py [--grid 16] [--segments 96] [--steps 6]"""
import argparse
import importlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=16)
    ap.add_argument("--segments", type=int, default=96)
    ap.add_argument("--steps", type=int, default=6)
    args = ap.parse_args()
    rb = importlib.import_module("reina-vk_b200")
    wl = rb.configs.instanced(width=1280, height=720, grid=args.grid, segments=args.segments, rings=args.segments // 2,
                              samples_per_pixel=8, max_bounces=8)
    for name, extra in (("flattened", 0), ("two_level", rb.RB200_FLAG_TWO_LEVEL)):
        t0 = time.time()
        r = rb.Renderer(wl.width, wl.height, wl.tables, flags=rb.RB200_FLAG_NEE | extra)
        r.synchronize()
        create_s = time.time() - t0
        info = r.bvh_info()
        for b in range(3):
            r.render_batch(wl.push_constants(b))
        r.synchronize()
        _, c0 = r.stats()
        t0 = time.time()
        for b in range(args.steps):
            r.render_batch(wl.push_constants(3 + b))
        r.synchronize()
        dt = time.time() - t0
        _, c1 = r.stats()
        rays = (c1["extendRays"] - c0["extendRays"]) + (c1["shadowRays"] - c0["shadowRays"])
        print(json.dumps({"mode": name, "instances": int(wl.tables.numInstances), "instanced_triangles": int(wl.tables.num_triangles()),
                          "stored_triangles": info["numTriangles"], "wide_nodes": info["numWideNodes"],
                          "hierarchy_bytes": info["nodeBytes"] + info["triangleBytes"],
                          "shading_record_bytes": 192 * info["numTriangles"], "build_ms": info["buildMs"],
                          "scene_create_s": create_s, "mrays_s": rays / dt / 1e6, "ms_per_step": dt / args.steps * 1e3,
                          "timing": "host wall clock around %d render_batch calls + synchronize" % args.steps}), flush=True)
        r.close()


if __name__ == "__main__":
    main()
