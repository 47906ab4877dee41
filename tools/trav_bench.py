"""Traversal micro-benchmark (GPU box): the traversal kernel alone (rb200_bench_trace) on the headline scene —
camera rays, and incoherent secondary rays leaving the surfaces the camera sees in random directions (closest hit) or
towards the light panel (any hit). Prints one JSON line; tools/sweep_variants.sh runs it for every build variant and
compares the checksums (every variant must return the same hits).
"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rb = importlib.import_module("reina-vk_b200")


def camera_rays(pc, W, H):
    inv_view = np.array(pc.invView, np.float32).reshape(4, 4).T
    inv_proj = np.array(pc.invProjection, np.float32).reshape(4, 4).T
    ys, xs = np.mgrid[0:H, 0:W]
    ndc = np.stack([(xs + 0.5) / W * 2 - 1, -((ys + 0.5) / H * 2 - 1), -np.ones_like(xs, float), np.ones_like(xs, float)], -1).reshape(-1, 4)
    d = ndc @ inv_proj.T
    d = d[:, :3] / d[:, 3:4]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d @ inv_view[:3, :3].T
    o = np.broadcast_to(inv_view[:3, 3], d.shape)
    return o.astype(np.float32), d.astype(np.float32)


def main():
    small = bool(os.environ.get("RB200_BENCH_SMALL"))
    W, H = (480, 270) if small else (1920, 1080)
    wl = rb.configs.dragon(W, H, n_along=1500, n_ring=16) if small else rb.configs.dragon(W, H)
    r = rb.Renderer(W, H, wl.tables, flags=rb.RB200_FLAG_NEE)
    pc = wl.push_constants(0)
    hits = r.trace_primary(pc)
    o, d = camera_rays(pc, W, H)
    ok = hits["t"] > 0
    P = (o + d * hits["t"][:, None])[ok]
    rng = np.random.RandomState(1)
    v = rng.normal(size=P.shape)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    v = np.where(((v * d[ok]).sum(1) > 0)[:, None], -v, v)        # leave on the side the ray came from
    P = (P + v * 1e-3).astype(np.float32)
    v = v.astype(np.float32)
    # shadow rays: towards random points of the light panel (y = 1.989, |x|, |z| <= 0.25)
    L = np.stack([rng.uniform(-0.25, 0.25, len(P)), np.full(len(P), 1.989), rng.uniform(-0.25, 0.25, len(P))], 1).astype(np.float32)
    sd = L - P
    dist = np.linalg.norm(sd, axis=1)
    sd = (sd / dist[:, None]).astype(np.float32)
    res = {"library": os.environ.get("RB200_LIBRARY", "default"), "rays": int(len(P)), "camera_rays": int(len(o))}
    reps = 3 if small else 10
    ms, ck = r.bench_trace(o, d, 1e4, reps=reps)
    res["camera_ms"], res["camera_mrays_s"], res["camera_checksum"] = ms, len(o) / ms / 1e3, ck
    ms, ck = r.bench_trace(P, v, 1e4, reps=reps)
    res["secondary_ms"], res["secondary_mrays_s"], res["secondary_checksum"] = ms, len(P) / ms / 1e3, ck
    ms, ck = r.bench_trace(P, sd, dist - 1e-3, any_hit=True, reps=reps)
    res["shadow_ms"], res["shadow_mrays_s"], res["shadow_checksum"] = ms, len(P) / ms / 1e3, ck
    info = r.bvh_info()
    res["wide_nodes"], res["node_bytes"], res["build_ms"] = info["numWideNodes"], info["nodeBytes"], info["buildMs"]
    r.close()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
