"""Probe run on the GPU box under ncu: how much does the ORDER of incoherent rays matter to the traversal kernel?
Traces the same set of secondary rays (origins on the headline scene's surfaces, random outgoing directions) in
(a) pixel order, (b) a random permutation, (c) sorted by Morton code of the origin + direction octant, and prints the
wall time of each rb200_trace_rays call (copies included); the kernel times come from the ncu launch list:
  ncu --metrics gpu__time_duration.sum -k regex:k_trace_query --csv python tools/coherence_probe.py
"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rb = importlib.import_module("reina-vk_b200")


def part1by2(v):
    v = v.astype(np.uint64) & np.uint64(0x3FF)
    v = (v | (v << np.uint64(16))) & np.uint64(0x30000FF)
    v = (v | (v << np.uint64(8))) & np.uint64(0x300F00F)
    v = (v | (v << np.uint64(4))) & np.uint64(0x30C30C3)
    v = (v | (v << np.uint64(2))) & np.uint64(0x9249249)
    return v


def main():
    W, H = 1920, 1080
    wl = rb.configs.dragon(W, H)
    r = rb.Renderer(W, H, wl.tables, flags=rb.RB200_FLAG_NEE)
    pc = wl.push_constants(0)
    hits = r.trace_primary(pc)
    # camera rays without jitter (good enough to place origins)
    inv_view = np.array(pc.invView, np.float32).reshape(4, 4).T
    inv_proj = np.array(pc.invProjection, np.float32).reshape(4, 4).T
    ys, xs = np.mgrid[0:H, 0:W]
    ndc = np.stack([(xs + 0.5) / W * 2 - 1, -((ys + 0.5) / H * 2 - 1), -np.ones_like(xs, float), np.ones_like(xs, float)], -1).reshape(-1, 4)
    d = ndc @ inv_proj.T
    d = d[:, :3] / d[:, 3:4]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d @ inv_view[:3, :3].T
    o = np.broadcast_to(inv_view[:3, 3], d.shape)
    ok = hits["t"] > 0
    P = (o + d * hits["t"][:, None])[ok]
    din = d[ok]
    rng = np.random.RandomState(1)
    v = rng.normal(size=P.shape)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    v = np.where(((v * din).sum(1) > 0)[:, None], -v, v)        # leave on the side the ray came from
    P = (P + v * 1e-3).astype(np.float32)
    v = v.astype(np.float32)
    n = P.shape[0]
    lo, hi = P.min(0), P.max(0)
    q = np.clip(((P - lo) / (hi - lo) * 1023), 0, 1023).astype(np.uint64)
    morton = (part1by2(q[:, 0]) << np.uint64(2)) | (part1by2(q[:, 1]) << np.uint64(1)) | part1by2(q[:, 2])
    octant = ((v[:, 0] < 0).astype(np.uint64) << np.uint64(2)) | ((v[:, 1] < 0).astype(np.uint64) << np.uint64(1)) | (v[:, 2] < 0).astype(np.uint64)
    orders = {
        "pixel": np.arange(n),
        "random": rng.permutation(n),
        "morton+octant": np.argsort((morton << np.uint64(3)) | octant, kind="stable"),
        "octant+morton": np.argsort((octant << np.uint64(30)) | morton, kind="stable"),
        "morton20+octant": np.argsort(((morton >> np.uint64(12)) << np.uint64(3)) | octant, kind="stable"),
    }
    ref = None
    for name, idx in orders.items():
        for rep in range(2):
            t0 = time.time()
            h = r.trace_rays(P[idx], v[idx], 1e4)
            dt = time.time() - t0
        back = np.empty_like(h)
        back[idx] = h
        if ref is None:
            ref = back
        assert (back["primitive"] == ref["primitive"]).all() and (back["t"].view(np.uint32) == ref["t"].view(np.uint32)).all()
        print("%-16s n=%d  call %.1f ms  hit %.1f %%" % (name, n, dt * 1e3, (h["t"] > 0).mean() * 100), flush=True)
    r.close()


if __name__ == "__main__":
    main()
