"""Run on the GPU box: what one device of an 8-device latency-mode group does per frame, measured on ONE GPU — a context
that traces the tiles of rank 0 of 8 (1/8 of the pixels), one batch at a time with nothing in flight behind it: time per
batch (CUDA events), then the per-wave queue counters and kernel times of one such batch (RB200_WAVE_LOG).
usage: python tools/tile_wave_probe.py [ranks] [out.csv]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ranks = int(sys.argv[1]) if len(sys.argv) > 1 else 8
out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/tile_wave_log.csv"
rb = importlib.import_module("reina-vk_b200")
import torch  # noqa: E402

wl = rb.configs.dragon(1920, 1080)
r = rb.Renderer(wl.width, wl.height, wl.tables, flags=rb.RB200_FLAG_NEE)
r.set_tiles(0, ranks)
ms = []
for b in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    r.render_batch(wl.push_constants(b))
    r.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
last, _ = r.stats()
r.close()
print("rank 0 of %d: ms per batch (one at a time)" % ranks, [round(x, 2) for x in ms], "rays per batch", last["extendRays"] + last["shadowRays"], "waves", last["waves"], "launches", last["kernelLaunches"])
os.environ["RB200_WAVE_LOG"] = out
r = rb.Renderer(wl.width, wl.height, wl.tables, flags=rb.RB200_FLAG_NEE | rb.RB200_FLAG_TIME_KERNELS)
r.set_tiles(0, ranks)
for b in range(2):
    r.render_batch(wl.push_constants(b))
    r.synchronize()
kt = r.kernel_times()
r.close()
print("kernel times of the last batch:", kt)
rows = np.genfromtxt(out, delimiter=",", names=True)
tot = sum(rows[c].sum() for c in rows.dtype.names if c.endswith("_us"))
print("waves logged", len(rows), "sum of kernel us", round(float(tot)), "waves with rays", int((rows["rays"] > 0).sum()), "waves with < 2000 rays", int(((rows["rays"] > 0) & (rows["rays"] < 2000)).sum()))
