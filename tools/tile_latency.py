"""Latency mode on N devices of one box, through rb200_group_* — one process, a context and a host
thread per device, interleaved 32 x 32 tiles; a frame is assembled by peer stores into device 0's image (default when the
devices have peer access) or by one ncclReduce per presented frame (RB200_GROUP_TILES_REDUCE=1).

For N in --gpus: time from queueing one batch (8 spp) to its tonemapped frame on the host, one frame at a time (the
interactive case: nothing in flight behind it), and the frame rate of the same loop; the N-device frame must equal the
1-device frame byte for byte (every pixel is accumulated by exactly one device in batch order, x + 0 = x).
One JSON line per N on stdout.
usage: python tools/tile_latency.py --gpus 1,2,4,8 [--config c3] [--frames 12]"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1,2")
    ap.add_argument("--config", default="c3")
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=4)
    args = ap.parse_args()
    rb = importlib.import_module("reina-vk_b200")
    wl = bench.build_workload(rb, config=args.config)
    ref_ldr = ref_hdr = None
    for n in [int(x) for x in args.gpus.split(",")]:
        g = rb.Group(wl.width, wl.height, wl.tables, list(range(n)), flags=rb.RB200_FLAG_NEE, tiles=True)
        hashes = {g.bvh_info(i)["hash"] for i in range(n)}
        peer = g.uses_peer_stores()
        # one frame at a time: batch b of every device's tiles -> reduce -> bloom + tonemap -> host
        lat = []
        for b in range(args.warmup + args.frames):
            t0 = time.perf_counter()
            g.render_batches(wl.push_constants(b), b, 1)
            g.present()
            ldr = g.read_ldr()
            lat.append((time.perf_counter() - t0) * 1e3)
        hdr = g.read_hdr()
        st = g.stats()
        lat = np.array(lat[args.warmup:])
        if ref_ldr is None:
            ref_ldr, ref_hdr = ldr, hdr
        out = {"mode": "interleaved tiles 32x32 (rb200_group_*, RB200_FLAG_GROUP_TILES)", "config": args.config, "n_gpus": n,
               "frame_assembly": "peer stores from k_accumulate into device 0's image (no collective)" if peer else
                                 ("one ncclReduce per frame" if n > 1 else "single device"),
               "width": wl.width, "height": wl.height, "frames": args.frames, "spp_per_frame": bench.SPP,
               "frame_latency_ms_median": float(np.median(lat)), "frame_latency_ms_min": float(lat.min()),
               "frame_latency_ms_max": float(lat.max()), "frames_per_s": float(1e3 / lat.mean()),
               "rays_total": int(st["extendRays"] + st["shadowRays"]), "bvh_hashes_equal": len(hashes) == 1,
               "ldr_identical_to_first_n": bool(np.array_equal(ldr, ref_ldr)),
               "hdr_identical_to_first_n": bool(np.array_equal(hdr.view(np.uint32), ref_hdr.view(np.uint32))),
               "timing": "host wall clock around render_batches + present + blocking read_ldr (a latency, so the host's view is the metric)"}
        print(json.dumps(out), flush=True)
        g.close()


if __name__ == "__main__":
    main()
