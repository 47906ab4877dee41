"""Drive rb200_postprocess at 1080p (for `ncu -k regex:"k_blur|k_combine"`) and print its time per call:
  python tools/post_profile.py [W H]      a synthetic HDR frame with bright pixels everywhere (no tile can be skipped)
  python tools/post_profile.py c3         the headline scene's frame after 4 batches (what a render presents)
"""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rb = importlib.import_module("reina-vk_b200")


def timed(r, n=50):
    for _ in range(3):
        r.postprocess()
    r.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r.postprocess()
    r.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


if len(sys.argv) > 1 and sys.argv[1] == "c3":
    wl = rb.configs.dragon(1920, 1080)
    r = rb.Renderer(wl.width, wl.height, wl.tables, flags=rb.RB200_FLAG_NEE)
    for b in range(4):
        r.render_batch(wl.push_constants(b))
    hdr = r.read_hdr()
    lum = hdr[..., 0] * 0.299 + hdr[..., 1] * 0.587 + hdr[..., 2] * 0.114
    print(f"C3 frame: {float((lum >= 1.0).mean()) * 100:.3f} % of the pixels are above the bloom threshold")
    print(f"postprocess 1920x1080 on the C3 frame: {timed(r):.4f} ms per call (blur_x + blur_y + combine_tonemap; 274 MB algorithmic)")
else:
    W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
    wl = rb.configs.small_mixed(W, H)
    r = rb.Renderer(W, H, wl.tables)
    rng = np.random.default_rng(3)
    img = rng.gamma(0.6, 0.8, (H, W, 4)).astype(np.float32)      # mostly dim, but pixels above the bloom threshold in every tile
    img[..., 3] = 1.0
    r.write_hdr(img)
    print(f"postprocess {W}x{H}, bright pixels in every tile: {timed(r, 20):.4f} ms per call")
r.close()
