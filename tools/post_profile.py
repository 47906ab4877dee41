"""Drive rb200_postprocess at 1080p on a synthetic HDR frame (for `ncu -k regex:"k_blur|k_combine"`) and print its
CUDA-event time per call."""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rb = importlib.import_module("reina-vk_b200")

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
wl = rb.configs.small_mixed(W, H)
r = rb.Renderer(W, H, wl.tables)
rng = np.random.default_rng(3)
img = rng.gamma(0.6, 0.8, (H, W, 4)).astype(np.float32)      # mostly dim, a few pixels above the bloom threshold
img[..., 3] = 1.0
r.write_hdr(img)
for _ in range(3):
    r.postprocess()
r.synchronize()
t0 = time.perf_counter()
n = 20
for _ in range(n):
    r.postprocess()
r.synchronize()
print(f"postprocess {W}x{H}: {(time.perf_counter() - t0) / n * 1e3:.3f} ms per call (blur_x + blur_y + combine_tonemap)")
r.close()
