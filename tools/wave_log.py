"""Run on the GPU box: per-wave queue counters and kernel times of one headline batch (RB200_WAVE_LOG developer aid).
usage: python tools/wave_log.py [out.csv]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/wave_log.csv"
os.environ["RB200_WAVE_LOG"] = out
rb = importlib.import_module("reina-vk_b200")
wl = rb.configs.dragon(1920, 1080)
r = rb.Renderer(wl.width, wl.height, wl.tables, flags=rb.RB200_FLAG_NEE | rb.RB200_FLAG_TIME_KERNELS)
for b in range(3):
    r.render_batch(wl.push_constants(b))
    r.synchronize()
print(r.kernel_times())
r.close()
