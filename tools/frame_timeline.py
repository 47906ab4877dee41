"""Developer aid (run on the GPU box): host- and device-side timeline of the single-GPU end-to-end frame loop of
bench.py (render_batch, postprocess, wait_ldr, read_ldr_async). RB200_NO_GRAPH=1 launches the wave loop kernel by kernel;
RB200_STAGGER_WAVE=-1 switches the lane stagger gate off. Shows whether the lanes run staggered or in lockstep (frames
completing in bursts); see profiles/r01_summary.md."""
import os, sys, time, importlib
import torch
sys.path.insert(0, os.getcwd())
rb = importlib.import_module("reina-vk_b200")
wl = rb.configs.dragon(1920, 1080)
stream = torch.cuda.Stream()
r = rb.Renderer(wl.width, wl.height, wl.tables, flags=rb.RB200_FLAG_NEE, stream=stream.cuda_stream)
depth = r.pipeline_depth()
frames = [r.pinned_frame() for _ in range(depth)]
K = 12
with torch.cuda.stream(stream):
    for i in range(5):
        r.render_batch(wl.push_constants(i))
    r.postprocess(); r.read_ldr()
    r.synchronize(); torch.cuda.synchronize()
    t0 = time.time(); log = []; evs = []
    for i in range(K):
        r.render_batch(wl.push_constants(5 + i)); ta = time.time()
        r.postprocess(); r.wait_ldr(depth - 1); tc = time.time(); r.read_ldr_async(frames[i % depth])
        e = torch.cuda.Event(enable_timing=True); e.record(stream); evs.append(e)
        log.append((i, round((ta - t0) * 1e3, 1), round((tc - t0) * 1e3, 1)))
    r.wait_ldr(); torch.cuda.synchronize(); t1 = time.time()
print("graph" if not os.environ.get("RB200_NO_GRAPH") else "nograph", "total ms", round((t1 - t0) * 1e3, 1), "per step", round((t1 - t0) * 1e3 / K, 1))
print(" host (i, after render_batch, after wait):", log)
print(" gpu frame done ms rel frame0:", [round(evs[0].elapsed_time(e), 1) for e in evs], flush=True)
r.close()
