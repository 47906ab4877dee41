"""Developer aid (run on the GPU box, alone or under torchrun with 2 ranks): host-side timestamps of every step of the
multi-GPU end-to-end frame loop (render_batch / snapshot + reduce / present + wait) and device-side completion times of
the frames. NO_NCCL=1 skips the reduce. This is how the 170 ms first-frame stall (kernel code loaded on first launch)
was found; see profiles/r01_summary.md."""
import os, sys, time, importlib
import torch, torch.distributed as dist
sys.path.insert(0, os.getcwd())
rb = importlib.import_module("reina-vk_b200")
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
wl = rb.configs.dragon(1920, 1080)
stream = torch.cuda.Stream()
r = rb.Renderer(wl.width, wl.height, wl.tables, flags=rb.RB200_FLAG_NEE | rb.RB200_FLAG_ACCUM_SUM, device=lr, stream=stream.cuda_stream)
hdr_t = torch.as_tensor(r.hdr_device_array(), device=f"cuda:{lr}")
depth = r.pipeline_depth()
frames = [r.pinned_frame() for _ in range(depth)]
snaps = [torch.empty_like(hdr_t) for _ in range(depth)]
K = 8
use_nccl = world > 1 and not os.environ.get("NO_NCCL")
with torch.cuda.stream(stream):
    for i in range(3):
        r.render_batch(wl.push_constants(rank + i * world))
    if world > 1:
        dist.reduce(hdr_t.clone(), dst=0); dist.barrier()
    r.synchronize(); torch.cuda.synchronize()
    t0 = time.time(); log = []
    evs = []
    for i in range(K):
        r.render_batch(wl.push_constants(rank + (3 + i) * world)); ta = time.time()
        snap = snaps[i % depth]; snap.copy_(hdr_t)
        if use_nccl:
            dist.reduce(snap, dst=0)
        tb = time.time()
        if rank == 0:
            r.present_sum((4 + i) * world, snap.data_ptr()); r.wait_ldr(depth - 1); tc = time.time(); r.read_ldr_async(frames[i % depth])
        else:
            tc = tb
        e = torch.cuda.Event(enable_timing=True); e.record(stream); evs.append(e)
        log.append((i, round((ta - t0) * 1e3, 1), round((tb - t0) * 1e3, 1), round((tc - t0) * 1e3, 1)))
    r.wait_ldr(); torch.cuda.synchronize(); t1 = time.time()
e0 = evs[0]
print(rank, "total ms", round((t1 - t0) * 1e3, 1), "host log (i, after render_batch, after reduce, after wait):", log, "gpu frame done ms rel frame0:", [round(e0.elapsed_time(e), 1) for e in evs], flush=True)
r.close()
