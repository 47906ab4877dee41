#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 path-tracing core.

Metric (BASELINE.json): Mrays/s (primary + bounce + shadow rays actually traced, counted by device counters) on the
Stanford-Dragon-class scene (procedural stand-in, 871,414 triangles, Disney BSDF, NEE + MIS) at 1920x1080.
One "step" = one rb200_render_batch call = one sample batch of the whole frame (samples_per_pixel = 8 and
max_bounces = 16, the reference's config/config.toml defaults) — i.e. one vkCmdTraceRaysKHR(W,H,1) of the reference.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA arm (under torchrun for N > 1)
  python bench.py --impl reference [--steps K] [--warmup W]      the reference's algorithm on the host cores (oracle)

Multi-GPU: sample split — rank r renders batches r, r+N, ... with the unmodified seed formula into a
local SUM image; one NCCL reduce of the float4 images to rank 0 closes the timed region. Weak scaling: every rank
renders K batches.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SPP, BOUNCES = 8, 16       # config/config.toml defaults: one step = one batch of 8 samples per pixel, <= 16 segments each
METRIC = "Mrays/s (primary+bounce+shadow), Stanford Dragon stand-in 1080p"

# BASELINE.json `configs`; c3 is the headline the metric is quoted on, the others are reported by --config for context
CONFIGS = {
    "c1": dict(name="C1 Cornell box + light, Lambertian, 800x600, 8 bounces, NEE+MIS (64 spp = 8 steps)",
               build=lambda rb: rb.configs.cornell(800, 600, samples_per_pixel=SPP, max_bounces=8), bounces=8),
    "c2": dict(name="C2 bunny stand-in (2 x 81,920 triangles) metal + glass with Beer's-law absorption, 1920x1080, showroom",
               build=lambda rb: rb.configs.bunny(1920, 1080, levels=6, samples_per_pixel=SPP, max_bounces=BOUNCES), bounces=BOUNCES),
    "c3": dict(name="C3 dragon stand-in (torus-knot tube + value-noise displacement), Disney BSDF, NEE+MIS, showroom + light panel",
               build=lambda rb: rb.configs.dragon(1920, 1080, samples_per_pixel=SPP, max_bounces=BOUNCES), bounces=BOUNCES),
    "c4": dict(name="C4 plant-class scene: 34,000 alpha-tested leaf cards, soil, normal-mapped Disney pot, 1920x1080, showroom",
               build=lambda rb: rb.configs.plant(1920, 1080, n_leaves=34000, samples_per_pixel=SPP, max_bounces=BOUNCES), bounces=BOUNCES),
    "c5": dict(name="C5 showroom + Max-Planck stand-in + one sphere per material, bloom + tonemap, 3840x2160",
               build=lambda rb: rb.configs.showroom_mixed(3840, 2160, levels=6, samples_per_pixel=SPP, max_bounces=BOUNCES), bounces=BOUNCES),
}


def build_workload(rb, small=False, config="c3"):
    if small:   # debugging aid only (RB200_BENCH_SMALL=1); never used for reported numbers
        return rb.configs.dragon(480, 270, n_along=1500, n_ring=16, samples_per_pixel=SPP, max_bounces=BOUNCES)
    return CONFIGS[config]["build"](rb)


def config_dict(wl, extra=None, config="c3"):
    d = {"workload": CONFIGS[config]["name"], "config": config,
         "triangles": wl.tables.num_triangles(), "width": wl.width, "height": wl.height,
         "samples_per_pixel_per_step": SPP, "max_bounces": CONFIGS[config]["bounces"],
         "l2_policy": "inputs larger than L2: every wave sweeps the path state of the batches in flight (0.5 GB per 2.07 M "
                      "slots); the BVH (48 MB for the headline scene) is meant to stay L2-resident"}
    if extra:
        d.update(extra)
    return d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                sm.append(float(f[1])); mx.append(float(f[2]))
                for n, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_baseline(rb, wl, seconds_hint=20.0):
    """Oracle (CPU restatement of the reference's shaders) on all host cores, bounded sample of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    sw, sh = 480, 270
    osc = ol.OracleScene(wl.tables)
    pc = wl.push_constants(0, samples_per_pixel=1)
    # same camera, reduced resolution
    from importlib import import_module
    cam = import_module("reina-vk_b200").camera
    kw = dict(wl.pc_kwargs); kw["samples_per_pixel"] = 1
    rays, t_total, b = 0, 0.0, 0
    while t_total < seconds_hint and b < 64:
        pc = cam.push_constants(sw, sh, total_emissive_weight=wl.tables.totalEmissiveWeight, sample_batch=b, **kw)
        t0 = time.time()
        _, cnt = osc.render_batch(sw, sh, rb.RB200_FLAG_NEE, pc)
        t_total += time.time() - t0
        rays += cnt["extendRays"] + cnt["shadowRays"]
        b += 1
    osc.close()
    return {"value": rays / t_total / 1e6, "unit": "Mrays/s", "cores": ol.NTHREADS, "kind": "port",
            "sample": f"same scene and camera at {sw}x{sh}, 1 spp x {b} batches, {BOUNCES} bounces, NEE on "
                      f"({rays} rays in {t_total:.1f} s)"}


def run_reference(args, rank, world):
    """--impl reference: the reference's own algorithm for this path on the host cores. The reference itself (Vulkan
    RT + GLSL) cannot be built or run in this image (no Vulkan loader / ICD / glslc), so this arm times the oracle —
    its line-by-line CPU restatement — with every host thread."""
    if rank != 0:
        return 0
    rb = importlib.import_module("reina-vk_b200")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    wl = build_workload(rb, small=bool(os.environ.get("RB200_BENCH_SMALL")))
    sw, sh = 480, 270
    kw = dict(wl.pc_kwargs); kw["samples_per_pixel"] = 1
    osc = ol.OracleScene(wl.tables)

    def step(b):
        pc = rb.camera.push_constants(sw, sh, total_emissive_weight=wl.tables.totalEmissiveWeight, sample_batch=b, **kw)
        _, cnt = osc.render_batch(sw, sh, rb.RB200_FLAG_NEE, pc)
        return cnt["extendRays"] + cnt["shadowRays"]
    for w in range(args.warmup):
        step(w)
    t0 = time.time()
    rays = sum(step(args.warmup + k) for k in range(args.steps))
    dt = time.time() - t0
    val = rays / dt / 1e6
    sample = f"each step = one 1-spp batch of the same scene/camera at {sw}x{sh} ({BOUNCES} bounces, NEE on)"
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": config_dict(wl, {"reference_sample": sample}),
           "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": ol.NTHREADS, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS), help="BASELINE.json config (default: c3, the headline)")
    ap.add_argument("--no-nee", action="store_true", help="the estimator as shipped upstream (next-event estimation compiled out)")
    ap.add_argument("--no-as-shipped", action="store_true", help="skip the extra as-shipped (NEE off) measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: reina-vk_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    rb = importlib.import_module("reina-vk_b200")
    wl = build_workload(rb, small=bool(os.environ.get("RB200_BENCH_SMALL")), config=args.config)
    stream = torch.cuda.Stream()
    nee_flag = 0 if args.no_nee else rb.RB200_FLAG_NEE
    flags = nee_flag | (rb.RB200_FLAG_ACCUM_SUM if world > 1 else 0)
    t0 = time.time()
    r = rb.Renderer(wl.width, wl.height, wl.tables, flags=flags, device=local_rank, stream=stream.cuda_stream)
    r.synchronize()
    scene_create_s = time.time() - t0
    bvh = r.bvh_info()
    parity_check = None
    total_batches = 0       # batches folded into the ranks' SUM images so far, over all ranks
    if world > 1:
        # the library's own NCCL communicator (rb200_context_comm_init -> ncclCommInitRank); torch.distributed only
        # carries the 128-byte unique id to the other ranks and the timing / counter reductions below
        uid = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local_rank}")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(rb.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, src=0)
        r.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
        # every replica of the BVH must be the same structure (deterministic build)
        hashes = [None] * world
        dist.all_gather_object(hashes, int(bvh["hash"]))
        assert len(set(hashes)) == 1, f"BVH hashes differ across ranks: {hashes}"
        # parity of the multi-GPU path: 2 batches per rank, reduced by the library, against the same 2 * world batches
        # rendered by ONE GPU (rank 0, a second context); only the fp32 summation order differs
        with torch.cuda.stream(stream):
            for i in range(2):
                r.render_batch(wl.push_constants(rank + i * world))
            total_batches += 2 * world
            r.reduce_present(total_batches)
            r.synchronize()
        if rank == 0:
            multi = torch.as_tensor(r.reduced_device_array(), device=f"cuda:{local_rank}").cpu().numpy() / float(total_batches)
            os.environ["RB200_LANES"], os.environ["RB200_ENGINES"] = "2", "1"
            r1 = rb.Renderer(wl.width, wl.height, wl.tables, flags=flags, device=local_rank, stream=stream.cuda_stream)
            del os.environ["RB200_LANES"], os.environ["RB200_ENGINES"]
            with torch.cuda.stream(stream):
                for b in range(2 * world):
                    r1.render_batch(wl.push_constants(b))
                single = r1.read_hdr() / float(total_batches)
            r1.close()
            num = np.abs(multi[..., :3] - single[..., :3]).max()
            den = max(1e-30, float(np.abs(single[..., :3]).max()))
            rel = np.abs(multi[..., :3] - single[..., :3]) / np.maximum(np.abs(single[..., :3]), 1e-3)
            parity_check = {"batches": 2 * world, "max_abs_diff": float(num), "max_abs_diff_over_max": float(num / den),
                            "max_rel_diff": float(rel.max()), "tolerance_rel": 1e-4, "bvh_hashes_equal": True,
                            "ok": bool(rel.max() <= 1e-4),
                            "what": "mean image of 2 batches per rank reduced by rb200_context_reduce_present vs the same batches "
                                    "rendered by one GPU (sum mode): identical samples, fp32 summation order differs"}
            assert parity_check["ok"], parity_check
        dist.barrier()

    def batch_index(i):       # rank r renders batches r, r + N, r + 2N, ... (behind the two of the parity check)
        return rank + (i + (2 if world > 1 else 0)) * world

    K, Wm = args.steps, args.warmup
    # The library traces the batches behind the one asked for speculatively (engines x lanes, context.cuh); its pipeline
    # is full from each engine's third call on. These untimed calls precede the W warm-up steps so that warm-up and
    # timed steps all run in the steady state a long render is in; they are ordinary batches of the same sequence.
    engines, lanes, _ = r.engine_config()
    fill = 3 * engines if lanes > 1 else engines
    _bi = batch_index

    def batch_index(i):
        return _bi(i + fill)
    with torch.cuda.stream(stream):
        for i in range(-fill, 0):
            r.render_batch(wl.push_constants(batch_index(i)))
        for i in range(Wm):
            r.render_batch(wl.push_constants(batch_index(i)))
        total_batches += (fill + Wm) * world
        # the warm-up also presents one frame (NCCL reduce, resolve, bloom + tonemap, read-back), result discarded
        if world > 1:
            r.reduce_present(total_batches)
        else:
            r.postprocess()
        if rank == 0:
            r.read_ldr()
        r.synchronize()
        _, cum0 = r.stats()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for i in range(K):
            r.render_batch(wl.push_constants(batch_index(Wm + i)))
        total_batches += K * world
        if world > 1:
            # the one collective of the path, issued by the library (ncclReduce from C++): every rank's SUM image (a
            # stream-ordered snapshot, so the accumulation images stay per-rank sums) reduced to rank 0, which resolves
            # and post-processes the reduced copy
            r.reduce_present(total_batches)
        ev1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = ev0.elapsed_time(ev1)
        clocks = sampler.stop() if rank == 0 else None
        _, cum1 = r.stats()
    rays = (cum1["extendRays"] - cum0["extendRays"]) + (cum1["shadowRays"] - cum0["shadowRays"])
    launches = cum1["kernelLaunches"] - cum0["kernelLaunches"]
    if world > 1:
        t = torch.tensor([ms], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        c = torch.tensor([rays, launches], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        rays, launches = int(c[0].item()), int(c[1].item())
    value = rays / (ms * 1e-3) / 1e6

    # ---- end to end through the C ABI with host buffers: push constants in, bloom + tonemap, RGBA8 frame out ----
    # Single GPU: a software pipeline as deep as the context's lanes (rb200_pipeline_depth), as a display loop with
    # that many frames in flight runs it — frame i's bloom + tonemap + device->host copy are queued behind batch i and
    # the host only blocks on frame i-depth+1 before queueing batch i+1, so the thin tails of the batches in flight
    # overlap the head of the next; every step still copies its inputs in and reads one frame out, and the last frames
    # are drained inside the timed region.
    depth = r.pipeline_depth()
    frames = [r.pinned_frame() for _ in range(depth)]
    ldr_host = frames[0]
    with torch.cuda.stream(stream):
        r.synchronize()
        _, c0 = r.stats()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(K):
            r.render_batch(wl.push_constants(batch_index(Wm + K + i)))
            if world > 1:
                # one presented frame per step, all inside the library: snapshot of this rank's SUM image behind batch i
                # (stream-ordered, no host wait), ONE ncclReduce to rank 0, which resolves + blooms + tonemaps the reduced
                # copy and reads the frame back asynchronously. The accumulation images are never touched, so the
                # engines keep tracing underneath.
                total_batches += world
                r.reduce_present(total_batches)
                if rank == 0:
                    r.wait_ldr(depth - 1)
                    r.read_ldr_async(frames[i % depth])
            else:
                r.postprocess()
                r.wait_ldr(depth - 1)                     # frame i-depth is on the host now: its buffer is free again
                r.read_ldr_async(frames[i % depth])       # frame i follows batch i on the device
        r.wait_ldr()
        e1.record(stream)
        torch.cuda.synchronize()
        ms_e2e = e0.elapsed_time(e1)
        _, c1 = r.stats()
    rays_e2e = (c1["extendRays"] - c0["extendRays"]) + (c1["shadowRays"] - c0["shadowRays"])
    if world > 1:
        t = torch.tensor([ms_e2e], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
        c = torch.tensor([rays_e2e], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        rays_e2e = int(c[0].item())
    e2e = {"value": rays_e2e / (ms_e2e * 1e-3) / 1e6, "unit": "Mrays/s",
           "h2d_bytes_per_step": C.sizeof(rb.abi.RtPushConsts) + C.sizeof(rb.abi.BloomPushConsts) + C.sizeof(rb.abi.TonemappingPushConsts),
           "d2h_bytes_per_step": int(ldr_host.nbytes), "ms_per_step": ms_e2e / K,
           "note": ("per step and rank: rb200_render_batch(host push constants) + rb200_context_reduce_present (snapshot of the SUM "
                    "image, one ncclReduce to rank 0 issued by the library; rank 0: resolve + bloom + tonemap of the reduced copy) "
                    "+ RGBA8 frame to pinned host memory "
                    if world > 1 else
                    "per step: rb200_render_batch(host push constants) + rb200_postprocess + RGBA8 frame to pinned host memory ") +
                   "(rb200_read_ldr_async, %d frames in flight = rb200_pipeline_depth, drained inside the timed region); " % depth +
                   "scene upload + BVH build happen once (scene_create_s)"}
    r.close()

    metric = METRIC if args.config == "c3" else "Mrays/s (primary+bounce+shadow), " + args.config.upper()
    out = {"metric": metric, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": Wm,
           "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "config": config_dict(wl, config=args.config, extra={"parallelism": f"sample-split x{world}", "nee": not args.no_nee,
                                                           "engines": engines, "lanes_per_engine": lanes,
                                                           "pipeline_fill_steps_before_warmup": fill}),
           "spp_per_s": SPP * K * world / (ms * 1e-3), "rays_per_step": rays / K / world,
           "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
           "bvh": {k: bvh[k] for k in ("numTriangles", "numWideNodes", "maxDepth", "nodeBytes", "triangleBytes", "buildMs")},
           "scene_create_s": scene_create_s}

    if parity_check is not None:
        out["parity_check"] = parity_check
    if rank == 0 and world == 1 and not args.no_roofline:
        out["roofline"] = roofline(rb, wl, local_rank, stream, nee_flag)
    if rank == 0 and world == 1 and not args.no_nee and not args.no_as_shipped:
        out["as_shipped"] = as_shipped(rb, wl, local_rank, stream, K)
        out["null_shadow_rays_skipped"] = null_shadow_skip(rb, wl, local_rank, stream, K)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(rb, wl)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


def as_shipped(rb, wl, device, stream, steps):
    """The same workload with next-event estimation off — the estimator the reference's shipped binaries run
    (raytrace.rgen.glsl:145-146, skipNEE hard-coded) and the one pinned to its compiled SPIR-V: device-timed Mrays/s."""
    import torch
    r = rb.Renderer(wl.width, wl.height, wl.tables, flags=0, device=device, stream=stream.cuda_stream)
    engines, lanes, _ = r.engine_config()
    fill = 3 * engines if lanes > 1 else engines
    with torch.cuda.stream(stream):
        for i in range(fill + 3):
            r.render_batch(wl.push_constants(i))
        r.synchronize()
        _, c0 = r.stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            r.render_batch(wl.push_constants(fill + 3 + i))
        e1.record(stream)
        torch.cuda.synchronize()
        _, c1 = r.stats()
    ms = e0.elapsed_time(e1)
    r.close()
    rays = c1["extendRays"] - c0["extendRays"]
    return {"value": rays / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": ms / steps, "rays_per_step": rays / steps,
            "spp_per_s": SPP * steps / (ms * 1e-3), "nee": False,
            "note": "flags = 0: no shadow rays, emission picked up by BRDF sampling only; same scene, camera, sample counts"}


def null_shadow_skip(rb, wl, device, stream, steps):
    """The same workload with RB200_FLAG_SKIP_NULL_SHADOW_RAYS (opt-in): shadow rays whose contribution is exactly zero before
    the visibility test are answered without a traversal (bit-identical images, tests/test_gpu_parity2.py). `value` counts
    only the rays that were traversed, so it can fall while spp/s rises."""
    import torch
    r = rb.Renderer(wl.width, wl.height, wl.tables, flags=rb.RB200_FLAG_NEE | rb.RB200_FLAG_SKIP_NULL_SHADOW_RAYS, device=device,
                    stream=stream.cuda_stream)
    engines, lanes, _ = r.engine_config()
    fill = 3 * engines if lanes > 1 else engines
    with torch.cuda.stream(stream):
        for i in range(fill + 3):
            r.render_batch(wl.push_constants(i))
        r.synchronize()
        _, c0 = r.stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            r.render_batch(wl.push_constants(fill + 3 + i))
        e1.record(stream)
        torch.cuda.synchronize()
        _, c1 = r.stats()
    ms = e0.elapsed_time(e1)
    r.close()
    d = {k: c1[k] - c0[k] for k in ("extendRays", "shadowRays", "shadowRaysSkipped")}
    rays = d["extendRays"] + d["shadowRays"] - d["shadowRaysSkipped"]
    return {"value": rays / (ms * 1e-3) / 1e6, "unit": "Mrays/s (traversed rays only)", "ms_per_step": ms / steps,
            "rays_traversed_per_step": rays / steps, "shadow_rays_skipped_per_step": d["shadowRaysSkipped"] / steps,
            "share_of_shadow_rays_skipped": d["shadowRaysSkipped"] / max(1, d["shadowRays"]),
            "spp_per_s": SPP * steps / (ms * 1e-3),
            "note": "opt-in flag; the headline value above traverses every shadow ray the reference's estimator casts"}


def roofline(rb, wl, device, stream, nee_flag):
    """Dominant kernel = k_extend (closest-hit traversal). Algorithmic bytes per ray:
    80 * N_node + 48 * N_tri + 32 (ray in) + 16 (hit out), N_node / N_tri measured by a counting build of the same kernel
    on the same batch; time = sum of the kernel's launch durations (CUDA events on the launching stream)."""
    peak, peak_src = measured_peak_hbm()
    res = {"bound": "hbm", "kernel": "k_extend", "unit": "GB/s", "peak": peak, "peak_source": peak_src}
    pc = wl.push_constants(1000)
    # timing pass (events around every kernel, launches serialised; no counters). Consecutive batches, so that the
    # engine is in its steady state — every launch of the last call carries the rays of all batches in flight — and
    # the launches timed are the launches the bench loop above issues
    rt = rb.Renderer(wl.width, wl.height, wl.tables, flags=nee_flag | rb.RB200_FLAG_TIME_KERNELS, device=device,
                     stream=stream.cuda_stream)
    engines, lanes, _ = rt.engine_config()
    for b in range(3 * engines + 1):
        rt.render_batch(wl.push_constants(1000 + b))
    kt = rt.kernel_times()
    last_t, _ = rt.stats()
    # post-processing of the frame just rendered: blur X + blur Y + fused combine / tonemap, CUDA events on the stream
    import torch
    with torch.cuda.stream(stream):
        for _ in range(3):
            rt.postprocess()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        for _ in range(20):
            rt.postprocess()
        p1.record(stream)
        torch.cuda.synchronize()
    post_ms = p0.elapsed_time(p1) / 20
    rt.close()
    # counting pass
    rc = rb.Renderer(wl.width, wl.height, wl.tables, flags=nee_flag | rb.RB200_FLAG_COUNT_BVH, device=device,
                     stream=stream.cuda_stream)
    rc.render_batch(pc)
    last_c, _ = rc.stats()
    # the bandwidth that bounds an L2-resident traversal is the L2 gather bandwidth; measure it on a
    # table the size of this BVH (independent random reads of whole 80-byte records, warm L2)
    info = rc.bvh_info()
    table_bytes = max(int(info["nodeBytes"] + info["triangleBytes"]), 1 << 20)     # (tiny scenes: the micro-benchmark needs >= 1024 records)
    l2_gather = rc.measure_gather(table_bytes, 80)
    rc.close()
    # algorithmic bytes of a record = what it carries (80-byte wide node, 48-byte triangle), not its stride in memory:
    # with RB_WIDE_LOADS the library pads records to 96 / 64 bytes for 256-bit loads, and padding is not information
    node_bytes = 80.0
    node_stride = info["nodeBytes"] / max(1, info["numWideNodes"])
    tri_stride = info["triangleBytes"] / max(1, info["numTriangles"])
    n_node = (last_c["nodeVisits"] - last_c["shadowNodeVisits"]) / max(1, last_c["extendRays"])
    n_tri = (last_c["triTests"] - last_c["shadowTriTests"]) / max(1, last_c["extendRays"])
    n_node_sh = last_c["shadowNodeVisits"] / max(1, last_c["shadowRays"])
    n_tri_sh = last_c["shadowTriTests"] / max(1, last_c["shadowRays"])
    bytes_per_ray = node_bytes * n_node + 48.0 * n_tri + 32.0 + 16.0
    bytes_per_shadow_ray = node_bytes * n_node_sh + 48.0 * n_tri_sh + 32.0 + 48.0 + 32.0   # ray record, 3 result records, radiance r/w
    ext_bytes = bytes_per_ray * kt["extendRays"]
    achieved = ext_bytes / (kt["extendMs"] * 1e-3) / 1e9
    total_ms = kt["generateMs"] + kt["extendMs"] + sum(kt["shadeMs"]) + kt["shadowMs"] + kt["finishMs"]
    traffic, traffic_source = None, None
    try:   # per-launch DRAM bytes of the dominant kernel: a STATIC figure from the committed ncu capture, not measured in this run
        tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        traffic, traffic_source = tj["k_extend_dram_bytes_per_launch"], "static ncu capture " + tj.get("source", "profiles/roofline_traffic.json")
    except Exception:
        pass

    def kernel_entry(items, unit_bytes, ms, what):
        if not ms or ms <= 0 or not items:
            return None
        ach = unit_bytes * items / (ms * 1e-3) / 1e9
        return {"bound": "hbm", "units_per_step": items, "bytes_per_unit": unit_bytes, "ms_per_step": ms, "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak, "what": what}
    px = wl.width * wl.height
    kernels = {
        "extend": kernel_entry(kt["extendRays"], bytes_per_ray, kt["extendMs"], "node_bytes * nodes/ray + 48 * tris/ray + 32 (ray in) + 16 (hit out), extend rays only"),
        "shadow": kernel_entry(kt["shadowRays"], bytes_per_shadow_ray, kt["shadowMs"], "same for any-hit rays: + 32 (ray record) + 48 (result records) + 32 (radiance read + write)"),
        "shade_disney": kernel_entry(kt["shadeItems"][3], 620.0, kt["shadeMs"][3], "SURVEY 8d: 0.62 KB gathered / written per shaded hit (with the shadow record)"),
        "shade_lambertian": kernel_entry(kt["shadeItems"][0], 620.0, kt["shadeMs"][0], "SURVEY 8d: 0.62 KB per shaded hit"),
        "shade_metal": kernel_entry(kt["shadeItems"][1], 580.0, kt["shadeMs"][1], "SURVEY 8d: 0.58 KB per shaded hit (no shadow record)"),
        "shade_dielectric": kernel_entry(kt["shadeItems"][2], 580.0, kt["shadeMs"][2], "SURVEY 8d: 0.58 KB per shaded hit (no shadow record)"),
        "miss": kernel_entry(kt["shadeItems"][4], 224.0, kt["shadeMs"][4], "96 B of path state in, 128 B out (the slot's next camera path)"),
        "post": kernel_entry(px, 132.0, post_ms, "SURVEY 8d: 132 B per pixel with combine + tonemap fused; units = pixels, per call"),
    }
    res.update({"achieved": achieved, "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_source,
                "kernels": {k: v for k, v in kernels.items() if v}, "post_ms_per_call": post_ms,
                "bytes_per_ray": bytes_per_ray, "nodes_per_ray": n_node, "tris_per_ray": n_tri, "node_bytes": node_bytes,
                "node_stride_bytes": node_stride, "triangle_stride_bytes": tri_stride,
                "shadow": {"bytes_per_ray": bytes_per_shadow_ray, "nodes_per_ray": n_node_sh, "tris_per_ray": n_tri_sh,
                           "rays_per_step": kt["shadowRays"], "ms_per_step": kt["shadowMs"], "launches": kt["shadowLaunches"],
                           "achieved": (bytes_per_shadow_ray * kt["shadowRays"] / (kt["shadowMs"] * 1e-3) / 1e9) if kt["shadowMs"] > 0 else None,
                           "frac": (bytes_per_shadow_ray * kt["shadowRays"] / (kt["shadowMs"] * 1e-3) / 1e9 / peak) if kt["shadowMs"] > 0 else None},
                "engine": {"engines": engines, "lanes": lanes},
                "extend_rays_per_batch": last_t["extendRays"], "extend_rays_per_step": kt["extendRays"],
                "extend_launches": kt["extendLaunches"],
                "extend_ms_per_batch": kt["extendMs"], "avg_launch_ms": kt["extendMs"] / max(1, kt["extendLaunches"]),
                "extend_mrays_s": kt["extendRays"] / (kt["extendMs"] * 1e-3) / 1e6,
                "shadow_mrays_s": (kt["shadowRays"] / (kt["shadowMs"] * 1e-3) / 1e6) if kt["shadowMs"] > 0 else None,
                "shade_items_per_step": {"lambertian": kt["shadeItems"][0], "metal": kt["shadeItems"][1],
                                         "dielectric": kt["shadeItems"][2], "disney": kt["shadeItems"][3],
                                         "miss": kt["shadeItems"][4], "finish": kt["finishItems"]},
                "kernel_ms_per_batch": {"generate": kt["generateMs"], "extend": kt["extendMs"], "shade_lambertian": kt["shadeMs"][0],
                                        "shade_metal": kt["shadeMs"][1], "shade_dielectric": kt["shadeMs"][2],
                                        "shade_disney": kt["shadeMs"][3], "miss": kt["shadeMs"][4], "shadow": kt["shadowMs"],
                                        "finish": kt["finishMs"]},
                "extend_share_of_step": kt["extendMs"] / total_ms if total_ms > 0 else None,
                # the same figure over the launches whose ray queue held >= half the pixels ("full waves", where the
                # kernel is throughput-bound; the thin waves behind them are bound by the latency of their longest ray
                # and run beside other lanes' full waves)
                "full_waves": ({"launches": kt["extendFullLaunches"], "rays": kt["extendFullRays"], "ms": kt["extendFullMs"],
                                "extend_mrays_s": kt["extendFullRays"] / (kt["extendFullMs"] * 1e-3) / 1e6,
                                "achieved": bytes_per_ray * kt["extendFullRays"] / (kt["extendFullMs"] * 1e-3) / 1e9,
                                "frac": bytes_per_ray * kt["extendFullRays"] / (kt["extendFullMs"] * 1e-3) / 1e9 / peak,
                                "share_of_extend_rays": kt["extendFullRays"] / max(1, kt["extendRays"]),
                                "shadow_mrays_s": (kt["shadowFullRays"] / (kt["shadowFullMs"] * 1e-3) / 1e6) if kt["shadowFullMs"] > 0 else None}
                               if kt["extendFullMs"] > 0 else None),
                "l2_gather": {"peak": l2_gather, "unit": "GB/s", "frac": achieved / l2_gather, "table_bytes": table_bytes,
                              "how": "rb200_measure_gather: independent random 80-byte record reads from a warm table of the "
                                     "BVH's size, measured in this run"},
                "note": "BVH traversal is an L2-resident pointer chase: the fraction is algorithmic bytes over the measured "
                        "HBM copy bandwidth, the only bandwidth peak MEASURED_PEAKS.json provides; l2_gather gives the same "
                        "numerator over the L2 gather bandwidth measured in this run (an upper bound: a traversal's next "
                        "address depends on the node it just read)"})
    return res


if __name__ == "__main__":
    sys.exit(main())
