"""reina-vk_b200 — B200-native path-tracing core for Reina (AlexanderJCS/reina-vk).

The product is csrc/librb200.so: hand-written sm_100a CUDA kernels (LBVH -> 8-wide BVH builder, wavefront
extend / shade / shadow tracer, Disney / Lambertian / metal / dielectric shading, accumulate, bloom, tonemap) behind
the C ABI of include/reina_b200.h. This package is the host-side mirror of the reference's scene layer plus a thin
ctypes binding; it contains no CPU rendering path and refuses to work without the CUDA library.

Import with importlib (the directory name carries the reference's hyphen):
    rb = importlib.import_module("reina-vk_b200")
"""
import ctypes as C

import numpy as np

from . import abi, camera, configs, gltf, meshes, scene
from .abi import (RB200_FLAG_ACCUM_SUM, RB200_FLAG_COUNT_BVH, RB200_FLAG_GROUP_TILES, RB200_FLAG_NEE, RB200_FLAG_TIME_KERNELS,
                  RB200_FLAG_SKIP_NULL_SHADOW_RAYS, RB200_FLAG_TWO_LEVEL,
                  BloomPushConsts, RB200Error, RtPushConsts, TonemappingPushConsts, load_library)
from .scene import Material, ModelData, Scene, SceneTables

__all__ = ["Renderer", "Group", "comm_unique_id", "RB200_FLAG_GROUP_TILES", "RB200_FLAG_TWO_LEVEL", "RB200_FLAG_SKIP_NULL_SHADOW_RAYS", "Material", "ModelData", "Scene", "SceneTables", "RtPushConsts", "BloomPushConsts",
           "TonemappingPushConsts", "RB200_FLAG_NEE", "RB200_FLAG_ACCUM_SUM", "RB200_FLAG_COUNT_BVH", "RB200_FLAG_TIME_KERNELS",
           "RB200Error",
           "abi", "camera", "configs", "gltf", "meshes", "scene", "load_library"]


class _DevicePtr:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can wrap it without a copy."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class Renderer:
    """One render context + one scene: the calls Reina's frame loop makes (src/Reina.cpp:296-392) —
    traceRays() -> applyBloom() -> applyTonemapping() -> save() — against the C ABI."""

    def __init__(self, width, height, tables, flags=0, device=0, stream=None):
        self.lib = load_library()
        self.width, self.height, self.flags = width, height, flags
        self._pinned = []
        self._ctx = C.c_void_p()
        abi.check(self.lib, self.lib.rb200_context_create(width, height, device, flags, C.byref(self._ctx)))
        if stream is not None:
            abi.check(self.lib, self.lib.rb200_context_set_stream(self._ctx, C.c_void_p(stream)))
        self._scene = C.c_void_p()
        self.tables = tables
        desc = tables.desc()
        if (flags & RB200_FLAG_NEE) and tables.numEmissive > 1:
            import warnings
            warnings.warn("%d emissive instances: the reference's light sampling (nee.h.glsl:97-105) addresses the triangles of "
                          "every emitter after the first through the concatenated triangle CDF, i.e. beyond the emitter's own "
                          "triangles; reproduced as is (DESIGN.md 2)" % tables.numEmissive)
        try:
            abi.check(self.lib, self.lib.rb200_scene_create(self._ctx, C.byref(desc), C.byref(self._scene)))
        except Exception:
            self.lib.rb200_context_destroy(self._ctx)
            self._ctx = None
            raise

    def set_tiles(self, tile_rank, tile_count, tile_size=32):
        """Interleaved-tile partition (latency mode over several GPUs): trace only this rank's tiles from now on."""
        abi.check(self.lib, self.lib.rb200_context_set_tiles(self._ctx, tile_rank, tile_count, tile_size))

    # -- the hot path ------------------------------------------------------------------------------------
    def render_batch(self, pc):
        abi.check(self.lib, self.lib.rb200_render_batch(self._ctx, self._scene, C.byref(pc)))

    def resolve_sum(self, num_batches):
        abi.check(self.lib, self.lib.rb200_resolve_sum(self._ctx, num_batches))

    def postprocess(self, bloom=None, tonemap=None):
        d = camera.DEFAULTS
        bloom = bloom or BloomPushConsts(d["bloom_radius"], d["bloom_threshold"], d["bloom_intensity"])
        tonemap = tonemap or TonemappingPushConsts(d["exposure"])
        abi.check(self.lib, self.lib.rb200_postprocess(self._ctx, C.byref(bloom), C.byref(tonemap)))

    def present_sum(self, num_batches, device_ptr=None, bloom=None, tonemap=None):
        """Resolve a SUM image (a device pointer, e.g. the NCCL-reduced copy of every rank's image; None = this context's
        own) into a staging image and bloom + tonemap it into the RGBA8 frame; the accumulation image is untouched."""
        d = camera.DEFAULTS
        bloom = bloom or BloomPushConsts(d["bloom_radius"], d["bloom_threshold"], d["bloom_intensity"])
        tonemap = tonemap or TonemappingPushConsts(d["exposure"])
        abi.check(self.lib, self.lib.rb200_present_sum(self._ctx, C.c_void_p(device_ptr) if device_ptr else None, num_batches,
                                                       C.byref(bloom), C.byref(tonemap)))

    def read_hdr(self, out=None):
        out = np.empty((self.height, self.width, 4), np.float32) if out is None else out
        abi.check(self.lib, self.lib.rb200_read_hdr(self._ctx, out.ctypes.data_as(C.c_void_p)))
        return out

    def read_ldr(self, out=None):
        out = np.empty((self.height, self.width, 4), np.uint8) if out is None else out
        abi.check(self.lib, self.lib.rb200_read_ldr(self._ctx, out.ctypes.data_as(C.c_void_p)))
        return out

    def pinned_frame(self):
        """A page-locked (H, W, 4) uint8 array from rb200_host_alloc, for read_ldr_async; freed by close()."""
        p = C.c_void_p()
        n = self.height * self.width * 4
        abi.check(self.lib, self.lib.rb200_host_alloc(n, C.byref(p)))
        self._pinned.append(p)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(self.height, self.width, 4))

    def read_ldr_async(self, out):
        """Enqueue the frame copy into `out` (use pinned_frame()) and return; wait_ldr() blocks until it landed."""
        assert out.dtype == np.uint8 and out.shape == (self.height, self.width, 4) and out.flags.c_contiguous
        abi.check(self.lib, self.lib.rb200_read_ldr_async(self._ctx, out.ctypes.data_as(C.c_void_p)))

    def wait_ldr(self, max_pending=0):
        """Block until at most `max_pending` read_ldr_async copies (the most recent ones) are still in flight."""
        abi.check(self.lib, self.lib.rb200_wait_ldr_pending(self._ctx, max_pending))

    def pipeline_depth(self):
        """Frames a display loop should keep outstanding so that the host never drains the device."""
        return int(self.lib.rb200_pipeline_depth())

    def write_hdr(self, img):
        img = np.ascontiguousarray(img, np.float32)
        assert img.shape == (self.height, self.width, 4)
        abi.check(self.lib, self.lib.rb200_write_hdr(self._ctx, img.ctypes.data_as(C.c_void_p)))

    def hdr_device_array(self):
        p = C.c_void_p()
        abi.check(self.lib, self.lib.rb200_hdr_device_ptr(self._ctx, C.byref(p)))
        return _DevicePtr(p.value, (self.height, self.width, 4))

    # -- parity / measurement helpers --------------------------------------------------------------------
    def trace_primary(self, pc):
        hits = np.empty(self.width * self.height, dtype=np.dtype(abi.PrimaryHit))
        abi.check(self.lib, self.lib.rb200_trace_primary(self._ctx, self._scene, C.byref(pc),
                                                         hits.ctypes.data_as(C.c_void_p)))
        return hits

    def trace_rays(self, origins, directions, tmax, any_hit=False):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(np.broadcast_to(np.asarray(tmax, np.float32), (o.shape[0],)), np.float32)
        hits = np.empty(o.shape[0], dtype=np.dtype(abi.PrimaryHit))
        abi.check(self.lib, self.lib.rb200_trace_rays(self._ctx, self._scene, o.shape[0], o.ctypes.data_as(C.c_void_p),
                                                      d.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p),
                                                      1 if any_hit else 0, hits.ctypes.data_as(C.c_void_p)))
        return hits

    def shade_hits(self, origins, directions, rng_states, inside, acc_dist):
        """One closest-hit traversal + one material-shader invocation per ray (rb200_shade_hits): structured array with the
        payload the shader leaves (color, albedo, origin, direction, emission, normal, pdf, accumulatedDistance, rngState,
        flags: 1 hit | 2 skip | 4 insideDielectric, material)."""
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, np.float32).reshape(-1, 3)
        n = o.shape[0]
        r = np.ascontiguousarray(rng_states, np.uint32).reshape(n)
        i = np.ascontiguousarray(inside, np.uint32).reshape(n)
        a = np.ascontiguousarray(acc_dist, np.float32).reshape(n)
        out = np.empty(n, dtype=np.dtype(abi.ShadeResult))
        p = lambda x: x.ctypes.data_as(C.c_void_p)
        abi.check(self.lib, self.lib.rb200_shade_hits(self._ctx, self._scene, n, p(o), p(d), p(r), p(i), p(a), p(out)))
        return out

    def bench_trace(self, origins, directions, tmax, any_hit=False, reps=10):
        """(ms per launch, checksum of the hits) of the traversal kernel alone on these rays (rb200_bench_trace)."""
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(np.broadcast_to(np.asarray(tmax, np.float32), (o.shape[0],)), np.float32)
        ms, ck = C.c_float(), C.c_uint64()
        abi.check(self.lib, self.lib.rb200_bench_trace(self._ctx, self._scene, o.shape[0], o.ctypes.data_as(C.c_void_p),
                                                       d.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p),
                                                       1 if any_hit else 0, reps, C.byref(ms), C.byref(ck)))
        return float(ms.value), int(ck.value)

    def engine_config(self):
        """(engines, lanes per engine, speculative batches discarded so far): see rb200_engine_config."""
        e, l, w = C.c_uint32(), C.c_uint32(), C.c_uint64()
        abi.check(self.lib, self.lib.rb200_engine_config(self._ctx, C.byref(e), C.byref(l), C.byref(w)))
        return int(e.value), int(l.value), int(w.value)

    def bvh_info(self):
        info = abi.BvhInfo()
        abi.check(self.lib, self.lib.rb200_scene_bvh_info(self._scene, C.byref(info)))
        return info.as_dict()

    def stats(self):
        last, cum = abi.Stats(), abi.Stats()
        abi.check(self.lib, self.lib.rb200_get_stats(self._ctx, C.byref(last), C.byref(cum)))
        return last.as_dict(), cum.as_dict()

    def kernel_times(self):
        kt = abi.KernelTimes()
        abi.check(self.lib, self.lib.rb200_get_kernel_times(self._ctx, C.byref(kt)))
        return kt.as_dict()

    def measure_gather(self, table_bytes, record_bytes=80):
        """GB/s of independent random record reads from an L2-resident table (the traversal roofline)."""
        g = C.c_float()
        abi.check(self.lib, self.lib.rb200_measure_gather(self._ctx, int(table_bytes), int(record_bytes), C.byref(g)))
        return float(g.value)

    def synchronize(self):
        abi.check(self.lib, self.lib.rb200_synchronize(self._ctx))

    # -- one process per GPU: the library's own NCCL communicator (rb200_context_comm_*) -------------------
    def comm_init(self, unique_id, rank, nranks):
        """ncclCommInitRank on this context's device; `unique_id` = the 128 bytes of comm_unique_id() from rank 0."""
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        abi.check(self.lib, self.lib.rb200_context_comm_init(self._ctx, C.cast(buf, C.c_void_p), rank, nranks))
        self._has_comm = True

    def reduce_present(self, total_batches, bloom=None, tonemap=None):
        """Snapshot of this rank's image -> one ncclReduce to rank 0 -> (rank 0) resolve + bloom + tonemap."""
        d = camera.DEFAULTS
        bloom = bloom or BloomPushConsts(d["bloom_radius"], d["bloom_threshold"], d["bloom_intensity"])
        tonemap = tonemap or TonemappingPushConsts(d["exposure"])
        abi.check(self.lib, self.lib.rb200_context_reduce_present(self._ctx, total_batches, C.byref(bloom), C.byref(tonemap)))

    def reduced_device_array(self):
        p = C.c_void_p()
        abi.check(self.lib, self.lib.rb200_context_reduced_device_ptr(self._ctx, C.byref(p)))
        return _DevicePtr(p.value, (self.height, self.width, 4))

    def close(self):
        if getattr(self, "_has_comm", False) and getattr(self, "_ctx", None):
            self.lib.rb200_context_comm_destroy(self._ctx)
            self._has_comm = False
        if getattr(self, "_scene", None):
            self.lib.rb200_scene_destroy(self._scene)
            self._scene = None
        if getattr(self, "_ctx", None):
            self.lib.rb200_context_destroy(self._ctx)
            self._ctx = None
        for p in getattr(self, "_pinned", []):
            self.lib.rb200_host_free(p)
        self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def comm_unique_id():
    """128 bytes from ncclGetUniqueId (rank 0 creates them and hands them to the other ranks)."""
    lib = load_library()
    buf = (C.c_char * 128)()
    abi.check(lib, lib.rb200_comm_unique_id(C.cast(buf, C.c_void_p)))
    return bytes(buf)


class Group:
    """Several GPUs in one process (rb200_group_*): a context and a host thread per device, the BVH replicated, one
    ncclReduce of the accumulation images per presented frame. Sample split by default, interleaved tiles with
    tiles=True."""

    def __init__(self, width, height, tables, devices, flags=0, tiles=False):
        self.lib = load_library()
        self.width, self.height = width, height
        devs = (C.c_int * len(devices))(*devices)
        self._g = C.c_void_p()
        abi.check(self.lib, self.lib.rb200_group_create(width, height, devs, len(devices),
                                                        flags | (RB200_FLAG_GROUP_TILES if tiles else 0), C.byref(self._g)))
        self._scene = C.c_void_p()
        self.tables = tables
        desc = tables.desc()
        try:
            abi.check(self.lib, self.lib.rb200_group_scene_create(self._g, C.byref(desc), C.byref(self._scene)))
        except Exception:
            self.lib.rb200_group_destroy(self._g)
            self._g = None
            raise

    def size(self):
        return int(self.lib.rb200_group_size(self._g))

    def uses_peer_stores(self):
        """Latency mode without a collective: the devices store their pixels into device 0's image over NVLink."""
        return bool(self.lib.rb200_group_uses_peer_stores(self._g))

    def render_batches(self, pc, first_batch, batches_per_device):
        abi.check(self.lib, self.lib.rb200_group_render_batches(self._g, self._scene, C.byref(pc), first_batch, batches_per_device))

    def present(self, bloom=None, tonemap=None):
        d = camera.DEFAULTS
        bloom = bloom or BloomPushConsts(d["bloom_radius"], d["bloom_threshold"], d["bloom_intensity"])
        tonemap = tonemap or TonemappingPushConsts(d["exposure"])
        abi.check(self.lib, self.lib.rb200_group_present(self._g, C.byref(bloom), C.byref(tonemap)))

    def read_ldr(self):
        out = np.empty((self.height, self.width, 4), np.uint8)
        abi.check(self.lib, self.lib.rb200_group_read_ldr(self._g, out.ctypes.data_as(C.c_void_p)))
        return out

    def read_hdr(self):
        out = np.empty((self.height, self.width, 4), np.float32)
        abi.check(self.lib, self.lib.rb200_group_read_hdr(self._g, out.ctypes.data_as(C.c_void_p)))
        return out

    def bvh_info(self, index=0):
        info = abi.BvhInfo()
        abi.check(self.lib, self.lib.rb200_group_scene_bvh_info(self._scene, index, C.byref(info)))
        return info.as_dict()

    def stats(self):
        cum = abi.Stats()
        abi.check(self.lib, self.lib.rb200_group_get_stats(self._g, C.byref(cum)))
        return cum.as_dict()

    def synchronize(self):
        abi.check(self.lib, self.lib.rb200_group_synchronize(self._g))

    def close(self):
        if getattr(self, "_scene", None):
            self.lib.rb200_group_scene_destroy(self._scene)
            self._scene = None
        if getattr(self, "_g", None):
            self.lib.rb200_group_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
