"""ctypes binding of oracle/liboracle.so — the CPU restatement used as the parity checker.
Test infrastructure only: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs."""
import ctypes as C
import importlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
rb = importlib.import_module("reina-vk_b200")
abi = rb.abi

ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")


class OracleCounters(C.Structure):
    _fields_ = [("extendRays", C.c_uint64), ("shadowRays", C.c_uint64), ("paths", C.c_uint64)]


_lib = None


def build_oracle():
    subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        _lib = C.CDLL(ORACLE_SO)
        _lib.oracle_scene_create.argtypes = [C.POINTER(abi.SceneDesc), C.c_int, C.POINTER(C.c_void_p)]
        _lib.oracle_scene_destroy.argtypes = [C.c_void_p]
        _lib.oracle_scene_set_two_level.argtypes = [C.c_void_p, C.c_int]
        _lib.oracle_scene_num_triangles.argtypes = [C.c_void_p]
        _lib.oracle_scene_num_triangles.restype = C.c_uint32
        _lib.oracle_render_batch.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(abi.RtPushConsts),
                                             C.c_void_p, C.c_int, C.POINTER(OracleCounters)]
        _lib.oracle_render_batch_tiles.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(abi.RtPushConsts),
                                                   C.c_void_p, C.c_int, C.POINTER(OracleCounters), C.c_uint32, C.c_uint32,
                                                   C.c_uint32]
        _lib.oracle_trace_primary.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(abi.RtPushConsts), C.c_void_p,
                                              C.c_int, C.c_int]
        _lib.oracle_trace_rays.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                           C.c_void_p, C.c_int, C.c_int]
        _lib.oracle_postprocess.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(abi.BloomPushConsts),
                                            C.POINTER(abi.TonemappingPushConsts), C.c_void_p, C.c_void_p, C.c_int]
        _lib.oracle_kat_random.argtypes = [C.POINTER(C.c_uint32)]
        _lib.oracle_kat_random.restype = C.c_float
        _lib.oracle_kat_power_heuristic.argtypes = [C.c_float, C.c_float]
        _lib.oracle_kat_power_heuristic.restype = C.c_float
        _lib.oracle_kat_offset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oracle_kat_bump.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oracle_kat_trace_main.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32), C.c_int, C.c_float,
                                               C.c_void_p, C.POINTER(C.c_uint32)]
        _lib.oracle_kat_sky.argtypes = [C.c_void_p, C.c_void_p]
        _lib.oracle_kat_bounce.argtypes = [C.c_void_p, C.POINTER(abi.RtPushConsts), C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32), C.c_int,
                                           C.c_float, C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p, C.POINTER(C.c_uint32)]
        _lib.oracle_kat_light_sample.argtypes = [C.c_void_p, C.POINTER(abi.RtPushConsts), C.POINTER(C.c_uint32), C.c_void_p]
        _lib.oracle_kat_direct_light_lambertian.argtypes = [C.c_void_p, C.POINTER(abi.RtPushConsts), C.c_void_p, C.c_void_p, C.c_void_p,
                                                            C.POINTER(C.c_uint32), C.c_void_p]
        _lib.oracle_kat_tonemap.argtypes = [C.c_void_p, C.c_float, C.c_void_p]
        _lib.oracle_kat_starting_ray.argtypes = [C.POINTER(abi.RtPushConsts), C.c_uint32, C.c_uint32, C.c_uint32,
                                                 C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]
        _lib.oracle_rb_math.argtypes = [C.c_int, C.c_uint32, C.c_void_p, C.c_void_p]
        _lib.oracle_rb_random.argtypes = [C.POINTER(C.c_uint32), C.c_uint32, C.c_void_p]
    return _lib


NTHREADS = max(1, os.cpu_count() or 1)


class OracleScene:
    def __init__(self, tables, bvh_threshold=1500):
        self.tables = tables
        self._h = C.c_void_p()
        d = tables.desc()
        rc = lib().oracle_scene_create(C.byref(d), bvh_threshold, C.byref(self._h))
        assert rc == 0

    def set_two_level(self, bvh_threshold=1500):
        """Two-level intersection: rays in object space per instance (Scene.cpp:93-111), see oracle/intersect.cpp."""
        rc = lib().oracle_scene_set_two_level(self._h, bvh_threshold)
        if rc != 0:
            raise RuntimeError("oracle_scene_set_two_level failed: %d" % rc)

    def num_triangles(self):
        return lib().oracle_scene_num_triangles(self._h)

    def render_batch(self, width, height, flags, pc, hdr=None, threads=NTHREADS, tiles=(0, 1, 32)):
        """tiles = (rank, count, tileSize): render only this rank's interleaved tiles, others untouched."""
        if hdr is None:
            hdr = np.zeros((height, width, 4), np.float32)
        cnt = OracleCounters()
        rc = lib().oracle_render_batch_tiles(self._h, width, height, flags, C.byref(pc), hdr.ctypes.data_as(C.c_void_p),
                                             threads, C.byref(cnt), tiles[0], tiles[1], tiles[2])
        assert rc == 0, rc
        return hdr, {"extendRays": cnt.extendRays, "shadowRays": cnt.shadowRays, "paths": cnt.paths}

    def trace_primary(self, width, height, pc, brute=False, threads=NTHREADS):
        hits = np.empty(width * height, dtype=np.dtype(abi.PrimaryHit))
        lib().oracle_trace_primary(self._h, width, height, C.byref(pc), hits.ctypes.data_as(C.c_void_p), threads,
                                   1 if brute else 0)
        return hits

    def trace_rays(self, origins, directions, tmax, any_hit=False, brute=False, threads=NTHREADS):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(np.broadcast_to(np.asarray(tmax, np.float32), (o.shape[0],)), np.float32)
        hits = np.empty(o.shape[0], dtype=np.dtype(abi.PrimaryHit))
        lib().oracle_trace_rays(self._h, o.shape[0], o.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p),
                                t.ctypes.data_as(C.c_void_p), 1 if any_hit else 0, hits.ctypes.data_as(C.c_void_p),
                                threads, 1 if brute else 0)
        return hits

    def close(self):
        if self._h:
            lib().oracle_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def postprocess(hdr, bloom=None, tonemap=None, threads=NTHREADS, want_combined=False):
    d = rb.camera.DEFAULTS
    bloom = bloom or abi.BloomPushConsts(d["bloom_radius"], d["bloom_threshold"], d["bloom_intensity"])
    tonemap = tonemap or abi.TonemappingPushConsts(d["exposure"])
    h, w = hdr.shape[:2]
    hdr = np.ascontiguousarray(hdr, np.float32)
    ldr = np.empty((h, w, 4), np.uint8)
    comb = np.empty((h, w, 4), np.float32) if want_combined else None
    lib().oracle_postprocess(w, h, hdr.ctypes.data_as(C.c_void_p), C.byref(bloom), C.byref(tonemap),
                             ldr.ctypes.data_as(C.c_void_p), comb.ctypes.data_as(C.c_void_p) if want_combined else None,
                             threads)
    return (ldr, comb) if want_combined else ldr


def rb_math(fn, x):
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty_like(x)
    lib().oracle_rb_math(fn, x.size, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p))
    return y
