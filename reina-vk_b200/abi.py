"""ctypes view of include/reina_b200.h — the C ABI of the B200 path-tracing core.

The structures are byte-for-byte the reference's polyglot structs (polyglot/raytrace.h:18-54,
polyglot/bloom.h:4-8, polyglot/tonemapping.h:4-6, src/scene/Instances.h:15-27). The CUDA library
(csrc/librb200.so) is mandatory: there is no CPU fallback, loading fails loudly when it has not been built.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RB200_LIBRARY (developer aid, tools/sweep_variants.sh): another build of the same library, e.g. csrc/variants/<name>.so
LIB_PATH = os.environ.get("RB200_LIBRARY") or os.path.join(_HERE, "csrc", "librb200.so")

RB200_OK = 0
RB200_FLAG_NEE = 1 << 0
RB200_FLAG_ACCUM_SUM = 1 << 1
RB200_FLAG_COUNT_BVH = 1 << 2
RB200_FLAG_TIME_KERNELS = 1 << 3
RB200_FLAG_GROUP_TILES = 1 << 4
RB200_FLAG_TWO_LEVEL = 1 << 5
RB200_FLAG_SKIP_NULL_SHADOW_RAYS = 1 << 6


class InstanceProperties(C.Structure):
    _fields_ = [
        ("indicesOffset", C.c_uint32), ("albedo", C.c_float * 3), ("emission", C.c_float * 3),
        ("tbnsIndicesOffset", C.c_uint32), ("texIndicesOffset", C.c_uint32), ("roughness", C.c_float),
        ("ior", C.c_float), ("interpNormals", C.c_uint32), ("absorption", C.c_float),
        ("textureID", C.c_int32), ("normalMapTexID", C.c_int32), ("bumpMapTexID", C.c_int32),
        ("cullBackface", C.c_uint32), ("anisotropic", C.c_float), ("subsurface", C.c_float),
        ("clearcoatGloss", C.c_float), ("sheenTint", C.c_float * 3), ("specularTint", C.c_float * 3),
        ("metallic", C.c_float), ("clearcoat", C.c_float), ("specularTransmission", C.c_float),
        ("sheen", C.c_float),
    ]


class RtPushConsts(C.Structure):
    _fields_ = [
        ("invView", C.c_float * 16), ("invProjection", C.c_float * 16), ("sampleBatch", C.c_uint32),
        ("totalEmissiveWeight", C.c_float), ("focusDist", C.c_float), ("defocusMultiplier", C.c_float),
        ("directClamp", C.c_float), ("indirectClamp", C.c_float), ("samplesPerPixel", C.c_uint32),
        ("maxBounces", C.c_uint32),
    ]


class BloomPushConsts(C.Structure):
    _fields_ = [("radius", C.c_float), ("threshold", C.c_float), ("intensity", C.c_float)]


class TonemappingPushConsts(C.Structure):
    _fields_ = [("exposure", C.c_float)]


class InstanceData(C.Structure):
    _fields_ = [
        ("transform", C.c_float * 16), ("materialOffset", C.c_uint32), ("cdfRangeStart", C.c_uint32),
        ("cdfRangeEnd", C.c_uint32), ("indexOffset", C.c_uint32), ("emission", C.c_float * 3),
        ("weight", C.c_float), ("area", C.c_float), ("cullBackface", C.c_uint32), ("padding", C.c_float * 2),
    ]


class Instance(C.Structure):
    _fields_ = [
        ("transform", C.c_float * 16), ("instancePropertiesID", C.c_uint32), ("materialIdx", C.c_uint32),
        ("indexOffset", C.c_uint32), ("triangleCount", C.c_uint32),
    ]


class Texture(C.Structure):
    _fields_ = [("rgba8", C.POINTER(C.c_uint8)), ("width", C.c_uint32), ("height", C.c_uint32)]


class SceneDesc(C.Structure):
    _fields_ = [
        ("vertices", C.POINTER(C.c_float)), ("numVertices", C.c_uint32),
        ("indices", C.POINTER(C.c_uint32)), ("numIndices", C.c_uint32),
        ("instanceProperties", C.POINTER(InstanceProperties)), ("numInstanceProperties", C.c_uint32),
        ("tbns", C.POINTER(C.c_float)), ("numTbns", C.c_uint32),
        ("tbnIndices", C.POINTER(C.c_uint32)), ("numTbnIndices", C.c_uint32),
        ("emissiveMetadata", C.POINTER(InstanceData)), ("numEmissive", C.c_uint32),
        ("cdfTriangles", C.POINTER(C.c_float)), ("numCdfTriangles", C.c_uint32),
        ("cdfInstances", C.POINTER(C.c_float)), ("numCdfInstances", C.c_uint32),
        ("texCoords", C.POINTER(C.c_float)), ("numTexCoords", C.c_uint32),
        ("texIndices", C.POINTER(C.c_uint32)), ("numTexIndices", C.c_uint32),
        ("textures", C.POINTER(Texture)), ("numTextures", C.c_uint32),
        ("instances", C.POINTER(Instance)), ("numInstances", C.c_uint32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("extendRays", C.c_uint64), ("shadowRays", C.c_uint64), ("paths", C.c_uint64),
        ("nodeVisits", C.c_uint64), ("triTests", C.c_uint64), ("waves", C.c_uint64),
        ("kernelLaunches", C.c_uint64), ("shadowNodeVisits", C.c_uint64), ("shadowTriTests", C.c_uint64),
        ("shadowRaysSkipped", C.c_uint64),
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class BvhInfo(C.Structure):
    _fields_ = [
        ("numTriangles", C.c_uint32), ("numWideNodes", C.c_uint32), ("maxDepth", C.c_uint32),
        ("reserved", C.c_uint32), ("nodeBytes", C.c_uint64), ("triangleBytes", C.c_uint64),
        ("hash", C.c_uint64), ("buildMs", C.c_float), ("sceneMin", C.c_float * 3), ("sceneMax", C.c_float * 3),
    ]

    def as_dict(self):
        d = {}
        for k, _ in self._fields_:
            v = getattr(self, k)
            d[k] = list(v) if hasattr(v, "__len__") else (float(v) if isinstance(v, float) else int(v))
        return d


class KernelTimes(C.Structure):
    _fields_ = [("generateMs", C.c_float), ("extendMs", C.c_float), ("shadeMs", C.c_float * 5), ("shadowMs", C.c_float),
                ("finishMs", C.c_float), ("extendLaunches", C.c_uint32), ("shadeLaunches", C.c_uint32),
                ("shadowLaunches", C.c_uint32), ("finishLaunches", C.c_uint32),
                ("extendFullMs", C.c_float), ("shadowFullMs", C.c_float), ("extendFullLaunches", C.c_uint32),
                ("shadowFullLaunches", C.c_uint32), ("extendFullRays", C.c_uint64), ("shadowFullRays", C.c_uint64),
                ("extendRays", C.c_uint64), ("shadowRays", C.c_uint64), ("shadeItems", C.c_uint64 * 5), ("finishItems", C.c_uint64)]

    def as_dict(self):
        return {"generateMs": self.generateMs, "extendMs": self.extendMs, "shadeMs": list(self.shadeMs),
                "shadowMs": self.shadowMs, "finishMs": self.finishMs, "extendLaunches": self.extendLaunches,
                "shadeLaunches": self.shadeLaunches, "shadowLaunches": self.shadowLaunches,
                "finishLaunches": self.finishLaunches, "extendFullMs": self.extendFullMs, "shadowFullMs": self.shadowFullMs,
                "extendFullLaunches": self.extendFullLaunches, "shadowFullLaunches": self.shadowFullLaunches,
                "extendFullRays": self.extendFullRays, "shadowFullRays": self.shadowFullRays,
                "extendRays": self.extendRays, "shadowRays": self.shadowRays, "shadeItems": list(self.shadeItems),
                "finishItems": self.finishItems}


class ShadeResult(C.Structure):
    _fields_ = [("color", C.c_float * 3), ("albedo", C.c_float * 3), ("origin", C.c_float * 3), ("direction", C.c_float * 3),
                ("emission", C.c_float * 3), ("normal", C.c_float * 3), ("pdf", C.c_float), ("accumulatedDistance", C.c_float),
                ("rngState", C.c_uint32), ("flags", C.c_uint32), ("material", C.c_uint32), ("reserved", C.c_uint32)]


class PrimaryHit(C.Structure):
    _fields_ = [("t", C.c_float), ("u", C.c_float), ("v", C.c_float), ("instance", C.c_uint32),
                ("primitive", C.c_uint32)]


assert C.sizeof(InstanceProperties) == 120
assert C.sizeof(RtPushConsts) == 160
assert C.sizeof(InstanceData) == 112
assert C.sizeof(Instance) == 80
assert C.sizeof(PrimaryHit) == 20

# name -> (restype, argtypes); every symbol include/reina_b200.h declares
SYMBOLS = {
    "rb200_version": (C.c_uint32, []),
    "rb200_last_error": (C.c_char_p, []),
    "rb200_context_create": (C.c_int, [C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.POINTER(C.c_void_p)]),
    "rb200_context_destroy": (C.c_int, [C.c_void_p]),
    "rb200_context_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rb200_scene_create": (C.c_int, [C.c_void_p, C.POINTER(SceneDesc), C.POINTER(C.c_void_p)]),
    "rb200_scene_destroy": (C.c_int, [C.c_void_p]),
    "rb200_scene_bvh_info": (C.c_int, [C.c_void_p, C.POINTER(BvhInfo)]),
    "rb200_render_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(RtPushConsts)]),
    "rb200_resolve_sum": (C.c_int, [C.c_void_p, C.c_uint32]),
    "rb200_wait_batches_pending": (C.c_int, [C.c_void_p, C.c_uint32]),
    "rb200_postprocess": (C.c_int, [C.c_void_p, C.POINTER(BloomPushConsts), C.POINTER(TonemappingPushConsts)]),
    "rb200_read_ldr": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rb200_read_hdr": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rb200_read_ldr_async": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rb200_context_set_tiles": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]),
    "rb200_present_sum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(BloomPushConsts), C.POINTER(TonemappingPushConsts)]),
    "rb200_wait_ldr": (C.c_int, [C.c_void_p]),
    "rb200_wait_ldr_pending": (C.c_int, [C.c_void_p, C.c_uint32]),
    "rb200_pipeline_depth": (C.c_uint32, []),
    "rb200_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "rb200_host_free": (C.c_int, [C.c_void_p]),
    "rb200_measure_gather": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint32, C.POINTER(C.c_float)]),
    "rb200_write_hdr": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rb200_hdr_device_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "rb200_trace_primary": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(RtPushConsts), C.c_void_p]),
    "rb200_trace_rays": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_void_p]),
    "rb200_shade_hits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p]),
    "rb200_bench_trace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint32,
                                    C.POINTER(C.c_float), C.POINTER(C.c_uint64)]),
    "rb200_engine_config": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]),
    "rb200_get_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats), C.POINTER(Stats)]),
    "rb200_get_kernel_times": (C.c_int, [C.c_void_p, C.POINTER(KernelTimes)]),
    "rb200_synchronize": (C.c_int, [C.c_void_p]),
    # several GPUs
    "rb200_group_create": (C.c_int, [C.c_uint32, C.c_uint32, C.POINTER(C.c_int), C.c_int, C.c_uint32, C.POINTER(C.c_void_p)]),
    "rb200_group_destroy": (C.c_int, [C.c_void_p]),
    "rb200_group_size": (C.c_int, [C.c_void_p]),
    "rb200_group_uses_peer_stores": (C.c_int, [C.c_void_p]),
    "rb200_group_context": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "rb200_group_set_tile_size": (C.c_int, [C.c_void_p, C.c_uint32]),
    "rb200_group_scene_create": (C.c_int, [C.c_void_p, C.POINTER(SceneDesc), C.POINTER(C.c_void_p)]),
    "rb200_group_scene_destroy": (C.c_int, [C.c_void_p]),
    "rb200_group_scene_bvh_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(BvhInfo)]),
    "rb200_group_render_batches": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(RtPushConsts), C.c_uint32, C.c_uint32]),
    "rb200_group_present": (C.c_int, [C.c_void_p, C.POINTER(BloomPushConsts), C.POINTER(TonemappingPushConsts)]),
    "rb200_group_read_ldr": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rb200_group_read_hdr": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rb200_group_synchronize": (C.c_int, [C.c_void_p]),
    "rb200_group_get_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "rb200_comm_unique_id": (C.c_int, [C.c_void_p]),
    "rb200_context_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "rb200_context_comm_destroy": (C.c_int, [C.c_void_p]),
    "rb200_context_reduce_present": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(BloomPushConsts), C.POINTER(TonemappingPushConsts)]),
    "rb200_context_reduced_device_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
}

_lib = None


class RB200Error(RuntimeError):
    pass


def load_library(path=None):
    """Load csrc/librb200.so and declare every entry point. Raises if the CUDA library is not built:
    this package has no CPU path."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RB200Error(
            f"{p} not found: the CUDA extension is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C reina-vk_b200/csrc`). reina-vk_b200 has no CPU fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def check(lib, status):
    if status != RB200_OK:
        msg = lib.rb200_last_error().decode("utf-8", "replace")
        raise RB200Error(f"rb200 error {status}: {msg}")
