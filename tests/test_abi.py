"""The drop-in boundary: include/reina_b200.h <-> csrc/librb200.so <-> the ctypes view. No GPU needed: nothing here
launches a kernel, it only checks that the CUDA library loads, exports every declared symbol and agrees on layouts
with the reference's polyglot structs (polyglot/raytrace.h:18-54, src/scene/Instances.h:15-27)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "reina_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"RB200_API\s+[\w\s\*]+?\b(rb200_\w+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for must in ("rb200_context_create", "rb200_scene_create", "rb200_render_batch", "rb200_postprocess", "rb200_read_ldr",
                 "rb200_read_hdr", "rb200_last_error", "rb200_trace_primary"):
        assert must in syms


def test_library_exports_every_declared_symbol(rb):
    lib = rb.load_library()
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in reina_b200.h but not exported by librb200.so"
    # and the ctypes table covers exactly the header
    assert sorted(rb.abi.SYMBOLS) == declared_symbols()


def test_missing_library_fails_loudly(rb, tmp_path):
    with pytest.raises(rb.RB200Error, match="no CPU fallback"):
        rb.abi.load_library(str(tmp_path / "nope.so"))


def test_struct_sizes_and_offsets(rb):
    a = rb.abi
    assert C.sizeof(a.InstanceProperties) == 120 and C.sizeof(a.RtPushConsts) == 160 and C.sizeof(a.InstanceData) == 112
    expect = {"indicesOffset": 0, "albedo": 4, "emission": 16, "tbnsIndicesOffset": 28, "texIndicesOffset": 32, "roughness": 36,
              "ior": 40, "interpNormals": 44, "absorption": 48, "textureID": 52, "normalMapTexID": 56, "bumpMapTexID": 60,
              "cullBackface": 64, "anisotropic": 68, "subsurface": 72, "clearcoatGloss": 76, "sheenTint": 80,
              "specularTint": 92, "metallic": 104, "clearcoat": 108, "specularTransmission": 112, "sheen": 116}
    for k, off in expect.items():
        assert getattr(a.InstanceProperties, k).offset == off, k
    pc = {"invView": 0, "invProjection": 64, "sampleBatch": 128, "totalEmissiveWeight": 132, "focusDist": 136,
          "defocusMultiplier": 140, "directClamp": 144, "indirectClamp": 148, "samplesPerPixel": 152, "maxBounces": 156}
    for k, off in pc.items():
        assert getattr(a.RtPushConsts, k).offset == off, k
    em = {"transform": 0, "materialOffset": 64, "cdfRangeStart": 68, "cdfRangeEnd": 72, "indexOffset": 76, "emission": 80,
          "weight": 92, "area": 96, "cullBackface": 100, "padding": 104}
    for k, off in em.items():
        assert getattr(a.InstanceData, k).offset == off, k


def test_header_offsets_match_comments():
    """The numeric offsets written next to every field in the header are the contract a cgo/JNI/ctypes binding relies on."""
    text = open(HEADER).read()
    block = text[text.index("typedef struct RB200InstanceProperties"):text.index("} RB200InstanceProperties;")]
    offs = [int(x) for x in re.findall(r"/\*\s+(\d+)\s", block)]
    assert offs == sorted(offs) and offs[0] == 0 and offs[-1] == 116


def test_no_device_is_an_error_not_a_fallback(rb):
    """On a box without a GPU context creation must fail with a message; with a GPU this is covered by the gpu tests."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = rb.load_library()
    ctx = C.c_void_p()
    rc = lib.rb200_context_create(64, 64, 0, 0, C.byref(ctx))
    assert rc != 0
    assert b"no CUDA device" in lib.rb200_last_error() or b"CUDA" in lib.rb200_last_error()


def test_header_compiles_as_plain_c99_and_cxx11(tmp_path):
    """The boundary is a C ABI: include/reina_b200.h must be usable from C (a cgo / JNI / ctypes-style binding) and from
    the C++ the reference is written in, with the POD sizes the reference's polyglot headers have."""
    import shutil
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "reina_b200.h"\n'
                   'int main(void) { RB200RtPushConsts pc; (void)pc;\n'
                   '  return sizeof(RB200InstanceProperties) == 120 && sizeof(RB200RtPushConsts) == 160 && sizeof(RB200InstanceData) == 112\n'
                   '         && sizeof(RB200Instance) == 80 ? 0 : 1; }\n')
    inc = os.path.join(ROOT, "include")
    cc, cxx = shutil.which("gcc"), shutil.which("g++")
    if not cc or not cxx:
        pytest.skip("no host compiler")
    exe = str(tmp_path / "hdr")
    subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-o", exe, str(src)], check=True)
    assert subprocess.run([exe]).returncode == 0
    subprocess.run([cxx, "-std=c++11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-x", "c++", "-fsyntax-only", str(src)], check=True)
