"""The C++ headless host (reina-vk_b200/host/): SURVEY.md §8(f) row 1.

CPU tests compare the C++ scene layer, OBJ importer, TOML reader, push-constant derivation, PNG writer and save
policy with the Python host (same reference interfaces: src/scene/Scene.cpp:6-125, src/scene/Models.cpp:117-175,
src/Reina.cpp:142-155, src/tools/SaveManager.cpp:6-44, src/tools/Clock.cpp:27-40) through the test hooks of
host/capi.cpp. The GPU test runs the reina_b200 binary end to end and compares its PNG with the frame the Python
host gets from the same library for the same push constants.
"""
import ctypes as C
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "reina-vk_b200", "host")

REFERENCE_SCHEMA = """
[camera.dof]
focus_dist = 2.2
defocus_multiplier = 1.5

[sampling]
samples_per_pixel = 8
max_bounces = 16
direct_clamp = 100
indirect_clamp = 10

[saving]
save_on_samples = [64, 256, 1024]
save_on_times = [60.0]

[postprocessing.bloom]
radius = 5.0
threshold = 1.0
intensity = 0.05

[postprocessing.tonemap]
exposure = 1
"""


class HostConfig(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("nee", C.c_uint32), ("samplesPerPixel", C.c_uint32),
                ("maxBounces", C.c_uint32), ("numSaveSamples", C.c_uint32), ("numSaveTimes", C.c_uint32),
                ("focusDist", C.c_float), ("defocusMultiplier", C.c_float), ("directClamp", C.c_float),
                ("indirectClamp", C.c_float), ("bloomRadius", C.c_float), ("bloomThreshold", C.c_float),
                ("bloomIntensity", C.c_float), ("exposure", C.c_float), ("cameraPos", C.c_double * 3),
                ("cameraLookAt", C.c_double * 3), ("fovYDegrees", C.c_double), ("saveSamples", C.c_int32 * 16),
                ("saveTimes", C.c_double * 16), ("scene", C.c_char * 64)]


@pytest.fixture(scope="module")
def host(rb):
    so = os.path.join(HOST, "libreina_host.so")
    if not os.path.exists(so) or not os.path.exists(os.path.join(HOST, "reina_b200")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "reina-vk_b200", "csrc")], check=True, capture_output=True)
        subprocess.run(["make", "-C", HOST], check=True, capture_output=True)
    lib = C.CDLL(so)
    lib.rbhost_last_error.restype = C.c_char_p
    lib.rbhost_png_encode.restype = C.c_int64
    lib.rbhost_png_encode.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64]
    lib.rbhost_tables_free.argtypes = [C.c_void_p]
    lib.rbhost_tables_desc.argtypes = [C.c_void_p, C.POINTER(rb.abi.SceneDesc), C.POINTER(C.c_float)]
    lib.rbhost_push_constants.argtypes = [C.c_char_p, C.c_float, C.POINTER(rb.abi.RtPushConsts)]
    return lib


def err(lib):
    return lib.rbhost_last_error().decode()


def arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * np.dtype(dtype).itemsize,)).view(dtype).copy()


def cpp_tables(lib, rb, handle):
    """Snapshot the C++ tables behind `handle` as numpy arrays, keyed like SceneTables."""
    d = rb.abi.SceneDesc()
    w = C.c_float()
    assert lib.rbhost_tables_desc(handle, C.byref(d), C.byref(w)) == 0, err(lib)
    t = dict(
        vertices=arr(d.vertices, d.numVertices * 4, np.float32), indices=arr(d.indices, d.numIndices, np.uint32),
        instanceProperties=arr(d.instanceProperties, d.numInstanceProperties * 120, np.uint8),
        tbns=arr(d.tbns, d.numTbns * 9, np.float32), tbnIndices=arr(d.tbnIndices, d.numTbnIndices, np.uint32),
        emissive=arr(d.emissiveMetadata, d.numEmissive * 112, np.uint8),
        cdfTriangles=arr(d.cdfTriangles, d.numCdfTriangles, np.float32),
        cdfInstances=arr(d.cdfInstances, d.numCdfInstances, np.float32),
        texCoords=arr(d.texCoords, d.numTexCoords * 2, np.float32), texIndices=arr(d.texIndices, d.numTexIndices, np.uint32),
        instances=arr(d.instances, d.numInstances * 80, np.uint8), totalEmissiveWeight=w.value,
        textures=[arr(d.textures[i].rgba8, d.textures[i].width * d.textures[i].height * 4, np.uint8)
                  .reshape(d.textures[i].height, d.textures[i].width, 4) for i in range(d.numTextures)])
    return t


def py_tables(tb):
    n_i, n_e, n_p = tb.numInstances, tb.numEmissive, tb.numInstanceProperties
    return dict(vertices=tb.vertices.reshape(-1), indices=tb.indices, instanceProperties=tb.instanceProperties[:n_p * 120],
                tbns=tb.tbns, tbnIndices=tb.tbnIndices, emissive=tb.emissive[:n_e * 112], cdfTriangles=tb.cdfTriangles,
                cdfInstances=tb.cdfInstances, texCoords=tb.texCoords, texIndices=tb.texIndices,
                instances=tb.instances[:n_i * 80], totalEmissiveWeight=tb.totalEmissiveWeight, textures=tb.textures)


def assert_tables_identical(a, b, approx=()):
    for k in a:
        if k == "textures":
            assert len(a[k]) == len(b[k])
            for x, y in zip(a[k], b[k]):
                assert x.shape == y.shape and (x == y).all()
        elif k == "totalEmissiveWeight":
            assert np.float32(a[k]) == np.float32(b[k])
        elif k in approx:
            assert a[k].shape == b[k].shape, k
            np.testing.assert_allclose(a[k], b[k], rtol=0, atol=2e-7, err_msg=k)
        else:
            assert a[k].shape == b[k].shape, k
            assert a[k].tobytes() == b[k].tobytes(), k


# ---------------------------------------------------------------------------------------------------------------
# scene layer
# ---------------------------------------------------------------------------------------------------------------
def test_cornell_tables_identical_to_python_host(host, rb):
    h = C.c_void_p()
    assert host.rbhost_tables_builtin(b"cornell", 1, C.byref(h)) == 0, err(host)
    wl = rb.configs.cornell(96, 72, nee=True)
    assert_tables_identical(cpp_tables(host, rb, h), py_tables(wl.tables))
    host.rbhost_tables_free(h)


def test_cornell_sphere_tables_match_python_host(host, rb):
    """The sphere goes through sin/cos of two different maths libraries: positions/frames to 2e-7, the rest exact."""
    h = C.c_void_p()
    assert host.rbhost_tables_builtin(b"cornell-sphere", 1, C.byref(h)) == 0, err(host)
    wl = rb.configs.cornell(96, 72, with_sphere=True, nee=True)
    assert_tables_identical(cpp_tables(host, rb, h), py_tables(wl.tables), approx=("vertices", "tbns", "texCoords"))
    host.rbhost_tables_free(h)


def test_unknown_scene_and_missing_obj_fail_loudly(host):
    h = C.c_void_p()
    assert host.rbhost_tables_builtin(b"sponza", 1, C.byref(h)) != 0
    assert "unknown built-in scene" in err(host)
    assert host.rbhost_tables_obj(b"/nonexistent/model.obj", 0, 1, C.byref(h)) != 0
    assert "Could not load model" in err(host)


OBJ_FULL = """# cube-ish test asset: quads and a pentagon, v/vt/vn triples, relative indices, shared corners
v -0.5 -0.5 0.5
v 0.5 -0.5 0.5
v 0.5 0.5 0.5
v -0.5 0.5 0.5
v -0.5 -0.5 -0.5
v 0.5 -0.5 -0.5
v 0.5 0.5 -0.5
v -0.5 0.5 -0.5
v 0.0 0.9 0.5   # apex of the pentagon
vt 0.0 0.0
vt 1.0 0.0
vt 1.0 1.0
vt 0.0 1.0
vt 0.5 1.25
vn 0 0 1
vn 0 0 -1
vn 1 0 0
vn -1 0 0
vn 0.0 0.70710678 0.70710678
f 1/1/1 2/2/1 3/3/1 9/5/5 4/4/1
f 6/1/2 5/2/2 8/3/2 7/4/2
f 2/1/3 6/2/3 7/3/3 3/4/3
f -5/1/4 -9/2/4 -6/3/4 -2/4/4
"""

OBJ_BARE = """v 0 0 0
v 1 0 0
v 1 1 0.1
v 0 1 0
v 0.5 0.5 1
f 1 2 3 4
f 1 2 5
f 2 3 5
f 3//  4// 5//
"""


@pytest.mark.parametrize("text, light", [(OBJ_FULL, 1), (OBJ_BARE, 0)])
def test_obj_importer_identical_to_python_importer(host, rb, tmp_path, text, light):
    import warnings
    p = tmp_path / "asset.obj"
    p.write_text(text)
    h = C.c_void_p()
    assert host.rbhost_tables_obj(str(p).encode(), 0, light, C.byref(h)) == 0, err(host)
    s = rb.scene.Scene()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        md = rb.meshes.load_obj(str(p))
    s.addObject(md, np.eye(4, dtype=np.float32), rb.scene.Material(albedo=(0.8, 0.8, 0.8), interpNormals=True))
    if light:
        s.addObject(rb.meshes.cornell_light(), np.eye(4, dtype=np.float32), rb.scene.Material(**rb.configs.LIGHT))
    assert_tables_identical(cpp_tables(host, rb, h), py_tables(s.build()))
    host.rbhost_tables_free(h)


# ---------------------------------------------------------------------------------------------------------------
# configuration
# ---------------------------------------------------------------------------------------------------------------
def test_config_reads_the_reference_schema(host):
    c = HostConfig()
    assert host.rbhost_config_parse(REFERENCE_SCHEMA.encode(), C.byref(c)) == 0, err(host)
    assert (c.samplesPerPixel, c.maxBounces) == (8, 16)
    assert (c.focusDist, c.defocusMultiplier) == (np.float32(2.2), np.float32(1.5))
    assert (c.directClamp, c.indirectClamp) == (100.0, 10.0)          # integers convert to float
    assert (c.bloomRadius, c.bloomThreshold, c.bloomIntensity, c.exposure) == (5.0, 1.0, np.float32(0.05), 1.0)
    assert list(c.saveSamples[:c.numSaveSamples]) == [64, 256, 1024]
    assert list(c.saveTimes[:c.numSaveTimes]) == [60.0]
    assert (c.width, c.height, c.scene) == (800, 600, b"cornell")     # [render] defaults


def test_repo_config_file_parses(host):
    c = HostConfig()
    text = open(os.path.join(ROOT, "config", "config.toml")).read()
    assert host.rbhost_config_parse(text.encode(), C.byref(c)) == 0, err(host)
    assert c.nee == 1 and list(c.cameraPos) == [0.0, 1.0, 3.9] and c.fovYDegrees == 40.0


def test_config_every_reference_key_is_required(host):
    """at_path(...).value<T>().value() on a missing key ends start-up (std::bad_optional_access in the reference)."""
    keys = ["focus_dist", "defocus_multiplier", "samples_per_pixel", "max_bounces", "direct_clamp", "indirect_clamp",
            "save_on_samples", "save_on_times", "radius", "threshold", "intensity", "exposure"]
    for k in keys:
        text = "\n".join(l for l in REFERENCE_SCHEMA.splitlines() if not l.startswith(k + " "))
        c = HostConfig()
        assert host.rbhost_config_parse(text.encode(), C.byref(c)) != 0, k
        assert "missing key" in err(host) and k in err(host)


def test_config_type_rules_and_syntax(host):
    c = HostConfig()
    bad = REFERENCE_SCHEMA.replace("samples_per_pixel = 8", "samples_per_pixel = 8.0")     # float -> uint32: refused
    assert host.rbhost_config_parse(bad.encode(), C.byref(c)) != 0 and "not an integer" in err(host)
    bad = REFERENCE_SCHEMA.replace("save_on_samples = [64, 256, 1024]", "save_on_samples = 64")
    assert host.rbhost_config_parse(bad.encode(), C.byref(c)) != 0 and "not an array" in err(host)
    bad = REFERENCE_SCHEMA.replace("radius = 5.0", "radius = 5.0\nradius = 6.0")
    assert host.rbhost_config_parse(bad.encode(), C.byref(c)) != 0 and "defined twice" in err(host)
    bad = REFERENCE_SCHEMA.replace("exposure = 1", "exposure = 1 2")
    assert host.rbhost_config_parse(bad.encode(), C.byref(c)) != 0 and "line" in err(host)
    ok = REFERENCE_SCHEMA.replace("save_on_samples = [64, 256, 1024]",
                                  "save_on_samples = [  # thresholds\n  1_024,\n  64, # small\n  256,\n]")
    ok += '\n[render]\nwidth = 0x140\nheight = 200\nscene = "cornell-sphere"\nnee = false\ncamera_pos = [1, 2.5, -3e0]\n'
    assert host.rbhost_config_parse(ok.encode(), C.byref(c)) == 0, err(host)
    assert list(c.saveSamples[:c.numSaveSamples]) == [1024, 64, 256]
    assert (c.width, c.height, c.scene, c.nee) == (320, 200, b"cornell-sphere", 0)
    assert list(c.cameraPos) == [1.0, 2.5, -3.0]


def test_push_constants_match_python_host(host, rb):
    pc = rb.abi.RtPushConsts()
    text = REFERENCE_SCHEMA + "\n[render]\nwidth = 640\nheight = 360\ncamera_pos = [-1.6899, 0.817017, -1.6386]\n" \
                              "camera_look_at = [0.0, 0.6, 0.0]\nfov_y_degrees = 30.0\n"
    assert host.rbhost_push_constants(text.encode(), 7.5, C.byref(pc)) == 0, err(host)
    ref = rb.camera.push_constants(640, 360, (-1.6899, 0.817017, -1.6386), (0.0, 0.6, 0.0), 30.0, total_emissive_weight=7.5)
    np.testing.assert_allclose(np.array(pc.invView[:]), np.array(ref.invView[:]), rtol=0, atol=1e-6)
    np.testing.assert_allclose(np.array(pc.invProjection[:]), np.array(ref.invProjection[:]), rtol=1e-6, atol=1e-6)
    for f in ("sampleBatch", "totalEmissiveWeight", "focusDist", "defocusMultiplier", "directClamp", "indirectClamp",
              "samplesPerPixel", "maxBounces"):
        assert getattr(pc, f) == getattr(ref, f), f
    assert pc.defocusMultiplier == np.float32(np.float32(1.5) / np.float32(100.0))    # src/Reina.cpp:150


# ---------------------------------------------------------------------------------------------------------------
# PNG + save policy
# ---------------------------------------------------------------------------------------------------------------
def decode_png(data):
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, []
    while pos < len(data):
        n, typ = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        crc, = struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])
        assert zlib.crc32(typ + body) == crc, typ
        chunks.append((typ, body))
        pos += 12 + n
    assert [c[0] for c in chunks][0] == b"IHDR" and chunks[-1] == (b"IEND", b"")
    w, h, depth, ctype, comp, filt, inter = struct.unpack(">IIBBBBB", chunks[0][1])
    assert (depth, ctype, comp, filt, inter) == (8, 6, 0, 0, 0)
    raw = zlib.decompress(b"".join(b for t, b in chunks if t == b"IDAT"))
    rows = np.frombuffer(raw, np.uint8).reshape(h, 1 + 4 * w)
    assert (rows[:, 0] == 0).all()
    return rows[:, 1:].reshape(h, w, 4)


def test_png_round_trip(host):
    rng = np.random.default_rng(5)
    for (h, w) in [(1, 1), (7, 13), (72, 96)]:
        img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        out = np.zeros(h * w * 4 + 4096, np.uint8)
        n = host.rbhost_png_encode(img.ctypes.data, w, h, out.ctypes.data, out.size)
        assert 0 < n <= out.size, err(host)
        assert (decode_png(out[:n].tobytes()) == img).all()
    assert host.rbhost_png_encode(None, 0, 0, out.ctypes.data, out.size) < 0 and "Could not save PNG" in err(host)


def schedule(host, samples, times, frames, spp, dt):
    s = (C.c_int32 * max(1, len(samples)))(*samples)
    t = (C.c_double * max(1, len(times)))(*times)
    buf = C.create_string_buffer(4096)
    assert host.rbhost_save_schedule(s, len(samples), t, len(times), frames, spp, C.c_double(dt), buf, 4096) == 0, err(host)
    return [(int(l.split(":")[0]), l.split(":")[1]) for l in buf.value.decode().splitlines()]


def test_save_policy_follows_the_reference_clock(host):
    """SaveManager fires on `threshold < samples` with Clock's counter, which skips the first frame's samples and is
    read before markFrame: with 8 spp the 64-sample file is written in frame 10 (0-based), the 256 one in frame 34."""
    got = schedule(host, [256, 64], [], 40, 8, 0.0)
    assert got == [(10, "output_64spp.png"), (34, "output_256spp.png")]
    # one save per frame, samples before times, times named by truncation, strictly-greater comparison
    got = schedule(host, [8], [0.5, 2.9], 8, 8, 0.5)
    assert got == [(1, "output_0sec.png"), (3, "output_8spp.png"), (5, "output_2sec.png")]
    assert schedule(host, [], [], 5, 8, 1.0) == []


# ---------------------------------------------------------------------------------------------------------------
# texture ingest
# ---------------------------------------------------------------------------------------------------------------
def png_load(host, path, flip):
    buf = np.zeros(1 << 20, np.uint8)
    w, h = C.c_uint32(), C.c_uint32()
    rc = host.rbhost_png_load(str(path).encode(), int(flip), buf.ctypes.data_as(C.c_void_p), C.c_uint64(buf.size),
                              C.byref(w), C.byref(h))
    if rc != 0:
        raise RuntimeError(err(host))
    return buf[:w.value * h.value * 4].reshape(h.value, w.value, 4).copy()


def test_png_decoder_matches_pil_on_every_colour_type(host, tmp_path):
    """8-bit RGBA the way stb_image returns it with 4 requested channels (src/graphics/Image.cpp:14-15): grey and
    palette expanded, tRNS honoured, file textures flipped vertically. PIL writes with adaptive scanline filters."""
    from PIL import Image
    rng = np.random.default_rng(9)
    h, w = 37, 53
    rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    smooth = (np.add.outer(np.arange(h) * 3, np.arange(w) * 2) % 256).astype(np.uint8)       # exercises Sub/Up/Paeth
    cases = {
        "rgba": Image.fromarray(rgba, "RGBA"),
        "rgb": Image.fromarray(rgba[..., :3].copy(), "RGB"),
        "grey": Image.fromarray(smooth, "L"),
        "grey_alpha": Image.fromarray(np.stack([smooth, rgba[..., 3]], axis=2), "LA"),
        "palette": Image.fromarray(rgba[..., :3].copy(), "RGB").quantize(64),
        "bilevel": Image.fromarray((smooth > 127).astype(np.uint8) * 255, "L").convert("1"),
    }
    for name, im in cases.items():
        p = tmp_path / f"{name}.png"
        im.save(p)
        want = np.asarray(Image.open(p).convert("RGBA"))
        assert (png_load(host, p, False) == want).all(), name
        assert (png_load(host, p, True) == want[::-1]).all(), name
    # palette with per-entry alpha (tRNS)
    pal = Image.fromarray(rgba[..., :3].copy(), "RGB").quantize(16)
    p = tmp_path / "palette_trns.png"
    pal.save(p, transparency=bytes([0, 128, 255, 7] + [255] * 12))
    assert (png_load(host, p, False) == np.asarray(Image.open(p).convert("RGBA"))).all()
    # 16-bit grey keeps the high byte
    g16 = (np.add.outer(np.arange(h), np.arange(w)) * 517 % 65536).astype(np.uint16)
    p = tmp_path / "grey16.png"
    Image.fromarray(g16).save(p)
    got = png_load(host, p, False)
    assert (got[..., 0] == (g16 >> 8)).all() and (got[..., 3] == 255).all()
    # our own writer round-trips through our own reader
    out = np.zeros(h * w * 4 + 4096, np.uint8)
    n = host.rbhost_png_encode(rgba.ctypes.data, w, h, out.ctypes.data, out.size)
    (tmp_path / "own.png").write_bytes(out[:n].tobytes())
    assert (png_load(host, tmp_path / "own.png", False) == rgba).all()


def _write_png(path, samples, ctype, depth, interlace, rng, palette=None, trns=None):
    """A PNG writer for the tests: `samples` is (h, w, channels) of full-depth sample values; scanlines get every filter
    type in turn; interlace = 1 splits the image into the seven Adam7 passes (PNG specification, section 8.2)."""
    import struct
    import zlib
    h, w, ch = samples.shape
    bpp = max(1, ch * depth // 8)

    def pack(rows):                                     # (n, pw, ch) sample rows -> list of packed byte rows
        out = []
        for r in rows:
            flat = r.reshape(-1)
            if depth == 16:
                out.append(flat.astype(">u2").tobytes())
            elif depth == 8:
                out.append(flat.astype(np.uint8).tobytes())
            else:
                bits = np.zeros(((flat.size * depth + 7) // 8) * 8, np.uint8)
                for k in range(depth):
                    bits[k:flat.size * depth:depth] = (flat >> (depth - 1 - k)) & 1
                out.append(np.packbits(bits).tobytes())
        return out

    def paeth(a, b, c):
        pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
        return a if pa <= pb and pa <= pc else (b if pb <= pc else c)

    def filtered(lines, first):
        out, prev = bytearray(), bytes(len(lines[0])) if lines else b""
        for i, line in enumerate(lines):
            f = (first + i) % 5
            out.append(f)
            for x, v in enumerate(line):
                a = line[x - bpp] if x >= bpp else 0
                b = prev[x]
                c = prev[x - bpp] if x >= bpp else 0
                out.append((v - (0, a, b, (a + b) >> 1, paeth(a, b, c))[f]) & 255)
            prev = line
        return bytes(out)
    passes = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)] if interlace else [(0, 0, 1, 1)]
    raw = b""
    for k, (x0, y0, dx, dy) in enumerate(passes):
        sub = samples[y0::dy, x0::dx]
        if sub.shape[0] and sub.shape[1]:
            raw += filtered(pack(sub), int(rng.integers(0, 5)) + k)

    def chunk(t, body):
        return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body))
    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, interlace))
    if palette is not None:
        data += chunk(b"PLTE", bytes(palette))
    if trns is not None:
        data += chunk(b"tRNS", bytes(trns))
    z = zlib.compress(raw, 6)
    half = len(z) // 2
    data += chunk(b"IDAT", z[:half]) + chunk(b"IDAT", z[half:]) + chunk(b"IEND", b"")
    path.write_bytes(data)


def test_png_decoder_adam7_interlace_matches_pil(host, tmp_path):
    """Interlaced PNGs (stb_image, the reference's decoder, reads them): seven passes with their own scanline lengths and
    filter histories, empty passes for images narrower / shorter than the 8 x 8 pattern, sub-byte depths packed per pass.
    PIL cannot write Adam7, so the files come from the writer above; PIL reads them and must agree byte for byte."""
    from PIL import Image
    rng = np.random.default_rng(31)
    p = tmp_path / "i.png"
    kinds = [(0, 1, 1), (0, 2, 1), (0, 4, 1), (0, 8, 1), (0, 16, 1), (2, 8, 3), (2, 16, 3), (3, 1, 1), (3, 2, 1), (3, 4, 1), (3, 8, 1), (4, 8, 2),
             (4, 16, 2), (6, 8, 4), (6, 16, 4)]
    cases = 0
    for (w, h) in [(1, 1), (2, 3), (3, 2), (4, 4), (5, 7), (8, 8), (9, 9), (33, 17), (16, 40)]:
        for ctype, depth, ch in kinds:
            top = (1 << depth) if ctype != 3 else min(1 << depth, 23)
            samples = rng.integers(0, top, (h, w, ch)).astype(np.uint32)
            palette = rng.integers(0, 256, 3 * top).astype(np.uint8) if ctype == 3 else None
            trns = rng.integers(0, 256, 5).astype(np.uint8) if ctype == 3 and depth >= 4 else None
            for interlace in (1, 0):
                _write_png(p, samples, ctype, depth, interlace, rng, palette, trns)
                got = png_load(host, p, False)
                if depth == 16:                          # PIL keeps 16 bits for grey and rescales; the host keeps the high byte
                    want = (samples >> 8).astype(np.uint8)
                    want = {1: lambda a: np.concatenate([a, a, a, np.full_like(a, 255)], 2), 2: lambda a: np.concatenate([a[..., :1]] * 3 + [a[..., 1:]], 2),
                            3: lambda a: np.concatenate([a, np.full_like(a[..., :1], 255)], 2), 4: lambda a: a}[ch](want)
                else:
                    want = np.asarray(Image.open(p).convert("RGBA"))
                assert (got == want).all(), (w, h, ctype, depth, interlace)
                assert (png_load(host, p, True) == want[::-1]).all()
                cases += 1
    assert cases == 9 * 15 * 2
    data = bytearray(p.read_bytes())
    data[28] = 2                                           # interlace method 2 does not exist (the checksum is checked first)
    import struct
    import zlib
    data[29:33] = struct.pack(">I", zlib.crc32(bytes(data[12:29])))
    p.write_bytes(bytes(data))
    with pytest.raises(RuntimeError, match="unknown interlace method"):
        png_load(host, p, False)


def test_png_decoder_refuses_what_it_does_not_support(host, tmp_path):
    from PIL import Image
    (tmp_path / "not.png").write_bytes(b"JFIF" * 10)
    with pytest.raises(RuntimeError, match="Could not load image at path"):
        png_load(host, tmp_path / "not.png", True)
    with pytest.raises(RuntimeError, match="Could not load image at path"):
        png_load(host, tmp_path / "absent.png", True)
    im = Image.fromarray(np.zeros((8, 8, 3), np.uint8), "RGB")
    good = tmp_path / "good.png"
    im.save(good)
    data = bytearray(good.read_bytes())
    data[40] ^= 0xFF                                       # corrupt a byte inside a chunk: checksum must catch it
    (tmp_path / "bad.png").write_bytes(bytes(data))
    with pytest.raises(RuntimeError, match="checksum|corrupt"):
        png_load(host, tmp_path / "bad.png", True)


# ---------------------------------------------------------------------------------------------------------------
# end to end on the GPU
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_cli_renders_cornell_like_the_python_host(host, rb, tmp_path):
    cfg = tmp_path / "config.toml"
    cfg.write_text(REFERENCE_SCHEMA.replace("samples_per_pixel = 8", "samples_per_pixel = 4")
                   .replace("save_on_samples = [64, 256, 1024]", "save_on_samples = [8]")
                   .replace("save_on_times = [60.0]", "save_on_times = []")
                   + "\n[render]\nwidth = 96\nheight = 72\nscene = \"cornell\"\nnee = true\n")
    out, pcfile = tmp_path / "final.png", tmp_path / "pc.bin"
    run = subprocess.run([os.path.join(HOST, "reina_b200"), "--config", str(cfg), "--spp", "24", "--out", str(out),
                          "--outdir", str(tmp_path), "--dump-pc", str(pcfile)], capture_output=True, text=True)
    assert run.returncode == 0, run.stderr + run.stdout
    assert "6 frames, 24 samples per pixel" in run.stdout
    assert (tmp_path / "output_8spp.png").exists()          # 8 < 4 * (k - 1)  ->  frame 4 of 0..5

    pc = rb.abi.RtPushConsts.from_buffer_copy(pcfile.read_bytes())
    wl = rb.configs.cornell(96, 72, nee=True)
    r = rb.Renderer(96, 72, wl.tables, flags=rb.RB200_FLAG_NEE)
    frames = {}
    for b in range(6):
        pc.sampleBatch = b
        r.render_batch(pc)
        if b in (4, 5):
            r.postprocess()
            frames[b] = r.read_ldr().copy()
    r.close()
    assert (decode_png(out.read_bytes()) == frames[5]).all()
    assert (decode_png((tmp_path / "output_8spp.png").read_bytes()) == frames[4]).all()


def _image_load(host, path, flip=0):
    host.rbhost_image_load.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    buf = np.zeros(512 * 512 * 4, np.uint8)
    w, h = C.c_uint32(), C.c_uint32()
    if host.rbhost_image_load(str(path).encode(), flip, buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(w), C.byref(h)) != 0:
        raise RuntimeError(err(host))
    return buf[: w.value * h.value * 4].reshape(h.value, w.value, 4).copy()


def test_jpeg_decoder_matches_pil_byte_for_byte(host, tmp_path):
    """host/jpeg.cpp (baseline and progressive Huffman JPEG: IJG integer IDCT, fancy chroma upsampling, fixed-point
    YCbCr -> RGB) against PIL / libjpeg-turbo: identical bytes for 4:4:4, 4:2:2 and 4:2:0, odd sizes down to 1x1, optimised
    Huffman tables, restart intervals, grayscale, quality 5..100, sequential and progressive (spectral selection +
    successive approximation: DC / AC first and refinement scans); file textures flipped vertically like the reference's."""
    from PIL import Image
    rng = np.random.default_rng(5)

    def picture(w, h):
        y, x = np.mgrid[0:h, 0:w]
        img = np.zeros((h, w, 3), np.uint8)
        img[..., 0] = (128 + 100 * np.sin(x / 7.0) * np.cos(y / 11.0)).astype(np.uint8)
        img[..., 1] = ((x * 5 + y * 3) % 256).astype(np.uint8)
        img[..., 2] = (rng.integers(0, 128, (h, w)) + ((x // 8 + y // 8) % 2) * 100).astype(np.uint8)
        return img
    p = tmp_path / "t.jpg"
    cases = 0
    for (w, h) in [(64, 48), (37, 29), (16, 16), (1, 1), (3, 5), (130, 67), (17, 33)]:
        for sub in (0, 1, 2):
            for kw in (dict(quality=30), dict(quality=75, optimize=True), dict(quality=95, restart_marker_blocks=2), dict(quality=100),
                       dict(quality=5, progressive=True), dict(quality=40, progressive=True), dict(quality=85, progressive=True, restart_marker_blocks=2),
                       dict(quality=98, progressive=True)):
                Image.fromarray(picture(w, h)).save(p, subsampling=sub, **kw)
                assert (b"\xff\xc2" in p.read_bytes()) == ("progressive" in kw)
                want = np.asarray(Image.open(p).convert("RGBA"))
                assert (_image_load(host, p) == want).all(), (w, h, sub, kw)
                cases += 1
        for prog in (False, True):                                                 # grayscale
            Image.fromarray(picture(w, h)[..., 0]).save(p, quality=80, progressive=prog)
            assert (_image_load(host, p) == np.asarray(Image.open(p).convert("RGBA"))).all()
    assert cases == 168
    Image.fromarray(picture(40, 30)).save(p, quality=90)
    assert (_image_load(host, p, flip=1) == np.asarray(Image.open(p).convert("RGBA"))[::-1]).all()
    # the same entry point still decodes PNG
    q = tmp_path / "t.png"
    Image.fromarray(picture(20, 10)).save(q)
    assert (_image_load(host, q) == np.asarray(Image.open(q).convert("RGBA"))).all()


def test_jpeg_decoder_refuses_what_it_does_not_support(host, tmp_path):
    from PIL import Image
    img = np.random.default_rng(1).integers(0, 256, (24, 24, 3), dtype=np.uint8)
    p = tmp_path / "bad.jpg"
    Image.fromarray(img).save(p, progressive=True)
    data = bytearray(p.read_bytes())
    sos = data.index(b"\xff\xda")
    ns = data[sos + 4]
    data[sos + 5 + 2 * ns] = 9                      # Ss > Se in the first scan header
    p.write_bytes(bytes(data))
    with pytest.raises(RuntimeError, match="Could not load image at path: .*bad progression parameters"):
        _image_load(host, p)
    Image.fromarray(img).convert("CMYK").save(p)
    with pytest.raises(RuntimeError, match="only grayscale and 3-component JPEG"):
        _image_load(host, p)
    p.write_bytes(b"GIF89a" + bytes(64))
    with pytest.raises(RuntimeError, match="neither a PNG nor a JPEG"):
        _image_load(host, p)
    with pytest.raises(RuntimeError, match="Could not load image at path"):
        _image_load(host, tmp_path / "absent.jpg")


@pytest.mark.gpu
def test_cli_textured_obj_like_the_python_host(host, rb, tmp_path):
    """--obj + --texture / --normal-map / --bump-map: the C++ importer and PNG decoder feed the same tables as load_obj +
    PIL-decoded, vertically flipped textures on the Python side (albedo, tangent-space normals, parallax height map);
    the rendered frame must be identical."""
    from PIL import Image
    (tmp_path / "asset.obj").write_text(OBJ_FULL)
    rng = np.random.default_rng(21)
    tex = rng.integers(0, 256, (32, 48, 4), dtype=np.uint8)
    tex[..., 3] = 255
    Image.fromarray(tex, "RGBA").save(tmp_path / "tex.png")
    nmap = rb.configs._bumpy_normal_map(32)
    hmap = rb.configs.brick_height_map(32)
    Image.fromarray(nmap, "RGBA").save(tmp_path / "nmap.png")
    Image.fromarray(hmap[..., 0], "L").save(tmp_path / "hmap.png")          # greyscale PNG: expanded to RGBA by the ingest
    cfg = tmp_path / "config.toml"
    cfg.write_text(REFERENCE_SCHEMA.replace("save_on_samples = [64, 256, 1024]", "save_on_samples = []")
                   .replace("save_on_times = [60.0]", "save_on_times = []")
                   + "\n[render]\nwidth = 80\nheight = 60\ncamera_pos = [0.4, 0.9, 2.6]\ncamera_look_at = [0.0, 0.1, 0.0]\n")
    out, pcfile = tmp_path / "final.png", tmp_path / "pc.bin"
    run = subprocess.run([os.path.join(HOST, "reina_b200"), "--config", str(cfg), "--texture", str(tmp_path / "tex.png"),
                          "--normal-map", str(tmp_path / "nmap.png"), "--bump-map", str(tmp_path / "hmap.png"),
                          "--obj", str(tmp_path / "asset.obj"), "--spp", "16", "--out", str(out), "--dump-pc", str(pcfile),
                          "--quiet"], capture_output=True, text=True)
    assert run.returncode == 0, run.stderr + run.stdout

    s = rb.scene.Scene()
    tid = s.defineTexture(np.ascontiguousarray(tex[::-1]))
    nid = s.defineTexture(np.ascontiguousarray(nmap[::-1]))
    hid = s.defineTexture(np.ascontiguousarray(hmap[::-1]))
    s.addObject(rb.meshes.load_obj(str(tmp_path / "asset.obj")), np.eye(4, dtype=np.float32),
                rb.scene.Material(albedo=(0.8, 0.8, 0.8), interpNormals=True, textureID=tid, normalMapID=nid, bumpMapID=hid))
    s.addObject(rb.meshes.cornell_light(), np.eye(4, dtype=np.float32), rb.scene.Material(**rb.configs.LIGHT))
    tables = s.build(require_emitter=True)
    pc = rb.abi.RtPushConsts.from_buffer_copy(pcfile.read_bytes())
    r = rb.Renderer(80, 60, tables, flags=rb.RB200_FLAG_NEE)
    for b in range(2):
        pc.sampleBatch = b
        r.render_batch(pc)
    r.postprocess()
    want = r.read_ldr().copy()
    r.close()
    got = decode_png(out.read_bytes())
    assert (got == want).all()
    assert got[..., :3].max() > 0                              # the frame is not empty


@pytest.mark.gpu
def test_cli_reports_errors(host, tmp_path):
    run = subprocess.run([os.path.join(HOST, "reina_b200"), "--config", str(tmp_path / "absent.toml")],
                         capture_output=True, text=True)
    assert run.returncode == 1 and "cannot open" in run.stderr
