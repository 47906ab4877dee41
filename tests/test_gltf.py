"""glTF 2.0 / GLB scene import: the Python mirror (reina-vk_b200/gltf.py) and the C++ host
(reina-vk_b200/host/gltf.cpp) of reina::scene::gltf (src/scene/gltf/gltfloader.cpp:72-456) on assets written by
tests/gltf_fixture.py. CPU: what the importer produces (objects, instances, materials, textures, transforms), that
both importers hand identical tables to rb200_scene_create, and the reference's error messages. GPU: the reina_b200
binary renders --gltf to the same frame the Python host gets from the same library."""
import ctypes as C
import importlib
import json
import os
import struct
import subprocess

import numpy as np
import pytest

import gltf_fixture as gf
from test_cpp_host import HOST, REFERENCE_SCHEMA, assert_tables_identical, cpp_tables, decode_png, err, host, py_tables  # noqa: F401


@pytest.fixture(scope="module")
def gl(rb):
    return importlib.import_module("reina-vk_b200.gltf")


@pytest.mark.filterwarnings("ignore:falling back to UV")
def test_glb_import_follows_the_reference_stages(rb, gl, tmp_path):
    path, imgs = gf.build(tmp_path)
    sc = gl.loadScene(path)
    # meshes 0, 1 (two primitives), 2 are used by the default scene; mesh 3 belongs to another scene
    assert len(sc.modelRanges) == 4
    tris = [r.indexCount for r in sc.modelRanges]
    assert tris[0] == 32 and tris[1] + tris[2] == 16 * 10 * 2 - 2 * 16 and tris[3] == 2
    # instances in depth-first node order: floor, ball (2 primitives), ball2 (2 primitives), lamp
    assert [i[2] for i in sc.instancesToCreate] == [0, 1, 2, 1, 2, 3]
    tiles, glass, none_, _, _, lamp = sc.materials
    f32 = lambda *v: tuple(float(np.float32(x)) for x in v)
    assert all(m.materialIdx == 3 and m.interpNormals and m.bumpMapID == -1 for m in sc.materials)
    assert tiles.albedo == f32(0.9, 0.8, 0.7) and tiles.roughness == f32(0.7)[0] and tiles.metallic == f32(0.1)[0]   # 0.95 clamped to 0.7
    assert (tiles.textureID, tiles.normalMapID, tiles.cullBackface) == (0, 1, False)                                 # doubleSided
    assert glass.roughness == f32(0.1)[0] and glass.ior == f32(1.33)[0] and glass.specularTransmission == f32(0.8)[0]  # 0.02 clamped up
    assert glass.cullBackface is False                                       # transmission switches culling off
    assert (none_.albedo, none_.roughness, none_.ior, none_.cullBackface, none_.metallic) == ((1.0, 1.0, 1.0), 0.0, 1.5, False, 0.0)
    assert lamp.emission == f32(12.5, 0.9 * 12.5, 0.8 * 12.5) or np.allclose(lamp.emission, (12.5, 11.25, 10.0), rtol=1e-6)
    assert lamp.metallic == 1.0 and lamp.roughness == f32(0.7)[0]            # glTF defaults 1.0 / 1.0 (clamped)
    # embedded images are not flipped (src/graphics/Image.cpp:28)
    assert (sc.texturesToCreate[0] == imgs["tex0"]).all() and (sc.texturesToCreate[1] == imgs["nmap"]).all()
    # world transform of "ball": root (T * R) * local (T * S), checked against numpy in float64
    a = 0.2
    R = np.array([[np.cos(2 * a), 0, np.sin(2 * a), 0], [0, 1, 0, 0], [-np.sin(2 * a), 0, np.cos(2 * a), 0], [0, 0, 0, 1]])
    T0 = np.eye(4); T0[:3, 3] = (0, 0.1, 0)
    T1 = np.eye(4); T1[:3, 3] = (-0.5, 0.42, 0.1)
    S1 = np.diag([1.0, 1.3, 0.8, 1.0])
    want = (T0 @ R @ T1 @ S1).T.astype(np.float32)        # glm layout m[c][r]
    np.testing.assert_allclose(sc.instancesToCreate[1][3], want, rtol=0, atol=1e-6)
    t = sc.build(require_emitter=True)
    assert t.numEmissive == 1 and t.num_triangles() == 32 + 2 * (tris[1] + tris[2]) + 2
    # de-quantised UVs of the floor: ushort / 65535
    uv = t.texCoords.reshape(-1, 2)[:25]
    assert uv.min() == 0.0 and uv.max() == 1.0 and np.allclose(uv[1], (0.25, 0.0), atol=1e-4)
    # TANGENT w = -1: bitangent = cross(n, t) * w = cross((0,1,0), (1,0,0)) * -1 = (0, 0, 1)
    assert (t.tbns.reshape(-1, 3, 3)[0] == np.array([[1, 0, 0], [0, 0, 1], [0, 1, 0]], np.float32)).all()


@pytest.mark.filterwarnings("ignore:falling back to UV")
def test_external_gltf_equals_glb_except_for_the_flipped_file_texture(rb, gl, tmp_path):
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    p_glb, imgs = gf.build(tmp_path / "a", external=False)
    p_ext, _ = gf.build(tmp_path / "b", external=True)
    a = py_tables(gl.loadScene(p_glb).build(require_emitter=True))
    b = py_tables(gl.loadScene(p_ext).build(require_emitter=True))
    # file textures are flipped vertically on load (stbi_set_flip_vertically_on_load(true), Image.cpp:14), data-URI
    # and buffer-view images are not (:28)
    assert (b["textures"][0] == imgs["tex0"][::-1]).all() and (b["textures"][1] == imgs["nmap"]).all()
    b["textures"][0] = a["textures"][0]
    assert_tables_identical(a, b)


@pytest.mark.filterwarnings("ignore:falling back to UV")
@pytest.mark.parametrize("external", [False, True])
def test_cpp_importer_tables_identical_to_python(host, rb, gl, tmp_path, external):
    path, _ = gf.build(tmp_path, external=external)
    host.rbhost_tables_gltf.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    h = C.c_void_p()
    assert host.rbhost_tables_gltf(path.encode(), 1, C.byref(h)) == 0, err(host)
    assert_tables_identical(cpp_tables(host, rb, h), py_tables(gl.loadScene(path).build(require_emitter=True)))
    host.rbhost_tables_free(h)


@pytest.mark.filterwarnings("ignore:falling back to UV")
def test_several_emitters_give_identical_light_tables_in_both_hosts(host, rb, gl, tmp_path):
    """Instances.cpp:52-114 with more than one emitter: emissive metadata (transform, CDF range, index offset, weight, area,
    cull flag), the concatenated triangle CDF, the instance CDF and totalEmissiveWeight of the C++ host equal the Python
    host's; two of the three emitters are instances of the same primitive."""
    path, _ = gf.build(tmp_path, glowing_glass=True)
    py = gl.loadScene(path).build(require_emitter=True)
    assert py.numEmissive == 3 and py.cdfInstances.size == 3 and py.cdfInstances[-1] == 1.0
    host.rbhost_tables_gltf.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    h = C.c_void_p()
    assert host.rbhost_tables_gltf(path.encode(), 1, C.byref(h)) == 0, err(host)
    assert_tables_identical(cpp_tables(host, rb, h), py_tables(py))
    host.rbhost_tables_free(h)


@pytest.mark.filterwarnings("ignore:falling back to UV")
def test_sparse_accessors_resolve_to_the_dense_scene(host, rb, gl, tmp_path):
    """glTF 2.0 3.6.2.3 (fastgltf resolves sparse accessors for the reference): base elements — or zeros when the accessor
    has no buffer view — with the listed elements replaced. The sparse file must give the tables of the dense one, in
    both importers; malformed index lists are refused."""
    (tmp_path / "d").mkdir()
    (tmp_path / "s").mkdir()
    dense, _ = gf.build(tmp_path / "d")
    sparse, _ = gf.build(tmp_path / "s", sparse=True)
    want = py_tables(gl.loadScene(dense).build(require_emitter=True))
    assert_tables_identical(py_tables(gl.loadScene(sparse).build(require_emitter=True)), want)
    host.rbhost_tables_gltf.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    h = C.c_void_p()
    assert host.rbhost_tables_gltf(sparse.encode(), 1, C.byref(h)) == 0, err(host)
    assert_tables_identical(cpp_tables(host, rb, h), want)
    host.rbhost_tables_free(h)
    # indices out of order
    (tmp_path / "e").mkdir()
    ext, _ = gf.build(tmp_path / "e", external=True, sparse=True)
    doc = json.loads(open(ext).read())
    acc = [a for a in doc["accessors"] if "sparse" in a and "bufferView" in a][0]
    bv = doc["bufferViews"][acc["sparse"]["indices"]["bufferView"]]
    raw = bytearray((tmp_path / "e" / "scene.bin").read_bytes())
    raw[bv["byteOffset"]:bv["byteOffset"] + 4] = raw[bv["byteOffset"] + 2:bv["byteOffset"] + 4] + raw[bv["byteOffset"]:bv["byteOffset"] + 2]
    (tmp_path / "e" / "scene.bin").write_bytes(bytes(raw))
    with pytest.raises(RuntimeError, match="bad sparse accessor"):
        gl.loadScene(ext)
    assert host.rbhost_tables_gltf(ext.encode(), 1, C.byref(h)) != 0 and "bad sparse accessor" in err(host)


@pytest.mark.filterwarnings("ignore:falling back to UV")
def test_jpeg_textures_give_identical_tables_too(host, rb, gl, tmp_path):
    """Most real .glb files embed JPEG images: the C++ host decodes baseline JPEG itself (host/jpeg.cpp, the IJG integer
    pipeline) and must hand over the same texels PIL gives the Python host."""
    path, _ = gf.build(tmp_path, external=True, jpeg=True)
    host.rbhost_tables_gltf.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    h = C.c_void_p()
    assert host.rbhost_tables_gltf(path.encode(), 1, C.byref(h)) == 0, err(host)
    py = py_tables(gl.loadScene(path).build(require_emitter=True))
    assert py["textures"][0].shape == (16, 24, 4) and py["textures"][0][..., :3].std() > 10
    assert_tables_identical(cpp_tables(host, rb, h), py)
    host.rbhost_tables_free(h)


def _write_gltf(tmp_path, doc, name="bad.gltf"):
    p = tmp_path / name
    p.write_text(json.dumps(doc))
    return str(p)


def test_errors_match_the_reference_messages(host, rb, gl, tmp_path):
    host.rbhost_tables_gltf.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]

    def both(path, needle):
        with pytest.raises(RuntimeError, match=needle):
            gl.loadScene(path).build()
        h = C.c_void_p()
        assert host.rbhost_tables_gltf(path.encode(), 0, C.byref(h)) != 0 and needle in err(host), err(host)
    both(str(tmp_path / "absent.glb"), "Failed to find glTF file")
    both(_write_gltf(tmp_path, {"asset": {"version": "2.0"}}), "No scenes supplied in gLTF file")
    both(_write_gltf(tmp_path, {"asset": {"version": "2.0"}, "extensionsRequired": ["KHR_mesh_quantization", "KHR_draco_mesh_compression"],
                                "scenes": [{"nodes": []}]}, "draco.gltf"), "required extension KHR_draco_mesh_compression is not enabled")
    (tmp_path / "junk.gltf").write_text("{ not json")
    both(str(tmp_path / "junk.gltf"), "Failed to parse glTF")
    tri = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    import base64
    uri = "data:application/octet-stream;base64," + base64.b64encode(tri.tobytes()).decode()
    doc = {"asset": {"version": "2.0"}, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
           "meshes": [{"primitives": [{"attributes": {"POSITION": 0}}]}],
           "buffers": [{"byteLength": 36, "uri": uri}], "bufferViews": [{"buffer": 0, "byteLength": 36}],
           "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"}]}
    both(_write_gltf(tmp_path, doc, "nonormal.gltf"), "Meshes without vertex normals are not supported")
    doc["meshes"][0]["primitives"][0]["attributes"]["NORMAL"] = 0
    doc["meshes"][0]["primitives"][0]["mode"] = 1
    both(_write_gltf(tmp_path, doc, "lines.gltf"), "only TRIANGLES primitives are supported")
    doc["meshes"][0]["primitives"][0]["mode"] = 4
    doc["accessors"][0]["count"] = 4
    both(_write_gltf(tmp_path, doc, "overrun.gltf"), "reads past its buffer view")
    doc["accessors"][0]["count"] = 3
    doc["images"] = [{"uri": "data:image/gif;base64," + base64.b64encode(b"GIF89a" + b"0" * 32).decode()}]
    both(_write_gltf(tmp_path, doc, "gif.gltf"), "neither a PNG nor a JPEG file")
    # a truncated GLB
    glb, _ = gf.build(tmp_path)
    raw = open(glb, "rb").read()
    (tmp_path / "cut.glb").write_bytes(raw[: len(raw) // 2])
    both(str(tmp_path / "cut.glb"), "Failed to parse glTF")
    (tmp_path / "v1.glb").write_bytes(b"glTF" + struct.pack("<II", 1, 20) + b"\0" * 8)
    both(str(tmp_path / "v1.glb"), "unsupported GLB version")


def test_hostile_numeric_fields_are_refused_not_trusted(host, rb, gl, tmp_path):
    """Counts, offsets, strides and lengths are untrusted JSON numbers: negative or huge values (which would wrap the
    importer's size_t address arithmetic — count = -2^63 once passed the bounds test and wrote out of bounds) must be
    refused with an error by the C++ importer, whatever field carries them; the Python importer must raise as well."""
    import base64
    host.rbhost_tables_gltf.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    tri = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    uv = np.zeros((3, 2), np.float32)
    blob = tri.tobytes() + uv.tobytes()
    uri = "data:application/octet-stream;base64," + base64.b64encode(blob).decode()

    def doc():
        return {"asset": {"version": "2.0"}, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
                "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "NORMAL": 0, "TEXCOORD_0": 1}}]}],
                "buffers": [{"byteLength": len(blob), "uri": uri}],
                "bufferViews": [{"buffer": 0, "byteLength": 36}, {"buffer": 0, "byteOffset": 36, "byteLength": 24}],
                "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"},
                              {"bufferView": 1, "componentType": 5126, "count": 3, "type": "VEC2"}]}
    h = C.c_void_p()
    good = _write_gltf(tmp_path, doc(), "good.gltf")
    assert host.rbhost_tables_gltf(good.encode(), 0, C.byref(h)) == 0, err(host)
    host.rbhost_tables_free(h)
    hostile = [-1, -2 ** 63, 2 ** 63 + 1024, 2 ** 32, 2 ** 31, 2 ** 62, -2 ** 31]
    cases = 0
    for where, field in (("accessors", "count"), ("accessors", "byteOffset"), ("bufferViews", "byteOffset"),
                         ("bufferViews", "byteStride"), ("bufferViews", "byteLength"), ("accessors", "bufferView")):
        for idx in (0, 1):
            for v in hostile:
                d = doc()
                d[where][idx][field] = v
                path = _write_gltf(tmp_path, d, "hostile.gltf")
                rc = host.rbhost_tables_gltf(path.encode(), 0, C.byref(h))
                if rc == 0:      # only harmless values may load (none of these are)
                    host.rbhost_tables_free(h)
                assert rc != 0 and "Failed to parse glTF" in err(host), (where, field, idx, v, err(host))
                with pytest.raises(Exception):
                    gl.loadScene(path).build()
                cases += 1
    assert cases == 84


@pytest.mark.filterwarnings("ignore:falling back to UV")
def test_oracle_renders_the_imported_scene(ol, rb, gl, tmp_path):
    """The imported tables are a valid scene for the path tracer: finite image, light reaches the floor."""
    path, _ = gf.build(tmp_path)
    t = gl.loadScene(path).build(require_emitter=True)
    pc = rb.camera.push_constants(64, 48, (0.0, 1.3, 3.4), (0.0, 0.4, 0.0), 40.0, total_emissive_weight=t.totalEmissiveWeight,
                                  samples_per_pixel=4, max_bounces=5)
    hdr, cnt = ol.OracleScene(t).render_batch(64, 48, rb.RB200_FLAG_NEE, pc)
    assert np.isfinite(hdr).all() and hdr[30:, :, :3].mean() > 0.01 and cnt["shadowRays"] > 0


@pytest.mark.gpu
@pytest.mark.filterwarnings("ignore:falling back to UV")
def test_cli_renders_glb_like_the_python_host(host, ol, rb, gl, tmp_path):
    """reina_b200 --gltf scene.glb: the frame equals the one the Python host renders from its own import of the same
    file through the same library, and the HDR image of that scene is bit-identical to the oracle's."""
    path, _ = gf.build(tmp_path)
    cfg = tmp_path / "config.toml"
    cfg.write_text(REFERENCE_SCHEMA.replace("save_on_samples = [64, 256, 1024]", "save_on_samples = []")
                   .replace("save_on_times = [60.0]", "save_on_times = []")
                   + "\n[render]\nwidth = 96\nheight = 72\ncamera_pos = [0.0, 1.3, 3.4]\ncamera_look_at = [0.0, 0.4, 0.0]\n")
    out, pcfile = tmp_path / "final.png", tmp_path / "pc.bin"
    run = subprocess.run([os.path.join(HOST, "reina_b200"), "--config", str(cfg), "--gltf", path, "--spp", "16", "--out", str(out),
                          "--dump-pc", str(pcfile), "--quiet"], capture_output=True, text=True)
    assert run.returncode == 0, run.stderr + run.stdout
    tables = gl.loadScene(path).build(require_emitter=True)
    pc = rb.abi.RtPushConsts.from_buffer_copy(pcfile.read_bytes())
    r = rb.Renderer(96, 72, tables, flags=rb.RB200_FLAG_NEE)
    sc = ol.OracleScene(tables)
    hdr_o = np.zeros((72, 96, 4), np.float32)
    for b in range(2):
        pc.sampleBatch = b
        r.render_batch(pc)
        hdr_o, _ = sc.render_batch(96, 72, rb.RB200_FLAG_NEE, pc, hdr_o)
    assert (r.read_hdr().view(np.uint32) == hdr_o.view(np.uint32)).all()
    r.postprocess()
    want = r.read_ldr().copy()
    r.close()
    got = decode_png(out.read_bytes())
    assert (got == want).all() and got[..., :3].max() > 0


def _emitters_gltf(tmp_path, order):
    """A flat scene of quads, one node per letter: A = lamp (instances of ONE mesh, translated: equal CDFs), D = lamp of
    another shape (unequal triangles: a different CDF), F = non-emissive floor."""
    B = gf._Bin()
    quad = np.array([[-0.4, 0, -0.4], [0.4, 0, 0.4], [0.4, 0, -0.4], [-0.4, 0, -0.4], [-0.4, 0, 0.4], [0.4, 0, 0.4]], np.float32)
    odd = np.array([[-0.4, 0, -0.4], [0.0, 0, 0.4], [0.4, 0, -0.4], [-0.4, 0, -0.4], [-0.4, 0, 0.4], [0.0, 0, 0.4]], np.float32)
    nrm = np.tile(np.array([[0, -1, 0]], np.float32), (6, 1))
    a_n = B.accessor(B.view(nrm.tobytes()), 5126, 6, "VEC3")
    a_quad = B.accessor(B.view(quad.tobytes()), 5126, 6, "VEC3", minmax=(quad.min(0), quad.max(0)))
    a_odd = B.accessor(B.view(odd.tobytes()), 5126, 6, "VEC3", minmax=(odd.min(0), odd.max(0)))
    lamp = {"emissiveFactor": [1.0, 0.9, 0.8], "doubleSided": True}
    doc = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": list(range(len(order)))}],
           "nodes": [{"mesh": "AFD".index(c), "translation": [1.0 * i, 1.5 if c != "F" else 0.0, 0.0]} for i, c in enumerate(order)],
           "meshes": [{"primitives": [{"attributes": {"POSITION": a_quad, "NORMAL": a_n}, "material": 0}]},
                      {"primitives": [{"attributes": {"POSITION": a_quad, "NORMAL": a_n}, "material": 1}]},
                      {"primitives": [{"attributes": {"POSITION": a_odd, "NORMAL": a_n}, "material": 0}]}],
           "materials": [lamp, {"pbrMetallicRoughness": {"baseColorFactor": [0.8, 0.8, 0.8, 1.0]}}],
           "accessors": B.accessors, "bufferViews": B.views}
    while len(B.data) % 4:
        B.data.append(0)
    doc["buffers"] = [{"byteLength": len(B.data)}]
    js = json.dumps(doc, separators=(",", ":")).encode()
    js += b" " * (-len(js) % 4)
    path = tmp_path / ("emitters_%s.glb" % order)
    path.write_bytes(struct.pack("<III", 0x46546C67, 2, 12 + 8 + len(js) + 8 + len(B.data)) + struct.pack("<II", len(js), 0x4E4F534A) + js +
                     struct.pack("<II", len(B.data), 0x004E4942) + bytes(B.data))
    return str(path)


@pytest.mark.filterwarnings("ignore:falling back to UV")
@pytest.mark.filterwarnings("ignore:.*several emitters")
def test_emissive_cdf_sharing_is_reproduced_as_written(host, rb, gl, tmp_path):
    """Instances::computeEmissiveDuplicates / computeSamplingDataEmissives (src/scene/Instances.cpp:27-50, 52-114): emitters
    with equal triangle CDFs share one range of cdfTriangles — but the map is keyed by instance index and consulted with
    the position in the emissive list (:72 vs :86), so what happens depends on where the emitters sit among the instances.
    Both hosts reproduce every outcome: sharing (AAF, AAAF), no sharing (FAA), another emitter's range (FAAA), and the
    std::out_of_range upstream dies with (FAAD)."""
    abi = importlib.import_module("reina-vk_b200.abi")
    host.rbhost_tables_gltf.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]

    def ranges(order):
        path = _emitters_gltf(tmp_path, order)
        py = gl.loadScene(path).build(require_emitter=True)
        h = C.c_void_p()
        assert host.rbhost_tables_gltf(path.encode(), 1, C.byref(h)) == 0, err(host)
        assert_tables_identical(cpp_tables(host, rb, h), py_tables(py))
        host.rbhost_tables_free(h)
        recs = (abi.InstanceData * py.numEmissive).from_buffer_copy(py.emissive.tobytes())
        return [(r.cdfRangeStart, r.cdfRangeEnd) for r in recs], py

    r, py = ranges("AAF")          # instance index == emissive position: the second lamp shares the first one's CDF
    assert r == [(0, 1), (0, 1)] and py.cdfTriangles.size == 2
    assert np.array_equal(py.cdfInstances, np.array([0.5, 1.0], np.float32))
    r, py = ranges("AAAF")
    assert r == [(0, 1)] * 3 and py.cdfTriangles.size == 2
    r, py = ranges("FAA")          # keys {2}, asked for positions 0 and 1: nothing is shared
    assert r == [(0, 1), (2, 3)] and py.cdfTriangles.size == 4
    r, py = ranges("FAAA")         # keys {2, 3}: position 2 (instance 3) is a key, at(3) = 1 -> emissive record 1 (instance 2)
    assert r == [(0, 1), (2, 3), (2, 3)] and py.cdfTriangles.size == 4
    r, py = ranges("AFA")          # keys {2}: position 1 (instance 2) is not asked about with its own index
    assert r == [(0, 1), (2, 3)]
    # FAAD: keys {2}; position 2 is instance 3 (the odd lamp): contains(2) is true, at(3) throws
    path = _emitters_gltf(tmp_path, "FAAD")
    with pytest.raises(IndexError, match="unordered_map::at"):
        gl.loadScene(path).build(require_emitter=True)
    h = C.c_void_p()
    assert host.rbhost_tables_gltf(path.encode(), 1, C.byref(h)) != 0 and "unordered_map::at" in err(host)


@pytest.mark.filterwarnings("ignore:falling back to UV")
def test_missing_tangents_follow_the_reference_call_as_written(host, rb, gl, tmp_path):
    """gltfloader.cpp:207-222 hands MikkTSpace the vertex array as un-indexed triples whose UVs are still all zero
    (TEXCOORD_0 is read at :225-249). For that input the algorithm opens no group and every corner keeps its initial
    tangent space (gltf.mikktspace_as_called): tangent (1,0,0), bitangent (0,1,0) for the vertices of complete triples,
    zeros for the left-over ones; the normal row is the vertex normal. Default rule of both importers; "uv" is the
    importer's own rule and both importers agree on it too."""
    t, b = gl.mikktspace_as_called(8)
    assert (t[:6] == (1, 0, 0)).all() and (b[:6] == (0, 1, 0)).all() and not t[6:].any() and not b[6:].any()
    t, b = gl.mikktspace_as_called(2)                   # fewer than three vertices: genTangSpace returns before writing
    assert not t.any() and not b.any()

    path, _ = gf.build(tmp_path)
    asset = gl.loadGltf(path)
    doc_prims = [p for mi in (0, 1, 2) for p in asset.doc["meshes"][mi]["primitives"]]
    prims = [p for _, ps in sorted(gl.loadPrimitives(asset).items()) for p in ps]
    uvprims = [p for _, ps in sorted(gl.loadPrimitives(asset, tangents="uv").items()) for p in ps]
    seen = 0
    for dp, p, q in zip(doc_prims, prims, uvprims):
        if "TANGENT" in dp["attributes"]:
            assert (p.tangent == q.tangent).all() and (p.bitangent == q.bitangent).all()
            continue
        seen += 1
        n = p.position.shape[0]
        full = 3 * (n // 3)
        assert (p.tangent[:full] == np.float32([1, 0, 0])).all() and (p.bitangent[:full] == np.float32([0, 1, 0])).all()
        assert not p.tangent[full:].any() and not p.bitangent[full:].any()
        tbn = p.toModelData().tbns.reshape(-1, 3, 3)
        assert (tbn[:, 2, :] == p.normal).all() and (tbn[:full, 0, :] == (1, 0, 0)).all()
        assert np.abs(q.tangent - p.tangent).max() > 0.1        # the UV rule really is a different frame
    assert seen >= 2
    with pytest.raises(ValueError):
        gl.loadPrimitives(asset, tangents="mikk")

    host.rbhost_tables_gltf_tangents.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    for rule, name in ((0, "reference"), (1, "uv")):
        h = C.c_void_p()
        assert host.rbhost_tables_gltf_tangents(path.encode(), 1, rule, C.byref(h)) == 0, err(host)
        assert_tables_identical(cpp_tables(host, rb, h), py_tables(gl.loadScene(path, tangents=name).build(require_emitter=True)))
        host.rbhost_tables_free(h)
    h = C.c_void_p()
    assert host.rbhost_tables_gltf_tangents(path.encode(), 1, 2, C.byref(h)) != 0
