"""Host-side scene layer: the Python mirror of Models::addModel / Scene::addInstance / Instance::computeCDF /
Instances::computeSamplingDataEmissives (src/scene/Models.cpp:24-92, Scene.cpp:27-54, Instance.cpp:15-53,
Instances.cpp:52-114). Checks the flat-table contract the kernels and the oracle consume."""
import ctypes as C

import numpy as np
import pytest


def test_model_ranges_and_offset_indices(rb):
    s = rb.Scene()
    box, light = rb.meshes.cornell_box(), rb.meshes.cornell_light()
    a = s.defineObject(box)
    b = s.defineObject(light)
    ra, rbg = s.modelRanges[a], s.modelRanges[b]
    assert (ra.firstVertex, ra.indexOffset, ra.indexCount) == (0, 0, 12)
    assert (rbg.firstVertex, rbg.indexOffset, rbg.indexCount) == (24, 36, 2)
    s.addInstance(a, np.eye(4, dtype=np.float32), rb.Material(**rb.configs.CORNELL_WALL))
    s.addInstance(b, np.eye(4, dtype=np.float32), rb.Material(**rb.configs.LIGHT))
    t = s.build(require_emitter=True)
    assert t.vertices.shape == (28, 4) and (t.vertices[:, 3] == 1).all()
    assert t.indices.size == 42 and t.indices[36:].min() >= 24        # second model's indices are offset by firstVertex
    assert (t.tbnIndices == t.indices).all() and (t.texIndices == t.indices).all()
    props = np.frombuffer(t.instanceProperties.tobytes(), dtype=np.uint32).reshape(2, 30)
    assert props[1, 0] == 36 and props[1, 7] == 36 and props[1, 8] == 36   # indicesOffset, tbnsIndicesOffset, texIndicesOffset
    assert t.num_triangles() == 14


def test_model_without_uvs_gets_sentinel(rb):
    s = rb.Scene()
    m = rb.meshes.cornell_light()
    m.texCoords = np.zeros((0, 2), np.float32)
    s.addObject(m, np.eye(4, dtype=np.float32), rb.Material(**rb.configs.LIGHT))
    t = s.build()
    props = np.frombuffer(t.instanceProperties.tobytes(), dtype=np.uint32).reshape(1, 30)
    assert props[0, 8] == 0xFFFFFFFF                      # ModelRange.texIndexOffset = -1 (Models.cpp:51)
    assert t.texCoords.size == 2                          # buildBuffers substitutes {0} (Models.cpp:112)


def test_light_cdf(rb):
    wl = rb.configs.cornell(64, 48)
    t = wl.tables
    assert t.numEmissive == 1 and t.cdfInstances.tolist() == [1.0]
    assert t.cdfTriangles.size == 2 and t.cdfTriangles[-1] == 1.0 and 0 < t.cdfTriangles[0] < 1
    em = rb.abi.InstanceData.from_buffer_copy(t.emissive.tobytes())
    area = 0.47 * 0.38
    assert abs(em.area - area) < 1e-5
    lum = 16.0 * (0.2126 + 0.7152 + 0.0722)
    assert abs(em.weight - area * lum) < 1e-3 and abs(t.totalEmissiveWeight - em.weight) < 1e-6
    assert (em.cdfRangeStart, em.cdfRangeEnd) == (0, 1) and em.indexOffset == 36 and em.cullBackface == 1


def test_scaled_instance_area_uses_row_vector_product(rb):
    s = rb.Scene()
    s.addObject(rb.meshes.cornell_light(), rb.camera.scale(2.0), rb.Material(**rb.configs.LIGHT))
    t = s.build(require_emitter=True)
    em = rb.abi.InstanceData.from_buffer_copy(t.emissive.tobytes())
    assert abs(em.area - 4 * 0.47 * 0.38) < 1e-4        # uniform scale: exact either way (Instance.cpp:23)


def test_errors_match_reference_messages(rb):
    s = rb.Scene()
    s.addObject(rb.meshes.cornell_box(), np.eye(4, dtype=np.float32), rb.Material(**rb.configs.CORNELL_WALL))
    with pytest.raises(RuntimeError, match="at least one emissive object"):
        s.build(require_emitter=True)
    with pytest.raises(RuntimeError, match="buffers are already built"):
        s.defineObject(rb.meshes.cornell_light())
    s2 = rb.Scene()
    with pytest.raises(RuntimeError, match="out of range for models"):
        s2.addInstance(3, np.eye(4, dtype=np.float32), rb.Material())


def test_scene_desc_roundtrip(rb):
    wl = rb.configs.small_mixed(32, 24)
    d = wl.tables.desc()
    assert d.numInstances == 6 and d.numTextures == 3 and d.numEmissive == 1
    assert d.numIndices == wl.tables.indices.size and d.numVertices == wl.tables.vertices.shape[0]
    inst0 = d.instances[0]
    assert inst0.triangleCount == 12 and inst0.materialIdx == 0 and inst0.transform[0] == 1.0 and inst0.transform[15] == 1.0
    assert C.sizeof(d) > 0


def test_obj_loader_contract(rb, tmp_path):
    p = tmp_path / "quad.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvn 0 0 1\n"
                 "f 1/1/1 2/2/1 3/3/1 4/4/1\nf 1/1/1 3/3/1 4/4/1\n")
    m = rb.meshes.load_obj(str(p))
    assert m.vertices.shape == (4, 4)                      # identical (v, vt, vn) triples are joined
    assert m.indices.reshape(-1, 3).tolist() == [[0, 1, 2], [0, 2, 3], [0, 2, 3]]    # fan triangulation, file order
    assert np.allclose(m.texCoords[0], [0, 1]) and np.allclose(m.texCoords[2], [1, 0])   # FlipUVs
    tb = m.tbns.reshape(-1, 3, 3)
    assert np.allclose(tb[:, 2], [0, 0, 1]) and np.allclose(np.abs(tb[:, 0]), [1, 0, 0])
    q = tmp_path / "nouv.obj"
    q.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    with pytest.warns(UserWarning, match="no texture coordinates"):
        m2 = rb.meshes.load_obj(str(q))
    assert m2.texCoords.shape[0] == 0 and np.allclose(m2.tbns.reshape(-1, 3, 3)[:, 2], [0, 0, 1])
    assert (m2.tbns.reshape(-1, 3, 3)[:, 0] == 0).all()    # zero tangents, like the reference (Models.cpp:144-152)


def test_procedural_meshes(rb):
    knot = rb.meshes.torus_knot(n_along=200, n_ring=12)
    assert knot.indices.size // 3 == 200 * 12 * 2
    assert rb.meshes.torus_knot.__defaults__[2] * rb.meshes.torus_knot.__defaults__[3] * 2 == 871414   # Stanford dragon's count
    tb = knot.tbns.reshape(-1, 3, 3)
    assert np.isfinite(tb).all() and np.allclose(np.linalg.norm(tb[:, 2], axis=1), 1, atol=1e-4)
    room = rb.meshes.showroom()
    v = room.vertices[:, :3]
    assert v[:, 0].max() <= 1.0001 and v[:, 2].max() <= 1.0001 and v[:, 1].min() >= -1e-6 and v[:, 1].max() > 1.9
    blob = rb.meshes.subdivided_blob(levels=2)
    assert blob.indices.size // 3 == 20 * 16


def test_camera_matrices(rb):
    pc = rb.camera.push_constants(800, 600, (0, 1, 3.9), (0, 1, 0), 40.0)
    iv = np.array(list(pc.invView), np.float32).reshape(4, 4)
    assert np.allclose(iv[3, :3], [0, 1, 3.9]) and np.allclose(iv[2, :3], [0, 0, 1], atol=1e-6)     # camera looks down -z
    ip = np.array(list(pc.invProjection), np.float32).reshape(4, 4)
    t = np.tan(np.radians(40.0) / 2)
    assert abs(ip[1, 1] - t) < 1e-6 and abs(ip[0, 0] - t * 800 / 600) < 1e-6
    assert abs(pc.defocusMultiplier - 0.015) < 1e-9 and pc.maxBounces == 16 and pc.samplesPerPixel == 8
