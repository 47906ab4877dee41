"""The reference's own assets through this repository's importers (CPU only, in the build container: /root/reference does
not exist on the GPU box, so everything here is skipped there and nothing under -m gpu touches it).
  * every texture the reference ships (13 PNG of five colour types, 2 baseline JPEG at 2048x2048) decoded by the C++
    host and compared byte for byte with PIL, flipped vertically as reina::graphics::Image's file constructor does;
  * every OBJ the reference ships imported by the C++ and the Python importer: identical tables;
  * the measured facts SURVEY.md quotes about those assets (triangle counts, the alpha channel of
    Hammered_Metal_Albedo.png)."""
import ctypes as C
import glob
import os
import warnings

import numpy as np
import pytest

from test_cpp_host import assert_tables_identical, cpp_tables, err, host, py_tables  # noqa: F401

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "textures")), reason="reference checkout not present")


def _load(host, path, flip):
    host.rbhost_image_load.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    buf = np.zeros(2048 * 2048 * 4, np.uint8)
    w, h = C.c_uint32(), C.c_uint32()
    assert host.rbhost_image_load(path.encode(), flip, buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(w), C.byref(h)) == 0, err(host)
    return buf[: w.value * h.value * 4].reshape(h.value, w.value, 4)


def test_every_reference_texture_decodes_like_pil(host):
    from PIL import Image
    files = sorted(glob.glob(os.path.join(REF, "textures", "**", "*.*"), recursive=True))
    assert len(files) == 15 and sum(f.endswith(".jpg") for f in files) == 2
    for f in files:
        want = np.asarray(Image.open(f).convert("RGBA"))[::-1]          # stbi_set_flip_vertically_on_load(true)
        assert (_load(host, f, 1) == want).all(), f


def test_hammered_metal_albedo_alpha_is_mostly_zero(host):
    """binding this texture as textureID makes the stochastic alpha test skip most hits."""
    a = _load(host, os.path.join(REF, "textures", "Hammered_Metal_Albedo.png"), 1)[..., 3]
    assert set(np.unique(a)) <= {0, 255} and abs((a == 0).mean() - 0.683) < 0.005


@pytest.mark.parametrize("name,triangles", [("cornell_box.obj", 12), ("cornell_light.obj", 2), ("empty_cornell_box.obj", 12),
                                            ("plant_pot.obj", 1408), ("plant_soil.obj", 240), ("plant_leaves_1.obj", 32808),
                                            ("plant_leaves_2.obj", 32808), ("showroom.obj", 1176), ("lowpoly_suzanne.obj", 319),
                                            ("uv_sphere.obj", 960), ("uv_sphere_highres.obj", 16128),
                                            ("ico_sphere_highres.obj", 5120), ("quad.obj", 2), ("blender_cube.obj", 12)])
def test_reference_obj_imports_identically_in_both_hosts(host, rb, name, triangles):
    path = os.path.join(REF, "models", name)
    h = C.c_void_p()
    assert host.rbhost_tables_obj(path.encode(), 0, 0, C.byref(h)) == 0, err(host)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        md = rb.meshes.load_obj(path)
    s = rb.scene.Scene()
    s.addObject(md, np.eye(4, dtype=np.float32), rb.scene.Material(albedo=(0.8, 0.8, 0.8), interpNormals=True))
    t = s.build()
    assert_tables_identical(cpp_tables(host, rb, h), py_tables(t))
    host.rbhost_tables_free(h)
    if triangles is not None:
        assert t.num_triangles() == triangles


def test_reference_plant_scene_renders_through_the_importers_and_the_oracle(ol, rb):
    """BASELINE config C4 with the reference's real assets: the four plant OBJs (67,264 triangles) in showroom.obj under
    cornell_light.obj, leaves with the 2K JPEG albedo + normal textures, pot with the hammered-metal normal map — host
    importers, table builder and the oracle end to end (the GPU box cannot see these files; there C4 runs on the
    stand-in geometry of configs.plant)."""
    from PIL import Image
    m = lambda n: rb.meshes.load_obj(os.path.join(REF, "models", n))
    tex = lambda n: np.ascontiguousarray(np.asarray(Image.open(os.path.join(REF, "textures", n)).convert("RGBA"))[::-1])
    s = rb.Scene()
    leaf_albedo, leaf_normal, pot_normal = (s.defineTexture(tex(n)) for n in ("qgdpH_2K_Albedo.jpg", "qgdpH_2K_Normal.jpg", "Hammered_Metal_normal.png"))
    I = np.eye(4, dtype=np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        s.addObject(m("showroom.obj"), I, rb.Material(materialIdx=0, albedo=(0.8, 0.8, 0.8), interpNormals=True))
        s.addObject(m("cornell_light.obj"), I, rb.Material(**rb.configs.LIGHT))
        s.addObject(m("plant_pot.obj"), I, rb.Material(materialIdx=3, albedo=(0.72, 0.45, 0.2), roughness=0.4, ior=1.5, interpNormals=True,
                                                       metallic=0.6, normalMapID=pot_normal, sheenTint=(1, 1, 1), specularTint=(1, 1, 1)))
        s.addObject(m("plant_soil.obj"), I, rb.Material(materialIdx=0, albedo=(0.25, 0.18, 0.12), interpNormals=True))
        for n in ("plant_leaves_1.obj", "plant_leaves_2.obj"):
            s.addObject(m(n), I, rb.Material(materialIdx=0, albedo=(0.9, 0.9, 0.9), textureID=leaf_albedo, normalMapID=leaf_normal, interpNormals=True))
    t = s.build(require_emitter=True)
    assert t.num_triangles() == 1176 + 2 + 1408 + 240 + 32808 + 32808
    W, H = 96, 72
    pc = rb.camera.push_constants(W, H, (-1.6899, 0.817, -1.6386), (0.0, 0.55, 0.0), 30.0, total_emissive_weight=t.totalEmissiveWeight,
                                  samples_per_pixel=2, max_bounces=6)
    sc = ol.OracleScene(t)
    hdr, cnt = sc.render_batch(W, H, rb.RB200_FLAG_NEE, pc)
    assert np.isfinite(hdr).all() and cnt["paths"] == W * H * 2 and cnt["shadowRays"] > 0
    hits = sc.trace_primary(W, H, pc)
    seen = set(np.unique(hits["instance"][hits["t"] > 0]).tolist())
    assert {0, 2, 4}.issubset(seen) or {0, 2, 5}.issubset(seen)           # showroom, pot and leaves are in view
    assert hdr[..., :3].mean() > 1e-3
