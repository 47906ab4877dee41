"""Builds small glTF 2.0 assets for the importer tests (tests/test_gltf.py): a .glb with everything embedded and the
same scene as .gltf + external .bin + external / data-URI PNGs. Exercises: node hierarchy with TRS and matrix
transforms, a mesh instanced twice, TANGENT (VEC4, handedness -1) and missing TANGENT, interleaved buffer views
(byteStride), normalized UNSIGNED_SHORT texture coordinates and UNSIGNED_BYTE / UNSIGNED_SHORT / absent indices
(KHR_mesh_quantization-style accessors), materials with emissive strength, transmission + ior, base colour and normal
textures, a material with an emissive texture (emission ignored), a primitive without material, an unused mesh."""
import base64
import io
import json
import struct

import numpy as np


def _png_bytes(img):
    from PIL import Image
    b = io.BytesIO()
    Image.fromarray(img, "RGBA").save(b, format="PNG")
    return b.getvalue()


def _grid(n, size, y=0.0):
    """(n+1)^2-vertex square in the xz plane facing +y, with UVs and analytic tangents."""
    g = np.linspace(-size, size, n + 1, dtype=np.float32)
    xs, zs = np.meshgrid(g, g)
    pos = np.stack([xs.ravel(), np.full(xs.size, y, np.float32), zs.ravel()], 1).astype(np.float32)
    uv = np.stack([(xs.ravel() + size) / (2 * size), (zs.ravel() + size) / (2 * size)], 1).astype(np.float32)
    nrm = np.tile(np.array([[0, 1, 0]], np.float32), (pos.shape[0], 1))
    tan = np.tile(np.array([[1, 0, 0, -1]], np.float32), (pos.shape[0], 1))
    tris = []
    for j in range(n):
        for i in range(n):
            a = j * (n + 1) + i
            tris += [(a, a + n + 1, a + 1), (a + 1, a + n + 1, a + n + 2)]
    return pos, uv, nrm, tan, np.array(tris, np.uint32)


def _sphere(seg, rings, radius):
    pos, uv, nrm = [], [], []
    for r in range(rings + 1):
        th = np.pi * r / rings
        for s in range(seg + 1):
            ph = 2 * np.pi * s / seg
            n = np.array([np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)])
            pos.append(n * radius); nrm.append(n); uv.append((s / seg, r / rings))
    tris = []
    w = seg + 1
    for r in range(rings):
        for s in range(seg):
            a, b, c, d = r * w + s, r * w + s + 1, (r + 1) * w + s + 1, (r + 1) * w + s
            if r != 0:
                tris.append((a, b, c))
            if r != rings - 1:
                tris.append((a, c, d))
    return (np.array(pos, np.float32), np.array(uv, np.float32), np.array(nrm, np.float32), np.array(tris, np.uint32))


class _Bin:
    def __init__(self):
        self.data = bytearray()
        self.views, self.accessors = [], []

    def view(self, raw, stride=None):
        while len(self.data) % 4:
            self.data.append(0)
        off = len(self.data)
        self.data += raw
        v = {"buffer": 0, "byteOffset": off, "byteLength": len(raw)}
        if stride:
            v["byteStride"] = stride
        self.views.append(v)
        return len(self.views) - 1

    def accessor(self, view, ctype, count, typ, offset=0, normalized=False, minmax=None):
        a = {"bufferView": view, "componentType": ctype, "count": int(count), "type": typ}
        if offset:
            a["byteOffset"] = offset
        if normalized:
            a["normalized"] = True
        if minmax is not None:
            a["min"], a["max"] = [float(x) for x in minmax[0]], [float(x) for x in minmax[1]]
        self.accessors.append(a)
        return len(self.accessors) - 1


def _jpeg_bytes(img, **kw):
    from PIL import Image
    b = io.BytesIO()
    Image.fromarray(img[..., :3], "RGB").save(b, format="JPEG", **kw)
    return b.getvalue()


def build(tmp_path, external=False, jpeg=False, sparse=False, glowing_glass=False):
    """Writes scene.glb (external=False) or scene.gltf + scene.bin + tex0.png (external=True) and returns the path.
    sparse=True stores the same scene through sparse accessors: the sphere's POSITION over a perturbed base array, the
    lamp's NORMAL over no buffer view at all (zeros + substitutions). glowing_glass=True makes the glass material emissive
    as well: three emitters, two of them instances of the same primitive."""
    rng = np.random.RandomState(4)
    B = _Bin()
    # mesh 0: floor grid — interleaved POSITION/NORMAL (stride 24), TANGENT vec4, normalized ushort UVs, ubyte indices
    pos, uv, nrm, tan, tris = _grid(4, 1.5)
    inter = np.concatenate([pos, nrm], 1).astype(np.float32).tobytes()
    v_int = B.view(inter, stride=24)
    a_pos = B.accessor(v_int, 5126, pos.shape[0], "VEC3", minmax=(pos.min(0), pos.max(0)))
    a_nrm = B.accessor(v_int, 5126, pos.shape[0], "VEC3", offset=12)
    a_tan = B.accessor(B.view(tan.tobytes()), 5126, pos.shape[0], "VEC4")
    uvq = np.round(uv * 65535).astype(np.uint16)
    a_uv = B.accessor(B.view(uvq.tobytes()), 5123, pos.shape[0], "VEC2", normalized=True)
    a_idx = B.accessor(B.view(tris.astype(np.uint8).tobytes()), 5121, tris.size, "SCALAR")
    # mesh 1: sphere, two primitives (upper / lower half) — no TANGENT, float UVs, ushort indices; the second
    # primitive has no material
    sp, suv, sn, st = _sphere(16, 10, 0.3)
    upper = st[: st.shape[0] // 2]
    lower = st[st.shape[0] // 2:]
    if sparse:
        k = np.array([1, 5, 17, 40, sp.shape[0] - 1], np.int64)
        base = sp.copy()
        base[k] += 7.0
        a_spos = B.accessor(B.view(base.tobytes()), 5126, sp.shape[0], "VEC3", minmax=(sp.min(0), sp.max(0)))
        B.accessors[a_spos]["sparse"] = {"count": len(k), "indices": {"bufferView": B.view(k.astype(np.uint16).tobytes()), "componentType": 5123},
                                         "values": {"bufferView": B.view(sp[k].tobytes())}}
    else:
        a_spos = B.accessor(B.view(sp.tobytes()), 5126, sp.shape[0], "VEC3", minmax=(sp.min(0), sp.max(0)))
    a_snrm = B.accessor(B.view(sn.tobytes()), 5126, sp.shape[0], "VEC3")
    a_suv = B.accessor(B.view(suv.tobytes()), 5126, sp.shape[0], "VEC2")
    a_si0 = B.accessor(B.view(upper.astype(np.uint16).tobytes()), 5123, upper.size, "SCALAR")
    a_si1 = B.accessor(B.view(lower.astype(np.uint16).tobytes()), 5123, lower.size, "SCALAR")
    # mesh 2: emissive quad as an un-indexed triangle list (indices are generated)
    q = np.array([[-0.4, 0, -0.4], [0.4, 0, 0.4], [0.4, 0, -0.4], [-0.4, 0, -0.4], [-0.4, 0, 0.4], [0.4, 0, 0.4]], np.float32)
    qn = np.tile(np.array([[0, -1, 0]], np.float32), (6, 1))
    a_qpos = B.accessor(B.view(q.tobytes()), 5126, 6, "VEC3", minmax=(q.min(0), q.max(0)))
    if sparse:
        pad = B.view(bytes(8))                                   # values start at a byteOffset inside their view
        B.accessors.append({"componentType": 5126, "count": 6, "type": "VEC3",
                            "sparse": {"count": 6, "indices": {"bufferView": B.view(np.arange(6, dtype=np.uint8).tobytes()), "componentType": 5121},
                                       "values": {"bufferView": B.view(bytes(12) + qn.tobytes()), "byteOffset": 12}}})
        a_qnrm = len(B.accessors) - 1
        assert pad >= 0
    else:
        a_qnrm = B.accessor(B.view(qn.tobytes()), 5126, 6, "VEC3")
    # mesh 3: never referenced by the default scene
    a_upos = B.accessor(B.view(q.tobytes()), 5126, 6, "VEC3", minmax=(q.min(0), q.max(0)))

    tex0 = rng.randint(0, 256, (16, 24, 4)).astype(np.uint8); tex0[..., 3] = 255
    y, x = np.mgrid[0:32, 0:32].astype(np.float32) / 32
    nmap = np.zeros((32, 32, 4), np.uint8)
    nmap[..., 0] = 128 + 60 * np.sin(x * 12); nmap[..., 1] = 128 + 60 * np.cos(y * 12); nmap[..., 2] = 230; nmap[..., 3] = 255
    images = []
    if external and jpeg:       # a JPEG file (4:2:0, flipped on load) and a progressive data-URI JPEG (4:4:4, not flipped)
        (tmp_path / "tex0.jpg").write_bytes(_jpeg_bytes(tex0, quality=85, subsampling=2))
        images.append({"uri": "tex0.jpg"})
        images.append({"uri": "data:image/jpeg;base64," + base64.b64encode(_jpeg_bytes(nmap, quality=95, subsampling=0, progressive=True)).decode()})
    elif external:
        (tmp_path / "tex0.png").write_bytes(_png_bytes(tex0))
        images.append({"uri": "tex0.png"})
        images.append({"uri": "data:image/png;base64," + base64.b64encode(_png_bytes(nmap)).decode()})
    else:
        images.append({"bufferView": B.view(_png_bytes(tex0)), "mimeType": "image/png"})
        images.append({"bufferView": B.view(_png_bytes(nmap)), "mimeType": "image/png"})

    c, s = np.cos(0.3), np.sin(0.3)
    doc = {
        "asset": {"version": "2.0"},
        "extensionsUsed": ["KHR_materials_emissive_strength", "KHR_materials_transmission", "KHR_materials_ior"],
        "scene": 0,
        "scenes": [{"nodes": [0, 4]}, {"nodes": [5]}],
        "nodes": [
            {"name": "root", "translation": [0.0, 0.1, 0.0], "rotation": [0.0, float(np.sin(0.2)), 0.0, float(np.cos(0.2))],
             "children": [1, 2, 3]},
            {"name": "floor", "mesh": 0},
            {"name": "ball", "mesh": 1, "translation": [-0.5, 0.42, 0.1], "scale": [1.0, 1.3, 0.8]},
            {"name": "ball2", "mesh": 1, "matrix": [c, 0, -s, 0, 0, 1, 0, 0, s, 0, c, 0, 0.55, 0.35, -0.2, 1]},
            {"name": "lamp", "mesh": 2, "translation": [0.0, 1.9, 0.0]},
            {"name": "other scene", "mesh": 3},
        ],
        "meshes": [
            {"primitives": [{"attributes": {"POSITION": a_pos, "NORMAL": a_nrm, "TANGENT": a_tan, "TEXCOORD_0": a_uv},
                             "indices": a_idx, "material": 0}]},
            {"primitives": [{"attributes": {"POSITION": a_spos, "NORMAL": a_snrm, "TEXCOORD_0": a_suv}, "indices": a_si0, "material": 1},
                            {"attributes": {"POSITION": a_spos, "NORMAL": a_snrm, "TEXCOORD_0": a_suv}, "indices": a_si1}]},
            {"primitives": [{"attributes": {"POSITION": a_qpos, "NORMAL": a_qnrm}, "material": 2, "mode": 4}]},
            {"primitives": [{"attributes": {"POSITION": a_upos, "NORMAL": a_qnrm}, "material": 3}]},
        ],
        "materials": [
            {"name": "tiles", "pbrMetallicRoughness": {"baseColorFactor": [0.9, 0.8, 0.7, 1.0], "metallicFactor": 0.1,
                                                      "roughnessFactor": 0.95, "baseColorTexture": {"index": 0}},
             "normalTexture": {"index": 1}, "doubleSided": True},
            {"name": "glass", "pbrMetallicRoughness": {"baseColorFactor": [0.6, 0.9, 0.7, 1.0], "metallicFactor": 0.0,
                                                      "roughnessFactor": 0.02},
             "extensions": {"KHR_materials_transmission": {"transmissionFactor": 0.8}, "KHR_materials_ior": {"ior": 1.33}}},
            {"name": "lamp", "emissiveFactor": [1.0, 0.9, 0.8], "doubleSided": True,
             "extensions": {"KHR_materials_emissive_strength": {"emissiveStrength": 12.5}}},
            {"name": "ignored emission", "emissiveFactor": [1.0, 1.0, 1.0], "emissiveTexture": {"index": 0}},
        ],
        "textures": [{"source": 0}, {"source": 1}],
        "images": images,
        "accessors": B.accessors,
        "bufferViews": B.views,
    }
    if glowing_glass:
        doc["materials"][1]["emissiveFactor"] = [0.2, 0.1, 0.05]
    while len(B.data) % 4:
        B.data.append(0)
    if external:
        doc["buffers"] = [{"byteLength": len(B.data), "uri": "scene.bin"}]
        (tmp_path / "scene.bin").write_bytes(bytes(B.data))
        path = tmp_path / "scene.gltf"
        path.write_text(json.dumps(doc, indent=1))
    else:
        doc["buffers"] = [{"byteLength": len(B.data)}]
        js = json.dumps(doc, separators=(",", ":")).encode()
        js += b" " * (-len(js) % 4)
        total = 12 + 8 + len(js) + 8 + len(B.data)
        path = tmp_path / "scene.glb"
        path.write_bytes(struct.pack("<III", 0x46546C67, 2, total) + struct.pack("<II", len(js), 0x4E4F534A) + js +
                         struct.pack("<II", len(B.data), 0x004E4942) + bytes(B.data))
    return str(path), {"tex0": tex0, "nmap": nmap}
