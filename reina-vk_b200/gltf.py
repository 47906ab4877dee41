"""glTF 2.0 / GLB scene import — the host-side mirror of reina::scene::gltf (src/scene/gltf/gltfloader.cpp:72-456),
the reference's default scene path (src/Reina.cpp:91). SURVEY.md 8f row 3.

Same stages and names as the reference:
  loadGltf               :72-120    parse .gltf (JSON + external / data-URI buffers) or .glb (JSON + BIN chunks)
  loadPrimitives         :122-255   meshes used by the default scene -> Primitive (POSITION, NORMAL, TANGENT,
                                    TEXCOORD_0, indices; integer attributes are de-quantised, KHR_mesh_quantization)
  Primitive.toModelData  :257-285   one ModelData per primitive, indices == tbnsIndices == texIndices
  addTexturesToScene     :287-343   every image becomes a scene texture: files are flipped vertically, embedded
                                    (bufferView / data-URI) images are not (src/graphics/Image.cpp:14,28)
  materialsFromMeshTBNs  :387-438   glTF PBR -> Disney material (materialIdx 3): roughness clamped to [0.1, 0.7],
                                    cullBackface = !doubleSided, KHR_materials_transmission => specularTransmission and no
                                    culling, KHR_materials_ior, emission = emissiveFactor * emissive_strength only when
                                    the material has NO emissive texture, baseColor / normal textures
  addInstancesToScene    :357-381   depth-first walk of the default scene, world = parent * local, one instance per
                                    primitive of the node's mesh
  loadScene              :440-456   the stages in the reference's order: meshes, textures, materials, instances

The reference parses with fastgltf@42d26b2 and computes missing tangents with MikkTSpace@3e895b4; neither is
vendored in the reference tree (parity unpinned). Its MikkTSpace call (:207-222) reads the vertex array as an
un-indexed triangle soup BEFORE TEXCOORD_0 and the indices have been loaded (:225-249 run afterwards), so every UV it
sees is (0, 0) — and for that input MikkTSpace's published algorithm has exactly one outcome, which
`mikktspace_as_called` below restates: tangent (1, 0, 0) and bitangent (0, 1, 0) for every vertex of a complete
triple, zeros for the one or two left over. That is what `tangents="reference"` (the default) yields; `tangents="uv"`
derives per-vertex frames from the UV derivatives instead (meshes._tangent_frames: what the call was presumably
meant to do). Rules of this importer where the reference leaves a choice:
  * used meshes are defined in ascending mesh index (the reference iterates an unordered_set);
  * TRS node transforms are composed in float64 (T * R * S, quaternion -> matrix by the standard formula), products
    are accumulated in float64 in a fixed order and rounded to float32 once per instance;
  * images: PNG and baseline / progressive JPEG (CMYK JPEG is refused loudly);
  * sparse accessors are resolved (base elements or zeros, then the substitutions);
  * primitive modes other than TRIANGLES are refused loudly.
The C++ host (host/gltf.cpp) implements the same rules; tests/test_gltf.py checks both yield identical tables.
"""
import base64
import io
import json
import os
import struct
import warnings

import numpy as np

from . import meshes
from .scene import Material, Scene

_COMPONENT = {5120: (np.int8, 1), 5121: (np.uint8, 1), 5122: (np.int16, 2), 5123: (np.uint16, 2), 5125: (np.uint32, 4),
              5126: (np.float32, 4)}
_ENABLED_EXTENSIONS = ("KHR_mesh_quantization", "KHR_texture_transform", "KHR_materials_variants", "KHR_materials_transmission",
                       "KHR_materials_clearcoat", "KHR_materials_emissive_strength")
_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT2": 4, "MAT3": 9, "MAT4": 16}


class Asset:
    """Parsed document + resolved buffers (what fastgltf::Asset holds after LoadExternalBuffers)."""

    def __init__(self, doc, buffers, base_dir):
        self.doc, self.buffers, self.base_dir = doc, buffers, base_dir


def _decode_data_uri(uri):
    head, _, payload = uri.partition(",")
    if not head.endswith(";base64"):
        raise RuntimeError("Failed to parse glTF: only base64 data URIs are supported")
    return base64.b64decode(payload)


def loadGltf(filepath):
    """gltfloader.cpp:72-120."""
    if not os.path.exists(filepath):
        raise RuntimeError("Failed to find glTF file: " + filepath)
    with open(filepath, "rb") as f:
        raw = f.read()
    base_dir = os.path.dirname(os.path.abspath(filepath))
    bin_chunk = None
    if raw[:4] == b"glTF":
        if len(raw) < 20:
            raise RuntimeError("Failed to parse glTF: truncated GLB header")
        _, version, length = struct.unpack_from("<III", raw, 0)
        if version != 2:
            raise RuntimeError("Failed to parse glTF: unsupported GLB version %d" % version)
        off, doc = 12, None
        while off + 8 <= min(length, len(raw)):
            clen, ctype = struct.unpack_from("<II", raw, off)
            body = raw[off + 8: off + 8 + clen]
            if len(body) != clen:
                raise RuntimeError("Failed to parse glTF: truncated GLB chunk")
            if ctype == 0x4E4F534A and doc is None:
                doc = json.loads(body.decode("utf-8"))
            elif ctype == 0x004E4942 and bin_chunk is None:
                bin_chunk = bytes(body)
            off += 8 + ((clen + 3) & ~3)
        if doc is None:
            raise RuntimeError("Failed to parse glTF: GLB without a JSON chunk")
    else:
        try:
            doc = json.loads(raw.decode("utf-8"))
        except ValueError as e:
            raise RuntimeError("Failed to parse glTF: " + str(e))
    # the parser of the reference is created with exactly these extensions (gltfloader.cpp:87-93) and rejects a file that
    # REQUIRES any other one (compressed geometry, basis textures, ...)
    for e in doc.get("extensionsRequired", []):
        if e not in _ENABLED_EXTENSIONS:
            raise RuntimeError("Failed to parse glTF: required extension " + str(e) + " is not enabled")
    buffers = []
    for i, b in enumerate(doc.get("buffers", [])):
        uri = b.get("uri")
        if uri is None:
            if i != 0 or bin_chunk is None:
                raise RuntimeError("Failed to parse glTF: buffer %d has no uri and no GLB BIN chunk" % i)
            data = bin_chunk
        elif uri.startswith("data:"):
            data = _decode_data_uri(uri)
        else:
            p = os.path.join(base_dir, uri)
            if not os.path.exists(p):
                raise RuntimeError("Failed to parse glTF: missing external buffer " + uri)
            with open(p, "rb") as f:
                data = f.read()
        if len(data) < int(b.get("byteLength", 0)):
            raise RuntimeError("Failed to parse glTF: buffer %d is shorter than its byteLength" % i)
        buffers.append(data)
    return Asset(doc, buffers, base_dir)


def _checked(v, what, index):
    """Untrusted JSON number -> int in [0, 2^31): negative or huge counts / offsets / strides are refused (host/gltf.cpp
    applies the same rule)."""
    if isinstance(v, bool) or not isinstance(v, int) or v < 0 or v > 0x7FFFFFFF:
        raise RuntimeError("Failed to parse glTF: %s of accessor %d out of range" % (what, index))
    return v


def _read_accessor(asset, index, want_float=True):
    """Accessor -> (count, ncomp) array. Floats as stored; integers either raw (indices) or de-quantised to fp32:
    normalized -> c / max (signed: max(c / max, -1)), as the glTF specification and KHR_mesh_quantization define."""
    doc = asset.doc
    acc = doc["accessors"][_checked(index, "index", index)]
    dt, size = _COMPONENT[acc["componentType"]]
    ncomp = _NCOMP[acc["type"]]
    count = _checked(acc["count"], "count", index)
    if "bufferView" not in acc:
        out = np.zeros((count, ncomp), dt)
    else:
        bv = doc["bufferViews"][_checked(acc["bufferView"], "bufferView", index)]
        buf = asset.buffers[_checked(bv["buffer"], "buffer", index)]
        voff = _checked(bv.get("byteOffset", 0), "bufferView.byteOffset", index)
        start = voff + _checked(acc.get("byteOffset", 0), "byteOffset", index)
        stride = _checked(bv.get("byteStride", 0), "byteStride", index) or size * ncomp
        need = start + (count - 1) * stride + size * ncomp if count else start
        if need > len(buf) or need > voff + _checked(bv["byteLength"], "bufferView.byteLength", index):
            raise RuntimeError("Failed to parse glTF: accessor %d reads past its buffer view" % index)
        raw = np.frombuffer(buf, np.uint8)
        idx = start + np.arange(count, dtype=np.int64)[:, None] * stride + np.arange(size * ncomp, dtype=np.int64)[None, :]
        out = np.ascontiguousarray(raw[idx]).view(dt).reshape(count, ncomp)
    if "sparse" in acc:
        # glTF 2.0, 3.6.2.3: `count` elements replaced; strictly increasing element numbers, tightly packed values
        sp = acc["sparse"]
        n = int(sp["count"])
        idt, isize = _COMPONENT[sp["indices"]["componentType"]]
        if idt not in (np.uint8, np.uint16, np.uint32):
            raise RuntimeError("Failed to parse glTF: bad sparse index type")

        def view(ref, nbytes):
            bv = doc["bufferViews"][ref["bufferView"]]
            buf = asset.buffers[bv["buffer"]]
            voff = int(bv.get("byteOffset", 0))
            start = voff + int(ref.get("byteOffset", 0))
            if start + nbytes > len(buf) or start + nbytes > voff + int(bv["byteLength"]):
                raise RuntimeError("Failed to parse glTF: accessor %d reads past its buffer view" % index)
            return bytes(buf[start:start + nbytes])
        if n > count:
            raise RuntimeError("Failed to parse glTF: bad sparse accessor")
        out = out.copy()
        if n:
            at = np.frombuffer(view(sp["indices"], n * isize), idt).astype(np.int64)
            vals = np.frombuffer(view(sp["values"], n * size * ncomp), dt).reshape(n, ncomp)
            if at.max() >= count or (np.diff(at) <= 0).any():
                raise RuntimeError("Failed to parse glTF: bad sparse accessor")
            out[at] = vals
    if not want_float:
        return out
    if dt == np.float32:
        return out.astype(np.float32)
    f = out.astype(np.float32)
    if acc.get("normalized", False):
        mx = np.float32({np.int8: 127, np.uint8: 255, np.int16: 32767, np.uint16: 65535}.get(dt, 1))
        f = f / mx
        if dt in (np.int8, np.int16):
            f = np.maximum(f, np.float32(-1.0))
    return f.astype(np.float32)


class Primitive:
    """reina::scene::gltf::Primitive (gltfloader.h:19-25): per-vertex position, normal, tangent, bitangent, uv."""

    def __init__(self):
        self.position = self.normal = self.tangent = self.bitangent = self.uv = self.indices = None
        self.materialIdx = -1
        self.hasTexCoords = False

    def toModelData(self):
        """gltfloader.cpp:257-285."""
        n = self.position.shape[0]
        tbn = np.zeros((n, 3, 3), np.float32)
        tbn[:, 0, :], tbn[:, 1, :], tbn[:, 2, :] = self.tangent, self.bitangent, self.normal
        tris = self.indices.reshape(-1, 3)
        return meshes.make_model(self.position, self.uv, self.normal, tris, tbn=tbn)


def mikktspace_as_called(n):
    """What genTangSpaceDefault (MikkTSpace@3e895b4, mikktspace.c) writes through the reference's callbacks
    (gltfloader.cpp:16-67, 207-222) for a primitive of n vertices. Returns (tangent, bitangent), float32 [n, 3].

    The call sees faces f = vertices 3f .. 3f+2 (getNumFaces = n / 3, no indices) whose texture coordinates are all
    (0, 0): `m.vertices.resize(count)` value-initialises VertexTBN and TEXCOORD_0 is only read afterwards. Following
    genTangSpace step by step with that input:
      * InitTriInfo marks every triangle GROUP_WITH_ANY ("assumed bad") and clears the mark only where the signed UV
        area t21x * t31y - t21y * t31x is non-zero — here it is 0 for every triangle, so vOs = vOt = 0 and the mark stays;
      * Build4RuleGroups opens a group only at a triangle WITHOUT that mark: no group is ever opened, so
        GenerateTSpaces has nothing to evaluate;
      * every per-corner tangent space therefore keeps the value it was initialised with before GenerateTSpaces:
        vOs = (1, 0, 0), vOt = (0, 1, 0), fMagS = fMagT = 1; DegenEpilogue copies such a space between corners of
        position-degenerate triangles (same values) and the final loop hands vOs / vOt to setTSpace for the three
        corners of every face.
    Vertices 3 * (n / 3) .. n - 1 belong to no face and keep the zeros of the value-initialisation; with n < 3 the
    call returns before writing anything. Positions, normals, welding and the angular threshold cannot change any of
    this: they only enter through groups, and there are none."""
    t = np.zeros((n, 3), np.float32)
    b = np.zeros((n, 3), np.float32)
    full = 3 * (n // 3)
    t[:full, 0] = 1.0
    b[:full, 1] = 1.0
    return t, b


def _cross32(a, b):
    return np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                     a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], axis=1).astype(np.float32)


def _default_scene_nodes(doc):
    scenes = doc.get("scenes", [])
    if not scenes:
        raise RuntimeError("No scenes supplied in gLTF file")
    return scenes[int(doc.get("scene", 0))].get("nodes", [])


def _walk(doc, visit):
    """fastgltf::iterateSceneNodes: depth-first, parents before children, world = parent * local (float64 here)."""
    nodes = doc.get("nodes", [])

    def rec(i, parent, depth):
        if depth > 256:
            raise RuntimeError("Failed to parse glTF: node hierarchy too deep (cycle?)")
        node = nodes[i]
        world = _mat_mul(parent, _local_matrix(node))
        visit(node, world)
        for c in node.get("children", []):
            rec(c, world, depth + 1)
    ident = [[1.0 if r == c else 0.0 for r in range(4)] for c in range(4)]     # [column][row]
    for i in _default_scene_nodes(doc):
        rec(i, ident, 0)


def _mat_mul(a, b):
    """Column-major 4x4 product in float64, sum over k in ascending order (no BLAS: the order is part of the rule)."""
    out = [[0.0] * 4 for _ in range(4)]
    for c in range(4):
        for r in range(4):
            s = 0.0
            for k in range(4):
                s = s + a[k][r] * b[c][k]
            out[c][r] = s
    return out


def _local_matrix(node):
    if "matrix" in node:
        m = [float(x) for x in node["matrix"]]
        return [[m[4 * c + r] for r in range(4)] for c in range(4)]
    t = [float(x) for x in node.get("translation", [0, 0, 0])]
    q = [float(x) for x in node.get("rotation", [0, 0, 0, 1])]
    s = [float(x) for x in node.get("scale", [1, 1, 1])]
    x, y, z, w = q
    # rotation matrix of a unit quaternion, rows r0..r2
    r0 = [1.0 - 2.0 * (y * y + z * z), 2.0 * (x * y - z * w), 2.0 * (x * z + y * w)]
    r1 = [2.0 * (x * y + z * w), 1.0 - 2.0 * (x * x + z * z), 2.0 * (y * z - x * w)]
    r2 = [2.0 * (x * z - y * w), 2.0 * (y * z + x * w), 1.0 - 2.0 * (x * x + y * y)]
    rows = [r0, r1, r2]
    m = [[0.0] * 4 for _ in range(4)]
    for c in range(3):
        for r in range(3):
            m[c][r] = rows[r][c] * s[c]
    m[3][0], m[3][1], m[3][2], m[3][3] = t[0], t[1], t[2], 1.0
    return m


TANGENT_RULES = ("reference", "uv")


def loadPrimitives(asset, tangents="reference"):
    """gltfloader.cpp:122-255. Returns {mesh index: [Primitive, ...]} for the meshes the default scene uses, in
    ascending mesh index. `tangents`: what a primitive without TANGENT gets — "reference": the outcome of the reference's
    MikkTSpace call as written (mikktspace_as_called), "uv": frames from the UV derivatives."""
    if tangents not in TANGENT_RULES:
        raise ValueError("tangents must be one of " + ", ".join(TANGENT_RULES))
    doc = asset.doc
    used = set()
    _walk(doc, lambda node, world: used.add(node["mesh"]) if "mesh" in node else None)
    out = {}
    for mi in sorted(used):
        prims = []
        for prim in doc["meshes"][mi].get("primitives", []):
            if int(prim.get("mode", 4)) != 4:
                raise RuntimeError("Failed to parse glTF: only TRIANGLES primitives are supported")
            attrs = prim.get("attributes", {})
            m = Primitive()
            m.materialIdx = int(prim.get("material", -1))
            if "POSITION" not in attrs:
                raise RuntimeError("Failed to parse glTF: primitive without POSITION")
            m.position = _read_accessor(asset, attrs["POSITION"])[:, :3]
            n = m.position.shape[0]
            if "NORMAL" not in attrs:
                raise RuntimeError("Meshes without vertex normals are not supported")
            m.normal = _read_accessor(asset, attrs["NORMAL"])[:, :3]
            if "TEXCOORD_0" in attrs:
                m.uv = _read_accessor(asset, attrs["TEXCOORD_0"])[:, :2]
                m.hasTexCoords = True
            else:
                warnings.warn("falling back to UV coords (0, 0) since none were found")
                m.uv = np.zeros((n, 2), np.float32)
            if "indices" in prim:
                m.indices = _read_accessor(asset, prim["indices"], want_float=False).astype(np.uint32).reshape(-1)
            else:
                m.indices = np.arange(n, dtype=np.uint32)           # fastgltf::Options::GenerateMeshIndices
            if m.indices.size % 3 or (m.indices.size and int(m.indices.max()) >= n):
                raise RuntimeError("Failed to parse glTF: bad index data")
            if "TANGENT" in attrs:
                t = _read_accessor(asset, attrs["TANGENT"])
                m.tangent = np.ascontiguousarray(t[:, :3], np.float32)
                w = t[:, 3] if t.shape[1] >= 4 else np.ones(n, np.float32)     # Vec3: assume w = +1
                m.bitangent = (_cross32(m.normal, m.tangent) * w[:, None].astype(np.float32)).astype(np.float32)
            elif tangents == "reference":
                m.tangent, m.bitangent = mikktspace_as_called(n)
            else:
                tb = meshes._tangent_frames(m.position, m.uv, m.normal, m.indices.reshape(-1, 3))
                m.tangent, m.bitangent = tb[:, 0, :].copy(), tb[:, 1, :].copy()
            prims.append(m)
        out[mi] = prims
    return out


def addMeshesToScene(scene, meshIdToPrimitives):
    """gltfloader.cpp:345-355."""
    return {mi: [scene.defineObject(p.toModelData()) for p in prims] for mi, prims in meshIdToPrimitives.items()}


def _decode_image(data, flip, name):
    """PNG or JPEG (baseline or progressive), by signature (the C++ host decodes both itself: host/texture.cpp, host/jpeg.cpp; its JPEG
    path follows the IJG integer pipeline, i.e. it yields PIL's bytes)."""
    from PIL import Image
    if data[:8] != b"\x89PNG\r\n\x1a\n" and data[:3] != b"\xff\xd8\xff":
        raise RuntimeError("Could not load image at path: " + name + ": neither a PNG nor a JPEG file")
    im = Image.open(io.BytesIO(data))
    if im.format == "JPEG" and im.mode == "CMYK":
        raise RuntimeError("Could not load image at path: " + name + ": unsupported JPEG (CMYK)")
    img = np.asarray(im.convert("RGBA"), np.uint8)
    return np.ascontiguousarray(img[::-1] if flip else img)


def addTexturesToScene(asset, scene):
    """gltfloader.cpp:287-343: every image, in order. File images are flipped vertically, embedded ones are not."""
    doc = asset.doc
    ids = {}
    for i, img in enumerate(doc.get("images", [])):
        if "uri" in img and not img["uri"].startswith("data:"):
            p = os.path.join(asset.base_dir, img["uri"])
            if not os.path.exists(p):
                raise RuntimeError("Could not load image at path: " + p)
            with open(p, "rb") as f:
                ids[i] = scene.defineTexture(_decode_image(f.read(), True, p))
        elif "uri" in img:
            ids[i] = scene.defineTexture(_decode_image(_decode_data_uri(img["uri"]), False, "image %d" % i))
        elif "bufferView" in img:
            bv = doc["bufferViews"][img["bufferView"]]
            buf = asset.buffers[bv["buffer"]]
            off = int(bv.get("byteOffset", 0))
            ids[i] = scene.defineTexture(_decode_image(buf[off: off + int(bv["byteLength"])], False, "image %d" % i))
        else:
            raise RuntimeError("Could not parse texture; internal gLTF data type not supported")
    return ids


def _f32(x):
    return float(np.float32(x))


def materialsFromMeshTBNs(asset, meshIdToPrimitives, gltfTexIdToSceneId):
    """gltfloader.cpp:387-438."""
    doc = asset.doc
    out = {}

    def tex_id(info, what):
        try:
            return int(gltfTexIdToSceneId[int(doc["textures"][int(info["index"])]["source"])])
        except (KeyError, IndexError):
            warnings.warn(what + " ID not found")
            return -1
    for mi, prims in meshIdToPrimitives.items():
        mats = []
        for p in prims:
            mat = Material(materialIdx=3, textureID=-1, normalMapID=-1, bumpMapID=-1, albedo=(1.0, 1.0, 1.0),
                           emission=(0.0, 0.0, 0.0), roughness=0.0, ior=1.5, interpNormals=True, absorption=0.0,
                           cullBackface=False, anisotropic=0.0, subsurface=0.0, clearcoatGloss=0.0,
                           sheenTint=(1.0, 1.0, 1.0), specularTint=(1.0, 1.0, 1.0), metallic=0.0, clearcoat=0.0,
                           specularTransmission=0.0, sheen=0.0)
            if p.materialIdx != -1:
                g = doc["materials"][p.materialIdx]
                ext = g.get("extensions", {})
                ef = [_f32(x) for x in g.get("emissiveFactor", [0.0, 0.0, 0.0])]
                if any(e > 0 for e in ef) and "emissiveTexture" not in g:
                    strength = np.float32(ext.get("KHR_materials_emissive_strength", {}).get("emissiveStrength", 1.0))
                    mat.emission = tuple(float(np.float32(e) * strength) for e in ef)
                pbr = g.get("pbrMetallicRoughness", {})
                mat.metallic = _f32(pbr.get("metallicFactor", 1.0))
                mat.roughness = float(np.maximum(np.minimum(np.float32(pbr.get("roughnessFactor", 1.0)), np.float32(0.7)),
                                                 np.float32(0.1)))
                mat.albedo = tuple(_f32(x) for x in pbr.get("baseColorFactor", [1.0, 1.0, 1.0, 1.0])[:3])
                mat.cullBackface = not bool(g.get("doubleSided", False))
                mat.ior = _f32(ext.get("KHR_materials_ior", {}).get("ior", 1.5))
                if "KHR_materials_transmission" in ext:
                    mat.specularTransmission = _f32(ext["KHR_materials_transmission"].get("transmissionFactor", 0.0))
                    mat.cullBackface = False
                if "baseColorTexture" in pbr:
                    mat.textureID = tex_id(pbr["baseColorTexture"], "Texture")
                if "normalTexture" in g:
                    mat.normalMapID = tex_id(g["normalTexture"], "Normal Texture")
            mats.append(mat)
        out[mi] = mats
    return out


def addInstancesToScene(asset, scene, gltfIdToSceneId, gltfModelIdToMaterials):
    """gltfloader.cpp:357-381."""
    def visit(node, world):
        if "mesh" not in node:
            return
        m = np.array(world, np.float64).astype(np.float32)          # [column][row] -> column-major 16 floats
        for oid, mat in zip(gltfIdToSceneId[node["mesh"]], gltfModelIdToMaterials[node["mesh"]]):
            scene.addInstance(oid, m, mat)
    _walk(asset.doc, visit)


def loadScene(filepath, tangents="reference"):
    """gltfloader.cpp:440-456 without the Vulkan handles; the caller builds the tables (Scene.build)."""
    asset = loadGltf(filepath)
    prims = loadPrimitives(asset, tangents)
    scene = Scene()
    model_ids = addMeshesToScene(scene, prims)
    tex_ids = addTexturesToScene(asset, scene)
    mats = materialsFromMeshTBNs(asset, prims, tex_ids)
    addInstancesToScene(asset, scene, model_ids, mats)
    return scene


def has_emitter(scene):
    return any(float(np.dot(np.float32(m.emission), np.float32(m.emission))) > 1e-10 for m in scene.materials)
