"""Host-side scene description: the Python mirror of the reference's scene layer, minus Vulkan.

Same names, argument meaning and error behaviour as
  reina::scene::Material      src/scene/Scene.h:21-46           (20 fields, same order)
  reina::scene::ModelData     src/scene/Models.h:23-30
  Models::addModel            src/scene/Models.cpp:24-92        (flat table concatenation, ModelRange)
  Scene::defineObject / defineTexture / addInstance / addObject / build   src/scene/Scene.cpp:6-125
  Instance::computeCDF        src/scene/Instance.cpp:15-53      (per-instance emissive area CDF)
  Instances::computeSamplingDataEmissives   src/scene/Instances.cpp:52-114
The output of build() is the set of tables the reference binds to its ray-tracing descriptor set
(src/Reina.cpp:394-407), packed into the RB200SceneDesc of include/reina_b200.h.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import abi


@dataclass
class Material:
    """reina::scene::Material (src/scene/Scene.h:21-46), aggregate order preserved."""
    materialIdx: int = 0          # 0 lambertian, 1 metal, 2 dielectric, 3 disney
    textureID: int = -1
    normalMapID: int = -1
    bumpMapID: int = -1
    albedo: tuple = (1.0, 1.0, 1.0)
    emission: tuple = (0.0, 0.0, 0.0)
    roughness: float = 0.0
    ior: float = 0.0
    interpNormals: bool = False
    absorption: float = 0.0
    cullBackface: bool = False
    anisotropic: float = 0.0
    subsurface: float = 0.0
    clearcoatGloss: float = 0.0
    sheenTint: tuple = (0.0, 0.0, 0.0)
    specularTint: tuple = (1.0, 1.0, 1.0)
    metallic: float = 0.0
    clearcoat: float = 0.0
    specularTransmission: float = 0.0
    sheen: float = 0.0


@dataclass
class ModelData:
    """reina::scene::ModelData (src/scene/Models.h:23-30). vertices: (n,4) float32 with w = 1;
    tbns: (m,3,3) float32 where tbns[i][c] is column c in (T, B, N); texCoords: (k,2) float32."""
    vertices: np.ndarray
    indices: np.ndarray
    tbns: np.ndarray
    tbnsIndices: np.ndarray
    texCoords: np.ndarray
    texIndices: np.ndarray


@dataclass
class ModelRange:   # src/scene/Models.h:12-21
    firstVertex: int
    firstNormal: int
    indexOffset: int
    tbnsIndexOffset: int
    texIndexOffset: int
    indexCount: int
    tbnsIndexCount: int
    texIndexCount: int


@dataclass
class SceneTables:
    """Everything rb200_scene_create needs, as contiguous numpy arrays (kept alive by this object)."""
    vertices: np.ndarray
    indices: np.ndarray
    instanceProperties: np.ndarray      # structured bytes, 120 B each
    tbns: np.ndarray
    tbnIndices: np.ndarray
    emissive: np.ndarray                # bytes, 112 B each
    cdfTriangles: np.ndarray
    cdfInstances: np.ndarray
    texCoords: np.ndarray
    texIndices: np.ndarray
    textures: list                      # list of (H,W,4) uint8
    instances: np.ndarray               # bytes, 80 B each
    totalEmissiveWeight: float
    numInstanceProperties: int
    numEmissive: int
    numInstances: int
    _keep: list = field(default_factory=list)

    def desc(self):
        d = abi.SceneDesc()

        def fp(a):
            return a.ctypes.data_as(C.POINTER(C.c_float))

        def up(a):
            return a.ctypes.data_as(C.POINTER(C.c_uint32))

        d.vertices, d.numVertices = fp(self.vertices), self.vertices.shape[0]
        d.indices, d.numIndices = up(self.indices), self.indices.size
        d.instanceProperties = self.instanceProperties.ctypes.data_as(C.POINTER(abi.InstanceProperties))
        d.numInstanceProperties = self.numInstanceProperties
        d.tbns, d.numTbns = fp(self.tbns), self.tbns.size // 9
        d.tbnIndices, d.numTbnIndices = up(self.tbnIndices), self.tbnIndices.size
        d.emissiveMetadata = self.emissive.ctypes.data_as(C.POINTER(abi.InstanceData))
        d.numEmissive = self.numEmissive
        d.cdfTriangles, d.numCdfTriangles = fp(self.cdfTriangles), self.cdfTriangles.size
        d.cdfInstances, d.numCdfInstances = fp(self.cdfInstances), self.cdfInstances.size
        d.texCoords, d.numTexCoords = fp(self.texCoords), self.texCoords.size // 2
        d.texIndices, d.numTexIndices = up(self.texIndices), self.texIndices.size
        texs = (abi.Texture * max(1, len(self.textures)))()
        for i, t in enumerate(self.textures):
            texs[i].rgba8 = t.ctypes.data_as(C.POINTER(C.c_uint8))
            texs[i].width, texs[i].height = t.shape[1], t.shape[0]
        self._keep.append(texs)
        d.textures, d.numTextures = texs, len(self.textures)
        d.instances = self.instances.ctypes.data_as(C.POINTER(abi.Instance))
        d.numInstances = self.numInstances
        return d

    def num_triangles(self):
        inst = np.frombuffer(self.instances.tobytes(), dtype=np.uint32).reshape(self.numInstances, 20)
        return int(inst[:, 19].sum())


def _emissive_duplicates(emissive_ids, inst_cdf):
    """Instances::computeEmissiveDuplicates (src/scene/Instances.cpp:27-50): instance index of a later emitter -> instance
    index of the first earlier emitter whose triangle CDF has the same length and agrees within FLT_EPSILON (:12-25)."""
    eps = np.float32(np.finfo(np.float32).eps)
    dup = {}
    for a, i in enumerate(emissive_ids):
        ci = inst_cdf[i][0]
        for j in emissive_ids[a + 1:]:
            cj = inst_cdf[j][0]
            if ci.size != cj.size or bool(np.any(np.abs(ci - cj) > eps)):
                continue
            dup.setdefault(j, i)
    return dup


def _is_emissive(e):
    # Instance::isEmissive, src/scene/Instance.cpp:75-77
    e = np.asarray(e, dtype=np.float32)
    return float(np.dot(e, e)) > np.float32(0.00001) * np.float32(0.00001)


class Scene:
    """reina::scene::Scene without the Vulkan objects (src/scene/Scene.h:62-138)."""

    def __init__(self):
        self.modelData = []
        self.modelRanges = []
        self._allVertices = []
        self._allTBNs = []
        self._allTexCoords = []
        self._allIndicesOffset = []
        self._allTexIndicesOffset = []
        self._allTBNsIndicesOffset = []
        self._nVerts = 0
        self._nTbns = 0
        self._nTex = 0
        self._nIdx = 0
        self._nTbnIdx = 0
        self._nTexIdx = 0
        self.texturesToCreate = []
        self.instancesToCreate = []     # (instancePropertiesID, materialIdx, objectID, transform)
        self.instanceProperties = []
        self.materials = []
        self._built = False

    # -- Models::addModel, src/scene/Models.cpp:24-92 -------------------------------------------------
    def defineObject(self, modelData):
        if self._built:
            raise RuntimeError("Could not add model; buffers are already built")
        md = modelData
        verts = np.ascontiguousarray(md.vertices, dtype=np.float32).reshape(-1, 4)
        tbns = np.ascontiguousarray(md.tbns, dtype=np.float32).reshape(-1, 9)
        tex = np.ascontiguousarray(md.texCoords, dtype=np.float32).reshape(-1, 2)
        idx = np.ascontiguousarray(md.indices, dtype=np.uint32).ravel()
        tbnIdx = np.ascontiguousarray(md.tbnsIndices, dtype=np.uint32).ravel()
        texIdx = np.ascontiguousarray(md.texIndices, dtype=np.uint32).ravel()
        rng = ModelRange(
            firstVertex=self._nVerts, firstNormal=self._nTbns, indexOffset=self._nIdx,
            tbnsIndexOffset=self._nTbnIdx,
            texIndexOffset=0xFFFFFFFF if tex.shape[0] == 0 else self._nTexIdx,
            indexCount=idx.size // 3, tbnsIndexCount=idx.size // 3, texIndexCount=texIdx.size // 3)
        self.modelRanges.append(rng)
        self.modelData.append(md)
        self._allVertices.append(verts)
        self._allTBNs.append(tbns)
        self._allTexCoords.append(tex)
        self._allIndicesOffset.append(idx + np.uint32(self._nVerts))
        self._allTBNsIndicesOffset.append(tbnIdx + np.uint32(self._nTbns))
        t = texIdx.copy()
        m = t != np.uint32(0xFFFFFFFF)
        t[m] = t[m] + np.uint32(self._nTex)
        self._allTexIndicesOffset.append(t)
        self._nVerts += verts.shape[0]
        self._nTbns += tbns.shape[0]
        self._nTex += tex.shape[0]
        self._nIdx += idx.size
        self._nTbnIdx += tbnIdx.size
        self._nTexIdx += texIdx.size
        return len(self.modelRanges) - 1

    def defineTexture(self, rgba8):
        """Takes decoded RGBA8 (H,W,4) uint8 (the reference decodes with stb_image, src/graphics/Image.cpp:10-41;
        decoding stays on the caller's side of the ABI)."""
        a = np.ascontiguousarray(rgba8, dtype=np.uint8)
        if a.ndim != 3 or a.shape[2] != 4:
            raise ValueError("texture must be (H, W, 4) uint8")
        self.texturesToCreate.append(a)
        return len(self.texturesToCreate) - 1

    # -- Scene::addInstance, src/scene/Scene.cpp:27-54 ------------------------------------------------
    def addInstance(self, objectID, transform, mat):
        if objectID >= len(self.modelRanges):
            raise RuntimeError(f"Index {objectID} out of range for models")
        r = self.modelRanges[objectID]
        p = abi.InstanceProperties()
        p.indicesOffset = r.indexOffset
        p.albedo[:] = [float(x) for x in mat.albedo]
        p.emission[:] = [float(x) for x in mat.emission]
        p.tbnsIndicesOffset = r.tbnsIndexOffset
        p.texIndicesOffset = r.texIndexOffset
        p.roughness, p.ior = mat.roughness, mat.ior
        p.interpNormals = 1 if mat.interpNormals else 0
        p.absorption = mat.absorption
        p.textureID, p.normalMapTexID, p.bumpMapTexID = mat.textureID, mat.normalMapID, mat.bumpMapID
        p.cullBackface = 1 if mat.cullBackface else 0
        p.anisotropic, p.subsurface, p.clearcoatGloss = mat.anisotropic, mat.subsurface, mat.clearcoatGloss
        p.sheenTint[:] = [float(x) for x in mat.sheenTint]
        p.specularTint[:] = [float(x) for x in mat.specularTint]
        p.metallic, p.clearcoat = mat.metallic, mat.clearcoat
        p.specularTransmission, p.sheen = mat.specularTransmission, mat.sheen
        self.instanceProperties.append(p)
        self.materials.append(mat)
        t = np.ascontiguousarray(transform, dtype=np.float32).reshape(4, 4)
        self.instancesToCreate.append((len(self.instanceProperties) - 1, int(mat.materialIdx), objectID, t))

    def addObject(self, modelData, transform, mat):
        oid = self.defineObject(modelData)
        self.addInstance(oid, transform, mat)
        return oid

    # -- Instance::computeCDF, src/scene/Instance.cpp:15-53 -------------------------------------------
    @staticmethod
    def _compute_cdf(md, transform, brightness):
        v = np.ascontiguousarray(md.vertices, dtype=np.float32).reshape(-1, 4)
        # glm `vec4 * mat4` is the row-vector product (= transpose(M) * v): reproduced as written (:23)
        M = transform.astype(np.float32)              # M[c] is column c (column-major)
        v4 = np.concatenate([v[:, :3], np.ones((v.shape[0], 1), np.float32)], axis=1)
        tv = np.stack([(v4 * M[c][None, :]).sum(axis=1, dtype=np.float32) for c in range(3)], axis=1).astype(np.float32)
        idx = np.ascontiguousarray(md.indices, dtype=np.uint32).reshape(-1, 3)
        ab = tv[idx[:, 1]] - tv[idx[:, 0]]
        ac = tv[idx[:, 2]] - tv[idx[:, 0]]
        tri_area = (np.linalg.norm(np.cross(ab, ac).astype(np.float32), axis=1).astype(np.float32) / np.float32(2))
        area = np.float32(0)
        cum = np.float32(0)
        cdf = np.zeros(idx.shape[0], np.float32)
        b = np.float32(brightness)
        for i in range(idx.shape[0]):     # sequential fp32 accumulation, as the reference
            area = np.float32(area + tri_area[i])
            cum = np.float32(cum + np.float32(tri_area[i] * b))
            cdf[i] = cum
        if cum == 0.0:
            raise RuntimeError("Cannot calculate CDF for a mesh because the cumulative area is 0")
        cdf = (cdf / cum).astype(np.float32)
        return cdf, float(area), float(cum)

    # -- Scene::build, src/scene/Scene.cpp:56-125 (steps 1,2,4,6; 3 and 5 happen inside rb200_scene_create) ----
    def build(self, require_emitter=False):
        self._built = True

        def cat(lst, dtype, shape):
            if not lst:
                return np.zeros(shape, dtype)
            return np.ascontiguousarray(np.concatenate(lst, axis=0), dtype=dtype)

        vertices = cat(self._allVertices, np.float32, (0, 4))
        tbns = cat(self._allTBNs, np.float32, (0, 9))
        tex = cat(self._allTexCoords, np.float32, (0, 2))
        if tex.shape[0] == 0:
            tex = np.zeros((1, 2), np.float32)      # Models::buildBuffers substitutes {0} (src/scene/Models.cpp:112)
        indices = cat(self._allIndicesOffset, np.uint32, (0,))
        tbnIdx = cat(self._allTBNsIndicesOffset, np.uint32, (0,))
        texIdx = cat(self._allTexIndicesOffset, np.uint32, (0,))
        if texIdx.size == 0:
            texIdx = np.zeros(1, np.uint32)

        nprops = len(self.instanceProperties)
        props = np.frombuffer(b"".join(bytes(p) for p in self.instanceProperties), dtype=np.uint8).copy() \
            if nprops else np.zeros(120, np.uint8)

        # instances (TLAS records) + emissive sampling data
        inst_bytes = []
        emissive_ids = []
        inst_cdf = {}
        for k, (pid, matIdx, oid, T) in enumerate(self.instancesToCreate):
            r = self.modelRanges[oid]
            rec = abi.Instance()
            rec.transform[:] = [float(x) for x in T.reshape(-1)]
            rec.instancePropertiesID, rec.materialIdx = pid, matIdx
            rec.indexOffset, rec.triangleCount = r.indexOffset, r.indexCount
            inst_bytes.append(bytes(rec))
            e = self.materials[pid].emission
            if _is_emissive(e):
                bright = np.float32(0.2126) * np.float32(e[0]) + np.float32(0.7152) * np.float32(e[1]) + \
                    np.float32(0.0722) * np.float32(e[2])
                inst_cdf[k] = self._compute_cdf(self.modelData[oid], T, bright)
                emissive_ids.append(k)
        instances = np.frombuffer(b"".join(inst_bytes), dtype=np.uint8).copy() if inst_bytes else np.zeros(80, np.uint8)

        # Instances::computeSamplingDataEmissives (src/scene/Instances.cpp:52-114), duplicate-CDF sharing included and
        # reproduced AS WRITTEN: computeEmissiveDuplicates (:27-50) keys its map by INSTANCE index, the loop asks it
        # `contains(position in the emissive list)` (:72) and then reads `.at(instance index)` (:86) and indexes the
        # emissive records with the mapped INSTANCE index. The two index spaces coincide only while every instance in
        # front of the duplicate is emissive; otherwise upstream shares nothing, shares another emitter's range, throws
        # std::out_of_range or reads past the vector (refused here) — each outcome is kept.
        dup = _emissive_duplicates(emissive_ids, inst_cdf)
        cdfTriangles = []
        recs = [abi.InstanceData() for _ in emissive_ids]        # value-initialised like std::vector<InstanceData>(n)
        offset = 0
        for pos, k in enumerate(emissive_ids):
            pid, matIdx, oid, T = self.instancesToCreate[k]
            cdf, area, weight = inst_cdf[k]
            d = recs[pos]
            d.transform[:] = [float(x) for x in T.reshape(-1)]
            d.materialOffset = matIdx
            d.indexOffset = self.modelRanges[oid].indexOffset
            d.emission[:] = [float(x) for x in self.materials[pid].emission]
            d.weight, d.area = weight, area
            d.cullBackface = 1 if self.materials[pid].cullBackface else 0
            if pos in dup:
                if k not in dup:
                    raise IndexError("unordered_map::at")        # what std::unordered_map::at throws upstream (:86)
                src = dup[k]
                if src >= len(recs):
                    raise RuntimeError("emissive duplicate mapping points outside the emissive list "
                                       "(undefined behaviour in the reference, src/scene/Instances.cpp:86)")
                d.cdfRangeStart, d.cdfRangeEnd = recs[src].cdfRangeStart, recs[src].cdfRangeEnd
                continue
            d.cdfRangeStart = offset
            d.cdfRangeEnd = offset + cdf.size - 1
            cdfTriangles.append(cdf)
            offset += cdf.size
        em_recs = [bytes(d) for d in recs]
        cum = np.float32(0)
        cdfInstances = np.zeros(len(emissive_ids), np.float32)
        for i, k in enumerate(emissive_ids):
            cum = np.float32(cum + np.float32(inst_cdf[k][2]))
            cdfInstances[i] = cum
        if len(emissive_ids):
            cdfInstances = (cdfInstances / cum).astype(np.float32)
        elif require_emitter:
            raise RuntimeError("Scene must have at least one emissive object")   # src/scene/Instances.cpp:125-127
        emissive = np.frombuffer(b"".join(em_recs), dtype=np.uint8).copy() if em_recs else np.zeros(112, np.uint8)
        cdfT = np.ascontiguousarray(np.concatenate(cdfTriangles), np.float32) if cdfTriangles else np.zeros(1, np.float32)
        if not len(emissive_ids):
            cdfT = np.zeros(0, np.float32)

        return SceneTables(
            vertices=vertices, indices=indices, instanceProperties=props, tbns=tbns.reshape(-1), tbnIndices=tbnIdx,
            emissive=emissive, cdfTriangles=cdfT, cdfInstances=cdfInstances, texCoords=tex.reshape(-1),
            texIndices=texIdx, textures=list(self.texturesToCreate), instances=instances,
            totalEmissiveWeight=float(cum), numInstanceProperties=nprops, numEmissive=len(emissive_ids),
            numInstances=len(self.instancesToCreate))
