// See scene.h. fp32 evaluation order matches reina-vk_b200/scene.py (built with -ffp-contract=off).
#include "scene.h"

#include <cmath>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <unordered_map>

namespace rbhost {

RB200SceneDesc SceneTables::desc() {
    RB200SceneDesc d{};
    d.vertices = vertices.data();            d.numVertices = uint32_t(vertices.size() / 4);
    d.indices = indices.data();              d.numIndices = uint32_t(indices.size());
    d.instanceProperties = instanceProperties.data();
    d.numInstanceProperties = uint32_t(instanceProperties.size());
    d.tbns = tbns.data();                    d.numTbns = uint32_t(tbns.size() / 9);
    d.tbnIndices = tbnIndices.data();        d.numTbnIndices = uint32_t(tbnIndices.size());
    d.emissiveMetadata = emissive.data();    d.numEmissive = uint32_t(emissive.size());
    d.cdfTriangles = cdfTriangles.data();    d.numCdfTriangles = uint32_t(cdfTriangles.size());
    d.cdfInstances = cdfInstances.data();    d.numCdfInstances = uint32_t(cdfInstances.size());
    d.texCoords = texCoords.data();          d.numTexCoords = uint32_t(texCoords.size() / 2);
    d.texIndices = texIndices.data();        d.numTexIndices = uint32_t(texIndices.size());
    textureRecords.clear();
    for (const Image8& t : textures) textureRecords.push_back({t.rgba.data(), uint32_t(t.width), uint32_t(t.height)});
    d.textures = textureRecords.data();      d.numTextures = uint32_t(textureRecords.size());
    d.instances = instances.data();          d.numInstances = uint32_t(instances.size());
    return d;
}

uint64_t SceneTables::numTriangles() const {
    uint64_t n = 0;
    for (const RB200Instance& i : instances) n += i.triangleCount;
    return n;
}

// Models::addModel, src/scene/Models.cpp:24-92
uint32_t Scene::defineObject(const ModelData& md) {
    if (built) throw std::runtime_error("Could not add model; buffers are already built");
    const uint32_t nVerts = uint32_t(allVertices.size() / 4), nTbns = uint32_t(allTBNs.size() / 9);
    const uint32_t nTex = uint32_t(allTexCoords.size() / 2);
    ModelRange r{};
    r.firstVertex = nVerts;
    r.firstNormal = nTbns;
    r.indexOffset = uint32_t(allIndices.size());
    r.tbnsIndexOffset = uint32_t(allTBNsIndices.size());
    r.texIndexOffset = md.texCoords.empty() ? 0xFFFFFFFFu : uint32_t(allTexIndices.size());
    r.indexCount = r.tbnsIndexCount = uint32_t(md.indices.size() / 3);
    r.texIndexCount = uint32_t(md.texIndices.size() / 3);
    modelRanges.push_back(r);
    modelData.push_back(md);
    allVertices.insert(allVertices.end(), md.vertices.begin(), md.vertices.end());
    allTBNs.insert(allTBNs.end(), md.tbns.begin(), md.tbns.end());
    allTexCoords.insert(allTexCoords.end(), md.texCoords.begin(), md.texCoords.end());
    for (uint32_t i : md.indices) allIndices.push_back(i + nVerts);
    for (uint32_t i : md.tbnsIndices) allTBNsIndices.push_back(i + nTbns);
    for (uint32_t i : md.texIndices) allTexIndices.push_back(i == 0xFFFFFFFFu ? i : i + nTex);
    return uint32_t(modelRanges.size() - 1);
}

uint32_t Scene::defineTexture(Image8 rgba8) {
    if (rgba8.width <= 0 || rgba8.height <= 0 || rgba8.rgba.size() != size_t(rgba8.width) * rgba8.height * 4)
        throw std::invalid_argument("texture must be width x height RGBA8");
    texturesToCreate.push_back(std::move(rgba8));
    return uint32_t(texturesToCreate.size() - 1);
}

// Scene::addInstance, src/scene/Scene.cpp:27-54
void Scene::addInstance(uint32_t objectID, const Mat4f& transform, const Material& mat) {
    if (objectID >= modelRanges.size())
        throw std::runtime_error("Index " + std::to_string(objectID) + " out of range for models");
    const ModelRange& r = modelRanges[objectID];
    RB200InstanceProperties p{};
    p.indicesOffset = r.indexOffset;
    for (int k = 0; k < 3; k++) {
        p.albedo[k] = mat.albedo[k];
        p.emission[k] = mat.emission[k];
        p.sheenTint[k] = mat.sheenTint[k];
        p.specularTint[k] = mat.specularTint[k];
    }
    p.tbnsIndicesOffset = r.tbnsIndexOffset;
    p.texIndicesOffset = r.texIndexOffset;
    p.roughness = mat.roughness;
    p.ior = mat.ior;
    p.interpNormals = mat.interpNormals ? 1u : 0u;
    p.absorption = mat.absorption;
    p.textureID = mat.textureID;
    p.normalMapTexID = mat.normalMapID;
    p.bumpMapTexID = mat.bumpMapID;
    p.cullBackface = mat.cullBackface ? 1u : 0u;
    p.anisotropic = mat.anisotropic;
    p.subsurface = mat.subsurface;
    p.clearcoatGloss = mat.clearcoatGloss;
    p.metallic = mat.metallic;
    p.clearcoat = mat.clearcoat;
    p.specularTransmission = mat.specularTransmission;
    p.sheen = mat.sheen;
    instanceProperties.push_back(p);
    materials.push_back(mat);
    instancesToCreate.push_back({uint32_t(instanceProperties.size() - 1), mat.materialIdx, objectID, transform});
}

uint32_t Scene::addObject(const ModelData& md, const Mat4f& transform, const Material& mat) {
    const uint32_t id = defineObject(md);
    addInstance(id, transform, mat);
    return id;
}

// Instance::computeCDF, src/scene/Instance.cpp:15-53. `vec4 * mat4` there is glm's row-vector product, i.e.
// component c of the result is dot(v, column c): reproduced as written.
Scene::Cdf Scene::computeCDF(const ModelData& md, const Mat4f& M, float brightness) {
    const size_t nv = md.numVertices(), nt = md.numTriangles();
    std::vector<float> tv(nv * 3);
    for (size_t v = 0; v < nv; v++) {
        const float p[4] = {md.vertices[4 * v], md.vertices[4 * v + 1], md.vertices[4 * v + 2], 1.0f};
        for (int c = 0; c < 3; c++) {
            const float p0 = p[0] * M[c * 4], p1 = p[1] * M[c * 4 + 1], p2 = p[2] * M[c * 4 + 2], p3 = p[3] * M[c * 4 + 3];
            tv[3 * v + c] = ((p0 + p1) + p2) + p3;
        }
    }
    Cdf out;
    out.cdf.resize(nt);
    float area = 0.0f, cum = 0.0f;
    for (size_t i = 0; i < nt; i++) {
        const float* a = &tv[3 * md.indices[3 * i]];
        const float* b = &tv[3 * md.indices[3 * i + 1]];
        const float* c = &tv[3 * md.indices[3 * i + 2]];
        const float ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
        const float ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
        const float m0 = ab[1] * ac[2], s0 = ab[2] * ac[1];
        const float m1 = ab[2] * ac[0], s1 = ab[0] * ac[2];
        const float m2 = ab[0] * ac[1], s2 = ab[1] * ac[0];
        const float cx = m0 - s0, cy = m1 - s1, cz = m2 - s2;
        const float x2 = cx * cx, y2 = cy * cy, z2 = cz * cz;
        const float triArea = std::sqrt((x2 + y2) + z2) / 2.0f;
        area = area + triArea;
        const float w = triArea * brightness;
        cum = cum + w;
        out.cdf[i] = cum;
    }
    if (cum == 0.0f) throw std::runtime_error("Cannot calculate CDF for a mesh because the cumulative area is 0");
    for (float& x : out.cdf) x = x / cum;
    out.area = area;
    out.weight = cum;
    return out;
}

// Scene::build, src/scene/Scene.cpp:56-125: steps 1, 2, 4, 6 (acceleration structures are built by the library)
SceneTables Scene::build(bool requireEmitter) {
    built = true;
    SceneTables t;
    t.vertices = allVertices;
    t.tbns = allTBNs;
    t.texCoords = allTexCoords;
    if (t.texCoords.empty()) t.texCoords.assign(2, 0.0f);   // Models::buildBuffers substitutes {0} (Models.cpp:112)
    t.indices = allIndices;
    t.tbnIndices = allTBNsIndices;
    t.texIndices = allTexIndices;
    if (t.texIndices.empty()) t.texIndices.assign(1, 0u);
    t.instanceProperties = instanceProperties;
    t.textures = texturesToCreate;

    std::vector<size_t> emissiveIds;
    std::vector<Cdf> cdfs;
    for (size_t k = 0; k < instancesToCreate.size(); k++) {
        const PendingInstance& pi = instancesToCreate[k];
        const ModelRange& r = modelRanges[pi.objectID];
        RB200Instance rec{};
        std::memcpy(rec.transform, pi.transform.data(), sizeof rec.transform);
        rec.instancePropertiesID = pi.propertiesID;
        rec.materialIdx = pi.materialIdx;
        rec.indexOffset = r.indexOffset;
        rec.triangleCount = r.indexCount;
        t.instances.push_back(rec);
        const std::array<float, 3>& e = materials[pi.propertiesID].emission;
        const float e2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
        if (e2 > 0.00001f * 0.00001f) {   // Instance::isEmissive, src/scene/Instance.cpp:75-77
            const float bright = (0.2126f * e[0] + 0.7152f * e[1]) + 0.0722f * e[2];
            cdfs.push_back(computeCDF(modelData[pi.objectID], pi.transform, bright));
            emissiveIds.push_back(k);
        }
    }

    // Instances::computeEmissiveDuplicates (src/scene/Instances.cpp:27-50): instance index of a later emitter -> instance
    // index of the first earlier emitter whose CDF has the same length and agrees within FLT_EPSILON (:12-25)
    std::unordered_map<size_t, size_t> duplicates;
    for (size_t a = 0; a < emissiveIds.size(); a++) {
        for (size_t b = a + 1; b < emissiveIds.size(); b++) {
            const std::vector<float>&ca = cdfs[a].cdf, &cb = cdfs[b].cdf;
            if (ca.size() != cb.size()) continue;
            bool same = true;
            for (size_t i = 0; i < ca.size() && same; i++) same = !(std::fabs(ca[i] - cb[i]) > std::numeric_limits<float>::epsilon());
            if (same) duplicates.emplace(emissiveIds[b], emissiveIds[a]);      // emplace keeps the first assignment
        }
    }

    // Instances::computeSamplingDataEmissives (src/scene/Instances.cpp:52-114), duplicate sharing reproduced AS WRITTEN:
    // the map is keyed by instance index, but the loop asks `contains(position in the emissive list)` (:72), reads
    // `.at(instance index)` (:86) and indexes the emissive records with the mapped INSTANCE index. The index spaces
    // coincide only while every instance in front of the duplicate is emissive; otherwise upstream shares nothing,
    // shares another emitter's range, throws std::out_of_range or reads past the vector (refused here).
    uint32_t offset = 0;
    float cum = 0.0f;
    t.emissive.assign(emissiveIds.size(), RB200InstanceData{});
    for (size_t i = 0; i < emissiveIds.size(); i++) {
        const PendingInstance& pi = instancesToCreate[emissiveIds[i]];
        const Material& m = materials[pi.propertiesID];
        RB200InstanceData& d = t.emissive[i];
        std::memcpy(d.transform, pi.transform.data(), sizeof d.transform);
        d.materialOffset = pi.materialIdx;
        d.indexOffset = modelRanges[pi.objectID].indexOffset;
        for (int k = 0; k < 3; k++) d.emission[k] = m.emission[k];
        d.weight = cdfs[i].weight;
        d.area = cdfs[i].area;
        d.cullBackface = m.cullBackface ? 1u : 0u;
        cum = cum + cdfs[i].weight;
        t.cdfInstances.push_back(cum);
        if (duplicates.count(i)) {
            const auto it = duplicates.find(emissiveIds[i]);
            if (it == duplicates.end()) throw std::out_of_range("unordered_map::at");
            if (it->second >= t.emissive.size())
                throw std::runtime_error("emissive duplicate mapping points outside the emissive list "
                                         "(undefined behaviour in the reference, src/scene/Instances.cpp:86)");
            d.cdfRangeStart = t.emissive[it->second].cdfRangeStart;
            d.cdfRangeEnd = t.emissive[it->second].cdfRangeEnd;
            continue;
        }
        d.cdfRangeStart = offset;
        d.cdfRangeEnd = offset + uint32_t(cdfs[i].cdf.size()) - 1;
        t.cdfTriangles.insert(t.cdfTriangles.end(), cdfs[i].cdf.begin(), cdfs[i].cdf.end());
        offset += uint32_t(cdfs[i].cdf.size());
    }
    if (!emissiveIds.empty()) {
        for (float& x : t.cdfInstances) x = x / cum;
    } else if (requireEmitter) {
        throw std::runtime_error("Scene must have at least one emissive object");
    }
    t.totalEmissiveWeight = cum;
    return t;
}

}  // namespace rbhost
