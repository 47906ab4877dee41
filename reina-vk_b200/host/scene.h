// Host-side scene layer of the headless renderer: reina::scene::Scene without the Vulkan objects.
//   Material                       src/scene/Scene.h:21-46      (20 fields, same order and defaults)
//   Scene::defineObject / defineTexture / addInstance / addObject / build     src/scene/Scene.cpp:6-125
//   Models::addModel               src/scene/Models.cpp:24-92   (flat table concatenation, ModelRange)
//   Instance::computeCDF           src/scene/Instance.cpp:15-53
//   Instances::computeSamplingDataEmissives                     src/scene/Instances.cpp:52-114
// build() yields the tables of Reina::writeDescriptorSets (src/Reina.cpp:394-407) as an RB200SceneDesc.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/reina_b200.h"
#include "model.h"
#include "rh_math.h"

namespace rbhost {

struct Material {
    uint32_t materialIdx = 0;   // 0 lambertian, 1 metal, 2 dielectric, 3 disney
    int textureID = -1;
    int normalMapID = -1;
    int bumpMapID = -1;
    std::array<float, 3> albedo{1.0f, 1.0f, 1.0f};
    std::array<float, 3> emission{0.0f, 0.0f, 0.0f};
    float roughness = 0.0f;
    float ior = 0.0f;
    bool interpNormals = false;
    float absorption = 0.0f;
    bool cullBackface = false;
    float anisotropic = 0.0f;
    float subsurface = 0.0f;
    float clearcoatGloss = 0.0f;
    std::array<float, 3> sheenTint{0.0f, 0.0f, 0.0f};
    std::array<float, 3> specularTint{1.0f, 1.0f, 1.0f};
    float metallic = 0.0f;
    float clearcoat = 0.0f;
    float specularTransmission = 0.0f;
    float sheen = 0.0f;
};

struct ModelRange {   // src/scene/Models.h:12-21
    uint32_t firstVertex, firstNormal, indexOffset, tbnsIndexOffset, texIndexOffset;
    uint32_t indexCount, tbnsIndexCount, texIndexCount;
};

// Everything rb200_scene_create reads; desc() points into these vectors, so keep the object alive during the call.
struct SceneTables {
    std::vector<float> vertices;
    std::vector<uint32_t> indices;
    std::vector<RB200InstanceProperties> instanceProperties;
    std::vector<float> tbns;
    std::vector<uint32_t> tbnIndices;
    std::vector<RB200InstanceData> emissive;
    std::vector<float> cdfTriangles;
    std::vector<float> cdfInstances;
    std::vector<float> texCoords;
    std::vector<uint32_t> texIndices;
    std::vector<Image8> textures;
    std::vector<RB200Texture> textureRecords;
    std::vector<RB200Instance> instances;
    float totalEmissiveWeight = 0.0f;
    RB200SceneDesc desc();
    uint64_t numTriangles() const;
};

class Scene {
public:
    uint32_t defineObject(const ModelData& modelData);
    uint32_t defineObject(const std::string& objPath) { return defineObject(load_obj(objPath)); }
    uint32_t defineTexture(Image8 rgba8);
    void addInstance(uint32_t objectID, const Mat4f& transform, const Material& mat);
    uint32_t addObject(const ModelData& modelData, const Mat4f& transform, const Material& mat);
    // requireEmitter: "Scene must have at least one emissive object" (src/scene/Instances.cpp:125-127), only
    // meaningful when next-event estimation will run
    SceneTables build(bool requireEmitter = false);

private:
    struct PendingInstance { uint32_t propertiesID, materialIdx, objectID; Mat4f transform; };
    struct Cdf { std::vector<float> cdf; float area, weight; };
    static Cdf computeCDF(const ModelData& md, const Mat4f& transform, float brightness);

    std::vector<ModelData> modelData;
    std::vector<ModelRange> modelRanges;
    std::vector<float> allVertices, allTBNs, allTexCoords;
    std::vector<uint32_t> allIndices, allTBNsIndices, allTexIndices;
    std::vector<Image8> texturesToCreate;
    std::vector<PendingInstance> instancesToCreate;
    std::vector<RB200InstanceProperties> instanceProperties;
    std::vector<Material> materials;
    bool built = false;
};

}  // namespace rbhost
