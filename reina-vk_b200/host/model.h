// Geometry sources of the headless host: the ModelData contract of the reference's importer
// (reina::scene::ModelData, src/scene/Models.h:23-30; Models::getObjData, src/scene/Models.cpp:117-175) and the
// built-in Cornell assets. Rules are the ones documented in reina-vk_b200/meshes.py; tests/test_cpp_host.py checks
// that both importers produce identical tables from the same file.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace rbhost {

struct ModelData {
    std::vector<float> vertices;        // xyzw, w = 1
    std::vector<uint32_t> indices;
    std::vector<float> tbns;            // 9 floats per entry: T, B, N columns
    std::vector<uint32_t> tbnsIndices;
    std::vector<float> texCoords;       // uv pairs (V already flipped); empty when the mesh has none
    std::vector<uint32_t> texIndices;
    size_t numVertices() const { return vertices.size() / 4; }
    size_t numTriangles() const { return indices.size() / 3; }
};

// pos: 3 floats per vertex; uv: 2 per vertex or empty; nrm: 3 per vertex; tris: 3 indices per triangle
ModelData make_model(const std::vector<float>& pos, const std::vector<float>& uv, const std::vector<float>& nrm,
                     const std::vector<uint32_t>& tris);

std::vector<float> smooth_normals(const std::vector<float>& pos, const std::vector<uint32_t>& tris);

// Wavefront OBJ: fan triangulation in file order, one vertex per distinct v/vt/vn triple, V flipped, smooth normals
// generated when absent, tangent frames from UV derivatives. Throws std::runtime_error when the file cannot be
// read or holds no faces ("Could not load model", as Models.cpp:123-125).
ModelData load_obj(const std::string& path, bool* hasTexCoords = nullptr);

ModelData cornell_box();
ModelData cornell_light();
ModelData uv_sphere(int segments, int rings, double radius);

struct Image8 {
    int width = 0, height = 0;
    std::vector<uint8_t> rgba;
};
Image8 cornell_texture(int width, int height);

}  // namespace rbhost
