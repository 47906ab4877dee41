// glTF 2.0 / GLB scene import for the headless host — reina::scene::gltf (src/scene/gltf/gltfloader.cpp:72-456),
// the reference's default scene path (src/Reina.cpp:91), without fastgltf, MikkTSpace and the Vulkan handles.
// SURVEY.md 8f row 3. The rules (mesh order, the two tangent rules, TRS composition in double precision, PNG-only images,
// file images flipped / embedded images not) are documented in reina-vk_b200/gltf.py; tests/test_gltf.py checks that
// both importers produce identical tables from the same asset.
#pragma once
#include <string>

#include "scene.h"

namespace rbhost {

// loadScene (gltfloader.cpp:440-456): meshes of the default scene -> objects, images -> textures, glTF PBR materials
// -> Disney materials, one instance per primitive per node. Throws std::runtime_error with the reference's messages
// ("Failed to find glTF file: ...", "Failed to parse glTF: ...", "No scenes supplied in gLTF file",
// "Meshes without vertex normals are not supported", ...). hasEmitter (optional) reports whether any instance emits.
//
// tangents: what a primitive without a TANGENT attribute gets. Reference = the outcome of the reference's MikkTSpace call
// exactly as it is written (gltfloader.cpp:207-222: un-indexed vertex triples, UVs still zero, hence the constant frame
// (1,0,0) / (0,1,0) — see Primitive::mikktspace_as_called); Uv = per-vertex frames from the UV derivatives.
enum class GltfTangents { Reference = 0, Uv = 1 };
Scene load_gltf_scene(const std::string& path, bool* hasEmitter = nullptr, GltfTangents tangents = GltfTangents::Reference);

}  // namespace rbhost
