// reina_b200: headless command-line renderer on top of librb200.so.
//
// Stands where the reference's main() + Reina::Reina() + renderLoop() stand (src/main.cpp, src/Reina.cpp:54-392),
// minus GLFW / the swap chain: it reads config.toml, builds the scene tables, renders sample batches and writes
// PNGs under the SaveManager policy (or one final image with --spp/--out).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "config.h"
#include "host.h"

using namespace rbhost;

static void usage() {
    std::puts(
        "usage: reina_b200 [options]\n"
        "  --config FILE        TOML configuration (default: config/config.toml)\n"
        "  --scene NAME         built-in scene: cornell | cornell-sphere (overrides [render].scene)\n"
        "  --obj FILE           render this OBJ instead of a built-in scene (repeatable); the Cornell light panel is added\n"
        "  --gltf FILE          render the default scene of a glTF 2.0 file (.gltf or .glb), the reference's default scene path;\n"
        "                       when none of its materials emits and NEE is on, the Cornell light panel is added\n"
        "  --gltf-tangents RULE tangent frames of glTF primitives without TANGENT: reference (default: what the reference's\n"
        "                       MikkTSpace call yields as written, the constant frame (1,0,0) / (0,1,0)) | uv (from the UV derivatives)\n"
        "  --material NAME      material of the following --obj: lambertian | metal | dielectric | disney (default lambertian)\n"
        "  --albedo R,G,B       albedo of the following --obj (default 0.8,0.8,0.8)\n"
        "  --roughness X  --ior X  --metallic X     parameters of the following --obj\n"
        "  --texture FILE       PNG / JPEG albedo texture of the following --obj (flipped vertically like the reference's file textures)\n"
        "  --normal-map FILE    PNG / JPEG tangent-space normal map of the following --obj\n"
        "  --bump-map FILE      PNG / JPEG height map of the following --obj (parallax mapping; black = surface, white = deepest)\n"
        "  --width N --height N image size (overrides [render])\n"
        "  --spp N              stop after N samples per pixel in total\n"
        "  --seconds S          stop after S seconds\n"
        "  --out FILE           write the final image to FILE (PNG)\n"
        "  --outdir DIR         directory for the output_<N>spp.png / output_<T>sec.png files of [saving] (default .)\n"
        "  --nee | --no-nee     next-event estimation on / off (overrides [render].nee)\n"
        "  --device N           CUDA device (default 0)\n"
        "  --gpus N             render on devices 0 .. N-1 of this machine: the BVH is replicated, device g renders sample batches\n"
        "                       g, g + N, ... and one NCCL reduce of the accumulation images closes every saved frame\n"
        "  --tiles              with --gpus: latency mode, every device renders its interleaved 32 x 32 tiles of every batch\n"
        "  --two-level          keep one hierarchy per object and a top level over the instances (the reference's BLAS / TLAS,\n"
        "                       src/scene/Scene.cpp:93-111) instead of flattening the instances: less memory, slower rays\n"
        "  --skip-null-shadow-rays  with NEE: shadow rays whose contribution is exactly zero before the visibility test (light facing\n"
        "                       away or below the horizon) are answered without a traversal; same image, less work\n"
        "  --dump-pc FILE       write the 160-byte push-constant block of batch 0 to FILE\n"
        "  --quiet              no progress output\n");
}

static bool parse_vec3(const char* s, std::array<float, 3>& out) {
    float a, b, c;
    if (std::sscanf(s, "%f,%f,%f", &a, &b, &c) != 3) return false;
    out = {a, b, c};
    return true;
}

int main(int argc, char** argv) {
    std::string configPath = "config/config.toml", sceneName, finalOut, outDir = ".", dumpPc, gltfPath;
    bool gltfTangentsUv = false;
    std::vector<ObjRequest> objs;
    std::string nextTexture, nextNormalMap, nextBumpMap;
    Material nextMat;
    nextMat.albedo = {0.8f, 0.8f, 0.8f};
    nextMat.interpNormals = true;
    long width = -1, height = -1, spp = 0, device = 0, gpus = 1;
    bool tiles = false, twoLevel = false, skipNullShadow = false;
    double seconds = 0.0;
    int nee = -1;
    bool quiet = false;

    try {
        for (int i = 1; i < argc; i++) {
            const std::string a = argv[i];
            auto value = [&]() -> const char* {
                if (i + 1 >= argc) throw std::runtime_error("option " + a + " needs a value");
                return argv[++i];
            };
            if (a == "--help" || a == "-h") { usage(); return 0; }
            else if (a == "--config") configPath = value();
            else if (a == "--scene") sceneName = value();
            else if (a == "--obj") {
                objs.push_back({value(), nextMat, nextTexture, nextNormalMap, nextBumpMap});
                nextTexture.clear(); nextNormalMap.clear(); nextBumpMap.clear();
            }
            else if (a == "--gltf") gltfPath = value();
            else if (a == "--gltf-tangents") {
                const std::string r = value();
                if (r == "reference") gltfTangentsUv = false;
                else if (r == "uv") gltfTangentsUv = true;
                else throw std::runtime_error("--gltf-tangents: reference | uv");
            }
            else if (a == "--texture") nextTexture = value();
            else if (a == "--normal-map") nextNormalMap = value();
            else if (a == "--bump-map") nextBumpMap = value();
            else if (a == "--material") {
                const std::string m = value();
                if (m == "lambertian") nextMat.materialIdx = 0;
                else if (m == "metal") nextMat.materialIdx = 1;
                else if (m == "dielectric") nextMat.materialIdx = 2;
                else if (m == "disney") nextMat.materialIdx = 3;
                else throw std::runtime_error("unknown material '" + m + "'");
            }
            else if (a == "--albedo") { if (!parse_vec3(value(), nextMat.albedo)) throw std::runtime_error("--albedo wants R,G,B"); }
            else if (a == "--roughness") nextMat.roughness = std::strtof(value(), nullptr);
            else if (a == "--ior") nextMat.ior = std::strtof(value(), nullptr);
            else if (a == "--metallic") nextMat.metallic = std::strtof(value(), nullptr);
            else if (a == "--width") width = std::strtol(value(), nullptr, 10);
            else if (a == "--height") height = std::strtol(value(), nullptr, 10);
            else if (a == "--spp") spp = std::strtol(value(), nullptr, 10);
            else if (a == "--seconds") seconds = std::strtod(value(), nullptr);
            else if (a == "--out") finalOut = value();
            else if (a == "--outdir") outDir = value();
            else if (a == "--nee") nee = 1;
            else if (a == "--no-nee") nee = 0;
            else if (a == "--device") device = std::strtol(value(), nullptr, 10);
            else if (a == "--gpus") gpus = std::strtol(value(), nullptr, 10);
            else if (a == "--tiles") tiles = true;
            else if (a == "--two-level") twoLevel = true;
            else if (a == "--skip-null-shadow-rays") skipNullShadow = true;
            else if (a == "--dump-pc") dumpPc = value();
            else if (a == "--quiet") quiet = true;
            else throw std::runtime_error("unknown option " + a + " (see --help)");
        }

        Config cfg = Config::from_toml(TomlDoc::parse_file(configPath));
        if (width > 0) cfg.width = uint32_t(width);
        if (height > 0) cfg.height = uint32_t(height);
        if (!sceneName.empty()) cfg.scene = sceneName;
        if (nee >= 0) cfg.nee = nee != 0;
        if (spp < 0 || seconds < 0) throw std::runtime_error("--spp and --seconds must not be negative");

        if (!gltfPath.empty() && !objs.empty()) throw std::runtime_error("--gltf and --obj cannot be combined");
        Scene scene = !gltfPath.empty() ? make_gltf_scene(gltfPath, cfg.nee, gltfTangentsUv)
                      : objs.empty()    ? make_builtin_scene(cfg.scene)
                                        : make_obj_scene(objs, true);
        SceneTables tables = scene.build(cfg.nee);
        RB200RtPushConsts pc = make_push_constants(cfg, tables.totalEmissiveWeight);
        if (!dumpPc.empty()) {
            std::ofstream f(dumpPc, std::ios::binary);
            f.write(reinterpret_cast<const char*>(&pc), sizeof pc);
            if (!f) throw std::runtime_error("cannot write " + dumpPc);
        }

        LoopOptions opt;
        opt.totalSamples = uint32_t(spp);
        opt.maxSeconds = seconds;
        opt.outputDir = outDir;
        opt.finalOutput = finalOut;
        opt.quiet = quiet;
        if (gpus < 1) throw std::runtime_error("--gpus must be at least 1");
        if (gpus > 1 || tiles) {
            GroupRenderer group(cfg.width, cfg.height, int(gpus), (cfg.nee ? uint32_t(RB200_FLAG_NEE) : 0u) | (twoLevel ? uint32_t(RB200_FLAG_TWO_LEVEL) : 0u) | (skipNullShadow ? uint32_t(RB200_FLAG_SKIP_NULL_SHADOW_RAYS) : 0u), tiles);
            group.setScene(tables);
            if (!quiet) {
                const RB200BvhInfo bvh = group.bvhInfo();
                std::printf("scene: %llu triangles, %u wide nodes, depth %u, built in %.2f ms on each of %ld devices (equal hashes)\n",
                            static_cast<unsigned long long>(tables.numTriangles()), bvh.numWideNodes, bvh.maxDepth, bvh.buildMs, gpus);
            }
            const LoopResult res = render_loop_group(group, cfg, pc, opt);
            if (!quiet) {
                const double rays = double(res.stats.extendRays + res.stats.shadowRays);
                std::printf("%u frames, %u samples per pixel, %.3f s, %.1f Mrays/s on %ld devices, %zu file(s) written\n", res.frames,
                            res.samples, res.seconds, res.seconds > 0 ? rays / res.seconds * 1e-6 : 0.0, gpus, res.filesWritten.size());
            }
            return 0;
        }

        Renderer renderer(cfg.width, cfg.height, int(device), (cfg.nee ? uint32_t(RB200_FLAG_NEE) : 0u) | (twoLevel ? uint32_t(RB200_FLAG_TWO_LEVEL) : 0u) | (skipNullShadow ? uint32_t(RB200_FLAG_SKIP_NULL_SHADOW_RAYS) : 0u));
        renderer.setScene(tables);
        if (!quiet) {
            const RB200BvhInfo bvh = renderer.bvhInfo();
            std::printf("scene: %llu triangles, %u wide nodes, depth %u, built in %.2f ms\n",
                        static_cast<unsigned long long>(tables.numTriangles()), bvh.numWideNodes, bvh.maxDepth, bvh.buildMs);
        }

        const LoopResult res = render_loop(renderer, cfg, pc, opt);
        if (!quiet) {
            const double rays = double(res.stats.extendRays + res.stats.shadowRays);
            std::printf("%u frames, %u samples per pixel, %.3f s, %.1f Mrays/s, %zu file(s) written\n", res.frames, res.samples,
                        res.seconds, res.seconds > 0 ? rays / res.seconds * 1e-6 : 0.0, res.filesWritten.size());
        }
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "reina_b200: %s\n", e.what());
        return 1;
    }
}
