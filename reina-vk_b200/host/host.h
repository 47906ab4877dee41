// The headless render loop: Reina::Reina() + Reina::renderLoop() (src/Reina.cpp:54-292, 295-392) with the window,
// swap chain and Vulkan plumbing removed and the GPU work delegated to the C ABI of include/reina_b200.h.
#pragma once
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

#include "../../include/reina_b200.h"
#include "config.h"
#include "save_manager.h"
#include "scene.h"

namespace rbhost {

// reina::tools::Clock's frame/sample bookkeeping (src/tools/Clock.cpp:27-40): the first markFrame() only arms the
// timer and does NOT count its samples, so getSampleCount() trails the image by two frames when the save check
// runs. Kept because the save thresholds of config.toml are defined against this counter.
class FrameClock {
public:
    explicit FrameClock(std::function<double()> now);
    double getAge() const { return now() - creationTime; }
    void markFrame(uint32_t samples);
    uint32_t getSampleCount() const { return samplesRecorded; }
    uint32_t getFrameCount() const { return frames; }
    double averageFrameTime() const { return frames ? frameTimeSum / frames : 0.0; }

private:
    std::function<double()> now;
    double creationTime;
    double lastFrameTime = -1.0;
    double frameTimeSum = 0.0;
    uint32_t samplesRecorded = 0, frames = 0;
};

// camera matrices + config values -> the push-constant block (src/Reina.cpp:142-155, src/graphics/Camera.cpp:10-11)
RB200RtPushConsts make_push_constants(const Config& cfg, float totalEmissiveWeight);

// built-in scenes: "cornell" (box + light, the worked example of src/Reina.cpp:108-111) and "cornell-sphere"
// (adds the Disney sphere of src/Reina.cpp:106); anything else throws
Scene make_builtin_scene(const std::string& name);

struct ObjRequest {
    std::string path;
    Material material;
    std::string texturePath;     // PNG albedo texture (Scene::defineTexture of a file: flipped vertically), may be empty
    std::string normalMapPath;   // PNG tangent-space normal map -> Material::normalMapID, may be empty
    std::string bumpMapPath;     // PNG height map for parallax mapping -> Material::bumpMapID, may be empty
};
// showroom-less generic scene: every OBJ at identity + the Cornell light panel (so NEE has an emitter)
Scene make_obj_scene(const std::vector<ObjRequest>& objs, bool addLight);

// reina::scene::gltf::loadScene (src/scene/gltf/gltfloader.cpp:440-456): the default scene of a .gltf / .glb file.
// The reference leaves the light panel commented out (:452) and then fails in Scene::build when nothing emits; here
// the Cornell light panel is added in that case when addLightIfDark is set (a warning says so).
Scene make_gltf_scene(const std::string& path, bool addLightIfDark, bool tangentsFromUv = false);

// RAII over RB200Context / RB200Scene; every failing call throws std::runtime_error(rb200_last_error())
class Renderer {
public:
    Renderer(uint32_t width, uint32_t height, int device, uint32_t flags);
    ~Renderer();
    Renderer(const Renderer&) = delete;
    Renderer& operator=(const Renderer&) = delete;

    void setScene(SceneTables& tables);
    void renderBatch(const RB200RtPushConsts& pc);
    void postprocess(const RB200BloomPushConsts& bloom, const RB200TonemappingPushConsts& tonemap);
    std::vector<uint8_t> readLdr();
    std::vector<float> readHdr();
    RB200BvhInfo bvhInfo() const;
    RB200Stats cumulativeStats();
    uint32_t width() const { return w; }
    uint32_t height() const { return h; }

private:
    uint32_t w, h;
    RB200Context* ctx = nullptr;
    RB200Scene* scene = nullptr;
};

struct LoopOptions {
    uint32_t totalSamples = 0;       // stop after this many samples per pixel (0: stop when every save threshold fired)
    double maxSeconds = 0.0;         // stop after this much wall time (0: no limit)
    std::string outputDir = ".";     // where SaveManager files go
    std::string finalOutput;         // written after the last frame when not empty
    bool quiet = false;
};

struct LoopResult {
    uint32_t frames = 0;
    uint32_t samples = 0;                    // samples per pixel actually accumulated
    double seconds = 0.0;
    std::vector<std::string> filesWritten;
    RB200Stats stats{};
};

LoopResult render_loop(Renderer& r, const Config& cfg, RB200RtPushConsts pc, const LoopOptions& opt);

// The same loop on several GPUs of this machine (rb200_group_*: a context and a host thread per device, the BVH
// replicated, one ncclReduce of the accumulation images per saved frame). Sample split: a frame is one batch per device
// (device g renders batch frame * n + g), so a frame adds n * samples_per_pixel samples; with `tiles` every device renders
// its interleaved tiles of one batch per frame instead (latency mode).
class GroupRenderer {
public:
    GroupRenderer(uint32_t width, uint32_t height, int numDevices, uint32_t flags, bool tiles);
    ~GroupRenderer();
    GroupRenderer(const GroupRenderer&) = delete;
    GroupRenderer& operator=(const GroupRenderer&) = delete;
    void setScene(SceneTables& tables);
    void renderFrame(const RB200RtPushConsts& pc, uint32_t firstBatch);
    void present(const RB200BloomPushConsts& bloom, const RB200TonemappingPushConsts& tonemap);
    std::vector<uint8_t> readLdr();
    RB200BvhInfo bvhInfo() const;
    RB200Stats cumulativeStats();
    int devices() const { return n; }
    bool tileMode() const { return tiles; }
    uint32_t width() const { return w; }
    uint32_t height() const { return h; }

private:
    uint32_t w, h;
    int n;
    bool tiles;
    RB200Group* group = nullptr;
    RB200GroupScene* scene = nullptr;
};

LoopResult render_loop_group(GroupRenderer& r, const Config& cfg, RB200RtPushConsts pc, const LoopOptions& opt);

}  // namespace rbhost
