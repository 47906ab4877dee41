// See host.h.
#include "host.h"

#include <chrono>
#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "png.h"
#include "gltf.h"
#include "texture.h"
#include "rh_math.h"

namespace rbhost {

FrameClock::FrameClock(std::function<double()> nowFn) : now(std::move(nowFn)), creationTime(now()) {}

void FrameClock::markFrame(uint32_t samples) {
    if (lastFrameTime < 0.0) {   // first call: arm only (src/tools/Clock.cpp:30-33)
        lastFrameTime = now();
        return;
    }
    const double t = now();
    frameTimeSum += t - lastFrameTime;
    lastFrameTime = t;
    frames++;
    samplesRecorded += samples;
}

RB200RtPushConsts make_push_constants(const Config& cfg, float totalEmissiveWeight) {
    RB200RtPushConsts pc{};
    const Vec3d pos{cfg.cameraPos[0], cfg.cameraPos[1], cfg.cameraPos[2]};
    const Vec3d look{cfg.cameraLookAt[0], cfg.cameraLookAt[1], cfg.cameraLookAt[2]};
    const double fovy = cfg.fovYDegrees * (3.141592653589793 / 180.0);
    const Mat4f invView = to_float(inverse(look_at(pos, look)));
    const Mat4f invProj = to_float(inverse(perspective(fovy, double(cfg.width) / double(cfg.height), 0.1, 100.0)));
    std::memcpy(pc.invView, invView.data(), sizeof pc.invView);
    std::memcpy(pc.invProjection, invProj.data(), sizeof pc.invProjection);
    pc.sampleBatch = 0;
    pc.totalEmissiveWeight = totalEmissiveWeight;
    pc.focusDist = cfg.focusDist;
    pc.defocusMultiplier = cfg.defocusMultiplier / 100.0f;   // src/Reina.cpp:150
    pc.directClamp = cfg.directClamp;
    pc.indirectClamp = cfg.indirectClamp;
    pc.samplesPerPixel = cfg.samplesPerPixel;
    pc.maxBounces = cfg.maxBounces;
    return pc;
}

static Material cornell_wall(int textureID) {
    Material m;   // `cornellWall`, src/Reina.cpp:108
    m.materialIdx = 0;
    m.textureID = textureID;
    m.albedo = {0.9f, 0.9f, 0.9f};
    m.cullBackface = true;
    return m;
}

static Material light_material() {
    Material m;   // `lightMaterial`, src/Reina.cpp:109
    m.materialIdx = 0;
    m.albedo = {0.9f, 0.9f, 0.9f};
    m.emission = {16.0f, 16.0f, 16.0f};
    m.cullBackface = true;
    return m;
}

Scene make_builtin_scene(const std::string& name) {
    if (name != "cornell" && name != "cornell-sphere")
        throw std::runtime_error("unknown built-in scene '" + name + "' (cornell | cornell-sphere)");
    Scene s;
    const uint32_t tex = s.defineTexture(cornell_texture(256, 384));
    s.addObject(cornell_box(), identity(), cornell_wall(int(tex)));
    s.addObject(cornell_light(), identity(), light_material());
    if (name == "cornell-sphere") {
        // glm::translate(glm::scale(I, .25), (0, 2, 0)), src/Reina.cpp:106
        const Mat4f t = to_float(mul(scale_d(0.25, 0.25, 0.25), translate_d(0.0, 2.0, 0.0)));
        Material m;
        m.materialIdx = 3;
        m.albedo = {1.0f, 1.0f, 1.0f};
        m.roughness = 0.5f;
        m.ior = 1.5f;
        m.interpNormals = true;
        m.sheenTint = {1.0f, 1.0f, 1.0f};
        s.addObject(uv_sphere(48, 24, 1.0), t, m);
    }
    return s;
}

Scene make_obj_scene(const std::vector<ObjRequest>& objs, bool addLight) {
    Scene s;
    for (const ObjRequest& o : objs) {
        bool hasUv = true;
        ModelData md = load_obj(o.path, &hasUv);
        if (!hasUv && o.material.materialIdx == 3)
            std::fprintf(stderr, "warning: %s has no texture coordinates; tangents are zero and Disney shading will be NaN on it "
                                 "(same as the reference, src/scene/Models.cpp:144-152)\n", o.path.c_str());
        Material m = o.material;
        if (!o.texturePath.empty()) m.textureID = int(s.defineTexture(load_image_rgba8(o.texturePath, true)));
        if (!o.normalMapPath.empty()) m.normalMapID = int(s.defineTexture(load_image_rgba8(o.normalMapPath, true)));
        if (!o.bumpMapPath.empty()) m.bumpMapID = int(s.defineTexture(load_image_rgba8(o.bumpMapPath, true)));
        s.addObject(md, identity(), m);
    }
    if (addLight) s.addObject(cornell_light(), identity(), light_material());
    return s;
}

Scene make_gltf_scene(const std::string& path, bool addLightIfDark, bool tangentsFromUv) {
    bool emits = false;
    Scene s = load_gltf_scene(path, &emits, tangentsFromUv ? GltfTangents::Uv : GltfTangents::Reference);
    if (!emits && addLightIfDark) {
        std::fprintf(stderr, "warning: %s has no emissive material; adding the Cornell light panel so that next-event "
                             "estimation has an emitter\n", path.c_str());
        s.addObject(cornell_light(), identity(), light_material());
    }
    return s;
}

static void check(int rc, const char* what) {
    if (rc != RB200_OK) throw std::runtime_error(std::string(what) + ": " + rb200_last_error());
}

Renderer::Renderer(uint32_t width, uint32_t height, int device, uint32_t flags) : w(width), h(height) {
    check(rb200_context_create(width, height, device, flags, &ctx), "rb200_context_create");
}

Renderer::~Renderer() {
    if (scene) rb200_scene_destroy(scene);
    if (ctx) rb200_context_destroy(ctx);
}

void Renderer::setScene(SceneTables& tables) {
    if (scene) { rb200_scene_destroy(scene); scene = nullptr; }
    RB200SceneDesc d = tables.desc();
    if (d.numEmissive > 1)
        std::fprintf(stderr, "warning: %u emissive instances: with next-event estimation the reference's light sampling "
                             "(nee.h.glsl:97-105) addresses the triangles of every emitter after the first through the "
                             "concatenated triangle CDF, i.e. beyond the emitter's own triangles; reproduced as is (DESIGN.md 2)\n",
                     d.numEmissive);
    check(rb200_scene_create(ctx, &d, &scene), "rb200_scene_create");
}

void Renderer::renderBatch(const RB200RtPushConsts& pc) {
    if (!scene) throw std::runtime_error("renderBatch: no scene set");
    check(rb200_render_batch(ctx, scene, &pc), "rb200_render_batch");
    // calls are asynchronous: without back-pressure the host clock (--seconds, the save_on_times thresholds) would run
    // ahead of the device by the whole driver queue. Keep at most rb200_pipeline_depth() batches in flight.
    check(rb200_wait_batches_pending(ctx, rb200_pipeline_depth()), "rb200_wait_batches_pending");
}

void Renderer::postprocess(const RB200BloomPushConsts& bloom, const RB200TonemappingPushConsts& tonemap) {
    check(rb200_postprocess(ctx, &bloom, &tonemap), "rb200_postprocess");
}

std::vector<uint8_t> Renderer::readLdr() {
    std::vector<uint8_t> px(size_t(w) * h * 4);
    check(rb200_read_ldr(ctx, px.data()), "rb200_read_ldr");
    return px;
}

std::vector<float> Renderer::readHdr() {
    std::vector<float> px(size_t(w) * h * 4);
    check(rb200_read_hdr(ctx, px.data()), "rb200_read_hdr");
    return px;
}

RB200BvhInfo Renderer::bvhInfo() const {
    RB200BvhInfo info{};
    if (!scene) throw std::runtime_error("bvhInfo: no scene set");
    check(rb200_scene_bvh_info(scene, &info), "rb200_scene_bvh_info");
    return info;
}

RB200Stats Renderer::cumulativeStats() {
    RB200Stats last{}, cum{};
    check(rb200_get_stats(ctx, &last, &cum), "rb200_get_stats");
    return cum;
}

// Reina::renderLoop (src/Reina.cpp:295-392): trace -> bloom -> tonemap -> save check -> markFrame, with
// sampleBatch incremented after every trace (src/Reina.cpp:431-432). Differences: no window or camera input, the
// loop ends on a sample / time budget instead of the window closing, and post-processing only runs on frames
// that are going to be read (it does not feed back into the accumulation image).
LoopResult render_loop(Renderer& r, const Config& cfg, RB200RtPushConsts pc, const LoopOptions& opt) {
    using clk = std::chrono::steady_clock;
    const auto t0 = clk::now();
    auto now = [t0] { return std::chrono::duration<double>(clk::now() - t0).count(); };
    FrameClock clock(now);
    SaveManager saves(cfg.saveOnSamples, cfg.saveOnTimes);
    LoopResult res;
    const std::string dir = opt.outputDir.empty() ? std::string(".") : opt.outputDir;
    if (pc.samplesPerPixel == 0) throw std::runtime_error("sampling.samples_per_pixel must be at least 1");

    for (;;) {
        r.renderBatch(pc);
        pc.sampleBatch++;
        res.frames++;
        res.samples += pc.samplesPerPixel;

        SaveInfo info = saves.shouldSave(clock.getSampleCount(), clock.getAge());
        if (info.shouldSave) {
            r.postprocess(cfg.bloom, cfg.tonemap);
            std::vector<uint8_t> px = r.readLdr();
            const std::string file = dir + "/" + info.filename;
            write_png_rgba8(file, px.data(), r.width(), r.height());
            res.filesWritten.push_back(file);
            if (!opt.quiet) std::printf("saved %s (%u samples per pixel in the image)\n", file.c_str(), res.samples);
        }
        clock.markFrame(pc.samplesPerPixel);

        if (opt.totalSamples && res.samples >= opt.totalSamples) break;
        if (opt.maxSeconds > 0.0 && now() >= opt.maxSeconds) break;
        if (!opt.totalSamples && opt.maxSeconds <= 0.0 && !saves.pending()) break;
    }
    if (!opt.finalOutput.empty()) {
        r.postprocess(cfg.bloom, cfg.tonemap);
        std::vector<uint8_t> px = r.readLdr();
        write_png_rgba8(opt.finalOutput, px.data(), r.width(), r.height());
        res.filesWritten.push_back(opt.finalOutput);
    }
    res.stats = r.cumulativeStats();   // synchronises
    res.seconds = now();
    return res;
}

GroupRenderer::GroupRenderer(uint32_t width, uint32_t height, int numDevices, uint32_t flags, bool tiles_)
    : w(width), h(height), n(numDevices), tiles(tiles_) {
    std::vector<int> devs(size_t(numDevices > 0 ? numDevices : 0));
    for (int i = 0; i < numDevices; i++) devs[size_t(i)] = i;
    check(rb200_group_create(width, height, devs.data(), numDevices, flags | (tiles ? uint32_t(RB200_FLAG_GROUP_TILES) : 0u), &group),
          "rb200_group_create");
}

GroupRenderer::~GroupRenderer() {
    if (scene) rb200_group_scene_destroy(scene);
    if (group) rb200_group_destroy(group);
}

void GroupRenderer::setScene(SceneTables& tables) {
    if (scene) { rb200_group_scene_destroy(scene); scene = nullptr; }
    RB200SceneDesc d = tables.desc();
    check(rb200_group_scene_create(group, &d, &scene), "rb200_group_scene_create");
}

void GroupRenderer::renderFrame(const RB200RtPushConsts& pc, uint32_t firstBatch) {
    if (!scene) throw std::runtime_error("renderFrame: no scene set");
    check(rb200_group_render_batches(group, scene, &pc, firstBatch, 1), "rb200_group_render_batches");
}

void GroupRenderer::present(const RB200BloomPushConsts& bloom, const RB200TonemappingPushConsts& tonemap) {
    check(rb200_group_present(group, &bloom, &tonemap), "rb200_group_present");
}

std::vector<uint8_t> GroupRenderer::readLdr() {
    std::vector<uint8_t> px(size_t(w) * h * 4);
    check(rb200_group_read_ldr(group, px.data()), "rb200_group_read_ldr");
    return px;
}

RB200BvhInfo GroupRenderer::bvhInfo() const {
    RB200BvhInfo info{};
    if (!scene) throw std::runtime_error("bvhInfo: no scene set");
    check(rb200_group_scene_bvh_info(scene, 0, &info), "rb200_group_scene_bvh_info");
    return info;
}

RB200Stats GroupRenderer::cumulativeStats() {
    RB200Stats cum{};
    check(rb200_group_get_stats(group, &cum), "rb200_group_get_stats");
    return cum;
}

LoopResult render_loop_group(GroupRenderer& r, const Config& cfg, RB200RtPushConsts pc, const LoopOptions& opt) {
    using clk = std::chrono::steady_clock;
    const auto t0 = clk::now();
    auto now = [t0] { return std::chrono::duration<double>(clk::now() - t0).count(); };
    FrameClock clock(now);
    SaveManager saves(cfg.saveOnSamples, cfg.saveOnTimes);
    LoopResult res;
    const std::string dir = opt.outputDir.empty() ? std::string(".") : opt.outputDir;
    if (pc.samplesPerPixel == 0) throw std::runtime_error("sampling.samples_per_pixel must be at least 1");
    const uint32_t batchesPerFrame = r.tileMode() ? 1u : uint32_t(r.devices());
    const uint32_t samplesPerFrame = pc.samplesPerPixel * batchesPerFrame;
    uint32_t nextBatch = 0;
    for (;;) {
        r.renderFrame(pc, nextBatch);
        nextBatch += batchesPerFrame;
        res.frames++;
        res.samples += samplesPerFrame;
        SaveInfo info = saves.shouldSave(clock.getSampleCount(), clock.getAge());
        if (info.shouldSave) {
            r.present(cfg.bloom, cfg.tonemap);
            std::vector<uint8_t> px = r.readLdr();
            const std::string file = dir + "/" + info.filename;
            write_png_rgba8(file, px.data(), r.width(), r.height());
            res.filesWritten.push_back(file);
            if (!opt.quiet) std::printf("saved %s (%u samples per pixel in the image)\n", file.c_str(), res.samples);
        }
        clock.markFrame(samplesPerFrame);
        if (opt.totalSamples && res.samples >= opt.totalSamples) break;
        if (opt.maxSeconds > 0.0 && now() >= opt.maxSeconds) break;
        if (!opt.totalSamples && opt.maxSeconds <= 0.0 && !saves.pending()) break;
    }
    if (!opt.finalOutput.empty()) {
        r.present(cfg.bloom, cfg.tonemap);
        std::vector<uint8_t> px = r.readLdr();
        write_png_rgba8(opt.finalOutput, px.data(), r.width(), r.height());
        res.filesWritten.push_back(opt.finalOutput);
    }
    res.stats = r.cumulativeStats();   // synchronises
    res.seconds = now();
    return res;
}

}  // namespace rbhost
