// Test hooks: a small extern "C" surface over the host classes so tests/test_cpp_host.py can compare the C++ host
// with the Python host table by table. Not part of the product ABI (that is include/reina_b200.h).
#include <cstring>
#include <string>

#include "config.h"
#include "gltf.h"
#include "host.h"
#include "png.h"
#include "texture.h"

using namespace rbhost;

namespace {
thread_local std::string g_error;

template <class F>
int guarded(F&& f) {
    try {
        f();
        g_error.clear();
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return -1;
    }
}
}  // namespace

extern "C" {

struct RBHostConfig {
    uint32_t width, height, nee, samplesPerPixel, maxBounces, numSaveSamples, numSaveTimes;
    float focusDist, defocusMultiplier, directClamp, indirectClamp;
    float bloomRadius, bloomThreshold, bloomIntensity, exposure;
    double cameraPos[3], cameraLookAt[3], fovYDegrees;
    int32_t saveSamples[16];
    double saveTimes[16];
    char scene[64];
};

const char* rbhost_last_error() { return g_error.c_str(); }

int rbhost_config_parse(const char* tomlText, RBHostConfig* out) {
    return guarded([&] {
        Config c = Config::from_toml(TomlDoc::parse(tomlText));
        std::memset(out, 0, sizeof *out);
        out->width = c.width; out->height = c.height; out->nee = c.nee;
        out->samplesPerPixel = c.samplesPerPixel; out->maxBounces = c.maxBounces;
        out->focusDist = c.focusDist; out->defocusMultiplier = c.defocusMultiplier;
        out->directClamp = c.directClamp; out->indirectClamp = c.indirectClamp;
        out->bloomRadius = c.bloom.radius; out->bloomThreshold = c.bloom.threshold;
        out->bloomIntensity = c.bloom.intensity; out->exposure = c.tonemap.exposure;
        for (int k = 0; k < 3; k++) { out->cameraPos[k] = c.cameraPos[k]; out->cameraLookAt[k] = c.cameraLookAt[k]; }
        out->fovYDegrees = c.fovYDegrees;
        out->numSaveSamples = uint32_t(c.saveOnSamples.size());
        out->numSaveTimes = uint32_t(c.saveOnTimes.size());
        for (size_t i = 0; i < c.saveOnSamples.size() && i < 16; i++) out->saveSamples[i] = c.saveOnSamples[i];
        for (size_t i = 0; i < c.saveOnTimes.size() && i < 16; i++) out->saveTimes[i] = c.saveOnTimes[i];
        std::strncpy(out->scene, c.scene.c_str(), sizeof out->scene - 1);
    });
}

int rbhost_push_constants(const char* tomlText, float totalEmissiveWeight, RB200RtPushConsts* out) {
    return guarded([&] { *out = make_push_constants(Config::from_toml(TomlDoc::parse(tomlText)), totalEmissiveWeight); });
}

int rbhost_tables_builtin(const char* name, int requireEmitter, SceneTables** out) {
    return guarded([&] {
        Scene s = make_builtin_scene(name);
        *out = new SceneTables(s.build(requireEmitter != 0));
    });
}

int rbhost_tables_obj(const char* path, uint32_t materialIdx, int addLight, SceneTables** out) {
    return guarded([&] {
        Material m;
        m.materialIdx = materialIdx;
        m.albedo = {0.8f, 0.8f, 0.8f};
        m.interpNormals = true;
        Scene s = make_obj_scene({{path, m, "", "", ""}}, addLight != 0);
        *out = new SceneTables(s.build(false));
    });
}

int rbhost_tables_gltf(const char* path, int requireEmitter, SceneTables** out) {
    return guarded([&] {
        Scene s = load_gltf_scene(path);
        *out = new SceneTables(s.build(requireEmitter != 0));
    });
}

// tangentRule: 0 = the reference's MikkTSpace call as written, 1 = frames from the UV derivatives (host/gltf.h)
int rbhost_tables_gltf_tangents(const char* path, int requireEmitter, int tangentRule, SceneTables** out) {
    return guarded([&] {
        if (tangentRule != 0 && tangentRule != 1) throw std::runtime_error("tangentRule must be 0 (reference) or 1 (uv)");
        Scene s = load_gltf_scene(path, nullptr, tangentRule == 0 ? GltfTangents::Reference : GltfTangents::Uv);
        *out = new SceneTables(s.build(requireEmitter != 0));
    });
}

int rbhost_tables_desc(SceneTables* t, RB200SceneDesc* out, float* totalEmissiveWeight) {
    return guarded([&] {
        *out = t->desc();
        *totalEmissiveWeight = t->totalEmissiveWeight;
    });
}

void rbhost_tables_free(SceneTables* t) { delete t; }

// returns the encoded size; writes at most `capacity` bytes
int64_t rbhost_png_encode(const uint8_t* rgba, uint32_t width, uint32_t height, uint8_t* out, uint64_t capacity) {
    int64_t n = -1;
    guarded([&] {
        std::vector<uint8_t> png = encode_png_rgba8(rgba, width, height);
        n = int64_t(png.size());
        std::memcpy(out, png.data(), png.size() < capacity ? png.size() : capacity);
    });
    return n;
}

// decodes a PNG file to RGBA8; returns 0 and the size, writes at most `capacity` bytes
int rbhost_png_load(const char* path, int flip, uint8_t* out, uint64_t capacity, uint32_t* width, uint32_t* height) {
    return guarded([&] {
        Image8 img = load_png_rgba8(path, flip != 0);
        *width = uint32_t(img.width);
        *height = uint32_t(img.height);
        std::memcpy(out, img.rgba.data(), img.rgba.size() < capacity ? img.rgba.size() : capacity);
    });
}

// decodes a PNG or JPEG file (by signature) to RGBA8
int rbhost_image_load(const char* path, int flip, uint8_t* out, uint64_t capacity, uint32_t* width, uint32_t* height) {
    return guarded([&] {
        Image8 img = load_image_rgba8(path, flip != 0);
        *width = uint32_t(img.width);
        *height = uint32_t(img.height);
        std::memcpy(out, img.rgba.data(), img.rgba.size() < capacity ? img.rgba.size() : capacity);
    });
}

// Drives SaveManager + FrameClock through `frames` frames of `spp` samples, `secondsPerFrame` apart, in the order of
// the render loop (save check, then markFrame). Writes "frame:filename\n" lines into `out`.
int rbhost_save_schedule(const int32_t* saveSamples, uint32_t nSamples, const double* saveTimes, uint32_t nTimes,
                         uint32_t frames, uint32_t spp, double secondsPerFrame, char* out, uint64_t capacity) {
    return guarded([&] {
        double t = 0.0;
        FrameClock clock([&t] { return t; });
        SaveManager mgr(std::vector<int>(saveSamples, saveSamples + nSamples), std::vector<double>(saveTimes, saveTimes + nTimes));
        std::string log;
        for (uint32_t f = 0; f < frames; f++) {
            t += secondsPerFrame;
            SaveInfo info = mgr.shouldSave(clock.getSampleCount(), clock.getAge());
            if (info.shouldSave) log += std::to_string(f) + ":" + info.filename + "\n";
            clock.markFrame(spp);
        }
        if (log.size() + 1 > capacity) throw std::runtime_error("schedule log does not fit");
        std::memcpy(out, log.c_str(), log.size() + 1);
    });
}

}  // extern "C"
