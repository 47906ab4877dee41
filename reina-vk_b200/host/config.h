// Configuration of the headless renderer.
//
// The reference reads config/config.toml with toml++ (an un-vendored dependency) and pulls every value with
// `config.at_path("a.b").value<T>().value()` (src/Reina.cpp:142-155, 246-262, src/tools/SaveManager.cpp:6-25): a missing
// key or a value of the wrong type ends start-up with std::bad_optional_access. TomlDoc is a reader for the subset of
// TOML that file uses — comments, [dotted.table] headers, `key = value` with integers, floats, booleans, basic
// strings and (possibly multi-line) arrays of those — and get<T>() keeps the all-keys-required rule, raising
// ConfigError that names the key instead.
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/reina_b200.h"

namespace rbhost {

struct ConfigError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct TomlValue {
    enum Kind { Integer, Float, Boolean, String, Array } kind = Integer;
    int64_t i = 0;
    double f = 0.0;
    bool b = false;
    std::string s;
    std::vector<TomlValue> items;
};

class TomlDoc {
public:
    static TomlDoc parse(const std::string& text);
    static TomlDoc parse_file(const std::string& path);

    bool has(const std::string& path) const { return values.count(path) != 0; }
    const TomlValue& at_path(const std::string& path) const;   // throws ConfigError when absent

    // toml++ value<T>() conversions: integers convert to floating point, not the reverse; no string coercion
    double get_float(const std::string& path) const;
    int64_t get_int(const std::string& path) const;
    uint32_t get_u32(const std::string& path) const;
    bool get_bool(const std::string& path) const;
    std::string get_string(const std::string& path) const;
    std::vector<double> get_float_array(const std::string& path) const;
    std::vector<int64_t> get_int_array(const std::string& path) const;

private:
    std::map<std::string, TomlValue> values;   // "table.sub.key" -> value
};

// The values the reference takes from its file, plus the [render] table this host adds (config/config.toml)
struct Config {
    // [render]
    uint32_t width = 800, height = 600;
    std::string scene = "cornell";
    bool nee = true;
    double cameraPos[3] = {0.0, 1.0, 3.9};
    double cameraLookAt[3] = {0.0, 1.0, 0.0};
    double fovYDegrees = 40.0;
    // [camera.dof], [sampling]
    float focusDist = 0, defocusMultiplier = 0;   // defocusMultiplier as written in the file (divided by 100 later)
    uint32_t samplesPerPixel = 0, maxBounces = 0;
    float directClamp = 0, indirectClamp = 0;
    // [saving]
    std::vector<int> saveOnSamples;
    std::vector<double> saveOnTimes;
    // [postprocessing.*]
    RB200BloomPushConsts bloom{};
    RB200TonemappingPushConsts tonemap{};

    // every key of the reference's tables is required; [render] keys are optional
    static Config from_toml(const TomlDoc& doc);
};

}  // namespace rbhost
