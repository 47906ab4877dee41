// See config.h.
#include "config.h"

#include <cctype>
#include <cerrno>
#include <cstdlib>
#include <fstream>
#include <sstream>

namespace rbhost {

namespace {

struct Cursor {
    const std::string& s;
    size_t p = 0;
    int line = 1;

    bool eof() const { return p >= s.size(); }
    char peek() const { return eof() ? '\0' : s[p]; }
    char next() {
        char c = s[p++];
        if (c == '\n') line++;
        return c;
    }
    [[noreturn]] void fail(const std::string& what) const {
        throw ConfigError("config: line " + std::to_string(line) + ": " + what);
    }
    // spaces and tabs; with `newlines` also line breaks and comments (inside arrays)
    void skip_blank(bool newlines) {
        while (!eof()) {
            char c = peek();
            if (c == ' ' || c == '\t' || c == '\r') { next(); continue; }
            if (newlines && c == '\n') { next(); continue; }
            if (newlines && c == '#') { while (!eof() && peek() != '\n') next(); continue; }
            break;
        }
    }
};

bool bare_key_char(char c) { return std::isalnum(static_cast<unsigned char>(c)) || c == '_' || c == '-'; }

std::string parse_key(Cursor& c) {
    std::string key;
    for (;;) {
        c.skip_blank(false);
        std::string part;
        if (c.peek() == '"') {
            c.next();
            while (!c.eof() && c.peek() != '"' && c.peek() != '\n') part += c.next();
            if (c.peek() != '"') c.fail("unterminated quoted key");
            c.next();
        } else {
            while (bare_key_char(c.peek())) part += c.next();
        }
        if (part.empty()) c.fail("expected a key");
        key += part;
        c.skip_blank(false);
        if (c.peek() == '.') { c.next(); key += '.'; continue; }
        return key;
    }
}

TomlValue parse_value(Cursor& c);

TomlValue parse_string(Cursor& c) {
    TomlValue v;
    v.kind = TomlValue::String;
    c.next();   // opening quote
    while (!c.eof() && c.peek() != '"') {
        char ch = c.next();
        if (ch == '\n') c.fail("unterminated string");
        if (ch == '\\') {
            if (c.eof()) c.fail("unterminated string");
            char e = c.next();
            switch (e) {
                case 'n': v.s += '\n'; break;
                case 't': v.s += '\t'; break;
                case '\\': v.s += '\\'; break;
                case '"': v.s += '"'; break;
                default: c.fail(std::string("unsupported escape \\") + e);
            }
        } else {
            v.s += ch;
        }
    }
    if (c.eof()) c.fail("unterminated string");
    c.next();
    return v;
}

TomlValue parse_scalar(Cursor& c) {
    std::string tok;
    while (!c.eof()) {
        char ch = c.peek();
        if (std::isalnum(static_cast<unsigned char>(ch)) || ch == '+' || ch == '-' || ch == '.' || ch == '_') tok += c.next();
        else break;
    }
    if (tok.empty()) c.fail("expected a value");
    TomlValue v;
    if (tok == "true" || tok == "false") {
        v.kind = TomlValue::Boolean;
        v.b = tok == "true";
        return v;
    }
    std::string digits;
    for (char ch : tok)
        if (ch != '_') digits += ch;
    if (tok == "inf" || tok == "+inf" || tok == "-inf" || tok == "nan" || tok == "+nan" || tok == "-nan") {
        v.kind = TomlValue::Float;
        v.f = std::strtod(digits.c_str(), nullptr);
        return v;
    }
    const bool isFloat = digits.find_first_of(".eE") != std::string::npos &&
                         digits.compare(0, 2, "0x") != 0;
    char* end = nullptr;
    errno = 0;
    if (isFloat) {
        v.kind = TomlValue::Float;
        v.f = std::strtod(digits.c_str(), &end);
    } else {
        v.kind = TomlValue::Integer;
        int base = 10;
        const char* start = digits.c_str();
        if (digits.compare(0, 2, "0x") == 0) { base = 16; start += 2; }
        else if (digits.compare(0, 2, "0o") == 0) { base = 8; start += 2; }
        else if (digits.compare(0, 2, "0b") == 0) { base = 2; start += 2; }
        v.i = std::strtoll(start, &end, base);
        v.f = static_cast<double>(v.i);
    }
    if (end == nullptr || *end != '\0' || errno == ERANGE) c.fail("cannot read value '" + tok + "'");
    return v;
}

TomlValue parse_array(Cursor& c) {
    TomlValue v;
    v.kind = TomlValue::Array;
    c.next();   // '['
    for (;;) {
        c.skip_blank(true);
        if (c.eof()) c.fail("unterminated array");
        if (c.peek() == ']') { c.next(); return v; }
        v.items.push_back(parse_value(c));
        c.skip_blank(true);
        if (c.peek() == ',') { c.next(); continue; }
        if (c.peek() == ']') { c.next(); return v; }
        c.fail("expected ',' or ']' in array");
    }
}

TomlValue parse_value(Cursor& c) {
    c.skip_blank(false);
    if (c.peek() == '"') return parse_string(c);
    if (c.peek() == '[') return parse_array(c);
    return parse_scalar(c);
}

void end_of_line(Cursor& c) {
    c.skip_blank(false);
    if (c.peek() == '#')
        while (!c.eof() && c.peek() != '\n') c.next();
    if (!c.eof() && c.peek() != '\n') c.fail("unexpected text after value");
    if (!c.eof()) c.next();
}

}  // namespace

TomlDoc TomlDoc::parse(const std::string& text) {
    TomlDoc doc;
    Cursor c{text};
    std::string table;
    while (!c.eof()) {
        c.skip_blank(true);
        if (c.eof()) break;
        if (c.peek() == '[') {
            c.next();
            if (c.peek() == '[') c.fail("arrays of tables are not supported");
            table = parse_key(c);
            if (c.peek() != ']') c.fail("expected ']'");
            c.next();
            end_of_line(c);
            continue;
        }
        std::string key = parse_key(c);
        if (c.peek() != '=') c.fail("expected '=' after key '" + key + "'");
        c.next();
        TomlValue v = parse_value(c);
        end_of_line(c);
        const std::string full = table.empty() ? key : table + "." + key;
        if (!doc.values.emplace(full, std::move(v)).second) c.fail("key '" + full + "' defined twice");
    }
    return doc;
}

TomlDoc TomlDoc::parse_file(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw ConfigError("config: cannot open " + path);
    std::ostringstream ss;
    ss << f.rdbuf();
    return parse(ss.str());
}

const TomlValue& TomlDoc::at_path(const std::string& path) const {
    auto it = values.find(path);
    if (it == values.end()) throw ConfigError("config: missing key '" + path + "'");
    return it->second;
}

static double as_float(const TomlValue& v, const std::string& path) {
    if (v.kind == TomlValue::Float) return v.f;
    if (v.kind == TomlValue::Integer) return static_cast<double>(v.i);
    throw ConfigError("config: key '" + path + "' is not a number");
}

static int64_t as_int(const TomlValue& v, const std::string& path) {
    if (v.kind == TomlValue::Integer) return v.i;
    throw ConfigError("config: key '" + path + "' is not an integer");
}

double TomlDoc::get_float(const std::string& path) const { return as_float(at_path(path), path); }
int64_t TomlDoc::get_int(const std::string& path) const { return as_int(at_path(path), path); }

uint32_t TomlDoc::get_u32(const std::string& path) const {
    int64_t v = get_int(path);
    if (v < 0 || v > 0xFFFFFFFFll) throw ConfigError("config: key '" + path + "' does not fit an unsigned 32-bit integer");
    return static_cast<uint32_t>(v);
}

bool TomlDoc::get_bool(const std::string& path) const {
    const TomlValue& v = at_path(path);
    if (v.kind != TomlValue::Boolean) throw ConfigError("config: key '" + path + "' is not a boolean");
    return v.b;
}

std::string TomlDoc::get_string(const std::string& path) const {
    const TomlValue& v = at_path(path);
    if (v.kind != TomlValue::String) throw ConfigError("config: key '" + path + "' is not a string");
    return v.s;
}

std::vector<double> TomlDoc::get_float_array(const std::string& path) const {
    const TomlValue& v = at_path(path);
    if (v.kind != TomlValue::Array) throw ConfigError("config: key '" + path + "' is not an array");
    std::vector<double> out;
    for (const TomlValue& x : v.items) out.push_back(as_float(x, path));
    return out;
}

std::vector<int64_t> TomlDoc::get_int_array(const std::string& path) const {
    const TomlValue& v = at_path(path);
    if (v.kind != TomlValue::Array) throw ConfigError("config: key '" + path + "' is not an array");
    std::vector<int64_t> out;
    for (const TomlValue& x : v.items) out.push_back(as_int(x, path));
    return out;
}

Config Config::from_toml(const TomlDoc& d) {
    Config c;
    // src/Reina.cpp:142-155
    c.focusDist = static_cast<float>(d.get_float("camera.dof.focus_dist"));
    c.defocusMultiplier = static_cast<float>(d.get_float("camera.dof.defocus_multiplier"));
    c.directClamp = static_cast<float>(d.get_float("sampling.direct_clamp"));
    c.indirectClamp = static_cast<float>(d.get_float("sampling.indirect_clamp"));
    c.samplesPerPixel = d.get_u32("sampling.samples_per_pixel");
    c.maxBounces = d.get_u32("sampling.max_bounces");
    // src/tools/SaveManager.cpp:6-25
    for (int64_t s : d.get_int_array("saving.save_on_samples")) c.saveOnSamples.push_back(static_cast<int>(s));
    c.saveOnTimes = d.get_float_array("saving.save_on_times");
    // src/Reina.cpp:246-262
    c.bloom.radius = static_cast<float>(d.get_float("postprocessing.bloom.radius"));
    c.bloom.threshold = static_cast<float>(d.get_float("postprocessing.bloom.threshold"));
    c.bloom.intensity = static_cast<float>(d.get_float("postprocessing.bloom.intensity"));
    c.tonemap.exposure = static_cast<float>(d.get_float("postprocessing.tonemap.exposure"));
    // [render]: this host's additions, all optional
    if (d.has("render.width")) c.width = d.get_u32("render.width");
    if (d.has("render.height")) c.height = d.get_u32("render.height");
    if (d.has("render.scene")) c.scene = d.get_string("render.scene");
    if (d.has("render.nee")) c.nee = d.get_bool("render.nee");
    if (d.has("render.fov_y_degrees")) c.fovYDegrees = d.get_float("render.fov_y_degrees");
    auto vec3 = [&](const char* key, double* out) {
        if (!d.has(key)) return;
        std::vector<double> v = d.get_float_array(key);
        if (v.size() != 3) throw ConfigError(std::string("config: key '") + key + "' must hold three numbers");
        for (int k = 0; k < 3; k++) out[k] = v[k];
    };
    vec3("render.camera_pos", c.cameraPos);
    vec3("render.camera_look_at", c.cameraLookAt);
    return c;
}

}  // namespace rbhost
