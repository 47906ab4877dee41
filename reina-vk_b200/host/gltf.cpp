// See gltf.h. Stage by stage the counterpart of reina-vk_b200/gltf.py (same rules, same arithmetic: fp32 where the
// reference's importer works in fp32, fp64 for node transforms with a fixed summation order).
#include "gltf.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>

#include "texture.h"

namespace rbhost {

namespace {

// ---------------------------------------------------------------------------------------------------------------
// a small JSON reader (RFC 8259): numbers as double (strtod), strings with \uXXXX escapes as UTF-8
// ---------------------------------------------------------------------------------------------------------------
struct Json {
    enum Type { Null, Bool, Number, String, Array, Object } type = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;

    const Json* find(const char* key) const {
        if (type != Object) return nullptr;
        for (const auto& kv : obj) if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool has(const char* key) const { return find(key) != nullptr; }
    const Json& at(const char* key) const {
        const Json* j = find(key);
        if (!j) throw std::runtime_error(std::string("Failed to parse glTF: missing \"") + key + "\"");
        return *j;
    }
    const Json& at(size_t i) const {
        if (type != Array || i >= arr.size()) throw std::runtime_error("Failed to parse glTF: index out of range");
        return arr[i];
    }
    size_t size() const { return type == Array ? arr.size() : 0; }
    double number() const {
        if (type != Number) throw std::runtime_error("Failed to parse glTF: number expected");
        return num;
    }
    long integer() const { return long(number()); }
    const std::string& string() const {
        if (type != String) throw std::runtime_error("Failed to parse glTF: string expected");
        return str;
    }
    double number_or(const char* key, double dflt) const { const Json* j = find(key); return j ? j->number() : dflt; }
    long integer_or(const char* key, long dflt) const { const Json* j = find(key); return j ? j->integer() : dflt; }
    bool bool_or(const char* key, bool dflt) const { const Json* j = find(key); return j && j->type == Bool ? j->b : dflt; }
};

class JsonParser {
public:
    JsonParser(const char* begin, const char* end) : p(begin), e(end) {}
    Json parse() {
        Json v = value(0);
        ws();
        if (p != e) fail("trailing characters");
        return v;
    }

private:
    const char* p;
    const char* e;
    [[noreturn]] void fail(const char* what) { throw std::runtime_error(std::string("Failed to parse glTF: JSON: ") + what); }
    void ws() { while (p < e && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++; }
    bool lit(const char* s) {
        size_t n = std::strlen(s);
        if (size_t(e - p) >= n && std::memcmp(p, s, n) == 0) { p += n; return true; }
        return false;
    }
    static void utf8(std::string& out, uint32_t c) {
        if (c < 0x80) out += char(c);
        else if (c < 0x800) { out += char(0xC0 | (c >> 6)); out += char(0x80 | (c & 0x3F)); }
        else if (c < 0x10000) { out += char(0xE0 | (c >> 12)); out += char(0x80 | ((c >> 6) & 0x3F)); out += char(0x80 | (c & 0x3F)); }
        else { out += char(0xF0 | (c >> 18)); out += char(0x80 | ((c >> 12) & 0x3F)); out += char(0x80 | ((c >> 6) & 0x3F)); out += char(0x80 | (c & 0x3F)); }
    }
    uint32_t hex4() {
        if (e - p < 4) fail("bad \\u escape");
        uint32_t v = 0;
        for (int i = 0; i < 4; i++, p++) {
            char c = *p;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= uint32_t(c - '0');
            else if (c >= 'a' && c <= 'f') v |= uint32_t(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= uint32_t(c - 'A' + 10);
            else fail("bad \\u escape");
        }
        return v;
    }
    std::string string() {
        std::string out;
        p++;   // opening quote
        for (;;) {
            if (p >= e) fail("unterminated string");
            char c = *p++;
            if (c == '"') break;
            if (c != '\\') { out += c; continue; }
            if (p >= e) fail("unterminated string");
            char esc = *p++;
            switch (esc) {
                case '"': out += '"'; break;
                case '\\': out += '\\'; break;
                case '/': out += '/'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'n': out += '\n'; break;
                case 'r': out += '\r'; break;
                case 't': out += '\t'; break;
                case 'u': {
                    uint32_t c1 = hex4();
                    if (c1 >= 0xD800 && c1 < 0xDC00 && e - p >= 6 && p[0] == '\\' && p[1] == 'u') {
                        p += 2;
                        uint32_t c2 = hex4();
                        c1 = 0x10000 + ((c1 - 0xD800) << 10) + (c2 - 0xDC00);
                    }
                    utf8(out, c1);
                    break;
                }
                default: fail("bad escape");
            }
        }
        return out;
    }
    Json value(int depth) {
        if (depth > 200) fail("nesting too deep");
        ws();
        if (p >= e) fail("unexpected end");
        Json v;
        if (*p == '{') {
            v.type = Json::Object;
            p++; ws();
            if (p < e && *p == '}') { p++; return v; }
            for (;;) {
                ws();
                if (p >= e || *p != '"') fail("object key expected");
                std::string k = string();
                ws();
                if (p >= e || *p != ':') fail("':' expected");
                p++;
                v.obj.emplace_back(std::move(k), value(depth + 1));
                ws();
                if (p < e && *p == ',') { p++; continue; }
                if (p < e && *p == '}') { p++; break; }
                fail("',' or '}' expected");
            }
        } else if (*p == '[') {
            v.type = Json::Array;
            p++; ws();
            if (p < e && *p == ']') { p++; return v; }
            for (;;) {
                v.arr.push_back(value(depth + 1));
                ws();
                if (p < e && *p == ',') { p++; continue; }
                if (p < e && *p == ']') { p++; break; }
                fail("',' or ']' expected");
            }
        } else if (*p == '"') {
            v.type = Json::String;
            v.str = string();
        } else if (lit("true")) { v.type = Json::Bool; v.b = true; }
        else if (lit("false")) { v.type = Json::Bool; v.b = false; }
        else if (lit("null")) { v.type = Json::Null; }
        else {
            const char* s = p;
            while (p < e && (std::strchr("+-.eE", *p) || (*p >= '0' && *p <= '9'))) p++;
            if (p == s) fail("unexpected character");
            std::string tok(s, p);
            char* endp = nullptr;
            v.type = Json::Number;
            v.num = std::strtod(tok.c_str(), &endp);
            if (endp == tok.c_str() || *endp) fail("bad number");
        }
        return v;
    }
};

std::vector<uint8_t> read_file(const std::string& path, const char* errPrefix) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error(errPrefix + path);
    std::vector<uint8_t> data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    return data;
}

bool file_exists(const std::string& path) { return bool(std::ifstream(path, std::ios::binary)); }

std::vector<uint8_t> decode_data_uri(const std::string& uri) {
    const size_t comma = uri.find(',');
    const std::string head = uri.substr(0, comma == std::string::npos ? uri.size() : comma);
    const char* suffix = ";base64";
    if (comma == std::string::npos || head.size() < 7 || head.compare(head.size() - 7, 7, suffix) != 0)
        throw std::runtime_error("Failed to parse glTF: only base64 data URIs are supported");
    std::vector<uint8_t> out;
    uint32_t acc = 0;
    int bits = 0;
    for (size_t i = comma + 1; i < uri.size(); i++) {
        const char c = uri[i];
        int v;
        if (c >= 'A' && c <= 'Z') v = c - 'A';
        else if (c >= 'a' && c <= 'z') v = c - 'a' + 26;
        else if (c >= '0' && c <= '9') v = c - '0' + 52;
        else if (c == '+' || c == '-') v = 62;
        else if (c == '/' || c == '_') v = 63;
        else continue;   // '=', whitespace
        acc = (acc << 6) | uint32_t(v);
        bits += 6;
        if (bits >= 8) { bits -= 8; out.push_back(uint8_t((acc >> bits) & 0xFF)); }
    }
    return out;
}

std::string dir_of(const std::string& path) {
    const size_t s = path.find_last_of('/');
    return s == std::string::npos ? std::string(".") : path.substr(0, s);
}

struct Asset {
    Json doc;
    std::vector<std::vector<uint8_t>> buffers;
    std::string baseDir;
};

uint32_t rd32(const uint8_t* p) { return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24); }

// loadGltf, gltfloader.cpp:72-120
Asset load_gltf(const std::string& filepath) {
    if (!file_exists(filepath)) throw std::runtime_error("Failed to find glTF file: " + filepath);
    const std::vector<uint8_t> raw = read_file(filepath, "Failed to open glTF file: ");
    Asset a;
    a.baseDir = dir_of(filepath);
    std::vector<uint8_t> binChunk;
    bool haveBin = false;
    if (raw.size() >= 4 && std::memcmp(raw.data(), "glTF", 4) == 0) {
        if (raw.size() < 20) throw std::runtime_error("Failed to parse glTF: truncated GLB header");
        const uint32_t version = rd32(&raw[4]), length = rd32(&raw[8]);
        if (version != 2) throw std::runtime_error("Failed to parse glTF: unsupported GLB version " + std::to_string(version));
        size_t off = 12;
        bool haveDoc = false;
        const size_t limit = std::min<size_t>(length, raw.size());
        while (off + 8 <= limit) {
            const uint32_t clen = rd32(&raw[off]), ctype = rd32(&raw[off + 4]);
            if (off + 8 + size_t(clen) > raw.size()) throw std::runtime_error("Failed to parse glTF: truncated GLB chunk");
            const uint8_t* body = &raw[off + 8];
            if (ctype == 0x4E4F534Au && !haveDoc) {
                a.doc = JsonParser(reinterpret_cast<const char*>(body), reinterpret_cast<const char*>(body) + clen).parse();
                haveDoc = true;
            } else if (ctype == 0x004E4942u && !haveBin) {
                binChunk.assign(body, body + clen);
                haveBin = true;
            }
            off += 8 + ((size_t(clen) + 3) & ~size_t(3));
        }
        if (!haveDoc) throw std::runtime_error("Failed to parse glTF: GLB without a JSON chunk");
    } else {
        a.doc = JsonParser(reinterpret_cast<const char*>(raw.data()), reinterpret_cast<const char*>(raw.data()) + raw.size()).parse();
    }
    if (a.doc.type != Json::Object) throw std::runtime_error("Failed to parse glTF: JSON: top level is not an object");
    // the parser of the reference is created with exactly these extensions (gltfloader.cpp:87-93) and rejects a file
    // that REQUIRES any other one (compressed geometry, basis textures, ...)
    if (const Json* req = a.doc.find("extensionsRequired")) {
        static const char* kEnabled[] = {"KHR_mesh_quantization", "KHR_texture_transform", "KHR_materials_variants",
                                         "KHR_materials_transmission", "KHR_materials_clearcoat", "KHR_materials_emissive_strength"};
        for (size_t i = 0; i < req->size(); i++) {
            const std::string& e = req->at(i).string();
            bool ok = false;
            for (const char* k : kEnabled) ok = ok || e == k;
            if (!ok) throw std::runtime_error("Failed to parse glTF: required extension " + e + " is not enabled");
        }
    }
    if (const Json* bufs = a.doc.find("buffers")) {
        for (size_t i = 0; i < bufs->size(); i++) {
            const Json& b = bufs->at(i);
            std::vector<uint8_t> data;
            if (const Json* uri = b.find("uri")) {
                const std::string& u = uri->string();
                if (u.compare(0, 5, "data:") == 0) data = decode_data_uri(u);
                else {
                    const std::string p = a.baseDir + "/" + u;
                    if (!file_exists(p)) throw std::runtime_error("Failed to parse glTF: missing external buffer " + u);
                    data = read_file(p, "Failed to parse glTF: missing external buffer ");
                }
            } else {
                if (i != 0 || !haveBin) throw std::runtime_error("Failed to parse glTF: buffer " + std::to_string(i) + " has no uri and no GLB BIN chunk");
                data = binChunk;
            }
            if (data.size() < size_t(b.integer_or("byteLength", 0)))
                throw std::runtime_error("Failed to parse glTF: buffer " + std::to_string(i) + " is shorter than its byteLength");
            a.buffers.push_back(std::move(data));
        }
    }
    return a;
}

int component_size(long ct) {
    switch (ct) {
        case 5120: case 5121: return 1;
        case 5122: case 5123: return 2;
        case 5125: case 5126: return 4;
        default: throw std::runtime_error("Failed to parse glTF: unknown componentType " + std::to_string(ct));
    }
}
int type_components(const std::string& t) {
    if (t == "SCALAR") return 1;
    if (t == "VEC2") return 2;
    if (t == "VEC3") return 3;
    if (t == "VEC4" || t == "MAT2") return 4;
    if (t == "MAT3") return 9;
    if (t == "MAT4") return 16;
    throw std::runtime_error("Failed to parse glTF: unknown accessor type " + t);
}

// Counts, offsets and strides come from untrusted JSON: every one is range-checked before it is used in address
// arithmetic (a negative or huge value would wrap the size_t expressions below), and the bounds tests are written so
// that they cannot overflow. 2^31 elements / bytes is far beyond anything a scene for this renderer holds.
size_t checked_size(long v, const char* what, long index) {
    if (v < 0 || v > 0x7FFFFFFFL)
        throw std::runtime_error(std::string("Failed to parse glTF: ") + what + " of accessor " + std::to_string(index) + " out of range");
    return size_t(v);
}
// true iff `count` elements of `elem` bytes, `stride` apart, starting at `start`, lie inside [0, limit)
bool range_fits(size_t start, size_t count, size_t stride, size_t elem, size_t limit) {
    if (count == 0) return start <= limit;
    if (start > limit || elem > limit - start) return false;
    if (stride == 0) return count == 1;
    return (count - 1) <= (limit - start - elem) / stride;
}

struct AccessorView {
    const uint8_t* base = nullptr;   // nullptr: accessor without a buffer view (all zeros)
    size_t stride = 0, count = 0;
    int ncomp = 0, csize = 0;
    long ctype = 0;
    bool normalized = false;
    std::shared_ptr<std::vector<uint8_t>> owned;   // sparse accessors: a packed copy with the substitutions applied
};

AccessorView open_accessor(const Asset& a, long index) {
    if (index < 0) throw std::runtime_error("Failed to parse glTF: negative accessor index");
    const Json& acc = a.doc.at("accessors").at(size_t(index));
    AccessorView v;
    v.ctype = acc.at("componentType").integer();
    v.csize = component_size(v.ctype);
    v.ncomp = type_components(acc.at("type").string());
    v.count = checked_size(acc.at("count").integer(), "count", index);
    v.normalized = acc.bool_or("normalized", false);
    if (const Json* bvi = acc.find("bufferView")) {
        const Json& bv = a.doc.at("bufferViews").at(checked_size(bvi->integer(), "bufferView", index));
        const size_t bi = checked_size(bv.at("buffer").integer(), "buffer", index);
        if (bi >= a.buffers.size()) throw std::runtime_error("Failed to parse glTF: buffer index out of range");
        const std::vector<uint8_t>& buf = a.buffers[bi];
        const size_t viewOff = checked_size(bv.integer_or("byteOffset", 0), "bufferView.byteOffset", index);
        const size_t viewLen = checked_size(bv.at("byteLength").integer(), "bufferView.byteLength", index);
        const size_t start = viewOff + checked_size(acc.integer_or("byteOffset", 0), "byteOffset", index);
        const size_t elem = size_t(v.csize) * size_t(v.ncomp);
        const size_t st = checked_size(bv.integer_or("byteStride", 0), "byteStride", index);
        v.stride = st ? st : elem;
        const size_t limit = std::min(buf.size(), viewOff + viewLen);       // both terms < 2^32: no overflow
        if (!range_fits(start, v.count, v.stride, elem, limit))
            throw std::runtime_error("Failed to parse glTF: accessor " + std::to_string(index) + " reads past its buffer view");
        v.base = buf.data() + start;
    }
    if (const Json* sp = acc.find("sparse")) {
        // glTF 2.0, 3.6.2.3: the accessor's elements (zeros without a buffer view) with `count` of them replaced;
        // indices are strictly increasing element numbers, values are tightly packed elements of the accessor's type
        const size_t elem = size_t(v.csize) * size_t(v.ncomp);
        auto packed = std::make_shared<std::vector<uint8_t>>(v.count * elem, uint8_t(0));
        if (v.base) for (size_t i = 0; i < v.count; i++) std::memcpy(packed->data() + i * elem, v.base + i * v.stride, elem);
        const size_t n = checked_size(sp->at("count").integer(), "sparse.count", index);
        const Json& si = sp->at("indices");
        const Json& sv = sp->at("values");
        const long ict = si.at("componentType").integer();
        if (ict != 5121 && ict != 5123 && ict != 5125) throw std::runtime_error("Failed to parse glTF: bad sparse index type");
        const size_t isize = size_t(component_size(ict));
        auto view = [&](const Json& ref, size_t bytes) -> const uint8_t* {
            const Json& bv = a.doc.at("bufferViews").at(checked_size(ref.at("bufferView").integer(), "sparse bufferView", index));
            const size_t bi = checked_size(bv.at("buffer").integer(), "buffer", index);
            if (bi >= a.buffers.size()) throw std::runtime_error("Failed to parse glTF: buffer index out of range");
            const size_t viewOff = checked_size(bv.integer_or("byteOffset", 0), "bufferView.byteOffset", index);
            const size_t start = viewOff + checked_size(ref.integer_or("byteOffset", 0), "sparse byteOffset", index);
            const size_t limit = std::min(a.buffers[bi].size(), viewOff + checked_size(bv.at("byteLength").integer(), "bufferView.byteLength", index));
            if (!range_fits(start, 1, 0, bytes, limit))
                throw std::runtime_error("Failed to parse glTF: accessor " + std::to_string(index) + " reads past its buffer view");
            return a.buffers[bi].data() + start;
        };
        if (n > v.count) throw std::runtime_error("Failed to parse glTF: bad sparse accessor");
        const uint8_t* ip = n ? view(si, n * isize) : nullptr;
        const uint8_t* vp = n ? view(sv, n * elem) : nullptr;
        size_t previous = 0;
        for (size_t k = 0; k < n; k++) {
            size_t at = 0;
            for (size_t b = 0; b < isize; b++) at |= size_t(ip[k * isize + b]) << (8 * b);
            if (at >= v.count || (k && at <= previous)) throw std::runtime_error("Failed to parse glTF: bad sparse accessor");
            previous = at;
            std::memcpy(packed->data() + at * elem, vp + k * elem, elem);
        }
        v.owned = packed;
        v.base = packed->data();
        v.stride = elem;
    }
    return v;
}

// one component as fp32: floats as stored, normalized integers c / max (signed: max(c / max, -1)), others converted
float component_f32(const AccessorView& v, size_t i, int c) {
    if (!v.base) return 0.0f;
    const uint8_t* p = v.base + i * v.stride + size_t(c) * size_t(v.csize);
    float f, mx;
    switch (v.ctype) {
        case 5126: std::memcpy(&f, p, 4); return f;
        case 5120: f = float(int8_t(p[0])); mx = 127.0f; break;
        case 5121: f = float(p[0]); mx = 255.0f; break;
        case 5122: { int16_t s; std::memcpy(&s, p, 2); f = float(s); mx = 32767.0f; break; }
        case 5123: { uint16_t s; std::memcpy(&s, p, 2); f = float(s); mx = 65535.0f; break; }
        default: { uint32_t s; std::memcpy(&s, p, 4); f = float(s); mx = 1.0f; break; }
    }
    if (v.normalized) {
        f = f / mx;
        if (v.ctype == 5120 || v.ctype == 5122) f = std::fmax(f, -1.0f);
    }
    return f;
}
uint32_t component_u32(const AccessorView& v, size_t i) {
    if (!v.base) return 0;
    const uint8_t* p = v.base + i * v.stride;
    switch (v.ctype) {
        case 5121: return p[0];
        case 5123: { uint16_t s; std::memcpy(&s, p, 2); return s; }
        case 5125: { uint32_t s; std::memcpy(&s, p, 4); return s; }
        case 5120: return uint32_t(int8_t(p[0]));
        case 5122: { int16_t s; std::memcpy(&s, p, 2); return uint32_t(s); }
        default: { float f; std::memcpy(&f, p, 4); return uint32_t(f); }
    }
}

std::vector<float> read_floats(const Asset& a, long index, int want) {
    const AccessorView v = open_accessor(a, index);
    if (v.ncomp < want) throw std::runtime_error("Failed to parse glTF: accessor " + std::to_string(index) + " has too few components");
    std::vector<float> out(v.count * size_t(want));
    for (size_t i = 0; i < v.count; i++)
        for (int c = 0; c < want; c++) out[i * size_t(want) + size_t(c)] = component_f32(v, i, c);
    return out;
}

// gltfloader.h:19-25
struct Primitive {
    std::vector<float> position, normal, tangent, bitangent, uv;   // 3,3,3,3,2 floats per vertex
    std::vector<uint32_t> indices;
    long materialIdx = -1;
    bool hasTangents = false;

    // What genTangSpaceDefault (MikkTSpace@3e895b4) writes through the reference's callbacks (gltfloader.cpp:16-67,
    // 207-222): the call sees faces f = vertices 3f .. 3f+2 whose UVs are all (0, 0) (TEXCOORD_0 is read afterwards), so
    // every triangle keeps InitTriInfo's GROUP_WITH_ANY mark, Build4RuleGroups opens no group and every corner keeps the
    // tangent space it was initialised with — (1, 0, 0) / (0, 1, 0); vertices beyond the last complete triple keep the zeros
    // of their value-initialisation. The derivation is written out in reina-vk_b200/gltf.py (mikktspace_as_called).
    void mikktspace_as_called() {
        const size_t n = position.size() / 3, full = 3 * (n / 3);
        tangent.assign(n * 3, 0.0f);
        bitangent.assign(n * 3, 0.0f);
        for (size_t v = 0; v < full; v++) { tangent[3 * v] = 1.0f; bitangent[3 * v + 1] = 1.0f; }
        hasTangents = true;
    }

    // toModelData, gltfloader.cpp:257-285
    ModelData toModelData() const {
        if (!hasTangents) return make_model(position, uv, normal, indices);     // GltfTangents::Uv: frames from the UV derivatives
        const size_t n = position.size() / 3;
        ModelData md;
        md.vertices.resize(n * 4);
        md.tbns.resize(n * 9);
        for (size_t v = 0; v < n; v++) {
            md.vertices[4 * v] = position[3 * v]; md.vertices[4 * v + 1] = position[3 * v + 1];
            md.vertices[4 * v + 2] = position[3 * v + 2]; md.vertices[4 * v + 3] = 1.0f;
            for (int k = 0; k < 3; k++) {
                md.tbns[9 * v + size_t(k)] = tangent[3 * v + size_t(k)];
                md.tbns[9 * v + 3 + size_t(k)] = bitangent[3 * v + size_t(k)];
                md.tbns[9 * v + 6 + size_t(k)] = normal[3 * v + size_t(k)];
            }
        }
        md.indices = indices; md.tbnsIndices = indices; md.texIndices = indices;
        md.texCoords = uv;
        return md;
    }
};

using M4 = std::array<std::array<double, 4>, 4>;   // [column][row]

M4 mat_mul(const M4& a, const M4& b) {
    M4 out{};
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) {
            double s = 0.0;
            for (int k = 0; k < 4; k++) s = s + a[size_t(k)][size_t(r)] * b[size_t(c)][size_t(k)];
            out[size_t(c)][size_t(r)] = s;
        }
    return out;
}

M4 local_matrix(const Json& node) {
    M4 m{};
    if (const Json* mj = node.find("matrix")) {
        if (mj->size() != 16) throw std::runtime_error("Failed to parse glTF: node matrix must have 16 elements");
        for (int c = 0; c < 4; c++)
            for (int r = 0; r < 4; r++) m[size_t(c)][size_t(r)] = mj->at(size_t(4 * c + r)).number();
        return m;
    }
    double t[3] = {0, 0, 0}, q[4] = {0, 0, 0, 1}, s[3] = {1, 1, 1};
    if (const Json* j = node.find("translation")) for (int k = 0; k < 3; k++) t[k] = j->at(size_t(k)).number();
    if (const Json* j = node.find("rotation")) for (int k = 0; k < 4; k++) q[k] = j->at(size_t(k)).number();
    if (const Json* j = node.find("scale")) for (int k = 0; k < 3; k++) s[k] = j->at(size_t(k)).number();
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double rows[3][3] = {
        {1.0 - 2.0 * (y * y + z * z), 2.0 * (x * y - z * w), 2.0 * (x * z + y * w)},
        {2.0 * (x * y + z * w), 1.0 - 2.0 * (x * x + z * z), 2.0 * (y * z - x * w)},
        {2.0 * (x * z - y * w), 2.0 * (y * z + x * w), 1.0 - 2.0 * (x * x + y * y)}};
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) m[size_t(c)][size_t(r)] = rows[r][c] * s[c];
    m[3][0] = t[0]; m[3][1] = t[1]; m[3][2] = t[2]; m[3][3] = 1.0;
    return m;
}

// fastgltf::iterateSceneNodes: depth-first, parents first, world = parent * local
template <class F>
void walk(const Json& doc, F&& visit) {
    const Json* scenes = doc.find("scenes");
    if (!scenes || scenes->size() == 0) throw std::runtime_error("No scenes supplied in gLTF file");
    const Json& scene = scenes->at(size_t(doc.integer_or("scene", 0)));
    const Json* nodes = doc.find("nodes");
    struct Rec {
        const Json* nodes; F& visit;
        void run(long i, const M4& parent, int depth) {
            if (depth > 256) throw std::runtime_error("Failed to parse glTF: node hierarchy too deep (cycle?)");
            if (!nodes) throw std::runtime_error("Failed to parse glTF: scene refers to missing nodes");
            const Json& node = nodes->at(size_t(i));
            const M4 world = mat_mul(parent, local_matrix(node));
            visit(node, world);
            if (const Json* ch = node.find("children"))
                for (size_t k = 0; k < ch->size(); k++) run(ch->at(k).integer(), world, depth + 1);
        }
    } rec{nodes, visit};
    M4 ident{};
    for (int i = 0; i < 4; i++) ident[size_t(i)][size_t(i)] = 1.0;
    if (const Json* roots = scene.find("nodes"))
        for (size_t k = 0; k < roots->size(); k++) rec.run(roots->at(k).integer(), ident, 0);
}

// loadPrimitives, gltfloader.cpp:122-255
std::map<long, std::vector<Primitive>> load_primitives(const Asset& a, GltfTangents tangents) {
    std::set<long> used;
    walk(a.doc, [&](const Json& node, const M4&) { if (const Json* m = node.find("mesh")) used.insert(m->integer()); });
    std::map<long, std::vector<Primitive>> out;
    for (long mi : used) {
        std::vector<Primitive> prims;
        const Json& mesh = a.doc.at("meshes").at(size_t(mi));
        const Json* pj = mesh.find("primitives");
        for (size_t pi = 0; pj && pi < pj->size(); pi++) {
            const Json& prim = pj->at(pi);
            if (prim.integer_or("mode", 4) != 4) throw std::runtime_error("Failed to parse glTF: only TRIANGLES primitives are supported");
            static const Json kEmpty;
            const Json* attrsP = prim.find("attributes");
            const Json& attrs = attrsP ? *attrsP : kEmpty;
            Primitive m;
            m.materialIdx = prim.integer_or("material", -1);
            if (!attrs.has("POSITION")) throw std::runtime_error("Failed to parse glTF: primitive without POSITION");
            m.position = read_floats(a, attrs.at("POSITION").integer(), 3);
            const size_t n = m.position.size() / 3;
            if (!attrs.has("NORMAL")) throw std::runtime_error("Meshes without vertex normals are not supported");
            m.normal = read_floats(a, attrs.at("NORMAL").integer(), 3);
            if (m.normal.size() != n * 3) throw std::runtime_error("Failed to parse glTF: NORMAL count differs from POSITION");
            if (attrs.has("TEXCOORD_0")) {
                m.uv = read_floats(a, attrs.at("TEXCOORD_0").integer(), 2);
                if (m.uv.size() != n * 2) throw std::runtime_error("Failed to parse glTF: TEXCOORD_0 count differs from POSITION");
            } else {
                std::fprintf(stderr, "Warning: falling back to UV coords (0, 0) since none were found\n");
                m.uv.assign(n * 2, 0.0f);
            }
            if (const Json* ij = prim.find("indices")) {
                const AccessorView v = open_accessor(a, ij->integer());
                m.indices.resize(v.count);
                for (size_t i = 0; i < v.count; i++) m.indices[i] = component_u32(v, i);
            } else {
                m.indices.resize(n);                               // fastgltf::Options::GenerateMeshIndices
                for (size_t i = 0; i < n; i++) m.indices[i] = uint32_t(i);
            }
            bool bad = m.indices.size() % 3 != 0;
            for (uint32_t ix : m.indices) bad = bad || ix >= n;
            if (bad) throw std::runtime_error("Failed to parse glTF: bad index data");
            if (attrs.has("TANGENT")) {
                const AccessorView v = open_accessor(a, attrs.at("TANGENT").integer());
                if (v.count != n || v.ncomp < 3) throw std::runtime_error("Failed to parse glTF: TANGENT count differs from POSITION");
                m.hasTangents = true;
                m.tangent.resize(n * 3);
                m.bitangent.resize(n * 3);
                for (size_t i = 0; i < n; i++) {
                    const float tx = component_f32(v, i, 0), ty = component_f32(v, i, 1), tz = component_f32(v, i, 2);
                    const float w = v.ncomp >= 4 ? component_f32(v, i, 3) : 1.0f;     // Vec3: assume w = +1
                    const float nx = m.normal[3 * i], ny = m.normal[3 * i + 1], nz = m.normal[3 * i + 2];
                    m.tangent[3 * i] = tx; m.tangent[3 * i + 1] = ty; m.tangent[3 * i + 2] = tz;
                    // bitangent = cross(normal, tangent) * w
                    const float cx = ny * tz - nz * ty, cy = nz * tx - nx * tz, cz = nx * ty - ny * tx;
                    m.bitangent[3 * i] = cx * w; m.bitangent[3 * i + 1] = cy * w; m.bitangent[3 * i + 2] = cz * w;
                }
            } else if (tangents == GltfTangents::Reference) {
                m.mikktspace_as_called();
            }
            prims.push_back(std::move(m));
        }
        out[mi] = std::move(prims);
    }
    return out;
}

// addTexturesToScene, gltfloader.cpp:287-343
std::map<long, int> add_textures(const Asset& a, Scene& scene) {
    std::map<long, int> ids;
    const Json* images = a.doc.find("images");
    for (size_t i = 0; images && i < images->size(); i++) {
        const Json& img = images->at(i);
        const std::string name = "image " + std::to_string(i);
        auto decode = [&](const uint8_t* data, size_t size, bool flip, const std::string& nm) {
            return decode_image_rgba8(data, size, flip, nm);      // PNG or baseline JPEG, by signature
        };
        if (const Json* uri = img.find("uri")) {
            const std::string& u = uri->string();
            if (u.compare(0, 5, "data:") == 0) {
                const std::vector<uint8_t> bytes = decode_data_uri(u);
                ids[long(i)] = int(scene.defineTexture(decode(bytes.data(), bytes.size(), false, name)));
            } else {
                const std::string p = a.baseDir + "/" + u;
                if (!file_exists(p)) throw std::runtime_error("Could not load image at path: " + p);
                const std::vector<uint8_t> bytes = read_file(p, "Could not load image at path: ");
                ids[long(i)] = int(scene.defineTexture(decode(bytes.data(), bytes.size(), true, p)));   // file: flipped
            }
        } else if (const Json* bvi = img.find("bufferView")) {
            const Json& bv = a.doc.at("bufferViews").at(size_t(bvi->integer()));
            const size_t bi = size_t(bv.at("buffer").integer());
            if (bi >= a.buffers.size()) throw std::runtime_error("Failed to parse glTF: buffer index out of range");
            const size_t off = size_t(bv.integer_or("byteOffset", 0)), len = size_t(bv.at("byteLength").integer());
            if (off + len > a.buffers[bi].size()) throw std::runtime_error("Failed to parse glTF: image buffer view out of range");
            ids[long(i)] = int(scene.defineTexture(decode(a.buffers[bi].data() + off, len, false, name)));
        } else {
            throw std::runtime_error("Could not parse texture; internal gLTF data type not supported");
        }
    }
    return ids;
}

// materialsFromMeshTBNs, gltfloader.cpp:387-438
Material material_of(const Asset& a, const Primitive& p, const std::map<long, int>& texIds) {
    Material mat;                       // Material{3, -1, -1, -1, vec3(1), vec3(0), 0, 1.5, true, 0, false, 0, 0, 0, vec3(1), vec3(1), 0, 0, 0, 0}
    mat.materialIdx = 3;
    mat.albedo = {1.0f, 1.0f, 1.0f};
    mat.ior = 1.5f;
    mat.interpNormals = true;
    mat.sheenTint = {1.0f, 1.0f, 1.0f};
    mat.specularTint = {1.0f, 1.0f, 1.0f};
    if (p.materialIdx == -1) return mat;
    const Json& g = a.doc.at("materials").at(size_t(p.materialIdx));
    static const Json kEmpty;
    const Json* extP = g.find("extensions");
    const Json& ext = extP ? *extP : kEmpty;
    auto tex_id = [&](const Json& info, const char* what) -> int {
        const Json* textures = a.doc.find("textures");
        const Json* idx = info.find("index");
        if (textures && idx && size_t(idx->integer()) < textures->size()) {
            if (const Json* src = textures->at(size_t(idx->integer())).find("source")) {
                auto it = texIds.find(src->integer());
                if (it != texIds.end()) return it->second;
            }
        }
        std::fprintf(stderr, "Warning: %s ID not found\n", what);
        return -1;
    };
    float ef[3] = {0.0f, 0.0f, 0.0f};
    if (const Json* e = g.find("emissiveFactor")) for (int k = 0; k < 3; k++) ef[k] = float(e->at(size_t(k)).number());
    if ((ef[0] > 0 || ef[1] > 0 || ef[2] > 0) && !g.has("emissiveTexture")) {
        double strength = 1.0;
        if (const Json* es = ext.find("KHR_materials_emissive_strength")) strength = es->number_or("emissiveStrength", 1.0);
        const float s = float(strength);
        mat.emission = {ef[0] * s, ef[1] * s, ef[2] * s};
    }
    const Json* pbrP = g.find("pbrMetallicRoughness");
    const Json& pbr = pbrP ? *pbrP : kEmpty;
    mat.metallic = float(pbr.number_or("metallicFactor", 1.0));
    mat.roughness = std::fmax(std::fmin(float(pbr.number_or("roughnessFactor", 1.0)), 0.7f), 0.1f);
    if (const Json* bc = pbr.find("baseColorFactor")) for (int k = 0; k < 3; k++) mat.albedo[size_t(k)] = float(bc->at(size_t(k)).number());
    mat.cullBackface = !g.bool_or("doubleSided", false);
    if (const Json* io = ext.find("KHR_materials_ior")) mat.ior = float(io->number_or("ior", 1.5));
    if (const Json* tr = ext.find("KHR_materials_transmission")) {
        mat.specularTransmission = float(tr->number_or("transmissionFactor", 0.0));
        mat.cullBackface = false;       // thin transmissive materials are not supported
    }
    if (const Json* t = pbr.find("baseColorTexture")) mat.textureID = tex_id(*t, "Texture");
    if (const Json* t = g.find("normalTexture")) mat.normalMapID = tex_id(*t, "Normal Texture");
    return mat;
}

}  // namespace

Scene load_gltf_scene(const std::string& path, bool* hasEmitter, GltfTangents tangents) {
    const Asset asset = load_gltf(path);
    const std::map<long, std::vector<Primitive>> prims = load_primitives(asset, tangents);
    Scene scene;
    std::map<long, std::vector<uint32_t>> objectIds;                      // addMeshesToScene, :345-355
    for (const auto& kv : prims)
        for (const Primitive& p : kv.second) objectIds[kv.first].push_back(scene.defineObject(p.toModelData()));
    const std::map<long, int> texIds = add_textures(asset, scene);
    std::map<long, std::vector<Material>> materials;
    for (const auto& kv : prims)
        for (const Primitive& p : kv.second) materials[kv.first].push_back(material_of(asset, p, texIds));
    bool emits = false;
    walk(asset.doc, [&](const Json& node, const M4& world) {               // addInstancesToScene, :357-381
        const Json* m = node.find("mesh");
        if (!m) return;
        Mat4f t;
        for (int c = 0; c < 4; c++)
            for (int r = 0; r < 4; r++) t[size_t(4 * c + r)] = float(world[size_t(c)][size_t(r)]);
        const std::vector<uint32_t>& ids = objectIds.at(m->integer());
        const std::vector<Material>& mats = materials.at(m->integer());
        for (size_t k = 0; k < ids.size(); k++) {
            scene.addInstance(ids[k], t, mats[k]);
            const auto& e = mats[k].emission;
            emits = emits || (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) > 0.00001f * 0.00001f;
        }
    });
    if (hasEmitter) *hasEmitter = emits;
    return scene;
}

}  // namespace rbhost
