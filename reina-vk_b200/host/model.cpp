// See model.h. Arithmetic follows reina-vk_b200/meshes.py step by step (fp32 differences, fp64 accumulation, the
// same accumulation order) so that the C++ host and the Python host hand identical tables to rb200_scene_create.
#include "model.h"

#include <array>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <tuple>

namespace rbhost {

namespace {

struct D3 { double x = 0, y = 0, z = 0; };

inline void normalise_or_zero(D3& v) {
    double l = std::sqrt((v.x * v.x + v.y * v.y) + v.z * v.z);
    if (l > 0) { v.x /= l; v.y /= l; v.z /= l; } else { v = D3{}; }
}

// per-vertex tangent / bitangent from the UV derivatives of the adjacent triangles
void tangent_frames(const std::vector<float>& pos, const std::vector<float>& uv, const std::vector<uint32_t>& tris,
                    std::vector<D3>& T, std::vector<D3>& B) {
    const size_t n = pos.size() / 3, m = tris.size() / 3;
    T.assign(n, D3{});
    B.assign(n, D3{});
    if (uv.size() != n * 2) return;
    std::vector<D3> t(m), b(m);
    for (size_t i = 0; i < m; i++) {
        const uint32_t i0 = tris[3 * i], i1 = tris[3 * i + 1], i2 = tris[3 * i + 2];
        // differences are taken in fp32 (the inputs are fp32 arrays), everything after in fp64
        const double e1[3] = {double(pos[3 * i1] - pos[3 * i0]), double(pos[3 * i1 + 1] - pos[3 * i0 + 1]),
                              double(pos[3 * i1 + 2] - pos[3 * i0 + 2])};
        const double e2[3] = {double(pos[3 * i2] - pos[3 * i0]), double(pos[3 * i2 + 1] - pos[3 * i0 + 1]),
                              double(pos[3 * i2 + 2] - pos[3 * i0 + 2])};
        const double d1[2] = {double(uv[2 * i1] - uv[2 * i0]), double(uv[2 * i1 + 1] - uv[2 * i0 + 1])};
        const double d2[2] = {double(uv[2 * i2] - uv[2 * i0]), double(uv[2 * i2 + 1] - uv[2 * i0 + 1])};
        const double det = d1[0] * d2[1] - d2[0] * d1[1];
        const double r = std::fabs(det) > 1e-20 ? 1.0 / det : 0.0;
        t[i] = {(e1[0] * d2[1] - e2[0] * d1[1]) * r, (e1[1] * d2[1] - e2[1] * d1[1]) * r, (e1[2] * d2[1] - e2[2] * d1[1]) * r};
        b[i] = {(e2[0] * d1[0] - e1[0] * d2[0]) * r, (e2[1] * d1[0] - e1[1] * d2[0]) * r, (e2[2] * d1[0] - e1[2] * d2[0]) * r};
    }
    for (int k = 0; k < 3; k++)          // corner-major accumulation order
        for (size_t i = 0; i < m; i++) {
            const uint32_t v = tris[3 * i + k];
            T[v].x += t[i].x; T[v].y += t[i].y; T[v].z += t[i].z;
            B[v].x += b[i].x; B[v].y += b[i].y; B[v].z += b[i].z;
        }
    for (size_t v = 0; v < n; v++) { normalise_or_zero(T[v]); normalise_or_zero(B[v]); }
}

}  // namespace

ModelData make_model(const std::vector<float>& pos, const std::vector<float>& uv, const std::vector<float>& nrm,
                     const std::vector<uint32_t>& tris) {
    const size_t n = pos.size() / 3;
    ModelData md;
    md.vertices.resize(n * 4);
    for (size_t v = 0; v < n; v++) {
        md.vertices[4 * v] = pos[3 * v];
        md.vertices[4 * v + 1] = pos[3 * v + 1];
        md.vertices[4 * v + 2] = pos[3 * v + 2];
        md.vertices[4 * v + 3] = 1.0f;
    }
    std::vector<D3> T, B;
    tangent_frames(pos, uv, tris, T, B);
    md.tbns.resize(n * 9);
    for (size_t v = 0; v < n; v++) {
        float* o = &md.tbns[9 * v];
        o[0] = float(T[v].x); o[1] = float(T[v].y); o[2] = float(T[v].z);
        o[3] = float(B[v].x); o[4] = float(B[v].y); o[5] = float(B[v].z);
        o[6] = nrm[3 * v]; o[7] = nrm[3 * v + 1]; o[8] = nrm[3 * v + 2];
    }
    md.indices = tris;
    md.tbnsIndices = tris;
    md.texIndices = tris;
    md.texCoords = uv;
    return md;
}

std::vector<float> smooth_normals(const std::vector<float>& pos, const std::vector<uint32_t>& tris) {
    const size_t n = pos.size() / 3, m = tris.size() / 3;
    std::vector<D3> acc(n), fn(m);
    for (size_t i = 0; i < m; i++) {
        const float* p0 = &pos[3 * tris[3 * i]];
        const float* p1 = &pos[3 * tris[3 * i + 1]];
        const float* p2 = &pos[3 * tris[3 * i + 2]];
        const float a[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
        const float b[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
        const float m0 = a[1] * b[2], s0 = a[2] * b[1];
        const float m1 = a[2] * b[0], s1 = a[0] * b[2];
        const float m2 = a[0] * b[1], s2 = a[1] * b[0];
        fn[i] = {double(m0 - s0), double(m1 - s1), double(m2 - s2)};
    }
    for (int k = 0; k < 3; k++)
        for (size_t i = 0; i < m; i++) {
            D3& d = acc[tris[3 * i + k]];
            d.x += fn[i].x; d.y += fn[i].y; d.z += fn[i].z;
        }
    std::vector<float> out(n * 3);
    for (size_t v = 0; v < n; v++) {
        double l = std::sqrt((acc[v].x * acc[v].x + acc[v].y * acc[v].y) + acc[v].z * acc[v].z);
        if (!(l > 0)) l = 1.0;
        out[3 * v] = float(acc[v].x / l);
        out[3 * v + 1] = float(acc[v].y / l);
        out[3 * v + 2] = float(acc[v].z / l);
    }
    return out;
}

ModelData load_obj(const std::string& path, bool* hasTexCoords) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("Could not load model: cannot open " + path);
    std::vector<std::array<double, 3>> vs, vns;
    std::vector<std::array<double, 2>> vts;
    std::map<std::tuple<long, long, long>, uint32_t> keymap;
    std::vector<float> pos, uv, nrm;
    std::vector<uint32_t> tris;
    bool has_uv = true, has_n = true;
    std::string line;
    while (std::getline(f, line)) {
        const size_t hash = line.find('#');
        if (hash != std::string::npos) line.resize(hash);
        std::istringstream ss(line);
        std::string tag;
        if (!(ss >> tag)) continue;
        if (tag == "v" || tag == "vn") {
            std::string a, b, c;
            if (!(ss >> a >> b >> c)) throw std::runtime_error("Could not load model: malformed '" + tag + "' line in " + path);
            std::array<double, 3> p = {std::strtod(a.c_str(), nullptr), std::strtod(b.c_str(), nullptr),
                                       std::strtod(c.c_str(), nullptr)};
            (tag == "v" ? vs : vns).push_back(p);
        } else if (tag == "vt") {
            std::string a, b;
            if (!(ss >> a >> b)) throw std::runtime_error("Could not load model: malformed 'vt' line in " + path);
            vts.push_back({std::strtod(a.c_str(), nullptr), std::strtod(b.c_str(), nullptr)});
        } else if (tag == "f") {
            std::vector<uint32_t> corner;
            std::string tok;
            while (ss >> tok) {
                long idx[3] = {0, 0, 0};
                size_t start = 0;
                for (int k = 0; k < 3 && start <= tok.size(); k++) {
                    size_t slash = tok.find('/', start);
                    std::string part = tok.substr(start, slash == std::string::npos ? std::string::npos : slash - start);
                    if (!part.empty()) idx[k] = std::strtol(part.c_str(), nullptr, 10);
                    if (slash == std::string::npos) break;
                    start = slash + 1;
                }
                if (idx[0] == 0) throw std::runtime_error("Could not load model: bad face token '" + tok + "' in " + path);
                const long vi = idx[0] > 0 ? idx[0] - 1 : long(vs.size()) + idx[0];
                const long ti = idx[1] ? (idx[1] > 0 ? idx[1] - 1 : long(vts.size()) + idx[1]) : -1;
                const long ni = idx[2] ? (idx[2] > 0 ? idx[2] - 1 : long(vns.size()) + idx[2]) : -1;
                if (vi < 0 || vi >= long(vs.size()) || ti >= long(vts.size()) || ni >= long(vns.size()))
                    throw std::runtime_error("Could not load model: face index out of range in " + path);
                if (ti < 0) has_uv = false;
                if (ni < 0) has_n = false;
                const auto key = std::make_tuple(vi, ti, ni);
                auto it = keymap.find(key);
                if (it == keymap.end()) {
                    it = keymap.emplace(key, uint32_t(pos.size() / 3)).first;
                    pos.push_back(float(vs[vi][0])); pos.push_back(float(vs[vi][1])); pos.push_back(float(vs[vi][2]));
                    if (ti >= 0) { uv.push_back(float(vts[ti][0])); uv.push_back(float(1.0 - vts[ti][1])); }   // FlipUVs
                    else { uv.push_back(0.0f); uv.push_back(0.0f); }
                    if (ni >= 0) { nrm.push_back(float(vns[ni][0])); nrm.push_back(float(vns[ni][1])); nrm.push_back(float(vns[ni][2])); }
                    else { nrm.push_back(0.0f); nrm.push_back(0.0f); nrm.push_back(0.0f); }
                }
                corner.push_back(it->second);
            }
            for (size_t k = 1; k + 1 < corner.size(); k++) {
                tris.push_back(corner[0]); tris.push_back(corner[k]); tris.push_back(corner[k + 1]);
            }
        }
    }
    if (tris.empty()) throw std::runtime_error("Could not load model: no faces in " + path);
    if (!has_uv) uv.clear();
    if (!has_n) nrm = smooth_normals(pos, tris);
    if (hasTexCoords) *hasTexCoords = has_uv;
    return make_model(pos, uv, nrm, tris);
}

// models/cornell_box.obj restated: corners of [-1,1] x [0,2] x [-1,1], six inward-facing quads, one 2 x 3 atlas
// cell each; face order ceiling, z=+1 wall, x=-1 wall, floor, x=+1 wall, z=-1 wall.
ModelData cornell_box() {
    static const double c[9][3] = {{0, 0, 0}, {1, 2, -1}, {1, 0, -1}, {1, 2, 1}, {1, 0, 1}, {-1, 2, -1}, {-1, 0, -1},
                                   {-1, 2, 1}, {-1, 0, 1}};
    struct Face { int corners[4]; double n[3]; int cx, cy; };
    static const Face faces[6] = {{{1, 3, 7, 5}, {0, -1, 0}, 0, 0}, {{4, 8, 7, 3}, {0, 0, -1}, 1, 0},
                                  {{8, 6, 5, 7}, {1, 0, 0}, 0, 1},  {{6, 8, 4, 2}, {0, 1, 0}, 1, 1},
                                  {{2, 4, 3, 1}, {-1, 0, 0}, 0, 2}, {{6, 2, 1, 5}, {0, 0, 1}, 1, 2}};
    static const double third[4] = {0.0, 0.333333, 0.666667, 1.0};
    std::vector<float> pos, uv, nrm;
    std::vector<uint32_t> tris;
    for (const Face& fc : faces) {
        const double u0 = 0.5 * fc.cx, u1 = 0.5 * fc.cx + 0.5, v0 = third[fc.cy], v1 = third[fc.cy + 1];
        const double cell[4][2] = {{u0, v0}, {u1, v0}, {u1, v1}, {u0, v1}};
        const uint32_t b = uint32_t(pos.size() / 3);
        for (int k = 0; k < 4; k++) {
            for (int a = 0; a < 3; a++) pos.push_back(float(c[fc.corners[k]][a]));
            uv.push_back(float(cell[k][0]));
            uv.push_back(float(1.0 - cell[k][1]));
            for (int a = 0; a < 3; a++) nrm.push_back(float(fc.n[a]));
        }
        const uint32_t t[6] = {b, b + 1, b + 2, b, b + 2, b + 3};
        tris.insert(tris.end(), t, t + 6);
    }
    return make_model(pos, uv, nrm, tris);
}

// models/cornell_light.obj restated: a 0.47 x 0.38 panel at y = 1.989 facing -y
ModelData cornell_light() {
    const double p[4][3] = {{-0.24, 1.989, -0.22}, {0.23, 1.989, 0.16}, {-0.24, 1.989, 0.16}, {0.23, 1.989, -0.22}};
    const double t[4][2] = {{0.0, 0.0}, {1.0, 1.0}, {0.0, 1.0}, {1.0, 0.0}};
    std::vector<float> pos, uv, nrm;
    for (int k = 0; k < 4; k++) {
        for (int a = 0; a < 3; a++) pos.push_back(float(p[k][a]));
        uv.push_back(float(t[k][0]));
        uv.push_back(float(1.0 - t[k][1]));
        nrm.push_back(0.0f); nrm.push_back(-1.0f); nrm.push_back(0.0f);
    }
    return make_model(pos, uv, nrm, {0, 1, 2, 0, 3, 1});
}

ModelData uv_sphere(int segments, int rings, double radius) {
    const double pi = 3.141592653589793;
    std::vector<float> pos, uv, nrm;
    std::vector<uint32_t> tris;
    for (int r = 0; r <= rings; r++) {
        const double th = pi * r / rings;
        for (int s = 0; s <= segments; s++) {
            const double ph = 2 * pi * s / segments;
            const double n[3] = {std::sin(th) * std::cos(ph), std::cos(th), std::sin(th) * std::sin(ph)};
            for (int a = 0; a < 3; a++) { pos.push_back(float(radius * n[a])); nrm.push_back(float(n[a])); }
            uv.push_back(float(double(s) / segments));
            uv.push_back(float(double(r) / rings));
        }
    }
    const uint32_t w = uint32_t(segments + 1);
    for (uint32_t r = 0; r < uint32_t(rings); r++)
        for (uint32_t s = 0; s < uint32_t(segments); s++) {
            const uint32_t a = r * w + s, b = a + 1, c = (r + 1) * w + s + 1, d = (r + 1) * w + s;
            if (r != 0) { tris.push_back(a); tris.push_back(b); tris.push_back(c); }
            if (r != uint32_t(rings - 1)) { tris.push_back(a); tris.push_back(c); tris.push_back(d); }
        }
    return make_model(pos, uv, nrm, tris);
}

// stand-in for textures/cornell_texture.png: 2 x 3 atlas, rows stored bottom-up like the reference's loader
// (stbi_set_flip_vertically_on_load, src/graphics/Image.cpp:14)
Image8 cornell_texture(int width, int height) {
    Image8 img;
    img.width = width;
    img.height = height;
    img.rgba.assign(size_t(width) * height * 4, 255);
    const int h3 = height / 3, w2 = width / 2;
    for (int y = h3; y < height; y++) {
        const bool red = y < 2 * h3;
        const int row = height - 1 - y;
        for (int x = 0; x < w2; x++) {
            uint8_t* p = &img.rgba[(size_t(row) * width + x) * 4];
            if (red) { p[0] = 255; p[1] = 63; p[2] = 63; } else { p[0] = 119; p[1] = 203; p[2] = 63; }
        }
    }
    return img;
}

}  // namespace rbhost
