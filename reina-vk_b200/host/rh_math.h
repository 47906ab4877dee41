// Host-side matrix helpers for the headless renderer: the handful of glm calls the reference's host makes
// (src/graphics/Camera.cpp:10-11 perspective / lookAt / inverse, src/Reina.cpp:106 translate / scale), restated.
// Matrices use glm's memory layout: m[c][r] is column c, row r; the 16 floats go to the kernels unchanged.
#pragma once
#include <array>
#include <cmath>
#include <stdexcept>

namespace rbhost {

struct Vec3d { double x, y, z; };
inline Vec3d operator-(Vec3d a, Vec3d b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline double dot(Vec3d a, Vec3d b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3d cross(Vec3d a, Vec3d b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline Vec3d normalize(Vec3d a) { double l = std::sqrt(dot(a, a)); return {a.x / l, a.y / l, a.z / l}; }

using Mat4d = std::array<std::array<double, 4>, 4>;   // [column][row]
using Mat4f = std::array<float, 16>;                  // column-major, what RB200Instance.transform holds

inline Mat4d identity_d() {
    Mat4d m{};
    for (int i = 0; i < 4; i++) m[i][i] = 1.0;
    return m;
}

inline Mat4f identity() {
    Mat4f m{};
    m[0] = m[5] = m[10] = m[15] = 1.0f;
    return m;
}

// glm::perspective, right-handed, depth -1..1
inline Mat4d perspective(double fovy, double aspect, double znear, double zfar) {
    double t = std::tan(fovy / 2.0);
    Mat4d m{};
    m[0][0] = 1.0 / (aspect * t);
    m[1][1] = 1.0 / t;
    m[2][2] = -(zfar + znear) / (zfar - znear);
    m[2][3] = -1.0;
    m[3][2] = -(2.0 * zfar * znear) / (zfar - znear);
    return m;
}

// glm::lookAt, right-handed
inline Mat4d look_at(Vec3d eye, Vec3d center, Vec3d up = {0.0, 1.0, 0.0}) {
    Vec3d f = normalize(center - eye);
    Vec3d s = normalize(cross(f, up));
    Vec3d u = cross(s, f);
    Mat4d m = identity_d();
    m[0][0] = s.x; m[1][0] = s.y; m[2][0] = s.z;
    m[0][1] = u.x; m[1][1] = u.y; m[2][1] = u.z;
    m[0][2] = -f.x; m[1][2] = -f.y; m[2][2] = -f.z;
    m[3][0] = -dot(s, eye); m[3][1] = -dot(u, eye); m[3][2] = dot(f, eye);
    return m;
}

// general inverse by Gauss-Jordan elimination with partial pivoting (the layout's transpose commutes with it)
inline Mat4d inverse(const Mat4d& in) {
    double a[4][8];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
            a[r][c] = in[c][r];
            a[r][4 + c] = r == c ? 1.0 : 0.0;
        }
    for (int col = 0; col < 4; col++) {
        int piv = col;
        for (int r = col + 1; r < 4; r++)
            if (std::fabs(a[r][col]) > std::fabs(a[piv][col])) piv = r;
        if (a[piv][col] == 0.0) throw std::runtime_error("singular matrix");
        if (piv != col)
            for (int c = 0; c < 8; c++) std::swap(a[piv][c], a[col][c]);
        double inv = 1.0 / a[col][col];
        for (int c = 0; c < 8; c++) a[col][c] *= inv;
        for (int r = 0; r < 4; r++) {
            if (r == col) continue;
            double f = a[r][col];
            if (f != 0.0)
                for (int c = 0; c < 8; c++) a[r][c] -= f * a[col][c];
        }
    }
    Mat4d out{};
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) out[c][r] = a[r][4 + c];
    return out;
}

inline Mat4f to_float(const Mat4d& m) {
    Mat4f o{};
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) o[c * 4 + r] = static_cast<float>(m[c][r]);
    return o;
}

inline Mat4d translate_d(double x, double y, double z) {
    Mat4d m = identity_d();
    m[3][0] = x; m[3][1] = y; m[3][2] = z;
    return m;
}

inline Mat4d scale_d(double x, double y, double z) {
    Mat4d m = identity_d();
    m[0][0] = x; m[1][1] = y; m[2][2] = z;
    return m;
}

// mathematical product A * B
inline Mat4d mul(const Mat4d& A, const Mat4d& B) {
    Mat4d o{};
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) {
            double s = 0.0;
            for (int k = 0; k < 4; k++) s += A[k][r] * B[c][k];
            o[c][r] = s;
        }
    return o;
}

}  // namespace rbhost
