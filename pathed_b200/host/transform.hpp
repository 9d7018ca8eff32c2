// Host-side Point/Vector/Transform arithmetic the parsers need, in the reference's exact operation
// order so that vertices, normals and environment matrices come out bit-identical:
//   Transform::apply      /root/reference/src/transform.cpp:61-100
//   matrix::{scale,rotateX,rotateY,rotateZ,translate,multiply}   src/matrix.cpp:40-152
//   TRS composition + analytic inverse                          src/scene_parser.cpp:716-812
#pragma once

#include <cmath>
#include <cstring>

namespace pathed {

struct Vec3 {
    float x = 0.f, y = 0.f, z = 0.f;
    Vec3() {}
    Vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    Vec3 operator-(const Vec3 &o) const { return Vec3(x - o.x, y - o.y, z - o.z); }
    Vec3 operator+(const Vec3 &o) const { return Vec3(x + o.x, y + o.y, z + o.z); }
    Vec3 operator*(float t) const { return Vec3(x * t, y * t, z * t); }
    float dot(const Vec3 &o) const { return x * o.x + y * o.y + z * o.z; }
    Vec3 cross(const Vec3 &o) const { return Vec3((y * o.z) - (z * o.y), (z * o.x) - (x * o.z), (x * o.y) - (y * o.x)); }
    float length() const { return sqrtf(x * x + y * y + z * z); }
    Vec3 normalized() const { const float n = sqrtf(x * x + y * y + z * z); return Vec3(x / n, y / n, z / n); }
};

struct Transform {
    float m[4][4];
    float inv[4][4];

    Transform() { identity(m); identity(inv); }

    static void identity(float (&a)[4][4])
    {
        for (int r = 0; r < 4; r++) { for (int c = 0; c < 4; c++) { a[r][c] = r == c ? 1.f : 0.f; } }
    }
    // result = left * result   (src/matrix.cpp: every builder pre-multiplies)
    static void premultiply(float (&result)[4][4], const float (&left)[4][4])
    {
        float original[4][4];
        memcpy(original, result, sizeof(original));
        for (int row = 0; row < 4; row++) {
            for (int col = 0; col < 4; col++) {
                result[row][col] = 0.f;
                for (int i = 0; i < 4; i++) { result[row][col] += left[row][i] * original[i][col]; }
            }
        }
    }
    static void scale(float (&a)[4][4], float x, float y, float z)
    {
        float s[4][4]; identity(s); s[0][0] = x; s[1][1] = y; s[2][2] = z; premultiply(a, s);
    }
    static void translate(float (&a)[4][4], float x, float y, float z)
    {
        float t[4][4]; identity(t); t[0][3] = x; t[1][3] = y; t[2][3] = z; premultiply(a, t);
    }
    static void rotateX(float (&a)[4][4], float theta)
    {
        float r[4][4]; identity(r);
        r[1][1] = cosf(theta); r[1][2] = -sinf(theta); r[2][1] = sinf(theta); r[2][2] = cosf(theta);
        premultiply(a, r);
    }
    static void rotateY(float (&a)[4][4], float theta)
    {
        float r[4][4]; identity(r);
        r[0][0] = cosf(theta); r[0][2] = sinf(theta); r[2][0] = -sinf(theta); r[2][2] = cosf(theta);
        premultiply(a, r);
    }
    static void rotateZ(float (&a)[4][4], float theta)
    {
        float r[4][4]; identity(r);
        r[0][0] = cosf(theta); r[0][1] = -sinf(theta); r[1][0] = sinf(theta); r[1][1] = cosf(theta);
        premultiply(a, r);
    }

    Vec3 applyPoint(const Vec3 &p) const
    {
        return Vec3(m[0][0] * p.x + m[0][1] * p.y + m[0][2] * p.z + m[0][3],
                    m[1][0] * p.x + m[1][1] * p.y + m[1][2] * p.z + m[1][3],
                    m[2][0] * p.x + m[2][1] * p.y + m[2][2] * p.z + m[2][3]);
    }
    Vec3 applyVector(const Vec3 &v) const
    {
        return Vec3(m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z,
                    m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z,
                    m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z);
    }
};

} // namespace pathed
