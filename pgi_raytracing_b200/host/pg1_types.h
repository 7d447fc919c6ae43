// pg1_types.h -- value types of the reference's host surface, restated (pg1/vector3.h, pg1/matrix3x3.h, pg1/structs.h,
// embree3/rtcore_ray.h).  Host code only; the device has its own copies in csrc/common.cuh.
#pragma once
#include <cmath>
#include <cstdint>

// pg1/vector3.h:23-129.  Normalize() scales by 1/sqrt(len^2) and leaves the zero vector untouched (pg1/vector3.cpp:24-36).
struct Vector3 {
    float x = 0.0f, y = 0.0f, z = 0.0f;
    Vector3() = default;
    Vector3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    float SqrL2Norm() const { return x * x + y * y + z * z; }
    float L2Norm() const { return std::sqrt(SqrL2Norm()); }
    void Normalize() {
        const float n = SqrL2Norm();
        if (n != 0) { const float rn = 1 / std::sqrt(n); x *= rn; y *= rn; z *= rn; }
    }
    Vector3 CrossProduct(const Vector3& v) const { return Vector3(y * v.z - z * v.y, z * v.x - x * v.z, x * v.y - y * v.x); }
    float DotProduct(const Vector3& v) const { return x * v.x + y * v.y + z * v.z; }
};
inline Vector3 operator+(const Vector3& a, const Vector3& b) { return Vector3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vector3 operator-(const Vector3& a, const Vector3& b) { return Vector3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vector3 operator*(const Vector3& a, float s) { return Vector3(s * a.x, s * a.y, s * a.z); }
inline Vector3 operator*(float s, const Vector3& a) { return Vector3(s * a.x, s * a.y, s * a.z); }

// pg1/matrix3x3.h:13-86: row major; the three-vector constructor takes basis COLUMNS (pg1/matrix3x3.cpp:32-45).
struct Matrix3x3 {
    float m[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    Matrix3x3() = default;
    Matrix3x3(const Vector3& ax, const Vector3& ay, const Vector3& az) {
        m[0][0] = ax.x; m[0][1] = ay.x; m[0][2] = az.x;
        m[1][0] = ax.y; m[1][1] = ay.y; m[1][2] = az.y;
        m[2][0] = ax.z; m[2][1] = ay.z; m[2][2] = az.z;
    }
};
inline Vector3 operator*(const Matrix3x3& a, const Vector3& b) {   // pg1/matrix3x3.cpp:68-73
    return Vector3(a.m[0][0] * b.x + a.m[0][1] * b.y + a.m[0][2] * b.z, a.m[1][0] * b.x + a.m[1][1] * b.y + a.m[1][2] * b.z,
                   a.m[2][0] * b.x + a.m[2][1] * b.y + a.m[2][2] * b.z);
}

inline float deg2rad(float d) { return d * 3.14159265358979323846f / 180.0f; }   // pg1/mymath.h:24-27

// pg1/structs.h:3-16
struct Coord2f { float u, v; };
struct Color3f { float r, g, b; };
struct alignas(16) Color4f { float r, g, b, a; };

// embree3/rtcore_ray.h:11-27 (RTCRay, 48 bytes); `time` carries the IOR of the medium (pg1/raytracer.cpp:200,228,416)
struct alignas(16) RTCRay {
    float org_x, org_y, org_z, tnear;
    float dir_x, dir_y, dir_z, time;
    float tfar;
    unsigned int mask, id, flags;
};
