// lmath.hpp -- the small fp32 vector/matrix kit of the Luz host mirror.
//
// The reference does its host-side math with glm 0.9.9.8 (GLM_FORCE_RADIANS,
// GLM_FORCE_DEPTH_ZERO_TO_ONE, right-handed; source/Core/Luzpch.hpp:20-30).  These functions
// restate the published formulas of the handful of glm entry points the lighting path depends on
// (translate / scale / quaternion-from-Euler / mat4_cast / perspectiveRH_ZO / lookAtRH / inverse /
// operator*), keeping glm's operation order so that SceneBlock matrices come out bit-identical
// to the reference's (checked against tests/golden/ref_host_*.json, which is produced by the
// reference's own compiled code).  Build with -ffp-contract=off.
#pragma once

#include <cmath>
#include <cstring>

namespace lm {

struct vec2 {
    float x = 0, y = 0;
};
struct vec3 {
    float x = 0, y = 0, z = 0;
    vec3() = default;
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit vec3(float a) : x(a), y(a), z(a) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};
struct vec4 {
    float x = 0, y = 0, z = 0, w = 0;
    vec4() = default;
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    vec4(vec3 v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};
struct mat4 { // column-major: c[col][row]
    vec4 c[4];
    mat4() : mat4(1.0f) {}
    explicit mat4(float d) {
        c[0] = vec4(d, 0, 0, 0);
        c[1] = vec4(0, d, 0, 0);
        c[2] = vec4(0, 0, d, 0);
        c[3] = vec4(0, 0, 0, d);
    }
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
    const float* data() const { return &c[0].x; }
    float* data() { return &c[0].x; }
};
struct quat {
    float w = 1, x = 0, y = 0, z = 0;
};

inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(vec3 a, vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline vec4 operator+(vec4 a, vec4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline vec4 operator-(vec4 a, vec4 b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline vec4 operator*(vec4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline vec4 operator*(vec4 a, vec4 b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }

inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }
inline vec3 radians(vec3 d) { return {radians(d.x), radians(d.y), radians(d.z)}; }

inline float dot(vec3 a, vec3 b) {
    const vec3 t = a * b;
    return t.x + t.y + t.z;
}
inline vec3 cross(vec3 a, vec3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline vec3 normalize(vec3 v) { return v * (1.0f / std::sqrt(dot(v, v))); }

// mat4 * vec4: ((c0*x + c1*y) + (c2*z + c3*w))
inline vec4 operator*(const mat4& m, vec4 v) {
    const vec4 a0 = m[0] * v.x, a1 = m[1] * v.y, a2 = m[2] * v.z, a3 = m[3] * v.w;
    return (a0 + a1) + (a2 + a3);
}
// mat4 * mat4: each result column summed left to right
inline mat4 operator*(const mat4& a, const mat4& b) {
    mat4 r(0.0f);
    for (int j = 0; j < 4; j++) r[j] = ((a[0] * b[j].x + a[1] * b[j].y) + a[2] * b[j].z) + a[3] * b[j].w;
    return r;
}

inline mat4 translate(const mat4& m, vec3 v) {
    mat4 r = m;
    r[3] = ((m[0] * v.x + m[1] * v.y) + m[2] * v.z) + m[3];
    return r;
}
inline mat4 translate(vec3 v) { return translate(mat4(1.0f), v); }
inline mat4 scale(vec3 v) {
    const mat4 m(1.0f);
    mat4 r(0.0f);
    r[0] = m[0] * v.x;
    r[1] = m[1] * v.y;
    r[2] = m[2] * v.z;
    r[3] = m[3];
    return r;
}

// quaternion from Euler angles (radians), pitch=x yaw=y roll=z
inline quat quat_from_euler(vec3 e) {
    const vec3 h = e * 0.5f;
    const vec3 c(std::cos(h.x), std::cos(h.y), std::cos(h.z));
    const vec3 s(std::sin(h.x), std::sin(h.y), std::sin(h.z));
    quat q;
    q.w = c.x * c.y * c.z + s.x * s.y * s.z;
    q.x = s.x * c.y * c.z - c.x * s.y * s.z;
    q.y = c.x * s.y * c.z + s.x * c.y * s.z;
    q.z = c.x * c.y * s.z - s.x * s.y * c.z;
    return q;
}
inline mat4 mat4_cast(quat q) {
    mat4 r(1.0f);
    const float qxx = q.x * q.x, qyy = q.y * q.y, qzz = q.z * q.z;
    const float qxz = q.x * q.z, qxy = q.x * q.y, qyz = q.y * q.z;
    const float qwx = q.w * q.x, qwy = q.w * q.y, qwz = q.w * q.z;
    r[0][0] = 1.0f - 2.0f * (qyy + qzz);
    r[0][1] = 2.0f * (qxy + qwz);
    r[0][2] = 2.0f * (qxz - qwy);
    r[1][0] = 2.0f * (qxy - qwz);
    r[1][1] = 1.0f - 2.0f * (qxx + qzz);
    r[1][2] = 2.0f * (qyz + qwx);
    r[2][0] = 2.0f * (qxz + qwy);
    r[2][1] = 2.0f * (qyz - qwx);
    r[2][2] = 1.0f - 2.0f * (qxx + qyy);
    return r;
}

// right-handed, depth 0..1
inline mat4 perspective(float fovy, float aspect, float zNear, float zFar) {
    const float t = std::tan(fovy / 2.0f);
    mat4 r(0.0f);
    r[0][0] = 1.0f / (aspect * t);
    r[1][1] = 1.0f / t;
    r[2][2] = zFar / (zNear - zFar);
    r[2][3] = -1.0f;
    r[3][2] = -(zFar * zNear) / (zFar - zNear);
    return r;
}
inline mat4 ortho(float l, float r_, float b, float t, float n, float f) {
    mat4 r(1.0f);
    r[0][0] = 2.0f / (r_ - l);
    r[1][1] = 2.0f / (t - b);
    r[2][2] = -1.0f / (f - n);
    r[3][0] = -(r_ + l) / (r_ - l);
    r[3][1] = -(t + b) / (t - b);
    r[3][2] = -n / (f - n);
    return r;
}
inline mat4 look_at(vec3 eye, vec3 center, vec3 up) {
    const vec3 f = normalize(center - eye);
    const vec3 s = normalize(cross(f, up));
    const vec3 u = cross(s, f);
    mat4 r(1.0f);
    r[0][0] = s.x;
    r[1][0] = s.y;
    r[2][0] = s.z;
    r[0][1] = u.x;
    r[1][1] = u.y;
    r[2][1] = u.z;
    r[0][2] = -f.x;
    r[1][2] = -f.y;
    r[2][2] = -f.z;
    r[3][0] = -dot(s, eye);
    r[3][1] = -dot(u, eye);
    r[3][2] = dot(f, eye);
    return r;
}
inline mat4 rotate(const mat4& m, float angle, vec3 axis_in) {
    const float c = std::cos(angle), s = std::sin(angle);
    const vec3 axis = normalize(axis_in);
    const vec3 temp = axis * (1.0f - c);
    float R[3][3];
    R[0][0] = c + temp[0] * axis[0];
    R[0][1] = temp[0] * axis[1] + s * axis[2];
    R[0][2] = temp[0] * axis[2] - s * axis[1];
    R[1][0] = temp[1] * axis[0] - s * axis[2];
    R[1][1] = c + temp[1] * axis[1];
    R[1][2] = temp[1] * axis[2] + s * axis[0];
    R[2][0] = temp[2] * axis[0] + s * axis[1];
    R[2][1] = temp[2] * axis[1] - s * axis[0];
    R[2][2] = c + temp[2] * axis[2];
    mat4 r(0.0f);
    r[0] = (m[0] * R[0][0] + m[1] * R[0][1]) + m[2] * R[0][2];
    r[1] = (m[0] * R[1][0] + m[1] * R[1][1]) + m[2] * R[1][2];
    r[2] = (m[0] * R[2][0] + m[1] * R[2][1]) + m[2] * R[2][2];
    r[3] = m[3];
    return r;
}

// cofactor inverse, in the operation order of the classic 2x2-sub-determinant formulation
inline mat4 inverse(const mat4& m) {
    const float c00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
    const float c02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
    const float c03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
    const float c04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
    const float c06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
    const float c07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
    const float c08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
    const float c10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
    const float c11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
    const float c12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
    const float c14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
    const float c15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
    const float c16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
    const float c18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
    const float c19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
    const float c20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
    const float c22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
    const float c23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
    const vec4 f0(c00, c00, c02, c03), f1(c04, c04, c06, c07), f2(c08, c08, c10, c11);
    const vec4 f3(c12, c12, c14, c15), f4(c16, c16, c18, c19), f5(c20, c20, c22, c23);
    const vec4 v0(m[1][0], m[0][0], m[0][0], m[0][0]), v1(m[1][1], m[0][1], m[0][1], m[0][1]);
    const vec4 v2(m[1][2], m[0][2], m[0][2], m[0][2]), v3(m[1][3], m[0][3], m[0][3], m[0][3]);
    const vec4 i0 = (v1 * f0 - v2 * f1) + v3 * f2;
    const vec4 i1 = (v0 * f0 - v2 * f3) + v3 * f4;
    const vec4 i2 = (v0 * f1 - v1 * f3) + v3 * f5;
    const vec4 i3 = (v0 * f2 - v1 * f4) + v2 * f5;
    const vec4 sa(+1, -1, +1, -1), sb(-1, +1, -1, +1);
    mat4 inv(0.0f);
    inv[0] = i0 * sa;
    inv[1] = i1 * sb;
    inv[2] = i2 * sa;
    inv[3] = i3 * sb;
    const vec4 row0(inv[0][0], inv[1][0], inv[2][0], inv[3][0]);
    const vec4 d0 = m[0] * row0;
    const float d1 = (d0.x + d0.y) + (d0.z + d0.w);
    const float ood = 1.0f / d1;
    mat4 r(0.0f);
    for (int k = 0; k < 4; k++) r[k] = inv[k] * ood;
    return r;
}

} // namespace lm
