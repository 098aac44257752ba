// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// Minimal header-only stand-in for the subset of g-truc/glm (pinned by the
// reference at 5c46b9c07008ae65cb81ab79cd677ecc1934b903, see
// /root/reference/install.sh:32-33) that the reference rasterizer
// (ext/diff_gaussian_rasterization_hair/cuda_rasterizer/*.cu) touches.
// glm is cloned by the reference's install script and is NOT vendored in
// /root/reference; there is no network here, so the reference sources are
// compiled in place against this shim to obtain oracle/_ref (Oracle-A).
//
// Written from glm's published semantics (column-major mat3, m[c][r]);
// evaluation order of every product follows glm's scalar code path
//   (A*B)[j][i] = A[0][i]*B[j][0] + A[1][i]*B[j][1] + A[2][i]*B[j][2]
//   dot(a,b)    = a.x*b.x + a.y*b.y + a.z*b.z
// so that FMA contraction by nvcc produces the same roundings.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define GLM_SHIM_FN __host__ __device__ inline
#else
#define GLM_SHIM_FN inline
#endif

namespace glm {

struct vec3 {
    float x, y, z;
    GLM_SHIM_FN vec3() : x(0.f), y(0.f), z(0.f) {}
    template <typename A, typename B, typename C>
    GLM_SHIM_FN vec3(A a, B b, C c) : x(float(a)), y(float(b)), z(float(c)) {}
    GLM_SHIM_FN explicit vec3(float s) : x(s), y(s), z(s) {}
    GLM_SHIM_FN float& operator[](int i) { return (&x)[i]; }
    GLM_SHIM_FN const float& operator[](int i) const { return (&x)[i]; }
    GLM_SHIM_FN vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    GLM_SHIM_FN vec3& operator+=(float s) { x += s; y += s; z += s; return *this; }
    GLM_SHIM_FN vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
};

struct vec4 {
    float x, y, z, w;
    GLM_SHIM_FN vec4() : x(0.f), y(0.f), z(0.f), w(0.f) {}
    template <typename A, typename B, typename C, typename D>
    GLM_SHIM_FN vec4(A a, B b, C c, D d) : x(float(a)), y(float(b)), z(float(c)), w(float(d)) {}
    GLM_SHIM_FN float& operator[](int i) { return (&x)[i]; }
    GLM_SHIM_FN const float& operator[](int i) const { return (&x)[i]; }
};

GLM_SHIM_FN vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
GLM_SHIM_FN vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
GLM_SHIM_FN vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
GLM_SHIM_FN vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
GLM_SHIM_FN vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
GLM_SHIM_FN vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
GLM_SHIM_FN vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }

GLM_SHIM_FN float dot(const vec3& a, const vec3& b) {
    vec3 t(a.x * b.x, a.y * b.y, a.z * b.z);
    return t.x + t.y + t.z;
}
GLM_SHIM_FN float length(const vec3& a) { return sqrtf(dot(a, a)); }
GLM_SHIM_FN vec3 max(const vec3& a, float s) {
    return vec3(a.x < s ? s : a.x, a.y < s ? s : a.y, a.z < s ? s : a.z);
}

// column-major 3x3: m[c] is column c, m[c][r] row r of that column
struct mat3 {
    vec3 c[3];
    GLM_SHIM_FN mat3() { c[0] = vec3(1, 0, 0); c[1] = vec3(0, 1, 0); c[2] = vec3(0, 0, 1); }
    GLM_SHIM_FN explicit mat3(float s) { c[0] = vec3(s, 0, 0); c[1] = vec3(0, s, 0); c[2] = vec3(0, 0, s); }
    template <typename X0, typename Y0, typename Z0, typename X1, typename Y1, typename Z1,
              typename X2, typename Y2, typename Z2>
    GLM_SHIM_FN mat3(X0 x0, Y0 y0, Z0 z0, X1 x1, Y1 y1, Z1 z1, X2 x2, Y2 y2, Z2 z2) {
        c[0] = vec3(x0, y0, z0); c[1] = vec3(x1, y1, z1); c[2] = vec3(x2, y2, z2);
    }
    GLM_SHIM_FN vec3& operator[](int i) { return c[i]; }
    GLM_SHIM_FN const vec3& operator[](int i) const { return c[i]; }
};

GLM_SHIM_FN mat3 operator*(const mat3& m1, const mat3& m2) {
    const float A00 = m1[0][0], A01 = m1[0][1], A02 = m1[0][2];
    const float A10 = m1[1][0], A11 = m1[1][1], A12 = m1[1][2];
    const float A20 = m1[2][0], A21 = m1[2][1], A22 = m1[2][2];
    const float B00 = m2[0][0], B01 = m2[0][1], B02 = m2[0][2];
    const float B10 = m2[1][0], B11 = m2[1][1], B12 = m2[1][2];
    const float B20 = m2[2][0], B21 = m2[2][1], B22 = m2[2][2];
    mat3 r;
    r[0][0] = A00 * B00 + A10 * B01 + A20 * B02;
    r[0][1] = A01 * B00 + A11 * B01 + A21 * B02;
    r[0][2] = A02 * B00 + A12 * B01 + A22 * B02;
    r[1][0] = A00 * B10 + A10 * B11 + A20 * B12;
    r[1][1] = A01 * B10 + A11 * B11 + A21 * B12;
    r[1][2] = A02 * B10 + A12 * B11 + A22 * B12;
    r[2][0] = A00 * B20 + A10 * B21 + A20 * B22;
    r[2][1] = A01 * B20 + A11 * B21 + A21 * B22;
    r[2][2] = A02 * B20 + A12 * B21 + A22 * B22;
    return r;
}
GLM_SHIM_FN mat3 operator*(float s, const mat3& m) {
    mat3 r;
    r[0] = m[0] * s; r[1] = m[1] * s; r[2] = m[2] * s;
    return r;
}
GLM_SHIM_FN mat3 operator*(const mat3& m, float s) { return s * m; }
GLM_SHIM_FN mat3 transpose(const mat3& m) {
    return mat3(m[0][0], m[1][0], m[2][0],
                m[0][1], m[1][1], m[2][1],
                m[0][2], m[1][2], m[2][2]);
}

}  // namespace glm
