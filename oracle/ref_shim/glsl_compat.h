/*
 * glsl_compat.h — just enough GLSL (types, swizzles, operators, builtins, ray-tracing / image / buffer stand-ins) for
 * the reference's own shader sources src/shaders/rt/{rtcommon.glsl, restir.glsl, rt.rgen} to compile as C++
 * (TEST INFRASTRUCTURE, oracle/_ref only).
 *
 * The shader text is not copied into this repository: oracle/make_glsl_ref.py reads it where it lies under
 * /root/reference and writes a lightly rewritten copy into oracle/_ref/ for the duration of the compile (removed afterwards; only the compiled library stays): `inout T x` / `out T x` become
 * `T& x`, decimal literals get an `f` suffix so that arithmetic stays fp32 as in GLSL, layout / binding declarations
 * are dropped (their storage is declared here), constructor calls with two random draws are brace-initialised
 * (GLSL evaluates arguments left to right, C++ does not promise to).
 *
 * What GLSL leaves to the implementation is taken from the numeric contract the oracle and the CUDA kernels use
 * (DESIGN.md §3 N8): dot / cross / matrix * vector with explicit fma, sin / cos / pow from gpurt_detmath.h.  With
 * that, the reference's text and the oracle's restatement must agree bit for bit — whole frames included.
 */
#pragma once
#include <cmath>
#include <cstdint>

#include "../../include/gpurt_detmath.h"

namespace glsl {

typedef unsigned int uint;
struct vec2;
struct vec3;
struct vec4;

struct vec2 {
    float x, y;
    vec2() : x(0), y(0) {}
    explicit vec2(float s) : x(s), y(s) {}
    vec2(float a, float b) : x(a), y(b) {}
    template <class U, class = decltype(U::x)> explicit vec2(const U& u) : x((float)u.x), y((float)u.y) {} /* uvec2 / ivec2 / swizzle */
};
struct ivec2 {
    int x, y;
    ivec2() : x(0), y(0) {}
    ivec2(int a, int b) : x(a), y(b) {}
    template <class U, class = decltype(U::x)> explicit ivec2(const U& u) : x((int)u.x), y((int)u.y) {}
};
struct ivec3 {
    int x, y, z;
    template <class A, class B, class C> ivec3(A a, B b, C c) : x((int)a), y((int)b), z((int)c) {}
};
struct vec3 {
    union {
        struct { float x, y, z; };
        struct { float r, g, b; };
    };
    vec3() : x(0), y(0), z(0) {}
    vec3(float s) : x(s), y(s), z(s) {}
    vec3(float a, float b_, float c) : x(a), y(b_), z(c) {}
    explicit vec3(const vec4& v);
};
/* vec4 with the two swizzles the shaders use on it as l-values and r-values */
struct vec4 {
    struct SwzXY {
        float x, y;
        operator vec2() const { return vec2(x, y); }
        SwzXY& operator=(vec2 v) { return x = v.x, y = v.y, *this; }
    };
    struct SwzXYZ {
        float x, y, z;
        operator vec3() const { return vec3(x, y, z); }
        SwzXYZ& operator/=(float s) { return x = x / s, y = y / s, z = z / s, *this; }
    };
    union {
        struct { float x, y, z, w; };
        struct { float r, g, b, a; };
        SwzXY xy;
        SwzXYZ xyz, rgb;
    };
    vec4() : x(0), y(0), z(0), w(0) {}
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    vec4(vec3 v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    vec4(vec2 v, float c, float d) : x(v.x), y(v.y), z(c), w(d) {}
};
inline vec3::vec3(const vec4& v) : x(v.x), y(v.y), z(v.z) {}
struct mat4 {
    float m[16]; /* column-major */
};
struct mat3 {
    vec3 c0, c1, c2;
    mat3(vec3 a, vec3 b, vec3 c) : c0(a), c1(b), c2(c) {}
};

inline vec2 operator+(vec2 a, vec2 b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator+(vec2 a, float s) { return vec2(a.x + s, a.y + s); }
inline vec2 operator-(vec2 a, float s) { return vec2(a.x - s, a.y - s); }
inline vec2 operator*(vec2 a, float s) { return vec2(a.x * s, a.y * s); }
inline vec2 operator*(vec2 a, vec2 b) { return vec2(a.x * b.x, a.y * b.y); }
inline vec2 operator/(vec2 a, vec2 b) { return vec2(a.x / b.x, a.y / b.y); }

inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator-(vec3 a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 operator*(vec3 a, vec3 b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(vec3 a, vec3 b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator*(vec3 a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, vec3 a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(vec3 a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator/(float s, vec3 a) { return vec3(s / a.x, s / a.y, s / a.z); }
inline vec3 operator-(float s, vec3 a) { return vec3(s - a.x, s - a.y, s - a.z); }
inline vec3 operator-(vec3 a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }
inline vec3 operator+(vec3 a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3& operator+=(vec3& a, vec3 b) { return a = a + b; }
inline vec3& operator*=(vec3& a, vec3 b) { return a = a * b; }
inline vec3& operator/=(vec3& a, float s) { return a = a / s; }

/* N8: fma inside dot / cross / matrix * vector */
inline float dot(vec3 a, vec3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
inline vec3 cross(vec3 a, vec3 b) {
    return vec3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
inline vec4 operator*(const mat4& M, vec4 v) {
    const float* m = M.m;
    return vec4(fmaf(m[0], v.x, fmaf(m[4], v.y, fmaf(m[8], v.z, m[12] * v.w))), fmaf(m[1], v.x, fmaf(m[5], v.y, fmaf(m[9], v.z, m[13] * v.w))),
                fmaf(m[2], v.x, fmaf(m[6], v.y, fmaf(m[10], v.z, m[14] * v.w))), fmaf(m[3], v.x, fmaf(m[7], v.y, fmaf(m[11], v.z, m[15] * v.w))));
}
inline vec3 operator*(const mat3& M, vec3 v) { return M.c0 * v.x + M.c1 * v.y + M.c2 * v.z; }
inline float length(vec3 a) { return sqrtf(dot(a, a)); }
inline vec3 normalize(vec3 a) { return a / length(a); }
inline vec3 reflect(vec3 I, vec3 N) { return I - (2.0f * dot(N, I)) * N; }
inline vec3 mix(vec3 a, vec3 b, float t) { return a * (1.0f - t) + b * t; }
inline float sqrt(float x) { return sqrtf(x); }
inline float cos(float x) { return dm_cos(x); }
inline float sin(float x) { return dm_sin(x); }
inline float pow(float x, float y) { return dm_pow(x, y); }
inline vec3 pow(vec3 x, vec3 y) { return vec3(dm_pow(x.x, y.x), dm_pow(x.y, y.y), dm_pow(x.z, y.z)); }
inline vec3 exp(vec3 x) { return vec3(dm_exp(x.x), dm_exp(x.y), dm_exp(x.z)); }
inline float abs(float x) { return fabsf(x); }
inline float max(float a, float b) { return fmaxf(a, b); }
inline float min(float a, float b) { return fminf(a, b); }
inline uint min(uint a, uint b) { return a < b ? a : b; }
inline vec3 max(vec3 a, vec3 b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline vec3 min(vec3 a, vec3 b) { return vec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
struct bvec2 { bool x, y; };
struct bvec3 { bool x, y, z; };
inline bvec3 greaterThan(vec3 a, vec3 b) { return bvec3{a.x > b.x, a.y > b.y, a.z > b.z}; }
inline bvec2 greaterThan(vec2 a, vec2 b) { return bvec2{a.x > b.x, a.y > b.y}; }
inline bvec2 lessThan(vec2 a, vec2 b) { return bvec2{a.x < b.x, a.y < b.y}; }
inline bool any(bvec3 b) { return b.x || b.y || b.z; }
inline bool all(bvec2 b) { return b.x && b.y; }

/* ---- stand-ins for the descriptor bindings of rt.rgen (rt.rgen:10-56), filled per frame by glsl_ref.cpp ---------- */
struct LaunchVec { /* gl_LaunchIDEXT / gl_LaunchSizeEXT: .x .y and .xy */
    union {
        struct { uint x, y, z; };
        struct { uint x, y; } xy;
    };
};
struct sampler2D { int id; };            /* Textures[i]: scene texture i; ppos / pnorm / palb: previous G-buffers */
struct image2D { float* px; };           /* RGBA32F */
struct accelerationStructureEXT {};
constexpr uint gl_RayFlagsOpaqueEXT = 1u, gl_RayFlagsTerminateOnFirstHitEXT = 4u, gl_RayFlagsSkipClosestHitShaderEXT = 8u;

} // namespace glsl
