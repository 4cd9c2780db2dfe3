/*
 * common.cuh — device-side records and fp32 helpers shared by every kernel of libgpurt.so.
 *
 * Everything here is compiled for sm_100a with -fmad=false: the compiler never contracts a*b+c on
 * its own, so each fused multiply-add in the primitive tests is an explicit fmaf() and matches the
 * numeric contract of DESIGN.md §3 bit for bit (the CPU oracle is written against the same contract
 * independently).  Functions marked GPURT_HD also compile for the host so that tests/emu can replay
 * kernel logic on the CPU while debugging — that harness is test-only and never linked into the
 * product.
 */
#pragma once
#include <cstdint>
#include <cmath>

#if defined(__CUDACC__)
#define GPURT_HD __host__ __device__ __forceinline__
#define GPURT_D __device__ __forceinline__
#else
#define GPURT_HD inline
#define GPURT_D inline
#endif

namespace gpurt {

#if !defined(__CUDACC__)
struct float4 { float x, y, z, w; };
struct float3 { float x, y, z; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct int2 { int x, y; };
#endif

struct F3 {
    float x, y, z;
};
GPURT_HD F3 f3(float x, float y, float z) { return F3{x, y, z}; }
GPURT_HD F3 operator-(F3 a, F3 b) { return F3{a.x - b.x, a.y - b.y, a.z - b.z}; }
GPURT_HD F3 operator+(F3 a, F3 b) { return F3{a.x + b.x, a.y + b.y, a.z + b.z}; }
GPURT_HD F3 operator*(F3 a, float s) { return F3{a.x * s, a.y * s, a.z * s}; }
/* contract: dot = fma(ax,bx, fma(ay,by, az*bz)); cross.x = fma(ay,bz, -(az*by)) */
GPURT_HD float dot3(F3 a, F3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
GPURT_HD F3 cross3(F3 a, F3 b) {
    return F3{fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x))};
}

GPURT_HD unsigned f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    unsigned u;
    __builtin_memcpy(&u, &f, 4);
    return u;
#endif
}
GPURT_HD float u2f(unsigned u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    __builtin_memcpy(&f, &u, 4);
    return f;
#endif
}

constexpr unsigned kNoHit = 0xFFFFFFFFu;
#define GPURT_INF (u2f(0x7f800000u))

/* ---- N3 ray/triangle (Moller-Trumbore, no det epsilon, exclusive interval) -------------------- */
GPURT_HD bool intersect_tri(F3 o, F3 d, float tmin, float tmax, F3 v0, F3 e1, F3 e2, float& t,
                            float& u, float& v) {
    F3 p = cross3(d, e2);
    float det = dot3(e1, p);
    if(det == 0.0f) return false;
    float inv = 1.0f / det;
    F3 s = o - v0;
    u = dot3(s, p) * inv;
    if(!(u >= 0.0f && u <= 1.0f)) return false;
    F3 q = cross3(s, e1);
    v = dot3(d, q) * inv;
    if(!(v >= 0.0f && u + v <= 1.0f)) return false;
    t = dot3(e2, q) * inv;
    return t > tmin && t < tmax;
}

/* ---- N5 closest point on triangle (Ericson 5.1.5 in a, ab, ac form) --------------------------- */
GPURT_HD float closest_point_tri(F3 p, F3 a, F3 ab, F3 ac, float& v, float& w) {
    F3 ap = p - a;
    float d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    F3 bp = ap - ab;
    float d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    F3 cp = ap - ac;
    float d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    float vc = fmaf(d1, d4, -(d3 * d2));
    float vb = fmaf(d5, d2, -(d1 * d6));
    float va = fmaf(d3, d6, -(d5 * d4));
    if(d1 <= 0.0f && d2 <= 0.0f) {
        v = 0.0f, w = 0.0f;
    } else if(d3 >= 0.0f && d4 <= d3) {
        v = 1.0f, w = 0.0f;
    } else if(vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) {
        v = d1 / (d1 - d3), w = 0.0f;
    } else if(d6 >= 0.0f && d5 <= d6) {
        v = 0.0f, w = 1.0f;
    } else if(vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) {
        v = 0.0f, w = d2 / (d2 - d6);
    } else if(va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) {
        w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        v = 1.0f - w;
    } else {
        float denom = 1.0f / (va + vb + vc);
        v = vb * denom;
        w = vc * denom;
    }
    F3 del = F3{fmaf(-w, ac.x, fmaf(-v, ab.x, ap.x)), fmaf(-w, ac.y, fmaf(-v, ab.y, ap.y)),
                fmaf(-w, ac.z, fmaf(-v, ab.z, ap.z))};
    return dot3(del, del);
}
GPURT_HD F3 tri_point(F3 a, F3 ab, F3 ac, float v, float w) {
    return F3{fmaf(w, ac.x, fmaf(v, ab.x, a.x)), fmaf(w, ac.y, fmaf(v, ab.y, a.y)),
              fmaf(w, ac.z, fmaf(v, ab.z, a.z))};
}

/* ---- N6 Morton -------------------------------------------------------------------------------- */
GPURT_HD unsigned long long expand21(unsigned long long x) {
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
GPURT_HD unsigned quant21(float c, float lo, float inv) {
    float n = (c - lo) * inv;
    float q = fminf(fmaxf(n * 2097152.0f, 0.0f), 2097151.0f);
    return (unsigned)q;
}

/* ---- records ---------------------------------------------------------------------------------- */
/* Triangle: 3 x float4 = 48 B.  r0 = v0.xyz | gid, r1 = e1.xyz | obj, r2 = e2.xyz | prim. */
constexpr int kTriVec4 = 3;
/* Wide node: 5 x float4 = 80 B (layout in bvh8.cuh). */
constexpr int kNodeVec4 = 5;
constexpr int kMaxLeafTris = 3;

struct Box3 {
    F3 lo, hi;
};
GPURT_HD float box_area(const Box3& b) {
    F3 e = b.hi - b.lo;
    return 2.0f * (e.x * e.y + e.y * e.z + e.z * e.x);
}

} // namespace gpurt
