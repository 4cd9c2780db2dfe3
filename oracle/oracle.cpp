/*
 * oracle.cpp — CPU oracle, geometry half (TEST INFRASTRUCTURE, NOT PRODUCT; see oracle.h).
 *
 * Compiled with -ffp-contract=off: every fused multiply-add below is an explicit fmaf() and
 * everything else is a single-rounded fp32 op, which is the numeric contract the
 * CUDA kernels follow independently.  Reference citations are relative to /root/reference.
 */
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

const uint32_t NONE = 0xFFFFFFFFu;
const float INF = std::numeric_limits<float>::infinity();

struct V3 {
    float x, y, z;
};
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
/* dot(a,b) = fma(ax,bx, fma(ay,by, az*bz)) */
inline float dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
/* cross(a,b).x = fma(ay,bz, -(az*by)) etc. */
inline V3 cross(V3 a, V3 b) {
    return {fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x))};
}
inline uint32_t f2u(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}

template <typename F> void parallel_for(uint64_t n, int threads, F f) {
    if(threads <= 0) threads = orc_hw_threads();
    if(threads == 1 || n < 1024) {
        f(0, n);
        return;
    }
    std::vector<std::thread> pool;
    uint64_t chunk = (n + threads - 1) / threads;
    for(int t = 0; t < threads; t++) {
        uint64_t b = t * chunk, e = std::min(n, b + chunk);
        if(b >= e) break;
        pool.emplace_back([=] { f(b, e); });
    }
    for(auto& th : pool) th.join();
}

/* ---- N3: ray/triangle. Contract at the reference boundary: rt.rgen:257-270 (closest, opaque,
 * cull mask 0xFF), vulkan.cpp:802 (no face culling), vulkan.cpp:904 (opaque geometry); Vulkan ray
 * interval is exclusive (tmin,tmax). attribs = (u,v) -> payload bary (1-u-v,u,v), rt.rchit:13.
 * Moller-Trumbore with NO determinant epsilon (rtcommon.glsl:181-202 `triangle_hit` is a
 * different function used only for light pdfs). */
inline bool intersect(V3 o, V3 d, float tmin, float tmax, V3 v0, V3 e1, V3 e2, float& t, float& u,
                      float& v) {
    V3 p = cross(d, e2);
    float det = dot(e1, p);
    if(det == 0.0f) return false;
    float inv = 1.0f / det;
    V3 s = sub(o, v0);
    u = dot(s, p) * inv;
    if(!(u >= 0.0f && u <= 1.0f)) return false;
    V3 q = cross(s, e1);
    v = dot(d, q) * inv;
    if(!(v >= 0.0f && u + v <= 1.0f)) return false;
    t = dot(e2, q) * inv;
    return t > tmin && t < tmax;
}

/* ---- N5: closest point on triangle (Ericson, Real-Time Collision Detection 5.1.5) expressed in
 * a, ab, ac only. FCPW semantics (README.md:6-8; upstream not vendored -> parity unpinned). */
inline float closest_point(V3 p, V3 a, V3 ab, V3 ac, float& v, float& w) {
    V3 ap = sub(p, a);
    float d1 = dot(ab, ap), d2 = dot(ac, ap);
    V3 bp = sub(ap, ab);
    float d3 = dot(ab, bp), d4 = dot(ac, bp);
    V3 cp = sub(ap, ac);
    float d5 = dot(ab, cp), d6 = dot(ac, cp);
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
    V3 del = {fmaf(-w, ac.x, fmaf(-v, ab.x, ap.x)), fmaf(-w, ac.y, fmaf(-v, ab.y, ap.y)),
              fmaf(-w, ac.z, fmaf(-v, ab.z, ap.z))};
    return dot(del, del);
}
inline V3 point_at(V3 a, V3 ab, V3 ac, float v, float w) {
    return {fmaf(w, ac.x, fmaf(v, ab.x, a.x)), fmaf(w, ac.y, fmaf(v, ab.y, a.y)),
            fmaf(w, ac.z, fmaf(v, ab.z, a.z))};
}

struct Tri {
    V3 v0, e1, e2;
};
inline Tri load_tri(const float* t9) {
    V3 v0{t9[0], t9[1], t9[2]}, v1{t9[3], t9[4], t9[5]}, v2{t9[6], t9[7], t9[8]};
    return {v0, sub(v1, v0), sub(v2, v0)};
}

struct Best {
    float t = INF, u = 0, v = 0;
    uint32_t gid = NONE;
};
/* N4: nearest t wins; equal t -> lowest gid */
inline void consider_hit(Best& b, float t, float u, float v, uint32_t gid) {
    if(t < b.t || (t == b.t && gid < b.gid)) b = {t, u, v, gid};
}
inline void store_hit(uint32_t* out, const Best& b) {
    out[0] = f2u(b.gid == NONE ? INF : b.t);
    out[1] = f2u(b.u);
    out[2] = f2u(b.v);
    out[3] = b.gid;
}

struct BestCP {
    float d2, v = 0, w = 0;
    uint32_t gid = NONE;
};
inline void consider_cp(BestCP& b, float d2, float v, float w, uint32_t gid) {
    if(d2 < b.d2 || (d2 == b.d2 && gid < b.gid)) b = {d2, v, w, gid};
}
inline void store_cp(uint32_t* out, const BestCP& b, const float* tris9) {
    if(b.gid == NONE) {
        out[0] = out[1] = out[2] = 0;
        out[3] = f2u(INF);
        out[4] = NONE;
        out[5] = 0;
        out[6] = out[7] = 0;
        return;
    }
    Tri t = load_tri(tris9 + 9ull * b.gid);
    V3 c = point_at(t.v0, t.e1, t.e2, b.v, b.w);
    out[0] = f2u(c.x), out[1] = f2u(c.y), out[2] = f2u(c.z);
    out[3] = f2u(sqrtf(b.d2));
    out[4] = b.gid;
    out[5] = 0;
    out[6] = f2u(b.v), out[7] = f2u(b.w);
}

inline uint64_t expand21(uint64_t x) {
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
inline uint32_t quant21(float c, float lo, float inv) {
    float n = (c - lo) * inv;
    float q = fminf(fmaxf(n * 2097152.0f, 0.0f), 2097151.0f);
    return (uint32_t)q;
}

struct Box {
    V3 lo{INF, INF, INF}, hi{-INF, -INF, -INF};
    void grow(V3 p) {
        lo = {fminf(lo.x, p.x), fminf(lo.y, p.y), fminf(lo.z, p.z)};
        hi = {fmaxf(hi.x, p.x), fmaxf(hi.y, p.y), fmaxf(hi.z, p.z)};
    }
    void grow(const Box& b) {
        lo = {fminf(lo.x, b.lo.x), fminf(lo.y, b.lo.y), fminf(lo.z, b.lo.z)};
        hi = {fmaxf(hi.x, b.hi.x), fmaxf(hi.y, b.hi.y), fmaxf(hi.z, b.hi.z)};
    }
};

} // namespace

struct orc_bvh {
    uint32_t n = 0;
    const float* tris9 = nullptr; /* borrowed */
    std::vector<Tri> tris;        /* in gid order */
    std::vector<Box> tri_box;     /* exact per-gid AABB */
    Box scene;
    float inflate = 0;
    std::vector<uint64_t> keys;  /* sorted */
    std::vector<uint32_t> order; /* sorted position -> gid */
    std::vector<int32_t> left, right;
    std::vector<Box> node_box; /* exact */
    /* traversal copy: per internal node, inflated boxes of both children */
    struct TNode {
        Box cb[2];
        int32_t c[2];
    };
    std::vector<TNode> tnodes;
};

extern "C" {

int orc_hw_threads(void) {
    unsigned n = std::thread::hardware_concurrency();
    return n ? (int)n : 1;
}

/* rtcommon.glsl:99-109 */
uint32_t orc_tea(uint32_t val0, uint32_t val1) {
    uint32_t v0 = val0, v1 = val1, s0 = 0;
    for(uint32_t n = 0; n < 16; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}
/* rtcommon.glsl:111-116 */
uint32_t orc_lcg(uint32_t* prev) {
    *prev = 1664525u * *prev + 1013904223u;
    return *prev & 0x00FFFFFFu;
}
/* rtcommon.glsl:118-120 */
float orc_randf(uint32_t* prev) { return (float)orc_lcg(prev) / (float)0x01000000; }
/* rtcommon.glsl:128-135 */
float orc_radical_inverse(uint32_t bits) {
    bits = (bits << 16u) | (bits >> 16u);
    bits = ((bits & 0x55555555u) << 1u) | ((bits & 0xAAAAAAAAu) >> 1u);
    bits = ((bits & 0x33333333u) << 2u) | ((bits & 0xCCCCCCCCu) >> 2u);
    bits = ((bits & 0x0F0F0F0Fu) << 4u) | ((bits & 0xF0F0F0F0u) >> 4u);
    bits = ((bits & 0x00FF00FFu) << 8u) | ((bits & 0xFF00FF00u) >> 8u);
    return (float)bits * 2.3283064365386963e-10f;
}

/* N1. model is column-major (lib/mat4.h:272-275): world.x = m0*x + m4*y + m8*z + m12 evaluated
 * as fma(m0,x, fma(m4,y, fma(m8,z, m12))). */
void orc_flatten_object(const float* verts48, const uint32_t* idx, uint32_t n_tris,
                        const float m[16], float* out) {
    for(uint32_t t = 0; t < n_tris; t++) {
        for(int k = 0; k < 3; k++) {
            const float* p = verts48 + 12ull * idx[3 * t + k];
            float x = p[0], y = p[1], z = p[2];
            out[9ull * t + 3 * k + 0] = fmaf(m[0], x, fmaf(m[4], y, fmaf(m[8], z, m[12])));
            out[9ull * t + 3 * k + 1] = fmaf(m[1], x, fmaf(m[5], y, fmaf(m[9], z, m[13])));
            out[9ull * t + 3 * k + 2] = fmaf(m[2], x, fmaf(m[6], y, fmaf(m[10], z, m[14])));
        }
    }
}

int orc_intersect(const float r[8], const float tri9[9], float* t, float* u, float* v) {
    Tri tr = load_tri(tri9);
    return intersect({r[0], r[1], r[2]}, {r[4], r[5], r[6]}, r[3], r[7], tr.v0, tr.e1, tr.e2, *t, *u,
                     *v)
               ? 1
               : 0;
}

float orc_closest_point_tri(const float p[3], const float tri9[9], float c[3], float* v, float* w) {
    Tri tr = load_tri(tri9);
    float d2 = closest_point({p[0], p[1], p[2]}, tr.v0, tr.e1, tr.e2, *v, *w);
    V3 q = point_at(tr.v0, tr.e1, tr.e2, *v, *w);
    c[0] = q.x, c[1] = q.y, c[2] = q.z;
    return d2;
}

void orc_closest_hit_brute(const float* tris9, uint32_t n_tris, const float* rays, uint64_t n,
                           uint32_t* hits, int threads) {
    std::vector<Tri> tris(n_tris);
    for(uint32_t i = 0; i < n_tris; i++) tris[i] = load_tri(tris9 + 9ull * i);
    parallel_for(n, threads, [&](uint64_t b, uint64_t e) {
        for(uint64_t i = b; i < e; i++) {
            const float* r = rays + 8 * i;
            V3 o{r[0], r[1], r[2]}, d{r[4], r[5], r[6]};
            Best best;
            for(uint32_t g = 0; g < n_tris; g++) {
                float t, u, v;
                if(intersect(o, d, r[3], r[7], tris[g].v0, tris[g].e1, tris[g].e2, t, u, v))
                    consider_hit(best, t, u, v, g);
            }
            store_hit(hits + 4 * i, best);
        }
    });
}

/* rt.rgen:272-291: occluded iff any triangle hit in (tmin,tmax) */
void orc_any_hit_brute(const float* tris9, uint32_t n_tris, const float* rays, uint64_t n,
                       uint8_t* occ, int threads) {
    std::vector<Tri> tris(n_tris);
    for(uint32_t i = 0; i < n_tris; i++) tris[i] = load_tri(tris9 + 9ull * i);
    parallel_for(n, threads, [&](uint64_t b, uint64_t e) {
        for(uint64_t i = b; i < e; i++) {
            const float* r = rays + 8 * i;
            V3 o{r[0], r[1], r[2]}, d{r[4], r[5], r[6]};
            uint8_t hit = 0;
            for(uint32_t g = 0; g < n_tris && !hit; g++) {
                float t, u, v;
                hit = intersect(o, d, r[3], r[7], tris[g].v0, tris[g].e1, tris[g].e2, t, u, v);
            }
            occ[i] = hit;
        }
    });
}

void orc_closest_point_brute(const float* tris9, uint32_t n_tris, const float* q, uint64_t n,
                             uint32_t* res, int threads) {
    std::vector<Tri> tris(n_tris);
    for(uint32_t i = 0; i < n_tris; i++) tris[i] = load_tri(tris9 + 9ull * i);
    parallel_for(n, threads, [&](uint64_t b, uint64_t e) {
        for(uint64_t i = b; i < e; i++) {
            V3 p{q[4 * i], q[4 * i + 1], q[4 * i + 2]};
            BestCP best;
            best.d2 = q[4 * i + 3];
            for(uint32_t g = 0; g < n_tris; g++) {
                float v, w;
                float d2 = closest_point(p, tris[g].v0, tris[g].e1, tris[g].e2, v, w);
                consider_cp(best, d2, v, w, g);
            }
            store_cp(res + 8 * i, best, tris9);
        }
    });
}

/* ---- N6 + Karras 2012 ------------------------------------------------------------------------ */
static inline int delta(const std::vector<uint64_t>& k, int64_t n, int64_t i, int64_t j) {
    if(j < 0 || j >= n) return -1;
    uint64_t a = k[i], b = k[j];
    if(a == b) return 64 + __builtin_clz((uint32_t)i ^ (uint32_t)j);
    return __builtin_clzll(a ^ b);
}

orc_bvh* orc_bvh_build(const float* tris9, uint32_t n) {
    orc_bvh* B = new orc_bvh;
    B->n = n;
    B->tris9 = tris9;
    B->tris.resize(n);
    B->tri_box.resize(n);
    for(uint32_t g = 0; g < n; g++) {
        const float* t = tris9 + 9ull * g;
        B->tris[g] = load_tri(t);
        Box b;
        b.grow(V3{t[0], t[1], t[2]});
        b.grow(V3{t[3], t[4], t[5]});
        b.grow(V3{t[6], t[7], t[8]});
        B->tri_box[g] = b;
        B->scene.grow(b);
    }
    if(n == 0) return B;
    V3 lo = B->scene.lo, hi = B->scene.hi;
    float maxabs = fmaxf(fmaxf(fmaxf(fabsf(lo.x), fabsf(hi.x)), fmaxf(fabsf(lo.y), fabsf(hi.y))),
                         fmaxf(fabsf(lo.z), fabsf(hi.z)));
    B->inflate = fmaxf(maxabs, 1e-30f) * 1.9073486328125e-06f; /* 2^-19 */
    V3 ext = sub(hi, lo);
    V3 inv{ext.x > 0 ? 1.0f / ext.x : 0.0f, ext.y > 0 ? 1.0f / ext.y : 0.0f,
           ext.z > 0 ? 1.0f / ext.z : 0.0f};
    std::vector<std::pair<uint64_t, uint32_t>> kv(n);
    for(uint32_t g = 0; g < n; g++) {
        const Box& b = B->tri_box[g];
        V3 c{(b.lo.x + b.hi.x) * 0.5f, (b.lo.y + b.hi.y) * 0.5f, (b.lo.z + b.hi.z) * 0.5f};
        uint64_t code = (expand21(quant21(c.x, lo.x, inv.x)) << 2) |
                        (expand21(quant21(c.y, lo.y, inv.y)) << 1) |
                        expand21(quant21(c.z, lo.z, inv.z));
        kv[g] = {code, g};
    }
    std::stable_sort(kv.begin(), kv.end(),
                     [](const auto& a, const auto& b) { return a.first < b.first; });
    B->keys.resize(n);
    B->order.resize(n);
    for(uint32_t i = 0; i < n; i++) B->keys[i] = kv[i].first, B->order[i] = kv[i].second;

    if(n < 2) return B;
    int64_t N = n;
    B->left.resize(n - 1);
    B->right.resize(n - 1);
    B->node_box.resize(n - 1);
    std::vector<int32_t> parent_of_internal(n - 1, -1), parent_of_leaf(n, -1);
    const auto& k = B->keys;
    for(int64_t i = 0; i < N - 1; i++) {
        int d = (delta(k, N, i, i + 1) - delta(k, N, i, i - 1)) >= 0 ? 1 : -1;
        int dmin = delta(k, N, i, i - d);
        int64_t lmax = 2;
        while(delta(k, N, i, i + lmax * d) > dmin) lmax *= 2;
        int64_t l = 0;
        for(int64_t t = lmax / 2; t >= 1; t /= 2)
            if(delta(k, N, i, i + (l + t) * d) > dmin) l += t;
        int64_t j = i + l * d;
        int dnode = delta(k, N, i, j);
        int64_t s = 0, t = l;
        do {
            t = (t + 1) / 2;
            if(delta(k, N, i, i + (s + t) * d) > dnode) s += t;
        } while(t > 1);
        int64_t gamma = i + s * d + std::min(d, 0);
        int64_t lo_i = std::min(i, j), hi_i = std::max(i, j);
        int32_t L = (lo_i == gamma) ? ~(int32_t)gamma : (int32_t)gamma;
        int32_t R = (hi_i == gamma + 1) ? ~(int32_t)(gamma + 1) : (int32_t)(gamma + 1);
        B->left[i] = L;
        B->right[i] = R;
        if(L < 0) parent_of_leaf[~L] = (int32_t)i; else parent_of_internal[L] = (int32_t)i;
        if(R < 0) parent_of_leaf[~R] = (int32_t)i; else parent_of_internal[R] = (int32_t)i;
    }
    /* bottom-up refit (the GPU does the same with atomic arrival flags) */
    std::vector<uint8_t> arrived(n - 1, 0);
    auto child_box = [&](int32_t c) -> Box { return c < 0 ? B->tri_box[B->order[~c]] : B->node_box[c]; };
    for(uint32_t leaf = 0; leaf < n; leaf++) {
        int32_t p = parent_of_leaf[leaf];
        while(p >= 0) {
            if(!arrived[p]) {
                arrived[p] = 1;
                break;
            }
            Box b = child_box(B->left[p]);
            b.grow(child_box(B->right[p]));
            B->node_box[p] = b;
            p = parent_of_internal[p];
        }
    }
    float e = B->inflate;
    auto inflated = [&](Box b) {
        b.lo = {b.lo.x - e, b.lo.y - e, b.lo.z - e};
        b.hi = {b.hi.x + e, b.hi.y + e, b.hi.z + e};
        return b;
    };
    B->tnodes.resize(n - 1);
    for(uint32_t i = 0; i + 1 < n; i++) {
        B->tnodes[i].c[0] = B->left[i];
        B->tnodes[i].c[1] = B->right[i];
        B->tnodes[i].cb[0] = inflated(child_box(B->left[i]));
        B->tnodes[i].cb[1] = inflated(child_box(B->right[i]));
    }
    return B;
}

/* N6' — the order and topology of the default ("fast trace") build: a top-down 16-bin surface-area-heuristic split of
 * the triangle boxes, restated here from its written definition (the comment at the top of gpu-rt_b200/host/sah_split.h)
 * as a plain sequential recursion:
 *   segment S of items (initially all triangles in id order); centroid c = (lo + hi) * 0.5 of the triangle's AABB;
 *   cb = bounds of the centroids (starting from +-3.0e38); for each axis with ext = cb.hi - cb.lo > 0:
 *   bin = min(15, (int)((c - cb.lo) * (16 / ext))), 0 if that product is negative or NaN; for k = 1..15 with both sides
 *   non-empty: cost = area(union of boxes in bins < k) * count + area(union in bins >= k) * count, area(b) = 2 (ex ey + ey ez + ez ex);
 *   lowest cost below 3.0e38 wins, ties to the lower axis then the lower k; stable partition; no candidate: split in the
 *   middle.  Inner nodes are numbered in preorder, leaves are single triangles at their final position. */
orc_bvh* orc_bvh_build_sah(const float* tris9, uint32_t n) {
    orc_bvh* B = orc_bvh_build(tris9, n); /* boxes, scene box, inflation; order / topology are replaced below */
    if(n < 2) {
        for(uint32_t i = 0; i < n; i++) B->keys[i] = i;
        return B;
    }
    const float BIG = 3.0e38f;
    std::vector<uint32_t> items(n);
    for(uint32_t g = 0; g < n; g++) items[g] = g;
    auto cen = [&](uint32_t g, int ax) {
        const Box& b = B->tri_box[g];
        return ax == 0 ? (b.lo.x + b.hi.x) * 0.5f : ax == 1 ? (b.lo.y + b.hi.y) * 0.5f : (b.lo.z + b.hi.z) * 0.5f;
    };
    struct B6 {
        float lo[3], hi[3];
    };
    auto empty = [&]() { return B6{{BIG, BIG, BIG}, {-BIG, -BIG, -BIG}}; };
    auto add = [&](B6& acc, const Box& b) {
        const float l[3] = {b.lo.x, b.lo.y, b.lo.z}, h[3] = {b.hi.x, b.hi.y, b.hi.z};
        for(int q = 0; q < 3; q++) {
            if(l[q] < acc.lo[q]) acc.lo[q] = l[q];
            if(h[q] > acc.hi[q]) acc.hi[q] = h[q];
        }
    };
    auto merge = [&](B6& acc, const B6& b) {
        for(int q = 0; q < 3; q++) {
            if(b.lo[q] < acc.lo[q]) acc.lo[q] = b.lo[q];
            if(b.hi[q] > acc.hi[q]) acc.hi[q] = b.hi[q];
        }
    };
    auto area = [](const B6& b) {
        float ex = b.hi[0] - b.lo[0], ey = b.hi[1] - b.lo[1], ez = b.hi[2] - b.lo[2];
        return 2.0f * (ex * ey + ey * ez + ez * ex);
    };
    auto bin_of = [](float c, float lo, float scale) {
        float f = (c - lo) * scale;
        if(!(f >= 0.0f)) return 0;
        return f < 16.0f ? (int)f : 15;
    };
    std::vector<int32_t> parent_of_internal(n - 1, -1), parent_of_leaf(n, -1);
    struct Frame {
        uint32_t a, b;
        int32_t me, parent;
        int side;
    };
    std::vector<Frame> todo{{0, n, 0, -1, 0}};
    while(!todo.empty()) {
        Frame f = todo.back();
        todo.pop_back();
        int32_t ref;
        if(f.b - f.a == 1) {
            ref = ~(int32_t)f.a;
            B->order[f.a] = items[f.a];
            parent_of_leaf[f.a] = f.parent;
        } else {
            ref = f.me;
            parent_of_internal[f.me] = f.parent;
            float clo[3] = {BIG, BIG, BIG}, chi[3] = {-BIG, -BIG, -BIG};
            for(uint32_t i = f.a; i < f.b; i++)
                for(int ax = 0; ax < 3; ax++) {
                    float c = cen(items[i], ax);
                    if(c < clo[ax]) clo[ax] = c;
                    if(c > chi[ax]) chi[ax] = c;
                }
            int best_ax = -1, best_k = 0;
            float best = BIG;
            float scale[3] = {0, 0, 0};
            for(int ax = 0; ax < 3; ax++) {
                float ext = chi[ax] - clo[ax];
                if(!(ext > 0.0f)) continue;
                scale[ax] = 16.0f / ext;
                B6 bb[16];
                uint32_t cnt[16] = {0};
                for(int k = 0; k < 16; k++) bb[k] = empty();
                for(uint32_t i = f.a; i < f.b; i++) {
                    int k = bin_of(cen(items[i], ax), clo[ax], scale[ax]);
                    cnt[k]++;
                    add(bb[k], B->tri_box[items[i]]);
                }
                for(int k = 1; k < 16; k++) {
                    B6 L = empty(), R = empty();
                    uint32_t cl = 0, cr = 0;
                    for(int j = 0; j < k; j++)
                        if(cnt[j]) merge(L, bb[j]), cl += cnt[j];
                    for(int j = k; j < 16; j++)
                        if(cnt[j]) merge(R, bb[j]), cr += cnt[j];
                    if(!cl || !cr) continue;
                    float cost = area(L) * (float)cl + area(R) * (float)cr;
                    if(cost < best) best = cost, best_ax = ax, best_k = k;
                }
            }
            uint32_t mid;
            if(best_ax < 0) mid = f.a + (f.b - f.a) / 2;
            else {
                auto it = std::stable_partition(items.begin() + f.a, items.begin() + f.b, [&](uint32_t g) {
                    return bin_of(cen(g, best_ax), clo[best_ax], scale[best_ax]) < best_k;
                });
                mid = (uint32_t)(it - items.begin());
            }
            todo.push_back({mid, f.b, f.me + (int32_t)(mid - f.a), f.me, 1});
            todo.push_back({f.a, mid, f.me + 1, f.me, 0});
        }
        if(f.parent >= 0) (f.side ? B->right : B->left)[f.parent] = ref;
    }
    for(uint32_t i = 0; i < n; i++) B->keys[i] = i; /* the key of a primitive in this build is its position */
    /* exact boxes: preorder numbering puts children after their parent */
    auto child_box = [&](int32_t c) -> Box { return c < 0 ? B->tri_box[B->order[~c]] : B->node_box[c]; };
    for(int64_t i = (int64_t)n - 2; i >= 0; i--) {
        Box b = child_box(B->left[i]);
        b.grow(child_box(B->right[i]));
        B->node_box[i] = b;
    }
    float e = B->inflate;
    auto inflated = [&](Box b) {
        b.lo = {b.lo.x - e, b.lo.y - e, b.lo.z - e};
        b.hi = {b.hi.x + e, b.hi.y + e, b.hi.z + e};
        return b;
    };
    for(uint32_t i = 0; i + 1 < n; i++) {
        B->tnodes[i].c[0] = B->left[i];
        B->tnodes[i].c[1] = B->right[i];
        B->tnodes[i].cb[0] = inflated(child_box(B->left[i]));
        B->tnodes[i].cb[1] = inflated(child_box(B->right[i]));
    }
    return B;
}

void orc_bvh_free(orc_bvh* b) { delete b; }
uint32_t orc_bvh_n_tris(const orc_bvh* b) { return b->n; }
float orc_bvh_inflation(const orc_bvh* b) { return b->inflate; }
void orc_bvh_scene_box(const orc_bvh* b, float o[6]) {
    o[0] = b->scene.lo.x, o[1] = b->scene.lo.y, o[2] = b->scene.lo.z;
    o[3] = b->scene.hi.x, o[4] = b->scene.hi.y, o[5] = b->scene.hi.z;
}
void orc_bvh_morton_keys(const orc_bvh* b, uint64_t* out) {
    memcpy(out, b->keys.data(), 8ull * b->n);
}
void orc_bvh_prim_order(const orc_bvh* b, uint32_t* out) {
    memcpy(out, b->order.data(), 4ull * b->n);
}
void orc_bvh_get_bvh2(const orc_bvh* b, int32_t* l, int32_t* r, float* boxes) {
    if(b->n < 2) return;
    memcpy(l, b->left.data(), 4ull * (b->n - 1));
    memcpy(r, b->right.data(), 4ull * (b->n - 1));
    for(uint32_t i = 0; i + 1 < b->n; i++) {
        const Box& x = b->node_box[i];
        float* o = boxes + 6ull * i;
        o[0] = x.lo.x, o[1] = x.lo.y, o[2] = x.lo.z, o[3] = x.hi.x, o[4] = x.hi.y, o[5] = x.hi.z;
    }
}

} /* extern "C" */

/* conservative slab test; fminf/fmaxf drop NaN operands (0*inf), which only widens the interval */
static inline bool slab(const Box& b, V3 o, V3 id, float tmin, float tmax, float& tn) {
    float x0 = (b.lo.x - o.x) * id.x, x1 = (b.hi.x - o.x) * id.x;
    float y0 = (b.lo.y - o.y) * id.y, y1 = (b.hi.y - o.y) * id.y;
    float z0 = (b.lo.z - o.z) * id.z, z1 = (b.hi.z - o.z) * id.z;
    float n = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), tmin));
    float f = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmax));
    tn = n;
    /* widen by 2 ulp-ish relative slack on both ends (Ize 2013) */
    return n * 0.9999995f <= f * 1.0000005f || n <= f;
}

template <bool ANY>
static void bvh_trace(const orc_bvh* B, const float* rays, uint64_t n, uint32_t* hits, uint8_t* occ,
                      int threads) {
    parallel_for(n, threads, [&](uint64_t b, uint64_t e) {
        struct Ent {
            int32_t node;
            float tn;
        };
        static thread_local std::vector<Ent> stack(256);
        for(uint64_t i = b; i < e; i++) {
            const float* r = rays + 8 * i;
            V3 o{r[0], r[1], r[2]}, d{r[4], r[5], r[6]};
            V3 id{1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
            float tmin = r[3], tmax = r[7];
            Best best;
            bool any = false;
            int sp = 0;
            if(B->n == 1) stack[sp++] = {~0, tmin};
            else if(B->n > 1) stack[sp++] = {0, tmin};
            while(sp && !any) {
                Ent en = stack[--sp];
                float lim = best.gid == NONE ? tmax : best.t;
                if(en.tn > lim) continue;
                if(en.node < 0) {
                    uint32_t g = B->order[~en.node];
                    float t, u, v;
                    const Tri& tr = B->tris[g];
                    if(intersect(o, d, tmin, tmax, tr.v0, tr.e1, tr.e2, t, u, v)) {
                        if(ANY) any = true;
                        else consider_hit(best, t, u, v, g);
                    }
                    continue;
                }
                const orc_bvh::TNode& nd = B->tnodes[en.node];
                float t0, t1;
                bool h0 = slab(nd.cb[0], o, id, tmin, lim, t0);
                bool h1 = slab(nd.cb[1], o, id, tmin, lim, t1);
                if(sp + 2 > (int)stack.size()) stack.resize(stack.size() * 2);
                if(h0 && h1) {
                    if(t0 <= t1) stack[sp++] = {nd.c[1], t1}, stack[sp++] = {nd.c[0], t0};
                    else stack[sp++] = {nd.c[0], t0}, stack[sp++] = {nd.c[1], t1};
                } else if(h0) stack[sp++] = {nd.c[0], t0};
                else if(h1) stack[sp++] = {nd.c[1], t1};
            }
            if(ANY) occ[i] = any;
            else store_hit(hits + 4 * i, best);
        }
    });
}

extern "C" {

void orc_bvh_closest_hit(const orc_bvh* B, const float* rays, uint64_t n, uint32_t* hits, int th) {
    bvh_trace<false>(B, rays, n, hits, nullptr, th);
}
void orc_bvh_any_hit(const orc_bvh* B, const float* rays, uint64_t n, uint8_t* occ, int th) {
    bvh_trace<true>(B, rays, n, nullptr, occ, th);
}

/* single-ray versions for the CPU integrator (oracle_render.cpp); hit4 = t,u,v,gid */
int orc_bvh_trace_one(const orc_bvh* B, const float* ray8, uint32_t* hit4) {
    bvh_trace<false>(B, ray8, 1, hit4, nullptr, 1);
    return hit4[3] != NONE;
}
int orc_bvh_occluded_one(const orc_bvh* B, const float* ray8) {
    uint8_t o = 0;
    bvh_trace<true>(B, ray8, 1, nullptr, &o, 1);
    return o;
}

static inline float box_d2(const Box& b, V3 p) {
    float dx = fmaxf(fmaxf(b.lo.x - p.x, p.x - b.hi.x), 0.0f);
    float dy = fmaxf(fmaxf(b.lo.y - p.y, p.y - b.hi.y), 0.0f);
    float dz = fmaxf(fmaxf(b.lo.z - p.z, p.z - b.hi.z), 0.0f);
    return fmaf(dx, dx, fmaf(dy, dy, dz * dz));
}

void orc_bvh_closest_point(const orc_bvh* B, const float* q, uint64_t n, uint32_t* res, int threads) {
    parallel_for(n, threads, [&](uint64_t b, uint64_t e) {
        struct Ent {
            int32_t node;
            float d2;
        };
        std::vector<Ent> stack(256);
        for(uint64_t i = b; i < e; i++) {
            V3 p{q[4 * i], q[4 * i + 1], q[4 * i + 2]};
            BestCP best;
            best.d2 = q[4 * i + 3];
            int sp = 0;
            if(B->n == 1) stack[sp++] = {~0, 0.0f};
            else if(B->n > 1) stack[sp++] = {0, 0.0f};
            while(sp) {
                Ent en = stack[--sp];
                if(en.d2 > best.d2) continue;
                if(en.node < 0) {
                    uint32_t g = B->order[~en.node];
                    const Tri& tr = B->tris[g];
                    float v, w;
                    float d2 = closest_point(p, tr.v0, tr.e1, tr.e2, v, w);
                    consider_cp(best, d2, v, w, g);
                    continue;
                }
                const orc_bvh::TNode& nd = B->tnodes[en.node];
                float d0 = box_d2(nd.cb[0], p) * 0.999999f, d1 = box_d2(nd.cb[1], p) * 0.999999f;
                if(sp + 2 > (int)stack.size()) stack.resize(stack.size() * 2);
                if(d0 <= d1) {
                    if(d1 <= best.d2) stack[sp++] = {nd.c[1], d1};
                    if(d0 <= best.d2) stack[sp++] = {nd.c[0], d0};
                } else {
                    if(d0 <= best.d2) stack[sp++] = {nd.c[0], d0};
                    if(d1 <= best.d2) stack[sp++] = {nd.c[1], d1};
                }
            }
            store_cp(res + 8 * i, best, B->tris9);
        }
    });
}

/* ray i: s = tea(i, seed); 5 randf draws (rtcommon.glsl:99-120) */
void orc_gen_random_rays(uint64_t n, uint32_t seed, const float box[6], float frac, float tmin,
                         float tmax, float* rays) {
    float lo[3], ext[3];
    for(int a = 0; a < 3; a++) {
        float e = box[3 + a] - box[a];
        lo[a] = box[a] - frac * e;
        ext[a] = e + 2.0f * frac * e;
    }
    for(uint64_t i = 0; i < n; i++) {
        uint32_t s = orc_tea((uint32_t)i, seed);
        float* r = rays + 8 * i;
        for(int a = 0; a < 3; a++) r[a] = lo[a] + ext[a] * orc_randf(&s);
        float z = 1.0f - 2.0f * orc_randf(&s);
        float phi = 6.283185307179586f * orc_randf(&s);
        float rr = sqrtf(fmaxf(0.0f, 1.0f - z * z));
        r[3] = tmin;
        r[4] = rr * cosf(phi), r[5] = rr * sinf(phi), r[6] = z;
        r[7] = tmax;
    }
}

void orc_gen_random_points(uint64_t n, uint32_t seed, const float box[6], float frac, float r2,
                           float* q) {
    float lo[3], ext[3];
    for(int a = 0; a < 3; a++) {
        float e = box[3 + a] - box[a];
        lo[a] = box[a] - frac * e;
        ext[a] = e + 2.0f * frac * e;
    }
    for(uint64_t i = 0; i < n; i++) {
        uint32_t s = orc_tea((uint32_t)i, seed);
        for(int a = 0; a < 3; a++) q[4 * i + a] = lo[a] + ext[a] * orc_randf(&s);
        q[4 * i + 3] = r2;
    }
}

} /* extern "C" */
