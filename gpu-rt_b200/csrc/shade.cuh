/*
 * shade.cuh — device-side shading library of the wavefront integrator: everything rt.rgen does
 * between two traceRayEXT calls (src/shaders/rt/rt.rgen:62-549, rtcommon.glsl:99-369,
 * restir.glsl:2-35), as __device__ functions over the flat scene arrays.
 *
 * fp32 contract N8: component-wise single-rounded ops in GLSL source order, fma only
 * inside dot3 / cross3 / mat*vec, normalize(v) = v / sqrt(dot3(v,v)), sin/cos/pow from
 * gpurt_detmath.h; compiled with -fmad=false.  Shadow / light rays inside an integrator are traced
 * inline with the same traverse8 core as the batch kernels.
 */
#pragma once
#include "../../include/gpurt_detmath.h"
#include "scene_view.cuh"
#include "traverse.cuh"

namespace gpurt {

/* Like traverse.cuh this file also compiles for the host: tests/emu replays the per-pixel functions
 * below on the CPU against the oracle (test-only; the product launches them from render.cu). */
#if defined(__CUDACC__)
#define SH_D __device__ __forceinline__
#define SH_D_CALL __device__ __noinline__ /* large bodies with more than one call site: one copy, called */
#define SH_CONST __constant__
#else
#define SH_D inline
#define SH_D_CALL inline
#define SH_CONST static
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
#endif

constexpr float kPiGlsl = 3.141592f;       /* rtcommon.glsl:6 */
constexpr float kLargeDist = 10000000.0f;  /* rtcommon.glsl:7 */
constexpr float kEps = 0.00001f;           /* rtcommon.glsl:8 */

SH_D F3 f3s(float s) { return F3{s, s, s}; }
SH_D F3 operator-(F3 a) { return F3{-a.x, -a.y, -a.z}; }
SH_D F3 operator*(F3 a, F3 b) { return F3{a.x * b.x, a.y * b.y, a.z * b.z}; }
SH_D F3 operator*(float s, F3 a) { return F3{s * a.x, s * a.y, s * a.z}; }
SH_D F3 operator/(F3 a, float s) { return F3{a.x / s, a.y / s, a.z / s}; }
SH_D F3 operator/(F3 a, F3 b) { return F3{a.x / b.x, a.y / b.y, a.z / b.z}; }
SH_D float length3(F3 a) { return sqrtf(dot3(a, a)); }
SH_D F3 normalize3(F3 a) { return a / length3(a); }
SH_D F3 reflect3(F3 I, F3 N) { return I - (2.0f * dot3(N, I)) * N; }
SH_D F3 mix3(F3 a, F3 b, float t) { return a * (1.0f - t) + b * t; }
SH_D bool any_gt0(F3 a) { return a.x > 0 || a.y > 0 || a.z > 0; }

struct F4 {
    float x, y, z, w;
};
/* column-major mat4 * vec4 */
SH_D F4 mul4(const float* m, float x, float y, float z, float w) {
    F4 r;
    r.x = fmaf(m[0], x, fmaf(m[4], y, fmaf(m[8], z, m[12] * w)));
    r.y = fmaf(m[1], x, fmaf(m[5], y, fmaf(m[9], z, m[13] * w)));
    r.z = fmaf(m[2], x, fmaf(m[6], y, fmaf(m[10], z, m[14] * w)));
    r.w = fmaf(m[3], x, fmaf(m[7], y, fmaf(m[11], z, m[15] * w)));
    return r;
}
SH_D F3 xform_point(const float* m, F3 p) {
    return F3{fmaf(m[0], p.x, fmaf(m[4], p.y, fmaf(m[8], p.z, m[12]))),
              fmaf(m[1], p.x, fmaf(m[5], p.y, fmaf(m[9], p.z, m[13]))),
              fmaf(m[2], p.x, fmaf(m[6], p.y, fmaf(m[10], p.z, m[14])))};
}
SH_D F3 xform_dir(const float* m, F3 p) {
    return F3{fmaf(m[0], p.x, fmaf(m[4], p.y, m[8] * p.z)), fmaf(m[1], p.x, fmaf(m[5], p.y, m[9] * p.z)),
              fmaf(m[2], p.x, fmaf(m[6], p.y, m[10] * p.z))};
}

/* push constants + UBO of one frame (rt.h:85-117) */
struct FrameParams {
    GpurtConstants c;
    GpurtCamera cam;
    uint32_t W, H, seed_val;
    /* multi-GPU sharding of the frame: this pipe renders the bands of `band_rows` rows whose
     * band index is congruent to `shard` modulo `n_shards`; (H, 1, 0) = the whole frame */
    uint32_t band_rows, n_shards, shard, n_local;
    /* extension (GpurtPipeParams): ReSTIR spatial reuse; 0 samples = the reference's estimator */
    uint32_t spatial_samples;
    float spatial_radius;
    /* extension (GpurtPipeParams::light_sampling): 0 = the reference's uniform light, uniform triangle; 1 = a light triangle
     * with probability proportional to area x luma(emissive), light_pdf weighted to match */
    uint32_t light_sampling;
};

/* i-th locally rendered pixel -> global pixel index y*W + x (RNG, image and G-buffers are indexed by
 * the global pixel, so results do not depend on how the frame is sharded) */
SH_D uint32_t shard_pixel(const FrameParams& P, uint32_t i) {
    uint32_t per_band = P.band_rows * P.W;
    uint32_t band = i / per_band, in_band = i - band * per_band;
    /* inside a band pixels are enumerated in 8x4 tiles, so the 32 camera rays a warp traces together
     * cover a compact screen region (better node / triangle sharing than a 32x1 scanline segment);
     * only the ORDER of the ray queue changes, never a pixel's result */
    if((P.W & 15u) == 0u && (P.band_rows & 7u) == 0u && (P.H & 7u) == 0u) {
        /* 128 consecutive paths (one CTA of the trace kernel) = a 16x8 pixel block of 2x2 warp tiles */
        uint32_t blk = in_band >> 7, t = in_band & 127u;
        uint32_t blocks_x = P.W >> 4;
        uint32_t by = blk / blocks_x, bx = blk - by * blocks_x;
        uint32_t w = t >> 5, l = t & 31u;
        in_band = (by * 8u + (w >> 1) * 4u + (l >> 3)) * P.W + bx * 16u + (w & 1u) * 8u + (l & 7u);
    } else if((P.W & 7u) == 0u && (P.band_rows & 3u) == 0u && (P.H & 3u) == 0u) {
        uint32_t tile = in_band >> 5, t = in_band & 31u;
        uint32_t tiles_x = P.W >> 3;
        uint32_t ty = tile / tiles_x, tx = tile - ty * tiles_x;
        in_band = (ty * 4u + (t >> 3)) * P.W + tx * 8u + (t & 7u);
    }
    return (band * P.n_shards + P.shard) * per_band + in_band;
}

/* j-th locally rendered row (bands in ascending order, rows inside a band in ascending order) -> image row */
SH_D uint32_t shard_row(const FrameParams& P, uint32_t j) {
    uint32_t band = j / P.band_rows;
    return (band * P.n_shards + P.shard) * P.band_rows + (j - band * P.band_rows);
}
/* ReSTIR on a sharded frame (gpurt_pipe_history_peers): which shards read row y of the previous frame's G-buffers and
 * reservoirs when their temporal / spatial look-ups stay within `halo` rows of the pixel being shaded — every shard that
 * owns a row of [y - halo, y + halo].  Bit s of the result = shard s; the owner's own bit is included.
 * halo = 0xFFFFFFFF: every shard (any camera motion).  n_shards <= 64. */
SH_D unsigned long long history_row_readers(const FrameParams& P, uint32_t y, uint32_t halo) {
    const unsigned long long all = P.n_shards >= 64u ? ~0ull : ((1ull << P.n_shards) - 1ull);
    if(halo == 0xFFFFFFFFu) return all;
    uint32_t y0 = y > halo ? y - halo : 0u, y1 = y + halo < P.H - 1u ? y + halo : P.H - 1u;
    uint32_t b0 = y0 / P.band_rows, b1 = y1 / P.band_rows;
    if(b1 - b0 + 1u >= P.n_shards) return all;
    unsigned long long m = 0;
    for(uint32_t b = b0; b <= b1; b++) m |= 1ull << (b % P.n_shards);
    return m;
}

struct Reservoir { /* restir.glsl:2-9; stored as 3 float4: pos|w_sum, normal|w, emissive|n_seen */
    F3 pos, normal, emissive;
    float w_sum, w;
    uint32_t n_seen;
};
struct TraceInfo {
    F3 o, d, acc;
    uint32_t depth;
    F3 throughput;
    float mis;
};
struct Payload {
    F3 bary;
    uint32_t obj_id, prim_id;
    bool hit;
};
struct HitInfo {
    F3 pos, normal, tangent;
    float tc[2];
};
struct MatInfo {
    F3 albedo, emissive, tanspaceNormal;
    float roughness;
    bool use_tanspace;
};
struct ShadeInfo {
    F3 wo, T, B, N;
};
struct LightSample {
    F3 pos, normal, emissive;
    float pdf;
};

SH_CONST float c_srgb_lut[256];

/* ---- light groups: an implicit two-level hierarchy over each light's triangles in index order ----------------
 * light_pdf (rt.rgen:222-255) sums triangle_pdf over EVERY triangle of every light whose box the ray hits; all but
 * the one or two triangles the ray actually crosses contribute +0.  The boxes below let it skip runs of 8 / 64
 * triangles the ray cannot touch.  The tests that remain run in the same ascending order and x + 0 == x for the
 * non-negative partial sums, so the result has the same bits as the full loop.  The boxes only have to be
 * conservative for the fp32 triangle test: they are padded by 2^-12 of the largest scene coordinate (128 x the BVH's
 * own N7 margin; the rounding of (lo - o) * (1/d) over a scene-sized distance is ~2^-21 of it). */
constexpr uint32_t kLightRun = 8;
constexpr float kLightPadScale = 128.0f; /* x accel inflation (= 2^-19 max|coord|) */

SH_D void light_run_box(const float4* tp, uint32_t count, float pad, float4& lo, float4& hi) {
    const float inf = GPURT_INF;
    F3 a = F3{inf, inf, inf}, b = F3{-inf, -inf, -inf};
    for(uint32_t t = 0; t < count; t++, tp += 3) {
        float4 r0 = tp[0], r1 = tp[1], r2 = tp[2];
        F3 p0 = F3{r0.x, r0.y, r0.z}, p1 = p0 + F3{r1.x, r1.y, r1.z}, p2 = p0 + F3{r2.x, r2.y, r2.z};
        a = F3{fminf(a.x, fminf(p0.x, fminf(p1.x, p2.x))), fminf(a.y, fminf(p0.y, fminf(p1.y, p2.y))),
               fminf(a.z, fminf(p0.z, fminf(p1.z, p2.z)))};
        b = F3{fmaxf(b.x, fmaxf(p0.x, fmaxf(p1.x, p2.x))), fmaxf(b.y, fmaxf(p0.y, fmaxf(p1.y, p2.y))),
               fmaxf(b.z, fmaxf(p0.z, fmaxf(p1.z, p2.z)))};
    }
    lo = make_float4(a.x - pad, a.y - pad, a.z - pad, 0.0f);
    hi = make_float4(b.x + pad, b.y + pad, b.z + pad, 0.0f);
}
/* box r of one light: r < n_runs64 -> triangles [64 r, 64 r + 64), else run g = r - n_runs64 -> [8 g, 8 g + 8) */
SH_D void light_box_record(const float4* light_tris, uint32_t n_tris, uint32_t r, float pad, float4* out2) {
    uint32_t ng = (n_tris + kLightRun - 1) / kLightRun, nsg = (ng + kLightRun - 1) / kLightRun;
    uint32_t span = r < nsg ? kLightRun * kLightRun : kLightRun, first = r < nsg ? r * span : (r - nsg) * span;
    uint32_t count = n_tris - first < span ? n_tris - first : span;
    light_run_box(light_tris + 3ull * first, count, pad, out2[0], out2[1]);
}
SH_D uint32_t light_box_records(uint32_t n_tris) {
    uint32_t ng = (n_tris + kLightRun - 1) / kLightRun;
    return ng + (ng + kLightRun - 1) / kLightRun;
}
/* does the ray o + t d, t >= 0, touch the (padded) box?  axes the ray is parallel to are decided by the origin */
SH_D bool ray_touches_box(F3 o, F3 d, F3 inv, float4 lo, float4 hi) {
    const float inf = GPURT_INF;
    bool px = fabsf(d.x) < 1e-20f, py = fabsf(d.y) < 1e-20f, pz = fabsf(d.z) < 1e-20f;
    float ax = px ? -inf : (lo.x - o.x) * inv.x, bx = px ? inf : (hi.x - o.x) * inv.x;
    float ay = py ? -inf : (lo.y - o.y) * inv.y, by = py ? inf : (hi.y - o.y) * inv.y;
    float az = pz ? -inf : (lo.z - o.z) * inv.z, bz = pz ? inf : (hi.z - o.z) * inv.z;
    float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
    float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    bool inside = (!px || (o.x >= lo.x && o.x <= hi.x)) && (!py || (o.y >= lo.y && o.y <= hi.y)) &&
                  (!pz || (o.z >= lo.z && o.z <= hi.z));
    return inside && tn <= tf;
}

/* everything a shading thread can see */
struct ShadeCtx {
    DeviceScene S;
    const float4* nodes;
    const float4* tris;
    const float4* tri_world; /* world-space triangles in global primitive order: v0 | e1 = v1-v0 | e2 = v2-v0 (k_flatten, N1) */
    unsigned n_nodes;
    /* light groups (render.cu k_light_groups): padded boxes over runs of 8 and 64 consecutive triangles of every
     * light, 2 float4 (lo, hi) per box; lgrp_off[l] = {first 64-run box, first 8-run box} of light l.  NULL: light_pdf
     * tests every triangle like the GLSL */
    const float4* lgrp;
    const uint2* lgrp_off;
    /* light BVH (render.cu pipe_light_accel): a second wide BVH over the triangles of the lights only, primitive id =
     * ltri_off[light] + triangle; n_lnodes == 0: light_pdf scans the light-run boxes instead */
    const float4* lnodes;
    const float4* ltris;
    const uint32_t* ltri_off;
    unsigned n_lnodes;
    /* world-space vertices of the light triangles (render.cu k_light_verts): 3 float4 per triangle = model * v0, v1, v2
     * exactly as light_sample computes them, light l starts at triangle lvert_off[l]; NULL: light_sample transforms
     * the three vertices itself like the GLSL */
    const float4* lverts;
    const uint32_t* lvert_off;
    /* power-proportional light sampling (GpurtPipeParams::light_sampling = 1): running sums of light_tri_power over the
     * light triangles in (light, triangle) order, light l starting at lcdf_off[l] (n_lights + 1 entries); NULL: off */
    const float* lcdf;
    const uint32_t* lcdf_off;
    unsigned n_ltris;
    const float4* prev_res;  /* previous frame reservoirs */
    const float4* ppos;      /* previous frame G-buffers */
    const float4* pnorm;
    const float4* palb;
    unsigned long long* ray_counts; /* [0] closest, [1] any */
};

/* weight of a light triangle for power-proportional sampling: area x luma of the object's emissive factor (an emissive
 * texture is not looked at: any positive weights give an unbiased estimator, the pdf carries them) */
GPURT_HD float light_tri_power(F3 v0, F3 v1, F3 v2, F3 emissive) { /* host + device: the pipe evaluates it on the host */
    F3 c = cross3(v1 - v0, v2 - v0);
    return 0.5f * sqrtf(dot3(c, c)) * (0.299f * emissive.x + 0.587f * emissive.y + 0.114f * emissive.z);
}

/* model * vec4(v, 1) of the three vertices of triangle t of object obj (rt.rgen:172-174) */
SH_D void light_world_tri(const DeviceScene& S, uint32_t obj, uint32_t t, float4* out3) {
    const uint32_t* ip = S.idx + 3ull * (S.tri_off[obj] + t);
    const float* m = reinterpret_cast<const float*>(S.descs + obj);
    for(int k = 0; k < 3; k++) {
        const float* v = reinterpret_cast<const float*>(S.verts + (S.vert_off[obj] + ip[k]));
        F3 w = xform_point(m, F3{v[0], v[1], v[2]});
        out3[k] = make_float4(w.x, w.y, w.z, 0.0f);
    }
}

struct Shader {
    const ShadeCtx& X;
    const FrameParams& P;
    uint32_t seed;
    Reservoir prev_res;
    unsigned n_closest, n_any;
    /* deferred shadow ray of integrate_direct (render.cu's shadow stage): instead of tracing `visibility` inline the
     * segment and the term it gates are handed back, and k_shadow_resolve finishes the pixel */
    uint32_t pix_x = 0, pix_y = 0; /* the pixel being shaded (spatial reuse looks around it) */
    bool defer_shadow = false, shadow_pending = false;
    F3 shadow_a, shadow_b, shadow_term;

    SH_D Shader(const ShadeCtx& x, const FrameParams& p) : X(x), P(p), seed(0), n_closest(0), n_any(0) {}

    /* rtcommon.glsl:111-124 */
    SH_D uint32_t lcg() {
        seed = 1664525u * seed + 1013904223u;
        return seed & 0x00FFFFFFu;
    }
    SH_D float randf() { return (float)lcg() / (float)0x01000000; }
    SH_D uint32_t randu(uint32_t a, uint32_t b) { return lcg() % (b - a) + a; }

    SH_D const float* model(uint32_t o) const { return reinterpret_cast<const float*>(X.S.descs + o); }
    SH_D const float* modelIT(uint32_t o) const { return reinterpret_cast<const float*>(X.S.descs + o) + 16; }
    SH_D F3 desc_vec(uint32_t o, int word) const {
        const float* f = reinterpret_cast<const float*>(X.S.descs + o) + word;
        return F3{f[0], f[1], f[2]};
    }
    SH_D int desc_int(uint32_t o, int word) const {
        return reinterpret_cast<const int*>(X.S.descs + o)[word];
    }
    SH_D const float* vertex(uint32_t obj, uint32_t i) const {
        return reinterpret_cast<const float*>(X.S.verts + (X.S.vert_off[obj] + i));
    }
    SH_D void tri_indices(uint32_t obj, uint32_t prim, uint32_t ind[3]) const {
        const uint32_t* p = X.S.idx + 3ull * (X.S.tri_off[obj] + prim);
        ind[0] = p[0], ind[1] = p[1], ind[2] = p[2];
    }
    SH_D void payload_from_hit(float u, float v, uint32_t gid, Payload& pl) const {
        uint32_t a = 0, b = X.S.n_objs;
        while(b - a > 1) {
            uint32_t m = (a + b) >> 1;
            if(X.S.tri_off[m] <= gid) a = m; else b = m;
        }
        pl.hit = true;
        pl.bary = F3{1.0f - u - v, u, v}; /* rt.rchit:13 */
        pl.obj_id = a;
        pl.prim_id = gid - X.S.tri_off[a];
    }

    /* texture(): R8G8B8A8_SRGB, linear, REPEAT, one mip (rt.cpp:439-446, vulkan.cpp:515-537) */
    SH_D F3 texture(int t, const float tc[2]) const {
        uint4 info = X.S.tex_info[t];
        int w = (int)info.y, h = (int)info.z;
        const uint8_t* base = X.S.texels + 4ull * info.x;
        float x = tc[0] * (float)w - 0.5f, y = tc[1] * (float)h - 0.5f;
        float fx0 = floorf(x), fy0 = floorf(y);
        float ax = x - fx0, ay = y - fy0;
        int x0 = (int)fx0, y0 = (int)fy0;
        int xa = x0 % w, xb = (x0 + 1) % w, ya = y0 % h, yb = (y0 + 1) % h;
        xa = xa < 0 ? xa + w : xa, xb = xb < 0 ? xb + w : xb, ya = ya < 0 ? ya + h : ya, yb = yb < 0 ? yb + h : yb;
        auto tex = [&](int xx, int yy) {
            const uint8_t* p = base + 4ull * ((size_t)yy * w + xx);
            return F3{c_srgb_lut[p[0]], c_srgb_lut[p[1]], c_srgb_lut[p[2]]};
        };
        F3 top = tex(xa, ya) * (1.0f - ax) + tex(xb, ya) * ax;
        F3 bot = tex(xa, yb) * (1.0f - ax) + tex(xb, yb) * ax;
        return top * (1.0f - ay) + bot * ay;
    }
    /* NEAREST fetch of a previous-frame G-buffer (rt.cpp:449-450) */
    SH_D F3 gbuf_fetch(const float4* img, float u, float v) const {
        int W = (int)P.W, H = (int)P.H;
        int x = (int)floorf(u * (float)W), y = (int)floorf(v * (float)H);
        x = ((x % W) + W) % W, y = ((y % H) + H) % H;
        float4 p = img[(size_t)y * W + x];
        return F3{p.x, p.y, p.z};
    }

    /* traceRayEXT closest, inline (rt.rgen:257-270) */
    SH_D void trace_ray(F3 o, F3 d, Payload& pl) {
        HitRec h;
        h.gid = kNoHit;
        n_closest++;
        if(X.n_nodes && traverse8<false, false>(X.nodes, X.tris, o, d, kEps, kLargeDist, h, nullptr))
            payload_from_hit(h.u, h.v, h.gid, pl);
        else
            pl.hit = false;
    }
    /* rt.rgen:272-291 */
    SH_D bool visibility(F3 a, F3 b) {
        F3 dir = b - a;
        float d = length3(dir);
        HitRec h;
        n_any++;
        return X.n_nodes && traverse8<true, false>(X.nodes, X.tris, a, dir / d, kEps, d - kEps, h, nullptr);
    }

    /* rt.rgen:62-95 */
    SH_D HitInfo hit_info(const Payload& pl) const {
        uint32_t obj = pl.obj_id;
        const float *mIT = modelIT(obj), *m = model(obj);
        F3 bary = pl.bary;
        uint32_t ind[3];
        tri_indices(obj, pl.prim_id, ind);
        const float *v0 = vertex(obj, ind[0]), *v1 = vertex(obj, ind[1]), *v2 = vertex(obj, ind[2]);
        HitInfo hit;
        F3 n = F3{v0[4], v0[5], v0[6]} * bary.x + F3{v1[4], v1[5], v1[6]} * bary.y + F3{v2[4], v2[5], v2[6]} * bary.z;
        hit.normal = normalize3(xform_dir(mIT, n));
        F3 t0 = F3{v0[8], v0[9], v0[10]} * v0[11], t1 = F3{v1[8], v1[9], v1[10]} * v1[11],
           t2 = F3{v2[8], v2[9], v2[10]} * v2[11];
        F3 t = t0 * bary.x + t1 * bary.y + t2 * bary.z;
        hit.tangent = normalize3(xform_dir(mIT, t));
        F3 p = F3{v0[0], v0[1], v0[2]} * bary.x + F3{v1[0], v1[1], v1[2]} * bary.y + F3{v2[0], v2[1], v2[2]} * bary.z;
        hit.pos = xform_point(m, p);
        hit.tc[0] = v0[3] * bary.x + v1[3] * bary.y + v2[3] * bary.z;
        hit.tc[1] = v0[7] * bary.x + v1[7] * bary.y + v2[7] * bary.z;
        return hit;
    }
    /* rt.rgen:97-130; Scene_Desc words: albedo 32, emissive 36, metal_rough 40, textures 44..47 */
    SH_D MatInfo mat_info(const Payload& pl, const HitInfo& hit) const {
        uint32_t obj = pl.obj_id;
        MatInfo mat;
        int albedoIdx = desc_int(obj, 44);
        mat.albedo = desc_vec(obj, 32);
        if(albedoIdx >= 0) mat.albedo = texture(albedoIdx, hit.tc);
        int emissiveIdx = desc_int(obj, 45);
        mat.emissive = desc_vec(obj, 36);
        if(emissiveIdx >= 0) mat.emissive = texture(emissiveIdx, hit.tc);
        int mrIdx = desc_int(obj, 46);
        F3 mr = desc_vec(obj, 40);
        if(mrIdx >= 0) mr = texture(mrIdx, hit.tc);
        mat.roughness = mr.y;
        if(P.c.use_metalness == 1) mat.albedo = mix3(f3s(0.04f), mat.albedo, mr.x);
        int nIdx = desc_int(obj, 47);
        mat.use_tanspace = nIdx >= 0;
        mat.tanspaceNormal = f3s(0.0f);
        if(mat.use_tanspace) mat.tanspaceNormal = texture(nIdx, hit.tc) * 2.0f - f3s(1.0f);
        return mat;
    }
    /* rtcommon.glsl:218-224 */
    SH_D static void make_tanspace(F3 N, F3& Nt, F3& Nb) {
        if(fabsf(N.x) > fabsf(N.y)) Nt = F3{N.z, 0.0f, -N.x} / sqrtf(N.x * N.x + N.z * N.z);
        else Nt = F3{0.0f, -N.z, N.y} / sqrtf(N.y * N.y + N.z * N.z);
        Nb = cross3(N, Nt);
    }
    /* rt.rgen:132-149 */
    SH_D ShadeInfo shade_info(F3 wo, HitInfo hit, const MatInfo& mat) const {
        ShadeInfo shade;
        shade.wo = wo;
        if(dot3(shade.wo, hit.normal) > 0) hit.normal = -hit.normal;
        shade.T = hit.tangent;
        shade.N = hit.normal;
        if(mat.use_tanspace && P.c.use_normal_map == 1) {
            shade.B = cross3(shade.N, shade.T);
            F3 tn = mat.tanspaceNormal;
            shade.N = normalize3(shade.T * tn.x + shade.B * tn.y + shade.N * tn.z);
        }
        make_tanspace(shade.N, shade.T, shade.B);
        return shade;
    }

    /* rtcommon.glsl:160-179 */
    SH_D F3 cospow_hemisphere(float exponent, F3 x, F3 y, F3 z) {
        float phi = (2 * kPiGlsl) * randf();
        float cosT = dm_pow(randf(), 1.0f / (exponent + 1.0f));
        float sinT = sqrtf(1.0f - cosT * cosT);
        F3 dir = F3{dm_cos(phi) * sinT, dm_sin(phi) * sinT, cosT};
        return dir.x * x + dir.y * y + dir.z * z;
    }
    SH_D F3 triangle_sample() {
        float u = sqrtf(randf());
        float v = randf();
        float a = u * (1 - v);
        float b = u * v;
        return F3{a, b, 1 - a - b};
    }

    /* rtcommon.glsl:255-369 */
    SH_D float bp_pdf(const MatInfo& mat, const ShadeInfo& sh, F3 wi) const {
        float oDn = dot3(-sh.wo, sh.N), iDn = dot3(wi, sh.N);
        if(oDn <= 0 || iDn <= 0) return 0;
        float ex = 1 / mat.roughness;
        F3 Hh = normalize3(wi - sh.wo);
        float cosine = fmaxf(dot3(Hh, sh.N), 0.0f);
        float N_pdf = (ex + 1) / (2 * kPiGlsl) * dm_pow(cosine, ex);
        return N_pdf / (4 * dot3(-sh.wo, Hh));
    }
    SH_D static F3 GGX_F(F3 r0, float iDn) {
        float cos5 = dm_pow(1 - iDn, 5.0f);
        return r0 + (f3s(1.0f) - r0) * cos5;
    }
    SH_D static float GGX_G(float oDn, float iDn, float a2) {
        float sqr0 = sqrtf(a2 + (1 - a2) * iDn * iDn);
        float sqr1 = sqrtf(a2 + (1 - a2) * oDn * oDn);
        return 2 * oDn * iDn / (oDn * sqr0 + iDn * sqr1);
    }
    SH_D static float GGX_D(float nDh, float a2) {
        float b = nDh * nDh * (a2 - 1) + 1;
        return a2 / (kPiGlsl * b * b);
    }
    SH_D float GGX_pdf(const MatInfo& mat, const ShadeInfo& sh, F3 wi) const {
        float oDn = dot3(-sh.wo, sh.N), iDn = dot3(wi, sh.N);
        if(oDn <= 0 || iDn <= 0) return 0;
        F3 Hh = normalize3(wi - sh.wo);
        float nDh = fmaxf(dot3(Hh, sh.N), 0.0f);
        float oDh = fmaxf(dot3(wi, Hh), 0.0f);
        float a2 = mat.roughness * mat.roughness;
        return GGX_D(nDh, a2) * nDh / (4 * oDh);
    }
    SH_D F3 GGX_eval(const MatInfo& mat, const ShadeInfo& sh, F3 wi) const {
        float oDn = dot3(-sh.wo, sh.N), iDn = dot3(wi, sh.N);
        if(oDn <= 0 || iDn <= 0) return f3s(0.0f);
        F3 Hh = normalize3(wi - sh.wo);
        float nDh = fmaxf(dot3(Hh, sh.N), 0.0f);
        float a2 = mat.roughness * mat.roughness;
        return GGX_F(mat.albedo, iDn) * GGX_D(nDh, a2) * GGX_G(oDn, iDn, a2) / (4 * oDn);
    }
    SH_D float MAT_pdf(const MatInfo& mat, const ShadeInfo& sh, F3 wi) const {
        if(P.c.brdf == 0) return bp_pdf(mat, sh, wi);
        if(P.c.brdf == 1) return GGX_pdf(mat, sh, wi);
        return 0;
    }
    SH_D F3 MAT_eval(const MatInfo& mat, const ShadeInfo& sh, F3 wi) const {
        if(P.c.brdf == 0) return mat.albedo * bp_pdf(mat, sh, wi);
        if(P.c.brdf == 1) return GGX_eval(mat, sh, wi);
        return f3s(0.0f);
    }
    SH_D bool MAT_sample(const MatInfo& mat, const ShadeInfo& sh, F3& wi) {
        if(P.c.brdf == 0) {
            float ex = 1 / mat.roughness;
            F3 Hh = cospow_hemisphere(ex, sh.T, sh.B, sh.N);
            wi = reflect3(sh.wo, Hh);
            return dot3(wi, sh.N) > 0;
        }
        if(P.c.brdf == 1) {
            float a2 = mat.roughness * mat.roughness;
            float Xi_x = randf(), Xi_y = randf();
            float phi = (2.0f * kPiGlsl) * Xi_x;
            float cosTheta = sqrtf((1.0f - Xi_y) / (1.0f + (a2 - 1.0f) * Xi_y));
            float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
            F3 dir = F3{dm_cos(phi) * sinTheta, dm_sin(phi) * sinTheta, cosTheta};
            F3 Hh = sh.T * dir.x + sh.B * dir.y + sh.N * dir.z;
            wi = reflect3(sh.wo, Hh);
            return dot3(wi, sh.N) > 0;
        }
        wi = f3s(0.0f);
        return false;
    }

    /* ---- lights ---- */
    /* rt.rgen:151-198 (Q4: no lights -> pdf 0, nothing drawn; Q5 texcoord quirk kept) */
    SH_D LightSample light_sample(F3 p) {
        LightSample s;
        if(P.c.n_lights <= 0) {
            s.pos = s.normal = s.emissive = f3s(0.0f);
            s.pdf = 0;
            return s;
        }
        uint32_t l_idx, t_idx;
        float prob = 0;
        const bool by_power = power_sampling();
        if(by_power) light_pick(l_idx, t_idx, prob);
        else l_idx = randu(0, (uint32_t)P.c.n_lights);
        uint32_t o_idx = X.S.lights[l_idx].index, n_tris = X.S.lights[l_idx].n_triangles;
        if(!by_power) t_idx = randu(0, n_tris);
        int emissiveIdx = desc_int(o_idx, 45);
        s.emissive = desc_vec(o_idx, 36);
        F3 _v0, _v1, _v2, bary;
        if(X.lverts && emissiveIdx < 0) { /* the three transformed vertices were computed once per build: one 48-byte read */
            const float4* q = X.lverts + 3ull * (X.lvert_off[l_idx] + t_idx);
            float4 a = GPURT_LDG(q), b = GPURT_LDG(q + 1), c = GPURT_LDG(q + 2);
            _v0 = F3{a.x, a.y, a.z}, _v1 = F3{b.x, b.y, b.z}, _v2 = F3{c.x, c.y, c.z};
            bary = triangle_sample();
            s.pos = _v0 * bary.x + _v1 * bary.y + _v2 * bary.z;
        } else {
            uint32_t ind[3];
            tri_indices(o_idx, t_idx, ind);
            const float *v0 = vertex(o_idx, ind[0]), *v1 = vertex(o_idx, ind[1]), *v2 = vertex(o_idx, ind[2]);
            const float* m = model(o_idx);
            _v0 = xform_point(m, F3{v0[0], v0[1], v0[2]}), _v1 = xform_point(m, F3{v1[0], v1[1], v1[2]}),
            _v2 = xform_point(m, F3{v2[0], v2[1], v2[2]});
            bary = triangle_sample();
            float tc[2] = {v0[3] * bary.x + v1[3] * bary.y + v2[3] * bary.z, v1[7] * bary.x + v1[7] * bary.y + v1[7] * bary.z};
            s.pos = _v0 * bary.x + _v1 * bary.y + _v2 * bary.z;
            if(emissiveIdx >= 0) s.emissive = texture(emissiveIdx, tc);
        }
        F3 Narea = cross3(_v1 - _v0, _v2 - _v0);
        float a = 2 / length3(Narea);
        F3 dist = s.pos - p;
        F3 N = normalize3(Narea);
        F3 d = normalize3(dist);
        float g = dot3(dist, dist) / fabsf(dot3(N, d));
        s.normal = N;
        s.pdf = by_power ? a * g * prob : a * g / (float)(n_tris * (uint32_t)P.c.n_lights);
        return s;
    }
    /* ---- extension: power-proportional choice of the light triangle (GpurtPipeParams::light_sampling = 1) ---- */
    SH_D bool power_sampling() const { return P.light_sampling == 1u && X.lcdf && X.n_ltris && X.lcdf[X.n_ltris - 1u] > 0; }
    /* probability of light triangle j (index in (light, triangle) order) */
    SH_D float light_tri_prob(uint32_t j) const {
        float prev = j ? X.lcdf[j - 1u] : 0.0f;
        return (X.lcdf[j] - prev) / X.lcdf[X.n_ltris - 1u];
    }
    /* one randf(): the first triangle whose running sum exceeds u x total (triangles of zero weight are never chosen) */
    SH_D void light_pick(uint32_t& l_idx, uint32_t& t_idx, float& prob) {
        float u = randf() * X.lcdf[X.n_ltris - 1u];
        uint32_t lo = 0, hi = X.n_ltris - 1u;
        while(lo < hi) {
            uint32_t mid = (lo + hi) >> 1;
            if(X.lcdf[mid] > u) hi = mid;
            else lo = mid + 1u;
        }
        prob = light_tri_prob(lo);
        l_idx = 0;
        while(l_idx + 1u < (uint32_t)P.c.n_lights && X.lcdf_off[l_idx + 1u] <= lo) l_idx++;
        t_idx = lo - X.lcdf_off[l_idx];
    }
    /* rt.rgen:200-220 */
    SH_D F3 light_sample_dir(F3 p) {
        uint32_t l_idx, t_idx;
        if(power_sampling()) {
            float prob;
            light_pick(l_idx, t_idx, prob);
        } else {
            l_idx = randu(0, (uint32_t)P.c.n_lights);
            t_idx = randu(0, X.S.lights[l_idx].n_triangles);
        }
        uint32_t o_idx = X.S.lights[l_idx].index;
        uint32_t ind[3];
        tri_indices(o_idx, t_idx, ind);
        const float *v0 = vertex(o_idx, ind[0]), *v1 = vertex(o_idx, ind[1]), *v2 = vertex(o_idx, ind[2]);
        F3 bary = triangle_sample();
        F3 point = F3{v0[0], v0[1], v0[2]} * bary.x + F3{v1[0], v1[1], v1[2]} * bary.y + F3{v2[0], v2[1], v2[2]} * bary.z;
        point = xform_point(model(o_idx), point);
        return normalize3(point - p);
    }
    /* rtcommon.glsl:181-214 */
    SH_D static bool triangle_hit(F3 o, F3 d, F3 pa, F3 pb, F3 pc, F3& hitp) {
        F3 v1 = pb - pa, v2 = pc - pa;
        F3 p = cross3(d, v2);
        float det = dot3(v1, p);
        if(fabsf(det) < kEps) return false;
        float invDet = 1 / det;
        F3 s = o - pa;
        float u = dot3(s, p) * invDet;
        if(u < 0 || u > 1) return false;
        F3 q = cross3(s, v1);
        float v = dot3(d, q) * invDet;
        if(v < 0 || u + v > 1) return false;
        float t = dot3(v2, q) * invDet;
        hitp = o + t * d;
        return t >= 0;
    }
    /* triangle_hit / triangle_pdf on a pre-flattened triangle (pa, e1 = pb - pa, e2 = pc - pa).  k_flatten evaluates
     * model * vec4(v, 1) and the two edge subtractions with exactly the operations of the GLSL (xform_point, then
     * `pb - pa`), so these return the same bits as the vertex forms below. */
    SH_D static float triangle_pdf_flat(F3 o, F3 d, F3 pa, F3 v1, F3 v2) {
        F3 p = cross3(d, v2);
        float det = dot3(v1, p);
        if(fabsf(det) < kEps) return 0;
        float invDet = 1 / det;
        F3 s = o - pa;
        float u = dot3(s, p) * invDet;
        if(u < 0 || u > 1) return 0;
        F3 q = cross3(s, v1);
        float v = dot3(d, q) * invDet;
        if(v < 0 || u + v > 1) return 0;
        float t = dot3(v2, q) * invDet;
        if(!(t >= 0)) return 0;
        F3 hitp = o + t * d;
        F3 c = cross3(v1, v2);
        float a = 2 / length3(c);
        F3 dist = hitp - o;
        F3 N = normalize3(c);
        float g = dot3(dist, dist) / fabsf(dot3(N, d));
        return a * g;
    }
    SH_D static float triangle_pdf(F3 o, F3 d, F3 v0, F3 v1, F3 v2) {
        F3 hitp;
        if(triangle_hit(o, d, v0, v1, v2, hitp)) {
            float a = 2 / length3(cross3(v1 - v0, v2 - v0));
            F3 dist = hitp - o;
            F3 N = normalize3(cross3(v1 - v0, v2 - v0));
            float g = dot3(dist, dist) / fabsf(dot3(N, d));
            return a * g;
        }
        return 0;
    }
    /* rtcommon.glsl:242-251 */
    SH_D static bool hit_bbox(F3 o, F3 d, F3 bmin, F3 bmax) {
        F3 invD = f3s(1.0f) / d;
        F3 t0 = (bmin - o) * invD, t1 = (bmax - o) * invD;
        F3 tNear = F3{fminf(t0.x, t1.x), fminf(t0.y, t1.y), fminf(t0.z, t1.z)};
        F3 tFar = F3{fmaxf(t0.x, t1.x), fmaxf(t0.y, t1.y), fmaxf(t0.z, t1.z)};
        float tNearMax = fmaxf(fmaxf(tNear.x, tNear.y), fmaxf(tNear.z, 0.0f));
        float tFarMin = fminf(fminf(tFar.x, tFar.y), tFar.z);
        return tNearMax <= tFarMin;
    }
    /* rt.rgen:222-255: every triangle of every light whose bbox the ray hits.
     * Written as ONE loop whose iteration is a single step of this lane's scan — move to the next light, skip a run of
     * 64 or 8 triangles whose box the ray misses (see light_run_box), or test one triangle — instead of the GLSL's loop
     * nest: in a nest, a warp whose lanes look at different runs executes every run's triangle loop with one or two
     * live lanes (ncu r01n: 70 % of the MIS shade kernel's instructions at 2.8 of 32 lanes); here the lanes that have a
     * triangle to test all test it in the same instruction stream.  Same tests, same order, same sums per lane. */
    SH_D float light_pdf_scan(F3 p, F3 d) const {
        const uint32_t n_lights = P.c.n_lights > 0 ? (uint32_t)P.c.n_lights : 0u;
        if(n_lights == 0) return 0;
        const F3 inv = f3s(1.0f) / d;
        const bool boxes = X.lgrp != nullptr;
        const bool by_power = power_sampling(); /* every term weighted by its triangle's probability instead of 1 / (n_tris n_lights) */
        float oacc = 0, tacc = 0;
        uint32_t l = 0, t = 0, n_tris = 0;
        bool in_light = false;
        const float4 *tp = nullptr, *b64 = nullptr, *b8 = nullptr;
        for(;;) {
            if(!in_light) { /* next light: `if(!hit_bbox(...)) continue;` */
                if(l >= n_lights) break;
                const SceneLight& L = X.S.lights[l];
                if(hit_bbox(p, d, F3{L.bmin[0], L.bmin[1], L.bmin[2]}, F3{L.bmax[0], L.bmax[1], L.bmax[2]})) {
                    /* the GLSL re-fetches 3 indices + 3 vertices and re-transforms them for every triangle of every
                     * call; the build already holds the same world-space triangles, contiguous per object */
                    tp = X.tri_world + 3ull * X.S.tri_off[L.index];
                    n_tris = L.n_triangles, t = 0, tacc = 0, in_light = true;
                    if(boxes) {
                        const uint2 off = X.lgrp_off[l];
                        b64 = X.lgrp + 2ull * off.x, b8 = X.lgrp + 2ull * off.y;
                    }
                } else
                    l++;
            } else if(t >= n_tris) { /* end of the triangle loop: `oacc += tacc / float(n_tris)` */
                oacc += by_power ? tacc : tacc / (float)n_tris;
                in_light = false, l++;
            } else {
                bool test = true;
                if(boxes && (t & (kLightRun * kLightRun - 1)) == 0) {
                    const float4* b = b64 + 2 * (t / (kLightRun * kLightRun));
                    if(!ray_touches_box(p, d, inv, GPURT_LDG(b), GPURT_LDG(b + 1))) t += kLightRun * kLightRun, test = false;
                }
                if(boxes && test && (t & (kLightRun - 1)) == 0) {
                    const float4* b = b8 + 2 * (t / kLightRun);
                    if(!ray_touches_box(p, d, inv, GPURT_LDG(b), GPURT_LDG(b + 1))) t += kLightRun, test = false;
                }
                if(test) {
                    const float4* q = tp + 3ull * t;
                    float4 r0 = GPURT_LDG(q), r1 = GPURT_LDG(q + 1), r2 = GPURT_LDG(q + 2);
                    float term = triangle_pdf_flat(p, d, F3{r0.x, r0.y, r0.z}, F3{r1.x, r1.y, r1.z}, F3{r2.x, r2.y, r2.z});
                    if(by_power && term != 0) term = term * light_tri_prob(X.lcdf_off[l] + t);
                    tacc += term;
                    t++;
                }
            }
        }
        return by_power ? oacc : oacc / (float)n_lights;
    }
    /* light_pdf through the light BVH.  Only the triangles the ray actually crosses contribute a non-zero term to the
     * GLSL's sums, and index-order runs make poor boxes (64 consecutive triangles of a sphere are one latitude ring: its
     * box is the whole disc).  So: collect every light triangle with a non-zero triangle_pdf along the ray [0, inf) with
     * the wide-BVH traversal core — any order —, sort the handful of hits by (light, triangle) and add them the way the
     * GLSL loop does: per light in ascending triangle order from 0, `oacc += tacc / n` only for lights whose hit_bbox
     * test passes.  Zero terms and lights without a hit add +0 to non-negative sums, so the bits are the GLSL's.  More
     * than kLightHits hits (a ray through a stack of lights): fall back to the scan. */
    static constexpr int kLightHits = 8;
    SH_D float light_pdf_bvh(F3 p, F3 d) const {
        unsigned hg[kLightHits];
        float hp[kLightHits];
        int nh = 0;
        {
            const RaySetup rs = make_ray_setup(p, d, 0.0f);
            const float tmax = GPURT_INF;
            uint2 stack[kStack];
            int sp = 0;
            uint2 ng;
            ng.x = 0u, ng.y = 0x80000000u;
            for(;;) { /* traverse8's node loop without a shrinking interval */
                if(ng.y <= 0x00ffffffu) {
                    if(sp == 0) break;
                    ng = stack[--sp];
                }
                unsigned hits = ng.y;
                unsigned bit = 31u - gpurt_clz(hits);
                ng.y &= ~(1u << bit);
                if(ng.y > 0x00ffffffu) stack[sp++] = ng;
                unsigned slot = (bit - 24u) ^ rs.octinv;
                unsigned rel = gpurt_popc(hits & 0xffu & ((1u << slot) - 1u));
                const float4* np = X.lnodes + (size_t)(ng.x + rel) * kNodeVec4;
                Node8 node;
#pragma unroll
                for(int k = 0; k < 5; k++) node.v[k] = GPURT_LDG(np + k);
                unsigned hit8 = node_hits8(node, rs, tmax);
                unsigned imask = f2u(node.v[0].w) >> 24;
                ng.x = f2u(node.v[1].x);
                ng.y = (octant_permute8(hit8 & imask, rs.octinv) << 24) | imask;
                unsigned leaf = hit8 & ~imask;
                const float4* tp = X.ltris + (size_t)f2u(node.v[1].y) * kTriVec4;
                unsigned m_lo = f2u(node.v[1].z), m_hi = f2u(node.v[1].w);
                while(leaf) {
                    unsigned ls = gpurt_ctz(leaf);
                    leaf &= leaf - 1u;
                    unsigned meta = ((ls & 4u ? m_hi : m_lo) >> (8u * (ls & 3u))) & 0xffu;
                    unsigned k = meta & 31u, kend = k + gpurt_popc(meta >> 5);
                    for(; k < kend; k++) {
                        float4 r0 = GPURT_LDG(tp + 3 * k), r1 = GPURT_LDG(tp + 3 * k + 1), r2 = GPURT_LDG(tp + 3 * k + 2);
                        float pdf = triangle_pdf_flat(p, d, F3{r0.x, r0.y, r0.z}, F3{r1.x, r1.y, r1.z}, F3{r2.x, r2.y, r2.z});
                        if(pdf != 0) {
                            if(nh == kLightHits) return light_pdf_scan(p, d);
                            unsigned gid = f2u(r0.w);
                            int i = nh++;
                            for(; i > 0 && hg[i - 1] > gid; i--) hg[i] = hg[i - 1], hp[i] = hp[i - 1];
                            hg[i] = gid, hp[i] = pdf;
                        }
                    }
                }
            }
        }
        const bool by_power = power_sampling(); /* hg = index in (light, triangle) order = index into lcdf */
        float oacc = 0;
        uint32_t l = 0;
        for(int i = 0; i < nh;) {
            while(X.ltri_off[l + 1] <= hg[i]) l++;
            const SceneLight& L = X.S.lights[l];
            const uint32_t end = X.ltri_off[l + 1];
            float tacc = 0;
            for(; i < nh && hg[i] < end; i++) tacc += by_power ? hp[i] * light_tri_prob(hg[i]) : hp[i];
            if(hit_bbox(p, d, F3{L.bmin[0], L.bmin[1], L.bmin[2]}, F3{L.bmax[0], L.bmax[1], L.bmax[2]}))
                oacc += by_power ? tacc : tacc / (float)L.n_triangles;
        }
        return by_power ? oacc : oacc / (float)P.c.n_lights;
    }
#if defined(GPURT_LIGHT_PDF_INLINE) /* A/B build: both call sites of integrate_mis get their own copy */
    SH_D
#else
    SH_D_CALL
#endif
    float light_pdf(F3 p, F3 d) const {
        if(P.c.n_lights <= 0) return 0;
        return X.n_lnodes ? light_pdf_bvh(p, d) : light_pdf_scan(p, d);
    }
    /* rt.rgen:293-301 */
    SH_D F3 direct_light(F3 o, F3 d) {
        Payload pl;
        trace_ray(o, d, pl);
        if(!pl.hit) return F3{P.c.env_light[0], P.c.env_light[1], P.c.env_light[2]};
        HitInfo hit = hit_info(pl);
        MatInfo mat = mat_info(pl, hit);
        return mat.emissive;
    }
    SH_D static float power_heuristic(float a, float b) { return a * a / (a * a + b * b); }
    SH_D static float luma(F3 rgb) { return 0.299f * rgb.x + 0.587f * rgb.y + 0.114f * rgb.z; }

    /* ---- integrators ---- */
    /* rt.rgen:303-353 */
    SH_D void integrate_mis(TraceInfo& trace, const HitInfo& hit, const MatInfo& mat, const ShadeInfo& shade) {
        if(any_gt0(mat.emissive)) {
            trace.acc = trace.acc + trace.throughput * trace.mis * mat.emissive;
            trace.depth = P.c.max_depth;
            return;
        }
        trace.o = hit.pos;
        if(mat.roughness == 0) {
            trace.d = reflect3(shade.wo, shade.N);
            trace.throughput = trace.throughput * mat.albedo;
            trace.mis = 1;
        } else {
            if(P.c.n_lights > 0) {
                F3 wi_light = light_sample_dir(hit.pos);
                float light_pdf_l = light_pdf(hit.pos, wi_light);
                if(light_pdf_l != 0) {
                    float light_pdf_m = MAT_pdf(mat, shade, wi_light);
                    F3 light_atten = MAT_eval(mat, shade, wi_light);
                    F3 weight = light_atten / light_pdf_l * power_heuristic(light_pdf_l, light_pdf_m);
                    trace.acc = trace.acc + trace.throughput * weight * direct_light(hit.pos, wi_light);
                }
            }
            F3 wi_brdf;
            if(!MAT_sample(mat, shade, wi_brdf)) {
                trace.depth = P.c.max_depth;
                return;
            }
            float brdf_pdf_m = MAT_pdf(mat, shade, wi_brdf);
            if(brdf_pdf_m != 0) {
                float brdf_pdf_l = light_pdf(hit.pos, wi_brdf);
                F3 brdf_atten = MAT_eval(mat, shade, wi_brdf);
                trace.throughput = trace.throughput * (brdf_atten / brdf_pdf_m);
                trace.mis = power_heuristic(brdf_pdf_m, brdf_pdf_l);
            } else {
                trace.depth = P.c.max_depth;
                return;
            }
            trace.d = wi_brdf;
        }
    }
    /* rt.rgen:355-389 */
    SH_D void integrate_mats(TraceInfo& trace, const HitInfo& hit, const MatInfo& mat, const ShadeInfo& shade) {
        if(any_gt0(mat.emissive)) {
            trace.acc = trace.acc + mat.emissive * trace.throughput;
            trace.depth = P.c.max_depth;
            return;
        }
        trace.o = hit.pos;
        if(mat.roughness == 0) {
            trace.d = reflect3(shade.wo, shade.N);
            trace.throughput = trace.throughput * mat.albedo;
        } else {
            F3 wi;
            if(!MAT_sample(mat, shade, wi)) {
                trace.depth = P.c.max_depth;
                return;
            }
            float pdf = MAT_pdf(mat, shade, wi);
            F3 atten = MAT_eval(mat, shade, wi);
            if(pdf != 0) trace.throughput = trace.throughput * (atten / pdf);
            else {
                trace.depth = P.c.max_depth;
                return;
            }
            trace.d = wi;
        }
    }
    /* rt.rgen:391-411 */
    SH_D void integrate_direct(TraceInfo& trace, const HitInfo& hit, const MatInfo& mat, const ShadeInfo& shade) {
        trace.depth = P.c.max_depth;
        if(any_gt0(mat.emissive)) {
            trace.acc = trace.acc + mat.emissive;
            return;
        }
        if(mat.roughness != 0) {
            LightSample light = light_sample(hit.pos);
            F3 wi = normalize3(light.pos - hit.pos);
            F3 light_atten = MAT_eval(mat, shade, wi);
            if(light.pdf != 0) {
                if(defer_shadow) { /* same arithmetic, finished by shadow_finish() once the segment has been traced */
                    shadow_pending = true, n_any++;
                    shadow_a = hit.pos, shadow_b = light.pos, shadow_term = light_atten / light.pdf * light.emissive;
                    return;
                }
                float shadow = visibility(hit.pos, light.pos) ? 0.0f : 1.0f;
                trace.acc = trace.acc + light_atten / light.pdf * light.emissive * shadow;
            }
        }
    }

    /* the last line of integrate_direct for a deferred shadow ray */
    SH_D static F3 shadow_finish(F3 acc, F3 term, bool occluded) { return acc + term * (occluded ? 0.0f : 1.0f); }

    /* restir.glsl:17-35 */
    SH_D void res_update(Reservoir& res, float weight, F3 pos, F3 normal, F3 emissive) {
        res.n_seen++;
        res.w_sum += weight;
        if(randf() < weight / res.w_sum) {
            res.pos = pos;
            res.normal = normal;
            res.emissive = emissive;
        }
    }
    SH_D static Reservoir res_new() {
        Reservoir r;
        r.pos = r.normal = r.emissive = F3{0.0f, 0.0f, 0.0f};
        r.w_sum = 0, r.w = 0, r.n_seen = 0;
        return r;
    }
    SH_D static Reservoir res_load(const float4* p) {
        float4 a = p[0], b = p[1], c = p[2];
        Reservoir r;
        r.pos = F3{a.x, a.y, a.z}, r.w_sum = a.w;
        r.normal = F3{b.x, b.y, b.z}, r.w = b.w;
        r.emissive = F3{c.x, c.y, c.z}, r.n_seen = f2u(c.w);
        return r;
    }
    SH_D static void res_store(float4* p, const Reservoir& r) {
        p[0] = make_float4(r.pos.x, r.pos.y, r.pos.z, r.w_sum);
        p[1] = make_float4(r.normal.x, r.normal.y, r.normal.z, r.w);
        p[2] = make_float4(r.emissive.x, r.emissive.y, r.emissive.z, u2f(r.n_seen));
    }
    /* rt.rgen:415-433 */
    SH_D float update_weight(Reservoir& res, const HitInfo& hit, const MatInfo& mat, const ShadeInfo& shade) const {
        if(res.n_seen == 0) {
            res.w = 0;
            return 0;
        }
        F3 dir = res.pos - hit.pos;
        F3 wi = normalize3(dir);
        F3 light_atten = MAT_eval(mat, shade, wi);
        float g = fabsf(dot3(res.normal, wi)) / dot3(dir, dir);
        F3 contrib = g * light_atten * res.emissive;
        float pHat = luma(contrib);
        res.w = (1 / pHat) * (res.w_sum / (float)res.n_seen);
        if(pHat == 0) res.w = 0;
        return pHat;
    }
    /* rt.rgen:435-505 */
    SH_D void reservoir_sample(TraceInfo& trace, const HitInfo& hit, const MatInfo& mat, const ShadeInfo& shade, bool first) {
        Reservoir new_res = res_new();
        if(P.c.n_lights > 0)
            for(uint32_t i = 0; i < P.cam.new_samples; i++) {
                LightSample light = light_sample(hit.pos);
                F3 wi = normalize3(light.pos - hit.pos);
                F3 light_atten = MAT_eval(mat, shade, wi);
                F3 contrib = light_atten * light.emissive / light.pdf;
                res_update(new_res, luma(contrib), light.pos, light.normal, light.emissive);
            }
        float new_pHat = update_weight(new_res, hit, mat, shade);
        if(new_pHat != 0 && visibility(hit.pos, new_res.pos)) new_res.w = 0;
        for(;;) {
            if(first && P.c.use_temporal == 1) {
                F4 pp = mul4(P.cam.prev_PV, hit.pos.x, hit.pos.y, hit.pos.z, 1.0f);
                pp.x /= pp.w, pp.y /= pp.w, pp.z /= pp.w;
                pp.x = (pp.x + 1.0f) * 0.5f, pp.y = (pp.y + 1.0f) * 0.5f;
                if(!((pp.x > 0 && pp.y > 0) && (pp.x < 1 && pp.y < 1))) break;
                F3 old_pos = gbuf_fetch(X.ppos, pp.x, pp.y);
                F3 old_norm = gbuf_fetch(X.pnorm, pp.x, pp.y);
                F3 old_alb = gbuf_fetch(X.palb, pp.x, pp.y);
                F3 posdiff = old_pos - hit.pos;
                if(dot3(posdiff, posdiff) > 0.01f) break;
                F3 albdiff = old_alb - mat.albedo;
                if(dot3(albdiff, albdiff) > 0.01f) break;
                if(dot3(old_norm, shade.N) < 0.5f) break;
                int fx = (int)(pp.x * (float)P.W), fy = (int)(pp.y * (float)P.H);
                prev_res = res_load(X.prev_res + 3ull * ((size_t)fy * P.W + fx));
            }
            Reservoir temporal_res = res_new();
            res_update(temporal_res, new_pHat * new_res.w * (float)new_res.n_seen, new_res.pos, new_res.normal, new_res.emissive);
            float old_pHat = update_weight(prev_res, hit, mat, shade);
            uint32_t cap = P.cam.temporal_multiplier * new_res.n_seen;
            prev_res.n_seen = cap < prev_res.n_seen ? cap : prev_res.n_seen;
            res_update(temporal_res, old_pHat * prev_res.w * (float)prev_res.n_seen, prev_res.pos, prev_res.normal, prev_res.emissive);
            temporal_res.n_seen = new_res.n_seen + prev_res.n_seen;
            update_weight(temporal_res, hit, mat, shade);
            new_res = temporal_res;
            break;
        }
        if(first && P.spatial_samples > 0) spatial_reuse(new_res, hit, mat, shade);
        if(new_res.w != 0) {
            F3 dir = new_res.pos - hit.pos;
            F3 wi = normalize3(dir);
            F3 light_atten = MAT_eval(mat, shade, wi);
            F3 contrib = light_atten * new_res.emissive;
            float g = fabsf(dot3(new_res.normal, wi)) / dot3(dir, dir);
            trace.acc = trace.acc + new_res.w * contrib * g;
        }
        prev_res = new_res;
    }
    /* Extension (include/gpurt.h GpurtPipeParams::spatial_samples; no reference counterpart): combine the pixel's reservoir
     * with reservoirs of the previous frame from a disc around the pixel.  Neighbours are treated the way the temporal
     * step treats prev_res (rt.rgen:489-497): their weight is re-derived at this shading point by update_weight, their
     * history is capped at temporal_multiplier x the new samples, and they enter with pHat * W * M.  The survivor is
     * tested for visibility once. */
    SH_D void spatial_reuse(Reservoir& cur, const HitInfo& hit, const MatInfo& mat, const ShadeInfo& shade) {
        Reservoir comb = res_new();
        float cur_pHat = update_weight(cur, hit, mat, shade);
        res_update(comb, cur_pHat * cur.w * (float)cur.n_seen, cur.pos, cur.normal, cur.emissive);
        uint32_t total = cur.n_seen;
        const uint32_t cap = P.cam.temporal_multiplier * P.cam.new_samples;
        for(uint32_t i = 0; i < P.spatial_samples; i++) {
            float r = P.spatial_radius * sqrtf(randf());
            float phi = 2.0f * kPiGlsl * randf();
            int qx = (int)pix_x + (int)floorf(r * dm_cos(phi) + 0.5f), qy = (int)pix_y + (int)floorf(r * dm_sin(phi) + 0.5f);
            if(qx < 0 || qy < 0 || qx >= (int)P.W || qy >= (int)P.H) continue;
            size_t q = (size_t)qy * P.W + (size_t)qx;
            float4 np = X.ppos[q], nn = X.pnorm[q];
            F3 d = F3{np.x, np.y, np.z} - hit.pos;
            if(!(dot3(F3{nn.x, nn.y, nn.z}, shade.N) >= 0.9f)) continue;                    /* same orientation */
            if(!(fabsf(dot3(d, shade.N)) <= 0.05f * sqrtf(dot3(d, d)) + 1.0e-3f)) continue; /* same plane */
            Reservoir nb = res_load(X.prev_res + 3ull * q);
            if(nb.n_seen == 0) continue;
            float nb_pHat = update_weight(nb, hit, mat, shade);
            nb.n_seen = cap < nb.n_seen ? cap : nb.n_seen;
            res_update(comb, nb_pHat * nb.w * (float)nb.n_seen, nb.pos, nb.normal, nb.emissive);
            total += nb.n_seen;
        }
        comb.n_seen = total;
        float pHat = update_weight(comb, hit, mat, shade);
        if(pHat != 0 && comb.w != 0 && visibility(hit.pos, comb.pos)) comb.w = 0;
        cur = comb;
    }
    /* rt.rgen:507-549 */
    SH_D void integrate_restir(TraceInfo& trace, const HitInfo& hit, const MatInfo& mat, const ShadeInfo& shade, bool d_only, bool first) {
        if(any_gt0(mat.emissive)) {
            trace.acc = trace.acc + mat.emissive * trace.throughput * trace.mis;
            trace.depth = P.c.max_depth;
            return;
        }
        trace.o = hit.pos;
        if(mat.roughness == 0) {
            trace.d = reflect3(shade.wo, shade.N);
            trace.throughput = trace.throughput * mat.albedo;
            trace.mis = 1;
        } else {
            if(trace.depth == 0) reservoir_sample(trace, hit, mat, shade, first);
            F3 wi_brdf;
            if(!MAT_sample(mat, shade, wi_brdf)) {
                trace.depth = P.c.max_depth;
                return;
            }
            float brdf_pdf = MAT_pdf(mat, shade, wi_brdf);
            if(brdf_pdf != 0) {
                F3 brdf_atten = MAT_eval(mat, shade, wi_brdf);
                trace.throughput = trace.throughput * (brdf_atten / brdf_pdf);
                trace.mis = trace.depth == 0 ? 0.0f : 1.0f;
            } else {
                trace.depth = P.c.max_depth;
                return;
            }
            trace.d = wi_brdf;
        }
        if(d_only) trace.depth = P.c.max_depth;
    }

    /* rt.rgen:551-565 */
    SH_D F3 make_camera_ray(uint32_t s, uint32_t px, uint32_t py) {
        float jx, jy;
        if(P.c.qmc == 0) {
            if(P.c.frame == 0) jx = jy = 0.5f;
            else {
                jx = randf();
                jy = randf();
            }
        } else {
            uint32_t i = s + (uint32_t)(P.c.samples * P.c.frame), N = (uint32_t)(P.c.samples * P.c.max_frame);
            jx = (float)i / (float)N;
            uint32_t bits = i; /* rtcommon.glsl:128-135 */
            bits = (bits << 16u) | (bits >> 16u);
            bits = ((bits & 0x55555555u) << 1u) | ((bits & 0xAAAAAAAAu) >> 1u);
            bits = ((bits & 0x33333333u) << 2u) | ((bits & 0xCCCCCCCCu) >> 2u);
            bits = ((bits & 0x0F0F0F0Fu) << 4u) | ((bits & 0xF0F0F0F0u) >> 4u);
            bits = ((bits & 0x00FF00FFu) << 8u) | ((bits & 0xFF00FF00u) >> 8u);
            jy = (float)bits * 2.3283064365386963e-10f;
        }
        float pcx = (float)px + jx, pcy = (float)py + jy;
        float ux = pcx / (float)P.W, uy = pcy / (float)P.H;
        F4 target = mul4(P.cam.iP, ux * 2.0f - 1.0f, uy * 2.0f - 1.0f, 0.0f, 1.0f);
        F4 direction = mul4(P.cam.iV, target.x, target.y, target.z, 0.0f);
        return normalize3(F3{direction.x, direction.y, direction.z});
    }
};

/* ---- per-pixel bodies of the frame kernels (render.cu launches them; tests/emu replays them) ---------------- */

/* k_frame_begin: tea(pixel, seed) — rtcommon.glsl:99-109; Q1: seed = user seed ^ frame replaces clockARB() */
SH_D void pixel_begin(const FrameParams& P, int restir, uint32_t i, float4* acc, float4* pathB, float4* gpos,
                      float4* gnorm, float4* galb, float4* res_out) {
    uint32_t v0 = i, v1 = P.seed_val, s0 = 0;
    for(uint32_t k = 0; k < 16; k++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    acc[i] = make_float4(0, 0, 0, 0);
    pathB[i] = make_float4(1, 1, 1, u2f(v0));
    gpos[i] = gnorm[i] = galb[i] = make_float4(0, 0, 0, 1); /* rt.rgen:573, :674-676 */
    if(restir) {
        res_out[3ull * i] = res_out[3ull * i + 1] = make_float4(0, 0, 0, 0);
        res_out[3ull * i + 2] = make_float4(0, 0, 0, u2f(0u));
    }
}

/* k_begin_camera: pixel_begin + pixel_gen_camera(s = 0) in one pass — the seed goes from tea() straight into the camera
 * ray instead of through pathB, and when every pixel is shaded at (s = 0, depth = 0) the G-buffer clear is left to
 * shade_step (a hit overwrites it anyway; a miss writes the cleared values there): ~96 bytes per pixel less than the two
 * kernels.  clear_gbuf: nothing will be shaded (max_depth 0), clear here. */
SH_D void pixel_begin_camera(const FrameParams& P, int restir, int clear_gbuf, uint32_t i, uint32_t li, float4* acc,
                             float4* pathA, float4* pathB, float4* gpos, float4* gnorm, float4* galb, float4* res_out,
                             float4* rays, uint32_t* queue) {
    uint32_t v0 = i, v1 = P.seed_val, s0 = 0;
    for(uint32_t k = 0; k < 16; k++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    acc[i] = make_float4(0, 0, 0, 0);
    if(clear_gbuf) gpos[i] = gnorm[i] = galb[i] = make_float4(0, 0, 0, 1);
    if(restir) {
        res_out[3ull * i] = res_out[3ull * i + 1] = make_float4(0, 0, 0, 0);
        res_out[3ull * i + 2] = make_float4(0, 0, 0, u2f(0u));
    }
    ShadeCtx dummy{};
    Shader sh(dummy, P);
    sh.seed = v0;
    F3 d = sh.make_camera_ray(0, i % P.W, i / P.W);
    F4 co = mul4(P.cam.iV, 0.0f, 0.0f, 0.0f, 1.0f); /* rt.rgen:572 */
    rays[2ull * li] = make_float4(co.x, co.y, co.z, kEps);
    rays[2ull * li + 1] = make_float4(d.x, d.y, d.z, kLargeDist);
    queue[li] = i;
    pathA[i] = make_float4(0, 0, 0, 1.0f);
    pathB[i] = make_float4(1.0f, 1.0f, 1.0f, u2f(sh.seed));
}

/* k_gen_camera: sample s of pixel i -> ray slot li of queue 0 (rt.rgen:551-565, :572, :579-585) */
SH_D void pixel_gen_camera(const FrameParams& P, uint32_t s, uint32_t i, uint32_t li, float4* pathA, float4* pathB,
                           float4* rays, uint32_t* queue) {
    ShadeCtx dummy{};
    Shader sh(dummy, P);
    float4 B = pathB[i];
    sh.seed = f2u(B.w);
    F3 d = sh.make_camera_ray(s, i % P.W, i / P.W);
    F4 co = mul4(P.cam.iV, 0.0f, 0.0f, 0.0f, 1.0f); /* rt.rgen:572 */
    rays[2ull * li] = make_float4(co.x, co.y, co.z, kEps);
    rays[2ull * li + 1] = make_float4(d.x, d.y, d.z, kLargeDist);
    queue[li] = i;
    pathA[i] = make_float4(0, 0, 0, 1.0f);                /* trace.acc, trace.mis */
    pathB[i] = make_float4(1.0f, 1.0f, 1.0f, u2f(sh.seed)); /* trace.throughput, rng */
}

/* One iteration of rt.rgen's bounce loop body after traceRayEXT (rt.rgen:591-627) for the path of pixel
 * `pix`: miss handling, hit_info / mat_info / shade_info, G-buffer capture, the selected integrator and
 * Russian roulette.  Returns true when the path ends here (`break` in the shader). */
/* INTEG = the integrator, fixed at compile time: one kernel per integrator instead of one kernel carrying all five
 * (the five-way kernel needs 128 registers -> 23 % occupancy; ncu capture prof_frame_r1k) */
template <int INTEG>
SH_D bool shade_step(const FrameParams& P, Shader& sh, TraceInfo& trace, uint32_t s, uint32_t depth, uint32_t pix,
                     float4 h, float4* gpos, float4* gnorm, float4* galb, float4* res_cur) {
    const bool restir = INTEG == 3 || INTEG == 4;
    bool broke = false;
    uint32_t gid = f2u(h.w);
    if(gid == kNoHit) { /* rt.rgen:591-598 */
        /* the G-buffer of a pixel whose first ray misses keeps its cleared value (rt.rgen:573); written here because
         * k_begin_camera leaves the clear to the first shading pass */
        if(s == 0 && depth == 0) gpos[pix] = gnorm[pix] = galb[pix] = make_float4(0, 0, 0, 1);
        if(depth == 0) trace.acc = F3{P.c.clear_col[0], P.c.clear_col[1], P.c.clear_col[2]};
        else trace.acc = trace.acc + F3{P.c.env_light[0], P.c.env_light[1], P.c.env_light[2]} * trace.throughput;
        return true;
    }
    Payload pl;
    sh.payload_from_hit(h.y, h.z, gid, pl);
    HitInfo hit = sh.hit_info(pl);
    MatInfo mat = sh.mat_info(pl, hit);
    ShadeInfo shade = sh.shade_info(trace.d, hit, mat);
    if(s == 0 && depth == 0) { /* rt.rgen:604-608 */
        gpos[pix] = make_float4(hit.pos.x, hit.pos.y, hit.pos.z, 1.0f);
        gnorm[pix] = make_float4(shade.N.x, shade.N.y, shade.N.z, 1.0f);
        galb[pix] = make_float4(mat.albedo.x, mat.albedo.y, mat.albedo.z, 1.0f);
    }
    if(restir && depth == 0) sh.prev_res = Shader::res_load(res_cur + 3ull * pix), sh.pix_x = pix % P.W, sh.pix_y = pix / P.W;
    if(INTEG == 0) sh.integrate_direct(trace, hit, mat, shade);
    else if(INTEG == 1) sh.integrate_mats(trace, hit, mat, shade);
    else if(INTEG == 2) sh.integrate_mis(trace, hit, mat, shade);
    else if(INTEG == 3) sh.integrate_restir(trace, hit, mat, shade, true, s == 0);
    else if(INTEG == 4) sh.integrate_restir(trace, hit, mat, shade, false, s == 0);
    if(restir && depth == 0) Shader::res_store(res_cur + 3ull * pix, sh.prev_res);
    if(P.c.use_rr == 1) { /* rt.rgen:622-627 */
        float pcont = fminf(fmaxf(trace.throughput.x, fmaxf(trace.throughput.y, trace.throughput.z)) + 0.001f, 0.95f);
        if(sh.randf() >= pcont) broke = true;
        else trace.throughput = trace.throughput / pcont;
    }
    return broke;
}

/* k_tail's per-path loop: the remaining bounces of pixel `pix` from depth0 on, traced inline; returns the number of
 * closest-hit rays it traced for the bounces themselves (shadow / light rays are counted inside `sh`) */
template <int INTEG>
SH_D unsigned path_tail(const FrameParams& P, const ShadeCtx& X, Shader& sh, uint32_t s, uint32_t depth0, uint32_t pix,
                        float4 r0, float4 r1, float4* pathA, float4* pathB, float4* acc, float4* gpos, float4* gnorm,
                        float4* galb, float4* res_cur) {
    unsigned n_wave = 0;
    float4 A = pathA[pix], B = pathB[pix];
    TraceInfo trace;
    trace.o = F3{r0.x, r0.y, r0.z}, trace.d = F3{r1.x, r1.y, r1.z};
    trace.acc = F3{A.x, A.y, A.z}, trace.mis = A.w;
    trace.throughput = F3{B.x, B.y, B.z};
    sh.seed = f2u(B.w);
    for(uint32_t depth = depth0;; depth++) {
        trace.depth = depth;
        HitRec hr;
        hr.gid = kNoHit, hr.t = 0, hr.u = hr.v = 0;
        n_wave++;
        if(X.n_nodes) traverse8<false, false>(X.nodes, X.tris, trace.o, trace.d, kEps, kLargeDist, hr, nullptr);
        float4 h = make_float4(hr.t, hr.u, hr.v, u2f(hr.gid));
        bool broke = shade_step<INTEG>(P, sh, trace, s, depth, pix, h, gpos, gnorm, galb, res_cur);
        if(broke || trace.depth + 1 >= (uint32_t)P.c.max_depth) break;
    }
    float4 a = acc[pix]; /* rt.rgen:630: acc += trace.acc; the RNG stream continues into the next sample */
    acc[pix] = make_float4(a.x + trace.acc.x, a.y + trace.acc.y, a.z + trace.acc.z, 0.0f);
    pathB[pix] = make_float4(1.0f, 1.0f, 1.0f, u2f(sh.seed));
    return n_wave;
}

/* rt.rgen:640-645: progressive mean over frames */
SH_D float4 accumulate_frame(float4 old, F3 avg, int frame) {
    if(frame > 0) {
        F3 m = mix3(F3{old.x, old.y, old.z}, avg, 1.0f / (float)(frame + 1));
        return make_float4(m.x, m.y, m.z, 1.0f);
    }
    return make_float4(avg.x, avg.y, avg.z, 1.0f);
}

/* k_frame_end: per-frame mean, progressive accumulation, debug views (rt.rgen:638-672) */
SH_D void pixel_end(const FrameParams& P, uint32_t i, const float4* acc, float4* image, const float4* gpos,
                    const float4* gnorm, const float4* ppos, const float4* pnorm, const float4* palb, float4* mean_out) {
    float4 a = acc[i];
    F3 avg = F3{a.x, a.y, a.z} / (float)P.c.samples; /* rt.rgen:638 */
    if(mean_out) { /* frame-parallel sharding: hand the frame mean to the accumulating rank, leave the image alone */
        mean_out[i] = make_float4(avg.x, avg.y, avg.z, 1.0f);
        return;
    }
    float4 out;
    out = accumulate_frame(P.c.frame > 0 ? image[i] : make_float4(0, 0, 0, 0), avg, P.c.frame);
    if(P.c.debug_view > 0) { /* rt.rgen:647-672 */
        float4 gp = gpos[i], gn = gnorm[i];
        F4 pp = mul4(P.cam.prev_PV, gp.x, gp.y, gp.z, 1.0f);
        pp.x /= pp.w, pp.y /= pp.w, pp.z /= pp.w;
        pp.x = (pp.x + 1.0f) * 0.5f, pp.y = (pp.y + 1.0f) * 0.5f;
        F3 n = F3{gn.x, gn.y, gn.z};
        if(dot3(n, n) > 0.5f && (pp.x > 0 && pp.y > 0) && (pp.x < 1 && pp.y < 1)) {
            int W = (int)P.W, H = (int)P.H;
            int x = (int)floorf(pp.x * (float)W), y = (int)floorf(pp.y * (float)H);
            x = ((x % W) + W) % W, y = ((y % H) + H) % H;
            const float4* img = P.c.debug_view == 1 ? ppos : P.c.debug_view == 2 ? pnorm : palb;
            float4 v = img[(size_t)y * W + x];
            if(P.c.debug_view <= 3) out = make_float4(v.x, v.y, v.z, 1.0f);
        } else
            out = make_float4(0, 0, 0, 1.0f);
    }
    image[i] = out;
}

} // namespace gpurt
