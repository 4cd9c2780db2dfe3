/*
 * oracle_render.cpp — CPU restatement of the reference integrator (TEST INFRASTRUCTURE, NOT
 * PRODUCT).  One sequential invocation per pixel, following src/shaders/rt/rt.rgen,
 * rtcommon.glsl and restir.glsl function by function (line citations at each function; paths are
 * relative to the reference root).  traceRayEXT is served by the oracle BVH (oracle.cpp), which is
 * itself pinned against brute force.
 *
 * fp32 contract N8: vector ops are component-wise single-rounded ops evaluated in
 * GLSL source order; dot / cross / mat*vec use the fma forms below; normalize(v) = v / sqrt(dot);
 * sin / cos / pow / exp come from include/gpurt_detmath.h; no other contraction
 * (-ffp-contract=off).  Documented deviations from the GLSL: Q1 seed, Q4 no-light
 * guard.
 *
 * PINNED: the reference's own rt.rgen / rtcommon.glsl / restir.glsl / tonemap.frag, compiled as C++ under the same
 * contract (oracle/_ref/libglsl_ref.so, oracle/ref_shim/glsl_*.{h,cpp}), produce the same bits as this file on 16
 * functions x 400 vectors and on whole frames of every integrator (tests/test_oracle.py, tests/golden/).
 */
#include "oracle_render.h"

#include <cmath>
#include <cstring>
#include <atomic>
#include <thread>
#include <vector>

#include "../include/gpurt_detmath.h"

extern "C" int orc_bvh_trace_one(const orc_bvh* B, const float* ray8, uint32_t* hit4);
extern "C" int orc_bvh_occluded_one(const orc_bvh* B, const float* ray8);

namespace {

const float M_PI_F = 3.141592f;      /* rtcommon.glsl:6 */
const float LARGE_DIST = 10000000.0f; /* rtcommon.glsl:7 */
const float EPS = 0.00001f;          /* rtcommon.glsl:8 */

struct vec3 {
    float x, y, z;
};
inline vec3 v3(float s) { return {s, s, s}; }
inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator-(vec3 a) { return {-a.x, -a.y, -a.z}; }
inline vec3 operator*(vec3 a, vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(float s, vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline vec3 operator/(vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline vec3 operator/(vec3 a, vec3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline float dot(vec3 a, vec3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
inline vec3 cross(vec3 a, vec3 b) {
    return {fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x))};
}
inline float length(vec3 a) { return sqrtf(dot(a, a)); }
inline vec3 normalize(vec3 a) { return a / length(a); }
inline vec3 reflect(vec3 I, vec3 N) { return I - (2.0f * dot(N, I)) * N; }
inline vec3 mix(vec3 a, vec3 b, float t) { return a * (1.0f - t) + b * t; }
inline bool any_gt0(vec3 a) { return a.x > 0 || a.y > 0 || a.z > 0; }
inline float u2f(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}
inline uint32_t f2u(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}

struct vec4 {
    float x, y, z, w;
};
/* column-major mat4 * vec4: r_i = fma(m[i],x, fma(m[4+i],y, fma(m[8+i],z, m[12+i]*w))) */
inline vec4 mul4(const float* m, float x, float y, float z, float w) {
    vec4 r;
    r.x = fmaf(m[0], x, fmaf(m[4], y, fmaf(m[8], z, m[12] * w)));
    r.y = fmaf(m[1], x, fmaf(m[5], y, fmaf(m[9], z, m[13] * w)));
    r.z = fmaf(m[2], x, fmaf(m[6], y, fmaf(m[10], z, m[14] * w)));
    r.w = fmaf(m[3], x, fmaf(m[7], y, fmaf(m[11], z, m[15] * w)));
    return r;
}
inline vec3 xform_point(const float* m, vec3 p) { /* vec3(m * vec4(p,1)) */
    return {fmaf(m[0], p.x, fmaf(m[4], p.y, fmaf(m[8], p.z, m[12]))),
            fmaf(m[1], p.x, fmaf(m[5], p.y, fmaf(m[9], p.z, m[13]))),
            fmaf(m[2], p.x, fmaf(m[6], p.y, fmaf(m[10], p.z, m[14])))};
}
inline vec3 xform_dir(const float* m, vec3 p) { /* vec3(m * vec4(p,0)) */
    return {fmaf(m[0], p.x, fmaf(m[4], p.y, m[8] * p.z)), fmaf(m[1], p.x, fmaf(m[5], p.y, m[9] * p.z)),
            fmaf(m[2], p.x, fmaf(m[6], p.y, m[10] * p.z))};
}

/* GpurtConstants (rt.h:85-102) */
struct Consts {
    float clear_col[4], env_light[4];
    int frame, samples, max_frame, qmc, max_depth, use_normal_map, use_metalness, use_temporal, integrator,
        brdf, debug_view, use_rr, n_lights, n_objs;
};
/* GpurtCamera (rt.h:104-117) */
struct Camera {
    float V[16], P[16], iV[16], iP[16], prev_PV[16];
    uint32_t new_samples, temporal_multiplier;
};
static_assert(sizeof(Consts) == 88 && sizeof(Camera) == 328, "layouts");

struct Reservoir { /* restir.glsl:2-9 */
    vec3 pos, normal, emissive;
    float w_sum, w;
    uint32_t n_seen;
};
struct TraceInfo { /* rtcommon.glsl:46-53 */
    vec3 o, d, acc;
    uint32_t depth;
    vec3 throughput;
    float mis;
};
struct Payload { /* rtcommon.glsl:55-60 */
    vec3 barycentrics;
    uint32_t obj_id, prim_id;
    bool hit;
};
struct HitInfo {
    vec3 pos, normal, tangent;
    float tc[2];
};
struct MatInfo {
    vec3 albedo, emissive, tanspaceNormal;
    float roughness;
    bool use_tanspace;
};
struct ShadeInfo {
    vec3 wo, T, B, N;
};
struct LightSample { /* rtcommon.glsl:30-38 */
    uint32_t l_idx, o_idx, t_idx;
    vec3 pos, normal, emissive;
    float pdf;
};

float srgb_lut[256];
struct LutInit {
    LutInit() {
        for(int i = 0; i < 256; i++) {
            double c = i / 255.0;
            srgb_lut[i] = (float)(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
        }
    }
} lut_init;

} // namespace

struct orc_scene {
    uint32_t n_objs, n_lights, n_tex;
    const uint32_t *descs, *tri_off, *vert_off, *idx, *lights, *tex_info;
    const float* verts;
    const uint8_t* texels;
    const orc_bvh* bvh;
};

namespace {

/* per-invocation state of rt.rgen (its globals: seed, payload, prev_res) */
/* optional recorder of every closest-hit ray a frame traces (bench.py's CPU arm replays them) */
static float* g_rec_buf = nullptr;
static uint64_t g_rec_cap = 0;
static std::atomic<uint64_t> g_rec_count{0};

struct Invocation {
    const orc_scene* S;
    const Consts* c;
    const Camera* cam;
    uint32_t W, H, px, py;
    uint32_t seed;
    Payload payload;
    Reservoir prev_res;
    const uint32_t* prev_res_buf;
    const float *ppos, *pnorm, *palb;
    uint32_t spatial_samples = 0; /* extension, see orc_render_set_spatial */
    float spatial_radius = 16.0f;
    /* extension, see orc_render_set_light_sampling: running sums of the light triangles' weights in (light, triangle)
     * order and the first index of every light; NULL = rt.rgen as written */
    const float* lcdf = nullptr;
    const uint32_t* lcdf_off = nullptr;
    uint32_t n_ltris = 0;
    bool power_sampling() const { return lcdf && n_ltris && lcdf[n_ltris - 1] > 0; }
    float light_tri_prob(uint32_t j) const {
        float prev = j ? lcdf[j - 1] : 0.0f;
        return (lcdf[j] - prev) / lcdf[n_ltris - 1];
    }
    void light_pick(uint32_t& l_idx, uint32_t& t_idx, float& prob) {
        float u = randf() * lcdf[n_ltris - 1];
        uint32_t j = 0; /* the first triangle whose running sum exceeds u (linear here, a binary search in the product) */
        while(j + 1 < n_ltris && !(lcdf[j] > u)) j++;
        prob = light_tri_prob(j);
        l_idx = 0;
        while(l_idx + 1 < (uint32_t)c->n_lights && lcdf_off[l_idx + 1] <= j) l_idx++;
        t_idx = j - lcdf_off[l_idx];
    }
    uint64_t n_closest = 0, n_any = 0;

    /* rtcommon.glsl:111-124 */
    uint32_t lcg() {
        seed = 1664525u * seed + 1013904223u;
        return seed & 0x00FFFFFFu;
    }
    float randf() { return (float)lcg() / (float)0x01000000; }
    uint32_t randu(uint32_t a, uint32_t b) { return lcg() % (b - a) + a; }

    const float* model(uint32_t o) const { return (const float*)(S->descs + 52ull * o); }
    const float* modelIT(uint32_t o) const { return (const float*)(S->descs + 52ull * o + 16); }
    vec3 desc_vec(uint32_t o, int word) const {
        const float* f = (const float*)(S->descs + 52ull * o + word);
        return {f[0], f[1], f[2]};
    }
    int desc_int(uint32_t o, int word) const { return (int)S->descs[52ull * o + word]; }
    const float* vertex(uint32_t obj, uint32_t i) const { return S->verts + 12ull * (S->vert_off[obj] + i); }
    void tri_indices(uint32_t obj, uint32_t prim, uint32_t ind[3]) const {
        const uint32_t* p = S->idx + 3ull * (S->tri_off[obj] + prim);
        ind[0] = p[0], ind[1] = p[1], ind[2] = p[2];
    }

    /* texture(): R8G8B8A8_SRGB, linear filter, REPEAT, single mip (rt.cpp:439-446, vulkan.cpp:515-537) */
    vec3 texture(int t, const float tc[2]) const {
        const uint32_t* info = S->tex_info + 4ull * t;
        int w = (int)info[1], h = (int)info[2];
        const uint8_t* base = S->texels + 4ull * info[0];
        float x = tc[0] * (float)w - 0.5f, y = tc[1] * (float)h - 0.5f;
        float fx0 = floorf(x), fy0 = floorf(y);
        float ax = x - fx0, ay = y - fy0;
        int x0 = (int)fx0, y0 = (int)fy0;
        auto wrap = [](int i, int n) {
            int m = i % n;
            return m < 0 ? m + n : m;
        };
        int xa = wrap(x0, w), xb = wrap(x0 + 1, w), ya = wrap(y0, h), yb = wrap(y0 + 1, h);
        auto tex = [&](int xx, int yy) {
            const uint8_t* p = base + 4ull * ((size_t)yy * w + xx);
            return vec3{srgb_lut[p[0]], srgb_lut[p[1]], srgb_lut[p[2]]};
        };
        vec3 top = tex(xa, ya) * (1.0f - ax) + tex(xb, ya) * ax;
        vec3 bot = tex(xa, yb) * (1.0f - ax) + tex(xb, yb) * ax;
        return top * (1.0f - ay) + bot * ay;
    }
    /* NEAREST fetch of a previous-frame G-buffer (rt.cpp:449-450), REPEAT */
    vec3 gbuf_fetch(const float* img, float u, float v) const {
        int x = (int)floorf(u * (float)W), y = (int)floorf(v * (float)H);
        x = ((x % (int)W) + (int)W) % (int)W, y = ((y % (int)H) + (int)H) % (int)H;
        const float* p = img + 4ull * ((size_t)y * W + x);
        return {p[0], p[1], p[2]};
    }

    /* traceRayEXT closest (rt.rgen:257-270) + rt.rchit:11-16 + rt.rmiss:9-11 */
    void trace_ray(vec3 o, vec3 d) {
        float ray[8] = {o.x, o.y, o.z, EPS, d.x, d.y, d.z, LARGE_DIST};
        uint32_t hit[4];
        n_closest++;
        if(g_rec_buf) {
            uint64_t k = g_rec_count.fetch_add(1);
            if(k < g_rec_cap) memcpy(g_rec_buf + 8 * k, ray, sizeof(ray));
        }
        if(orc_bvh_trace_one(S->bvh, ray, hit)) {
            float u = u2f(hit[1]), v = u2f(hit[2]);
            uint32_t gid = hit[3];
            uint32_t a = 0, b = S->n_objs;
            while(b - a > 1) {
                uint32_t m = (a + b) >> 1;
                if(S->tri_off[m] <= gid) a = m; else b = m;
            }
            payload.hit = true;
            payload.barycentrics = {1.0f - u - v, u, v};
            payload.obj_id = a;
            payload.prim_id = gid - S->tri_off[a];
        } else
            payload.hit = false;
    }
    /* rt.rgen:272-291 */
    bool visibility(vec3 a, vec3 b) {
        vec3 dir = b - a;
        float d = length(dir);
        vec3 nd = dir / d;
        float ray[8] = {a.x, a.y, a.z, EPS, nd.x, nd.y, nd.z, d - EPS};
        n_any++;
        return orc_bvh_occluded_one(S->bvh, ray) != 0;
    }

    /* rt.rgen:62-95 */
    HitInfo hit_info() const {
        uint32_t obj = payload.obj_id; /* objects[obj].index == obj (rt.cpp:33) */
        const float *mIT = modelIT(obj), *m = model(obj);
        vec3 bary = payload.barycentrics;
        uint32_t ind[3];
        tri_indices(obj, payload.prim_id, ind);
        const float *v0 = vertex(obj, ind[0]), *v1 = vertex(obj, ind[1]), *v2 = vertex(obj, ind[2]);
        HitInfo hit;
        vec3 n = vec3{v0[4], v0[5], v0[6]} * bary.x + vec3{v1[4], v1[5], v1[6]} * bary.y + vec3{v2[4], v2[5], v2[6]} * bary.z;
        hit.normal = normalize(xform_dir(mIT, n));
        vec3 t0 = vec3{v0[8], v0[9], v0[10]} * v0[11], t1 = vec3{v1[8], v1[9], v1[10]} * v1[11],
             t2 = vec3{v2[8], v2[9], v2[10]} * v2[11];
        vec3 t = t0 * bary.x + t1 * bary.y + t2 * bary.z;
        hit.tangent = normalize(xform_dir(mIT, t));
        vec3 p = vec3{v0[0], v0[1], v0[2]} * bary.x + vec3{v1[0], v1[1], v1[2]} * bary.y + vec3{v2[0], v2[1], v2[2]} * bary.z;
        hit.pos = xform_point(m, p);
        hit.tc[0] = v0[3] * bary.x + v1[3] * bary.y + v2[3] * bary.z;
        hit.tc[1] = v0[7] * bary.x + v1[7] * bary.y + v2[7] * bary.z;
        return hit;
    }
    /* rt.rgen:97-130. Scene_Desc words: albedo 32, emissive 36, metal_rough 40, tex ids 44..47 */
    MatInfo mat_info(const HitInfo& hit) const {
        uint32_t obj = payload.obj_id;
        MatInfo mat;
        int albedoIdx = desc_int(obj, 44);
        mat.albedo = desc_vec(obj, 32);
        if(albedoIdx >= 0) mat.albedo = texture(albedoIdx, hit.tc);
        int emissiveIdx = desc_int(obj, 45);
        mat.emissive = desc_vec(obj, 36);
        if(emissiveIdx >= 0) mat.emissive = texture(emissiveIdx, hit.tc);
        int mrIdx = desc_int(obj, 46);
        vec3 mr = desc_vec(obj, 40);
        if(mrIdx >= 0) mr = texture(mrIdx, hit.tc);
        mat.roughness = mr.y;
        if(c->use_metalness == 1) mat.albedo = mix(v3(0.04f), mat.albedo, mr.x);
        int nIdx = desc_int(obj, 47);
        mat.use_tanspace = nIdx >= 0;
        mat.tanspaceNormal = v3(0);
        if(mat.use_tanspace) mat.tanspaceNormal = texture(nIdx, hit.tc) * 2.0f - v3(1.0f);
        return mat;
    }
    /* rtcommon.glsl:218-224 */
    static void make_tanspace(vec3 N, vec3& Nt, vec3& Nb) {
        if(fabsf(N.x) > fabsf(N.y)) Nt = vec3{N.z, 0, -N.x} / sqrtf(N.x * N.x + N.z * N.z);
        else Nt = vec3{0, -N.z, N.y} / sqrtf(N.y * N.y + N.z * N.z);
        Nb = cross(N, Nt);
    }
    /* rt.rgen:132-149 */
    ShadeInfo shade_info(const TraceInfo& trace, HitInfo hit, const MatInfo& mat) const {
        ShadeInfo shade;
        shade.wo = trace.d;
        if(dot(shade.wo, hit.normal) > 0) hit.normal = -hit.normal;
        shade.T = hit.tangent;
        shade.N = hit.normal;
        if(mat.use_tanspace && c->use_normal_map == 1) {
            shade.B = cross(shade.N, shade.T);
            vec3 tn = mat.tanspaceNormal;
            shade.N = normalize(shade.T * tn.x + shade.B * tn.y + shade.N * tn.z);
        }
        make_tanspace(shade.N, shade.T, shade.B);
        return shade;
    }

    /* ---- sampling: rtcommon.glsl:137-179 ---- */
    vec3 cospow_hemisphere(float exponent, vec3 x, vec3 y, vec3 z) {
        float phi = (2 * M_PI_F) * randf();
        float cosT = dm_pow(randf(), 1.0f / (exponent + 1.0f));
        float sinT = sqrtf(1.0f - cosT * cosT);
        vec3 dir = {dm_cos(phi) * sinT, dm_sin(phi) * sinT, cosT};
        return dir.x * x + dir.y * y + dir.z * z;
    }
    vec3 triangle_sample() {
        float u = sqrtf(randf());
        float v = randf();
        float a = u * (1 - v);
        float b = u * v;
        return {a, b, 1 - a - b};
    }

    /* ---- materials: rtcommon.glsl:255-369 ---- */
    float bp_pdf(const MatInfo& mat, const ShadeInfo& sh, vec3 wi) const {
        float oDn = dot(-sh.wo, sh.N), iDn = dot(wi, sh.N);
        if(oDn <= 0 || iDn <= 0) return 0;
        float ex = 1 / mat.roughness;
        vec3 Hh = normalize(wi - sh.wo);
        float cosine = fmaxf(dot(Hh, sh.N), 0.0f);
        float N_pdf = (ex + 1) / (2 * M_PI_F) * dm_pow(cosine, ex);
        return N_pdf / (4 * dot(-sh.wo, Hh));
    }
    static vec3 GGX_F(vec3 r0, float iDn) {
        float cos5 = dm_pow(1 - iDn, 5.0f);
        return r0 + (v3(1) - r0) * cos5;
    }
    static float GGX_G(float oDn, float iDn, float a2) {
        float sqr0 = sqrtf(a2 + (1 - a2) * iDn * iDn);
        float sqr1 = sqrtf(a2 + (1 - a2) * oDn * oDn);
        return 2 * oDn * iDn / (oDn * sqr0 + iDn * sqr1);
    }
    static float GGX_D(float nDh, float a2) {
        float b = nDh * nDh * (a2 - 1) + 1;
        return a2 / (M_PI_F * b * b);
    }
    float GGX_pdf(const MatInfo& mat, const ShadeInfo& sh, vec3 wi) const {
        float oDn = dot(-sh.wo, sh.N), iDn = dot(wi, sh.N);
        if(oDn <= 0 || iDn <= 0) return 0;
        vec3 Hh = normalize(wi - sh.wo);
        float nDh = fmaxf(dot(Hh, sh.N), 0.0f);
        float oDh = fmaxf(dot(wi, Hh), 0.0f);
        float a2 = mat.roughness * mat.roughness;
        return GGX_D(nDh, a2) * nDh / (4 * oDh);
    }
    vec3 GGX_eval(const MatInfo& mat, const ShadeInfo& sh, vec3 wi) const {
        float oDn = dot(-sh.wo, sh.N), iDn = dot(wi, sh.N);
        if(oDn <= 0 || iDn <= 0) return v3(0);
        vec3 Hh = normalize(wi - sh.wo);
        float nDh = fmaxf(dot(Hh, sh.N), 0.0f);
        float a2 = mat.roughness * mat.roughness;
        return GGX_F(mat.albedo, iDn) * GGX_D(nDh, a2) * GGX_G(oDn, iDn, a2) / (4 * oDn);
    }
    float MAT_pdf(const MatInfo& mat, const ShadeInfo& sh, vec3 wi) const {
        if(c->brdf == 0) return bp_pdf(mat, sh, wi);
        if(c->brdf == 1) return GGX_pdf(mat, sh, wi);
        return 0;
    }
    vec3 MAT_eval(const MatInfo& mat, const ShadeInfo& sh, vec3 wi) const {
        if(c->brdf == 0) return mat.albedo * bp_pdf(mat, sh, wi);
        if(c->brdf == 1) return GGX_eval(mat, sh, wi);
        return v3(0);
    }
    bool MAT_sample(const MatInfo& mat, const ShadeInfo& sh, vec3& wi) {
        if(c->brdf == 0) {
            float ex = 1 / mat.roughness;
            vec3 Hh = cospow_hemisphere(ex, sh.T, sh.B, sh.N);
            wi = reflect(sh.wo, Hh);
            return dot(wi, sh.N) > 0;
        }
        if(c->brdf == 1) {
            float a2 = mat.roughness * mat.roughness;
            float Xi_x = randf(), Xi_y = randf();
            float phi = (2.0f * M_PI_F) * Xi_x;
            float cosTheta = sqrtf((1.0f - Xi_y) / (1.0f + (a2 - 1.0f) * Xi_y));
            float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
            vec3 dir = {dm_cos(phi) * sinTheta, dm_sin(phi) * sinTheta, cosTheta};
            vec3 Hh = sh.T * dir.x + sh.B * dir.y + sh.N * dir.z;
            wi = reflect(sh.wo, Hh);
            return dot(wi, sh.N) > 0;
        }
        wi = v3(0);
        return false;
    }

    /* ---- lights ---- */
    const uint32_t* light(uint32_t l) const { return S->lights + 12ull * l; }
    /* rt.rgen:151-198.  Q4: with no lights the GLSL takes lcg % 0; here the sample has pdf 0 and
     * draws nothing. */
    LightSample light_sample(vec3 p) {
        LightSample s;
        if(c->n_lights <= 0) {
            s.l_idx = s.o_idx = s.t_idx = 0;
            s.pos = s.normal = s.emissive = v3(0);
            s.pdf = 0;
            return s;
        }
        float prob = 0;
        const bool by_power = power_sampling();
        if(by_power) light_pick(s.l_idx, s.t_idx, prob);
        else s.l_idx = randu(0, (uint32_t)c->n_lights);
        s.o_idx = light(s.l_idx)[8];
        uint32_t n_tris = light(s.l_idx)[9];
        if(!by_power) s.t_idx = randu(0, n_tris);
        uint32_t ind[3];
        tri_indices(s.o_idx, s.t_idx, ind);
        const float *v0 = vertex(s.o_idx, ind[0]), *v1 = vertex(s.o_idx, ind[1]), *v2 = vertex(s.o_idx, ind[2]);
        const float* m = model(s.o_idx);
        vec3 _v0 = xform_point(m, {v0[0], v0[1], v0[2]}), _v1 = xform_point(m, {v1[0], v1[1], v1[2]}),
             _v2 = xform_point(m, {v2[0], v2[1], v2[2]});
        vec3 bary = triangle_sample();
        /* Q5: the v coordinate is v1's for all three vertices (rt.rgen:177) */
        float tc[2] = {v0[3] * bary.x + v1[3] * bary.y + v2[3] * bary.z, v1[7] * bary.x + v1[7] * bary.y + v1[7] * bary.z};
        s.pos = _v0 * bary.x + _v1 * bary.y + _v2 * bary.z;
        int emissiveIdx = desc_int(s.o_idx, 45);
        s.emissive = desc_vec(s.o_idx, 36);
        if(emissiveIdx >= 0) s.emissive = texture(emissiveIdx, tc);
        vec3 Narea = cross(_v1 - _v0, _v2 - _v0);
        float a = 2 / length(Narea);
        vec3 dist = s.pos - p;
        vec3 N = normalize(Narea);
        vec3 d = normalize(dist);
        float g = dot(dist, dist) / fabsf(dot(N, d));
        s.normal = N;
        s.pdf = by_power ? a * g * prob : a * g / (float)(n_tris * (uint32_t)c->n_lights);
        return s;
    }
    /* rt.rgen:200-220 */
    vec3 light_sample_dir(vec3 p) {
        uint32_t l_idx, t_idx;
        if(power_sampling()) {
            float prob;
            light_pick(l_idx, t_idx, prob);
        } else {
            l_idx = randu(0, (uint32_t)c->n_lights);
            t_idx = randu(0, light(l_idx)[9]);
        }
        uint32_t o_idx = light(l_idx)[8];
        uint32_t ind[3];
        tri_indices(o_idx, t_idx, ind);
        const float *v0 = vertex(o_idx, ind[0]), *v1 = vertex(o_idx, ind[1]), *v2 = vertex(o_idx, ind[2]);
        vec3 bary = triangle_sample();
        vec3 point = vec3{v0[0], v0[1], v0[2]} * bary.x + vec3{v1[0], v1[1], v1[2]} * bary.y + vec3{v2[0], v2[1], v2[2]} * bary.z;
        point = xform_point(model(o_idx), point);
        return normalize(point - p);
    }
    /* rtcommon.glsl:181-214 */
    static bool triangle_hit(vec3 o, vec3 d, vec3 pa, vec3 pb, vec3 pc, vec3& hitp) {
        vec3 v1 = pb - pa, v2 = pc - pa;
        vec3 p = cross(d, v2);
        float det = dot(v1, p);
        if(fabsf(det) < EPS) return false;
        float invDet = 1 / det;
        vec3 s = o - pa;
        float u = dot(s, p) * invDet;
        if(u < 0 || u > 1) return false;
        vec3 q = cross(s, v1);
        float v = dot(d, q) * invDet;
        if(v < 0 || u + v > 1) return false;
        float t = dot(v2, q) * invDet;
        hitp = o + t * d;
        return t >= 0;
    }
    static float triangle_pdf(vec3 o, vec3 d, vec3 v0, vec3 v1, vec3 v2) {
        vec3 hitp;
        if(triangle_hit(o, d, v0, v1, v2, hitp)) {
            float a = 2 / length(cross(v1 - v0, v2 - v0));
            vec3 dist = hitp - o;
            vec3 N = normalize(cross(v1 - v0, v2 - v0));
            float g = dot(dist, dist) / fabsf(dot(N, d));
            return a * g;
        }
        return 0;
    }
    /* rtcommon.glsl:242-251 */
    static bool hit_bbox(vec3 o, vec3 d, vec3 bmin, vec3 bmax) {
        vec3 invD = v3(1) / d;
        vec3 t0 = (bmin - o) * invD, t1 = (bmax - o) * invD;
        vec3 tNear = {fminf(t0.x, t1.x), fminf(t0.y, t1.y), fminf(t0.z, t1.z)};
        vec3 tFar = {fmaxf(t0.x, t1.x), fmaxf(t0.y, t1.y), fmaxf(t0.z, t1.z)};
        float tNearMax = fmaxf(fmaxf(tNear.x, tNear.y), fmaxf(tNear.z, 0.0f));
        float tFarMin = fminf(fminf(tFar.x, tFar.y), tFar.z);
        return tNearMax <= tFarMin;
    }
    /* rt.rgen:222-255 */
    float light_pdf(vec3 p, vec3 d) const {
        if(c->n_lights <= 0) return 0; /* Q4 */
        const bool by_power = power_sampling(); /* extension: every term weighted by its triangle's probability */
        float oacc = 0;
        for(uint32_t l = 0; l < (uint32_t)c->n_lights; l++) {
            float tacc = 0;
            const uint32_t* L = light(l);
            uint32_t o_idx = L[8], n_tris = L[9];
            const float* lf = (const float*)L;
            if(!hit_bbox(p, d, {lf[0], lf[1], lf[2]}, {lf[4], lf[5], lf[6]})) continue;
            const float* m = model(o_idx);
            for(uint32_t t = 0; t < n_tris; t++) {
                uint32_t ind[3];
                tri_indices(o_idx, t, ind);
                const float *a = vertex(o_idx, ind[0]), *b = vertex(o_idx, ind[1]), *cc = vertex(o_idx, ind[2]);
                vec3 v0 = xform_point(m, {a[0], a[1], a[2]}), v1 = xform_point(m, {b[0], b[1], b[2]}),
                     v2 = xform_point(m, {cc[0], cc[1], cc[2]});
                float term = triangle_pdf(p, d, v0, v1, v2);
                if(by_power) term = term * light_tri_prob(lcdf_off[l] + t);
                tacc += term;
            }
            oacc += by_power ? tacc : tacc / (float)n_tris;
        }
        return by_power ? oacc : oacc / (float)c->n_lights;
    }
    /* rt.rgen:293-301 */
    vec3 direct_light(vec3 o, vec3 d) {
        trace_ray(o, d);
        if(!payload.hit) return {c->env_light[0], c->env_light[1], c->env_light[2]};
        HitInfo hit = hit_info();
        MatInfo mat = mat_info(hit);
        return mat.emissive;
    }
    static float power_heuristic(float a, float b) { return a * a / (a * a + b * b); }
    static float luma(vec3 rgb) { return 0.299f * rgb.x + 0.587f * rgb.y + 0.114f * rgb.z; }

    /* ---- integrators ---- */
    /* rt.rgen:303-353 */
    void integrate_mis(TraceInfo& trace, const HitInfo& hit, const MatInfo& mat, const ShadeInfo& shade) {
        if(any_gt0(mat.emissive)) {
            trace.acc = trace.acc + trace.throughput * trace.mis * mat.emissive;
            trace.depth = c->max_depth;
            return;
        }
        trace.o = hit.pos;
        if(mat.roughness == 0) {
            trace.d = reflect(shade.wo, shade.N);
            trace.throughput = trace.throughput * mat.albedo;
            trace.mis = 1;
        } else {
            if(c->n_lights > 0) { /* Q4 */
                vec3 wi_light = light_sample_dir(hit.pos);
                float light_pdf_l = light_pdf(hit.pos, wi_light);
                if(light_pdf_l != 0) {
                    float light_pdf_m = MAT_pdf(mat, shade, wi_light);
                    vec3 light_atten = MAT_eval(mat, shade, wi_light);
                    vec3 weight = light_atten / light_pdf_l * power_heuristic(light_pdf_l, light_pdf_m);
                    trace.acc = trace.acc + trace.throughput * weight * direct_light(hit.pos, wi_light);
                }
            }
            vec3 wi_brdf;
            if(!MAT_sample(mat, shade, wi_brdf)) {
                trace.depth = c->max_depth;
                return;
            }
            float brdf_pdf_m = MAT_pdf(mat, shade, wi_brdf);
            if(brdf_pdf_m != 0) {
                float brdf_pdf_l = light_pdf(hit.pos, wi_brdf);
                vec3 brdf_atten = MAT_eval(mat, shade, wi_brdf);
                trace.throughput = trace.throughput * (brdf_atten / brdf_pdf_m);
                trace.mis = power_heuristic(brdf_pdf_m, brdf_pdf_l);
            } else {
                trace.depth = c->max_depth;
                return;
            }
            trace.d = wi_brdf;
        }
    }
    /* rt.rgen:355-389 */
    void integrate_mats(TraceInfo& trace, const HitInfo& hit, const MatInfo& mat, const ShadeInfo& shade) {
        if(any_gt0(mat.emissive)) {
            trace.acc = trace.acc + mat.emissive * trace.throughput;
            trace.depth = c->max_depth;
            return;
        }
        trace.o = hit.pos;
        if(mat.roughness == 0) {
            trace.d = reflect(shade.wo, shade.N);
            trace.throughput = trace.throughput * mat.albedo;
        } else {
            vec3 wi;
            if(!MAT_sample(mat, shade, wi)) {
                trace.depth = c->max_depth;
                return;
            }
            float pdf = MAT_pdf(mat, shade, wi);
            vec3 atten = MAT_eval(mat, shade, wi);
            if(pdf != 0) trace.throughput = trace.throughput * (atten / pdf);
            else {
                trace.depth = c->max_depth;
                return;
            }
            trace.d = wi;
        }
    }
    /* rt.rgen:391-411 */
    void integrate_direct(TraceInfo& trace, const HitInfo& hit, const MatInfo& mat, const ShadeInfo& shade) {
        trace.depth = c->max_depth;
        if(any_gt0(mat.emissive)) {
            trace.acc = trace.acc + mat.emissive;
            return;
        }
        if(mat.roughness != 0) {
            LightSample light = light_sample(hit.pos);
            vec3 wi = normalize(light.pos - hit.pos);
            vec3 light_atten = MAT_eval(mat, shade, wi);
            if(light.pdf != 0) {
                float shadow = visibility(hit.pos, light.pos) ? 0.0f : 1.0f;
                trace.acc = trace.acc + light_atten / light.pdf * light.emissive * shadow;
            }
        }
    }

    /* restir.glsl:17-35 */
    void res_update(Reservoir& res, float weight, vec3 pos, vec3 normal, vec3 emissive) {
        res.n_seen++;
        res.w_sum += weight;
        if(randf() < weight / res.w_sum) {
            res.pos = pos;
            res.normal = normal;
            res.emissive = emissive;
        }
    }
    static Reservoir res_new() {
        Reservoir r;
        r.pos = r.normal = r.emissive = v3(0);
        r.w_sum = 0, r.w = 0, r.n_seen = 0;
        return r;
    }
    /* rt.rgen:415-433 */
    float update_weight(Reservoir& res, const HitInfo& hit, const MatInfo& mat, const ShadeInfo& shade) const {
        if(res.n_seen == 0) {
            res.w = 0;
            return 0;
        }
        vec3 dir = res.pos - hit.pos;
        vec3 wi = normalize(dir);
        vec3 light_atten = MAT_eval(mat, shade, wi);
        float g = fabsf(dot(res.normal, wi)) / dot(dir, dir);
        vec3 contrib = g * light_atten * res.emissive;
        float pHat = luma(contrib);
        res.w = (1 / pHat) * (res.w_sum / (float)res.n_seen);
        if(pHat == 0) res.w = 0;
        return pHat;
    }
    /* rt.rgen:435-505 */
    void reservoir_sample(TraceInfo& trace, const HitInfo& hit, const MatInfo& mat, const ShadeInfo& shade, bool first) {
        Reservoir new_res = res_new();
        if(c->n_lights > 0) /* Q4 */
            for(uint32_t i = 0; i < cam->new_samples; i++) {
                LightSample light = light_sample(hit.pos);
                vec3 wi = normalize(light.pos - hit.pos);
                vec3 light_atten = MAT_eval(mat, shade, wi);
                vec3 contrib = light_atten * light.emissive / light.pdf;
                res_update(new_res, luma(contrib), light.pos, light.normal, light.emissive);
            }
        float new_pHat = update_weight(new_res, hit, mat, shade);
        if(new_pHat != 0 && visibility(hit.pos, new_res.pos)) new_res.w = 0;
        for(;;) {
            if(first && c->use_temporal == 1) {
                vec4 pp = mul4(cam->prev_PV, hit.pos.x, hit.pos.y, hit.pos.z, 1.0f);
                pp.x /= pp.w, pp.y /= pp.w, pp.z /= pp.w;
                pp.x = (pp.x + 1.0f) * 0.5f, pp.y = (pp.y + 1.0f) * 0.5f;
                if(!((pp.x > 0 && pp.y > 0) && (pp.x < 1 && pp.y < 1))) break;
                vec3 old_pos = gbuf_fetch(ppos, pp.x, pp.y);
                vec3 old_norm = gbuf_fetch(pnorm, pp.x, pp.y);
                vec3 old_alb = gbuf_fetch(palb, pp.x, pp.y);
                vec3 posdiff = old_pos - hit.pos;
                if(dot(posdiff, posdiff) > 0.01f) break;
                vec3 albdiff = old_alb - mat.albedo;
                if(dot(albdiff, albdiff) > 0.01f) break;
                if(dot(old_norm, shade.N) < 0.5f) break;
                int fx = (int)(pp.x * (float)W), fy = (int)(pp.y * (float)H);
                const uint32_t* r = prev_res_buf + 12ull * ((size_t)fy * W + fx);
                prev_res.pos = {u2f(r[0]), u2f(r[1]), u2f(r[2])}, prev_res.w_sum = u2f(r[3]);
                prev_res.normal = {u2f(r[4]), u2f(r[5]), u2f(r[6])}, prev_res.w = u2f(r[7]);
                prev_res.emissive = {u2f(r[8]), u2f(r[9]), u2f(r[10])}, prev_res.n_seen = r[11];
            }
            Reservoir temporal_res = res_new();
            res_update(temporal_res, new_pHat * new_res.w * (float)new_res.n_seen, new_res.pos, new_res.normal, new_res.emissive);
            float old_pHat = update_weight(prev_res, hit, mat, shade);
            uint32_t cap = cam->temporal_multiplier * new_res.n_seen;
            prev_res.n_seen = cap < prev_res.n_seen ? cap : prev_res.n_seen;
            res_update(temporal_res, old_pHat * prev_res.w * (float)prev_res.n_seen, prev_res.pos, prev_res.normal, prev_res.emissive);
            temporal_res.n_seen = new_res.n_seen + prev_res.n_seen;
            update_weight(temporal_res, hit, mat, shade);
            new_res = temporal_res;
            break;
        }
        if(first && spatial_samples > 0) { /* extension, see orc_render_set_spatial */
            /* the pixel's reservoir and `spatial_samples` reservoirs of the previous frame from a disc of `spatial_radius`
             * pixels: each enters the way prev_res enters the temporal step (weight re-derived here, history capped) */
            Reservoir comb = res_new();
            float cur_pHat = update_weight(new_res, hit, mat, shade);
            res_update(comb, cur_pHat * new_res.w * (float)new_res.n_seen, new_res.pos, new_res.normal, new_res.emissive);
            uint32_t total = new_res.n_seen;
            const uint32_t cap = cam->temporal_multiplier * cam->new_samples;
            for(uint32_t i = 0; i < spatial_samples; i++) {
                float r = spatial_radius * sqrtf(randf());
                float phi = 2.0f * M_PI_F * randf();
                int qx = (int)px + (int)floorf(r * dm_cos(phi) + 0.5f), qy = (int)py + (int)floorf(r * dm_sin(phi) + 0.5f);
                if(qx < 0 || qy < 0 || qx >= (int)W || qy >= (int)H) continue;
                size_t q = (size_t)qy * W + (size_t)qx;
                vec3 npos = {ppos[4 * q], ppos[4 * q + 1], ppos[4 * q + 2]}, nnorm = {pnorm[4 * q], pnorm[4 * q + 1], pnorm[4 * q + 2]};
                vec3 d = npos - hit.pos;
                if(!(dot(nnorm, shade.N) >= 0.9f)) continue;
                if(!(fabsf(dot(d, shade.N)) <= 0.05f * sqrtf(dot(d, d)) + 1.0e-3f)) continue;
                const uint32_t* rr = prev_res_buf + 12ull * q;
                Reservoir nb;
                nb.pos = {u2f(rr[0]), u2f(rr[1]), u2f(rr[2])}, nb.w_sum = u2f(rr[3]);
                nb.normal = {u2f(rr[4]), u2f(rr[5]), u2f(rr[6])}, nb.w = u2f(rr[7]);
                nb.emissive = {u2f(rr[8]), u2f(rr[9]), u2f(rr[10])}, nb.n_seen = rr[11];
                if(nb.n_seen == 0) continue;
                float nb_pHat = update_weight(nb, hit, mat, shade);
                nb.n_seen = cap < nb.n_seen ? cap : nb.n_seen;
                res_update(comb, nb_pHat * nb.w * (float)nb.n_seen, nb.pos, nb.normal, nb.emissive);
                total += nb.n_seen;
            }
            comb.n_seen = total;
            float pHat = update_weight(comb, hit, mat, shade);
            if(pHat != 0 && comb.w != 0 && visibility(hit.pos, comb.pos)) comb.w = 0;
            new_res = comb;
        }
        if(new_res.w != 0) {
            vec3 dir = new_res.pos - hit.pos;
            vec3 wi = normalize(dir);
            vec3 light_atten = MAT_eval(mat, shade, wi);
            vec3 contrib = light_atten * new_res.emissive;
            float g = fabsf(dot(new_res.normal, wi)) / dot(dir, dir);
            trace.acc = trace.acc + new_res.w * contrib * g;
        }
        prev_res = new_res;
    }
    /* rt.rgen:507-549 */
    void integrate_restir(TraceInfo& trace, const HitInfo& hit, const MatInfo& mat, const ShadeInfo& shade, bool d_only, bool first) {
        if(any_gt0(mat.emissive)) {
            trace.acc = trace.acc + mat.emissive * trace.throughput * trace.mis;
            trace.depth = c->max_depth;
            return;
        }
        trace.o = hit.pos;
        if(mat.roughness == 0) {
            trace.d = reflect(shade.wo, shade.N);
            trace.throughput = trace.throughput * mat.albedo;
            trace.mis = 1;
        } else {
            if(trace.depth == 0) reservoir_sample(trace, hit, mat, shade, first);
            vec3 wi_brdf;
            if(!MAT_sample(mat, shade, wi_brdf)) {
                trace.depth = c->max_depth;
                return;
            }
            float brdf_pdf = MAT_pdf(mat, shade, wi_brdf);
            if(brdf_pdf != 0) {
                vec3 brdf_atten = MAT_eval(mat, shade, wi_brdf);
                trace.throughput = trace.throughput * (brdf_atten / brdf_pdf);
                trace.mis = trace.depth == 0 ? 0.0f : 1.0f;
            } else {
                trace.depth = c->max_depth;
                return;
            }
            trace.d = wi_brdf;
        }
        if(d_only) trace.depth = c->max_depth;
    }

    /* rt.rgen:551-565 */
    void make_camera_ray(uint32_t s, vec3& d) {
        float jx, jy;
        if(c->qmc == 0) {
            if(c->frame == 0) jx = jy = 0.5f;
            else {
                jx = randf();
                jy = randf();
            }
        } else {
            uint32_t i = s + (uint32_t)(c->samples * c->frame), N = (uint32_t)(c->samples * c->max_frame);
            jx = (float)i / (float)N;
            jy = orc_radical_inverse(i);
        }
        float pcx = (float)px + jx, pcy = (float)py + jy;
        float ux = pcx / (float)W, uy = pcy / (float)H;
        vec4 target = mul4(cam->iP, ux * 2.0f - 1.0f, uy * 2.0f - 1.0f, 0.0f, 1.0f);
        vec4 direction = mul4(cam->iV, target.x, target.y, target.z, 0.0f);
        d = normalize({direction.x, direction.y, direction.z});
    }

    /* rt.rgen:567-677 */
    void main_(uint32_t seed_val, float* image, uint32_t* out_res, float* pos_img, float* norm_img, float* alb_img) {
        size_t pix = (size_t)py * W + px;
        seed = orc_tea((uint32_t)pix, seed_val);
        vec3 acc = v3(0);
        vec4 co = mul4(cam->iV, 0, 0, 0, 1);
        vec3 camera_o = {co.x, co.y, co.z};
        vec3 gbuf_pos = v3(0), gbuf_norm = v3(0), gbuf_albedo = v3(0);
        prev_res = res_new();
        for(int s = 0; s < c->samples; s++) {
            TraceInfo trace;
            trace.o = camera_o;
            make_camera_ray((uint32_t)s, trace.d);
            trace.acc = v3(0);
            trace.throughput = v3(1);
            trace.depth = 0;
            trace.mis = 1;
            for(; trace.depth < (uint32_t)c->max_depth; trace.depth++) {
                trace_ray(trace.o, trace.d);
                if(!payload.hit) {
                    if(trace.depth == 0) trace.acc = {c->clear_col[0], c->clear_col[1], c->clear_col[2]};
                    else trace.acc = trace.acc + vec3{c->env_light[0], c->env_light[1], c->env_light[2]} * trace.throughput;
                    break;
                }
                HitInfo hit = hit_info();
                MatInfo mat = mat_info(hit);
                ShadeInfo shade = shade_info(trace, hit, mat);
                if(s == 0 && trace.depth == 0) {
                    gbuf_pos = hit.pos;
                    gbuf_norm = shade.N;
                    gbuf_albedo = mat.albedo;
                }
                if(c->integrator == 0) integrate_direct(trace, hit, mat, shade);
                else if(c->integrator == 1) integrate_mats(trace, hit, mat, shade);
                else if(c->integrator == 2) integrate_mis(trace, hit, mat, shade);
                else if(c->integrator == 3) integrate_restir(trace, hit, mat, shade, true, s == 0);
                else if(c->integrator == 4) integrate_restir(trace, hit, mat, shade, false, s == 0);
                if(c->use_rr == 1) {
                    float pcont = fminf(fmaxf(trace.throughput.x, fmaxf(trace.throughput.y, trace.throughput.z)) + 0.001f, 0.95f);
                    if(randf() >= pcont) break;
                    trace.throughput = trace.throughput / pcont;
                }
            }
            acc = acc + trace.acc;
        }
        if(c->integrator == 3 || c->integrator == 4) {
            uint32_t* r = out_res + 12ull * pix;
            r[0] = f2u(prev_res.pos.x), r[1] = f2u(prev_res.pos.y), r[2] = f2u(prev_res.pos.z), r[3] = f2u(prev_res.w_sum);
            r[4] = f2u(prev_res.normal.x), r[5] = f2u(prev_res.normal.y), r[6] = f2u(prev_res.normal.z), r[7] = f2u(prev_res.w);
            r[8] = f2u(prev_res.emissive.x), r[9] = f2u(prev_res.emissive.y), r[10] = f2u(prev_res.emissive.z), r[11] = prev_res.n_seen;
        }
        vec3 avg = acc / (float)c->samples;
        float* px4 = image + 4 * pix;
        if(c->frame > 0) {
            float a = 1.0f / (float)(c->frame + 1);
            vec3 old_color = {px4[0], px4[1], px4[2]};
            vec3 m = mix(old_color, avg, a);
            px4[0] = m.x, px4[1] = m.y, px4[2] = m.z, px4[3] = 1;
        } else
            px4[0] = avg.x, px4[1] = avg.y, px4[2] = avg.z, px4[3] = 1;
        if(c->debug_view > 0) { /* rt.rgen:647-672 */
            vec4 pp = mul4(cam->prev_PV, gbuf_pos.x, gbuf_pos.y, gbuf_pos.z, 1.0f);
            pp.x /= pp.w, pp.y /= pp.w, pp.z /= pp.w;
            pp.x = (pp.x + 1.0f) * 0.5f, pp.y = (pp.y + 1.0f) * 0.5f;
            if(dot(gbuf_norm, gbuf_norm) > 0.5f && (pp.x > 0 && pp.y > 0) && (pp.x < 1 && pp.y < 1)) {
                vec3 v = c->debug_view == 1 ? gbuf_fetch(ppos, pp.x, pp.y)
                         : c->debug_view == 2 ? gbuf_fetch(pnorm, pp.x, pp.y)
                                              : gbuf_fetch(palb, pp.x, pp.y);
                if(c->debug_view <= 3) px4[0] = v.x, px4[1] = v.y, px4[2] = v.z, px4[3] = 1;
            } else
                px4[0] = px4[1] = px4[2] = 0, px4[3] = 1;
        }
        float *gp = pos_img + 4 * pix, *gn = norm_img + 4 * pix, *ga = alb_img + 4 * pix;
        gp[0] = gbuf_pos.x, gp[1] = gbuf_pos.y, gp[2] = gbuf_pos.z, gp[3] = 1;
        gn[0] = gbuf_norm.x, gn[1] = gbuf_norm.y, gn[2] = gbuf_norm.z, gn[3] = 1;
        ga[0] = gbuf_albedo.x, ga[1] = gbuf_albedo.y, ga[2] = gbuf_albedo.z, ga[3] = 1;
    }
};

} // namespace

extern "C" {

orc_scene* orc_scene_create(uint32_t n_objs, const uint32_t* descs, const uint32_t* tri_off,
                            const uint32_t* vert_off, const float* verts, const uint32_t* idx,
                            uint32_t n_lights, const uint32_t* lights, uint32_t n_tex,
                            const uint32_t* tex_info, const uint8_t* texels, const orc_bvh* bvh) {
    orc_scene* s = new orc_scene;
    s->n_objs = n_objs, s->n_lights = n_lights, s->n_tex = n_tex;
    s->descs = descs, s->tri_off = tri_off, s->vert_off = vert_off, s->idx = idx, s->lights = lights;
    s->tex_info = tex_info, s->verts = verts, s->texels = texels, s->bvh = bvh;
    return s;
}
void orc_scene_free(orc_scene* s) { delete s; }

void orc_record_rays(float* rays8, uint64_t capacity) {
    g_rec_buf = rays8, g_rec_cap = capacity;
    g_rec_count = 0;
}
uint64_t orc_recorded_rays(void) { return std::min<uint64_t>(g_rec_count.load(), g_rec_cap); }

/* Extension shared with the product (include/gpurt.h GpurtPipeParams::spatial_samples / spatial_radius; NOT part of the
 * reference): ReSTIR spatial reuse.  0 samples = rt.rgen as written.  Set before orc_render_frame. */
static uint32_t g_spatial_samples = 0;
static float g_spatial_radius = 16.0f;
void orc_render_set_spatial(uint32_t samples, float radius) { g_spatial_samples = samples, g_spatial_radius = radius; }

/* Extension shared with the product (GpurtPipeParams::light_sampling, NOT in the reference): 1 = light_sample /
 * light_sample_dir choose a light triangle with probability proportional to area x luma(emissive factor) and light_pdf
 * weights every triangle's term by that probability; 0 (default) = rt.rgen as written. */
static uint32_t g_light_sampling = 0;
void orc_render_set_light_sampling(uint32_t mode) { g_light_sampling = mode; }
/* the table: weight of triangle t of light l = 0.5 |(v1 - v0) x (v2 - v0)| * luma(emissive) on the world-space vertices
 * light_sample computes, summed in (light, triangle) order */
static void light_power_table(const orc_scene* S, const Consts& c, std::vector<float>& cdf, std::vector<uint32_t>& off) {
    Invocation inv;
    inv.S = S, inv.c = &c;
    const uint32_t nl = c.n_lights > 0 ? (uint32_t)c.n_lights : 0u;
    off.assign(nl + 1, 0);
    for(uint32_t l = 0; l < nl; l++) off[l + 1] = off[l] + inv.light(l)[9];
    cdf.assign(std::max<uint32_t>(off[nl], 1u), 0.0f);
    float run = 0.0f;
    for(uint32_t l = 0; l < nl; l++) {
        const uint32_t o_idx = inv.light(l)[8], n_tris = inv.light(l)[9];
        const float* m = inv.model(o_idx);
        const vec3 e = inv.desc_vec(o_idx, 36);
        for(uint32_t t = 0; t < n_tris; t++) {
            uint32_t ind[3];
            inv.tri_indices(o_idx, t, ind);
            const float *a = inv.vertex(o_idx, ind[0]), *b = inv.vertex(o_idx, ind[1]), *cc = inv.vertex(o_idx, ind[2]);
            vec3 v0 = xform_point(m, {a[0], a[1], a[2]}), v1 = xform_point(m, {b[0], b[1], b[2]}),
                 v2 = xform_point(m, {cc[0], cc[1], cc[2]});
            float w = 0.5f * length(cross(v1 - v0, v2 - v0)) * Invocation::luma(e);
            if(!(w > 0.0f) || w > 3.0e38f) w = 0.0f;
            run += w;
            cdf[off[l] + t] = run;
        }
    }
}

void orc_render_frame(const orc_scene* S, const uint32_t* consts, const uint32_t* camera, uint32_t w,
                      uint32_t h, uint32_t seed, float* image, const uint32_t* prev_res,
                      uint32_t* out_res, const float* ppos, const float* pnorm, const float* palb,
                      float* pos, float* norm, float* alb, uint64_t* ray_counts2, int threads) {
    Consts c;
    Camera cam;
    memcpy(&c, consts, sizeof(c));
    memcpy(&cam, camera, sizeof(cam));
    if(threads <= 0) threads = orc_hw_threads();
    uint32_t seed_val = seed ^ (uint32_t)c.frame;
    std::vector<float> lcdf;
    std::vector<uint32_t> lcdf_off;
    if(g_light_sampling == 1 && c.n_lights > 0) light_power_table(S, c, lcdf, lcdf_off);
    std::vector<std::thread> pool;
    std::vector<uint64_t> nc(threads, 0), na(threads, 0);
    for(int t = 0; t < threads; t++)
        pool.emplace_back([&, t] {
            for(uint32_t y = t; y < h; y += threads)
                for(uint32_t x = 0; x < w; x++) {
                    Invocation inv;
                    inv.S = S, inv.c = &c, inv.cam = &cam, inv.W = w, inv.H = h, inv.px = x, inv.py = y;
                    inv.prev_res_buf = prev_res, inv.ppos = ppos, inv.pnorm = pnorm, inv.palb = palb;
                    inv.spatial_samples = g_spatial_samples, inv.spatial_radius = g_spatial_radius;
                    if(!lcdf_off.empty()) inv.lcdf = lcdf.data(), inv.lcdf_off = lcdf_off.data(), inv.n_ltris = lcdf_off.back();
                    inv.main_(seed_val, image, out_res, pos, norm, alb);
                    nc[t] += inv.n_closest, na[t] += inv.n_any;
                }
        });
    for(auto& th : pool) th.join();
    if(ray_counts2) {
        ray_counts2[0] = ray_counts2[1] = 0;
        for(int t = 0; t < threads; t++) ray_counts2[0] += nc[t], ray_counts2[1] += na[t];
    }
}

void orc_camera_ray(const uint32_t* consts, const uint32_t* camera, uint32_t w, uint32_t h, uint32_t px,
                    uint32_t py, uint32_t s, uint32_t* rng, float* o3, float* d3) {
    Consts c;
    Camera cam;
    memcpy(&c, consts, sizeof(c));
    memcpy(&cam, camera, sizeof(cam));
    Invocation inv;
    inv.S = nullptr, inv.c = &c, inv.cam = &cam, inv.W = w, inv.H = h, inv.px = px, inv.py = py;
    inv.seed = *rng;
    vec3 d;
    inv.make_camera_ray(s, d);
    vec4 co = mul4(cam.iV, 0, 0, 0, 1);
    o3[0] = co.x, o3[1] = co.y, o3[2] = co.z;
    d3[0] = d.x, d3[1] = d.y, d3[2] = d.z;
    *rng = inv.seed;
}

static void fill_shade(const float* wo3, const float* n3, ShadeInfo& sh) {
    sh.wo = {wo3[0], wo3[1], wo3[2]};
    sh.N = {n3[0], n3[1], n3[2]};
    Invocation::make_tanspace(sh.N, sh.T, sh.B);
}
float orc_mat_pdf(int brdf, float roughness, const float* wo3, const float* n3, const float* wi3) {
    Consts c;
    memset(&c, 0, sizeof(c));
    c.brdf = brdf;
    Invocation inv;
    inv.c = &c;
    MatInfo m;
    m.albedo = v3(1), m.roughness = roughness;
    ShadeInfo sh;
    fill_shade(wo3, n3, sh);
    return inv.MAT_pdf(m, sh, {wi3[0], wi3[1], wi3[2]});
}
void orc_mat_eval(int brdf, const float* albedo3, float roughness, const float* wo3, const float* n3,
                  const float* wi3, float* out3) {
    Consts c;
    memset(&c, 0, sizeof(c));
    c.brdf = brdf;
    Invocation inv;
    inv.c = &c;
    MatInfo m;
    m.albedo = {albedo3[0], albedo3[1], albedo3[2]}, m.roughness = roughness;
    ShadeInfo sh;
    fill_shade(wo3, n3, sh);
    vec3 r = inv.MAT_eval(m, sh, {wi3[0], wi3[1], wi3[2]});
    out3[0] = r.x, out3[1] = r.y, out3[2] = r.z;
}

/* the oracle's texture unit for the compiled reference shader (oracle/ref_shim/glsl_ref.cpp): texture() is hardware
 * behaviour (R8G8B8A8_SRGB decode, linear filter, REPEAT), not shader text */
void orc_texture_fetch(const orc_scene* S, int tex, float u, float v, float* rgb) {
    Invocation inv;
    inv.S = S;
    const float tc[2] = {u, v};
    vec3 c = inv.texture(tex, tc);
    rgb[0] = c.x, rgb[1] = c.y, rgb[2] = c.z;
}

/* One call of one restated rtcommon.glsl / restir.glsl function, by id — the same ids and argument layout as
 * ref_glsl_unit in oracle/ref_shim/glsl_ref.cpp, which runs the reference's own shader text compiled as C++.
 * in: floats, u[0]: RNG state (in / out), u[1], u[2]: unsigned arguments / results, out: floats. */
void orc_glsl_unit(int fn, const float* in, uint32_t* u, float* out) {
    Consts c;
    memset(&c, 0, sizeof(c));
    Invocation inv;
    inv.S = nullptr, inv.c = &c, inv.cam = nullptr;
    inv.seed = u[0];
    auto V = [&](int k) { return vec3{in[k], in[k + 1], in[k + 2]}; };
    auto put = [&](int k, vec3 v) { out[k] = v.x, out[k + 1] = v.y, out[k + 2] = v.z; };
    MatInfo mat;
    ShadeInfo sh;
    auto shading = [&]() {
        c.brdf = (int)in[0];
        mat.roughness = in[1];
        mat.albedo = V(2), mat.emissive = v3(0), mat.tanspaceNormal = v3(0), mat.use_tanspace = false;
        sh.wo = V(5), sh.N = V(8);
        Invocation::make_tanspace(sh.N, sh.T, sh.B);
    };
    switch(fn) {
    case 0: u[0] = orc_tea(u[1], u[2]); return;
    case 1: out[0] = inv.randf(); break;
    case 2: u[1] = inv.randu(u[1], u[2]); break;
    case 3: put(0, inv.cospow_hemisphere(in[0], V(1), V(4), V(7))); break;
    case 4: put(0, inv.triangle_sample()); break;
    case 5: {
        vec3 hitp = v3(0);
        out[0] = Invocation::triangle_hit(V(0), V(3), V(6), V(9), V(12), hitp) ? 1.0f : 0.0f;
        put(1, hitp);
        break;
    }
    case 6: out[0] = Invocation::triangle_pdf(V(0), V(3), V(6), V(9), V(12)); break;
    case 7: {
        vec3 t, b;
        Invocation::make_tanspace(V(0), t, b);
        put(0, t), put(3, b);
        break;
    }
    case 8: out[0] = Invocation::hit_bbox(V(0), V(3), V(6), V(9)) ? 1.0f : 0.0f; break;
    case 9: shading(), out[0] = inv.MAT_pdf(mat, sh, V(11)); break;
    case 10: shading(), put(0, inv.MAT_eval(mat, sh, V(11))); break;
    case 11: {
        shading();
        vec3 wi = v3(0);
        out[0] = inv.MAT_sample(mat, sh, wi) ? 1.0f : 0.0f;
        put(1, wi);
        break;
    }
    case 12: {
        Reservoir r;
        r.pos = V(0), r.normal = V(3), r.emissive = V(6), r.w_sum = in[9], r.w = in[10], r.n_seen = u[1];
        inv.res_update(r, in[11], V(12), V(15), V(18));
        put(0, r.pos), put(3, r.normal), put(6, r.emissive), out[9] = r.w_sum, out[10] = r.w;
        u[1] = r.n_seen;
        break;
    }
    case 13: out[0] = Invocation::power_heuristic(in[0], in[1]); break;
    case 14: out[0] = Invocation::luma(V(0)); break;
    case 15: /* hammersley(i, N) as make_camera_ray evaluates it (rtcommon.glsl:128-139) */
        out[0] = (float)u[1] / (float)u[2], out[1] = orc_radical_inverse(u[1]);
        break;
    default: break;
    }
    u[0] = inv.seed;
}

/* tonemap.frag:17-48, then the R8G8B8A8_SRGB framebuffer encode (gpurt.cpp:176, :258-262) */
void orc_tonemap(const float* rgba, uint64_t n, int op, float exposure, float gamma, uint8_t* out) {
    auto u2 = [](float x) {
        const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
        return ((x * (A * x + C * B) + D * E) / (x * (A * x + B) + D * F)) - E / F;
    };
    float white = 1.0f / u2(11.2f);
    float ig = 1.0f / gamma;
    for(uint64_t i = 0; i < n; i++) {
        float o[4] = {rgba[4 * i], rgba[4 * i + 1], rgba[4 * i + 2], rgba[4 * i + 3]};
        for(int k = 0; k < 3; k++) {
            float x = o[k];
            if(op == 0) x = dm_pow(u2(x * exposure) * white, ig);
            else if(op == 1) x = dm_pow(1.0f - dm_exp(-x * exposure), ig);
            o[k] = x;
        }
        for(int k = 0; k < 4; k++) {
            float x = o[k];
            x = x != x ? 0.0f : fminf(fmaxf(x, 0.0f), 1.0f);
            if(k < 3) x = x <= 0.0031308f ? 12.92f * x : 1.055f * dm_pow(x, 1.0f / 2.4f) - 0.055f;
            out[4 * i + k] = (uint8_t)(x * 255.0f + 0.5f);
        }
    }
}

} /* extern "C" */
