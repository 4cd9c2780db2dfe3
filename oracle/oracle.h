/*
 * oracle.h — CPU oracle for the GPU-RT hot path (TEST INFRASTRUCTURE, NOT PRODUCT).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library. The product (libgpurt.so) never links or calls it.
 *
 * PARITY STATUS: the reference's BVH build and traversal live inside the NVIDIA Vulkan
 * driver (src/vk/vulkan.cpp:877 vkCmdBuildAccelerationStructuresKHR, src/shaders/rt/rt.rgen:258
 * traceRayEXT) and its closest-point query lives on a branch that is not in the snapshot
 * (README.md:6-8).  The reference ships no tests or golden vectors.  For those three pieces this
 * oracle is "parity unpinned": it restates the *observable contract* at the reference's own call
 * sites (ray interval, flags, payload, instance transforms) with a documented fp32 numeric
 * contract.  Everything that IS reference source is restated line by line and pinned
 * against that source compiled in oracle/_ref: host matrices, camera, scene loading / packing and
 * texture decode against the reference's C++ (libgpurt_ref.so); RNG, sampling, BRDFs, light
 * sampling, the five integrators, ReSTIR, accumulation, debug views and the tonemap against the
 * reference's GLSL compiled as C++ (libglsl_ref.so) — bit for bit, whole frames included
 * (tests/golden/glsl_*_golden.*, tests/test_oracle.py).
 *
 * All functions are extern "C", plain pointers and sizes, so tests drive them with ctypes.
 */
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- layouts (little-endian fp32 / u32) -------------------------------------------------------
 * world triangle : 9 floats  v0.xyz v1.xyz v2.xyz            (gid = index in this array)
 * ray            : 8 floats  o.xyz tmin d.xyz tmax            (rt.rgen:257-270: tmin=1e-5,tmax=1e7)
 * hit            : 4 words   t u v gid    miss: t=+inf,u=v=0,gid=0xFFFFFFFF   (rt.rchit:11-16)
 * query          : 4 floats  p.xyz r2max
 * cpq result     : 8 words   c.xyz dist gid obj(unused here =0) u v ; none: dist=+inf,gid=~0
 */

/* rtcommon.glsl:99-124 */
uint32_t orc_tea(uint32_t v0, uint32_t v1);
uint32_t orc_lcg(uint32_t* state);
float orc_randf(uint32_t* state);
/* rtcommon.glsl:128-136 */
float orc_radical_inverse(uint32_t bits);

/* N1: world-space flattening of one object (rt.rgen:172-174 `vec3(model * vec4(v,1))`,
 * instance transform vulkan.cpp:794-799). verts: 48-byte stride (mesh.h:16-22). */
void orc_flatten_object(const float* verts48, const uint32_t* idx, uint32_t n_tris,
                        const float model[16], float* out_tris9);

/* N3: one ray / one triangle. returns 1 on hit and writes t,u,v. */
int orc_intersect(const float ray[8], const float tri9[9], float* t, float* u, float* v);
/* N5: one point / one triangle: closest point c, barycentrics (v,w), returns d^2. */
float orc_closest_point_tri(const float p[3], const float tri9[9], float c[3], float* v, float* w);

/* brute-force references, O(n_tris * n): the ground truth for every BVH path */
void orc_closest_hit_brute(const float* tris9, uint32_t n_tris, const float* rays, uint64_t n,
                           uint32_t* hits4, int threads);
void orc_any_hit_brute(const float* tris9, uint32_t n_tris, const float* rays, uint64_t n,
                       uint8_t* occluded, int threads);
void orc_closest_point_brute(const float* tris9, uint32_t n_tris, const float* queries,
                             uint64_t n, uint32_t* results8, int threads);

/* N6: canonical primitive order: 63-bit Morton of AABB centroid, stable sort (ties by gid). */
typedef struct orc_bvh orc_bvh;
orc_bvh* orc_bvh_build(const float* tris9, uint32_t n_tris);
/* N6': the default ("fast trace") build's order and topology: top-down 16-bin SAH split (definition restated in oracle.cpp);
 * morton_keys then returns positions 0..n-1.  The same query functions work on either tree. */
orc_bvh* orc_bvh_build_sah(const float* tris9, uint32_t n_tris);
void orc_bvh_free(orc_bvh*);
uint32_t orc_bvh_n_tris(const orc_bvh*);
void orc_bvh_scene_box(const orc_bvh*, float out6[6]);
void orc_bvh_morton_keys(const orc_bvh*, uint64_t* out_sorted_keys);
void orc_bvh_prim_order(const orc_bvh*, uint32_t* out_order);
/* Karras-2012 binary hierarchy over the sorted keys: n-1 internal nodes.
 * child encoding: >=0 internal index, <0 leaf ~sorted_position. boxes are the exact (uninflated)
 * union of triangle AABBs: 6 floats min.xyz max.xyz per internal node. */
void orc_bvh_get_bvh2(const orc_bvh*, int32_t* left, int32_t* right, float* boxes6);
float orc_bvh_inflation(const orc_bvh*);

/* CPU BVH traversal (the timed "host path") — must equal brute force exactly */
void orc_bvh_closest_hit(const orc_bvh*, const float* rays, uint64_t n, uint32_t* hits4, int threads);
void orc_bvh_any_hit(const orc_bvh*, const float* rays, uint64_t n, uint8_t* occluded, int threads);
void orc_bvh_closest_point(const orc_bvh*, const float* queries, uint64_t n, uint32_t* results8,
                           int threads);

/* SURVEY §8d config-1 generators (reference RNG) */
void orc_gen_random_rays(uint64_t n, uint32_t seed, const float box6[6], float inflate_frac,
                         float tmin, float tmax, float* rays);
void orc_gen_random_points(uint64_t n, uint32_t seed, const float box6[6], float inflate_frac,
                           float r2, float* queries);

int orc_hw_threads(void);

#ifdef __cplusplus
}
#endif
