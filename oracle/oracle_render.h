/*
 * oracle_render.h — CPU restatement of the reference's ray-generation shader (TEST INFRASTRUCTURE).
 * Follows src/shaders/rt/rt.rgen, rtcommon.glsl and restir.glsl function by function; see
 * oracle_render.cpp for the line citations and DESIGN.md §3 (N8) for the fp32 contract.
 */
#pragma once
#include <stdint.h>

#include "oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_scene orc_scene;

/* All arrays are borrowed (must outlive the scene). Layouts = include/gpurt.h:
 * descs: 52 words per object (GpurtSceneDesc), lights: 12 words (GpurtSceneLight),
 * verts: 12 floats per vertex (Mesh::Vertex), idx: object-local, tri_off/vert_off: n_objs+1.
 * tex_info: per texture {texel offset, w, h, 0}; texels RGBA8. bvh: over the world triangles. */
orc_scene* orc_scene_create(uint32_t n_objs, const uint32_t* descs, const uint32_t* tri_off,
                            const uint32_t* vert_off, const float* verts, const uint32_t* idx,
                            uint32_t n_lights, const uint32_t* lights, uint32_t n_tex,
                            const uint32_t* tex_info, const uint8_t* texels, const orc_bvh* bvh);
void orc_scene_free(orc_scene*);

/* One invocation of rt.rgen `main` per pixel (rt.rgen:567-677).
 * consts: GpurtConstants (22 words, frame already incremented as RTPipe::trace does);
 * camera: GpurtCamera (82 words). seed: replaces clockARB(): tea(pixel, seed ^ frame).
 * image: RGBA32F in/out (progressive accumulation, rt.rgen:638-645).
 * reservoirs: 12 words each {pos.xyz,w_sum, normal.xyz,w, emissive.xyz,n_seen}; prev_* are the
 * previous frame's buffers (read), out_* this frame's (written). G-buffers RGBA32F. */
/* Extension shared with the product (GpurtPipeParams::spatial_samples / spatial_radius, NOT in the reference): ReSTIR spatial
 * reuse for the following orc_render_frame calls; 0 samples (the default) = rt.rgen as written. */
void orc_render_set_spatial(uint32_t samples, float radius);
/* Extension shared with the product (GpurtPipeParams::light_sampling, NOT in the reference): 1 = light triangles chosen in
 * proportion to area x luma(emissive), light_pdf weighted to match; 0 (default) = rt.rgen as written. */
void orc_render_set_light_sampling(uint32_t mode);
void orc_render_frame(const orc_scene* S, const uint32_t* consts, const uint32_t* camera, uint32_t w,
                      uint32_t h, uint32_t seed, float* image, const uint32_t* prev_res,
                      uint32_t* out_res, const float* ppos, const float* pnorm, const float* palb,
                      float* pos, float* norm, float* alb, uint64_t* ray_counts2, int threads);

/* Record the closest-hit rays (8 floats each, traversal order of the worker threads) traced by the
 * following orc_render_frame calls into rays8[capacity]; NULL stops recording. */
void orc_record_rays(float* rays8, uint64_t capacity);
uint64_t orc_recorded_rays(void);

/* unit-level entry points for tests (each restates one GLSL function) */
void orc_camera_ray(const uint32_t* consts, const uint32_t* camera, uint32_t w, uint32_t h, uint32_t px,
                    uint32_t py, uint32_t s, uint32_t* rng, float* o3, float* d3);
float orc_mat_pdf(int brdf, float roughness, const float* wo3, const float* n3, const float* wi3);
void orc_mat_eval(int brdf, const float* albedo3, float roughness, const float* wo3, const float* n3,
                  const float* wi3, float* out3);
void orc_texture_fetch(const orc_scene* S, int tex, float u, float v, float* rgb);
/* one restated rtcommon.glsl / restir.glsl function by id (ids and layout: oracle/ref_shim/glsl_ref.cpp, which runs the
 * reference's own shader text): the hook tests/test_oracle.py pins against tests/golden/glsl_unit_golden.npz */
void orc_glsl_unit(int fn, const float* in, uint32_t* u, float* out);
/* tonemap.frag:17-48 on RGBA32F -> RGBA8 (op 0 Uncharted2, 1 exponential, 2 passthrough) */
void orc_tonemap(const float* rgba, uint64_t n, int op, float exposure, float gamma, uint8_t* out);

#ifdef __cplusplus
}
#endif
