/*
 * gpurt.h — C ABI of libgpurt.so: the B200-native replacement for GPU-RT's ray-tracing hot path.
 *
 * Drop-in boundary.  Each entry point names the reference interface it replaces;
 * citations are relative to the reference root.  The reference calls these interfaces as C++
 * classes in namespace VK from GPURT (src/gpurt.cpp:39-45, :216-241); the shim that re-creates
 * VK::Accel / VK::RTPipe on top of this header is gpu-rt_b200/host/rtpipe.h and the maintainer-side
 * binding is shown in INTEGRATION.md.
 *
 * Conventions: every function returns 0 on success and a negative GPURT_E_* code on failure, never
 * exits (the reference's VK_CHECK -> die() -> exit(), src/vk/vulkan.h:26-33, is replaced by error
 * codes + gpurt_last_error()).  Handles are opaque.  One gpurt_ctx per GPU, used from one host
 * thread at a time.  `mem` says where ray/query/result buffers live.  GPURT_MEM_HOST: the call returns when the results
 * are in the caller's array; page-locked arrays (cudaHostAlloc / cudaHostRegister) are read and written in place by
 * the kernels over PCIe (one launch, no staging), pageable ones go through a chunked copy pipeline.
 * GPURT_MEM_DEVICE buffers are used in place and the call is asynchronous on the context's stream.  There is NO CPU fallback: compute entry points fail
 * with GPURT_E_NO_DEVICE when no B200 is present.
 */
#ifndef GPURT_H
#define GPURT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPURT_OK 0
#define GPURT_E_INVALID (-1)   /* bad argument */
#define GPURT_E_NO_DEVICE (-2) /* CUDA device missing / wrong ordinal */
#define GPURT_E_CUDA (-3)      /* CUDA runtime error, see gpurt_last_error() */
#define GPURT_E_IO (-4)        /* scene file could not be read / parsed */
#define GPURT_E_STATE (-5)     /* call order (e.g. render before accel build) */

#define GPURT_MEM_HOST 0
#define GPURT_MEM_DEVICE 1

#define GPURT_NO_HIT 0xFFFFFFFFu

typedef struct gpurt_ctx gpurt_ctx;
typedef struct gpurt_scene gpurt_scene;
typedef struct gpurt_accel gpurt_accel;
typedef struct gpurt_pipe gpurt_pipe;

/* ---- POD records ----------------------------------------------------------------------------- */

/* traceRayEXT arguments (src/shaders/rt/rt.rgen:257-270): origin, tmin, direction, tmax. 32 B. */
typedef struct GpurtRay {
    float o[3], tmin;
    float d[3], tmax;
} GpurtRay;

/* Ray_Payload (src/shaders/rt/rtcommon.glsl:55-60) as written by rt.rchit:11-16 / rt.rmiss:9-11.
 * barycentrics = (1-u-v, u, v).  prim is the GLOBAL triangle id: objects in Scene order
 * (Scene::for_objs), triangles in index-buffer order; gpurt_scene_tri_offsets() maps it back to
 * (gl_InstanceCustomIndexEXT, gl_PrimitiveID).  Miss: t=+inf, prim=GPURT_NO_HIT. 16 B. */
typedef struct GpurtHit {
    float t, u, v;
    uint32_t prim;
} GpurtHit;

/* Closest-point query (FCPW-GPU branch, README.md:6-8): point + squared search radius. 16 B. */
typedef struct GpurtQuery {
    float p[3], r2;
} GpurtQuery;

/* Closest point, its distance, the triangle (global id, object id) and barycentrics (weights of
 * vertex 1 and 2).  Nothing within the radius: dist=+inf, prim=GPURT_NO_HIT. 32 B. */
typedef struct GpurtClosestPoint {
    float p[3], dist;
    uint32_t prim, obj;
    float u, v;
} GpurtClosestPoint;

/* Material (src/scene/material.h:15-21), packed like Scene_Desc's tail (src/vk/rt.h:70-77). */
typedef struct GpurtMaterial {
    float albedo[3];
    int32_t albedo_tex;
    float emissive[3];
    int32_t emissive_tex;
    float metal_rough[2];
    int32_t metal_rough_tex;
    int32_t normal_tex;
} GpurtMaterial;

/* Scene_Desc (src/vk/rt.h:67-78, GLSL Scene_Obj rtcommon.glsl:10-21). 208 B. */
typedef struct GpurtSceneDesc {
    float model[16], modelIT[16]; /* column-major */
    float albedo[4], emissive[4], metal_rough[4];
    int32_t albedo_tex, emissive_tex, metal_rough_tex, normal_tex;
    uint32_t index;
    uint32_t _pad[3];
} GpurtSceneDesc;

/* Scene_Light (src/vk/rt.h:79-84). 48 B. */
typedef struct GpurtSceneLight {
    float bmin[4], bmax[4];
    uint32_t index, n_triangles;
    uint32_t _pad[2];
} GpurtSceneLight;

/* RTPipe_Constants (src/vk/rt.h:85-102; GLSL push constants rtcommon.glsl:77-95). 88 B.
 * `frame` is maintained by the pipe exactly like RTPipe::trace (src/vk/rt.cpp:353-368);
 * n_lights / n_objs are filled from the scene.  `seed` replaces clockARB() in
 * rt.rgen:569: the per-pixel stream is tea(pixel, seed ^ frame). */
typedef struct GpurtConstants {
    float clear_col[4];
    float env_light[4];
    int32_t frame, samples, max_frame, qmc, max_depth, use_normal_map, use_metalness, use_temporal,
        integrator, brdf, debug_view, use_rr, n_lights, n_objs;
} GpurtConstants;

/* UBO (src/vk/rt.h:104-117): camera matrices + ReSTIR constants, all column-major. */
typedef struct GpurtCamera {
    float V[16], P[16], iV[16], iP[16];
    float prev_PV[16];
    uint32_t new_samples, temporal_multiplier;
} GpurtCamera;

/* RTPipe public tunables (src/vk/rt.h:38-53), same defaults. */
typedef struct GpurtPipeParams {
    int32_t max_frames, samples_per_frame, max_depth;
    float clear[3], env[3], env_scale;
    int32_t use_normal_map, use_rr, use_metalness, use_qmc, use_temporal, integrator,
        temporal_scale, brdf, debug_view, res_samples;
    uint32_t seed; /* SURVEY Q1 */
    /* Extension, off by default (no reference counterpart: todo.txt:10-12 "work on ReSTIR"): ReSTIR
     * spatial reuse.  After the temporal combination of the frame's first sample the pixel's reservoir is combined with
     * `spatial_samples` reservoirs of the PREVIOUS frame taken from a disc of `spatial_radius` pixels around it (neighbours
     * whose previous-frame normal / plane disagree are skipped), and the surviving sample gets one visibility ray.  Changes
     * the estimator (less noise per frame, slightly biased like any unshadowed-neighbour reuse), hence a flag. */
    int32_t spatial_samples;
    float spatial_radius;
    /* Extension, off by default (SURVEY §8f rank 4 "light-sampling acceleration ... only behind a flag"): 0 = the reference's
     * light_sample / light_sample_dir (uniform light, uniform triangle, rt.rgen:151-220); 1 = one light triangle chosen with
     * probability proportional to area x luma(emissive factor) from a table of running sums (one random number, binary
     * search), with light_pdf (rt.rgen:222-255) weighting every triangle's term by the same probability, so direct, MIS and
     * ReSTIR estimators stay unbiased and small or dim emitters stop getting as many samples as large bright ones. */
    int32_t light_sampling;
} GpurtPipeParams;

typedef struct GpurtAccelInfo {
    uint32_t n_tris, n_objs;
    uint32_t n_bvh2_nodes;       /* n_tris-1 */
    uint32_t n_wide_nodes;       /* 80-byte 8-wide nodes */
    uint32_t wide_depth;         /* levels of the wide tree */
    float scene_min[3], scene_max[3];
    float inflation;             /* conservative AABB padding, DESIGN.md §3 N7 */
    float build_ms;              /* device time of the last build */
    uint64_t node_bytes, tri_bytes;
    float tree_cost;             /* sum of the binary inner nodes' box areas (SAH cost up to constants) */
    float tree_cost_at_build;    /* ... as of the last full build */
    uint32_t refits;             /* gpurt_accel_refit calls since the last full build */
    uint32_t reserved;
} GpurtAccelInfo;

/* traversal statistics from the instrumented kernel (N_node, N_tri per ray) */
typedef struct GpurtTraceStats {
    uint64_t rays, nodes_visited, tris_tested, hits;
} GpurtTraceStats;

/* ---- errors ---------------------------------------------------------------------------------- */
const char* gpurt_last_error(void);
const char* gpurt_version(void);

/* ---- context: replaces the VK::Manager singleton (src/vk/vulkan.cpp:17-20, :1136-1160) -------- */
int gpurt_ctx_create(int device_ordinal, gpurt_ctx** out);
int gpurt_ctx_destroy(gpurt_ctx* ctx);
/* Launch on a caller-owned cudaStream_t (e.g. torch's current stream). NULL is CUDA's legacy default
 * stream (what torch uses unless told otherwise); GPURT_STREAM_OWN goes back to the context's own. */
#define GPURT_STREAM_OWN ((void*)(intptr_t)-1)
int gpurt_ctx_set_stream(gpurt_ctx* ctx, void* cuda_stream);
int gpurt_ctx_synchronize(gpurt_ctx* ctx);

/* ---- scene: replaces Scene (src/scene/scene.h:18-42) + RTPipe::build_desc (src/vk/rt.cpp:26-76) */
/* ctx may be NULL for host-only use (loading / packing); such a scene cannot be built. */
int gpurt_scene_create(gpurt_ctx* ctx, gpurt_scene** out);
int gpurt_scene_destroy(gpurt_scene* scene);
/* Scene::load (src/scene/scene.cpp:317-391): glTF / GLB -> objects (one per primitive) + textures;
 * `scale` is Scene::scale (scene.h:42). Object order = Scene::for_objs order. */
int gpurt_scene_load_gltf(gpurt_scene* scene, const char* path, float scale);
/* Procedural stand-in for media/sponza (Sponza.bin is missing from the reference snapshot):
 * same triangle count (262,267), object count (103) and object-space extent. */
int gpurt_scene_make_sponza_standin(gpurt_scene* scene);
/* Object(id,pose,mesh,material) (src/scene/object.h:15): verts are Mesh::Vertex, 48-byte stride
 * (src/vk/mesh.h:16-22); model is scale*pose.transform() column-major (src/gpurt.cpp:228-231). */
int gpurt_scene_add_object(gpurt_scene* scene, const void* verts48, uint32_t n_verts,
                           const uint32_t* indices, uint32_t n_indices, const float model[16],
                           const GpurtMaterial* material, uint32_t* out_obj_index);
/* Pose edit (GPURT::edit_scene -> rebuild_tlas, src/gpurt.cpp:286-289, :378-385): replaces the model
 * matrix of object `obj` (index in Scene::for_objs order).  Follow with gpurt_accel_update(). */
int gpurt_scene_set_transform(gpurt_scene* scene, uint32_t obj, const float model[16]);
/* Material edit (the ImGui material editor, src/gpurt.cpp:378-385, followed by rt_pipe.recreate(scene) -> build_desc,
 * src/vk/rt.cpp:26-76): replaces the material of object `obj`.  Follow with gpurt_accel_sync_scene() — the BVH is unaffected. */
int gpurt_scene_set_material(gpurt_scene* scene, uint32_t obj, const GpurtMaterial* material);
/* Object order.  0 (default): the iteration order of the reference's std::unordered_map<id, Object>, i.e. what
 * Scene::for_objs gives for ids handed out from 1.  1: insertion order — for callers that already iterate the
 * reference's Scene themselves and pass instance i as the i-th add_object call (GPURT::build_accel's BLAS / BLAS_T). */
int gpurt_scene_set_ordered(gpurt_scene* scene, int ordered);
/* RTPipe::build_textures starts from an empty table (rt.cpp:432-434). */
int gpurt_scene_clear_textures(gpurt_scene* scene);
/* RTPipe::build_textures (src/vk/rt.cpp:430-455): RGBA8, sampled as sRGB, linear, repeat. */
int gpurt_scene_add_texture(gpurt_scene* scene, const uint8_t* rgba8, uint32_t w, uint32_t h,
                            int32_t* out_tex_index);
/* Texture `tex` as uploaded (RGBA8, row-major): sizes first (rgba8_out may be NULL), then the texels. */
int gpurt_scene_get_texture(const gpurt_scene* scene, uint32_t tex, uint32_t* w, uint32_t* h, uint8_t* rgba8_out);
int gpurt_scene_counts(const gpurt_scene* scene, uint32_t* n_objs, uint32_t* n_tris,
                       uint32_t* n_lights, uint32_t* n_textures);
/* n_objs+1 prefix offsets: global prim id -> (object, primitive). */
int gpurt_scene_tri_offsets(const gpurt_scene* scene, uint32_t* out_offsets);
int gpurt_scene_get_descs(const gpurt_scene* scene, GpurtSceneDesc* out_descs);
int gpurt_scene_get_lights(const gpurt_scene* scene, GpurtSceneLight* out_lights);
int gpurt_scene_object_sizes(const gpurt_scene* scene, uint32_t obj, uint32_t* n_verts,
                             uint32_t* n_indices);
int gpurt_scene_get_object(const gpurt_scene* scene, uint32_t obj, void* out_verts48,
                           uint32_t* out_indices);

/* ---- camera: Camera (src/util/camera.cpp:58-71, :150-155) + RTPipe::update_uniforms ------------ */
/* mode 0: Camera::reset() defaults; mode 1: look_at(center,pos) + vfov. Fills V,P,iV,iP. */
int gpurt_camera_make(int mode, float width, float height, const float pos[3],
                      const float center[3], float vfov_deg, GpurtCamera* out);

/* ---- acceleration structure: replaces VK::Accel (src/vk/vulkan.h:256-284) ---------------------- */
/* = every BLAS build (src/vk/vulkan.cpp:881-936) + the TLAS build (:777-856) of GPURT::build_accel (src/gpurt.cpp:220-241)
 * in one call: a binary BVH over the world-space triangles of all instances, collapsed to 80-byte 8-wide compressed nodes.
 *
 * GPURT_BUILD_DEFAULT — what the reference asks its driver for (PREFER_FAST_TRACE, vulkan.cpp:780, :884): primitive order
 * and binary topology from a top-down 16-bin surface-area-heuristic split of the triangle boxes, built on the device
 * (csrc/sah_build.cu; the tree is defined in gpu-rt_b200/host/sah_split.h and restated in oracle/oracle.cpp).  Measured on
 * media/cbox against the Morton build: 36 % fewer node visits per random ray, random rays +30 %, shadow rays +36 %,
 * camera rays +18 %, closest points +14 %; build 1.8x (16.7 k triangles: 1.06 vs 0.60 ms; 262 k: 1.84 vs 0.81 ms).
 * gpurt_accel_get_morton_keys returns positions 0..n-1 for this build. */
#define GPURT_BUILD_DEFAULT 0u
#define GPURT_BUILD_KEEP_BVH2 1u /* accepted; the binary tree is always kept (gpurt_accel_update rebuilds into it) */
/* SAH-optimal wide collapse (dynamic programming over the binary tree, Ylitie-Karras-Laine 2017) instead of the
 * greedy largest-area rule.  Same query results.  Measured on the Sponza stand-in: 31 % fewer wide nodes, primary
 * rays +5 %, bounce rays -3 %, closest-point queries -15 %, build +15..40 % — an option for camera-ray-heavy use. */
#define GPURT_BUILD_SAH_COLLAPSE 2u
#define GPURT_BUILD_SAH_SPLIT 4u /* accepted: the SAH split is the default build (round 1 had it behind this flag) */
/* Morton-order LBVH instead of the SAH split (63-bit keys of the AABB centroids, radix sort, Karras 2012): the fastest
 * build (PREFER_FAST_BUILD), for scenes rebuilt every frame and for regular procedural geometry, where the Morton order
 * is as good (Sponza stand-in: rays +6 % with this flag, closest points -20 %).  Same query results. */
#define GPURT_BUILD_LBVH 8u
int gpurt_accel_build(gpurt_scene* scene, uint32_t flags, gpurt_accel** out);
/* Rebuild after scene edits (GPURT::build_accel with rebuild_tlas / rebuild_blas, src/gpurt.cpp:220-241).
 * Pose-only edits re-upload the 208-byte Scene_Desc records and rebuild on the device in the buffers
 * the accel already owns (no geometry upload, no allocation); geometry edits fall back to a full build.
 * Pipes created on this accel stay valid (call gpurt_pipe_reset_frame, like GPURT::build_accel does). */
int gpurt_accel_update(gpurt_accel* accel);
/* Pose-only edit (GPURT::edit_scene sets rebuild_tlas and nothing else, src/gpurt.cpp:286-289; the reference then rebuilds
 * only the TLAS over the unchanged BLASes, :228-237): keep the primitive order and the binary topology, flatten the
 * triangles under the new instance matrices, refit the node boxes, collapse and encode the wide tree again — no sort, no
 * SAH split (262 k triangles: about 0.6 ms instead of 1.8 ms).  Query results are those of a fresh build (any valid tree
 * gives the same hits); what degrades when objects move far is the tree's quality, see tree_cost in GpurtAccelInfo.
 * GPURT_E_STATE if geometry changed since the last build. */
int gpurt_accel_refit(gpurt_accel* accel);
/* gpurt_accel_refit when only poses changed and the refitted tree's cost stays within max_cost_growth (<= 0: 1.25) times the
 * cost at the last full build; gpurt_accel_update otherwise.  What the drop-in shim calls for TLAS->recreate. */
int gpurt_accel_update_auto(gpurt_accel* accel, float max_cost_growth);
/* Bring the device copy of Scene_Desc / Scene_Light / textures up to date after material or texture edits WITHOUT
 * rebuilding the BVH (= rt_pipe.recreate(scene) with an unchanged TLAS).  Geometry or pose edits need gpurt_accel_update. */
int gpurt_accel_sync_scene(gpurt_accel* accel);
int gpurt_accel_destroy(gpurt_accel* accel);
int gpurt_accel_info(const gpurt_accel* accel, GpurtAccelInfo* out);
/* Canonical primitive order (sorted Morton position -> global prim id), host buffers. */
int gpurt_accel_get_prim_order(const gpurt_accel* accel, uint32_t* out_order);
int gpurt_accel_get_morton_keys(const gpurt_accel* accel, uint64_t* out_sorted_keys);
/* n_tris-1 internal nodes: child >= 0 internal, < 0 leaf ~sorted_position; boxes 6 floats. */
int gpurt_accel_get_bvh2(const gpurt_accel* accel, int32_t* left, int32_t* right, float* boxes6);

/* ---- queries ---------------------------------------------------------------------------------- */
/* Processing order: device batches of >= 2^20 rays / points on scenes whose BVH exceeds 64 MB are keyed by the Morton
 * code of the ray origin / query point; unless neighbouring elements already share a cell they are processed through
 * a radix-sorted index (10 M triangles, random input: rays 1.3x, points 1.5x).  Results land at their original
 * positions and never depend on the processing order; such a call synchronises the stream once (coherence counter).
 * GPURT_SPATIAL_ORDER=0 in the environment disables it. */
/* traceRayEXT closest hit (rt.rgen:257-270 + rt.rchit + rt.rmiss) for a batch of rays. */
int gpurt_trace_closest(gpurt_accel* accel, const GpurtRay* rays, uint64_t n, GpurtHit* hits, int mem);
/* `visibility` (rt.rgen:272-291): 1 = occluded. */
int gpurt_trace_any(gpurt_accel* accel, const GpurtRay* rays, uint64_t n, uint8_t* occluded, int mem);
/* FCPW closest-point query (README.md:6-8) over the same BVH. */
int gpurt_closest_points(gpurt_accel* accel, const GpurtQuery* queries, uint64_t n,
                         GpurtClosestPoint* results, int mem);
/* Same as gpurt_trace_closest but through the binary LBVH (debug / cross-check path). */
int gpurt_trace_closest_bvh2(gpurt_accel* accel, const GpurtRay* rays, uint64_t n, GpurtHit* hits, int mem);
/* Instrumented closest-hit: same results + visit counters (device buffers only). */
int gpurt_trace_closest_stats(gpurt_accel* accel, const GpurtRay* rays, uint64_t n, GpurtHit* hits,
                              GpurtTraceStats* out_host_stats);
/* Instrumented closest-point query: visit counters only (device queries; `rays` = queries, `hits` = queries answered). */
int gpurt_closest_points_stats(gpurt_accel* accel, const GpurtQuery* queries, uint64_t n, GpurtTraceStats* out_host_stats);
/* Device time (ms) of the last query / render call on this accel's context (CUDA events). */
int gpurt_last_kernel_ms(gpurt_ctx* ctx, float* out_ms);

/* ---- multi-GPU result placement (no reference counterpart — GPU-RT is single-GPU) --- */
/* One process per GPU, scene replicated, queries sharded.  Instead of a gather collective after the
 * query, the rank that owns the full result array exports it once; every other rank maps it over
 * NVLink and passes `mapped + first_element` as the device `hits` / `results` pointer of its query
 * call, so the traversal kernel's own result stores land in the owner's HBM while it computes.
 * The handle is 64 opaque bytes (cudaIpcMemHandle_t); ship it with any byte transport. */
#define GPURT_IPC_HANDLE_BYTES 64
/* owner rank: a dedicated device allocation + its handle */
int gpurt_shared_alloc(gpurt_ctx* ctx, uint64_t bytes, void** out_device_ptr,
                       uint8_t handle_out[GPURT_IPC_HANDLE_BYTES]);
int gpurt_shared_free(gpurt_ctx* ctx, void* device_ptr);
/* other ranks (other processes): map / unmap the owner's allocation */
int gpurt_shared_open(gpurt_ctx* ctx, const uint8_t handle[GPURT_IPC_HANDLE_BYTES], void** out_device_ptr);
int gpurt_shared_close(gpurt_ctx* ctx, void* mapped_device_ptr);

/* ---- multi-GPU gather of sharded result arrays without a collective ------------------------------------
 * n_records results of record_bytes (16: GpurtHit, 32: GpurtClosestPoint) live in ONE array on the owner rank; rank r
 * answers the records [first_record[r], first_record[r + 1]) with ONE query call per round, passing the pointer
 * gpurt_gather_results gives it as the device `hits` / `results` argument.  A large incoherent batch on a large scene is
 * traversed in Morton order (csrc/order.cu), so its results come out in processing order: the sender traverses the batch in
 * slices of that order and, after each slice, its copy engine moves the slice's results and storage indices — coalesced —
 * into an inbox next to the owner's array and raises a flag there; kernels the owner queued with gpurt_gather_begin wait for
 * the flags and scatter the slices to their storage positions locally while the owner's own batch is still being traversed.
 * Small or already coherent batches are stored by the traversal kernel directly, as with gpurt_shared_open.  Measured on
 * config 4 at 8 GPUs: see csrc/gather.cu.
 *   owner:   gpurt_gather_create -> handle to the other ranks;   per round: gpurt_gather_begin, its own query call (results =
 *            the pointer from gpurt_gather_results), gpurt_gather_end (the context's stream then waits for all scatters)
 *   others:  gpurt_gather_open (handle from another process, or same_process_base = the owner's base pointer inside one
 *            process);   per round: one query call with results = the pointer from gpurt_gather_results
 * Rounds are counted on both sides: every rank issues exactly one call per round.  A sender does not overwrite the inbox before
 * the owner has scattered the previous round (acknowledged over NVLink); waits give up after 20 s (gpurt_gather_end reports
 * the count). */
typedef struct gpurt_gather gpurt_gather;
int gpurt_gather_create(gpurt_ctx* ctx, uint64_t n_records, uint32_t record_bytes, uint32_t n_ranks,
                        const uint64_t* first_record /* [n_ranks + 1] */, uint32_t owner_rank, gpurt_gather** out,
                        uint8_t handle_out[GPURT_IPC_HANDLE_BYTES], uint64_t* out_bytes);
int gpurt_gather_open(gpurt_ctx* ctx, const uint8_t handle[GPURT_IPC_HANDLE_BYTES], void* same_process_base, uint64_t n_records,
                      uint32_t record_bytes, uint32_t n_ranks, const uint64_t* first_record, uint32_t owner_rank,
                      uint32_t my_rank, gpurt_gather** out);
/* out_array: record 0 of the whole array (the owner's memory or its mapping); out_mine: this rank's first record */
int gpurt_gather_results(gpurt_gather* g, void** out_array, void** out_mine);
int gpurt_gather_base(gpurt_gather* g, void** out_base); /* owner: what same_process_base takes */
int gpurt_gather_begin(gpurt_gather* g);
int gpurt_gather_end(gpurt_gather* g, uint32_t* out_timeouts /* NULL: do not synchronise */);
int gpurt_gather_destroy(gpurt_gather* g);

/* ---- integrator: replaces VK::RTPipe (src/vk/rt.h:14-142) -------------------------------------- */
int gpurt_pipe_params_default(GpurtPipeParams* out); /* rt.h:38-53 */
/* RTPipe::recreate(scene) + use_accel(tlas) (src/vk/rt.cpp:16-24, :159-176) */
int gpurt_pipe_create(gpurt_scene* scene, gpurt_accel* accel, gpurt_pipe** out);
int gpurt_pipe_destroy(gpurt_pipe* pipe);
/* RTPipe::reset_frame (src/vk/rt.cpp:396-398) */
int gpurt_pipe_reset_frame(gpurt_pipe* pipe);
/* RTPipe::update_uniforms + RTPipe::trace (src/vk/rt.cpp:121-138, :346-394): renders one
 * progressive frame of width x height; returns 1 when frame >= max_frames (nothing rendered),
 * like trace() returning false. cam->prev_PV is filled by the pipe from the previous call. */
int gpurt_pipe_render_frame(gpurt_pipe* pipe, const GpurtPipeParams* params, const GpurtCamera* cam,
                            uint32_t width, uint32_t height);
int gpurt_pipe_frame_index(const gpurt_pipe* pipe, int32_t* out_frame);
/* Multi-GPU sharding of a frame (the reference is single-GPU): this pipe renders only the
 * bands of `band_rows` rows whose band index is congruent to `shard` modulo `n_shards`.  RNG streams,
 * image and G-buffers stay indexed by the global pixel, so the union of all shards is bit-identical to
 * an unsharded frame (integrators 0-2; ReSTIR's temporal pass reads neighbouring pixels of the previous
 * frame and needs the whole previous frame on the rank: gpurt_pipe_history_peers below).  band_rows = 0 restores
 * whole-frame rendering. */
int gpurt_pipe_set_shard(gpurt_pipe* pipe, uint32_t band_rows, uint32_t n_shards, uint32_t shard);
/* ReSTIR (integrators 3, 4) on a sharded frame (SURVEY §8e "halo exchange or all-gather of the previous frame"; the temporal
 * pass reprojects the hit point into the previous frame, rt.rgen:454-472, and reads G-buffers + reservoir there): every
 * shard keeps buffers for the whole frame and, at the end of a frame, stores the rows it rendered straight into the other
 * shards' buffers over peer memory (NVLink), then raises a per-shard flag there; the next frame's first kernel waits for the
 * flags.  No host synchronisation and no collective library call per frame; frames stay asynchronous on the context's stream.
 *   1. every shard: gpurt_pipe_set_shard, then gpurt_pipe_history_export(w, h) -> its history block (device pointer, 64-byte
 *      IPC handle to ship to the other processes, size).  The block is re-allocated when the frame size changes: export again.
 *   2. every shard: gpurt_pipe_history_peers(n_shards, blocks[shard index], halo_rows) with the other shards' blocks mapped
 *      by gpurt_shared_open (or plain device pointers inside one process); the own entry is ignored.
 *      halo_rows = GPURT_HISTORY_ALL_ROWS: every shard receives every row — bit-identical to the unsharded frame for any
 *      camera motion.  Otherwise a row goes only to the shards owning a row within halo_rows of it: bit-identical as long as
 *      reprojection (and spatial_radius) stay within halo_rows rows — e.g. a static or slowly moving camera with a few
 *      contiguous bands (band_rows = ceil(h / n_shards)) moves (n-1)/n less data.
 *   3. all shards render the same sequence of frames (same count, same sizes); a shard that is 20 s late is not waited for
 *      any longer (the frame proceeds on stale rows and gpurt_pipe_history_status counts the time-out).
 * Inside one process on one stream, render frame f on every shard before frame f + 1 on any.
 * Lifetime: a shard's block is freed by gpurt_pipe_destroy and re-allocated when the frame size changes — the other shards
 * call gpurt_pipe_history_peers(pipe, 0, NULL, 0) and unmap it (gpurt_shared_close) first. */
#define GPURT_HISTORY_ALL_ROWS 0xFFFFFFFFu
int gpurt_pipe_history_export(gpurt_pipe* pipe, uint32_t width, uint32_t height, void** out_device_ptr,
                              uint8_t handle_out[GPURT_IPC_HANDLE_BYTES], uint64_t* out_bytes);
int gpurt_pipe_history_peers(gpurt_pipe* pipe, uint32_t n_shards, void* const* blocks, uint32_t halo_rows);
int gpurt_pipe_history_status(gpurt_pipe* pipe, uint32_t* out_frames_pushed, uint32_t* out_timeouts);
/* Frame-parallel sharding (integrators 0-2, whose frames are independent given the frame index): render frame
 * `frame` of the progressive sequence — same RNG streams as RTPipe::trace would use for it — and write the
 * per-pixel mean of its samples (rt.rgen:638) to mean_out_device (w*h RGBA32F, may be a gpurt_shared_open
 * mapping of another GPU's memory) without touching this pipe's accumulated image or frame counter. */
int gpurt_pipe_render_frame_mean(gpurt_pipe* pipe, const GpurtPipeParams* params, const GpurtCamera* cam,
                                 uint32_t width, uint32_t height, int32_t frame, void* mean_out_device);
/* image = frame == 0 ? mean : mix(image, mean, 1/(frame+1)) (rt.rgen:640-645); call for frame = 0, 1, 2, ...
 * in order: the result is bit-identical to rendering the frames one after the other on one GPU. */
int gpurt_pipe_accumulate_mean(gpurt_pipe* pipe, const void* mean_device, int32_t frame, uint32_t width,
                               uint32_t height);
/* rt_target (RGBA32F, src/gpurt.cpp:189-193) -> caller buffer of width*height*4 floats. */
int gpurt_pipe_read_image(gpurt_pipe* pipe, float* out_rgba, int mem);
/* The same copy without stalling the render loop (GPURT::render hands rt_target to the next pass on the GPU,
 * src/gpurt.cpp:60-66; a headless consumer wants it in host memory): snapshot of the image as of the frames queued so
 * far, copied to page-locked host memory by the pipe's own copy stream while the following frames trace and shade — only
 * their accumulation into rt_target waits for the copy.  gpurt_pipe_read_image_wait returns once the last requested copy
 * has landed (earlier ones have, too: they are ordered).  Pageable memory works but makes the call itself synchronous. */
int gpurt_pipe_read_image_async(gpurt_pipe* pipe, float* out_rgba_host);
int gpurt_pipe_read_image_wait(gpurt_pipe* pipe);
/* which: 0 position, 1 normal, 2 albedo (rt.rgen:674-676) */
int gpurt_pipe_read_gbuffer(gpurt_pipe* pipe, int which, float* out_rgba, int mem);
/* rays traced by the last frame: [0] closest-hit, [1] any-hit */
int gpurt_pipe_ray_counts(const gpurt_pipe* pipe, uint64_t out_counts[2]);
/* Device pointer of rt_target (RGBA32F) for zero-copy consumers (e.g. an NCCL gather of tiles). */
int gpurt_pipe_device_image(gpurt_pipe* pipe, void** out_dev_rgba);
/* Push constants (rt.h:85-102) and UBO (rt.h:104-117) exactly as used by the last rendered frame,
 * plus the per-frame seed word that replaces clockARB(). For parity tests / debugging. */
int gpurt_pipe_last_uniforms(const gpurt_pipe* pipe, GpurtConstants* out_consts, GpurtCamera* out_ubo,
                             uint32_t* out_seed_word);
/* The reservoir buffer written by the last frame (rt.rgen:633-636): 12 words per pixel
 * {pos.xyz, w_sum, normal.xyz, w, emissive.xyz, n_seen}. */
int gpurt_pipe_read_reservoirs(gpurt_pipe* pipe, float* out_words, int mem);
/* Device pointer + count of the ray queue (GpurtRay) traced at `bounce` in the last sample of the
 * last frame; lets a benchmark re-trace exactly the rays the integrator produced. */
int gpurt_pipe_bounce_rays(gpurt_pipe* pipe, uint32_t bounce, void** out_dev_rays, uint32_t* out_count);

/* tonemap.frag:17-48 (exposure/Uncharted2 + gamma) -> RGBA8, host or device output */
int gpurt_tonemap(gpurt_pipe* pipe, int op, float exposure, float gamma, uint8_t* out_rgba8, int mem);

/* rt_target as a file in linear radiance (the reference vendors tinyexr next to stb_image_write,
 * deps/sf_libs, and GPURT::save_rt, src/gpurt.cpp:258-262, writes the tonemapped PNG): OpenEXR 2, one part, scanlines,
 * no compression, channels A B G R as 32-bit float.  rgba = width*height*4 floats in HOST memory, row 0 on top. */
int gpurt_write_exr(const char* path, const float* rgba, uint32_t width, uint32_t height);

#ifdef __cplusplus
}
#endif
#endif /* GPURT_H */
