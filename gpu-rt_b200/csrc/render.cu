/*
 * render.cu — wavefront path-tracing / ReSTIR-DI integrator: replaces VK::RTPipe
 * (src/vk/rt.h:14-142, rt.cpp) and the ray-generation shader rt.rgen (one mega-kernel thread per
 * pixel, rt.rgen:567-677) with a sequence of kernels per progressive frame:
 *
 *   k_frame_begin                      per pixel: RNG seed, zero accumulators / G-buffer / reservoir
 *   for s < samples:
 *     k_gen_camera                     make_camera_ray (rt.rgen:551-565) -> ray queue 0 (all pixels)
 *     for depth < wave:                (wave = max_depth, or 4 for small shards)
 *       k_trace_closest_indirect       the batch closest-hit kernel over the compacted queue
 *       k_shade                        miss / hit_info / mat_info / shade_info / integrator / Russian
 *                                      roulette; surviving paths are compacted into the next queue
 *                                      with a warp ballot + one atomicAdd per warp
 *     k_tail                           remaining bounces of the surviving paths, one thread per path
 *   k_frame_end                        reservoir / progressive accumulation / debug view
 *
 * Per-pixel state (RNG, throughput, radiance, mis weight) lives in two float4 arrays indexed by
 * pixel, so a path keeps the RNG stream of its pixel across bounces and samples exactly like the
 * sequential shader (results do not depend on queue order).  Shadow and light rays inside an
 * integrator are traced inline by the shading thread.
 */
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../host/camera.h"
#include "device.cuh"
#include "shade.cuh"

using namespace gpurt;

/* Minimum CTAs per SM asked of the shade / tail kernels of each integrator (ptxas caps the registers accordingly).  Measured
 * on mis_test 1080p (tools/ab_restir.sh, variants built side by side): MIS 5 -> 6 CTAs (96 -> 80 registers, 40 B of spills)
 * 0.434 -> 0.386 ms per frame, 7: 0.395, 4: 0.446; ReSTIR 4 -> 7 CTAs (128 -> 72 registers, 504 B of spills) 0.283 -> 0.265 and
 * 0.331 -> 0.308 ms, 6: 0.270 / 0.316, 8: 0.262 / 0.311; direct 5 -> 8 CTAs (96 -> 64 registers, 60 B of spills): the reference's
 * default workload (cbox 1280x720, 8 spp, depth 8) 1.873 -> 1.720 ms per frame, 7 CTAs: 1.742; material (64 registers as it
 * is) 9 / 10 CTAs: -1 % / +0.5 %, left alone.  These kernels wait on dependent gathers; more resident warps beat fewer spills. */
#ifndef GPURT_MIS_MINB
#define GPURT_MIS_MINB 6
#endif
#ifndef GPURT_RESTIR_MINB
#define GPURT_RESTIR_MINB 7
#endif
#ifndef GPURT_MAT_MINB
#define GPURT_MAT_MINB 0
#endif
#ifndef GPURT_DIRECT_MINB
#define GPURT_DIRECT_MINB 8
#endif

struct gpurt_pipe {
    gpurt_ctx* ctx = nullptr;
    gpurt_scene* scene = nullptr;
    gpurt_accel* accel = nullptr;
    uint32_t w = 0, h = 0;
    float4* image = nullptr;                /* rt_target, RGBA32F (gpurt.cpp:189-193) */
    float4* res[2] = {nullptr, nullptr};    /* ping-pong reservoirs (rt.cpp:183-184), 3 float4 / pixel */
    float4* gbuf[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}}; /* rt.cpp:193-198 */
    float4 *acc = nullptr, *pathA = nullptr, *pathB = nullptr, *rays[2] = {nullptr, nullptr}, *hits = nullptr;
    uint32_t* queue[2] = {nullptr, nullptr};
    uint32_t* counts = nullptr;             /* [2*i], [2*i+1]: queue sizes, zeroed per frame */
    unsigned long long* ray_counts = nullptr;
    int parity = 0;
    int frame = -1;                         /* consts.frame (rt.h:88) */
    GpurtCamera old_cam;                    /* rt.h:132; identity matrices until the first change */
    bool old_cam_init = false;
    FrameParams last;                       /* uniforms of the last rendered frame */
    uint64_t last_counts[2] = {0, 0};
    uint32_t max_counts = 0;
    uint32_t band_rows = 0, n_shards = 1, shard = 0; /* gpurt_pipe_set_shard; 0 = whole frame */
    /* res[] and gbuf[][] live in one block (exportable as a whole, gpurt_pipe_history_export):
     * [kHistHeader: flag of shard s at byte 128*s, time-out counter at byte 8192][parity 0: res 48n, pos, normal, albedo 16n each][parity 1] */
    char* hist = nullptr;
    char** d_peers = nullptr;                         /* device array [64]: history blocks of all shards (own entry = hist) */
    uint32_t n_peers = 0, halo_rows = 0xFFFFFFFFu;    /* gpurt_pipe_history_peers; 0 peers = no exchange */
    uint32_t hist_seq = 0;                            /* frames pushed so far (= the flag value peers wait for) */
    uint32_t wave_depth = 0;                          /* bounces run as wavefronts before k_tail; 0 = by size */
    /* queue sizes of the previous frame, copied back without a sync and used to size this frame's launches */
    uint32_t* h_counts = nullptr;                     /* pinned */
    cudaEvent_t ev_counts = nullptr;
    bool counts_pending = false;
    std::vector<uint32_t> est_counts;
    bool use_est = true;                              /* GPURT_WAVE_ESTIMATE=0: one thread per pixel for every launch */
    /* light groups for light_pdf (shade.cuh light_run_box), rebuilt when the accel's triangles change */
    float4* lgrp = nullptr;
    uint2* lgrp_off = nullptr;
    size_t lgrp_cap = 0, lgrp_off_cap = 0;
    uint64_t lgrp_version = ~0ull;
    const float4* lgrp_tris = nullptr;
    bool use_lgrp = true;                             /* GPURT_LIGHT_GROUPS=0: every light triangle, like the GLSL */
    /* world-space vertices of the light triangles for light_sample (shade.cuh), rebuilt with the light groups */
    float4* lverts = nullptr;
    uint32_t* lvert_off = nullptr;
    size_t lverts_cap = 0, lvert_off_cap = 0;
    uint64_t lverts_version = ~0ull;
    const float4* lverts_tris = nullptr;
    bool use_lverts = true;                           /* GPURT_LIGHT_VERTS=0: light_sample transforms its vertices itself */
    /* power-proportional light sampling (GpurtPipeParams::light_sampling): running sums over the light triangles */
    float* lcdf = nullptr;
    uint32_t* lcdf_off = nullptr;
    size_t lcdf_cap = 0, lcdf_off_cap = 0;
    uint32_t n_ltris = 0;
    uint64_t lcdf_version = ~0ull;
    const float4* lcdf_tris = nullptr;
    bool use_shadow_queue = false;                    /* GPURT_SHADOW_QUEUE=1: integrator 0 queues its shadow rays for k_shadow_resolve
                                                       * (measured: 10 % slower than the inline trace; bit-identical) */
    /* light BVH for light_pdf (shade.cuh light_pdf_bvh): a second accel over the lights' triangles only */
    gpurt_scene* lscene = nullptr;
    gpurt_accel* laccel = nullptr;
    uint64_t laccel_version = ~0ull, laccel_geom = ~0ull;
    bool use_lbvh = true;                             /* GPURT_LIGHT_BVH=0: light-run boxes only */
    /* gpurt_pipe_read_image_async: the image crosses PCIe on a second stream while the next frame traces and shades;
     * the next writer of `image` (k_frame_end, k_accumulate_mean) waits for ev_copy */
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_frame = nullptr, ev_copy = nullptr;
    bool copy_pending = false;
};

namespace gpurt {

static inline unsigned cdivu(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

__global__ void __launch_bounds__(256) k_frame_begin(const __grid_constant__ FrameParams P, int restir, float4* acc,
                                                     float4* pathB, float4* gpos, float4* gnorm, float4* galb,
                                                     float4* res_out) {
    uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    if(li >= P.n_local) return;
    pixel_begin(P, restir, shard_pixel(P, li), acc, pathB, gpos, gnorm, galb, res_out);
}

/* padded boxes over runs of 8 / 64 consecutive triangles of every light (blockIdx.y = light) */
__global__ void __launch_bounds__(128) k_light_groups(const float4* __restrict__ tri_world, const uint32_t* __restrict__ tri_off,
                                                      const SceneLight* __restrict__ lights, const uint2* __restrict__ off,
                                                      float pad, float4* __restrict__ out) {
    const uint32_t l = blockIdx.y, r = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_tris = lights[l].n_triangles;
    if(r >= light_box_records(n_tris)) return;
    light_box_record(tri_world + 3ull * tri_off[lights[l].index], n_tris, r, pad, out + 2ull * (off[l].x + r));
}

/* model * v of the three vertices of every light triangle, as light_sample computes them (blockIdx.y = light) */
__global__ void __launch_bounds__(128) k_light_verts(DeviceScene S, const uint32_t* __restrict__ off, float4* __restrict__ out) {
    const uint32_t l = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    if(t >= S.lights[l].n_triangles) return;
    light_world_tri(S, S.lights[l].index, t, out + 3ull * (off[l] + t));
}

__global__ void __launch_bounds__(256) k_gen_camera(const __grid_constant__ FrameParams P, uint32_t s,
                                                    float4* pathA, float4* pathB, float4* rays,
                                                    uint32_t* queue, uint32_t* count) {
    uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    if(li == 0) *count = P.n_local;
    if(li >= P.n_local) return;
    pixel_gen_camera(P, s, shard_pixel(P, li), li, pathA, pathB, rays, queue);
}

__global__ void __launch_bounds__(256) k_begin_camera(const __grid_constant__ FrameParams P, int restir, int clear_gbuf, float4* acc,
                                                      float4* pathA, float4* pathB, float4* gpos, float4* gnorm, float4* galb,
                                                      float4* res_out, float4* rays, uint32_t* queue, uint32_t* counts,
                                                      uint32_t n_counts, unsigned long long* ray_counts) {
    uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    /* the frame's counters start here too (n_counts = 0: the host cleared them): queue 0 holds every local pixel, the
     * later queues are empty, no rays counted yet */
    if(li == 0) counts[0] = P.n_local;
    else if(li < n_counts) counts[li] = 0u;
    if(ray_counts && li < 2u) ray_counts[li] = 0ull;
    if(li >= P.n_local) return;
    pixel_begin_camera(P, restir, clear_gbuf, shard_pixel(P, li), li, acc, pathA, pathB, gpos, gnorm, galb, res_out, rays, queue);
}

/* STRIDE = false: one ray per thread, the grid covers the queue (first bounce, or no size estimate yet).
 * STRIDE = true: launches sized from an estimate stride over the queue in case it is larger than estimated; wrapping
 * the traversal in that loop costs 2-4 % (ncu r01v), hence the two instantiations. */
template <bool STRIDE>
__global__ void __launch_bounds__(128) k_trace_closest_indirect(const float4* __restrict__ nodes,
                                                                const float4* __restrict__ tris,
                                                                const float4* __restrict__ rays,
                                                                const uint32_t* __restrict__ count,
                                                                float4* __restrict__ hits, unsigned n_nodes) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(!STRIDE) {
        if(i >= *count) return;
        float4 a = __ldg(rays + 2ull * i), b = __ldg(rays + 2ull * i + 1);
        HitRec h;
        h.t = a.w, h.u = h.v = 0, h.gid = kNoHit;
        if(n_nodes) traverse8<false, false>(nodes, tris, f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), a.w, b.w, h, nullptr);
        hits[i] = make_float4(h.gid == kNoHit ? GPURT_INF : h.t, h.u, h.v, u2f(h.gid));
        return;
    }
    const uint32_t cnt = *count;
    for(; i < cnt; i += gridDim.x * blockDim.x) {
        float4 a = __ldg(rays + 2ull * i), b = __ldg(rays + 2ull * i + 1);
        HitRec h;
        h.t = a.w, h.u = h.v = 0, h.gid = kNoHit;
        if(n_nodes) traverse8<false, false>(nodes, tris, f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), a.w, b.w, h, nullptr);
        hits[i] = make_float4(h.gid == kNoHit ? GPURT_INF : h.t, h.u, h.v, u2f(h.gid));
    }
}

/* DEFER (integrator 0 only): the shadow ray of integrate_direct is not traced here — the pixel's segment goes to the
 * (otherwise unused) next-bounce queue together with the term it gates, and k_shadow_resolve traces it and finishes
 * the pixel.  Paths of integrator 0 end at their first hit, so nothing else competes for that queue. */
template <int INTEG, bool DEFER = false>
__global__ void __launch_bounds__(128, INTEG == 2 ? GPURT_MIS_MINB : (INTEG >= 3 ? GPURT_RESTIR_MINB : (INTEG == 1 ? GPURT_MAT_MINB : GPURT_DIRECT_MINB))) k_shade(const __grid_constant__ FrameParams P,
                                               const __grid_constant__ ShadeCtx X, uint32_t s, uint32_t depth,
                                               const uint32_t* __restrict__ count_in, const uint32_t* __restrict__ queue_in,
                                               const float4* __restrict__ rays_in, const float4* __restrict__ hits,
                                               float4* pathA, float4* pathB, float4* acc, float4* gpos, float4* gnorm,
                                               float4* galb, float4* res_cur, uint32_t* count_out, uint32_t* queue_out,
                                               float4* rays_out) {
    /* grid-stride over the queue, one warp-aligned slice of 32 paths per iteration (the trip count is the same for the
     * lanes of a warp, so the full-mask ballot / reduce below are safe) */
    const uint32_t cnt = *count_in, lane = threadIdx.x & 31u;
    unsigned nc = 0, na = 0;
    for(uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < cnt; base += gridDim.x * blockDim.x) {
        const uint32_t k = base + lane;
        const bool live = k < cnt;
        bool cont = false;
        uint32_t pix = 0;
        TraceInfo trace;
        Shader sh(X, P);
        sh.defer_shadow = DEFER;
        if(live) {
            pix = queue_in[k];
            float4 r0 = rays_in[2ull * k], r1 = rays_in[2ull * k + 1], h = hits[k];
            float4 A = pathA[pix], B = pathB[pix];
            trace.o = F3{r0.x, r0.y, r0.z}, trace.d = F3{r1.x, r1.y, r1.z};
            trace.acc = F3{A.x, A.y, A.z}, trace.mis = A.w;
            trace.throughput = F3{B.x, B.y, B.z}, trace.depth = depth;
            sh.seed = f2u(B.w);
            bool broke = shade_step<INTEG>(P, sh, trace, s, depth, pix, h, gpos, gnorm, galb, res_cur);
            cont = !broke && trace.depth + 1 < (uint32_t)P.c.max_depth;
            if(DEFER && sh.shadow_pending) { /* finished by k_shadow_resolve: radiance so far | term, RNG state */
                pathA[pix] = make_float4(trace.acc.x, trace.acc.y, trace.acc.z, 0.0f);
                pathB[pix] = make_float4(sh.shadow_term.x, sh.shadow_term.y, sh.shadow_term.z, u2f(sh.seed));
            } else if(cont) {
                pathA[pix] = make_float4(trace.acc.x, trace.acc.y, trace.acc.z, trace.mis);
                pathB[pix] = make_float4(trace.throughput.x, trace.throughput.y, trace.throughput.z, u2f(sh.seed));
            } else { /* rt.rgen:630: acc += trace.acc; the RNG stream continues into the next sample */
                float4 a = acc[pix];
                acc[pix] = make_float4(a.x + trace.acc.x, a.y + trace.acc.y, a.z + trace.acc.z, 0.0f);
                pathB[pix] = make_float4(1.0f, 1.0f, 1.0f, u2f(sh.seed));
            }
            nc += 1u + sh.n_closest, na += sh.n_any; /* this thread's wavefront ray + its inline rays */
        }
        /* warp-aggregated compaction of the surviving paths (DEFER: of the pixels with a shadow ray to trace) */
        if(DEFER) cont = live && sh.shadow_pending;
        unsigned m = __ballot_sync(0xffffffffu, cont);
        uint32_t slot0 = 0;
        if(m) {
            if(lane == (unsigned)__ffs(m) - 1) slot0 = atomicAdd(count_out, (uint32_t)__popc(m));
            slot0 = __shfl_sync(0xffffffffu, slot0, __ffs(m) - 1);
        }
        if(cont) {
            uint32_t slot = slot0 + __popc(m & ((1u << lane) - 1u));
            queue_out[slot] = pix;
            if(DEFER) { /* the segment of `visibility` (rt.rgen:272-291): direction (b - a) / |b - a|, (EPS, |b - a| - EPS) */
                F3 dir = sh.shadow_b - sh.shadow_a;
                float dl = length3(dir);
                dir = dir / dl;
                rays_out[2ull * slot] = make_float4(sh.shadow_a.x, sh.shadow_a.y, sh.shadow_a.z, kEps);
                rays_out[2ull * slot + 1] = make_float4(dir.x, dir.y, dir.z, dl - kEps);
            } else {
                rays_out[2ull * slot] = make_float4(trace.o.x, trace.o.y, trace.o.z, kEps);
                rays_out[2ull * slot + 1] = make_float4(trace.d.x, trace.d.y, trace.d.z, kLargeDist);
            }
        }
    }
    /* ray accounting, one atomic per warp */
    nc = __reduce_add_sync(0xffffffffu, nc);
    na = __reduce_add_sync(0xffffffffu, na);
    if(lane == 0) {
        if(nc) atomicAdd(X.ray_counts + 0, (unsigned long long)nc);
        if(na) atomicAdd(X.ray_counts + 1, (unsigned long long)na);
    }
}

/* Shadow stage of integrator 0: any-hit trace of the queued segments, then the last line of integrate_direct and the
 * end-of-path accumulation k_shade left undone (rt.rgen:630).  One thread per queued pixel. */
__global__ void __launch_bounds__(128) k_shadow_resolve(const float4* __restrict__ nodes, const float4* __restrict__ tris,
                                                        unsigned n_nodes, const uint32_t* __restrict__ count,
                                                        const uint32_t* __restrict__ queue, const float4* __restrict__ rays,
                                                        float4* pathA, float4* pathB, float4* acc) {
    const uint32_t cnt = *count;
    for(uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < cnt; k += gridDim.x * blockDim.x) {
        const uint32_t pix = queue[k];
        const float4 a = __ldg(rays + 2ull * k), b = __ldg(rays + 2ull * k + 1);
        HitRec h;
        const bool occluded = n_nodes && traverse8<true, false>(nodes, tris, f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), a.w, b.w, h, nullptr);
        const float4 A = pathA[pix], B = pathB[pix];
        const F3 t = Shader::shadow_finish(F3{A.x, A.y, A.z}, F3{B.x, B.y, B.z}, occluded);
        const float4 o = acc[pix];
        acc[pix] = make_float4(o.x + t.x, o.y + t.y, o.z + t.z, 0.0f);
        pathB[pix] = make_float4(1.0f, 1.0f, 1.0f, B.w);
    }
}

/* Tail of the wavefront: after the first bounces only a fraction of the paths is alive and one
 * trace + one shade launch per bounce become latency-bound (worst when a frame is sharded over 8 GPUs).
 * Here every remaining path runs its bounce loop to the end in one thread: same shade_step, same
 * traverse8, same per-pixel RNG stream, so the image does not change — only the launch count does. */
template <int INTEG>
__global__ void __launch_bounds__(128, INTEG == 2 ? GPURT_MIS_MINB : (INTEG >= 3 ? GPURT_RESTIR_MINB : (INTEG == 1 ? GPURT_MAT_MINB : GPURT_DIRECT_MINB))) k_tail(const __grid_constant__ FrameParams P, const __grid_constant__ ShadeCtx X,
                                              uint32_t s, uint32_t depth0, const uint32_t* __restrict__ count_in,
                                              const uint32_t* __restrict__ queue_in, const float4* __restrict__ rays_in,
                                              float4* pathA, float4* pathB, float4* acc, float4* gpos, float4* gnorm,
                                              float4* galb, float4* res_cur) {
    const uint32_t cnt = *count_in;
    unsigned nc = 0, na = 0;
    for(uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < cnt; k += gridDim.x * blockDim.x) {
        Shader sh(X, P);
        nc += path_tail<INTEG>(P, X, sh, s, depth0, queue_in[k], rays_in[2ull * k], rays_in[2ull * k + 1], pathA, pathB, acc,
                               gpos, gnorm, galb, res_cur);
        nc += sh.n_closest, na += sh.n_any;
    }
    nc = __reduce_add_sync(0xffffffffu, nc); /* every lane gets here, whatever its trip count was */
    na = __reduce_add_sync(0xffffffffu, na);
    if((threadIdx.x & 31) == 0) {
        if(nc) atomicAdd(X.ray_counts + 0, (unsigned long long)nc);
        if(na) atomicAdd(X.ray_counts + 1, (unsigned long long)na);
    }
}

/* frame-parallel sharding: fold the frame mean another GPU rendered into this pipe's image */
__global__ void __launch_bounds__(256) k_accumulate_mean(float4* __restrict__ image, const float4* __restrict__ mean,
                                                         uint32_t n, int frame) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    float4 m = mean[i];
    image[i] = accumulate_frame(frame > 0 ? image[i] : m, F3{m.x, m.y, m.z}, frame);
}

__global__ void __launch_bounds__(256) k_frame_end(const __grid_constant__ FrameParams P, const float4* acc,
                                                   float4* image, const float4* gpos, const float4* gnorm,
                                                   const float4* ppos, const float4* pnorm, const float4* palb,
                                                   float4* mean_out, const uint32_t* __restrict__ counts,
                                                   uint32_t* host_counts, uint32_t n_counts) {
    uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    /* this frame's queue sizes, written straight into pinned host memory for the next frame's launch sizes (a D2H copy
     * in the stream would put a copy-engine hop between the last shading kernel and this one) */
    if(host_counts && li < n_counts) host_counts[li] = counts[li];
    if(li >= P.n_local) return;
    pixel_end(P, shard_pixel(P, li), acc, image, gpos, gnorm, ppos, pnorm, palb, mean_out);
}

/* ---- ReSTIR on a sharded frame: the previous frame travels between the shards' history blocks over peer memory ---- */
constexpr size_t kHistHeader = 16384;
constexpr uint32_t kHistMaxShards = 64;

/* One float4 per thread of the rows this shard rendered (reservoirs, position, normal, albedo of the frame just
 * finished), stored to the same offset of every shard that will read the row next frame.  Row-major over the local rows,
 * so loads and remote stores are fully coalesced. */
__global__ void __launch_bounds__(256) k_history_push(const __grid_constant__ FrameParams P, uint32_t halo, const char* own,
                                                      char* const* __restrict__ peers, size_t parity_off) {
    /* blockIdx.y = local row, blockIdx.x * 256 + threadIdx.x = float4 within the row's 6 W float4s: no index division, and
     * the row's readers are worked out once per block (rows nobody reads leave at once) */
    __shared__ unsigned long long s_readers;
    const uint32_t y = shard_row(P, blockIdx.y);
    if(threadIdx.x == 0) s_readers = y < P.H ? (history_row_readers(P, y, halo) & ~(1ull << P.shard)) : 0ull;
    __syncthreads();
    unsigned long long readers = s_readers;
    const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
    if(!readers || o >= 6u * P.W) return;
    const size_t n = (size_t)P.W * P.H;
    size_t off;
    if(o < 3u * P.W) off = 3ull * y * P.W + o;
    else {
        const uint32_t g = o >= 5u * P.W ? 2u : o >= 4u * P.W ? 1u : 0u, x = o - (g + 3u) * P.W;
        off = (3ull + g) * n + (size_t)y * P.W + x;
    }
    float4 v = ((const float4*)(own + parity_off))[off];
    while(readers) {
        int s = __ffsll((long long)readers) - 1;
        readers &= readers - 1;
        ((float4*)(peers[s] + parity_off))[off] = v;
    }
    __threadfence_system();
}
/* "this shard's rows of frame number `seq` are in your block" — after k_history_push in stream order */
__global__ void k_history_signal(char* const* __restrict__ peers, uint32_t n_shards, uint32_t shard, uint32_t seq) {
    uint32_t s = threadIdx.x;
    if(s >= n_shards || !peers[s]) return;
    __threadfence_system();
    *(volatile uint32_t*)(peers[s] + 128ull * shard) = seq;
}
/* wait until every shard has delivered `seq` frames (also means: every shard is done reading the buffers this frame is
 * about to overwrite).  Gives up after 20 s and counts the time-out instead of hanging the device. */
__global__ void k_history_wait(char* own, uint32_t n_shards, uint32_t seq) {
    uint32_t s = threadIdx.x;
    if(s >= n_shards) return;
    const volatile uint32_t* f = (const volatile uint32_t*)(own + 128ull * s);
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while((int32_t)(*f - seq) < 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if(t - t0 > 20000000000ull) {
            atomicAdd((uint32_t*)(own + 8192), 1u);
            break;
        }
        __nanosleep(100);
    }
    __threadfence_system();
}

/* tonemap.frag:17-48 followed by the R8G8B8A8_SRGB framebuffer encode (gpurt.cpp:176) */
__global__ void __launch_bounds__(256) k_tonemap(const float4* __restrict__ img, uint32_t n, int op, float exposure,
                                                 float gamma, uchar4* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    float4 c = img[i];
    float v[4] = {c.x, c.y, c.z, c.w};
    const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
    float x11 = 11.2f;
    float white = 1.0f / (((x11 * (A * x11 + C * B) + D * E) / (x11 * (A * x11 + B) + D * F)) - E / F);
    float ig = 1.0f / gamma;
    unsigned char q[4];
    for(int k = 0; k < 4; k++) {
        float x = v[k];
        if(k < 3) {
            if(op == 0) {
                float y = x * exposure;
                x = dm_pow((((y * (A * y + C * B) + D * E) / (y * (A * y + B) + D * F)) - E / F) * white, ig);
            } else if(op == 1)
                x = dm_pow(1.0f - dm_exp(-x * exposure), ig);
        }
        x = x != x ? 0.0f : fminf(fmaxf(x, 0.0f), 1.0f);
        if(k < 3) x = x <= 0.0031308f ? 12.92f * x : 1.055f * dm_pow(x, 1.0f / 2.4f) - 0.055f;
        q[k] = (unsigned char)(x * 255.0f + 0.5f);
    }
    out[i] = make_uchar4(q[0], q[1], q[2], q[3]);
}

static int pipe_free(gpurt_pipe* p) {
    void* ptrs[] = {p->image, p->hist, p->d_peers, p->acc, p->pathA, p->pathB, p->rays[0], p->rays[1], p->hits,
                    p->queue[0], p->queue[1], p->counts, p->ray_counts, p->lgrp, p->lgrp_off, p->lverts, p->lvert_off, p->lcdf, p->lcdf_off};
    for(void* q : ptrs)
        if(q) cudaFree(q);
    return GPURT_OK;
}

/* RTPipe::resize_temporal_stuff (rt.cpp:178-220) + rt_target (gpurt.cpp:189-193) */
static int pipe_resize_impl(gpurt_pipe* p, uint32_t w, uint32_t h, uint32_t max_depth);
static int pipe_resize(gpurt_pipe* p, uint32_t w, uint32_t h, uint32_t max_depth) {
    int rc = pipe_resize_impl(p, w, h, max_depth);
    if(rc) p->w = p->h = 0, p->max_counts = 0; /* some buffers have the new size, some are gone: the next call starts over */
    return rc;
}
static int pipe_resize_impl(gpurt_pipe* p, uint32_t w, uint32_t h, uint32_t max_depth) {
    uint32_t need_counts = 2 * (max_depth + 2);
    if(p->w == w && p->h == h && p->max_counts >= need_counts) return GPURT_OK;
    bool dims = !(p->w == w && p->h == h);
    size_t n = (size_t)w * h;
    cudaStream_t st = p->ctx->stream;
    auto alloc = [&](auto*& ptr, size_t bytes) -> int {
        if(ptr) cudaFree(ptr);
        ptr = nullptr;
        GPURT_CUDA(cudaMalloc((void**)&ptr, bytes ? bytes : 16));
        GPURT_CUDA(cudaMemsetAsync(ptr, 0, bytes ? bytes : 16, st));
        return GPURT_OK;
    };
    int rc;
    if(dims) {
        if((rc = alloc(p->image, n * 16))) return rc;
        p->n_peers = 0, p->hist_seq = 0; /* a new block: the shards exchange handles again (gpurt_pipe_history_export) */
        for(int k = 0; k < 2; k++) {
            p->res[k] = nullptr;
            for(int g = 0; g < 3; g++) p->gbuf[k][g] = nullptr;
        }
        if((rc = alloc(p->hist, kHistHeader + 2 * n * 96))) return rc;
        for(int k = 0; k < 2; k++) {
            char* base = p->hist + kHistHeader + (size_t)k * n * 96;
            p->res[k] = (float4*)base;
            for(int g = 0; g < 3; g++) p->gbuf[k][g] = (float4*)(base + n * 48 + (size_t)g * n * 16);
            if((rc = alloc(p->rays[k], n * 32))) return rc;
            if((rc = alloc(p->queue[k], n * 4))) return rc;
        }
        if((rc = alloc(p->acc, n * 16))) return rc;
        if((rc = alloc(p->pathA, n * 16))) return rc;
        if((rc = alloc(p->pathB, n * 16))) return rc;
        if((rc = alloc(p->hits, n * 16))) return rc;
        if((rc = alloc(p->ray_counts, 16))) return rc;
        p->parity = 0;
    }
    if(p->max_counts < need_counts) {
        if((rc = alloc(p->counts, need_counts * 4))) return rc;
        GPURT_CUDA(cudaStreamSynchronize(st)); /* a copy into the old pinned block may be in flight */
        if(p->h_counts) cudaFreeHost(p->h_counts);
        p->h_counts = nullptr, p->counts_pending = false;
        GPURT_CUDA(cudaMallocHost((void**)&p->h_counts, need_counts * 4));
        if(!p->ev_counts) GPURT_CUDA(cudaEventCreateWithFlags(&p->ev_counts, cudaEventDisableTiming));
        p->max_counts = need_counts;
    }
    if(dims) p->est_counts.clear();
    p->w = w, p->h = h;
    return GPURT_OK;
}

/* (re)build the light groups when the accel's world-space triangles changed (build / update) */
static int pipe_light_groups(gpurt_pipe* p) {
    const gpurt_accel* A = p->accel;
    const std::vector<SceneLight>& L = p->scene->packed.lights;
    if(!p->use_lgrp || L.empty() || !A->tri_gid || L.size() != A->dscene.n_lights) {
        p->lgrp_tris = nullptr;
        return GPURT_OK;
    }
    if(p->lgrp_tris == A->tri_gid && p->lgrp_version == A->dscene.version) return GPURT_OK;
    cudaStream_t st = p->ctx->stream;
    std::vector<uint2> off(L.size());
    uint32_t total = 0, most = 0;
    for(size_t l = 0; l < L.size(); l++) {
        uint32_t ng = (L[l].n_triangles + kLightRun - 1) / kLightRun, nsg = (ng + kLightRun - 1) / kLightRun;
        off[l] = make_uint2(total, total + nsg);
        total += nsg + ng;
        most = std::max(most, nsg + ng);
    }
    if(p->lgrp_cap < total || p->lgrp_off_cap < L.size()) {
        GPURT_CUDA(cudaStreamSynchronize(st));
        if(p->lgrp) cudaFree(p->lgrp);
        if(p->lgrp_off) cudaFree(p->lgrp_off);
        p->lgrp = nullptr, p->lgrp_off = nullptr, p->lgrp_cap = p->lgrp_off_cap = 0;
        GPURT_CUDA(cudaMalloc((void**)&p->lgrp, (size_t)std::max(total, 1u) * 32));
        GPURT_CUDA(cudaMalloc((void**)&p->lgrp_off, L.size() * sizeof(uint2)));
        p->lgrp_cap = total, p->lgrp_off_cap = L.size();
    }
    /* pageable source: the copy is staged before the call returns, so `off` may go out of scope */
    GPURT_CUDA(cudaMemcpyAsync(p->lgrp_off, off.data(), L.size() * sizeof(uint2), cudaMemcpyHostToDevice, st));
    if(most)
        k_light_groups<<<dim3(cdivu(most, 128), (unsigned)L.size()), 128, 0, st>>>(
            A->tri_gid, A->dscene.tri_off, A->dscene.lights, p->lgrp_off, A->inflate * kLightPadScale, p->lgrp);
    GPURT_CUDA(cudaGetLastError());
    p->lgrp_tris = A->tri_gid, p->lgrp_version = A->dscene.version;
    return GPURT_OK;
}

/* (re)build the world-space light vertices when the scene's geometry or poses changed */
static int pipe_light_verts(gpurt_pipe* p) {
    const gpurt_accel* A = p->accel;
    const PackedScene& M = p->scene->packed;
    const std::vector<SceneLight>& L = M.lights;
    bool usable = p->use_lverts && !L.empty() && A->tri_gid && L.size() == A->dscene.n_lights;
    for(const SceneLight& l : L)
        usable = usable && l.index < M.descs.size() && l.n_triangles == M.tri_off[l.index + 1] - M.tri_off[l.index];
    if(!usable) {
        p->lverts_tris = nullptr;
        return GPURT_OK;
    }
    if(p->lverts_tris == A->tri_gid && p->lverts_version == A->dscene.version) return GPURT_OK;
    cudaStream_t st = p->ctx->stream;
    std::vector<uint32_t> off(L.size());
    uint32_t total = 0, most = 0;
    for(size_t l = 0; l < L.size(); l++) off[l] = total, total += L[l].n_triangles, most = std::max(most, L[l].n_triangles);
    if(p->lverts_cap < total || p->lvert_off_cap < L.size()) {
        GPURT_CUDA(cudaStreamSynchronize(st));
        if(p->lverts) cudaFree(p->lverts);
        if(p->lvert_off) cudaFree(p->lvert_off);
        p->lverts = nullptr, p->lvert_off = nullptr, p->lverts_cap = p->lvert_off_cap = 0;
        GPURT_CUDA(cudaMalloc((void**)&p->lverts, (size_t)std::max(total, 1u) * 48));
        GPURT_CUDA(cudaMalloc((void**)&p->lvert_off, L.size() * sizeof(uint32_t)));
        p->lverts_cap = total, p->lvert_off_cap = L.size();
    }
    GPURT_CUDA(cudaMemcpyAsync(p->lvert_off, off.data(), L.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    if(most) k_light_verts<<<dim3(cdivu(most, 128), (unsigned)L.size()), 128, 0, st>>>(A->dscene, p->lvert_off, p->lverts);
    GPURT_CUDA(cudaGetLastError());
    p->lverts_tris = A->tri_gid, p->lverts_version = A->dscene.version;
    return GPURT_OK;
}

/* Running sums of light_tri_power over the light triangles, in (light, triangle) order, for light_sampling = 1.  The
 * weights come from the world-space vertices of the light vertex table (read back: a few thousand triangles) and the
 * objects' emissive factors; the sum is taken on the host in index order, as the oracle takes it.  Rebuilt with the table. */
static int pipe_light_cdf(gpurt_pipe* p) {
    int rc = pipe_light_verts(p);
    if(rc) return rc;
    if(!p->lverts_tris) return set_error("light_sampling = 1 needs the light vertex table (GPURT_LIGHT_VERTS=0, or lights that are not whole objects)"), GPURT_E_STATE;
    if(p->lcdf_tris == p->lverts_tris && p->lcdf_version == p->lverts_version) return GPURT_OK;
    const PackedScene& M = p->scene->packed;
    const std::vector<SceneLight>& L = M.lights;
    cudaStream_t st = p->ctx->stream;
    std::vector<uint32_t> off(L.size() + 1, 0);
    for(size_t l = 0; l < L.size(); l++) off[l + 1] = off[l] + L[l].n_triangles;
    const uint32_t total = off.back();
    std::vector<float4> v(3ull * std::max(total, 1u));
    if(total) GPURT_CUDA(cudaMemcpyAsync(v.data(), p->lverts, (size_t)total * 48, cudaMemcpyDeviceToHost, st));
    GPURT_CUDA(cudaStreamSynchronize(st));
    std::vector<float> cdf(std::max(total, 1u), 0.0f);
    float run = 0.0f;
    for(size_t l = 0; l < L.size(); l++) {
        const float* e = reinterpret_cast<const float*>(&M.descs[L[l].index]) + 36; /* Scene_Desc::emissive (rt.h:68-77) */
        for(uint32_t t = 0; t < L[l].n_triangles; t++) {
            const float4 *q = &v[3ull * (off[l] + t)];
            float w = light_tri_power(F3{q[0].x, q[0].y, q[0].z}, F3{q[1].x, q[1].y, q[1].z}, F3{q[2].x, q[2].y, q[2].z}, F3{e[0], e[1], e[2]});
            if(!(w > 0.0f) || w > 3.0e38f) w = 0.0f; /* NaN / negative / infinite weights never get chosen */
            run += w;
            cdf[off[l] + t] = run;
        }
    }
    if(p->lcdf_cap < total || p->lcdf_off_cap < L.size() + 1) {
        if(p->lcdf) cudaFree(p->lcdf);
        if(p->lcdf_off) cudaFree(p->lcdf_off);
        p->lcdf = nullptr, p->lcdf_off = nullptr, p->lcdf_cap = p->lcdf_off_cap = 0;
        GPURT_CUDA(cudaMalloc((void**)&p->lcdf, (size_t)std::max(total, 1u) * 4));
        GPURT_CUDA(cudaMalloc((void**)&p->lcdf_off, (L.size() + 1) * 4));
        p->lcdf_cap = total, p->lcdf_off_cap = L.size() + 1;
    }
    GPURT_CUDA(cudaMemcpyAsync(p->lcdf, cdf.data(), (size_t)std::max(total, 1u) * 4, cudaMemcpyHostToDevice, st));
    GPURT_CUDA(cudaMemcpyAsync(p->lcdf_off, off.data(), (L.size() + 1) * 4, cudaMemcpyHostToDevice, st));
    GPURT_CUDA(cudaStreamSynchronize(st)); /* the host vectors leave scope */
    p->n_ltris = total, p->lcdf_tris = p->lverts_tris, p->lcdf_version = p->lverts_version;
    return GPURT_OK;
}

static void pipe_drop_light_accel(gpurt_pipe* p) {
    if(p->laccel) gpurt_accel_destroy(p->laccel);
    delete p->lscene;
    p->laccel = nullptr, p->lscene = nullptr;
}

/* (re)build the light BVH: the emissive objects of the scene, in light order, as a scene of their own (same vertices,
 * indices and model matrices -> k_flatten produces the same world-space triangles), built by the ordinary pipeline.
 * Its culling boxes are padded for ray origins anywhere in the MAIN scene (min_inflate). */
static int pipe_light_accel(gpurt_pipe* p) {
    const gpurt_accel* A = p->accel;
    const PackedScene& M = p->scene->packed;
    bool usable = p->use_lbvh && !M.lights.empty() && M.lights.size() == A->dscene.n_lights && A->n > 0;
    for(const SceneLight& L : M.lights)
        usable = usable && L.n_triangles > 0 && L.index < M.descs.size() && L.n_triangles == M.tri_off[L.index + 1] - M.tri_off[L.index];
    if(!usable) { /* a light without triangles makes the GLSL's 0/0: leave that to the scan, which reproduces it */
        pipe_drop_light_accel(p);
        return GPURT_OK;
    }
    const bool same_geom = p->laccel && p->laccel_geom == A->dscene.geom_version && p->lscene->packed.descs.size() == M.lights.size();
    if(same_geom && p->laccel_version == A->dscene.version) return GPURT_OK;
    int rc;
    if(!same_geom) {
        pipe_drop_light_accel(p);
        p->lscene = new gpurt_scene;
        p->lscene->ctx = p->ctx;
        p->lscene->label = "lights";
        PackedScene& S = p->lscene->packed;
        S.tri_off.push_back(0), S.vert_off.push_back(0);
        for(const SceneLight& L : M.lights) {
            const uint32_t o = L.index;
            S.descs.push_back(M.descs[o]);
            S.verts.insert(S.verts.end(), M.verts.begin() + M.vert_off[o], M.verts.begin() + M.vert_off[o + 1]);
            S.idx.insert(S.idx.end(), M.idx.begin() + 3ull * M.tri_off[o], M.idx.begin() + 3ull * M.tri_off[o + 1]);
            S.tri_off.push_back((uint32_t)(S.idx.size() / 3)), S.vert_off.push_back((uint32_t)S.verts.size());
        }
        p->lscene->dirty = false; /* `packed` is authoritative: there is no host Scene behind it */
        p->lscene->version = p->lscene->geom_version = 1;
        p->laccel = new gpurt_accel;
        p->laccel->ctx = p->ctx, p->laccel->scene = p->lscene, p->laccel->flags = GPURT_BUILD_LBVH; /* rebuilt on every pose edit */
    } else { /* pose edit: same geometry, new model matrices */
        for(size_t l = 0; l < M.lights.size(); l++) p->lscene->packed.descs[l] = M.descs[M.lights[l].index];
        p->lscene->version++;
    }
    p->laccel->min_inflate = A->inflate * kLightPadScale;
    rc = build_accel_device(p->laccel);
    if(rc == GPURT_OK && p->laccel->depth > 60) rc = GPURT_E_STATE;
    if(rc) {
        pipe_drop_light_accel(p);
        return rc == GPURT_E_STATE ? GPURT_OK : rc; /* too deep for the stack: scan instead */
    }
    p->laccel_version = A->dscene.version, p->laccel_geom = A->dscene.geom_version;
    return GPURT_OK;
}

/* __constant__ symbols exist once per device: the table is uploaded once per device ordinal */
static cudaError_t upload_lut_once(int device) {
    static std::mutex mu;
    static bool done[256] = {};
    std::lock_guard<std::mutex> lock(mu);
    if(device < 0 || device >= 256) return cudaErrorInvalidDevice;
    if(done[device]) return cudaSuccess;
    float lut[256];
    for(int i = 0; i < 256; i++) {
        double c = i / 255.0;
        lut[i] = (float)(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
    }
    cudaError_t e = cudaMemcpyToSymbol(c_srgb_lut, lut, sizeof(lut));
    if(e == cudaSuccess) done[device] = true;
    return e;
}

} // namespace gpurt

extern "C" {

int gpurt_pipe_create(gpurt_scene* scene, gpurt_accel* accel, gpurt_pipe** out) {
    if(!scene || !accel || !out) return set_error("NULL argument"), GPURT_E_INVALID;
    if(accel->scene != scene) return set_error("accel was built from a different scene"), GPURT_E_STATE;
    GPURT_CUDA(cudaSetDevice(accel->ctx->device));
    GPURT_CUDA(upload_lut_once(accel->ctx->device));
    gpurt_pipe* p = new gpurt_pipe;
    p->ctx = accel->ctx, p->scene = scene, p->accel = accel;
    std::memset(&p->old_cam, 0, sizeof(p->old_cam));
    std::memset(&p->last, 0, sizeof(p->last));
    if(const char* e = getenv("GPURT_WAVE_DEPTH")) p->wave_depth = (uint32_t)std::max(1, atoi(e)); /* tuning knob */
    if(const char* e = getenv("GPURT_WAVE_ESTIMATE")) p->use_est = atoi(e) != 0;                   /* A/B knob */
    if(const char* e = getenv("GPURT_LIGHT_GROUPS")) p->use_lgrp = atoi(e) != 0;                   /* A/B knob */
    if(const char* e = getenv("GPURT_LIGHT_BVH")) p->use_lbvh = atoi(e) != 0;                      /* A/B knob */
    if(const char* e = getenv("GPURT_LIGHT_VERTS")) p->use_lverts = atoi(e) != 0;                  /* A/B knob */
    if(const char* e = getenv("GPURT_SHADOW_QUEUE")) p->use_shadow_queue = atoi(e) != 0;           /* A/B knob */
    *out = p;
    return GPURT_OK;
}
int gpurt_pipe_destroy(gpurt_pipe* p) {
    if(!p) return GPURT_OK;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    pipe_free(p);
    pipe_drop_light_accel(p);
    if(p->h_counts) cudaFreeHost(p->h_counts);
    if(p->ev_counts) cudaEventDestroy(p->ev_counts);
    if(p->copy_stream) {
        cudaStreamSynchronize(p->copy_stream);
        cudaStreamDestroy(p->copy_stream);
        cudaEventDestroy(p->ev_frame);
        cudaEventDestroy(p->ev_copy);
    }
    delete p;
    return GPURT_OK;
}
int gpurt_pipe_reset_frame(gpurt_pipe* p) { /* rt.cpp:396-398 */
    if(!p) return set_error("NULL argument"), GPURT_E_INVALID;
    p->frame = -1;
    return GPURT_OK;
}
/* Multi-GPU sharding (no reference counterpart: the reference is single-GPU). */
int gpurt_pipe_set_shard(gpurt_pipe* p, uint32_t band_rows, uint32_t n_shards, uint32_t shard) {
    if(!p || (band_rows && (!n_shards || shard >= n_shards))) return set_error("bad shard arguments"), GPURT_E_INVALID;
    p->band_rows = band_rows, p->n_shards = band_rows ? n_shards : 1, p->shard = band_rows ? shard : 0;
    return GPURT_OK;
}
/* ReSTIR on a sharded frame: every shard keeps the whole previous frame (G-buffers + reservoirs) because the temporal
 * pass reprojects into it (rt.rgen:454-472); each shard stores the rows it rendered into the other shards' blocks. */
int gpurt_pipe_history_export(gpurt_pipe* p, uint32_t w, uint32_t h, void** out_ptr, uint8_t* handle, uint64_t* out_bytes) {
    if(!p || !w || !h) return set_error("bad argument"), GPURT_E_INVALID;
    GPURT_CUDA(cudaSetDevice(p->ctx->device));
    int rc = pipe_resize(p, w, h, 0);
    if(rc) return rc;
    if(out_ptr) *out_ptr = p->hist;
    if(out_bytes) *out_bytes = kHistHeader + 2ull * w * h * 96;
    if(handle) {
        cudaIpcMemHandle_t hd;
        cudaError_t e = cudaIpcGetMemHandle(&hd, p->hist);
        if(e != cudaSuccess) return set_error(std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)), GPURT_E_CUDA;
        memcpy(handle, &hd, sizeof hd);
    }
    return GPURT_OK;
}
int gpurt_pipe_history_peers(gpurt_pipe* p, uint32_t n, void* const* blocks, uint32_t halo_rows) {
    if(!p) return set_error("NULL argument"), GPURT_E_INVALID;
    if(n == 0) {
        p->n_peers = 0;
        return GPURT_OK;
    }
    if(!blocks || !p->band_rows || n != p->n_shards || n > kHistMaxShards)
        return set_error("history_peers: one block per shard of gpurt_pipe_set_shard (at most 64)"), GPURT_E_INVALID;
    if(!p->hist) return set_error("history_peers: call gpurt_pipe_history_export first"), GPURT_E_STATE;
    GPURT_CUDA(cudaSetDevice(p->ctx->device));
    cudaStream_t st = p->ctx->stream;
    if(!p->d_peers) GPURT_CUDA(cudaMalloc((void**)&p->d_peers, kHistMaxShards * sizeof(char*)));
    char* host[kHistMaxShards] = {};
    for(uint32_t s = 0; s < n; s++) {
        host[s] = s == p->shard ? p->hist : (char*)blocks[s];
        if(!host[s]) return set_error("history_peers: NULL block"), GPURT_E_INVALID;
    }
    GPURT_CUDA(cudaMemcpyAsync(p->d_peers, host, sizeof host, cudaMemcpyHostToDevice, st));
    GPURT_CUDA(cudaStreamSynchronize(st)); /* `host` leaves scope */
    p->n_peers = n, p->halo_rows = halo_rows;
    return GPURT_OK;
}
int gpurt_pipe_history_status(gpurt_pipe* p, uint32_t* out_frames_pushed, uint32_t* out_timeouts) {
    if(!p) return set_error("NULL argument"), GPURT_E_INVALID;
    if(out_frames_pushed) *out_frames_pushed = p->hist_seq;
    if(out_timeouts) {
        *out_timeouts = 0;
        if(p->hist) {
            GPURT_CUDA(cudaSetDevice(p->ctx->device));
            GPURT_CUDA(cudaStreamSynchronize(p->ctx->stream));
            GPURT_CUDA(cudaMemcpy(out_timeouts, p->hist + 8192, 4, cudaMemcpyDeviceToHost));
        }
    }
    return GPURT_OK;
}
int gpurt_pipe_frame_index(const gpurt_pipe* p, int32_t* f) {
    if(!p || !f) return set_error("NULL argument"), GPURT_E_INVALID;
    *f = p->frame;
    return GPURT_OK;
}

/* forced_frame < 0: the reference's frame logic (update_uniforms + trace).  forced_frame >= 0: render exactly that
 * frame of the progressive sequence and write its per-pixel mean to mean_out instead of accumulating. */
static int render_core(gpurt_pipe* p, const GpurtPipeParams* prm, const GpurtCamera* cam, uint32_t w, uint32_t h,
                       int forced_frame, float4* mean_out) {
    if(!p || !prm || !cam || !w || !h) return set_error("bad argument"), GPURT_E_INVALID;
    if(prm->samples_per_frame < 1 || prm->max_depth < 0) return set_error("bad sample / depth count"), GPURT_E_INVALID;
    gpurt_ctx* ctx = p->ctx;
    GPURT_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;

    /* ---- RTPipe::update_uniforms (rt.cpp:121-138) ---- */
    FrameParams F;
    std::memset(&F, 0, sizeof(F));
    std::memcpy(F.cam.V, cam->V, 64 * 4); /* V, P, iV, iP */
    F.cam.new_samples = (uint32_t)prm->res_samples;
    F.cam.temporal_multiplier = (uint32_t)prm->temporal_scale;
    {
        Mat4 oP, oV; /* identity until old_cam is first assigned, like `CameraConstants old_cam = {}` */
        if(p->old_cam_init) {
            std::memcpy(oP.data(), p->old_cam.P, 64);
            std::memcpy(oV.data(), p->old_cam.V, 64);
        }
        Mat4 pv = oP * oV;
        std::memcpy(F.cam.prev_PV, pv.data(), 64);
    }
    if(forced_frame < 0 && p->frame >= 0 && (!p->old_cam_init || std::memcmp(cam->V, p->old_cam.V, 64 * 4) != 0)) {
        p->frame = -1; /* reset_frame() */
        std::memcpy(p->old_cam.V, cam->V, 64 * 4);
        p->old_cam_init = true;
    }

    /* ---- RTPipe::trace (rt.cpp:346-394) ---- */
    int rc = pipe_resize(p, w, h, (uint32_t)prm->max_depth);
    if(rc) return rc;
    if(forced_frame < 0 && p->frame >= prm->max_frames) return 1;
    GpurtConstants& c = F.c;
    c.clear_col[0] = prm->clear[0], c.clear_col[1] = prm->clear[1], c.clear_col[2] = prm->clear[2], c.clear_col[3] = 1.0f;
    c.env_light[0] = prm->env_scale * prm->env[0], c.env_light[1] = prm->env_scale * prm->env[1];
    c.env_light[2] = prm->env_scale * prm->env[2], c.env_light[3] = 1.0f;
    c.samples = prm->samples_per_frame, c.max_depth = prm->max_depth, c.use_normal_map = prm->use_normal_map;
    c.use_metalness = prm->use_metalness, c.integrator = prm->integrator, c.brdf = prm->brdf, c.use_rr = prm->use_rr;
    c.max_frame = prm->max_frames, c.qmc = prm->use_qmc, c.use_temporal = prm->use_temporal, c.debug_view = prm->debug_view;
    c.n_lights = (int)p->accel->dscene.n_lights, c.n_objs = (int)p->accel->dscene.n_objs;
    c.frame = forced_frame < 0 ? ++p->frame : forced_frame;
    F.W = w, F.H = h;
    F.seed_val = prm->seed ^ (uint32_t)c.frame;
    F.spatial_samples = prm->spatial_samples > 0 ? (uint32_t)prm->spatial_samples : 0u, F.spatial_radius = prm->spatial_radius;
    F.light_sampling = prm->light_sampling == 1 ? 1u : 0u;

    F.band_rows = p->band_rows ? p->band_rows : h, F.n_shards = p->band_rows ? p->n_shards : 1, F.shard = p->band_rows ? p->shard : 0;
    {
        uint32_t n_local = 0, bands = (h + F.band_rows - 1) / F.band_rows;
        for(uint32_t g = F.shard; g < bands; g += F.n_shards) n_local += std::min(F.band_rows, h - g * F.band_rows) * w;
        F.n_local = n_local;
    }
    const uint32_t n = F.n_local; /* pixels rendered by this pipe */
    const int cur = p->parity, prev = cur ^ 1; /* bind_temporal_stuff ping-pong (rt.cpp:222-344) */
    const bool restir = c.integrator == 3 || c.integrator == 4;
    /* ReSTIR on a sharded frame: wait for the other shards' rows of the previous frame before reading them, and hand
     * this frame's rows over at the end (gpurt_pipe_history_peers) */
    const bool exchange = restir && p->band_rows && p->n_peers && !mean_out;
    auto history_wait = [&] {
        if(exchange && p->hist_seq > 0) k_history_wait<<<1, kHistMaxShards, 0, st>>>(p->hist, F.n_shards, p->hist_seq);
    };
    auto history_push = [&] {
        if(!exchange) return;
        const uint32_t local_rows = n / w;
        const size_t parity_off = kHistHeader + (size_t)cur * w * h * 96;
        if(local_rows)
            k_history_push<<<dim3(cdivu(6 * (size_t)w, 256), local_rows), 256, 0, st>>>(F, p->halo_rows, p->hist, p->d_peers, parity_off);
        k_history_signal<<<1, kHistMaxShards, 0, st>>>(p->d_peers, F.n_shards, F.shard, ++p->hist_seq);
    };
    if(n == 0) { /* more shards than bands: nothing to do on this rank */
        history_wait(), history_push();
        GPURT_CUDA(cudaGetLastError());
        p->parity ^= 1;
        p->last = F;
        return GPURT_OK;
    }
    ShadeCtx X;
    X.S = p->accel->dscene;
    X.nodes = (const float4*)p->accel->nodes, X.tris = p->accel->tri_wide, X.n_nodes = p->accel->n_nodes;
    X.tri_world = p->accel->tri_gid;
    if((rc = pipe_light_groups(p))) return rc;
    X.lgrp = p->lgrp_tris ? p->lgrp : nullptr, X.lgrp_off = p->lgrp_off;
    X.lnodes = nullptr, X.ltris = nullptr, X.ltri_off = nullptr, X.n_lnodes = 0;
    X.lverts = nullptr, X.lvert_off = nullptr;
    if(c.integrator != 1 && c.integrator != 2) { /* the integrators that call light_sample */
        if((rc = pipe_light_verts(p))) return rc;
        if(p->lverts_tris) X.lverts = p->lverts, X.lvert_off = p->lvert_off;
    }
    if(c.integrator == 2) { /* only MIS evaluates light_pdf */
        if((rc = pipe_light_accel(p))) return rc;
        if(p->laccel && p->laccel->n_nodes)
            X.lnodes = (const float4*)p->laccel->nodes, X.ltris = p->laccel->tri_wide, X.ltri_off = p->laccel->dscene.tri_off,
            X.n_lnodes = p->laccel->n_nodes;
    }
    X.lcdf = nullptr, X.lcdf_off = nullptr, X.n_ltris = 0;
    if(F.light_sampling && c.n_lights > 0 && c.integrator != 1) { /* every integrator that samples or evaluates lights */
        if((rc = pipe_light_cdf(p))) return rc;
        X.lcdf = p->lcdf, X.lcdf_off = p->lcdf_off, X.n_ltris = p->n_ltris;
    }
    X.prev_res = p->res[prev], X.ppos = p->gbuf[prev][0], X.pnorm = p->gbuf[prev][1], X.palb = p->gbuf[prev][2];
    X.ray_counts = p->ray_counts;

    GPURT_CUDA(cudaEventRecord(ctx->ev0, st));
    /* the first sample's camera rays are generated by the same kernel that starts the frame (k_begin_camera) unless nothing
     * will be shaded (then k_frame_begin clears the G-buffers itself); GPURT_FUSED_BEGIN=0 keeps the separate kernels.  That
     * kernel also zeroes the ray counters and the queue sizes when its grid covers them (two memset launches less). */
    static const bool fused_begin_ok = !(getenv("GPURT_FUSED_BEGIN") && atoi(getenv("GPURT_FUSED_BEGIN")) == 0);
    const bool shades = c.samples > 0 && c.max_depth > 0;
    const bool fused_begin = fused_begin_ok && shades;
    const bool fused_clear = fused_begin && p->max_counts <= n;
    if(!fused_clear) GPURT_CUDA(cudaMemsetAsync(p->ray_counts, 0, 16, st));
    history_wait();
    if(!fused_begin)
        k_frame_begin<<<cdivu(n, 256), 256, 0, st>>>(F, restir ? 1 : 0, p->acc, p->pathB, p->gbuf[cur][0],
                                                    p->gbuf[cur][1], p->gbuf[cur][2], p->res[cur]);
    /* integrate_direct and the direct-only ReSTIR end every path at its first hit (rt.rgen:393, :548): the queues of
     * the later bounces are always empty, so they are not launched */
    const uint32_t D = (c.integrator == 0 || c.integrator == 3) ? std::min(1u, (uint32_t)c.max_depth) : (uint32_t)c.max_depth;
    /* Queue sizes after the first bounce are only known on the device.  A grid sized for every pixel costs 12-14 us per
     * launch even when the queue is empty (16,200 CTAs that read the count and exit at 1080p; ncu r01r) — a fifth of a
     * ReSTIR frame on mis_test —, while a small fixed grid striding over a large queue is slower than one thread per
     * path (config-2 frame +5 %).  A progressive render traces nearly the same queues frame after frame, so the sizes of
     * the previous frame's last sample (copied back without a sync at the end of every frame) size these launches,
     * with 12.5 % + 1024 paths of head room; the kernels stride over the queue, so a low estimate is only slower. */
    const unsigned full_grid = cdivu(n, 128);
    if(p->counts_pending) {
        if(cudaEventQuery(p->ev_counts) == cudaSuccess) {
            p->est_counts.assign(p->h_counts, p->h_counts + p->max_counts);
            p->counts_pending = false;
        } else
            (void)cudaGetLastError(); /* cudaErrorNotReady is not a failure: keep the older estimate */
    }
    auto late_grid = [&](uint32_t d) -> unsigned {
        if(!p->use_est || d >= p->est_counts.size()) return full_grid;
        uint64_t est = (uint64_t)p->est_counts[d] + p->est_counts[d] / 8 + 1024;
        uint64_t g = std::max<uint64_t>((uint64_t)ctx->sm_count, (est + 127) / 128);
        return 2 * g > full_grid ? full_grid : (unsigned)g; /* more than half of the pixels alive: the plain full-size launch */
    };
    for(uint32_t s = 0; s < (uint32_t)c.samples && D > 0; s++) {
        if(!(s == 0 && fused_clear)) GPURT_CUDA(cudaMemsetAsync(p->counts, 0, (size_t)p->max_counts * 4, st));
        if(s == 0 && fused_begin)
            k_begin_camera<<<cdivu(n, 256), 256, 0, st>>>(F, restir ? 1 : 0, 0, p->acc, p->pathA, p->pathB, p->gbuf[cur][0], p->gbuf[cur][1],
                                                          p->gbuf[cur][2], p->res[cur], p->rays[0], p->queue[0], p->counts,
                                                          fused_clear ? p->max_counts : 0u, fused_clear ? p->ray_counts : nullptr);
        else
            k_gen_camera<<<cdivu(n, 256), 256, 0, st>>>(F, s, p->pathA, p->pathB, p->rays[0], p->queue[0], p->counts + 0);
        /* wavefront for the first `wave` bounces, then one tail kernel for whatever is still alive */
        /* measured (profiles/r01_tuning.md): with >= ~2.5 M paths per launch the full wavefront is fastest
         * (tail after 3 bounces: -9 %, mega-kernel: -73 %); for small shards (4K frame over 8 GPUs) every
         * launch is latency-bound and a tail after 4 bounces is 10 % faster */
        const uint32_t wave = p->wave_depth ? std::min(D, p->wave_depth) : (n > 2500000u ? D : std::min(D, 4u));
        for(uint32_t d = 0; d < wave; d++) {
            int qi = d & 1, qo = qi ^ 1;
            const unsigned grid = d == 0 ? full_grid : late_grid(d);
            if(grid == full_grid)
                k_trace_closest_indirect<false><<<grid, 128, 0, st>>>(X.nodes, X.tris, p->rays[qi], p->counts + d, p->hits, X.n_nodes);
            else
                k_trace_closest_indirect<true><<<grid, 128, 0, st>>>(X.nodes, X.tris, p->rays[qi], p->counts + d, p->hits, X.n_nodes);
#define GPURT_SHADE(I)                                                                                                   \
    k_shade<I><<<grid, 128, 0, st>>>(F, X, s, d, p->counts + d, p->queue[qi], p->rays[qi], p->hits, p->pathA,  \
                                              p->pathB, p->acc, p->gbuf[cur][0], p->gbuf[cur][1], p->gbuf[cur][2],      \
                                              p->res[cur], p->counts + d + 1, p->queue[qo], p->rays[qo])
            switch(c.integrator) {
            case 0:
                if(p->use_shadow_queue) {
                    k_shade<0, true><<<grid, 128, 0, st>>>(F, X, s, d, p->counts + d, p->queue[qi], p->rays[qi], p->hits, p->pathA,
                                                           p->pathB, p->acc, p->gbuf[cur][0], p->gbuf[cur][1], p->gbuf[cur][2],
                                                           p->res[cur], p->counts + d + 1, p->queue[qo], p->rays[qo]);
                    k_shadow_resolve<<<grid, 128, 0, st>>>(X.nodes, X.tris, X.n_nodes, p->counts + d + 1, p->queue[qo], p->rays[qo],
                                                           p->pathA, p->pathB, p->acc);
                } else
                    GPURT_SHADE(0);
                break;
            case 1: GPURT_SHADE(1); break;
            case 2: GPURT_SHADE(2); break;
            case 3: GPURT_SHADE(3); break;
            default: GPURT_SHADE(4); break;
            }
#undef GPURT_SHADE
        }
        if(wave < D) {
#define GPURT_TAIL(I)                                                                                                    \
    k_tail<I><<<late_grid(wave), 128, 0, st>>>(F, X, s, wave, p->counts + wave, p->queue[wave & 1], p->rays[wave & 1],     \
                                             p->pathA, p->pathB, p->acc, p->gbuf[cur][0], p->gbuf[cur][1],              \
                                             p->gbuf[cur][2], p->res[cur])
            switch(c.integrator) {
            case 0: GPURT_TAIL(0); break;
            case 1: GPURT_TAIL(1); break;
            case 2: GPURT_TAIL(2); break;
            case 3: GPURT_TAIL(3); break;
            default: GPURT_TAIL(4); break;
            }
#undef GPURT_TAIL
        }
        /* closest-hit rays of the wavefront = sum of queue sizes */
    }
    const bool report = D > 0 && c.samples > 0 && p->h_counts && p->max_counts <= n; /* this frame's queue sizes -> next frame */
    if(p->copy_pending) GPURT_CUDA(cudaStreamWaitEvent(st, p->ev_copy, 0)); /* an asynchronous read-back still reads the image */
    k_frame_end<<<cdivu(n, 256), 256, 0, st>>>(F, p->acc, p->image, p->gbuf[cur][0], p->gbuf[cur][1], p->gbuf[prev][0],
                                              p->gbuf[prev][1], p->gbuf[prev][2], mean_out, p->counts,
                                              report ? p->h_counts : nullptr, p->max_counts);
    if(report) {
        GPURT_CUDA(cudaEventRecord(p->ev_counts, st));
        p->counts_pending = true;
    }
    history_push();
    GPURT_CUDA(cudaEventRecord(ctx->ev1, st));
    GPURT_CUDA(cudaGetLastError());
    p->parity ^= 1;
    p->last = F;
    return GPURT_OK;
}

int gpurt_pipe_render_frame(gpurt_pipe* p, const GpurtPipeParams* prm, const GpurtCamera* cam, uint32_t w, uint32_t h) {
    return render_core(p, prm, cam, w, h, -1, nullptr);
}
int gpurt_pipe_render_frame_mean(gpurt_pipe* p, const GpurtPipeParams* prm, const GpurtCamera* cam, uint32_t w, uint32_t h,
                                 int32_t frame, void* mean_out_device) {
    if(!prm || frame < 0 || !mean_out_device) return set_error("bad argument"), GPURT_E_INVALID;
    if(prm->integrator > 2) return set_error("frame-parallel rendering needs an integrator without temporal reuse (0-2)"), GPURT_E_INVALID;
    return render_core(p, prm, cam, w, h, frame, (float4*)mean_out_device);
}
int gpurt_pipe_accumulate_mean(gpurt_pipe* p, const void* mean_device, int32_t frame, uint32_t w, uint32_t h) {
    if(!p || !mean_device || frame < 0 || !w || !h) return set_error("bad argument"), GPURT_E_INVALID;
    GPURT_CUDA(cudaSetDevice(p->ctx->device));
    if(!p->image || p->w != w || p->h != h) return set_error("accumulate: the pipe has no image of this size (render or resize first)"), GPURT_E_STATE;
    cudaStream_t st = p->ctx->stream;
    if(p->copy_pending) GPURT_CUDA(cudaStreamWaitEvent(st, p->ev_copy, 0));
    GPURT_CUDA(cudaEventRecord(p->ctx->ev0, st));
    k_accumulate_mean<<<cdivu(w * h, 256), 256, 0, st>>>(p->image, (const float4*)mean_device, w * h, frame);
    GPURT_CUDA(cudaEventRecord(p->ctx->ev1, st));
    GPURT_CUDA(cudaGetLastError());
    p->frame = frame;
    return GPURT_OK;
}

int gpurt_pipe_last_uniforms(const gpurt_pipe* p, GpurtConstants* c, GpurtCamera* cam, uint32_t* seed_val) {
    if(!p) return set_error("NULL argument"), GPURT_E_INVALID;
    if(c) *c = p->last.c;
    if(cam) *cam = p->last.cam;
    if(seed_val) *seed_val = p->last.seed_val;
    return GPURT_OK;
}

static int copy_out(gpurt_pipe* p, const void* src, size_t bytes, void* dst, int mem) {
    GPURT_CUDA(cudaSetDevice(p->ctx->device));
    cudaStream_t st = p->ctx->stream;
    if(mem == GPURT_MEM_DEVICE) {
        GPURT_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st));
        return GPURT_OK;
    }
    GPURT_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
    GPURT_CUDA(cudaStreamSynchronize(st));
    return GPURT_OK;
}
int gpurt_pipe_read_image(gpurt_pipe* p, float* out, int mem) {
    if(!p || !out || !p->image) return set_error("no image"), GPURT_E_STATE;
    return copy_out(p, p->image, (size_t)p->w * p->h * 16, out, mem);
}
/* Progressive display without a stall (the reference hands rt_target to the tonemap pass on the GPU, src/gpurt.cpp:60-66;
 * a headless consumer wants it in host memory): the copy is ordered after the frames rendered so far, runs on the pipe's
 * own copy stream and overlaps the tracing / shading of the following frames; only their accumulation kernel waits for it. */
int gpurt_pipe_read_image_async(gpurt_pipe* p, float* out_host) {
    if(!p || !out_host || !p->image) return set_error("no image"), GPURT_E_STATE;
    GPURT_CUDA(cudaSetDevice(p->ctx->device));
    if(!p->copy_stream) {
        GPURT_CUDA(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
        GPURT_CUDA(cudaEventCreateWithFlags(&p->ev_frame, cudaEventDisableTiming));
        GPURT_CUDA(cudaEventCreateWithFlags(&p->ev_copy, cudaEventDisableTiming));
    }
    GPURT_CUDA(cudaEventRecord(p->ev_frame, p->ctx->stream));
    GPURT_CUDA(cudaStreamWaitEvent(p->copy_stream, p->ev_frame, 0));
    GPURT_CUDA(cudaMemcpyAsync(out_host, p->image, (size_t)p->w * p->h * 16, cudaMemcpyDeviceToHost, p->copy_stream));
    GPURT_CUDA(cudaEventRecord(p->ev_copy, p->copy_stream));
    p->copy_pending = true;
    return GPURT_OK;
}
int gpurt_pipe_read_image_wait(gpurt_pipe* p) {
    if(!p) return set_error("NULL argument"), GPURT_E_INVALID;
    if(!p->copy_pending) return GPURT_OK;
    GPURT_CUDA(cudaSetDevice(p->ctx->device));
    GPURT_CUDA(cudaEventSynchronize(p->ev_copy));
    p->copy_pending = false;
    return GPURT_OK;
}
int gpurt_pipe_read_gbuffer(gpurt_pipe* p, int which, float* out, int mem) {
    if(!p || !out || !p->image || which < 0 || which > 2) return set_error("bad g-buffer request"), GPURT_E_STATE;
    return copy_out(p, p->gbuf[p->parity ^ 1][which], (size_t)p->w * p->h * 16, out, mem); /* last written */
}
int gpurt_pipe_read_reservoirs(gpurt_pipe* p, float* out, int mem) {
    if(!p || !out || !p->image) return set_error("no reservoirs"), GPURT_E_STATE;
    return copy_out(p, p->res[p->parity ^ 1], (size_t)p->w * p->h * 48, out, mem);
}
int gpurt_pipe_ray_counts(const gpurt_pipe* cp, uint64_t out[2]) {
    gpurt_pipe* p = const_cast<gpurt_pipe*>(cp);
    if(!p || !out || !p->image) return set_error("no frame rendered"), GPURT_E_STATE;
    GPURT_CUDA(cudaSetDevice(p->ctx->device));
    GPURT_CUDA(cudaStreamSynchronize(p->ctx->stream));
    unsigned long long c[2];
    GPURT_CUDA(cudaMemcpy(c, p->ray_counts, 16, cudaMemcpyDeviceToHost));
    out[0] = c[0], out[1] = c[1];
    return GPURT_OK;
}
int gpurt_pipe_device_image(gpurt_pipe* p, void** out) {
    if(!p || !out || !p->image) return set_error("no image"), GPURT_E_STATE;
    *out = p->image;
    return GPURT_OK;
}
/* device pointers of the ray queue traced at `bounce` in the last sample of the last frame */
int gpurt_pipe_bounce_rays(gpurt_pipe* p, uint32_t bounce, void** out_rays, uint32_t* out_count) {
    if(!p || !out_rays || !out_count || !p->image) return set_error("no frame rendered"), GPURT_E_STATE;
    if(bounce >= (uint32_t)p->last.c.max_depth) return set_error("bounce out of range"), GPURT_E_INVALID;
    GPURT_CUDA(cudaSetDevice(p->ctx->device));
    GPURT_CUDA(cudaStreamSynchronize(p->ctx->stream));
    GPURT_CUDA(cudaMemcpy(out_count, p->counts + bounce, 4, cudaMemcpyDeviceToHost));
    *out_rays = p->rays[bounce & 1];
    return GPURT_OK;
}

int gpurt_tonemap(gpurt_pipe* p, int op, float exposure, float gamma, uint8_t* out, int mem) {
    if(!p || !out || !p->image) return set_error("no image"), GPURT_E_STATE;
    GPURT_CUDA(cudaSetDevice(p->ctx->device));
    uint32_t n = p->w * p->h;
    int rc = p->ctx->scratch.reserve((size_t)n * 4);
    if(rc) return rc;
    k_tonemap<<<cdivu(n, 256), 256, 0, p->ctx->stream>>>(p->image, n, op, exposure, gamma, p->ctx->scratch.as<uchar4>());
    GPURT_CUDA(cudaGetLastError());
    return copy_out(p, p->ctx->scratch.p, (size_t)n * 4, out, mem);
}

} /* extern "C" */
