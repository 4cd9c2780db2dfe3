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

#include "../host/camera.h"
#include "shade.cuh"

using namespace gpurt;

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
    uint32_t wave_depth = 0;                          /* bounces run as wavefronts before k_tail; 0 = by size */
};

namespace gpurt {

static inline unsigned cdivu(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

__global__ void __launch_bounds__(256) k_frame_begin(const __grid_constant__ FrameParams P, int restir, float4* acc,
                                                     float4* pathB, float4* gpos, float4* gnorm, float4* galb,
                                                     float4* res_out) {
    uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    if(li >= P.n_local) return;
    const uint32_t i = shard_pixel(P, li);
    /* tea(pixel, seed) — rtcommon.glsl:99-109; Q1: seed = user seed ^ frame replaces clockARB() */
    uint32_t v0 = i, v1 = P.seed_val, s0 = 0;
    for(uint32_t k = 0; k < 16; k++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    acc[i] = make_float4(0, 0, 0, 0);
    pathB[i] = make_float4(1, 1, 1, __uint_as_float(v0));
    gpos[i] = gnorm[i] = galb[i] = make_float4(0, 0, 0, 1); /* rt.rgen:573, :674-676 */
    if(restir) {
        res_out[3ull * i] = res_out[3ull * i + 1] = make_float4(0, 0, 0, 0);
        res_out[3ull * i + 2] = make_float4(0, 0, 0, __uint_as_float(0u));
    }
}

__global__ void __launch_bounds__(256) k_gen_camera(const __grid_constant__ FrameParams P, uint32_t s,
                                                    float4* pathA, float4* pathB, float4* rays,
                                                    uint32_t* queue, uint32_t* count) {
    uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    if(li == 0) *count = P.n_local;
    if(li >= P.n_local) return;
    const uint32_t i = shard_pixel(P, li);
    ShadeCtx dummy{};
    Shader sh(dummy, P);
    float4 B = pathB[i];
    sh.seed = __float_as_uint(B.w);
    F3 d = sh.make_camera_ray(s, i % P.W, i / P.W);
    F4 co = mul4(P.cam.iV, 0.0f, 0.0f, 0.0f, 1.0f); /* rt.rgen:572 */
    rays[2ull * li] = make_float4(co.x, co.y, co.z, kEps);
    rays[2ull * li + 1] = make_float4(d.x, d.y, d.z, kLargeDist);
    queue[li] = i;
    pathA[i] = make_float4(0, 0, 0, 1.0f);                           /* trace.acc, trace.mis */
    pathB[i] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(sh.seed)); /* trace.throughput, rng */
}

__global__ void __launch_bounds__(128) k_trace_closest_indirect(const float4* __restrict__ nodes,
                                                                const float4* __restrict__ tris,
                                                                const float4* __restrict__ rays,
                                                                const uint32_t* __restrict__ count,
                                                                float4* __restrict__ hits, unsigned n_nodes) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= *count) return;
    float4 a = __ldg(rays + 2ull * i), b = __ldg(rays + 2ull * i + 1);
    HitRec h;
    h.t = a.w, h.u = h.v = 0, h.gid = kNoHit;
    if(n_nodes) traverse8<false, false>(nodes, tris, f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), a.w, b.w, h, nullptr);
    hits[i] = make_float4(h.gid == kNoHit ? GPURT_INF : h.t, h.u, h.v, u2f(h.gid));
}

/* One iteration of rt.rgen's bounce loop body after traceRayEXT (rt.rgen:591-627) for the path of pixel
 * `pix`: miss handling, hit_info / mat_info / shade_info, G-buffer capture, the selected integrator and
 * Russian roulette.  Returns true when the path ends here (`break` in the shader). */
/* INTEG = the integrator, fixed at compile time: one kernel per integrator instead of one kernel carrying all five
 * (the five-way kernel needs 128 registers -> 23 % occupancy; ncu capture prof_frame_r1k) */
template <int INTEG>
__device__ __forceinline__ bool shade_step(const FrameParams& P, Shader& sh, TraceInfo& trace, uint32_t s, uint32_t depth,
                                           uint32_t pix, float4 h, float4* gpos, float4* gnorm, float4* galb,
                                           float4* res_cur) {
    const bool restir = INTEG == 3 || INTEG == 4;
    bool broke = false;
    uint32_t gid = f2u(h.w);
    if(gid == kNoHit) { /* rt.rgen:591-598 */
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
    if(restir && depth == 0) sh.prev_res = Shader::res_load(res_cur + 3ull * pix);
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

template <int INTEG>
__global__ void __launch_bounds__(128) k_shade(const __grid_constant__ FrameParams P,
                                               const __grid_constant__ ShadeCtx X, uint32_t s, uint32_t depth,
                                               const uint32_t* __restrict__ count_in, const uint32_t* __restrict__ queue_in,
                                               const float4* __restrict__ rays_in, const float4* __restrict__ hits,
                                               float4* pathA, float4* pathB, float4* acc, float4* gpos, float4* gnorm,
                                               float4* galb, float4* res_cur, uint32_t* count_out, uint32_t* queue_out,
                                               float4* rays_out) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    bool live = k < *count_in;
    bool cont = false;
    uint32_t pix = 0;
    TraceInfo trace;
    Shader sh(X, P);
    if(live) {
        pix = queue_in[k];
        float4 r0 = rays_in[2ull * k], r1 = rays_in[2ull * k + 1], h = hits[k];
        float4 A = pathA[pix], B = pathB[pix];
        trace.o = F3{r0.x, r0.y, r0.z}, trace.d = F3{r1.x, r1.y, r1.z};
        trace.acc = F3{A.x, A.y, A.z}, trace.mis = A.w;
        trace.throughput = F3{B.x, B.y, B.z}, trace.depth = depth;
        sh.seed = __float_as_uint(B.w);
        bool broke = shade_step<INTEG>(P, sh, trace, s, depth, pix, h, gpos, gnorm, galb, res_cur);
        cont = !broke && trace.depth + 1 < (uint32_t)P.c.max_depth;
        if(cont) {
            pathA[pix] = make_float4(trace.acc.x, trace.acc.y, trace.acc.z, trace.mis);
            pathB[pix] = make_float4(trace.throughput.x, trace.throughput.y, trace.throughput.z, __uint_as_float(sh.seed));
        } else { /* rt.rgen:630: acc += trace.acc; the RNG stream continues into the next sample */
            float4 a = acc[pix];
            acc[pix] = make_float4(a.x + trace.acc.x, a.y + trace.acc.y, a.z + trace.acc.z, 0.0f);
            pathB[pix] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(sh.seed));
        }
    }
    /* warp-aggregated compaction of the surviving paths */
    unsigned m = __ballot_sync(0xffffffffu, cont);
    unsigned lane = threadIdx.x & 31;
    uint32_t base = 0;
    if(m) {
        if(lane == (unsigned)__ffs(m) - 1) base = atomicAdd(count_out, (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    }
    if(cont) {
        uint32_t slot = base + __popc(m & ((1u << lane) - 1u));
        queue_out[slot] = pix;
        rays_out[2ull * slot] = make_float4(trace.o.x, trace.o.y, trace.o.z, kEps);
        rays_out[2ull * slot + 1] = make_float4(trace.d.x, trace.d.y, trace.d.z, kLargeDist);
    }
    /* ray accounting: this thread's wavefront ray + its inline rays, one atomic per warp */
    unsigned nc = __reduce_add_sync(0xffffffffu, live ? 1u + sh.n_closest : 0u);
    unsigned na = __reduce_add_sync(0xffffffffu, live ? sh.n_any : 0u);
    if(lane == 0) {
        if(nc) atomicAdd(X.ray_counts + 0, (unsigned long long)nc);
        if(na) atomicAdd(X.ray_counts + 1, (unsigned long long)na);
    }
}

/* Tail of the wavefront: after the first bounces only a fraction of the paths is alive and one
 * trace + one shade launch per bounce become latency-bound (worst when a frame is sharded over 8 GPUs).
 * Here every remaining path runs its bounce loop to the end in one thread: same shade_step, same
 * traverse8, same per-pixel RNG stream, so the image does not change — only the launch count does. */
template <int INTEG>
__global__ void __launch_bounds__(128) k_tail(const __grid_constant__ FrameParams P, const __grid_constant__ ShadeCtx X,
                                              uint32_t s, uint32_t depth0, const uint32_t* __restrict__ count_in,
                                              const uint32_t* __restrict__ queue_in, const float4* __restrict__ rays_in,
                                              float4* pathA, float4* pathB, float4* acc, float4* gpos, float4* gnorm,
                                              float4* galb, float4* res_cur) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    bool live = k < *count_in;
    Shader sh(X, P);
    unsigned n_wave = 0;
    if(live) {
        uint32_t pix = queue_in[k];
        float4 r0 = rays_in[2ull * k], r1 = rays_in[2ull * k + 1];
        float4 A = pathA[pix], B = pathB[pix];
        TraceInfo trace;
        trace.o = F3{r0.x, r0.y, r0.z}, trace.d = F3{r1.x, r1.y, r1.z};
        trace.acc = F3{A.x, A.y, A.z}, trace.mis = A.w;
        trace.throughput = F3{B.x, B.y, B.z};
        sh.seed = __float_as_uint(B.w);
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
        float4 a = acc[pix];
        acc[pix] = make_float4(a.x + trace.acc.x, a.y + trace.acc.y, a.z + trace.acc.z, 0.0f);
        pathB[pix] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(sh.seed));
    }
    unsigned nc = __reduce_add_sync(0xffffffffu, live ? n_wave + sh.n_closest : 0u);
    unsigned na = __reduce_add_sync(0xffffffffu, live ? sh.n_any : 0u);
    if((threadIdx.x & 31) == 0) {
        if(nc) atomicAdd(X.ray_counts + 0, (unsigned long long)nc);
        if(na) atomicAdd(X.ray_counts + 1, (unsigned long long)na);
    }
}

/* rt.rgen:640-645: progressive mean over frames */
__device__ __forceinline__ float4 accumulate_frame(float4 old, F3 avg, int frame) {
    if(frame > 0) {
        F3 m = mix3(F3{old.x, old.y, old.z}, avg, 1.0f / (float)(frame + 1));
        return make_float4(m.x, m.y, m.z, 1.0f);
    }
    return make_float4(avg.x, avg.y, avg.z, 1.0f);
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
                                                   float4* mean_out) {
    uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    if(li >= P.n_local) return;
    const uint32_t i = shard_pixel(P, li);
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
    void* ptrs[] = {p->image, p->res[0], p->res[1], p->gbuf[0][0], p->gbuf[0][1], p->gbuf[0][2], p->gbuf[1][0],
                    p->gbuf[1][1], p->gbuf[1][2], p->acc, p->pathA, p->pathB, p->rays[0], p->rays[1], p->hits,
                    p->queue[0], p->queue[1], p->counts, p->ray_counts};
    for(void* q : ptrs)
        if(q) cudaFree(q);
    return GPURT_OK;
}

/* RTPipe::resize_temporal_stuff (rt.cpp:178-220) + rt_target (gpurt.cpp:189-193) */
static int pipe_resize(gpurt_pipe* p, uint32_t w, uint32_t h, uint32_t max_depth) {
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
        for(int k = 0; k < 2; k++) {
            if((rc = alloc(p->res[k], n * 48))) return rc;
            for(int g = 0; g < 3; g++)
                if((rc = alloc(p->gbuf[k][g], n * 16))) return rc;
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
        p->max_counts = need_counts;
    }
    p->w = w, p->h = h;
    return GPURT_OK;
}

static void upload_lut_once() {
    static bool done = false;
    if(done) return;
    float lut[256];
    for(int i = 0; i < 256; i++) {
        double c = i / 255.0;
        lut[i] = (float)(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
    }
    cudaMemcpyToSymbol(c_srgb_lut, lut, sizeof(lut));
    done = true;
}

} // namespace gpurt

extern "C" {

int gpurt_pipe_create(gpurt_scene* scene, gpurt_accel* accel, gpurt_pipe** out) {
    if(!scene || !accel || !out) return set_error("NULL argument"), GPURT_E_INVALID;
    if(accel->scene != scene) return set_error("accel was built from a different scene"), GPURT_E_STATE;
    gpurt_pipe* p = new gpurt_pipe;
    p->ctx = accel->ctx, p->scene = scene, p->accel = accel;
    std::memset(&p->old_cam, 0, sizeof(p->old_cam));
    std::memset(&p->last, 0, sizeof(p->last));
    GPURT_CUDA(cudaSetDevice(p->ctx->device));
    upload_lut_once();
    if(const char* e = getenv("GPURT_WAVE_DEPTH")) p->wave_depth = (uint32_t)std::max(1, atoi(e)); /* tuning knob */
    *out = p;
    return GPURT_OK;
}
int gpurt_pipe_destroy(gpurt_pipe* p) {
    if(!p) return GPURT_OK;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    pipe_free(p);
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

    F.band_rows = p->band_rows ? p->band_rows : h, F.n_shards = p->band_rows ? p->n_shards : 1, F.shard = p->band_rows ? p->shard : 0;
    {
        uint32_t n_local = 0, bands = (h + F.band_rows - 1) / F.band_rows;
        for(uint32_t g = F.shard; g < bands; g += F.n_shards) n_local += std::min(F.band_rows, h - g * F.band_rows) * w;
        F.n_local = n_local;
    }
    const uint32_t n = F.n_local; /* pixels rendered by this pipe */
    if(n == 0) { /* more shards than bands: nothing to do on this rank */
        p->parity ^= 1;
        p->last = F;
        return GPURT_OK;
    }
    const int cur = p->parity, prev = cur ^ 1; /* bind_temporal_stuff ping-pong (rt.cpp:222-344) */
    const bool restir = c.integrator == 3 || c.integrator == 4;
    ShadeCtx X;
    X.S = p->accel->dscene;
    X.nodes = (const float4*)p->accel->nodes, X.tris = p->accel->tri_wide, X.n_nodes = p->accel->n_nodes;
    X.tri_world = p->accel->tri_gid;
    X.prev_res = p->res[prev], X.ppos = p->gbuf[prev][0], X.pnorm = p->gbuf[prev][1], X.palb = p->gbuf[prev][2];
    X.ray_counts = p->ray_counts;

    GPURT_CUDA(cudaEventRecord(ctx->ev0, st));
    GPURT_CUDA(cudaMemsetAsync(p->ray_counts, 0, 16, st));
    k_frame_begin<<<cdivu(n, 256), 256, 0, st>>>(F, restir ? 1 : 0, p->acc, p->pathB, p->gbuf[cur][0],
                                                p->gbuf[cur][1], p->gbuf[cur][2], p->res[cur]);
    const uint32_t D = (uint32_t)c.max_depth;
    for(uint32_t s = 0; s < (uint32_t)c.samples && D > 0; s++) {
        GPURT_CUDA(cudaMemsetAsync(p->counts, 0, (size_t)p->max_counts * 4, st));
        k_gen_camera<<<cdivu(n, 256), 256, 0, st>>>(F, s, p->pathA, p->pathB, p->rays[0], p->queue[0], p->counts + 0);
        /* wavefront for the first `wave` bounces, then one tail kernel for whatever is still alive */
        /* measured (profiles/r01_tuning.md): with >= ~2.5 M paths per launch the full wavefront is fastest
         * (tail after 3 bounces: -9 %, mega-kernel: -73 %); for small shards (4K frame over 8 GPUs) every
         * launch is latency-bound and a tail after 4 bounces is 10 % faster */
        const uint32_t wave = p->wave_depth ? std::min(D, p->wave_depth) : (n > 2500000u ? D : std::min(D, 4u));
        for(uint32_t d = 0; d < wave; d++) {
            int qi = d & 1, qo = qi ^ 1;
            k_trace_closest_indirect<<<cdivu(n, 128), 128, 0, st>>>(X.nodes, X.tris, p->rays[qi], p->counts + d, p->hits,
                                                                   X.n_nodes);
#define GPURT_SHADE(I)                                                                                                   \
    k_shade<I><<<cdivu(n, 128), 128, 0, st>>>(F, X, s, d, p->counts + d, p->queue[qi], p->rays[qi], p->hits, p->pathA,  \
                                              p->pathB, p->acc, p->gbuf[cur][0], p->gbuf[cur][1], p->gbuf[cur][2],      \
                                              p->res[cur], p->counts + d + 1, p->queue[qo], p->rays[qo])
            switch(c.integrator) {
            case 0: GPURT_SHADE(0); break;
            case 1: GPURT_SHADE(1); break;
            case 2: GPURT_SHADE(2); break;
            case 3: GPURT_SHADE(3); break;
            default: GPURT_SHADE(4); break;
            }
#undef GPURT_SHADE
        }
        if(wave < D) {
#define GPURT_TAIL(I)                                                                                                    \
    k_tail<I><<<cdivu(n, 128), 128, 0, st>>>(F, X, s, wave, p->counts + wave, p->queue[wave & 1], p->rays[wave & 1],     \
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
    k_frame_end<<<cdivu(n, 256), 256, 0, st>>>(F, p->acc, p->image, p->gbuf[cur][0], p->gbuf[cur][1], p->gbuf[prev][0],
                                              p->gbuf[prev][1], p->gbuf[prev][2], mean_out);
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
