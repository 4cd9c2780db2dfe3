/*
 * api.cu — CUDA half of the C ABI (include/gpurt.h): context, scene upload, accel build, queries.
 * The host-only half (scene loading / packing, camera) is host/host_api.cpp.
 */
#include <cstdlib>
#include <cstring>

#include "device.cuh"

using namespace gpurt;

namespace gpurt {

template <typename T> static int upload(cudaStream_t st, T*& dst, const std::vector<T>& src) {
    dst = nullptr;
    size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
    GPURT_CUDA(cudaMalloc((void**)&dst, bytes));
    if(!src.empty())
        GPURT_CUDA(cudaMemcpyAsync(dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    return GPURT_OK;
}

void free_scene(DeviceScene& d) {
    void* ptrs[] = {d.verts, d.idx, d.tri_off, d.vert_off, d.descs, d.lights, d.texels, d.tex_info};
    for(void* p : ptrs)
        if(p) cudaFree(p);
    d = DeviceScene();
}

int upload_scene(gpurt_ctx* ctx, gpurt_scene* s, DeviceScene& d) {
    if(!s->pack()) return GPURT_E_INVALID;
    const PackedScene& P = s->packed;
    cudaStream_t st = ctx->stream;
    if(d.verts && d.geom_version == s->geom_version && d.n_objs == P.descs.size() && d.n_lights == P.lights.size()) {
        /* pose-only edit: geometry stays resident, Scene_Desc / Scene_Light are rewritten in place */
        if(d.version == s->version) return GPURT_OK;
        if(!P.descs.empty())
            GPURT_CUDA(cudaMemcpyAsync(d.descs, P.descs.data(), P.descs.size() * sizeof(SceneDesc), cudaMemcpyHostToDevice, st));
        if(!P.lights.empty())
            GPURT_CUDA(cudaMemcpyAsync(d.lights, P.lights.data(), P.lights.size() * sizeof(SceneLight), cudaMemcpyHostToDevice, st));
        d.version = s->version;
        GPURT_CUDA(cudaStreamSynchronize(st)); /* P.descs may be rebuilt by the next edit */
        return GPURT_OK;
    }
    free_scene(d);
    int rc;
    if((rc = upload(st, d.verts, P.verts))) return rc;
    if((rc = upload(st, d.idx, P.idx))) return rc;
    if((rc = upload(st, d.tri_off, P.tri_off))) return rc;
    if((rc = upload(st, d.vert_off, P.vert_off))) return rc;
    if((rc = upload(st, d.descs, P.descs))) return rc;
    if((rc = upload(st, d.lights, P.lights))) return rc;
    d.n_objs = (uint32_t)P.descs.size();
    d.n_tris = P.tri_off.back();
    d.n_lights = (uint32_t)P.lights.size();
    d.n_verts = (uint32_t)P.verts.size();
    d.version = s->version, d.geom_version = s->geom_version;
    /* textures: RTPipe::build_textures (src/vk/rt.cpp:430-455) */
    std::vector<uint8_t> texels;
    std::vector<uint4> info;
    for(const Texture& t : s->scene.textures) {
        info.push_back(make_uint4((unsigned)(texels.size() / 4), t.w, t.h, 0));
        texels.insert(texels.end(), t.rgba.begin(), t.rgba.end());
    }
    d.n_textures = (uint32_t)info.size();
    if((rc = upload(st, d.texels, texels))) return rc;
    if((rc = upload(st, d.tex_info, info))) return rc;
    GPURT_CUDA(cudaStreamSynchronize(st));
    return GPURT_OK;
}

/* Run `launch(d_in, d_out, count)` over n elements with caller buffers in host or device memory.
 * Device buffers: one launch, asynchronous on the context's stream.  Host buffers: the batch is cut
 * into chunks and pipelined over three streams (H2D copy -> kernel -> D2H copy), so that PCIe traffic
 * in both directions overlaps the traversal; with pinned caller memory the call approaches
 * max(H2D, kernel, D2H) instead of their sum.  ev0/ev1 bracket the device work for gpurt_last_kernel_ms. */
template <typename F>
static int run_query(gpurt_ctx* ctx, const void* in, size_t in_stride, void* out, size_t out_stride, uint64_t n,
                     int mem, F launch, bool single_pass = false) {
    GPURT_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if(mem == GPURT_MEM_DEVICE) {
        GPURT_CUDA(cudaEventRecord(ctx->ev0, st));
        int rc = launch(in, out, n);
        if(rc) return rc;
        GPURT_CUDA(cudaEventRecord(ctx->ev1, st));
        return GPURT_OK;
    }
    if(mem != GPURT_MEM_HOST) return set_error("mem must be GPURT_MEM_HOST or GPURT_MEM_DEVICE"), GPURT_E_INVALID;
    int rc;
    /* Pinned (page-locked, hence device-mapped) caller buffers and a launch that reads every element once: the kernel
     * streams its input over PCIe itself and stores its results straight into the caller's array — one launch, no
     * staging copies, no chunk pipeline with its un-overlapped first copy in and last copy out.  Every ray still crosses
     * PCIe once in and its hit once out.  GPURT_ZERO_COPY=0 keeps the staged pipeline (pageable memory always uses it). */
    static const int zc_mode = getenv("GPURT_ZERO_COPY") ? atoi(getenv("GPURT_ZERO_COPY")) : 1;
    const bool zero_copy = zc_mode == 1;
    char* mapped_out = nullptr; /* GPURT_ZERO_COPY=2 (A/B): copy engine in, results stored straight into the caller's array */
    if(zc_mode == 2 && single_pass && n) {
        cudaPointerAttributes ao;
        if(cudaPointerGetAttributes(&ao, out) == cudaSuccess && ao.type == cudaMemoryTypeHost && ao.devicePointer)
            mapped_out = (char*)ao.devicePointer;
        (void)cudaGetLastError();
    }
    if(zero_copy && single_pass && n) {
        cudaPointerAttributes ai, ao;
        const bool ok = cudaPointerGetAttributes(&ai, in) == cudaSuccess && ai.type == cudaMemoryTypeHost && ai.devicePointer &&
                        cudaPointerGetAttributes(&ao, out) == cudaSuccess && ao.type == cudaMemoryTypeHost && ao.devicePointer;
        (void)cudaGetLastError(); /* pageable memory answers with an error on older drivers: not a failure */
        if(ok) {
            GPURT_CUDA(cudaEventRecord(ctx->ev0, st));
            if((rc = launch(ai.devicePointer, ao.devicePointer, n))) return rc;
            GPURT_CUDA(cudaEventRecord(ctx->ev1, st));
            GPURT_CUDA(cudaStreamSynchronize(st));
            return GPURT_OK;
        }
    }
    if((rc = ctx->d_in.reserve(n * in_stride))) return rc;
    if((rc = ctx->d_out.reserve(n * out_stride))) return rc;
    /* chunk = a quarter of the batch, between 64 Ki and 512 Ki elements (16 MB of rays per copy): each chunk costs
     * ~35 us of host-side issue (2 copies, 1 launch, 4 event calls), the first and last chunk are not overlapped.
     * Measured on the 3.49 M-ray bench step: 64 Ki 987, 128 Ki 1250, 256 Ki 1255, 512 Ki 1364, 1 Mi 1267 Mrays/s
     * (PCIe bound at 55.4 GB/s: 1731). */
    uint64_t chunk = std::min<uint64_t>(1u << 19, std::max<uint64_t>(1u << 16, (n / 4 + 1023) & ~1023ull));
    if(const char* e = getenv("GPURT_HOST_CHUNK")) chunk = std::max<uint64_t>(1024, strtoull(e, nullptr, 10));
    GPURT_CUDA(cudaEventRecord(ctx->ev0, st));
    GPURT_CUDA(cudaStreamWaitEvent(ctx->s_h2d, ctx->ev0, 0)); /* staging buffers may still be in use on st */
    GPURT_CUDA(cudaStreamWaitEvent(ctx->s_d2h, ctx->ev0, 0));
    /* after the first copy is queued the caller's buffers are in flight: every exit drains the three streams */
    /* GPURT_HOST_TAPER=1 (A/B knob): chunks halve towards the end so that less work is left after the last host-to-device
     * copy.  Measured on the 3.49 M-ray step: 2.66 ms against 2.60 ms with equal chunks — every chunk costs ~40 us of
     * pipeline time, which outweighs the shorter tail — so it is off. */
    static const bool taper = getenv("GPURT_HOST_TAPER") && atoi(getenv("GPURT_HOST_TAPER")) != 0;
    const uint64_t min_chunk = std::min<uint64_t>(chunk, 1u << 15);
    auto body = [&]() -> int {
        for(uint64_t off = 0, cnt = 0; off < n; off += cnt) {
            const uint64_t rem = n - off;
            cnt = std::min<uint64_t>(chunk, rem);
            if(taper && rem <= 2 * chunk) cnt = std::min<uint64_t>(rem, std::max<uint64_t>(min_chunk, (rem / 2 + 1023) & ~1023ull));
            if(rem - cnt < min_chunk / 2) cnt = rem;
            char* di = (char*)ctx->d_in.p + off * in_stride;
            char* dout = mapped_out ? mapped_out + off * out_stride : (char*)ctx->d_out.p + off * out_stride;
            GPURT_CUDA(cudaMemcpyAsync(di, (const char*)in + off * in_stride, cnt * in_stride, cudaMemcpyHostToDevice, ctx->s_h2d));
            GPURT_CUDA(cudaEventRecord(ctx->ev_copy, ctx->s_h2d));
            GPURT_CUDA(cudaStreamWaitEvent(st, ctx->ev_copy, 0));
            if(int lrc = launch(di, dout, cnt)) return lrc;
            if(mapped_out) continue;
            GPURT_CUDA(cudaEventRecord(ctx->ev_kernel, st));
            GPURT_CUDA(cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_kernel, 0));
            GPURT_CUDA(cudaMemcpyAsync((char*)out + off * out_stride, dout, cnt * out_stride, cudaMemcpyDeviceToHost, ctx->s_d2h));
        }
        GPURT_CUDA(cudaEventRecord(ctx->ev1, st));
        return GPURT_OK;
    };
    rc = body();
    cudaError_t e0 = cudaStreamSynchronize(ctx->s_h2d), e1 = cudaStreamSynchronize(st), e2 = cudaStreamSynchronize(ctx->s_d2h);
    if(rc) return rc;
    for(cudaError_t e : {e0, e1, e2})
        if(e != cudaSuccess) return set_error(std::string("host-buffer query: ") + cudaGetErrorString(e)), GPURT_E_CUDA;
    return GPURT_OK;
}

} // namespace gpurt

extern "C" {

int gpurt_ctx_create(int device, gpurt_ctx** out) {
    if(!out) return set_error("out is NULL"), GPURT_E_INVALID;
    int count = 0;
    if(cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return set_error("no CUDA device: libgpurt has no CPU fallback"), GPURT_E_NO_DEVICE;
    if(device < 0 || device >= count) return set_error("device ordinal out of range"), GPURT_E_NO_DEVICE;
    GPURT_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    GPURT_CUDA(cudaGetDeviceProperties(&prop, device));
    if(prop.major < 10)
        return set_error(std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                         "; libgpurt is built for sm_100a only"),
               GPURT_E_NO_DEVICE;
    gpurt_ctx* c = new gpurt_ctx;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    { /* keep freed build temporaries cached in the stream-ordered pool instead of returning them to the OS */
        cudaMemPool_t pool;
        if(cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    cudaError_t e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    c->stream = c->own_stream;
    if(e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if(e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if(e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking);
    if(e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking);
    if(e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming);
    if(e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_kernel, cudaEventDisableTiming);
    if(e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_switch, cudaEventDisableTiming);
    if(e != cudaSuccess) { /* release whatever was created before the failure */
        set_error(std::string("gpurt_ctx_create: ") + cudaGetErrorString(e));
        gpurt_ctx_destroy(c);
        return GPURT_E_CUDA;
    }
    *out = c;
    return GPURT_OK;
}
int gpurt_ctx_destroy(gpurt_ctx* c) {
    if(!c) return GPURT_OK;
    cudaSetDevice(c->device);
    if(c->stream) cudaStreamSynchronize(c->stream);
    while(!c->gathers.empty()) gpurt_gather_destroy(c->gathers.back());
    c->d_in.release(), c->d_out.release(), c->scratch.release(), c->build_arena.release();
    if(c->pinned_word) cudaFreeHost(c->pinned_word);
    for(cudaEvent_t ev : {c->ev0, c->ev1, c->ev_copy, c->ev_kernel, c->ev_switch, c->ev_place})
        if(ev) cudaEventDestroy(ev);
    for(cudaStream_t st : {c->s_h2d, c->s_d2h, c->s_place, c->own_stream})
        if(st) cudaStreamDestroy(st);
    delete c;
    return GPURT_OK;
}
int gpurt_ctx_set_stream(gpurt_ctx* c, void* stream) {
    if(!c) return set_error("NULL context"), GPURT_E_INVALID;
    cudaStream_t next = stream == GPURT_STREAM_OWN ? c->own_stream : (cudaStream_t)stream;
    if(next != c->stream) {
        /* work queued on the old stream may still use the context's arenas / staging buffers: the new stream starts after it */
        GPURT_CUDA(cudaSetDevice(c->device));
        GPURT_CUDA(cudaEventRecord(c->ev_switch, c->stream));
        GPURT_CUDA(cudaStreamWaitEvent(next, c->ev_switch, 0));
        c->stream = next;
    }
    return GPURT_OK;
}
int gpurt_ctx_synchronize(gpurt_ctx* c) {
    if(!c) return set_error("NULL context"), GPURT_E_INVALID;
    GPURT_CUDA(cudaSetDevice(c->device));
    GPURT_CUDA(cudaStreamSynchronize(c->stream));
    return GPURT_OK;
}
int gpurt_last_kernel_ms(gpurt_ctx* c, float* ms) {
    if(!c || !ms) return set_error("NULL argument"), GPURT_E_INVALID;
    GPURT_CUDA(cudaEventSynchronize(c->ev1));
    GPURT_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return GPURT_OK;
}

/* ---- cross-process result buffers (multi-GPU sharding, include/gpurt.h) ----------------------- */
static_assert(sizeof(cudaIpcMemHandle_t) == GPURT_IPC_HANDLE_BYTES, "handle size");
int gpurt_shared_alloc(gpurt_ctx* c, uint64_t bytes, void** out, uint8_t* handle) {
    if(!c || !out || !handle || !bytes) return set_error("NULL argument"), GPURT_E_INVALID;
    GPURT_CUDA(cudaSetDevice(c->device));
    void* p = nullptr;
    GPURT_CUDA(cudaMalloc(&p, bytes)); /* a dedicated allocation: the handle names its base address */
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if(e != cudaSuccess) {
        cudaFree(p);
        return set_error(std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)), GPURT_E_CUDA;
    }
    memcpy(handle, &h, sizeof h);
    *out = p;
    return GPURT_OK;
}
int gpurt_shared_free(gpurt_ctx* c, void* p) {
    if(!c) return set_error("NULL context"), GPURT_E_INVALID;
    GPURT_CUDA(cudaSetDevice(c->device));
    GPURT_CUDA(cudaFree(p));
    return GPURT_OK;
}
int gpurt_shared_open(gpurt_ctx* c, const uint8_t* handle, void** out) {
    if(!c || !handle || !out) return set_error("NULL argument"), GPURT_E_INVALID;
    GPURT_CUDA(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    GPURT_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return GPURT_OK;
}
int gpurt_shared_close(gpurt_ctx* c, void* p) {
    if(!c) return set_error("NULL context"), GPURT_E_INVALID;
    GPURT_CUDA(cudaSetDevice(c->device));
    GPURT_CUDA(cudaStreamSynchronize(c->stream));
    GPURT_CUDA(cudaIpcCloseMemHandle(p));
    return GPURT_OK;
}

/* ---- accel ------------------------------------------------------------------------------------ */
int gpurt_accel_build(gpurt_scene* s, uint32_t flags, gpurt_accel** out) {
    if(!s || !out) return set_error("NULL argument"), GPURT_E_INVALID;
    if(!s->ctx) return set_error("scene was created without a context: no device to build on"), GPURT_E_NO_DEVICE;
    gpurt_accel* A = new gpurt_accel;
    A->ctx = s->ctx;
    A->scene = s;
    A->flags = flags;
    int rc = build_accel_device(A);
    if(rc == GPURT_OK && A->depth > 60) rc = (set_error("wide BVH deeper than the traversal stack"), GPURT_E_STATE);
    if(rc) {
        free_accel_device(A);
        delete A;
        return rc;
    }
    *out = A;
    return GPURT_OK;
}
int gpurt_accel_update(gpurt_accel* A) {
    if(!A) return set_error("NULL argument"), GPURT_E_INVALID;
    GPURT_CUDA(cudaSetDevice(A->ctx->device));
    if(!A->scene->pack()) return GPURT_E_INVALID;
    if(A->scene->geom_version != A->dscene.geom_version) { /* geometry changed: buffer sizes change too */
        GPURT_CUDA(cudaStreamSynchronize(A->ctx->stream));
        free_accel_device(A);
    }
    int rc = build_accel_device(A);
    if(rc == GPURT_OK && A->depth > 60) rc = (set_error("wide BVH deeper than the traversal stack"), GPURT_E_STATE);
    return rc;
}
/* pose-only edit, keep order and topology */
int gpurt_accel_refit(gpurt_accel* A) {
    if(!A) return set_error("NULL argument"), GPURT_E_INVALID;
    GPURT_CUDA(cudaSetDevice(A->ctx->device));
    if(!A->scene->pack()) return GPURT_E_INVALID;
    if(A->scene->geom_version != A->dscene.geom_version || A->scene->packed.tri_off.back() != A->n || !A->parent)
        return set_error("geometry changed since the last build: gpurt_accel_update() is needed"), GPURT_E_STATE;
    if(A->n < 2) return gpurt_accel_update(A);
    int rc = build_accel_device(A, true);
    if(rc == GPURT_OK && A->depth > 60) rc = (set_error("wide BVH deeper than the traversal stack"), GPURT_E_STATE);
    return rc;
}
int gpurt_accel_update_auto(gpurt_accel* A, float max_cost_growth) {
    if(!A) return set_error("NULL argument"), GPURT_E_INVALID;
    if(!A->scene->pack()) return GPURT_E_INVALID;
    const bool pose_only = A->scene->geom_version == A->dscene.geom_version && A->scene->packed.tri_off.back() == A->n && A->parent && A->n >= 2;
    if(!pose_only) return gpurt_accel_update(A);
    int rc = gpurt_accel_refit(A);
    if(rc) return rc;
    if(!(max_cost_growth > 0.0f)) max_cost_growth = 1.25f;
    if(A->tree_cost > max_cost_growth * A->tree_cost_at_build) return gpurt_accel_update(A); /* the old topology no longer fits */
    return GPURT_OK;
}
int gpurt_accel_sync_scene(gpurt_accel* A) {
    if(!A) return set_error("NULL argument"), GPURT_E_INVALID;
    GPURT_CUDA(cudaSetDevice(A->ctx->device));
    if(!A->scene->pack()) return GPURT_E_INVALID;
    if(A->scene->packed.tri_off.back() != A->n) return set_error("geometry changed: gpurt_accel_update() is needed"), GPURT_E_STATE;
    if(A->scene->geom_version != A->dscene.geom_version) GPURT_CUDA(cudaStreamSynchronize(A->ctx->stream)); /* buffers are replaced */
    return upload_scene(A->ctx, A->scene, A->dscene);
}
int gpurt_accel_destroy(gpurt_accel* A) {
    if(!A) return GPURT_OK;
    cudaSetDevice(A->ctx->device);
    cudaStreamSynchronize(A->ctx->stream);
    free_accel_device(A);
    delete A;
    return GPURT_OK;
}
int gpurt_accel_info(const gpurt_accel* A, GpurtAccelInfo* o) {
    if(!A || !o) return set_error("NULL argument"), GPURT_E_INVALID;
    std::memset(o, 0, sizeof(*o));
    o->n_tris = A->n;
    o->n_objs = A->dscene.n_objs;
    o->n_bvh2_nodes = A->n > 1 ? A->n - 1 : 0;
    o->n_wide_nodes = A->n_nodes;
    o->wide_depth = A->depth;
    for(int k = 0; k < 3; k++) o->scene_min[k] = A->scene_box[k], o->scene_max[k] = A->scene_box[3 + k];
    o->inflation = A->inflate;
    o->build_ms = A->build_ms;
    o->node_bytes = (uint64_t)A->n_nodes * sizeof(Node8);
    o->tri_bytes = (uint64_t)A->n * 48;
    o->tree_cost = A->tree_cost, o->tree_cost_at_build = A->tree_cost_at_build, o->refits = A->refits;
    return GPURT_OK;
}
static int d2h(const gpurt_accel* A, void* dst, const void* src, size_t bytes) {
    if(!bytes) return GPURT_OK;
    GPURT_CUDA(cudaSetDevice(A->ctx->device));
    GPURT_CUDA(cudaStreamSynchronize(A->ctx->stream));
    GPURT_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return GPURT_OK;
}
int gpurt_accel_get_prim_order(const gpurt_accel* A, uint32_t* out) {
    if(!A || !out) return set_error("NULL argument"), GPURT_E_INVALID;
    return d2h(A, out, A->order, (size_t)A->n * 4);
}
int gpurt_accel_get_morton_keys(const gpurt_accel* A, uint64_t* out) {
    if(!A || !out) return set_error("NULL argument"), GPURT_E_INVALID;
    return d2h(A, out, A->keys, (size_t)A->n * 8);
}
int gpurt_accel_get_bvh2(const gpurt_accel* A, int32_t* left, int32_t* right, float* boxes6) {
    if(!A || !left || !right || !boxes6) return set_error("NULL argument"), GPURT_E_INVALID;
    if(A->n < 2) return GPURT_OK;
    size_t m = A->n - 1;
    int rc;
    if((rc = d2h(A, left, A->left, m * 4))) return rc;
    if((rc = d2h(A, right, A->right, m * 4))) return rc;
    std::vector<float4> lo(m), hi(m);
    if((rc = d2h(A, lo.data(), A->node_lo, m * 16))) return rc;
    if((rc = d2h(A, hi.data(), A->node_hi, m * 16))) return rc;
    for(size_t i = 0; i < m; i++) {
        float* o = boxes6 + 6 * i;
        o[0] = lo[i].x, o[1] = lo[i].y, o[2] = lo[i].z, o[3] = hi[i].x, o[4] = hi[i].y, o[5] = hi[i].z;
    }
    return GPURT_OK;
}

/* ---- queries ---------------------------------------------------------------------------------- */
/* true when a batch of n elements is answered by one pass over its input (order.cu does not re-order it) */
static bool single_pass(const gpurt_accel* A, uint64_t n, bool any_bvh_size = false) {
    const size_t bvh_bytes = (size_t)A->n_nodes * sizeof(Node8) + (size_t)A->n * 48;
    return n < (1u << 20) || bvh_bytes <= (any_bvh_size ? (4u << 20) : (64u << 20)); /* order.cu kOrderMinBvhBytes[Points] */
}
int gpurt_trace_closest(gpurt_accel* A, const GpurtRay* rays, uint64_t n, GpurtHit* hits, int mem) {
    if(!A || (n && (!rays || !hits))) return set_error("NULL argument"), GPURT_E_INVALID;
    return run_query(A->ctx, rays, sizeof(GpurtRay), hits, sizeof(GpurtHit), n, mem,
                     [&](const void* i, void* o, uint64_t c) { return launch_trace_closest(A, (const float4*)i, c, (float4*)o); },
                     single_pass(A, n));
}
int gpurt_trace_closest_bvh2(gpurt_accel* A, const GpurtRay* rays, uint64_t n, GpurtHit* hits, int mem) {
    if(!A || (n && (!rays || !hits))) return set_error("NULL argument"), GPURT_E_INVALID;
    return run_query(A->ctx, rays, sizeof(GpurtRay), hits, sizeof(GpurtHit), n, mem,
                     [&](const void* i, void* o, uint64_t c) { return launch_trace_closest_bvh2(A, (const float4*)i, c, (float4*)o); });
}
int gpurt_trace_any(gpurt_accel* A, const GpurtRay* rays, uint64_t n, uint8_t* occ, int mem) {
    if(!A || (n && (!rays || !occ))) return set_error("NULL argument"), GPURT_E_INVALID;
    return run_query(A->ctx, rays, sizeof(GpurtRay), occ, 1, n, mem,
                     [&](const void* i, void* o, uint64_t c) { return launch_trace_any(A, (const float4*)i, c, (uint8_t*)o); },
                     single_pass(A, n));
}
int gpurt_closest_points(gpurt_accel* A, const GpurtQuery* q, uint64_t n, GpurtClosestPoint* res, int mem) {
    if(!A || (n && (!q || !res))) return set_error("NULL argument"), GPURT_E_INVALID;
    return run_query(A->ctx, q, sizeof(GpurtQuery), res, sizeof(GpurtClosestPoint), n, mem,
                     [&](const void* i, void* o, uint64_t c) { return launch_closest_points(A, (const float4*)i, c, (float4*)o); },
                     single_pass(A, n, true)); /* point batches may be re-ordered on scenes > 4 MB (cpq.cu) */
}
int gpurt_trace_closest_stats(gpurt_accel* A, const GpurtRay* rays, uint64_t n, GpurtHit* hits,
                              GpurtTraceStats* out) {
    if(!A || !out || (n && (!rays || !hits))) return set_error("NULL argument"), GPURT_E_INVALID;
    gpurt_ctx* ctx = A->ctx;
    GPURT_CUDA(cudaSetDevice(ctx->device));
    int rc = ctx->scratch.reserve(64);
    if(rc) return rc;
    unsigned long long* d = ctx->scratch.as<unsigned long long>();
    GPURT_CUDA(cudaMemsetAsync(d, 0, 32, ctx->stream));
    rc = run_query(ctx, rays, sizeof(GpurtRay), hits, sizeof(GpurtHit), n, GPURT_MEM_DEVICE,
                   [&](const void* i, void* o, uint64_t c) { return launch_trace_closest_stats(A, (const float4*)i, c, (float4*)o, d); });
    if(rc) return rc;
    unsigned long long h[4];
    GPURT_CUDA(cudaMemcpyAsync(h, d, 32, cudaMemcpyDeviceToHost, ctx->stream));
    GPURT_CUDA(cudaStreamSynchronize(ctx->stream));
    out->rays = n, out->nodes_visited = h[0], out->tris_tested = h[1], out->hits = h[2];
    return GPURT_OK;
}
int gpurt_closest_points_stats(gpurt_accel* A, const GpurtQuery* q, uint64_t n, GpurtTraceStats* out) {
    if(!A || !out || (n && !q)) return set_error("NULL argument"), GPURT_E_INVALID;
    gpurt_ctx* ctx = A->ctx;
    GPURT_CUDA(cudaSetDevice(ctx->device));
    int rc = ctx->scratch.reserve(64);
    if(rc) return rc;
    unsigned long long* d = ctx->scratch.as<unsigned long long>();
    GPURT_CUDA(cudaMemsetAsync(d, 0, 32, ctx->stream));
    if((rc = launch_closest_points_stats(A, (const float4*)q, n, d))) return rc;
    unsigned long long h[4];
    GPURT_CUDA(cudaMemcpyAsync(h, d, 32, cudaMemcpyDeviceToHost, ctx->stream));
    GPURT_CUDA(cudaStreamSynchronize(ctx->stream));
    out->rays = n, out->nodes_visited = h[0], out->tris_tested = h[1], out->hits = h[2];
    return GPURT_OK;
}

} /* extern "C" */
