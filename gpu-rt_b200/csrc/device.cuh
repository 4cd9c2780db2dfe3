/*
 * device.cuh — CUDA-side internals of libgpurt.so: context, device scene, acceleration structure.
 */
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../host/internal.h"
#include "bvh8.cuh"
#include "scene_view.cuh"

struct gpurt_gather;

namespace gpurt {

#define GPURT_CUDA(call)                                                                           \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if(e_ != cudaSuccess) {                                                                    \
            gpurt::set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                  \
            return GPURT_E_CUDA;                                                                   \
        }                                                                                          \
    } while(0)

/* grow-only device / pinned buffers */
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <typename T> T* as() const { return (T*)p; }
};

} // namespace gpurt

struct gpurt_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    /* host-buffer calls: copy streams and hand-over events of the H2D -> kernel -> D2H pipeline */
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_copy = nullptr, ev_kernel = nullptr;
    cudaEvent_t ev_switch = nullptr; /* orders a newly selected stream after the old one (gpurt_ctx_set_stream) */
    std::vector<gpurt_gather*> gathers; /* gpurt_gather_create / _open on this context */
    cudaStream_t s_place = nullptr;  /* result placement into another GPU's memory while the next slice is computed (order.cu) */
    cudaEvent_t ev_place = nullptr;
    /* staging for GPURT_MEM_HOST calls */
    gpurt::DevBuf d_in, d_out;
    gpurt::DevBuf scratch;
    gpurt::DevBuf build_arena; /* temporaries of gpurt_accel_build / gpurt_accel_update */
    unsigned* pinned_word = nullptr; /* mapped host memory for single-word read-backs (sah_build.cu) */
};

namespace gpurt {

int upload_scene(gpurt_ctx* ctx, gpurt_scene* s, DeviceScene& out);
void free_scene(DeviceScene& d);

size_t scan_tmp_bytes(size_t n);

/* stable LSD radix sort of 64-bit keys with 32-bit payload, 8 bits per pass */
int radix_sort_u64(cudaStream_t st, uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp,
                   uint32_t* vals_tmp, size_t n, int passes, DevBuf& tmp, int sm_count);

} // namespace gpurt

struct gpurt_accel {
    gpurt_ctx* ctx = nullptr;
    gpurt_scene* scene = nullptr;
    gpurt::DeviceScene dscene;
    uint32_t n = 0;
    uint32_t flags = 0;
    /* gid order */
    float4* tri_gid = nullptr; /* 3 x float4 per triangle */
    float4* tri_lo = nullptr;
    float4* tri_hi = nullptr;
    /* canonical order */
    uint64_t* keys = nullptr;
    uint32_t* order = nullptr;
    /* binary tree (kept for gpurt_accel_refit: parents and leaf ranges too) */
    int *left = nullptr, *right = nullptr;
    int *parent = nullptr, *range_first = nullptr, *range_last = nullptr;
    float* tree_cost_dev = nullptr; /* sum of the inner nodes' box areas (the SAH cost up to constants), one float */
    float tree_cost = 0, tree_cost_at_build = 0;
    uint32_t refits = 0; /* since the last full build */
    float4 *node_lo = nullptr, *node_hi = nullptr;
    /* wide BVH */
    gpurt::Node8* nodes = nullptr;
    float4* tri_wide = nullptr;
    uint32_t n_nodes = 0, depth = 0;
    float scene_box[6] = {0, 0, 0, 0, 0, 0};
    float inflate = 0;
    float min_inflate = 0; /* lower bound for `inflate` (a light BVH is queried from anywhere in the main scene) */
    float build_ms = 0;
};

namespace gpurt {
int build_accel_device(gpurt_accel* A, bool refit_only = false);
/* sah_build.cu: the binned-SAH tree of host/sah_split.h on the device */
size_t sah_split_tmp_bytes(size_t n, int sm_count);
int build_sah_split_device(gpurt_ctx* ctx, const float4* tri_lo, const float4* tri_hi, unsigned n, uint32_t* order, uint64_t* keys,
                           int* left, int* right, int* parent, int* range_first, int* range_last, void* tmp, size_t tmp_bytes,
                           unsigned* levels_out);
void free_accel_device(gpurt_accel* A);

/* order.cu: processing order for large incoherent device batches.  `order` (may be NULL) maps processing slot ->
 * storage index; when `unperm` is set the kernel writes slot-indexed records to `out` (local staging) and
 * finish_spatial_order() moves them to the caller's (remote) array. */
struct OrderPlan {
    gpurt_gather* gather = nullptr; /* results go to a gather's array on another GPU (gather.cu) */
    uint64_t n = 0;
    const uint32_t* order = nullptr;
    const uint32_t* unperm = nullptr;
    void* out = nullptr;
    bool scatter = false; /* results on another GPU: slices staged in processing order, scattered by a second stream */
};
int plan_spatial_order(gpurt_accel* A, const float4* pos, unsigned stride_vec4, uint64_t n, void* results,
                       size_t result_bytes, OrderPlan& P, bool sliced_scatter = false, bool any_bvh_size = false);
int finish_spatial_order(gpurt_accel* A, const OrderPlan& P, uint64_t n, void* results, size_t result_bytes);
/* remote results of an ordered batch: the slice [off, off + m) of the staging array (processing order) goes to its storage
 * positions in `results` on the placement stream, after everything queued on the context's stream so far */
void order_slices(uint64_t n, std::vector<uint64_t>& ends);
size_t order_arena_bytes(uint64_t n, size_t result_bytes, bool staged);
int scatter_slice_async(gpurt_accel* A, const OrderPlan& P, uint64_t off, uint64_t m, void* results, size_t result_bytes);
int scatter_join(gpurt_accel* A, const OrderPlan& P); /* the context's stream waits for the placement stream */
void preload_sort_kernels();
void preload_order_kernels();
void preload_cpq_kernels();
void preload_trace_kernels();
/* gather.cu */
gpurt_gather* gather_find(gpurt_ctx* ctx, const void* results, uint64_t n, size_t record_bytes);
int gather_begin_batch(gpurt_gather* g);
int gather_push_slice(gpurt_gather* g, const void* staged, const uint32_t* order, uint64_t off, uint64_t m);
int gather_join(gpurt_gather* g);
int gather_signal_direct(gpurt_gather* g);

/* query launchers (device pointers, async on ctx->stream) */
int launch_trace_closest(gpurt_accel* A, const float4* rays, uint64_t n, float4* hits);
int launch_trace_any(gpurt_accel* A, const float4* rays, uint64_t n, uint8_t* occ);
int launch_trace_closest_bvh2(gpurt_accel* A, const float4* rays, uint64_t n, float4* hits);
int launch_trace_closest_stats(gpurt_accel* A, const float4* rays, uint64_t n, float4* hits,
                               unsigned long long* d_counters);
int launch_closest_points(gpurt_accel* A, const float4* queries, uint64_t n, float4* results);
int launch_closest_points_stats(gpurt_accel* A, const float4* queries, uint64_t n, unsigned long long* d_counters);
} // namespace gpurt
