/*
 * order.cu — spatial processing order for large, incoherent device batches of rays / query points.
 *
 * One thread per ray or point walks the BVH.  When the BVH is much larger than L2 and the batch arrives in no
 * spatial order, the 32 lanes of a warp touch unrelated nodes: every fetch misses L1, and for closest-point
 * descents the lanes also shrink their radius at different paces (9 of 32 lanes busy).  Measured on the
 * 10 M-triangle config-4 soup (tools/cpq_sort_probe.py, tools/ray_sort_probe.py): uniform random points
 * 807 -> 1355 Mq/s and uniform random rays 676 -> 970 Mrays/s when processed in Morton order of the position.
 * So such batches get a 30-bit Morton key per element, a 3- or 4-pass radix sort of (key, index) pairs, and the
 * traversal kernel reads its element and writes its result through the sorted index.  Skipped when the batch is small
 * (< 2^20), when a probe on 1 / 16 of the batch finds that neighbouring elements already share a cell of a 16^3 grid
 * (32^3 on small scenes; e.g. rays and points generated per pixel), and — for rays — when the scene is small (BVH < 64 MB:
 * ray order hardly matters once the tree sits in L2; stand-in bounce rays gain 10-18 % BEFORE paying for the sort).
 * Closest-point batches are re-ordered on every scene larger than the L1s (4 MB) since the end of round 2: the descent
 * diverges with the spread of a warp's points even when the tree sits in L2 (stand-in, 2 M points near the visible
 * surfaces, kernel + sort:
 * jittered by +-30 units 1431 -> 1453 Mq/s: the probe orders them and the sort takes back what the kernel gains (1.41 ->
 * 1.19 ms); +-200 units 738 -> 1090 Mq/s; points in pixel order without jitter lose 8 % when sorted -> left alone by the
 * probe; profiles/r04b_cpq_order.log, r04c_cpq_order.log).
 * Results never depend on the processing order (contracts N4 / N5).
 *
 * If the result array lives on another GPU (gpurt_shared_open mapping), scattered 16/32-byte stores over NVLink issued by
 * the traversal kernel itself would cost ~70 % more than the traversal, so results are staged locally in processing order.
 * Closest-hit and closest-point batches are then traversed in slices of the processing order: while slice k + 1 is
 * traversed, a small kernel on a second stream stores slice k to its storage positions on the other GPU, so only the last
 * slice's transfer is exposed and the batch keeps the coherence of ONE sort over all its elements (chunking the batch by
 * storage position instead, to overlap copies, sorts each chunk separately: 3.1 M-point chunks of config 4 run 17 % slower
 * than one 12.5 M-point batch).  Any-hit results (1 byte) are written front to back by a second kernel.
 */
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "device.cuh"
#include "traverse.cuh"

namespace gpurt {

GPURT_HD unsigned order_key30(float4 q, float lx, float ly, float lz, float ix, float iy, float iz) {
    unsigned x = (unsigned)fminf(fmaxf((q.x - lx) * ix * 1024.0f, 0.0f), 1023.0f);
    unsigned y = (unsigned)fminf(fmaxf((q.y - ly) * iy * 1024.0f, 0.0f), 1023.0f);
    unsigned z = (unsigned)fminf(fmaxf((q.z - lz) * iz * 1024.0f, 0.0f), 1023.0f);
    return (unsigned)((expand21(x) << 2) | (expand21(y) << 1) | expand21(z));
}
__global__ void __launch_bounds__(256) k_order_keys(const float4* __restrict__ pos, unsigned stride, uint64_t n, float lx,
                                                    float ly, float lz, float ix, float iy, float iz,
                                                    unsigned key_shift, uint64_t* __restrict__ keys,
                                                    uint32_t* __restrict__ vals) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    keys[i] = order_key30(__ldg(pos + (size_t)stride * i), lx, ly, lz, ix, iy, iz) >> key_shift;
    vals[i] = (uint32_t)i;
}
/* Coherence probe on a sample of the batch (every block_stride-th run of 256 elements; nothing is written but two
 * counters): does the next element fall into the same cell of a 2^bits-per-axis grid (the top 3 * bits key bits)?
 * counters[0] += pairs that do, counters[1] += pairs looked at. */
__global__ void __launch_bounds__(256) k_order_probe(const float4* __restrict__ pos, unsigned stride, uint64_t n, float lx,
                                                     float ly, float lz, float ix, float iy, float iz, unsigned block_stride,
                                                     unsigned probe_shift, unsigned* __restrict__ counters) {
    uint64_t i = (uint64_t)blockIdx.x * block_stride * blockDim.x + threadIdx.x;
    unsigned key = 0xffffffffu;
    if(i < n) key = order_key30(__ldg(pos + (size_t)stride * i), lx, ly, lz, ix, iy, iz);
    unsigned next = __shfl_down_sync(0xffffffffu, key, 1);
    bool pair = (threadIdx.x & 31) != 31 && i + 1 < n;
    bool same = pair && (key >> probe_shift) == (next >> probe_shift);
    unsigned cs = __popc(__ballot_sync(0xffffffffu, same)), cp = __popc(__ballot_sync(0xffffffffu, pair));
    if((threadIdx.x & 31) == 0 && cp) {
        if(cs) atomicAdd(counters, cs);
        atomicAdd(counters + 1, cp);
    }
}

__global__ void __launch_bounds__(256) k_order_invert(const uint32_t* __restrict__ order, uint64_t n, uint32_t* __restrict__ inv) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) inv[order[i]] = (uint32_t)i;
}
/* results[j] = staged[inv[j]] for records of VEC4 float4s; one thread per float4 */
template <int VEC4>
__global__ void __launch_bounds__(256) k_order_unpermute(const float4* __restrict__ staged, const uint32_t* __restrict__ inv,
                                                         uint64_t n, float4* __restrict__ results) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(t >= (uint64_t)VEC4 * n) return;
    results[t] = staged[(uint64_t)VEC4 * inv[t / VEC4] + (t % VEC4)];
}
/* results[order[i]] = staged[i] for the records of one slice; one thread per float4, the VEC4 float4s of a record by
 * neighbouring lanes */
template <int VEC4>
__global__ void __launch_bounds__(256) k_order_scatter(const float4* __restrict__ staged, const uint32_t* __restrict__ order,
                                                       uint64_t m, float4* __restrict__ results) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(t >= (uint64_t)VEC4 * m) return;
    results[(uint64_t)VEC4 * order[t / VEC4] + (t % VEC4)] = staged[t];
}
__global__ void __launch_bounds__(256) k_order_unpermute_u8(const uint8_t* __restrict__ staged, const uint32_t* __restrict__ inv,
                                                            uint64_t n, uint8_t* __restrict__ results) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(t < n) results[t] = staged[inv[t]];
}

constexpr uint64_t kOrderMinBatch = 1u << 20;
/* 8-bit radix passes over the 30-bit Morton key: the top 24 bits (2^24 cells) order a batch of up to 2^24 elements as well as
 * all 30 do and save a pass — stand-in, 2 M points: 1428 -> 1485 Mq/s; config 4 in calls of 12.5 M: 1378 -> 1392; one call of
 * 100 M points (6 per cell) loses 5 % with 3 passes and 35 % with 2, so dense batches keep 4
 * (tools/order_passes_probe.py, profiles/r05c_order_passes.log) */
constexpr uint64_t kOrderCoarseBatch = 1u << 24;
constexpr size_t kOrderMinBvhBytes = 64u << 20;
/* closest-point batches: re-ordered whenever the tree does not fit the L1s.  Measured with 2^20 uniform random points: on
 * the stand-in (16 MB of nodes + triangles, L2-resident) 738 -> 1101 Mq/s including the sort; on cbox (1 MB: served by
 * L1 whatever the order) 677 -> 563 Mq/s — the kernel gains nothing there and the sort is a sixth of the call. */
constexpr size_t kOrderMinBvhBytesPoints = 4u << 20;

/* bytes plan_spatial_order takes from the context's build arena for a batch of n elements (with a staging array) */
size_t order_arena_bytes(uint64_t n, size_t result_bytes, bool staged) {
    const size_t kb = ((size_t)n * 8 + 255) & ~(size_t)255, vb = ((size_t)n * 4 + 255) & ~(size_t)255;
    return 2 * kb + 2 * vb + 256 + (staged ? (((size_t)n * result_bytes + 255) & ~(size_t)255) : 0);
}

int plan_spatial_order(gpurt_accel* A, const float4* pos, unsigned stride_vec4, uint64_t n, void* results,
                       size_t result_bytes, OrderPlan& P, bool sliced_scatter, bool any_bvh_size) {
    P = OrderPlan();
    P.out = results, P.n = n;
    gpurt_ctx* ctx = A->ctx;
    cudaStream_t st = ctx->stream;
    if(sliced_scatter && !ctx->gathers.empty() && (P.gather = gather_find(ctx, results, n, result_bytes))) {
        int grc = gather_begin_batch(P.gather);
        if(grc) return grc;
    }
    static const bool allow = !(getenv("GPURT_SPATIAL_ORDER") && atoi(getenv("GPURT_SPATIAL_ORDER")) == 0);
    const size_t bvh_bytes = (size_t)A->n_nodes * sizeof(Node8) + (size_t)A->n * 48;
    /* test hooks (sanitizer runs, small-scene tests of the ordered paths): GPURT_ORDER_MIN_BATCH / GPURT_ORDER_MIN_BVH_BYTES */
    const char *eb = getenv("GPURT_ORDER_MIN_BATCH"), *ev = getenv("GPURT_ORDER_MIN_BVH_BYTES");
    const uint64_t min_batch = eb ? (uint64_t)atoll(eb) : kOrderMinBatch;
    const size_t min_bvh = ev ? (size_t)atoll(ev) : (any_bvh_size ? kOrderMinBvhBytesPoints : kOrderMinBvhBytes);
    if(!allow || n < min_batch || n >= (1ull << 30) || bvh_bytes <= min_bvh || !A->n_nodes) return GPURT_OK;
    /* scratch in the build arena (no build runs concurrently on this stream): keys | keys_tmp | vals | vals_tmp | counter | staging */
    const size_t kb = ((size_t)n * 8 + 255) & ~(size_t)255, vb = ((size_t)n * 4 + 255) & ~(size_t)255;
    cudaPointerAttributes pa;
    /* GPURT_PLACE_FORCE=1 (test hook): treat local results like another GPU's, to run the placement path on one GPU */
    const bool force_remote = getenv("GPURT_PLACE_FORCE") && atoi(getenv("GPURT_PLACE_FORCE")) != 0;
    const bool remote = (cudaPointerGetAttributes(&pa, results) == cudaSuccess && pa.type == cudaMemoryTypeDevice &&
                         pa.device != ctx->device) || force_remote;
    (void)cudaGetLastError();
    const size_t used = 2 * kb + 2 * vb + 256, stage_bytes = (remote || P.gather) ? (((size_t)n * result_bytes + 255) & ~(size_t)255) : 0;
    int rc = ctx->build_arena.reserve(used + stage_bytes);
    if(rc) return rc;
    char* base = (char*)ctx->build_arena.p;
    uint64_t *keys = (uint64_t*)base, *keys_tmp = (uint64_t*)(base + kb);
    uint32_t *vals = (uint32_t*)(base + 2 * kb), *vals_tmp = (uint32_t*)(base + 2 * kb + vb);
    unsigned* counter = (unsigned*)(base + 2 * kb + 2 * vb);
    const float* sb = A->scene_box;
    float inv[3];
    for(int k = 0; k < 3; k++) inv[k] = sb[3 + k] > sb[k] ? 1.0f / (sb[3 + k] - sb[k]) : 0.0f;
    /* the probe reads 1 / 16 of the batch (every 16th run of 256 elements) and costs a few microseconds plus one 8-byte
     * read-back; only batches it finds incoherent pay for keys and sort */
    const char* epb = getenv("GPURT_ORDER_PROBE_BITS"); /* experiment hook: bits per axis of the probe grid */
    const unsigned probe_bits = epb ? (unsigned)std::min(10, std::max(1, atoi(epb))) : (bvh_bytes > kOrderMinBvhBytes ? 4u : 5u);
    const unsigned nb = (unsigned)((n + 255) / 256), probe_stride = 16;
    GPURT_CUDA(cudaMemsetAsync(counter, 0, 8, st));
    k_order_probe<<<(nb + probe_stride - 1) / probe_stride, 256, 0, st>>>(pos, stride_vec4, n, sb[0], sb[1], sb[2], inv[0], inv[1], inv[2],
                                                                          probe_stride, 30u - 3u * probe_bits, counter);
    unsigned same[2] = {0, 0};
    GPURT_CUDA(cudaMemcpyAsync(same, counter, 8, cudaMemcpyDeviceToHost, st));
    GPURT_CUDA(cudaStreamSynchronize(st));
    if((double)same[0] >= 0.5 * (double)same[1]) return GPURT_OK; /* already coherent */
    /* the sort looks at the top 8 * passes bits of the 30-bit key (GPURT_ORDER_PASSES, experiment hook: 2..4) */
    const char* eps = getenv("GPURT_ORDER_PASSES");
    const int passes = eps ? std::min(4, std::max(2, atoi(eps))) : (n <= kOrderCoarseBatch ? 3 : 4);
    k_order_keys<<<nb, 256, 0, st>>>(pos, stride_vec4, n, sb[0], sb[1], sb[2], inv[0], inv[1], inv[2], passes >= 4 ? 0u : 30u - 8u * passes,
                                     keys, vals);
    rc = radix_sort_u64(st, keys, vals, keys_tmp, vals_tmp, n, passes, ctx->scratch, ctx->sm_count);
    if(rc) return rc;
    if(passes & 1) { /* an odd number of passes leaves the result in the secondary buffers; the roles swap */
        std::swap(keys, keys_tmp);
        std::swap(vals, vals_tmp);
    }
    P.order = vals;
    /* opt-in (GPURT_PLACE_SLICES=<slices>): fine for one or two senders (2 GPUs, config 4: 2616 Mq/s against 2655 with the
     * results left local), but scattered 32-byte stores from 7 senders into one GPU arrive at only ~175 GB/s (8 GPUs: 6039
     * Mq/s against 8105 for caller-side chunks + coalesced copies), so the default stays "stage, then one coalesced pass" */
    const bool allow_slices = getenv("GPURT_PLACE_SLICES") && atoi(getenv("GPURT_PLACE_SLICES")) > 0;
    if(P.gather) { /* slices go to the owner's inbox by copy engine; the owner scatters them (gather.cu) */
        P.out = base + used;
        P.scatter = true;
    } else if(remote && sliced_scatter && allow_slices) {
        P.out = base + used;
        P.scatter = true;
        if(!ctx->s_place) {
            GPURT_CUDA(cudaStreamCreateWithFlags(&ctx->s_place, cudaStreamNonBlocking));
            GPURT_CUDA(cudaEventCreateWithFlags(&ctx->ev_place, cudaEventDisableTiming));
        }
    } else if(remote) {
        P.out = base + used;
        uint32_t* invw = (uint32_t*)keys_tmp; /* the sort is finished with its key scratch */
        k_order_invert<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P.order, n, invw);
        P.unperm = invw;
    }
    GPURT_CUDA(cudaGetLastError());
    return GPURT_OK;
}

/* Slices of the processing order, as ascending end positions.  The transfer of slice k overlaps the traversal of slice
 * k + 1, so only the last slice's transfer is exposed, and every slice boundary costs a drain of the traversal kernel: a few
 * large slices first, small ones at the end (GPURT_PLACE_SLICES=<k>: k equal slices instead). */
void order_slices(uint64_t n, std::vector<uint64_t>& ends) {
    ends.clear();
    const char* env = getenv("GPURT_PLACE_SLICES");
    const uint64_t min_slice = getenv("GPURT_ORDER_MIN_BATCH") ? 256u : 1u << 18; /* small slices only under the test hook */
    if(env && atoi(env) > 0) {
        const uint64_t k = (uint64_t)atoi(env);
        uint64_t s = std::max<uint64_t>((n + k - 1) / k, min_slice);
        s = (s + 127) & ~(uint64_t)127;
        for(uint64_t e = s; e < n; e += s) ends.push_back(e);
    } else {
        static const double cum[] = {0.30, 0.55, 0.74, 0.87, 0.95};
        uint64_t prev = 0;
        for(double c : cum) {
            uint64_t e = ((uint64_t)((double)n * c) + 127) & ~(uint64_t)127;
            if(e >= n || n - e < min_slice) break;
            if(e - prev < min_slice) continue;
            ends.push_back(e), prev = e;
        }
    }
    ends.push_back(n);
}
int scatter_slice_async(gpurt_accel* A, const OrderPlan& P, uint64_t off, uint64_t m, void* results, size_t result_bytes) {
    gpurt_ctx* ctx = A->ctx;
    if(P.gather) return gather_push_slice(P.gather, P.out, P.order, off, m);
    GPURT_CUDA(cudaEventRecord(ctx->ev_place, ctx->stream));
    GPURT_CUDA(cudaStreamWaitEvent(ctx->s_place, ctx->ev_place, 0));
    const float4* staged = (const float4*)((const char*)P.out + off * result_bytes);
    if(result_bytes == 32)
        k_order_scatter<2><<<(unsigned)((2 * m + 255) / 256), 256, 0, ctx->s_place>>>(staged, P.order + off, m, (float4*)results);
    else if(result_bytes == 16)
        k_order_scatter<1><<<(unsigned)((m + 255) / 256), 256, 0, ctx->s_place>>>(staged, P.order + off, m, (float4*)results);
    else return set_error("scatter_slice_async: record size"), GPURT_E_STATE;
    GPURT_CUDA(cudaGetLastError());
    return GPURT_OK;
}
int scatter_join(gpurt_accel* A, const OrderPlan& P) {
    gpurt_ctx* ctx = A->ctx;
    if(P.gather) return gather_join(P.gather);
    GPURT_CUDA(cudaEventRecord(ctx->ev_place, ctx->s_place));
    GPURT_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_place, 0));
    return GPURT_OK;
}

int finish_spatial_order(gpurt_accel* A, const OrderPlan& P, uint64_t n, void* results, size_t result_bytes) {
    if(P.gather && !P.scatter) return gather_signal_direct(P.gather); /* answered with the kernel's own stores: tell the owner */
    if(!P.unperm) return GPURT_OK;
    cudaStream_t st = A->ctx->stream;
    if(result_bytes == 32)
        k_order_unpermute<2><<<(unsigned)((2 * n + 255) / 256), 256, 0, st>>>((const float4*)P.out, P.unperm, n, (float4*)results);
    else if(result_bytes == 16)
        k_order_unpermute<1><<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const float4*)P.out, P.unperm, n, (float4*)results);
    else if(result_bytes == 1)
        k_order_unpermute_u8<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const uint8_t*)P.out, P.unperm, n, (uint8_t*)results);
    else return set_error("finish_spatial_order: record size"), GPURT_E_STATE;
    GPURT_CUDA(cudaGetLastError());
    return GPURT_OK;
}

/* CUDA loads a kernel at its first launch, and that can wait for the device to go idle — forever, if a kernel of
 * gather.cu is spinning on a flag only this launch would raise.  gpurt_gather_create / _open load everything a round uses. */
static void preload(const void* f) {
    cudaFuncAttributes a;
    (void)cudaFuncGetAttributes(&a, f);
}
void preload_order_kernels() {
    preload((const void*)k_order_probe), preload((const void*)k_order_keys), preload((const void*)k_order_invert), preload((const void*)k_order_unpermute<1>);
    preload((const void*)k_order_unpermute<2>), preload((const void*)k_order_scatter<1>), preload((const void*)k_order_scatter<2>);
    preload((const void*)k_order_unpermute_u8);
    preload_sort_kernels();
}

} // namespace gpurt
