/*
 * gather.cu — gathering the result arrays of sharded query batches on one GPU without a collective (the
 * reference is single-GPU and has no counterpart).
 *
 * Every rank answers a contiguous range of one big result array that lives on the owner rank.  A large incoherent batch is
 * traversed in Morton order of its elements (order.cu), so its results come out in PROCESSING order and belong at scattered
 * STORAGE positions.  Three ways to get them to the owner were measured on config 4 (10 M triangles, 100 M points, 8 GPUs):
 *   - the traversal's own stores, or a scatter kernel per slice, over NVLink: 87.5 M scattered 32-byte stores from 7 senders
 *     arrive at ~175 GB/s — 6039 Mq/s;
 *   - chunks by storage position, each sorted on its own, chunk k copied while chunk k + 1 is traversed (bench r02): small
 *     chunks are 17 % less coherent and the last chunk's copy is exposed — 8105 Mq/s, 8786 with a tapered chunk schedule;
 *   - this file: the sender sorts its whole range ONCE, traverses it in slices of the processing order, and after each
 *     slice its copy engine moves the slice's results and storage indices — coalesced — into an inbox on the owner, then a
 *     flag; kernels the owner queued on a side stream wait for the flag and do the scatter locally, at HBM speed, while the
 *     owner's own batch is still being traversed.  Only the last slice's copy + scatter is exposed.
 */
#include <algorithm>
#include <cstring>

#include "device.cuh"

using namespace gpurt;

struct gpurt_gather {
    gpurt_ctx* ctx = nullptr;
    bool owner = false, ipc = false;
    uint32_t n_ranks = 0, owner_rank = 0, my_rank = 0, record_bytes = 0;
    uint64_t n_records = 0;
    std::vector<uint64_t> first;          /* [n_ranks + 1] */
    char* base = nullptr;                 /* owner: allocation; others: mapping */
    uint64_t off_results = 0, off_staging = 0, off_order = 0, bytes = 0;
    uint32_t seq = 0;                     /* rounds begun (owner) / batches pushed (sender) */
    uint32_t slices_pushed = 0;           /* sender: slices of the current batch handed over so far */
    cudaStream_t s_side = nullptr;        /* owner: waits + scatters */
    cudaEvent_t ev = nullptr;
};

namespace gpurt {

constexpr uint64_t kGatherHeader = 65536; /* progress word of rank r at byte 128 r; direct word at 8192 + 128 r; skip words
                                             at 16384 + 128 r; time-out counter at 32768; rounds fully scattered at 40960 */
constexpr uint32_t kGatherMaxRanks = 64;

static uint64_t align256(uint64_t x) { return (x + 255) & ~(uint64_t)255; }

static void gather_layout(gpurt_gather* g) {
    const uint64_t foreign = g->n_records - (g->first[g->owner_rank + 1] - g->first[g->owner_rank]);
    g->off_results = kGatherHeader;
    g->off_staging = g->off_results + align256(g->n_records * g->record_bytes);
    g->off_order = g->off_staging + align256(foreign * g->record_bytes);
    g->bytes = g->off_order + align256(foreign * 4);
}
/* position of rank r's first record among the records that do NOT belong to the owner (= inside the inbox) */
static uint64_t foreign_first(const gpurt_gather* g, uint32_t r) {
    uint64_t f = g->first[r];
    if(r > g->owner_rank) f -= g->first[g->owner_rank + 1] - g->first[g->owner_rank];
    return f;
}

__global__ void k_gather_signal(char* base, uint32_t rank, unsigned long long progress, uint32_t direct_seq) {
    if(threadIdx.x != 0) return;
    __threadfence_system();
    if(direct_seq) *(volatile uint32_t*)(base + 8192 + 128ull * rank) = direct_seq;
    else *(volatile unsigned long long*)(base + 128ull * rank) = progress;
}
/* owner, after the last scatter of a round: the inbox may be overwritten */
__global__ void k_gather_ack(char* base, uint32_t seq) {
    if(threadIdx.x != 0) return;
    __threadfence_system();
    *(volatile uint32_t*)(base + 40960) = seq;
}
/* sender, before the first copy of round `seq`: the owner has scattered round seq - 1 out of the inbox */
__global__ void k_gather_wait_ack(char* base, uint32_t seq) {
    if(threadIdx.x != 0 || seq <= 1u) return;
    const volatile uint32_t* ack = (const volatile uint32_t*)(base + 40960);
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while((int32_t)(*ack - (seq - 1u)) < 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if(t - t0 > 20000000000ull) break;
        __nanosleep(500);
    }
}
/* owner: wait until rank `r` has delivered slice `k` of round `seq` (or answered the round with direct stores) */
__global__ void k_gather_wait(char* base, uint32_t r, uint32_t seq, uint32_t k) {
    if(threadIdx.x != 0) return;
    const volatile unsigned long long* prog = (const volatile unsigned long long*)(base + 128ull * r);
    const volatile uint32_t* direct = (const volatile uint32_t*)(base + 8192 + 128ull * r);
    uint32_t* skip = (uint32_t*)(base + 16384 + 128ull * r);
    const unsigned long long want = ((unsigned long long)seq << 16) | (k + 1u);
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for(;;) {
        if(*direct == seq) {
            *skip = seq;
            break;
        }
        if(*prog >= want) break;
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if(t - t0 > 20000000000ull) {
            atomicAdd((uint32_t*)(base + 32768), 1u);
            *skip = seq; /* nothing trustworthy to scatter */
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}
/* results[first + order[i]] = staged[i], unless the sender answered this round with direct stores */
template <int VEC4>
__global__ void __launch_bounds__(256) k_gather_scatter(const char* base, uint32_t r, uint32_t seq, const float4* __restrict__ staged,
                                                        const uint32_t* __restrict__ order, uint64_t m, float4* __restrict__ results) {
    if(*(const uint32_t*)(base + 16384 + 128ull * r) == seq) return;
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(t >= (uint64_t)VEC4 * m) return;
    results[(uint64_t)VEC4 * __ldcg(order + t / VEC4) + (t % VEC4)] = __ldcg(staged + t);
}

static void gather_preload() {
    cudaFuncAttributes a;
    for(const void* f : {(const void*)k_gather_signal, (const void*)k_gather_ack, (const void*)k_gather_wait_ack, (const void*)k_gather_wait,
                         (const void*)k_gather_scatter<1>, (const void*)k_gather_scatter<2>})
        (void)cudaFuncGetAttributes(&a, f);
    preload_order_kernels(), preload_cpq_kernels(), preload_trace_kernels();
}

gpurt_gather* gather_find(gpurt_ctx* ctx, const void* results, uint64_t n, size_t record_bytes) {
    for(gpurt_gather* g : ctx->gathers) {
        if(g->owner || g->record_bytes != record_bytes) continue;
        const char* want = g->base + g->off_results + g->first[g->my_rank] * g->record_bytes;
        if((const char*)results == want && n == g->first[g->my_rank + 1] - g->first[g->my_rank]) return g;
    }
    return nullptr;
}
int gather_begin_batch(gpurt_gather* g) {
    g->seq++, g->slices_pushed = 0;
    if(!g->s_side) {
        GPURT_CUDA(cudaStreamCreateWithFlags(&g->s_side, cudaStreamNonBlocking));
        GPURT_CUDA(cudaEventCreateWithFlags(&g->ev, cudaEventDisableTiming));
    }
    return GPURT_OK;
}
/* sender, after slice [off, off + m) of the processing order was traversed on the context's stream: results + storage
 * indices to the owner's inbox on the side stream, then the progress flag */
int gather_push_slice(gpurt_gather* g, const void* staged, const uint32_t* order, uint64_t off, uint64_t m) {
    gpurt_ctx* ctx = g->ctx;
    const uint32_t slice = g->slices_pushed++;
    GPURT_CUDA(cudaEventRecord(g->ev, ctx->stream));
    GPURT_CUDA(cudaStreamWaitEvent(g->s_side, g->ev, 0));
    const uint64_t f = foreign_first(g, g->my_rank) + off;
    if(slice == 0) k_gather_wait_ack<<<1, 32, 0, g->s_side>>>(g->base, g->seq);
    GPURT_CUDA(cudaMemcpyAsync(g->base + g->off_staging + f * g->record_bytes, (const char*)staged + off * g->record_bytes,
                               m * g->record_bytes, cudaMemcpyDeviceToDevice, g->s_side));
    GPURT_CUDA(cudaMemcpyAsync(g->base + g->off_order + f * 4, order + off, m * 4, cudaMemcpyDeviceToDevice, g->s_side));
    k_gather_signal<<<1, 32, 0, g->s_side>>>(g->base, g->my_rank, ((unsigned long long)g->seq << 16) | (slice + 1u), 0u);
    GPURT_CUDA(cudaGetLastError());
    return GPURT_OK;
}
int gather_join(gpurt_gather* g) { /* the staging area of the sender is free again once its copies are done */
    GPURT_CUDA(cudaEventRecord(g->ev, g->s_side));
    GPURT_CUDA(cudaStreamWaitEvent(g->ctx->stream, g->ev, 0));
    return GPURT_OK;
}
/* sender: the batch was answered with direct stores into the owner's array (small or already coherent batch) */
int gather_signal_direct(gpurt_gather* g) {
    k_gather_signal<<<1, 32, 0, g->ctx->stream>>>(g->base, g->my_rank, 0ull, g->seq);
    GPURT_CUDA(cudaGetLastError());
    return GPURT_OK;
}

} // namespace gpurt

extern "C" {

static int gather_check(uint64_t n_records, uint32_t record_bytes, uint32_t n_ranks, const uint64_t* first, uint32_t owner_rank) {
    if(!n_records || (record_bytes != 16 && record_bytes != 32) || !n_ranks || n_ranks > kGatherMaxRanks || !first || owner_rank >= n_ranks)
        return set_error("gather: bad arguments (records of 16 or 32 bytes, at most 64 ranks)"), GPURT_E_INVALID;
    if(first[0] != 0 || first[n_ranks] != n_records) return set_error("gather: first_record must run from 0 to n_records"), GPURT_E_INVALID;
    for(uint32_t r = 0; r < n_ranks; r++)
        if(first[r] > first[r + 1] || first[r + 1] - first[r] >= (1ull << 30))
            return set_error("gather: ranges must ascend and hold fewer than 2^30 records each"), GPURT_E_INVALID;
    return GPURT_OK;
}

int gpurt_gather_create(gpurt_ctx* ctx, uint64_t n_records, uint32_t record_bytes, uint32_t n_ranks, const uint64_t* first,
                        uint32_t owner_rank, gpurt_gather** out, uint8_t* handle, uint64_t* out_bytes) {
    if(!ctx || !out) return set_error("NULL argument"), GPURT_E_INVALID;
    int rc = gather_check(n_records, record_bytes, n_ranks, first, owner_rank);
    if(rc) return rc;
    GPURT_CUDA(cudaSetDevice(ctx->device));
    gather_preload();
    gpurt_gather* g = new gpurt_gather;
    g->ctx = ctx, g->owner = true, g->n_ranks = n_ranks, g->owner_rank = g->my_rank = owner_rank, g->record_bytes = record_bytes;
    g->n_records = n_records, g->first.assign(first, first + n_ranks + 1);
    gather_layout(g);
    cudaError_t e = cudaMalloc((void**)&g->base, g->bytes);
    if(e == cudaSuccess) e = cudaMemset(g->base, 0, kGatherHeader);
    if(e == cudaSuccess) e = cudaStreamCreateWithPriority(&g->s_side, cudaStreamNonBlocking, -1);
    if(e == cudaSuccess) e = cudaEventCreateWithFlags(&g->ev, cudaEventDisableTiming);
    if(e == cudaSuccess && handle) {
        cudaIpcMemHandle_t h;
        e = cudaIpcGetMemHandle(&h, g->base);
        if(e == cudaSuccess) memcpy(handle, &h, sizeof h);
    }
    if(e != cudaSuccess) {
        set_error(std::string("gpurt_gather_create: ") + cudaGetErrorString(e));
        if(g->base) cudaFree(g->base);
        if(g->s_side) cudaStreamDestroy(g->s_side);
        if(g->ev) cudaEventDestroy(g->ev);
        delete g;
        return GPURT_E_CUDA;
    }
    /* allocate now what the owner's own batch will need: an allocation inside a round would wait for the receivers queued
     * by gpurt_gather_begin (cudaMalloc synchronises the device), i.e. for every sender */
    if((rc = ctx->build_arena.reserve(order_arena_bytes(first[owner_rank + 1] - first[owner_rank], record_bytes, false)))) {
        gpurt_gather_destroy(g);
        return rc;
    }
    if(out_bytes) *out_bytes = g->bytes;
    ctx->gathers.push_back(g);
    *out = g;
    return GPURT_OK;
}
int gpurt_gather_open(gpurt_ctx* ctx, const uint8_t* handle, void* same_process_base, uint64_t n_records, uint32_t record_bytes,
                      uint32_t n_ranks, const uint64_t* first, uint32_t owner_rank, uint32_t my_rank, gpurt_gather** out) {
    if(!ctx || !out || (!handle && !same_process_base)) return set_error("NULL argument"), GPURT_E_INVALID;
    int rc = gather_check(n_records, record_bytes, n_ranks, first, owner_rank);
    if(rc) return rc;
    if(my_rank >= n_ranks || my_rank == owner_rank) return set_error("gather_open: my_rank must be one of the other ranks"), GPURT_E_INVALID;
    GPURT_CUDA(cudaSetDevice(ctx->device));
    gather_preload();
    gpurt_gather* g = new gpurt_gather;
    g->ctx = ctx, g->owner = false, g->n_ranks = n_ranks, g->owner_rank = owner_rank, g->my_rank = my_rank, g->record_bytes = record_bytes;
    g->n_records = n_records, g->first.assign(first, first + n_ranks + 1);
    gather_layout(g);
    if(same_process_base) g->base = (char*)same_process_base;
    else {
        cudaIpcMemHandle_t h;
        memcpy(&h, handle, sizeof h);
        cudaError_t e = cudaIpcOpenMemHandle((void**)&g->base, h, cudaIpcMemLazyEnablePeerAccess);
        if(e != cudaSuccess) {
            delete g;
            return set_error(std::string("gpurt_gather_open: ") + cudaGetErrorString(e)), GPURT_E_CUDA;
        }
    }
    g->ipc = !same_process_base; /* same-process bases are not unmapped on destroy */
    cudaError_t e = cudaStreamCreateWithFlags(&g->s_side, cudaStreamNonBlocking);
    if(e == cudaSuccess) e = cudaEventCreateWithFlags(&g->ev, cudaEventDisableTiming);
    if(e != cudaSuccess) {
        set_error(std::string("gpurt_gather_open: ") + cudaGetErrorString(e));
        gpurt_gather_destroy(g);
        return GPURT_E_CUDA;
    }
    /* sort scratch + staging of this rank's batch, allocated outside the rounds (see gpurt_gather_create) */
    if((rc = ctx->build_arena.reserve(order_arena_bytes(first[my_rank + 1] - first[my_rank], record_bytes, true)))) {
        gpurt_gather_destroy(g);
        return rc;
    }
    ctx->gathers.push_back(g);
    *out = g;
    return GPURT_OK;
}
int gpurt_gather_results(gpurt_gather* g, void** out_base, void** out_mine) {
    if(!g) return set_error("NULL argument"), GPURT_E_INVALID;
    if(out_base) *out_base = g->base + g->off_results;
    if(out_mine) *out_mine = g->base + g->off_results + g->first[g->my_rank] * g->record_bytes;
    return GPURT_OK;
}
int gpurt_gather_base(gpurt_gather* g, void** out_base) {
    if(!g || !out_base) return set_error("NULL argument"), GPURT_E_INVALID;
    *out_base = g->base;
    return GPURT_OK;
}
int gpurt_gather_begin(gpurt_gather* g) {
    if(!g || !g->owner) return set_error("gather_begin: owner only"), GPURT_E_INVALID;
    gpurt_ctx* ctx = g->ctx;
    GPURT_CUDA(cudaSetDevice(ctx->device));
    g->seq++;
    /* the receivers start after whatever the context's stream holds so far (e.g. a consumer of the previous round) */
    GPURT_CUDA(cudaEventRecord(g->ev, ctx->stream));
    GPURT_CUDA(cudaStreamWaitEvent(g->s_side, g->ev, 0));
    const uint32_t n_ranks = g->n_ranks;
    size_t max_slices = 0;
    std::vector<std::vector<uint64_t>> ends(n_ranks);
    for(uint32_t r = 0; r < n_ranks; r++) {
        const uint64_t n = g->first[r + 1] - g->first[r];
        if(r == g->owner_rank || !n) continue;
        order_slices(n, ends[r]);
        max_slices = std::max(max_slices, ends[r].size());
    }
    for(size_t k = 0; k < max_slices; k++)
        for(uint32_t r = 0; r < n_ranks; r++) {
            if(k >= ends[r].size()) continue;
            const uint64_t off = k ? ends[r][k - 1] : 0, m = ends[r][k] - off, f = foreign_first(g, r) + off;
            k_gather_wait<<<1, 32, 0, g->s_side>>>(g->base, r, g->seq, (uint32_t)k);
            const float4* staged = (const float4*)(g->base + g->off_staging + f * g->record_bytes);
            const uint32_t* order = (const uint32_t*)(g->base + g->off_order + f * 4);
            float4* results = (float4*)(g->base + g->off_results + g->first[r] * g->record_bytes);
            if(g->record_bytes == 32)
                k_gather_scatter<2><<<(unsigned)((2 * m + 255) / 256), 256, 0, g->s_side>>>(g->base, r, g->seq, staged, order, m, results);
            else
                k_gather_scatter<1><<<(unsigned)((m + 255) / 256), 256, 0, g->s_side>>>(g->base, r, g->seq, staged, order, m, results);
        }
    k_gather_ack<<<1, 32, 0, g->s_side>>>(g->base, g->seq);
    GPURT_CUDA(cudaGetLastError());
    return GPURT_OK;
}
int gpurt_gather_end(gpurt_gather* g, uint32_t* out_timeouts) {
    if(!g || !g->owner) return set_error("gather_end: owner only"), GPURT_E_INVALID;
    GPURT_CUDA(cudaSetDevice(g->ctx->device));
    GPURT_CUDA(cudaEventRecord(g->ev, g->s_side));
    GPURT_CUDA(cudaStreamWaitEvent(g->ctx->stream, g->ev, 0));
    if(out_timeouts) { /* synchronises: for tests and diagnostics */
        GPURT_CUDA(cudaStreamSynchronize(g->ctx->stream));
        GPURT_CUDA(cudaMemcpy(out_timeouts, g->base + 32768, 4, cudaMemcpyDeviceToHost));
    }
    return GPURT_OK;
}
int gpurt_gather_destroy(gpurt_gather* g) {
    if(!g) return GPURT_OK;
    gpurt_ctx* ctx = g->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if(g->s_side) cudaStreamSynchronize(g->s_side);
    ctx->gathers.erase(std::remove(ctx->gathers.begin(), ctx->gathers.end(), g), ctx->gathers.end());
    if(g->owner) cudaFree(g->base);
    else if(g->ipc) cudaIpcCloseMemHandle(g->base);
    if(g->s_side) cudaStreamDestroy(g->s_side);
    if(g->ev) cudaEventDestroy(g->ev);
    delete g;
    return GPURT_OK;
}

} /* extern "C" */
