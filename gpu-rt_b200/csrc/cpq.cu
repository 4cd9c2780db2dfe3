/*
 * cpq.cu — closest-point query over the same 8-wide BVH (the reference's FCPW-GPU work,
 * README.md:6-8; no source in the snapshot, FCPW semantics: nearest point on any triangle within a
 * search radius, its distance and primitive).
 *
 * Priority-ordered descent (traverse.cuh: closest_point8).  N5 tie rule: smallest d^2 wins, equal
 * d^2 -> lowest global primitive id; the radius is inclusive.
 */
#include <cstdlib>

#include "device.cuh"
#include "traverse.cuh"

namespace gpurt {

/* ---- query re-ordering ------------------------------------------------------------------------ */
/* Unlike rays, closest-point descents of neighbouring points visit the same nodes and shrink their radius at the
 * same pace; a warp of 32 unrelated points runs at ~9 of 32 lanes.  Large device batches whose points arrive in
 * no spatial order (config 4: 100 M uniform random points) are therefore processed in Morton order of the query
 * position — a key per query, one 4-pass radix sort of (key, index) pairs, and the descent kernel reads its
 * query and writes its result through the sorted index — which is 1.5x faster end to end on that workload.
 * Batches that are already coherent (measured: neighbours fall into the same 16^3 cell) skip the sort, and so do
 * scenes whose BVH fits in L2 with room to spare (< 64 MB), where order hardly matters.
 * Results do not depend on the processing order. */
__global__ void __launch_bounds__(256) k_cpq_keys(const float4* __restrict__ queries, uint64_t n, float lx, float ly,
                                                  float lz, float ix, float iy, float iz, uint64_t* __restrict__ keys,
                                                  uint32_t* __restrict__ vals, unsigned* __restrict__ same_cell) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned key = 0xffffffffu;
    if(i < n) {
        float4 q = __ldg(queries + i);
        unsigned x = (unsigned)fminf(fmaxf((q.x - lx) * ix * 1024.0f, 0.0f), 1023.0f);
        unsigned y = (unsigned)fminf(fmaxf((q.y - ly) * iy * 1024.0f, 0.0f), 1023.0f);
        unsigned z = (unsigned)fminf(fmaxf((q.z - lz) * iz * 1024.0f, 0.0f), 1023.0f);
        key = (unsigned)((expand21(x) << 2) | (expand21(y) << 1) | expand21(z));
        keys[i] = key;
        vals[i] = (uint32_t)i;
    }
    /* coherence probe: does the next query fall into the same cell of a 16^3 grid (top 12 key bits)? */
    unsigned next = __shfl_down_sync(0xffffffffu, key, 1);
    bool same = (threadIdx.x & 31) != 31 && i + 1 < n && (key >> 18) == (next >> 18);
    unsigned cnt = __popc(__ballot_sync(0xffffffffu, same));
    if((threadIdx.x & 31) == 0 && cnt) atomicAdd(same_cell, cnt);
}

template <int STACK>
__global__ void __launch_bounds__(128) k_closest_points(const float4* __restrict__ nodes,
                                                        const float4* __restrict__ tris,
                                                        const float4* __restrict__ queries, uint64_t n,
                                                        float4* __restrict__ results, unsigned n_nodes,
                                                        const uint32_t* __restrict__ order, int staged) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const uint64_t slot = i;   /* processing position */
    if(order) i = order[i];    /* storage position (processing order != storage order) */
    float4 q = __ldg(queries + i);
    CpRec best;
    best.gid = kNoHit;
    if(n_nodes) closest_point8<STACK>(nodes, tris, f3(q.x, q.y, q.z), q.w, best);
    float4 o0, o1;
    if(best.gid == kNoHit) {
        o0 = make_float4(0.0f, 0.0f, 0.0f, GPURT_INF);
        o1 = make_float4(u2f(kNoHit), u2f(0u), 0.0f, 0.0f);
    } else {
        const float4* tp = tris + (size_t)best.idx * kTriVec4;
        float4 r0 = __ldg(tp), r1 = __ldg(tp + 1), r2 = __ldg(tp + 2);
        F3 c = tri_point(f3(r0.x, r0.y, r0.z), f3(r1.x, r1.y, r1.z), f3(r2.x, r2.y, r2.z), best.v, best.w);
        o0 = make_float4(c.x, c.y, c.z, sqrtf(best.d2));
        o1 = make_float4(u2f(best.gid), r1.w, best.v, best.w);
    }
    if(staged) i = slot; /* results go to a local staging array in processing order, k_cpq_unpermute moves them */
    results[2 * i] = o0;
    results[2 * i + 1] = o1;
}

/* Sorted batches whose result array lives on another GPU (gpurt_shared_open mapping): 32-byte stores scattered over
 * NVLink are slow (config 4 on 8 GPUs: 16.8 ms vs 10.0 ms with local results), so the descent writes a local
 * staging array in processing order and this kernel writes the caller's array front to back, fully coalesced. */
__global__ void __launch_bounds__(256) k_cpq_invert(const uint32_t* __restrict__ order, uint64_t n, uint32_t* __restrict__ inv) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) inv[order[i]] = (uint32_t)i;
}
__global__ void __launch_bounds__(256) k_cpq_unpermute(const float4* __restrict__ staged, const uint32_t* __restrict__ inv,
                                                       uint64_t n, float4* __restrict__ results) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; /* one thread per float4 */
    if(t >= 2 * n) return;
    results[t] = staged[2ull * inv[t >> 1] + (t & 1)];
}

constexpr uint64_t kCpqSortMin = 1u << 20; /* smaller batches are not worth a sort */

int launch_closest_points(gpurt_accel* A, const float4* queries, uint64_t n, float4* results) {
    if(!n) return GPURT_OK;
    unsigned nb = (unsigned)((n + 127) / 128);
    const float4* nodes = (const float4*)A->nodes;
    gpurt_ctx* ctx = A->ctx;
    cudaStream_t st = ctx->stream;
    const uint32_t* order = nullptr;
    float4* out = results;
    const uint32_t* unperm = nullptr; /* storage position -> processing position, when results are staged */
    static const bool allow_sort = !(getenv("GPURT_CPQ_SORT") && atoi(getenv("GPURT_CPQ_SORT")) == 0);
    /* only when the BVH does not sit in L2 anyway: on the 16 k-triangle Cornell box the sort costs 8 % and gains nothing */
    const size_t bvh_bytes = (size_t)A->n_nodes * sizeof(Node8) + (size_t)A->n * 48;
    if(allow_sort && n >= kCpqSortMin && n < (1ull << 30) && bvh_bytes > (64u << 20)) {
        /* scratch: keys | vals | keys_tmp | vals_tmp | counter, in the build arena (no build runs concurrently on this stream) */
        size_t kb = ((size_t)n * 8 + 255) & ~(size_t)255, vb = ((size_t)n * 4 + 255) & ~(size_t)255;
        /* is the result array on another GPU (a gpurt_shared_open mapping)?  then the results are staged, see below */
        cudaPointerAttributes pa;
        const bool remote = cudaPointerGetAttributes(&pa, results) == cudaSuccess && pa.type == cudaMemoryTypeDevice &&
                            pa.device != ctx->device;
        (void)cudaGetLastError();
        const size_t used = 2 * kb + 2 * vb + 256, stage_bytes = remote ? (((size_t)n * 32 + 255) & ~(size_t)255) : 0;
        int rc = ctx->build_arena.reserve(used + stage_bytes);
        if(rc) return rc;
        char* base = (char*)ctx->build_arena.p;
        uint64_t *keys = (uint64_t*)base, *keys_tmp = (uint64_t*)(base + kb);
        uint32_t *vals = (uint32_t*)(base + 2 * kb), *vals_tmp = (uint32_t*)(base + 2 * kb + vb);
        unsigned* counter = (unsigned*)(base + 2 * kb + 2 * vb);
        const float* sb = A->scene_box;
        float inv[3];
        for(int k = 0; k < 3; k++) inv[k] = sb[3 + k] > sb[k] ? 1.0f / (sb[3 + k] - sb[k]) : 0.0f;
        GPURT_CUDA(cudaMemsetAsync(counter, 0, 4, st));
        k_cpq_keys<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(queries, n, sb[0], sb[1], sb[2], inv[0], inv[1], inv[2], keys,
                                                                 vals, counter);
        unsigned same = 0;
        GPURT_CUDA(cudaMemcpyAsync(&same, counter, 4, cudaMemcpyDeviceToHost, st));
        GPURT_CUDA(cudaStreamSynchronize(st));
        if((double)same < 0.5 * (double)n) { /* incoherent batch */
            rc = radix_sort_u64(st, keys, vals, keys_tmp, vals_tmp, n, 4, ctx->scratch, ctx->sm_count);
            if(rc) return rc;
            order = vals; /* 4 passes: the result is back in the primary buffers */
            if(remote) { /* stage locally, write the remote array coalesced afterwards */
                out = (float4*)(base + used);
                uint32_t* invw = (uint32_t*)keys_tmp; /* the sort is finished with its key scratch */
                k_cpq_invert<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(order, n, invw);
                unperm = invw;
            }
        }
    }
    unsigned need = 7u * A->depth + 1u;
    if(need <= 64) k_closest_points<64><<<nb, 128, 0, st>>>(nodes, A->tri_wide, queries, n, out, A->n_nodes, order, unperm ? 1 : 0);
    else if(need <= 128) k_closest_points<128><<<nb, 128, 0, st>>>(nodes, A->tri_wide, queries, n, out, A->n_nodes, order, unperm ? 1 : 0);
    else if(need <= 256) k_closest_points<256><<<nb, 128, 0, st>>>(nodes, A->tri_wide, queries, n, out, A->n_nodes, order, unperm ? 1 : 0);
    else if(need <= 512) k_closest_points<512><<<nb, 128, 0, st>>>(nodes, A->tri_wide, queries, n, out, A->n_nodes, order, unperm ? 1 : 0);
    else return set_error("wide BVH too deep for the closest-point stack"), GPURT_E_STATE;
    if(unperm) k_cpq_unpermute<<<(unsigned)((2 * n + 255) / 256), 256, 0, st>>>(out, unperm, n, results);
    GPURT_CUDA(cudaGetLastError());
    return GPURT_OK;
}

} // namespace gpurt
