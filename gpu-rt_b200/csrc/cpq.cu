/*
 * cpq.cu — closest-point query over the same 8-wide BVH (the reference's FCPW-GPU work,
 * README.md:6-8; no source in the snapshot, FCPW semantics: nearest point on any triangle within a
 * search radius, its distance and primitive).
 *
 * Priority-ordered descent (traverse.cuh: closest_point8).  N5 tie rule: smallest d^2 wins, equal
 * d^2 -> lowest global primitive id; the radius is inclusive.
 */
#include <algorithm>

#include "device.cuh"
#include "traverse.cuh"

namespace gpurt {

#ifndef GPURT_CPQ_MINB
#define GPURT_CPQ_MINB 10 /* minimum CTAs per SM asked of k_closest_points: 48 registers instead of 54 (8 B of spills); bench
                            queries 1388 -> 1430 Mq/s; 8 CTAs (58 registers) 1315, 12 CTAs (40 registers, 140 B spills) 1350
                            (tools/ab_cpq.sh, variants built side by side, interleaved x 3) */
#endif
#ifndef GPURT_CPQ_BLOCK
#define GPURT_CPQ_BLOCK 128 /* threads per CTA (experiment hook, with GPURT_CPQ_MINB: tools/build_variant.sh) */
#endif
template <int STACK, bool ORDERED>
__global__ void __launch_bounds__(GPURT_CPQ_BLOCK, GPURT_CPQ_MINB) k_closest_points(const float4* __restrict__ nodes,
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
    if(ORDERED && staged) i = slot; /* results go to a local staging array in processing order (order.cu) */
    results[2 * i] = o0;
    results[2 * i + 1] = o1;
}

/* instrumented launch for the roofline (mean nodes visited / triangles tested per query); results unchanged */
__global__ void __launch_bounds__(128) k_closest_points_stats(const float4* __restrict__ nodes, const float4* __restrict__ tris,
                                                              const float4* __restrict__ queries, uint64_t n,
                                                              unsigned n_nodes, unsigned long long* counters) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    float4 q = __ldg(queries + i);
    CpRec best;
    best.gid = kNoHit;
    unsigned vc[2] = {0u, 0u};
    if(n_nodes) closest_point8<512>(nodes, tris, f3(q.x, q.y, q.z), q.w, best, vc);
    atomicAdd(counters + 0, (unsigned long long)vc[0]);
    atomicAdd(counters + 1, (unsigned long long)vc[1]);
    if(best.gid != kNoHit) atomicAdd(counters + 2, 1ull);
}
int launch_closest_points_stats(gpurt_accel* A, const float4* queries, uint64_t n, unsigned long long* d_counters) {
    if(!n) return GPURT_OK;
    if(7u * A->depth + 1u > 512u) return set_error("wide BVH too deep for the closest-point stack"), GPURT_E_STATE;
    k_closest_points_stats<<<(unsigned)((n + 127) / 128), 128, 0, A->ctx->stream>>>((const float4*)A->nodes, A->tri_wide, queries, n,
                                                                                   A->n_nodes, d_counters);
    GPURT_CUDA(cudaGetLastError());
    return GPURT_OK;
}

int launch_closest_points(gpurt_accel* A, const float4* queries, uint64_t n, float4* results) {
    if(!n) return GPURT_OK;
    const float4* nodes = (const float4*)A->nodes;
    cudaStream_t st = A->ctx->stream;
    /* large incoherent batches are processed in Morton order of the query point (order.cu) — on every scene larger than
     * the L1s: the descent diverges with the spread of a warp's points even when the whole tree sits in L2 (stand-in, 2 M
     * points jittered by +-200 units: 738 -> 1101 Mq/s including the sort; profiles/r04c_cpq_order.log) */
    OrderPlan P;
    int rc = plan_spatial_order(A, queries, 1, n, results, 32, P, true, true);
    if(rc) return rc;
    unsigned need = 7u * A->depth + 1u;
    if(need > 512) return set_error("wide BVH too deep for the closest-point stack"), GPURT_E_STATE;
    /* [off, off + m) of the processing order; the whole batch unless results are scattered to another GPU slice by slice */
    auto launch = [&](uint64_t off, uint64_t m) {
        const unsigned nb = (unsigned)((m + GPURT_CPQ_BLOCK - 1) / GPURT_CPQ_BLOCK);
        float4* out = (float4*)P.out + (P.scatter ? 2 * off : 0);
        const uint32_t* order = P.order ? P.order + off : nullptr;
        const int staged = (P.unperm || P.scatter) ? 1 : 0;
#define GPURT_CPQ_LAUNCH(S)                                                                                                  \
    (order ? k_closest_points<S, true><<<nb, GPURT_CPQ_BLOCK, 0, st>>>(nodes, A->tri_wide, queries, m, out, A->n_nodes, order, staged) \
           : k_closest_points<S, false><<<nb, GPURT_CPQ_BLOCK, 0, st>>>(nodes, A->tri_wide, queries, m, out, A->n_nodes, nullptr, 0))
        if(need <= 64) GPURT_CPQ_LAUNCH(64);
        else if(need <= 128) GPURT_CPQ_LAUNCH(128);
        else if(need <= 256) GPURT_CPQ_LAUNCH(256);
        else GPURT_CPQ_LAUNCH(512);
#undef GPURT_CPQ_LAUNCH
    };
    if(P.scatter) {
        std::vector<uint64_t> ends;
        order_slices(n, ends);
        uint64_t off = 0;
        for(uint64_t e : ends) {
            const uint64_t m = e - off;
            launch(off, m);
            GPURT_CUDA(cudaGetLastError());
            if((rc = scatter_slice_async(A, P, off, m, results, 32))) return rc;
            off = e;
        }
        return scatter_join(A, P);
    }
    launch(0, n);
    GPURT_CUDA(cudaGetLastError());
    return finish_spatial_order(A, P, n, results, 32);
}

void preload_cpq_kernels() {
    cudaFuncAttributes a;
#define GPURT_PRELOAD(S) (void)cudaFuncGetAttributes(&a, (const void*)k_closest_points<S, true>), (void)cudaFuncGetAttributes(&a, (const void*)k_closest_points<S, false>)
    GPURT_PRELOAD(64), GPURT_PRELOAD(128), GPURT_PRELOAD(256), GPURT_PRELOAD(512);
#undef GPURT_PRELOAD
}

} // namespace gpurt
