/*
 * cpq.cu — closest-point query over the same 8-wide BVH (the reference's FCPW-GPU work,
 * README.md:6-8; no source in the snapshot, FCPW semantics: nearest point on any triangle within a
 * search radius, its distance and primitive).
 *
 * Priority-ordered descent (traverse.cuh: closest_point8).  N5 tie rule: smallest d^2 wins, equal
 * d^2 -> lowest global primitive id; the radius is inclusive.
 */
#include "device.cuh"
#include "traverse.cuh"

namespace gpurt {

template <int STACK>
__global__ void __launch_bounds__(128) k_closest_points(const float4* __restrict__ nodes,
                                                        const float4* __restrict__ tris,
                                                        const float4* __restrict__ queries, uint64_t n,
                                                        float4* __restrict__ results, unsigned n_nodes) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
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
    results[2 * i] = o0;
    results[2 * i + 1] = o1;
}

int launch_closest_points(gpurt_accel* A, const float4* queries, uint64_t n, float4* results) {
    if(!n) return GPURT_OK;
    unsigned nb = (unsigned)((n + 127) / 128);
    const float4* nodes = (const float4*)A->nodes;
    cudaStream_t st = A->ctx->stream;
    unsigned need = 7u * A->depth + 1u;
    if(need <= 64) k_closest_points<64><<<nb, 128, 0, st>>>(nodes, A->tri_wide, queries, n, results, A->n_nodes);
    else if(need <= 128) k_closest_points<128><<<nb, 128, 0, st>>>(nodes, A->tri_wide, queries, n, results, A->n_nodes);
    else if(need <= 256) k_closest_points<256><<<nb, 128, 0, st>>>(nodes, A->tri_wide, queries, n, results, A->n_nodes);
    else if(need <= 512) k_closest_points<512><<<nb, 128, 0, st>>>(nodes, A->tri_wide, queries, n, results, A->n_nodes);
    else return set_error("wide BVH too deep for the closest-point stack"), GPURT_E_STATE;
    GPURT_CUDA(cudaGetLastError());
    return GPURT_OK;
}

} // namespace gpurt
