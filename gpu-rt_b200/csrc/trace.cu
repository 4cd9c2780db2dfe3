/*
 * trace.cu — closest-hit and any-hit traversal of the 8-wide compressed BVH.
 *
 * Replaces traceRayEXT as used by trace_ray (src/shaders/rt/rt.rgen:257-270: opaque, cull mask
 * 0xFF, interval (tmin,tmax) exclusive) together with rt.rchit:11-16 / rt.rmiss:9-11 (payload
 * write), and `visibility` (rt.rgen:272-291: terminate on first hit, skip closest-hit shader).
 * B200 has no RT cores, so this is a software traversal: one ray per thread, 5 x 16-byte loads per
 * node, 3 x 16-byte loads per triangle, an 8-byte-per-entry traversal stack, octant-ordered
 * front-to-back descent without distance sorting (bvh8.cuh).
 *
 * N4 tie rule: nearest t wins, equal t -> lowest global primitive id, so the result does not depend
 * on traversal order and equals the brute-force oracle bit for bit.
 */
#include <algorithm>

#include "device.cuh"
#include "traverse.cuh"

namespace gpurt {

/* experiment hooks (tools/build_variant.sh): threads per CTA and minimum CTAs per SM of k_trace_closest */
#ifndef GPURT_TRACE_BLOCK
#define GPURT_TRACE_BLOCK 128
#endif
#ifdef GPURT_TRACE_MINB
#define GPURT_TRACE_BOUNDS __launch_bounds__(GPURT_TRACE_BLOCK, GPURT_TRACE_MINB)
#else
#define GPURT_TRACE_BOUNDS __launch_bounds__(GPURT_TRACE_BLOCK)
#endif

/* ORDERED: the batch is processed through a sorted index (order.cu); the plain instantiation is the headline kernel */
template <bool STATS, bool ORDERED>
__global__ void GPURT_TRACE_BOUNDS k_trace_closest(const float4* __restrict__ nodes,
                                                       const float4* __restrict__ tris,
                                                       const float4* __restrict__ rays, uint64_t n,
                                                       float4* __restrict__ hits, unsigned n_nodes,
                                                       unsigned long long* counters,
                                                       const uint32_t* __restrict__ order, int staged) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const uint64_t slot = i;        /* processing position */
    if(ORDERED) i = order[i];       /* storage position */
    float4 a = __ldg(rays + 2 * i), b = __ldg(rays + 2 * i + 1);
    HitRec h;
    h.t = a.w, h.u = h.v = 0, h.gid = kNoHit;
    if(n_nodes)
        traverse8<false, STATS>(nodes, tris, f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), a.w, b.w, h, counters);
    float4 out;
    out.x = h.gid == kNoHit ? GPURT_INF : h.t;
    out.y = h.u, out.z = h.v, out.w = u2f(h.gid);
    hits[ORDERED && staged ? slot : i] = out;
}

template <bool ORDERED>
__global__ void __launch_bounds__(128) k_trace_any(const float4* __restrict__ nodes,
                                                   const float4* __restrict__ tris,
                                                   const float4* __restrict__ rays, uint64_t n,
                                                   uint8_t* __restrict__ occ, unsigned n_nodes,
                                                   const uint32_t* __restrict__ order, int staged) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const uint64_t slot = i;
    if(ORDERED) i = order[i];
    float4 a = __ldg(rays + 2 * i), b = __ldg(rays + 2 * i + 1);
    HitRec h;
    bool hit = n_nodes && traverse8<true, false>(nodes, tris, f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), a.w,
                                                  b.w, h, nullptr);
    occ[ORDERED && staged ? slot : i] = hit ? 1 : 0;
}

/* ---- binary-LBVH traversal (debug / baseline for the wide path) ------------------------------- */
__device__ __forceinline__ bool slab_test(float4 lo, float4 hi, float e, F3 o, F3 id, float tmin,
                                          float tmax, float& tn) {
    float x0 = (lo.x - e - o.x) * id.x, x1 = (hi.x + e - o.x) * id.x;
    float y0 = (lo.y - e - o.y) * id.y, y1 = (hi.y + e - o.y) * id.y;
    float z0 = (lo.z - e - o.z) * id.z, z1 = (hi.z + e - o.z) * id.z;
    tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), tmin));
    float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmax));
    return tn <= tf;
}

__global__ void __launch_bounds__(128) k_trace_closest_bvh2(
    const int* __restrict__ left, const int* __restrict__ right, const float4* __restrict__ node_lo,
    const float4* __restrict__ node_hi, const float4* __restrict__ tri_lo,
    const float4* __restrict__ tri_hi, const uint32_t* __restrict__ order,
    const float4* __restrict__ tri_gid, float inflate, unsigned n_tris,
    const float4* __restrict__ rays, uint64_t n, float4* __restrict__ hits) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    float4 a = __ldg(rays + 2 * i), b = __ldg(rays + 2 * i + 1);
    F3 o = f3(a.x, a.y, a.z), d = f3(b.x, b.y, b.z);
    F3 id = f3(safe_rcp_dir(d.x), safe_rcp_dir(d.y), safe_rcp_dir(d.z));
    float tmin = a.w, tmax = b.w;
    HitRec best;
    best.t = tmax, best.u = best.v = 0, best.gid = kNoHit;
    int stack[128];
    int sp = 0;
    if(n_tris == 1) stack[sp++] = ~0;
    else if(n_tris > 1) stack[sp++] = 0;
    while(sp) {
        int c = stack[--sp];
        if(c < 0) {
            unsigned g = order[~c];
            float4 r0 = tri_gid[3ull * g], r1 = tri_gid[3ull * g + 1], r2 = tri_gid[3ull * g + 2];
            float t, u, v;
            if(intersect_tri(o, d, tmin, tmax, f3(r0.x, r0.y, r0.z), f3(r1.x, r1.y, r1.z),
                             f3(r2.x, r2.y, r2.z), t, u, v))
                if(t < best.t || (t == best.t && g < best.gid)) best.t = t, best.u = u, best.v = v, best.gid = g;
            continue;
        }
        int ch[2] = {left[c], right[c]};
        float tn[2];
        bool h[2];
#pragma unroll
        for(int k = 0; k < 2; k++) {
            float4 lo, hi;
            if(ch[k] < 0) { unsigned g = order[~ch[k]]; lo = tri_lo[g], hi = tri_hi[g]; }
            else lo = node_lo[ch[k]], hi = node_hi[ch[k]];
            h[k] = slab_test(lo, hi, inflate, o, id, tmin, best.t, tn[k]);
        }
        if(sp + 2 > 128) break; /* cannot happen: depth <= 96 */
        if(h[0] && h[1]) {
            if(tn[0] <= tn[1]) stack[sp++] = ch[1], stack[sp++] = ch[0];
            else stack[sp++] = ch[0], stack[sp++] = ch[1];
        } else if(h[0]) stack[sp++] = ch[0];
        else if(h[1]) stack[sp++] = ch[1];
    }
    float4 out;
    out.x = best.gid == kNoHit ? GPURT_INF : best.t;
    out.y = best.u, out.z = best.v, out.w = u2f(best.gid);
    hits[i] = out;
}

/* ---- launchers -------------------------------------------------------------------------------- */
static inline unsigned blocks_for(uint64_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }

int launch_trace_closest(gpurt_accel* A, const float4* rays, uint64_t n, float4* hits) {
    if(!n) return GPURT_OK;
    OrderPlan P; /* large incoherent batches on large scenes are processed in Morton order of the ray origin (order.cu) */
    int rc = plan_spatial_order(A, rays, 2, n, hits, 16, P, true);
    if(rc) return rc;
    if(P.scatter) { /* hits live on another GPU: slice k is stored there while slice k + 1 is traced (order.cu) */
        std::vector<uint64_t> ends;
        order_slices(n, ends);
        uint64_t off = 0;
        for(uint64_t e : ends) {
            const uint64_t m = e - off;
            k_trace_closest<false, true><<<blocks_for(m, GPURT_TRACE_BLOCK), GPURT_TRACE_BLOCK, 0, A->ctx->stream>>>(
                (const float4*)A->nodes, A->tri_wide, rays, m, (float4*)P.out + off, A->n_nodes, nullptr, P.order + off, 1);
            GPURT_CUDA(cudaGetLastError());
            if((rc = scatter_slice_async(A, P, off, m, hits, 16))) return rc;
            off = e;
        }
        return scatter_join(A, P);
    }
    if(P.order)
        k_trace_closest<false, true><<<blocks_for(n, GPURT_TRACE_BLOCK), GPURT_TRACE_BLOCK, 0, A->ctx->stream>>>(
            (const float4*)A->nodes, A->tri_wide, rays, n, (float4*)P.out, A->n_nodes, nullptr, P.order, P.unperm ? 1 : 0);
    else
        k_trace_closest<false, false><<<blocks_for(n, GPURT_TRACE_BLOCK), GPURT_TRACE_BLOCK, 0, A->ctx->stream>>>(
            (const float4*)A->nodes, A->tri_wide, rays, n, hits, A->n_nodes, nullptr, nullptr, 0);
    GPURT_CUDA(cudaGetLastError());
    return finish_spatial_order(A, P, n, hits, 16);
}
int launch_trace_closest_stats(gpurt_accel* A, const float4* rays, uint64_t n, float4* hits,
                               unsigned long long* d_counters) {
    if(!n) return GPURT_OK;
    k_trace_closest<true, false><<<blocks_for(n, GPURT_TRACE_BLOCK), GPURT_TRACE_BLOCK, 0, A->ctx->stream>>>(
        (const float4*)A->nodes, A->tri_wide, rays, n, hits, A->n_nodes, d_counters, nullptr, 0);
    GPURT_CUDA(cudaGetLastError());
    return GPURT_OK;
}
int launch_trace_any(gpurt_accel* A, const float4* rays, uint64_t n, uint8_t* occ) {
    if(!n) return GPURT_OK;
    OrderPlan P;
    int rc = plan_spatial_order(A, rays, 2, n, occ, 1, P);
    if(rc) return rc;
    if(P.order)
        k_trace_any<true><<<blocks_for(n, 128), 128, 0, A->ctx->stream>>>((const float4*)A->nodes, A->tri_wide, rays, n,
                                                                        (uint8_t*)P.out, A->n_nodes, P.order, P.unperm ? 1 : 0);
    else
        k_trace_any<false><<<blocks_for(n, 128), 128, 0, A->ctx->stream>>>((const float4*)A->nodes, A->tri_wide, rays, n, occ,
                                                                         A->n_nodes, nullptr, 0);
    GPURT_CUDA(cudaGetLastError());
    return finish_spatial_order(A, P, n, occ, 1);
}
int launch_trace_closest_bvh2(gpurt_accel* A, const float4* rays, uint64_t n, float4* hits) {
    if(!n) return GPURT_OK;
    k_trace_closest_bvh2<<<blocks_for(n, 128), 128, 0, A->ctx->stream>>>(
        A->left, A->right, A->node_lo, A->node_hi, A->tri_lo, A->tri_hi, A->order, A->tri_gid, A->inflate,
        A->n, rays, n, hits);
    GPURT_CUDA(cudaGetLastError());
    return GPURT_OK;
}

void preload_trace_kernels() {
    cudaFuncAttributes a;
    (void)cudaFuncGetAttributes(&a, (const void*)k_trace_closest<false, true>);
    (void)cudaFuncGetAttributes(&a, (const void*)k_trace_closest<false, false>);
}

} // namespace gpurt
