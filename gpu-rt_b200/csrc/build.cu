/*
 * build.cu — acceleration-structure build on the GPU.
 *
 * Replaces VK::Accel: every per-object BLAS build (src/vk/vulkan.cpp:881-936) plus the TLAS build
 * over instance transforms (src/vk/vulkan.cpp:777-856) issued by GPURT::build_accel
 * (src/gpurt.cpp:220-241).  Pipeline (all kernels hand-written, one stream):
 *
 *   k_flatten        instances -> world-space triangles (N1), exact AABBs, scene box (block-reduced)
 *   k_morton         63-bit Morton key of the AABB centroid (N6)
 *   k_rs_digit_hist, stable 8-bit LSD radix sort over 64-bit keys, "onesweep": one kernel per pass with
 *   k_rs_onesweep    decoupled look-back between tiles (ties keep gid order)
 *   k_karras         binary radix tree (Karras 2012), one thread per internal node
 *   k_refit          bottom-up AABB union with arrival counters (+ SAH-optimal collapse tables when asked for)
 *   k_collapse_*     level-synchronous collapse into 80-byte 8-wide nodes (bvh8.cuh): greedy largest-area
 *                    opening by default, runs of small levels inside one CTA
 *   k_tri_reorder    triangles re-laid out in node order
 *
 * Temporaries live in one arena owned by the context; gpurt_accel_update re-runs the pipeline in the buffers
 * the accel already owns.
 */
#include <cstdlib>
#include <cstring>

#include "../host/sah_split.h"
#include "device.cuh"

namespace gpurt {

/* ---------------------------------------------------------------------------------------------- */
int DevBuf::reserve(size_t bytes) {
    if(bytes <= cap) return GPURT_OK;
    if(p) cudaFree(p);
    p = nullptr, cap = 0;
    GPURT_CUDA(cudaMalloc(&p, bytes));
    cap = bytes;
    return GPURT_OK;
}
void DevBuf::release() {
    if(p) cudaFree(p);
    p = nullptr, cap = 0;
}
static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

/* ---- exclusive scan ---------------------------------------------------------------------------- */
constexpr int SCAN_T = 256, SCAN_I = 8, SCAN_TILE = SCAN_T * SCAN_I;

/* T = uint32_t, or uint64_t holding two 32-bit counters that are scanned together */
template <typename T> __device__ __forceinline__ T block_exclusive_scan(T v, T& total) {
    __shared__ T wsum[SCAN_T / 32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for(int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, inc, o);
        if(lane >= o) inc += t;
    }
    if(lane == 31) wsum[w] = inc;
    __syncthreads();
    if(w == 0) {
        T s = lane < SCAN_T / 32 ? wsum[lane] : 0;
#pragma unroll
        for(int o = 1; o < SCAN_T / 32; o <<= 1) {
            T t = __shfl_up_sync(0xffffffffu, s, o);
            if(lane >= o) s += t;
        }
        if(lane < SCAN_T / 32) wsum[lane] = s;
    }
    __syncthreads();
    T base = w ? wsum[w - 1] : 0;
    total = wsum[SCAN_T / 32 - 1];
    __syncthreads();
    return base + inc - v;
}

template <typename T>
__global__ void __launch_bounds__(SCAN_T) k_scan_local(const T* __restrict__ in, T* __restrict__ out, size_t n,
                                                       T* __restrict__ sums) {
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_I;
    T v[SCAN_I], s = 0;
#pragma unroll
    for(int i = 0; i < SCAN_I; i++) {
        v[i] = base + i < n ? in[base + i] : 0;
        s += v[i];
    }
    T total;
    T ex = block_exclusive_scan<T>(s, total);
#pragma unroll
    for(int i = 0; i < SCAN_I; i++) {
        if(base + i < n) out[base + i] = ex;
        ex += v[i];
    }
    if(threadIdx.x == 0) sums[blockIdx.x] = total;
}
template <typename T> __global__ void k_scan_add(T* __restrict__ out, size_t n, const T* __restrict__ offs) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) out[i] += offs[i / SCAN_TILE];
}

size_t scan_tmp_bytes(size_t n) {
    size_t total = 0;
    while(n > 1) {
        n = (n + SCAN_TILE - 1) / SCAN_TILE;
        total += (n + 1) * sizeof(uint64_t);
    }
    return total + 64;
}

template <typename T> static int scan_rec(cudaStream_t st, const T* in, T* out, size_t n, T* tmp) {
    unsigned nb = cdiv(n, SCAN_TILE);
    k_scan_local<T><<<nb, SCAN_T, 0, st>>>(in, out, n, tmp);
    if(nb > 1) {
        int rc = scan_rec<T>(st, tmp, tmp, nb, tmp + nb + 1);
        if(rc) return rc;
        k_scan_add<T><<<cdiv(n, 256), 256, 0, st>>>(out, n, tmp);
    }
    GPURT_CUDA(cudaGetLastError());
    return GPURT_OK;
}
/* ---- radix sort ------------------------------------------------------------------------------- */
/* Stable LSD radix sort of 64-bit keys with a 32-bit payload, 8 bits per pass, "onesweep" style:
 * one kernel histograms all 8 digits of every key, then each pass is ONE kernel that ranks a tile of
 * 4096 keys, obtains the tile's per-digit base with a decoupled look-back over the tiles before it
 * (tiles are numbered by an atomic ticket, so every predecessor is already running) and scatters.
 * Per pass every key and payload is read once and written once. */
constexpr int RS_T = 256, RS_I = 16, RS_TILE = RS_T * RS_I, RS_W = RS_T / 32;
constexpr uint32_t RS_FLAG_AGG = 1u << 30, RS_FLAG_PREFIX = 2u << 30, RS_VALUE = (1u << 30) - 1u;

__global__ void __launch_bounds__(RS_T) k_rs_digit_hist(const uint64_t* __restrict__ keys, size_t n, int passes,
                                                        uint32_t* __restrict__ ghist /* [passes][256] */) {
    __shared__ uint32_t h[8][256];
    for(int i = threadIdx.x; i < 8 * 256; i += RS_T) (&h[0][0])[i] = 0;
    __syncthreads();
    for(size_t i = (size_t)blockIdx.x * RS_T + threadIdx.x; i < n; i += (size_t)gridDim.x * RS_T) {
        uint64_t k = keys[i];
        for(int p = 0; p < passes; p++) atomicAdd(&h[p][(unsigned)(k >> (8 * p)) & 255u], 1u);
    }
    __syncthreads();
    for(int p = 0; p < passes; p++) {
        uint32_t c = h[p][threadIdx.x];
        if(c) atomicAdd(&ghist[p * 256 + threadIdx.x], c);
    }
}

constexpr size_t RS_SMEM = (size_t)RS_TILE * 12 + (size_t)RS_W * 256 * 4 + 2 * 256 * 4 + 16;

__global__ void __launch_bounds__(RS_T) k_rs_onesweep(const uint64_t* __restrict__ kin, const uint32_t* __restrict__ vin,
                                                      uint64_t* __restrict__ kout, uint32_t* __restrict__ vout,
                                                      size_t n, int shift, const uint32_t* __restrict__ ghist,
                                                      uint32_t* __restrict__ desc /* [tiles][256], zeroed */,
                                                      unsigned* __restrict__ ticket) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    uint64_t* s_key = reinterpret_cast<uint64_t*>(rs_smem);                       /* [RS_TILE] tile in digit order */
    uint32_t* s_val = reinterpret_cast<uint32_t*>(rs_smem + (size_t)RS_TILE * 8);  /* [RS_TILE] */
    uint32_t(*cnt)[256] = reinterpret_cast<uint32_t(*)[256]>(rs_smem + (size_t)RS_TILE * 12); /* [RS_W][256] */
    uint32_t* s_lstart = &cnt[RS_W][0];  /* [256] start of digit d inside the tile */
    uint32_t* s_gbase = s_lstart + 256;  /* [256] start of this tile's digit d in the output */
    unsigned* s_tile = s_gbase + 256;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if(threadIdx.x == 0) *s_tile = atomicAdd(ticket, 1u);
    for(int i = threadIdx.x; i < RS_W * 256; i += RS_T) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const unsigned tile = *s_tile;
    const size_t tile0 = (size_t)tile * RS_TILE;
    const size_t seg = tile0 + (size_t)w * (32 * RS_I);
    const unsigned tile_n = (unsigned)min((size_t)RS_TILE, n - tile0);
    uint64_t key[RS_I];
    uint32_t rank[RS_I];
    const unsigned lt = (1u << l) - 1u;
#pragma unroll
    for(int r = 0; r < RS_I; r++) {
        size_t i = seg + (size_t)r * 32 + l;
        bool valid = i < n;
        unsigned vm = __ballot_sync(0xffffffffu, valid);
        key[r] = 0;
        rank[r] = 0;
        if(valid) {
            key[r] = kin[i];
            unsigned d = (unsigned)(key[r] >> shift) & 255u;
            unsigned peers = __match_any_sync(vm, d);
            uint32_t before = cnt[w][d];
            rank[r] = before + __popc(peers & lt);
            __syncwarp(vm);
            if(l == __ffs(peers) - 1) cnt[w][d] = before + __popc(peers);
            __syncwarp(vm);
        }
    }
    __syncthreads();
    {
        /* thread d owns digit d: tile total, publish, look back, turn the per-warp counts into offsets */
        const unsigned d = threadIdx.x;
        uint32_t total = 0;
#pragma unroll
        for(int ww = 0; ww < RS_W; ww++) total += cnt[ww][d];
        volatile uint32_t* vd = desc;
        if(tile > 0) atomicExch(&desc[(size_t)tile * 256 + d], total | RS_FLAG_AGG);
        uint32_t excl = 0;
        for(int t = (int)tile - 1; t >= 0; t--) {
            uint32_t v;
            do { v = vd[(size_t)t * 256 + d]; } while((v >> 30) == 0u);
            excl += v & RS_VALUE;
            if(v & RS_FLAG_PREFIX) break;
        }
        atomicExch(&desc[(size_t)tile * 256 + d], (excl + total) | RS_FLAG_PREFIX);
        uint32_t gtot, ltot;
        uint32_t gbase = block_exclusive_scan<uint32_t>(ghist[d], gtot); /* start of digit d in the output */
        uint32_t lstart = block_exclusive_scan<uint32_t>(total, ltot);   /* start of digit d in the tile */
        s_gbase[d] = gbase + excl;
        s_lstart[d] = lstart;
        uint32_t run = lstart;
#pragma unroll
        for(int ww = 0; ww < RS_W; ww++) {
            uint32_t c = cnt[ww][d];
            cnt[ww][d] = run;
            run += c;
        }
    }
    __syncthreads();
    /* stage the tile in digit order, then write runs of equal digits with consecutive threads */
#pragma unroll
    for(int r = 0; r < RS_I; r++) {
        size_t i = seg + (size_t)r * 32 + l;
        if(i < n) {
            unsigned d = (unsigned)(key[r] >> shift) & 255u;
            uint32_t lp = cnt[w][d] + rank[r];
            s_key[lp] = key[r];
            s_val[lp] = vin[i];
        }
    }
    __syncthreads();
#pragma unroll
    for(int r = 0; r < RS_I; r++) {
        unsigned j = (unsigned)r * RS_T + threadIdx.x;
        if(j < tile_n) {
            uint64_t k = s_key[j];
            unsigned d = (unsigned)(k >> shift) & 255u;
            uint32_t pos = s_gbase[d] + (j - s_lstart[d]);
            kout[pos] = k;
            vout[pos] = s_val[j];
        }
    }
}

size_t radix_sort_tmp_bytes(size_t n, int passes) {
    size_t nb = (n + RS_TILE - 1) / RS_TILE;
    return (size_t)passes * (256 + 64) * 4 + (size_t)passes * nb * 256 * 4;
}

void preload_sort_kernels() {
    cudaFuncAttributes a;
    (void)cudaFuncGetAttributes(&a, (const void*)k_rs_digit_hist);
    (void)cudaFuncGetAttributes(&a, (const void*)k_rs_onesweep);
}
int radix_sort_u64(cudaStream_t st, uint64_t* keys, uint32_t* vals, uint64_t* kt, uint32_t* vt,
                   size_t n, int passes, DevBuf& tmp, int sm_count) {
    if(n == 0) return GPURT_OK;
    if(n > RS_VALUE || passes > 8) return set_error("radix sort: more than 2^30-1 keys"), GPURT_E_INVALID;
    unsigned nb = cdiv(n, RS_TILE);
    size_t bytes = radix_sort_tmp_bytes(n, passes);
    int rc = tmp.reserve(bytes);
    if(rc) return rc;
    uint32_t* ghist = tmp.as<uint32_t>();                 /* [passes][256] */
    unsigned* tickets = ghist + (size_t)passes * 256;      /* [passes] (64 reserved) */
    uint32_t* desc = ghist + (size_t)passes * (256 + 64);  /* [passes][nb][256] */
    GPURT_CUDA(cudaMemsetAsync(tmp.p, 0, bytes, st));
    GPURT_CUDA(cudaFuncSetAttribute(k_rs_onesweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_SMEM));
    k_rs_digit_hist<<<std::min(nb * (unsigned)RS_I, (unsigned)sm_count * 8u), RS_T, 0, st>>>(keys, n, passes, ghist);
    for(int p = 0; p < passes; p++) {
        k_rs_onesweep<<<nb, RS_T, RS_SMEM, st>>>(keys, vals, kt, vt, n, 8 * p, ghist + p * 256,
                                          desc + (size_t)p * nb * 256, tickets + p);
        uint64_t* a = keys;
        keys = kt, kt = a;
        uint32_t* b = vals;
        vals = vt, vt = b;
    }
    GPURT_CUDA(cudaGetLastError());
    return GPURT_OK; /* passes even -> result back in the caller's primary buffers */
}

/* ---- flatten ---------------------------------------------------------------------------------- */
__device__ __forceinline__ void atomic_min_f(float* a, float v) {
    if(v >= 0.0f) atomicMin((int*)a, __float_as_int(v));
    else atomicMax((unsigned*)a, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* a, float v) {
    if(v >= 0.0f) atomicMax((int*)a, __float_as_int(v));
    else atomicMin((unsigned*)a, __float_as_uint(v));
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for(int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for(int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

/* N1: world = model * vec4(v,1) evaluated as fma(m0,x, fma(m4,y, fma(m8,z, m12))).
 * Grid-stride over the triangles; the scene box is reduced per thread, per warp, per block and only then
 * with 6 atomics per block (the first version issued 6 same-address atomics per warp: 1.9 ms of the
 * 10 M-triangle build). */
__global__ void __launch_bounds__(256) k_flatten(DeviceScene S, float4* __restrict__ tri,
                                                 float4* __restrict__ tlo, float4* __restrict__ thi,
                                                 float* __restrict__ scene_box) {
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for(unsigned gid = blockIdx.x * blockDim.x + threadIdx.x; gid < S.n_tris; gid += gridDim.x * blockDim.x) {
        /* object of this triangle: last o with tri_off[o] <= gid */
        unsigned a = 0, b = S.n_objs;
        while(b - a > 1) {
            unsigned m = (a + b) >> 1;
            if(S.tri_off[m] <= gid) a = m; else b = m;
        }
        unsigned obj = a, prim = gid - S.tri_off[obj];
        const float* m = reinterpret_cast<const float*>(S.descs + obj); /* model is the first member */
        const float4* vb = reinterpret_cast<const float4*>(S.verts + S.vert_off[obj]); /* 3 x float4 per vertex */
        float w[3][3], tl[3], th[3];
#pragma unroll
        for(int k = 0; k < 3; k++) {
            float4 v = vb[3ull * S.idx[3ull * gid + k]]; /* pos.xyz | u */
            float x = v.x, y = v.y, z = v.z;
            w[k][0] = fmaf(m[0], x, fmaf(m[4], y, fmaf(m[8], z, m[12])));
            w[k][1] = fmaf(m[1], x, fmaf(m[5], y, fmaf(m[9], z, m[13])));
            w[k][2] = fmaf(m[2], x, fmaf(m[6], y, fmaf(m[10], z, m[14])));
        }
#pragma unroll
        for(int c = 0; c < 3; c++) {
            tl[c] = fminf(fminf(w[0][c], w[1][c]), w[2][c]);
            th[c] = fmaxf(fmaxf(w[0][c], w[1][c]), w[2][c]);
            lo[c] = fminf(lo[c], tl[c]);
            hi[c] = fmaxf(hi[c], th[c]);
        }
        tri[3ull * gid + 0] = make_float4(w[0][0], w[0][1], w[0][2], __uint_as_float(gid));
        tri[3ull * gid + 1] = make_float4(w[1][0] - w[0][0], w[1][1] - w[0][1], w[1][2] - w[0][2],
                                          __uint_as_float(obj));
        tri[3ull * gid + 2] = make_float4(w[2][0] - w[0][0], w[2][1] - w[0][1], w[2][2] - w[0][2],
                                          __uint_as_float(prim));
        tlo[gid] = make_float4(tl[0], tl[1], tl[2], 0.0f);
        thi[gid] = make_float4(th[0], th[1], th[2], 0.0f);
    }
    __shared__ float red[6][8];
#pragma unroll
    for(int c = 0; c < 3; c++) {
        float l = warp_min(lo[c]), h = warp_max(hi[c]);
        if((threadIdx.x & 31) == 0) red[c][threadIdx.x >> 5] = l, red[3 + c][threadIdx.x >> 5] = h;
    }
    __syncthreads();
    if(threadIdx.x < 6) {
        const int c = threadIdx.x;
        float v = red[c][0];
        for(int k = 1; k < 8; k++) v = c < 3 ? fminf(v, red[c][k]) : fmaxf(v, red[c][k]);
        if(c < 3) { if(v < 3.0e38f) atomic_min_f(scene_box + c, v); }
        else if(v > -3.0e38f) atomic_max_f(scene_box + c, v);
    }
}

__global__ void __launch_bounds__(256) k_morton(const float4* __restrict__ tlo,
                                                const float4* __restrict__ thi, unsigned n, float lx,
                                                float ly, float lz, float ix, float iy, float iz,
                                                uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
    if(g >= n) return;
    float4 lo = tlo[g], hi = thi[g];
    float cx = (lo.x + hi.x) * 0.5f, cy = (lo.y + hi.y) * 0.5f, cz = (lo.z + hi.z) * 0.5f;
    keys[g] = (expand21(quant21(cx, lx, ix)) << 2) | (expand21(quant21(cy, ly, iy)) << 1) |
              expand21(quant21(cz, lz, iz));
    vals[g] = g;
}

/* ---- Karras 2012 ------------------------------------------------------------------------------ */
__device__ __forceinline__ int key_delta(const uint64_t* __restrict__ k, int n, int i, int j) {
    if(j < 0 || j >= n) return -1;
    uint64_t a = k[i], b = k[j];
    if(a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

__global__ void __launch_bounds__(256) k_karras(const uint64_t* __restrict__ k, int n,
                                                int* __restrict__ left, int* __restrict__ right,
                                                int* __restrict__ parent /* [n-1 internal | n leaves] */,
                                                int* __restrict__ rf, int* __restrict__ rl) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n - 1) return;
    int d = (key_delta(k, n, i, i + 1) - key_delta(k, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = key_delta(k, n, i, i - d);
    int lmax = 2;
    while(key_delta(k, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for(int t = lmax / 2; t >= 1; t /= 2)
        if(key_delta(k, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = key_delta(k, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) / 2;
        if(key_delta(k, n, i, i + (s + t) * d) > dnode) s += t;
    } while(t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int L = (lo == gamma) ? ~gamma : gamma;
    int R = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    left[i] = L, right[i] = R;
    rf[i] = lo, rl[i] = hi;
    if(L < 0) parent[(n - 1) + ~L] = i; else parent[L] = i;
    if(R < 0) parent[(n - 1) + ~R] = i; else parent[R] = i;
    if(i == 0) parent[0] = -1;
}

template <bool SAH_TABLES>
__global__ void __launch_bounds__(256) k_refit(int n, const int* __restrict__ left,
                                               const int* __restrict__ right,
                                               const int* __restrict__ parent,
                                               const uint32_t* __restrict__ order,
                                               const float4* __restrict__ tlo,
                                               const float4* __restrict__ thi, float4* node_lo,
                                               float4* node_hi, unsigned* __restrict__ arrive, Bvh2View B,
                                               float* dp_cost, unsigned char* dp_dec) {
    int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if(leaf >= n) return;
    int p = parent[(n - 1) + leaf];
    while(p >= 0) {
        __threadfence();
        if(atomicAdd(&arrive[p], 1u) == 0u) return;
        int L = left[p], R = right[p];
        float4 al, ah, bl, bh;
        if(L < 0) { unsigned g = order[~L]; al = tlo[g], ah = thi[g]; }
        else { al = __ldcg(&node_lo[L]), ah = __ldcg(&node_hi[L]); }
        if(R < 0) { unsigned g = order[~R]; bl = tlo[g], bh = thi[g]; }
        else { bl = __ldcg(&node_lo[R]), bh = __ldcg(&node_hi[R]); }
        Box3 nb;
        nb.lo = f3(fminf(al.x, bl.x), fminf(al.y, bl.y), fminf(al.z, bl.z));
        nb.hi = f3(fmaxf(ah.x, bh.x), fmaxf(ah.y, bh.y), fmaxf(ah.z, bh.z));
        __stcg(&node_lo[p], make_float4(nb.lo.x, nb.lo.y, nb.lo.z, 0.0f));
        __stcg(&node_hi[p], make_float4(nb.hi.x, nb.hi.y, nb.hi.z, 0.0f));
        /* both subtrees are final: SAH-optimal collapse tables of this node (bvh8.cuh, dp_node) */
        if(SAH_TABLES) dp_node(B, dp_cost, dp_dec, p, nb);
        p = parent[p];
    }
}

/* ---- collapse --------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(128) k_collapse_count(Bvh2View B, const int* __restrict__ items,
                                                        unsigned n_items, int* __restrict__ children,
                                                        uint64_t* __restrict__ counts /* inner | tris << 32 */) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i == 0) counts[n_items] = 0; /* the exclusive scan leaves the level totals here */
    if(i >= n_items) return;
    int ch[8], nt;
    int ni = collapse_node(B, items[i], ch, nt);
#pragma unroll
    for(int s = 0; s < 8; s++) children[8ull * i + s] = ch[s];
    counts[i] = (uint64_t)(uint32_t)ni | ((uint64_t)(uint32_t)nt << 32);
}

__global__ void __launch_bounds__(128) k_collapse_emit(Bvh2View B, unsigned n_items,
                                                       const int* __restrict__ children,
                                                       const uint64_t* __restrict__ offsets,
                                                       unsigned level_base, unsigned next_base,
                                                       unsigned tri_cursor, Node8* __restrict__ nodes,
                                                       int* __restrict__ next_items,
                                                       uint32_t* __restrict__ wide_order) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n_items) return;
    int ch[8];
#pragma unroll
    for(int s = 0; s < 8; s++) ch[s] = children[8ull * i + s];
    const uint64_t off = offsets[i];
    const unsigned off_inner = (unsigned)off;
    unsigned child_base = next_base + off_inner;
    unsigned tri_base = tri_cursor + (unsigned)(off >> 32);
    Node8 node;
    encode_node(B, ch, child_base, tri_base, node);
    Node8* dst = nodes + (level_base + i);
#pragma unroll
    for(int k = 0; k < 5; k++) dst->v[k] = node.v[k];
    unsigned r = 0, t = 0;
    for(int s = 0; s < 8; s++) {
        int c = ch[s];
        if(c == kEmptyChild) continue;
        if(c >= 0) next_items[off_inner + r++] = c;
        else {
            unsigned first, count;
            decode_leaf_range(c, first, count);
            for(unsigned k = 0; k < count; k++, t++) wide_order[tri_base + t] = B.order[first + k];
        }
    }
}

/* Levels with few items (the top of the tree, and its last levels) are launch-latency bound when every level costs
 * three kernels and a host read-back: one CTA walks them instead — count, block scan, emit and the level switch all
 * inside the kernel — and stops at the first level with more than COLLAPSE_SMALL items (or none).  Same node and
 * triangle layout as the level-per-launch path. */
constexpr int COLLAPSE_SMALL = 512;
struct CollapseState {
    unsigned n_items, level_base, tri_cursor, depth;
};
__global__ void __launch_bounds__(COLLAPSE_SMALL) k_collapse_small(Bvh2View B, int* items_a, int* items_b,
                                                                   CollapseState* __restrict__ state,
                                                                   Node8* __restrict__ nodes,
                                                                   uint32_t* __restrict__ wide_order) {
    __shared__ unsigned long long wsum[COLLAPSE_SMALL / 32];
    __shared__ unsigned long long s_total;
    CollapseState S = *state;
    const unsigned i = threadIdx.x, lane = i & 31u, w = i >> 5;
    while(S.n_items > 0 && S.n_items <= (unsigned)COLLAPSE_SMALL && S.depth < 64u) {
        int ch[8], nt = 0, ni = 0;
        const bool valid = i < S.n_items;
        if(valid) ni = collapse_node(B, items_a[i], ch, nt);
        unsigned long long v = valid ? ((unsigned long long)(unsigned)ni | ((unsigned long long)(unsigned)nt << 32)) : 0ull;
        unsigned long long inc = v;
#pragma unroll
        for(int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if(lane >= (unsigned)o) inc += t;
        }
        if(lane == 31) wsum[w] = inc;
        __syncthreads();
        if(w == 0) {
            unsigned long long x = lane < COLLAPSE_SMALL / 32 ? wsum[lane] : 0ull;
#pragma unroll
            for(int o = 1; o < COLLAPSE_SMALL / 32; o <<= 1) {
                unsigned long long t = __shfl_up_sync(0xffffffffu, x, o);
                if(lane >= (unsigned)o) x += t;
            }
            if(lane < COLLAPSE_SMALL / 32) wsum[lane] = x;
            if(lane == COLLAPSE_SMALL / 32 - 1) s_total = x;
        }
        __syncthreads();
        const unsigned long long off = (w ? wsum[w - 1] : 0ull) + inc - v, total = s_total;
        const unsigned next_base = S.level_base + S.n_items;
        if(valid) {
            const unsigned off_inner = (unsigned)off, tri_base = S.tri_cursor + (unsigned)(off >> 32);
            Node8 node;
            encode_node(B, ch, next_base + off_inner, tri_base, node);
            Node8* dst = nodes + (S.level_base + i);
#pragma unroll
            for(int k = 0; k < 5; k++) dst->v[k] = node.v[k];
            unsigned r = 0, t = 0;
            for(int s = 0; s < 8; s++) {
                int c = ch[s];
                if(c == kEmptyChild) continue;
                if(c >= 0) items_b[off_inner + r++] = c;
                else {
                    unsigned first, count;
                    decode_leaf_range(c, first, count);
                    for(unsigned k = 0; k < count; k++, t++) wide_order[tri_base + t] = B.order[first + k];
                }
            }
        }
        S.level_base = next_base, S.n_items = (unsigned)total, S.tri_cursor += (unsigned)(total >> 32), S.depth++;
        int* sw = items_a;
        items_a = items_b, items_b = sw;
        __syncthreads(); /* next level reads what this one wrote; wsum / s_total are reused */
    }
    if(i == 0) *state = S;
}

/* The whole collapse in ONE launch (default; GPURT_COLLAPSE_LEVELS=1 keeps the level-per-launch path above): a grid of
 * co-resident CTAs (cooperative launch) walks the wide tree level by level.  Per level: every CTA takes a contiguous run
 * of the level's items, opens them (collapse_node) and scans their (children, triangles) counts locally; grid barrier;
 * every CTA adds up the CTA totals before it (at most a few hundred numbers) for its base offsets and the level totals,
 * encodes its nodes and writes the next level's items; grid barrier.  No host read-back, no launch per level: the 262 k
 * triangle collapse drops from ~0.47 ms (9 levels x 3 kernels + a stream synchronise each) to the time of its gathers.
 * Node, item and triangle order are those of the level-per-launch path (items stay in parent order, slots in slot order). */
constexpr int COLLAPSE_T = 128;
struct CollapseAll {
    int *items_a, *items_b, *children;
    uint64_t* offs;            /* per item: exclusive prefix of (children | tris << 32) inside its CTA's run */
    uint64_t* cta_tot;         /* per CTA: totals of its run */
    unsigned* barrier;         /* zeroed before the launch */
    CollapseState* state;      /* out: {0, n_nodes, n_tris placed, depth}; depth = 0xffffffff on overflow */
    unsigned max_nodes;
};
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& epoch) {
    __syncthreads();
    if(threadIdx.x == 0) {
        epoch++;
        __threadfence();
        atomicAdd(counter, 1u);
        const unsigned want = epoch * gridDim.x;
        while(*(volatile unsigned*)counter < want) {}
        __threadfence();
    }
    __syncthreads();
}
__global__ void __launch_bounds__(COLLAPSE_T) k_collapse_all(Bvh2View B, CollapseAll C, Node8* __restrict__ nodes,
                                                             uint32_t* __restrict__ wide_order) {
    __shared__ unsigned long long wsum[COLLAPSE_T / 32];
    __shared__ unsigned long long s_carry, s_base, s_total;
    const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    unsigned epoch = 0;
    unsigned n_items = 1, level_base = 0, tri_cursor = 0, depth = 0;
    int *items_a = C.items_a, *items_b = C.items_b;
    while(n_items) {
        const unsigned per = (n_items + gridDim.x - 1) / gridDim.x;
        const unsigned i0 = min(n_items, blockIdx.x * per), i1 = min(n_items, i0 + per);
        /* phase 1: open my items, local exclusive scan of their counts */
        if(tid == 0) s_carry = 0ull;
        __syncthreads();
        for(unsigned t0 = i0; t0 < i1; t0 += COLLAPSE_T) {
            const unsigned i = t0 + tid;
            const bool valid = i < i1;
            int ch[8], nt = 0, ni = 0;
            if(valid) {
                ni = collapse_node(B, __ldcg(items_a + i), ch, nt);
#pragma unroll
                for(int q = 0; q < 8; q++) C.children[8ull * i + q] = ch[q];
            }
            const unsigned long long v = valid ? ((unsigned long long)(unsigned)ni | ((unsigned long long)(unsigned)nt << 32)) : 0ull;
            unsigned long long inc = v;
#pragma unroll
            for(int o = 1; o < 32; o <<= 1) {
                unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
                if(lane >= (unsigned)o) inc += t;
            }
            if(lane == 31) wsum[w] = inc;
            __syncthreads();
            unsigned long long before = s_carry;
            for(unsigned k = 0; k < w; k++) before += wsum[k];
            if(valid) C.offs[i] = before + inc - v;
            __syncthreads();
            if(tid == COLLAPSE_T - 1) s_carry = before + inc;
            __syncthreads();
        }
        if(tid == 0) C.cta_tot[blockIdx.x] = s_carry;
        grid_barrier(C.barrier, epoch);
        /* phase 2: my base = totals of the CTAs before me; level totals = all of them */
        {
            unsigned long long mine = 0ull, all = 0ull;
            for(unsigned b = tid; b < gridDim.x; b += COLLAPSE_T) {
                const unsigned long long t = __ldcg(C.cta_tot + b);
                all += t;
                if(b < blockIdx.x) mine += t;
            }
#pragma unroll
            for(int o = 16; o > 0; o >>= 1) {
                mine += __shfl_xor_sync(0xffffffffu, mine, o);
                all += __shfl_xor_sync(0xffffffffu, all, o);
            }
            __shared__ unsigned long long r_mine[COLLAPSE_T / 32], r_all[COLLAPSE_T / 32];
            if(lane == 0) r_mine[w] = mine, r_all[w] = all;
            __syncthreads();
            if(tid == 0) {
                unsigned long long m = 0ull, a = 0ull;
                for(int k = 0; k < COLLAPSE_T / 32; k++) m += r_mine[k], a += r_all[k];
                s_base = m, s_total = a;
            }
            __syncthreads();
        }
        const unsigned long long base = s_base, total = s_total;
        const unsigned next_base = level_base + n_items;
        const unsigned n_next = (unsigned)total;
        if((size_t)next_base + n_next > C.max_nodes || depth + 1u >= 64u) { /* every CTA sees the same numbers */
            if(blockIdx.x == 0 && tid == 0) *C.state = CollapseState{0u, next_base, tri_cursor, 0xffffffffu};
            return;
        }
        for(unsigned i = i0 + tid; i < i1; i += COLLAPSE_T) {
            int ch[8];
#pragma unroll
            for(int q = 0; q < 8; q++) ch[q] = __ldcg(C.children + 8ull * i + q);
            const unsigned long long off = base + __ldcg(C.offs + i);
            const unsigned off_inner = (unsigned)off, tri_base = tri_cursor + (unsigned)(off >> 32);
            Node8 node;
            encode_node(B, ch, next_base + off_inner, tri_base, node);
            Node8* dst = nodes + (level_base + i);
#pragma unroll
            for(int k = 0; k < 5; k++) dst->v[k] = node.v[k];
            unsigned r = 0, t = 0;
            for(int q = 0; q < 8; q++) {
                int c = ch[q];
                if(c == kEmptyChild) continue;
                if(c >= 0) items_b[off_inner + r++] = c;
                else {
                    unsigned first, count;
                    decode_leaf_range(c, first, count);
                    for(unsigned k = 0; k < count; k++, t++) wide_order[tri_base + t] = B.order[first + k];
                }
            }
        }
        level_base = next_base, n_items = n_next, tri_cursor += (unsigned)(total >> 32), depth++;
        int* sw = items_a;
        items_a = items_b, items_b = sw;
        if(n_items) grid_barrier(C.barrier, epoch); /* the next level reads the items written above (through L2) */
    }
    if(blockIdx.x == 0 && tid == 0) *C.state = CollapseState{0u, level_base, tri_cursor, depth};
}

/* triangles into node order: slot j of the wide layout holds triangle wide_order[j]; one thread per float4 */
__global__ void __launch_bounds__(256) k_tri_reorder(const uint32_t* __restrict__ wide_order, size_t n,
                                                     const float4* __restrict__ tri_gid, float4* __restrict__ tri_wide) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= 3 * n) return;
    size_t j = i / 3;
    unsigned q = (unsigned)(i - 3 * j);
    tri_wide[i] = tri_gid[3ull * wide_order[j] + q];
}

/* n <= kMaxLeafTris: a single node whose slot 0 holds every triangle */
__global__ void k_single_leaf(Bvh2View B, unsigned n, Node8* nodes, const float4* tri_gid, float4* tri_wide) {
    if(threadIdx.x || blockIdx.x) return;
    int ch[8];
    for(int s = 0; s < 8; s++) ch[s] = kEmptyChild;
    ch[0] = encode_leaf_range(0, n);
    Node8 node;
    encode_node(B, ch, 0, 0, node);
    nodes[0] = node;
    for(unsigned k = 0; k < n; k++)
        for(int q = 0; q < 3; q++) tri_wide[3 * k + q] = tri_gid[3 * B.order[k] + q];
}

/* ---- orchestration ---------------------------------------------------------------------------- */
/* Buffers the accel keeps (allocated by the first build, reused by gpurt_accel_update): small ones come
 * from the device's stream-ordered pool, which keeps freed memory cached (a second build of a 262 k-triangle
 * scene finds all of them there), buffers of 32 MB and more use plain cudaMalloc, which is faster than
 * growing the pool for a one-off 10 M-triangle build.  Build temporaries are carved from one arena
 * owned by the context (grown on demand, never shrunk): no allocation call at all in a rebuild. */
static thread_local cudaStream_t t_alloc_stream = nullptr;
template <typename T> static int dkeep(T*& p, size_t count) {
    if(p) return GPURT_OK;
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    if(bytes >= (32u << 20)) GPURT_CUDA(cudaMalloc((void**)&p, bytes));
    else GPURT_CUDA(cudaMallocAsync((void**)&p, bytes, t_alloc_stream));
    return GPURT_OK;
}
struct Arena {
    char* base = nullptr;
    size_t used = 0, cap = 0;
    template <typename T> T* take(size_t count) {
        size_t bytes = (std::max<size_t>(count, 1) * sizeof(T) + 255) & ~(size_t)255;
        T* p = (T*)(base + used);
        used += bytes;
        return used <= cap ? p : nullptr;
    }
};
#define TRY(x)                                                                                     \
    do {                                                                                           \
        int rc_ = (x);                                                                             \
        if(rc_) return rc_;                                                                        \
    } while(0)

void free_accel_device(gpurt_accel* A) {
    void* ptrs[] = {A->tri_gid, A->tri_lo, A->tri_hi, A->keys, A->order, A->left, A->right,
                    A->node_lo, A->node_hi, A->nodes, A->tri_wide, A->parent, A->range_first, A->range_last, A->tree_cost_dev};
    for(void* p : ptrs)
        if(p) cudaFree(p);
    A->tri_gid = A->tri_lo = A->tri_hi = A->node_lo = A->node_hi = A->tri_wide = nullptr;
    A->keys = nullptr, A->order = nullptr, A->nodes = nullptr;
    A->left = A->right = A->parent = A->range_first = A->range_last = nullptr;
    A->tree_cost_dev = nullptr;
    free_scene(A->dscene);
}

/* sum over the inner nodes of 2 (ex ey + ey ez + ez ex): the surface-area-heuristic cost of the binary tree up to constant
 * factors; gpurt_accel_update_auto compares it with the value at the last full build to decide whether a refit still serves */
__global__ void __launch_bounds__(256) k_tree_cost(const float4* __restrict__ lo, const float4* __restrict__ hi, unsigned ni, float* out) {
    __shared__ float s[8];
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    float a = 0.0f;
    if(i < ni) {
        float4 l = lo[i], h = hi[i];
        float ex = h.x - l.x, ey = h.y - l.y, ez = h.z - l.z;
        a = 2.0f * (ex * ey + ey * ez + ez * ex);
        if(!(a == a) || a > 3.0e38f) a = 0.0f;
    }
#pragma unroll
    for(int d = 16; d >= 1; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
    if((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = a;
    __syncthreads();
    if(threadIdx.x == 0) {
        float t = 0.0f;
        for(int w = 0; w < 8; w++) t += s[w];
        atomicAdd(out, t);
    }
}

/* refit_only: poses changed, geometry did not, and the caller accepts the existing primitive order and topology: triangles
 * are flattened again, node boxes refitted, the wide tree collapsed and encoded again — the order / topology stage (the
 * SAH split or the Morton sort, most of a build) is skipped.  Any valid tree gives the same query results. */
int build_accel_device(gpurt_accel* A, bool refit_only) {
    gpurt_ctx* ctx = A->ctx;
    cudaStream_t st = ctx->stream;
    GPURT_CUDA(cudaSetDevice(ctx->device));
    t_alloc_stream = st;
    TRY(upload_scene(ctx, A->scene, A->dscene));
    const unsigned n = A->dscene.n_tris;
    const unsigned ni = n > 1 ? n - 1 : 0;
    const size_t max_nodes = (size_t)n / 2 + 2; /* every inner wide node owns > kMaxLeafTris triangles, >= 2 children */
    A->n = n;

    TRY(dkeep(A->tri_gid, 3ull * n));
    TRY(dkeep(A->tri_lo, n));
    TRY(dkeep(A->tri_hi, n));
    TRY(dkeep(A->keys, n));
    TRY(dkeep(A->order, n));
    TRY(dkeep(A->tri_wide, 3ull * n));
    TRY(dkeep(A->left, ni));
    TRY(dkeep(A->right, ni));
    TRY(dkeep(A->node_lo, ni));
    TRY(dkeep(A->node_hi, ni));
    TRY(dkeep(A->nodes, max_nodes));
    TRY(dkeep(A->parent, (size_t)ni + n));
    TRY(dkeep(A->range_first, ni));
    TRY(dkeep(A->range_last, ni));
    TRY(dkeep(A->tree_cost_dev, 1));

    /* temporaries: [box | range_first | range_last] live through the whole build; behind them phase 1
     * (sort + tree: keys_tmp, vals_tmp / arrival counters, parent) and phase 2 (collapse: item lists,
     * child lists, counters, scan scratch) share the same bytes */
    const size_t pad = 256;
    /* SAH-optimal collapse (GPURT_BUILD_SAH_COLLAPSE, or GPURT_COLLAPSE=sah in the environment) instead of the greedy one */
    static const bool env_sah = getenv("GPURT_COLLAPSE") && !strcmp(getenv("GPURT_COLLAPSE"), "sah");
    const bool sah = env_sah || (A->flags & GPURT_BUILD_SAH_COLLAPSE);
    size_t fixed = 4 * pad + 2 * ((size_t)ni * 4 + pad) + (sah ? (size_t)ni * 8 : 0);
    /* GPURT_BUILD_SAH_SPLIT: the binned-SAH tree is built on the device (sah_build.cu); GPURT_SAH_HOST=1 keeps the host
     * builder of host/sah_split.h (the definition both follow) for A/B runs and the device == host test */
    const bool sah_split = (A->flags & GPURT_BUILD_LBVH) == 0; /* the default build */
    const bool sah_host = getenv("GPURT_SAH_HOST") && atoi(getenv("GPURT_SAH_HOST")) != 0; /* read per build */
    const size_t sah_tmp = sah_split && !sah_host && n > 1 ? sah_split_tmp_bytes(n, ctx->sm_count) : 0;
    size_t phase1 = (size_t)n * 8 + (size_t)n * 4 + ((size_t)ni + n) * 4 + (sah ? (size_t)ni * 28 : 0) + sah_tmp + 6 * pad;
    size_t phase2 = 2 * max_nodes * 4 + 8 * max_nodes * 4 + (max_nodes + 1) * 8 + scan_tmp_bytes(max_nodes + 1) +
                    (size_t)n * 4 + 8 * pad;
    TRY(ctx->build_arena.reserve(fixed + std::max(phase1, phase2)));
    Arena ar;
    ar.base = (char*)ctx->build_arena.p, ar.cap = ctx->build_arena.cap;
    float* d_box = ar.take<float>(8);
    int* range_first = A->range_first;
    int* range_last = A->range_last;
    unsigned char* dp_dec = ar.take<unsigned char>(sah ? (size_t)ni * 8 : 1);
    const size_t phase_mark = ar.used;
    uint64_t* keys_tmp = ar.take<uint64_t>(n);
    uint32_t* vals_tmp = ar.take<uint32_t>(n);
    int* parent = A->parent;
    float* dp_cost = ar.take<float>(sah ? (size_t)ni * 7 : 1);
    char* sah_buf = ar.take<char>(sah_tmp);
    if(!dp_cost || !sah_buf) return set_error("build arena layout"), GPURT_E_STATE;

    struct EventPair { /* destroyed on every return path */
        cudaEvent_t a = nullptr, b = nullptr;
        ~EventPair() {
            if(a) cudaEventDestroy(a);
            if(b) cudaEventDestroy(b);
        }
    } ev;
    GPURT_CUDA(cudaEventCreate(&ev.a));
    GPURT_CUDA(cudaEventCreate(&ev.b));
    cudaEvent_t e0 = ev.a, e1 = ev.b;
    GPURT_CUDA(cudaEventRecord(e0, st));

    float init[6] = {3.0e38f, 3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f};
    GPURT_CUDA(cudaMemcpyAsync(d_box, init, sizeof(init), cudaMemcpyHostToDevice, st));
    if(n) k_flatten<<<std::min(cdiv(n, 256), (unsigned)ctx->sm_count * 16u), 256, 0, st>>>(A->dscene, A->tri_gid, A->tri_lo, A->tri_hi, d_box);
    GPURT_CUDA(cudaMemcpyAsync(A->scene_box, d_box, sizeof(init), cudaMemcpyDeviceToHost, st));
    GPURT_CUDA(cudaStreamSynchronize(st));
    if(n == 0) {
        for(float& f : A->scene_box) f = 0;
        A->n_nodes = 0, A->depth = 0, A->build_ms = 0;
        return GPURT_OK;
    }
    const float* sb = A->scene_box;
    float maxabs = 0;
    for(int k = 0; k < 6; k++) maxabs = fmaxf(maxabs, fabsf(sb[k]));
    /* N7, global part: 2^-19 * max|coord| (32 ulp of the largest coordinate) covers fp32 rounding of the
     * primitive tests; the node-relative part is added in encode_node (bvh8.cuh). */
    A->inflate = fmaxf(fmaxf(maxabs, 1e-30f) * 1.9073486328125e-06f, A->min_inflate);
    float ext[3] = {sb[3] - sb[0], sb[4] - sb[1], sb[5] - sb[2]}, inv[3];
    for(int k = 0; k < 3; k++) inv[k] = ext[k] > 0 ? 1.0f / ext[k] : 0.0f;

    if(refit_only) {
        /* order, keys and topology stay */
    } else if(!sah_split) {
        /* keys + sort */
        k_morton<<<cdiv(n, 256), 256, 0, st>>>(A->tri_lo, A->tri_hi, n, sb[0], sb[1], sb[2], inv[0], inv[1],
                                              inv[2], A->keys, A->order);
        TRY(radix_sort_u64(st, A->keys, A->order, keys_tmp, vals_tmp, n, 8, ctx->scratch, ctx->sm_count));
    } else if(!sah_host && n > 1) {
        TRY(build_sah_split_device(ctx, A->tri_lo, A->tri_hi, n, A->order, A->keys, A->left, A->right, parent, range_first,
                                   range_last, sah_buf, sah_tmp, nullptr));
    } else {
        /* GPURT_SAH_HOST=1: primitive order and binary topology from the host-side binned-SAH definition
         * (host/sah_split.h) over the boxes k_flatten just wrote; everything after it (refit, collapse, node encoding,
         * triangle re-layout) is the same device code */
        std::vector<float> h_lo(4ull * n), h_hi(4ull * n);
        GPURT_CUDA(cudaMemcpyAsync(h_lo.data(), A->tri_lo, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
        GPURT_CUDA(cudaMemcpyAsync(h_hi.data(), A->tri_hi, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
        GPURT_CUDA(cudaStreamSynchronize(st));
        SahSplitTree T;
        build_sah_split(h_lo.data(), h_hi.data(), 4, n, T);
        std::vector<uint64_t> h_keys(n);
        for(unsigned i = 0; i < n; i++) h_keys[i] = i; /* the "key" of a primitive is its position in the SAH order */
        GPURT_CUDA(cudaMemcpyAsync(A->order, T.order.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
        GPURT_CUDA(cudaMemcpyAsync(A->keys, h_keys.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
        if(ni) {
            GPURT_CUDA(cudaMemcpyAsync(A->left, T.left.data(), (size_t)ni * 4, cudaMemcpyHostToDevice, st));
            GPURT_CUDA(cudaMemcpyAsync(A->right, T.right.data(), (size_t)ni * 4, cudaMemcpyHostToDevice, st));
            GPURT_CUDA(cudaMemcpyAsync(parent, T.parent.data(), ((size_t)ni + n) * 4, cudaMemcpyHostToDevice, st));
            GPURT_CUDA(cudaMemcpyAsync(range_first, T.range_first.data(), (size_t)ni * 4, cudaMemcpyHostToDevice, st));
            GPURT_CUDA(cudaMemcpyAsync(range_last, T.range_last.data(), (size_t)ni * 4, cudaMemcpyHostToDevice, st));
        }
        GPURT_CUDA(cudaStreamSynchronize(st)); /* the host vectors go out of scope */
    }

    /* binary tree */
    Bvh2View B;
    B.left = A->left, B.right = A->right, B.range_first = range_first, B.range_last = range_last;
    B.node_lo = A->node_lo, B.node_hi = A->node_hi, B.tri_lo = A->tri_lo, B.tri_hi = A->tri_hi;
    B.order = A->order, B.inflate = A->inflate;
    if(const char* e = getenv("GPURT_SAH_CPRIM")) B.dp_cprim = (float)atof(e);
    unsigned* arrive = (unsigned*)vals_tmp; /* the sort is done with it */
    if(ni) {
        GPURT_CUDA(cudaMemsetAsync(arrive, 0, (size_t)ni * 4, st));
        if(!sah_split && !refit_only)
            k_karras<<<cdiv(ni, 256), 256, 0, st>>>(A->keys, (int)n, A->left, A->right, parent, range_first, range_last);
        if(sah)
            k_refit<true><<<cdiv(n, 256), 256, 0, st>>>((int)n, A->left, A->right, parent, A->order, A->tri_lo, A->tri_hi,
                                                       A->node_lo, A->node_hi, arrive, B, dp_cost, dp_dec);
        else
            k_refit<false><<<cdiv(n, 256), 256, 0, st>>>((int)n, A->left, A->right, parent, A->order, A->tri_lo, A->tri_hi,
                                                        A->node_lo, A->node_hi, arrive, B, dp_cost, dp_dec);
    }
    GPURT_CUDA(cudaMemsetAsync(A->tree_cost_dev, 0, 4, st));
    if(ni) k_tree_cost<<<cdiv(ni, 256), 256, 0, st>>>(A->node_lo, A->node_hi, ni, A->tree_cost_dev);
    GPURT_CUDA(cudaGetLastError());
    if(sah) B.dp_dec = dp_dec; /* the collapse follows the SAH-optimal decisions */

    /* wide collapse, one level of the wide tree per iteration */
    bool one_launch = false;
    CollapseState hs_all = {0u, 0u, 0u, 0u};
    if(n <= (unsigned)kMaxLeafTris) {
        k_single_leaf<<<1, 32, 0, st>>>(B, n, A->nodes, A->tri_gid, A->tri_wide);
        A->n_nodes = 1, A->depth = 1;
    } else {
        ar.used = phase_mark; /* phase 2 reuses the bytes of phase 1 (stream order makes that safe) */
        int* items_a = ar.take<int>(max_nodes);
        int* items_b = ar.take<int>(max_nodes);
        int* children = ar.take<int>(8 * max_nodes);
        uint64_t* cnt = ar.take<uint64_t>(max_nodes + 1);
        uint64_t* scan_tmp = ar.take<uint64_t>(scan_tmp_bytes(max_nodes + 1) / 8 + 1);
        uint32_t* wide_order = ar.take<uint32_t>(n);
        if(!wide_order) return set_error("build arena layout"), GPURT_E_STATE;
        CollapseState* d_state = ar.take<CollapseState>(1);
        if(!d_state) return set_error("build arena layout"), GPURT_E_STATE;
        int root = 0;
        GPURT_CUDA(cudaMemcpyAsync(items_a, &root, 4, cudaMemcpyHostToDevice, st));
        unsigned n_items = 1, level_base = 0, tri_cursor = 0, depth = 0;
        static const bool per_level = getenv("GPURT_COLLAPSE_LEVELS") && atoi(getenv("GPURT_COLLAPSE_LEVELS")) != 0;
        if(!per_level) { /* the whole collapse in one cooperative launch (k_collapse_all) */
            int per_sm = 0;
            GPURT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_collapse_all, COLLAPSE_T, 0));
            const unsigned want = (unsigned)std::max<size_t>(1, ((size_t)n / 6 + COLLAPSE_T - 1) / COLLAPSE_T); /* ~ the widest level */
            const unsigned grid = std::max(1u, std::min((unsigned)ctx->sm_count * (unsigned)std::max(1, std::min(per_sm, 8)), want));
            uint64_t* cta_tot = ar.take<uint64_t>(grid);
            unsigned* barrier = ar.take<unsigned>(64);
            if(per_sm > 0 && cta_tot && barrier) {
                GPURT_CUDA(cudaMemsetAsync(barrier, 0, 4, st));
                CollapseAll C{items_a, items_b, children, cnt, cta_tot, barrier, d_state, (unsigned)max_nodes};
                Node8* nodes_arg = A->nodes;
                void* args[] = {&B, &C, &nodes_arg, &wide_order};
                GPURT_CUDA(cudaLaunchCooperativeKernel((const void*)k_collapse_all, dim3(grid), dim3(COLLAPSE_T), args, 0, st));
                GPURT_CUDA(cudaMemcpyAsync(&hs_all, d_state, sizeof(hs_all), cudaMemcpyDeviceToHost, st));
                one_launch = true;
                n_items = 0;
            }
        }
        while(n_items) {
            if(n_items <= (unsigned)COLLAPSE_SMALL) { /* a run of small levels in one launch */
                CollapseState hs = {n_items, level_base, tri_cursor, depth};
                GPURT_CUDA(cudaMemcpyAsync(d_state, &hs, sizeof(hs), cudaMemcpyHostToDevice, st));
                k_collapse_small<<<1, COLLAPSE_SMALL, 0, st>>>(B, items_a, items_b, d_state, A->nodes, wide_order);
                GPURT_CUDA(cudaMemcpyAsync(&hs, d_state, sizeof(hs), cudaMemcpyDeviceToHost, st));
                GPURT_CUDA(cudaStreamSynchronize(st));
                if((hs.depth - depth) & 1u) { /* an odd number of levels swapped the item lists once more */
                    int* sw = items_a;
                    items_a = items_b, items_b = sw;
                }
                n_items = hs.n_items, level_base = hs.level_base, tri_cursor = hs.tri_cursor, depth = hs.depth;
                if(depth >= 64u || (size_t)level_base + n_items > max_nodes) return set_error("wide node bound exceeded"), GPURT_E_STATE;
                continue;
            }
            k_collapse_count<<<cdiv(n_items, 128), 128, 0, st>>>(B, items_a, n_items, children, cnt);
            TRY(scan_rec<uint64_t>(st, cnt, cnt, n_items + 1, scan_tmp));
            unsigned next_base = level_base + n_items;
            k_collapse_emit<<<cdiv(n_items, 128), 128, 0, st>>>(B, n_items, children, cnt, level_base, next_base,
                                                               tri_cursor, A->nodes, items_b, wide_order);
            uint32_t tot[2]; /* {inner children, triangles} emitted by this level */
            GPURT_CUDA(cudaMemcpyAsync(tot, cnt + n_items, 8, cudaMemcpyDeviceToHost, st));
            GPURT_CUDA(cudaStreamSynchronize(st));
            level_base = next_base;
            tri_cursor += tot[1];
            n_items = tot[0];
            int* sw = items_a;
            items_a = items_b, items_b = sw;
            depth++;
            if((size_t)level_base + n_items > max_nodes) return set_error("wide node bound exceeded"), GPURT_E_STATE;
        }
        k_tri_reorder<<<cdiv(3ull * n, 256), 256, 0, st>>>(wide_order, n, A->tri_gid, A->tri_wide);
        if(!one_launch) {
            A->n_nodes = level_base, A->depth = depth;
            if(tri_cursor != n)
                return set_error("collapse lost triangles: " + std::to_string(tri_cursor) + " of " + std::to_string(n)), GPURT_E_STATE;
        }
    }
    GPURT_CUDA(cudaEventRecord(e1, st));
    GPURT_CUDA(cudaMemcpyAsync(&A->tree_cost, A->tree_cost_dev, 4, cudaMemcpyDeviceToHost, st));
    GPURT_CUDA(cudaStreamSynchronize(st)); /* the only host wait of the collapse when it ran as one launch */
    if(one_launch) {
        if(hs_all.depth == 0xffffffffu) return set_error("wide node bound exceeded"), GPURT_E_STATE;
        A->n_nodes = hs_all.level_base, A->depth = hs_all.depth;
        if(hs_all.tri_cursor != n)
            return set_error("collapse lost triangles: " + std::to_string(hs_all.tri_cursor) + " of " + std::to_string(n)), GPURT_E_STATE;
    }
    if(refit_only) A->refits++;
    else A->tree_cost_at_build = A->tree_cost, A->refits = 0;
    cudaEventElapsedTime(&A->build_ms, e0, e1);
    GPURT_CUDA(cudaGetLastError());
    return GPURT_OK;
}

} // namespace gpurt
