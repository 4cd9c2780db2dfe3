/*
 * sah_build.cu — the binned-SAH binary tree of host/sah_split.h, built on the device.
 *
 * The reference asks its driver for a PREFER_FAST_TRACE acceleration structure (src/vk/vulkan.cpp:780, :884); here that is
 * a top-down 16-bin SAH split of the triangle boxes k_flatten wrote, and this file reproduces the tree DEFINED in
 * host/sah_split.h bit for bit (same centroid bounds, bins, costs, tie rules, stable partitions, preorder node ids):
 * every reduction is a min / max / integer count, so the order in which a parallel machine performs it does not matter.
 *
 * Two regimes:
 *
 *   segments of more than 1024 items ("big"), level by level over the whole grid, fixed tiles of 1024 positions:
 *     k_sah_bins     per warp, lane L owns bin L&15 of axis L>>4 (third axis in a second register set): the 32 items of a
 *                    step are broadcast by shuffle and each lane folds the ones that fall into its bin — no atomics, no
 *                    conflicts; one flush of 48 x 7 words per warp and segment into the segment's global bins.  The last
 *                    tile to finish a segment evaluates the 45 split candidates.
 *     k_sah_count    left-going items per tile and segment
 *     k_sah_level    ONE block: exclusive scan of the tile counts, then per segment: boundary, node record, children
 *                    (leaf / small job / next level's big segment), bins and centroid bounds of the new big segments
 *     k_sah_scatter  stable partition into the other record array, centroid bounds of big children (match_any +
 *                    redux), primitive ids of single-item children
 *   segments of at most 1024 items: k_sah_small, persistent warps over a ticket queue.  A warp owns one segment: bounds by
 *     warp reduction, the same transposed binning, split, stable partition by ballot, then it keeps the left child and
 *     queues the right one.
 *
 * Item records (box + primitive id, 32 bytes) are physically partitioned between two arrays, so every pass reads
 * contiguous memory.
 *
 * Measured and not kept: binning by match.any + group-masked redux.sync (group-masked redux serialises per group: 2.7x
 * slower than the register-transposed loop); <= 32-item subtrees level by level with every lane walking the members of
 * its segment (bit-identical, 3-8 % slower than node by node: ~70 warp instructions per member and axis).
 * -DSAH_PROFILE prints where k_sah_small's cycles go: on 262 k triangles 79 % in <= 32-item subtrees (~10 k warp
 * instructions each), which is also the tail of the build's critical path.
 */
#include <cstdio>
#include <cstdlib>

#include "device.cuh"

namespace gpurt {

namespace {

constexpr unsigned SAH_TILE = 1024;      /* positions per tile of the grid-wide kernels */
constexpr unsigned SAH_SMALL_MAX = 1024; /* default: segments up to this size are built by one warp of k_sah_small (>= SAH_TILE) */
constexpr int SAH_BINS = 16;
constexpr float SAH_BIG = 3.0e38f;
constexpr unsigned FULL = 0xffffffffu;
constexpr int SAH_BIN_WORDS = 3 * SAH_BINS * 7; /* per segment: [axis][bin]{count, lo xyz, hi xyz} */

struct SahRec {
    float4 lo; /* w: primitive id */
    float4 hi;
};
struct SahSeg {
    unsigned a, b;
    int me, parent;
};
struct SahSplit {
    unsigned a, b;
    int axis, k; /* axis < 0: split at the middle */
    float lo, scale;
    unsigned nl, base;
    int child[2]; /* >= 0: big segment of the next level, -1: small job, -2: single item */
};
struct SahJob {
    unsigned a, b;
    int me, parent;
    unsigned src;
};
struct SahState { /* the hot counters of k_sah_small live on cache lines of their own */
    unsigned n_seg, n_seg_next, small_total, pad0[29];
    unsigned q_tail, pad1[31];
    unsigned q_head, pad2[31];
    unsigned leaves_done, pad3[31];
    unsigned finished, stalled, pad4[30];
    unsigned long long prof[4]; /* -DSAH_PROFILE: cycles in <= 32-item subtrees, in larger nodes, waiting for tickets; subtrees */
};

/* order-preserving float <-> int for atomicMin / atomicMax */
__device__ __forceinline__ int enc(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float dec(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

/* sah_split.h bin_of */
__device__ __forceinline__ int sah_bin(float c, float lo, float scale) {
    const float f = (c - lo) * scale;
    return f >= 0.0f ? (f < (float)SAH_BINS ? (int)f : SAH_BINS - 1) : 0;
}
__device__ __forceinline__ float sah_area(const float lo[3], const float hi[3]) {
    float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
    return 2.0f * (ex * ey + ey * ez + ez * ex);
}
__device__ __forceinline__ float axis_of(float4 v, int ax) { return ax == 0 ? v.x : (ax == 1 ? v.y : v.z); }

/* bins of a segment spread over a warp: set A = (axis lane>>4, bin lane&15), set B = (axis 2, bin lane&15; lanes < 16) */
struct LaneBins {
    unsigned cnt[2];
    float lo[2][3], hi[2][3];
    __device__ __forceinline__ void reset() {
#pragma unroll
        for(int s = 0; s < 2; s++) {
            cnt[s] = 0;
#pragma unroll
            for(int k = 0; k < 3; k++) lo[s][k] = SAH_BIG, hi[s][k] = -SAH_BIG;
        }
    }
};
struct SegFrame { /* centroid bounds of a segment and what follows from them */
    float lo[3], scale[3];
    bool use[3];
    __device__ __forceinline__ void set(const float cl[3], const float ch[3]) {
#pragma unroll
        for(int k = 0; k < 3; k++) {
            const float ext = ch[k] - cl[k];
            use[k] = ext > 0.0f;
            lo[k] = cl[k];
            scale[k] = use[k] ? (float)SAH_BINS / ext : 0.0f;
        }
    }
    /* the three bins of an item, one byte each; 0xff = axis not used */
    __device__ __forceinline__ unsigned bins(float4 l, float4 h) const {
        unsigned p = 0;
#pragma unroll
        for(int k = 0; k < 3; k++) {
            const float c = (axis_of(l, k) + axis_of(h, k)) * 0.5f;
            p |= (use[k] ? (unsigned)sah_bin(c, lo[k], scale[k]) : 0xffu) << (8 * k);
        }
        return p;
    }
};

/* fold the items of the lanes in `group` into the lane-owned bins (every lane of the warp calls this) */
__device__ __forceinline__ void bins_add(LaneBins& B, unsigned group, float4 l, float4 h, unsigned packed, int lane) {
    const unsigned sh = 8u * (unsigned)(lane >> 4), mine = (unsigned)(lane & 15);
    for(unsigned m = group; m; m &= m - 1) {
        const int j = __ffs((int)m) - 1;
        const float lx = __shfl_sync(FULL, l.x, j), ly = __shfl_sync(FULL, l.y, j), lz = __shfl_sync(FULL, l.z, j);
        const float hx = __shfl_sync(FULL, h.x, j), hy = __shfl_sync(FULL, h.y, j), hz = __shfl_sync(FULL, h.z, j);
        const unsigned pk = __shfl_sync(FULL, packed, j);
        if(((pk >> sh) & 0xffu) == mine) {
            B.cnt[0]++;
            B.lo[0][0] = fminf(B.lo[0][0], lx), B.lo[0][1] = fminf(B.lo[0][1], ly), B.lo[0][2] = fminf(B.lo[0][2], lz);
            B.hi[0][0] = fmaxf(B.hi[0][0], hx), B.hi[0][1] = fmaxf(B.hi[0][1], hy), B.hi[0][2] = fmaxf(B.hi[0][2], hz);
        }
        if(((pk >> 16) & 0xffu) == mine) { /* lanes >= 16 keep a redundant copy */
            B.cnt[1]++;
            B.lo[1][0] = fminf(B.lo[1][0], lx), B.lo[1][1] = fminf(B.lo[1][1], ly), B.lo[1][2] = fminf(B.lo[1][2], lz);
            B.hi[1][0] = fmaxf(B.hi[1][0], hx), B.hi[1][1] = fmaxf(B.hi[1][1], hy), B.hi[1][2] = fmaxf(B.hi[1][2], hz);
        }
    }
}


/* One step of 32 items (lane = item) into bins in shared memory (phase A: the tile's bins, shared by its warps; small
 * segments: the warp's own).  Per axis: when every item of the step falls into the same bin — the usual case near the top
 * of a spatially coherent mesh — the warp reduces the boxes with six full-mask redux.sync and one lane updates the bin;
 * otherwise every lane updates its bin with shared-memory atomics (conflicts only between lanes that share a bin).  About
 * 1-3 warp instructions per item instead of the 17-28 of the register-transposed loop, which remains for <= 32-item
 * subtrees.  packed: the item's three bins, 0xff = none (lane beyond the segment, axis not used). */
__device__ __forceinline__ void bins_step(int* sb, unsigned packed, float4 l, float4 h, int lane) {
    const float lv[3] = {l.x, l.y, l.z}, hv[3] = {h.x, h.y, h.z};
    int el[3], eh[3];
#pragma unroll
    for(int q = 0; q < 3; q++) el[q] = enc(lv[q] == lv[q] ? lv[q] : SAH_BIG), eh[q] = enc(hv[q] == hv[q] ? hv[q] : -SAH_BIG);
#pragma unroll
    for(int ax = 0; ax < 3; ax++) {
        const unsigned b = (packed >> (8 * ax)) & 0xffu;
        const unsigned valid = __ballot_sync(FULL, b != 0xffu);
        if(!valid) continue;
        const unsigned bref = __shfl_sync(FULL, b, __ffs((int)valid) - 1);
        if(__all_sync(FULL, b == 0xffu || b == bref)) {
            int mn[3], mx[3];
#pragma unroll
            for(int q = 0; q < 3; q++) {
                mn[q] = __reduce_min_sync(FULL, b != 0xffu ? el[q] : enc(SAH_BIG));
                mx[q] = __reduce_max_sync(FULL, b != 0xffu ? eh[q] : enc(-SAH_BIG));
            }
            if(lane == 0) {
                int* w = sb + (ax * SAH_BINS + (int)bref) * 7;
                atomicAdd((unsigned*)w, (unsigned)__popc(valid));
#pragma unroll
                for(int q = 0; q < 3; q++) atomicMin(w + 1 + q, mn[q]), atomicMax(w + 4 + q, mx[q]);
            }
        } else if(b != 0xffu) {
            int* w = sb + (ax * SAH_BINS + (int)b) * 7;
            atomicAdd((unsigned*)w, 1u);
#pragma unroll
            for(int q = 0; q < 3; q++) atomicMin(w + 1 + q, el[q]), atomicMax(w + 4 + q, eh[q]);
        }
    }
}
__device__ __forceinline__ void bins_clear(int* sb, int lane) {
    for(int i = lane; i < SAH_BIN_WORDS; i += 32) sb[i] = (i % 7) == 0 ? 0 : ((i % 7) < 4 ? enc(SAH_BIG) : enc(-SAH_BIG));
    __syncwarp();
}
/* shared bins -> the lane-owned layout eval_split works on */
__device__ __forceinline__ void bins_load(LaneBins& B, const int* sb, int lane) {
    __syncwarp();
#pragma unroll
    for(int q = 0; q < 2; q++) {
        const int* w = sb + ((q == 0 ? (lane >> 4) : 2) * SAH_BINS + (lane & 15)) * 7;
        B.cnt[q] = (unsigned)w[0];
#pragma unroll
        for(int c = 0; c < 3; c++) B.lo[q][c] = dec(w[1 + c]), B.hi[q][c] = dec(w[4 + c]);
    }
}

/* The split of a segment from its bins (sah_split.h: lowest cost, ties to the lower axis, then the lower boundary;
 * candidates need both sides non-empty and a cost below 3.0e38).  Returns axis (-1: none), boundary k and the left count,
 * the same in every lane. */
__device__ __forceinline__ void eval_split(const LaneBins& B, int lane, int& axis, int& k, unsigned& nl) {
    const int b = lane & 15;
    float best_cost = 0.0f;
    int best_idx = 0x7fffffff;
    unsigned nl_set[2];
#pragma unroll
    for(int s = 0; s < 2; s++) {
        /* inclusive prefix (bins 0..b) and suffix (bins b..15) unions inside each half-warp */
        unsigned pc = B.cnt[s], sc = B.cnt[s];
        float pl[3], ph[3], sl[3], shh[3];
#pragma unroll
        for(int q = 0; q < 3; q++) pl[q] = sl[q] = B.lo[s][q], ph[q] = shh[q] = B.hi[s][q];
#pragma unroll
        for(int d = 1; d < 16; d <<= 1) {
            unsigned c2 = __shfl_up_sync(FULL, pc, d, 16);
            float l2[3], h2[3];
#pragma unroll
            for(int q = 0; q < 3; q++) l2[q] = __shfl_up_sync(FULL, pl[q], d, 16), h2[q] = __shfl_up_sync(FULL, ph[q], d, 16);
            if(b >= d) {
                pc += c2;
#pragma unroll
                for(int q = 0; q < 3; q++) pl[q] = fminf(pl[q], l2[q]), ph[q] = fmaxf(ph[q], h2[q]);
            }
            c2 = __shfl_down_sync(FULL, sc, d, 16);
#pragma unroll
            for(int q = 0; q < 3; q++) l2[q] = __shfl_down_sync(FULL, sl[q], d, 16), h2[q] = __shfl_down_sync(FULL, shh[q], d, 16);
            if(b + d < 16) {
                sc += c2;
#pragma unroll
                for(int q = 0; q < 3; q++) sl[q] = fminf(sl[q], l2[q]), shh[q] = fmaxf(shh[q], h2[q]);
            }
        }
        /* lane b evaluates the boundary k = b + 1: left = bins 0..b (its prefix), right = bins b+1..15 (the next lane's suffix) */
        const unsigned rc = __shfl_down_sync(FULL, sc, 1, 16);
        float rl[3], rh[3];
#pragma unroll
        for(int q = 0; q < 3; q++) rl[q] = __shfl_down_sync(FULL, sl[q], 1, 16), rh[q] = __shfl_down_sync(FULL, shh[q], 1, 16);
        nl_set[s] = pc;
        const int ax = s == 0 ? (lane >> 4) : 2;
        if(b < 15 && pc != 0 && rc != 0 && (s == 0 || lane < 16)) {
            const float cost = sah_area(pl, ph) * (float)pc + sah_area(rl, rh) * (float)rc;
            const int idx = ax * 16 + b + 1;
            if(cost < SAH_BIG && (best_idx == 0x7fffffff || cost < best_cost || (cost == best_cost && idx < best_idx)))
                best_cost = cost, best_idx = idx;
        }
    }
#pragma unroll
    for(int d = 16; d >= 1; d >>= 1) {
        const float c2 = __shfl_xor_sync(FULL, best_cost, d);
        const int i2 = __shfl_xor_sync(FULL, best_idx, d);
        if(i2 != 0x7fffffff && (best_idx == 0x7fffffff || c2 < best_cost || (c2 == best_cost && i2 < best_idx)))
            best_cost = c2, best_idx = i2;
    }
    if(best_idx == 0x7fffffff) {
        axis = -1, k = 0, nl = 0;
        return;
    }
    axis = best_idx >> 4, k = best_idx & 15;
    const int src = (axis < 2 ? axis * 16 : 0) + (k - 1);
    const unsigned n0 = __shfl_sync(FULL, nl_set[0], src), n1 = __shfl_sync(FULL, nl_set[1], src);
    nl = axis < 2 ? n0 : n1;
}

/* ---- big segments ------------------------------------------------------------------------------------------------- */
/* root: either the first big segment (with empty bins / bounds) or the first small job */
__global__ void k_sah_setup(unsigned n, unsigned small_max, SahState* st, SahSeg* segs, SahJob* jobs, unsigned* ready, int* cb, int* bins,
                            unsigned* tiles_done) {
    const bool big_root = n > small_max;
    for(int i = threadIdx.x; i < SAH_BIN_WORDS; i += blockDim.x) bins[i] = (i % 7) == 0 ? 0 : ((i % 7) < 4 ? enc(SAH_BIG) : enc(-SAH_BIG));
    if(threadIdx.x < 6) cb[threadIdx.x] = threadIdx.x < 3 ? enc(SAH_BIG) : enc(-SAH_BIG);
    if(threadIdx.x == 0) {
        tiles_done[0] = 0;
        st->n_seg = big_root ? 1u : 0u, st->n_seg_next = 0u, st->small_total = big_root ? 0u : n;
        st->q_tail = big_root ? 0u : 1u, st->q_head = 0u, st->leaves_done = 0u, st->finished = 0u, st->stalled = 0u;
        st->prof[0] = st->prof[1] = st->prof[2] = st->prof[3] = 0ull;
        if(big_root) segs[0] = SahSeg{0u, n, 0, -1};
        else jobs[0] = SahJob{0u, n, 0, -1, 0u}, ready[0] = 1u;
    }
}

__global__ void __launch_bounds__(256) k_sah_init(const float4* __restrict__ tri_lo, const float4* __restrict__ tri_hi, unsigned n,
                                                   SahRec* rec, int* seg_of, int seg_value, uint64_t* keys, int* cb) {
    __shared__ int s_e[6];
    /* identities and clamps are those of sah_split.h's empty_box(): +-3.0e38 */
    int e[6] = {enc(SAH_BIG), enc(SAH_BIG), enc(SAH_BIG), enc(-SAH_BIG), enc(-SAH_BIG), enc(-SAH_BIG)};
    if(threadIdx.x < 6) s_e[threadIdx.x] = e[threadIdx.x];
    __syncthreads();
    for(unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 l = tri_lo[i], h = tri_hi[i];
        l.w = __uint_as_float(i);
        rec[i].lo = l, rec[i].hi = h;
        seg_of[i] = seg_value;
        keys[i] = i; /* the "key" of a primitive in this build is its position in the SAH order */
        const float c[3] = {(l.x + h.x) * 0.5f, (l.y + h.y) * 0.5f, (l.z + h.z) * 0.5f};
#pragma unroll
        for(int q = 0; q < 3; q++)
            if(c[q] == c[q]) e[q] = min(e[q], enc(fminf(c[q], SAH_BIG))), e[3 + q] = max(e[3 + q], enc(fmaxf(c[q], -SAH_BIG)));
    }
    if(seg_value < 0) return;
    /* centroid bounds of the root: warp, block, then one global update per block */
#pragma unroll
    for(int q = 0; q < 3; q++) {
        const int mn = __reduce_min_sync(FULL, e[q]), mx = __reduce_max_sync(FULL, e[3 + q]);
        if((threadIdx.x & 31) == 0) atomicMin(s_e + q, mn), atomicMax(s_e + 3 + q, mx);
    }
    __syncthreads();
    if(threadIdx.x < 3) atomicMin(cb + threadIdx.x, s_e[threadIdx.x]);
    else if(threadIdx.x < 6) atomicMax(cb + threadIdx.x, s_e[threadIdx.x]);
}

__global__ void __launch_bounds__(256) k_sah_bins(const SahRec* __restrict__ rec, const int* __restrict__ seg_of, unsigned n,
                                                   const SahSeg* __restrict__ segs, const int* __restrict__ cb, int* bins,
                                                   unsigned* tiles_done, SahSplit* split) {
    __shared__ int s_slot[2];
    __shared__ int s_bins[2][SAH_BIN_WORDS]; /* the tile's share of the (at most two) segments it touches */
    const unsigned tile0 = blockIdx.x * SAH_TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if(threadIdx.x < 2) s_slot[threadIdx.x] = -1;
    for(int i = threadIdx.x; i < 2 * SAH_BIN_WORDS; i += blockDim.x)
        (&s_bins[0][0])[i] = (i % 7) == 0 ? 0 : ((i % 7) < 4 ? enc(SAH_BIG) : enc(-SAH_BIG));
    __syncthreads();
    LaneBins B;
    SegFrame F;
    int cur = -1, cur_slot = 0;
    for(int it = 0; it < 4; it++) {
        const unsigned pos = tile0 + warp * 128 + it * 32 + lane;
        const int s = pos < n ? seg_of[pos] : -1;
        float4 l = make_float4(0, 0, 0, 0), h = l;
        if(s >= 0) {
            l = rec[pos].lo, h = rec[pos].hi;
            const unsigned a = segs[s].a;
            if(pos == tile0) s_slot[0] = s;                 /* the segment that covers the tile's first position */
            else if(pos == a) s_slot[1] = s;                /* a segment that starts inside the tile (at most one is big) */
        }
        unsigned todo = __ballot_sync(FULL, s >= 0);
        while(todo) { /* at most two big segments meet in a warp's 32 positions */
            const int sj = __shfl_sync(FULL, s, __ffs((int)todo) - 1);
            const unsigned group = __ballot_sync(FULL, s == sj);
            if(sj != cur) {
                cur = sj, cur_slot = segs[sj].a > tile0 ? 1 : 0;
                float cl[3], ch[3];
#pragma unroll
                for(int q = 0; q < 3; q++) cl[q] = dec(cb[6 * sj + q]), ch[q] = dec(cb[6 * sj + 3 + q]);
                F.set(cl, ch);
            }
            bins_step(s_bins[cur_slot], s == sj ? F.bins(l, h) : 0xffffffffu, l, h, lane);
            todo &= ~group;
        }
    }
    __syncthreads();
    for(int i = threadIdx.x; i < 2 * SAH_BIN_WORDS; i += blockDim.x) { /* one global update per tile, segment and non-empty bin */
        const int slot = i / SAH_BIN_WORDS, w = i % SAH_BIN_WORDS, sg = s_slot[slot];
        if(sg < 0 || s_bins[slot][w - w % 7] == 0) continue;
        int* g = bins + (size_t)sg * SAH_BIN_WORDS + w;
        if(w % 7 == 0) atomicAdd((unsigned*)g, (unsigned)s_bins[slot][w]);
        else if(w % 7 < 4) atomicMin(g, s_bins[slot][w]);
        else atomicMax(g, s_bins[slot][w]);
    }
    /* the last tile of a segment to get here evaluates its split (warp 0: slot 0, warp 1: slot 1) */
    __threadfence();
    __syncthreads();
    if(warp >= 2) return;
    const int s = s_slot[warp];
    if(s < 0) return;
    const SahSeg sg = segs[s];
    const unsigned ntiles = (sg.b - 1) / SAH_TILE - sg.a / SAH_TILE + 1;
    unsigned old = 0;
    if(lane == 0) old = atomicAdd(tiles_done + s, 1u);
    old = __shfl_sync(FULL, old, 0);
    if(old != ntiles - 1) return;
    __threadfence();
    {
        const int* base = bins + (size_t)s * SAH_BIN_WORDS;
#pragma unroll
        for(int q = 0; q < 2; q++) {
            const int* w = base + ((q == 0 ? (lane >> 4) : 2) * SAH_BINS + (lane & 15)) * 7;
            B.cnt[q] = (unsigned)__ldcg(w);
#pragma unroll
            for(int c = 0; c < 3; c++) B.lo[q][c] = dec(__ldcg(w + 1 + c)), B.hi[q][c] = dec(__ldcg(w + 4 + c));
        }
    }
    int axis, k;
    unsigned nl;
    eval_split(B, lane, axis, k, nl);
    if(lane == 0) {
        float cl[3], ch[3];
#pragma unroll
        for(int q = 0; q < 3; q++) cl[q] = dec(cb[6 * s + q]), ch[q] = dec(cb[6 * s + 3 + q]);
        F.set(cl, ch);
        SahSplit sp;
        sp.a = sg.a, sp.b = sg.b, sp.axis = axis, sp.k = k;
        sp.lo = axis >= 0 ? F.lo[axis] : 0.0f, sp.scale = axis >= 0 ? F.scale[axis] : 0.0f;
        sp.nl = axis >= 0 ? nl : (sg.b - sg.a) / 2, sp.base = 0, sp.child[0] = sp.child[1] = -1;
        split[s] = sp;
    }
}

__device__ __forceinline__ bool goes_left(const SahSplit& sp, float4 l, float4 h, unsigned pos) {
    if(sp.axis < 0) return pos < sp.a + (sp.b - sp.a) / 2;
    const float c = (axis_of(l, sp.axis) + axis_of(h, sp.axis)) * 0.5f;
    return sah_bin(c, sp.lo, sp.scale) < sp.k;
}

/* per thread 4 consecutive positions; returns the left flags (bit r) and slots (bit 4 + r) of its items.  The splits of
 * the (at most two) big segments that meet in the tile are staged in shared memory first (block-wide: contains barriers);
 * s_sp[slot] is left valid for the caller. */
__device__ __forceinline__ unsigned tile_flags(const SahRec* __restrict__ rec, const int* __restrict__ seg_of, unsigned n,
                                               const SahSplit* __restrict__ split, unsigned tile0, unsigned p0, int seg[4],
                                               SahSplit* s_sp, int* s_seg) {
    if(threadIdx.x == 0) s_seg[0] = tile0 < n ? seg_of[tile0] : -1, s_seg[1] = -1;
#pragma unroll
    for(int r = 0; r < 4; r++) seg[r] = p0 + r < n ? seg_of[p0 + r] : -1;
    __syncthreads();
    const int s0 = s_seg[0];
#pragma unroll
    for(int r = 0; r < 4; r++)
        if(seg[r] >= 0 && seg[r] != s0) s_seg[1] = seg[r]; /* every writer stores the same value */
    __syncthreads();
    if(threadIdx.x < 2 * (sizeof(SahSplit) / 4)) {
        const int slot = threadIdx.x / (sizeof(SahSplit) / 4), w = threadIdx.x % (sizeof(SahSplit) / 4);
        if(s_seg[slot] >= 0) ((unsigned*)(s_sp + slot))[w] = ((const unsigned*)(split + s_seg[slot]))[w];
    }
    __syncthreads();
    unsigned f = 0;
#pragma unroll
    for(int r = 0; r < 4; r++) {
        if(seg[r] < 0) continue;
        const unsigned pos = p0 + r, slot = seg[r] == s0 ? 0u : 1u;
        if(goes_left(s_sp[slot], rec[pos].lo, rec[pos].hi, pos)) f |= 1u << r;
        f |= (16u * slot) << r;
    }
    return f;
}

__global__ void __launch_bounds__(256) k_sah_count(const SahRec* __restrict__ rec, const int* __restrict__ seg_of, unsigned n,
                                                    const SahSplit* __restrict__ split, unsigned* tile_left) {
    __shared__ unsigned s_sum[8];
    __shared__ SahSplit s_sp[2];
    __shared__ int s_seg[2];
    const unsigned tile0 = blockIdx.x * SAH_TILE;
    int seg[4];
    const unsigned f = tile_flags(rec, seg_of, n, split, tile0, tile0 + 4 * threadIdx.x, seg, s_sp, s_seg);
    unsigned c = 0; /* slot 0 lefts in the low half, slot 1 lefts in the high half */
#pragma unroll
    for(int r = 0; r < 4; r++)
        if(f & (1u << r)) c += (f & (16u << r)) ? 0x10000u : 1u;
    c = __reduce_add_sync(FULL, c);
    if((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = c;
    __syncthreads();
    if(threadIdx.x == 0) {
        unsigned t = 0;
        for(int w = 0; w < 8; w++) t += s_sum[w];
        tile_left[2 * blockIdx.x] = t & 0xffffu, tile_left[2 * blockIdx.x + 1] = t >> 16;
    }
}

/* one block: scan of the tile counts (in place, exclusive, total at [2 * ntiles]) and the bookkeeping of the level */
__global__ void __launch_bounds__(1024) k_sah_level(unsigned* tile_left, unsigned n_entries, const SahSeg* __restrict__ segs,
                                                     SahSeg* segs_next, SahSplit* split, SahState* st, int* cb, int* bins,
                                                     unsigned* tiles_done, SahJob* jobs, int* ready, unsigned dst_parity,
                                                     int* left, int* right, int* parent, int* range_first, int* range_last,
                                                     unsigned ni, unsigned* host_flag, unsigned small_max) {
    __shared__ unsigned s_w[32];
    __shared__ unsigned s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if(threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for(unsigned base = 0; base < n_entries; base += 1024) {
        const unsigned i = base + threadIdx.x;
        const unsigned v = i < n_entries ? tile_left[i] : 0;
        unsigned inc = v;
#pragma unroll
        for(int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(FULL, inc, d);
            if(lane >= d) inc += t;
        }
        if(lane == 31) s_w[warp] = inc;
        __syncthreads();
        if(warp == 0) {
            unsigned w = s_w[lane], winc = w;
#pragma unroll
            for(int d = 1; d < 32; d <<= 1) {
                const unsigned t = __shfl_up_sync(FULL, winc, d);
                if(lane >= d) winc += t;
            }
            s_w[lane] = winc - w;
        }
        __syncthreads();
        const unsigned carry = s_carry;
        if(i < n_entries) tile_left[i] = carry + s_w[warp] + inc - v;
        __syncthreads();
        if(threadIdx.x == 1023) s_carry = carry + s_w[31] + inc;
        __syncthreads();
    }
    if(threadIdx.x == 0) tile_left[n_entries] = s_carry;
    __syncthreads();
    const unsigned n_seg = st->n_seg;
    for(unsigned s = threadIdx.x; s < n_seg; s += 1024) {
        const SahSeg sg = segs[s];
        SahSplit sp = split[s];
        const unsigned first = 2 * (sg.a / SAH_TILE) + ((sg.a % SAH_TILE) ? 1u : 0u);
        const unsigned last = 2 * ((sg.b - 1) / SAH_TILE);
        const unsigned nl = tile_left[last + 1] - tile_left[first];
        sp.nl = nl, sp.base = tile_left[first];
        const unsigned m = sg.b - sg.a, mid = sg.a + nl;
        const unsigned ca[2] = {sg.a, mid}, cbnd[2] = {mid, sg.b};
        const int cme[2] = {sg.me + 1, sg.me + (int)nl};
        int ref[2];
        for(int side = 0; side < 2; side++) {
            const unsigned c = cbnd[side] - ca[side];
            if(c == 1) {
                ref[side] = ~(int)ca[side];
                parent[(size_t)ni + ca[side]] = sg.me;
                sp.child[side] = -2;
            } else if(c <= small_max) {
                ref[side] = cme[side];
                const unsigned idx = atomicAdd(&st->q_tail, 1u);
                jobs[idx] = SahJob{ca[side], cbnd[side], cme[side], sg.me, dst_parity};
                ready[idx] = 1;
                atomicAdd(&st->small_total, c);
                sp.child[side] = -1;
            } else {
                ref[side] = cme[side];
                const unsigned idx = atomicAdd(&st->n_seg_next, 1u);
                segs_next[idx] = SahSeg{ca[side], cbnd[side], cme[side], sg.me};
                sp.child[side] = (int)idx;
            }
        }
        (void)m;
        left[sg.me] = ref[0], right[sg.me] = ref[1], parent[sg.me] = sg.parent;
        range_first[sg.me] = (int)sg.a, range_last[sg.me] = (int)sg.b - 1;
        split[s] = sp;
    }
    __syncthreads();
    /* working state of the next level's big segments (this level's bins / bounds / arrival counters are dead by now) */
    const unsigned nb = st->n_seg_next;
    for(unsigned i = threadIdx.x; i < nb * (unsigned)SAH_BIN_WORDS; i += 1024) {
        const unsigned w = i % 7;
        bins[i] = w == 0 ? 0 : (w < 4 ? enc(SAH_BIG) : enc(-SAH_BIG));
    }
    for(unsigned i = threadIdx.x; i < nb * 6; i += 1024) cb[i] = (i % 6) < 3 ? enc(SAH_BIG) : enc(-SAH_BIG);
    for(unsigned i = threadIdx.x; i < nb; i += 1024) tiles_done[i] = 0;
    __syncthreads();
    if(threadIdx.x == 0) {
        st->n_seg = nb, st->n_seg_next = 0;
        *host_flag = nb;
    }
}

__global__ void __launch_bounds__(256) k_sah_scatter(const SahRec* __restrict__ rec, const int* __restrict__ seg_of, unsigned n,
                                                      const SahSplit* __restrict__ split, const unsigned* __restrict__ scan,
                                                      SahRec* rec_out, int* seg_out, int* cb, uint32_t* order) {
    __shared__ unsigned s_w[8];
    __shared__ int s_cb[4][6], s_child[4]; /* centroid bounds of the (slot, side) children that are big segments */
    const unsigned tile0 = blockIdx.x * SAH_TILE, p0 = tile0 + 4 * threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if(threadIdx.x < 24) s_cb[threadIdx.x / 6][threadIdx.x % 6] = (threadIdx.x % 6) < 3 ? enc(SAH_BIG) : enc(-SAH_BIG);
    if(threadIdx.x < 4) s_child[threadIdx.x] = -1;
    __shared__ SahSplit s_sp[2];
    __shared__ int s_seg[2];
    int seg[4];
    const unsigned f = tile_flags(rec, seg_of, n, split, tile0, p0, seg, s_sp, s_seg);
    unsigned c = 0;
#pragma unroll
    for(int r = 0; r < 4; r++)
        if(f & (1u << r)) c += (f & (16u << r)) ? 0x10000u : 1u;
    unsigned inc = c;
#pragma unroll
    for(int d = 1; d < 32; d <<= 1) {
        const unsigned t = __shfl_up_sync(FULL, inc, d);
        if(lane >= d) inc += t;
    }
    if(lane == 31) s_w[warp] = inc;
    __syncthreads();
    unsigned before = inc - c; /* lefts of both slots in this tile before this thread's items */
    for(int w = 0; w < warp; w++) before += s_w[w];
#pragma unroll
    for(int r = 0; r < 4; r++) {
        const unsigned pos = p0 + r;
        const bool valid = pos < n;
        const int s = seg[r];
        int child = -1, cidx = 0;
        unsigned dest = pos;
        float4 l = make_float4(0, 0, 0, 0), h = l;
        if(valid && s >= 0) {
            const unsigned slot = (f >> (4 + r)) & 1u;
            const SahSplit& sp = s_sp[slot];
            const bool is_left = (f >> r) & 1u;
            const unsigned in_tile = slot ? (before >> 16) : (before & 0xffffu);
            const unsigned lefts_before = scan[2 * blockIdx.x + slot] - sp.base + in_tile;
            dest = is_left ? sp.a + lefts_before : sp.a + sp.nl + (pos - sp.a) - lefts_before;
            child = sp.child[is_left ? 0 : 1], cidx = (int)slot * 2 + (is_left ? 0 : 1);
            l = rec[pos].lo, h = rec[pos].hi;
            rec_out[dest].lo = l, rec_out[dest].hi = h;
            if(is_left) before += slot ? 0x10000u : 1u;
            if(child == -2) order[dest] = __float_as_uint(l.w);
        }
        if(valid) seg_out[dest] = child >= 0 ? child : -1;
        /* centroid bounds of big children: reduced per warp and child (every lane takes part in the match), merged per
         * tile in shared memory, one global update per tile and child */
        const unsigned grp = __match_any_sync(FULL, child);
        if(child >= 0) {
            const float cen[3] = {(l.x + h.x) * 0.5f, (l.y + h.y) * 0.5f, (l.z + h.z) * 0.5f};
#pragma unroll
            for(int q = 0; q < 3; q++) {
                const bool ok = cen[q] == cen[q];
                const int mn = __reduce_min_sync(grp, enc(ok ? fminf(cen[q], SAH_BIG) : SAH_BIG));
                const int mx = __reduce_max_sync(grp, enc(ok ? fmaxf(cen[q], -SAH_BIG) : -SAH_BIG));
                if(lane == __ffs((int)grp) - 1) atomicMin(&s_cb[cidx][q], mn), atomicMax(&s_cb[cidx][3 + q], mx);
            }
            if(lane == __ffs((int)grp) - 1) s_child[cidx] = child;
        }
    }
    __syncthreads();
    if(threadIdx.x < 24 && s_child[threadIdx.x / 6] >= 0) {
        int* g = cb + 6 * s_child[threadIdx.x / 6] + threadIdx.x % 6;
        if(threadIdx.x % 6 < 3) atomicMin(g, s_cb[threadIdx.x / 6][threadIdx.x % 6]);
        else atomicMax(g, s_cb[threadIdx.x / 6][threadIdx.x % 6]);
    }
}

/* ---- small segments: persistent warps over a ticket queue ---------------------------------------------------------- */
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

/* leaves finished: the warp that completes the last one raises `finished` for the waiting warps */
__device__ __forceinline__ void leaves_add(SahState* st, unsigned k, unsigned total) {
    __threadfence();
    if(atomicAdd(&st->leaves_done, k) + k >= total) st_release(&st->finished, 1u);
}

struct SmallOut {
    uint32_t* order;
    int *left, *right, *parent, *range_first, *range_last;
    unsigned ni;
};

/* A subtree of at most 32 items, entirely in registers: lane i holds the item at position a + i.  Nodes are taken from
 * a warp-private stack; a partition permutes the lanes through shared memory; only the node records and, at the end,
 * the primitive ids of the final lane order go to global memory. */
__device__ __forceinline__ void subtree32(unsigned a, unsigned m, int me0, int par0, float4 l, float4 h, float4* stage, uint4* stack,
                                          const SmallOut& O, int lane) {
    const unsigned lt = (1u << lane) - 1u;
    int sp = 0;
    if(lane == 0) stack[0] = make_uint4(0u, m, (unsigned)me0, (unsigned)par0);
    sp = 1;
    __syncwarp();
    while(sp) {
        const uint4 nd = stack[--sp];
        __syncwarp();
        const unsigned off = nd.x, cnt = nd.y;
        const int me = (int)nd.z, par = (int)nd.w;
        const bool in = (unsigned)lane - off < cnt;
        if(cnt == 2) {
            /* Two items.  With finite centroids the definition reduces to: the first axis on which the centroids differ, the
             * boundary k = 1 (bins 0 and >= 1), cost = area(first) + area(second); a cost of 3.0e38 or more, or no such axis,
             * means the middle split — which is the same node, possibly without the swap. */
            const float c[3] = {(l.x + h.x) * 0.5f, (l.y + h.y) * 0.5f, (l.z + h.z) * 0.5f};
            float o[3];
#pragma unroll
            for(int q = 0; q < 3; q++) o[q] = __shfl_sync(FULL, c[q], (int)off + 1 - (lane - (int)off)); /* the partner's */
            bool fin = true;
            int ax = -1;
#pragma unroll
            for(int q = 2; q >= 0; q--) {
                fin = fin && fabsf(c[q]) < SAH_BIG && fabsf(o[q]) < SAH_BIG && fabsf(c[q] - o[q]) < SAH_BIG;
                if(c[q] != o[q]) ax = q;
            }
            const float lo3[3] = {l.x, l.y, l.z}, hi3[3] = {h.x, h.y, h.z};
            const float my_area = sah_area(lo3, hi3);
            const float a0 = __shfl_sync(FULL, my_area, (int)off), a1 = __shfl_sync(FULL, my_area, (int)off + 1);
            fin = __shfl_sync(FULL, fin, (int)off) && __shfl_sync(FULL, fin, (int)off + 1);
            ax = __shfl_sync(FULL, ax, (int)off);
            if(fin) {
                /* first item (lane off) goes right iff a valid candidate exists and its centroid is the larger one */
                const float c_first = __shfl_sync(FULL, ax >= 0 ? c[ax] : 0.0f, (int)off);
                const float c_second = __shfl_sync(FULL, ax >= 0 ? c[ax] : 0.0f, (int)off + 1);
                const bool first_is_low = c_first < c_second;
                const float cost = first_is_low ? a0 * 1.0f + a1 * 1.0f : a1 * 1.0f + a0 * 1.0f;
                const bool swap = ax >= 0 && cost < SAH_BIG && !first_is_low;
                if(swap) {
                    const int partner = (int)off + 1 - (lane - (int)off);
                    const float4 l2 = make_float4(__shfl_sync(FULL, l.x, partner & 31), __shfl_sync(FULL, l.y, partner & 31),
                                                  __shfl_sync(FULL, l.z, partner & 31), __shfl_sync(FULL, l.w, partner & 31));
                    const float4 h2 = make_float4(__shfl_sync(FULL, h.x, partner & 31), __shfl_sync(FULL, h.y, partner & 31),
                                                  __shfl_sync(FULL, h.z, partner & 31), __shfl_sync(FULL, h.w, partner & 31));
                    if(in) l = l2, h = h2;
                }
                if(lane == 0) {
                    const unsigned p0 = a + off;
                    O.left[me] = ~(int)p0, O.right[me] = ~(int)(p0 + 1), O.parent[me] = par;
                    O.range_first[me] = (int)p0, O.range_last[me] = (int)p0 + 1;
                    O.parent[(size_t)O.ni + p0] = me, O.parent[(size_t)O.ni + p0 + 1] = me;
                }
                continue;
            }
        }
        float cl[3], ch[3];
        {
            const float c[3] = {(l.x + h.x) * 0.5f, (l.y + h.y) * 0.5f, (l.z + h.z) * 0.5f};
#pragma unroll
            for(int q = 0; q < 3; q++) cl[q] = in ? fminf(SAH_BIG, c[q]) : SAH_BIG, ch[q] = in ? fmaxf(-SAH_BIG, c[q]) : -SAH_BIG;
        }
#pragma unroll
        for(int d = 16; d >= 1; d >>= 1)
#pragma unroll
            for(int q = 0; q < 3; q++)
                cl[q] = fminf(cl[q], __shfl_xor_sync(FULL, cl[q], d)), ch[q] = fmaxf(ch[q], __shfl_xor_sync(FULL, ch[q], d));
        SegFrame F;
        F.set(cl, ch);
        LaneBins B;
        B.reset();
        const unsigned group = __ballot_sync(FULL, in);
        bins_add(B, group, l, h, in ? F.bins(l, h) : 0xffffffffu, lane);
        int axis, k;
        unsigned nl;
        eval_split(B, lane, axis, k, nl);
        if(axis < 0) nl = cnt / 2;
        const unsigned nr = cnt - nl;
        bool is_left = false;
        if(in) is_left = axis < 0 ? (unsigned)lane < off + nl : sah_bin((axis_of(l, axis) + axis_of(h, axis)) * 0.5f, F.lo[axis], F.scale[axis]) < k;
        const unsigned ml = __ballot_sync(FULL, in && is_left), mr = __ballot_sync(FULL, in && !is_left);
        if(in) {
            const unsigned dest = is_left ? off + __popc(ml & lt) : off + nl + __popc(mr & lt);
            stage[2 * dest] = l, stage[2 * dest + 1] = h;
        }
        __syncwarp();
        if(in) l = stage[2 * lane], h = stage[2 * lane + 1];
        if(lane == 0) {
            const unsigned p0 = a + off;
            O.left[me] = nl == 1 ? ~(int)p0 : me + 1;
            O.right[me] = nr == 1 ? ~(int)(p0 + nl) : me + (int)nl;
            O.parent[me] = par;
            O.range_first[me] = (int)p0, O.range_last[me] = (int)(p0 + cnt) - 1;
            if(nl == 1) O.parent[(size_t)O.ni + p0] = me;
            if(nr == 1) O.parent[(size_t)O.ni + p0 + nl] = me;
            if(nr >= 2) stack[sp] = make_uint4(off + nl, nr, (unsigned)(me + (int)nl), (unsigned)me);
            if(nl >= 2) stack[sp + (nr >= 2 ? 1 : 0)] = make_uint4(off, nl, (unsigned)(me + 1), (unsigned)me);
        }
        sp += (nr >= 2 ? 1 : 0) + (nl >= 2 ? 1 : 0);
        __syncwarp();
    }
    if((unsigned)lane < m) O.order[a + lane] = __float_as_uint(l.w);
}

#ifndef GPURT_SAH_SMALL_MINB
#define GPURT_SAH_SMALL_MINB 0 /* minimum CTAs per SM asked of k_sah_small (A/B builds, tools/ab_build.sh): 72 registers as compiled;
                                  8 / 10 / 12 CTAs (64 / 48 / 40 registers, spills) make the stand-in build 1.40 -> 1.43 / 1.53 / 1.60 ms */
#endif
__global__ void __launch_bounds__(128, GPURT_SAH_SMALL_MINB) k_sah_small(SahRec* rec0, SahRec* rec1, SahJob* jobs, unsigned* ready, SahState* st,
                                                    unsigned cap, SmallOut O) {
    __shared__ float4 s_stage[4][64];
    __shared__ uint4 s_stack[4][32];
    __shared__ int s_bins[4][SAH_BIN_WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* stage = s_stage[warp];
    uint4* stack = s_stack[warp];
    int* sb = s_bins[warp];
    const unsigned lt = (1u << lane) - 1u;
    const unsigned total = st->small_total;
    for(;;) {
        unsigned ticket = 0;
        if(lane == 0) ticket = atomicAdd(&st->q_head, 1u);
        ticket = __shfl_sync(FULL, ticket, 0);
#ifdef SAH_PROFILE
        const long long tw0 = clock64();
#endif
        int got = 0;
        if(lane == 0) {
            /* a ticket is served as soon as some warp queues its job; the warp that finishes the last leaf raises
             * `finished`.  A warp gives up after ~2 s without either (cannot happen unless a job was lost: the host then
             * reports the build as failed instead of hanging the device) */
            unsigned ns = 32;
            for(unsigned spins = 0;; spins++) {
                if(ticket < cap && ld_acquire(ready + ticket)) {
                    got = 1;
                    break;
                }
                if((spins & 3u) == 3u && (ld_acquire(&st->finished) || ld_acquire(&st->stalled))) break;
                if(spins > (1u << 21)) {
                    atomicExch(&st->stalled, 1u);
                    break;
                }
                __nanosleep(ns);
                if(ns < 1024) ns += ns;
            }
        }
#ifdef SAH_PROFILE
        if(lane == 0 && got) atomicAdd(&st->prof[2], (unsigned long long)(clock64() - tw0));
#endif
        got = __shfl_sync(FULL, got, 0);
        if(!got) return;
        __syncwarp();
        const unsigned* jw = (const unsigned*)(jobs + ticket);
        unsigned a = __ldcg(jw), b = __ldcg(jw + 1), src = __ldcg(jw + 4);
        int me = (int)__ldcg(jw + 2), par = (int)__ldcg(jw + 3);
        bool have_bounds = false; /* a child continued by this warp inherits the bounds its parent's partition reduced */
        float cl[3], ch[3];
        for(;;) { /* one inner node per iteration: [a, b), at least two items */
            const SahRec* in = src ? rec1 : rec0;
            SahRec* out = src ? rec0 : rec1;
            const unsigned m = b - a;
            if(m <= 32) {
                float4 l = make_float4(0, 0, 0, 0), h = l;
                if((unsigned)lane < m) l = __ldcg(&in[a + lane].lo), h = __ldcg(&in[a + lane].hi);
#ifdef SAH_PROFILE
                const long long t0 = clock64();
#endif
                subtree32(a, m, me, par, l, h, stage, stack, O, lane);
#ifdef SAH_PROFILE
                if(lane == 0) atomicAdd(&st->prof[0], (unsigned long long)(clock64() - t0)), atomicAdd(&st->prof[3], 1ull);
#endif
                if(lane == 0) leaves_add(st, m, total);
                break;
            }
#ifdef SAH_PROFILE
            const long long tn0 = clock64();
#endif
            if(!have_bounds) {
#pragma unroll
                for(int q = 0; q < 3; q++) cl[q] = SAH_BIG, ch[q] = -SAH_BIG;
                for(unsigned i = a + lane; i < b; i += 32) {
                    const float4 l = __ldcg(&in[i].lo), h = __ldcg(&in[i].hi);
                    const float c[3] = {(l.x + h.x) * 0.5f, (l.y + h.y) * 0.5f, (l.z + h.z) * 0.5f};
#pragma unroll
                    for(int q = 0; q < 3; q++) cl[q] = fminf(cl[q], c[q]), ch[q] = fmaxf(ch[q], c[q]);
                }
#pragma unroll
                for(int d = 16; d >= 1; d >>= 1)
#pragma unroll
                    for(int q = 0; q < 3; q++)
                        cl[q] = fminf(cl[q], __shfl_xor_sync(FULL, cl[q], d)), ch[q] = fmaxf(ch[q], __shfl_xor_sync(FULL, ch[q], d));
            }
            SegFrame F;
            F.set(cl, ch);
            bins_clear(sb, lane);
            {
                /* the next step's records are in flight while this step's 32 items are binned */
                unsigned i = a + lane;
                float4 l = make_float4(0, 0, 0, 0), h = l;
                if(i < b) l = __ldcg(&in[i].lo), h = __ldcg(&in[i].hi);
                for(unsigned base = a; base < b; base += 32) {
                    const bool valid = base + lane < b;
                    const float4 l0 = l, h0 = h;
                    i = base + 32 + lane;
                    if(i < b) l = __ldcg(&in[i].lo), h = __ldcg(&in[i].hi);
                    bins_step(sb, valid ? F.bins(l0, h0) : 0xffffffffu, l0, h0, lane);
                }
            }
            LaneBins B;
            bins_load(B, sb, lane);
            int axis, k;
            unsigned nl;
            eval_split(B, lane, axis, k, nl);
            if(axis < 0) nl = m / 2;
            const float s_lo = axis >= 0 ? F.lo[axis] : 0.0f, s_scale = axis >= 0 ? F.scale[axis] : 0.0f;
            const unsigned nr = m - nl;
            unsigned run_l = 0, run_r = 0;
            int eb[2][6]; /* centroid bounds of the two children, reduced while their items pass by */
#pragma unroll
            for(int q = 0; q < 6; q++) eb[0][q] = eb[1][q] = q < 3 ? enc(SAH_BIG) : enc(-SAH_BIG);
            {
                unsigned i = a + lane;
                float4 l = make_float4(0, 0, 0, 0), h = l;
                if(i < b) l = __ldcg(&in[i].lo), h = __ldcg(&in[i].hi);
                for(unsigned base = a; base < b; base += 32) {
                    const unsigned pos = base + lane;
                    const bool valid = pos < b;
                    const float4 l0 = l, h0 = h;
                    i = base + 32 + lane;
                    if(i < b) l = __ldcg(&in[i].lo), h = __ldcg(&in[i].hi);
                    const float c[3] = {(l0.x + h0.x) * 0.5f, (l0.y + h0.y) * 0.5f, (l0.z + h0.z) * 0.5f};
                    bool is_left = false;
                    if(valid) is_left = axis < 0 ? pos < a + nl : sah_bin(axis == 0 ? c[0] : (axis == 1 ? c[1] : c[2]), s_lo, s_scale) < k;
                    const unsigned ml = __ballot_sync(FULL, valid && is_left), mr = __ballot_sync(FULL, valid && !is_left);
                    if(valid) {
                        const unsigned dest = is_left ? a + run_l + __popc(ml & lt) : a + nl + run_r + __popc(mr & lt);
                        out[dest].lo = l0, out[dest].hi = h0;
                        if((is_left && nl == 1) || (!is_left && nr == 1)) O.order[dest] = __float_as_uint(l0.w);
                    }
                    run_l += __popc(ml), run_r += __popc(mr);
#pragma unroll
                    for(int q = 0; q < 3; q++) {
                        const bool ok = valid && c[q] == c[q];
                        const int lo_e = enc(ok ? fminf(c[q], SAH_BIG) : SAH_BIG), hi_e = enc(ok ? fmaxf(c[q], -SAH_BIG) : -SAH_BIG);
                        eb[0][q] = min(eb[0][q], __reduce_min_sync(FULL, is_left ? lo_e : enc(SAH_BIG)));
                        eb[0][3 + q] = max(eb[0][3 + q], __reduce_max_sync(FULL, is_left ? hi_e : enc(-SAH_BIG)));
                        eb[1][q] = min(eb[1][q], __reduce_min_sync(FULL, valid && !is_left ? lo_e : enc(SAH_BIG)));
                        eb[1][3 + q] = max(eb[1][3 + q], __reduce_max_sync(FULL, valid && !is_left ? hi_e : enc(-SAH_BIG)));
                    }
                }
            }
            if(lane == 0) {
                O.left[me] = nl == 1 ? ~(int)a : me + 1;
                O.right[me] = nr == 1 ? ~(int)(a + nl) : me + (int)nl;
                O.parent[me] = par;
                O.range_first[me] = (int)a, O.range_last[me] = (int)b - 1;
                if(nl == 1) O.parent[(size_t)O.ni + a] = me;
                if(nr == 1) O.parent[(size_t)O.ni + a + nl] = me;
            }
            const unsigned leaves = (nl == 1 ? 1u : 0u) + (nr == 1 ? 1u : 0u);
            __threadfence();
            __syncwarp(); /* the children read what the lanes of this warp just wrote */
            /* two children to build: the smaller one goes to the queue (thousands of warps are waiting for tickets; finishing
             * a <= 32-item subtree here would put ~40 us of single-warp work on this chain), the larger one stays */
            const unsigned done_now = leaves;
            bool go_l = nl >= 2, go_r = nr >= 2;
            if(go_l && go_r) {
                const bool keep_left = nl >= nr;
                if(lane == 0) {
                    const unsigned idx = atomicAdd(&st->q_tail, 1u);
                    jobs[idx] = keep_left ? SahJob{a + nl, b, me + (int)nl, me, src ^ 1u} : SahJob{a, a + nl, me + 1, me, src ^ 1u};
                    __threadfence();
                    st_release(ready + idx, 1u);
                }
                go_l = keep_left, go_r = !keep_left;
            }
            if(done_now && lane == 0) leaves_add(st, done_now, total);
#ifdef SAH_PROFILE
            if(lane == 0) atomicAdd(&st->prof[1], (unsigned long long)(clock64() - tn0));
#endif
            if(!go_l && !go_r) break;
            const int side = go_l ? 0 : 1;
#pragma unroll
            for(int q = 0; q < 3; q++) cl[q] = dec(eb[side][q]), ch[q] = dec(eb[side][3 + q]);
            have_bounds = true;
            if(go_l) par = me, b = a + nl, me = me + 1, src ^= 1u;
            else par = me, a = a + nl, me = me + (int)nl, src ^= 1u;
        }
    }
}

} // namespace

/* bytes of temporaries build_sah_split_device() needs for n triangles */
size_t sah_split_tmp_bytes(size_t n, int sm_count) {
    const size_t ntiles = (n + SAH_TILE - 1) / SAH_TILE, max_big = n / SAH_TILE + 2, cap = n + (size_t)sm_count * 64 + 64;
    return 2 * n * sizeof(SahRec) + 2 * n * 4 + cap * (sizeof(SahJob) + 4) + (2 * ntiles + 2) * 4 +
           max_big * (2 * sizeof(SahSeg) + sizeof(SahSplit) + 24 + SAH_BIN_WORDS * 4 + 4) + sizeof(SahState) + 16 * 256;
}

/* order / keys / left / right / parent / range_first / range_last of the SAH-split tree over tri_lo / tri_hi (n >= 2).
 * `tmp` = sah_split_tmp_bytes(n) bytes of device memory.  Synchronises the stream once per level of big segments. */
int build_sah_split_device(gpurt_ctx* ctx, const float4* tri_lo, const float4* tri_hi, unsigned n, uint32_t* order, uint64_t* keys,
                           int* left, int* right, int* parent, int* range_first, int* range_last, void* tmp, size_t tmp_bytes,
                           unsigned* levels_out) {
    cudaStream_t st = ctx->stream;
    const unsigned ni = n - 1;
    const unsigned ntiles = (n + SAH_TILE - 1) / SAH_TILE, max_big = n / SAH_TILE + 2;
    int occ = 0;
    GPURT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sah_small, 128, 0));
    occ = std::max(1, std::min(occ, 16));
    const unsigned grid_small = (unsigned)ctx->sm_count * (unsigned)occ;
    const size_t cap = (size_t)n + (size_t)grid_small * 4 + 64;
    char* p = (char*)tmp;
    auto take = [&](size_t bytes) {
        char* q = p;
        p += (bytes + 255) & ~(size_t)255;
        return (void*)q;
    };
    SahRec* rec[2] = {(SahRec*)take((size_t)n * sizeof(SahRec)), (SahRec*)take((size_t)n * sizeof(SahRec))};
    int* seg_of[2] = {(int*)take((size_t)n * 4), (int*)take((size_t)n * 4)};
    SahJob* jobs = (SahJob*)take(cap * sizeof(SahJob));
    unsigned* ready = (unsigned*)take(cap * 4);
    unsigned* tile_left = (unsigned*)take(((size_t)2 * ntiles + 2) * 4);
    SahSeg* segs[2] = {(SahSeg*)take(max_big * sizeof(SahSeg)), (SahSeg*)take(max_big * sizeof(SahSeg))};
    SahSplit* split = (SahSplit*)take(max_big * sizeof(SahSplit));
    int* cb = (int*)take(max_big * 24);
    int* bins = (int*)take(max_big * SAH_BIN_WORDS * 4);
    unsigned* tiles_done = (unsigned*)take(max_big * 4);
    SahState* state = (SahState*)take(sizeof(SahState));
    if((size_t)(p - (char*)tmp) > tmp_bytes) return set_error("SAH build: temporary buffer too small"), GPURT_E_STATE;

    if(!ctx->pinned_word) GPURT_CUDA(cudaHostAlloc((void**)&ctx->pinned_word, 1024, cudaHostAllocMapped));
    volatile unsigned* h_flag = ctx->pinned_word; /* next level's big-segment count, written by k_sah_level */
    unsigned* d_flag = nullptr;
    GPURT_CUDA(cudaHostGetDevicePointer((void**)&d_flag, ctx->pinned_word, 0));

    unsigned small_max = SAH_SMALL_MAX; /* GPURT_SAH_SMALL_MAX: A/B knob, the tree does not depend on it */
    if(const char* e = getenv("GPURT_SAH_SMALL_MAX")) small_max = std::max<unsigned>(SAH_TILE, (unsigned)atoi(e));
    const bool big_root = n > small_max;
    GPURT_CUDA(cudaMemsetAsync(ready, 0, cap * 4, st));
    GPURT_CUDA(cudaMemsetAsync(state, 0, sizeof(SahState), st)); /* the padding words travel to the host with the rest */
    k_sah_setup<<<1, 128, 0, st>>>(n, small_max, state, segs[0], jobs, ready, cb, bins, tiles_done);
    k_sah_init<<<std::min((n + 255) / 256, (unsigned)ctx->sm_count * 8u), 256, 0, st>>>(tri_lo, tri_hi, n, rec[0], seg_of[0],
                                                                                   big_root ? 0 : -1, keys, cb);
    /* Levels of big segments.  The host learns from k_sah_level whether another level follows; so that this read-back does
     * not leave the device idle, level L + 1 is already queued when the host waits for level L's count (a level without
     * big segments touches nothing: its kernels find seg_of == -1 everywhere).  Word 0 / 1 of the pinned block alternate. */
    unsigned par = 0, levels = 0;
    if(big_root) {
        cudaEvent_t ev[2] = {ctx->ev_copy, ctx->ev_kernel}; /* free between host-buffer calls */
        auto launch_level = [&](unsigned which) {
            k_sah_bins<<<ntiles, 256, 0, st>>>(rec[par], seg_of[par], n, segs[par], cb, bins, tiles_done, split);
            k_sah_count<<<ntiles, 256, 0, st>>>(rec[par], seg_of[par], n, split, tile_left);
            k_sah_level<<<1, 1024, 0, st>>>(tile_left, 2 * ntiles, segs[par], segs[par ^ 1], split, state, cb, bins, tiles_done, jobs,
                                            (int*)ready, par ^ 1u, left, right, parent, range_first, range_last, ni, d_flag + which,
                                            small_max);
            cudaEventRecord(ev[which], st);
            k_sah_scatter<<<ntiles, 256, 0, st>>>(rec[par], seg_of[par], n, split, tile_left, rec[par ^ 1], seg_of[par ^ 1], cb, order);
            par ^= 1;
            levels++;
        };
        launch_level(0);
        for(unsigned which = 0;; which ^= 1) {
            if(levels > 512) return set_error("SAH build: more than 512 levels of big segments"), GPURT_E_STATE;
            launch_level(which ^ 1);                       /* speculative */
            GPURT_CUDA(cudaEventSynchronize(ev[which]));
            const unsigned n_big = h_flag[which];
            if(n_big > max_big) return set_error("SAH build: segment bound exceeded"), GPURT_E_STATE;
            if(n_big == 0) break;                          /* the level just queued is the empty one */
        }
    }
    k_sah_small<<<grid_small, 128, 0, st>>>(rec[0], rec[1], jobs, ready, state, (unsigned)cap,
                                            SmallOut{order, left, right, parent, range_first, range_last, ni});
    GPURT_CUDA(cudaGetLastError());
    GPURT_CUDA(cudaMemcpyAsync(ctx->pinned_word + 16, state, sizeof(SahState), cudaMemcpyDeviceToHost, st));
    GPURT_CUDA(cudaStreamSynchronize(st));
    const SahState* hs = (const SahState*)(ctx->pinned_word + 16);
    if(hs->stalled || hs->leaves_done != hs->small_total)
        return set_error("SAH build: the small-segment queue stalled (" + std::to_string(hs->leaves_done) + " of " +
                         std::to_string(hs->small_total) + " leaves)"), GPURT_E_STATE;
#ifdef SAH_PROFILE
    fprintf(stderr, "[sah profile] n %u: <=32 subtrees %llu (%.1f Mcycles), larger nodes %.1f Mcycles, ticket waits of served warps %.1f Mcycles\n", n,
            hs->prof[3], hs->prof[0] * 1e-6, hs->prof[1] * 1e-6, hs->prof[2] * 1e-6);
#endif
    if(levels_out) *levels_out = levels;
    return GPURT_OK;
}

} // namespace gpurt
