/*
 * traverse.cuh — per-ray / per-query traversal cores of the 8-wide BVH, written as host+device
 * code: the kernels in trace.cu / cpq.cu / render.cu call them per thread; tests/emu replays the
 * very same code on the CPU (debug aid only — never a product path).
 */
#pragma once
#include "bvh8.cuh"

#if defined(__CUDA_ARCH__)
#define GPURT_LDG(p) __ldg(p)
#define GPURT_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#else
#define GPURT_LDG(p) (*(p))
#define GPURT_ATOMIC_ADD(p, v) (*(p) += (v))
#endif

namespace gpurt {

GPURT_HD unsigned gpurt_clz(unsigned x) {
#if defined(__CUDA_ARCH__)
    return (unsigned)__clz((int)x);
#else
    return x ? (unsigned)__builtin_clz(x) : 32u;
#endif
}
GPURT_HD unsigned gpurt_ctz(unsigned x) {
#if defined(__CUDA_ARCH__)
    return (unsigned)__ffs((int)x) - 1u;
#else
    return (unsigned)__builtin_ctz(x);
#endif
}
GPURT_HD unsigned gpurt_popc(unsigned x) {
#if defined(__CUDA_ARCH__)
    return (unsigned)__popc(x);
#else
    return (unsigned)__builtin_popcount(x);
#endif
}

constexpr int kStack = 64; /* uint2 entries; wide depth is checked against this at build time */

struct HitRec {
    float t, u, v;
    unsigned gid;
};

template <bool ANY, bool STATS>
GPURT_HD bool traverse8(const float4* __restrict__ nodes,
                                          const float4* __restrict__ tris, F3 o, F3 d, float tmin,
                                          float tmax, HitRec& best, unsigned long long* counters) {
    RaySetup rs = make_ray_setup(o, d, tmin);
    uint2 stack[kStack];
    int sp = 0;
    uint2 ng;
    ng.x = 0u, ng.y = 0x80000000u;
    best.t = tmax, best.u = 0.0f, best.v = 0.0f, best.gid = kNoHit;
    unsigned n_nodes = 0, n_tris = 0;
    for(;;) {
        if(ng.y <= 0x00ffffffu) {
            if(sp == 0) break;
            ng = stack[--sp];
        }
        /* pop the nearest hit child of the current node group */
        unsigned hits = ng.y;
        unsigned bit = 31u - gpurt_clz(hits);
        ng.y &= ~(1u << bit);
        if(ng.y > 0x00ffffffu) stack[sp++] = ng;
        unsigned slot = (bit - 24u) ^ rs.octinv;
        unsigned rel = gpurt_popc(hits & 0xffu & ((1u << slot) - 1u));
        const float4* np = nodes + (size_t)(ng.x + rel) * kNodeVec4;
        Node8 node;
#pragma unroll
        for(int k = 0; k < 5; k++) node.v[k] = GPURT_LDG(np + k);
        if(STATS) n_nodes++;
        unsigned hit8 = node_hits8(node, rs, best.t);
        unsigned imask = f2u(node.v[0].w) >> 24;
        ng.x = f2u(node.v[1].x);
        ng.y = (octant_permute8(hit8 & imask, rs.octinv) << 24) | imask;
        /* hit leaf slots: meta byte = unary(count) << 5 | first triangle offset */
        unsigned leaf = hit8 & ~imask;
        const float4* tp = tris + (size_t)f2u(node.v[1].y) * kTriVec4;
        unsigned m_lo = f2u(node.v[1].z), m_hi = f2u(node.v[1].w);
        while(leaf) {
            unsigned slot = gpurt_ctz(leaf);
            leaf &= leaf - 1u;
            unsigned meta = ((slot & 4u ? m_hi : m_lo) >> (8u * (slot & 3u))) & 0xffu;
            unsigned k = meta & 31u, kend = k + gpurt_popc(meta >> 5);
            for(; k < kend; k++) {
                float4 r0 = GPURT_LDG(tp + 3 * k), r1 = GPURT_LDG(tp + 3 * k + 1), r2 = GPURT_LDG(tp + 3 * k + 2);
                if(STATS) n_tris++;
                float t, u, v;
                if(intersect_tri(o, d, tmin, tmax, f3(r0.x, r0.y, r0.z), f3(r1.x, r1.y, r1.z),
                                 f3(r2.x, r2.y, r2.z), t, u, v)) {
                    if(ANY) return true;
                    unsigned gid = f2u(r0.w);
                    if(t < best.t || (t == best.t && gid < best.gid)) best.t = t, best.u = u, best.v = v, best.gid = gid;
                }
            }
        }
    }
    if(STATS) {
        GPURT_ATOMIC_ADD(counters + 0, (unsigned long long)n_nodes);
        GPURT_ATOMIC_ADD(counters + 1, (unsigned long long)n_tris);
        if(best.gid != kNoHit) GPURT_ATOMIC_ADD(counters + 2, 1ull);
    }
    return best.gid != kNoHit;
}


/* ---- closest-point descent (cpq.cu) ----------------------------------------------------------- */
constexpr unsigned kLeafBit = 0x80000000u;

struct CpRec {
    float d2, v, w;
    unsigned gid, idx; /* idx: position in the wide triangle array */
};

/* Nearest child first, the rest pushed with their squared box distance and re-checked against the
 * shrinking radius when popped.  At most 7 pushes per level: STACK >= 7*depth+1. */
template <int STACK>
GPURT_HD void closest_point8(const float4* __restrict__ nodes, const float4* __restrict__ tris, F3 p,
                             float r2, CpRec& best, unsigned* visit_counts = nullptr) {
    best.d2 = r2, best.v = 0.0f, best.w = 0.0f, best.gid = kNoHit, best.idx = 0;
#if defined(__CUDA_ARCH__)
    const unsigned one = c_one_bits; /* see byte_as_unit_float */
#else
    const unsigned one = 0x3f800000u;
#endif
    uint2 stack[STACK];
    int sp = 0;
    unsigned cur = 0;      /* entry in hand: node index, or kLeafBit | first triangle << 2 | count */
    float cur_d2 = 0.0f;
    bool have = true;
    /* One iteration = (pop if nothing is in hand) + (expand the node in hand) + (test the leaf in hand).  A lane
     * whose nearest child is a leaf does node and leaf work in the same iteration, so a warp with a mix of
     * node and leaf entries keeps more lanes busy per pass (ncu: 9 of 32 lanes with one entry per iteration). */
    for(;;) {
        if(!have) {
            while(sp > 0) { /* entries pushed earlier may have fallen outside the shrinking radius */
                uint2 e = stack[--sp];
                if(u2f(e.y) <= best.d2) {
                    cur = e.x, cur_d2 = u2f(e.y), have = true;
                    break;
                }
            }
            if(!have) break;
        }
        if(!(cur & kLeafBit)) {
            have = false;
            if(cur_d2 <= best.d2) {
                const float4* np = nodes + (size_t)cur * kNodeVec4;
                Node8 node;
#pragma unroll
                for(int k = 0; k < 5; k++) node.v[k] = GPURT_LDG(np + k);
                if(visit_counts) visit_counts[0]++;
                unsigned imask = f2u(node.v[0].w) >> 24;
                unsigned child_base = f2u(node.v[1].x), tri_base = f2u(node.v[1].y);
                unsigned m_lo = f2u(node.v[1].z), m_hi = f2u(node.v[1].w);
                /* Nearest child continues immediately; the others are pushed (unsorted) with their box distance.
                 * A full distance sort (19-comparator network) was measured: same visit counts on surface-near
                 * queries (17.5 nodes / 21 triangles per query either way) and 10 % slower, so it is not used. */
                unsigned near_ref = 0;
                float near_d2 = 0.0f;
                bool near_ok = false;
                ChildDist cd = make_child_dist(node, p, one);
#pragma unroll
                for(int s = 0; s < 8; s++) {
                    unsigned meta = byte_of(m_lo, m_hi, s);
                    if(meta == 0) continue;
                    float d2 = child_dist2(node, cd, s, one);
                    if(d2 > best.d2) continue;
                    unsigned ref;
                    if((imask >> s) & 1u) ref = child_base + gpurt_popc(imask & ((1u << s) - 1u));
                    else ref = kLeafBit | ((tri_base + (meta & 31u)) << 2) | gpurt_popc(meta >> 5);
                    uint2 e;
                    if(!near_ok) {
                        near_ref = ref, near_d2 = d2, near_ok = true;
                        continue;
                    }
                    if(d2 < near_d2) {
                        e.x = near_ref, e.y = f2u(near_d2);
                        near_ref = ref, near_d2 = d2;
                    } else
                        e.x = ref, e.y = f2u(d2);
                    stack[sp++] = e;
                }
                if(near_ok) cur = near_ref, cur_d2 = near_d2, have = true;
            }
        }
        if(have && (cur & kLeafBit)) {
            have = false;
            if(cur_d2 <= best.d2) {
                unsigned first = (cur & ~kLeafBit) >> 2, count = cur & 3u;
                if(visit_counts) visit_counts[1] += count;
                for(unsigned k = 0; k < count; k++) {
                    const float4* tp = tris + (size_t)(first + k) * kTriVec4;
                    float4 r0 = GPURT_LDG(tp), r1 = GPURT_LDG(tp + 1), r2v = GPURT_LDG(tp + 2);
                    float v, w;
                    float d2 = closest_point_tri(p, f3(r0.x, r0.y, r0.z), f3(r1.x, r1.y, r1.z),
                                                 f3(r2v.x, r2v.y, r2v.z), v, w);
                    unsigned gid = f2u(r0.w);
                    if(d2 < best.d2 || (d2 == best.d2 && gid < best.gid))
                        best.d2 = d2, best.v = v, best.w = w, best.gid = gid, best.idx = first + k;
                }
            }
        }
    }
}

} // namespace gpurt
