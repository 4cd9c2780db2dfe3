/*
 * bvh8.cuh — the compressed 8-wide BVH node: layout, collapse from the binary LBVH, child tests.
 *
 * Replaces the opaque NVIDIA acceleration structure built at src/vk/vulkan.cpp:877 and walked by
 * traceRayEXT (src/shaders/rt/rt.rgen:258).  Layout follows the compressed-wide-BVH idea
 * (Ylitie, Karras, Laine 2017): 80 bytes = 5 x 16-byte loads per 8 children.
 *
 *   vec0: p.x p.y p.z | ex ey ez imask          quantisation origin, per-axis exponents, inner mask
 *   vec1: child_base | tri_base | meta[0..3] | meta[4..7]
 *   vec2: qlo_x[0..7] | qlo_y[0..7]             8-bit child boxes on the grid p + q * 2^(e-127)
 *   vec3: qlo_z[0..7] | qhi_x[0..7]
 *   vec4: qhi_y[0..7] | qhi_z[0..7]
 *
 *   meta[i] == 0            empty slot
 *   inner child  (imask i)  meta = 0x20 | (24 + i); the child node is child_base + popc(imask & ((1<<i)-1))
 *   leaf child              meta = unary(count) << 5 | first triangle offset from tri_base (count<=3,
 *                           at most 24 triangles per node, stored contiguously in wide order)
 *
 * Slot i encodes the octant of the child relative to the node centre (bit0:+x bit1:+y bit2:+z), so a
 * ray visits hit children in front-to-back order by popping the highest bit of
 * (24 + (i ^ octinv)) without sorting distances.
 */
#pragma once
#include "common.cuh"

namespace gpurt {

struct Node8 {
    float4 v[5];
};

/* children produced by the collapse, before encoding */
constexpr int kEmptyChild = (int)0x80000000;
/* >= 0: BVH2 internal node that becomes an inner child; < 0 (and != empty): leaf range */
GPURT_HD int encode_leaf_range(unsigned first, unsigned count) { return ~(int)((first << 2) | count); }
GPURT_HD void decode_leaf_range(int c, unsigned& first, unsigned& count) {
    unsigned u = (unsigned)~c;
    first = u >> 2;
    count = u & 3u;
}

/* Read-only view of the binary LBVH (Karras 2012) the collapse consumes. */
struct Bvh2View {
    const int* left;        /* >=0 internal, <0 leaf ~sorted_pos */
    const int* right;
    const int* range_first; /* per internal node: first sorted position covered */
    const int* range_last;
    const float4* node_lo;  /* exact boxes of internal nodes */
    const float4* node_hi;
    const float4* tri_lo;   /* exact boxes of triangles, gid order */
    const float4* tri_hi;
    const unsigned* order;  /* sorted position -> gid */
    float inflate;
    /* SAH-optimal collapse decisions (dp_node), 8 bytes per internal node; nullptr = greedy collapse */
    const unsigned char* dp_dec = nullptr;
    float dp_cprim = 0.3f; /* cost of a triangle test relative to a node visit */
};

GPURT_HD Box3 bvh2_child_box(const Bvh2View& B, int c) {
    Box3 b;
    if(c >= 0) {
        float4 lo = B.node_lo[c], hi = B.node_hi[c];
        b.lo = f3(lo.x, lo.y, lo.z), b.hi = f3(hi.x, hi.y, hi.z);
    } else {
        unsigned g = B.order[~c];
        float4 lo = B.tri_lo[g], hi = B.tri_hi[g];
        b.lo = f3(lo.x, lo.y, lo.z), b.hi = f3(hi.x, hi.y, hi.z);
    }
    return b;
}
GPURT_HD int bvh2_count(const Bvh2View& B, int c) {
    return c < 0 ? 1 : B.range_last[c] - B.range_first[c] + 1;
}
/* leaf-like: at most kMaxLeafTris triangles, contiguous in sorted order */
GPURT_HD int bvh2_to_child(const Bvh2View& B, int c) {
    if(c < 0) return encode_leaf_range((unsigned)~c, 1);
    int n = bvh2_count(B, c);
    if(n <= kMaxLeafTris) return encode_leaf_range((unsigned)B.range_first[c], (unsigned)n);
    return c;
}

/* Assign the chosen children to octant slots (greedy on the signed centroid projection) and convert them to
 * collapse references.  cand[i] is a raw BVH2 ref, box[i] its box, as_leaf bit i = the subtree becomes a leaf slot.
 * out_child[8] is indexed by slot.  Returns the number of inner children; n_leaf_tris gets the triangles referenced
 * by leaf slots. */
GPURT_HD int collapse_assign(const Bvh2View& B, const int cand[8], const Box3 box[8], unsigned as_leaf, int n,
                             int out_child[8], int& n_leaf_tris) {
    /* node centre from the union of candidate boxes */
    Box3 nb = box[0];
    F3 cen[8];
    for(int i = 0; i < n; i++) {
        const Box3& b = box[i];
        cen[i] = (b.lo + b.hi) * 0.5f;
        nb.lo = f3(fminf(nb.lo.x, b.lo.x), fminf(nb.lo.y, b.lo.y), fminf(nb.lo.z, b.lo.z));
        nb.hi = f3(fmaxf(nb.hi.x, b.hi.x), fmaxf(nb.hi.y, b.hi.y), fmaxf(nb.hi.z, b.hi.z));
    }
    F3 nc = (nb.lo + nb.hi) * 0.5f;
    for(int s = 0; s < 8; s++) out_child[s] = kEmptyChild;
#ifndef GPURT_ASSIGN_VARIANT
#define GPURT_ASSIGN_VARIANT 0
#endif
#if GPURT_ASSIGN_VARIANT == 0
    unsigned used_child = 0, used_slot = 0;
    for(int round = 0; round < n; round++) {
        float best = -3.0e38f;
        int bc = -1, bs = -1;
        for(int i = 0; i < n; i++) {
            if(used_child & (1u << i)) continue;
            F3 dlt = cen[i] - nc;
            for(int s = 0; s < 8; s++) {
                if(used_slot & (1u << s)) continue;
                float cost = ((s & 1) ? dlt.x : -dlt.x) + ((s & 2) ? dlt.y : -dlt.y) +
                             ((s & 4) ? dlt.z : -dlt.z);
                if(cost > best) best = cost, bc = i, bs = s;
            }
        }
        used_child |= 1u << bc;
        used_slot |= 1u << bs;
        int c = cand[bc];
        if(c < 0) out_child[bs] = encode_leaf_range((unsigned)~c, 1);
        else if((as_leaf >> bc) & 1u) out_child[bs] = encode_leaf_range((unsigned)B.range_first[c], (unsigned)bvh2_count(B, c));
        else out_child[bs] = c;
    }
#else
    /* experiment (tools/assign_probe.py): the assignment that maximises the SUM of the projections, by dynamic programming
     * over slot subsets (child i takes one of the slots not used by children 0..i-1); variant 2 / 3 divide the offsets by the
     * node's extent first */
    float cost[8][8];
    F3 ext = nb.hi - nb.lo;
    for(int i = 0; i < n; i++) {
        F3 dlt = cen[i] - nc;
#if GPURT_ASSIGN_VARIANT >= 2
        dlt = F3{ext.x > 0 ? dlt.x / ext.x : 0.0f, ext.y > 0 ? dlt.y / ext.y : 0.0f, ext.z > 0 ? dlt.z / ext.z : 0.0f};
#endif
        for(int s = 0; s < 8; s++) cost[i][s] = ((s & 1) ? dlt.x : -dlt.x) + ((s & 2) ? dlt.y : -dlt.y) + ((s & 4) ? dlt.z : -dlt.z);
    }
    (void)ext;
    int slot_of[8];
#if GPURT_ASSIGN_VARIANT == 2
    { /* greedy as variant 0, on the normalised offsets */
        unsigned used_child = 0, used_slot = 0;
        for(int round = 0; round < n; round++) {
            float best = -3.0e38f;
            int bc = -1, bs = -1;
            for(int i = 0; i < n; i++) {
                if(used_child & (1u << i)) continue;
                for(int s = 0; s < 8; s++)
                    if(!(used_slot & (1u << s)) && cost[i][s] > best) best = cost[i][s], bc = i, bs = s;
            }
            used_child |= 1u << bc, used_slot |= 1u << bs;
            slot_of[bc] = bs;
        }
    }
#else
    {
        float dp[256];
        signed char choice[8][256];
        for(int m = 0; m < 256; m++) dp[m] = -3.0e38f;
        dp[0] = 0.0f;
        for(int m = 0; m < 256; m++) {
            int i = 0;
            for(int b = m; b; b &= b - 1) i++;
            if(i >= n || dp[m] < -1.0e38f) continue;
            for(int s = 0; s < 8; s++) {
                if(m & (1 << s)) continue;
                float v = dp[m] + cost[i][s];
                int m2 = m | (1 << s);
                if(v > dp[m2]) dp[m2] = v, choice[i][m2] = (signed char)s;
            }
        }
        int bestm = -1;
        float bestv = -3.0e38f;
        for(int m = 0; m < 256; m++) {
            int i = 0;
            for(int b = m; b; b &= b - 1) i++;
            if(i == n && dp[m] > bestv) bestv = dp[m], bestm = m;
        }
        for(int i = n - 1, m = bestm; i >= 0; i--) {
            int sl = choice[i][m];
            slot_of[i] = sl;
            m &= ~(1 << sl);
        }
    }
#endif
    for(int i = 0; i < n; i++) {
        int c = cand[i], bs = slot_of[i];
        if(c < 0) out_child[bs] = encode_leaf_range((unsigned)~c, 1);
        else if((as_leaf >> i) & 1u) out_child[bs] = encode_leaf_range((unsigned)B.range_first[c], (unsigned)bvh2_count(B, c));
        else out_child[bs] = c;
    }
#endif
    int n_inner = 0;
    n_leaf_tris = 0;
    for(int s = 0; s < 8; s++) {
        int c = out_child[s];
        if(c == kEmptyChild) continue;
        if(c >= 0) n_inner++;
        else {
            unsigned f, k;
            decode_leaf_range(c, f, k);
            n_leaf_tris += (int)k;
        }
    }
    return n_inner;
}

/* Greedy surface-area collapse: start from the two children of BVH2 node `root`, repeatedly open
 * the inner candidate with the largest area until 8 children. */
GPURT_HD int collapse_node_greedy(const Bvh2View& B, int root, int out_child[8], int& n_leaf_tris) {
    int cand[8];    /* raw BVH2 refs */
    Box3 box[8];    /* their boxes, loaded once */
    float area[8];  /* surface area, or -1 when the candidate cannot be opened (leaf / small subtree) */
    unsigned as_leaf = 0;
    auto load = [&](int i, int c) {
        cand[i] = c;
        box[i] = bvh2_child_box(B, c);
        bool leaf = c < 0 || bvh2_count(B, c) <= kMaxLeafTris;
        area[i] = leaf ? -1.0f : box_area(box[i]);
        as_leaf = (as_leaf & ~(1u << i)) | ((leaf ? 1u : 0u) << i);
    };
    int n = 2;
    load(0, B.left[root]);
    load(1, B.right[root]);
    while(n < 8) {
        int best = -1;
        float best_area = -1.0f;
        for(int i = 0; i < n; i++)
            if(area[i] > best_area) best_area = area[i], best = i;
        if(best < 0) break;
        int c = cand[best];
        load(best, B.left[c]);
        load(n++, B.right[c]);
    }
    return collapse_assign(B, cand, box, as_leaf, n, out_child, n_leaf_tris);
}

/* ---- SAH-optimal collapse (dynamic programming over the binary tree, after Ylitie, Karras, Laine 2017) ---------
 * C(n, i) = least SAH cost of representing the subtree of BVH2 node n by at most i roots that sit in the slots of
 * one wide node:   C(n,1) = min(leaf: A(n) P(n) c_prim if P(n) <= 3,  inner: A(n) c_node + D(n,8)),
 * D(n,j) = min_k C(left,k) + C(right,j-k),   C(n,i) = min(D(n,i), C(n,i-1)).   One bottom-up pass (dp_node, called
 * from the refit kernel when both children are final) stores the costs and the 8 decision bytes per node; the
 * collapse then follows the decisions instead of the greedy largest-area rule.  Against the greedy rule on the
 * Sponza stand-in: 32 % fewer wide nodes, 5 % / 2 % / 4 % fewer node visits for primary / bounce / random rays. */
constexpr float kDpCostNode = 1.0f;
#if defined(__CUDA_ARCH__)
#define GPURT_LDCG_F(p) __ldcg(p)
#else
#define GPURT_LDCG_F(p) (*(p))
#endif
GPURT_HD float dp_ref_cost(const Bvh2View& B, const float* cost, int c, int i) { /* C(c, i), i = 1..7 */
    if(c < 0) return box_area(bvh2_child_box(B, c)) * B.dp_cprim;
    return GPURT_LDCG_F(cost + 7ull * c + (i - 1));
}
/* dec[0] = split of D(n,8); dec[1] = 1 leaf / 2 inner; dec[i], i = 2..7: 0 = same as i-1, else the split k */
GPURT_HD void dp_node(const Bvh2View& B, float* cost, unsigned char* dec, int n, const Box3& box) {
    const int L = B.left[n], R = B.right[n];
    float cl[7], cr[7];
    for(int i = 1; i <= 7; i++) cl[i - 1] = dp_ref_cost(B, cost, L, i), cr[i - 1] = dp_ref_cost(B, cost, R, i);
    const float A = box_area(box);
    const int P = bvh2_count(B, n);
    float out[7];
    unsigned char d[8];
    {
        float best = 3.0e38f;
        int kb = 1;
        for(int k = 1; k <= 7; k++) {
            float c = cl[k - 1] + cr[7 - k];
            if(c < best) best = c, kb = k;
        }
        d[0] = (unsigned char)kb;
        float c_int = A * kDpCostNode + best;
        float c_leaf = P <= kMaxLeafTris ? A * (float)P * B.dp_cprim : 3.0e38f;
        out[0] = fminf(c_leaf, c_int);
        d[1] = c_leaf <= c_int ? 1 : 2;
    }
    for(int i = 2; i <= 7; i++) {
        float best = 3.0e38f;
        int kb = 1;
        for(int k = 1; k < i; k++) {
            float c = cl[k - 1] + cr[i - k - 1];
            if(c < best) best = c, kb = k;
        }
        if(best < out[i - 2]) out[i - 1] = best, d[i] = (unsigned char)kb;
        else out[i - 1] = out[i - 2], d[i] = 0;
    }
    for(int i = 0; i < 7; i++) cost[7ull * n + i] = out[i];
    for(int i = 0; i < 8; i++) dec[8ull * n + i] = d[i];
}
GPURT_HD int collapse_node_dp(const Bvh2View& B, int root, int out_child[8], int& n_leaf_tris) {
    int cand[8];
    Box3 box[8];
    unsigned as_leaf = 0;
    int n = 0;
    int sc[16], si[16], sp = 0; /* pending (ref, budget) pairs, left child on top so candidates come out left to right */
    const int k8 = B.dp_dec[8ull * root];
    sc[sp] = B.right[root], si[sp++] = 8 - k8;
    sc[sp] = B.left[root], si[sp++] = k8;
    while(sp > 0 && n < 8) {
        int c = sc[--sp], i = si[sp];
        bool leaf = true;
        if(c >= 0) {
            const unsigned char* d = B.dp_dec + 8ull * c;
            while(i > 1 && d[i] == 0) i--;
            if(i > 1) {
                sc[sp] = B.right[c], si[sp++] = i - d[i];
                sc[sp] = B.left[c], si[sp++] = d[i];
                continue;
            }
            leaf = d[1] == 1;
        }
        cand[n] = c;
        box[n] = bvh2_child_box(B, c);
        as_leaf |= (leaf ? 1u : 0u) << n;
        n++;
    }
    return collapse_assign(B, cand, box, as_leaf, n, out_child, n_leaf_tris);
}
GPURT_HD int collapse_node(const Bvh2View& B, int root, int out_child[8], int& n_leaf_tris) {
    return B.dp_dec ? collapse_node_dp(B, root, out_child, n_leaf_tris) : collapse_node_greedy(B, root, out_child, n_leaf_tris);
}

/* box of a collapsed child (exact), from the BVH2 */
GPURT_HD Box3 collapsed_child_box(const Bvh2View& B, int c) {
    if(c >= 0) return bvh2_child_box(B, c);
    unsigned first, count;
    decode_leaf_range(c, first, count);
    Box3 b = bvh2_child_box(B, ~(int)first);
    for(unsigned k = 1; k < count; k++) {
        Box3 t = bvh2_child_box(B, ~(int)(first + k));
        b.lo = f3(fminf(b.lo.x, t.lo.x), fminf(b.lo.y, t.lo.y), fminf(b.lo.z, t.lo.z));
        b.hi = f3(fmaxf(b.hi.x, t.hi.x), fmaxf(b.hi.y, t.hi.y), fmaxf(b.hi.z, t.hi.z));
    }
    return b;
}

GPURT_HD unsigned pick_exponent(float extent) {
    /* smallest e with extent <= 255 * 2^(e-127); extent >= 0 */
    if(!(extent > 0.0f)) return 1u;
    unsigned e = (f2u(extent) >> 23) & 0xffu; /* 2^(e-127) <= extent < 2^(e-126) */
    e = e > 7u ? e - 7u : 1u;                 /* extent/2^(e-7-127) < 256 */
    while(extent > 255.0f * u2f(e << 23)) e++;
    if(e < 1u) e = 1u;
    if(e > 254u) e = 254u;
    return e;
}

/* Encode one node. child[8] by slot; child boxes are inflated by B.inflate (N7) and quantised
 * outward.  child_base / tri_base are absolute indices. */
GPURT_HD void encode_node(const Bvh2View& B, const int child[8], unsigned child_base,
                          unsigned tri_base, Node8& out) {
    Box3 cb[8];
    Box3 nb;
    nb.lo = f3(3.0e38f, 3.0e38f, 3.0e38f);
    nb.hi = f3(-3.0e38f, -3.0e38f, -3.0e38f);
    float e = B.inflate;
    for(int s = 0; s < 8; s++) {
        if(child[s] == kEmptyChild) continue;
        Box3 b = collapsed_child_box(B, child[s]);
        b.lo = f3(b.lo.x - e, b.lo.y - e, b.lo.z - e);
        b.hi = f3(b.hi.x + e, b.hi.y + e, b.hi.z + e);
        cb[s] = b;
        nb.lo = f3(fminf(nb.lo.x, b.lo.x), fminf(nb.lo.y, b.lo.y), fminf(nb.lo.z, b.lo.z));
        nb.hi = f3(fmaxf(nb.hi.x, b.hi.x), fmaxf(nb.hi.y, b.hi.y), fmaxf(nb.hi.z, b.hi.z));
    }
    /* N7, node-relative part: the conversion-free plane evaluation (byte_as_unit_float) rounds
     * (origin - 2^15*scale) once, an error of at most 2^-24 * 257 * node extent per axis; pad every child
     * by 2^-14 * node extent (4x that bound).  Small nodes get a negligible pad, only the few large ones
     * near the root a visible one. */
    {
        F3 pad = (nb.hi - nb.lo) * 6.103515625e-05f;
        for(int s = 0; s < 8; s++) {
            if(child[s] == kEmptyChild) continue;
            cb[s].lo = cb[s].lo - pad;
            cb[s].hi = cb[s].hi + pad;
        }
        nb.lo = nb.lo - pad;
        nb.hi = nb.hi + pad;
    }
    unsigned ex = pick_exponent(nb.hi.x - nb.lo.x), ey = pick_exponent(nb.hi.y - nb.lo.y),
             ez = pick_exponent(nb.hi.z - nb.lo.z);
    float sx = u2f(ex << 23), sy = u2f(ey << 23), sz = u2f(ez << 23);
    float ix = u2f((254u - ex) << 23), iy = u2f((254u - ey) << 23), iz = u2f((254u - ez) << 23);
    unsigned imask = 0, meta[8], q[6][8];
    unsigned tri_off = 0;
    for(int s = 0; s < 8; s++) {
        meta[s] = 0;
        for(int k = 0; k < 6; k++) q[k][s] = k < 3 ? 255u : 0u;
        int c = child[s];
        if(c == kEmptyChild) continue;
        if(c >= 0) {
            imask |= 1u << s;
            meta[s] = 0x20u | (24u + (unsigned)s);
        } else {
            unsigned first, count;
            decode_leaf_range(c, first, count);
            unsigned unary = count == 1 ? 1u : count == 2 ? 3u : 7u;
            meta[s] = (unary << 5) | tri_off;
            tri_off += count;
        }
        const Box3& b = cb[s];
        float lo[3] = {b.lo.x, b.lo.y, b.lo.z}, hi[3] = {b.hi.x, b.hi.y, b.hi.z};
        float p[3] = {nb.lo.x, nb.lo.y, nb.lo.z}, sc[3] = {sx, sy, sz}, isc[3] = {ix, iy, iz};
        for(int a = 0; a < 3; a++) {
            float fl = floorf((lo[a] - p[a]) * isc[a]);
            float fh = ceilf((hi[a] - p[a]) * isc[a]);
            fl = fminf(fmaxf(fl, 0.0f), 255.0f);
            fh = fminf(fmaxf(fh, 0.0f), 255.0f);
            /* keep the decoded box a superset under fp32 rounding of (x - p) */
            if(fl > 0.0f && p[a] + fl * sc[a] > lo[a]) fl -= 1.0f;
            if(fh < 255.0f && p[a] + fh * sc[a] < hi[a]) fh += 1.0f;
            q[a][s] = (unsigned)fl;
            q[3 + a][s] = (unsigned)fh;
        }
    }
    auto pack4 = [](const unsigned* b) { return b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24); };
    out.v[0].x = nb.lo.x, out.v[0].y = nb.lo.y, out.v[0].z = nb.lo.z;
    out.v[0].w = u2f(ex | (ey << 8) | (ez << 16) | (imask << 24));
    out.v[1].x = u2f(child_base), out.v[1].y = u2f(tri_base);
    out.v[1].z = u2f(pack4(meta)), out.v[1].w = u2f(pack4(meta + 4));
    out.v[2].x = u2f(pack4(q[0])), out.v[2].y = u2f(pack4(q[0] + 4));
    out.v[2].z = u2f(pack4(q[1])), out.v[2].w = u2f(pack4(q[1] + 4));
    out.v[3].x = u2f(pack4(q[2])), out.v[3].y = u2f(pack4(q[2] + 4));
    out.v[3].z = u2f(pack4(q[3])), out.v[3].w = u2f(pack4(q[3] + 4));
    out.v[4].x = u2f(pack4(q[4])), out.v[4].y = u2f(pack4(q[4] + 4));
    out.v[4].z = u2f(pack4(q[5])), out.v[4].w = u2f(pack4(q[5] + 4));
}

/* ---- traversal-side decoding ------------------------------------------------------------------- */
#if defined(__CUDACC__)
static __constant__ unsigned c_one_bits = 0x3f800000u;
#endif

struct RaySetup {
    F3 o, d, idir;
    float tmin;
    unsigned octinv; /* bit set where the ray travels towards + */
    unsigned one;    /* 0x3f800000 held in a register for byte_as_unit_float */
};
GPURT_HD float safe_rcp_dir(float d) {
    float a = fabsf(d) > 1e-20f ? d : copysignf(1e-20f, d);
    return 1.0f / a;
}
GPURT_HD RaySetup make_ray_setup(F3 o, F3 d, float tmin) {
    RaySetup r;
    r.o = o, r.d = d, r.tmin = tmin;
    r.idir = f3(safe_rcp_dir(d.x), safe_rcp_dir(d.y), safe_rcp_dir(d.z));
    r.octinv = (r.idir.x < 0.0f ? 0u : 1u) | (r.idir.y < 0.0f ? 0u : 2u) | (r.idir.z < 0.0f ? 0u : 4u);
#if defined(__CUDA_ARCH__)
    r.one = c_one_bits; /* read from the constant bank so ptxas cannot fold it into PRMT's immediate slot */
#else
    r.one = 0x3f800000u;
#endif
    return r;
}
GPURT_HD unsigned byte_of(unsigned lo4, unsigned hi4, int i) {
    return ((i < 4 ? lo4 : hi4) >> (8 * (i & 3))) & 0xffu;
}

/* 1 + q * 2^-15 for byte `i` (0..3) of `word`, exactly, with one byte-permute: the byte lands in
 * mantissa bits 8..15 of 1.0f.  fma(that, 2^15*s, o - 2^15*s) == o + q*s up to one rounding of
 * (o - 2^15*s), i.e. <= 2^-24 * 257 * node extent in position units — covered by the node-relative
 * N7 padding (2^-14 * node extent, encode_node).  Replaces shift + mask + I2F (quarter-rate). */
GPURT_HD float byte_as_unit_float(unsigned word, unsigned one, int i) {
#if defined(__CUDA_ARCH__)
    /* `one` (0x3f800000) is kept in a register so the selector can be an immediate: one PRMT, no
     * selector materialisation per use */
    unsigned d;
    switch(i) {
    case 0: asm("prmt.b32 %0, %1, %2, 0x7604;" : "=r"(d) : "r"(word), "r"(one)); break;
    case 1: asm("prmt.b32 %0, %1, %2, 0x7614;" : "=r"(d) : "r"(word), "r"(one)); break;
    case 2: asm("prmt.b32 %0, %1, %2, 0x7624;" : "=r"(d) : "r"(word), "r"(one)); break;
    default: asm("prmt.b32 %0, %1, %2, 0x7634;" : "=r"(d) : "r"(word), "r"(one)); break;
    }
    return __uint_as_float(d);
#else
    return u2f(one | (((word >> (8 * i)) & 0xffu) << 8));
#endif
}

/* Test the 8 children of `n` against the ray interval [tmin, tmax]; returns one bit per SLOT
 * (bit i = child i hit).  Empty slots hold an inverted box (lo=255, hi=0) and always fail.
 * Cost per child: 4 PRMT + 2 I2F.U8 conversions, 6 FFMA, 4 FMNMX, 1 FSETP, 1 predicated OR. */
GPURT_HD unsigned node_hits8(const Node8& n, const RaySetup& r, float tmax) {
    unsigned eb = f2u(n.v[0].w);
    /* x,y: scale = 2^(e-127) * 2^15 * idir (exponent byte moved up by 15), PRMT conversion */
    float sx = u2f(((eb & 0xffu) + 15u) << 23) * r.idir.x;
    float sy = u2f((((eb >> 8) & 0xffu) + 15u) << 23) * r.idir.y;
    /* z planes go through I2F.U8 (XU pipe) to take load off the ALU pipe, which bounds this loop */
    float sz = u2f(((eb >> 16) & 0xffu) << 23) * r.idir.z;
    float ox = (n.v[0].x - r.o.x) * r.idir.x - sx;
    float oy = (n.v[0].y - r.o.y) * r.idir.y - sy;
    float oz = (n.v[0].z - r.o.z) * r.idir.z;
    /* near / far byte planes per axis according to the ray octant */
    bool px = r.idir.x >= 0.0f, py = r.idir.y >= 0.0f, pz = r.idir.z >= 0.0f;
    unsigned nx0 = f2u(px ? n.v[2].x : n.v[3].z), nx1 = f2u(px ? n.v[2].y : n.v[3].w);
    unsigned fx0 = f2u(px ? n.v[3].z : n.v[2].x), fx1 = f2u(px ? n.v[3].w : n.v[2].y);
    unsigned ny0 = f2u(py ? n.v[2].z : n.v[4].x), ny1 = f2u(py ? n.v[2].w : n.v[4].y);
    unsigned fy0 = f2u(py ? n.v[4].x : n.v[2].z), fy1 = f2u(py ? n.v[4].y : n.v[2].w);
    unsigned nz0 = f2u(pz ? n.v[3].x : n.v[4].z), nz1 = f2u(pz ? n.v[3].y : n.v[4].w);
    unsigned fz0 = f2u(pz ? n.v[4].z : n.v[3].x), fz1 = f2u(pz ? n.v[4].w : n.v[3].y);
    unsigned one = r.one;
#ifndef GPURT_NODE_TEST_SIGN
#define GPURT_NODE_TEST_SIGN 1
#endif
#if GPURT_NODE_TEST_SIGN
    /* The interval test without the two clamps and without predicates: the child is missed iff one of
     *   tf3 - tn3,   tmax - tn3,   tf3 - tmin        (tn3 / tf3 = latest entry / earliest exit over the three slabs)
     * is negative, i.e. iff the OR of the three differences has its sign bit set; a funnel shift moves that bit into the
     * mask.  Per child: 2 FMNMX3 + LOP3 + SHF on the ALU pipe and 3 FADD on the FMA pipe, instead of 2 FMNMX + 2 FMNMX3 +
     * FSETP + SEL + IADD3 on the ALU pipe, which bounds this loop (ncu: ALU 64-70 %, FMA 27 %).  Same decisions as
     * `max(tn3, tmin) <= min(tf3, tmax)` except where a difference is exactly -0 (a plane through the origin at tmin = 0)
     * or tmax < tmin: a child that cannot hold an accepted hit (t > tmin, padded boxes) in either form. */
    unsigned acc = 0;
#pragma unroll
    for(int i = 7; i >= 0; i--) {
        const int b = i & 3;
        float tnx = fmaf(byte_as_unit_float(i < 4 ? nx0 : nx1, one, b), sx, ox);
        float tny = fmaf(byte_as_unit_float(i < 4 ? ny0 : ny1, one, b), sy, oy);
        float tnz = fmaf((float)(((i < 4 ? nz0 : nz1) >> (8 * b)) & 0xffu), sz, oz);
        float tfx = fmaf(byte_as_unit_float(i < 4 ? fx0 : fx1, one, b), sx, ox);
        float tfy = fmaf(byte_as_unit_float(i < 4 ? fy0 : fy1, one, b), sy, oy);
        float tfz = fmaf((float)(((i < 4 ? fz0 : fz1) >> (8 * b)) & 0xffu), sz, oz);
        float tn3 = fmaxf(fmaxf(tnx, tny), tnz);
        float tf3 = fminf(fminf(tfx, tfy), tfz);
        unsigned miss = f2u(tf3 - tn3) | f2u(tmax - tn3) | f2u(tf3 - r.tmin);
#if defined(__CUDA_ARCH__)
        acc = __funnelshift_l(miss, acc, 1); /* (acc << 1) | (miss >> 31) */
#else
        acc = (acc << 1) | (miss >> 31);
#endif
    }
    return ~acc & 0xffu;
#else
    unsigned hits = 0;
#pragma unroll
    for(int i = 0; i < 8; i++) {
        const int b = i & 3;
        float tnx = fmaf(byte_as_unit_float(i < 4 ? nx0 : nx1, one, b), sx, ox);
        float tny = fmaf(byte_as_unit_float(i < 4 ? ny0 : ny1, one, b), sy, oy);
        float tnz = fmaf((float)(((i < 4 ? nz0 : nz1) >> (8 * b)) & 0xffu), sz, oz);
        float tfx = fmaf(byte_as_unit_float(i < 4 ? fx0 : fx1, one, b), sx, ox);
        float tfy = fmaf(byte_as_unit_float(i < 4 ? fy0 : fy1, one, b), sy, oy);
        float tfz = fmaf((float)(((i < 4 ? fz0 : fz1) >> (8 * b)) & 0xffu), sz, oz);
        float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, r.tmin));
        float tf = fminf(fminf(tfx, tfy), fminf(tfz, tmax));
        if(tn <= tf) hits |= 1u << i;
    }
    return hits;
#endif
}

/* slot i -> traversal priority i ^ octinv (three conditional bit-swap stages on a byte), so that
 * the highest set bit is the nearest child */
GPURT_HD unsigned octant_permute8(unsigned x, unsigned octinv) {
    if(octinv & 1u) x = ((x & 0x55u) << 1) | ((x & 0xaau) >> 1);
    if(octinv & 2u) x = ((x & 0x33u) << 2) | ((x & 0xccu) >> 2);
    if(octinv & 4u) x = ((x & 0x0fu) << 4) | ((x & 0xf0u) >> 4);
    return x;
}

/* Per-node constants of the closest-point child test.  With u = 1 + q*2^-15 (byte_as_unit_float) and
 * S = 2^(e-127+15):  lo - p = fma(u_lo, S, a)   and   p - hi = fma(u_hi, -S, -a),  a = origin - S - p,
 * so each signed plane distance is one PRMT + one FFMA (no I2F, no extra subtraction). */
struct ChildDist {
    float sx, sy, sz, ax, ay, az;
};
GPURT_HD ChildDist make_child_dist(const Node8& n, F3 p, unsigned one) {
    (void)one;
    unsigned eb = f2u(n.v[0].w);
    ChildDist c;
    c.sx = u2f(((eb & 0xffu) + 15u) << 23);
    c.sy = u2f((((eb >> 8) & 0xffu) + 15u) << 23);
    c.sz = u2f((((eb >> 16) & 0xffu) + 15u) << 23);
    c.ax = (n.v[0].x - c.sx) - p.x;
    c.ay = (n.v[0].y - c.sy) - p.y;
    c.az = (n.v[0].z - c.sz) - p.z;
    return c;
}
GPURT_HD float child_dist2(const Node8& n, const ChildDist& c, int i, unsigned one) {
    const int b = i & 3;
    float lx = fmaf(byte_as_unit_float(f2u(i < 4 ? n.v[2].x : n.v[2].y), one, b), c.sx, c.ax);
    float ly = fmaf(byte_as_unit_float(f2u(i < 4 ? n.v[2].z : n.v[2].w), one, b), c.sy, c.ay);
    float lz = fmaf(byte_as_unit_float(f2u(i < 4 ? n.v[3].x : n.v[3].y), one, b), c.sz, c.az);
    float hx = fmaf(byte_as_unit_float(f2u(i < 4 ? n.v[3].z : n.v[3].w), one, b), -c.sx, -c.ax);
    float hy = fmaf(byte_as_unit_float(f2u(i < 4 ? n.v[4].x : n.v[4].y), one, b), -c.sy, -c.ay);
    float hz = fmaf(byte_as_unit_float(f2u(i < 4 ? n.v[4].z : n.v[4].w), one, b), -c.sz, -c.az);
    float dx = fmaxf(fmaxf(lx, hx), 0.0f);
    float dy = fmaxf(fmaxf(ly, hy), 0.0f);
    float dz = fmaxf(fmaxf(lz, hz), 0.0f);
    return fmaf(dx, dx, fmaf(dy, dy, dz * dz));
}

} // namespace gpurt
