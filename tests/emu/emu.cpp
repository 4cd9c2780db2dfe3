/*
 * emu.cpp — CPU replay of the product's host+device traversal/collapse code (csrc/bvh8.cuh,
 * csrc/traverse.cuh) for debugging without a GPU.  TEST-ONLY: built by tests/test_emu.py into
 * tests/emu/libemu.so, never linked into libgpurt.so and never used as a fallback.
 * The binary LBVH it collapses comes from the oracle (tests pass it in).
 */
#include <cstring>
#include <vector>

#include "../../gpu-rt_b200/csrc/traverse.cuh"

using namespace gpurt;

struct Emu {
    std::vector<Node8> nodes;
    std::vector<float4> tri_wide;
    unsigned depth = 0;
};

static void fill_ranges(const int* left, const int* right, int node, std::vector<int>& rf, std::vector<int>& rl) {
    /* iterative post-order */
    std::vector<std::pair<int, int>> st{{node, 0}};
    while(!st.empty()) {
        auto [c, phase] = st.back();
        st.pop_back();
        if(phase == 0) {
            st.push_back({c, 1});
            if(left[c] >= 0) st.push_back({left[c], 0});
            if(right[c] >= 0) st.push_back({right[c], 0});
        } else {
            rf[c] = left[c] < 0 ? ~left[c] : rf[left[c]];
            rl[c] = right[c] < 0 ? ~right[c] : rl[right[c]];
        }
    }
}

extern "C" {

int g_greedy = 1; /* 1: the greedy largest-area collapse (the default); 0: SAH-optimal collapse (GPURT_BUILD_SAH_COLLAPSE) */
void emu_set_greedy(int g) { g_greedy = g; }

void* emu_build(const float* tris9, unsigned n, const unsigned* order, const int* left, const int* right,
                const float* boxes6, float inflate) {
    Emu* E = new Emu;
    std::vector<float4> tri_gid(3ull * n), tlo(n), thi(n), nlo(n ? n - 1 : 0), nhi(n ? n - 1 : 0);
    for(unsigned g = 0; g < n; g++) {
        const float* t = tris9 + 9ull * g;
        tri_gid[3 * g + 0] = {t[0], t[1], t[2], u2f(g)};
        tri_gid[3 * g + 1] = {t[3] - t[0], t[4] - t[1], t[5] - t[2], u2f(0)};
        tri_gid[3 * g + 2] = {t[6] - t[0], t[7] - t[1], t[8] - t[2], u2f(g)};
        tlo[g] = {fminf(fminf(t[0], t[3]), t[6]), fminf(fminf(t[1], t[4]), t[7]), fminf(fminf(t[2], t[5]), t[8]), 0};
        thi[g] = {fmaxf(fmaxf(t[0], t[3]), t[6]), fmaxf(fmaxf(t[1], t[4]), t[7]), fmaxf(fmaxf(t[2], t[5]), t[8]), 0};
    }
    for(unsigned i = 0; i + 1 < n; i++) {
        nlo[i] = {boxes6[6 * i], boxes6[6 * i + 1], boxes6[6 * i + 2], 0};
        nhi[i] = {boxes6[6 * i + 3], boxes6[6 * i + 4], boxes6[6 * i + 5], 0};
    }
    std::vector<int> rf(n ? n - 1 : 0), rl(n ? n - 1 : 0);
    if(n > 1) fill_ranges(left, right, 0, rf, rl);
    Bvh2View B{left, right, rf.data(), rl.data(), nlo.data(), nhi.data(), tlo.data(), thi.data(), order, inflate};
    /* SAH-optimal collapse tables, children before parents (the GPU computes them in the refit kernel) */
    std::vector<float> dp_cost(n > 1 ? 7ull * (n - 1) : 0);
    std::vector<unsigned char> dp_dec(n > 1 ? 8ull * (n - 1) : 0);
    if(n > 1 && !g_greedy) {
        std::vector<std::pair<int, int>> st{{0, 0}};
        while(!st.empty()) {
            auto [c, phase] = st.back();
            st.pop_back();
            if(phase == 0) {
                st.push_back({c, 1});
                if(left[c] >= 0) st.push_back({left[c], 0});
                if(right[c] >= 0) st.push_back({right[c], 0});
            } else
                dp_node(B, dp_cost.data(), dp_dec.data(), c, bvh2_child_box(B, c));
        }
        B.dp_dec = dp_dec.data();
    }
    E->tri_wide.resize(3ull * n);
    if(n == 0) return E;
    if(n <= (unsigned)kMaxLeafTris) {
        int ch[8];
        for(int s = 0; s < 8; s++) ch[s] = kEmptyChild;
        ch[0] = encode_leaf_range(0, n);
        Node8 node;
        encode_node(B, ch, 0, 0, node);
        E->nodes.push_back(node);
        for(unsigned k = 0; k < n; k++)
            for(int q = 0; q < 3; q++) E->tri_wide[3 * k + q] = tri_gid[3 * order[k] + q];
        E->depth = 1;
        return E;
    }
    std::vector<int> items{0};
    unsigned level_base = 0, tri_cursor = 0;
    while(!items.empty()) {
        size_t m = items.size();
        std::vector<int> children(8 * m);
        std::vector<unsigned> oi(m + 1, 0), ot(m + 1, 0);
        for(size_t i = 0; i < m; i++) {
            int nt;
            int ni = collapse_node(B, items[i], &children[8 * i], nt);
            oi[i + 1] = oi[i] + ni, ot[i + 1] = ot[i] + nt;
        }
        unsigned next_base = level_base + (unsigned)m;
        std::vector<int> next(oi[m]);
        E->nodes.resize(next_base);
        for(size_t i = 0; i < m; i++) {
            const int* ch = &children[8 * i];
            unsigned child_base = next_base + oi[i], tri_base = tri_cursor + ot[i];
            encode_node(B, ch, child_base, tri_base, E->nodes[level_base + i]);
            unsigned r = 0, t = 0;
            for(int s = 0; s < 8; s++) {
                int c = ch[s];
                if(c == kEmptyChild) continue;
                if(c >= 0) next[oi[i] + r++] = c;
                else {
                    unsigned first, count;
                    decode_leaf_range(c, first, count);
                    for(unsigned k = 0; k < count; k++, t++)
                        for(int q = 0; q < 3; q++)
                            E->tri_wide[3ull * (tri_base + t) + q] = tri_gid[3ull * order[first + k] + q];
                }
            }
        }
        level_base = next_base;
        tri_cursor += ot[m];
        items.swap(next);
        E->depth++;
    }
    if(tri_cursor != n) E->depth = 0xFFFFFFFFu; /* flag: lost triangles */
    return E;
}
void emu_free(void* h) { delete(Emu*)h; }
unsigned emu_n_nodes(void* h) { return (unsigned)((Emu*)h)->nodes.size(); }
unsigned emu_depth(void* h) { return ((Emu*)h)->depth; }

void emu_trace(void* h, const float* rays, unsigned long long n, unsigned* hits4, unsigned char* occ,
               unsigned long long* counters) {
    Emu* E = (Emu*)h;
    for(unsigned long long i = 0; i < n; i++) {
        const float* r = rays + 8 * i;
        HitRec b;
        b.t = r[7], b.u = b.v = 0, b.gid = kNoHit;
        if(occ) {
            HitRec a;
            occ[i] = !E->nodes.empty() && traverse8<true, false>((const float4*)E->nodes.data(), E->tri_wide.data(),
                                                                  f3(r[0], r[1], r[2]), f3(r[4], r[5], r[6]), r[3], r[7], a, nullptr);
        }
        if(!E->nodes.empty())
            traverse8<false, true>((const float4*)E->nodes.data(), E->tri_wide.data(), f3(r[0], r[1], r[2]),
                                   f3(r[4], r[5], r[6]), r[3], r[7], b, counters);
        hits4[4 * i + 0] = f2u(b.gid == kNoHit ? GPURT_INF : b.t);
        hits4[4 * i + 1] = f2u(b.u), hits4[4 * i + 2] = f2u(b.v), hits4[4 * i + 3] = b.gid;
    }
}

unsigned g_cpq_counts[2] = {0, 0};
void emu_cpq_counts(unsigned long long* out) { out[0] = g_cpq_counts[0], out[1] = g_cpq_counts[1], g_cpq_counts[0] = g_cpq_counts[1] = 0; }
void emu_cpq(void* h, const float* q, unsigned long long n, unsigned* res8) {
    Emu* E = (Emu*)h;
    for(unsigned long long i = 0; i < n; i++) {
        CpRec b;
        b.gid = kNoHit;
        if(!E->nodes.empty())
            closest_point8<512>((const float4*)E->nodes.data(), E->tri_wide.data(), f3(q[4 * i], q[4 * i + 1], q[4 * i + 2]),
                                q[4 * i + 3], b, g_cpq_counts);
        unsigned* o = res8 + 8 * i;
        if(b.gid == kNoHit) {
            o[0] = o[1] = o[2] = 0, o[3] = f2u(GPURT_INF), o[4] = kNoHit, o[5] = 0, o[6] = o[7] = 0;
        } else {
            const float4* tp = E->tri_wide.data() + 3ull * b.idx;
            F3 c = tri_point(f3(tp[0].x, tp[0].y, tp[0].z), f3(tp[1].x, tp[1].y, tp[1].z), f3(tp[2].x, tp[2].y, tp[2].z), b.v, b.w);
            o[0] = f2u(c.x), o[1] = f2u(c.y), o[2] = f2u(c.z), o[3] = f2u(sqrtf(b.d2));
            o[4] = b.gid, o[5] = 0, o[6] = f2u(b.v), o[7] = f2u(b.w);
        }
    }
}
}
