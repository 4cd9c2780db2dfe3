/*
 * emu.cpp — CPU replay of the product's host+device traversal/collapse code (csrc/bvh8.cuh,
 * csrc/traverse.cuh) and of the wavefront integrator's per-pixel shading code (csrc/shade.cuh) for debugging without a GPU.  TEST-ONLY: built by tests/test_emu.py into
 * tests/emu/libemu.so, never linked into libgpurt.so and never used as a fallback.
 * The binary LBVH it collapses comes from the oracle (tests pass it in).
 */
#include <cstring>
#include <vector>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <thread>

#include "../../gpu-rt_b200/csrc/shade.cuh"
#include "../../gpu-rt_b200/host/sah_split.h"

using namespace gpurt;

struct Emu {
    std::vector<Node8> nodes;
    std::vector<float4> tri_wide;
    std::vector<float4> tri_gid; /* world-space triangles in global primitive order (k_flatten's output) */
    unsigned depth = 0;
    float inflate = 0;
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

/* ---- experiment support for tools/sah_probe.py: alternative binary trees in the layout emu_build() takes, to answer
 * "what would this tree buy the same collapse + traversal" without a GPU. ---- */
namespace {
struct SahBuilder { /* box helpers of the experimental builders below */
    static float area(const float* b) {
        float ex = b[3] - b[0], ey = b[4] - b[1], ez = b[5] - b[2];
        return 2.0f * (ex * ey + ey * ez + ez * ex);
    }
    static void grow(float* b, const float* l, const float* h) {
        for(int k = 0; k < 3; k++) b[k] = fminf(b[k], l[k]), b[3 + k] = fmaxf(b[3 + k], h[k]);
    }
};
} // namespace

extern "C" {

/* Hybrid: keep the Morton LBVH below "cluster roots" (maximal subtrees with at most `cluster` triangles) and rebuild
 * only the tree ABOVE them by binned SAH over the cluster boxes — what a host-side top-level pass over a few thousand
 * boxes could do after the GPU LBVH.  in_*: the LBVH (oracle conventions); out_*: the re-assembled tree. */
namespace {
struct Hybrid {
    const unsigned* in_order;
    const int *in_left, *in_right;
    const float* in_boxes;
    const float* tris;
    int bins;
    unsigned* order;
    int *left, *right;
    float* boxes6;
    int next_node = 0;
    unsigned next_pos = 0;
    struct Item { int ref; float box[6]; float cen[3]; };
    std::vector<Item> items;

    void ref_box(int ref, float* b) const {
        if(ref >= 0) {
            for(int k = 0; k < 6; k++) b[k] = in_boxes[6 * ref + k];
        } else {
            const float* t = tris + 9ull * in_order[~ref];
            for(int k = 0; k < 3; k++) b[k] = fminf(fminf(t[k], t[3 + k]), t[6 + k]), b[3 + k] = fmaxf(fmaxf(t[k], t[3 + k]), t[6 + k]);
        }
    }
    unsigned count(int ref, std::vector<unsigned>& memo) const {
        if(ref < 0) return 1;
        if(memo[ref]) return memo[ref];
        return memo[ref] = count(in_left[ref], memo) + count(in_right[ref], memo);
    }
    void collect(int ref, unsigned cluster, std::vector<unsigned>& memo) {
        if(ref < 0 || count(ref, memo) <= cluster) {
            Item it;
            it.ref = ref;
            ref_box(ref, it.box);
            for(int k = 0; k < 3; k++) it.cen[k] = (it.box[k] + it.box[3 + k]) * 0.5f;
            items.push_back(it);
            return;
        }
        collect(in_left[ref], cluster, memo);
        collect(in_right[ref], cluster, memo);
    }
    /* copy an LBVH subtree, renumbering nodes and positions in DFS order */
    int copy(int ref) {
        if(ref < 0) {
            order[next_pos] = in_order[~ref];
            return ~(int)(next_pos++);
        }
        const int me = next_node++;
        for(int k = 0; k < 6; k++) boxes6[6 * me + k] = in_boxes[6 * ref + k];
        int l = copy(in_left[ref]);
        int r = copy(in_right[ref]);
        left[me] = l, right[me] = r;
        return me;
    }
    int build(unsigned a, unsigned b) {
        if(b - a == 1) return copy(items[a].ref);
        const int me = next_node++;
        float box[6] = {3e38f, 3e38f, 3e38f, -3e38f, -3e38f, -3e38f}, cb[6] = {3e38f, 3e38f, 3e38f, -3e38f, -3e38f, -3e38f};
        for(unsigned i = a; i < b; i++) SahBuilder::grow(box, items[i].box, items[i].box + 3), SahBuilder::grow(cb, items[i].cen, items[i].cen);
        for(int k = 0; k < 6; k++) boxes6[6 * me + k] = box[k];
        int best_axis = -1, best_bin = 0;
        float best_cost = 3e38f;
        std::vector<float> bb(6 * bins), left_area(bins);
        std::vector<unsigned> cnt(bins), left_cnt(bins);
        for(int ax = 0; ax < 3; ax++) {
            float ext = cb[3 + ax] - cb[ax];
            if(!(ext > 0)) continue;
            for(int i = 0; i < bins; i++) {
                cnt[i] = 0;
                for(int k = 0; k < 3; k++) bb[6 * i + k] = 3e38f, bb[6 * i + 3 + k] = -3e38f;
            }
            float scale = (float)bins / ext;
            for(unsigned i = a; i < b; i++) {
                int bi = std::min(bins - 1, (int)((items[i].cen[ax] - cb[ax]) * scale));
                cnt[bi]++;
                SahBuilder::grow(&bb[6 * bi], items[i].box, items[i].box + 3);
            }
            float run[6] = {3e38f, 3e38f, 3e38f, -3e38f, -3e38f, -3e38f};
            unsigned c = 0;
            for(int i = 0; i < bins - 1; i++) {
                if(cnt[i]) SahBuilder::grow(run, &bb[6 * i], &bb[6 * i + 3]);
                c += cnt[i];
                left_cnt[i] = c, left_area[i] = c ? SahBuilder::area(run) : 0.0f;
            }
            float rrun[6] = {3e38f, 3e38f, 3e38f, -3e38f, -3e38f, -3e38f};
            c = 0;
            for(int i = bins - 1; i > 0; i--) {
                if(cnt[i]) SahBuilder::grow(rrun, &bb[6 * i], &bb[6 * i + 3]);
                c += cnt[i];
                if(left_cnt[i - 1] == 0 || c == 0) continue;
                /* clusters hold similar triangle counts, so the item count stands in for the triangle count */
                float cost = left_area[i - 1] * (float)left_cnt[i - 1] + SahBuilder::area(rrun) * (float)c;
                if(cost < best_cost) best_cost = cost, best_axis = ax, best_bin = i;
            }
        }
        unsigned mid;
        if(best_axis >= 0) {
            float ext = cb[3 + best_axis] - cb[best_axis], scale = (float)bins / ext;
            auto it = std::stable_partition(items.begin() + a, items.begin() + b, [&](const Item& q) {
                return std::min(bins - 1, (int)((q.cen[best_axis] - cb[best_axis]) * scale)) < best_bin;
            });
            mid = (unsigned)(it - items.begin());
        } else
            mid = a + (b - a) / 2;
        if(mid == a || mid == b) mid = a + (b - a) / 2;
        int l = build(a, mid);
        int r = build(mid, b);
        left[me] = l, right[me] = r;
        return me;
    }
};
} // namespace

unsigned emu_hybrid_bvh2(const float* tris9, unsigned n, const unsigned* in_order, const int* in_left, const int* in_right,
                         const float* in_boxes6, unsigned cluster, int bins, unsigned* order, int* left, int* right,
                         float* boxes6) {
    Hybrid H;
    H.tris = tris9, H.in_order = in_order, H.in_left = in_left, H.in_right = in_right, H.in_boxes = in_boxes6;
    H.bins = bins, H.order = order, H.left = left, H.right = right, H.boxes6 = boxes6;
    if(n < 2) {
        if(n) order[0] = in_order[0];
        return n;
    }
    std::vector<unsigned> memo(n - 1, 0);
    H.collect(0, cluster, memo);
    H.build(0, (unsigned)H.items.size());
    return (unsigned)H.items.size();
}

/* PLOC (parallel locally-ordered clustering, Meister & Bittner 2018) over the Morton order: every round each cluster
 * looks `radius` neighbours to each side for the partner with the smallest merged surface area, mutual choices merge.
 * The GPU-friendly bottom-up alternative to a top-down binned-SAH build: rounds of search / merge / compact kernels
 * over the array the LBVH sort already produces.  The tree is renumbered depth-first at the end so that every node
 * covers a contiguous range of `order`, as the collapse expects. */
unsigned emu_ploc_bvh2(const float* tris9, unsigned n, const unsigned* morton_order, int radius, unsigned* order, int* left,
                       int* right, float* boxes6) {
    if(n < 2) {
        if(n) order[0] = morton_order[0];
        return 0;
    }
    struct Cl { int ref; float box[6]; };
    std::vector<Cl> cur(n), nxt;
    for(unsigned i = 0; i < n; i++) {
        unsigned g = morton_order[i];
        const float* t = tris9 + 9ull * g;
        cur[i].ref = ~(int)g;
        for(int k = 0; k < 3; k++)
            cur[i].box[k] = fminf(fminf(t[k], t[3 + k]), t[6 + k]), cur[i].box[3 + k] = fmaxf(fmaxf(t[k], t[3 + k]), t[6 + k]);
    }
    std::vector<int> tl(n - 1), tr(n - 1); /* temporary tree: refs >= 0 internal (creation order), < 0: ~gid */
    std::vector<float> tb(6ull * (n - 1));
    int made = 0;
    unsigned rounds = 0;
    std::vector<int> nn;
    while(cur.size() > 1) {
        const int m = (int)cur.size();
        nn.assign(m, -1);
        for(int i = 0; i < m; i++) {
            float best = 3e38f;
            for(int j = std::max(0, i - radius); j <= std::min(m - 1, i + radius); j++) {
                if(j == i) continue;
                float b[6];
                for(int k = 0; k < 3; k++) b[k] = fminf(cur[i].box[k], cur[j].box[k]), b[3 + k] = fmaxf(cur[i].box[3 + k], cur[j].box[3 + k]);
                float a = SahBuilder::area(b);
                if(a < best) best = a, nn[i] = j;
            }
        }
        nxt.clear();
        for(int i = 0; i < m; i++) {
            int j = nn[i];
            if(nn[j] == i) {
                if(i < j) {
                    Cl c;
                    c.ref = made;
                    for(int k = 0; k < 3; k++) c.box[k] = fminf(cur[i].box[k], cur[j].box[k]), c.box[3 + k] = fmaxf(cur[i].box[3 + k], cur[j].box[3 + k]);
                    tl[made] = cur[i].ref, tr[made] = cur[j].ref;
                    for(int k = 0; k < 6; k++) tb[6ull * made + k] = c.box[k];
                    made++;
                    nxt.push_back(c);
                }
            } else
                nxt.push_back(cur[i]);
        }
        cur.swap(nxt);
        rounds++;
    }
    /* depth-first renumbering from the root (the last node made) */
    int next_node = 0;
    unsigned next_pos = 0;
    struct Fr { int ref; int me; int phase; };
    std::vector<Fr> st;
    std::vector<int> newid(n - 1, -1);
    /* first pass: assign new ids in preorder, positions to leaves in left-to-right order */
    std::vector<int> stack{cur[0].ref};
    std::vector<int> pre;
    while(!stack.empty()) {
        int r = stack.back();
        stack.pop_back();
        if(r < 0) {
            order[next_pos++] = (unsigned)~r;
            continue;
        }
        newid[r] = next_node++;
        stack.push_back(tr[r]);
        stack.push_back(tl[r]);
    }
    /* second pass: leaf refs become positions; walk again in the same order to number them */
    next_pos = 0;
    stack.assign(1, cur[0].ref);
    std::vector<std::pair<int, int>> fix; /* (new node, side) waiting for a leaf position */
    std::vector<std::pair<int, int>> parent_slot{{-1, 0}};
    std::vector<std::pair<int, int>> pstack{{-1, 0}};
    while(!stack.empty()) {
        int r = stack.back();
        stack.pop_back();
        std::pair<int, int> ps = pstack.back();
        pstack.pop_back();
        int ref_out;
        if(r < 0) ref_out = ~(int)(next_pos++);
        else {
            ref_out = newid[r];
            for(int k = 0; k < 6; k++) boxes6[6ull * ref_out + k] = tb[6ull * r + k];
            stack.push_back(tr[r]), pstack.push_back({ref_out, 1});
            stack.push_back(tl[r]), pstack.push_back({ref_out, 0});
        }
        if(ps.first >= 0) (ps.second ? right : left)[ps.first] = ref_out;
    }
    return rounds;
}

/* The product's SAH-split tree (host/sah_split.h, GPURT_BUILD_SAH_SPLIT) in the arrays emu_build() takes:
 * order[n], left/right[n-1], boxes6[6 (n-1)] — same conventions as the oracle's orc_bvh_get_bvh2.  `bins` is ignored
 * (the definition fixes 16).  Returns the depth of the binary tree. */
unsigned emu_sah_bvh2(const float* tris9, unsigned n, int bins, unsigned* order, int* left, int* right, float* boxes6) {
    (void)bins;
    std::vector<float> lo(4ull * n), hi(4ull * n);
    for(unsigned g = 0; g < n; g++) {
        const float* t = tris9 + 9ull * g;
        for(int k = 0; k < 3; k++) {
            lo[4ull * g + k] = fminf(fminf(t[k], t[3 + k]), t[6 + k]);
            hi[4ull * g + k] = fmaxf(fmaxf(t[k], t[3 + k]), t[6 + k]);
        }
    }
    SahSplitTree T;
    build_sah_split(lo.data(), hi.data(), 4, n, T);
    for(unsigned i = 0; i < n; i++) order[i] = T.order[i];
    if(n < 2) return 0;
    /* node boxes = exact unions, children before parents: preorder numbering means children have larger ids */
    for(int i = (int)n - 2; i >= 0; i--) {
        left[i] = T.left[i], right[i] = T.right[i];
        float* b = boxes6 + 6ull * i;
        for(int k = 0; k < 3; k++) b[k] = 3e38f, b[3 + k] = -3e38f;
        for(int c : {T.left[i], T.right[i]}) {
            const float *cl, *ch;
            if(c >= 0) cl = boxes6 + 6ull * c, ch = cl + 3;
            else cl = &lo[4ull * T.order[~c]], ch = &hi[4ull * T.order[~c]];
            for(int k = 0; k < 3; k++) b[k] = fminf(b[k], cl[k]), b[3 + k] = fmaxf(b[3 + k], ch[k]);
        }
    }
    return T.depth;
}

/* milliseconds of build_sah_split() alone with a given thread count (0 = all cores) */
double emu_sah_split_ms(const float* tris9, unsigned n, unsigned threads) {
    std::vector<float> lo(4ull * n), hi(4ull * n);
    for(unsigned g = 0; g < n; g++)
        for(int k = 0; k < 3; k++) {
            const float* t = tris9 + 9ull * g;
            lo[4ull * g + k] = fminf(fminf(t[k], t[3 + k]), t[6 + k]), hi[4ull * g + k] = fmaxf(fmaxf(t[k], t[3 + k]), t[6 + k]);
        }
    SahSplitTree T;
    auto t0 = std::chrono::steady_clock::now();
    build_sah_split(lo.data(), hi.data(), 4, n, T, threads);
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

/* consistency of every array build_sah_split() hands to the device stages: 0 = fine, else the number of the failed check */
int emu_sah_split_check(const float* tris9, unsigned n) {
    std::vector<float> lo(4ull * n), hi(4ull * n);
    for(unsigned g = 0; g < n; g++)
        for(int k = 0; k < 3; k++) {
            const float* t = tris9 + 9ull * g;
            lo[4ull * g + k] = fminf(fminf(t[k], t[3 + k]), t[6 + k]), hi[4ull * g + k] = fmaxf(fmaxf(t[k], t[3 + k]), t[6 + k]);
        }
    SahSplitTree T, U;
    build_sah_split(lo.data(), hi.data(), 4, n, T, 1);
    build_sah_split(lo.data(), hi.data(), 4, n, U, 7);
    if(T.order != U.order || T.left != U.left || T.right != U.right || T.parent != U.parent || T.range_first != U.range_first ||
       T.range_last != U.range_last || T.depth != U.depth)
        return 1; /* deterministic, whatever the number of threads */
    if(T.order.size() != n) return 2;
    std::vector<char> seen(n, 0);
    for(unsigned g : T.order) {
        if(g >= n || seen[g]) return 3;
        seen[g] = 1;
    }
    if(n < 2) return 0;
    const unsigned ni = n - 1;
    if(T.left.size() != ni || T.right.size() != ni || T.parent.size() != (size_t)ni + n || T.range_first.size() != ni) return 4;
    if(T.parent[0] != -1) return 5;
    std::vector<int> child_count(ni, 0);
    for(unsigned i = 0; i < ni; i++) {
        const int L = T.left[i], R = T.right[i];
        for(int c : {L, R}) {
            int p = c >= 0 ? T.parent[c] : T.parent[(size_t)ni + (unsigned)~c];
            if(p != (int)i) return 6;
            if(c >= 0 && (c <= (int)i || c >= (int)ni)) return 7; /* preorder: children after their parent */
            if(c < 0 && (unsigned)~c >= n) return 8;
        }
        /* the left subtree covers the lower positions, together they tile the node's range */
        int lf = L >= 0 ? T.range_first[L] : ~L, ll = L >= 0 ? T.range_last[L] : ~L;
        int rf = R >= 0 ? T.range_first[R] : ~R, rl = R >= 0 ? T.range_last[R] : ~R;
        if(lf != T.range_first[i] || rl != T.range_last[i] || ll + 1 != rf) return 9;
    }
    if(T.range_first[0] != 0 || T.range_last[0] != (int)n - 1) return 10;
    return 0;
}

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
    E->tri_gid = tri_gid;
    E->inflate = inflate;
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
} /* extern "C" */

/* ---- replay of render.cu's frame: k_frame_begin, then per sample k_gen_camera + the bounce loop (k_tail's per-path
 * loop from depth 0 — the same shade_step / traverse8 calls the k_trace_closest_indirect + k_shade wavefront makes, in
 * the same per-pixel order), then k_frame_end.  Pixels only interact through the previous frame's buffers. ---- */
struct EmuFrameArgs {
    const uint32_t *descs, *tri_off, *vert_off;
    const float* verts;
    const uint32_t* idx;
    const uint32_t* lights;
    const uint32_t* tex_info;
    const uint8_t* texels;
    uint32_t n_objs, n_lights, n_tex;
    /* light BVH (render.cu pipe_light_accel): an Emu built over the lights' triangles, and their prefix offsets */
    void* light_bvh;
    const uint32_t* ltri_off;
};

/* pipe_light_groups() + k_light_groups of render.cu */
static int g_light_groups = 1, g_light_bvh = 1, g_light_verts = 1;
static void build_light_groups(const Emu* E, const EmuFrameArgs* A, std::vector<float4>& boxes, std::vector<uint2>& off) {
    const SceneLight* L = (const SceneLight*)A->lights;
    off.resize(A->n_lights);
    uint32_t total = 0;
    for(uint32_t l = 0; l < A->n_lights; l++) {
        uint32_t ng = (L[l].n_triangles + kLightRun - 1) / kLightRun, nsg = (ng + kLightRun - 1) / kLightRun;
        off[l] = uint2{total, total + nsg};
        total += nsg + ng;
    }
    boxes.resize(2ull * total);
    for(uint32_t l = 0; l < A->n_lights; l++)
        for(uint32_t r = 0; r < light_box_records(L[l].n_triangles); r++)
            light_box_record(E->tri_gid.data() + 3ull * A->tri_off[L[l].index], L[l].n_triangles, r, E->inflate * kLightPadScale,
                             boxes.data() + 2ull * (off[l].x + r));
}
static void fill_ctx(const Emu* E, const EmuFrameArgs* A, ShadeCtx& X) {
    X.S.verts = (Vertex*)A->verts, X.S.idx = (uint32_t*)A->idx, X.S.tri_off = (uint32_t*)A->tri_off;
    X.S.vert_off = (uint32_t*)A->vert_off, X.S.descs = (SceneDesc*)A->descs, X.S.lights = (SceneLight*)A->lights;
    X.S.n_objs = A->n_objs, X.S.n_lights = A->n_lights, X.S.n_tris = A->tri_off[A->n_objs];
    X.S.texels = (uint8_t*)A->texels, X.S.tex_info = (uint4*)A->tex_info, X.S.n_textures = A->n_tex;
    X.nodes = (const float4*)E->nodes.data(), X.tris = E->tri_wide.data(), X.tri_world = E->tri_gid.data();
    X.n_nodes = (unsigned)E->nodes.size();
    if(g_light_bvh && A->light_bvh) {
        const Emu* LB = (const Emu*)A->light_bvh;
        X.lnodes = (const float4*)LB->nodes.data(), X.ltris = LB->tri_wide.data(), X.ltri_off = A->ltri_off;
        X.n_lnodes = (unsigned)LB->nodes.size();
    }
}

template <int I>
static void render_rows(const FrameParams& P, const ShadeCtx& X, uint32_t i0, uint32_t i1, float4* image, float4* res_cur,
                        float4* gpos, float4* gnorm, float4* galb, float4* acc, float4* pathA, float4* pathB,
                        unsigned long long* counts) {
    const bool restir = I == 3 || I == 4;
    unsigned long long nc = 0, na = 0;
    for(uint32_t li = i0; li < i1; li++) {
        const uint32_t i = shard_pixel(P, li);
        pixel_begin(P, restir ? 1 : 0, i, acc, pathB, gpos, gnorm, galb, res_cur);
        for(uint32_t s = 0; s < (uint32_t)P.c.samples && P.c.max_depth > 0; s++) {
            float4 ray[2];
            uint32_t q;
            pixel_gen_camera(P, s, i, 0, pathA, pathB, ray, &q);
            Shader sh(X, P);
            nc += path_tail<I>(P, X, sh, s, 0, i, ray[0], ray[1], pathA, pathB, acc, gpos, gnorm, galb, res_cur);
            nc += sh.n_closest, na += sh.n_any;
        }
        pixel_end(P, i, acc, image, gpos, gnorm, X.ppos, X.pnorm, X.palb, nullptr);
    }
    counts[0] = nc, counts[1] = na;
}

extern "C" {
/* shard_pixel() of shade.cuh for every local index of one shard: out[n_local] global pixel indices; returns n_local
 * computed the way render_core does */
unsigned emu_shard_pixels(unsigned w, unsigned h, unsigned band_rows, unsigned n_shards, unsigned shard, unsigned* out) {
    FrameParams P;
    std::memset(&P, 0, sizeof(P));
    P.W = w, P.H = h, P.band_rows = band_rows ? band_rows : h, P.n_shards = band_rows ? n_shards : 1, P.shard = band_rows ? shard : 0;
    unsigned n_local = 0, bands = (h + P.band_rows - 1) / P.band_rows;
    for(unsigned g = P.shard; g < bands; g += P.n_shards) n_local += std::min(P.band_rows, h - g * P.band_rows) * w;
    P.n_local = n_local;
    if(out)
        for(unsigned i = 0; i < n_local; i++) out[i] = shard_pixel(P, i);
    return n_local;
}

void emu_set_light_groups(int on) { g_light_groups = on; }
void emu_set_light_bvh(int on) { g_light_bvh = on; }
void emu_set_light_verts(int on) { g_light_verts = on; }

/* light_pdf(p, d) for n rays (6 floats each): through the light-run boxes, testing every triangle, through the light
 * BVH -> out3[3 i .. 3 i + 2] */
void emu_light_pdf(void* bvh, const EmuFrameArgs* A, const float* rays6, unsigned long long n, float* out3) {
    Emu* E = (Emu*)bvh;
    FrameParams P;
    std::memset(&P, 0, sizeof(P));
    P.c.n_lights = (int)A->n_lights, P.c.n_objs = (int)A->n_objs;
    ShadeCtx X{};
    fill_ctx(E, A, X);
    std::vector<float4> lboxes;
    std::vector<uint2> loff;
    build_light_groups(E, A, lboxes, loff);
    const unsigned n_lnodes = X.n_lnodes;
    for(int mode = 0; mode < 3; mode++) {
        X.lgrp = mode != 1 ? lboxes.data() : nullptr, X.lgrp_off = loff.data();
        X.n_lnodes = mode == 2 ? n_lnodes : 0;
        Shader sh(X, P);
        for(unsigned long long i = 0; i < n; i++) {
            const float* r = rays6 + 6 * i;
            out3[3 * i + mode] = sh.light_pdf(F3{r[0], r[1], r[2]}, F3{r[3], r[4], r[5]});
        }
    }
}

/* GpurtPipeParams::light_sampling of the following emu_render_frame calls (render.cu pipe_light_cdf builds the table) */
static uint32_t g_light_sampling = 0;
void emu_set_light_sampling(uint32_t mode) { g_light_sampling = mode; }

/* band sharding of the following emu_render_frame calls (gpurt_pipe_set_shard); 0 rows = the whole frame */
static uint32_t g_band_rows = 0, g_n_shards = 1, g_shard = 0;
void emu_set_shard(uint32_t band_rows, uint32_t n_shards, uint32_t shard) { g_band_rows = band_rows, g_n_shards = n_shards, g_shard = shard; }
/* history_row_readers() of shade.cuh: the shards that receive row y of `shard`'s previous frame (k_history_push) */
unsigned long long emu_history_row_readers(uint32_t w, uint32_t h, uint32_t band_rows, uint32_t n_shards, uint32_t shard, uint32_t y,
                                           uint32_t halo) {
    FrameParams P;
    std::memset(&P, 0, sizeof(P));
    P.W = w, P.H = h, P.band_rows = band_rows, P.n_shards = n_shards, P.shard = shard;
    return history_row_readers(P, y, halo);
}
/* shard_row() of shade.cuh */
uint32_t emu_shard_row(uint32_t w, uint32_t h, uint32_t band_rows, uint32_t n_shards, uint32_t shard, uint32_t j) {
    FrameParams P;
    std::memset(&P, 0, sizeof(P));
    P.W = w, P.H = h, P.band_rows = band_rows, P.n_shards = n_shards, P.shard = shard;
    return shard_row(P, j);
}

/* ReSTIR spatial-reuse extension of the following emu_render_frame calls (GpurtPipeParams::spatial_samples / spatial_radius) */
static uint32_t g_spatial_samples = 0;
static float g_spatial_radius = 16.0f;
void emu_set_spatial(uint32_t samples, float radius) { g_spatial_samples = samples, g_spatial_radius = radius; }

void emu_render_frame(void* bvh, const EmuFrameArgs* A, const uint32_t* consts, const uint32_t* camera, uint32_t w, uint32_t h,
                      uint32_t seed_val, float* image, const uint32_t* prev_res, uint32_t* out_res, const float* ppos,
                      const float* pnorm, const float* palb, float* pos, float* norm, float* alb,
                      unsigned long long* ray_counts2, int threads) {
    Emu* E = (Emu*)bvh;
    static bool lut_done = false;
    if(!lut_done) { /* upload_lut_once() of render.cu */
        for(int i = 0; i < 256; i++) {
            double c = i / 255.0;
            c_srgb_lut[i] = (float)(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
        }
        lut_done = true;
    }
    FrameParams P;
    std::memset(&P, 0, sizeof(P));
    std::memcpy(&P.c, consts, sizeof(P.c));
    std::memcpy(&P.cam, camera, sizeof(P.cam));
    P.W = w, P.H = h, P.seed_val = seed_val;
    P.band_rows = g_band_rows ? g_band_rows : h, P.n_shards = g_band_rows ? g_n_shards : 1, P.shard = g_band_rows ? g_shard : 0;
    P.n_local = emu_shard_pixels(w, h, g_band_rows, g_n_shards, g_shard, nullptr);
    P.spatial_samples = g_spatial_samples, P.spatial_radius = g_spatial_radius;
    ShadeCtx X{};
    fill_ctx(E, A, X);
    std::vector<float4> lboxes;
    std::vector<uint2> loff;
    if(g_light_groups && A->n_lights) {
        build_light_groups(E, A, lboxes, loff);
        X.lgrp = lboxes.data(), X.lgrp_off = loff.data();
    }
    std::vector<float4> lverts; /* pipe_light_verts() + k_light_verts of render.cu */
    std::vector<uint32_t> lvoff(A->n_lights);
    if(g_light_verts && A->n_lights) {
        const SceneLight* L = (const SceneLight*)A->lights;
        uint32_t total = 0;
        for(uint32_t l = 0; l < A->n_lights; l++) lvoff[l] = total, total += L[l].n_triangles;
        lverts.resize(3ull * total);
        for(uint32_t l = 0; l < A->n_lights; l++)
            for(uint32_t t = 0; t < L[l].n_triangles; t++) light_world_tri(X.S, L[l].index, t, lverts.data() + 3ull * (lvoff[l] + t));
        X.lverts = lverts.data(), X.lvert_off = lvoff.data();
    }
    std::vector<float> lcdf; /* pipe_light_cdf() of render.cu */
    std::vector<uint32_t> lcdf_off(A->n_lights + 1, 0);
    P.light_sampling = g_light_sampling == 1 ? 1u : 0u;
    if(P.light_sampling && A->n_lights && P.c.integrator != 1) {
        const SceneLight* L = (const SceneLight*)A->lights;
        for(uint32_t l = 0; l < A->n_lights; l++) lcdf_off[l + 1] = lcdf_off[l] + L[l].n_triangles;
        lcdf.assign(std::max(lcdf_off[A->n_lights], 1u), 0.0f);
        float run = 0.0f;
        for(uint32_t l = 0; l < A->n_lights; l++) {
            const float* e = reinterpret_cast<const float*>(X.S.descs + L[l].index) + 36;
            for(uint32_t t = 0; t < L[l].n_triangles; t++) {
                float4 q[3];
                light_world_tri(X.S, L[l].index, t, q);
                float w = light_tri_power(F3{q[0].x, q[0].y, q[0].z}, F3{q[1].x, q[1].y, q[1].z}, F3{q[2].x, q[2].y, q[2].z}, F3{e[0], e[1], e[2]});
                if(!(w > 0.0f) || w > 3.0e38f) w = 0.0f;
                run += w;
                lcdf[lcdf_off[l] + t] = run;
            }
        }
        X.lcdf = lcdf.data(), X.lcdf_off = lcdf_off.data(), X.n_ltris = lcdf_off[A->n_lights];
    }
    X.prev_res = (const float4*)prev_res, X.ppos = (const float4*)ppos, X.pnorm = (const float4*)pnorm, X.palb = (const float4*)palb;
    const uint32_t n = P.n_local; /* thread ranges are over local indices; buffers are indexed by the global pixel */
    std::vector<float4> acc((size_t)w * h), pathA((size_t)w * h), pathB((size_t)w * h);
    int T = threads > 0 ? threads : (int)std::max(1u, std::thread::hardware_concurrency());
    T = std::min<int>(T, (int)std::max(1u, n / 64));
    std::vector<unsigned long long> cnt(2 * T, 0);
    std::vector<std::thread> pool;
    for(int t = 0; t < T; t++) {
        uint32_t i0 = (uint32_t)((unsigned long long)n * t / T), i1 = (uint32_t)((unsigned long long)n * (t + 1) / T);
        auto run = [&, i0, i1, t] {
#define EMU_ROWS(I)                                                                                                      \
    render_rows<I>(P, X, i0, i1, (float4*)image, (float4*)out_res, (float4*)pos, (float4*)norm, (float4*)alb, acc.data(), \
                   pathA.data(), pathB.data(), &cnt[2 * t])
            switch(P.c.integrator) {
            case 0: EMU_ROWS(0); break;
            case 1: EMU_ROWS(1); break;
            case 2: EMU_ROWS(2); break;
            case 3: EMU_ROWS(3); break;
            default: EMU_ROWS(4); break;
            }
#undef EMU_ROWS
        };
        pool.emplace_back(run);
    }
    for(auto& th : pool) th.join();
    if(ray_counts2) {
        ray_counts2[0] = ray_counts2[1] = 0;
        for(int t = 0; t < T; t++) ray_counts2[0] += cnt[2 * t], ray_counts2[1] += cnt[2 * t + 1];
    }
}
}
