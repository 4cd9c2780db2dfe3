/*
 * sah_split.h — top-down binned-SAH binary tree over triangle boxes (GPURT_BUILD_SAH_SPLIT), host side.
 *
 * The default build orders primitives along a Morton curve (contract N6) because that is a sort plus two linear
 * passes on the GPU.  On real meshes a tree split by the surface-area heuristic is visited about 30 % less per ray
 * (tools/sah_probe.py), so a static scene — the reference's use case: BLAS built once with
 * PREFER_FAST_TRACE, src/vk/vulkan.cpp:881-936 — can ask for one.  This header is the *definition* of that tree, written
 * so that a parallel builder can reproduce it bit for bit: every reduction is a min / max or an integer count (order
 * independent), every partition is stable, every float expression is written out and compiled without contraction.
 *
 *   items      triangles, initially in global primitive id order
 *   box(t)     the world-space AABB k_flatten wrote (tri_lo / tri_hi), centroid c(t) = (lo + hi) * 0.5f (as in N6)
 *   node(S)    cb = centroid bounds of S.  For axis = x, y, z with ext = cb.hi - cb.lo > 0:
 *                bin(t) = min(BINS - 1, (int)((c(t)[axis] - cb.lo[axis]) * ((float)BINS / ext))), 0 if that is NaN
 *                for each boundary k = 1..BINS-1 with both sides non-empty:
 *                  cost = area(union of bins < k) * count(bins < k) + area(union of bins >= k) * count(bins >= k)
 *              area(b) = 2 * (ex * ey + ey * ez + ez * ex);  the lowest cost wins, ties -> lower axis, then lower k.
 *              S is stably partitioned into bins < k | bins >= k.  No axis with ext > 0: split at the middle.
 *   leaves     single triangles; nodes are numbered in preorder (root 0), leaf positions in depth-first order
 *              (= the final order of `items`: a subtree over items[a, b) owns positions a..b-1).
 * Output conventions are those of the LBVH stage (build.cu k_karras): child >= 0 inner node, < 0: ~sorted position.
 */
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <thread>
#include <vector>

namespace gpurt {

struct SahSplitTree {
    std::vector<uint32_t> order;                 /* sorted position -> primitive id */
    std::vector<int> left, right;                /* n - 1 inner nodes */
    std::vector<int> parent;                     /* [n - 1 inner | n leaf positions], root -1 */
    std::vector<int> range_first, range_last;    /* sorted positions covered by each inner node */
    unsigned depth = 0;
};

namespace sah_detail {
constexpr int kBins = 16;
struct Box {
    float lo[3], hi[3];
};
inline Box empty_box() { return Box{{3.0e38f, 3.0e38f, 3.0e38f}, {-3.0e38f, -3.0e38f, -3.0e38f}}; }
inline void grow(Box& b, const float* lo, const float* hi) {
    for(int k = 0; k < 3; k++) b.lo[k] = std::min(b.lo[k], lo[k]), b.hi[k] = std::max(b.hi[k], hi[k]);
}
/* bin of a centroid coordinate; NaN / negative products (non-finite input) land in bin 0 instead of being undefined */
inline int bin_of(float c, float lo, float scale) {
    const float f = (c - lo) * scale;
    return f >= 0.0f ? (f < (float)kBins ? (int)f : kBins - 1) : 0;
}
inline float area(const Box& b) {
    float ex = b.hi[0] - b.lo[0], ey = b.hi[1] - b.lo[1], ez = b.hi[2] - b.lo[2];
    return 2.0f * (ex * ey + ey * ez + ez * ex);
}
} // namespace sah_detail

/* tri_lo / tri_hi: n records of `stride` floats, xyz first (the float4 arrays of the accel).
 * threads: 0 = std::thread::hardware_concurrency().  Node ids and leaf positions follow from subtree sizes alone (a
 * subtree over m items owns m - 1 consecutive preorder ids and m consecutive positions), so subtrees are independent
 * work items and the result does not depend on the number of threads. */
inline void build_sah_split(const float* tri_lo, const float* tri_hi, size_t stride, uint32_t n, SahSplitTree& T,
                            unsigned threads = 0) {
    using namespace sah_detail;
    T = SahSplitTree();
    T.order.resize(n);
    if(n == 0) return;
    if(n == 1) {
        T.order[0] = 0;
        T.parent.assign(1, -1);
        return;
    }
    const uint32_t ni = n - 1;
    T.left.resize(ni), T.right.resize(ni), T.range_first.resize(ni), T.range_last.resize(ni);
    T.parent.assign((size_t)ni + n, -1);
    std::vector<uint32_t> items(n), scratch(n);
    std::vector<float> cen(3ull * n);
    for(uint32_t g = 0; g < n; g++) {
        items[g] = g;
        for(int k = 0; k < 3; k++) cen[3ull * g + k] = (tri_lo[stride * g + k] + tri_hi[stride * g + k]) * 0.5f;
    }
    struct Job {
        uint32_t a, b;  /* segment of `items` = the leaf positions this subtree will own */
        int me;         /* preorder id of the subtree's root if it is an inner node */
        int parent;     /* inner node that waits for this subtree, -1 for the root */
        int side;       /* 0 left, 1 right */
        unsigned depth;
    };
    /* split one inner node: partitions items[a, b) and returns the boundary.  `par` > 1: the two reductions over the
     * segment (centroid bounds; per-axis bin counts and boxes) are cut into `par` slices and merged — min / max / integer
     * sums, so the merged values are the sequential ones */
    struct Bins {
        Box bb[3][kBins];
        uint32_t cnt[3][kBins];
    };
    auto split = [&](const Job& j, unsigned par) -> uint32_t {
        const uint32_t m = j.b - j.a;
        par = std::max(1u, std::min(par, m / 16384u));
        auto slice = [&](unsigned t, uint32_t& lo, uint32_t& hi) {
            lo = j.a + (uint32_t)((uint64_t)m * t / par), hi = j.a + (uint32_t)((uint64_t)m * (t + 1) / par);
        };
        auto run = [&](auto&& fn) {
            if(par == 1) return fn(0u);
            std::vector<std::thread> pool;
            for(unsigned t = 0; t < par; t++) pool.emplace_back(fn, t);
            for(auto& th : pool) th.join();
        };
        Box cb_one;
        Bins bins_one; /* the common case (par == 1) stays on the stack */
        std::vector<Box> cb_many(par > 1 ? par : 0);
        std::vector<Bins> bins_many(par > 1 ? par : 0);
        Box* cbs = par > 1 ? cb_many.data() : &cb_one;
        Bins* part = par > 1 ? bins_many.data() : &bins_one;
        run([&](unsigned t) {
            uint32_t lo, hi;
            slice(t, lo, hi);
            Box c = empty_box();
            for(uint32_t i = lo; i < hi; i++) grow(c, &cen[3ull * items[i]], &cen[3ull * items[i]]);
            cbs[t] = c;
        });
        Box cb = empty_box();
        for(unsigned t = 0; t < par; t++) grow(cb, cbs[t].lo, cbs[t].hi);
        float scale[3];
        bool use[3];
        for(int ax = 0; ax < 3; ax++) {
            const float ext = cb.hi[ax] - cb.lo[ax];
            use[ax] = ext > 0.0f;
            scale[ax] = use[ax] ? (float)kBins / ext : 0.0f;
        }
        run([&](unsigned t) {
            uint32_t lo, hi;
            slice(t, lo, hi);
            Bins& B = part[t];
            for(int ax = 0; ax < 3; ax++)
                for(int k = 0; k < kBins; k++) B.bb[ax][k] = empty_box(), B.cnt[ax][k] = 0;
            for(uint32_t i = lo; i < hi; i++) {
                const uint32_t g = items[i];
                for(int ax = 0; ax < 3; ax++) {
                    if(!use[ax]) continue;
                    const int bi = bin_of(cen[3ull * g + ax], cb.lo[ax], scale[ax]);
                    B.cnt[ax][bi]++;
                    grow(B.bb[ax][bi], tri_lo + stride * g, tri_hi + stride * g);
                }
            }
        });
        int best_axis = -1, best_k = 0;
        float best_cost = 3.0e38f;
        for(int ax = 0; ax < 3; ax++) {
            if(!use[ax]) continue;
            Box bb[kBins];
            uint32_t cnt[kBins];
            for(int k = 0; k < kBins; k++) {
                bb[k] = empty_box(), cnt[k] = 0;
                for(unsigned t = 0; t < par; t++) {
                    const Bins& B = part[t];
                    cnt[k] += B.cnt[ax][k];
                    if(B.cnt[ax][k]) grow(bb[k], B.bb[ax][k].lo, B.bb[ax][k].hi);
                }
            }
            float right_area[kBins];
            uint32_t right_cnt[kBins];
            Box run_box = empty_box();
            uint32_t c = 0;
            for(int k = kBins - 1; k >= 1; k--) {
                if(cnt[k]) grow(run_box, bb[k].lo, bb[k].hi);
                c += cnt[k];
                right_cnt[k] = c, right_area[k] = c ? area(run_box) : 0.0f;
            }
            run_box = empty_box(), c = 0;
            for(int k = 1; k < kBins; k++) {
                if(cnt[k - 1]) grow(run_box, bb[k - 1].lo, bb[k - 1].hi);
                c += cnt[k - 1];
                if(c == 0 || right_cnt[k] == 0) continue;
                const float cost = area(run_box) * (float)c + right_area[k] * (float)right_cnt[k];
                if(cost < best_cost) best_cost = cost, best_axis = ax, best_k = k;
            }
        }
        if(best_axis < 0) return j.a + (j.b - j.a) / 2;
        const float lo = cb.lo[best_axis], sc = scale[best_axis];
        uint32_t nl = 0, nr = 0;
        for(uint32_t i = j.a; i < j.b; i++) { /* stable partition through this segment's slice of the scratch array */
            const uint32_t g = items[i];
            const int bi = bin_of(cen[3ull * g + best_axis], lo, sc);
            if(bi < best_k) items[j.a + nl++] = g;
            else scratch[j.a + nr++] = g;
        }
        for(uint32_t i = 0; i < nr; i++) items[j.a + nl + i] = scratch[j.a + i];
        return j.a + nl; /* both sides are non-empty by construction of best_k */
    };
    /* emit the node or leaf of job j; inner nodes return their two child jobs */
    auto emit = [&](const Job& j, Job out[2], unsigned par) -> int {
        int ref, n_out = 0;
        if(j.b - j.a == 1) {
            T.order[j.a] = items[j.a];
            T.parent[(size_t)ni + j.a] = j.parent;
            ref = ~(int)j.a;
        } else {
            ref = j.me;
            T.parent[j.me] = j.parent;
            T.range_first[j.me] = (int)j.a, T.range_last[j.me] = (int)j.b - 1;
            const uint32_t mid = split(j, par);
            /* preorder: the left subtree takes the ids right after this node, (mid - a) - 1 of them */
            out[0] = Job{j.a, mid, j.me + 1, j.me, 0, j.depth + 1};
            out[1] = Job{mid, j.b, j.me + (int)(mid - j.a), j.me, 1, j.depth + 1};
            n_out = 2;
        }
        if(j.parent >= 0) (j.side ? T.right : T.left)[j.parent] = ref;
        return n_out;
    };
    auto run_subtree = [&](Job root, unsigned& max_depth) {
        std::vector<Job> stack{root};
        while(!stack.empty()) {
            const Job j = stack.back();
            stack.pop_back();
            if(j.b - j.a > 1) max_depth = std::max(max_depth, j.depth);
            Job out[2];
            if(emit(j, out, 1) == 2) stack.push_back(out[1]), stack.push_back(out[0]);
        }
    };
    if(threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
    const uint32_t grain = std::max<uint32_t>(4096, n / (8 * threads));
    /* top of the tree on this thread until the segments are small enough, then the subtrees in parallel */
    std::vector<Job> top{Job{0, n, 0, -1, 0, 1}}, tasks;
    unsigned depth = 0;
    while(!top.empty()) {
        const Job j = top.back();
        top.pop_back();
        if(threads > 1 && j.b - j.a <= grain) {
            tasks.push_back(j);
            continue;
        }
        if(j.b - j.a > 1) depth = std::max(depth, j.depth);
        Job out[2];
        if(emit(j, out, threads) == 2) top.push_back(out[1]), top.push_back(out[0]);
    }
    if(!tasks.empty()) {
        std::atomic<size_t> next{0};
        std::vector<unsigned> depths(threads, 0);
        std::vector<std::thread> pool;
        for(unsigned t = 0; t < std::min<size_t>(threads, tasks.size()); t++)
            pool.emplace_back([&, t] {
                for(size_t k; (k = next.fetch_add(1)) < tasks.size();) run_subtree(tasks[k], depths[t]);
            });
        for(auto& th : pool) th.join();
        for(unsigned d : depths) depth = std::max(depth, d);
    }
    T.depth = depth;
}

} // namespace gpurt
