// rm_bvh.cpp -- host-side builder of the hierarchy in rm_bvh.cuh: binned surface-area heuristic (16 bins, three
// axes), leaves of at most kBvhLeafMax primitives, median splits when the heuristic cannot separate a range or
// the tree gets deeper than 40 levels (so the traversal stack of kBvhStack entries always suffices).
// Pure C++: runs once per scene at upload (~70 ms for 10^5 primitives on one core; scenes of 8192 primitives and more
// sweep the top of the tree and build its subtrees on a thread pool -- the tree is the same for any number of threads),
// shared with the host emulation.
#include <algorithm>
#include <atomic>
#include <climits>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <thread>

#include "rm_scene.h"

namespace rm {

namespace {

struct Box {
    float lo[3], hi[3];
    void clear() {
        for (int a = 0; a < 3; a++) {
            lo[a] = std::numeric_limits<float>::infinity();
            hi[a] = -std::numeric_limits<float>::infinity();
        }
    }
    void grow(const Box& b) {
        for (int a = 0; a < 3; a++) {
            lo[a] = std::min(lo[a], b.lo[a]);
            hi[a] = std::max(hi[a], b.hi[a]);
        }
    }
    double half_area() const {
        const double x = (double)hi[0] - lo[0], y = (double)hi[1] - lo[1], z = (double)hi[2] - lo[2];
        return x * y + y * z + z * x;
    }
};

struct Item {
    Box box;
    float c[3];     // centroid
    int code;
};

float round_down(double v) {
    float f = (float)v;
    if ((double)f > v) f = std::nextafterf(f, -std::numeric_limits<float>::infinity());
    return f;
}
float round_up(double v) {
    float f = (float)v;
    if ((double)f < v) f = std::nextafterf(f, std::numeric_limits<float>::infinity());
    return f;
}

// A subtree handed to the thread pool: items [begin, end), found at `depth`.  Child codes below kDeferred + 2^20 mark
// such a subtree in the part of the tree built up front (leaf codes stay above -2^30, see make_leaf).
struct Task {
    int begin, end, depth;
};
constexpr int kDeferred = INT_MIN;

struct Builder {
    Buf<Item>& items;                  // shared; a builder only touches the ranges it is given
    std::vector<R4<float>>& nodes;
    std::vector<int>& prims;
    int max_depth = 0;
    std::vector<Task>* defer = nullptr;   // when set: ranges of at most `grain` items become tasks instead of subtrees
    int grain = 0;
    HostPool* pool = nullptr;             // when set: large ranges are swept on its threads

    Builder(Buf<Item>& i, std::vector<R4<float>>& n, std::vector<int>& p) : items(i), nodes(n), prims(p) {}

    int make_leaf(int begin, int end) {
        const int first = (int)prims.size();
        for (int i = begin; i < end; i++) prims.push_back(items[i].code);
        return ~((first << 3) | (end - begin));
    }

    // returns the child code of the subtree over items [begin, end) and its box
    int build(int begin, int end, int depth, Box& box) {
        max_depth = std::max(max_depth, depth);
        const int n = end - begin;
        // ranges of the top of a large tree are swept in blocks on the pool's threads; min / max and counts combine
        // exactly, so the splits are those of the sequential sweep
        constexpr int kBlock = 4096;
        const bool wide = pool != nullptr && n >= 4 * kBlock;
        const int n_blocks = (n + kBlock - 1) / kBlock;
        box.clear();
        Box cb;
        cb.clear();
        auto bounds_of = [&](int b0, int b1, Box& bx, Box& cx) {
            for (int i = b0; i < b1; i++) {
                bx.grow(items[i].box);
                for (int a = 0; a < 3; a++) {
                    cx.lo[a] = std::min(cx.lo[a], items[i].c[a]);
                    cx.hi[a] = std::max(cx.hi[a], items[i].c[a]);
                }
            }
        };
        if (wide) {
            std::vector<Box> part(2 * (size_t)n_blocks);
            pool->run(n_blocks, [&](int k) {
                part[2 * k].clear();
                part[2 * k + 1].clear();
                bounds_of(begin + k * kBlock, std::min(end, begin + (k + 1) * kBlock), part[2 * k], part[2 * k + 1]);
            });
            for (int k = 0; k < n_blocks; k++) {
                box.grow(part[2 * k]);
                cb.grow(part[2 * k + 1]);
            }
        } else {
            bounds_of(begin, end, box, cb);
        }
        if (n <= kBvhLeafMax) return make_leaf(begin, end);
        if (defer && n <= grain) {
            defer->push_back({begin, end, depth});
            return kDeferred + (int)defer->size() - 1;
        }

        int mid = -1;
        if (depth < 40) {
            constexpr int kBins = 16;
            struct Bins {
                Box bb[3][kBins];
                int cnt[3][kBins];
            };
            double scale[3];
            bool use[3];
            for (int a = 0; a < 3; a++) {
                const double ext = (double)cb.hi[a] - cb.lo[a];
                use[a] = ext > 0. && std::isfinite(ext);
                scale[a] = use[a] ? kBins / ext : 0.;
            }
            auto clear_bins = [](Bins& B) {
                for (int a = 0; a < 3; a++)
                    for (int k = 0; k < kBins; k++) {
                        B.bb[a][k].clear();
                        B.cnt[a][k] = 0;
                    }
            };
            auto bin_range = [&](int b0, int b1, Bins& B) {
                for (int i = b0; i < b1; i++)
                    for (int a = 0; a < 3; a++) {
                        if (!use[a]) continue;
                        int k = (int)(((double)items[i].c[a] - cb.lo[a]) * scale[a]);
                        k = std::min(std::max(k, 0), kBins - 1);
                        B.bb[a][k].grow(items[i].box);
                        B.cnt[a][k]++;
                    }
            };
            Bins all;
            clear_bins(all);
            if (wide) {
                std::vector<Bins> part((size_t)n_blocks);
                pool->run(n_blocks, [&](int k) {
                    clear_bins(part[k]);
                    bin_range(begin + k * kBlock, std::min(end, begin + (k + 1) * kBlock), part[k]);
                });
                for (int k = 0; k < n_blocks; k++)
                    for (int a = 0; a < 3; a++)
                        for (int j = 0; j < kBins; j++) {
                            all.bb[a][j].grow(part[k].bb[a][j]);
                            all.cnt[a][j] += part[k].cnt[a][j];
                        }
            } else {
                bin_range(begin, end, all);
            }
            double best = std::numeric_limits<double>::infinity();
            int best_axis = -1, best_bin = -1;
            for (int a = 0; a < 3; a++) {
                if (!use[a]) continue;
                const Box* bb = all.bb[a];
                const int* cnt = all.cnt[a];
                double right_area[kBins];
                Box acc;
                acc.clear();
                int right_cnt[kBins];
                int c = 0;
                for (int k = kBins - 1; k > 0; k--) {
                    acc.grow(bb[k]);
                    c += cnt[k];
                    right_area[k] = c ? acc.half_area() : 0.;
                    right_cnt[k] = c;
                }
                acc.clear();
                c = 0;
                for (int k = 0; k + 1 < kBins; k++) {
                    acc.grow(bb[k]);
                    c += cnt[k];
                    if (c == 0 || right_cnt[k + 1] == 0) continue;
                    const double cost = acc.half_area() * c + right_area[k + 1] * right_cnt[k + 1];
                    if (cost < best) {
                        best = cost;
                        best_axis = a;
                        best_bin = k;
                    }
                }
            }
            if (best_axis >= 0 && std::isfinite(best)) {
                const int a = best_axis;
                const double ext = (double)cb.hi[a] - cb.lo[a], scale = kBins / ext;
                auto it = std::partition(items.begin() + begin, items.begin() + end, [&](const Item& it_) {
                    int k = (int)(((double)it_.c[a] - cb.lo[a]) * scale);
                    k = std::min(std::max(k, 0), kBins - 1);
                    return k <= best_bin;
                });
                mid = (int)(it - items.begin());
                if (mid == begin || mid == end) mid = -1;
            }
        }
        if (mid < 0) {
            // median split along the widest centroid axis (by index when all centroids coincide)
            int a = 0;
            for (int k = 1; k < 3; k++)
                if ((double)cb.hi[k] - cb.lo[k] > (double)cb.hi[a] - cb.lo[a]) a = k;
            mid = begin + n / 2;
            std::nth_element(items.begin() + begin, items.begin() + mid, items.begin() + end,
                             [a](const Item& x, const Item& y) { return x.c[a] < y.c[a] || (x.c[a] == y.c[a] && x.code < y.code); });
        }
        const size_t me = nodes.size() / 4;
        nodes.resize(nodes.size() + 4);
        Box b0, b1;
        const int c0 = build(begin, mid, depth + 1, b0);
        const int c1 = build(mid, end, depth + 1, b1);
        write_node(me, b0, c0, b1, c1);
        return (int)me;
    }

    void write_node(size_t me, const Box& b0, int c0, const Box& b1, int c1) {
        float f0, f1;
        std::memcpy(&f0, &c0, 4);
        std::memcpy(&f1, &c1, 4);
        nodes[4 * me + 0] = {b0.lo[0], b0.hi[0], b0.lo[1], b0.hi[1]};
        nodes[4 * me + 1] = {b1.lo[0], b1.hi[0], b1.lo[1], b1.hi[1]};
        nodes[4 * me + 2] = {b0.lo[2], b0.hi[2], b1.lo[2], b1.hi[2]};
        nodes[4 * me + 3] = {f0, f1, 0.f, 0.f};
    }
};

}  // namespace

int build_bvh(const Buf<BvhPrimBox>& in, std::vector<R4<float>>& nodes, std::vector<int>& prims, HostPool* shared_pool) {
    nodes.clear();
    prims.clear();
    if (in.empty()) return 0;
    const int n = (int)in.size();
    // large scenes: the caller's threads or, without any, some of this call's own
    std::unique_ptr<HostPool> own_pool;
    HostPool* pool = nullptr;
    if (n >= 8192) {
        pool = shared_pool;
        if (!pool) {
            own_pool.reset(new HostPool(host_thread_count(16)));
            pool = own_pool.get();
        }
    }
    // S: largest finite coordinate magnitude; every box grows by 2^-14 S on each side (see rm_bvh.cuh)
    double S = 0.;
    for (const BvhPrimBox& p : in)
        for (int a = 0; a < 3; a++) {
            if (std::isfinite(p.lo[a])) S = std::max(S, std::fabs(p.lo[a]));
            if (std::isfinite(p.hi[a])) S = std::max(S, std::fabs(p.hi[a]));
        }
    const double pad = std::max(S, 1e-30) * (1.0 / 16384.0);
    const double big = 1e30;
    Buf<Item> items((size_t)n);
    auto make_items = [&](int b0, int b1) {
        for (int i = b0; i < b1; i++) {
            const BvhPrimBox& p = in[i];
            Item& it = items[i];
            for (int a = 0; a < 3; a++) {
                double lo = p.lo[a], hi = p.hi[a];
                if (!(lo >= -big)) lo = -big;      // also catches NaN: an unbounded box is conservative
                if (!(hi <= big)) hi = big;
                if (!(lo <= hi)) { lo = -big; hi = big; }
                it.box.lo[a] = round_down(lo - pad);
                it.box.hi[a] = round_up(hi + pad);
                it.c[a] = (float)(0.5 * (lo + hi));
            }
            it.code = p.code;
        }
    };
    if (pool) {
        constexpr int kBlock = 4096;
        pool->run((n + kBlock - 1) / kBlock, [&](int k) { make_items(k * kBlock, std::min(n, (k + 1) * kBlock)); });
    } else {
        make_items(0, n);
    }
    Builder b(items, nodes, prims);
    if (n <= kBvhLeafMax) {
        // the root is always an inner node: one real leaf and one empty one behind the same box
        Box box;
        box.clear();
        for (auto& it : items) box.grow(it.box);
        nodes.resize(4);
        const int leaf = b.make_leaf(0, n);
        b.write_node(0, box, leaf, box, ~0);
        return 1;
    }
    Box root;
    if (!pool) {
        b.build(0, n, 1, root);
        return b.max_depth;
    }
    // Large scenes: the top of the tree here, the subtrees of at most n/64 items on a pool of threads, each into
    // arrays of its own, appended in task order afterwards -- the same splits as the sequential build (a subtree is a
    // function of its item range only), a numbering that does not depend on thread timing.
    std::vector<Task> tasks;
    b.defer = &tasks;
    b.grain = std::max(n / 64, 1024);
    b.pool = pool;
    b.build(0, n, 1, root);
    struct Sub {
        std::vector<R4<float>> nodes;
        std::vector<int> prims;
        int root = 0, depth = 0;
    };
    std::vector<Sub> subs(tasks.size());
    pool->run((int)tasks.size(), [&](int k) {
        // arrays local to the worker while they grow (neighbouring Sub records share cache lines)
        std::vector<R4<float>> ln;
        std::vector<int> lp;
        Builder sb(items, ln, lp);
        Box box;
        const int root_code = sb.build(tasks[k].begin, tasks[k].end, tasks[k].depth, box);
        subs[k].nodes = std::move(ln);
        subs[k].prims = std::move(lp);
        subs[k].root = root_code;
        subs[k].depth = sb.max_depth;
    });
    int depth = b.max_depth;
    const size_t n_top = nodes.size() / 4;
    std::vector<int> mapped(tasks.size());
    for (size_t k = 0; k < tasks.size(); k++) {
        const int node_base = (int)(nodes.size() / 4), prim_base = (int)prims.size();
        auto remap = [&](int c) {
            if (c >= 0) return c + node_base;
            const int code = ~c;
            return ~((((code >> 3) + prim_base) << 3) | (code & 7));
        };
        for (size_t i = 0; i + 3 < subs[k].nodes.size(); i += 4) {
            R4<float> links = subs[k].nodes[i + 3];
            int c0, c1;
            std::memcpy(&c0, &links.x, 4);
            std::memcpy(&c1, &links.y, 4);
            c0 = remap(c0);
            c1 = remap(c1);
            std::memcpy(&links.x, &c0, 4);
            std::memcpy(&links.y, &c1, 4);
            subs[k].nodes[i + 3] = links;
        }
        mapped[k] = remap(subs[k].root);
        nodes.insert(nodes.end(), subs[k].nodes.begin(), subs[k].nodes.end());
        prims.insert(prims.end(), subs[k].prims.begin(), subs[k].prims.end());
        depth = std::max(depth, subs[k].depth);
    }
    for (size_t i = 0; i < n_top; i++) {
        R4<float>& links = nodes[4 * i + 3];
        int c[2];
        std::memcpy(&c[0], &links.x, 4);
        std::memcpy(&c[1], &links.y, 4);
        for (int j = 0; j < 2; j++)
            if (c[j] < kDeferred + (1 << 20)) c[j] = mapped[c[j] - kDeferred];
        std::memcpy(&links.x, &c[0], 4);
        std::memcpy(&links.y, &c[1], 4);
    }
    return depth;
}

}  // namespace rm
