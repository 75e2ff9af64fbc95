// Host-side (CPU) pieces of the GeneClust-ps tail behind the same C ABI: single-linkage labelling of the scaled MST,
// condensed tree (minimum cluster size), cluster stabilities and excess-of-mass selection with point labelling -
// hdbscan 0.8.29 `_hdbscan_linkage.label`, `_hdbscan_tree.condense_tree / compute_stability / get_clusters('eom') /
// do_labelling` as called at scGeneClust/tl/cluster.py:127-130.  Sequential union-find / tree walks over N - 1 edges:
// no GPU work, but ~35 ms of Python loops at N = 4000 in the reference-shaped host code; here ~1 ms.
// (GLOSH `outlier_scores` stays in numpy: it depends on numpy's own argsort tie order, see tl/_hdbscan_tail.py.)
// No CUDA in this translation unit; it is compiled by nvcc only because the library has one build recipe.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

#include "common.cuh"

namespace {

struct Row { int64_t parent, child; double lambda; int64_t size; };

// hdbscan.label: N = n_edges + 1 points; out rows = (left root, right root, distance, merged size)
void single_linkage(const double *mst, int64_t n_edges, std::vector<double> &hier) {
    const int64_t n = n_edges + 1;
    std::vector<int64_t> parent(2 * n - 1, -1), size(2 * n - 1, 0);
    for (int64_t i = 0; i < n; ++i) size[i] = 1;
    auto find = [&](int64_t x) {
        int64_t root = x;
        while (parent[root] != -1) root = parent[root];
        while (parent[x] != -1 && parent[x] != root) { const int64_t nx = parent[x]; parent[x] = root; x = nx; }
        return root;
    };
    hier.assign(n_edges * 4, 0.0);
    int64_t next_label = n;
    for (int64_t i = 0; i < n_edges; ++i) {
        const int64_t a = (int64_t)mst[3 * i], b = (int64_t)mst[3 * i + 1];
        const int64_t aa = find(a), bb = find(b);
        hier[4 * i] = (double)aa; hier[4 * i + 1] = (double)bb; hier[4 * i + 2] = mst[3 * i + 2];
        hier[4 * i + 3] = (double)(size[aa] + size[bb]);
        size[next_label] = size[aa] + size[bb];
        parent[aa] = parent[bb] = next_label;
        ++next_label;
    }
}

// breadth-first node list below `root` in hdbscan's order (level by level, left child before right child)
void bfs(const std::vector<double> &hier, int64_t num_points, int64_t root, std::vector<int64_t> &out) {
    out.clear();
    std::vector<int64_t> frontier{root}, next;
    while (!frontier.empty()) {
        out.insert(out.end(), frontier.begin(), frontier.end());
        next.clear();
        for (int64_t x : frontier)
            if (x >= num_points) {
                next.push_back((int64_t)hier[4 * (x - num_points)]);
                next.push_back((int64_t)hier[4 * (x - num_points) + 1]);
            }
        frontier.swap(next);
    }
}

void condense(const std::vector<double> &hier, int64_t n_edges, int64_t min_cluster_size, std::vector<Row> &rows) {
    const int64_t root = 2 * n_edges, num_points = root / 2 + 1;
    int64_t next_label = num_points + 1;
    std::vector<int64_t> relabel(root + 1, 0), nodes, sub;
    std::vector<char> ignore(root + 1, 0);
    relabel[root] = num_points;
    bfs(hier, num_points, root, nodes);
    rows.clear();
    for (int64_t node : nodes) {
        if (ignore[node] || node < num_points) continue;
        const double *ch = &hier[4 * (node - num_points)];
        const int64_t left = (int64_t)ch[0], right = (int64_t)ch[1];
        const double lam = ch[2] > 0.0 ? 1.0 / ch[2] : std::numeric_limits<double>::infinity();
        const int64_t lc = left >= num_points ? (int64_t)hier[4 * (left - num_points) + 3] : 1;
        const int64_t rc = right >= num_points ? (int64_t)hier[4 * (right - num_points) + 3] : 1;
        const bool bl = lc >= min_cluster_size, br = rc >= min_cluster_size;
        if (bl && br) {
            relabel[left] = next_label++;
            rows.push_back({relabel[node], relabel[left], lam, lc});
            relabel[right] = next_label++;
            rows.push_back({relabel[node], relabel[right], lam, rc});
            continue;
        }
        if (bl) relabel[left] = relabel[node];
        if (br) relabel[right] = relabel[node];
        const int64_t kids[2] = {left, right};
        const bool big[2] = {bl, br};
        for (int k = 0; k < 2; ++k) {
            if (big[k]) continue;
            bfs(hier, num_points, kids[k], sub);
            for (int64_t s : sub) {
                if (s < num_points) rows.push_back({relabel[node], s, lam, 1});
                ignore[s] = 1;
            }
        }
    }
}

}  // namespace

// mst_sorted: n_edges x 3 (i, j, weight) ascending by weight.  Outputs: the condensed tree (capacity 2 * n_edges rows;
// n_rows_out[0] rows written) and the EOM point labels (n_edges + 1 entries, -1 = noise).
extern "C" int gc_host_hdbscan_tree(const double *mst_sorted, int64_t n_edges, int64_t min_cluster_size, int64_t *parent,
                                    int64_t *child, double *lambda, int64_t *size, int64_t *n_rows_out, int64_t *labels) {
    GC_REQUIRE(mst_sorted && parent && child && lambda && size && n_rows_out && labels, "null pointer");
    GC_REQUIRE(n_edges >= 1 && min_cluster_size >= 2, "bad arguments");
    std::vector<double> hier;
    single_linkage(mst_sorted, n_edges, hier);
    std::vector<Row> rows;
    condense(hier, n_edges, min_cluster_size, rows);
    const int64_t n_rows = (int64_t)rows.size();
    GC_REQUIRE(n_rows >= 1 && n_rows <= 2 * n_edges, "condensed tree size out of range");
    for (int64_t i = 0; i < n_rows; ++i) { parent[i] = rows[i].parent; child[i] = rows[i].child; lambda[i] = rows[i].lambda; size[i] = rows[i].size; }
    n_rows_out[0] = n_rows;

    // compute_stability: births = min lambda at which a node appears as a child; stability += (lambda - birth) * size
    int64_t smallest = rows[0].parent, largest_parent = rows[0].parent, largest_child = rows[0].child;
    for (const Row &r : rows) {
        smallest = std::min(smallest, r.parent); largest_parent = std::max(largest_parent, r.parent);
        largest_child = std::max(largest_child, r.child);
    }
    largest_child = std::max(largest_child, smallest);
    const double nan = std::numeric_limits<double>::quiet_NaN();
    std::vector<double> births(largest_child + 1, nan);
    for (const Row &r : rows) births[r.child] = std::isnan(births[r.child]) ? r.lambda : std::fmin(births[r.child], r.lambda);
    births[smallest] = 0.0;
    const int64_t n_clusters = largest_parent - smallest + 1;
    std::vector<double> stability(n_clusters, 0.0);
    for (const Row &r : rows) stability[r.parent - smallest] += (r.lambda - births[r.parent]) * (double)r.size;

    // excess of mass over the cluster tree, children before parents (descending id), the root excluded
    std::vector<std::vector<int64_t>> kids(n_clusters);
    for (const Row &r : rows)
        if (r.size > 1) kids[r.parent - smallest].push_back(r.child);
    std::vector<char> is_cluster(n_clusters, 1);
    is_cluster[0] = 0;                                           // allow_single_cluster = False
    std::vector<int64_t> stack;
    for (int64_t c = largest_parent; c > smallest; --c) {
        double subtree = 0.0;
        for (int64_t k : kids[c - smallest]) subtree += stability[k - smallest];
        if (subtree > stability[c - smallest]) {
            is_cluster[c - smallest] = 0;
            stability[c - smallest] = subtree;
        } else {
            stack.assign(kids[c - smallest].begin(), kids[c - smallest].end());
            while (!stack.empty()) {
                const int64_t s = stack.back();
                stack.pop_back();
                is_cluster[s - smallest] = 0;
                for (int64_t k : kids[s - smallest]) stack.push_back(k);
            }
        }
    }
    std::vector<int64_t> label_of(n_clusters, -1);
    int64_t next = 0;
    for (int64_t c = smallest + 1; c <= largest_parent; ++c)
        if (is_cluster[c - smallest]) label_of[c - smallest] = next++;

    // do_labelling: union every non-selected child into its parent; a point's representative decides its label
    std::vector<int64_t> uf(largest_parent + 1);
    for (int64_t i = 0; i <= largest_parent; ++i) uf[i] = i;
    auto find = [&](int64_t x) {
        int64_t root = x;
        while (uf[root] != root) root = uf[root];
        while (uf[x] != root) { const int64_t nx = uf[x]; uf[x] = root; x = nx; }
        return root;
    };
    for (const Row &r : rows) {
        const bool selected = r.child >= smallest && is_cluster[r.child - smallest];
        if (!selected) uf[find(r.child)] = find(r.parent);
    }
    for (int64_t p = 0; p < smallest; ++p) {
        const int64_t c = find(p);
        labels[p] = c > smallest ? label_of[c - smallest] : -1;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Singleton rescue: summed isolation depths of sklearn's IsolationForest(random_state=seed) on n <= 256 one-dimensional
// float32 samples (scGeneClust/tl/selection.py:51), the restatement documented in tl/selection.py::_isolation_forest_1d:
// numpy legacy MT19937 seeding chain -> sklearn's xorshift rand_r -> depth-first random splits, left child first.
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct Mt19937 {
    uint32_t mt[624];
    int idx;
    explicit Mt19937(uint32_t seed) {                                  // init_genrand (numpy _legacy_seeding for an int seed)
        mt[0] = seed;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    uint32_t next() {
        if (idx >= 624) {
            for (int k = 0; k < 624; ++k) {
                const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
        return y;
    }
    // RandomState.randint(2**31 - 1): masked rejection on 32-bit words (range 2**31 - 2, mask 0x7fffffff)
    uint32_t randint31() {
        uint32_t v;
        do { v = next() & 0x7fffffffu; } while (v > 0x7ffffffeu);
        return v;
    }
    double random_sample() {                                           // rk_double
        const uint32_t a = next() >> 5, b = next() >> 6;
        return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
    }
};

inline uint32_t rand_r_step(uint32_t &s) {                             // sklearn utils/_random.pxd our_rand_r
    if (s == 0) s = 1;
    s ^= s << 13; s ^= s >> 17; s ^= s << 5;
    return s & 0x7fffffffu;
}

}  // namespace

// depths_host[i] = sum over the trees of (node depth + c_of[leaf size]); c_of_host[m] = sklearn's _average_path_length(m),
// m = 0..n, computed by the caller with numpy so that its log() is numpy's.
extern "C" int gc_host_iforest_depths(const float *x_host, int n, uint32_t seed, int n_estimators, const double *c_of_host,
                                      double *depths_host) {
    GC_REQUIRE(x_host && c_of_host && depths_host, "null pointer");
    GC_REQUIRE(n >= 1 && n <= 256 && n_estimators >= 1, "n must be in [1, 256] (max_samples='auto' keeps every sample)");
    std::vector<double> y(n);
    {
        Mt19937 r(seed);                                               // y = RandomState(seed).uniform(size=n): impurity only
        for (int i = 0; i < n; ++i) y[i] = r.random_sample();
    }
    std::vector<uint32_t> seeds(n_estimators);
    {
        Mt19937 r(seed);                                               // BaseBagging._fit: randint(MAX_INT, size=n_estimators)
        for (int t = 0; t < n_estimators; ++t) seeds[t] = r.randint31();
    }
    int max_depth = 0;
    while ((1 << max_depth) < std::max(n, 2)) ++max_depth;             // ceil(log2(max(n, 2)))
    const double eps = std::numeric_limits<double>::epsilon();
    for (int i = 0; i < n; ++i) depths_host[i] = 0.0;
    std::vector<double> contrib(n);
    struct Node { std::vector<int> ids; int depth; };
    std::vector<Node> stack;
    for (int t = 0; t < n_estimators; ++t) {
        const uint32_t tree_seed = Mt19937(seeds[t]).randint31();      // _set_random_states
        uint32_t state = Mt19937(tree_seed).randint31();               // Splitter.__cinit__: randint(0, RAND_R_MAX)
        stack.clear();
        Node root;
        root.ids.resize(n);
        for (int i = 0; i < n; ++i) root.ids[i] = i;
        root.depth = 0;
        stack.push_back(std::move(root));
        while (!stack.empty()) {
            Node node = std::move(stack.back());
            stack.pop_back();
            const int m = (int)node.ids.size();
            bool leaf = node.depth >= max_depth || m < 2;
            if (!leaf) {
                double s1 = 0.0, s2 = 0.0;
                for (int i : node.ids) { s1 += y[i]; s2 += y[i] * y[i]; }
                const double mean = s1 / m;
                leaf = s2 / m - mean * mean <= eps;
            }
            if (!leaf) {
                rand_r_step(state);                                    // rand_int(0, 1): the only feature is drawn
                float lo = x_host[node.ids[0]], hi = lo;
                for (int i : node.ids) { const float v = x_host[i]; if (v < lo) lo = v; else if (v > hi) hi = v; }
                if (hi <= lo + 1e-7f) {                                // FEATURE_THRESHOLD, float32 arithmetic
                    leaf = true;
                } else {
                    double thr = ((double)hi - (double)lo) * (double)rand_r_step(state) / 2147483647.0 + (double)lo;
                    if (thr == (double)hi) thr = (double)lo;
                    Node left, right;
                    left.depth = right.depth = node.depth + 1;
                    for (int i : node.ids) ((double)x_host[i] <= thr ? left : right).ids.push_back(i);
                    stack.push_back(std::move(right));                 // left child is built first
                    stack.push_back(std::move(left));
                }
            }
            if (leaf) {
                const double d = (double)(node.depth + 1) + c_of_host[m] - 1.0;
                for (int i : node.ids) contrib[i] = d;
            }
        }
        for (int i = 0; i < n; ++i) depths_host[i] += contrib[i];
    }
    return 0;
}
