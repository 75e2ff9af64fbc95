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
