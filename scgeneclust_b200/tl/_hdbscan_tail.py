"""MST -> gene clusters + GLOSH outlier scores: the HDBSCAN tail of GeneClust-ps
(scGeneClust/tl/cluster.py:122-131: `label`, `condense_tree(., 3)`, `compute_stability`, `get_clusters(., 'eom', ...)`,
`outlier_scores` of hdbscan 0.8.29 `_hdbscan_linkage.pyx` / `_hdbscan_tree.pyx`).

Sequential union-find / tree walks over N - 1 edges of a few thousand genes: host side by nature (SURVEY.md section 8f
rank 1).  hdbscan is not installable in the build image, so this is a restatement of its published algorithm - PARITY
UNPINNED against hdbscan itself; the cluster labels are cross-checked against scikit-learn's fork of the same functions
(`sklearn.cluster._hdbscan._tree.tree_to_labels`) in tests/test_host_logic.py.
"""
from __future__ import annotations

import numpy as np

CONDENSED_DTYPE = [('parent', np.intp), ('child', np.intp), ('lambda_val', np.float64), ('child_size', np.intp)]


def single_linkage_label(mst: np.ndarray) -> np.ndarray:
    """`hdbscan._hdbscan_linkage.label`: (n_edges, 3) edges sorted by weight -> (n_edges, 4) scipy-style linkage rows
    (left root, right root, distance, merged size); N = n_edges + 1 points."""
    n_edges = mst.shape[0]
    n = n_edges + 1
    parent = np.full(2 * n - 1, -1, dtype=np.intp)
    size = np.concatenate([np.ones(n, dtype=np.intp), np.zeros(n - 1, dtype=np.intp)])
    next_label = n
    out = np.zeros((n_edges, 4))

    def find(x):
        root = x
        while parent[root] != -1:
            root = parent[root]
        while parent[x] != -1 and parent[x] != root:      # path compression
            parent[x], x = root, parent[x]
        return root

    for i in range(n_edges):
        a, b, delta = int(mst[i, 0]), int(mst[i, 1]), mst[i, 2]
        aa, bb = find(a), find(b)
        out[i] = (aa, bb, delta, size[aa] + size[bb])
        size[next_label] = size[aa] + size[bb]
        parent[aa] = parent[bb] = next_label
        next_label += 1
    return out


def _bfs_from_hierarchy(hierarchy: np.ndarray, root: int, num_points: int) -> list:
    result, frontier = [], [root]
    while frontier:
        result.extend(frontier)
        inner = [x - num_points for x in frontier if x >= num_points]
        frontier = hierarchy[inner, :2].flatten().astype(np.intp).tolist() if inner else []
    return result


def condense_tree(hierarchy: np.ndarray, min_cluster_size: int = 3) -> np.ndarray:
    """`hdbscan._hdbscan_tree.condense_tree`: drop splits that shed fewer than `min_cluster_size` points."""
    root = 2 * hierarchy.shape[0]
    num_points = root // 2 + 1
    next_label = num_points + 1
    relabel = np.empty(root + 1, dtype=np.intp)
    relabel[root] = num_points
    ignore = np.zeros(root + 1, dtype=bool)
    rows = []
    for node in _bfs_from_hierarchy(hierarchy, root, num_points):
        if ignore[node] or node < num_points:
            continue
        left, right, dist = int(hierarchy[node - num_points, 0]), int(hierarchy[node - num_points, 1]), hierarchy[node - num_points, 2]
        lam = 1.0 / dist if dist > 0.0 else np.inf
        left_count = int(hierarchy[left - num_points, 3]) if left >= num_points else 1
        right_count = int(hierarchy[right - num_points, 3]) if right >= num_points else 1
        big_l, big_r = left_count >= min_cluster_size, right_count >= min_cluster_size
        if big_l and big_r:
            for child, cnt in ((left, left_count), (right, right_count)):
                relabel[child] = next_label
                next_label += 1
                rows.append((relabel[node], relabel[child], lam, cnt))
            continue
        if big_l:
            relabel[left] = relabel[node]
        if big_r:
            relabel[right] = relabel[node]
        for child, big in ((left, big_l), (right, big_r)):
            if big:
                continue
            for sub in _bfs_from_hierarchy(hierarchy, child, num_points):
                if sub < num_points:
                    rows.append((relabel[node], sub, lam, 1))
                ignore[sub] = True
    return np.array(rows, dtype=CONDENSED_DTYPE)


def compute_stability(tree: np.ndarray) -> dict:
    """`hdbscan._hdbscan_tree.compute_stability`: sum over a cluster's rows of (lambda - birth lambda) * child size."""
    smallest = int(tree['parent'].min())
    largest_parent = int(tree['parent'].max())
    largest_child = max(int(tree['child'].max()), smallest)
    births = np.full(largest_child + 1, np.nan)
    np.fmin.at(births, tree['child'], tree['lambda_val'])          # min lambda per child (NaN start = identity of fmin)
    births[smallest] = 0.0
    result = np.zeros(largest_parent - smallest + 1)
    for parent, lam, size in zip(tree['parent'], tree['lambda_val'], tree['child_size']):   # row order as in hdbscan
        result[parent - smallest] += (lam - births[parent]) * size
    return {float(c): result[c - smallest] for c in range(smallest, largest_parent + 1)}


def get_clusters_eom(tree: np.ndarray, stability: dict) -> np.ndarray:
    """`get_clusters(tree, stability, 'eom', allow_single_cluster=False, match_reference_implementation=False,
    cluster_selection_epsilon=0.0, max_cluster_size=0)` -> point labels (-1 = noise)."""
    node_list = sorted(stability.keys(), reverse=True)[:-1]          # all clusters but the root
    cluster_tree = tree[tree['child_size'] > 1]
    is_cluster = {c: True for c in node_list}
    children_of = {}
    for parent, child in zip(cluster_tree['parent'], cluster_tree['child']):
        children_of.setdefault(int(parent), []).append(int(child))
    for node in node_list:
        kids = children_of.get(int(node), [])
        subtree = np.sum([stability[float(c)] for c in kids]) if kids else 0.0
        if subtree > stability[node]:
            is_cluster[node] = False
            stability[node] = subtree
        else:
            stack = list(kids)
            while stack:                                             # every descendant loses cluster status
                sub = stack.pop()
                is_cluster[float(sub)] = False
                stack.extend(children_of.get(sub, []))
    clusters = sorted(int(c) for c, keep in is_cluster.items() if keep)
    label_of = {c: i for i, c in enumerate(clusters)}
    selected = set(clusters)
    # do_labelling: union every non-selected child into its parent, then read off each point's representative
    root_cluster = int(tree['parent'].min())
    uf = np.arange(int(tree['parent'].max()) + 1, dtype=np.intp)

    def find(x):
        root = x
        while uf[root] != root:
            root = uf[root]
        while uf[x] != root:
            uf[x], x = root, uf[x]
        return root

    for parent, child in zip(tree['parent'], tree['child']):
        if int(child) not in selected:
            uf[find(int(child))] = find(int(parent))
    labels = np.empty(root_cluster, dtype=np.intp)
    for p in range(root_cluster):
        c = find(p)
        labels[p] = label_of[c] if c > root_cluster else -1
    return labels


def outlier_scores(tree: np.ndarray) -> np.ndarray:
    """`hdbscan._hdbscan_tree.outlier_scores` (GLOSH): (lambda_max(cluster) - lambda(point)) / lambda_max(cluster), with
    lambda_max propagated up the cluster tree in hdbscan's own (argsort-by-parent, break at the first point row) order."""
    parent, child, lam = tree['parent'], tree['child'], tree['lambda_val']
    deaths = np.zeros(int(parent.max()) + 1)
    np.maximum.at(deaths, parent, lam)                               # max_lambdas
    root_cluster = int(parent.min())
    for n in np.argsort(parent):
        cluster = int(child[n])
        if cluster < root_cluster:
            break
        if deaths[cluster] > deaths[parent[n]]:
            deaths[parent[n]] = deaths[cluster]
    result = np.zeros(root_cluster)
    is_point = child < root_cluster
    lam_max = deaths[parent[is_point]]
    lam_p = lam[is_point]
    with np.errstate(invalid='ignore', divide='ignore'):
        score = np.where((lam_max == 0.0) | ~np.isfinite(lam_p), 0.0, (lam_max - lam_p) / lam_max)
    result[child[is_point]] = score
    return result


def native_tree_and_labels(mst_sorted: np.ndarray, min_cluster_size: int = 3):
    """Condensed tree + EOM labels from the C++ host routine `gc_host_hdbscan_tree` of libgeneclust_b200.so (the same
    algorithm as `single_linkage_label` / `condense_tree` / `compute_stability` / `get_clusters_eom` above, ~30x
    faster than the Python loops; tests compare the two)."""
    from .._lib import lib
    mst = np.ascontiguousarray(mst_sorted, dtype=np.float64)
    n_edges = mst.shape[0]
    parent = np.empty(2 * n_edges, dtype=np.int64)
    child = np.empty(2 * n_edges, dtype=np.int64)
    lam = np.empty(2 * n_edges, dtype=np.float64)
    size = np.empty(2 * n_edges, dtype=np.int64)
    n_rows = np.zeros(1, dtype=np.int64)
    labels = np.empty(n_edges + 1, dtype=np.int64)
    lib().host_hdbscan_tree(mst.ctypes.data, n_edges, int(min_cluster_size), parent.ctypes.data, child.ctypes.data,
                            lam.ctypes.data, size.ctypes.data, n_rows.ctypes.data, labels.ctypes.data)
    k = int(n_rows[0])
    tree = np.empty(k, dtype=CONDENSED_DTYPE)
    tree['parent'], tree['child'], tree['lambda_val'], tree['child_size'] = parent[:k], child[:k], lam[:k], size[:k]
    return tree, labels.astype(np.intp)


def mst_to_clusters(mst_sorted: np.ndarray, n_vars: int, min_cluster_size: int = 3, native: bool = True):
    """The five calls of scGeneClust/tl/cluster.py:127-131 -> (labels, outlier_scores)."""
    if mst_sorted.shape[0] + 1 != n_vars:
        # the reference feeds a spanning FOREST to `label`, which assumes N - 1 edges, and then fails when it assigns
        # the mis-sized result to adata.var ("MST graph is not complete")
        raise ValueError(f"the MI > 0 redundancy graph is disconnected: {mst_sorted.shape[0]} MST edges for {n_vars} "
                         "genes (the reference fails here as well: Length of values does not match length of index)")
    if native:
        tree, labels = native_tree_and_labels(mst_sorted, min_cluster_size)
    else:
        slt = single_linkage_label(mst_sorted)
        tree = condense_tree(slt, min_cluster_size)
        labels = get_clusters_eom(tree, compute_stability(tree))
    return labels, outlier_scores(tree)
