"""CPU restatement of the reference's stage functions (test infrastructure; see oracle/__init__.py).

Every function cites the reference lines it follows (paths relative to /root/reference) or, where the
arithmetic lives in scikit-learn, the installed sklearn 1.9.0 call the reference makes.
"""
from __future__ import annotations

from itertools import combinations

import numpy as np
import pandas as pd
from sklearn.cluster import MiniBatchKMeans
from sklearn.decomposition import PCA
from sklearn.ensemble import IsolationForest
from sklearn.feature_selection import mutual_info_classif, mutual_info_regression
from sklearn.metrics.pairwise import paired_distances
from sklearn.preprocessing import minmax_scale

THETA = 100.0  # scanpy default overdispersion (normalize_pearson_residuals(theta=100))
N_PCS = 50     # scanpy.settings.N_PCS


# ---------------------------------------------------------------------------------------------------------
# a1  pp.normalize, sc branch                       scGeneClust/pp/_preprocessing.py:15-29
# ---------------------------------------------------------------------------------------------------------
def pearson_residuals(X: np.ndarray, theta: float = THETA) -> np.ndarray:
    """Analytic Pearson residuals, dense path of scanpy 1.9 `_pearson_residuals` (PARITY UNPINNED: scanpy is
    not installed; restated from the published algorithm).  `X` is cells x genes; the result keeps X's dtype
    (float32 for 10x data), clip = sqrt(n_cells)."""
    X = np.asarray(X)
    dt = X.dtype if X.dtype.kind == "f" else np.dtype(np.float32)
    X = X.astype(dt, copy=False)
    clip = dt.type(np.sqrt(X.shape[0]))
    sums_genes = np.sum(X, axis=0, keepdims=True)
    sums_cells = np.sum(X, axis=1, keepdims=True)
    sum_total = np.sum(sums_genes)
    mu = np.array(sums_cells @ sums_genes / sum_total)
    diff = np.array(X - mu)
    residuals = diff / np.sqrt(mu + mu ** 2 / dt.type(theta))
    return np.clip(residuals, -clip, clip).astype(dt, copy=False)


def pearson_residuals_block(X_block: np.ndarray, sums_cells: np.ndarray, sums_genes: np.ndarray, sum_total: float,
                            n_cells_total: int, theta: float = THETA) -> np.ndarray:
    """The same formula for a sub-block of a larger matrix whose sums are given (`sums_cells`: the block's cells,
    `sums_genes`: the block's genes, both over the FULL matrix; clip = sqrt(number of cells of the full matrix)).
    Lets the full-size tests check sampled blocks of a matrix that never exists on the host."""
    X = np.asarray(X_block, dtype=np.float32)
    clip = np.float32(np.sqrt(n_cells_total))
    mu = np.asarray(sums_cells, dtype=np.float32).reshape(-1, 1) @ np.asarray(sums_genes, dtype=np.float32).reshape(1, -1)
    mu = mu / np.float32(sum_total)
    residuals = (X - mu) / np.sqrt(mu + mu ** 2 / np.float32(theta))
    return np.clip(residuals, -clip, clip).astype(np.float32, copy=False)


# ---------------------------------------------------------------------------------------------------------
# a2  pp.reduce_dim                                  scGeneClust/pp/_preprocessing.py:35-55
# ---------------------------------------------------------------------------------------------------------
def _sc_pca(data: np.ndarray, random_state: int) -> np.ndarray:
    """`sc.pp.pca(ndarray, svd_solver='auto', random_state=...)` for dense input (scanpy 1.9
    `preprocessing/_pca.py`; PARITY UNPINNED on the scanpy wrapper): n_comps = 50, or min(shape)-1 when
    min(shape) <= 50; zero-centred sklearn PCA with the solver string passed through; cast to float32."""
    min_dim = min(data.shape)
    n_comps = min_dim - 1 if N_PCS >= min_dim else N_PCS
    pca = PCA(n_components=n_comps, svd_solver="auto", random_state=random_state)
    return pca.fit_transform(data).astype(np.float32, copy=False)


def gene_pca(Z: np.ndarray, random_state: int = 0) -> np.ndarray:
    """`adata.varm['X_pca'] = sc.pp.pca(norm_counts.T, ...)` (pp/_preprocessing.py:52) -> genes x 50 float32."""
    return _sc_pca(np.asanyarray(Z).T, random_state)


def cell_pca(Z: np.ndarray, random_state: int = 0) -> np.ndarray:
    """`adata.obsm['X_pca'] = sc.pp.pca(norm_counts, ...)` (pp/_preprocessing.py:54), ps only."""
    return _sc_pca(np.asanyarray(Z), random_state)


# ---------------------------------------------------------------------------------------------------------
# a3 + a4  cluster_genes fast branch, compute_gene_closeness     scGeneClust/tl/cluster.py:65-71, 84-105
# ---------------------------------------------------------------------------------------------------------
def kmeans_genes(X_pca: np.ndarray, n_gene_clusters: int, random_state: int = 0):
    """MiniBatchKMeans(K, batch_size=max(1024, G//10), random_state, n_init='auto').fit_predict (cluster.py:66-70).
    Returns (labels int32[G], centers float32[K, d], n_steps)."""
    km = MiniBatchKMeans(n_clusters=n_gene_clusters, batch_size=max(1024, X_pca.shape[0] // 10),
                         random_state=random_state, n_init="auto")
    labels = km.fit_predict(X_pca)
    return labels, km.cluster_centers_, km.n_steps_


def gene_closeness(X_pca: np.ndarray, labels: np.ndarray, centers: np.ndarray) -> np.ndarray:
    """cluster.py:101-104: euclidean distance to the own centre, then 1 - minmax_scale within each cluster."""
    all_distances = paired_distances(X_pca, centers[labels])
    for gene_cluster in np.unique(labels):
        mask = labels == gene_cluster
        all_distances[mask] = 1 - minmax_scale(all_distances[mask])
    return all_distances


# ---------------------------------------------------------------------------------------------------------
# a5  select_from_clusters fast branch + compute_deviance        scGeneClust/tl/selection.py:46-60, 82-89
# ---------------------------------------------------------------------------------------------------------
def compute_deviance(X: np.ndarray) -> np.ndarray:
    """selection.py:82-89, verbatim arithmetic (binomial deviance of the *residual* columns; NaNs dropped)."""
    pi = X.sum(0) / X.sum()
    n = X.sum(1)[:, np.newaxis]
    with np.errstate(all="ignore"):
        left_half = np.nansum(X * np.log(X / n.dot(pi[np.newaxis, :])), axis=0)
        right_half = np.nansum((n - X) * (np.log((n - X) / (n.dot((1 - pi)[np.newaxis, :])))), axis=0)
    return 2 * (left_half + right_half)


def select_fast(var_names, labels, closeness, Z=None, post_hoc_filtering=True, random_state=0) -> np.ndarray:
    """selection.py:46-60.  Order of the result: max-closeness gene of each non-singleton cluster (ascending
    cluster id), then the min-closeness genes, then outlier genes among singleton clusters."""
    var_names = pd.Index(var_names)
    var = pd.DataFrame({"cluster": labels, "closeness": closeness}, index=var_names)
    counts = var["cluster"].value_counts()
    is_single = var["cluster"].isin(counts[counts == 1].index)
    if post_hoc_filtering and is_single.sum() > 0:
        single_genes = var_names[is_single.to_numpy()]
        deviances = compute_deviance(np.asarray(Z)[:, is_single.to_numpy()])
        is_outlier = IsolationForest(random_state=random_state).fit_predict(deviances.reshape(-1, 1)) == -1
        outlier_genes = single_genes[is_outlier]
    else:
        outlier_genes = np.array([])
    grouped = var.loc[~is_single, ["cluster", "closeness"]].rename_axis("gene").groupby(by="cluster")["closeness"]
    max_genes = grouped.nlargest(1).reset_index(level=1)["gene"].values
    min_genes = grouped.nsmallest(1).reset_index(level=1)["gene"].values
    return np.hstack((max_genes, min_genes, outlier_genes))


# ---------------------------------------------------------------------------------------------------------
# a6  find_relevant_genes                                 scGeneClust/tl/information.py:21-57
# ---------------------------------------------------------------------------------------------------------
def gene_relevance(Z: np.ndarray, labels, random_state: int = 0) -> np.ndarray:
    """information.py:21-24 per gene: mutual_info_classif(column, labels, discrete_features=False, seed)."""
    labels = np.asarray(labels)
    return np.array([
        mutual_info_classif(Z[:, i].reshape(-1, 1), labels, discrete_features=False, random_state=random_state)[0]
        for i in range(Z.shape[1])
    ])


def relevant_mask(relevance: np.ndarray, top_pct: int) -> np.ndarray:
    """information.py:50-52: strict > percentile(relevance, 100 - top_pct)."""
    return relevance > np.percentile(relevance, 100 - top_pct)


# ---------------------------------------------------------------------------------------------------------
# a7  compute_gene_redundancy                             scGeneClust/tl/information.py:60-85
# ---------------------------------------------------------------------------------------------------------
def redundancy_pair(E: np.ndarray, i: int, j: int, random_state: int = 0) -> float:
    """information.py:60-63 for one pair (rows i<j of varm['X_pca'])."""
    g1, g2 = E[i, :].reshape(-1, 1), E[j, :]
    return mutual_info_regression(g1, g2, discrete_features=False, random_state=random_state)[0]


def gene_redundancy(E: np.ndarray, random_state: int = 0, pairs=None) -> np.ndarray:
    """information.py:83-84.  With `pairs=None` returns the dense symmetric N x N float64 matrix
    (squareform of the upper triangle in itertools.combinations order); with an explicit (P, 2) list returns
    the P values."""
    from scipy.spatial.distance import squareform
    if pairs is None:
        vals = [redundancy_pair(E, i, j, random_state) for i, j in combinations(range(E.shape[0]), 2)]
        return squareform(vals)
    return np.array([redundancy_pair(E, int(i), int(j), random_state) for i, j in pairs])


# ---------------------------------------------------------------------------------------------------------
# a8  compute_gene_complementarity (per MST edge)         scGeneClust/tl/information.py:88-98
# ---------------------------------------------------------------------------------------------------------
def complementarity_edge(Z: np.ndarray, clusters, i: int, j: int, random_state: int = 0) -> float:
    clusters = np.asarray(clusters)
    cmi = 0
    for clus in np.unique(clusters):
        m = clusters == clus
        g1, g2 = Z[m, i].reshape(-1, 1), Z[m, j]
        rlv = mutual_info_regression(g1, g2, discrete_features=False, n_neighbors=3, random_state=random_state)[0]
        cmi += rlv * m.sum() / Z.shape[0]
    return cmi


def gene_complementarity(Z: np.ndarray, clusters, edges, random_state: int = 0) -> np.ndarray:
    return np.array([complementarity_edge(Z, clusters, int(i), int(j), random_state) for i, j in edges])


# ---------------------------------------------------------------------------------------------------------
# f2  high-confidence cells: mixture fits + co-membership            scGeneClust/tl/confidence.py:55-60, 82-115
# ---------------------------------------------------------------------------------------------------------
def cell_co_membership(labels: np.ndarray, probas: np.ndarray, p: float = 0.95) -> np.ndarray:
    """confidence.py:110-114 for one mixture fit: co[i, j > i] = (proba_i >= p) and (proba_j > p) and label_i == label_j."""
    n = labels.shape[0]
    co = np.zeros((n, n))
    for i in range(n - 1):
        if probas[i] >= p:
            co[i, i + 1:] = np.logical_and(probas[i + 1:] > p, labels[i + 1:] == labels[i])
    return co


def mixture_labels_probas(X_pca: np.ndarray, n_clusters: int, idx: int, random_state: int = 0):
    """confidence.py:98-104: the reference's GaussianMixture call on the first `idx` cell PCs."""
    from sklearn.mixture import GaussianMixture
    X = X_pca[:, :idx]
    gmm = GaussianMixture(n_components=n_clusters, init_params='k-means++', random_state=random_state).fit(X)
    return gmm.predict(X), gmm.predict_proba(X).max(1)


def co_membership_frequency(X_pca: np.ndarray, n_clusters: int, n_components: int = 10, random_state: int = 0):
    """confidence.py:55-60: sum of the co-membership matrices of the fits on PC prefixes 2 .. n_components + 1."""
    mats = [cell_co_membership(*mixture_labels_probas(X_pca, n_clusters, idx, random_state))
            for idx in range(2, 2 + n_components)]
    return np.sum(mats, axis=0)


# ---------------------------------------------------------------------------------------------------------
# whole fast path on the CPU (timed as the cpu_baseline / reference arm)     scGeneClust/_model.py:123-133
# ---------------------------------------------------------------------------------------------------------
def geneclust_fast(counts: np.ndarray, var_names, n_var_clusters: int, random_state: int = 0,
                   post_hoc_filtering: bool = True, timings: dict | None = None):
    import time
    t0 = time.perf_counter()
    Z = pearson_residuals(np.asarray(counts, dtype=np.float32))
    t1 = time.perf_counter()
    X_pca = gene_pca(Z, random_state)
    t2 = time.perf_counter()
    labels, centers, n_steps = kmeans_genes(X_pca, n_var_clusters, random_state)
    closeness = gene_closeness(X_pca, labels, centers)
    t3 = time.perf_counter()
    selected = select_fast(var_names, labels, closeness, Z, post_hoc_filtering, random_state)
    t4 = time.perf_counter()
    if timings is not None:
        timings.update(normalize=t1 - t0, pca=t2 - t1, kmeans=t3 - t2, select=t4 - t3, total=t4 - t0,
                       kmeans_steps=int(n_steps))
    return dict(Z=Z, X_pca=X_pca, labels=labels, centers=centers, closeness=closeness, selected=selected)


# ---------------------------------------------------------------------------------------------------------
# link-wise parity of a GeneClust-fast run (what the tests / smoke / bench assert; see tests/test_gpu_subsample.py)
# ---------------------------------------------------------------------------------------------------------
def fast_parity(gpu: dict, ref: dict, var_names, n_var_clusters: int, random_state: int = 0, Z=None) -> dict:
    """`gpu` / `ref`: dicts with X_pca, labels, closeness, selected (ref = this module's `geneclust_fast` result or a
    fixture of it).  Returns
      pca_rel_err_max / _first10 : link 1 (float): per-component relative error of the PCA scores;
      tail_exact                 : link 2 (exact): scikit-learn's MiniBatchKMeans + the reference's closeness / selection
                                   run on the CPU from the GPU's OWN float32 embedding reproduce the GPU's labels, closeness
                                   (1e-5) and - when the residuals `Z` are given or no singleton cluster needs them - list;
      labels_equal / selected_equal / ari : end-to-end agreement, REPORTED (needs all ~1400 k-means++ draws to coincide, which
                                   two float32 PCA implementations only achieve by chance - scripts/cpu_noise_floor.py)."""
    from sklearn.metrics import adjusted_rand_score
    rel = np.linalg.norm(gpu['X_pca'] - ref['X_pca'], axis=0) / np.linalg.norm(ref['X_pca'], axis=0)
    labels_cpu, centers_cpu, _ = kmeans_genes(np.ascontiguousarray(gpu['X_pca']), n_var_clusters, random_state)
    close_cpu = gene_closeness(np.ascontiguousarray(gpu['X_pca']), labels_cpu, centers_cpu)
    same_labels = bool(np.array_equal(labels_cpu, gpu['labels']))
    same_close = bool(np.allclose(close_cpu, gpu['closeness'], rtol=0, atol=1e-5))
    has_single = bool((np.bincount(labels_cpu, minlength=n_var_clusters) == 1).any())
    list_checked = Z is not None or not has_single
    same_list = True
    if list_checked:
        same_list = bool(np.array_equal(np.asarray(gpu['selected']).astype(str),
                                        select_fast(var_names, labels_cpu, np.asarray(gpu['closeness']), Z, True,
                                                    random_state).astype(str)))
    return {"pca_rel_err_max": float(rel.max()), "pca_rel_err_first10": float(rel[:10].max()),
            "tail_exact": bool(same_labels and same_close and same_list),
            "tail_labels_equal": same_labels, "tail_closeness_equal": same_close,
            "tail_list_equal": same_list if list_checked else None,
            "labels_equal": bool(np.array_equal(gpu['labels'], ref['labels'])),
            "ari": float(adjusted_rand_score(gpu['labels'], ref['labels'])),
            "selected_equal": bool(np.array_equal(np.asarray(gpu['selected']).astype(str), np.asarray(ref['selected']).astype(str))),
            "selected_common": int(len(set(np.asarray(gpu['selected']).astype(str)) & set(np.asarray(ref['selected']).astype(str)))),
            "n_selected": int(len(gpu['selected'])), "n_selected_ref": int(len(ref['selected']))}

