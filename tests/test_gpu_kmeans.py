"""GPU parity of the GeneClust-fast gene k-means (a3), closeness (a4) and selection (a5) against the golden
vectors produced by the reference's `cluster_genes('fast')` / `select_from_clusters('fast')`.
Bar: labels identical where the float32 GEMM summation order allows it, otherwise ARI (north_star); closeness
within 1e-6 absolute given the same labels/centres; selected-gene list identical."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests")
    return torch


def test_assign_and_update_kernels_vs_numpy(torch_cuda):
    torch = torch_cuda
    from scgeneclust_b200._lib import lib, ptr, stream_handle
    rs = np.random.RandomState(0)
    n, d, K = 5000, 50, 37
    X = rs.standard_normal((n, d)).astype(np.float32)
    Cn = X[rs.choice(n, K, replace=False)] + 0.01
    dX, dC = torch.from_numpy(X).cuda(), torch.from_numpy(Cn).cuda()
    labels = torch.empty(n, dtype=torch.int32, device="cuda")
    sq = torch.empty(n, dtype=torch.float32, device="cuda")
    L, st = lib(), stream_handle()
    L.kmeans_assign(ptr(dX), d, None, n, ptr(dC), K, d, ptr(labels), ptr(sq), st)
    D = ((X[:, None, :].astype(np.float64) - Cn[None].astype(np.float64)) ** 2).sum(-1)
    lab_ref = D.argmin(1)
    lab = labels.cpu().numpy()
    gap = np.sort(D, 1)
    ambiguous = (gap[:, 1] - gap[:, 0]) < 1e-4 * gap[:, 0]
    assert np.array_equal(lab[~ambiguous], lab_ref[~ambiguous])
    assert np.allclose(sq.cpu().numpy(), D[np.arange(n), lab], rtol=1e-5)
    # minibatch update == sequential float32 running mean in sample order (_k_means_minibatch.pyx:60-108)
    w = rs.randint(0, 5, K).astype(np.float32)
    dW = torch.from_numpy(w.copy()).cuda()
    newC = torch.empty_like(dC)
    L.kmeans_minibatch_update(ptr(dX), d, None, n, ptr(labels), ptr(dC), ptr(newC), ptr(dW), K, d, st)
    ref = Cn.copy()
    w_ref = w.copy()
    for k in range(K):
        idx = np.flatnonzero(lab == k)
        if len(idx):
            acc = Cn[k] * w[k]
            for s in idx:
                acc = acc + X[s]
            w_ref[k] = w[k] + np.float32(len(idx))
            ref[k] = acc * (np.float32(1.0) / w_ref[k])
    assert np.array_equal(newC.cpu().numpy(), ref)
    assert np.array_equal(dW.cpu().numpy(), w_ref)


@pytest.mark.parametrize("kind,n", [("normal", 3000), ("normal", 7000), ("grid", 6000), ("grid", 7000), ("dups", 6000)])
def test_kmeanspp_matches_sklearn(torch_cuda, kind, n):
    """Device k-means++ draws the same centre indices as sklearn's `_kmeans_plusplus` from the same RandomState.
    The candidate draw is `searchsorted(cumsum(closest_dist_sq), u * pot)` on a sequential float32 cumsum, which the
    round kernel evaluates speculatively (kmeans_kernels.cu): 'grid' (half-integer coordinates: squared distances with
    few mantissa bits whose running sum passes 2^24, so round-half-to-even ties are frequent) and 'dups' (duplicated
    rows: exact zeros, equal values) stress its parity / binade logic; n = 7000 spans two staged chunks."""
    torch = torch_cuda
    from sklearn.cluster._kmeans import _kmeans_plusplus
    from sklearn.utils.extmath import row_norms
    from scgeneclust_b200.tl.cluster import DeviceMiniBatchKMeans
    rs = np.random.RandomState(4)
    if kind == "normal":
        X = (rs.standard_normal((n, 50)) * (3.0 / np.sqrt(1 + np.arange(50)))).astype(np.float32)
    elif kind == "grid":
        X = (rs.randint(-64, 65, (n, 50)) * 0.5).astype(np.float32)
    else:
        base = (rs.standard_normal((n // 4, 50)) * 2.0).astype(np.float32)
        X = base[rs.randint(0, len(base), n)]
    K = 60
    _, idx_ref = _kmeans_plusplus(X, K, row_norms(X, squared=True), np.ones(n, np.float32),
                                  np.random.RandomState(7))
    km = DeviceMiniBatchKMeans(K, 1024, 7)
    centers = km._init_centroids(torch.from_numpy(X).cuda(), np.random.RandomState(7), n)
    idx = km.init_indices_.cpu().numpy()
    assert np.array_equal(idx, idx_ref)
    assert np.array_equal(centers.cpu().numpy(), X[idx_ref])


def _ari(a, b):
    from sklearn.metrics import adjusted_rand_score
    return adjusted_rand_score(a, b)


def test_fast_clustering_golden(torch_cuda, golden_dir):
    torch = torch_cuda
    from scgeneclust_b200 import AnnDataLite
    from scgeneclust_b200.tl import cluster_genes, select_from_clusters
    g = np.load(os.path.join(golden_dir, "fast_cluster_select_small.npz"), allow_pickle=False)
    ad = AnnDataLite(g["Zres"].copy(), var_names=g["var_names"].astype(object))
    ad.varm["X_pca"] = g["X_pca"]
    cluster_genes(ad, None, "fast", n_gene_clusters=int(g["n_gene_clusters"]), random_state=int(g["seed"]))
    labels = ad.var["cluster"].to_numpy()
    ari = _ari(labels, g["labels"])
    # golden = the reference's own cluster_genes('fast') + select_from_clusters (tests/golden/make_golden.py): the gene
    # clusters, the closeness and the selected-gene list must be THE SAME, not merely similar
    assert np.array_equal(labels, g["labels"]), f"k-means labels differ from the reference's (ARI {ari:.6f})"
    assert np.allclose(ad.var["closeness"].to_numpy(), g["closeness"], rtol=0, atol=2e-6)
    sel = select_from_clusters(ad, "fast", True, int(g["seed"]))
    assert np.array_equal(sel.astype(str), g["selected_posthoc"])


def test_closeness_given_labels(torch_cuda, golden_dir):
    torch = torch_cuda
    from oracle import stages
    from scgeneclust_b200.tl.cluster import compute_gene_closeness
    g = np.load(os.path.join(golden_dir, "fast_cluster_select_small.npz"), allow_pickle=False)
    X, labels = g["X_pca"], g["labels"].astype(np.int32)
    K = int(g["n_gene_clusters"])
    centers = np.stack([X[labels == k].mean(0) for k in range(K)]).astype(np.float32)
    ref = stages.gene_closeness(X, labels, centers)
    got = compute_gene_closeness(torch.from_numpy(X).cuda(), torch.from_numpy(labels).cuda(),
                                 torch.from_numpy(centers).cuda()).cpu().numpy()
    assert np.allclose(got, ref, rtol=0, atol=2e-6)
    # one gene at closeness 1 per cluster; singleton clusters map to 1 (cluster.py:104)
    for k in range(K):
        assert got[labels == k].max() == 1.0


def test_kmeans_full_size_ari(torch_cuda):
    """Config-c3 size of the k-means stage (G = 20 000 genes x 50 PCs, K = 200) against sklearn on the same input."""
    torch = torch_cuda
    from oracle import stages
    from scgeneclust_b200.tl.cluster import DeviceMiniBatchKMeans
    rs = np.random.RandomState(1)
    G, K = 20000, 200
    centers = rs.standard_normal((150, 50)) * 3
    X = ((centers[rs.randint(0, 150, G)] + rs.standard_normal((G, 50))) / np.sqrt(1 + np.arange(50))).astype(np.float32)
    ref_labels, _, ref_steps = stages.kmeans_genes(X, K, 0)
    km = DeviceMiniBatchKMeans(K, max(1024, G // 10), 0)
    labels = km.fit_predict(torch.from_numpy(X).cuda()).cpu().numpy()
    ari = _ari(labels, ref_labels)
    assert km.n_steps_ == ref_steps, f"mini-batch steps {km.n_steps_} vs sklearn {ref_steps} (ARI {ari:.4f})"
    assert np.array_equal(labels, ref_labels), f"k-means labels differ from sklearn's at config-c3 size (ARI {ari:.4f})"
