"""High-confidence-cell stage (scGeneClust/tl/confidence.py) on the device: the co-membership frequency kernel is pinned
bit for bit by the reference's own `_compute_cell_co_membership` (tests/golden/confidence_small.npz); the device Gaussian
mixture is compared with scikit-learn's to float tolerance (the reference fits in float32: identical posteriors are not
defined); Leiden is third-party and absent, so the threshold sweep is exercised with a stand-in partitioner."""
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


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "confidence_small.npz"), allow_pickle=False)


def test_comembership_kernel_equals_reference(torch_cuda, gold):
    torch = torch_cuda
    from scgeneclust_b200.tl.confidence import co_membership_frequency
    freq = co_membership_frequency(torch.from_numpy(gold["labels"]).cuda(), torch.from_numpy(gold["probas"]).cuda(),
                                   float(gold["p"]))
    assert freq.dtype == torch.uint8
    assert np.array_equal(freq.cpu().numpy(), gold["frequency"])
    # thresholds are asymmetric (>= p for the row cell, > p for the column cell): a posterior exactly at p
    lab = np.zeros((1, 5), np.int32)
    pr = np.array([[0.95, 0.95, 0.96, 0.94, 1.0]])
    f = co_membership_frequency(torch.from_numpy(lab).cuda(), torch.from_numpy(pr).cuda(), 0.95).cpu().numpy()
    expect = np.zeros((5, 5), np.uint8)
    expect[0, [2, 4]] = 1          # row 0 passes (>=), column 1 does not (> fails at exactly p)
    expect[1, [2, 4]] = 1
    expect[2, 4] = 1
    assert np.array_equal(f, expect)


def test_device_mixture_close_to_sklearn(torch_cuda, gold):
    torch = torch_cuda
    from sklearn.metrics import adjusted_rand_score
    from scgeneclust_b200.tl.confidence import gaussian_mixture_fit_predict
    X = gold["X_pca"]
    K, seed = int(gold["n_clusters"]), int(gold["seed"])
    for idx, lab_ref, pr_ref in zip(gold["prefixes"], gold["labels"], gold["probas"]):
        lab, pr = gaussian_mixture_fit_predict(torch.from_numpy(np.ascontiguousarray(X[:, :idx])).cuda(), K, seed)
        lab, pr = lab.cpu().numpy(), pr.cpu().numpy()
        ari = adjusted_rand_score(lab, lab_ref)
        same_conf = np.mean((pr >= 0.95) == (pr_ref >= 0.95))
        print(f"GMM on {idx} PCs: ARI vs sklearn {ari:.4f}, |dp| median {np.median(np.abs(pr - pr_ref)):.1e}, "
              f"same confident status {same_conf:.4f}")
        assert ari >= 0.98
        assert same_conf >= 0.98
        assert np.median(np.abs(pr - pr_ref)) <= 1e-3


def test_frequency_from_device_fits_and_threshold_sweep(torch_cuda, gold):
    """End to end without Leiden: frequency matrix from the device fits agrees with the reference's on almost every
    pair, and `find_high_confidence_cells` with a stand-in partitioner (connected components of the cut graph) subsets
    the AnnData and writes string labels like the reference does."""
    torch = torch_cuda
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import connected_components
    from scgeneclust_b200 import AnnDataLite
    from scgeneclust_b200.tl.confidence import find_high_confidence_cells, mixture_co_membership
    X = gold["X_pca"]
    K, seed = int(gold["n_clusters"]), int(gold["seed"])
    freq, labels, probas = mixture_co_membership(torch.from_numpy(X).cuda(), K, n_components=len(gold["prefixes"]),
                                                 random_state=seed)
    diff = np.mean(freq.cpu().numpy() != gold["frequency"])
    print(f"frequency matrix from device fits: {100 * diff:.3f} % of the cell pairs differ from the reference's")
    assert diff <= 0.02

    def components(adjacency, seed_):
        return connected_components(csr_matrix(adjacency), directed=False)[1]

    ad = AnnDataLite(np.zeros((X.shape[0], 7), np.float32))
    ad.obsm['X_pca'] = X
    find_high_confidence_cells(ad, K, n_components=4, random_state=seed, leiden=components)
    assert 0.1 * X.shape[0] < ad.n_obs <= X.shape[0]
    assert ad.obs['cluster'].to_numpy().dtype.kind in "UO" and len(np.unique(ad.obs['cluster'])) <= K
    with pytest.raises(NotImplementedError, match="pseudo cell types"):
        find_high_confidence_cells(AnnDataLite(np.zeros((X.shape[0], 7), np.float32)), K)
