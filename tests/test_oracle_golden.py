"""The CPU oracle against the golden vectors produced by the reference's own functions
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np

from oracle import mi_bruteforce as mb
from oracle import stages


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_redundancy_matches_reference(golden_dir):
    g = _load(golden_dir, "redundancy_small.npz")
    E, seed, ref = g["E"], int(g["seed"]), g["redundancy"]
    N, n = E.shape
    # sklearn-calling restatement
    got = stages.gene_redundancy(E, seed)
    assert np.array_equal(got, ref)
    # brute-force restatement (exposes the integer counts the GPU kernel is checked on)
    xi1, xi2 = mb.noise_vectors(n, seed)
    XR = np.stack([mb.prep_feature_explicit(E[i], xi1) for i in range(N)])
    YR = np.stack([mb.prep_target_explicit(E[i], xi2) for i in range(N)])
    iu = np.triu_indices(N, 1)
    pairs = np.stack(iu, 1)
    nx, ny, mi = mb.ksg_batch(XR, YR, pairs)
    assert np.array_equal(mi, ref[iu])
    assert nx.min() >= 2 and ny.min() >= 0 and nx.max() <= n - 1
    # reference invariants: symmetric, zero diagonal, non-negative (tests/test_GeneClust.py:23)
    assert np.array_equal(ref, ref.T) and not ref.diagonal().any() and (ref >= 0).all()


def test_relevance_matches_reference(golden_dir):
    g = _load(golden_dir, "relevance_small.npz")
    Z, labels, seed = g["Z"], g["labels"], int(g["seed"])
    ref = g["relevance_all"]
    assert np.array_equal(stages.gene_relevance(Z, labels, seed), ref)
    xi1, _ = mb.noise_vectors(Z.shape[0], seed, continuous_target=False)
    brute = np.array([mb.ross_mi(mb.prep_feature_explicit(Z[:, j], xi1), labels) for j in range(Z.shape[1])])
    assert np.array_equal(brute, ref)
    mask = stages.relevant_mask(ref, int(g["top_pct"]))
    assert np.array_equal(np.array([f"gene_{i}" for i in np.flatnonzero(mask)]), g["kept_genes"])
    assert np.array_equal(ref[mask], g["kept_relevance"])


def test_complementarity_matches_reference(golden_dir):
    g = _load(golden_dir, "complementarity_small.npz")
    Z, labels, seed, edges, ref = g["Z"], g["labels"], int(g["seed"]), g["edges"], g["complementarity"]
    assert np.array_equal(stages.gene_complementarity(Z, labels, edges, seed), ref)
    # brute force: per cluster slice, fresh noise of the slice length
    out = []
    for i, j in edges:
        cmi = 0
        for clus in np.unique(labels):
            m = labels == clus
            xi1, xi2 = mb.noise_vectors(int(m.sum()), seed)
            x = mb.prep_feature_explicit(Z[m, i], xi1)
            y = mb.prep_target_explicit(Z[m, j], xi2)
            cmi += mb.ksg_mi(x, y) * m.sum() / Z.shape[0]
        out.append(cmi)
    assert np.array_equal(np.array(out), ref)


def test_fast_cluster_and_selection_match_reference(golden_dir):
    g = _load(golden_dir, "fast_cluster_select_small.npz")
    X_pca, K, seed = g["X_pca"], int(g["n_gene_clusters"]), int(g["seed"])
    labels, centers, _ = stages.kmeans_genes(X_pca, K, seed)
    assert np.array_equal(labels, g["labels"])
    closeness = stages.gene_closeness(X_pca, labels, centers)
    assert np.array_equal(closeness, g["closeness"])
    # closeness invariants (cluster.py:104): in [0, 1] up to one float32 ulp of minmax_scale's x*scale+min form
    assert closeness.min() >= -2e-7 and closeness.max() <= 1
    sel = stages.select_fast(g["var_names"], labels, closeness, g["Zres"], True, seed)
    assert np.array_equal(sel.astype(str), g["selected_posthoc"])
    sel = stages.select_fast(g["var_names"], labels, closeness, None, False, seed)
    assert np.array_equal(sel.astype(str), g["selected_plain"])
    assert np.allclose(stages.compute_deviance(g["Zres"][:, :8]), g["deviance_first8"], rtol=0, atol=0,
                       equal_nan=True)


def test_cell_co_membership_matches_reference_golden(golden_dir):
    """f2: the oracle's co-membership restatement against the reference's own `_compute_cell_co_membership`
    (tests/golden/make_golden_confidence.py), from the stored mixture labels / posteriors."""
    import os
    import numpy as np
    from oracle import stages
    g = np.load(os.path.join(golden_dir, "confidence_small.npz"), allow_pickle=False)
    freq = np.sum([stages.cell_co_membership(lab, pr, float(g["p"])) for lab, pr in zip(g["labels"], g["probas"])], axis=0)
    assert np.array_equal(freq, g["frequency"].astype(np.float64))
    assert not np.tril(freq).any()
