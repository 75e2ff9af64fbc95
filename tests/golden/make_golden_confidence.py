"""Golden fixture for the high-confidence-cell stage (scGeneClust/tl/confidence.py:82-115), produced by executing the
REFERENCE'S OWN `_compute_cell_co_membership` (import-time-only dependencies stubbed, see make_golden.py):

    python tests/golden/make_golden_confidence.py      ->  tests/golden/confidence_small.npz

Stored: the cells x PCs matrix, the co-membership matrices the reference returns for the PC prefixes 2 .. 5 (summed:
the "frequency matrix" of confidence.py:58-60), and the labels / max posterior of the identical
`GaussianMixture(n_components, init_params='k-means++', random_state)` fits (the generator asserts that rebuilding the
co-membership from them reproduces the reference's matrices bit for bit), so the device kernel is pinned independently
of BLAS-level float differences in the mixture fits between machines.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, versions  # noqa: E402

import_reference()
import scGeneClust.tl.confidence as ref_conf  # noqa: E402
from sklearn.mixture import GaussianMixture  # noqa: E402

rs = np.random.RandomState(21)
C, d, K, seed = 700, 12, 5, 0
centers = rs.standard_normal((K, d)) * np.array([6, 5, 4, 3, 2, 2, 1, 1, 1, 1, 1, 1])
types_ = rs.randint(0, K, C)
X_pca = (centers[types_] + rs.standard_normal((C, d)) * 1.3).astype(np.float32)
ref_conf.X_pca = X_pca
prefixes = list(range(2, 6))
co = [ref_conf._compute_cell_co_membership(idx, n_clusters=K, random_state=seed) for idx in prefixes]
labels, probas = [], []
for idx, ref_mat in zip(prefixes, co):
    gmm = GaussianMixture(n_components=K, init_params='k-means++', random_state=seed).fit(X_pca[:, :idx])
    lab, pr = gmm.predict(X_pca[:, :idx]), gmm.predict_proba(X_pca[:, :idx]).max(1)
    rebuilt = np.zeros((C, C))
    for i in range(C - 1):
        if pr[i] >= 0.95:
            rebuilt[i, i + 1:] = np.logical_and(pr[i + 1:] > 0.95, lab[i + 1:] == lab[i])
    assert np.array_equal(rebuilt, ref_mat), idx
    labels.append(lab.astype(np.int32))
    probas.append(pr)
freq = np.sum(co, axis=0)
print("frequency matrix: nonzero", int((freq > 0).sum()), "max", freq.max(), "confident per run",
      [int((p >= 0.95).sum()) for p in probas])
np.savez_compressed(os.path.join(HERE, "confidence_small.npz"), X_pca=X_pca, n_clusters=K, seed=seed,
                    prefixes=np.array(prefixes), labels=np.stack(labels), probas=np.stack(probas),
                    frequency=freq.astype(np.uint8), p=0.95, versions=versions())
