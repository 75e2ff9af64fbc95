"""The north_star subsample harness: the GPU path against the CPU oracle on
synth(20 000 cells, 20 000 genes) - the "cells[0:20 000] of the same counter-based stream" block that verifies the
inputs of the configurations too large for the CPU (c5).  The device-generated block must equal the oracle generator's
bit for bit, PCA scores agree within 1e-3 relative, and the gene clusters and the selected-gene list are THE SAME."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CELLS, GENES, K = 20_000, 20_000, 200


def test_subsample_20k_gpu_vs_cpu_oracle():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests")
    from sklearn.metrics import adjusted_rand_score
    from oracle import stages
    from oracle import synth as osy
    from scgeneclust_b200 import AnnDataLite, pp, scGeneClust, tl
    from scgeneclust_b200.synth import DeviceSynth, synth_params

    counts = DeviceSynth(synth_params(GENES, 0)).counts(0, CELLS)
    X = osy.synth_counts_c(osy.synth_params(GENES, 0), 0, CELLS, 0, GENES)
    assert np.array_equal(counts.to_numpy(), X), "device generator and oracle generator disagree"
    names = np.array([f"g{i}" for i in range(GENES)], dtype=object)
    ref = stages.geneclust_fast(X, names, K, 0)

    sel = scGeneClust(AnnDataLite(counts, var_names=names), n_var_clusters=K, version='fast', verbosity=0)
    ad = AnnDataLite(counts, var_names=names)
    pp.normalize(ad, 'sc')
    pp.reduce_dim(ad, 'fast', 0)
    tl.cluster_genes(ad, None, 'fast', n_gene_clusters=K, random_state=0)
    rel = np.linalg.norm(ad.varm['X_pca'] - ref['X_pca'], axis=0) / np.linalg.norm(ref['X_pca'], axis=0)
    labels = ad.var['cluster'].to_numpy()
    ari = adjusted_rand_score(labels, ref['labels'])
    print(f"20k x 20k subsample: PCA rel err max {rel.max():.2e}, ARI {ari:.4f}, labels identical "
          f"{np.array_equal(labels, ref['labels'])}, list identical {np.array_equal(sel, ref['selected'])}")
    assert rel.max() <= 1e-3, f"PCA rel err {rel.max():.2e}"
    assert np.array_equal(labels, ref['labels']), f"gene clusters differ (ARI {ari:.4f})"
    assert np.allclose(ad.var['closeness'].to_numpy(), ref['closeness'], rtol=0, atol=1e-4)
    assert np.array_equal(sel, ref['selected']), "selected-gene list differs from the CPU oracle's"
    # the same block from host float32 CSR (what an AnnData holds): identical result to the device-resident run
    import scipy.sparse as sp
    sel_csr = scGeneClust(AnnDataLite(sp.csr_matrix(X[:, :].astype(np.float32)), var_names=names), n_var_clusters=K,
                          version='fast', verbosity=0)
    assert np.array_equal(sel_csr, sel)
