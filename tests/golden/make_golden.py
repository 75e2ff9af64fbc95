"""Generate the golden fixtures in this directory by executing the REFERENCE'S OWN functions.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py

The reference package cannot be imported as-is (anndata, scanpy, squidpy, SpaGCN, igraph, leidenalg, hdbscan,
cv2 are not installed).  Those modules are needed only at import time for the functions pinned here (type
annotations, loaders, the MST and HDBSCAN tails), so they are replaced by empty stub modules and the
reference's stage functions are then run unmodified on seeded inputs wrapped in `AnnDataLite`:

  tl/information.py : _compute_relevance, find_relevant_genes, _compute_redundancy, compute_gene_redundancy,
                      _compute_complementarity
  tl/cluster.py     : cluster_genes(version='fast') (MiniBatchKMeans + compute_gene_closeness)
  tl/selection.py   : select_from_clusters(version='fast'), compute_deviance

`pp.normalize` / `pp.reduce_dim` call scanpy and cannot be executed; they stay "parity unpinned".
Library versions that produced the fixtures are stored in each file (`versions`).
"""
import importlib
import os
import sys
import types

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from scgeneclust_b200._anndata_lite import AnnDataLite  # noqa: E402


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__path__ = []  # behave as a package so `import a.b` works
    sys.modules[name] = mod
    return mod


def import_reference():
    _stub("anndata", AnnData=AnnDataLite)
    for name in ("scanpy", "squidpy", "SpaGCN", "igraph", "leidenalg", "cv2",
                 "hdbscan", "hdbscan._hdbscan_linkage", "hdbscan._hdbscan_tree"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:  # noqa: BLE001
                _stub(name)
    sys.modules["hdbscan._hdbscan_linkage"].__dict__.setdefault("label", None)
    for fn in ("condense_tree", "compute_stability", "get_clusters", "outlier_scores"):
        sys.modules["hdbscan._hdbscan_tree"].__dict__.setdefault(fn, None)
    import scGeneClust.tl.cluster as ref_cluster
    import scGeneClust.tl.information as ref_info
    import scGeneClust.tl.selection as ref_sel
    return ref_info, ref_cluster, ref_sel


def versions():
    import scipy
    import sklearn
    return np.array([f"numpy={np.__version__}", f"scipy={scipy.__version__}", f"sklearn={sklearn.__version__}",
                     f"pandas={pd.__version__}"])


def structured_embedding(rs, G, d=50, n_groups=12):
    """Gene embedding with cluster structure and decaying component scales (looks like PCA scores)."""
    centers = rs.standard_normal((n_groups, d)) * 4.0
    scales = (1.0 / np.sqrt(1.0 + np.arange(d)))[None, :]
    grp = rs.randint(0, n_groups, G)
    return ((centers[grp] + rs.standard_normal((G, d))) * scales).astype(np.float32)


def main():
    ref_info, ref_cluster, ref_sel = import_reference()
    from loguru import logger
    logger.remove()
    seed = 0

    # ---- a7 redundancy: reference compute_gene_redundancy on a 24 x 50 float32 embedding ------------------
    rs = np.random.RandomState(11)
    E = structured_embedding(rs, 24)
    E[5, 10:14] = E[5, 3]          # duplicated coordinates inside one gene
    E[6] = np.round(E[6] * 2) / 2  # heavy ties
    E[7] = E[2]                    # identical genes
    ad = AnnDataLite(np.zeros((3, 24), np.float32))
    ad.varm["X_pca"] = E
    ref_info.compute_gene_redundancy(ad, max_workers=2, random_state=seed)
    np.savez_compressed(os.path.join(HERE, "redundancy_small.npz"), E=E, seed=seed,
                        redundancy=ad.varp["redundancy"], versions=versions())

    # ---- a6 relevance: reference _compute_relevance per gene + find_relevant_genes ------------------------
    rs = np.random.RandomState(12)
    n_hc, G = 400, 60
    types_ = rs.randint(0, 6, n_hc)
    labels = np.array([str(t) for t in types_], dtype=object)
    labels[types_ == 5] = "5"
    labels[:1] = "lonely"           # a unique label -> sample dropped by the estimator
    Z = rs.standard_normal((n_hc, G)).astype(np.float32)
    Z[:, :20] += (types_[:, None] * rs.standard_normal((1, 20))).astype(np.float32)  # informative genes
    Z[:, 20] = np.round(Z[:, 20])   # ties
    Z[:, 21] = -0.25                # constant column (std -> 1 branch)
    Z[:, 22] = np.where(rs.rand(n_hc) < 0.8, -0.1, Z[:, 22])  # zero-inflated-like
    ad = AnnDataLite(Z.copy())
    ad.obs["cluster"] = labels
    ref_info.expr_mtx, ref_info.clusters, ref_info.seed = ad.X, ad.obs["cluster"], seed
    relevance_all = np.array([ref_info._compute_relevance(i) for i in range(G)])
    ref_info.find_relevant_genes(ad, 20, max_workers=2, random_state=seed)
    np.savez_compressed(os.path.join(HERE, "relevance_small.npz"), Z=Z, labels=labels.astype(str), seed=seed,
                        top_pct=20, relevance_all=relevance_all, kept_genes=np.asarray(ad.var_names, dtype=str),
                        kept_relevance=ad.var["relevance"].to_numpy(), versions=versions())

    # ---- a8 complementarity: reference _compute_complementarity on explicit edges -------------------------
    edges = np.array([(0, 1), (2, 17), (20, 3), (22, 5), (7, 40), (59, 58)])
    ad = AnnDataLite(Z.copy())
    lab2 = np.array([str(t) for t in types_], dtype=object)  # no unique label: every cluster >= 10 cells
    ref_info.expr_mtx, ref_info.clusters, ref_info.seed = ad.X, pd.Series(lab2), seed
    complm = np.array([ref_info._compute_complementarity((int(i), int(j))) for i, j in edges])
    np.savez_compressed(os.path.join(HERE, "complementarity_small.npz"), Z=Z, labels=lab2.astype(str), seed=seed,
                        edges=edges, complementarity=complm, versions=versions())

    # ---- a3 + a4 + a5: reference cluster_genes('fast') and select_from_clusters('fast') -------------------
    rs = np.random.RandomState(13)
    G, C = 2500, 40
    X_pca = structured_embedding(rs, G, n_groups=30)
    # far-away genes in different directions -> singleton clusters (exercises the IsolationForest rescue)
    X_pca[:6] = 0.0
    for g, (dim, val) in enumerate([(0, 300.0), (0, -300.0), (1, 300.0), (1, -300.0), (2, 300.0), (2, -300.0)]):
        X_pca[g, dim] = val
    Zres = rs.standard_normal((C, G)).astype(np.float32) * 2.0
    names = np.array([f"g{i}" for i in range(G)], dtype=object)
    ad = AnnDataLite(Zres.copy(), var_names=names)
    ad.varm["X_pca"] = X_pca
    ref_cluster.cluster_genes(ad, None, "fast", n_gene_clusters=40, random_state=seed)
    labels_km = ad.var["cluster"].to_numpy()
    closeness = ad.var["closeness"].to_numpy()
    sel_posthoc = ref_sel.select_from_clusters(ad, "fast", True, seed)
    sel_plain = ref_sel.select_from_clusters(ad, "fast", False, seed)
    sizes = np.bincount(labels_km, minlength=40)
    print("fast: cluster sizes min/max", sizes.min(), sizes.max(), "singletons", int((sizes == 1).sum()),
          "selected", len(sel_posthoc), len(sel_plain))
    np.savez_compressed(os.path.join(HERE, "fast_cluster_select_small.npz"), X_pca=X_pca, Zres=Zres,
                        var_names=names.astype(str), n_gene_clusters=40, seed=seed, labels=labels_km,
                        closeness=closeness, selected_posthoc=np.asarray(sel_posthoc, dtype=str),
                        selected_plain=np.asarray(sel_plain, dtype=str),
                        deviance_first8=ref_sel.compute_deviance(Zres[:, :8]), versions=versions())
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
