"""CPU oracle for the GeneClust gene-grouping hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under `scgeneclust_b200/` may import this package.  It is imported by `tests/`, by
`__graft_entry__.smoke()` and by the `cpu_baseline` / `--impl reference` legs of `bench.py`, always as the
checker or as the timed CPU reference, never as the thing shipped.

What it is: a restatement of the reference's CPU algorithm for the path SURVEY.md §8(a) lists, built from the
*same third-party calls the reference makes* (scikit-learn 1.9.0 `PCA`, `MiniBatchKMeans`,
`mutual_info_classif`, `mutual_info_regression`, `IsolationForest`; pandas groupby) plus numpy restatements of
the pieces whose source is not in the container (scanpy's analytic Pearson residuals) and brute-force fp64
restatements of the two kNN mutual-information estimators that expose their integer neighbour counts.

Pinning status (see DESIGN.md "Oracle"):
  * a3/a4/a5/a6/a7/a8 (k-means + closeness, fast selection, relevance, redundancy, complementarity) are PINNED:
    `tests/golden/make_golden.py` executes the reference's own functions from /root/reference (with the
    import-time-only dependencies anndata/igraph/hdbscan stubbed) on seeded inputs and the fixtures under
    `tests/golden/` are their outputs; `tests/test_oracle_golden.py` checks this package against them.
  * a1 (scanpy `normalize_pearson_residuals`) and the scanpy side of a2 (`sc.pp.pca` wrapper: n_comps rule,
    float32 cast) are "PARITY UNPINNED": scanpy is absent, the formula is restated from its published
    algorithm (Lause et al. 2021; scanpy 1.9 `experimental/pp/_normalization.py`).  The sklearn side of a2 is
    the installed `sklearn.decomposition.PCA` itself.
"""
