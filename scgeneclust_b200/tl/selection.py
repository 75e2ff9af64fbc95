"""Representative-gene selection with the reference's name and semantics (scGeneClust/tl/selection.py:14-89).
G-length vector work: stays on the host (numpy), except the residual columns of singleton-cluster genes, which are
produced on the device from the resident counts."""
from __future__ import annotations

from typing import Literal

import numpy as np
from loguru import logger
from sklearn.ensemble import IsolationForest

from ..pp._preprocessing import get_state


def compute_deviance(X: np.ndarray) -> np.ndarray:
    """Binomial deviance of each column (scGeneClust/tl/selection.py:82-89).  The reference applies it to the
    *residual* columns, so logs of negative numbers give NaNs that `nansum` drops; kept as is."""
    pi = X.sum(0) / X.sum()
    n = X.sum(1)[:, np.newaxis]
    with np.errstate(all='ignore'):
        left_half = np.nansum(X * np.log(X / n.dot(pi[np.newaxis, :])), axis=0)
        right_half = np.nansum((n - X) * (np.log((n - X) / (n.dot((1 - pi)[np.newaxis, :])))), axis=0)
    return 2 * (left_half + right_half)


def _residual_columns(adata, mask: np.ndarray) -> np.ndarray:
    """Pearson-residual columns of the masked genes as a host float32 (cells x n_masked) array."""
    st = get_state(adata)
    if st.Z is not None:
        import torch
        idx = torch.from_numpy(np.flatnonzero(mask)).to(st.Z.device)
        return st.Z.index_select(1, idx).cpu().numpy()
    if st.op is not None:
        return st.op.materialize_columns(np.flatnonzero(mask)).cpu().numpy()
    return np.asarray(adata.X)[:, mask]


def _isolation_forest_outliers(values: np.ndarray, random_state: int) -> np.ndarray:
    """`IsolationForest(random_state=...).fit_predict(values.reshape(-1, 1)) == -1` (scGeneClust/tl/selection.py:51).

    With one or two samples the forest's answer does not depend on its random draws, so it is written down instead
    of fitting 100 trees (~0.1 s of Python overhead): max_samples_ = n, so every tree holds all samples; for n = 1
    the normaliser c(1) = 0 makes every score exactly -0.5, for n = 2 every tree isolates both samples at depth 1
    (or keeps two equal values in one leaf, path length c(2) = 1) and every score is -2^(-1) = -0.5 as well.
    `offset_` is -0.5 (contamination='auto'), the decision function is 0 and `predict` flags only values < 0:
    no outlier.  (sklearn 1.9.0 ensemble/_iforest.py:297-380, :529-640.)  Non-finite input other than NaN makes
    sklearn raise, so that case - and every n >= 3 - still goes through sklearn itself."""
    if values.shape[0] <= 2 and not np.isinf(values).any():
        return np.zeros(values.shape[0], dtype=bool)
    return IsolationForest(random_state=random_state).fit_predict(values.reshape(-1, 1)) == -1


def _first_extremes_per_cluster(cluster: np.ndarray, value: np.ndarray):
    """Positions of the first maximum and of the first minimum of `value` inside every cluster, ascending cluster id -
    what `groupby('cluster')[col].nlargest(1)` / `.nsmallest(1)` (keep='first') pick.  One stable sort by cluster
    (original order inside a cluster) serves both."""
    order = np.argsort(cluster, kind='stable')
    c, v = cluster[order], value[order]
    starts = np.flatnonzero(np.r_[True, c[1:] != c[:-1]])
    seg = np.repeat(np.arange(len(starts)), np.diff(np.r_[starts, len(c)]))
    out = []
    for reduce_ in (np.maximum, np.minimum):
        ext = reduce_.reduceat(v, starts)
        hit = np.flatnonzero(v == ext[seg])                       # sorted positions attaining the extreme
        out.append(order[hit[np.r_[True, seg[hit][1:] != seg[hit][:-1]]]])
    return out[0], out[1]


def _first_extreme_per_cluster(cluster: np.ndarray, value: np.ndarray, largest: bool) -> np.ndarray:
    first_max, first_min = _first_extremes_per_cluster(cluster, value)
    return first_max if largest else first_min


def select_from_clusters(adata, version: Literal['fast', 'ps'], post_hoc_filtering: bool = True,
                         random_state: int = 0) -> np.ndarray:
    """Same contract as the reference: returns the ndarray of selected gene names.
    fast: max-closeness gene of every non-singleton cluster (ascending cluster id), then the min-closeness genes,
    then IsolationForest outliers among singleton-cluster genes.  ps: top-relevance gene per dense cluster plus
    GLOSH-thresholded low-density genes."""
    cluster = adata.var['cluster'].to_numpy()
    var_names = np.asarray(adata.var_names)
    if version == 'fast':
        ids, counts = np.unique(cluster, return_counts=True)
        is_single = np.isin(cluster, ids[counts == 1])
        if post_hoc_filtering and (n_single := int(is_single.sum())) > 0:
            single_genes = var_names[is_single]
            deviances = compute_deviance(_residual_columns(adata, is_single))
            is_outlier = _isolation_forest_outliers(deviances, random_state)
            logger.opt(colors=True).debug(
                f"Found <yellow>{is_outlier.sum()}</yellow> outlier(s) in <yellow>{n_single}</yellow> single gene cluster(s).")
            outlier_genes = single_genes[is_outlier]
        else:
            outlier_genes = np.array([])
        pos = np.flatnonzero(~is_single)
        closeness = adata.var['closeness'].to_numpy()[pos]
        first_max, first_min = _first_extremes_per_cluster(cluster[pos], closeness)
        max_genes, min_genes = var_names[pos[first_max]], var_names[pos[first_min]]
        return np.hstack((max_genes, min_genes, outlier_genes))
    is_low_density = cluster == -1
    if (n_low := int(is_low_density.sum())) > 0:
        logger.opt(colors=True).debug(f"Found <yellow>{n_low}</yellow> low density genes.")
        adata.var['representative'] = False
        pos = np.flatnonzero(~is_low_density)
        relevance = adata.var['relevance'].to_numpy()[pos]
        top = pos[_first_extreme_per_cluster(cluster[pos], relevance, True)]
        representative = np.zeros(len(cluster), bool)
        representative[top] = True
        if post_hoc_filtering:
            score = adata.var['outlier_score'].to_numpy()
            dense_scores = score[~is_low_density]
            threshold = np.median(dense_scores[dense_scores > 0]) if (dense_scores > 0).any() else np.nan
            representative[is_low_density] = score[is_low_density] >= threshold
        adata.var['representative'] = representative
    # like the reference, this raises KeyError when there is no low-density gene (var['representative'] is unset)
    return var_names[adata.var['representative'].to_numpy()]
