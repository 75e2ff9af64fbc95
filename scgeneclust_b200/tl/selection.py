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


def device_deviance(Xd) -> np.ndarray:
    """`compute_deviance` of a device-resident float32 (cells x m) residual block, on the device
    (gc_binomial_deviance: numpy's float32 operation and summation order); returns the host float32[m] vector."""
    import torch
    from .._lib import lib, ptr, stream_handle
    L = lib()
    Xd = Xd.contiguous()
    C_, m = Xd.shape
    out = torch.empty(m, dtype=torch.float32, device=Xd.device)
    scratch = torch.empty(L.binomial_deviance_scratch_bytes(C_, m), dtype=torch.uint8, device=Xd.device)
    L.binomial_deviance(ptr(Xd), C_, m, ptr(out), ptr(scratch), stream_handle())
    return out.cpu().numpy()


def _singleton_deviances(adata, mask: np.ndarray) -> np.ndarray:
    """Deviance of the masked (singleton-cluster) genes: on the device when the residuals live there (no C x m
    host copy, no host log pass), else the reference's numpy expression on `adata.X`."""
    st = get_state(adata)
    if st.Z is not None:
        import torch
        idx = torch.from_numpy(np.flatnonzero(mask)).to(st.Z.device)
        return device_deviance(st.Z.index_select(1, idx))
    if st.op is not None:
        return device_deviance(st.op.materialize_columns(np.flatnonzero(mask)))
    return compute_deviance(np.asarray(adata.X)[:, mask])


_RAND_R_MAX = 2147483647
_EULER_GAMMA = float(np.euler_gamma)


def _avg_path_length(n: int) -> float:
    """`_average_path_length` of sklearn ensemble/_iforest.py:647-681 for one leaf size."""
    if n <= 1:
        return 0.0
    if n == 2:
        return 1.0
    return float(2.0 * (np.log(n - 1.0) + _EULER_GAMMA) - 2.0 * (n - 1.0) / n)


def _first_randint31(seeds: np.ndarray) -> np.ndarray:
    """`[RandomState(s).randint(2**31 - 1) for s in seeds]` (equally `.randint(0, 2**31 - 1)`), vectorised over the
    seeds: MT19937 `init_genrand(s)`, the first regenerated word (needs init words 0, 1 and 397 only), tempering,
    then numpy's masked rejection for a 31-bit range (keep `word & 0x7fffffff` unless it equals 2**31 - 1 - the one
    value in 2**31 that needs a second word; those seeds go through RandomState itself)."""
    seeds = np.asarray(seeds, dtype=np.uint64)
    mt_prev = seeds & np.uint64(0xFFFFFFFF)
    mt0 = mt_prev.copy()
    mt1 = mt397 = None
    for i in range(1, 398):
        mt_prev = (np.uint64(1812433253) * (mt_prev ^ (mt_prev >> np.uint64(30))) + np.uint64(i)) & np.uint64(0xFFFFFFFF)
        if i == 1:
            mt1 = mt_prev
        if i == 397:
            mt397 = mt_prev
    yv = (mt0 & np.uint64(0x80000000)) | (mt1 & np.uint64(0x7FFFFFFF))
    word = mt397 ^ (yv >> np.uint64(1)) ^ np.where(yv & np.uint64(1), np.uint64(0x9908B0DF), np.uint64(0))
    word ^= word >> np.uint64(11)
    word ^= (word << np.uint64(7)) & np.uint64(0x9D2C5680)
    word ^= (word << np.uint64(15)) & np.uint64(0xEFC60000)
    word ^= word >> np.uint64(18)
    out = (word & np.uint64(0x7FFFFFFF)).astype(np.int64)
    for k in np.flatnonzero(out == 0x7FFFFFFF):                       # rejected draw (probability 2**-31)
        out[k] = np.random.RandomState(int(seeds[k])).randint(np.iinfo(np.int32).max)
    return out


def _isolation_forest_1d(values: np.ndarray, random_state: int, n_estimators: int = 100,
                         return_scores: bool = False) -> np.ndarray:
    """`IsolationForest(random_state=seed).fit_predict(values.reshape(-1, 1)) == -1` for at most 256 finite values,
    restated from scikit-learn 1.9.0 so that the ~0.7 ms of Python overhead sklearn spends per tree (clone, parameter
    validation, 100 estimator objects) is not paid for a handful of scalars:

    * ensemble/_bagging.py:398-520: seeds = RandomState(seed).randint(2**31 - 1, size=100); tree i gets
      random_state = RandomState(seeds[i]).randint(2**31 - 1) (ensemble/_base.py:49-99); with max_samples =
      min(256, n) = n and bootstrap=False every tree sees every sample (unit weights);
    * tree/_splitter.pyx Splitter.__cinit__: rand_r state = RandomState(tree seed).randint(0, 2**31 - 1);
    * tree/_tree.pyx:140-330 depth-first build, left child first, max_depth = ceil(log2(max(n, 2))); a node with >= 2
      samples, depth < max_depth and impure y draws (tree/_splitter.pyx:507-700, one feature): rand_int(0, 1) [one
      xorshift step, utils/_random.pxd:20-34], then, unless max <= min + 1e-7 in float32, threshold =
      (max - min) * rand / 2147483647 + min in float64 (tree/_utils.pyx:57-61), reset to min when it equals max;
      samples with value <= threshold go left;
    * ensemble/_iforest.py:529-640: depth of a sample in a tree = node depth + c(leaf size); score =
      -2 ** (-sum(depths) / (100 c(n))); outlier iff score - (-0.5) < 0.

    The regression target y = RandomState(seed).uniform(size=n) (ensemble/_iforest.py:350) only enters through the
    "impurity <= eps" leaf test, which is evaluated here with the plain variance formula.
    Checked against sklearn itself (flags and score_samples) on random inputs in tests/test_host_logic.py."""
    x32 = np.asarray(values, dtype=np.float64).astype(np.float32)
    n = x32.shape[0]
    xs = [float(v) for v in x32]                      # float32 values as Python floats (exact)
    ys = np.random.RandomState(random_state).uniform(size=n).tolist()
    max_depth = int(np.ceil(np.log2(max(n, 2))))
    seeds = np.random.RandomState(random_state).randint(np.iinfo(np.int32).max, size=n_estimators)
    states = _first_randint31(_first_randint31(seeds)).tolist()
    eps = float(np.finfo(np.float64).eps)
    c_of = [_avg_path_length(m) for m in range(n + 1)]
    depths = [0.0] * n
    for state in states:
        contrib = [0.0] * n
        stack = [(list(range(n)), 0)]                 # (sample ids, depth); left child is pushed last = built first
        while stack:
            ids, depth = stack.pop()
            m = len(ids)
            leaf = depth >= max_depth or m < 2
            if not leaf:
                s1 = s2 = 0.0
                for i in ids:
                    s1 += ys[i]
                    s2 += ys[i] * ys[i]
                leaf = s2 / m - (s1 / m) ** 2 <= eps
            if not leaf:
                # rand_int(0, 1): the (only) feature is drawn - one xorshift step of our_rand_r
                if state == 0:
                    state = 1
                state ^= (state << 13) & 0xFFFFFFFF
                state ^= state >> 17
                state ^= (state << 5) & 0xFFFFFFFF
                lo = hi = xs[ids[0]]
                for i in ids:
                    v = xs[i]
                    if v < lo:
                        lo = v
                    elif v > hi:
                        hi = v
                if hi - lo < 1e-6 and np.float32(hi) <= np.float32(lo) + np.float32(1e-7):
                    leaf = True                       # constant feature (float32 test of the splitter): no split
                else:
                    if state == 0:
                        state = 1
                    state ^= (state << 13) & 0xFFFFFFFF
                    state ^= state >> 17
                    state ^= (state << 5) & 0xFFFFFFFF
                    thr = (hi - lo) * float(state & 0x7FFFFFFF) / 2147483647.0 + lo
                    if thr == hi:
                        thr = lo
                    left, right = [], []
                    for i in ids:
                        (left if xs[i] <= thr else right).append(i)
                    stack.append((right, depth + 1))
                    stack.append((left, depth + 1))
            if leaf:
                d = (depth + 1) + c_of[m] - 1.0
                for i in ids:
                    contrib[i] = d
        for i in range(n):
            depths[i] += contrib[i]
    depths = np.asarray(depths)
    denominator = n_estimators * c_of[n]
    scores = -(2 ** (-np.divide(depths, denominator, out=np.ones_like(depths), where=denominator != 0)))
    if return_scores:
        return scores
    return scores - (-0.5) < 0


def _isolation_forest_1d_native(values: np.ndarray, random_state: int, n_estimators: int = 100,
                                return_scores: bool = False) -> np.ndarray:
    """The same restatement as `_isolation_forest_1d`, with the forest built by the C++ host routine
    `gc_host_iforest_depths` of libgeneclust_b200.so (n <= 256 samples, seeds < 2**32); the path-length table and the
    final score formula stay in numpy so that every transcendental is numpy's."""
    from .._lib import lib
    x32 = np.ascontiguousarray(np.asarray(values, dtype=np.float64).astype(np.float32))
    n = x32.shape[0]
    c_of = np.array([_avg_path_length(m) for m in range(n + 1)], dtype=np.float64)
    depths = np.empty(n, dtype=np.float64)
    lib().host_iforest_depths(x32.ctypes.data, n, int(random_state), n_estimators, c_of.ctypes.data, depths.ctypes.data)
    denominator = n_estimators * c_of[n]
    scores = -(2 ** (-np.divide(depths, denominator, out=np.ones_like(depths), where=denominator != 0)))
    if return_scores:
        return scores
    return scores - (-0.5) < 0


def _isolation_forest_outliers(values: np.ndarray, random_state: int) -> np.ndarray:
    """`IsolationForest(random_state=...).fit_predict(values.reshape(-1, 1)) == -1` (scGeneClust/tl/selection.py:51).

    With one or two samples the forest's answer does not depend on its random draws: max_samples_ = n, so every tree
    holds all samples; for n = 1 the normaliser c(1) = 0 makes every score exactly -0.5, for n = 2 every tree isolates
    both samples at depth 1 (or keeps two equal values in one leaf, path length c(2) = 1) and every score is
    -2^(-1) = -0.5 as well; `offset_` is -0.5, the decision function is 0 and `predict` flags only values < 0: no
    outlier.  Up to 256 finite values (sklearn's max_samples: every tree still sees every sample) use the line-by-line
    restatement above in its native form; anything else (more values, NaN / inf,
    which make sklearn take its missing-value path or raise) goes through sklearn itself."""
    n = values.shape[0]
    finite = bool(np.isfinite(values).all())
    if n <= 2 and not np.isinf(values).any():
        return np.zeros(n, dtype=bool)
    if n <= 256 and finite and bool(np.isfinite(values.astype(np.float32)).all()) and 0 <= random_state < 2 ** 32:
        return _isolation_forest_1d_native(values, random_state)
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


def _select_fast_device(adata, fc, cluster, var_names, post_hoc_filtering, random_state) -> np.ndarray:
    """GeneClust-fast selection (scGeneClust/tl/selection.py:46-60) with the per-cluster reductions on the device
    (gc_cluster_extremes: cluster sizes, first arg-max / arg-min of the closeness): the host only maps 2 K indices to
    gene names.  Same output order: max genes by ascending cluster id, min genes, outliers among singleton clusters."""
    import torch
    from .._lib import lib, ptr, stream_handle
    labels, closeness, K = fc
    out = torch.empty(3 * K, dtype=torch.int32, device=labels.device)
    scratch = torch.empty(2 * K, dtype=torch.int32, device=labels.device)
    lib().cluster_extremes(ptr(closeness), ptr(labels), labels.numel(), K, ptr(out), ptr(scratch), stream_handle())
    o = out.cpu().numpy()
    counts, first_max, first_min = o[:K], o[K:2 * K], o[2 * K:]
    multi = counts > 1
    single_pos = np.sort(first_max[counts == 1])                   # a singleton's only gene (ascending gene order)
    if post_hoc_filtering and len(single_pos) > 0:
        is_single = np.zeros(len(cluster), dtype=bool)
        is_single[single_pos] = True
        deviances = _singleton_deviances(adata, is_single)
        is_outlier = _isolation_forest_outliers(deviances, random_state)
        logger.opt(colors=True).debug(
            f"Found <yellow>{is_outlier.sum()}</yellow> outlier(s) in <yellow>{len(single_pos)}</yellow> single gene cluster(s).")
        outlier_genes = var_names[single_pos][is_outlier]
    else:
        outlier_genes = np.array([])
    return np.hstack((var_names[first_max[multi]], var_names[first_min[multi]], outlier_genes))


def select_from_clusters(adata, version: Literal['fast', 'ps'], post_hoc_filtering: bool = True,
                         random_state: int = 0) -> np.ndarray:
    """Same contract as the reference: returns the ndarray of selected gene names.
    fast: max-closeness gene of every non-singleton cluster (ascending cluster id), then the min-closeness genes,
    then IsolationForest outliers among singleton-cluster genes.  ps: top-relevance gene per dense cluster plus
    GLOSH-thresholded low-density genes."""
    cluster = adata.var['cluster'].to_numpy()
    var_names = np.asarray(adata.var_names)
    if version == 'fast':
        st = get_state(adata)
        fc = st.fast_clusters
        if (fc is not None and fc[0].numel() == len(cluster) and np.array_equal(cluster, fc[3])
                and np.array_equal(adata.var['closeness'].to_numpy(), fc[4])):    # var columns untouched since clustering
            return _select_fast_device(adata, fc[:3], cluster, var_names, post_hoc_filtering, random_state)
        ids, counts = np.unique(cluster, return_counts=True)
        is_single = np.isin(cluster, ids[counts == 1])
        if post_hoc_filtering and (n_single := int(is_single.sum())) > 0:
            single_genes = var_names[is_single]
            deviances = _singleton_deviances(adata, is_single)
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
