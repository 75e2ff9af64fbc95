"""High-confidence cells (scGeneClust/tl/confidence.py:30-148) - the stage that supplies the pseudo cell types
`obs['cluster']` GeneClust-ps scores genes against.

What is built here:
  * the ten Gaussian-mixture fits on prefixes of the cell PCs (confidence.py:58-59, 102-104) on the device:
    k-means++ seeding by this library's exact restatement of scikit-learn's `kmeans_plusplus` (gc_kmeanspp_init, the same
    routine GeneClust-fast uses), EM in float64 with `torch.linalg` as the small-matrix library (Cholesky of <= 11 x 11
    covariances) - checked against `sklearn.mixture.GaussianMixture` to float tolerance (the reference runs it in
    float32, so identical posteriors are not defined);
  * the cells x cells co-membership frequency matrix (confidence.py:60, 104-115) as one integer kernel
    (gc_comembership_counts) - bit-exact against the reference's own `_compute_cell_co_membership`
    (tests/golden/confidence_small.npz);
  * the frequency-threshold sweep (confidence.py:61-76) around a `leiden` partitioner.
What is NOT: Leiden itself.  The reference calls leidenalg / igraph (confidence.py:118-148); neither is installable in the
build image, so the sweep uses them when they can be imported and otherwise asks for the pseudo cell types as an INPUT
(`adata.obs['cluster']`, optionally a boolean `adata.obs['highly_confident']`) - that path is what the parity tests and
the bench use.  Leiden parity is unpinned.
"""
from __future__ import annotations

import math
import os
from typing import Callable, Optional

import numpy as np
import torch
from loguru import logger

from .._lib import lib, ptr, require_cuda, stream_handle
from ..pp._preprocessing import get_state

P_CONFIDENT = 0.95          # confidence.py:82 default `p`


# -------------------------------------------------------------------------------------------------------------
# Gaussian mixture (sklearn.mixture.GaussianMixture(K, init_params='k-means++', random_state) as confidence.py:102 calls it:
# covariance_type='full', tol=1e-3, reg_covar=1e-6, max_iter=100, n_init=1)
# -------------------------------------------------------------------------------------------------------------
def _kmeanspp_indices(X32: torch.Tensor, K: int, rs: np.random.RandomState) -> torch.Tensor:
    """`kmeans_plusplus(X, K, random_state=rs)[1]` (sklearn cluster/_kmeans.py:70-275) on the device: the draws from `rs`
    are the ones scikit-learn makes (first centre by `choice`, then K - 1 rounds of 2 + log K uniforms)."""
    L, st = lib(), stream_handle()
    n, d = X32.shape
    n_trials = 2 + int(np.log(K))
    sw = np.ones(n, dtype=np.float32)              # kmeans_plusplus: sample_weight=None -> ones of X's dtype
    first = int(rs.choice(n, p=sw / sw.sum()))     # (_kmeans.py:225: choice with p, i.e. the cdf / random_sample path)
    uniforms = rs.uniform(size=(K - 1, n_trials)) if K > 1 else np.zeros((0, n_trials))
    d_uni = torch.from_numpy(np.ascontiguousarray(uniforms)).to(X32.device)
    ws = torch.empty(L.kmeanspp_scratch_bytes(n, n_trials), dtype=torch.uint8, device=X32.device)
    centers = torch.empty((K, d), dtype=torch.float32, device=X32.device)
    indices = torch.empty(K, dtype=torch.int32, device=X32.device)
    L.kmeanspp_init(ptr(X32), X32.stride(0), n, d, K, first, ptr(d_uni), n_trials, ptr(ws), ptr(centers), ptr(indices), st)
    return indices.to(torch.int64)


def _m_step(X: torch.Tensor, resp: torch.Tensor, reg_covar: float, init: bool = False):
    """`_estimate_gaussian_parameters` + `_compute_precision_cholesky` (mixture/_gaussian_mixture.py:270-340); weights are
    nk / n at initialisation (`_initialize`) and nk / sum(nk) in `_m_step`."""
    n, d = X.shape
    nk = resp.sum(0) + 10 * torch.finfo(resp.dtype).eps
    means = (resp.t() @ X) / nk.unsqueeze(1)
    diff = X.unsqueeze(0) - means.unsqueeze(1)                       # K x n x d
    cov = torch.einsum('kn,kni,knj->kij', resp.t(), diff, diff) / nk.view(-1, 1, 1)
    cov = cov + reg_covar * torch.eye(d, dtype=X.dtype, device=X.device)
    chol = torch.linalg.cholesky(cov)                                # lower
    eye = torch.eye(d, dtype=X.dtype, device=X.device).expand_as(chol)
    prec_chol = torch.linalg.solve_triangular(chol, eye, upper=False).transpose(1, 2)
    return (nk / n if init else nk / nk.sum()), means, prec_chol


def _log_resp(X: torch.Tensor, weights, means, prec_chol):
    """`_estimate_log_prob_resp`: weighted log probabilities, their logsumexp and the log responsibilities."""
    d = X.shape[1]
    log_det = torch.log(torch.diagonal(prec_chol, dim1=1, dim2=2)).sum(1)
    y = torch.einsum('ni,kij->knj', X, prec_chol) - torch.einsum('ki,kij->kj', means, prec_chol).unsqueeze(1)
    log_prob = -0.5 * (d * math.log(2 * math.pi) + (y * y).sum(2)) + log_det.unsqueeze(1)     # K x n
    weighted = (log_prob + torch.log(weights).unsqueeze(1)).t()                                # n x K
    norm = torch.logsumexp(weighted, dim=1)
    return norm, weighted - norm.unsqueeze(1)


def gaussian_mixture_fit_predict(X32: torch.Tensor, n_components: int, random_state: int, tol: float = 1e-3,
                                 reg_covar: float = 1e-6, max_iter: int = 100):
    """Labels and maximum posterior of every row of X32 (cells x dims, float32, device) under the fitted mixture - what
    `gmm.predict(X)` and `gmm.predict_proba(X).max(1)` return at confidence.py:103-104."""
    require_cuda()
    X32 = X32.contiguous()
    n = X32.shape[0]
    rs = np.random.RandomState(random_state)
    idx = _kmeanspp_indices(X32, n_components, rs)
    X = X32.to(torch.float64)
    resp = torch.zeros((n, n_components), dtype=torch.float64, device=X.device)
    resp[idx, torch.arange(n_components, device=X.device)] = 1.0    # _initialize_parameters, init_params='k-means++'
    weights, means, prec_chol = _m_step(X, resp, reg_covar, init=True)
    lower = -math.inf
    for _ in range(max_iter):
        prev = lower
        norm, log_resp = _log_resp(X, weights, means, prec_chol)
        weights, means, prec_chol = _m_step(X, torch.exp(log_resp), reg_covar)
        lower = float(norm.mean().item())
        if abs(lower - prev) < tol:
            break
    _, log_resp = _log_resp(X, weights, means, prec_chol)
    proba, labels = torch.exp(log_resp).max(dim=1)
    return labels.to(torch.int32), proba


# -------------------------------------------------------------------------------------------------------------
# co-membership frequency matrix
# -------------------------------------------------------------------------------------------------------------
def co_membership_frequency(labels: torch.Tensor, probas: torch.Tensor, p: float = P_CONFIDENT) -> torch.Tensor:
    """Sum over the R fits of the reference's co-membership matrices (confidence.py:60, 111-114): uint8 (C, C), entry
    (i, j > i) = number of fits with proba_i >= p, proba_j > p and label_i == label_j.  labels: R x C int32, probas:
    R x C float64, both on the device."""
    require_cuda()
    L, st = lib(), stream_handle()
    R, C_ = labels.shape
    labels = labels.to(torch.int32).contiguous()
    probas = probas.to(torch.float64).contiguous()
    freq = torch.empty((C_, C_), dtype=torch.uint8, device=labels.device)
    scratch = torch.empty(2 * R * C_, dtype=torch.int32, device=labels.device)
    L.comembership_counts(ptr(labels), ptr(probas), C_, R, float(p), ptr(freq), ptr(scratch), st)
    return freq


def mixture_co_membership(X_pca: torch.Tensor, n_cell_clusters: int, n_components: int = 10, random_state: int = 0):
    """confidence.py:55-60: one mixture per PC prefix 2 .. n_components + 1, then the frequency matrix.  Returns
    (freq uint8 C x C, labels R x C, probas R x C) on the device."""
    labs, prs = [], []
    for idx in range(2, 2 + n_components):
        lab, pr = gaussian_mixture_fit_predict(X_pca[:, :idx], n_cell_clusters, random_state)
        labs.append(lab)
        prs.append(pr)
    labels, probas = torch.stack(labs), torch.stack(prs)
    return co_membership_frequency(labels, probas), labels, probas


# -------------------------------------------------------------------------------------------------------------
# threshold sweep (confidence.py:61-76) around a Leiden partitioner
# -------------------------------------------------------------------------------------------------------------
def _leidenalg_partition(adjacency: np.ndarray, seed: int) -> np.ndarray:
    """confidence.py:118-148 verbatim in behaviour: weighted undirected graph from the upper-triangular adjacency,
    leidenalg RBConfigurationVertexPartition, resolution 1, n_iterations=-1."""
    import igraph as ig
    import leidenalg
    G = ig.Graph.Weighted_Adjacency(adjacency, mode="upper")
    partition = leidenalg.find_partition(G, partition_type=leidenalg.RBConfigurationVertexPartition,
                                         weights=G.es['weight'], n_iterations=-1, resolution_parameter=1.0, seed=seed)
    return np.array(partition.membership)


def sweep_frequency_threshold(freq: np.ndarray, n_cell_clusters: int, n_obs: int, seed: int,
                              leiden: Callable[[np.ndarray, int], np.ndarray]):
    """confidence.py:62-74: lower the frequency cut-off from 8 until more than 10 % of the cells sit in clusters of
    at least ten cells.  Returns (is_confident bool[C], cluster_labels int[C], freq_th)."""
    import pandas as pd
    is_confident = cluster_labels = None
    freq_th = 0
    for freq_th in range(8, 0, -1):
        cut_matrix = np.where(freq < freq_th, 0, freq)
        cluster_labels = leiden(cut_matrix, seed)
        cluster_counts = pd.Series(cluster_labels).value_counts(ascending=False)
        if (cluster_counts < 10).any():
            cut_k = min(n_cell_clusters, np.argwhere((cluster_counts < 10).values).squeeze(axis=1)[0])
        else:
            cut_k = cluster_counts.shape[0]
        is_confident = np.isin(cluster_labels, cluster_counts.index[:cut_k])
        if is_confident.sum() / n_obs > 0.1:
            break
    return is_confident, cluster_labels, freq_th


def _subset_cells(adata, st, keep: np.ndarray):
    from .._anndata_lite import DeviceMatrix, ShardedMatrix
    adata._inplace_subset_obs(keep)
    if isinstance(adata.X, (DeviceMatrix, ShardedMatrix)):
        st.Z = adata.X.tensor                    # the subset already happened on the device
    elif st.Z is not None:
        st.Z = st.Z.index_select(0, torch.from_numpy(np.flatnonzero(keep)).to(st.Z.device)).contiguous()


def find_high_confidence_cells(adata, n_obs_clusters: int, n_components: int = 10,
                               max_workers: int = os.cpu_count() - 1, random_state: int = 0,
                               leiden: Optional[Callable[[np.ndarray, int], np.ndarray]] = None):
    """Same name, arguments and AnnData side effects as the reference (confidence.py:30-80): subsets `adata` to the
    high-confidence cells and writes their pseudo cell types to `obs['cluster']`.  If `obs['cluster']` is already there
    (and optionally a boolean `obs['highly_confident']`), those are used as given."""
    st = get_state(adata)
    if 'cluster' in adata.obs:
        keep = np.asarray(adata.obs['highly_confident'], dtype=bool) if 'highly_confident' in adata.obs \
            else np.ones(adata.n_obs, bool)
        adata.obs['cluster'] = np.asarray(adata.obs['cluster']).astype(str)
        if not keep.all():
            _subset_cells(adata, st, keep)
        logger.opt(colors=True).info(
            f"Using <yellow>{adata.n_obs}</yellow> provided high-confidence cells in "
            f"<yellow>{len(np.unique(adata.obs['cluster']))}</yellow> pseudo cell types.")
        return
    logger.info("Finding high-confidence cells...")
    if leiden is None:
        try:
            import igraph  # noqa: F401
            import leidenalg  # noqa: F401
            leiden = _leidenalg_partition
        except Exception as exc:  # noqa: BLE001
            raise NotImplementedError(
                "GeneClust-ps needs pseudo cell types: either provide `adata.obs['cluster']` (and optionally a boolean "
                "`adata.obs['highly_confident']`), or install leidenalg + igraph - the Gaussian-mixture fits and the "
                "co-membership frequency matrix of the reference's finder run on the GPU (tl.confidence."
                "mixture_co_membership), its Leiden step is the third-party library") from exc
    X_pca = st.obsm_pca if st.obsm_pca is not None else \
        torch.from_numpy(np.ascontiguousarray(adata.obsm['X_pca'], dtype=np.float32)).cuda()
    n_obs = adata.n_obs
    freq, _, _ = mixture_co_membership(X_pca, n_obs_clusters, n_components, random_state)
    is_confident, cluster_labels, freq_th = sweep_frequency_threshold(freq.cpu().numpy().astype(np.float64), n_obs_clusters,
                                                                      n_obs, random_state, leiden)
    logger.opt(colors=True).debug(f"Final frequency cutoff: <yellow>{freq_th}</yellow>")
    _subset_cells(adata, st, is_confident)
    adata.obs['cluster'] = cluster_labels.astype(str)[is_confident]
    logger.opt(colors=True).info(
        f"Found <yellow>{adata.n_obs}</yellow> (<yellow>{np.round(adata.n_obs / n_obs * 100)}%</yellow>) high-confidence cells.")
