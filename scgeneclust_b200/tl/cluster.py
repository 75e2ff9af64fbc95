"""Gene clustering stage with the reference's names and semantics (scGeneClust/tl/cluster.py:22-131), running on
the device.  GeneClust-fast = mini-batch k-means on `varm['X_pca']` + per-cluster closeness; GeneClust-ps =
relevance -> redundancy -> MST + complementarity -> HDBSCAN-style gene clusters."""
from __future__ import annotations

import os
from typing import Literal, Optional

import numpy as np
import torch
from loguru import logger

from .._lib import lib, ptr, require_cuda, stream_handle
from ..pp._preprocessing import get_state


_PINNED: dict = {}


def _pinned(name: str, n: int, dtype) -> torch.Tensor:
    """Process-wide pinned staging buffers (cudaHostAlloc costs ~0.1 ms a call, more than a mini-batch step)."""
    key = (name, int(n), dtype)
    if key not in _PINNED:
        _PINNED[key] = torch.empty(int(n), dtype=dtype).pin_memory()
    return _PINNED[key]


_CDF: dict = {}


def _uniform_cdf(n: int, dev) -> torch.Tensor:
    """The float64 cdf `RandomState.choice(n, size, p=ones(n, float32) / n)` searches (numpy mtrand.pyx: cumsum of p,
    normalised by its last entry), on the device.  It depends on n only; kept for the process (160 KB at 20 000)."""
    key = (int(n), str(dev))
    if key not in _CDF:
        p = np.ones(n, dtype=np.float32)
        p = p / np.sum(p)
        cdf = np.asarray(p, dtype=np.float64).cumsum()
        cdf /= cdf[-1]
        if len(_CDF) > 8:
            _CDF.clear()
        _CDF[key] = torch.from_numpy(cdf).to(dev)
    return _CDF[key]


class DeviceMiniBatchKMeans:
    """`sklearn.cluster.MiniBatchKMeans(n_clusters, batch_size, random_state, n_init='auto')` as the reference
    configures it (scGeneClust/tl/cluster.py:66-69), re-expressed for the device.

    The control flow and every draw from `RandomState` follow scikit-learn 1.9.0 `MiniBatchKMeans.fit`
    (cluster/_kmeans.py:2056-2224): validation-set `randint`, init-subset `randint`, k-means++ `choice`/`uniform`,
    per-step `choice(n, batch, p)`, `choice(replace=False)` on random reassignment; defaults max_iter=100,
    max_no_improvement=10, reassignment_ratio=0.01, tol=0, init='k-means++', n_init=1.  The arithmetic (distances,
    argmin, running means, potentials) runs in kernels of libgeneclust_b200.so; the host keeps only the RNG, the
    K-length count vector and the early-stopping scalars.
    """

    def __init__(self, n_clusters: int, batch_size: int = 1024, random_state: int = 0, max_iter: int = 100,
                 max_no_improvement: int = 10, reassignment_ratio: float = 0.01):
        self.n_clusters, self.batch_size, self.random_state = int(n_clusters), int(batch_size), random_state
        self.max_iter, self.max_no_improvement, self.reassignment_ratio = max_iter, max_no_improvement, reassignment_ratio

    # -- k-means++ on a (sub)sample, cluster/_kmeans.py:180-275 ------------------------------------------------
    def _init_centroids(self, X: torch.Tensor, rs: np.random.RandomState, init_size: int) -> torch.Tensor:
        L, st = lib(), stream_handle()
        n, d = X.shape
        K = self.n_clusters
        if init_size < n:                                    # _init_centroids, cluster/_kmeans.py:1015-1020
            init_indices = rs.randint(0, n, init_size)
            Xs = X.index_select(0, torch.from_numpy(init_indices).to(X.device)).contiguous()
        else:
            Xs = X
        ns = Xs.shape[0]
        n_trials = 2 + int(np.log(K))
        sw = np.ones(ns, dtype=np.float32)
        first = int(rs.choice(ns, p=sw / sw.sum()))
        # one call draws the same stream as the K-1 per-round `uniform(size=n_trials)` calls of _kmeans_plusplus
        uniforms = rs.uniform(size=(K - 1, n_trials)) if K > 1 else np.zeros((0, n_trials))
        d_uni = torch.from_numpy(np.ascontiguousarray(uniforms)).to(X.device)
        ws = torch.empty(L.kmeanspp_scratch_bytes(ns, n_trials), dtype=torch.uint8, device=X.device)
        centers = torch.empty((K, d), dtype=torch.float32, device=X.device)
        indices = torch.empty(K, dtype=torch.int32, device=X.device)
        L.kmeanspp_init(ptr(Xs), Xs.stride(0), ns, d, K, first, ptr(d_uni), n_trials, ptr(ws), ptr(centers),
                        ptr(indices), st)
        self.init_indices_ = indices
        return centers

    def fit_predict(self, X: torch.Tensor) -> torch.Tensor:
        require_cuda()
        L, st = lib(), stream_handle()
        assert X.dtype == torch.float32 and X.is_cuda and X.is_contiguous()
        n, d = X.shape
        K = self.n_clusters
        if n < K:
            raise ValueError(f"n_samples={n} should be >= n_clusters={K}.")   # sklearn _check_params_vs_input
        dev = X.device
        rs = np.random.RandomState(self.random_state)
        batch = min(self.batch_size, n)
        init_size = 3 * batch
        if init_size < K:
            init_size = 3 * K
        init_size = min(init_size, n)

        rs.randint(0, n, init_size)                          # validation set (only ranks inits; n_init == 1)
        centers = self._init_centroids(X, rs, init_size)
        centers_new = torch.empty_like(centers)

        ewa, ewa_min, no_improvement, n_since_reassign = None, None, 0, 0
        n_steps = (self.max_iter * n) // batch
        # `random_state.choice(n, batch, p=sample_weight / sum)` (cluster/_kmeans.py:2169-2174) is, for replace=True
        # with p, `cdf.searchsorted(random_sample(batch), side='right')` with cdf = float64 cumsum of p normalised by
        # its last entry (numpy mtrand.pyx RandomState.choice); the cdf is the same every step, so it is built once
        cdf_dev = _uniform_cdf(n, dev)
        labels_b = torch.empty(batch, dtype=torch.int32, device=dev)
        sq_b = torch.empty(batch, dtype=torch.float32, device=dev)
        idx = torch.empty(batch, dtype=torch.int32, device=dev)
        # Host <-> device traffic of a step goes through pinned host buffers that the kernels address directly (unified
        # addressing: a cudaHostAlloc'ed buffer has the same pointer on the device): the uniforms are read, the per-cluster
        # counts and the batch inertia are kept, and the reassignment list is read there - no copy calls, one stream
        # synchronisation per step (inherent: the next step's RandomState draws depend on this step's counts).
        u_pin = _pinned('u', batch, torch.float64)
        counts = _pinned('counts', K, torch.float32)
        counts.zero_()
        counts_h = np.zeros(K, dtype=np.float32)
        counts_view = counts.numpy()
        inertia_pin = _pinned('inertia', 1, torch.float64)
        inertia_view = inertia_pin.numpy()
        re_pin = _pinned('reassign', 2 * K, torch.int32)
        re_np = re_pin.numpy()
        u_np = u_pin.numpy()
        stream = torch.cuda.current_stream()
        p_X, ldx, p_cdf, p_u, p_idx, p_lab, p_sq = ptr(X), X.stride(0), ptr(cdf_dev), ptr(u_pin), ptr(idx), ptr(labels_b), ptr(sq_b)
        p_counts, p_inertia, p_re = ptr(counts), ptr(inertia_pin), ptr(re_pin)
        p_c, p_cn = ptr(centers), ptr(centers_new)
        step = 0
        for step in range(n_steps):
            u_np[:] = rs.random_sample(batch)
            # _random_reassign(), evaluated before the step on the counts of the previous one
            n_since_reassign += batch
            random_reassign = bool((counts_h == 0).any() or n_since_reassign >= 10 * K)
            if random_reassign:
                n_since_reassign = 0
            # _mini_batch_step: one call enqueues batch selection (searchsorted of the uniforms against the float64 cdf),
            # assignment, inertia and centre update
            L.kmeans_minibatch_step(p_X, ldx, n, d, K, p_cdf, p_u, batch, p_idx, p_lab, p_sq, p_c, p_cn, p_counts,
                                    p_inertia, st)
            stream.synchronize()
            counts_h = counts_view.copy()
            batch_inertia = float(np.float32(inertia_view[0]))
            if random_reassign and self.reassignment_ratio > 0:
                to_reassign = counts_h < self.reassignment_ratio * counts_h.max()
                if to_reassign.sum() > 0.5 * batch:
                    keep = np.argsort(counts_h)[int(0.5 * batch):]
                    to_reassign[keep] = False
                n_reassigns = int(to_reassign.sum())
                if n_reassigns:
                    picks = rs.choice(batch, replace=False, size=n_reassigns)
                    # the "dirty hack" count reset of _mini_batch_step (cluster/_kmeans.py:1679-1682)
                    min_count = np.min(counts_h[~to_reassign])
                    counts_h[to_reassign] = min_count
                    re_np[:n_reassigns] = np.flatnonzero(to_reassign)
                    re_np[K:K + n_reassigns] = picks
                    L.kmeans_reassign(p_X, ldx, d, p_idx, p_re, p_re + 4 * K, n_reassigns, float(min_count), p_cn,
                                      p_counts, st)
            p_c, p_cn = p_cn, p_c
            centers, centers_new = centers_new, centers
            # _mini_batch_convergence (tol == 0)
            batch_inertia /= batch
            if step == 0:
                continue
            if ewa is None:
                ewa = batch_inertia
            else:
                alpha = min(batch * 2.0 / (n + 1), 1)
                ewa = ewa * (1 - alpha) + batch_inertia * alpha
            if ewa_min is None or ewa < ewa_min:
                no_improvement, ewa_min = 0, ewa
            else:
                no_improvement += 1
            if self.max_no_improvement is not None and no_improvement >= self.max_no_improvement:
                break
        self.cluster_centers_ = centers
        self.n_steps_ = step + 1
        self.labels_ = torch.empty(n, dtype=torch.int32, device=dev)
        L.kmeans_assign(ptr(X), X.stride(0), None, n, ptr(centers), K, d, ptr(self.labels_), None, st)
        return self.labels_


def compute_gene_closeness(X_pca: torch.Tensor, labels: torch.Tensor, centers: torch.Tensor) -> torch.Tensor:
    """scGeneClust/tl/cluster.py:84-105 on the device: euclidean distance of every gene to its cluster centre
    (`paired_distances`), then `1 - minmax_scale` within each cluster (float32, sklearn's x*scale + min form)."""
    L, st = lib(), stream_handle()
    n, d = X_pca.shape
    K = centers.shape[0]
    dist = torch.empty(n, dtype=torch.float32, device=X_pca.device)
    out = torch.empty_like(dist)
    scratch = torch.empty(2 * K, dtype=torch.int32, device=X_pca.device)
    L.paired_dist(ptr(X_pca), X_pca.stride(0), n, d, ptr(centers), ptr(labels), ptr(dist), st)
    L.closeness(ptr(dist), ptr(labels), n, K, ptr(scratch), ptr(out), st)
    return out


def cluster_genes(adata, img: Optional[np.ndarray], version: Literal['fast', 'ps'],
                  modality: Literal['sc', 'st'] = 'sc', shape: Literal['hexagon', 'square'] = 'hexagon',
                  n_gene_clusters: int = None, n_obs_clusters: int = None, n_components: int = 10,
                  relevant_gene_pct: int = 20, max_workers: int = os.cpu_count() - 1, random_state: int = 0):
    """Same signature and AnnData side effects as the reference (scGeneClust/tl/cluster.py:22-81).  `max_workers`
    is accepted for compatibility; the work runs on the GPU."""
    logger.info("Clustering genes...")
    st = get_state(adata)
    if version == 'fast':
        X = st.varm_pca
        if X is None:
            X = torch.from_numpy(np.ascontiguousarray(adata.varm['X_pca'], dtype=np.float32)).cuda()
        km = DeviceMiniBatchKMeans(n_clusters=n_gene_clusters, batch_size=max(1024, adata.n_vars // 10),
                                   random_state=random_state)
        labels = km.fit_predict(X)
        closeness = compute_gene_closeness(X, labels, km.cluster_centers_)
        # one synchronisation for both read-backs (pinned staging; the var columns get their own arrays)
        lab_pin, clo_pin = _pinned('labels', labels.numel(), torch.int32), _pinned('closeness', closeness.numel(), torch.float32)
        lab_pin.copy_(labels, non_blocking=True)
        clo_pin.copy_(closeness, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        labels_h, closeness_h = lab_pin.numpy().copy(), clo_pin.numpy().copy()
        adata.var['cluster'] = labels_h
        adata.var['closeness'] = closeness_h
        # the selection stage reduces these on the device as long as the var columns still hold these values
        st.fast_clusters = (labels, closeness, int(n_gene_clusters), labels_h, closeness_h)
        st.info['kmeans'] = dict(n_steps=km.n_steps_, centers=km.cluster_centers_)
    else:
        from .confidence import find_high_confidence_cells
        from .information import compute_gene_complementarity, compute_gene_redundancy, find_relevant_genes
        if modality != 'sc':
            raise NotImplementedError("modality='st' (SpaGCN spatial domains) is outside the B200 hot path")
        find_high_confidence_cells(adata, n_obs_clusters, n_components, max_workers, random_state)
        find_relevant_genes(adata, relevant_gene_pct, max_workers, random_state)
        compute_gene_redundancy(adata, max_workers, random_state)
        compute_gene_complementarity(adata, max_workers, random_state)
        generate_gene_clusters(adata)
    logger.info("Gene clustering done!")


def generate_gene_clusters(adata):
    """scGeneClust/tl/cluster.py:108-131: MST + edge scale -> single-linkage tree -> condensed tree (min cluster size
    3) -> EOM clusters + GLOSH outlier scores.  Sequential O(N) tree walks on a few thousand genes: host side."""
    from ._hdbscan_tail import mst_to_clusters
    g1, g2 = adata.uns['mst_edges'][:, 0], adata.uns['mst_edges'][:, 1]
    rel = adata.var['relevance'].values
    edge_min_relevance = np.minimum(rel[g1], rel[g2])
    forest = get_state(adata).forest
    if forest is not None:                                    # streamed redundancy stage: the forest carries its weights
        assert np.array_equal(forest[0], adata.uns['mst_edges'])
        edge_redundancy = forest[1]
    else:
        edge_redundancy = adata.varp['redundancy'][g1, g2]    # device-backed matrices gather the N - 1 entries there
    edge_scales = np.maximum(edge_min_relevance, adata.uns['mst_edges_complm']) / edge_redundancy
    MST = np.hstack((adata.uns['mst_edges'], edge_scales.reshape(-1, 1)))
    MST = MST[np.argsort(MST.T[2]), :]
    labels, scores = mst_to_clusters(MST, adata.n_vars, min_cluster_size=3)
    adata.var['outlier_score'], adata.var['cluster'] = scores, labels
