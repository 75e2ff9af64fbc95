"""Preprocessing stages with the reference's names and semantics (scGeneClust/pp/_preprocessing.py:15-55), running
on the device.  Both functions mutate `adata` like the reference; device-resident state lives in
`adata.uns['geneclust_b200']` (a `HotPathState`)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Literal, Optional

import numpy as np
import torch
from loguru import logger

from .._device import (DeviceCounts, HostCopy, ResidualOperator, ingest_counts, plan_pca, prefetch_omega,
                       randomized_pca)

STATE_KEY = "geneclust_b200"


@dataclass
class HotPathState:
    """Device-side companions of the AnnData slots the reference fills."""
    op: Optional[ResidualOperator] = None        # implicit adata.X (Pearson residuals)
    Z: Optional[torch.Tensor] = None             # materialised residuals (ps path), float32 C x G
    varm_pca: Optional[torch.Tensor] = None      # device copy of varm['X_pca']
    obsm_pca: Optional[torch.Tensor] = None
    redundancy: Optional[torch.Tensor] = None    # device copy of varp['redundancy']
    dense_redundancy: bool = True                # False: the entry point does not return varp['redundancy'] -> stream it
    forest: Optional[tuple] = None               # (edges (n, 2) int64, redundancy of each edge) from the streamed stage
    fast_clusters: Optional[tuple] = None        # (labels int32, closeness float32, K) on the device, GeneClust-fast
    comm: object = None                          # optional GeneShardComm
    omega: object = None                         # OmegaPrefetch started by `start_omega`
    host_copies: list = field(default_factory=list)   # HostCopy objects in flight (`finish_host_copies`)
    info: dict = field(default_factory=dict)


def get_state(adata) -> HotPathState:
    st = adata.uns.get(STATE_KEY)
    if st is None:
        st = HotPathState()
        adata.uns[STATE_KEY] = st
    return st


def start_omega(adata, random_state: int = 0):
    """Optional head start for `reduce_dim`: begin generating the randomized solver's test matrix (a sequential MT19937
    stream, one CTA) on a side stream before the counts are uploaded / summed.  Purely an ordering optimisation: without
    it `reduce_dim` generates the same matrix itself."""
    from .._lib import require_cuda
    require_cuda()
    n_obs, n_vars = adata.shape
    get_state(adata).omega = prefetch_omega(random_state, plan_pca(int(n_obs), int(n_vars), 'genes'), 'cuda')


def normalize(adata, modality: Literal['sc', 'st'] = 'sc', materialize: bool = False):
    """`adata.X` (raw counts) -> analytic Pearson residuals (reference: pp/_preprocessing.py:15-29).

    The counts go to the device once (uint16); the residual matrix is an implicit operator there.  `adata.X` is
    overwritten with the dense float32 residuals only when `materialize=True` (GeneClust-ps and
    `return_info=True` need them on the host); GeneClust-fast never materialises them.
    """
    logger.info("Preprocessing data...")
    if modality != 'sc':
        raise NotImplementedError("modality='st' (normalize_total + log1p, SpaGCN) is outside the B200 hot path")
    st = get_state(adata)
    counts = ingest_counts(adata.X)
    st.comm = counts.comm
    st.op = ResidualOperator.from_counts(counts)
    adata.uns['pearson_residuals_normalization'] = {'theta': st.op.theta, 'clip': None, 'computed_on': 'adata.X'}
    if materialize:
        st.Z = st.op.materialize()                   # gene-sharded input: the residual columns of this rank's genes
        from .._anndata_lite import AnnDataLite, DeviceMatrix, ShardedMatrix
        if st.comm is not None:
            if not isinstance(adata, AnnDataLite):
                raise NotImplementedError("gene-sharded dense residuals need the AnnDataLite container")
            adata.X = ShardedMatrix(st.Z, np.arange(st.comm.gene_lo, st.comm.gene_hi), st.comm.total_genes, st.comm)
        else:
            # AnnDataLite carries the device matrix itself (downloaded on first host use); a real AnnData needs an ndarray
            adata.X = DeviceMatrix(st.Z) if isinstance(adata, AnnDataLite) else st.Z.cpu().numpy()


def finish_host_copies(adata):
    """Wait for the host copies a stage left in flight (`reduce_dim(..., defer_host_copy=True)`)."""
    st = get_state(adata)
    for hc in st.host_copies:
        hc.wait()
    st.host_copies = []


def reduce_dim(adata, version: Literal['fast', 'ps'], random_state: int = 0, defer_host_copy: bool = False):
    """Gene-level PCA into `adata.varm['X_pca']` and, for GeneClust-ps, cell-level PCA into `adata.obsm['X_pca']`
    (reference: pp/_preprocessing.py:35-55; sc.pp.pca -> sklearn PCA(50, svd_solver='auto') -> randomized SVD).

    `defer_host_copy=True` (the entry point): the ndarray put into `varm['X_pca']` is filled on a side stream while the
    clustering already runs on the device copy; its contents are valid after `finish_host_copies(adata)`, which the
    entry point calls before it returns.  The default waits for the copy, like any stage-by-stage caller expects."""
    st = get_state(adata)
    if st.op is None:
        raise RuntimeError("reduce_dim called before normalize")
    info = {}
    st.varm_pca = randomized_pca(st.op, 'genes', random_state, info=info, omega=st.omega)
    st.omega = None
    st.info['gene_pca'] = info
    if st.comm is not None:
        st.varm_pca = st.comm.all_gather_rows(st.varm_pca)
    hc = HostCopy(st.varm_pca)
    adata.varm['X_pca'] = hc.array
    if defer_host_copy:
        st.host_copies.append(hc)
    else:
        hc.wait()
    if version == 'ps':
        st.obsm_pca = randomized_pca(st.op, 'cells', random_state)
        adata.obsm['X_pca'] = st.obsm_pca.cpu().numpy()
    logger.info("Data preprocessing done.")
