"""The single entry point, same signature / returns / error behaviour as the reference's
`scGeneClust.scGeneClust` (scGeneClust/_model.py:20-144)."""
from __future__ import annotations

import os
from typing import Literal, Optional, Tuple, Union

import numpy as np
from loguru import logger

from . import pp, tl
from ._utils import set_logger
from ._validation import check_all_genes_selected, check_args


def _working_copy(raw_adata):
    """`raw_adata.copy()` without duplicating / densifying the count matrix on the host (contrast
    scGeneClust/_model.py:123-124): the copy shares `X` until `pp.normalize` moves it to the device."""
    from ._anndata_lite import AnnDataLite
    if isinstance(raw_adata, AnnDataLite):
        new = AnnDataLite.__new__(AnnDataLite)
        new._X = raw_adata._X
        new.obs, new.var = raw_adata.obs.copy(), raw_adata.var.copy()
        for slot in ("obsm", "varm", "obsp", "varp"):
            setattr(new, slot, {k: np.array(v, copy=True) for k, v in getattr(raw_adata, slot).items()})
        new.uns = dict(raw_adata.uns)
        new.raw = None
        return new
    return raw_adata.copy()


def scGeneClust(
        raw_adata,
        image: np.ndarray = None,
        n_var_clusters: int = None,
        n_obs_clusters: int = None,
        n_components: int = 10,
        relevant_gene_pct: int = 20,
        post_hoc_filtering: bool = True,
        version: Literal['fast', 'ps'] = 'fast',
        modality: Literal['sc', 'st'] = 'sc',
        shape: Literal['hexagon', 'square'] = 'hexagon',
        return_info: bool = False,
        subset: bool = False,
        max_workers: int = os.cpu_count() - 1,
        verbosity: Literal[0, 1, 2] = 1,
        random_state: int = 0
) -> Optional[Union[Tuple[object, np.ndarray], np.ndarray]]:
    """GeneClust-fast / GeneClust-ps gene selection on a B200.

    Parameters, return modes (`subset` -> None and in-place subset; `return_info` -> (adata, genes); else genes)
    and exceptions are the reference's.  `raw_adata.X` holds integer-valued raw counts (dense ndarray, scipy
    sparse, or an already device-resident `DeviceCounts`).  `max_workers` is validated like the reference but
    unused: the stages run on the GPU.  With `return_info=True` the returned AnnData carries the reference's slots
    (`varm['X_pca']`, `var['cluster']`, `var['closeness']` for fast; `var['relevance']`, `varp['redundancy']`,
    `uns['mst_edges']`, `uns['mst_edges_complm']`, `var['outlier_score']`, `var['representative']` for ps) and
    `X` = the Pearson residuals.
    """
    check_args(raw_adata, image, version, n_var_clusters, n_obs_clusters, n_components, relevant_gene_pct,
               post_hoc_filtering, modality, shape, return_info, subset, max_workers, verbosity, random_state)
    set_logger(verbosity)
    logger.opt(colors=True).info(
        f"Performing <magenta>GeneClust-{version}</magenta> "
        f"on <magenta>{'scRNA-seq' if modality == 'sc' else 'SRT'}</magenta> data, on the GPU."
    )
    copied_adata = _working_copy(raw_adata)

    # head start for the gene PCA: its Omega test matrix is a sequential MT19937 stream (one CTA), so it is started on a
    # side stream before the counts are uploaded and summed
    pp.start_omega(copied_adata, random_state)
    # varp['redundancy'] (N x N) is part of the out-contract only with return_info: otherwise GeneClust-ps streams the
    # MI tiles into the spanning forest and never builds the matrix
    pp.get_state(copied_adata).dense_redundancy = bool(return_info)
    # preprocessing: residuals stay an implicit device operator for `fast`; `ps` and `return_info` need them dense
    pp.normalize(copied_adata, modality, materialize=(version == 'ps' or return_info))
    pp.reduce_dim(copied_adata, version, random_state, defer_host_copy=True)
    # gene clustering
    tl.cluster_genes(copied_adata, image, version, modality, shape, n_var_clusters, n_obs_clusters, n_components,
                     relevant_gene_pct, max_workers, random_state)
    # select features from gene clusters
    selected_genes = tl.select_from_clusters(copied_adata, version, post_hoc_filtering, random_state)
    pp.finish_host_copies(copied_adata)
    check_all_genes_selected(raw_adata, selected_genes)

    if subset:
        raw_adata._inplace_subset_var(selected_genes)
        logger.opt(colors=True).info(f"<magenta>GeneClust-{version}</magenta> finished.")
        return None

    logger.opt(colors=True).info(f"GeneClust-{version} finished.")
    if return_info:
        from ._anndata_lite import DeviceMatrix
        from .pp._preprocessing import STATE_KEY
        # the device-side state (counts, residuals, PCA buffers, N x N redundancy) is not part of the reference's
        # out-contract, cannot be pickled / written to h5ad and would pin gigabytes of HBM for the life of the result
        copied_adata.uns.pop(STATE_KEY, None)
        from ._anndata_lite import ShardedMatrix
        if isinstance(copied_adata.X, ShardedMatrix):         # gene-sharded run: gather the residual columns (collective)
            copied_adata._X = copied_adata.X.gather_columns().cpu().numpy()
        if isinstance(copied_adata.X, DeviceMatrix):          # hand the residuals back as the ndarray the reference returns
            copied_adata.X = np.asarray(copied_adata.X)
        for slot in (copied_adata.varp, copied_adata.varm, copied_adata.obsm):
            for key, value in list(slot.items()):
                if isinstance(value, DeviceMatrix):
                    slot[key] = np.asarray(value)
        return copied_adata, selected_genes
    return selected_genes
