"""Argument checks with the reference's error types and messages (scGeneClust/_validation.py:15-138).  The
integer-count test on `adata.X` (reference: host densify + `X % 1 == 0`, _validation.py:68-70) is fused into the
device ingest (`gc_csr_to_counts` / `gc_dense_to_counts` flags) and raises the same ValueError from there."""
from __future__ import annotations

import os

import numpy as np
from loguru import logger

from ._anndata_lite import is_anndata


def check_args(adata, img, version, n_gene_clusters, n_cell_clusters, n_components, relevant_gene_pct,
               post_hoc_filtering, modality, shape, return_info, subset, max_workers, verbosity, random_state):
    if verbosity not in (0, 1, 2):
        raise ValueError(f"Expected `verbosity` in {{0, 1, 2}}, got {verbosity}.")
    if version not in ('fast', 'ps'):
        raise ValueError(f"Expected `version` in {{'fast', 'ps'}}, got {version}.")
    if modality not in ('sc', 'st'):
        raise ValueError(f"Expected `modality` in {{'sc', 'st'}}, got {modality}.")
    if not isinstance(random_state, int):
        raise TypeError(f"Expected `random_stat` to be an integer, got {type(random_state)}.")
    if not isinstance(max_workers, int):
        raise TypeError(f"Expected `max_workers` to be a positive integer, got {type(max_workers)}.")
    if max_workers < 1:
        raise ValueError(f"Expected `max_workers` to be a positive integer, got {max_workers}.")
    if max_workers > os.cpu_count():
        raise ValueError(f"Worker limit exceeded. Maximum {os.cpu_count()}, got {max_workers}.")
    if not isinstance(return_info, bool):
        raise TypeError(f"Expected `return_info` to be bool, got {type(return_info)}.")
    if not isinstance(subset, bool):
        raise TypeError(f"Expected `subset` to be bool, got {type(subset)}.")
    check_raw_counts(adata)
    check_gene_clustering(img, version, n_gene_clusters, n_cell_clusters, n_components, relevant_gene_pct,
                          post_hoc_filtering, modality, shape)


def check_raw_counts(adata):
    """Type / `.raw` checks of the reference; the integer-valued test itself runs on the device during ingest."""
    if is_anndata(adata):
        if adata.raw is not None:
            raise ValueError("Expected raw counts in `adata.raw` as input.")
    else:
        raise TypeError(f"Expected an `AnnData` object, got {type(adata)}.")


def check_gene_clustering(img, version, n_gene_clusters, n_obs_clusters, n_components, relevant_gene_pct,
                          post_hoc_filtering, modality, shape):
    if not isinstance(post_hoc_filtering, bool):
        raise TypeError(f"Expected `post_hoc_filtering` to be bool, got {type(post_hoc_filtering)}.")
    if version == 'fast':
        if not isinstance(n_gene_clusters, int):
            raise TypeError(f"Expected `n_gene_clusters` to be an integer, got {type(n_gene_clusters)}.")
        if n_gene_clusters <= 1:
            raise ValueError(f"Expected `n_gene_clusters` > 1, got {n_gene_clusters}.")
        if modality == 'st':
            raise ValueError("GeneClust-fast does not support spatial transcriptomics. Please set `version == 'ps'`.")
        return
    if n_gene_clusters is not None:
        raise TypeError(f"Expected `n_gene_clusters` to be None, got {type(n_gene_clusters)}.")
    if modality == 'st':
        if img is None:
            raise RuntimeWarning("Image is not available. This could lower the accuracy of spaGCN.")
        if not isinstance(img, np.ndarray):
            raise TypeError(f"Expected `image` to be an ndarray, got {type(img)}.")
    if shape not in ('hexagon', 'square'):
        raise ValueError(f"Expected `shape` in {{'hexagon', 'square'}}, got {shape}.")
    if not isinstance(n_obs_clusters, int):
        raise TypeError(f"Expected `n_obs_clusters` to be an integer, got {type(n_obs_clusters)}.")
    if n_obs_clusters <= 1:
        raise ValueError(f"Expected `n_obs_clusters` > 1, got {n_obs_clusters}.")
    if not isinstance(n_components, int):
        raise TypeError(f"Expected `n_components` to be an integer, got {type(n_components)}.")
    if n_components <= 1:
        raise ValueError(f"Expected `n_components` > 1, got {n_components}.")
    if not isinstance(relevant_gene_pct, int):
        raise TypeError(f"Expected `relevant_gene_pct` to be an integer, got {type(relevant_gene_pct)}.")
    if relevant_gene_pct <= 0 or relevant_gene_pct >= 100:
        raise ValueError(f"Expected `relevant_gene_pct` between (0, 100), got {relevant_gene_pct}.")


def check_all_genes_selected(adata, selected_genes: np.ndarray):
    # hash lookup instead of the reference's np.isin on object arrays (sort-based, ~70 ms for 20 000 names)
    wanted = set(np.asarray(selected_genes).tolist())
    n_selected_genes = sum(map(wanted.__contains__, adata.var_names.tolist()))
    if n_selected_genes != selected_genes.shape[0]:
        msg = f"Found only {n_selected_genes} selected genes in `adata.var_names`, not {selected_genes.shape[0]}."
        raise RuntimeError(msg)
    logger.opt(colors=True).info(f"Selected <yellow>{n_selected_genes}</yellow> genes.")
