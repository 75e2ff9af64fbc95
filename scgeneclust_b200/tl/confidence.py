"""High-confidence cells (scGeneClust/tl/confidence.py:30-148).  OUT OF SCOPE for the B200 hot path (SURVEY.md
section 8f rank 2): the reference's finder needs 10 GMM fits, a dense cells x cells co-membership matrix and Leiden
(leidenalg/igraph, not installable here).  The hot path therefore takes the pseudo cell types as an INPUT: if the
AnnData already carries `obs['cluster']` (and optionally a boolean `obs['highly_confident']`), those are used."""
from __future__ import annotations

import os

import numpy as np
from loguru import logger

from ..pp._preprocessing import get_state


def find_high_confidence_cells(adata, n_obs_clusters: int, n_components: int = 10,
                               max_workers: int = os.cpu_count() - 1, random_state: int = 0):
    if 'cluster' not in adata.obs:
        raise NotImplementedError(
            "GeneClust-ps on the B200 hot path takes the pseudo cell types as input: provide `adata.obs['cluster']` "
            "(and optionally a boolean `adata.obs['highly_confident']`); the GMM + Leiden finder of the reference "
            "(tl/confidence.py) is outside the accelerated path")
    st = get_state(adata)
    if 'highly_confident' in adata.obs:
        keep = np.asarray(adata.obs['highly_confident'], dtype=bool)
    else:
        keep = np.ones(adata.n_obs, bool)
    adata.obs['cluster'] = np.asarray(adata.obs['cluster']).astype(str)
    if not keep.all():
        from .._anndata_lite import DeviceMatrix, ShardedMatrix
        adata._inplace_subset_obs(keep)
        if isinstance(adata.X, (DeviceMatrix, ShardedMatrix)):
            st.Z = adata.X.tensor                    # the subset already happened on the device
        elif st.Z is not None:
            import torch
            st.Z = st.Z.index_select(0, torch.from_numpy(np.flatnonzero(keep)).to(st.Z.device)).contiguous()
    logger.opt(colors=True).info(
        f"Using <yellow>{adata.n_obs}</yellow> provided high-confidence cells in "
        f"<yellow>{len(np.unique(adata.obs['cluster']))}</yellow> pseudo cell types.")
