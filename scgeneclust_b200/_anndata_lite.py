"""Minimal AnnData stand-in.

`anndata` is not installable in the build container (no network, no wheel), and the hot path only
touches a small slice of the AnnData surface (`/root/reference/scGeneClust/_model.py:123-143`,
`tl/information.py:53-54`, `tl/confidence.py:75-76`).  `AnnDataLite` implements exactly that slice with the
same attribute names and mutation semantics so the stage functions in this package are written against the
real contract.  When the real `anndata` package is importable it is accepted everywhere `AnnDataLite` is
(see `is_anndata`).
"""
from __future__ import annotations

import copy as _copy
from typing import Optional

import numpy as np
import pandas as pd
from scipy.sparse import issparse

try:  # pragma: no cover - not available in the build container
    import anndata as _ad

    _REAL_ANNDATA = (_ad.AnnData,)
except Exception:  # noqa: BLE001
    _REAL_ANNDATA = ()


class DeviceMatrix:
    """A dense matrix that lives on the device and stands in for `adata.X` (the Pearson residuals) or
    `adata.varp['redundancy']` until a host
    array is really needed: `np.asarray(x)` downloads it, `x[:, idx]` / `x[idx, :]` subset on the device.  GeneClust-ps
    subsets the residual matrix twice (high-confidence cells, relevant genes; scGeneClust/tl/confidence.py:75-76,
    tl/information.py:53-54) before anything reads it on the host - at config c4 the eager copy was 0.8 GB of
    pageable D2H plus two host fancy-index copies."""
    ndim = 2

    def __init__(self, tensor):
        self.tensor = tensor

    @property
    def shape(self):
        return tuple(self.tensor.shape)

    @property
    def dtype(self):
        return np.dtype(str(self.tensor.dtype).replace("torch.", ""))

    def __array__(self, dtype=None, copy=None):
        a = self.tensor.cpu().numpy()
        return a if dtype is None else a.astype(dtype, copy=False)

    def copy(self):
        return DeviceMatrix(self.tensor.clone())

    def __getitem__(self, key):
        """numpy indexing semantics.  `x[rows, :]` / `x[:, cols]` (one axis subset, the other a full slice) stay on the
        device and return a DeviceMatrix; two index arrays are PAIRED like numpy's (`x[i, j]` -> the elements
        x[i[k], j[k]], gathered on the device and returned as an ndarray); everything else goes through the host copy."""
        import torch
        if not (isinstance(key, tuple) and len(key) == 2):
            return np.asarray(self)[key]
        full = [isinstance(k, slice) and k == slice(None) for k in key]
        idx = []
        for k, is_full in zip(key, full):
            if is_full:
                idx.append(None)
                continue
            k = np.asarray(k)
            if k.dtype == bool and k.ndim == 1:
                k = np.flatnonzero(k)
            if k.ndim != 1 or k.dtype.kind not in "iu":
                return np.asarray(self)[key]
            idx.append(torch.from_numpy(k.astype(np.int64)).to(self.tensor.device))
        if all(full):
            return self
        if idx[0] is not None and idx[1] is not None:
            if idx[0].numel() != idx[1].numel():
                return np.asarray(self)[key]                     # let numpy broadcast or raise as it would
            return self.tensor[idx[0], idx[1]].cpu().numpy()
        axis = 0 if idx[0] is not None else 1
        return DeviceMatrix(self.tensor.index_select(axis, idx[axis]).contiguous())


class ShardedMatrix:
    """Gene-sharded stand-in for `adata.X` in a multi-rank GeneClust-ps run: this rank holds the residual columns of ITS
    genes (`tensor`: cells x local genes, device), the AnnData describes the global matrix.  `positions` are the indices
    of the local columns in the CURRENT global column order (ranks own ascending, disjoint index sets, so concatenating
    the shards in rank order gives the global order).  Row subsets apply to every shard alike; a column subset given in
    global indices keeps the local columns it names."""
    ndim = 2

    def __init__(self, tensor, positions: np.ndarray, n_cols_total: int, comm):
        self.tensor, self.positions, self.n_cols_total, self.comm = tensor, np.asarray(positions, dtype=np.int64), int(n_cols_total), comm

    @property
    def shape(self):
        return (int(self.tensor.shape[0]), self.n_cols_total)

    @property
    def dtype(self):
        return np.dtype(str(self.tensor.dtype).replace("torch.", ""))

    def copy(self):
        return ShardedMatrix(self.tensor.clone(), self.positions.copy(), self.n_cols_total, self.comm)

    def __array__(self, dtype=None, copy=None):
        raise TypeError("a gene-sharded matrix has no single host copy; use gather_columns() on the ranks")

    def __getitem__(self, key):
        import torch
        if not (isinstance(key, tuple) and len(key) == 2):
            raise IndexError("ShardedMatrix supports x[rows, :] and x[:, cols] only")
        rows, cols = key
        t, positions, total = self.tensor, self.positions, self.n_cols_total
        if not (isinstance(rows, slice) and rows == slice(None)):
            r = np.asarray(rows)
            r = np.flatnonzero(r) if r.dtype == bool else r.astype(np.int64)
            t = t.index_select(0, torch.from_numpy(r).to(t.device))
        if not (isinstance(cols, slice) and cols == slice(None)):
            c = np.asarray(cols)
            c = np.flatnonzero(c) if c.dtype == bool else c.astype(np.int64)
            if np.any(np.diff(c) <= 0):
                raise IndexError("ShardedMatrix column subsets must be strictly increasing global indices")
            new_pos = np.searchsorted(c, positions)                  # where each local column lands in the new order
            new_pos_c = np.minimum(new_pos, len(c) - 1) if len(c) else new_pos
            keep = (new_pos < len(c)) & (c[new_pos_c] == positions) if len(c) else np.zeros(len(positions), bool)
            t = t.index_select(1, torch.from_numpy(np.flatnonzero(keep)).to(t.device))
            positions, total = new_pos[keep], len(c)
        return ShardedMatrix(t.contiguous(), positions, total, self.comm)

    def gather_columns(self):
        """The full (cells x all current columns) matrix on every rank (all-gather over the gene shards)."""
        import torch
        import torch.distributed as dist
        comm = self.comm
        n_rows = int(self.tensor.shape[0])
        counts = [torch.zeros(1, dtype=torch.int64, device=self.tensor.device) for _ in range(comm.world)]
        dist.all_gather(counts, torch.tensor([self.tensor.shape[1]], dtype=torch.int64, device=self.tensor.device), group=comm.group)
        counts = [int(c.item()) for c in counts]
        width = max(max(counts), 1)
        pad = torch.zeros((width, n_rows), dtype=self.tensor.dtype, device=self.tensor.device)   # columns as rows: contiguous
        pad[:self.tensor.shape[1]] = self.tensor.t()
        parts = [torch.empty_like(pad) for _ in range(comm.world)]
        dist.all_gather(parts, pad, group=comm.group)
        full = torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0).t().contiguous()
        assert full.shape[1] == self.n_cols_total
        return full


class AnnDataLite:
    """`n_obs` x `n_vars` annotated matrix: `.X`, `.obs`, `.var`, `.obsm`, `.varm`, `.obsp`, `.varp`, `.uns`."""

    def __init__(self, X, obs: Optional[pd.DataFrame] = None, var: Optional[pd.DataFrame] = None,
                 obs_names=None, var_names=None):
        if not (issparse(X) or isinstance(X, (np.ndarray, DeviceMatrix, ShardedMatrix)) or getattr(X, "is_device_counts", False)):
            X = np.asarray(X)
        if X.ndim != 2:
            raise ValueError(f"X must be 2-D, got shape {X.shape}")
        self._X = X
        n_obs, n_vars = X.shape
        if obs is None:
            # names are kept as an object-dtype Index (pandas 3 would otherwise spend 1-2 ms per 20 000 names inferring
            # its arrow string dtype); default names ('cell_0', ... / 'gene_0', ...) are materialised on first use of `obs_names` / `var_names`
            # (building 10^5..10^6 strings up front would cost more than the whole device pipeline)
            index = pd.Index(np.asarray(obs_names, dtype=object), dtype=object) if obs_names is not None else pd.RangeIndex(n_obs)
            obs = pd.DataFrame(index=index)
        if var is None:
            index = pd.Index(np.asarray(var_names, dtype=object), dtype=object) if var_names is not None else pd.RangeIndex(n_vars)
            var = pd.DataFrame(index=index)
        if len(obs) != n_obs or len(var) != n_vars:
            raise ValueError("obs/var length does not match X")
        self.obs, self.var = obs, var
        self.obsm, self.varm, self.obsp, self.varp, self.uns = {}, {}, {}, {}, {}
        self.raw = None

    # -- shape / names --------------------------------------------------------------------------------------
    @property
    def X(self):
        return self._X

    @X.setter
    def X(self, value):
        if value.shape != self._X.shape:
            raise ValueError(f"X has shape {self._X.shape}, cannot assign shape {value.shape}")
        self._X = value

    @property
    def n_obs(self) -> int:
        return self._X.shape[0]

    @property
    def n_vars(self) -> int:
        return self._X.shape[1]

    @property
    def shape(self):
        return self._X.shape

    @staticmethod
    def _named(frame: pd.DataFrame, prefix: str) -> pd.Index:
        if isinstance(frame.index, pd.RangeIndex):
            frame.index = pd.Index(np.char.add(prefix, np.arange(len(frame)).astype(str)).astype(object), dtype=object)
        return frame.index

    @property
    def obs_names(self) -> pd.Index:
        return self._named(self.obs, "cell_")

    @property
    def var_names(self) -> pd.Index:
        return self._named(self.var, "gene_")

    # -- copy / subset --------------------------------------------------------------------------------------
    def copy(self) -> "AnnDataLite":
        new = AnnDataLite(self._X.copy(), self.obs.copy(), self.var.copy())
        for slot in ("obsm", "varm", "obsp", "varp"):
            setattr(new, slot, {k: np.array(v, copy=True) for k, v in getattr(self, slot).items()})
        new.uns = _copy.deepcopy(self.uns)
        return new

    def _resolve(self, index, names: pd.Index) -> np.ndarray:
        index = np.asarray(index)
        if index.dtype == bool:
            if index.shape[0] != len(names):
                raise IndexError("boolean index has wrong length")
            return np.flatnonzero(index)
        if index.dtype.kind in "iu":
            return index.astype(np.intp)
        pos = names.get_indexer(index)
        if (pos < 0).any():
            raise KeyError(f"names not found: {list(index[pos < 0][:5])}")
        return pos

    def _inplace_subset_var(self, index) -> None:
        pos = self._resolve(index, self.var_names)
        self._X = self._X[:, pos]
        self.var = self.var.iloc[pos].copy()
        self.varm = {k: np.asarray(v)[pos] for k, v in self.varm.items()}
        self.varp = {k: np.asarray(v)[np.ix_(pos, pos)] for k, v in self.varp.items()}

    def _inplace_subset_obs(self, index) -> None:
        pos = self._resolve(index, self.obs_names)
        self._X = self._X[pos, :]
        self.obs = self.obs.iloc[pos].copy()
        self.obsm = {k: np.asarray(v)[pos] for k, v in self.obsm.items()}
        self.obsp = {k: np.asarray(v)[np.ix_(pos, pos)] for k, v in self.obsp.items()}

    def __repr__(self) -> str:
        return (f"AnnDataLite object with n_obs x n_vars = {self.n_obs} x {self.n_vars}\n"
                f"    obs: {list(self.obs.columns)}\n    var: {list(self.var.columns)}\n"
                f"    uns: {list(self.uns)}\n    obsm: {list(self.obsm)}\n    varm: {list(self.varm)}\n"
                f"    varp: {list(self.varp)}")


def is_anndata(obj) -> bool:
    """True for `AnnDataLite` and, when importable, for real `anndata.AnnData`."""
    return isinstance(obj, (AnnDataLite,) + _REAL_ANNDATA)
