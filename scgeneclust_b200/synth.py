"""Synthetic negative-binomial count matrices for the bench and test harness, generated on
the device by `gc_synth_counts`, shard by shard; `oracle/synth.py` is the bit-identical numpy twin used by the
checker.  Not part of the reference's API (its loaders download real datasets, scGeneClust/_utils.py:17-74)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class SynthParams:
    seed: int
    G: int
    n_types: int
    gene_mean: np.ndarray      # float64 [G]
    type_fc: np.ndarray        # float64 [G, n_types]
    gene_module: np.ndarray    # int32 [G], -1 = no module
    size_levels: np.ndarray    # float64 [64]
    module_levels: np.ndarray  # float64 [64]
    n_modules: int


def synth_params(G: int, seed: int = 0, n_types: int = 8, n_modules: int = 40, module_size: int = 50) -> SynthParams:
    """Parameter tables of the generator: 8 cell types, per-gene base mean exp(N(-1.5, 1.5^2)) clipped to
    [0.01, 50], 5 % marker genes per type (log2 fold change N(0, 1.5^2)), 40 modules of 50 co-regulated genes,
    64 quantised levels each for the per-cell size factor exp(N(0, 0.3^2)) and the per-(cell, module) factor
    exp(N(0, 0.5^2))."""
    rs = np.random.RandomState(seed)
    gene_mean = np.clip(np.exp(rs.normal(-1.5, 1.5, G)), 0.01, 50.0)
    type_fc = np.ones((G, n_types))
    n_mark = max(1, int(0.05 * G))
    for t in range(n_types):
        markers = rs.choice(G, n_mark, replace=False)
        type_fc[markers, t] = 2.0 ** rs.normal(0.0, 1.5, n_mark)
    gene_module = np.full(G, -1, np.int32)
    n_in = min(G, n_modules * module_size)
    members = rs.permutation(G)[:n_in]
    gene_module[members] = (np.arange(n_in) // module_size).astype(np.int32)
    size_levels = np.exp(rs.normal(0.0, 0.3, 64))
    module_levels = np.exp(rs.normal(0.0, 0.5, 64))
    return SynthParams(seed, G, n_types, gene_mean, np.ascontiguousarray(type_fc), gene_module, size_levels,
                       module_levels, n_modules)


class DeviceSynth:
    """Parameter tables resident on the device; `counts()` fills any (cell, gene) block of the global matrix."""

    def __init__(self, params: SynthParams, device="cuda"):
        import torch
        self.p = params
        self.dev = torch.device(device)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)  # noqa: E731
        self.gene_mean, self.type_fc, self.gene_module = t(params.gene_mean), t(params.type_fc), t(params.gene_module)
        self.size_levels, self.module_levels = t(params.size_levels), t(params.module_levels)

    def counts(self, c0: int, nc: int, g0: int = 0, ng: int | None = None):
        """`DeviceCounts` holding cells [c0, c0+nc) x genes [g0, g0+ng) of the global synthetic matrix."""
        import torch
        from ._device import DeviceCounts, _round_up
        from ._lib import lib, ptr, stream_handle
        ng = self.p.G - g0 if ng is None else ng
        ld = _round_up(ng, 64)
        out = torch.zeros((nc, ld), dtype=torch.int16, device=self.dev)
        lib().synth_counts(self.p.seed, c0, nc, g0, ng, ptr(self.gene_mean), ptr(self.type_fc), self.p.n_types,
                           ptr(self.gene_module), ptr(self.size_levels), len(self.p.size_levels),
                           ptr(self.module_levels), len(self.p.module_levels), self.p.n_modules, ptr(out), ld,
                           stream_handle())
        return DeviceCounts(out, nc, ng)


def synth_cell_types(params: SynthParams, c0: int, nc: int) -> np.ndarray:
    """Cell type the generator assigned to cells [c0, c0+nc): type = hash(seed, stream 1, cell) mod n_types, the same
    integer hash `gc_synth_counts` evaluates on the device (csrc/synth_kernels.cu).  The bench uses these labels as the
    given pseudo cell types of GeneClust-ps."""
    u64 = np.uint64
    m1, m2, gold = u64(0xBF58476D1CE4E5B9), u64(0x94D049BB133111EB), u64(0x9E3779B97F4A7C15)

    def mix(z):
        z = (z ^ (z >> u64(30))) * m1
        z = (z ^ (z >> u64(27))) * m2
        return z ^ (z >> u64(31))
    cells = np.arange(c0, c0 + nc, dtype=u64)
    with np.errstate(over="ignore"):
        z = u64(params.seed) + gold * u64(2)
        z = mix(z ^ (cells * m1))
        z = mix(z ^ (np.zeros_like(cells) * m2))
    return (z % u64(params.n_types)).astype(np.int32)
