"""Counter-based synthetic negative-binomial count generator.  TEST/BENCH INFRASTRUCTURE.

numpy twin of `gc_synth_counts` (scgeneclust_b200/csrc/synth_kernels.cu): the two produce bit-identical uint16
counts for any (cell, gene) sub-block, so the CPU oracle can regenerate any subsample of a matrix that was
generated on the device shard by shard (config c5: 1M x 30k never exists on the host).

Model: 8 cell types (uniform), per-gene base mean exp(N(-1.5, 1.5^2)) clipped to [0.01, 50], 5 % marker genes per
type with log2 fold change N(0, 1.5^2), 40 co-regulated modules of 50 genes sharing a per-cell log-normal factor,
per-cell size factor exp(N(0, 0.3^2)) (the parameter tables are built by
`scgeneclust_b200.synth.synth_params` and passed in as `p`); counts ~ NB(mean, r = 2) (Gamma-Poisson with dispersion 0.5) by inversion
of the cdf.  Per-cell factors are quantised to 64 levels selected by hashing so the device needs no per-cell tables.
"""
from __future__ import annotations

import ctypes
import os
from types import SimpleNamespace

import numpy as np

U64 = np.uint64
_M1, _M2, _GOLD = U64(0xBF58476D1CE4E5B9), U64(0x94D049BB133111EB), U64(0x9E3779B97F4A7C15)


def _mix64(z):
    z = z ^ (z >> U64(30)); z = z * _M1
    z = z ^ (z >> U64(27)); z = z * _M2
    return z ^ (z >> U64(31))


def hash3(seed: int, stream: int, a, b):
    """splitmix-style hash of (seed, stream, a, b) -> uint64; same integer arithmetic as the device kernel."""
    with np.errstate(over="ignore"):
        z = U64(seed) + _GOLD * U64(stream + 1)
        z = _mix64(z ^ (np.asarray(a, dtype=U64) * _M1))
        z = _mix64(z ^ (np.asarray(b, dtype=U64) * _M2))
    return z


def cell_types(p, c0: int, nc: int) -> np.ndarray:
    """Generator's true cell type of cells [c0, c0+nc)."""
    cells = np.arange(c0, c0 + nc, dtype=U64)
    return (hash3(p.seed, 1, cells, 0) % U64(p.n_types)).astype(np.int32)


def synth_counts(p, c0: int, nc: int, g0: int, ng: int) -> np.ndarray:
    """uint16 counts of cells [c0, c0+nc) x genes [g0, g0+ng); bit-identical to gc_synth_counts."""
    cells = np.arange(c0, c0 + nc, dtype=U64)[:, None]
    genes = np.arange(g0, g0 + ng, dtype=U64)[None, :]
    gi = np.arange(g0, g0 + ng)
    t = (hash3(p.seed, 1, cells, 0) % U64(p.n_types)).astype(np.int64)[:, 0]
    sc = (hash3(p.seed, 2, cells, 0) % U64(len(p.size_levels))).astype(np.int64)[:, 0]
    mean = p.gene_mean[gi][None, :] * p.type_fc[gi][:, t].T
    mean = mean * p.size_levels[sc][:, None]
    mod = p.gene_module[gi]
    has = mod >= 0
    if has.any():
        lev = hash3(p.seed, 3, cells, mod[has].astype(U64)[None, :]) % U64(len(p.module_levels))
        mean[:, has] = mean[:, has] * p.module_levels[lev.astype(np.int64)]
    u = (hash3(p.seed, 4, cells, genes) >> U64(11)).astype(np.float64) * 1.1102230246251565e-16
    denom = 2.0 + mean
    pr, q = 2.0 / denom, mean / denom
    pmf = pr * pr
    cdf = pmf.copy()
    k = np.zeros(mean.shape, np.int64)
    active = (u > cdf) & (pmf > 0.0)
    while active.any():
        idx = np.nonzero(active)
        kk = k[idx] + 1
        pm = (pmf[idx] * q[idx]) * ((kk + 1).astype(np.float64) / kk.astype(np.float64))
        cd = cdf[idx] + pm
        k[idx], pmf[idx], cdf[idx] = kk, pm, cd
        still = (u[idx] > cd) & (kk < 65535) & (pm > 0.0)
        active[idx] = still
    return k.astype(np.uint16)


def synth_params(G: int, seed: int = 0, n_types: int = 8, n_modules: int = 40, module_size: int = 50):
    """The generator's parameter tables - the oracle's own copy of `scgeneclust_b200.synth.synth_params` (same draws
    from RandomState(seed) in the same order; equality is checked in tests/test_host_logic.py), so the CPU reference
    arm of bench.py builds its input without importing anything of the product package."""
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
    return SimpleNamespace(seed=seed, G=G, n_types=n_types, gene_mean=gene_mean,
                           type_fc=np.ascontiguousarray(type_fc), gene_module=gene_module, size_levels=size_levels,
                           module_levels=module_levels, n_modules=n_modules)


_CLIB = None


def _clib():
    """oracle/_build/liboracle_synth.so (oracle/synth_c.c), built on first use when gcc is present."""
    global _CLIB
    if _CLIB is None:
        from .build_c import LIB, build
        path = LIB if os.path.exists(LIB) else build()
        dll = ctypes.CDLL(path)
        p = ctypes.c_void_p
        dll.oracle_synth_counts.restype = None
        dll.oracle_synth_counts.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                            ctypes.c_int64, p, p, ctypes.c_int, p, p, ctypes.c_int, p, ctypes.c_int,
                                            p, ctypes.c_int64]
        _CLIB = dll
    return _CLIB


def synth_counts_c(p, c0: int, nc: int, g0: int = 0, ng: int | None = None, out: np.ndarray | None = None) -> np.ndarray:
    """Same block as `synth_counts`, from the C twin (oracle/synth_c.c, OpenMP over cells): what the full-size CPU
    arms use (the numpy version takes ~0.1 s per cell).  `out`: optional preallocated uint16 [nc, >= ng] C-order."""
    ng = p.G - g0 if ng is None else ng
    if out is None:
        out = np.empty((nc, ng), dtype=np.uint16)
    assert out.dtype == np.uint16 and out.flags.c_contiguous and out.shape[0] == nc and out.shape[1] >= ng
    gm = np.ascontiguousarray(p.gene_mean, dtype=np.float64)
    fc = np.ascontiguousarray(p.type_fc, dtype=np.float64)
    mod = np.ascontiguousarray(p.gene_module, dtype=np.int32)
    sl = np.ascontiguousarray(p.size_levels, dtype=np.float64)
    ml = np.ascontiguousarray(p.module_levels, dtype=np.float64)
    _clib().oracle_synth_counts(p.seed, c0, nc, g0, ng, gm.ctypes.data, fc.ctypes.data, p.n_types, mod.ctypes.data,
                                sl.ctypes.data, len(sl), ml.ctypes.data, len(ml), out.ctypes.data, out.shape[1])
    return out[:, :ng]
