"""Device-resident data for the hot path: the uint16 count matrix, the implicit Pearson-residual operator and the
randomized-SVD driver that runs on it.  torch is used for allocation, streams and tiny elementwise glue only; every
pass over the data is a kernel of libgeneclust_b200.so called through the C ABI.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch
from scipy.sparse import issparse

from ._lib import ResidualMatrix, lib, ptr, require_cuda, stream_handle

THETA = 100.0  # scanpy normalize_pearson_residuals default (call site: scGeneClust/pp/_preprocessing.py:29)
N_PCS = 50     # scanpy.settings.N_PCS, the n_comps default of sc.pp.pca (pp/_preprocessing.py:52)
N_OVERSAMPLES = 10
Q_LD = 64      # panel leading dimension used by the kernels


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# -------------------------------------------------------------------------------------------------------------
# ingest
# -------------------------------------------------------------------------------------------------------------
@dataclass
class DeviceCounts:
    """cells x genes raw counts on the device, uint16, row-major with leading dimension `ld` (multiple of 64,
    padding columns are zero).  Stored in an int16 tensor (bit pattern is what matters)."""
    data: torch.Tensor
    C: int
    G: int                       # genes held by THIS rank
    comm: object = None          # GeneShardComm when `data` is the gene shard [comm.gene_lo, comm.gene_hi)

    is_device_counts = True     # lets AnnDataLite carry the device-resident matrix as `.X`
    ndim = 2

    @property
    def shape(self):
        """Shape of the GLOBAL matrix (cells x all genes), which is what the AnnData describes."""
        return (self.C, self.G if self.comm is None else self.comm.total_genes)

    def copy(self) -> "DeviceCounts":
        return DeviceCounts(self.data.clone(), self.C, self.G, self.comm)

    @property
    def ld(self) -> int:
        return self.data.shape[1]

    @property
    def device(self):
        return self.data.device

    def to_numpy(self) -> np.ndarray:
        return self.data[:, :self.G].cpu().numpy().view(np.uint16)


@dataclass
class HostShard:
    """Host-resident gene shard of a count matrix for the multi-GPU entry: `array` is cells x local genes
    (uint16, ideally pinned), `comm` says which genes of the global matrix they are."""
    array: np.ndarray
    comm: object

    is_device_counts = True
    ndim = 2

    @property
    def shape(self):
        return (self.array.shape[0], self.comm.total_genes)


def _check_flags(flags: torch.Tensor, comm=None):
    """Raise the ingest errors.  With a gene-shard communicator the flags are max-reduced first, so that every rank
    raises together (a rank-local raise would leave the other ranks waiting in the next collective)."""
    if comm is not None and comm.world > 1:
        comm.all_reduce_max(flags)
    f = flags.cpu().numpy()
    if f[0]:
        # same message as the reference, scGeneClust/_validation.py:70
        raise ValueError("Expected raw counts, got normalized counts possibly.")
    if f[1]:
        raise ValueError("Counts above 65535 are not supported by the uint16 device layout.")
    if f[2]:
        raise ValueError("Sparse count matrix has column indices outside [0, n_vars).")
    if f[4]:
        raise ValueError("Expected raw counts, got negative values (int16 input).")


def ingest_counts(X, device="cuda", rows_per_chunk: int = 8192, comm=None) -> DeviceCounts:
    """Host (or device) count matrix -> `DeviceCounts`.  Never densifies on the host (contrast
    scGeneClust/_model.py:124, _validation.py:68): CSR is expanded and validated on the device; a CSR matrix that is
    not in canonical form (unsorted or duplicate column indices) has its duplicates summed, as `toarray()` does.
    Host arrays are uploaded asynchronously when they sit in pinned memory."""
    require_cuda()
    L = lib()
    dev = torch.device(device)
    if isinstance(X, DeviceCounts):
        return X
    if isinstance(X, HostShard):
        if X.array.shape[1] != X.comm.local_genes:
            raise ValueError(f"host shard has {X.array.shape[1]} genes, the communicator assigns {X.comm.local_genes}")
        local = ingest_counts(X.array, device, rows_per_chunk, comm=X.comm)
        local.comm = X.comm
        return local
    if isinstance(X, torch.Tensor):
        if X.dtype not in (torch.int16, torch.uint16):
            raise TypeError("device tensors must already be uint16/int16 counts")
        C_, G_ = X.shape
        ld = _round_up(G_, 64)
        if X.device.type == "cuda" and G_ == ld and X.is_contiguous():
            return DeviceCounts(X.view(torch.int16), C_, G_)
        out = torch.zeros((C_, ld), dtype=torch.int16, device=dev)
        out[:, :G_] = X.view(torch.int16).to(dev)
        return DeviceCounts(out, C_, G_)
    C_, G_ = X.shape
    ld = _round_up(G_, 64)
    counts = torch.zeros((C_, ld), dtype=torch.int16, device=dev)
    flags = torch.zeros(8, dtype=torch.int32, device=dev)
    st = stream_handle()
    if issparse(X):
        X = X.tocsr()
        indptr = np.ascontiguousarray(X.indptr, dtype=np.int64)
        data = X.data if X.data.dtype == np.float32 else X.data.astype(np.float32)
        indices = X.indices if X.indices.dtype == np.int32 else X.indices.astype(np.int32)
        d_indptr = torch.from_numpy(indptr).to(dev, non_blocking=True)
        # one bulk copy of values and indices (asynchronous when the caller's arrays are in pinned memory)
        d_data = torch.from_numpy(np.ascontiguousarray(data)).to(dev, non_blocking=True)
        d_idx = torch.from_numpy(np.ascontiguousarray(indices)).to(dev, non_blocking=True)
        L.csr_to_counts(ptr(d_data), ptr(d_idx), ptr(d_indptr), C_, 0, G_, ptr(counts), ld, 0, ptr(flags), st)
        if int(flags[3].item()):
            # unsorted / duplicate indices: redo with summing stores (flags 0-2 are re-derived by the second pass)
            counts.zero_()
            flags.zero_()
            L.csr_to_counts(ptr(d_data), ptr(d_idx), ptr(d_indptr), C_, 0, G_, ptr(counts), ld, 1, ptr(flags), st)
    else:
        Xh = np.asarray(X)
        if Xh.dtype == np.uint16 or Xh.dtype == np.int16:
            src = torch.from_numpy(np.ascontiguousarray(Xh).view(np.int16))
            if G_ == ld:
                counts.copy_(src, non_blocking=True)
            else:   # strided device-side placement of a straight (pinned-speed) upload
                counts[:, :G_] = src.to(dev, non_blocking=True)
            if Xh.dtype == np.int16:                              # the uint16 layout would read a negative as >= 32768
                flags[4:5] = (counts.min() < 0).to(torch.int32)
        else:
            # float32 C-order input is uploaded as is, block by block (no host conversion pass; asynchronous from pinned
            # memory); other dtypes are converted block-wise on the host first
            rows = max(1, min(rows_per_chunk, 65535))
            for r0 in range(0, C_, rows):
                blk = np.ascontiguousarray(Xh[r0:r0 + rows], dtype=np.float32)
                d_blk = torch.from_numpy(blk).to(dev, non_blocking=True)
                L.dense_to_counts(ptr(d_blk), G_, blk.shape[0], G_, r0, ptr(counts), ld, ptr(flags), st)
    _check_flags(flags, comm)
    return DeviceCounts(counts, C_, G_)


# -------------------------------------------------------------------------------------------------------------
# implicit Pearson-residual matrix
# -------------------------------------------------------------------------------------------------------------
@dataclass
class ResidualOperator:
    """Z[c,g] = clip((x - mu)/sqrt(mu + mu^2/theta), +-sqrt(C)), mu = r_c s_g / T, as an operator on the device
    counts; optional PCA centring shifts are applied implicitly by the pass kernels."""
    counts: DeviceCounts
    row_sum: torch.Tensor      # int64 [C]   exact
    col_sum: torch.Tensor      # int64 [G]   exact
    row_sum_f: torch.Tensor    # float32 [C]
    col_sum_f: torch.Tensor    # float32 [G]
    total: float
    theta: float
    clip: float
    upper_clip_inactive: bool = False      # proved from the data: no residual reaches +clip (see from_counts)
    _scratch: Optional[torch.Tensor] = field(default=None, repr=False)

    @property
    def C(self) -> int:
        return self.counts.C

    @property
    def G(self) -> int:
        return self.counts.G

    @classmethod
    def from_counts(cls, counts: DeviceCounts, theta: float = THETA) -> "ResidualOperator":
        """With a gene-sharded `counts` (counts.comm set) the per-cell sums and the total are all-reduced so every
        shard uses the statistics of the full matrix."""
        L = lib()
        comm = counts.comm
        dev = counts.device
        row_sum = torch.zeros(counts.C, dtype=torch.int64, device=dev)
        col_sum = torch.zeros(counts.G, dtype=torch.int64, device=dev)
        col_max = torch.zeros(counts.G, dtype=torch.int32, device=dev)
        L.count_sums(ptr(counts.data), counts.C, counts.G, counts.ld, ptr(row_sum), ptr(col_sum), ptr(col_max),
                     stream_handle())
        n_cells = counts.C
        if comm is not None:
            comm.all_reduce_sum(row_sum)
        # Can a residual reach the upper clip sqrt(C)?  z = (x - mu) / sqrt(mu (1 + mu/theta)) < x / sqrt(mu), and
        # mu = r_c s_g / T >= r_min s_g / T, so gene g cannot if  x_max(g) / sqrt(r_min s_g / T) < clip.  When this holds for
        # every gene (typical: a count that large needs an outlier cell), the pass kernels skip the per-entry clip test;
        # the result is the same, proved rather than assumed.
        clip = float(np.float32(np.sqrt(n_cells)))
        # one call, one read-back: float32 copies of the sums and head = [total, any all-zero gene, any all-zero cell,
        # clip proof failed]
        row_sum_f = torch.empty(counts.C, dtype=torch.float32, device=dev)
        col_sum_f = torch.empty(counts.G, dtype=torch.float32, device=dev)
        head_d = torch.empty(4, dtype=torch.int64, device=dev)
        scratch = torch.empty(3 * 148, dtype=torch.int64, device=dev)
        L.count_stats(ptr(row_sum), counts.C, ptr(col_sum), ptr(col_max), counts.G, clip, ptr(row_sum_f),
                      ptr(col_sum_f), ptr(head_d), ptr(scratch), stream_handle())
        if comm is not None:
            # the all-zero tests are rank-local on sharded input: every rank must raise together or the others hang in
            # the next collective (the clip flag may differ between ranks: each picks its own kernel variant)
            comm.all_reduce_max(head_d[1:3])
        head = head_d.cpu().numpy()
        total = float(head[0])
        if total <= 0:
            raise ValueError("count matrix is empty")
        if head[1] or head[2]:
            # the reference yields NaN residuals here (0/0) and sklearn PCA then raises on NaN input
            raise ValueError("Input contains all-zero genes or cells: Pearson residuals would be NaN. "
                             "Filter them first (the reference's loader uses min_cells=3).")
        return cls(counts, row_sum, col_sum, row_sum_f, col_sum_f, total, float(theta), clip,
                   upper_clip_inactive=not bool(head[3]))

    def struct(self, row_shift=None, col_shift=None) -> ResidualMatrix:
        return ResidualMatrix(ptr(self.counts.data), self.C, self.G, self.counts.ld, ptr(self.row_sum_f),
                              ptr(self.col_sum_f), self.total, self.theta, self.clip, ptr(row_shift), ptr(col_shift),
                              1 if self.upper_clip_inactive else 0)

    def _scratch_buf(self) -> torch.Tensor:
        need = lib().rsvd_scratch_bytes(self.C, self.G)
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.empty(need, dtype=torch.uint8, device=self.counts.device)
        return self._scratch

    def apply(self, trans: int, Q: torch.Tensor, row_shift=None, col_shift=None, out=None, simt: bool = False,
              kernel: Optional[str] = None):
        """trans=0: (C x 64) = M @ Q(G x 64);  trans=1: (G x 64) = M^T @ Q(C x 64).
        kernel: 'ts' (default; tcgen05 with the residual operand in tensor memory), 'ss' (tcgen05, both operands from
        shared memory) or 'simt' (exact float32 FFMA) - the last two are cross-checks."""
        n_in, n_out = (self.G, self.C) if trans == 0 else (self.C, self.G)
        assert Q.shape == (n_in, Q_LD) and Q.dtype == torch.float32
        if out is None:
            out = torch.empty((n_out, Q_LD), dtype=torch.float32, device=Q.device)
        m = self.struct(row_shift, col_shift)
        kernel = kernel or ('simt' if simt else 'ts')
        fn = {'ts': lib().rsvd_apply, 'ss': lib().rsvd_apply_ss, 'simt': lib().rsvd_apply_simt}[kernel]
        fn(C.byref(m), trans, ptr(Q), ptr(out), Q_LD, ptr(self._scratch_buf()), stream_handle())
        return out

    def materialize(self) -> torch.Tensor:
        """Dense float32 residual matrix (C x G) on the device (ps path / return_info)."""
        Z = torch.empty((self.C, self.G), dtype=torch.float32, device=self.counts.device)
        lib().pearson_residuals(ptr(self.counts.data), self.C, self.G, self.counts.ld, ptr(self.row_sum_f),
                                ptr(self.col_sum_f), self.total, self.theta, self.clip, ptr(Z), self.G,
                                stream_handle())
        return Z

    def materialize_columns(self, gene_idx) -> torch.Tensor:
        """Dense float32 residual columns (C x len(gene_idx)) of the given genes (singleton-cluster rescue,
        scGeneClust/tl/selection.py:50)."""
        dev = self.counts.device
        gene_idx = np.asarray(gene_idx, dtype=np.int64)
        comm = self.counts.comm
        if comm is not None:      # sharded: every rank fills the columns of its own genes, then they are summed
            mine = (gene_idx >= comm.gene_lo) & (gene_idx < comm.gene_hi)
            Zall = torch.zeros((self.C, len(gene_idx)), dtype=torch.float32, device=dev)
            if mine.any():
                Zall[:, torch.from_numpy(np.flatnonzero(mine)).to(dev)] = self._local_columns(gene_idx[mine] - comm.gene_lo)
            return comm.all_reduce_sum(Zall)
        return self._local_columns(gene_idx)

    def _local_columns(self, local_idx: np.ndarray) -> torch.Tensor:
        dev = self.counts.device
        idx = torch.as_tensor(local_idx, dtype=torch.int64, device=dev)
        ns = int(idx.numel())
        ld = _round_up(max(ns, 1), 64)
        sub = torch.zeros((self.C, ld), dtype=torch.int16, device=dev)
        sub[:, :ns] = self.counts.data.index_select(1, idx)
        col_f = self.col_sum_f.index_select(0, idx).contiguous()
        Z = torch.empty((self.C, ns), dtype=torch.float32, device=dev)
        lib().pearson_residuals(ptr(sub), self.C, ns, ld, ptr(self.row_sum_f), ptr(col_f), self.total, self.theta,
                                self.clip, ptr(Z), ns, stream_handle())
        return Z

    def residual_sums(self, rows: bool = True, cols: bool = False):
        dev = self.counts.device
        zr = torch.zeros(self.C, dtype=torch.float64, device=dev) if rows else None
        zc = torch.zeros(self.G, dtype=torch.float64, device=dev) if cols else None
        lib().residual_sums(ptr(self.counts.data), self.C, self.G, self.counts.ld, ptr(self.row_sum_f),
                            ptr(self.col_sum_f), self.total, self.theta, self.clip, ptr(zr), ptr(zc), stream_handle())
        return zr, zc


# -------------------------------------------------------------------------------------------------------------
# randomized PCA on the operator (sklearn PCA(svd_solver='auto') -> 'randomized'; utils/extmath.py:313-386,560-626)
# -------------------------------------------------------------------------------------------------------------
class PanelAlgebra:
    """64-column panel helpers (Gram, Cholesky-QR, thin eigendecomposition) - all on the device."""

    def __init__(self, device, max_rows: int):
        self.dev = device
        self.scratch = torch.empty(lib().gram_scratch_bytes(max_rows), dtype=torch.uint8, device=device)
        self.gram = torch.empty((Q_LD, Q_LD), dtype=torch.float64, device=device)
        self.rinv = torch.empty((Q_LD, Q_LD), dtype=torch.float32, device=device)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)

    def gram_of(self, P: torch.Tensor, comm=None, replicated: bool = False) -> torch.Tensor:
        """self.gram = P^T P in float64.  `comm` + rows sharded (gene-space panels): local Gram, all-reduced.
        `comm` + `replicated` (cell-space panels, identical on every rank after their all-reduce): every rank takes
        the Gram of its own slice of the rows and the slices are all-reduced - the panel has as many rows as there
        are cells, so the replicated version made every rank pay C x 64 x 64 float64 FMAs per normalisation."""
        L, st = lib(), stream_handle()
        if comm is not None and replicated and comm.world > 1:
            lo, hi = comm.split_range(P.shape[0])
            if hi > lo:
                L.panel_gram(ptr(P[lo:hi]), hi - lo, ptr(self.gram), ptr(self.scratch), st)
            else:
                self.gram.zero_()
            comm.all_reduce_sum(self.gram)
            return self.gram
        L.panel_gram(ptr(P), P.shape[0], ptr(self.gram), ptr(self.scratch), st)
        if comm is not None:
            comm.all_reduce_sum(self.gram)
        return self.gram

    def chol_qr(self, P: torch.Tensor, q: int, out: torch.Tensor, comm=None, replicated: bool = False) -> torch.Tensor:
        """out = P @ R^-1 with P^T P = R^T R (one Cholesky-QR step); `comm` / `replicated` as in `gram_of`."""
        L, st = lib(), stream_handle()
        self.gram_of(P, comm, replicated)
        L.chol_inv(ptr(self.gram), q, ptr(self.rinv), ptr(self.status), st)
        L.panel_matmul(ptr(P), P.shape[0], ptr(self.rinv), ptr(out), st)
        return out


_SIDE_STREAMS: dict = {}


def _side_stream(dev) -> "torch.cuda.Stream":
    """One long-lived side stream per device (a fresh stream per call would give the caching allocator a fresh pool,
    i.e. a cudaMalloc + device synchronisation, every call)."""
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=key)
    return _SIDE_STREAMS[key]


class HostCopy:
    """Device tensor -> ndarray in pinned host memory, copied on the side stream so that the kernels queued after it on
    the main stream do not wait for PCIe.  `array` is valid after `wait()`."""

    def __init__(self, t: torch.Tensor):
        t = t.contiguous()
        self._pin = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        side = _side_stream(t.device)
        side.wait_stream(torch.cuda.current_stream(t.device))
        with torch.cuda.stream(side):
            self._pin.copy_(t, non_blocking=True)
            self._done = torch.cuda.Event()
            self._done.record(side)
        t.record_stream(side)
        self.array = self._pin.numpy()          # shares the pinned buffer (kept alive by the array's base)

    def wait(self) -> np.ndarray:
        self._done.synchronize()
        return self.array


def device_normal_panel(random_state: int, rows: int, q: int, device) -> torch.Tensor:
    """`np.random.RandomState(random_state).normal(size=(rows, q)).astype(np.float32)` as a zero-padded rows x 64
    device panel, generated on the device (gc_mt19937_normal_panel)."""
    L, st = lib(), stream_handle()
    state = np.random.RandomState(random_state).get_state()
    assert state[0] == 'MT19937' and state[3] == 0      # fresh seed: no cached gaussian
    key = torch.from_numpy(np.ascontiguousarray(state[1], dtype=np.uint32).view(np.int32)).to(device)
    panel = torch.zeros((rows, Q_LD), dtype=torch.float32, device=device)
    scratch = torch.empty(L.mt19937_normal_scratch_bytes(rows, q), dtype=torch.uint8, device=device)
    status = torch.zeros(1, dtype=torch.int32, device=device)
    L.mt19937_normal_panel(ptr(key), int(state[2]), rows, q, ptr(panel), Q_LD, ptr(scratch), ptr(status), st)
    panel._gc_status = status           # checked together with the Cholesky status at the end of the solver
    return panel


def pca_solver(n_samples: int, n_features: int, n_comps: int) -> str:
    """sklearn PCA 'auto' policy (decomposition/_pca.py:524-536)."""
    if n_features <= 1000 and n_samples >= 10 * n_features:
        return "covariance_eigh"
    if max(n_samples, n_features) <= 500:
        return "full"
    if 1 <= n_comps < 0.8 * min(n_samples, n_features):
        return "randomized"
    return "full"


def _small_pca(op: ResidualOperator, samples: str, solver: str, n_comps: int) -> torch.Tensor:
    """The two exact solvers sklearn's 'auto' policy picks for SMALL inputs (decomposition/_pca.py:524-536, :540-700):
    'covariance_eigh' (<= 1000 features and >= 10x more samples, e.g. gene PCA of a data set with a few hundred cells)
    and 'full' (both dimensions <= 500, or n_comps >= 0.8 min_dim).  These sizes are far from the hot path: the residual
    matrix is materialised and the dense factorisations are library calls (torch.linalg on the device, float64), followed
    by sklearn's v-based sign flip and its transform formulas.  Returns float32 scores (n_samples x n_comps)."""
    Z = op.materialize().to(torch.float64)
    X = Z.t().contiguous() if samples == "genes" else Z
    n = X.shape[0]
    mean = X.mean(dim=0)
    if solver == "covariance_eigh":
        Cm = X.t() @ X
        Cm -= n * torch.outer(mean, mean)
        Cm /= (n - 1)
        evals, evecs = torch.linalg.eigh(Cm)                      # ascending
        Vt = evecs.t().flip(0)[:n_comps].contiguous()
        idx = Vt.abs().argmax(dim=1)
        Vt *= torch.sign(Vt[torch.arange(Vt.shape[0], device=Vt.device), idx]).unsqueeze(1)      # svd_flip(u=None, v-based)
        scores = X @ Vt.t() - (mean.unsqueeze(0) @ Vt.t())        # PCA._transform
    elif solver == "full":
        U, S, Vt = torch.linalg.svd(X - mean, full_matrices=False)
        idx = Vt.abs().argmax(dim=1)
        signs = torch.sign(Vt[torch.arange(Vt.shape[0], device=Vt.device), idx])
        scores = (U * signs.unsqueeze(0))[:, :n_comps] * S[:n_comps]
    else:
        raise NotImplementedError(f"PCA solver '{solver}'")
    return scores.to(torch.float32).contiguous()


@dataclass
class PcaPlan:
    """What sklearn's PCA(svd_solver='auto') + randomized_svd will do for this shape (decomposition/_pca.py:524-536,
    utils/extmath.py:313-386) - known from the shape alone, before any data is touched."""
    solver: str
    n_comps: int
    q: int
    n_iter: int
    transpose: bool
    omega_rows: int            # A.shape[1] after sklearn's transposition = rows of the test matrix Omega


def plan_pca(n_cells: int, n_genes_total: int, samples: str, n_comps: Optional[int] = None) -> PcaPlan:
    n_samples, n_features = (n_genes_total, n_cells) if samples == "genes" else (n_cells, n_genes_total)
    min_dim = min(n_samples, n_features)
    if n_comps is None:
        n_comps = min_dim - 1 if N_PCS >= min_dim else N_PCS     # scanpy's n_comps rule
    solver = pca_solver(n_samples, n_features, n_comps)
    n_iter = 7 if n_comps < 0.1 * min_dim else 4
    transpose = n_samples < n_features
    # A (after sklearn's transpose) has the stored cells x genes orientation iff ...
    a_is_stored = transpose if samples == "genes" else (not transpose)
    return PcaPlan(solver, n_comps, n_comps + N_OVERSAMPLES, n_iter, transpose, n_genes_total if a_is_stored else n_cells)


@dataclass
class OmegaPrefetch:
    """An Omega panel being generated on the side stream (see `prefetch_omega`)."""
    key: tuple
    panel: torch.Tensor
    status: torch.Tensor
    stream: "torch.cuda.Stream"


def prefetch_omega(random_state: int, plan: PcaPlan, device) -> Optional[OmegaPrefetch]:
    """Start generating the solver's test matrix on the side stream NOW (the entry point calls this before the counts are
    even uploaded): the MT19937 word stream comes from one CTA, so it is the one piece of the PCA worth starting early."""
    if plan.solver != "randomized" or plan.q > Q_LD:
        return None
    dev = torch.device(device)
    main_stream = torch.cuda.current_stream(dev)
    side = _side_stream(dev)
    side.wait_stream(main_stream)
    with torch.cuda.stream(side):
        panel = device_normal_panel(random_state, plan.omega_rows, plan.q, dev)
    return OmegaPrefetch((int(random_state), plan.omega_rows, plan.q), panel, panel._gc_status, side)


def randomized_pca(op: ResidualOperator, samples: str, random_state: int = 0, n_comps: Optional[int] = None,
                   simt: bool = False, info: Optional[dict] = None, omega: Optional[OmegaPrefetch] = None) -> torch.Tensor:
    """PCA scores (n_samples x n_comps, float32, device) of the residual matrix.

    samples='genes' is `sc.pp.pca(norm_counts.T)` (scGeneClust/pp/_preprocessing.py:52): samples are genes,
    features are cells, centring = per-cell mean over genes.  samples='cells' is `sc.pp.pca(norm_counts)` (:54).
    Follows sklearn's randomized solver step for step (same Omega, n_iter rule, transpose rule, v-based sign
    flip); the LU/QR panel normalisers are replaced by Cholesky-QR, which spans the same subspace.
    With `comm` (gene-sharded operator) panels in cell space are all-reduced and panels in gene space stay local.
    `omega`: a test matrix already being generated (`prefetch_omega`); used if it was made for the same draw.
    """
    L, st = lib(), stream_handle()
    dev = op.counts.device
    comm = op.counts.comm
    G_total = op.G if comm is None else comm.total_genes
    plan = plan_pca(op.C, G_total, samples, n_comps)
    n_comps, solver, q, n_iter, transpose = plan.n_comps, plan.solver, plan.q, plan.n_iter, plan.transpose
    if solver != "randomized":
        if comm is not None:
            raise NotImplementedError(f"PCA solver '{solver}' (small matrices) is not available on gene-sharded input")
        scores = _small_pca(op, samples, solver, n_comps)
        if info is not None:
            info.update(solver=solver, n_comps=n_comps, passes=0)
        return scores
    if q > Q_LD:
        raise NotImplementedError("n_components + 10 must be <= 64")

    # Omega exactly as sklearn draws it (utils/extmath.py:323-334): normal((A.shape[1], q)) cast to float32 - the
    # legacy RandomState stream is reproduced on the device from the seeded MT19937 state (20-40 ms on the host).
    # Its word stream comes from ONE CTA (MT19937 is sequential), so it runs on a side stream.
    main_stream = torch.cuda.current_stream()
    if omega is None or omega.key != (int(random_state), plan.omega_rows, q):
        omega = prefetch_omega(random_state, plan, dev)
    Qp_full, omega_status, side_stream = omega.panel, omega.status, omega.stream

    a_is_stored = transpose if samples == "genes" else (not transpose)
    fwd = 0 if a_is_stored else 1          # trans flag of  A @ Q ; trans=0 contracts genes, trans=1 contracts cells
    bwd = 1 - fwd
    # centring vector = mean over samples of every feature (decomposition/_pca.py:733-735).  It is indexed by cells for
    # samples='genes' (row_shift) and by genes for samples='cells' (col_shift).  When sklearn transposes, that is the
    # OUTPUT index of the first pass, and the sums come out of that pass for free: Omega gets a column of ones in a
    # padding slot, so column 63 of the raw product is  Z 1  (resp. Z^T 1); the rank-1 centring term is then applied
    # to the result.  Otherwise the shift is needed INSIDE the first pass and costs one extra pass over the counts.
    ones_trick = transpose and q < Q_LD
    row_shift = col_shift = None
    if not ones_trick:
        if samples == "genes":
            zr, _ = op.residual_sums(rows=True, cols=False)
            if comm is not None:
                comm.all_reduce_sum(zr)
            row_shift = (zr / float(G_total)).to(torch.float32)
        else:
            _, zc = op.residual_sums(rows=False, cols=True)
            col_shift = (zc / float(op.C)).to(torch.float32)

    main_stream.wait_stream(side_stream)
    Qp_full.record_stream(main_stream)
    omega_status.record_stream(main_stream)
    Qp = Qp_full
    if comm is not None and fwd == 0:
        Qp = Qp[comm.gene_lo:comm.gene_hi].contiguous()

    alg = PanelAlgebra(dev, max(op.C, op.G))
    # Gene-sharded input: a cell-space panel (cells x 64) is the SUM over the gene shards of every rank's partial product.
    # It is never all-reduced and normalised redundantly: reduce-scatter hands every rank one row chunk of the sum, the
    # Cholesky-QR runs on that chunk (64 x 64 Gram all-reduced), and the normalised chunks are all-gathered - the same
    # NVLink bytes as the all-reduce, 1/world of the panel algebra per rank.
    # GC_PANEL_EXCHANGE=allreduce keeps the previous scheme (all-reduce + replicated normalisation) for A/B timing
    import os as _os
    sliced = comm is not None and comm.world > 1 and _os.environ.get("GC_PANEL_EXCHANGE", "rs_ag") != "allreduce"
    replicated_ar = comm is not None and comm.world > 1 and not sliced
    if sliced:
        chunk = comm.row_chunk(op.C)
        c_pad = chunk * comm.world
        valid = max(0, min(chunk, op.C - comm.rank * chunk))           # real rows of this rank's chunk
        tmp_c_pad = torch.zeros((c_pad, Q_LD), dtype=torch.float32, device=dev)
        buf_c_pad = torch.zeros((c_pad, Q_LD), dtype=torch.float32, device=dev)
        tmp_c, buf_c = tmp_c_pad[:op.C], buf_c_pad[:op.C]
        sl_in = torch.zeros((chunk, Q_LD), dtype=torch.float32, device=dev)
        sl_a, sl_b = torch.zeros_like(sl_in), torch.zeros_like(sl_in)
    else:
        buf_c = torch.empty((op.C, Q_LD), dtype=torch.float32, device=dev)
        tmp_c = torch.empty_like(buf_c)
    buf_g = torch.empty((op.G, Q_LD), dtype=torch.float32, device=dev)
    tmp_g = torch.empty_like(buf_g)

    def apply(trans, Qin):
        """Raw pass.  Gene-sharded trans = 0 output is this rank's PARTIAL sum (in tmp_c): `normalise` reduces it."""
        out = tmp_c if trans == 0 else tmp_g
        op.apply(trans, Qin, row_shift, col_shift, out=out, simt=simt)
        if replicated_ar and trans == 0:
            comm.all_reduce_sum(out)
        return out

    def chol_qr_chunk(cur, n_steps=1):
        """Cholesky-QR of the row-sliced panel whose chunk on this rank is `cur` (chunk x 64, zero pad rows); returns the
        normalised FULL panel (all-gathered into buf_c)."""
        for _ in range(n_steps):
            nxt = sl_a if cur is not sl_a else sl_b
            if valid > 0:
                alg.chol_qr(cur[:valid], q, nxt[:valid], comm, replicated=False)
            else:                                    # a rank past the end of the panel still joins the Gram all-reduce
                alg.gram.zero_()
                comm.all_reduce_sum(alg.gram)
            cur = nxt
        comm.all_gather_chunks(cur, buf_c_pad)
        return buf_c

    def normalise(trans, Y, n_steps=1):
        """Orthonormalise the columns of the panel Y (Cholesky-QR, `n_steps` rounds).  trans == 1: rows = this rank's genes
        (sharded, local Gram all-reduced); trans == 0: rows = all cells."""
        if trans == 0 and sliced:
            comm.reduce_scatter_rows(tmp_c_pad, sl_in)         # Y is tmp_c: this rank's partial sums
            return chol_qr_chunk(sl_in, n_steps)
        cur, dst = Y, (buf_c if trans == 0 else buf_g)
        for step in range(n_steps):
            out = dst if step % 2 == 0 else Y                  # ping-pong: the second round writes back into Y's buffer
            if trans == 0 and replicated_ar:
                alg.chol_qr(cur, q, out, comm, replicated=True)     # every rank: Gram of its row slice, all-reduced
            else:
                alg.chol_qr(cur, q, out, comm if trans == 1 else None, replicated=False)
            cur = out
        return cur

    Qcur = Qp
    for it in range(n_iter):
        if it == 0 and ones_trick:
            # raw first pass with a column of ones in Omega's last padding slot: column 63 of the product is Z 1
            n_k_total = G_total if fwd == 0 else op.C
            Qp[:, Q_LD - 1] = 1.0
            Y = apply(fwd, Qp)
            colsum = Qp.sum(dim=0, dtype=torch.float64).to(torch.float32)
            if replicated_ar and fwd == 0:
                comm.all_reduce_sum(colsum)                    # (Y itself was all-reduced by apply)
            if sliced and fwd == 0:
                comm.all_reduce_sum(colsum)                    # Omega's rows are sharded like the genes
                comm.reduce_scatter_rows(tmp_c_pad, sl_in)
                part = sl_in                                   # this rank's chunk of the summed raw product
            else:
                part = Y
            shift = (part[:, Q_LD - 1] / float(n_k_total)).contiguous()
            torch.addcmul(part, shift.unsqueeze(1), colsum.unsqueeze(0), value=-1.0, out=part)
            part[:, q:] = 0.0
            if sliced and fwd == 0:
                shift_full = torch.empty(c_pad, dtype=torch.float32, device=dev)
                comm.all_gather_chunks(shift, shift_full)
                shift = shift_full[:op.C].contiguous()
                Qcur = chol_qr_chunk(part)
            else:
                Qcur = normalise(fwd, part)
            if samples == "genes":
                row_shift = shift
            else:
                col_shift = shift
        else:
            Qcur = normalise(fwd, apply(fwd, Qcur))
        Qcur = normalise(bwd, apply(bwd, Qcur))
    # final range: orthonormal basis by two Cholesky-QR steps (qr_normalizer of extmath.py:386)
    Qf = normalise(fwd, apply(fwd, Qcur), n_steps=2).clone()
    Bt = apply(bwd, Qf)                                  # B^T = A^T Q  (A.shape[1] x q)
    # thin SVD of B through the q x q eigenproblem  B B^T = Uhat S^2 Uhat^T  (float64, on the device)
    if sliced and bwd == 0:
        comm.all_reduce_sum(Bt)                          # cell-space B^T (sklearn did not transpose): needed in full
    # (replicated_ar: apply() already all-reduced it)
    alg.gram_of(Bt, comm, replicated=(bwd == 0))
    evals = torch.empty(Q_LD, dtype=torch.float64, device=dev)
    uhat = torch.empty((Q_LD, Q_LD), dtype=torch.float32, device=dev)
    L.sym_eig(ptr(alg.gram), q, ptr(evals), ptr(uhat), None, st)
    sing = torch.sqrt(torch.clamp(evals, min=0.0))
    sign_src = torch.empty(Q_LD, dtype=torch.float32, device=dev)
    if not transpose:
        # U = Q Uhat, Vt = S^-1 Uhat^T B ; scores = U S ; sign from the rows of Vt = columns of B^T Uhat
        L.panel_colabsmax(ptr(Bt), Bt.shape[0], ptr(uhat), ptr(sign_src), ptr(alg.scratch), st)
        if comm is not None and bwd == 1:
            sign_src = comm.absmax_merge(sign_src)
        L.scale_cols(ptr(uhat), ptr(sign_src), ptr(sing), n_comps, st)
        base = Qf
    else:
        # U = (S^-1 Uhat^T B)^T, Vt = (Q Uhat)^T ; scores = U S = B^T Uhat ; sign from the columns of Q Uhat
        L.panel_colabsmax(ptr(Qf), Qf.shape[0], ptr(uhat), ptr(sign_src), ptr(alg.scratch), st)
        if comm is not None and fwd == 1:
            sign_src = comm.absmax_merge(sign_src)
        L.scale_cols(ptr(uhat), ptr(sign_src), None, n_comps, st)
        base = Bt
    scores = torch.empty((base.shape[0], Q_LD), dtype=torch.float32, device=dev)
    L.panel_matmul(ptr(base), base.shape[0], ptr(uhat), ptr(scores), st)
    flags = int((alg.status + omega_status).item())
    if flags & 1:
        raise FloatingPointError("Cholesky-QR panel normalisation hit a non-positive pivot (rank-deficient panel)")
    if flags & 2:
        raise RuntimeError("device Omega generation ran out of pre-generated MT19937 words")
    if info is not None:
        info.update(n_iter=n_iter, transpose=bool(transpose), q=q, n_comps=n_comps,
                    singular_values=sing[:n_comps].cpu().numpy(), passes=2 * n_iter + 2)
    return scores[:, :n_comps].contiguous()
