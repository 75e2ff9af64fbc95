"""GeneClust-ps information-theoretic stages with the reference's names and semantics
(scGeneClust/tl/information.py:21-132) on the device: relevance (Ross kNN MI of every gene against the pseudo cell
types), redundancy (KSG kNN MI of every gene pair on the 50-d gene embedding), MST, and complementarity
(per-cell-cluster KSG MI on the MST edges).  The estimators are the exact fp64 ones scikit-learn 1.9.0 runs for the
reference (`mutual_info_classif` / `mutual_info_regression`, feature_selection/_mutual_info.py:20-153, 294-315);
their integer neighbour counts are bit-identical."""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch
from loguru import logger
from scipy.special import digamma

from .._lib import lib, ptr, require_cuda, stream_handle
from ..pp._preprocessing import get_state

K_NEIGHBORS = 3


def _psi_table(n_max: int, device) -> torch.Tensor:
    """psi[m] = digamma(m) for m = 1..n_max (psi[0] unused), from scipy like the reference's estimator."""
    t = np.zeros(n_max + 1, dtype=np.float64)
    t[1:] = digamma(np.arange(1, n_max + 1, dtype=np.float64))
    return torch.from_numpy(t).to(device)


def _noise(n: int, random_state: int, target: bool):
    """The normals `_estimate_mi` draws from a fresh RandomState(seed) (feature_selection/_mutual_info.py:296-315):
    xi1 for the single feature column, then xi2 for a continuous target."""
    rng = np.random.RandomState(random_state)
    xi1 = rng.standard_normal(size=(n, 1))[:, 0]
    xi2 = rng.standard_normal(size=n) if target else None
    return xi1, xi2


def _device_matrix(adata, st, attr: str, host) -> torch.Tensor:
    from .._anndata_lite import DeviceMatrix, ShardedMatrix
    if isinstance(host, (DeviceMatrix, ShardedMatrix)):
        return host.tensor                      # ShardedMatrix: this rank's gene columns
    t = getattr(st, attr, None)
    if t is not None and tuple(t.shape) == tuple(np.shape(host)):
        return t
    return torch.from_numpy(np.ascontiguousarray(host, dtype=np.float32)).cuda()


# -------------------------------------------------------------------------------------------------------------
# a6 relevance
# -------------------------------------------------------------------------------------------------------------
def gene_relevance(Z: torch.Tensor, labels, random_state: int = 0, return_counts: bool = False):
    """Ross MI of every column of Z (n cells x G genes, float32, device) against `labels` - what
    `_compute_relevance` (scGeneClust/tl/information.py:21-24) returns for every gene.  float64 tensor [G]."""
    require_cuda()
    L, stream = lib(), stream_handle()
    dev = Z.device
    n, G = Z.shape
    labels = np.asarray(labels)
    uniq, inv, cnt = np.unique(labels, return_inverse=True, return_counts=True)
    label_k = np.where(cnt > 1, np.minimum(K_NEIGHBORS, cnt - 1), 0).astype(np.int32)
    keep = cnt[inv] > 1
    n_kept = int(keep.sum())
    if n_kept < 2:
        raise ValueError("relevance needs at least one label with more than one cell")
    k_all = label_k[inv][keep].astype(np.float64)
    label_counts = cnt[inv][keep].astype(np.float64)
    const_term = (digamma(n_kept) + np.mean(digamma(k_all))) - np.mean(digamma(label_counts))
    xi1, _ = _noise(n, random_state, target=False)

    order = torch.arange(n, dtype=torch.int32, device=dev)
    seg_off = torch.tensor([0, n], dtype=torch.int32, device=dev)
    d_xi1 = torch.from_numpy(xi1).to(dev)
    XR = torch.empty((G, n), dtype=torch.float64, device=dev)
    L.mi_prepare(ptr(Z), Z.stride(0), ptr(order), ptr(seg_off), 1, G, n, ptr(d_xi1), None, ptr(XR), None, stream)
    d_label = torch.from_numpy(inv.astype(np.int32)).to(dev)
    d_label_k = torch.from_numpy(label_k).to(dev)
    m_all = torch.empty((n, G), dtype=torch.int32, device=dev)
    mi = torch.empty(G, dtype=torch.float64, device=dev)
    psi = _psi_table(n, dev)
    L.mi_cd(ptr(XR), n, G, n, ptr(d_label), ptr(d_label_k), len(uniq), float(const_term), ptr(psi),
            ptr(m_all), ptr(mi), stream)
    if return_counts:
        return mi, m_all, keep
    return mi


def find_relevant_genes(adata, top_pct: int, max_workers: int = os.cpu_count() - 1, random_state: int = 0):
    """scGeneClust/tl/information.py:27-57: `var['relevance']`, then in-place subset to the genes whose relevance is
    strictly above the (100 - top_pct) percentile."""
    logger.info("Finding relevant genes...")
    st = get_state(adata)
    Z = _device_matrix(adata, st, 'Z', adata.X)
    relevance = gene_relevance(Z, np.asarray(adata.obs['cluster']), random_state)
    if st.comm is not None and st.comm.world > 1:
        relevance = st.comm.all_gather_rows(relevance)          # every rank scored its own genes
    relevance = relevance.cpu().numpy()
    logger.debug("Gene relevance computed!")
    rlv_th = np.percentile(relevance, 100 - top_pct)
    logger.opt(colors=True).debug(f"Relevance threshold: <yellow>{rlv_th}</yellow>")
    is_relevant = relevance > rlv_th
    adata._inplace_subset_var(is_relevant)
    adata.var['relevance'] = relevance[is_relevant]
    keep = torch.from_numpy(np.flatnonzero(is_relevant)).to(Z.device)
    from .._anndata_lite import DeviceMatrix, ShardedMatrix
    st.Z = adata.X.tensor if isinstance(adata.X, (DeviceMatrix, ShardedMatrix)) else Z.index_select(1, keep).contiguous()
    if st.varm_pca is not None:
        st.varm_pca = st.varm_pca.index_select(0, keep).contiguous()
    logger.opt(colors=True).info(
        f"<yellow>{is_relevant.sum()}</yellow> (<yellow>{top_pct}%</yellow>) genes are marked as relevant genes."
    )


# -------------------------------------------------------------------------------------------------------------
# a7 redundancy
# -------------------------------------------------------------------------------------------------------------
def prepare_pair_roles(E: torch.Tensor, random_state: int = 0):
    """Feature-role (float64 path) and target-role (float32 path) versions of every row of E (N x n float32):
    what `_estimate_mi` hands to `_compute_mi_cc` as x resp. y (feature_selection/_mutual_info.py:294-315)."""
    L, stream = lib(), stream_handle()
    dev = E.device
    N, n = E.shape
    xi1, xi2 = _noise(n, random_state, target=True)
    Et = E.t().contiguous()                                   # samples x genes, the layout gc_mi_prepare reads
    order = torch.arange(n, dtype=torch.int32, device=dev)
    seg_off = torch.tensor([0, n], dtype=torch.int32, device=dev)
    XR = torch.empty((N, n), dtype=torch.float64, device=dev)
    YR = torch.empty((N, n), dtype=torch.float64, device=dev)
    d_xi1, d_xi2 = torch.from_numpy(xi1).to(dev), torch.from_numpy(xi2).to(dev)   # keep alive across the launch
    L.mi_prepare(ptr(Et), Et.stride(0), ptr(order), ptr(seg_off), 1, N, n, ptr(d_xi1), ptr(d_xi2), ptr(XR), ptr(YR),
                 stream)
    return XR, YR


def gene_redundancy(E: torch.Tensor, random_state: int = 0, p_begin: int = 0, p_end: Optional[int] = None,
                    dense: bool = True, return_counts: bool = False, sorted_scan: bool = False, roles=None):
    """KSG MI of the gene pairs [p_begin, p_end) in `itertools.combinations(range(N), 2)` order - the values
    `_compute_redundancy` (scGeneClust/tl/information.py:60-63) returns.  Returns (condensed[p_end-p_begin],
    dense N x N symmetric or None[, nx, ny uint8 counts])."""
    require_cuda()
    L, stream = lib(), stream_handle()
    dev = E.device
    N, n = E.shape
    n_pairs = N * (N - 1) // 2
    p_end = n_pairs if p_end is None else p_end
    XR, YR = roles if roles is not None else prepare_pair_roles(E, random_state)
    cond = torch.empty(p_end - p_begin, dtype=torch.float64, device=dev)
    D = torch.zeros((N, N), dtype=torch.float64, device=dev) if dense else None
    nx = torch.empty((p_end - p_begin, n), dtype=torch.uint8, device=dev) if return_counts else None
    ny = torch.empty_like(nx) if return_counts else None
    psi = _psi_table(n, dev)
    if not sorted_scan:
        # all n^2 distances per pair.  Measured faster than the outward scan on the 50-d PCA embeddings GeneClust feeds it
        # (37.7 vs 50.9 ms for 7 998 000 pairs): a few large-coordinate PCs are far from everything, their lanes walk the
        # whole sorted array, and the warp waits for its slowest lane
        L.mi_cc_pairs(ptr(XR), ptr(YR), N, n, n, ptr(psi), p_begin, p_end, ptr(cond), ptr(D), ptr(nx), ptr(ny), stream)
    else:                # independent second implementation of the same estimator: the cross-check of the tests
        scratch = torch.empty(L.mi_cc_pairs_scratch_bytes(N, n), dtype=torch.uint8, device=dev)
        L.mi_cc_pairs_sorted(ptr(XR), ptr(YR), N, n, n, ptr(psi), p_begin, p_end, ptr(cond), ptr(D), ptr(nx), ptr(ny),
                             ptr(scratch), stream)
    if return_counts:
        return cond, D, nx, ny
    return cond, D


TILE_PAIRS = 4 * 1024 * 1024        # gene pairs per streamed tile (32 MB of float64 MI values)


class _ForestBuffers:
    """Ping-pong device buffers of a spanning forest (<= N - 1 weighted edges) + the Boruvka scratch."""

    def __init__(self, N: int, device):
        L = lib()
        m = max(N - 1, 1)
        self.N, self.dev = N, device
        self.u = [torch.empty(m, dtype=torch.int32, device=device) for _ in range(2)]
        self.v = [torch.empty(m, dtype=torch.int32, device=device) for _ in range(2)]
        self.w = [torch.empty(m, dtype=torch.float64, device=device) for _ in range(2)]
        self.n = [torch.zeros(1, dtype=torch.int32, device=device) for _ in range(2)]
        self.scratch = torch.empty(L.mst_scratch_bytes(N), dtype=torch.uint8, device=device)
        self.cur, self.count = 0, 0

    def fold(self, cond: Optional[torch.Tensor], p_begin: int, extra=None):
        """forest <- MSF(forest u condensed tile [p_begin, p_begin + len(cond)) u extra edges (u, v, w tensors))."""
        L, stream = lib(), stream_handle()
        c, o = self.cur, self.cur ^ 1
        if extra is not None:
            eu = torch.cat([self.u[c][:self.count], extra[0]]).contiguous()
            ev = torch.cat([self.v[c][:self.count], extra[1]]).contiguous()
            ew = torch.cat([self.w[c][:self.count], extra[2]]).contiguous()
        else:
            eu, ev, ew = self.u[c], self.v[c], self.w[c]
        n_extra = self.count + (0 if extra is None else int(extra[0].numel()))
        n_cond = 0 if cond is None else int(cond.numel())
        L.mst_boruvka_edges(ptr(cond), p_begin, n_cond, ptr(eu), ptr(ev), ptr(ew), n_extra, self.N, ptr(self.u[o]),
                            ptr(self.v[o]), ptr(self.w[o]), ptr(self.n[o]), ptr(self.scratch), stream)
        self.cur = o
        self.count = int(self.n[o].item())

    def edges(self):
        c = self.cur
        return self.u[c][:self.count], self.v[c][:self.count], self.w[c][:self.count]


def redundancy_forest(E: torch.Tensor, random_state: int = 0, comm=None, tile_pairs: int = TILE_PAIRS):
    """Maximum-redundancy spanning forest of the relevant genes WITHOUT the N x N matrix (north_star: "no dense N x N
    materialisation ... top-k merge"): the KSG MI values of the gene pairs are produced tile by tile (gc_mi_cc_pairs)
    and each tile is folded into the forest found so far (gc_mst_boruvka_edges; MSF(E1 u E2) = MSF(MSF(E1) u E2)).
    With a gene-shard communicator every rank streams its own slice of the pair range, the per-rank forests
    (<= N - 1 edges each) are all-gathered and reduced once more, identically on every rank.
    Returns (edges int64 (n, 2) sorted by (source, target), weights float64 (n,), n_pairs)."""
    require_cuda()
    dev = E.device
    N = E.shape[0]
    n_pairs = N * (N - 1) // 2
    lo, hi = (0, n_pairs) if comm is None or comm.world == 1 else comm.split_range(n_pairs)
    roles = prepare_pair_roles(E, random_state)
    forest = _ForestBuffers(N, dev)
    for t0 in range(lo, hi, tile_pairs):
        t1 = min(hi, t0 + tile_pairs)
        cond, _ = gene_redundancy(E, random_state, t0, t1, dense=False, roles=roles)
        forest.fold(cond, t0)
    if comm is not None and comm.world > 1:
        import torch.distributed as dist
        m = max(N - 1, 1)
        u, v, w = forest.edges()
        pack = torch.zeros((m, 3), dtype=torch.float64, device=dev)           # weight 0 = padding (a non-edge)
        pack[:forest.count, 0], pack[:forest.count, 1], pack[:forest.count, 2] = u.double(), v.double(), w
        parts = [torch.empty_like(pack) for _ in range(comm.world)]
        dist.all_gather(parts, pack, group=comm.group)
        allp = torch.cat(parts)
        merged = _ForestBuffers(N, dev)
        merged.fold(None, 0, extra=(allp[:, 0].to(torch.int32).contiguous(), allp[:, 1].to(torch.int32).contiguous(),
                                    allp[:, 2].contiguous()))
        forest = merged
    u, v, w = forest.edges()
    e = torch.stack([u, v], 1).cpu().numpy().astype(np.int64)
    wh = w.cpu().numpy()
    order = np.lexsort((e[:, 1], e[:, 0]))
    return e[order], wh[order], n_pairs


def compute_gene_redundancy(adata, max_workers: int = os.cpu_count() - 1, random_state: int = 0):
    """scGeneClust/tl/information.py:66-85: `varp['redundancy']` = symmetric N x N float64, zero diagonal.  With a
    gene-shard communicator in the state, ranks compute disjoint condensed ranges and all-gather them.
    When the entry point does not have to hand the matrix back (`return_info=False`) it is never built: the MI tiles are
    folded straight into the spanning forest the next stage needs (`redundancy_forest`), which is kept in the state."""
    logger.info("Computing gene redundancy...")
    st = get_state(adata)
    E = _device_matrix(adata, st, 'varm_pca', adata.varm['X_pca'])
    N = E.shape[0]
    comm = st.comm
    if not st.dense_redundancy:
        st.forest = redundancy_forest(E, random_state, comm)[:2]
        st.redundancy = None
        logger.info("Gene redundancy computed (streamed into the spanning forest; no N x N matrix).")
        return
    if comm is None or comm.world == 1:
        _, D = gene_redundancy(E, random_state)
    else:
        n_pairs = N * (N - 1) // 2
        lo, hi = comm.split_range(n_pairs)
        cond, _ = gene_redundancy(E, random_state, lo, hi, dense=False)
        full = comm.all_gather_ranges(cond, n_pairs)
        D = torch.zeros((N, N), dtype=torch.float64, device=E.device)
        iu = torch.triu_indices(N, N, 1, device=E.device)
        D[iu[0], iu[1]] = full
        D = D + D.t()
    st.redundancy = D
    st.forest = None
    from .._anndata_lite import AnnDataLite, DeviceMatrix
    # 8 N^2 bytes (128 MB at N = 4000): stays on the device for the MST and the edge scaling; downloaded when read
    adata.varp['redundancy'] = DeviceMatrix(D) if isinstance(adata, AnnDataLite) else D.cpu().numpy()
    logger.info("Gene redundancy computed.")


# -------------------------------------------------------------------------------------------------------------
# MST (scGeneClust/tl/information.py:124-126) - "next" row of the scope table; host side for now
# -------------------------------------------------------------------------------------------------------------
def redundancy_mst(R: np.ndarray) -> np.ndarray:
    """Maximum-redundancy spanning tree (forest, if the MI > 0 graph is disconnected) of the graph that has an
    edge wherever R != 0 - igraph `Weighted_Adjacency(-R, 'undirected').spanning_tree(weights)`.  Returns the
    (n_edges, 2) int array ordered like igraph's edge ids (upper triangle, row-major), source < target."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import minimum_spanning_tree
    iu = np.triu_indices(R.shape[0], 1)
    w = R[iu]
    nz = w != 0
    graph = csr_matrix((-w[nz], (iu[0][nz], iu[1][nz])), shape=R.shape)
    mst = minimum_spanning_tree(graph).tocoo()
    src, dst = np.minimum(mst.row, mst.col), np.maximum(mst.row, mst.col)
    order = np.lexsort((dst, src))
    return np.stack([src[order], dst[order]], axis=1).astype(np.int64)


def redundancy_mst_device(D: torch.Tensor) -> np.ndarray:
    """The same spanning forest from the device-resident redundancy matrix (gc_mst_boruvka: Boruvka under the strict
    edge order "larger MI first, ties by smaller upper-triangle edge id", which makes the forest unique).  With distinct
    MI values it is the forest scipy / igraph find; among exactly tied values the three libraries may pick different,
    equally heavy edges.  Returns (n_edges, 2) int64 sorted by (source, target), source < target."""
    require_cuda()
    L, stream = lib(), stream_handle()
    N = D.shape[0]
    assert D.dtype == torch.float64 and D.is_contiguous() and D.shape == (N, N)
    edges = torch.empty((max(N - 1, 1), 2), dtype=torch.int32, device=D.device)
    n_edges = torch.zeros(1, dtype=torch.int32, device=D.device)
    scratch = torch.empty(L.mst_scratch_bytes(N), dtype=torch.uint8, device=D.device)
    L.mst_boruvka(ptr(D), N, ptr(edges), ptr(n_edges), ptr(scratch), stream)
    e = edges[:int(n_edges.item())].cpu().numpy().astype(np.int64)
    return e[np.lexsort((e[:, 1], e[:, 0]))]


# -------------------------------------------------------------------------------------------------------------
# a8 complementarity
# -------------------------------------------------------------------------------------------------------------
def gene_complementarity(Z: torch.Tensor, clusters, edges: np.ndarray, random_state: int = 0) -> np.ndarray:
    """`_compute_complementarity` (scGeneClust/tl/information.py:88-98) for every edge (i, j): sum over the cell
    clusters c of KSG-MI(Z[c, i], Z[c, j]) * n_c / n.  Z: n cells x N genes float32 on the device."""
    require_cuda()
    L, stream = lib(), stream_handle()
    dev = Z.device
    n, N = Z.shape
    clusters = np.asarray(clusters)
    uniq, inv, cnt = np.unique(clusters, return_inverse=True, return_counts=True)
    if cnt.min() <= K_NEIGHBORS:
        raise ValueError("every cell cluster needs more than 3 cells for the k = 3 neighbour estimator")
    order = np.argsort(inv, kind="stable").astype(np.int32)     # cells grouped by cluster, original order inside
    seg_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    xi1 = np.empty(n, np.float64)
    xi2 = np.empty(n, np.float64)
    for c, m in enumerate(cnt):                                 # fresh RandomState(seed) per call, length n_c
        a, b = _noise(int(m), random_state, target=True)
        xi1[seg_off[c]:seg_off[c + 1]], xi2[seg_off[c]:seg_off[c + 1]] = a, b
    edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    genes = np.unique(edges)                                    # only the genes on an edge are prepared
    pos = np.full(N, -1, np.int64)
    pos[genes] = np.arange(len(genes))
    Zs = Z.index_select(1, torch.from_numpy(genes).to(dev)).contiguous()
    Gs = len(genes)
    XR = torch.empty((Gs, n), dtype=torch.float64, device=dev)
    YR = torch.empty((Gs, n), dtype=torch.float64, device=dev)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    # every device argument is bound to a name: a temporary would be freed (and its block reused by the next
    # upload) before the kernel runs
    d_order, d_seg, d_xi1, d_xi2 = t(order), t(seg_off), t(xi1), t(xi2)
    L.mi_prepare(ptr(Zs), Zs.stride(0), ptr(d_order), ptr(d_seg), len(uniq), Gs, n, ptr(d_xi1), ptr(d_xi2), ptr(XR),
                 ptr(YR), stream)
    n_edges, n_c = len(edges), len(uniq)
    e_i = np.repeat(pos[edges[:, 0]], n_c)
    e_j = np.repeat(pos[edges[:, 1]], n_c)
    seg = np.tile(np.arange(n_c), n_edges)
    x_off = e_i * n + seg_off[seg]
    y_off = e_j * n + seg_off[seg]
    length = cnt[seg].astype(np.int32)
    out_off = np.concatenate([[0], np.cumsum(length[:-1], dtype=np.int64)]).astype(np.int64)
    total = int(length.sum(dtype=np.int64))
    nx = torch.empty(total, dtype=torch.int32, device=dev)
    ny = torch.empty(total, dtype=torch.int32, device=dev)
    mi = torch.empty(n_edges * n_c, dtype=torch.float64, device=dev)
    d_xoff, d_yoff, d_len, d_ooff = t(x_off.astype(np.int64)), t(y_off.astype(np.int64)), t(length), t(out_off)
    psi = _psi_table(int(cnt.max()), dev)
    L.mi_cc_problems(ptr(XR), ptr(YR), ptr(d_xoff), ptr(d_yoff), ptr(d_len), ptr(d_ooff), n_edges * n_c, ptr(psi),
                     ptr(nx), ptr(ny), ptr(mi), stream)
    mi_h = mi.cpu().numpy().reshape(n_edges, n_c)
    cmi = np.zeros(n_edges)
    for c in range(n_c):                                        # the reference's accumulation order (np.unique order)
        cmi = cmi + mi_h[:, c] * cnt[c] / n
    return cmi


def compute_gene_complementarity(adata, max_workers: int = os.cpu_count() - 1, random_state: int = 0):
    """scGeneClust/tl/information.py:101-132: `uns['mst_edges']`, `uns['mst_edges_complm']`."""
    logger.info("Computing gene complementarity...")
    st = get_state(adata)
    logger.debug("Building MST...")
    D = st.redundancy
    if st.forest is not None:                                   # streamed: the forest came out of the redundancy stage
        adata.uns['mst_edges'] = st.forest[0]
    elif D is not None and tuple(D.shape) == tuple(np.shape(adata.varp['redundancy'])):
        adata.uns['mst_edges'] = redundancy_mst_device(D)
    else:
        adata.uns['mst_edges'] = redundancy_mst(np.asarray(adata.varp['redundancy']))
    logger.debug("MST built.")
    from .._anndata_lite import ShardedMatrix
    edges = adata.uns['mst_edges']
    clusters = np.asarray(adata.obs['cluster'])
    comm = st.comm
    if isinstance(adata.X, ShardedMatrix) and comm is not None and comm.world > 1:
        # the two genes of an edge may live on different ranks: the residual columns of the RELEVANT genes (cells x N,
        # small) are all-gathered, the edges are split across the ranks and their values gathered
        Z = adata.X.gather_columns()
        lo, hi = comm.split_range(len(edges))
        local = gene_complementarity(Z, clusters, edges[lo:hi], random_state) if hi > lo else np.zeros(0)
        cm = comm.all_gather_ranges(torch.from_numpy(np.ascontiguousarray(local)).to(Z.device), len(edges))
        adata.uns['mst_edges_complm'] = cm.cpu().numpy()
    else:
        Z = _device_matrix(adata, st, 'Z', adata.X)
        adata.uns['mst_edges_complm'] = gene_complementarity(Z, clusters, edges, random_state)
    logger.info("Gene complementarity computed.")
