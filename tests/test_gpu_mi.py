"""GPU parity of the kNN mutual-information kernels (a6, a7, a8) against the golden vectors produced by the
reference's own functions and against the brute-force oracle.  Bar: integer neighbour counts bit-exact, MI bit-exact
(same fp64 operations in numpy's summation order)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests")
    from scgeneclust_b200._lib import lib
    lib()
    return torch


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_redundancy_golden_bit_exact(torch_cuda, golden_dir):
    torch = torch_cuda
    from oracle import mi_bruteforce as mb
    from scgeneclust_b200.tl.information import gene_redundancy, prepare_pair_roles
    g = _load(golden_dir, "redundancy_small.npz")
    E, seed, ref = g["E"], int(g["seed"]), g["redundancy"]
    N, n = E.shape
    dE = torch.from_numpy(E).cuda()
    XR, YR = prepare_pair_roles(dE, seed)
    xi1, xi2 = mb.noise_vectors(n, seed)
    XR_o = np.stack([mb.prep_feature(E[i], xi1) for i in range(N)])
    YR_o = np.stack([mb.prep_target(E[i], xi2) for i in range(N)])
    assert np.array_equal(XR.cpu().numpy(), XR_o)          # input preparation is bit-exact
    assert np.array_equal(YR.cpu().numpy(), YR_o)
    cond, D, nx, ny = gene_redundancy(dE, seed, return_counts=True)
    iu = np.triu_indices(N, 1)
    nx_o, ny_o, mi_o = mb.ksg_batch(XR_o, YR_o, np.stack(iu, 1))
    assert np.array_equal(nx.cpu().numpy().astype(np.int16), nx_o)
    assert np.array_equal(ny.cpu().numpy().astype(np.int16), ny_o)
    assert np.array_equal(cond.cpu().numpy(), ref[iu])
    Dh = D.cpu().numpy()
    assert np.array_equal(Dh, ref)
    assert np.array_equal(Dh, Dh.T) and not Dh.diagonal().any()   # tests/test_GeneClust.py:23


@pytest.mark.parametrize("N,n", [(257, 50), (64, 49), (33, 64), (40, 7)])
def test_redundancy_random_vs_oracle(torch_cuda, N, n):
    torch = torch_cuda
    from oracle import mi_bruteforce as mb
    from scgeneclust_b200.tl.information import gene_redundancy
    rs = np.random.RandomState(N * 100 + n)
    E = (rs.standard_normal((N, n)) * (1.0 / np.sqrt(1 + np.arange(n)))).astype(np.float32)
    E[3] = E[1]                      # identical genes
    E[5] = np.round(E[5] * 4) / 4    # ties
    E[7, :] = 0.125                  # constant gene (std -> 1 branches)
    seed = 3
    xi1, xi2 = mb.noise_vectors(n, seed)
    XR_o = np.stack([mb.prep_feature(E[i], xi1) for i in range(N)])
    YR_o = np.stack([mb.prep_target(E[i], xi2) for i in range(N)])
    iu = np.triu_indices(N, 1)
    nx_o, ny_o, mi_o = mb.ksg_batch(XR_o, YR_o, np.stack(iu, 1))
    cond, D, nx, ny = gene_redundancy(torch.from_numpy(E).cuda(), seed, return_counts=True)
    assert np.array_equal(nx.cpu().numpy().astype(np.int16), nx_o)
    assert np.array_equal(ny.cpu().numpy().astype(np.int16), ny_o)
    assert np.array_equal(cond.cpu().numpy(), mi_o)
    # partial ranges (the multi-GPU split) reproduce the same values
    P = len(mi_o)
    lo, hi = P // 3, P // 3 + P // 2
    part, _ = gene_redundancy(torch.from_numpy(E).cuda(), seed, lo, hi, dense=False)
    assert np.array_equal(part.cpu().numpy(), mi_o[lo:hi])


def test_redundancy_vs_sklearn_sample(torch_cuda):
    """Direct check against the call the reference makes (information.py:62-63) on a sample of pairs."""
    torch = torch_cuda
    from oracle import stages
    from scgeneclust_b200.tl.information import gene_redundancy
    rs = np.random.RandomState(5)
    N, n = 300, 50
    E = rs.standard_normal((N, n)).astype(np.float32) * 3
    _, D = gene_redundancy(torch.from_numpy(E).cuda(), 0)
    Dh = D.cpu().numpy()
    pairs = np.stack([rs.randint(0, N, 60), rs.randint(0, N, 60)], 1)
    pairs = np.sort(pairs[pairs[:, 0] != pairs[:, 1]], 1)
    ref = stages.gene_redundancy(E, 0, pairs)
    assert np.array_equal(Dh[pairs[:, 0], pairs[:, 1]], ref)


def test_redundancy_full_size_properties(torch_cuda):
    """Config-c4 size (N = 4000 -> 7 998 000 pairs): size-independent properties + a sample against the oracle."""
    torch = torch_cuda
    from oracle import mi_bruteforce as mb
    from scgeneclust_b200.tl.information import gene_redundancy
    rs = np.random.RandomState(9)
    N, n = 4000, 50
    E = (rs.standard_normal((N, n)) * (2.0 / np.sqrt(1 + np.arange(n)))).astype(np.float32)
    cond, D = gene_redundancy(torch.from_numpy(E).cuda(), 0)
    assert bool((D == D.t()).all()) and float(D.diagonal().abs().max()) == 0.0 and float(D.min()) >= 0.0
    iu = torch.triu_indices(N, N, 1, device=D.device)
    assert torch.equal(D[iu[0], iu[1]], cond)
    xi1, xi2 = mb.noise_vectors(n, 0)
    pairs = np.sort(np.stack([rs.randint(0, N, 400), rs.randint(0, N, 400)], 1), 1)
    pairs = pairs[pairs[:, 0] != pairs[:, 1]]
    XR_o = np.stack([mb.prep_feature(E[i], xi1) for i in pairs[:, 0]])
    YR_o = np.stack([mb.prep_target(E[j], xi2) for j in pairs[:, 1]])
    _, _, mi_o = mb.ksg_batch(XR_o, YR_o, np.stack([np.arange(len(pairs))] * 2, 1))
    assert np.array_equal(D.cpu().numpy()[pairs[:, 0], pairs[:, 1]], mi_o)


def test_relevance_golden_bit_exact(torch_cuda, golden_dir):
    torch = torch_cuda
    from oracle import mi_bruteforce as mb
    from scgeneclust_b200 import AnnDataLite
    from scgeneclust_b200.tl.information import find_relevant_genes, gene_relevance
    g = _load(golden_dir, "relevance_small.npz")
    Z, labels, seed = g["Z"], g["labels"], int(g["seed"])
    mi, m_all, keep = gene_relevance(torch.from_numpy(Z).cuda(), labels, seed, return_counts=True)
    assert np.array_equal(mi.cpu().numpy(), g["relevance_all"])
    xi1, _ = mb.noise_vectors(Z.shape[0], seed, continuous_target=False)
    m_h = m_all.cpu().numpy()
    for j in (0, 5, 20, 21, 22, 59):
        keep_o, _, _, m_o = mb.ross_counts(mb.prep_feature(Z[:, j], xi1), labels)
        assert np.array_equal(keep_o, keep)
        assert np.array_equal(m_h[keep, j], m_o)
    ad = AnnDataLite(Z.copy())
    ad.obs["cluster"] = labels
    find_relevant_genes(ad, int(g["top_pct"]), 1, seed)
    assert np.array_equal(np.asarray(ad.var_names, dtype=str), g["kept_genes"])
    assert np.array_equal(ad.var["relevance"].to_numpy(), g["kept_relevance"])


def test_relevance_large_vs_sklearn_sample(torch_cuda):
    """n_hc = 3000 cells (pairwise-sum recursion, multi-tile kernels) against mutual_info_classif."""
    torch = torch_cuda
    from oracle import stages
    from scgeneclust_b200.tl.information import gene_relevance
    rs = np.random.RandomState(21)
    n, G = 3000, 96
    types = rs.randint(0, 8, n)
    Z = rs.standard_normal((n, G)).astype(np.float32)
    Z[:, :30] += (types[:, None] * rs.standard_normal((1, 30)) * 0.3).astype(np.float32)
    Z[:, 40] = np.where(rs.rand(n) < 0.7, -0.2, Z[:, 40])     # many exact ties (zero-inflated gene)
    labels = np.array([f"t{t}" for t in types])
    mi = gene_relevance(torch.from_numpy(Z).cuda(), labels, 0).cpu().numpy()
    cols = [0, 7, 29, 40, 41, 95]
    assert np.array_equal(mi[cols], stages.gene_relevance(Z[:, cols], labels, 0))


def test_complementarity_golden_bit_exact(torch_cuda, golden_dir):
    torch = torch_cuda
    from scgeneclust_b200.tl.information import gene_complementarity
    g = _load(golden_dir, "complementarity_small.npz")
    got = gene_complementarity(torch.from_numpy(g["Z"]).cuda(), g["labels"], g["edges"], int(g["seed"]))
    assert np.array_equal(got, g["complementarity"])


@pytest.mark.parametrize("n", [50, 64, 33, 7])
def test_sorted_scan_pairs_equal_brute_force(torch_cuda, n):
    """Two independent device implementations of the KSG pair estimator - the production brute-force kernel and the
    sorted outward scan with early exit - against each other and the numpy oracle: neighbour counts nx, ny and MI
    bit-identical, incl. duplicate values (ties) and tiny n."""
    torch = torch_cuda
    from oracle import mi_bruteforce as mb
    from scgeneclust_b200.tl.information import gene_redundancy
    rs = np.random.RandomState(n)
    N = 300
    E = (rs.standard_normal((N, n)) * (2.0 / np.sqrt(1 + np.arange(n)))).astype(np.float32)
    E[:40] = np.round(E[:40], 1)                       # heavy ties in 40 genes
    E[40:50, : n // 2] = E[40:50, n // 2: 2 * (n // 2)]  # exact duplicates inside a gene
    Ed = torch.from_numpy(E).cuda()
    c_s, D_s, nx_s, ny_s = gene_redundancy(Ed, 3, return_counts=True, sorted_scan=True)
    c_b, D_b, nx_b, ny_b = gene_redundancy(Ed, 3, return_counts=True)
    assert torch.equal(nx_s, nx_b) and torch.equal(ny_s, ny_b)
    assert torch.equal(c_s, c_b) and torch.equal(D_s, D_b)
    xi1, xi2 = mb.noise_vectors(n, 3)
    XR = np.stack([mb.prep_feature(e, xi1) for e in E[:60]])
    YR = np.stack([mb.prep_target(e, xi2) for e in E[:60]])
    pairs = np.stack(np.triu_indices(60, 1), 1)
    nx_o, ny_o, mi_o = mb.ksg_batch(XR, YR, pairs)
    iu = np.triu_indices(N, 1)
    sel = np.flatnonzero((iu[0] < 60) & (iu[1] < 60))
    assert np.array_equal(nx_s.cpu().numpy()[sel], nx_o) and np.array_equal(ny_s.cpu().numpy()[sel], ny_o)
    assert np.array_equal(c_s.cpu().numpy()[sel], mi_o)


def _kruskal_total_order(R):
    """Spanning forest under the strict order (larger R first, ties by smaller upper-triangle edge id): host reference."""
    N = R.shape[0]
    iu = np.triu_indices(N, 1)
    w = R[iu]
    keep = np.flatnonzero(w != 0)
    order = keep[np.lexsort((keep, -w[keep]))]
    parent = list(range(N))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    out = []
    for k in order:
        a, b = int(iu[0][k]), int(iu[1][k])
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[ra] = rb
            out.append((a, b))
    out = np.array(sorted(out), dtype=np.int64).reshape(-1, 2)
    return out


@pytest.mark.parametrize("N,zero_frac,ties", [(300, 0.45, False), (257, 0.45, True), (120, 0.97, False), (2, 0.0, False)])
def test_device_mst_equals_total_order_kruskal(torch_cuda, N, zero_frac, ties):
    """Device Boruvka spanning forest of the MI > 0 graph: identical edge set to Kruskal under the same strict edge
    order (so ties are resolved identically), same total weight as scipy's MST, forests when the graph is disconnected."""
    torch = torch_cuda
    from scipy.sparse.csgraph import connected_components
    from scgeneclust_b200.tl.information import redundancy_mst, redundancy_mst_device
    rs = np.random.RandomState(N)
    A = rs.rand(N, N)
    if ties:
        A = np.round(A, 2)                                   # many exactly equal weights
    A[rs.rand(N, N) < zero_frac] = 0.0
    R = np.triu(A, 1)
    R = R + R.T
    got = redundancy_mst_device(torch.from_numpy(R).cuda())
    ref = _kruskal_total_order(R)
    assert np.array_equal(got, ref)
    n_comp, _ = connected_components(R != 0, directed=False)
    assert len(got) == N - n_comp
    sp_edges = redundancy_mst(R)
    assert len(sp_edges) == len(got)
    assert np.isclose(R[got[:, 0], got[:, 1]].sum(), R[sp_edges[:, 0], sp_edges[:, 1]].sum(), rtol=1e-12)
    if not ties:
        assert np.array_equal(got, sp_edges)


@pytest.mark.parametrize("N,tile,ties", [(64, 500, False), (700, 50_000, True), (1500, 300_000, False)])
def test_streamed_forest_equals_dense_forest(torch_cuda, N, tile, ties):
    """`redundancy_forest` (MI tiles folded into the spanning forest by gc_mst_boruvka_edges; no N x N matrix) finds the
    forest the dense routine finds on the materialised matrix - same edges, and the weights it carries are the matrix
    entries - for tile sizes that split the pair range unevenly; with many exactly equal MI values too (a coarse
    embedding makes ties), where only the shared strict edge order makes the forest unique."""
    torch = torch_cuda
    from scgeneclust_b200.tl.information import gene_redundancy, redundancy_forest, redundancy_mst_device
    rs = np.random.RandomState(N)
    E = (rs.standard_normal((N, 50)) * (3.0 / (1 + np.arange(50)))).astype(np.float32)
    if ties:
        E = np.round(E, 1)                                   # few distinct coordinates: many tied MI values
    Ed = torch.from_numpy(E).cuda()
    _, D = gene_redundancy(Ed, 0)
    ref = redundancy_mst_device(D)
    edges, weights, n_pairs = redundancy_forest(Ed, 0, None, tile_pairs=tile)
    assert n_pairs == N * (N - 1) // 2
    assert np.array_equal(edges, ref)
    assert np.array_equal(weights, D.cpu().numpy()[edges[:, 0], edges[:, 1]])


def test_ps_entry_streamed_equals_dense(torch_cuda):
    """`scGeneClust(version='ps')` without return_info streams the redundancy stage (no varp['redundancy']); the selected
    genes equal those of the return_info=True run, which builds the dense matrix like the reference."""
    from oracle.synth import cell_types, synth_counts
    from scgeneclust_b200 import AnnDataLite, scGeneClust
    from scgeneclust_b200.synth import synth_params
    p = synth_params(900, 0)
    X = synth_counts(p, 0, 1300, 0, 900)
    keep = (X > 0).sum(0) >= 3
    X = np.ascontiguousarray(X[:, keep]).astype(np.float32)
    names = np.array([f"g{i}" for i in np.flatnonzero(keep)], dtype=object)

    def make():
        ad = AnnDataLite(X, var_names=names)
        ad.obs['cluster'] = np.array([f"type{t}" for t in cell_types(p, 0, 1300)])
        return ad
    info, genes_dense = scGeneClust(make(), version='ps', n_obs_clusters=8, return_info=True, verbosity=0)
    genes_stream = scGeneClust(make(), version='ps', n_obs_clusters=8, verbosity=0)
    assert 'redundancy' in info.varp
    assert np.array_equal(genes_dense, genes_stream)
