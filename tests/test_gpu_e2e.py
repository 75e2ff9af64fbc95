"""End-to-end parity of `scGeneClust(adata, version='fast'|'ps')` through the public entry point against the CPU
oracle on the same seeded synthetic counts."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests")
    return torch


@pytest.fixture(scope="module")
def data():
    from oracle.synth import cell_types, synth_counts
    from scgeneclust_b200.synth import synth_params
    p = synth_params(2500, 0)
    X = synth_counts(p, 0, 3000, 0, 2500)
    keep = (X > 0).sum(0) >= 3
    names = np.array([f"gene{i}" for i in np.flatnonzero(keep)], dtype=object)
    return np.ascontiguousarray(X[:, keep]), names, cell_types(p, 0, 3000)


def test_fast_end_to_end_vs_oracle(torch_cuda, data):
    from sklearn.metrics import adjusted_rand_score
    from oracle import stages
    from scgeneclust_b200 import AnnDataLite, scGeneClust
    X, names, _ = data
    K = 40
    ref = stages.geneclust_fast(X, names, K, 0)
    for X_in in (X.astype(np.float32), sp.csr_matrix(X.astype(np.float32))):
        raw = AnnDataLite(X_in, var_names=names)
        info, sel = scGeneClust(raw, n_var_clusters=K, version='fast', return_info=True, verbosity=0)
        gpu = dict(X_pca=info.varm['X_pca'], labels=info.var['cluster'].to_numpy(),
                   closeness=info.var['closeness'].to_numpy(), selected=sel)
        par = stages.fast_parity(gpu, ref, names, K, 0, Z=ref['Z'])
        print(f"fast e2e: {par}")
        assert par["pca_rel_err_max"] <= 1e-3                            # north_star: PCA scores within 1e-3 relative
        assert par["tail_exact"], par                                    # k-means / closeness / selection: exact on the same embedding
        assert set(sel) <= set(names) and len(sel) <= 2 * K + (np.bincount(info.var['cluster']) == 1).sum()
        assert info.X.shape == X.shape and info.X.dtype == np.float32          # residuals (return_info contract)
        assert np.allclose(info.X, ref['Z'], rtol=2e-5, atol=2e-6)
        cl = info.var['closeness'].to_numpy()
        assert cl.max() <= 1.0 and cl.min() >= -1e-6
    # given the oracle's embedding, the k-means + selection tail must reproduce the oracle's list exactly
    from scgeneclust_b200.tl import cluster_genes, select_from_clusters
    ad = AnnDataLite(ref['Z'].copy(), var_names=names)
    ad.varm['X_pca'] = ref['X_pca']
    cluster_genes(ad, None, 'fast', n_gene_clusters=K, random_state=0)
    lab_same = np.array_equal(ad.var['cluster'].to_numpy(), ref['labels'])
    sel2 = select_from_clusters(ad, 'fast', True, 0)
    assert lab_same, f"labels differ (ARI {adjusted_rand_score(ad.var['cluster'].to_numpy(), ref['labels']):.4f})"
    assert np.array_equal(sel2, ref['selected'])


def test_return_modes(torch_cuda, data):
    from scgeneclust_b200 import AnnDataLite, scGeneClust
    X, names, _ = data
    raw = AnnDataLite(X[:, :800].astype(np.float32), var_names=names[:800])
    sel = scGeneClust(raw, n_var_clusters=20, verbosity=0)
    assert isinstance(sel, np.ndarray) and raw.n_vars == 800          # input untouched
    assert scGeneClust(raw, n_var_clusters=20, subset=True, verbosity=0) is None
    assert raw.n_vars == len(sel) and set(raw.var_names) == set(sel)


def test_ps_stages_vs_oracle(torch_cuda, data):
    """GeneClust-ps with the pseudo cell types given: relevance, redundancy, complementarity equal the oracle's
    (sklearn calls) on the device-computed residuals/embedding."""
    from oracle import stages
    from scgeneclust_b200 import AnnDataLite
    from scgeneclust_b200 import pp
    from scgeneclust_b200.tl.confidence import find_high_confidence_cells
    from scgeneclust_b200.tl.information import (compute_gene_complementarity, compute_gene_redundancy,
                                                 find_relevant_genes)
    X, names, types = data
    X, names = X[:1200, :600], names[:600]
    ok = (X > 0).sum(0) >= 3
    X, names = np.ascontiguousarray(X[:, ok]), names[ok]
    ad = AnnDataLite(X.astype(np.float32), var_names=names)
    ad.obs['cluster'] = np.array([f"type{t}" for t in types[:1200]])
    hc = np.ones(1200, bool)
    hc[::3] = False
    ad.obs['highly_confident'] = hc
    pp.normalize(ad, 'sc', materialize=True)
    pp.reduce_dim(ad, 'ps', 0)
    find_high_confidence_cells(ad, 8)
    assert ad.n_obs == hc.sum()
    Z = np.asarray(ad.X).copy()                  # adata.X is device-backed until a host array is asked for
    labels = np.asarray(ad.obs['cluster'])
    find_relevant_genes(ad, 20, 1, 0)
    rel_ref = stages.gene_relevance(Z, labels, 0)
    mask = stages.relevant_mask(rel_ref, 20)
    assert np.array_equal(np.asarray(ad.var_names), names[mask])
    assert np.array_equal(ad.var['relevance'].to_numpy(), rel_ref[mask])
    compute_gene_redundancy(ad, 1, 0)
    R = np.asarray(ad.varp['redundancy'])       # device-backed until read
    E = ad.varm['X_pca']
    rs = np.random.RandomState(0)
    pairs = np.sort(np.stack([rs.randint(0, len(E), 80), rs.randint(0, len(E), 80)], 1), 1)
    pairs = pairs[pairs[:, 0] != pairs[:, 1]]
    assert np.array_equal(R[pairs[:, 0], pairs[:, 1]], stages.gene_redundancy(E, 0, pairs))
    assert np.array_equal(R, R.T)
    compute_gene_complementarity(ad, 1, 0)
    edges = ad.uns['mst_edges']
    assert edges.shape[1] == 2 and len(edges) <= ad.n_vars - 1
    sample = edges[:: max(1, len(edges) // 12)]
    ref_c = stages.gene_complementarity(np.asarray(ad.X), labels, sample, 0)
    assert np.array_equal(ad.uns['mst_edges_complm'][:: max(1, len(edges) // 12)], ref_c)


def test_ps_entry_point_end_to_end(torch_cuda, data):
    """`scGeneClust(adata, version='ps', n_obs_clusters=8, return_info=True)` with the pseudo cell types supplied in
    `obs['cluster']`: every slot of the reference's out-contract is filled (scGeneClust/_model.py:86-99), the redundancy
    matrix is exactly symmetric (the reference's own test, tests/test_GeneClust.py:22-23), the MST spans the relevant
    genes, and the HDBSCAN tail agrees with scikit-learn's fork run on the same MST."""
    from sklearn.cluster._hdbscan import _linkage, _tree
    from scgeneclust_b200 import AnnDataLite, scGeneClust
    X, names, types = data
    X, names = X[:1500, :700], names[:700]
    ok = (X > 0).sum(0) >= 3
    X, names = np.ascontiguousarray(X[:, ok]), names[ok]
    raw = AnnDataLite(X.astype(np.float32), var_names=names)
    raw.obs['cluster'] = np.array([f"type{t}" for t in types[:1500]])
    info, genes = scGeneClust(raw, version='ps', n_obs_clusters=8, return_info=True, verbosity=0)
    assert genes.shape[0] > 0 and set(genes) <= set(names)                   # reference test: genes.shape[0] > 0
    n = info.n_vars
    R = info.varp['redundancy']
    assert R.shape == (n, n) and np.array_equal(R, R.T) and not np.diag(R).any()
    assert info.uns['mst_edges'].shape == (n - 1, 2) and info.uns['mst_edges_complm'].shape == (n - 1,)
    for col in ('relevance', 'outlier_score', 'cluster', 'representative'):
        assert col in info.var
    assert np.array_equal(np.asarray(info.var_names)[info.var['representative'].to_numpy()], genes)
    # the tail: rebuild the scaled MST exactly as scGeneClust/tl/cluster.py:122-126 does and hand it to sklearn's fork
    e = info.uns['mst_edges']
    rel = info.var['relevance'].to_numpy()
    scale = np.maximum(np.minimum(rel[e[:, 0]], rel[e[:, 1]]), info.uns['mst_edges_complm']) / R[e[:, 0], e[:, 1]]
    order = np.argsort(scale)
    rec = np.zeros(n - 1, dtype=_linkage.MST_edge_dtype)
    rec['current_node'], rec['next_node'], rec['distance'] = e[order, 0], e[order, 1], scale[order]
    ref_labels, _ = _tree.tree_to_labels(_linkage.make_single_linkage(rec), 3, "eom", False, 0.0, None)
    assert np.array_equal(info.var['cluster'].to_numpy(), ref_labels)
    s = info.var['outlier_score'].to_numpy()
    assert (s >= 0).all() and (s <= 1).all()


@pytest.mark.parametrize("C,m", [(3001, 1), (3001, 2), (50000, 1), (50000, 7), (20011, 40), (4099, 300), (131075, 5)])
def test_device_deviance_equals_numpy(torch_cuda, C, m):
    """Singleton rescue: the device binomial deviance must be the reference's numpy expression
    (scGeneClust/tl/selection.py:82-89) on float32 residual-like columns, including its NaN handling; equality up to
    the last-bit difference between numpy's SIMD logf and CUDA's logf summed over C terms (tolerance 2e-5 relative)."""
    torch = torch_cuda
    from scgeneclust_b200.tl.selection import compute_deviance, device_deviance
    rs = np.random.RandomState(C + m)
    X = (rs.standard_normal((C, m)) * rs.uniform(0.5, 3, size=m) + rs.uniform(-0.2, 0.4, size=m)).astype(np.float32)
    X[rs.rand(C, m) < 0.3] = np.float32(-0.31)                 # many equal negative residuals (zero counts)
    X[rs.rand(C, m) < 0.01] = 0.0
    ref = compute_deviance(X)
    got = device_deviance(torch.from_numpy(X).cuda())
    assert got.dtype == np.float32 and got.shape == (m,)
    # a 4-byte-aligned view of the same block takes the narrow-copy path of the column-sum kernel: same bits
    flat = torch.zeros(C * m + 1, dtype=torch.float32, device="cuda")
    flat[1:].view(C, m).copy_(torch.from_numpy(X))
    assert np.array_equal(device_deviance(flat[1:].view(C, m)).view(np.uint32), got.view(np.uint32))
    # left and right halves (~1e4) cancel to ~1e2..1e3, which amplifies the last-bit logf differences to ~1e-6 relative
    # (with 300 columns the worst column cancels harder: 1e-4)
    assert np.allclose(got, ref, rtol=2e-5 if m <= 64 else 1e-4, atol=0, equal_nan=True), (got, ref)
    print(f"deviance C={C} m={m}: identical {int((got == ref).sum())}/{m}, max rel {np.max(np.abs(got - ref) / np.abs(ref)):.1e}")


def test_fast_c1_shape_more_genes_than_cells_vs_oracle(torch_cuda):
    """BASELINE.json config c1 shape class (pbmc3k: 2 700 cells, ~13 700 genes after the loader's min_cells=3 filter;
    the real file is not shipped, so a synthetic matrix of that shape): sklearn does NOT transpose here, the randomized
    solver runs the passes the other way round; ragged sizes (not multiples of 64 / 128)."""
    from sklearn.metrics import adjusted_rand_score
    from oracle import stages
    from oracle.synth import synth_counts
    from scgeneclust_b200 import AnnDataLite, scGeneClust
    from scgeneclust_b200.synth import synth_params
    p = synth_params(6001, 1)
    X = np.concatenate([synth_counts(p, c0, min(450, 1350 - c0), 0, 6001) for c0 in range(0, 1350, 450)])
    keep = (X > 0).sum(0) >= 3
    X = np.ascontiguousarray(X[:, keep])
    assert X.shape[1] > X.shape[0]
    names = np.array([f"g{i}" for i in np.flatnonzero(keep)], dtype=object)
    K = 30
    ref = stages.geneclust_fast(X, names, K, 0)
    info, sel = scGeneClust(AnnDataLite(X.astype(np.float32), var_names=names), n_var_clusters=K, version='fast',
                            return_info=True, verbosity=0)
    rel = np.linalg.norm(info.varm['X_pca'] - ref['X_pca'], axis=0) / np.linalg.norm(ref['X_pca'], axis=0)
    ari = adjusted_rand_score(info.var['cluster'].to_numpy(), ref['labels'])
    assert rel.max() <= 1e-3, f"PCA rel err {rel.max():.2e}"
    assert np.array_equal(info.var['cluster'].to_numpy(), ref['labels']), f"gene clusters differ (ARI {ari:.4f})"
    assert np.array_equal(sel, ref['selected'])


def test_argument_errors_match_reference(torch_cuda):
    """Error types / messages of scGeneClust/_validation.py, raised before any device work."""
    from scgeneclust_b200 import AnnDataLite, scGeneClust
    ad = AnnDataLite(np.ones((4, 5), np.float32))
    with pytest.raises(ValueError, match="version"):
        scGeneClust(ad, version='slow', n_var_clusters=3)
    with pytest.raises(TypeError, match="n_gene_clusters"):
        scGeneClust(ad, version='fast')
    with pytest.raises(ValueError, match="> 1"):
        scGeneClust(ad, version='fast', n_var_clusters=1)
    with pytest.raises(ValueError, match="spatial"):
        scGeneClust(ad, version='fast', n_var_clusters=3, modality='st')
    with pytest.raises(TypeError, match="AnnData"):
        scGeneClust(np.ones((4, 5)), version='fast', n_var_clusters=3)
    with pytest.raises(TypeError, match="n_obs_clusters"):
        scGeneClust(ad, version='ps')
    ad.raw = object()
    with pytest.raises(ValueError, match="raw"):
        scGeneClust(ad, version='fast', n_var_clusters=3)


@pytest.fixture(scope="module")
def pbmc3k_substitute():
    """the bundled pbmc3k_raw.h5ad is absent, so configs c1 / c2 run on the sanctioned substitute
    `synth(2700, 32738, seed 0)` at pbmc3k's sequencing depth (gene means scaled to ~2.4 k counts per cell) with the
    reference loader's min_cells = 3 gene filter (scGeneClust/_utils.py:35)."""
    from oracle import synth as osy
    p = osy.synth_params(32738, 0)
    p.gene_mean = p.gene_mean * 0.1
    X = osy.synth_counts_c(p, 0, 2700, 0, 32738)
    keep = (X > 0).sum(0) >= 3
    names = np.array([f"g{i}" for i in np.flatnonzero(keep)], dtype=object)
    return np.ascontiguousarray(X[:, keep]), names, osy.cell_types(p, 0, 2700)


def test_config_c1_true_shape_fast(torch_cuda, pbmc3k_substitute):
    """BASELINE.json config c1 at its true shape: GeneClust-fast, 2 700 cells x 32 738 genes before the filter,
    n_var_clusters = 200; sklearn does not transpose (more genes than cells)."""
    from sklearn.metrics import adjusted_rand_score
    from oracle import stages
    from scgeneclust_b200 import AnnDataLite, scGeneClust
    X, names, _ = pbmc3k_substitute
    assert X.shape[0] == 2700 and X.shape[1] > 13_000
    ref = stages.geneclust_fast(X, names, 200, 0)
    info, sel = scGeneClust(AnnDataLite(sp.csr_matrix(X.astype(np.float32)), var_names=names), n_var_clusters=200,
                            version='fast', return_info=True, verbosity=0)
    gpu = dict(X_pca=info.varm['X_pca'], labels=info.var['cluster'].to_numpy(), closeness=info.var['closeness'].to_numpy(),
               selected=sel)
    par = stages.fast_parity(gpu, ref, names, 200, 0, Z=ref['Z'])
    print(f"c1 true shape {X.shape}: {par}")
    # Link-wise bars (tests/test_gpu_subsample.py explains why end-to-end label identity is reported, not asserted: at
    # this shape two runs of this test with numerically equivalent kernels gave ARI 0.9999 and 0.067 against the oracle -
    # one early k-means++ draw falling on a neighbouring sample re-partitions the mostly unstructured genes):
    assert par["pca_rel_err_max"] <= 1e-3 and par["pca_rel_err_first10"] <= 1e-4         # link 1: float tolerance
    assert par["tail_exact"], par                                                        # link 2 on the GPU's embedding
    from scgeneclust_b200.tl import cluster_genes, select_from_clusters                  # link 2 on the oracle's embedding
    ad = AnnDataLite(ref['Z'].copy(), var_names=names)
    ad.varm['X_pca'] = ref['X_pca']
    cluster_genes(ad, None, 'fast', n_gene_clusters=200, random_state=0)
    assert np.array_equal(ad.var['cluster'].to_numpy(), ref['labels'])
    assert np.allclose(ad.var['closeness'].to_numpy(), ref['closeness'], rtol=0, atol=1e-5)
    assert np.array_equal(select_from_clusters(ad, 'fast', True, 0), ref['selected'])


def test_config_c2_true_shape_ps(torch_cuda, pbmc3k_substitute):
    """BASELINE.json config c2 at its true shape: GeneClust-ps, n_obs_clusters = 8, 20 % relevant genes, pseudo cell
    types given (the generator's types on a fixed 60 % high-confidence subset).  Sampled genes / pairs / edges are checked
    bit for bit against the oracle's sklearn calls on the device-computed residuals and embedding."""
    from oracle import stages
    from scgeneclust_b200 import AnnDataLite, scGeneClust
    X, names, types = pbmc3k_substitute
    raw = AnnDataLite(X.astype(np.float32), var_names=names)
    raw.obs['cluster'] = np.array([f"type{t}" for t in types])
    raw.obs['highly_confident'] = np.random.RandomState(0).rand(2700) < 0.6
    info, genes = scGeneClust(raw, version='ps', n_obs_clusters=8, return_info=True, verbosity=0)
    n = info.n_vars
    assert genes.shape[0] > 0 and n <= int(0.2 * X.shape[1]) + 1
    R = info.varp['redundancy']
    assert R.shape == (n, n) and np.array_equal(R, R.T)                    # the reference's own test
    Z, labels = np.asarray(info.X), np.asarray(info.obs['cluster'])
    rs = np.random.RandomState(5)
    gs = rs.choice(n, 24, replace=False)
    assert np.array_equal(info.var['relevance'].to_numpy()[gs], stages.gene_relevance(Z[:, gs], labels, 0))
    pairs = np.sort(np.stack([rs.randint(0, n, 60), rs.randint(0, n, 60)], 1), 1)
    pairs = pairs[pairs[:, 0] != pairs[:, 1]]
    assert np.array_equal(R[pairs[:, 0], pairs[:, 1]], stages.gene_redundancy(info.varm['X_pca'], 0, pairs))
    edges = info.uns['mst_edges']
    pick = rs.choice(len(edges), 6, replace=False)
    assert np.array_equal(info.uns['mst_edges_complm'][pick], stages.gene_complementarity(Z, labels, edges[pick], 0))


@pytest.mark.gpu
@pytest.mark.parametrize("C,m,kind", [(70001, 7, "walk"), (70001, 3, "grow"), (20000, 5, "ties"), (5000, 64, "mixed"),
                                      (300, 2, "grow"), (400003, 7, "deviance"), (9000, 80, "walk")])
def test_sum_axis0_is_numpy_sequential_sum(torch_cuda, C, m, kind):
    """The piece-parallel column sum must return the bits of numpy's row-after-row float32 reduction, whatever the
    data does to the speculation (sign changes, binade crossings, exact ties, huge entries, inf)."""
    torch = torch_cuda
    from scgeneclust_b200._lib import lib, ptr, stream_handle
    rs = np.random.RandomState(C + m)
    if kind == "walk":                                   # wanders around zero: most pieces are replayed
        X = rs.standard_normal((C, m)).astype(np.float32)
    elif kind == "grow":
        X = np.abs(rs.standard_normal((C, m))).astype(np.float32) * 3
    elif kind == "ties":                                 # multiples of half an ulp of sums near 2^20, 2^12 and 1
        X = (rs.randint(-3, 4, (C, m)) * np.array([2.0 ** -4, 2.0 ** -12, 2.0 ** -24, 2.0 ** -4, 2.0 ** -13])).astype(np.float32)
        X[0] = [2 ** 20, -2 ** 12 - 1, 1.0, -2 ** 20 + 0.5, 2 ** 11]
    elif kind == "mixed":
        X = (rs.standard_normal((C, m)) * np.exp(rs.uniform(-12, 12, (C, m)))).astype(np.float32)
        X[rs.randint(C, size=3), rs.randint(m, size=3)] = [np.inf, 3e38, -3e38]
    else:                                                # deviance-like terms: mostly positive, many zeros
        v = rs.poisson(1.5, (C, m)).astype(np.float32)
        with np.errstate(all="ignore"):
            X = np.nan_to_num(v * np.log(v / 1.7), nan=0.0).astype(np.float32) - np.float32(0.05) * (v == 1)
    with np.errstate(all="ignore"):
        ref = X.sum(0)
    L, st = lib(), stream_handle()
    dX = torch.from_numpy(X).cuda()
    out = torch.empty(m, dtype=torch.float32, device="cuda")
    rep = torch.zeros(m, dtype=torch.int32, device="cuda")
    scratch = torch.empty(L.binomial_deviance_scratch_bytes(C, m) + 256, dtype=torch.uint8, device="cuda")
    L.sum_axis0_f32(ptr(dX), C, m, ptr(out), ptr(rep), ptr(scratch), st)
    got = out.cpu().numpy()
    print(f"sum_axis0 {kind} C={C} m={m}: replayed pieces {rep.cpu().numpy().tolist()[:8]} of {-(-C // 256)}")
    nan = np.isnan(ref)                                  # (x86 and CUDA encode an invalid-operation NaN differently)
    assert np.array_equal(np.isnan(got), nan)
    assert np.array_equal(got.view(np.uint32)[~nan], ref.view(np.uint32)[~nan]), (got, ref)
