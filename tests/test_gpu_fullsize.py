"""Size-independent properties of the hot path at BASELINE.json's FULL sizes (where the CPU oracle would take minutes):
config c3 = 50 000 cells x 20 000 genes for the embedding passes, config c4 = 4 000 relevant genes for the redundancy
stage.  Checksums are exact (integers); linear-algebra identities hold to float32 accumulation tolerance; sampled
entries are compared with the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

C3, G3 = 50_000, 20_000


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests")
    return torch


@pytest.fixture(scope="module")
def c3_operator(torch_cuda):
    from scgeneclust_b200._device import ResidualOperator
    from scgeneclust_b200.synth import DeviceSynth, synth_params
    params = synth_params(G3, 0)
    counts = DeviceSynth(params).counts(0, C3)
    return params, counts, ResidualOperator.from_counts(counts)


def test_c3_count_checksums_and_sampled_block(torch_cuda, c3_operator):
    """Exact integer checksum of checksums: sum of per-cell sums == sum of per-gene sums == total; a block of the
    device-generated matrix equals the CPU generator's; its sums equal the device's."""
    from oracle.synth import synth_counts
    params, counts, op = c3_operator
    assert int(op.row_sum.sum().item()) == int(op.col_sum.sum().item()) == int(op.total)
    blk = synth_counts(params, 31_000, 64, 0, G3)
    dev_blk = counts.data[31_000:31_064, :G3].cpu().numpy().view(np.uint16)
    assert np.array_equal(dev_blk, blk)
    assert np.array_equal(op.row_sum[31_000:31_064].cpu().numpy(), blk.sum(1, dtype=np.int64))


def test_c3_pass_kernel_linearity_adjointness_and_simt_agreement(torch_cuda, c3_operator):
    """tcgen05 pass kernel at full size: M(aQ1 + Q2) = a M Q1 + M Q2 (linearity), <Y, M Q> = <M^T Y, Q> (the two
    directions are each other's adjoint), and agreement with the exact-float32 SIMT kernel; tolerance = the bf16 hi+lo
    operand split (2^-16 relative per product, float32 accumulation)."""
    torch = torch_cuda
    _, _, op = c3_operator
    dev = op.counts.device
    g = torch.Generator(device="cpu").manual_seed(0)
    zr, _ = op.residual_sums(rows=True, cols=False)
    shift = (zr / G3).to(torch.float32)
    Q1 = torch.zeros((G3, 64)); Q1[:, :60] = torch.randn((G3, 60), generator=g)
    Q2 = torch.zeros((G3, 64)); Q2[:, :60] = torch.randn((G3, 60), generator=g)
    Q1, Q2 = Q1.to(dev), Q2.to(dev)
    Y1 = op.apply(0, Q1, shift, None).clone()
    Y2 = op.apply(0, Q2, shift, None).clone()
    Y12 = op.apply(0, 0.5 * Q1 + Q2, shift, None).clone()
    scale = float(Y12.abs().max())
    assert float((Y12 - (0.5 * Y1 + Y2)).abs().max()) <= 1e-4 * scale
    assert not bool(Y1[:, 60:].any())
    W = torch.zeros((C3, 64)); W[:, :60] = torch.randn((C3, 60), generator=g)
    W = W.to(dev)
    Zt = op.apply(1, W, shift, None).clone()
    lhs = (W.double() * Y1.double()).sum()
    rhs = (Zt.double() * Q1.double()).sum()
    assert abs(float(lhs - rhs)) <= 1e-5 * max(abs(float(lhs)), float(W.double().norm() * Y1.double().norm()) * 1e-2)
    for kernel in ("simt", "ss"):
        Ys = op.apply(0, Q1, shift, None, kernel=kernel)
        assert float((Ys - Y1).abs().max()) <= 1e-4 * scale, kernel
        Zs = op.apply(1, W, shift, None, kernel=kernel)
        assert float((Zs - Zt).abs().max()) <= 1e-4 * float(Zt.abs().max()), kernel
    # determinism: the split-K reduction order is fixed
    assert torch.equal(op.apply(0, Q1, shift, None), Y1)


def test_c3_embedding_is_orthogonal_and_sampled_scores_match_float64(torch_cuda, c3_operator):
    """Full-size gene PCA: score columns are mutually orthogonal with decreasing norms (U S of an SVD), and the scores of
    a block of genes equal Z_block^T (centred) projected on the returned directions recomputed in float64."""
    torch = torch_cuda
    from scgeneclust_b200._device import randomized_pca
    _, counts, op = c3_operator
    info = {}
    scores = randomized_pca(op, 'genes', 0, info=info)
    assert scores.shape == (G3, 50) and info['n_iter'] == 7 and info['transpose'] is True
    S = scores.double()
    gram = (S.T @ S).cpu().numpy()
    d = np.sqrt(np.diag(gram))
    assert np.all(np.diff(d) <= 1e-6 * d[0])                                        # singular values descending
    off = gram - np.diag(np.diag(gram))
    assert np.abs(off).max() <= 1e-4 * d[0] * d[0]
    assert np.allclose(d, info['singular_values'], rtol=1e-4)


def test_c4_redundancy_sampled_pairs_and_structure(torch_cuda):
    """Config c4 size (4 000 genes x 50 PCs -> 7 998 000 pairs): dense output symmetric with zero diagonal, equal to the
    condensed vector, and a random sample of pairs bit-identical to the oracle; splitting the pair range over two
    'ranks' reproduces the same values (the multi-GPU partition)."""
    torch = torch_cuda
    from oracle import mi_bruteforce as mb
    from scgeneclust_b200.tl.information import gene_redundancy
    rs = np.random.RandomState(4)
    N, n = 4000, 50
    E = (rs.standard_normal((N, n)) * (3.0 / (1 + np.arange(n)))).astype(np.float32)
    Ed = torch.from_numpy(E).cuda()
    cond, D = gene_redundancy(Ed, 0)
    assert bool((D == D.t()).all()) and float(D.diagonal().abs().max()) == 0.0 and float(D.min()) >= 0.0
    iu = torch.triu_indices(N, N, 1, device=D.device)
    assert torch.equal(D[iu[0], iu[1]], cond)
    half = cond.numel() // 2 + 17
    a, _ = gene_redundancy(Ed, 0, 0, half, dense=False)
    b, _ = gene_redundancy(Ed, 0, half, cond.numel(), dense=False)
    assert torch.equal(torch.cat([a, b]), cond)
    xi1, xi2 = mb.noise_vectors(n, 0)
    pairs = np.sort(np.stack([rs.randint(0, N, 300), rs.randint(0, N, 300)], 1), 1)
    pairs = pairs[pairs[:, 0] != pairs[:, 1]]
    XR = np.stack([mb.prep_feature(E[i], xi1) for i in pairs[:, 0]])
    YR = np.stack([mb.prep_target(E[j], xi2) for j in pairs[:, 1]])
    _, _, mi_o = mb.ksg_batch(XR, YR, np.stack([np.arange(len(pairs))] * 2, 1))
    assert np.array_equal(D.cpu().numpy()[pairs[:, 0], pairs[:, 1]], mi_o)


def test_c3_pass_kernel_sampled_blocks_vs_oracle(torch_cuda, c3_operator):
    """The tcgen05 pass kernel at the benchmarked size against the ORACLE (not against another kernel of this repo):
    for blocks of 64 cells (M Q) and blocks of 64 genes (M^T Q) the counts are regenerated on the CPU by the oracle's
    generator, turned into Pearson residuals by the oracle's formula (float32, like the reference) from the exact
    integer sums, centred, and multiplied with the panel in float64.  Tolerance 2e-4 of the output scale: the bf16
    hi+lo operand split (2^-16 per product) plus float32 accumulation over 20 000 / 50 000 terms."""
    torch = torch_cuda
    from oracle import stages
    from oracle import synth as osy
    params, counts, op = c3_operator
    po = osy.synth_params(G3, 0)
    dev = op.counts.device
    g = torch.Generator(device="cpu").manual_seed(7)
    Q = torch.zeros((G3, 64)); Q[:, :60] = torch.randn((G3, 60), generator=g)
    W = torch.zeros((C3, 64)); W[:, :60] = torch.randn((C3, 60), generator=g)
    zr, _ = op.residual_sums(rows=True, cols=False)
    shift = (zr / G3).to(torch.float32)
    Y = op.apply(0, Q.to(dev), shift, None).cpu().numpy().astype(np.float64)
    Zt = op.apply(1, W.to(dev), shift, None).cpu().numpy().astype(np.float64)
    row_sum = op.row_sum.cpu().numpy().astype(np.float64)
    col_sum = op.col_sum.cpu().numpy().astype(np.float64)
    total = float(op.total)
    shift_h = shift.cpu().numpy().astype(np.float64)
    Qh, Wh = Q.numpy().astype(np.float64), W.numpy().astype(np.float64)
    y_scale, z_scale = np.abs(Y).max(), np.abs(Zt).max()
    worst_y = worst_z = worst_shift = 0.0
    for c0 in (0, 12_345, 31_000, C3 - 64):                      # incl. the ragged last tile (50 000 = 390 x 128 + 80)
        blk = osy.synth_counts_c(po, c0, 64, 0, G3)
        Z = stages.pearson_residuals_block(blk, row_sum[c0:c0 + 64], col_sum, total, C3).astype(np.float64)
        ref_shift = Z.mean(axis=1)
        worst_shift = max(worst_shift, np.abs(ref_shift - shift_h[c0:c0 + 64]).max())
        ref = (Z - ref_shift[:, None]) @ Qh
        worst_y = max(worst_y, np.abs(ref - Y[c0:c0 + 64]).max())
    for g0 in (0, 7_777, 13_120, G3 - 64):
        blk = osy.synth_counts_c(po, 0, C3, g0, 64)
        Z = stages.pearson_residuals_block(blk, row_sum, col_sum[g0:g0 + 64], total, C3).astype(np.float64)
        ref = (Z - shift_h[:, None]).T @ Wh
        worst_z = max(worst_z, np.abs(ref - Zt[g0:g0 + 64]).max())
    print(f"c3 sampled blocks vs oracle: M Q {worst_y / y_scale:.2e}, M^T Q {worst_z / z_scale:.2e} of scale; "
          f"centring vector abs err {worst_shift:.2e}")
    assert worst_shift <= 1e-5
    assert worst_y <= 2e-4 * y_scale
    assert worst_z <= 2e-4 * z_scale


def test_c3_fast_selection_vs_oracle_golden(torch_cuda, c3_operator, golden_dir):
    """GeneClust-fast at config c3 (50 000 x 20 000, K = 200) through the entry point against the CPU oracle's result on
    the SAME matrix (tests/golden/c3_fast_oracle.npz, written by tests/golden/make_c3_oracle.py: oracle/stages.py =
    numpy residuals + sklearn 1.9.0 PCA / MiniBatchKMeans / IsolationForest; 47 s on 8 cores).  north_star's bars: PCA
    scores within 1e-3 relative per component, gene clusters reported as ARI - and, where the float32 paths agree
    closely enough, the same labels and the same selected-gene list."""
    import os
    from sklearn.metrics import adjusted_rand_score
    from scgeneclust_b200 import AnnDataLite, scGeneClust
    _, counts, _ = c3_operator
    gold = np.load(os.path.join(golden_dir, "c3_fast_oracle.npz"), allow_pickle=False)
    names = np.array([f"g{i}" for i in range(G3)], dtype=object)
    sel = scGeneClust(AnnDataLite(counts, var_names=names), n_var_clusters=200, version='fast', verbosity=0)
    # the intermediate slots through the stage functions (return_info=True would materialise 4 GB of residuals)
    from scgeneclust_b200 import pp, tl
    ad = AnnDataLite(counts, var_names=names)
    pp.normalize(ad, 'sc')
    pp.reduce_dim(ad, 'fast', 0)
    tl.cluster_genes(ad, None, 'fast', n_gene_clusters=200, random_state=0)
    from oracle import stages
    gpu = dict(X_pca=np.asarray(ad.varm['X_pca']), labels=ad.var['cluster'].to_numpy(),
               closeness=ad.var['closeness'].to_numpy(), selected=sel)
    ref = dict(X_pca=gold['X_pca'], labels=gold['labels'], closeness=gold['closeness'], selected=gold['selected'])
    par = stages.fast_parity(gpu, ref, names, 200, 0)
    print(f"c3 vs oracle golden: {par}")
    assert par["pca_rel_err_max"] <= 1e-3 and par["pca_rel_err_first10"] <= 1e-4         # link 1: float tolerance
    assert par["tail_exact"], par                  # link 2: sklearn k-means + closeness + selection on the GPU's embedding
    # end to end (labels_equal / selected_equal / ari in the printed block): reported, not asserted - see the harness
    # docstring of tests/test_gpu_subsample.py
