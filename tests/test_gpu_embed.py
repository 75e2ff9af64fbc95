"""GPU parity of ingest, Pearson residuals (a1) and the gene-by-cell embedding (a2) against the CPU oracle
(oracle/stages.py: numpy restatement of scanpy's residuals + the installed sklearn PCA the reference calls).
Tolerances: counts / integer sums bit-exact; residuals 1e-5 relative (float32 rsqrt vs sqrt+div); PCA scores
1e-3 relative per component (BASELINE.json north_star)."""
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
def small_counts():
    from oracle.synth import synth_counts
    from scgeneclust_b200.synth import synth_params
    p = synth_params(2500, 0)
    X = synth_counts(p, 0, 3000, 0, 2500)
    keep = (X > 0).sum(0) >= 3                      # the reference loader's min_cells=3 filter (_utils.py:35)
    return p, np.ascontiguousarray(X[:, keep]), keep


def test_device_synth_bit_identical_to_oracle(torch_cuda, small_counts):
    from scgeneclust_b200.synth import DeviceSynth
    p, _, _ = small_counts
    from oracle.synth import synth_counts
    ds = DeviceSynth(p)
    for (c0, nc, g0, ng) in [(0, 200, 0, 2500), (1234, 77, 1000, 333), (10 ** 6 - 5, 5, 2400, 100)]:
        got = ds.counts(c0, nc, g0, ng).to_numpy()
        assert np.array_equal(got, synth_counts(p, c0, nc, g0, ng))


def test_ingest_paths_and_validation(torch_cuda, small_counts):
    from scgeneclust_b200._device import ingest_counts
    _, X, _ = small_counts
    Xs = X[:500]
    a = ingest_counts(Xs.astype(np.float32)).to_numpy()
    b = ingest_counts(sp.csr_matrix(Xs.astype(np.float32))).to_numpy()
    c = ingest_counts(Xs).to_numpy()
    assert np.array_equal(a, Xs) and np.array_equal(b, Xs) and np.array_equal(c, Xs)
    bad = Xs.astype(np.float32).copy()
    bad[3, 4] = 0.5
    with pytest.raises(ValueError, match="Expected raw counts"):     # scGeneClust/_validation.py:70
        ingest_counts(bad)
    with pytest.raises(ValueError, match="Expected raw counts"):
        ingest_counts(sp.csr_matrix(bad))
    empty_row = sp.csr_matrix(np.array([[0, 0, 0], [1, 0, 2]], np.float32))   # ragged: an empty CSR row
    assert np.array_equal(ingest_counts(empty_row).to_numpy(), empty_row.toarray().astype(np.uint16))


def test_count_sums_and_residuals(torch_cuda, small_counts):
    from oracle import stages
    from scgeneclust_b200._device import ResidualOperator, ingest_counts
    _, X, _ = small_counts
    op = ResidualOperator.from_counts(ingest_counts(X))
    assert np.array_equal(op.row_sum.cpu().numpy(), X.sum(1, dtype=np.int64))
    assert np.array_equal(op.col_sum.cpu().numpy(), X.sum(0, dtype=np.int64))
    Z_ref = stages.pearson_residuals(X.astype(np.float32))
    Z = op.materialize().cpu().numpy()
    assert np.allclose(Z, Z_ref, rtol=2e-5, atol=2e-6)
    assert np.abs(Z).max() <= np.float32(np.sqrt(X.shape[0]))
    zr, zc = op.residual_sums(rows=True, cols=True)
    assert np.allclose(zr.cpu().numpy(), Z_ref.sum(1, dtype=np.float64), rtol=1e-6, atol=1e-3)
    assert np.allclose(zc.cpu().numpy(), Z_ref.sum(0, dtype=np.float64), rtol=1e-6, atol=1e-3)
    cols = np.array([0, 17, 1999])
    assert np.allclose(op.materialize_columns(cols).cpu().numpy(), Z_ref[:, cols], rtol=2e-5, atol=2e-6)
    with pytest.raises(ValueError, match="all-zero"):
        Xz = X.copy()
        Xz[:, 5] = 0
        ResidualOperator.from_counts(ingest_counts(Xz))


@pytest.mark.parametrize("kernel", ["ts", "ss", "simt"])
def test_rsvd_passes_vs_numpy(torch_cuda, small_counts, kernel):
    """Y = M Q and Y = M^T Q over the implicit centred residual matrix against float64 numpy, for the product kernel
    (tcgen05, residual operand in tensor memory, TMA-fed count tiles) and the two cross-check kernels; ragged sizes
    (1111 x ~1500: partial tiles in both directions, zero-filled by TMA / clamped loads)."""
    torch = torch_cuda
    from oracle import stages
    from scgeneclust_b200._device import ResidualOperator, ingest_counts
    _, X, _ = small_counts
    X = X[:1111, :1501]
    X = X[:, X.sum(0) > 0]
    C, G = X.shape
    op = ResidualOperator.from_counts(ingest_counts(X))
    Z = stages.pearson_residuals(X.astype(np.float32)).astype(np.float64)
    rs = np.random.RandomState(0)
    row_shift = Z.mean(1)
    col_shift = Z.mean(0)
    for trans, n_in in ((0, G), (1, C)):
        Q = np.zeros((n_in, 64), np.float32)
        Q[:, :60] = rs.standard_normal((n_in, 60))
        for rsft, csft in ((row_shift, None), (None, col_shift), (None, None)):
            M = Z - (rsft[:, None] if rsft is not None else 0) - (csft[None, :] if csft is not None else 0)
            ref = (M @ Q) if trans == 0 else (M.T @ Q)
            t = lambda a: None if a is None else torch.from_numpy(a.astype(np.float32)).cuda()  # noqa: E731
            got = op.apply(trans, torch.from_numpy(Q).cuda(), t(rsft), t(csft), kernel=kernel).cpu().numpy()
            scale = np.abs(ref).max()
            assert np.abs(got[:, :60] - ref[:, :60]).max() <= 2e-4 * scale, (trans, kernel)
            assert not got[:, 60:].any()


def test_panel_algebra(torch_cuda):
    torch = torch_cuda
    from scgeneclust_b200._device import PanelAlgebra, Q_LD
    from scgeneclust_b200._lib import lib, ptr, stream_handle
    rs = np.random.RandomState(1)
    rows, q = 7001, 60
    P = np.zeros((rows, Q_LD), np.float32)
    P[:, :q] = rs.standard_normal((rows, q)) * np.linspace(1, 30, q)
    dP = torch.from_numpy(P).cuda()
    alg = PanelAlgebra(dP.device, rows)
    out = torch.empty_like(dP)
    alg.chol_qr(dP, q, out)
    Qh = out.cpu().numpy().astype(np.float64)
    assert np.abs(Qh[:, :q].T @ Qh[:, :q] - np.eye(q)).max() < 5e-5
    assert int(alg.status.item()) == 0 and not Qh[:, q:].any()
    # symmetric eigensolver vs LAPACK
    L, st = lib(), stream_handle()
    Gm = np.zeros((Q_LD, Q_LD))
    Gm[:q, :q] = (P[:, :q].astype(np.float64).T @ P[:, :q].astype(np.float64))
    dG = torch.from_numpy(Gm).cuda()
    ev = torch.empty(Q_LD, dtype=torch.float64, device="cuda")
    U32 = torch.empty((Q_LD, Q_LD), dtype=torch.float32, device="cuda")
    U64 = torch.empty((Q_LD, Q_LD), dtype=torch.float64, device="cuda")
    L.sym_eig(ptr(dG), q, ptr(ev), ptr(U32), ptr(U64), st)
    w = np.linalg.eigvalsh(Gm[:q, :q])[::-1]
    assert np.allclose(ev.cpu().numpy()[:q], w, rtol=1e-12)
    U = U64.cpu().numpy()[:q, :q]
    assert np.abs(U.T @ U - np.eye(q)).max() < 1e-12
    assert np.abs(U.T @ Gm[:q, :q] @ U - np.diag(w)).max() < 1e-9 * w[0]


def test_gene_pca_vs_sklearn(torch_cuda, small_counts):
    """varm['X_pca'] against the oracle (sklearn PCA(50, 'auto', random_state) on the float32 residuals)."""
    from oracle import stages
    from scgeneclust_b200._device import ResidualOperator, ingest_counts, randomized_pca
    _, X, _ = small_counts
    Z = stages.pearson_residuals(X.astype(np.float32))
    ref = stages.gene_pca(Z, 0)
    op = ResidualOperator.from_counts(ingest_counts(X))
    for simt in (True, False):
        info = {}
        got = randomized_pca(op, 'genes', 0, simt=simt, info=info).cpu().numpy()
        assert got.shape == ref.shape == (X.shape[1], 50) and got.dtype == np.float32
        rel = np.linalg.norm(got - ref, axis=0) / np.linalg.norm(ref, axis=0)
        print(f"gene PCA (simt={simt}): n_iter={info['n_iter']} transpose={info['transpose']} "
              f"max per-component rel err {rel.max():.2e}")
        assert rel.max() <= 1e-3
    # cell-level PCA (ps path, pp/_preprocessing.py:54)
    refc = stages.cell_pca(Z, 0)
    gotc = randomized_pca(op, 'cells', 0).cpu().numpy()
    relc = np.linalg.norm(gotc - refc, axis=0) / np.linalg.norm(refc, axis=0)
    print(f"cell PCA: max per-component rel err {relc.max():.2e}")
    assert relc.max() <= 1e-3


def test_gene_pca_untransposed_shape(torch_cuda, small_counts):
    """More genes than cells (config c1 shape class): sklearn does not transpose, passes run the other way round."""
    from oracle import stages
    from scgeneclust_b200._device import ResidualOperator, ingest_counts, randomized_pca
    _, X, _ = small_counts
    X = X[:900]
    X = X[:, (X > 0).sum(0) >= 3]
    Z = stages.pearson_residuals(X.astype(np.float32))
    ref = stages.gene_pca(Z, 0)
    info = {}
    got = randomized_pca(ResidualOperator.from_counts(ingest_counts(X)), 'genes', 0, info=info).cpu().numpy()
    assert info['transpose'] is False
    rel = np.linalg.norm(got - ref, axis=0) / np.linalg.norm(ref, axis=0)
    print(f"gene PCA untransposed: max per-component rel err {rel.max():.2e}")
    assert rel.max() <= 1e-3


@pytest.mark.parametrize("seed,rows,q", [(0, 20000, 60), (7, 1237, 60), (123456, 50, 13), (2 ** 32 - 1, 3000, 60),
                                         (5, 50000, 60)])
def test_device_omega_equals_numpy_randomstate(torch_cuda, seed, rows, q):
    """The randomized solver's test matrix (sklearn utils/extmath.py:323-334) generated on the device from the seeded
    MT19937 state must be the float32 cast of numpy's own legacy stream: identical entries, zero padding columns."""
    from scgeneclust_b200._device import device_normal_panel
    panel = device_normal_panel(seed, rows, q, "cuda")
    assert int(panel._gc_status.item()) == 0
    got = panel.cpu().numpy()
    ref = np.random.RandomState(seed).normal(size=(rows, q)).astype(np.float32)
    assert np.array_equal(got[:, :q], ref)
    assert not got[:, q:].any()


@pytest.mark.parametrize("cells,genes,expect", [(300, 4000, "covariance_eigh"), (260, 320, "full")])
def test_small_input_pca_solvers_vs_sklearn(torch_cuda, cells, genes, expect):
    """Small inputs make sklearn's 'auto' policy leave the randomized solver (few-hundred-cell data sets: gene PCA runs
    'covariance_eigh'; tiny matrices: 'full'); the device path follows it and matches the oracle's scores."""
    from oracle import stages
    from oracle.synth import synth_counts
    from scgeneclust_b200._device import ResidualOperator, ingest_counts, pca_solver, randomized_pca
    from scgeneclust_b200.synth import synth_params
    p = synth_params(genes, 2)
    X = synth_counts(p, 0, cells, 0, genes)
    X = np.ascontiguousarray(X[:, (X > 0).sum(0) >= 3])
    assert pca_solver(X.shape[1], X.shape[0], 50) == expect
    op = ResidualOperator.from_counts(ingest_counts(X))
    info = {}
    got = randomized_pca(op, 'genes', 0, info=info).cpu().numpy()
    assert info['solver'] == expect
    ref = stages.gene_pca(stages.pearson_residuals(X.astype(np.float32)), 0)
    rel = np.linalg.norm(got - ref, axis=0) / np.linalg.norm(ref, axis=0)
    print(f"{expect}: {X.shape}, max per-component rel err {rel.max():.2e}")
    assert got.shape == ref.shape and rel.max() <= 1e-3


def test_upper_clip_is_skipped_only_when_proved_inactive(torch_cuda, small_counts):
    """The pass kernels drop the per-entry `min(z, sqrt(C))` only when `ResidualOperator.from_counts` has PROVED from the
    per-gene count maxima that no residual can reach it.  Soundness: whenever the flag is set, the oracle's residuals
    stay below the clip.  A well-expressed matrix passes the proof and gives the same product as the always-clipping
    kernels; the same matrix with one outlier count in its rarest gene reaches the clip: the proof must fail and the
    product must match the oracle's clipped residuals."""
    torch = torch_cuda
    from oracle import stages
    from scgeneclust_b200._device import ResidualOperator, ingest_counts
    _, Xs, _ = small_counts
    Xs = np.ascontiguousarray(Xs[:900, :1200])
    Xs = Xs[:, Xs.sum(0) > 0].copy()
    rs = np.random.RandomState(3)
    Xd = rs.poisson(rs.uniform(2.0, 9.0, size=300), size=(900, 300)).astype(np.uint16)     # every gene well expressed
    for X in (Xs, Xd):
        op = ResidualOperator.from_counts(ingest_counts(X))
        Z = stages.pearson_residuals(X.astype(np.float32))
        if op.upper_clip_inactive:
            assert Z.max() < np.float32(np.sqrt(X.shape[0]))
    X = Xd
    op = ResidualOperator.from_counts(ingest_counts(X))
    assert op.upper_clip_inactive
    Q = np.zeros((X.shape[1], 64), np.float32)
    Q[:, :60] = rs.standard_normal((X.shape[1], 60))
    Qd = torch.from_numpy(Q).cuda()
    y_ts = op.apply(0, Qd).cpu().numpy()
    y_simt = op.apply(0, Qd, kernel='simt').cpu().numpy()
    assert np.abs(y_ts - y_simt).max() <= 2e-4 * np.abs(y_simt).max()
    # an outlier: 4000 counts of the rarest gene in one cell -> residual far above sqrt(900) = 30
    g = int(np.argmin(X.sum(0)))
    Xo = X.copy()
    Xo[17, g] = 4000
    Zo = stages.pearson_residuals(Xo.astype(np.float32))
    assert Zo[17, g] == np.float32(np.sqrt(Xo.shape[0]))               # the oracle clipped it
    opo = ResidualOperator.from_counts(ingest_counts(Xo))
    assert not opo.upper_clip_inactive
    for trans, panel in ((0, Q), (1, np.pad(rs.standard_normal((Xo.shape[0], 60)).astype(np.float32), ((0, 0), (0, 4))))):
        ref = (Zo.astype(np.float64) @ panel) if trans == 0 else (Zo.astype(np.float64).T @ panel)
        got = opo.apply(trans, torch.from_numpy(np.ascontiguousarray(panel)).cuda()).cpu().numpy()
        assert np.abs(got[:, :60] - ref[:, :60]).max() <= 2e-4 * np.abs(ref).max()
