"""Host-side logic of the drop-in boundary, CPU only: argument validation with the reference's error types
(scGeneClust/_validation.py), fast/ps selection (tl/selection.py) against the golden vectors, MST helper."""
import os

import numpy as np
import pytest

from scgeneclust_b200 import AnnDataLite, scGeneClust
from scgeneclust_b200.tl.selection import compute_deviance, select_from_clusters
from scgeneclust_b200.tl.information import redundancy_mst


def _adata(n=6, g=5):
    return AnnDataLite(np.arange(n * g, dtype=np.float32).reshape(n, g))


@pytest.mark.parametrize("kwargs,exc", [
    (dict(verbosity=3), ValueError), (dict(version='slow'), ValueError), (dict(modality='xx'), ValueError),
    (dict(random_state=0.5), TypeError), (dict(max_workers=1.5), TypeError), (dict(max_workers=0), ValueError),
    (dict(max_workers=10 ** 6), ValueError), (dict(return_info=1), TypeError), (dict(subset=1), TypeError),
    (dict(post_hoc_filtering=1), TypeError), (dict(version='fast', n_var_clusters=None), TypeError),
    (dict(version='fast', n_var_clusters=1), ValueError),
    (dict(version='fast', n_var_clusters=5, modality='st'), ValueError),
    (dict(version='ps', n_var_clusters=5), TypeError), (dict(version='ps', n_obs_clusters=None), TypeError),
    (dict(version='ps', n_obs_clusters=1), ValueError),
    (dict(version='ps', n_obs_clusters=3, n_components=1), ValueError),
    (dict(version='ps', n_obs_clusters=3, relevant_gene_pct=100), ValueError),
    (dict(version='ps', n_obs_clusters=3, relevant_gene_pct=2.5), TypeError),
    (dict(version='ps', n_obs_clusters=3, shape='circle'), ValueError),
    (dict(version='ps', n_obs_clusters=3, modality='st'), RuntimeWarning),
    (dict(version='ps', n_obs_clusters=3, modality='st', image=[1]), TypeError),
])
def test_argument_errors_match_reference(kwargs, exc):
    kw = dict(version='fast', n_var_clusters=3)
    if kwargs.get('version') == 'ps':
        kw = {}
    kw.update(kwargs)
    with pytest.raises(exc):
        scGeneClust(_adata(), **kw)


def test_non_anndata_and_raw_rejected():
    with pytest.raises(TypeError):
        scGeneClust(np.zeros((3, 3)), n_var_clusters=2)
    ad = _adata()
    ad.raw = object()
    with pytest.raises(ValueError):
        scGeneClust(ad, n_var_clusters=2)


def test_fast_selection_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "fast_cluster_select_small.npz"), allow_pickle=False)
    ad = AnnDataLite(g["Zres"].copy(), var_names=g["var_names"].astype(object))
    ad.var['cluster'] = g["labels"]
    ad.var['closeness'] = g["closeness"]
    sel = select_from_clusters(ad, 'fast', True, int(g["seed"]))
    assert np.array_equal(sel.astype(str), g["selected_posthoc"])
    sel = select_from_clusters(ad, 'fast', False, int(g["seed"]))
    assert np.array_equal(sel.astype(str), g["selected_plain"])
    assert np.array_equal(compute_deviance(g["Zres"][:, :8]), g["deviance_first8"], equal_nan=True)


def test_fast_selection_tie_rule_first_occurrence():
    ad = AnnDataLite(np.zeros((2, 6), np.float32), var_names=np.array(list("abcdef"), dtype=object))
    ad.var['cluster'] = np.array([1, 0, 1, 0, 1, 2], np.int32)
    ad.var['closeness'] = np.array([1.0, 0.5, 1.0, 0.5, 0.0, 1.0], np.float32)
    sel = select_from_clusters(ad, 'fast', False, 0)
    # cluster 0: max first 'b', min first 'b'; cluster 1: max first 'a', min 'e'; cluster 2 singleton dropped
    assert list(sel) == ['b', 'a', 'b', 'e']


def test_ps_selection_rules():
    names = np.array([f"g{i}" for i in range(8)], dtype=object)
    ad = AnnDataLite(np.zeros((2, 8), np.float32), var_names=names)
    ad.var['cluster'] = np.array([0, 0, 0, 1, 1, 1, -1, -1])
    ad.var['relevance'] = np.array([.1, .9, .9, .3, .2, .1, .5, .6])
    ad.var['outlier_score'] = np.array([0., .2, .4, 0., .6, .0, .39, .41])
    sel = select_from_clusters(ad, 'ps', True, 0)   # dense positive scores: .2 .4 .6 -> median .4
    assert list(sel) == ['g1', 'g3', 'g7']
    sel = select_from_clusters(ad, 'ps', False, 0)
    assert list(sel) == ['g1', 'g3']


def test_redundancy_mst_is_max_spanning_forest():
    rs = np.random.RandomState(0)
    N = 30
    R = np.triu(rs.rand(N, N), 1)
    R[R < 0.4] = 0.0              # the reference's graph has no edge where MI == 0
    R[:, 7] = 0.0; R[7, :] = 0.0  # isolated vertex -> spanning forest
    R = R + R.T
    edges = redundancy_mst(R)
    assert edges.shape == (N - 2, 2) and (edges[:, 0] < edges[:, 1]).all()
    assert 7 not in edges
    # total weight equals Kruskal on sorted edges (independent check)
    iu = np.triu_indices(N, 1)
    order = np.argsort(-R[iu], kind="stable")
    parent = list(range(N))
    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a
    tot = 0.0
    for e in order:
        i, j, w = iu[0][e], iu[1][e], R[iu][e]
        if w == 0:
            break
        ri, rj = find(i), find(j)
        if ri != rj:
            parent[ri] = rj
            tot += w
    assert np.isclose(R[edges[:, 0], edges[:, 1]].sum(), tot, rtol=0, atol=1e-12)
    assert list(map(tuple, edges)) == sorted(map(tuple, edges))


def test_minibatch_sampling_restatement_equals_randomstate_choice():
    """DeviceMiniBatchKMeans draws its mini-batches as `cdf.searchsorted(rs.random_sample(batch), 'right')` with a cdf
    built once; that must be the index stream of `rs.choice(n, batch, p=p)` (sklearn cluster/_kmeans.py:2169-2174),
    and one `uniform(size=(K-1, t))` call must equal the K-1 per-round draws of k-means++ (cluster/_kmeans.py:249)."""
    import numpy as np
    for seed, n, batch in [(0, 20000, 2000), (3, 1371, 1024), (11, 50, 50)]:
        p = np.ones(n, dtype=np.float32)
        p = p / np.sum(p)
        cdf = np.asarray(p, dtype=np.float64).cumsum()
        cdf /= cdf[-1]
        a, b = np.random.RandomState(seed), np.random.RandomState(seed)
        for _ in range(5):
            assert np.array_equal(a.choice(n, batch, p=p, replace=True), cdf.searchsorted(b.random_sample(batch), side='right'))
        assert a.randint(1 << 30) == b.randint(1 << 30)          # both streams consumed the same amount
    a, b = np.random.RandomState(5), np.random.RandomState(5)
    assert np.array_equal(np.stack([a.uniform(size=7) for _ in range(9)]), b.uniform(size=(9, 7)))


def test_isolation_forest_restatement_equals_sklearn():
    """The singleton-rescue outlier test (scGeneClust/tl/selection.py:51) runs a line-by-line restatement of sklearn's
    1-D IsolationForest; its score_samples must be bit-identical to sklearn's own, incl. ties / near-constant data."""
    import numpy as np
    from sklearn.ensemble import IsolationForest
    from scgeneclust_b200.tl.selection import (_first_randint31, _isolation_forest_1d, _isolation_forest_1d_native,
                                                _isolation_forest_outliers)
    sd = np.array([0, 1, 5, 123456, 2 ** 31 - 2, 99999999])
    assert np.array_equal(_first_randint31(sd), [np.random.RandomState(int(s)).randint(2 ** 31 - 1) for s in sd])
    rs = np.random.RandomState(2)
    for trial in range(25):
        n, seed = int(rs.randint(3, 64)), int(rs.randint(0, 1000))
        kind = trial % 5
        if kind == 0:
            v = rs.normal(size=n) * 1e4
        elif kind == 1:
            v = np.round(rs.normal(size=n), 1) * 100                         # ties
        elif kind == 2:
            v = rs.lognormal(size=n) * 1e5
        elif kind == 3:
            v = np.concatenate([rs.normal(size=n - 1), [50.0]])              # one obvious outlier
        else:
            v = np.concatenate([np.full(n - 2, 3.25), rs.normal(size=2) * 1e-8 + 3.25])   # constant in float32
        forest = IsolationForest(random_state=seed).fit(v.reshape(-1, 1))
        ref_scores = forest.score_samples(v.reshape(-1, 1))
        assert np.array_equal(_isolation_forest_1d(v, seed, return_scores=True), ref_scores)            # Python restatement
        assert np.array_equal(_isolation_forest_1d_native(v, seed, return_scores=True), ref_scores)     # C++ host routine
        assert np.array_equal(_isolation_forest_outliers(v, seed), forest.predict(v.reshape(-1, 1)) == -1)
    for n in (150, 256):                                                     # up to sklearn's max_samples: still every sample per tree
        v = rs.standard_normal(n) * 50
        forest = IsolationForest(random_state=7).fit(v.reshape(-1, 1))
        assert np.array_equal(_isolation_forest_1d_native(v, 7, return_scores=True), forest.score_samples(v.reshape(-1, 1)))
    for n in (1, 2):                                                         # closed-form small cases
        v = rs.normal(size=n)
        assert np.array_equal(_isolation_forest_outliers(v, 0),
                              IsolationForest(random_state=0).fit_predict(v.reshape(-1, 1)) == -1)


def test_hdbscan_tail_labels_equal_sklearn_fork():
    """MST -> single linkage -> condensed tree (min cluster size 3) -> EOM labels: the restatement of hdbscan's
    functions (scGeneClust/tl/cluster.py:127-131) must agree with scikit-learn's fork of the same code
    (`_hdbscan._linkage.make_single_linkage`, `_hdbscan._tree.tree_to_labels`); GLOSH scores are in [0, 1]."""
    import numpy as np
    from scipy.sparse.csgraph import minimum_spanning_tree
    from sklearn.cluster._hdbscan import _linkage, _tree
    from scgeneclust_b200.tl._hdbscan_tail import mst_to_clusters, single_linkage_label
    rs = np.random.RandomState(0)
    for trial in range(4):
        n = [40, 120, 300, 77][trial]
        centers = rs.normal(size=(5, 3)) * 6
        pts = centers[rs.randint(0, 5, n)] + rs.normal(size=(n, 3))
        dmat = np.linalg.norm(pts[:, None] - pts[None], axis=2)
        mst = minimum_spanning_tree(dmat).tocoo()
        edges = np.stack([mst.row, mst.col, mst.data], axis=1)
        edges = edges[np.argsort(edges[:, 2])]
        labels, scores = mst_to_clusters(edges, n, min_cluster_size=3)
        labels_py, scores_py = mst_to_clusters(edges, n, min_cluster_size=3, native=False)    # Python restatement
        assert np.array_equal(labels, labels_py) and np.array_equal(scores, scores_py)
        from scgeneclust_b200.tl._hdbscan_tail import condense_tree, native_tree_and_labels
        tree_c, _ = native_tree_and_labels(edges, 3)
        tree_py = condense_tree(single_linkage_label(edges), 3)
        assert all(np.array_equal(tree_c[f], tree_py[f]) for f in ('parent', 'child', 'lambda_val', 'child_size'))
        rec = np.zeros(n - 1, dtype=_linkage.MST_edge_dtype)
        rec['current_node'], rec['next_node'], rec['distance'] = edges[:, 0].astype(np.intp), edges[:, 1].astype(np.intp), edges[:, 2]
        slt = _linkage.make_single_linkage(rec)
        ref_labels, _ = _tree.tree_to_labels(slt, 3, "eom", False, 0.0, None)
        mine = single_linkage_label(edges)
        assert np.array_equal(mine[:, 2], slt['value']) and np.array_equal(mine[:, 3], slt['cluster_size'])
        assert np.array_equal(labels, ref_labels)
        assert scores.shape == (n,) and (scores >= 0).all() and (scores <= 1).all()
    with __import__('pytest').raises(ValueError):
        mst_to_clusters(edges[:-1], n)                                        # spanning forest: the reference fails too


def test_device_matrix_stands_in_for_adata_x():
    """`DeviceMatrix` (adata.X / varp['redundancy'] kept on the device until read) behaves like the ndarray it replaces
    for everything the stage functions do with it: shape / dtype, column and row subsetting through AnnDataLite's
    in-place subset calls, np.asarray, copy.  Runs on a CPU tensor: the class is device-agnostic."""
    import numpy as np
    import torch
    from scgeneclust_b200._anndata_lite import AnnDataLite, DeviceMatrix
    rs = np.random.RandomState(0)
    A = rs.standard_normal((7, 5)).astype(np.float32)
    dm = DeviceMatrix(torch.from_numpy(A.copy()))
    assert dm.shape == (7, 5) and dm.dtype == np.float32 and dm.ndim == 2
    assert np.array_equal(np.asarray(dm), A)
    ad = AnnDataLite(dm, var_names=[f"g{i}" for i in range(5)], obs_names=[f"c{i}" for i in range(7)])
    keep_cells = np.array([True, False, True, True, False, True, True])
    ad._inplace_subset_obs(keep_cells)
    assert isinstance(ad.X, DeviceMatrix) and ad.X.shape == (5, 5) and list(ad.obs_names) == ["c0", "c2", "c3", "c5", "c6"]
    ad._inplace_subset_var(np.array(["g3", "g1"], dtype=object))
    assert isinstance(ad.X, DeviceMatrix) and np.array_equal(np.asarray(ad.X), A[keep_cells][:, [3, 1]])
    cp = ad.copy()
    cp.X.tensor.zero_()
    assert np.asarray(ad.X).any()                                      # copy is deep
    assert np.array_equal(dm[[0, 2], :].tensor.numpy(), A[[0, 2]])     # integer row lists
    assert np.array_equal(np.asarray(dm)[1:3, 2], dm[1:3, 2])          # anything else falls back to a host array
    d64 = DeviceMatrix(torch.zeros((3, 3), dtype=torch.float64))
    assert d64.dtype == np.float64


def test_anndata_lite_lazy_default_names():
    import numpy as np
    from scgeneclust_b200._anndata_lite import AnnDataLite
    ad = AnnDataLite(np.zeros((3, 4), np.float32))
    assert list(ad.var_names) == ["gene_0", "gene_1", "gene_2", "gene_3"] and list(ad.obs_names) == ["cell_0", "cell_1", "cell_2"]
    ad._inplace_subset_var(np.array([True, False, True, False]))
    assert list(ad.var_names) == ["gene_0", "gene_2"] and ad.shape == (3, 2)


def test_glosh_scores_and_noise_labels_on_separated_blobs():
    """GLOSH outlier scores from the condensed tree (scGeneClust/tl/cluster.py:131): two tight blobs become two
    clusters, the point that joins a blob at a much larger distance than the blob's own scale gets the top outlier
    score, points leaving the root are noise (-1); labels equal scikit-learn's fork on the same tree."""
    import numpy as np
    from scipy.sparse.csgraph import minimum_spanning_tree
    from sklearn.cluster._hdbscan import _linkage, _tree
    from scgeneclust_b200.tl._hdbscan_tail import mst_to_clusters
    rs = np.random.RandomState(3)
    blobs = np.concatenate([rs.normal(0, 0.05, size=(40, 2)), rs.normal(5, 0.05, size=(40, 2))])
    strays = np.array([[2.5, 2.5], [2.4, -3.0], [8.0, 0.0]])
    pts = np.concatenate([blobs, strays])
    dmat = np.linalg.norm(pts[:, None] - pts[None], axis=2)
    mst = minimum_spanning_tree(dmat).tocoo()
    edges = np.stack([mst.row, mst.col, mst.data], axis=1)
    edges = edges[np.argsort(edges[:, 2])]
    labels, scores = mst_to_clusters(edges, len(pts), min_cluster_size=3)
    assert set(labels[:40]) == {labels[0]} and set(labels[40:80]) == {labels[40]} and labels[0] != labels[40]
    assert (labels[81:] == -1).all()
    assert int(np.argmax(scores)) == 80 and scores[80] > 0.9          # joins blob 1 at ~3.5 against a blob scale of ~0.05
    assert np.median(scores[:80]) < 0.1
    assert scores.min() >= 0 and scores.max() <= 1
    rec = np.zeros(len(pts) - 1, dtype=_linkage.MST_edge_dtype)
    rec['current_node'], rec['next_node'], rec['distance'] = edges[:, 0].astype(np.intp), edges[:, 1].astype(np.intp), edges[:, 2]
    assert np.array_equal(labels, _tree.tree_to_labels(_linkage.make_single_linkage(rec), 3, "eom", False, 0.0, None)[0])


def test_speculative_cumsum_prototype_is_bit_exact():
    """The algorithm of the k-means++ round kernel's running sum (scripts/proto/spec_cumsum.py): piece-parallel
    evaluation of a sequential float32 cumsum from surrogate starts, verified piece by piece - must reproduce the
    dependent chain bit for bit, including ties (few mantissa bits), zeros, outliers and 20 binades of range."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "proto", "spec_cumsum.py")
    spec = importlib.util.spec_from_file_location("spec_cumsum", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rs = np.random.RandomState(11)
    reals = []
    for trial in range(24):
        cn = int(rs.choice([6000, 6144, 777, 65, 3]))
        v = [rs.gamma(2.0, 30.0, cn),
             rs.gamma(2.0, 30.0, cn) * (rs.rand(cn) < 0.5),
             np.ldexp(rs.randint(1, 64, cn).astype(np.float64), rs.randint(-6, 3, cn)),
             rs.lognormal(0, 4, cn)][trial % 4].astype(np.float32)
        carry = 0.0 if trial % 3 else float(np.float32(rs.gamma(2.0, 1e5)))
        reals.append(mod.check(v, carry))          # asserts bit equality at every piece boundary
    assert np.mean(reals) < 20                     # and most pieces are resolved without a real chain


def test_mt19937_jump_polynomial_header_is_current():
    """csrc/mt_jump_poly.h (the GF(2) jump polynomial that lets the Omega stream run in parallel segments) must be what
    scripts/mt_jump_poly.py derives - Berlekamp-Massey on the generator's own output, square-and-multiply, and a check of
    the jump identity against a directly generated sequence."""
    import importlib.util
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("mt_jump_poly", os.path.join(root, "scripts", "mt_jump_poly.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    header = open(os.path.join(root, "scgeneclust_b200", "csrc", "mt_jump_poly.h")).read()
    words = [int(w, 16) for w in re.findall(r"0x([0-9a-f]{8})u", header)]
    assert len(words) == 624 and f"kMtSegBlocks = {mod.SEG_BLOCKS}" in header
    g = sum(w << (32 * i) for i, w in enumerate(words))
    taps = np.array([i for i in range(mod.DEG) if (g >> i) & 1], dtype=np.int64)
    seq = mod.sequence(99, mod.SEG_BLOCKS + 40)
    for k in (0, 1000):
        for w in (0, 17, 623):
            assert np.bitwise_xor.reduce(seq[k + taps + w]) == seq[k + mod.J + w]
    # and the generator the script uses is numpy's: tempered words of the first block
    rs = np.random.RandomState(99)
    y = seq[:4].astype(np.uint64)
    y ^= y >> np.uint64(11)
    y ^= (y << np.uint64(7)) & np.uint64(0x9d2c5680)
    y ^= (y << np.uint64(15)) & np.uint64(0xefc60000)
    y ^= y >> np.uint64(18)
    assert np.array_equal(y.astype(np.uint32), rs.randint(0, 2 ** 32, 4, dtype=np.uint64).astype(np.uint32))


def test_piece_parallel_float32_sum_prototype_is_exact():
    """scripts/proto/spec_colsum.py: the integer-delta summary of a piece under a guessed binade, verified piece by piece,
    returns the bits of the sequential float32 sum on ties, sign changes and binade crossings (the device version is
    csrc/deviance_kernels.cu colsum_spec_*; its GPU test is tests/test_gpu_e2e.py::test_sum_axis0_is_numpy_sequential_sum)."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("spec_colsum", os.path.join(root, "scripts", "proto", "spec_colsum.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rs = np.random.RandomState(3)
    cases = [rs.standard_normal(700).astype(np.float32),
             (np.abs(rs.standard_normal(900)) * 3).astype(np.float32),
             np.concatenate([[np.float32(2 ** 20)], (rs.randint(-3, 4, 800) * 2.0 ** -4).astype(np.float32)]),
             np.concatenate([[np.float32(1.0)], (rs.randint(-1, 2, 600) * 2.0 ** -24).astype(np.float32)])]
    for x in cases:
        assert mod.seq_sum(x).view(np.uint32) == mod.spec_sum(x).view(np.uint32)
