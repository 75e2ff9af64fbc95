"""Brute-force fp64 restatements of the two kNN mutual-information estimators the reference calls through
scikit-learn, exposing the integer neighbour counts (test infrastructure; see oracle/__init__.py).

Reference call sites: scGeneClust/tl/information.py:24 (mutual_info_classif -> Ross estimator),
:63 and :96 (mutual_info_regression -> Kraskov-Stoegbauer-Grassberger estimator).
Arithmetic: sklearn 1.9.0 feature_selection/_mutual_info.py:20-80 (`_compute_mi_cc`), :83-153 (`_compute_mi_cd`),
:294-315 (`_estimate_mi` scaling + noise).
"""
from __future__ import annotations

import numpy as np
from scipy.special import digamma
from sklearn.preprocessing import scale

K_NEIGHBORS = 3


# ---------------------------------------------------------------------------------------------------------
# input preparation = what `_estimate_mi` does before the estimator sees the data (_mutual_info.py:294-315)
# ---------------------------------------------------------------------------------------------------------
def noise_vectors(n: int, random_state: int, continuous_target: bool = True):
    """The normals `_estimate_mi` draws from a fresh RandomState(seed): xi1 for the (single) feature column,
    then xi2 for a continuous target."""
    rng = np.random.RandomState(random_state)
    xi1 = rng.standard_normal(size=(n, 1))[:, 0]
    xi2 = rng.standard_normal(size=n) if continuous_target else None
    return xi1, xi2


def prep_feature(x: np.ndarray, xi1: np.ndarray) -> np.ndarray:
    """Feature column as the estimator sees it: float64 cast, /std (ddof 0, tiny std -> 1), + 1e-10*max(1,mean|x|)*xi1.
    Uses sklearn's own `scale` so the bits are the reference's."""
    X = np.asarray(x).reshape(-1, 1).astype(np.float64, copy=True)
    X[:, [0]] = scale(X[:, [0]], with_mean=False, copy=False)
    means = np.maximum(1, np.mean(np.abs(X), axis=0))
    X += 1e-10 * means * xi1.reshape(-1, 1)
    return X[:, 0]


def prep_target(y: np.ndarray, xi2: np.ndarray) -> np.ndarray:
    """Continuous target as the estimator sees it.  NOTE the dtype flow: `check_X_y` leaves a float32 `y`
    float32, `scale` keeps float32, and the in-place `y += noise` rounds back to float32 - so for the
    reference's float32 embeddings the target noise mostly vanishes.  Returned as float64 (the hstack in
    `_compute_mi_cc` promotes)."""
    y = np.array(y, copy=True)
    if y.dtype.kind != "f":
        y = y.astype(np.float64)
    y = scale(y, with_mean=False)
    y += 1e-10 * np.maximum(1, np.mean(np.abs(y))) * xi2
    return y.astype(np.float64)


# -- explicit restatement of the same preparation with numpy's pairwise summation spelled out --------------
def np_pairwise_sum(a: np.ndarray):
    """numpy's `pairwise_sum` for a contiguous 1-D float array (numpy/_core/src/umath/loops_utils.h.src):
    <8 elements sequential; <=128 elements eight interleaved accumulators combined as
    ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) then the tail sequentially; otherwise split at n/2 rounded down to a
    multiple of 8.  Bit-exact twin of `np.sum(a)`; the CUDA prep kernels mirror this function."""
    t = a.dtype.type
    n = a.shape[0]
    if n < 8:
        res = t(0.0)
        for i in range(n):
            res = t(res + a[i])
        return res
    if n <= 128:
        r = [t(a[j]) for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = t(r[j] + a[i + j])
            i += 8
        res = t(t(t(r[0] + r[1]) + t(r[2] + r[3])) + t(t(r[4] + r[5]) + t(r[6] + r[7])))
        while i < n:
            res = t(res + a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return t(np_pairwise_sum(a[:n2]) + np_pairwise_sum(a[n2:]))


def prep_feature_explicit(x: np.ndarray, xi1: np.ndarray) -> np.ndarray:
    """`prep_feature` with every floating-point operation spelled out (what the device kernel executes)."""
    x = np.asarray(x).astype(np.float64)
    n = x.shape[0]
    mean = np_pairwise_sum(x) / np.float64(n)
    dev = x - mean
    var = np_pairwise_sum(dev * dev) / np.float64(n)
    std = np.sqrt(var)
    if std < 10 * np.finfo(np.float64).eps:
        std = np.float64(1.0)
    x = x / std
    mabs = np_pairwise_sum(np.abs(x)) / np.float64(n)
    coef = np.float64(1e-10) * np.maximum(np.float64(1.0), mabs)
    return x + coef * xi1


def prep_target_explicit(y: np.ndarray, xi2: np.ndarray) -> np.ndarray:
    """`prep_target` for a float32 input with every operation spelled out."""
    y = np.asarray(y)
    assert y.dtype == np.float32
    n = y.shape[0]
    f = np.float32
    mean = f(np_pairwise_sum(y) / f(n))
    dev = (y - mean).astype(f)
    var = f(np_pairwise_sum((dev * dev).astype(f)) / f(n))
    std = np.sqrt(var)  # float32
    if std == 0.0:      # 1-D input: `_handle_zeros_in_scale` gets a scalar and only tests == 0
        std = f(1.0)
    y = (y / std).astype(f)
    mabs = f(np_pairwise_sum(np.abs(y)) / f(n))
    # numpy-2 promotion: python float 1e-10 is "weak", so 1e-10 * float32-scalar is a float32 product
    coef = np.float64(f(f(1e-10) * np.maximum(f(1.0), mabs)))
    return (y.astype(np.float64) + coef * xi2).astype(f).astype(np.float64)


# ---------------------------------------------------------------------------------------------------------
# KSG estimator (continuous-continuous), `_compute_mi_cc`
# ---------------------------------------------------------------------------------------------------------
def ksg_counts(x: np.ndarray, y: np.ndarray, k: int = K_NEIGHBORS):
    """x, y: prepared float64 vectors of equal length n.  Returns (eps, nx, ny) with
    eps_s = k-th smallest Chebyshev distance from s to the other points (= (k+1)-th smallest including the
    zero self-distance), nx_s = #{t != s : |x_s - x_t| <= nextafter(eps_s, 0)}, ny likewise."""
    dx = np.abs(x[:, None] - x[None, :])
    dy = np.abs(y[:, None] - y[None, :])
    d = np.maximum(dx, dy)
    eps = np.partition(d, k, axis=1)[:, k]
    r = np.nextafter(eps, 0)
    nx = (dx <= r[:, None]).sum(1) - 1
    ny = (dy <= r[:, None]).sum(1) - 1
    return eps, nx.astype(np.int64), ny.astype(np.int64)


def ksg_mi_from_counts(nx: np.ndarray, ny: np.ndarray, k: int = K_NEIGHBORS) -> float:
    n = nx.shape[0]
    mi = digamma(n) + digamma(k) - np.mean(digamma(nx + 1.0)) - np.mean(digamma(ny + 1.0))
    return max(0, mi)


def ksg_mi(x: np.ndarray, y: np.ndarray, k: int = K_NEIGHBORS) -> float:
    _, nx, ny = ksg_counts(x, y, k)
    return ksg_mi_from_counts(nx, ny, k)


def ksg_batch(XR: np.ndarray, YR: np.ndarray, pairs: np.ndarray, k: int = K_NEIGHBORS):
    """Vectorised KSG over many pairs.  XR/YR: (N, n) prepared x-role / y-role vectors; pairs: (P, 2) with
    pair p using XR[pairs[p,0]] and YR[pairs[p,1]].  Returns (nx, ny) int16 arrays of shape (P, n) and mi (P,)."""
    n = XR.shape[1]
    psi = digamma(np.arange(1, n + 1, dtype=np.float64))  # psi[m-1] = digamma(m)
    nx_all = np.empty((len(pairs), n), np.int16)
    ny_all = np.empty((len(pairs), n), np.int16)
    mi_all = np.empty(len(pairs), np.float64)
    step = max(1, (1 << 22) // (n * n))
    for s0 in range(0, len(pairs), step):
        pp = pairs[s0:s0 + step]
        x = XR[pp[:, 0]]
        y = YR[pp[:, 1]]
        dx = np.abs(x[:, :, None] - x[:, None, :])
        dy = np.abs(y[:, :, None] - y[:, None, :])
        d = np.maximum(dx, dy)
        eps = np.partition(d, k, axis=2)[:, :, k]
        r = np.nextafter(eps, 0)
        nx = (dx <= r[:, :, None]).sum(2) - 1
        ny = (dy <= r[:, :, None]).sum(2) - 1
        nx_all[s0:s0 + step] = nx
        ny_all[s0:s0 + step] = ny
        mi = digamma(n) + digamma(k) - psi[nx].mean(1) - psi[ny].mean(1)
        mi_all[s0:s0 + step] = np.maximum(0, mi)
    return nx_all, ny_all, mi_all


# ---------------------------------------------------------------------------------------------------------
# Ross estimator (continuous-discrete), `_compute_mi_cd`
# ---------------------------------------------------------------------------------------------------------
def ross_counts(c: np.ndarray, d: np.ndarray, k: int = K_NEIGHBORS):
    """c: prepared float64 vector, d: labels.  Returns (keep_mask, k_all, label_counts, m_all) restricted to
    the samples whose label is not unique; m_all includes the sample itself."""
    n = c.shape[0]
    radius = np.zeros(n)
    label_counts = np.zeros(n, np.int64)
    k_all = np.zeros(n, np.int64)
    for label in np.unique(d):
        mask = d == label
        count = int(mask.sum())
        if count > 1:
            kk = min(k, count - 1)
            cm = c[mask]
            dist = np.abs(cm[:, None] - cm[None, :])
            r = np.partition(dist, kk, axis=1)[:, kk]  # kk-th neighbour excluding self (self is index 0)
            radius[mask] = np.nextafter(r, 0)
            k_all[mask] = kk
        label_counts[mask] = count
    keep = label_counts > 1
    cc, rr = c[keep], radius[keep]
    m_all = (np.abs(cc[:, None] - cc[None, :]) <= rr[:, None]).sum(1)
    return keep, k_all[keep], label_counts[keep], m_all.astype(np.int64)


def ross_mi_from_counts(k_all, label_counts, m_all) -> float:
    n = m_all.shape[0]
    mi = (digamma(n) + np.mean(digamma(k_all.astype(np.float64)))
          - np.mean(digamma(label_counts.astype(np.float64))) - np.mean(digamma(m_all.astype(np.float64))))
    return max(0, mi)


def ross_mi(c: np.ndarray, d: np.ndarray, k: int = K_NEIGHBORS) -> float:
    _, k_all, label_counts, m_all = ross_counts(c, d, k)
    return ross_mi_from_counts(k_all, label_counts, m_all)
