"""NumPy/LAPACK restatement of the cGP per-cluster GP hot path (TEST INFRASTRUCTURE ONLY).

Citations: ``REF/`` = /root/reference (gptune/cGP), ``SK/`` = scikit-learn 1.9.0,
``SP/`` = SciPy 1.18.1 (the third-party packages the reference delegates to; see
``oracle/__init__.py`` for the pin statement).

theta layout (SK/gaussian_process/kernels.py:752,763-765): ``[log l_1 .. log l_d, log s2]``
for ``Matern(nu=1.5, ARD) + WhiteKernel`` as built at REF/cGP/cGP_constrained.py:169.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg
import scipy.optimize
from scipy.special import ndtr

LOG_BOUND_LO = math.log(1e-5)  # REF/cGP/cGP_constrained.py:169 length_scale_bounds / noise_level_bounds
LOG_BOUND_HI = math.log(1e5)
SQRT3 = math.sqrt(3.0)


# --------------------------------------------------------------------------------------
# kernels  (SK/gaussian_process/kernels.py:1713-1768 Matern, :1558-1585 RBF, :1406-1413 White)
# --------------------------------------------------------------------------------------
def _scaled_sqdist(A, B, length_scale):
    """sum_p (A_ip/l_p - B_jp/l_p)^2, divide-then-subtract as ``pdist(X / length_scale)``
    does (SK kernels.py:1716,1720), accumulated sequentially over p like scipy's C loop."""
    As = A / length_scale
    Bs = B / length_scale
    out = np.zeros((A.shape[0], B.shape[0]))
    for p in range(A.shape[1]):
        t = As[:, p, None] - Bs[None, :, p]
        out += t * t
    return out


def _stationary(sq, kind):
    if kind == "matern32":
        k = np.sqrt(sq) * SQRT3  # SK kernels.py:1725
        return (1.0 + k) * np.exp(-k)  # :1726
    if kind == "rbf":
        return np.exp(-0.5 * sq)  # SK kernels.py:1561-1562
    raise ValueError(kind)


def kernel_matrix(X, theta, kind="matern32", alpha=0.0):
    """K = k_stationary(X, X) + exp(theta[-1]) I + alpha I.

    SK kernels.py:1743 forces the stationary diagonal to 1; Sum adds the White term (:871,
    :1407); GPR adds alpha (SK/gaussian_process/_gpr.py:589,350)."""
    X = np.asarray(X, dtype=np.float64)
    d = X.shape[1]
    ls = np.exp(theta[:d])
    K = _stationary(_scaled_sqdist(X, X, ls), kind)
    np.fill_diagonal(K, 1.0)
    K[np.diag_indices_from(K)] += math.exp(theta[d])
    K[np.diag_indices_from(K)] += alpha
    return K


def kernel_cross(Xq, Xtrain, theta, kind="matern32"):
    """k(Xq, Xtrain): stationary part only, White contributes 0 (SK kernels.py:1418-1419)."""
    d = Xtrain.shape[1]
    ls = np.exp(theta[:d])
    return _stationary(_scaled_sqdist(np.asarray(Xq, float), np.asarray(Xtrain, float), ls), kind)


def kernel_matrix_grad(X, theta, kind="matern32"):
    """dK/dtheta [n, n, P] (SK kernels.py:1752-1768 Matern nu=1.5, :1579-1585 RBF, :1408-1413 White)."""
    X = np.asarray(X, dtype=np.float64)
    n, d = X.shape
    ls = np.exp(theta[:d])
    D = (X[:, None, :] - X[None, :, :]) ** 2 / (ls ** 2)  # :1753
    out = np.empty((n, n, d + 1))
    if kind == "matern32":
        out[:, :, :d] = 3.0 * D * np.exp(-np.sqrt(3.0 * D.sum(-1)))[..., None]  # :1768
    else:
        K = _stationary(_scaled_sqdist(X, X, ls), kind)
        np.fill_diagonal(K, 1.0)
        out[:, :, :d] = D * K[..., None]  # :1581-1584
    out[:, :, d] = math.exp(theta[d]) * np.eye(n)  # :1412
    return out


# --------------------------------------------------------------------------------------
# GPR  (SK/gaussian_process/_gpr.py)
# --------------------------------------------------------------------------------------
def normalize_y(y):
    """SK _gpr.py:275-280 with _handle_zeros_in_scale (SK/preprocessing/_data.py:127-132)."""
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    mean = float(np.mean(y))
    std = float(np.std(y))
    if std < 10 * np.finfo(np.float64).eps:
        std = 1.0
    return (y - mean) / std, mean, std


def lml_and_grad(X, y_norm, theta, alpha=1e-5, kind="matern32", want_grad=True):
    """log-marginal likelihood and gradient, SK _gpr.py:584-651.

    Memory-lean: the [n,n,P] gradient tensor is contracted one parameter at a time."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y_norm, dtype=np.float64).reshape(-1)
    theta = np.asarray(theta, dtype=np.float64)
    n, d = X.shape
    K = kernel_matrix(X, theta, kind, alpha)
    try:
        L = scipy.linalg.cholesky(K, lower=True, check_finite=False)  # :591
    except np.linalg.LinAlgError:
        return (-np.inf, np.zeros_like(theta)) if want_grad else -np.inf  # :592-593
    a = scipy.linalg.cho_solve((L, True), y, check_finite=False)  # :601
    lml = -0.5 * float(y @ a) - float(np.log(np.diag(L)).sum()) - n / 2 * math.log(2 * math.pi)  # :613-617
    if not want_grad:
        return lml
    Kinv = scipy.linalg.cho_solve((L, True), np.eye(n), check_finite=False)  # :631-633
    inner = np.outer(a, a) - Kinv  # :629,636
    ls = np.exp(theta[:d])
    grad = np.empty(d + 1)
    if kind == "matern32":
        Dsum = np.zeros((n, n))
        Dp = []
        for p in range(d):
            t = (X[:, p, None] - X[None, :, p]) ** 2 / (ls[p] ** 2)
            Dp.append(t)
            Dsum += t
        E = np.exp(-np.sqrt(3.0 * Dsum))
        for p in range(d):
            grad[p] = 0.5 * float(np.sum(inner * (3.0 * Dp[p] * E)))  # :648-650 with kernels.py:1768
    else:
        Ks = _stationary(_scaled_sqdist(X, X, ls), kind)
        np.fill_diagonal(Ks, 1.0)
        for p in range(d):
            t = (X[:, p, None] - X[None, :, p]) ** 2 / (ls[p] ** 2)
            grad[p] = 0.5 * float(np.sum(inner * (t * Ks)))
    grad[d] = 0.5 * math.exp(theta[d]) * float(np.trace(inner))
    return lml, grad


def restart_thetas(P, R, seed=0):
    """theta0 of runs 0..R: run 0 starts at theta=0 (kernel template l=1, s2=1, cloned each fit:
    SK _gpr.py:249,312-318); runs 1..R draw ``RandomState(seed).uniform(lo, hi)`` sequentially
    (SK _gpr.py:329-333).  Identical for every cluster because random_state=0 is re-seeded per fit
    (REF/cGP/cGP_constrained.py:315)."""
    rng = np.random.RandomState(seed)
    out = np.zeros((R + 1, P))
    lo = np.full(P, LOG_BOUND_LO)
    hi = np.full(P, LOG_BOUND_HI)
    for r in range(R):
        out[r + 1] = rng.uniform(lo, hi)
    return out


def fit_gpr(X, y_raw, n_restarts, alpha=1e-5, kind="matern32", seed=0):
    """GaussianProcessRegressor.fit as configured at REF/cGP/cGP_constrained.py:315-327
    (SK _gpr.py:233-368): normalise y, L-BFGS-B from R+1 starts, first argmin of -lml,
    final factorisation."""
    X = np.asarray(X, dtype=np.float64)
    y, y_mean, y_std = normalize_y(y_raw)
    P = X.shape[1] + 1
    starts = restart_thetas(P, n_restarts, seed)
    bounds = [(LOG_BOUND_LO, LOG_BOUND_HI)] * P
    runs = []
    for t0 in starts:
        def obj(th):
            lml, g = lml_and_grad(X, y, th, alpha, kind)
            return -lml, -g

        res = scipy.optimize.minimize(obj, t0, method="L-BFGS-B", jac=True, bounds=bounds)  # SK _gpr.py:660-666
        runs.append((res.x, res.fun, res.nfev))
    vals = [r[1] for r in runs]
    best = int(np.argmin(vals))  # SK _gpr.py:336-337 (first minimum)
    theta = runs[best][0]
    K = kernel_matrix(X, theta, kind, alpha)
    L = scipy.linalg.cholesky(K, lower=True, check_finite=False)  # SK _gpr.py:352
    a = scipy.linalg.cho_solve((L, True), y, check_finite=False)  # :363
    return dict(theta=theta, lml=-vals[best], L=L, alpha_vec=a, y_mean=y_mean, y_std=y_std,
                y_norm=y, X=X, kind=kind, run_thetas=np.array([r[0] for r in runs]),
                run_values=np.array(vals), run_nfev=np.array([r[2] for r in runs]), best_run=best)


def predict_mean_std(Xtrain, theta, L, alpha_vec, y_mean, y_std, Xq, kind="matern32"):
    """GaussianProcessRegressor.predict(return_std=True), SK _gpr.py:446-500."""
    d = Xtrain.shape[1]
    Ks = kernel_cross(Xq, Xtrain, theta, kind)  # :447
    mean = y_std * (Ks @ alpha_vec) + y_mean  # :448-451
    V = scipy.linalg.solve_triangular(L, Ks.T, lower=True, check_finite=False)  # :460
    var = np.full(Xq.shape[0], 1.0 + math.exp(theta[d]))  # kernel_.diag: Matern 1 + White s2 (kernels.py:890,1438)
    var -= np.einsum("ij,ji->i", V.T, V)  # :483
    var[var < 0] = 0.0  # :485-491
    return mean, np.sqrt(var * y_std ** 2)  # :494-500


# --------------------------------------------------------------------------------------
# k-NN router (REF/cGP/classify.py:4; SK/neighbors/_classification.py:245-312)
# --------------------------------------------------------------------------------------
def knn_labels(Xtrain, labels, Q, k=3, chunk=4096):
    """Exact k-NN vote.  rdist = sum_p (q_p - x_p)^2 sequential in p, separate mul/add roundings
    (SK/metrics/_dist_metrics.pxd.tp:44-49).  Equal distances: lowest training index wins (the
    north star's rule; a later point replaces a heap entry only if strictly smaller,
    SK/utils/_heap.pyx:46-47).  Vote ties: smallest label (scipy.stats.mode via SK/utils/fixes.py:58-59)."""
    Xtrain = np.asarray(Xtrain, dtype=np.float64)
    Q = np.asarray(Q, dtype=np.float64)
    labels = np.asarray(labels).astype(np.int64)
    n, d = Xtrain.shape
    k = min(k, n)
    ncls = int(labels.max()) + 1
    out = np.empty(Q.shape[0], dtype=np.int64)
    for s in range(0, Q.shape[0], chunk):
        q = Q[s:s + chunk]
        dist = np.zeros((q.shape[0], n))
        for p in range(d):
            t = q[:, p, None] - Xtrain[None, :, p]
            dist += t * t
        idx = np.argsort(dist, axis=1, kind="stable")[:, :k]
        votes = np.zeros((q.shape[0], ncls), dtype=np.int64)
        nl = labels[idx]
        for j in range(k):
            np.add.at(votes, (np.arange(q.shape[0]), nl[:, j]), 1)
        out[s:s + chunk] = np.argmax(votes, axis=1)  # first max = smallest label
    return out


# --------------------------------------------------------------------------------------
# acquisition (REF/cGP/acquisition.py:4-32) and cross-cluster selection (REF/cGP/cGP_constrained.py:305,376-378)
# --------------------------------------------------------------------------------------
def ei_scores(mean, std, y_max, n_c):
    """Vectorised expected_improvement / n_c.  REF/cGP/acquisition.py:23-32 with xi=0:
    imp = mu - max(Y); Z = imp/sigma; ei = imp*Phi(Z) + sigma*phi(Z); sigma<=0 -> 0; score = ei/n_c.
    norm.cdf == scipy.special.ndtr, norm.pdf == exp(-z^2/2)/sqrt(2 pi)."""
    mean = np.asarray(mean, dtype=np.float64)
    std = np.asarray(std, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        imp = mean - y_max
        Z = imp / std
        ei = imp * ndtr(Z) + std * (np.exp(-0.5 * Z * Z) / math.sqrt(2 * math.pi))
    ei[std <= 0.0] = 0.0
    return ei / n_c


def select_next(Xtrain, y_raw, labels, models, Q, k=3):
    """Dense candidate scan that replaces the reference's per-cluster trust-constr search
    (REF/cGP/cGP_constrained.py:359-378): route every candidate with k-NN (acquisition.py:7-9),
    score it with its own cluster's GP, take the max; ties -> lowest cluster id, then lowest
    candidate index (ascending-c strict '>' at cGP_constrained.py:305,376-378).

    models[c] = dict from fit_gpr for cluster c (clusters are 0..C-1 after recode)."""
    Xtrain = np.asarray(Xtrain, dtype=np.float64)
    y_raw = np.asarray(y_raw, dtype=np.float64).reshape(-1)
    labels = np.asarray(labels).astype(np.int64)
    lab = knn_labels(Xtrain, labels, Q, k)
    M = Q.shape[0]
    mean = np.zeros(M)
    std = np.zeros(M)
    score = np.zeros(M)
    best = (-np.inf, -1, -1)
    for c, m in enumerate(models):
        sel = np.nonzero(lab == c)[0]
        if sel.size == 0:
            continue
        mu, sd = predict_mean_std(m["X"], m["theta"], m["L"], m["alpha_vec"], m["y_mean"], m["y_std"],
                                  Q[sel], m["kind"])
        sc = ei_scores(mu, sd, float(np.max(y_raw[labels == c])), int(np.sum(labels == c)))
        mean[sel], std[sel], score[sel] = mu, sd, sc
        j = int(np.argmax(sc))  # first max
        if sc[j] > best[0]:
            best = (float(sc[j]), int(sel[j]), c)
    return dict(labels=lab, mean=mean, std=std, score=score, best_score=best[0], best_idx=best[1],
                best_cluster=best[2], x_next=Q[best[1]].copy())


# --------------------------------------------------------------------------------------
# host pre-steps that define the inputs (REF/cGP/cGP_constrained.py:249-255,279-284)
# --------------------------------------------------------------------------------------
def merge_small_clusters(labels, d):
    """Clusters with count <= d+2 are merged into the currently largest one (:279-282)."""
    labels = np.array(labels).astype(np.int64)
    for c in np.unique(labels):
        if np.sum(labels == c) <= d + 2:
            occ = np.bincount(labels)
            labels[labels == c] = np.argmax(occ)
    return labels


def recode(labels):
    """Relabel to 0..C-1 in ascending order of the original labels (:249-255)."""
    labels = np.array(labels)
    out = labels.copy()
    for ck, j in enumerate(np.unique(labels)):
        out[labels == j] = ck
    return out
