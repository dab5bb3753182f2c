"""Extended-precision truth for the posterior variance (TEST INFRASTRUCTURE ONLY).

``var = (1 + s2) - || L^-1 k* ||^2`` (SK/gaussian_process/_gpr.py:460-483) cancels where a candidate sits close to
training points, so two correct fp64 implementations (LAPACK ``dtrtrs`` inside scikit-learn, the CUDA path) can
differ from each other by more than 1e-8 relative while both are that close to the exact value.  This module
evaluates the same formula in x87 80-bit ``np.longdouble`` (eps = 1.1e-19: 3-4 decimal digits beyond fp64) from
the SAME fp64 kernel values, so that parity tests can say which side is off and by how much.

Pure NumPy loops, O(n^3) without BLAS: use for n <= ~600 (n = 512: < 1 s).
"""
from __future__ import annotations

import math

import numpy as np

from . import gp_oracle as O

LD = np.longdouble


def cholesky_ld(K):
    """Lower Cholesky factor of K in long double (column by column; SK _gpr.py:352 in higher precision)."""
    K = np.asarray(K, dtype=LD)
    n = K.shape[0]
    L = np.zeros((n, n), dtype=LD)
    for j in range(n):
        s = K[j, j] - np.dot(L[j, :j], L[j, :j])
        if not s > 0:
            raise np.linalg.LinAlgError(f"{j + 1}-th leading minor not positive definite")
        L[j, j] = np.sqrt(s)
        if j + 1 < n:
            L[j + 1:, j] = (K[j + 1:, j] - L[j + 1:, :j] @ L[j, :j]) / L[j, j]
    return L


def solve_lower_ld(L, B):
    """V = L^-1 B by forward substitution in long double (SK _gpr.py:460)."""
    L = np.asarray(L, dtype=LD)
    B = np.asarray(B, dtype=LD)
    V = np.zeros_like(B)
    for i in range(L.shape[0]):
        V[i] = (B[i] - L[i, :i] @ V[:i]) / L[i, i]
    return V


def predict_var_truth(Xtrain, theta, Xq, kind="matern32", alpha=1e-5, L64=None):
    """Posterior variance in NORMALISED target units, long double, from the fp64 kernel values.

    L64=None: the whole chain (Cholesky included) in long double  -> "exact" value of the formula for the fp64 K;
    L64 given: the triangular solve against that fp64 factor only  -> what a perfect solve with scikit-learn's own
    ``L_`` would return (scikit-learn's definition of the result, SK _gpr.py:460 uses self.L_)."""
    d = Xtrain.shape[1]
    Ks = O.kernel_cross(Xq, Xtrain, theta, kind)
    if L64 is None:
        L = cholesky_ld(O.kernel_matrix(Xtrain, theta, kind, alpha))
    else:
        L = np.asarray(L64, dtype=LD)
    V = solve_lower_ld(L, Ks.T)
    prior = LD(1.0) + LD(math.exp(theta[d]))  # kernel_.diag: Matern 1 + White s2, alpha not included
    return prior - np.sum(V * V, axis=0)


def mean_truth(Xtrain, y_norm, theta, Xq, kind="matern32", alpha=1e-5):
    """Posterior mean in normalised units, long double (alpha = K^-1 y through the long-double factor)."""
    Ks = O.kernel_cross(Xq, Xtrain, theta, kind)
    L = cholesky_ld(O.kernel_matrix(Xtrain, theta, kind, alpha))
    z = solve_lower_ld(L, np.asarray(y_norm, dtype=LD).reshape(-1, 1))
    return (solve_lower_ld(L, Ks.T).T @ z).reshape(-1)


def lml_truth(Xtrain, y_norm, theta, kind="matern32", alpha=1e-5):
    """Log-marginal likelihood (SK _gpr.py:613-617) in long double."""
    L = cholesky_ld(O.kernel_matrix(Xtrain, theta, kind, alpha))
    z = solve_lower_ld(L, np.asarray(y_norm, dtype=LD).reshape(-1, 1)).reshape(-1)
    n = Xtrain.shape[0]
    return -LD(0.5) * np.dot(z, z) - np.sum(np.log(np.diag(L))) - LD(n) / 2 * LD(math.log(2 * math.pi))
