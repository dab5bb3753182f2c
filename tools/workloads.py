"""Synthetic inputs of the BASELINE.json configs (data definitions: SURVEY.md §8d).  Shared by bench.py, the golden
generator (oracle/make_golden.py), tools/config_runs.py and the tests, so that every consumer sees the same bits."""
from __future__ import annotations

import numpy as np


def schaffer_n2(X):
    x, yy = X[:, 0], X[:, 1]
    return 0.5 + (np.sin(x * x - yy * yy) ** 2 - 0.5) / (1 + 0.001 * (x * x + yy * yy)) ** 2


def cfg2_samples(N=2000, seed=123):
    """configs[1]: Schaffer N2 on [-100, 100]^2, N_PILOT = 2000 drawn as the reference draws pilots
    (REF/cGP/cGP_constrained.py:196-199: np.random.seed(RND_SEED), one uniform column per dimension)."""
    np.random.seed(seed)
    d = 2
    X = np.zeros((N, d))
    for j in range(d):
        X[:, j] = np.random.uniform(-100, 100, size=(N, 1)).ravel()
    return X, schaffer_n2(X)


def cfg2_candidates(M=65536):
    return np.random.default_rng(2024).uniform(-100, 100, size=(M, 2))


def cfg3(n=32768, d=6, C=8, M=4194304):
    """configs[2]: X ~ default_rng(0).uniform[0,1]^6, r = 4[x0>.5]+2[x1>.5]+[x2>.5] (mod C),
    y = r + sin(2 pi (r+1) sum(x)/6) + 0.01 N(0,1), labels = r, candidates ~ default_rng(1).uniform."""
    rng = np.random.default_rng(0)
    X = rng.uniform(size=(n, d))
    r = (4 * (X[:, 0] > .5) + 2 * (X[:, 1] > .5) + (X[:, 2] > .5)).astype(np.int64) % C
    y = r + np.sin(2 * np.pi * (r + 1) * X.sum(1) / 6) + 0.01 * rng.standard_normal(n)
    Q = np.random.default_rng(1).uniform(size=(M, d)) if M else np.zeros((0, d))
    return X, y, r, Q


def cfg4(n=65536, M=8388608, d=4):
    """configs[3]: ONE cluster, 4-D (blocked DMMA Cholesky stress)."""
    rng = np.random.default_rng(0)
    X = rng.uniform(size=(n, d))
    y = np.sin(4 * np.pi * X[:, 0]) * np.cos(2 * np.pi * X[:, 1]) + X[:, 2] * X[:, 3] + 1e-2 * rng.standard_normal(n)
    Q = np.random.default_rng(1).uniform(size=(M, d)) if M else np.zeros((0, d))
    return X, y, np.zeros(n, dtype=np.int64), Q


def cfg5(C=64, M=16777216, d=8, lo=480, hi=544):
    """configs[4]: 8-D, C clusters of n_c in [lo, hi] (cells of a 2^6 split on x0..x5), candidates uniform."""
    rng = np.random.default_rng(0)
    sizes = rng.integers(lo, hi + 1, size=C)
    Xs, labs = [], []
    for c in range(C):
        x = rng.uniform(size=(sizes[c], d))
        for b in range(6):
            x[:, b] = 0.5 * x[:, b] + 0.5 * ((c >> b) & 1)
        Xs.append(x)
        labs.append(np.full(sizes[c], c))
    X, lab = np.concatenate(Xs), np.concatenate(labs)
    perm = rng.permutation(X.shape[0])
    X, lab = X[perm], lab[perm]
    y = lab % 7 + np.sin(2 * np.pi * X.sum(1) / 8) + 0.01 * rng.standard_normal(X.shape[0])
    Q = np.random.default_rng(1).uniform(size=(M, d)) if M else np.zeros((0, d))
    return X, y, lab.astype(np.int64), Q
