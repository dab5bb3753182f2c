"""Whole-model clustered-GP surface: the batched entry points the reference lacks.

``CGPRegressor.fit(X, y, labels)`` is the body of the per-cluster loop of REF/cGP/cGP_constrained.py:305-327
(all clusters x restarts fitted in lock-step on the GPU); ``predict`` is the whole-model predict defined by
REF/cGP/f_truth_dill.py:20-30 (classify, then the routed cluster's GP); ``select_next`` replaces the per-cluster
trust-constr EI search of REF/cGP/cGP_constrained.py:359-378 by a dense candidate scan with the same winner rule.
"""
from __future__ import annotations

import numpy as np

from .engine import ClusterGPEngine
from .gpr import GaussianProcessRegressor, KNeighborsClassifier, parse_kernel


class CGPRegressor:
    def __init__(self, kernel=None, *, alpha=1e-5, n_restarts_optimizer=None, normalize_y=True, random_state=0,
                 n_neighbors=3, optimizer="fmin_l_bfgs_b", max_lockstep=None, mem_fraction=0.6):
        self.kernel = kernel
        self.alpha = alpha
        self.n_restarts_optimizer = n_restarts_optimizer  # None -> 10*d, REF/cGP/cGP_constrained.py:316
        self.normalize_y = normalize_y
        self.random_state = random_state
        self.n_neighbors = n_neighbors
        self.optimizer = optimizer
        self.max_lockstep = max_lockstep  # None: run every L-BFGS-B run to convergence (engine.fit)
        self.mem_fraction = mem_fraction  # share of the free device memory the factor workspaces may take

    def fit(self, X, y, labels):
        X = np.asarray(X, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64).reshape(-1)
        labels = np.asarray(labels).astype(np.int64).reshape(-1)
        d = X.shape[1]
        kind, theta0, bounds, template = parse_kernel(self.kernel, d)
        R = 10 * d if self.n_restarts_optimizer is None else int(self.n_restarts_optimizer)
        eng = ClusterGPEngine(X, y, labels, kind, self.alpha, self.normalize_y, mem_fraction=self.mem_fraction)
        eng.fit(R, seed=self.random_state, theta0=theta0, bounds=bounds, optimize=self.optimizer is not None,
                max_lockstep=self.max_lockstep)
        self.engine_ = eng
        self._template = template
        self._bounds = bounds
        self.n_clusters_ = eng.C
        self.theta_ = eng.theta.copy()
        self.log_marginal_likelihood_value_ = eng.model["lml_host"].copy()
        self.classify_ = KNeighborsClassifier(n_neighbors=min(self.n_neighbors, X.shape[0])).fit(X, labels)
        self._gpr_list = None
        return self

    def partial_fit(self, x, y, label, reoptimize=False, n_restarts=0):
        """Incremental refit: add one sample to cluster `label`.  With reoptimize=False theta stays
        and only that cluster's factor grows by a row (exact: equals refactorising the extended data at the same theta);
        with reoptimize=True every cluster is re-optimised starting from its current theta (+ n_restarts random starts)."""
        eng = self.engine_
        eng.append(x, y, label)
        if reoptimize and self.optimizer is not None:
            eng.fit(int(n_restarts), seed=self.random_state, theta0=eng.theta, bounds=self._bounds)
        self.theta_ = eng.theta.copy()
        self.log_marginal_likelihood_value_ = eng.model["lml_host"].copy()
        self.classify_ = KNeighborsClassifier(n_neighbors=min(self.n_neighbors, eng.n)).fit(eng.X_host, eng.labels_host)
        self._gpr_list = None
        return self

    @property
    def GPR_list(self):
        """Per-cluster regressors with the scikit-learn attribute surface (the reference pickles this list,
        REF/cGP/cGP_constrained.py:539)."""
        if self._gpr_list is None:
            params = dict(kernel=self.kernel, alpha=self.alpha, optimizer=self.optimizer,
                          n_restarts_optimizer=self.n_restarts_optimizer or 0, normalize_y=self.normalize_y,
                          random_state=self.random_state)
            self._gpr_list = [GaussianProcessRegressor._from_engine(self.engine_, c, self._template, params)
                              for c in range(self.n_clusters_)]
        return self._gpr_list

    def predict(self, X, return_std=False):
        X = np.asarray(X, dtype=np.float64)
        mean, std = self.engine_.predict(X, k=self.n_neighbors, return_std=return_std)
        if return_std:
            return mean.cpu().numpy(), std.cpu().numpy()
        return mean.cpu().numpy()

    def predict_labels(self, X):
        return self.engine_.route(np.asarray(X, dtype=np.float64), self.n_neighbors).cpu().numpy()

    def score_candidates(self, candidates):
        """-> dict(labels, mean, std, score, best_score, best_idx, best_cluster) on the host (single rank)."""
        eng = self.engine_
        Q = eng._to_dev(candidates)
        ql = eng.route(Q, self.n_neighbors)
        (bs, bi, bc), (score, mean, std) = eng.ei_scan(Q, self.n_neighbors, want_arrays=True, qlabel=ql)
        return dict(labels=ql.cpu().numpy(), mean=mean.cpu().numpy(), std=std.cpu().numpy(), score=score.cpu().numpy(),
                    best_score=float(bs.item()), best_idx=int(bi.item()), best_cluster=int(bc.item()))

    def select_next(self, candidates, sharded=False, prune=True, boundary_penalty=None):
        """-> (x_next [d], score, idx).  Multi-GPU aware (see cgp_b200/dist.py).  prune=True (default): exact bound-pruned
        scan, same selection and score bits (candidates whose EI bound from the predictive mean cannot reach the running
        best skip the variance contraction); prune=False: exhaustive scan."""
        x, score, idx, _cl = self.engine_.select_next(candidates, self.n_neighbors, sharded=sharded, prune=prune,
                                                      boundary_penalty=boundary_penalty)
        return x, score, idx

    def select_next_mspe(self, candidates, sharded=False, boundary_penalty=None):
        """Dense scan of the reference's second acquisition (REF/cGP/acquisition.py:34-75, minimised):
        -> (x_next [d], (mspe + penalty) / n_c, idx).  Multi-GPU aware like select_next."""
        x, v, idx, _cl = self.engine_.select_next_mspe(candidates, self.n_neighbors, sharded=sharded,
                                                        boundary_penalty=boundary_penalty)
        return x, v, idx

    def mspe_candidates(self, candidates):
        """mspe / n_c of every candidate (host array), the batched form of acquisition.mean_square_prediction_error."""
        eng = self.engine_
        _, arr = eng.mspe_scan(eng._to_dev(candidates), self.n_neighbors, want_arrays=True)
        return arr.cpu().numpy()
