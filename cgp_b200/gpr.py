"""scikit-learn-shaped surfaces of the B200 backend.

``GaussianProcessRegressor`` and ``KNeighborsClassifier`` are drop-ins for the objects the reference builds at
REF/cGP/cGP_constrained.py:98,290,315-316 and passes to REF/cGP/acquisition.py:4-32: same constructor keywords,
``fit`` / ``predict`` / ``log_marginal_likelihood`` and the fitted attributes the reference reads or pickles
(``kernel_``, ``X_train_``, ``y_train_``, ``L_``, ``alpha_``, ``_y_train_mean``, ``_y_train_std``,
``log_marginal_likelihood_value_``).  scikit-learn kernel objects are accepted as *parameter containers*
(theta, bounds); no scikit-learn arithmetic runs on the hot path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .engine import ClusterGPEngine


def parse_kernel(kernel, d):
    """-> (kind, theta0 [d+1] in ARD layout, bounds [d+1,2], template).  Accepts 'matern32' / 'rbf' or a
    scikit-learn ``Matern(nu=1.5)|RBF + WhiteKernel`` sum (REF/cGP/cGP_constrained.py:169)."""
    from .engine import LOG_HI, LOG_LO

    if kernel is None:
        kernel = "matern32"
    if isinstance(kernel, str):
        if kernel not in _lib.KIND:
            raise ValueError(f"unknown kernel {kernel!r}")
        return kernel, np.zeros(d + 1), np.tile([LOG_LO, LOG_HI], (d + 1, 1)), None
    k1, k2 = getattr(kernel, "k1", None), getattr(kernel, "k2", None)
    n1, n2 = type(k1).__name__, type(k2).__name__
    if n2 != "WhiteKernel" or n1 not in ("Matern", "RBF"):
        raise NotImplementedError("the B200 backend supports Matern(nu=1.5)+WhiteKernel and RBF+WhiteKernel "
                                  f"(REF/cGP/cGP_constrained.py:169); got {kernel!r}")
    if n1 == "Matern" and abs(float(k1.nu) - 1.5) > 1e-12:
        raise NotImplementedError("only Matern nu=1.5 is implemented on the GPU")
    ls = np.atleast_1d(np.asarray(k1.length_scale, dtype=np.float64))
    if ls.shape[0] == 1 and d > 1:
        raise NotImplementedError("isotropic length_scale with d>1: pass length_scale=np.ones(d) (ARD) as the "
                                  "reference does")
    if ls.shape[0] != d:
        raise ValueError("length_scale dimension does not match X")
    b = np.asarray(kernel.bounds, dtype=np.float64)
    if b.shape != (d + 1, 2) or not np.all(np.isfinite(b)):
        raise NotImplementedError("all hyper-parameters must be free with finite bounds")
    return ("matern32" if n1 == "Matern" else "rbf"), np.asarray(kernel.theta, float).copy(), b, kernel


class _KernelView:
    """Minimal kernel_ stand-in when no scikit-learn template was given."""

    def __init__(self, kind, theta, bounds):
        self.kind, self.theta, self.bounds = kind, np.asarray(theta, float), np.asarray(bounds, float)

    def __repr__(self):
        d = len(self.theta) - 1
        name = "Matern(nu=1.5)" if self.kind == "matern32" else "RBF"
        return f"{name}(length_scale={np.exp(self.theta[:d])}) + WhiteKernel(noise_level={np.exp(self.theta[d]):.3g})"


class GaussianProcessRegressor:
    """Single-cluster GP on the GPU with scikit-learn's GaussianProcessRegressor surface (SK/_gpr.py:233-500)."""

    def __init__(self, kernel=None, *, alpha=1e-10, optimizer="fmin_l_bfgs_b", n_restarts_optimizer=0,
                 normalize_y=False, copy_X_train=True, n_targets=None, random_state=None):
        self.kernel = kernel
        self.alpha = alpha
        self.optimizer = optimizer
        self.n_restarts_optimizer = n_restarts_optimizer
        self.normalize_y = normalize_y
        self.copy_X_train = copy_X_train
        self.n_targets = n_targets
        self.random_state = random_state

    # -- construction from an already fitted cluster of a ClusterGPEngine (used by CGPRegressor)
    @classmethod
    def _from_engine(cls, eng, c, template, params):
        self = cls(**params)
        self._bind(eng, c, template)
        return self

    def _bind(self, eng, c, template):
        self._eng, self._c = eng, c
        s = slice(int(eng.offs[c]), int(eng.offs[c + 1]))
        self.X_train_ = eng.Xs[s].cpu().numpy()
        self.y_train_ = eng.yn_host[s].reshape(-1, 1) if self._y2d else eng.yn_host[s].copy()
        self._y_train_mean = np.array([eng.y_mean[c]]) if self._y2d else np.float64(eng.y_mean[c])
        self._y_train_std = np.array([eng.y_std[c]]) if self._y2d else np.float64(eng.y_std[c])
        theta = eng.theta[c]
        if template is not None:
            from sklearn.base import clone

            self.kernel_ = clone(template)
            self.kernel_.theta = theta
        else:
            self.kernel_ = _KernelView(eng.kind, theta, np.tile([np.log(1e-5), np.log(1e5)], (len(theta), 1)))
        self.log_marginal_likelihood_value_ = float(eng.model["lml_host"][c])
        self._L = self._alpha_ = None

    _y2d = False

    def fit(self, X, y):
        X = np.asarray(X, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        self._y2d = y.ndim == 2
        if self._y2d and y.shape[1] != 1:
            raise NotImplementedError("multi-output targets are not supported on the GPU path")
        if self.random_state is not None and not isinstance(self.random_state, (int, np.integer)):
            raise NotImplementedError("random_state must be None or an int")
        kind, theta0, bounds, template = parse_kernel(self.kernel, X.shape[1])
        eng = ClusterGPEngine(X, y.reshape(-1), np.zeros(X.shape[0], dtype=np.int64), kind, self.alpha, self.normalize_y)
        seed = self.random_state  # check_random_state(None) is the global RNG in sklearn; we use a fresh one
        eng.fit(self.n_restarts_optimizer, seed=seed, theta0=theta0, bounds=bounds,
                optimize=self.optimizer is not None)
        self._bind(eng, 0, template)
        return self

    @property
    def L_(self):
        if self._L is None:
            self._L = self._eng.L_of(self._c)
        return self._L

    @property
    def alpha_(self):
        if self._alpha_ is None:
            a = self._eng.alpha_of(self._c)
            self._alpha_ = a.reshape(-1, 1) if self._y2d else a
        return self._alpha_

    def predict(self, X, return_std=False, return_cov=False):
        if return_std and return_cov:
            raise RuntimeError("At most one of return_std or return_cov can be requested.")  # SK/_gpr.py:405-408
        X = np.asarray(X, dtype=np.float64)
        if X.ndim != 2:
            raise ValueError("Expected 2D array")
        eng = self._eng
        if return_cov:
            return self._predict_cov(X)
        ql = torch.full((X.shape[0],), self._c, dtype=torch.int64, device=eng.device)
        mean, std = eng.predict(X, qlabel=ql, return_std=return_std)
        mean = mean.cpu().numpy()
        if return_std:
            return mean, std.cpu().numpy()
        return mean

    def _predict_cov(self, X):
        """Joint posterior covariance, SK/_gpr.py:464-478: ``kernel_(X) - V^T V`` with ``V = L^-1 K*^T``, through the
        same DMMA tile GEMM as the factorisation (W = L^-1 is resident): Vt = K* W^T, V^T V = Vt Vt^T."""
        eng, c = self._eng, self._c
        m = X.shape[0]
        n, npad = int(eng.counts[c]), int(eng.npad[c])
        mpad = _lib.padded_n(max(m, 1))
        dev = eng.device
        Xd = torch.from_numpy(np.ascontiguousarray(X)).to(dev)
        th = eng.model["theta_dev"][c].contiguous()
        Xt = eng.Xs[int(eng.offs[c]):int(eng.offs[c + 1])]
        Ks = torch.zeros((mpad, npad), dtype=torch.float64, device=dev)
        Ks[:m, :n] = _lib.kernel_cross(Xd, Xt, th, eng.kind)
        W = eng.model["W"][int(eng.moff[c]):int(eng.moff[c + 1])].view(npad, npad)
        Vt = _lib.dgemm_nt(Ks, W)  # [mpad, npad], Vt[q, i] = sum_k K*[q, k] W[i, k]
        VtV = _lib.dgemm_nt(Vt, Vt)[:m, :m]
        ql = torch.full((m,), c, dtype=torch.int64, device=dev)
        mean, _ = eng.predict(Xd, qlabel=ql, return_std=False)
        cov =(_lib.kernel_matrix(Xd, th, eng.kind, 0.0) - VtV) * (eng.y_std[c] ** 2)
        return mean.cpu().numpy(), cov.cpu().numpy()

    def log_marginal_likelihood(self, theta=None, eval_gradient=False, clone_kernel=True):
        if theta is None:
            if eval_gradient:
                raise ValueError("Gradient can only be evaluated for theta!=None")  # SK/_gpr.py:573-576
            return self.log_marginal_likelihood_value_
        lml, grad, _ = self._eng.lml_grad([self._c], np.asarray(theta, float)[None, :], eval_gradient)
        if eval_gradient:
            return float(lml[0]), grad[0]
        return float(lml[0])

    # dill/pickle: keep host arrays only (device state is rebuilt lazily on predict)
    def __getstate__(self):
        st = {k: v for k, v in self.__dict__.items() if k not in ("_eng",)}
        st["_L"] = self.L_
        st["_alpha_"] = self.alpha_
        eng = self._eng
        s = slice(int(eng.offs[self._c]), int(eng.offs[self._c + 1]))
        st["_pk"] = dict(kind=eng.kind, alpha=eng.alpha, theta=eng.theta[self._c].copy(),
                         y_raw=eng.y_host[eng.order][s].copy())
        return st

    def __setstate__(self, st):
        pk = st.pop("_pk")
        self.__dict__.update(st)
        self._pk = pk
        self.__dict__["_eng_lazy"] = True

    def __getattr__(self, name):
        if name == "_eng" and self.__dict__.get("_eng_lazy"):
            pk = self._pk
            eng = ClusterGPEngine(self.X_train_, pk["y_raw"], np.zeros(len(pk["y_raw"]), dtype=np.int64), pk["kind"],
                                  pk["alpha"], self.normalize_y)
            eng.set_theta(pk["theta"][None, :])
            self.__dict__["_eng"] = eng
            self.__dict__["_c"] = 0
            self.__dict__["_eng_lazy"] = False
            return eng
        raise AttributeError(name)


class KNeighborsClassifier:
    """GPU k-NN router with the surface of the reference's classify plugin object
    (REF/cGP/classify.py:3-9: ``fit(X, labels)``, ``predict(X)``)."""

    def __init__(self, n_neighbors=3):
        self.n_neighbors = n_neighbors

    def fit(self, X, y):
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
        y = np.asarray(y).reshape(-1)
        if X.ndim != 2 or X.shape[0] != y.shape[0]:
            raise ValueError("X [n,d] and y [n] expected")
        if self.n_neighbors > X.shape[0]:
            raise ValueError(f"Expected n_neighbors <= n_samples_fit, but n_neighbors = {self.n_neighbors}, "
                             f"n_samples_fit = {X.shape[0]}")
        self.classes_, inv = np.unique(y, return_inverse=True)
        self._fit_X = X
        self._y = inv.astype(np.int64)
        self.n_samples_fit_ = X.shape[0]
        self._dev = None
        return self

    def _device_state(self):
        if self._dev is None:
            if not torch.cuda.is_available():
                raise _lib.CgpError("cgp_b200 needs a CUDA device (B200); there is no CPU fallback")
            dev = torch.device("cuda", torch.cuda.current_device())
            self._dev = (torch.from_numpy(self._fit_X).to(dev), torch.from_numpy(self._y).to(dev))
        return self._dev

    def predict(self, X):
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
        if X.ndim != 2 or X.shape[1] != self._fit_X.shape[1]:
            raise ValueError("X has a different number of features than the training data")
        Xd, yd = self._device_state()
        Q = torch.from_numpy(X).to(Xd.device)
        idx = _lib.knn_labels(Xd, yd, Q, self.n_neighbors, len(self.classes_)).cpu().numpy()
        return self.classes_[idx]

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_dev"] = None
        return st
