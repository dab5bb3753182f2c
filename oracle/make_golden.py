"""Generate tests/golden/*.npz from the REFERENCE's own code path (run in the build container only).

TEST INFRASTRUCTURE.  Reads /root/reference (cGP/BUKIN.pkl, cGP/acquisition.py) and the installed
scikit-learn 1.9.0 / SciPy 1.18.1 configured exactly as REF/cGP/cGP_constrained.py:169,290,315-316.
The GPU box has no /root/reference: tests only read the committed .npz files.

    python oracle/make_golden.py
"""
from __future__ import annotations

import io
import os
import pickle
import sys
import types
import warnings

import numpy as np

REF = "/root/reference/cGP"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


# ---------------------------------------------------------------- BUKIN.pkl (sklearn 0.24.1 pickle)
class _Stub:
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.state = state


def _load_bukin():
    """Permissive unpickler: sklearn.neighbors._kd_tree / _dist_metrics of 0.24 no longer exist."""
    import dill

    class U(dill.Unpickler):
        def find_class(self, module, name):
            if module.startswith("sklearn.neighbors._kd_tree") or module.startswith("sklearn.neighbors._dist_metrics") \
                    or module.startswith("sklearn.neighbors._ball_tree"):
                if name.startswith("newObj"):
                    return lambda cls, *a: cls.__new__(cls)
                return _Stub
            return super().find_class(module, name)

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with open(os.path.join(REF, "BUKIN.pkl"), "rb") as f:
            return U(f).load()


def bukin():
    obj = _load_bukin()
    out = {}
    for i, g in enumerate(obj["GPR_list"]):
        d = g.__dict__
        out[f"g{i}_X"] = np.asarray(d["X_train_"], float)
        out[f"g{i}_y"] = np.asarray(d["y_train_"], float).reshape(-1)
        out[f"g{i}_theta"] = np.asarray(d["kernel_"].theta, float)
        out[f"g{i}_L"] = np.asarray(d["L_"], float)
        out[f"g{i}_alpha"] = np.asarray(d["alpha_"], float).reshape(-1)
        out[f"g{i}_ymean"] = np.asarray(d["_y_train_mean"], float).reshape(-1)
        out[f"g{i}_ystd"] = np.asarray(d["_y_train_std"], float).reshape(-1)
        out[f"g{i}_lml"] = np.asarray(d["log_marginal_likelihood_value_"], float).reshape(-1)
    clf = obj["Classify"].__dict__
    out["knn_X"] = np.asarray(clf["_fit_X"], float)
    out["knn_y"] = np.asarray(clf["_y"]).astype(np.int64)
    out["knn_k"] = np.asarray([clf["n_neighbors"]])
    out["knn_classes"] = np.asarray(clf["classes_"]).astype(np.int64)
    # What REF/cGP/f_truth_dill.py:20-30 returns for a few inputs: k-NN label -> that cluster's GPR.predict.  The
    # 0.24.1 objects cannot predict under scikit-learn 1.9 (their KD-tree classes are gone), so the same objects are
    # rebuilt from the pickled arrays: GPR at the stored theta (optimizer=None) on the de-normalised targets.
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.neighbors import KNeighborsClassifier

    knn = KNeighborsClassifier(n_neighbors=int(clf["n_neighbors"])).fit(out["knn_X"], out["knn_classes"][out["knn_y"]])
    gprs = []
    for i, g in enumerate(obj["GPR_list"]):
        d = g.__dict__
        r = GaussianProcessRegressor(kernel=d["kernel_"].clone_with_theta(out[f"g{i}_theta"]), alpha=d["alpha"],
                                     normalize_y=True, optimizer=None, random_state=0)
        r.fit(out[f"g{i}_X"], (out[f"g{i}_y"] * out[f"g{i}_ystd"][0] + out[f"g{i}_ymean"][0]).reshape(-1, 1))
        assert np.max(np.abs(r.L_ - out[f"g{i}_L"])) < 1e-9, "rebuilt factor differs from the pickled one"
        gprs.append(r)
    rng = np.random.default_rng(5)
    xs = np.vstack([[-10.0, 1.0], [1.0, 1.0], np.column_stack([rng.uniform(-15, -5, 30), rng.uniform(-3, 3, 30)])])
    lab = knn.predict(xs)
    mu, sd = np.zeros(len(xs)), np.zeros(len(xs))
    for j, x in enumerate(xs):
        m_, s_ = gprs[int(lab[j])].predict(x.reshape(1, -1), return_std=True)
        mu[j], sd[j] = np.asarray(m_).reshape(-1)[0], s_[0]
    out.update(ft_x=xs, ft_label=lab.astype(np.int64), ft_mean=mu, ft_std=sd)
    np.savez_compressed(os.path.join(OUT, "bukin.npz"), **out)
    print("bukin.npz", {k: v.shape for k, v in out.items()})


# ---------------------------------------------------------------- reference-configured sklearn objects
def _ref_gpr(d, R, kind="matern32"):
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, Matern, WhiteKernel

    if kind == "matern32":  # REF/cGP/cGP_constrained.py:169
        k = Matern(length_scale=np.ones(d), length_scale_bounds=(1e-05, 100000.0), nu=3 / 2)
    else:
        k = RBF(length_scale=np.ones(d), length_scale_bounds=(1e-05, 100000.0))
    k = k + WhiteKernel(noise_level=1.0, noise_level_bounds=(1e-05, 100000.0))
    # REF/cGP/cGP_constrained.py:315-316
    return GaussianProcessRegressor(kernel=k, random_state=0, normalize_y=True, alpha=1e-5,
                                    optimizer="fmin_l_bfgs_b", n_restarts_optimizer=R)


def _ref_acquisition():
    sys.path.insert(0, REF)
    import acquisition  # REF/cGP/acquisition.py

    return acquisition


def _piecewise(X, rng):
    r = (X[:, 0] > 0.5).astype(float) * 2 + (X[:, 1] > 0.5)
    return r + np.sin(2 * np.pi * (r + 1) * X.sum(1) / X.shape[1]) + 0.01 * rng.standard_normal(X.shape[0])


def case_lml(name, n, d, kind, seed, nthetas=6):
    """LML value + gradient at fixed thetas, K, through sklearn's own log_marginal_likelihood."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(size=(n, d))
    y = _piecewise(X, rng)
    g = _ref_gpr(d, 0, kind)
    g.optimizer = None
    g.fit(X, y)  # normalises y, sets kernel_
    thetas = np.vstack([np.zeros(d + 1), rng.uniform(-3.0, 2.0, size=(nthetas - 1, d + 1))])
    thetas[2, -1] = np.log(1e-5)  # noise at its lower bound
    lml = np.zeros(nthetas)
    grad = np.zeros((nthetas, d + 1))
    for i, th in enumerate(thetas):
        lml[i], grad[i] = g.log_marginal_likelihood(th, eval_gradient=True)
    K0 = g.kernel_.clone_with_theta(thetas[1])(X)
    out = dict(X=X, y=y, y_norm=g.y_train_.reshape(-1), thetas=thetas, lml=lml, grad=grad, K1=K0,
               kind=np.array([kind]))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "lml", lml)


def case_fit_select(name, n, d, C, R, M, seed, kind="matern32", n_pointwise=64):
    """Full per-cluster fit + k-NN routing + predict + EI + argmax through sklearn objects and the
    reference's acquisition.expected_improvement (pointwise, on a subset)."""
    from sklearn.neighbors import KNeighborsClassifier

    acq = _ref_acquisition()
    rng = np.random.default_rng(seed)
    X = rng.uniform(size=(n, d))
    y = _piecewise(X, rng)
    # labels by construction: C strips along x0 (0..C-1 ascending, already recoded)
    labels = np.minimum((X[:, 0] * C).astype(np.int64), C - 1)
    Q = rng.uniform(size=(M, d))
    clf = KNeighborsClassifier(n_neighbors=3)  # REF/cGP/classify.py:4, cGP_constrained.py:290
    clf.fit(X, labels)  # :293
    qlab = clf.predict(Q)
    assert clf._fit_method == "kd_tree"
    out = dict(X=X, y=y, labels=labels, Q=Q, qlab=qlab.astype(np.int64), R=np.array([R]), kind=np.array([kind]))
    mean = np.zeros(M)
    std = np.zeros(M)
    score = np.zeros(M)
    best = (-np.inf, -1, -1)
    pw_idx, pw_val = [], []
    for c in range(C):
        Xt = X[labels == c]
        Yt = y[labels == c].reshape(-1, 1)
        g = _ref_gpr(d, R, kind)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            g.fit(Xt, Yt)  # :327
        out[f"c{c}_theta"] = g.kernel_.theta
        out[f"c{c}_lml"] = np.array([g.log_marginal_likelihood_value_])
        out[f"c{c}_alpha"] = g.alpha_.reshape(-1)
        out[f"c{c}_ymean"] = np.asarray(g._y_train_mean).reshape(-1)
        out[f"c{c}_ystd"] = np.asarray(g._y_train_std).reshape(-1)
        out[f"c{c}_Ldiag"] = np.diag(g.L_).copy()
        sel = np.nonzero(qlab == c)[0]
        mu, sd = g.predict(Q[sel], return_std=True)
        mu = mu.reshape(-1)
        imp = mu - np.max(Yt)
        from scipy.stats import norm

        with np.errstate(divide="ignore", invalid="ignore"):
            Z = imp / sd
            ei = imp * norm.cdf(Z) + sd * norm.pdf(Z)  # REF/cGP/acquisition.py:23-26
        ei[sd <= 0.0] = 0.0  # :27
        sc = ei / Xt.shape[0]  # :32
        mean[sel], std[sel], score[sel] = mu, sd, sc
        j = int(np.argmax(sc))
        if sc[j] > best[0]:  # cGP_constrained.py:376-378
            best = (float(sc[j]), int(sel[j]), c)
        # pointwise through the reference's own function on a subset (incl. the per-cluster argmax)
        sub = list(sel[:n_pointwise]) + [int(sel[j])]
        for i in sub:
            v = acq.expected_improvement(X=Q[i], surrogate=g, X_sample=Xt, Y_sample=Yt,
                                         correct_label=np.float64(c), classify=clf, USE_SKLEARN=True,
                                         VERBOSE=False)
            pw_idx.append(i)
            pw_val.append(-v)
    out.update(mean=mean, std=std, score=score, best_score=np.array([best[0]]), best_idx=np.array([best[1]]),
               best_cluster=np.array([best[2]]), pw_idx=np.array(pw_idx), pw_val=np.array(pw_val))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "best", best, "thetas", [out[f"c{c}_theta"] for c in range(C)])


def case_cov_mspe(name, n, d, seed, m=40, n_mspe=6):
    """predict(return_cov=True) and the reference's mean_square_prediction_error (REF/cGP/acquisition.py:34-75)
    through sklearn objects, at a fixed theta (optimizer=None)."""
    from sklearn.neighbors import KNeighborsClassifier

    acq = _ref_acquisition()
    rng = np.random.default_rng(seed)
    X = rng.uniform(size=(n, d))
    y = _piecewise(X, rng)
    theta = np.array([-0.8] * d + [-3.0])
    g = _ref_gpr(d, 0)
    g.kernel = g.kernel.clone_with_theta(theta)
    g.optimizer = None
    g.fit(X, y.reshape(-1, 1))
    Q = rng.uniform(size=(m, d))
    mean, cov = g.predict(Q, return_cov=True)
    clf = KNeighborsClassifier(n_neighbors=3).fit(X, np.zeros(n, dtype=np.int64))
    mspe = np.array([acq.mean_square_prediction_error(X=Q[i], surrogate=g, X_sample=X, Y_sample=y.reshape(-1, 1),
                                                      correct_label=np.float64(0), classify=clf, VERBOSE=False)
                     for i in range(n_mspe)])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), X=X, y=y, theta=theta, Q=Q, mean=np.asarray(mean).reshape(-1),
                        cov=cov, mspe=mspe)
    print(name, "cov diag", np.diag(cov)[:4], "mspe", mspe)


def _fit_clusters(X, y, labels, Q, R, kind="matern32", truth_cap=0, k=3):
    """Per-cluster sklearn fit + routed predict + EI + reference winner rule (REF/cGP/cGP_constrained.py:305-378) for a
    dense candidate set.  truth_cap > 0: also the long-double posterior variance / mean of up to truth_cap routed
    candidates per cluster (oracle/extended.py) in NORMALISED units."""
    from scipy.stats import norm
    from sklearn.neighbors import KNeighborsClassifier

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import extended as E

    C, d, M = int(labels.max()) + 1, X.shape[1], Q.shape[0]
    clf = KNeighborsClassifier(n_neighbors=k).fit(X, labels)
    qlab = clf.predict(Q)
    out = dict(qlab=qlab.astype(np.int16))
    mean, std, score = np.zeros(M), np.zeros(M), np.zeros(M)
    best = (-np.inf, -1, -1)
    t_idx, t_var, t_mean, t_lml = [], [], [], []
    for c in range(C):
        Xt, Yt = X[labels == c], y[labels == c].reshape(-1, 1)
        g = _ref_gpr(d, R, kind)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            g.fit(Xt, Yt)
        th = g.kernel_.theta
        out[f"c{c}_theta"] = th
        out[f"c{c}_lml"] = np.array([g.log_marginal_likelihood_value_])
        out[f"c{c}_grad"] = g.log_marginal_likelihood(th, eval_gradient=True)[1]
        out[f"c{c}_ymean"] = np.asarray(g._y_train_mean).reshape(-1)
        out[f"c{c}_ystd"] = np.asarray(g._y_train_std).reshape(-1)
        sel = np.nonzero(qlab == c)[0]
        mu, sd = g.predict(Q[sel], return_std=True)
        mu = mu.reshape(-1)
        imp = mu - np.max(Yt)
        with np.errstate(divide="ignore", invalid="ignore"):
            Z = imp / sd
            ei = imp * norm.cdf(Z) + sd * norm.pdf(Z)  # REF/cGP/acquisition.py:23-26
        ei[sd <= 0.0] = 0.0
        sc = ei / Xt.shape[0]
        mean[sel], std[sel], score[sel] = mu, sd, sc
        if len(sel):
            j = int(np.argmax(sc))
            if sc[j] > best[0]:
                best = (float(sc[j]), int(sel[j]), c)
        if truth_cap:
            ts = sel[:truth_cap]
            yn = g.y_train_.reshape(-1)
            t_idx.append(ts)
            t_var.append(np.asarray(E.predict_var_truth(Xt, th, Q[ts], kind), dtype=np.float64))
            t_mean.append(np.asarray(E.mean_truth(Xt, yn, th, Q[ts], kind), dtype=np.float64))
            t_lml.append(float(E.lml_truth(Xt, yn, th, kind)))
        print(f"  cluster {c}: n_c={Xt.shape[0]} theta={np.round(th, 3)} lml={g.log_marginal_likelihood_value_:.6f}", flush=True)
    out.update(mean=mean, std=std, score=score, best_score=np.array([best[0]]), best_idx=np.array([best[1]]),
               best_cluster=np.array([best[2]]))
    if truth_cap:
        out.update(truth_idx=np.concatenate(t_idx), truth_var_norm=np.concatenate(t_var),
                   truth_mean_norm=np.concatenate(t_mean), truth_lml=np.array(t_lml))
    return out


def case_cfg2_full(name="cfg2_full"):
    """BASELINE configs[1] at FULL size: Schaffer N2, N_PILOT = 2000, DGM4 clustering on [X, y] (REF/cGP/DGM4.py,
    cGP_constrained.py:214-232), R = 10 d = 20 restarts, 65 536 candidates.  X / y / Q are regenerated from seeds by
    tools/workloads.py; the clustering result is stored (it is host-side and sklearn-version dependent)."""
    from sklearn.mixture import BayesianGaussianMixture

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import oracle
    from tools import workloads as W

    X, y = W.cfg2_samples()
    Q = W.cfg2_candidates()
    lab = BayesianGaussianMixture(weight_concentration_prior_type="dirichlet_process", n_components=4,
                                  random_state=123).fit_predict(np.concatenate([X, y[:, None]], axis=1))
    lab = oracle.recode(oracle.merge_small_clusters(lab, 2))
    out = _fit_clusters(X, y, lab, Q, R=20)
    out.update(labels=lab.astype(np.int16), R=np.array([20]))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "sizes", np.bincount(lab), "best", out["best_score"], out["best_idx"], out["best_cluster"])


def case_cfg5_shape(name="cfg5_shape"):
    """Three clusters of the BASELINE configs[4] shape (8-D, n_c = 480 / 512 / 544: below, at and above a 128-row tile
    boundary), R = 2, 6144 candidates of which a quarter sit within 1e-3 of a training point (posterior variance
    cancels there: 1 + s2 - |L^-1 k*|^2 ~ 1e-5).  Long-double truth for the first 700 routed candidates per cluster."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    rng = np.random.default_rng(55)
    d, sizes = 8, (480, 512, 544)
    Xs, labs = [], []
    for c, n_c in enumerate(sizes):
        x = rng.uniform(size=(n_c, d))
        x[:, 0] = (x[:, 0] + c) / len(sizes)  # strips along x0
        Xs.append(x)
        labs.append(np.full(n_c, c))
    X, lab = np.concatenate(Xs), np.concatenate(labs)
    perm = rng.permutation(X.shape[0])
    X, lab = X[perm], lab[perm]
    y = lab % 7 + np.sin(2 * np.pi * X.sum(1) / 8) + 0.01 * rng.standard_normal(X.shape[0])
    M = 6144
    Q = rng.uniform(size=(M, d))
    near = rng.choice(M, M // 4, replace=False)
    Q[near] = np.clip(X[rng.integers(0, X.shape[0], size=len(near))] + 1e-3 * rng.standard_normal((len(near), d)), 0, 1)
    out = _fit_clusters(X, y, lab, Q, R=2, truth_cap=700)
    out.update(X=X, y=y, labels=lab.astype(np.int16), Q=Q, R=np.array([2]))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "best", out["best_score"], out["best_idx"], out["best_cluster"])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    if only:  # regenerate selected cases:  python oracle/make_golden.py case_cfg2_full case_cfg5_shape
        for fn in only:
            globals()[fn]()
        sys.exit(0)
    bukin()
    case_lml("lml_matern_n96_d2", 96, 2, "matern32", 11)
    case_lml("lml_matern_n300_d6", 300, 6, "matern32", 12)
    case_lml("lml_rbf_n150_d3", 150, 3, "rbf", 13)
    case_fit_select("fit_select_d2", n=120, d=2, C=2, R=20, M=2000, seed=21)
    case_fit_select("fit_select_d6", n=360, d=6, C=3, R=4, M=3000, seed=22)
    case_cov_mspe("cov_mspe_d3", n=150, d=3, seed=31)
    case_cfg2_full()
    case_cfg5_shape()
