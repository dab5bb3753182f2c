"""Incremental refit: appending one sample to a factorised cluster at unchanged theta must equal
factorising the extended data from scratch (which the other GPU tests pin against the oracle / scikit-learn), and must
agree with the oracle directly.  The reference has no such path (it refits everything, REF/cGP/cGP_constrained.py:276,
315-327), so the parity target is "same numbers as its from-scratch refit at the same theta"."""
import numpy as np
import pytest

import oracle
from conftest import relerr, var_excess

pytestmark = pytest.mark.gpu


def _data(n, d, seed):
    rng = np.random.default_rng(seed)
    X = rng.uniform(size=(n, d))
    lab = (X[:, 0] > 0.55).astype(np.int64)
    y = lab + np.sin(5 * X.sum(1)) + 0.01 * rng.standard_normal(n)
    return X, y, lab, rng


@pytest.mark.parametrize("kind", ["matern32", "rbf"])
def test_append_equals_refactorisation_and_oracle(kind):
    from cgp_b200.engine import ClusterGPEngine

    X, y, lab, rng = _data(400, 3, 5)
    # make cluster 0 hold exactly 128 samples so the first append crosses the 128-row padding boundary
    idx0 = np.nonzero(lab == 0)[0][:128]
    keep = np.concatenate([idx0, np.nonzero(lab == 1)[0]])
    keep.sort()
    X, y, lab = X[keep], y[keep], lab[keep]
    assert np.sum(lab == 0) == 128
    th = np.array([[-0.9, -0.4, 0.2, -5.0], [-0.2, -1.1, 0.4, -3.0]])
    eng = ClusterGPEngine(X, y, lab, kind, 1e-5).set_theta(th)
    new = [(np.array([0.2, 0.3, 0.7]), 0.4, 0), (np.array([0.9, 0.1, 0.5]), 1.7, 1), (np.array([0.31, 0.8, 0.15]), -0.3, 0),
           (np.array([0.21, 0.3, 0.7]), 2.5, 0)]  # the last one changes the incumbent and sits close to the first
    for x, yv, c in new:
        eng.append(x, yv, c)
        X, y, lab = np.vstack([X, x]), np.append(y, yv), np.append(lab, c)
    ref = ClusterGPEngine(X, y, lab, kind, 1e-5).set_theta(th)
    Q = rng.uniform(size=(3000, 3))
    for c in range(2):
        assert eng.counts[c] == ref.counts[c]
        assert relerr(eng.L_of(c), ref.L_of(c)) < 1e-9 or np.max(np.abs(eng.L_of(c) - ref.L_of(c))) < 1e-11
        assert np.max(np.abs(eng.W_of(c) - ref.W_of(c))) <= 1e-9 * np.max(np.abs(ref.W_of(c)))
        assert np.max(np.abs(eng.alpha_of(c) - ref.alpha_of(c))) <= 1e-8 * np.max(np.abs(ref.alpha_of(c)))
        assert abs(eng.model["lml_host"][c] - ref.model["lml_host"][c]) <= 1e-10 * abs(ref.model["lml_host"][c])
        # oracle (NumPy/LAPACK restatement of SK/_gpr.py:584-617) on the extended cluster
        Xt, yt = X[lab == c], y[lab == c]
        yn, ym, ys = oracle.normalize_y(yt)
        lml = oracle.lml_and_grad(Xt, yn, th[c], 1e-5, kind=kind, want_grad=False)
        assert abs(eng.model["lml_host"][c] - lml) <= 1e-8 * abs(lml)
        assert abs(eng.y_mean[c] - ym) < 1e-14 and abs(eng.y_std[c] - ys) < 1e-14 and eng.y_max[c] == yt.max()
    assert np.array_equal(eng.route(Q, 3).cpu().numpy(), ref.route(Q, 3).cpu().numpy())
    (bs, bi, bc), (sc, mu, sd) = eng.ei_scan(Q, 3, want_arrays=True)
    (rs, ri, rc), (rsc, rmu, rsd) = ref.ei_scan(Q, 3, want_arrays=True)
    assert relerr(mu.cpu().numpy(), rmu.cpu().numpy()) < 1e-8
    assert var_excess(sd.cpu().numpy(), rsd.cpu().numpy()) <= 1.0
    assert int(bi.item()) == int(ri.item()) and int(bc.item()) == int(rc.item())


def test_append_to_tiny_cluster_and_not_pd():
    from cgp_b200.engine import ClusterGPEngine

    rng = np.random.default_rng(3)
    X = rng.uniform(size=(9, 2))
    y = np.sin(3 * X[:, 0]) + X[:, 1]
    lab = np.array([0, 0, 0, 0, 0, 1, 1, 1, 1])
    th = np.array([[0.1, -0.3, -9.0], [0.0, 0.0, -2.0]])
    eng = ClusterGPEngine(X, y, lab, "matern32", 1e-5).set_theta(th)
    eng.append([0.5, 0.5], 0.9, 1)
    X2, y2, lab2 = np.vstack([X, [0.5, 0.5]]), np.append(y, 0.9), np.append(lab, 1)
    ref = ClusterGPEngine(X2, y2, lab2, "matern32", 1e-5).set_theta(th)
    assert np.max(np.abs(eng.L_of(1) - ref.L_of(1))) < 1e-12
    assert np.max(np.abs(eng.alpha_of(1) - ref.alpha_of(1))) <= 1e-9 * np.max(np.abs(ref.alpha_of(1)))
    # exact duplicates with no jitter (alpha = 0, noise 1e-30): the extended matrix is singular, the new pivot is
    # rounding noise of either sign -> info > 0 -> LinAlgError with scikit-learn's wording (SK/_gpr.py:353-361)
    eng0 = ClusterGPEngine(X, y, lab, "matern32", 0.0).set_theta(np.array([[0.0, 0.0, np.log(1e-30)], [0.0, 0.0, -2.0]]))
    with pytest.raises(np.linalg.LinAlgError):
        for _ in range(40):
            eng0.append(X[0], y[0], 0)
    with pytest.raises(ValueError):
        eng.append([0.1, 0.1], 0.0, 5)


def test_partial_fit_reoptimise_and_driver_incremental():
    from cgp_b200 import CGPRegressor
    from cgp_b200.driver import cGP_constrained
    from cgp_b200.plugins import f_bukin

    X, y, lab, rng = _data(260, 2, 11)
    m = CGPRegressor("matern32", alpha=1e-5, n_restarts_optimizer=1, random_state=0).fit(X, y, lab)
    th_old = m.theta_.copy()
    x_new, y_new = np.array([0.3, 0.6]), 0.25
    m.partial_fit(x_new, y_new, 0, reoptimize=True)
    X2, y2, lab2 = np.vstack([X, x_new]), np.append(y, y_new), np.append(lab, 0)
    assert m.engine_.n == 261 and m.classify_.predict(x_new[None, :])[0] in (0, 1)
    for c in range(2):  # warm-started optimum is at least as good as the old theta on the extended data
        Xt, yt = X2[lab2 == c], y2[lab2 == c]
        yn, _, _ = oracle.normalize_y(yt)
        at_old = oracle.lml_and_grad(Xt, yn, th_old[c], 1e-5, want_grad=False)
        at_new = oracle.lml_and_grad(Xt, yn, m.theta_[c], 1e-5, want_grad=False)
        assert at_new >= at_old - 1e-9 * abs(at_old)
        assert abs(m.log_marginal_likelihood_value_[c] - at_new) <= 1e-8 * abs(at_new)

    class Run(cGP_constrained):
        def f_truth(self, X):
            return f_bukin.f_truth(X)

        def get_bounds(self, restrict):
            return f_bukin.get_bounds(restrict)

    r = Run({"N_PILOT_CGP": 16, "N_SEQUENTIAL_CGP": 4, "RND_SEED_CGP": 3, "NO_CLUSTER_CGP": True, "N_CANDIDATES_CGP": 2048,
             "N_RESTARTS_CGP": 1, "INCREMENTAL_CGP": True})
    Xr, Yr = r.run()
    assert Xr.shape == (20, 2)
    assert [h["mode"] for h in r.history] == ["acquisition"] + ["acquisition+incremental"] * 3
    assert r.model_.engine_.n == 19  # the model holds every sample drawn before the last acquisition
