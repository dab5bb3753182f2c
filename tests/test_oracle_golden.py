"""CPU: pin the NumPy oracle against the reference's fixtures (BUKIN.pkl) and against outputs of the
reference's own scikit-learn code path committed under tests/golden/ (oracle/make_golden.py)."""
import numpy as np
import pytest

import oracle
from conftest import load_golden, relerr

RTOL = 1e-8  # north star tolerance for LML / mean / variance


def test_bukin_known_answers():
    g = load_golden("bukin")
    for i in range(2):
        X, y, th = g[f"g{i}_X"], g[f"g{i}_y"], g[f"g{i}_theta"]
        lml = oracle.lml_and_grad(X, y, th, alpha=1e-5, want_grad=False)
        assert abs(lml - g[f"g{i}_lml"][0]) <= 1e-12 * abs(lml)
        K = oracle.kernel_matrix(X, th, "matern32", alpha=1e-5)
        L = np.linalg.cholesky(K)
        assert np.max(np.abs(L - g[f"g{i}_L"])) < 1e-12
        a = np.linalg.solve(K, y)
        assert relerr(a, g[f"g{i}_alpha"]) < 1e-9
    # k-NN fixture: every training point with k=1 votes for itself
    lab = oracle.knn_labels(g["knn_X"], g["knn_y"], g["knn_X"], k=1)
    assert np.array_equal(lab, g["knn_y"])


@pytest.mark.parametrize("name", ["lml_matern_n96_d2", "lml_matern_n300_d6", "lml_rbf_n150_d3"])
def test_lml_and_grad_vs_sklearn(name):
    g = load_golden(name)
    kind = str(g["kind"][0])
    K = oracle.kernel_matrix(g["X"], g["thetas"][1], kind)
    assert np.max(np.abs(K - g["K1"])) < 1e-14
    for th, lml_ref, grad_ref in zip(g["thetas"], g["lml"], g["grad"]):
        lml, grad = oracle.lml_and_grad(g["X"], g["y_norm"], th, 1e-5, kind)
        assert abs(lml - lml_ref) <= RTOL * abs(lml_ref)
        assert np.max(np.abs(grad - grad_ref)) <= 1e-7 * max(1.0, np.max(np.abs(grad_ref)))
    yn, _, _ = oracle.normalize_y(g["y"])
    assert np.max(np.abs(yn - g["y_norm"])) < 1e-14


@pytest.mark.parametrize("name,C", [("fit_select_d2", 2), ("fit_select_d6", 3)])
def test_fit_predict_select_vs_reference(name, C):
    g = load_golden(name)
    X, y, labels, Q = g["X"], g["y"], g["labels"], g["Q"]
    kind = str(g["kind"][0])
    R = int(g["R"][0])
    # router: bit-exact labels vs sklearn's kd_tree classifier
    lab = oracle.knn_labels(X, labels, Q, k=3)
    assert np.array_equal(lab, g["qlab"])
    models = []
    for c in range(C):
        m = oracle.fit_gpr(X[labels == c], y[labels == c], R, 1e-5, kind)
        assert abs(m["lml"] - g[f"c{c}_lml"][0]) <= RTOL * abs(g[f"c{c}_lml"][0])
        models.append(m)
    sel = oracle.select_next(X, y, labels, models, Q, k=3)
    ok = g["std"] > 1e-6
    assert relerr(sel["mean"], g["mean"]) < 1e-6
    assert relerr(sel["std"][ok], g["std"][ok]) < 1e-6
    assert sel["best_idx"] == int(g["best_idx"][0])
    assert sel["best_cluster"] == int(g["best_cluster"][0])
    # reference's own pointwise expected_improvement on a subset
    assert np.max(np.abs(sel["score"][g["pw_idx"]] - g["pw_val"])) < 1e-9


def test_predict_at_reference_theta():
    """With theta fixed to sklearn's optimum, mean/std agree to the north-star tolerance."""
    g = load_golden("fit_select_d2")
    X, y, labels, Q = g["X"], g["y"], g["labels"], g["Q"]
    for c in range(2):
        Xt, yt = X[labels == c], y[labels == c]
        yn, ym, ys = oracle.normalize_y(yt)
        th = g[f"c{c}_theta"]
        K = oracle.kernel_matrix(Xt, th, "matern32", 1e-5)
        L = np.linalg.cholesky(K)
        a = np.linalg.solve(K, yn)
        sel = g["qlab"] == c
        mu, sd = oracle.predict_mean_std(Xt, th, L, a, ym, ys, Q[sel])
        assert relerr(mu, g["mean"][sel]) < RTOL
        big = g["std"][sel] > 1e-3
        assert relerr(sd[big], g["std"][sel][big]) < 1e-7


def test_merge_and_recode():
    lab = np.array([3, 3, 3, 3, 3, 3, 7, 7, 1, 1, 1, 1, 1, 1, 1])
    m = oracle.merge_small_clusters(lab, d=2)  # cluster 7 has 2 <= 4 samples -> merged into largest (1)
    assert set(np.unique(m)) == {1, 3}
    r = oracle.recode(m)
    assert set(np.unique(r)) == {0, 1} and r[0] == 1 and r[-1] == 0
