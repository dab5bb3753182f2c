"""GPU parity sweep over input dimension, cluster size around the tile edges, k and kernel kind (small sizes, fixed
seeds): every template instantiation of the kernels (DMAX 4/8/16, k-NN D = 1..16) is exercised against the oracle."""
import numpy as np
import pytest
import torch

import oracle
from conftest import relerr, var_excess

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("d", [1, 2, 5, 8, 9, 16])
@pytest.mark.parametrize("kind", ["matern32", "rbf"])
def test_lml_grad_predict_sweep(d, kind):
    from cgp_b200.engine import ClusterGPEngine

    rng = np.random.default_rng(100 + d)
    sizes = [max(d + 3, 7), 128, 129, 257]
    C = len(sizes)
    n = sum(sizes)
    X = rng.uniform(size=(n, d))
    lab = np.repeat(np.arange(C), sizes)
    perm = rng.permutation(n)
    X, lab = X[perm], lab[perm]
    y = np.sin(3 * X.sum(1)) + lab + 0.05 * rng.standard_normal(n)
    eng = ClusterGPEngine(X, y, lab, kind, 1e-5)
    pc = np.arange(2 * C) % C
    thetas = rng.uniform(-1.5, 1.0, size=(2 * C, d + 1))
    thetas[:, d] = rng.uniform(-5, -1, size=2 * C)
    lml, grad, info = eng.lml_grad(pc, thetas, True)
    assert np.all(info == 0)
    for i, c in enumerate(pc):
        yn, _, _ = oracle.normalize_y(y[lab == c])
        l, g = oracle.lml_and_grad(X[lab == c], yn, thetas[i], 1e-5, kind)
        assert abs(lml[i] - l) <= 1e-8 * abs(l), (d, kind, c)
        assert np.max(np.abs(grad[i] - g)) <= 1e-7 * max(1.0, np.max(np.abs(g))), (d, kind, c)
    eng.set_theta(thetas[:C])
    Q = rng.uniform(size=(700, d))
    models = []
    for c in range(C):
        yn, ym, ys = oracle.normalize_y(y[lab == c])
        K = oracle.kernel_matrix(X[lab == c], thetas[c], kind, 1e-5)
        models.append(dict(X=X[lab == c], theta=thetas[c], L=np.linalg.cholesky(K), alpha_vec=np.linalg.solve(K, yn),
                           y_mean=ym, y_std=ys, kind=kind))
    ref = oracle.select_next(X, y, lab, models, Q, 3)
    (bs, bi, bc), (score, mean, std) = eng.ei_scan(Q, 3, want_arrays=True)
    assert np.array_equal(eng.route(Q, 3).cpu().numpy(), ref["labels"])
    assert relerr(mean.cpu().numpy(), ref["mean"]) < 1e-7
    assert var_excess(std.cpu().numpy(), ref["std"]) <= 1.0
    assert int(bi.item()) == ref["best_idx"]


@pytest.mark.parametrize("d", list(range(1, 17)))
def test_knn_all_dims(d):
    from cgp_b200 import _lib

    rng = np.random.default_rng(d)
    n, m, C = 1500, 4000, 6
    X = np.round(rng.uniform(size=(n, d)), 2)  # coarse grid -> exact ties occur
    lab = rng.integers(0, C, size=n)
    Q = np.round(rng.uniform(size=(m, d)), 2)
    for k in (1, 3, 5):
        got = _lib.knn_labels(torch.from_numpy(X).cuda(), torch.from_numpy(lab).cuda(), torch.from_numpy(Q).cuda(), k, C)
        assert np.array_equal(got.cpu().numpy(), oracle.knn_labels(X, lab, Q, k)), (d, k)


def test_lml_over_full_bounds_range():
    """Random thetas over the WHOLE optimiser box [log 1e-5, log 1e5]^P (where L-BFGS-B restarts are drawn): finite
    values agree with LAPACK; where either side reports non-PD (ill-conditioned corners) both must be far from optimal."""
    from cgp_b200.engine import ClusterGPEngine

    rng = np.random.default_rng(77)
    n, d = 200, 3
    X = rng.uniform(size=(n, d))
    y = np.sin(4 * X[:, 0]) + X[:, 1] - X[:, 2] ** 2 + 0.05 * rng.standard_normal(n)
    eng = ClusterGPEngine(X, y, np.zeros(n, dtype=np.int64), "matern32", 1e-5)
    yn, _, _ = oracle.normalize_y(y)
    thetas = oracle.restart_thetas(d + 1, 60, seed=5)
    lml, grad, info = eng.lml_grad(np.zeros(len(thetas), dtype=np.int32), thetas, True)
    n_ok = 0
    for i, th in enumerate(thetas):
        l, g = oracle.lml_and_grad(X, yn, th, 1e-5, "matern32")
        if info[i] == 0 and np.isfinite(l):
            # conditioning: cond(K) <= n (1 + s2 + alpha) / (s2 + alpha); tolerance scales with it
            s2 = np.exp(th[d])
            cond = n * (1 + s2 + 1e-5) / (s2 + 1e-5)
            tol = max(1e-8, 1e-14 * cond)
            assert abs(lml[i] - l) <= tol * abs(l), (i, th, lml[i], l)
            assert np.max(np.abs(grad[i] - g)) <= max(1e-6, 1e-12 * cond) * max(1.0, np.max(np.abs(g))), (i, th)
            n_ok += 1
        else:
            assert (info[i] != 0 and lml[i] == -np.inf) or not np.isfinite(l)
    assert n_ok >= 40
