"""CPU: host-side logic — lock-step L-BFGS-B equals scipy.optimize.minimize bit for bit, restart draws equal
scikit-learn's, kernel parsing, plugin contracts."""
import numpy as np
import pytest
import scipy.optimize

import oracle
from cgp_b200.engine import LOG_HI, LOG_LO, restart_thetas
from cgp_b200.lbfgs import minimize_lockstep


def test_lockstep_equals_scipy():
    rng = np.random.default_rng(0)
    X = rng.uniform(size=(60, 3))
    y, _, _ = oracle.normalize_y(np.sin(4 * X[:, 0]) + X[:, 1] ** 2 + 0.05 * rng.standard_normal(60))
    starts = restart_thetas(4, 5)
    lo, hi = np.full(4, LOG_LO), np.full(4, LOG_HI)

    def evalb(idx, Xs):
        out = [oracle.lml_and_grad(X, y, th) for th in Xs]
        return np.array([-o[0] for o in out]), np.array([-o[1] for o in out])

    xs, fs, nfev, nit, steps = minimize_lockstep(evalb, starts, lo, hi)
    for i, t0 in enumerate(starts):
        res = scipy.optimize.minimize(lambda th: tuple(-v for v in oracle.lml_and_grad(X, y, th)), t0,
                                      method="L-BFGS-B", jac=True, bounds=list(zip(lo, hi)))
        assert np.array_equal(res.x, xs[i]) and res.fun == fs[i] and res.nit == nit[i]
    assert steps == nfev.max()


def test_lockstep_handles_inf():
    """Non-PD evaluations return (+inf, 0) like SK/_gpr.py:592-593; the state machine must terminate."""
    def evalb(idx, Xs):
        f = np.where(Xs[:, 0] > 1.0, np.inf, (Xs ** 2).sum(1))
        g = np.where((Xs[:, 0] > 1.0)[:, None], 0.0, 2 * Xs)
        return f, g

    xs, fs, *_ = minimize_lockstep(evalb, [np.array([0.9, 0.5]), np.array([2.0, 2.0])], np.full(2, -5.0), np.full(2, 5.0))
    assert np.all(np.isfinite(xs))
    assert abs(fs[0]) < 1e-8


def test_restart_draws_match_sklearn():
    from sklearn.utils import check_random_state

    rng = check_random_state(0)
    b = np.tile([LOG_LO, LOG_HI], (7, 1))
    want = [rng.uniform(b[:, 0], b[:, 1]) for _ in range(4)]
    got = restart_thetas(7, 4, 0)
    assert np.array_equal(got[0], np.zeros(7))
    assert np.array_equal(got[1:], np.array(want))
    assert np.array_equal(got, oracle.restart_thetas(7, 4, 0))


def test_parse_kernel():
    from sklearn.gaussian_process.kernels import RBF, Matern, WhiteKernel

    from cgp_b200.gpr import parse_kernel

    k = Matern(length_scale=np.ones(3), length_scale_bounds=(1e-05, 100000.0), nu=3 / 2) + \
        WhiteKernel(noise_level=1.0, noise_level_bounds=(1e-05, 100000.0))
    kind, th0, b, tmpl = parse_kernel(k, 3)
    assert kind == "matern32" and np.array_equal(th0, np.zeros(4)) and tmpl is k
    assert np.allclose(b, np.tile([LOG_LO, LOG_HI], (4, 1)))
    kind, *_ = parse_kernel(RBF(np.ones(2), (1e-5, 1e5)) + WhiteKernel(1.0, (1e-5, 1e5)), 2)
    assert kind == "rbf"
    import pytest

    with pytest.raises(NotImplementedError):
        parse_kernel(Matern(np.ones(2), nu=2.5) + WhiteKernel(), 2)
    with pytest.raises(NotImplementedError):
        parse_kernel(RBF(np.ones(2)), 2)


def test_plugin_contracts():
    import importlib

    for name, k in [("cluster", 3), ("DGM2", 2), ("DGM5", 5)]:
        obj = importlib.import_module(f"cgp_b200.plugins.{name}").f_Cluster(123)
        assert obj.n_components == k and obj.random_state == 123 and hasattr(obj, "fit_predict")
        assert obj.weight_concentration_prior_type == "dirichlet_process"
    for k in range(1, 6):
        obj = importlib.import_module(f"cgp_b200.plugins.KMeans{k}").f_Cluster(7)
        assert obj.n_clusters == k and obj.random_state == 7
    clf = importlib.import_module("cgp_b200.plugins.classify").f_Classify()
    assert clf.n_neighbors == 3 and hasattr(clf, "fit") and hasattr(clf, "predict")
    XY = np.random.default_rng(0).normal(size=(40, 3))
    lab = importlib.import_module("cgp_b200.plugins.KMeans2").f_Cluster(0).fit_predict(XY)
    assert lab.shape == (40,)


def test_pipelined_lockstep_equals_plain():
    """Two alternating groups (host/device overlap) give every run the same evaluation sequence as one group."""
    rng = np.random.default_rng(1)
    n_runs, dim = 40, 3
    A = rng.uniform(0.5, 2.0, size=(n_runs, dim))
    c = rng.uniform(-1, 1, size=(n_runs, dim))

    def fg(idx, Xs):
        a, cc = A[idx], c[idx]
        r = Xs - cc
        f = (a * r ** 2).sum(1) + 0.1 * np.sin(3 * Xs).sum(1)
        g = 2 * a * r + 0.3 * np.cos(3 * Xs)
        return f, g

    x0 = [rng.uniform(-2, 2, size=dim) for _ in range(n_runs)]
    lo, hi = np.full(dim, -3.0), np.full(dim, 3.0)
    plain = minimize_lockstep(fg, x0, lo, hi)
    tickets, counter = {}, [0]

    def submit(idx, Xs):
        counter[0] += 1
        tickets[counter[0]] = fg(idx, Xs)
        return counter[0]

    piped = minimize_lockstep(fg, x0, lo, hi, submit=submit, collect=lambda t: tickets.pop(t), pipeline_min=8)
    assert np.array_equal(plain[0], piped[0]) and np.array_equal(plain[1], piped[1])
    assert np.array_equal(plain[2], piped[2]) and np.array_equal(plain[3], piped[3])
    assert not tickets


def _toy_problem(n_runs=23, dim=4, seed=5):
    rng = np.random.default_rng(seed)
    A = rng.uniform(0.5, 2.0, size=(n_runs, dim))
    c = rng.uniform(-1, 1, size=(n_runs, dim))

    def fg(idx, Xs):
        r = Xs - c[idx]
        return (A[idx] * r ** 2).sum(1) + 0.1 * np.sin(3 * Xs).sum(1), 2 * A[idx] * r + 0.3 * np.cos(3 * Xs)

    x0 = [rng.uniform(-2, 2, size=dim) for _ in range(n_runs)]
    return fg, x0, np.full(dim, -3.0), np.full(dim, 3.0)


@pytest.mark.parametrize("n_lanes,policy", [(1, None), (2, None), (3, None), (2, lambda a: 2 if a <= 6 or a > 15 else 1)])
def test_async_lanes_equal_lockstep(n_lanes, policy):
    """minimize_async (independent evaluation lanes, no lock-step) gives every run exactly the result of the lock-step
    driver / scipy.optimize.minimize: a run's evaluation sequence does not depend on the batching."""
    import time

    from cgp_b200.lbfgs import minimize_async

    fg, x0, lo, hi = _toy_problem()
    plain = minimize_lockstep(fg, x0, lo, hi)
    in_flight = [0]
    max_in_flight = [0]

    def submit(lane, gid, X):
        in_flight[0] += 1
        max_in_flight[0] = max(max_in_flight[0], in_flight[0])
        return (time.perf_counter() + 2e-4 * (1 + lane), fg(gid, X))

    def collect(t):
        in_flight[0] -= 1
        return t[1]

    fin, st = minimize_async(x0, np.arange(len(x0)), lo, hi, submit, lambda t: time.perf_counter() >= t[0], collect,
                             n_lanes=n_lanes, lane_policy=policy)
    assert sorted(fin) == list(range(len(x0)))
    for i in range(len(x0)):
        x, f, nfev, nit = fin[i]
        assert np.array_equal(x, plain[0][i]) and f == plain[1][i] and nfev == plain[2][i] and nit == plain[3][i]
    assert st["evals"] == plain[2].sum() and max_in_flight[0] <= n_lanes
    if n_lanes > 1 and policy is None:
        assert max_in_flight[0] == n_lanes


def test_async_max_evals_keeps_best_evaluated_point():
    from cgp_b200.lbfgs import minimize_async

    fg, x0, lo, hi = _toy_problem(5)
    fin, st = minimize_async(x0, np.arange(5), lo, hi, lambda lane, gid, X: fg(gid, X), lambda t: True, lambda t: t,
                             n_lanes=2, max_evals=1)
    f0, _ = fg(np.arange(5), np.stack(x0))
    for i in range(5):
        assert np.array_equal(fin[i][0], x0[i]) and fin[i][1] == f0[i] and fin[i][2] == 1
    assert st["evals"] == 5
