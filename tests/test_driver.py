"""Driver surface: option keys / hooks / flags of the reference drivers on the B200 backend."""
import os

import numpy as np
import pytest


def test_driver_host_logic_random_branch(tmp_path):
    """EXPLORATION_RATE=0 never fits a surrogate, so the whole loop is host code: pilot sampling, clustering plugin,
    merge + recode, random search, output table format (REF/cGP/cGP_constrained.py:443-450)."""
    from cgp_b200.driver import cGP_constrained
    from cgp_b200.plugins import f_bukin

    class Run(cGP_constrained):
        def f_truth(self, X):
            return f_bukin.f_truth(X)

        def get_bounds(self, restrict):
            return f_bukin.get_bounds(restrict)

    out = str(tmp_path / "run")
    r = Run({"N_PILOT_CGP": 12, "N_SEQUENTIAL_CGP": 4, "RND_SEED_CGP": 7, "EXPLORATION_RATE_CGP": 0.0,
             "N_COMPONENTS_CGP": 2, "OUTPUT_NAME_CGP": out})
    X, Y = r.run()
    assert X.shape == (16, 2) and Y.shape == (16, 1)
    b = f_bukin.get_bounds(1)
    assert np.all(X >= b[:, 0]) and np.all(X <= b[:, 1])
    assert all(h["mode"] == "random" for h in r.history)
    tab = np.loadtxt(out + ".txt", delimiter=",", skiprows=1)
    assert open(out + ".txt").readline().strip() == "Y,X1,X2,label"
    assert tab.shape == (16, 4) and np.allclose(tab[:, 0], Y[:, 0], atol=1e-11) and np.allclose(tab[:, 1:3], X, atol=1e-11)
    assert os.path.exists(out + ".log")
    # same seed -> same trajectory
    X2, _ = Run({"N_PILOT_CGP": 12, "N_SEQUENTIAL_CGP": 4, "RND_SEED_CGP": 7, "EXPLORATION_RATE_CGP": 0.0,
                 "N_COMPONENTS_CGP": 2}).run()
    assert np.array_equal(X, X2)


def test_driver_rejects_unsupported():
    from cgp_b200.driver import cGP_constrained

    with pytest.raises(NotImplementedError):
        cGP_constrained({"ACQUISITION_CGP": "UCB"})
    with pytest.raises(NotImplementedError):
        cGP_constrained({"METHOD_CGP": "BAYESIAN"})


@pytest.mark.gpu
def test_driver_cli_end_to_end(tmp_path, monkeypatch):
    """`cGP.py -p 20 -s 3 -c 3 -n 3` shape (BASELINE config 1) through the GPU backend, with -C/-N plugins and -d."""
    import dill

    from cgp_b200 import driver

    monkeypatch.chdir(tmp_path)
    X, Y = driver.main(["-p", "20", "-s", "3", "-C", "KMeans2", "-N", "classify", "-f", "cgp_b200.plugins.f_bukin",
                        "-r", "123", "-o", "bukin_run", "-d", "bukin_model", "-X", "--candidates", "4096", "--restarts", "2"])
    assert X.shape == (23, 2) and Y.shape == (23, 1)
    # the exhaustive scan walks the same trajectory as the (default) exact pruned scan (same seed -> same candidates -> same selections)
    Xp, Yp = driver.main(["-p", "20", "-s", "3", "-C", "KMeans2", "-N", "classify", "-f", "cgp_b200.plugins.f_bukin",
                          "-r", "123", "-o", "bukin_run_p", "-X", "--candidates", "4096", "--restarts", "2", "--exhaustive-scan"])
    assert np.array_equal(X, Xp) and np.array_equal(Y, Yp)
    assert os.path.exists("bukin_run.txt") and os.path.exists("bukin_run.log") and os.path.exists("bukin_model.pkl")
    with open("bukin_model.pkl", "rb") as f:
        obj = dill.load(f)
    assert set(obj) == {"Cluster", "Classify", "GPR_list"}
    # the pickled objects predict like REF/cGP/f_truth_dill.py:20-30
    x = X[:5]
    lab = obj["Classify"].predict(x)
    mu = [obj["GPR_list"][int(l)].predict(x[i:i + 1], return_std=True)[0][0] for i, l in enumerate(lab)]
    assert np.all(np.isfinite(mu))


@pytest.mark.gpu
def test_driver_mspe_acquisition():
    """ACQUISITION_CGP='MSPE' (-A mean_square_prediction_error): the reference's second acquisition as a dense scan."""
    from cgp_b200.driver import cGP_constrained
    from cgp_b200.plugins import f_bukin

    class Run(cGP_constrained):
        def f_truth(self, X):
            return f_bukin.f_truth(X)

        def get_bounds(self, restrict):
            return f_bukin.get_bounds(restrict)

    r = Run({"N_PILOT_CGP": 18, "N_SEQUENTIAL_CGP": 2, "RND_SEED_CGP": 5, "NO_CLUSTER_CGP": True, "N_CANDIDATES_CGP": 1024,
             "N_RESTARTS_CGP": 1, "ACQUISITION_CGP": "MSPE"})
    X, Y = r.run()
    assert X.shape == (20, 2) and all(h["mode"] == "acquisition" and np.isfinite(h["score"]) for h in r.history)
    b = f_bukin.get_bounds(1)
    assert np.all(X >= b[:, 0]) and np.all(X <= b[:, 1])


def test_constraint_objects_become_a_candidate_mask():
    """The scipy constraint objects a reference -f module exports (REF/cGP/f_truth.py:44-59) and the class hooks
    (REF/cGP/cGP_constrained_module.py:202-222) filter the dense candidate set: lb <= A q <= ub, lb <= fun(q) <= ub,
    Bounds — vectorised fun, point-wise-only fun, and vector-valued fun."""
    from scipy.optimize import LinearConstraint, NonlinearConstraint

    from cgp_b200.driver import cGP_constrained

    rng = np.random.default_rng(0)
    Q = rng.uniform(-2, 2, size=(5000, 2))
    bounds = np.array([[-1.5, 1.5], [-2.0, 1.0]])

    class A(cGP_constrained):
        def get_linear_constraint(self):
            return LinearConstraint([[1, 2], [2, 1]], [-np.inf, -1.0], [1.0, np.inf])

        def get_nonlinear_constraint(self):
            return NonlinearConstraint(lambda x: x[0] ** 2 + x[1], -np.inf, 1.0)

    want = ((Q[:, 0] + 2 * Q[:, 1] <= 1) & (2 * Q[:, 0] + Q[:, 1] >= -1) & (Q[:, 0] ** 2 + Q[:, 1] <= 1)
            & (np.abs(Q[:, 0]) <= 1.5) & (Q[:, 1] >= -2) & (Q[:, 1] <= 1))
    assert np.array_equal(A({}).constraint_mask(Q, bounds), want)

    calls = [0]

    def pointwise_only(x):  # breaks on a [d, M] argument
        calls[0] += 1
        assert np.ndim(x) == 1
        return float(x[0] * x[1])

    class B(cGP_constrained):
        def get_nonlinear_constraint(self):
            return NonlinearConstraint(pointwise_only, -0.5, 0.5)

    got = B({}).constraint_mask(Q[:300], np.array([[-9.0, 9.0], [-9.0, 9.0]]))
    assert np.array_equal(got, np.abs(Q[:300, 0] * Q[:300, 1]) <= 0.5) and calls[0] >= 300

    class Cc(cGP_constrained):
        def get_nonlinear_constraint(self):  # vector-valued, the reference's documented example
            return NonlinearConstraint(lambda x: [x[0] ** 2 + x[1], x[0] ** 2 - x[1]], -np.inf, [1.0, 1.0])

    got = Cc({}).constraint_mask(Q, np.array([[-9.0, 9.0], [-9.0, 9.0]]))
    assert np.array_equal(got, (Q[:, 0] ** 2 + Q[:, 1] <= 1) & (Q[:, 0] ** 2 - Q[:, 1] <= 1))
    # the reference's placeholder constraints (cons_f = 0 within (-inf, 9), all-infinite linear bounds) keep everything
    from cgp_b200.plugins import f_bukin

    class D(cGP_constrained):
        def get_linear_constraint(self):
            return getattr(f_bukin, "linear_constraint", None)

        def get_nonlinear_constraint(self):
            return getattr(f_bukin, "nonlinear_constraint", None)

    b = f_bukin.get_bounds(1)
    Qb = rng.uniform(b[:, 0], b[:, 1], size=(100, 2))
    assert np.all(D({}).constraint_mask(Qb, b))


@pytest.mark.gpu
def test_boundary_penalty_inside_the_gpu_score():
    """(EI + boundary_penalty(x, X_c)) / n_c on the GPU (REF/cGP/acquisition.py:29-32) == the same expression evaluated
    on the host from the per-candidate EI arrays; pruned and exhaustive scans agree; a point-wise-only hook works."""
    import torch

    from cgp_b200 import CGPRegressor

    rng = np.random.default_rng(3)
    n, d = 500, 2
    X = rng.uniform(size=(n, d))
    lab = (X[:, 0] > 0.5).astype(np.int64)
    y = lab + np.sin(5 * X.sum(1)) + 0.01 * rng.standard_normal(n)
    Q = rng.uniform(size=(20000, d))
    m = CGPRegressor("matern32", alpha=1e-5, n_restarts_optimizer=1, random_state=0).fit(X, y, lab)
    seen = []

    def pen(Xq, data_X):  # vectorised: pull towards the centroid of the cluster's samples, strong enough to matter
        seen.append(data_X.shape[0])
        return -0.5 * np.sum((np.atleast_2d(Xq) - data_X.mean(0)) ** 2, axis=1)

    out = m.score_candidates(Q)
    counts = np.bincount(lab)
    host = np.empty(Q.shape[0])
    for c in range(2):
        sel = out["labels"] == c
        host[sel] = (out["score"][sel] * counts[c] + pen(Q[sel], X[lab == c])) / counts[c]
    want = int(np.lexsort((np.arange(Q.shape[0]), out["labels"], -host))[0])
    x1, s1, i1 = m.select_next(Q, boundary_penalty=pen)
    x2, s2, i2 = m.select_next(Q, boundary_penalty=pen, prune=False)
    assert i1 == i2 == want and s1 == s2 and np.array_equal(x1, Q[want])
    assert abs(s1 - host[want]) <= 1e-12 * abs(host[want])
    assert i1 != m.select_next(Q)[2]  # the penalty really changed the choice
    assert set(seen) == set(counts.tolist())  # the hook saw each cluster's own samples (data_X = X_c)

    def pen_point(Xq, data_X):
        assert Xq.shape == (1, d)
        return float(-0.5 * np.sum((Xq[0] - data_X.mean(0)) ** 2))

    Qs = Q[:1500]
    a = m.select_next(Qs, boundary_penalty=pen)
    b = m.select_next(Qs, boundary_penalty=pen_point)
    assert a[2] == b[2] and a[1] == b[1]
    assert m.select_next(Qs, boundary_penalty=lambda a_, b_: 0)[2] == m.select_next(Qs)[2]
    torch.cuda.synchronize()


@pytest.mark.gpu
def test_driver_honours_reference_constraints(tmp_path, monkeypatch):
    """A -f module with REAL constraint objects: every sequential sample satisfies them (the reference hands the same
    objects to trust-constr, REF/cGP/cGP_constrained.py:362-366)."""
    from cgp_b200 import driver

    mod = tmp_path / "f_constrained.py"
    mod.write_text(
        "import numpy as np\n"
        "from scipy.optimize import LinearConstraint, NonlinearConstraint\n"
        "EXAMPLE_NAME = 'constrained'\n"
        "def f_truth(X):\n    X = X.reshape(1, -1)\n    return -(X[:, 0] - 0.3) ** 2 - (X[:, 1] + 0.2) ** 2\n"
        "def get_bounds(restrict):\n    return np.array([[-1, 1], [-1, 1]]).astype(float)\n"
        "def boundary_penalty(X, data_X=None):\n    return 0\n"
        "def censor_function(Y):\n    return Y\n"
        "linear_constraint = LinearConstraint([[1, 1]], [-np.inf], [0.2])\n"
        "nonlinear_constraint = NonlinearConstraint(lambda x: x[0] ** 2 + x[1] ** 2, 0.04, np.inf)\n")
    monkeypatch.chdir(tmp_path)
    X, Y = driver.main(["-p", "15", "-s", "4", "-g", "1", "-f", "f_constrained", "-r", "3", "-X", "--candidates", "4096",
                        "--restarts", "1"])
    new = X[15:]
    assert new.shape == (4, 2)
    assert np.all(new.sum(1) <= 0.2 + 1e-12) and np.all((new ** 2).sum(1) >= 0.04 - 1e-12)
