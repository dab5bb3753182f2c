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
