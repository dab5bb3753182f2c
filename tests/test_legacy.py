"""Legacy model files of the reference (sklearn 0.24 pickles): parsing on the CPU; the reference tree only exists in the
build container, so these tests skip on the GPU box."""
import os

import numpy as np
import pytest

from conftest import load_golden

BUKIN = "/root/reference/cGP/BUKIN.pkl"


@pytest.mark.skipif(not os.path.exists(BUKIN), reason="reference tree not present")
def test_parse_bukin_pickle():
    from cgp_b200 import legacy

    arr = legacy.extract_arrays(legacy._permissive_load(BUKIN))
    g = load_golden("bukin")
    assert len(arr["gprs"]) == 2
    for i, a in enumerate(arr["gprs"]):
        assert np.array_equal(a["X"], g[f"g{i}_X"]) and np.array_equal(a["theta"], g[f"g{i}_theta"])
        assert abs(a["lml"] - g[f"g{i}_lml"][0]) < 1e-12 and a["alpha"] == 1e-5
    assert arr["classify"]["k"] == 3 and arr["classify"]["X"].shape == (19, 2)


def _bukin_arrays():
    """The arrays legacy.extract_arrays returns for BUKIN.pkl, from the committed golden (tests/golden/bukin.npz holds
    exactly those arrays, written by oracle/make_golden.py): the GPU box has no /root/reference."""
    g = load_golden("bukin")
    gprs = [dict(X=g[f"g{i}_X"], y_norm=g[f"g{i}_y"], y_mean=float(g[f"g{i}_ymean"][0]), y_std=float(g[f"g{i}_ystd"][0]),
                 theta=g[f"g{i}_theta"], alpha=1e-5, kind="matern32") for i in range(2)]
    clf = dict(X=g["knn_X"], labels=g["knn_classes"][g["knn_y"]], k=int(g["knn_k"][0]))
    return g, dict(gprs=gprs, classify=clf)


@pytest.mark.gpu
def test_bukin_predict_like_f_truth_dill():
    """REF/cGP/f_truth_dill.py:17-30 on the legacy BUKIN model: k-NN label, then that cluster's predict, against the
    values scikit-learn returns for the same arrays (incl. the x = [-10, 1] and the reference's own [1, 1] call)."""
    from cgp_b200 import legacy

    g, arr = _bukin_arrays()
    m = legacy.model_from_arrays(arr)
    assert np.array_equal(m["Classify"].predict(g["ft_x"]), g["ft_label"])
    for j, x in enumerate(g["ft_x"]):
        lab = int(m["Classify"].predict(x.reshape(1, -1))[0])
        mu, sd = m["GPR_list"][lab].predict(x.reshape(1, -1), return_std=True)
        mu = float(np.asarray(mu).reshape(-1)[0])
        assert abs(mu - g["ft_mean"][j]) <= 1e-8 * max(abs(g["ft_mean"][j]), 1e-3), (j, mu, g["ft_mean"][j])
        assert abs(sd[0] ** 2 - g["ft_std"][j] ** 2) <= 1e-8 * g["ft_std"][j] ** 2 + 1e-18, (j, sd[0], g["ft_std"][j])
        assert legacy.f_truth(m, x) == mu
    for i in range(2):  # the factor the pickle stores
        L = m["GPR_list"][i].L_
        assert np.max(np.abs(L - g[f"g{i}_L"])) < 1e-10


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(BUKIN), reason="reference tree not present")
def test_bukin_pickle_loader_on_gpu():
    from cgp_b200 import legacy

    m = legacy.load_reference_model(BUKIN)
    g = load_golden("bukin")
    assert abs(legacy.f_truth(m, g["ft_x"][0]) - g["ft_mean"][0]) <= 1e-8 * abs(g["ft_mean"][0])
