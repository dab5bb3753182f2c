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


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(BUKIN), reason="reference tree not present")
def test_bukin_predict_like_f_truth_dill():
    from cgp_b200 import legacy

    m = legacy.load_reference_model(BUKIN)
    x = np.array([[-10.0, 1.0]])
    lab = int(m["Classify"].predict(x)[0])
    mu, sd = m["GPR_list"][lab].predict(x, return_std=True)
    assert np.isfinite(mu[0]) and sd[0] >= 0
