import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def var_excess(std, std_ref, rtol=1e-8, cancel=1e-13):
    """Posterior-variance parity (north star: 1e-8 relative).  var = prior - |L^-1 k*|^2 is a difference of O(prior)
    numbers, so no fp64 implementation can hold 1e-8 RELATIVE below ~1e-5 of the prior; the bound is
    |var - var_ref| <= rtol * var_ref + cancel * max(var_ref)   (cancel = 1e-13 ~ 500 ulp of the quantity cancelled;
    max(var_ref) stands for the prior variance).  Returns the worst ratio error / bound (<= 1 passes).  No std floor."""
    v, r = np.asarray(std, dtype=np.float64) ** 2, np.asarray(std_ref, dtype=np.float64) ** 2
    if v.size == 0:
        return 0.0
    return float(np.max(np.abs(v - r) / (rtol * r + cancel * np.max(r) + 1e-300)))


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0
