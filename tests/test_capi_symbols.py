"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/cgp_b200.h declares
(no compute calls — there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge

    ge.build()
    return ge.SO


def _declared():
    src = open(os.path.join(ROOT, "include", "cgp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cgp_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built):
    names = _declared()
    assert len(names) >= 15
    lib = ctypes.CDLL(built)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cgp_b200.h but not exported"


def test_binding_covers_header(built):
    from cgp_b200 import _lib

    assert sorted(_lib.SIGNATURES) == _declared()
    L = _lib.lib()
    assert L.cgp_version() >= 100
    assert L.cgp_padded_n(1) == 128 and L.cgp_padded_n(128) == 128 and L.cgp_padded_n(129) == 256
    assert _lib.model_elems([0, 5, 135]) == 128 * 128 + 256 * 256
    assert _lib.lml_workspace_bytes([0, 5, 135], [0, 1, 1], 3) > 3 * 256 * 256 * 8


def test_null_handle_is_rejected_without_touching_a_device(built):
    """Every compute entry point returns -1 (argument 1 invalid, LAPACK style) for a NULL handle - checked here because
    it needs no GPU; the full argument-error matrix runs on the GPU box (tests/test_gpu_parity.py)."""
    import ctypes as C

    from cgp_b200 import _lib

    L = _lib.lib()
    for name, (res, args) in _lib.SIGNATURES.items():
        if res is not C.c_int or not args or args[0] is not C.c_void_p or name in ("cgp_destroy",):
            continue
        vals = []
        for a in args:
            if a in (C.c_void_p, _lib._pi64, _lib._pi32, C.POINTER(C.c_double)):
                vals.append(None)  # NULL pointer
            elif a is C.c_double:
                vals.append(0.0)
            else:
                vals.append(0)
        assert getattr(L, name)(*vals) == -1, name
    assert L.cgp_destroy(None) == -1
    assert L.cgp_append_workspace_bytes(0) == 0 and L.cgp_append_workspace_bytes(4100) > 3 * 4224 * 8
    assert L.cgp_score_workspace_bytes(None, 0, 2, 10) == 0


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import numpy as np
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cgp_b200 import CGPRegressor, KNeighborsClassifier, _lib

    X = np.random.rand(20, 2)
    with pytest.raises(_lib.CgpError):
        CGPRegressor("matern32", n_restarts_optimizer=0).fit(X, X[:, 0], np.zeros(20, dtype=int))
    clf = KNeighborsClassifier(3).fit(X, np.zeros(20, dtype=int))
    with pytest.raises(_lib.CgpError):
        clf.predict(X)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "cgp_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
