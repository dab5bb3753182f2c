"""GPU parity tests: the CUDA path (through the C-ABI / ctypes binding) against the oracle, the committed
golden fixtures (sklearn / reference outputs) and size-independent properties.

Tolerances (north star): k-NN labels bit-exact; LML / posterior mean / variance 1e-8 relative; chosen sample
identical."""
import numpy as np
import pytest
import torch

import oracle
from conftest import load_golden, relerr, var_excess

pytestmark = pytest.mark.gpu

RTOL = 1e-8


@pytest.fixture(scope="module")
def lib():
    from cgp_b200 import _lib

    assert torch.cuda.is_available()
    _lib.handle()
    return _lib


def dev(a, dtype=torch.float64):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype)


# ------------------------------------------------------------------------------------------------ DMMA mainloop
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (256, 384, 160), (512, 128, 2048)])
def test_dgemm_nt(lib, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn((M, K), dtype=torch.float64, device="cuda", generator=g)
    B = torch.randn((N, K), dtype=torch.float64, device="cuda", generator=g)
    C = lib.dgemm_nt(A, B)
    ref = A @ B.T
    assert float((C - ref).abs().max()) <= 1e-12 * K


# ------------------------------------------------------------------------------------------------ K1
@pytest.mark.parametrize("name", ["lml_matern_n96_d2", "lml_matern_n300_d6", "lml_rbf_n150_d3"])
def test_kernel_matrix_vs_sklearn(lib, name):
    g = load_golden(name)
    kind = str(g["kind"][0])
    K = lib.kernel_matrix(dev(g["X"]), dev(g["thetas"][1]), kind, 0.0).cpu().numpy()
    assert np.max(np.abs(K - g["K1"])) < 1e-13
    Kc = lib.kernel_cross(dev(g["X"][:50]), dev(g["X"]), dev(g["thetas"][1]), kind).cpu().numpy()
    off = ~np.eye(g["X"].shape[0], dtype=bool)[:50]
    assert np.max(np.abs(Kc[off] - g["K1"][:50][off])) < 1e-13


# ------------------------------------------------------------------------------------------------ K2..K6
def _lml_batch(lib, X, yn, thetas, kind, alpha=1e-5, offs=None, pc=None):
    n, d = X.shape
    offs = np.array([0, n]) if offs is None else offs
    pc = np.zeros(len(thetas), dtype=np.int32) if pc is None else pc
    ws = torch.empty(lib.lml_workspace_bytes(offs, pc, d), dtype=torch.uint8, device="cuda")
    lml, grad, info = lib.lml_grad_batched(dev(X), dev(yn), offs, pc, dev(thetas), alpha, kind, ws)
    return lml.cpu().numpy(), grad.cpu().numpy(), info.cpu().numpy()


@pytest.mark.parametrize("name", ["lml_matern_n96_d2", "lml_matern_n300_d6", "lml_rbf_n150_d3"])
def test_lml_grad_vs_sklearn(lib, name):
    g = load_golden(name)
    kind = str(g["kind"][0])
    lml, grad, info = _lml_batch(lib, g["X"], g["y_norm"], g["thetas"], kind)
    assert np.all(info == 0)
    assert relerr(lml, g["lml"]) < RTOL
    scale = np.maximum(np.abs(g["grad"]).max(axis=1, keepdims=True), 1.0)
    assert np.max(np.abs(grad - g["grad"]) / scale) < 1e-7


def test_bukin_known_answers(lib):
    g = load_golden("bukin")
    for i in range(2):
        lml, _, info = _lml_batch(lib, g[f"g{i}_X"], g[f"g{i}_y"], g[f"g{i}_theta"][None, :], "matern32")
        assert info[0] == 0
        assert abs(lml[0] - g[f"g{i}_lml"][0]) <= RTOL * abs(g[f"g{i}_lml"][0])
        X = g[f"g{i}_X"]
        m = lib.factorize(dev(X), dev(g[f"g{i}_y"]), np.array([0, X.shape[0]]), dev(g[f"g{i}_theta"][None, :]), 1e-5,
                          "matern32")
        n, npad = X.shape[0], lib.padded_n(X.shape[0])
        L = m["L"].view(npad, npad)[:n, :n].cpu().numpy()
        assert np.max(np.abs(L - g[f"g{i}_L"])) < 1e-10
        assert relerr(m["alpha_vec"].cpu().numpy(), g[f"g{i}_alpha"]) < 1e-7


def test_lml_ragged_batch_and_nonpd(lib):
    """Three clusters of ragged sizes (5, 130, 257 -> 1, 2, 3 tiles), several thetas each, one non-PD item."""
    rng = np.random.default_rng(3)
    sizes = [5, 130, 257]
    d = 3
    offs = np.concatenate([[0], np.cumsum(sizes)])
    X = rng.uniform(size=(offs[-1], d))
    y = np.concatenate([oracle.normalize_y(np.sin(5 * X[offs[c]:offs[c + 1]].sum(1)) + 0.1 * rng.standard_normal(sizes[c]))[0]
                        for c in range(3)])
    pc = np.array([0, 1, 2, 2, 1, 0, 2], dtype=np.int32)
    thetas = rng.uniform(-2, 1, size=(len(pc), d + 1))
    lml, grad, info = _lml_batch(lib, X, y, thetas, "matern32", offs=offs, pc=pc)
    for i, c in enumerate(pc):
        s = slice(offs[c], offs[c + 1])
        l, gr = oracle.lml_and_grad(X[s], y[s], thetas[i], 1e-5, "matern32")
        assert info[i] == 0
        assert abs(lml[i] - l) <= RTOL * abs(l)
        assert np.max(np.abs(grad[i] - gr)) <= 1e-7 * max(1.0, np.max(np.abs(gr)))
    # negative jitter makes K indefinite -> info > 0 (LAPACK convention), lml = -inf downstream
    lml, grad, info = _lml_batch(lib, X, y, thetas, "matern32", alpha=-10.0, offs=offs, pc=pc)
    assert np.all(info > 0)
    assert np.all(np.isinf(lml)) and np.all(lml < 0)
    assert np.all(grad == 0)
    # a tiny workspace forces the internal chunking path: identical results
    one = max(lib.lml_workspace_bytes(offs, np.array([c], dtype=np.int32), d) for c in range(3))
    ws = torch.empty(int(one * 1.5), dtype=torch.uint8, device="cuda")
    l2, g2, i2 = lib.lml_grad_batched(dev(X), dev(y), offs, pc, dev(thetas), 1e-5, "matern32", ws)
    l1, g1, _ = _lml_batch(lib, X, y, thetas, "matern32", offs=offs, pc=pc)
    assert np.array_equal(l2.cpu().numpy(), l1) and np.array_equal(g2.cpu().numpy(), g1)


def test_factorize_properties_large(lib):
    """n = 1500 (12 tiles): L L^T = K, W L = I, alpha = K^-1 y, lml vs LAPACK — size-independent properties."""
    rng = np.random.default_rng(5)
    n, d = 1500, 4
    X = rng.uniform(size=(n, d))
    y, _, _ = oracle.normalize_y(np.sin(4 * X[:, 0]) * np.cos(2 * X[:, 1]) + X[:, 2] * X[:, 3] + 0.01 * rng.standard_normal(n))
    th = np.array([-1.0, -0.5, 0.2, 0.0, -6.0])
    m = lib.factorize(dev(X), dev(y), np.array([0, n]), dev(th[None, :]), 1e-5, "matern32")
    npad = lib.padded_n(n)
    L = m["L"].view(npad, npad).cpu().numpy()
    W = m["W"].view(npad, npad).cpu().numpy()
    assert np.all(np.triu(L[:n, :n], 1) == 0)
    assert np.allclose(L[n:, n:], np.eye(npad - n)) and np.all(L[n:, :n] == 0)
    K = oracle.kernel_matrix(X, th, "matern32", 1e-5)
    assert np.max(np.abs(L[:n, :n] @ L[:n, :n].T - K)) < 1e-12
    assert np.max(np.abs(np.tril(W[:n, :n]) @ L[:n, :n] - np.eye(n))) < 1e-8
    a = np.linalg.solve(K, y)
    assert relerr(m["alpha_vec"].cpu().numpy(), a) < 1e-6
    lml = oracle.lml_and_grad(X, y, th, 1e-5, want_grad=False)
    assert abs(float(m["lml"][0]) - lml) <= RTOL * abs(lml)
    assert int(m["info"][0]) == 0


# ------------------------------------------------------------------------------------------------ K7
@pytest.fixture(params=["exact", "tc"])
def knn_mode(request, lib):
    """Both k-NN paths: the exact fp64 brute-force kernel and the tensor-core prefilter + exact finish (forced on even
    for training sets it would not normally be chosen for)."""
    old = lib.set_option("knn_mode", request.param)
    yield request.param
    lib.set_option("knn_mode", old)


@pytest.mark.parametrize("name", ["fit_select_d2", "fit_select_d6"])
def test_knn_bit_exact_vs_sklearn(lib, name, knn_mode):
    g = load_golden(name)
    C = int(g["labels"].max()) + 1
    lab = lib.knn_labels(dev(g["X"]), dev(g["labels"], torch.int64), dev(g["Q"]), 3, C).cpu().numpy()
    assert np.array_equal(lab, g["qlab"])


@pytest.mark.parametrize("d,n,m,k", [(8, 8192, 20000, 3), (6, 5000, 12000, 3), (11, 3000, 6000, 2), (2, 4000, 9000, 1),
                                     (16, 2100, 4100, 3)])
def test_knn_prefilter_bit_exact_random(lib, knn_mode, d, n, m, k):
    """Continuous data (the prefilter certifies almost every query) incl. near-duplicates of training points."""
    rng = np.random.default_rng(100 + d)
    X = rng.uniform(size=(n, d))
    labels = rng.integers(0, 9, size=n)
    Q = rng.uniform(size=(m, d))
    Q[: m // 8] = X[rng.integers(0, n, size=m // 8)] + 1e-9 * rng.standard_normal((m // 8, d))
    Q[m // 8: m // 6] = X[rng.integers(0, n, size=m // 6 - m // 8)]  # exact copies of training points
    got = lib.knn_labels(dev(X), dev(labels, torch.int64), dev(Q), k, 9).cpu().numpy()
    assert np.array_equal(got, oracle.knn_labels(X, labels, Q, k))


def test_knn_prefilter_offset_and_scaled_data(lib, knn_mode):
    """Coordinates far from the origin (the error bound of the fp32 ranking grows with |x|^2: most queries fall back
    to the exact kernel) and Schaffer-like [-100, 100] boxes: labels stay bit-exact."""
    rng = np.random.default_rng(77)
    for shift, scale in ((1000.0, 1.0), (0.0, 100.0), (-3.0e4, 10.0)):
        X = shift + scale * rng.uniform(-1, 1, size=(3000, 3))
        labels = rng.integers(0, 4, size=3000)
        Q = shift + scale * rng.uniform(-1, 1, size=(5000, 3))
        got = lib.knn_labels(dev(X), dev(labels, torch.int64), dev(Q), 3, 4).cpu().numpy()
        assert np.array_equal(got, oracle.knn_labels(X, labels, Q, 3))


def test_knn_ties_and_k(lib, knn_mode):
    """Integer grid -> many exact distance ties; lowest training index must win, vote ties -> smallest label."""
    rng = np.random.default_rng(9)
    X = rng.integers(0, 4, size=(600, 3)).astype(np.float64)
    labels = rng.integers(0, 5, size=600)
    Q = rng.integers(0, 4, size=(3000, 3)).astype(np.float64) + 0.5 * rng.integers(0, 2, size=(3000, 3))
    for k in (1, 3, 4, 8):
        got = lib.knn_labels(dev(X), dev(labels, torch.int64), dev(Q), k, 5).cpu().numpy()
        assert np.array_equal(got, oracle.knn_labels(X, labels, Q, k))
    # training set larger than one shared-memory tile
    X2 = rng.uniform(size=(2500, 6))
    l2 = rng.integers(0, 7, size=2500)
    Q2 = rng.uniform(size=(5000, 6))
    got = lib.knn_labels(dev(X2), dev(l2, torch.int64), dev(Q2), 3, 7).cpu().numpy()
    assert np.array_equal(got, oracle.knn_labels(X2, l2, Q2, 3))
    assert lib.knn_labels(dev(X2), dev(l2, torch.int64), dev(Q2[:0]), 3, 7).numel() == 0


# ------------------------------------------------------------------------------------------------ K8 + EI + argmax
@pytest.mark.parametrize("name,C", [("fit_select_d2", 2), ("fit_select_d6", 3)])
def test_predict_and_select_at_reference_theta(lib, name, C):
    from cgp_b200.engine import ClusterGPEngine

    g = load_golden(name)
    eng = ClusterGPEngine(g["X"], g["y"], g["labels"], str(g["kind"][0]), 1e-5)
    eng.set_theta(np.stack([g[f"c{c}_theta"] for c in range(C)]))
    for c in range(C):
        assert abs(eng.model["lml_host"][c] - g[f"c{c}_lml"][0]) <= RTOL * abs(g[f"c{c}_lml"][0])
        assert relerr(eng.alpha_of(c), g[f"c{c}_alpha"]) < 1e-6
        assert relerr(np.diag(eng.L_of(c)), g[f"c{c}_Ldiag"]) < 1e-9
    ql = eng.route(g["Q"], 3)
    assert np.array_equal(ql.cpu().numpy(), g["qlab"])
    (bs, bi, bc), (score, mean, std) = eng.ei_scan(g["Q"], 3, want_arrays=True)
    mean, std, score = mean.cpu().numpy(), std.cpu().numpy(), score.cpu().numpy()
    assert relerr(mean, g["mean"]) < RTOL
    assert var_excess(std, g["std"]) <= 1.0  # variance 1e-8 relative, every candidate (conftest.var_excess)
    assert np.max(np.abs(score - g["score"])) < 1e-9
    assert np.max(np.abs(score[g["pw_idx"]] - g["pw_val"])) < 1e-9  # reference's own pointwise function
    assert int(bi.item()) == int(g["best_idx"][0])
    assert int(bc.item()) == int(g["best_cluster"][0])
    assert abs(float(bs.item()) - float(g["best_score"][0])) < 1e-9 * abs(float(g["best_score"][0]))
    m2, s2 = eng.predict(g["Q"][:777], return_std=True)
    assert np.array_equal(m2.cpu().numpy(), mean[:777]) and np.array_equal(s2.cpu().numpy(), std[:777])


@pytest.mark.parametrize("name,C", [("fit_select_d2", 2), ("fit_select_d6", 3)])
def test_full_fit_vs_sklearn(name, C):
    """Lock-step L-BFGS-B over all clusters x restarts reproduces scikit-learn's fit (same optimum, same sample)."""
    from cgp_b200 import CGPRegressor

    g = load_golden(name)
    R = int(g["R"][0])
    m = CGPRegressor(str(g["kind"][0]), alpha=1e-5, n_restarts_optimizer=R, random_state=0).fit(g["X"], g["y"], g["labels"])
    for c in range(C):
        assert abs(m.log_marginal_likelihood_value_[c] - g[f"c{c}_lml"][0]) <= RTOL * abs(g[f"c{c}_lml"][0])
    x, score, idx = m.select_next(g["Q"])
    assert idx == int(g["best_idx"][0])
    assert np.array_equal(x, g["Q"][idx])
    out = m.score_candidates(g["Q"])
    assert relerr(out["mean"], g["mean"]) < 1e-5
    assert np.array_equal(out["labels"], g["qlab"])


def test_sklearn_shaped_gpr_surface():
    """The single-GP facade behaves like the object built at REF/cGP/cGP_constrained.py:315-327."""
    from sklearn.gaussian_process.kernels import Matern, WhiteKernel

    from cgp_b200 import GaussianProcessRegressor, KNeighborsClassifier
    from cgp_b200.acquisition import expected_improvement

    g = load_golden("fit_select_d2")
    X, y, labels = g["X"], g["y"], g["labels"]
    Xt, Yt = X[labels == 1], y[labels == 1].reshape(-1, 1)
    k = Matern(length_scale=np.ones(2), length_scale_bounds=(1e-05, 100000.0), nu=3 / 2) + \
        WhiteKernel(noise_level=1.0, noise_level_bounds=(1e-05, 100000.0))
    mt = GaussianProcessRegressor(kernel=k, random_state=0, normalize_y=True, alpha=1e-5, optimizer="fmin_l_bfgs_b",
                                  n_restarts_optimizer=20)
    mt.fit(Xt, Yt)
    assert abs(mt.log_marginal_likelihood_value_ - g["c1_lml"][0]) <= RTOL * abs(g["c1_lml"][0])
    assert abs(mt.log_marginal_likelihood(mt.kernel_.theta) - g["c1_lml"][0]) <= RTOL * abs(g["c1_lml"][0])
    assert mt.L_.shape == (Xt.shape[0], Xt.shape[0]) and mt.alpha_.shape == (Xt.shape[0], 1)
    clf = KNeighborsClassifier(n_neighbors=3).fit(X, labels)
    assert np.array_equal(clf.predict(g["Q"]), g["qlab"])
    for i, v in zip(g["pw_idx"], g["pw_val"]):
        if g["qlab"][i] != 1:
            continue
        got = -expected_improvement(X=g["Q"][i], surrogate=mt, X_sample=Xt, Y_sample=Yt, correct_label=np.float64(1),
                                    classify=clf, VERBOSE=False)
        assert abs(got - v) < 1e-7 * max(abs(v), 1e-3)
        break
    import dill

    mt2 = dill.loads(dill.dumps(mt))
    mu1 = mt.predict(g["Q"][:50])
    mu2 = mt2.predict(g["Q"][:50])
    assert np.allclose(mu1, mu2, rtol=1e-12, atol=0)


def test_roundtrip_properties_midsize():
    """n=3000 in 3 clusters (8..9 tiles each), M=20000: GPU scan == oracle scan at the GPU's own thetas."""
    from cgp_b200.engine import ClusterGPEngine

    rng = np.random.default_rng(17)
    n, d, C, M = 3000, 4, 3, 20000
    X = rng.uniform(size=(n, d))
    labels = np.minimum((X[:, 0] * C).astype(np.int64), C - 1)
    y = labels + np.sin(5 * X.sum(1)) + 0.02 * rng.standard_normal(n)
    Q = rng.uniform(size=(M, d))
    th = np.tile(np.array([-0.8, -0.3, 0.1, 0.4, -5.0]), (C, 1))
    eng = ClusterGPEngine(X, y, labels, "matern32", 1e-5).set_theta(th)
    (bs, bi, bc), (score, mean, std) = eng.ei_scan(Q, 3, want_arrays=True)
    models = []
    for c in range(C):
        Xt, yt = X[labels == c], y[labels == c]
        yn, ym, ys = oracle.normalize_y(yt)
        K = oracle.kernel_matrix(Xt, th[c], "matern32", 1e-5)
        L = np.linalg.cholesky(K)
        models.append(dict(X=Xt, theta=th[c], L=L, alpha_vec=np.linalg.solve(K, yn), y_mean=ym, y_std=ys, kind="matern32"))
    ref = oracle.select_next(X, y, labels, models, Q, 3)
    assert np.array_equal(eng.route(Q, 3).cpu().numpy(), ref["labels"])
    assert relerr(mean.cpu().numpy(), ref["mean"]) < RTOL
    assert var_excess(std.cpu().numpy(), ref["std"]) <= 1.0
    assert int(bi.item()) == ref["best_idx"] and int(bc.item()) == ref["best_cluster"]


def test_capi_argument_errors(lib):
    """LAPACK-style negative return codes for invalid arguments (include/cgp_b200.h conventions)."""
    import ctypes as C

    L = lib.lib()
    h = lib.handle()
    X = dev(np.random.default_rng(0).uniform(size=(10, 2)))
    th = dev(np.zeros(3))
    K = torch.empty((10, 10), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    assert L.cgp_kernel_matrix(None, 0, X.data_ptr(), 10, 2, th.data_ptr(), 0.0, K.data_ptr(), st) == -1
    assert L.cgp_kernel_matrix(h, 7, X.data_ptr(), 10, 2, th.data_ptr(), 0.0, K.data_ptr(), st) == -2
    assert L.cgp_kernel_matrix(h, 0, X.data_ptr(), 10, 0, th.data_ptr(), 0.0, K.data_ptr(), st) == -5
    assert L.cgp_kernel_matrix(h, 0, X.data_ptr(), 10, 17, th.data_ptr(), 0.0, K.data_ptr(), st) == -5
    assert L.cgp_kernel_matrix(h, 0, X.data_ptr(), 10, 2, None, 0.0, K.data_ptr(), st) == -6
    lab = dev(np.zeros(10), torch.int64)
    out = torch.empty(4, dtype=torch.int64, device="cuda")
    assert L.cgp_knn_labels(h, X.data_ptr(), lab.data_ptr(), 10, 2, 11, 1, X.data_ptr(), 4, out.data_ptr(), st) == -6  # k > n
    assert L.cgp_knn_labels(h, X.data_ptr(), lab.data_ptr(), 10, 2, 9, 1, X.data_ptr(), 4, out.data_ptr(), st) == -6   # k > CGP_MAX_K
    assert L.cgp_dgemm_nt(h, X.data_ptr(), X.data_ptr(), K.data_ptr(), 100, 128, 32, st) == -5
    # a workspace smaller than one pair is refused, not overrun
    offs = np.array([0, 10], dtype=np.int64)
    pc = np.zeros(1, dtype=np.int32)
    ws = torch.empty(1024, dtype=torch.uint8, device="cuda")
    with pytest.raises(lib.CgpError):
        lib.lml_grad_batched(X, dev(np.zeros(10)), offs, pc, dev(np.zeros((1, 3))), 1e-5, "matern32", ws)
    # the binding refuses host / wrong-dtype tensors instead of passing bad pointers
    with pytest.raises(lib.CgpError):
        lib.kernel_matrix(X.cpu(), th)
    with pytest.raises(lib.CgpError):
        lib.kernel_matrix(X.float(), th)


def test_return_cov_and_mspe_vs_reference():
    """predict(return_cov=True) (SK/_gpr.py:464-478) and the reference's mean_square_prediction_error
    (REF/cGP/acquisition.py:34-75) against outputs of the reference's own code (tests/golden/cov_mspe_d3.npz)."""
    from cgp_b200 import GaussianProcessRegressor, KNeighborsClassifier
    from cgp_b200.acquisition import mean_square_prediction_error
    from sklearn.gaussian_process.kernels import Matern, WhiteKernel

    g = load_golden("cov_mspe_d3")
    X, y, Q = g["X"], g["y"], g["Q"]
    k = (Matern(length_scale=np.ones(3), length_scale_bounds=(1e-05, 100000.0), nu=3 / 2) +
         WhiteKernel(noise_level=1.0, noise_level_bounds=(1e-05, 100000.0))).clone_with_theta(g["theta"])
    m = GaussianProcessRegressor(kernel=k, alpha=1e-5, normalize_y=True, optimizer=None, random_state=0).fit(X, y.reshape(-1, 1))
    mean, cov = m.predict(Q, return_cov=True)
    assert relerr(mean, g["mean"]) < RTOL
    assert np.max(np.abs(cov - g["cov"])) < 1e-9 * np.max(np.abs(g["cov"]))
    with pytest.raises(RuntimeError):
        m.predict(Q, return_std=True, return_cov=True)
    clf = KNeighborsClassifier(3).fit(X, np.zeros(X.shape[0], dtype=np.int64))
    for i, want in enumerate(g["mspe"]):
        got = mean_square_prediction_error(X=Q[i], surrogate=m, X_sample=X, Y_sample=y.reshape(-1, 1),
                                           correct_label=np.float64(0), classify=clf, VERBOSE=False)
        assert abs(got - want) <= 1e-7 * abs(want)


def test_dense_mspe_scan_vs_reference_pointwise():
    """The dense MSPE scan (engine.mspe_scan -> cgp_quadform_scan: one [m, n_c] x [n_c, n_c] DMMA product + row dot) against (a) the
    reference's own outputs for a single cluster (tests/golden/cov_mspe_d3.npz, made by REF/cGP/acquisition.py:34-75)
    and (b) the point-wise restatement with scikit-learn surrogates on a two-cluster model; selection = minimum,
    ties to the lowest cluster id then index."""
    from cgp_b200 import CGPRegressor
    from cgp_b200.acquisition import mean_square_prediction_error, mean_square_prediction_error_batch
    from sklearn.gaussian_process import GaussianProcessRegressor as SkGPR
    from sklearn.gaussian_process.kernels import Matern, WhiteKernel
    from sklearn.neighbors import KNeighborsClassifier as SkKNN

    g = load_golden("cov_mspe_d3")
    X, y, Q = g["X"], g["y"], g["Q"]
    k = (Matern(length_scale=np.ones(3), length_scale_bounds=(1e-05, 100000.0), nu=3 / 2) +
         WhiteKernel(noise_level=1.0, noise_level_bounds=(1e-05, 100000.0))).clone_with_theta(g["theta"])
    m1 = CGPRegressor(kernel=k, alpha=1e-5, optimizer=None, random_state=0, n_neighbors=3).fit(X, y, np.zeros(len(y), dtype=int))
    got = mean_square_prediction_error_batch(m1, Q)
    k_g = len(g["mspe"])  # the golden file holds the reference's value for the first few candidates
    assert got.shape == (Q.shape[0],) and np.max(np.abs(got[:k_g] - g["mspe"]) / np.abs(g["mspe"])) <= 1e-7
    # two clusters, routed
    rng = np.random.default_rng(9)
    n, d, M = 420, 2, 3000
    X = rng.uniform(size=(n, d))
    lab = (X[:, 0] > 0.45).astype(np.int64)
    y = lab + np.sin(5 * X.sum(1)) + 0.02 * rng.standard_normal(n)
    Qs = rng.uniform(size=(M, d))
    Qs[-100:] = Qs[:100]  # exact ties
    th = np.array([[-0.8, -0.5, -4.0], [-0.3, -1.0, -3.0]])
    from cgp_b200.engine import ClusterGPEngine

    eng = ClusterGPEngine(X, y, lab, "matern32", 1e-5).set_theta(th)
    (v, i, c), arr = eng.mspe_scan(Qs, 3, want_arrays=True)
    arr = arr.cpu().numpy()
    ql = eng.route(Qs, 3).cpu().numpy()
    clf = SkKNN(n_neighbors=3).fit(X, lab)
    assert np.array_equal(ql, clf.predict(Qs))
    for cc in range(2):
        kk = (Matern(length_scale=np.ones(d), length_scale_bounds=(1e-05, 100000.0), nu=3 / 2) +
              WhiteKernel(noise_level=1.0, noise_level_bounds=(1e-05, 100000.0))).clone_with_theta(th[cc])
        Xt, yt = X[lab == cc], y[lab == cc].reshape(-1, 1)
        sk = SkGPR(kernel=kk, alpha=1e-5, normalize_y=True, optimizer=None, random_state=0).fit(Xt, yt)
        for q in np.nonzero(ql == cc)[0][:25]:
            want = mean_square_prediction_error(X=Qs[q], surrogate=sk, X_sample=Xt, Y_sample=yt, correct_label=np.float64(cc),
                                                classify=clf, VERBOSE=False)
            assert abs(arr[q] - want) <= 1e-6 * max(abs(want), 1e-12), (cc, q, arr[q], want)
    order = np.lexsort((np.arange(M), ql, arr))  # minimum; ties -> lowest cluster id, then lowest index
    assert int(i.item()) == int(order[0]) and int(c.item()) == int(ql[order[0]]) and float(v.item()) == arr[order[0]]
    # additive soft-constraint term inside the scan: (mspe + penalty) / n_c, REF/cGP/acquisition.py:69,75
    pen = rng.uniform(0, 1e-3, size=M)
    (vp, ip, cp), arrp = eng.mspe_scan(Qs, 3, want_arrays=True, penalty=pen)
    n_c = np.bincount(lab)[ql]
    assert np.max(np.abs(arrp.cpu().numpy() - (arr + pen / n_c))) <= 1e-15 * np.max(np.abs(arr + pen / n_c)) + 1e-18
    op = np.lexsort((np.arange(M), ql, arrp.cpu().numpy()))
    assert int(ip.item()) == int(op[0])
    x, val, idx, cl = eng.select_next_mspe(Qs, 3)
    assert idx == int(order[0]) and val == arr[order[0]] and np.array_equal(x, Qs[idx])


def test_workspace_hygiene_nan_fill_and_canaries(lib):
    """compute-sanitizer is closed on this pool, so: (1) every workspace / output buffer is pre-filled with 0xFF bytes
    (NaN as fp64, -1 as int) — a kernel that consumed uninitialised memory would propagate NaN into the results;
    (2) a canary region behind each workspace must stay untouched (no out-of-bounds writes past the declared size)."""
    rng = np.random.default_rng(21)
    sizes = [37, 200, 300]
    d, C = 4, 3
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n = int(offs[-1])
    X = rng.uniform(size=(n, d))
    lab_sorted = np.repeat(np.arange(C), sizes)
    y = np.concatenate([oracle.normalize_y(np.sin(4 * X[offs[c]:offs[c + 1]].sum(1)) + 0.1 * rng.standard_normal(sizes[c]))[0]
                        for c in range(C)])
    pc = np.array([0, 1, 2, 1, 2], dtype=np.int32)
    th = rng.uniform(-1, 0.5, size=(len(pc), d + 1))
    th[:, d] = -3.0
    CAN = 4096

    def poisoned(nbytes):
        t = torch.full((nbytes + CAN,), 0xFF, dtype=torch.uint8, device="cuda")
        t[nbytes:] = 0xA5
        return t

    def canary_ok(t, nbytes):
        return bool((t[nbytes:] == 0xA5).all())

    need = lib.lml_workspace_bytes(offs, pc, d)
    ws = poisoned(need)
    Xd, yd, thd = dev(X), dev(y), dev(th)
    lml, grad, info = lib.lml_grad_batched(Xd, yd, offs, pc, thd, 1e-5, "matern32", ws[:need])
    assert canary_ok(ws, need)
    lml, grad = lml.cpu().numpy(), grad.cpu().numpy()
    assert np.all(np.isfinite(lml)) and np.all(np.isfinite(grad)) and np.all(info.cpu().numpy() == 0)
    for i, c in enumerate(pc):
        s = slice(offs[c], offs[c + 1])
        l, g = oracle.lml_and_grad(X[s], y[s], th[i], 1e-5, "matern32")
        assert abs(lml[i] - l) <= RTOL * abs(l) and np.max(np.abs(grad[i] - g)) <= 1e-7 * max(1.0, np.max(np.abs(g)))
    # factorize + scoring with poisoned workspaces
    need_f = lib.factorize_workspace_bytes(offs, d)
    wf = poisoned(need_f)
    m = lib.factorize(Xd, yd, offs, dev(th[:C]), 1e-5, "matern32", wf[:need_f])
    assert canary_ok(wf, need_f) and np.all(np.isfinite(m["alpha_vec"].cpu().numpy()))
    M = 1000
    Q = rng.uniform(size=(M, d))
    ql = dev(rng.integers(0, C, size=M), torch.int64)
    need_s = lib.score_workspace_bytes(offs, d, M)
    wsc = poisoned(need_s)
    ones = dev(np.ones(C))
    (bs, bi, bc), (score, mean, std) = lib.ei_argmax(Xd, offs, dev(th[:C]), m["W"], m["alpha_vec"], dev(np.zeros(C)), ones,
                                                      dev(np.zeros(C)), dev(Q), ql, "matern32", 0, True, wsc[:need_s])
    assert canary_ok(wsc, need_s)
    sc = score.cpu().numpy()
    assert np.all(np.isfinite(sc)) and np.all(np.isfinite(mean.cpu().numpy())) and np.all(std.cpu().numpy() >= 0)
    assert int(bi.item()) == int(np.argmax(sc)) or sc[int(bi.item())] == sc.max()
    (bs2, bi2, bc2), _ = lib.ei_argmax(Xd, offs, dev(th[:C]), m["W"], m["alpha_vec"], dev(np.zeros(C)), ones, dev(np.zeros(C)),
                                        dev(Q), ql, "matern32", 0, False)
    assert int(bi2.item()) == int(bi.item()) and float(bs2.item()) == float(bs.item())
    # pruned scan with a poisoned workspace: same winner, canary intact
    wsp = poisoned(need_s)
    (bs3, bi3, bc3), _ = lib.ei_argmax(Xd, offs, dev(th[:C]), m["W"], m["alpha_vec"], dev(np.zeros(C)), ones, dev(np.zeros(C)),
                                        dev(Q), ql, "matern32", 0, False, wsp[:need_s], True, alpha=1e-5)
    assert canary_ok(wsp, need_s)
    assert int(bi3.item()) == int(bi.item()) and float(bs3.item()) == float(bs.item()) and int(bc3.item()) == int(bc.item())
    # rank-1 append + alpha refresh on poisoned workspaces / poisoned new rows
    c, n_old = 1, sizes[1]
    n_new = n_old + 1
    npad = lib.padded_n(n_new)
    assert npad == lib.padded_n(n_old)
    o = int(sum(lib.padded_n(v) ** 2 for v in sizes[:c]))
    Lc = m["L"][o:o + npad * npad].view(npad, npad).clone()
    Wc = m["W"][o:o + npad * npad].view(npad, npad).clone()
    x_new = rng.uniform(size=(1, d))
    Xc = dev(np.vstack([X[offs[c]:offs[c + 1]], x_new]))
    need_a = lib.append_workspace_bytes(n_new)
    wa = poisoned(need_a)
    info = lib.append_sample(Xc, dev(th[c]), 1e-5, "matern32", Lc, Wc, wa[:need_a])
    assert canary_ok(wa, need_a) and int(info.item()) == 0
    Kn = oracle.kernel_matrix(np.vstack([X[offs[c]:offs[c + 1]], x_new]), th[c], "matern32", 1e-5)
    Ln = np.linalg.cholesky(Kn)
    assert np.max(np.abs(Lc[:n_new, :n_new].cpu().numpy() - Ln)) < 1e-11
    assert np.max(np.abs(Wc[:n_new, :n_new].cpu().numpy() @ Ln - np.eye(n_new))) < 1e-9
    assert float(Lc[n_new, n_new].item()) == 1.0 and float(Lc[n_new - 1, n_new].item()) == 0.0  # pad untouched
    y_c = dev(rng.standard_normal(n_new))
    a_out = torch.full((n_new + 8,), float("nan"), dtype=torch.float64, device="cuda")
    wa2 = poisoned(need_a)
    lml_c = lib.refresh_alpha(Lc, Wc, n_new, y_c, a_out[:n_new], wa2[:need_a])
    assert canary_ok(wa2, need_a) and bool(torch.isnan(a_out[n_new:]).all())
    a_ref = np.linalg.solve(Kn, y_c.cpu().numpy())
    assert np.max(np.abs(a_out[:n_new].cpu().numpy() - a_ref)) <= 1e-9 * np.max(np.abs(a_ref))
    z = np.linalg.solve(Ln, y_c.cpu().numpy())
    lml_ref = -0.5 * z @ z - np.log(np.diag(Ln)).sum() - 0.5 * n_new * np.log(2 * np.pi)
    assert abs(float(lml_c.item()) - lml_ref) <= 1e-10 * abs(lml_ref)
