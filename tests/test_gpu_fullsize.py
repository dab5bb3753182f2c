"""GPU parity at BASELINE sizes against scikit-learn (the arithmetic the reference delegates to,
REF/cGP/cGP_constrained.py:315-327, REF/cGP/acquisition.py:12-32) and against an extended-precision truth.

* ``cfg2_full``  : BASELINE configs[1] at FULL size (Schaffer N2, N_PILOT = 2000, DGM4 -> clusters of ~500, R = 20,
  65 536 candidates): fitted thetas / LML / gradient / routed mean / std / EI / selected candidate from scikit-learn
  objects configured as the reference (oracle/make_golden.py:case_cfg2_full).
* ``cfg5_shape`` : three clusters of the configs[4] shape (8-D, n_c = 480 / 512 / 544), a quarter of the candidates
  within 1e-3 of a training point (variance cancels to ~3e-5 of the prior), with the posterior variance / mean / LML
  recomputed in 80-bit long double (oracle/extended.py).

Tolerance (north star): k-NN labels bit-exact; LML, posterior mean and posterior VARIANCE 1e-8 relative; the chosen
candidate identical.  The variance is asserted at 1e-8 on std^2 for every candidate (no std floor): the truth table
printed by test_variance_truth_table shows what scikit-learn itself achieves against the long-double value.
"""
import numpy as np
import pytest

from conftest import load_golden, relerr

pytestmark = pytest.mark.gpu
RTOL = 1e-8


def _cfg2():
    from tools import workloads as W

    g = load_golden("cfg2_full")
    X, y = W.cfg2_samples()
    return g, X, y, g["labels"].astype(np.int64), W.cfg2_candidates()


def _thetas(g, C):
    return np.stack([g[f"c{c}_theta"] for c in range(C)])


def _check_scan(eng, g, Q, C, var_rtol=RTOL):
    for c in range(C):
        ref = float(g[f"c{c}_lml"][0])
        assert abs(eng.model["lml_host"][c] - ref) <= RTOL * abs(ref), (c, eng.model["lml_host"][c], ref)
    ql = eng.route(Q, 3)
    assert np.array_equal(ql.cpu().numpy(), g["qlab"].astype(np.int64)), "k-NN labels differ from scikit-learn's"
    (bs, bi, bc), (score, mean, std) = eng.ei_scan(Q, 3, want_arrays=True, qlabel=ql)
    mean, std, score = mean.cpu().numpy(), std.cpu().numpy(), score.cpu().numpy()
    # mean: 1e-8 relative; where the mean crosses zero the denominator is floored at 1e-3 (targets are O(1))
    assert np.max(np.abs(mean - g["mean"]) / np.maximum(np.abs(g["mean"]), 1e-3)) < RTOL
    assert relerr(std ** 2, g["std"] ** 2) < var_rtol  # every candidate, no floor
    assert np.max(np.abs(score - g["score"])) < 1e-9
    assert int(bi.item()) == int(g["best_idx"][0]) and int(bc.item()) == int(g["best_cluster"][0])
    assert abs(float(bs.item()) - float(g["best_score"][0])) <= RTOL * abs(float(g["best_score"][0]))
    # the default (pruned) scan picks the same candidate with the same score bits
    (ps, pi, pc), _ = eng.ei_scan(Q, 3, qlabel=ql, prune=True)
    assert int(pi.item()) == int(bi.item()) and int(pc.item()) == int(bc.item())
    assert ps.cpu().numpy().tobytes() == bs.cpu().numpy().tobytes()
    return mean, std


def _check_grad(eng, g, C):
    """LML gradient at the fitted theta: 1e-7 of the gradient's scale (it is ~0 at an interior optimum, so the scale is
    max(|g|, |lml|); same convention as tests/test_gpu_parity.py)."""
    th = _thetas(g, C)
    lml, grad, info = eng.lml_grad(np.arange(C, dtype=np.int32), th, True)
    assert not np.any(info)
    for c in range(C):
        ref = g[f"c{c}_grad"]
        scale = max(np.max(np.abs(ref)), abs(float(g[f"c{c}_lml"][0])) * 1e-3, 1.0)
        assert np.max(np.abs(grad[c] - ref)) <= 1e-7 * scale, (c, grad[c], ref)
        assert abs(lml[c] - float(g[f"c{c}_lml"][0])) <= RTOL * abs(float(g[f"c{c}_lml"][0]))


def test_cfg2_full_size_at_sklearn_theta():
    from cgp_b200.engine import ClusterGPEngine

    g, X, y, lab, Q = _cfg2()
    C = int(lab.max()) + 1
    eng = ClusterGPEngine(X, y, lab, "matern32", 1e-5).set_theta(_thetas(g, C))
    _check_scan(eng, g, Q, C)
    _check_grad(eng, g, C)


def test_cfg2_full_size_fit_reproduces_sklearn():
    """All clusters x 21 L-BFGS-B runs in lock-step find scikit-learn's optimum (LML 1e-8) and select the same candidate."""
    from cgp_b200 import CGPRegressor

    g, X, y, lab, Q = _cfg2()
    C = int(lab.max()) + 1
    m = CGPRegressor("matern32", alpha=1e-5, n_restarts_optimizer=int(g["R"][0]), random_state=0).fit(X, y, lab)
    for c in range(C):
        ref = float(g[f"c{c}_lml"][0])
        assert abs(m.log_marginal_likelihood_value_[c] - ref) <= RTOL * abs(ref), (c, m.log_marginal_likelihood_value_[c], ref)
    x, score, idx = m.select_next(Q)
    assert idx == int(g["best_idx"][0]) and np.array_equal(x, Q[idx])
    assert abs(score - float(g["best_score"][0])) <= 1e-6 * abs(float(g["best_score"][0]))


def test_cfg5_shape_at_sklearn_theta():
    from cgp_b200.engine import ClusterGPEngine

    g = load_golden("cfg5_shape")
    lab = g["labels"].astype(np.int64)
    eng = ClusterGPEngine(g["X"], g["y"], lab, "matern32", 1e-5).set_theta(_thetas(g, 3))
    _check_scan(eng, g, g["Q"], 3)
    _check_grad(eng, g, 3)


def test_variance_truth_table(capsys):
    """GPU and scikit-learn against the long-double value of  1 + s2 - |L^-1 k*|^2  (same fp64 kernel values), per decade
    of the variance; the GPU must be within 1e-8 of the TRUTH everywhere, and within 1e-8 of scikit-learn wherever
    scikit-learn itself is (it is, on this fixture: its worst error is ~1e-10)."""
    from cgp_b200.engine import ClusterGPEngine

    g = load_golden("cfg5_shape")
    lab = g["labels"].astype(np.int64)
    eng = ClusterGPEngine(g["X"], g["y"], lab, "matern32", 1e-5).set_theta(_thetas(g, 3))
    ti = g["truth_idx"]
    ql = g["qlab"].astype(np.int64)[ti]
    ys = np.array([g[f"c{c}_ystd"][0] for c in range(3)])[ql]
    ym = np.array([g[f"c{c}_ymean"][0] for c in range(3)])[ql]
    mean, std = eng.predict(g["Q"][ti], return_std=True)
    gpu_var = (std.cpu().numpy() / ys) ** 2
    sk_var = (g["std"][ti] / ys) ** 2
    truth = g["truth_var_norm"]
    e_gpu, e_sk = np.abs(gpu_var - truth) / truth, np.abs(sk_var - truth) / truth
    rows = []
    for lo, hi in ((1e-2, 10.0), (1e-3, 1e-2), (1e-4, 1e-3), (1e-5, 1e-4)):
        m = (truth > lo) & (truth <= hi)
        if m.any():
            rows.append((lo, hi, int(m.sum()), float(e_gpu[m].max()), float(e_sk[m].max())))
    with capsys.disabled():
        print("\nposterior variance vs long-double truth (normalised units), max relative error per decade")
        print("  variance in        count    GPU            scikit-learn")
        for lo, hi, cnt, a, b in rows:
            print(f"  ({lo:.0e}, {hi:.0e}]   {cnt:6d}   {a:.3e}      {b:.3e}")
    assert e_gpu.max() < RTOL, e_gpu.max()
    ok = e_sk < 0.5 * RTOL
    assert np.max(np.abs(gpu_var[ok] - sk_var[ok]) / sk_var[ok]) < RTOL
    # mean and LML against the truth as well
    gpu_mean_norm = (mean.cpu().numpy() - ym) / ys
    assert np.max(np.abs(gpu_mean_norm - g["truth_mean_norm"]) / np.maximum(np.abs(g["truth_mean_norm"]), 1e-3)) < RTOL
    for c in range(3):
        assert abs(eng.model["lml_host"][c] - g["truth_lml"][c]) <= RTOL * abs(g["truth_lml"][c])
