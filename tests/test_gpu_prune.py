"""Exact bound-pruned EI scan (cgp_ei_argmax_pruned) selects the same candidate, with the same score bits, as the
exhaustive scan (cgp_ei_argmax) — on landscapes where almost everything is pruned, where nothing can be pruned
(all scores underflow to 0: the tie rule decides), and across several chunks per cluster."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _engine(n, d, C, seed, noise=0.01, theta=None):
    from cgp_b200.engine import ClusterGPEngine

    rng = np.random.default_rng(seed)
    X = rng.uniform(size=(n, d))
    lab = (np.floor(X[:, 0] * C).astype(np.int64)) % C
    y = lab + np.sin(2 * np.pi * (lab + 1) * X.sum(1) / d) + noise * rng.standard_normal(n)
    eng = ClusterGPEngine(X, y, lab, "matern32", 1e-5)
    th = np.tile(np.array([-0.7] * d + [-4.0]), (C, 1)) if theta is None else theta
    eng.set_theta(th)
    return eng, rng


def _same(eng, Q, k=3, idx_base=0):
    ql = eng.route(Q, k)
    (fs, fi, fc), _ = eng.ei_scan(Q, k, idx_base=idx_base, qlabel=ql)
    (ps, pi, pc), _ = eng.ei_scan(Q, k, idx_base=idx_base, qlabel=ql, prune=True)
    assert int(fi.item()) == int(pi.item()) and int(fc.item()) == int(pc.item())
    assert fs.cpu().numpy().tobytes() == ps.cpu().numpy().tobytes()  # same bits
    return float(fs.item()), int(fi.item())


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_pruned_scan_selects_identically(seed):
    import torch

    eng, rng = _engine(1500, 4, 3, seed)
    Q = torch.from_numpy(rng.uniform(size=(40000, 4))).cuda()
    score, idx = _same(eng, Q, idx_base=1000 * seed)
    assert score > 0 and idx >= 1000 * seed


def test_pruned_scan_many_chunks_per_cluster():
    import torch

    eng, rng = _engine(256, 2, 2, 7)  # npad = 128/256 -> chunks of 65536 candidates, ~3 chunks per cluster
    Q = torch.from_numpy(rng.uniform(size=(400000, 2))).cuda()
    _same(eng, Q)


def test_pruned_scan_all_zero_scores_and_ties():
    """Targets far below the incumbent everywhere except at the incumbent itself and a tiny predictive variance: EI
    underflows to exactly 0 for (almost) every candidate, so the first maximum in (cluster, index) order must win in
    both scans; duplicated candidates tie exactly."""
    import torch

    from cgp_b200.engine import ClusterGPEngine

    rng = np.random.default_rng(3)
    X = rng.uniform(size=(600, 2))
    lab = (X[:, 0] > 0.5).astype(np.int64)
    y = np.where(np.arange(600) % 97 == 0, 50.0, 0.0) + 1e-3 * rng.standard_normal(600)  # isolated spikes = incumbents
    eng = ClusterGPEngine(X, y, lab, "matern32", 1e-5).set_theta(np.array([[-3.0, -3.0, -9.0], [-3.0, -3.0, -9.0]]))
    q = rng.uniform(size=(20000, 2))
    Q = torch.from_numpy(np.concatenate([q, q[:5000]])).cuda()  # exact duplicates -> exact score ties
    _same(eng, Q)


def test_pruned_scan_rejects_per_candidate_outputs():
    import torch

    from cgp_b200 import _lib

    eng, rng = _engine(300, 2, 2, 5)
    Q = torch.from_numpy(rng.uniform(size=(1000, 2))).cuda()
    with pytest.raises(_lib.CgpError):
        eng.ei_scan(Q, 3, want_arrays=True, prune=True)


def test_pruned_scan_candidates_on_top_of_samples():
    """The single-neighbour variance bound is tightest exactly where the exact variance cancels (candidates on or within
    1e-9..1e-3 of a training point, noise at its lower bound 1e-5): the bound must still dominate the exact score, i.e.
    the selection (and its score bits) must equal the exhaustive scan's."""
    import torch

    from cgp_b200.engine import ClusterGPEngine

    rng = np.random.default_rng(11)
    n, d, C = 900, 3, 3
    X = rng.uniform(size=(n, d))
    lab = np.minimum((X[:, 0] * C).astype(np.int64), C - 1)
    y = np.sin(7 * X.sum(1)) + lab + 1e-3 * rng.standard_normal(n)
    for noise in (np.log(1e-5), -6.0, -2.0):
        eng = ClusterGPEngine(X, y, lab, "matern32", 1e-5).set_theta(np.tile(np.array([-1.0, -0.8, -1.2, noise]), (C, 1)))
        parts = [X, X + 1e-9 * rng.standard_normal(X.shape), X + 1e-5 * rng.standard_normal(X.shape),
                 X + 1e-3 * rng.standard_normal(X.shape), rng.uniform(size=(20000, d))]
        Q = torch.from_numpy(np.clip(np.concatenate(parts), 0, 1)).cuda()
        _same(eng, Q)


def test_pruned_scan_flat_targets_and_incumbents_everywhere():
    """(a) constant targets in one cluster (y_std falls back to 1, EI of every candidate there is pure sigma * pdf(0));
    (b) every cluster's incumbent (its best sample) is itself a candidate; (c) the whole candidate set duplicated so that
    both halves (= the shards of a 2-rank run) hold the same maximum: the lowest index must win in both scans."""
    import torch

    from cgp_b200.engine import ClusterGPEngine

    rng = np.random.default_rng(12)
    n, d = 700, 2
    X = rng.uniform(size=(n, d))
    lab = (X[:, 1] > 0.5).astype(np.int64)
    y = np.where(lab == 0, 3.0, np.cos(6 * X[:, 0]) + 0.01 * rng.standard_normal(n))  # cluster 0: exactly flat
    eng = ClusterGPEngine(X, y, lab, "matern32", 1e-5).set_theta(np.array([[-1.0, -1.0, -5.0], [-1.2, -0.7, -6.0]]))
    inc = np.stack([X[lab == c][np.argmax(y[lab == c])] for c in range(2)])
    q = np.concatenate([rng.uniform(size=(15000, d)), inc])
    Q = torch.from_numpy(np.concatenate([q, q])).cuda()
    score, idx = _same(eng, Q)
    assert idx < q.shape[0]  # the first copy wins the exact tie
    # shard-wise: each half alone picks the same local candidate with the same score bits
    (s0, i0, _), _ = eng.ei_scan(Q[: q.shape[0]], 3, prune=True)
    (s1, i1, _), _ = eng.ei_scan(Q[q.shape[0]:], 3, prune=True, idx_base=q.shape[0])
    assert int(i0.item()) == idx and int(i1.item()) == idx + q.shape[0] and float(s0.item()) == float(s1.item()) == score
