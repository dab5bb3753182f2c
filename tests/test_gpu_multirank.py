"""Two ranks over NCCL on two GPUs of one box: the sharded fit (asynchronous runs + work stealing, and the two older
modes) and the sharded candidate scan select the SAME candidate with the SAME score bits and the same thetas as one rank
(REF/cGP/cGP_parallel.py:396-407 is the reference's precedent for treating clusters as independent work).
Skipped with fewer than two GPUs (run with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _problem():
    rng = np.random.default_rng(17)
    n, d, C = 1500, 3, 3
    X = rng.uniform(size=(n, d))
    lab = np.minimum((X[:, 0] * C).astype(np.int64), C - 1)
    y = lab + np.sin(2 * np.pi * (lab + 1) * X.sum(1) / d) + 0.01 * rng.standard_normal(n)
    Q = rng.uniform(size=(30011, d))
    return X, y, lab, Q


def _run(mode, R=5):
    from cgp_b200.engine import ClusterGPEngine

    X, y, lab, Q = _problem()
    eng = ClusterGPEngine(X, y, lab, "matern32", 1e-5)
    eng.fit(R, seed=0, shard_mode=mode)
    x, score, idx, cl = eng.select_next(Q, k=3)
    xe, se, ie, ce = eng.select_next(Q, k=3, prune=False)
    assert (idx, cl) == (ie, ce) and score == se
    xm, vm, im, cm = eng.select_next_mspe(Q, k=3)  # the second acquisition, candidates sharded the same way
    return dict(mspe=(vm, im, cm, xm.tolist()), theta=eng.theta.copy(), lml=eng.model["lml_host"].copy(), x=x, score=score, idx=idx, cluster=cl,
                evals=eng.fit_stats["evals"], nfev=eng.fit_stats["nfev"].copy(), stats={k: v for k, v in eng.fit_stats.items()
                                                                                       if k in ("stolen", "given", "lanes")})


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        out = {m: _run(m) for m in ("async", "balanced", "static")}
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_two_ranks_select_the_same_sample_as_one():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ref = _run("async")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank in (0, 1):
        for mode, got in res[rank].items():
            assert np.array_equal(got["theta"], ref["theta"]), (rank, mode)
            assert np.array_equal(got["lml"], ref["lml"]) and np.array_equal(got["nfev"], ref["nfev"])
            assert got["idx"] == ref["idx"] and got["cluster"] == ref["cluster"] and got["score"] == ref["score"]
            assert np.array_equal(got["x"], ref["x"])
            assert got["mspe"] == ref["mspe"], (rank, mode, got["mspe"], ref["mspe"])
    # the work was really shared: no rank evaluated everything, together they evaluated exactly the single-rank count
    for mode in ("async", "balanced", "static"):
        e0, e1 = res[0][mode]["evals"], res[1][mode]["evals"]
        assert e0 + e1 == ref["evals"] and max(e0, e1) < 0.8 * ref["evals"], (mode, e0, e1, ref["evals"])


def test_single_rank_modes_agree():
    """async (lanes) == the lock-step driver on one GPU: same thetas, same evaluation counts, same selection."""
    a, b = _run("async"), _run("static")
    assert np.array_equal(a["theta"], b["theta"]) and np.array_equal(a["nfev"], b["nfev"])
    assert a["idx"] == b["idx"] and a["score"] == b["score"] and a["evals"] == b["evals"]


def test_engine_on_a_non_current_device():
    """Tensors / engines on a GPU that is not the process' current device: every C entry point runs on the handle's
    device and on that device's stream, and restores the caller's device."""
    import torch

    from cgp_b200.engine import ClusterGPEngine

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    torch.cuda.set_device(0)
    X, y, lab, Q = _problem()
    th = np.tile(np.array([-0.7, -0.7, -0.7, -4.0]), (3, 1))
    e0 = ClusterGPEngine(X, y, lab, "matern32", 1e-5, device=0).set_theta(th)
    e1 = ClusterGPEngine(X, y, lab, "matern32", 1e-5, device=1).set_theta(th)
    assert torch.cuda.current_device() == 0
    assert e1.model["L"].device.index == 1
    assert np.array_equal(e0.model["lml_host"], e1.model["lml_host"])
    m0, s0 = e0.predict(Q[:3000])
    m1, s1 = e1.predict(Q[:3000])
    assert m1.device.index == 1 and torch.equal(m0.cpu(), m1.cpu()) and torch.equal(s0.cpu(), s1.cpu())
    l0, g0, _ = e0.lml_grad([0, 1, 2], th)
    l1, g1, _ = e1.lml_grad([0, 1, 2], th)
    assert np.array_equal(l0, l1) and np.array_equal(g0, g1)
    assert torch.cuda.current_device() == 0
