"""Round-1 tool (superseded by `bench.py --preset cfg2|cfg4|cfg5`): timings of the other BASELINE.json configs on ONE B200.
    python tools/config_runs.py cfg1 | cfg2 | cfg5 [--m M] | cfg4 [--n N --m M]
Data definitions: SURVEY.md §8d."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgp_b200 import _lib  # noqa: E402
from cgp_b200.engine import ClusterGPEngine  # noqa: E402


def sync_time(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return time.perf_counter() - t0, out


def arg(name, default):
    return type(default)(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def cfg2():
    """Schaffer N2 2-D, N_PILOT=2000, DGM4 clusters (host), R=20, 64k candidates (tools/workloads.py)."""
    import oracle
    from cgp_b200.plugins import DGM4
    from tools import workloads as W

    X, y = W.cfg2_samples()
    t_cl = time.perf_counter()
    lab = DGM4.f_Cluster(123).fit_predict(np.concatenate([X, y[:, None]], axis=1))
    lab = oracle.recode(oracle.merge_small_clusters(lab, 2))
    t_cl = time.perf_counter() - t_cl
    return X, y, lab, W.cfg2_candidates(), 20, dict(host_clustering_s=t_cl)


def cfg5():
    """8-D, 64 clusters of n_c in [480, 544], R=32, 16M candidates (tools/workloads.py)."""
    from tools import workloads as W

    X, y, lab, Q = W.cfg5(64, arg("--m", 16777216))
    return X, y, lab, Q, 32, {}


def run_fit_scan(name, X, y, lab, Q, R, extra):
    out = dict(config=name, n=int(X.shape[0]), d=int(X.shape[1]), clusters=int(lab.max()) + 1, R=R, M=int(Q.shape[0]), **extra)
    Qd = torch.from_numpy(Q).cuda()
    for rep in range(2):  # second repetition is the warm one
        eng = ClusterGPEngine(X, y, lab, "matern32", 1e-5)
        t_fit, _ = sync_time(lambda: eng.fit(R, seed=0))
        t_knn, ql = sync_time(lambda: eng.route(Qd, 3))
        t_scan, res = sync_time(lambda: eng.ei_scan(Qd, 3, qlabel=ql))
    out.update(fit_s=t_fit, evals=eng.fit_stats["evals"], lockstep=eng.fit_stats["lockstep"],
               fit_eval_wall_s=eng.fit_stats["eval_s"], fit_optimizer_host_s=eng.fit_stats["optimizer_host_s"], knn_s=t_knn, scan_s=t_scan,
               iteration_s=t_fit + t_knn + t_scan, cand_per_s_acq=Q.shape[0] / (t_knn + t_scan),
               best=dict(score=float(res[0][0].item()), idx=int(res[0][1].item()), cluster=int(res[0][2].item())))
    print(json.dumps(out))


def cfg1():
    """configs[0] through the GPU-backend driver: `cGP.py -p 20 -s 10 -c 3 -n 3` on Bukin N6, seed 123
    (the reference's own CPU run of this command takes 10.7 s of loop time in the build container)."""
    import tempfile

    from cgp_b200 import driver

    with tempfile.TemporaryDirectory() as tmp:
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            driver.main(["-p", "20", "-s", "2", "-c", "3", "-n", "3", "-f", "cgp_b200.plugins.f_bukin", "-r", "123", "-X", "-o", "warm"])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            X, Y = driver.main(["-p", "20", "-s", "10", "-c", "3", "-n", "3", "-f", "cgp_b200.plugins.f_bukin", "-r", "123", "-X",
                                "-o", "run"])
            torch.cuda.synchronize()
            t = time.perf_counter() - t0
        finally:
            os.chdir(cwd)
    print(json.dumps(dict(config="cfg1", command="driver -p 20 -s 10 -c 3 -n 3 -f f_bukin -r 123 (DGM3 clustering on the host, R = 10 d = 20, "
                          "65536 candidates per step)", loop_s=t, s_per_iteration=t / 10, samples=int(X.shape[0]),
                          best_y=float(Y.max()))))


def cfg4():
    """single cluster, 4-D Matern-3/2: blocked DMMA Cholesky stress (one LML+grad evaluation, final factorisation, scan)."""
    n = arg("--n", 65536)
    M = arg("--m", 1048576)
    d = 4
    rng = np.random.default_rng(0)
    X = rng.uniform(size=(n, d))
    y = np.sin(4 * np.pi * X[:, 0]) * np.cos(2 * np.pi * X[:, 1]) + X[:, 2] * X[:, 3] + 1e-2 * rng.standard_normal(n)
    lab = np.zeros(n, dtype=np.int64)
    eng = ClusterGPEngine(X, y, lab, "matern32", 1e-5, mem_fraction=0.9)
    th = np.array([-1.2, -0.9, 0.3, 0.1, -7.0])
    t_eval, (lml, grad, info) = sync_time(lambda: eng.lml_grad([0], th[None, :], True))
    t_eval2, _ = sync_time(lambda: eng.lml_grad([0], th[None, :] + 0.01, True))
    t_fac, _ = sync_time(lambda: eng.set_theta(th[None, :]))
    Qd = torch.from_numpy(np.random.default_rng(1).uniform(size=(M, d))).cuda()
    t_knn, ql = sync_time(lambda: eng.route(Qd, 3))
    t_scan, res = sync_time(lambda: eng.ei_scan(Qd, 3, qlabel=ql))
    fl = float(n) ** 3
    print(json.dumps(dict(config="cfg4", n=n, d=d, M=M, lml=float(lml[0]), info=int(info[0]), grad=grad[0].tolist(),
                          lml_grad_eval_s=t_eval2, first_eval_s=t_eval, tflops_alg=fl / t_eval2 / 1e12, factorize_s=t_fac,
                          knn_s=t_knn, scan_s=t_scan, scan_tflops_alg=M * float(n) ** 2 / t_scan / 1e12,
                          cand_per_s=M / (t_knn + t_scan), best_idx=int(res[0][1].item()))))


if __name__ == "__main__":
    _lib.handle()
    which = sys.argv[1]
    if which == "cfg1":
        cfg1()
    elif which == "cfg2":
        run_fit_scan("cfg2", *cfg2())
    elif which == "cfg5":
        run_fit_scan("cfg5", *cfg5())
    else:
        cfg4()
