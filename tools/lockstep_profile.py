"""Where does a lock-step go?  Per lock-step of one fit: active runs, host time to submit, wait for the device,
host time to advance the L-BFGS-B machines; plus per-kernel-class totals.
    python tools/lockstep_profile.py cfg2 | cfg5 | cfg3r8      (cfg3r8 = the pairs rank 0 of 8 owns in config 3)
    | cfg3 (all pairs, one repetition, also dumps nfev per pair)
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgp_b200 import _lib, dist as cdist  # noqa: E402
from cgp_b200.engine import ClusterGPEngine, restart_thetas, LOG_LO, LOG_HI  # noqa: E402
from cgp_b200.lbfgs import minimize_lockstep  # noqa: E402
from tools.config_runs import cfg2, cfg5  # noqa: E402


def cfg3():
    rng = np.random.default_rng(0)
    n, d = 32768, 6
    X = rng.uniform(size=(n, d))
    r = (4 * (X[:, 0] > .5) + 2 * (X[:, 1] > .5) + (X[:, 2] > .5)).astype(np.int64)
    y = r + np.sin(2 * np.pi * (r + 1) * X.sum(1) / 6) + 0.01 * rng.standard_normal(n)
    return X, y, r, None, 16, {}


def main():
    which = sys.argv[1]
    X, y, lab, _Q, R, _ = {"cfg2": cfg2, "cfg5": cfg5, "cfg3r8": cfg3, "cfg3": cfg3}[which]()
    eng = ClusterGPEngine(X, y, lab, "matern32", 1e-5)
    P, C = eng.P, eng.C
    lo, hi = np.full(P, LOG_LO), np.full(P, LOG_HI)
    starts = restart_thetas(P, R, 0, lo, hi)
    R1 = R + 1
    n_pairs = C * R1
    mine = cdist.my_pairs(n_pairs, 0, 8) if which == "cfg3r8" else list(range(n_pairs))
    pc_all = np.asarray([p // R1 for p in mine], dtype=np.int32)
    x0s = [starts[p % R1] for p in mine]
    out = dict(config=which, pairs=len(mine))
    prof = {}
    for rep in ([1] if which == "cfg3" else [0, 1]):
        steps = []
        spikes = []
        last = [time.perf_counter()]

        def eval_batch(idx, Xs):
            t0 = time.perf_counter()
            tk = eng.lml_grad_submit(pc_all[idx], Xs, True)
            t1 = time.perf_counter()
            lml, grad, _ = eng.lml_grad_collect(tk)
            t2 = time.perf_counter()
            steps.append([len(idx), (t0 - last[0]) * 1e3, (t1 - t0) * 1e3, (t2 - t1) * 1e3])
            if t1 - t0 > 3e-3:
                spikes.append([len(idx), round((t1 - t0) * 1e3, 2)])
            last[0] = t2
            return -lml, -grad

        _lib.profile_read(reset=True)
        _lib.profile_enable(rep == 0)  # class totals from the first repetition; the timed one runs unprofiled
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        last[0] = t0
        _xs, _fs, nfev, _nit, _st = minimize_lockstep(eval_batch, x0s, lo, hi)  # plain lock-step (no two-group pipelining) so the split is clean
        torch.cuda.synchronize()
        fit_s = time.perf_counter() - t0
        if rep == 0:
            prof = {k: [round(v[0], 2), v[1]] for k, v in _lib.profile_read(reset=True).items() if v[1]}
            _lib.profile_enable(False)
    s = np.asarray(steps)
    out.update(fit_s=fit_s, lockstep=len(steps), evals=int(s[:, 0].sum()), host_advance_ms=float(s[:, 1].sum()),
               host_submit_ms=float(s[:, 2].sum()), device_wait_ms=float(s[:, 3].sum()), class_ms=prof,
               steps_active_adv_submit_wait=[[int(a), round(b, 2), round(c, 2), round(d, 2)] for a, b, c, d in steps[::max(1, len(steps) // 40)]])
    out["submit_spikes_active_ms"] = spikes[:30]
    if which == "cfg3":
        out["nfev_by_pair"] = [int(v) for v in nfev]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
