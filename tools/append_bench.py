"""Incremental refit at config-3 scale on one B200: append one sample to a cluster of ~4096 vs refactorising.
    python tools/append_bench.py > gpurun_out/append_bench.json"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import PRESETS, make_workload  # noqa: E402
from cgp_b200 import _lib  # noqa: E402
from cgp_b200.engine import ClusterGPEngine  # noqa: E402


def main():
    p = PRESETS["cfg3"]
    X, y, labels, _ = make_workload(p)
    eng = ClusterGPEngine(X, y, labels, "matern32", 1e-5)
    th = np.tile(np.array([-0.7] * p["d"] + [-4.0]), (p["C"], 1))
    eng.set_theta(th)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng._factorize()
    torch.cuda.synchronize()
    t_fac = time.perf_counter() - t0
    rng = np.random.default_rng(5)
    ts, tk = [], []
    for i in range(6):
        x = rng.uniform(size=p["d"])
        c = int(4 * (x[0] > .5) + 2 * (x[1] > .5) + (x[2] > .5))
        _lib.profile_read(reset=True)
        _lib.profile_enable(True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.append(x, float(c + np.sin(x.sum())), c)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
        prof = _lib.profile_read(reset=True)
        _lib.profile_enable(False)
        tk.append(prof["append"][0] + prof["cross"][0])
    n_c = int(eng.counts.max())
    print(json.dumps(dict(n_c=n_c, refactorize_all_clusters_s=t_fac, append_wall_s=ts, append_kernel_ms=tk,
                          append_hbm_gbs=[4 * 8.0 * n_c * n_c / 2 / (k * 1e-3) / 1e9 for k in tk],
                          note="append reads the lower triangle of W four times (l = W k, W^T l, z = W y, W^T z): 16 n^2 bytes")))


if __name__ == "__main__":
    main()
