"""One rank's share of the configs[2] fit at 8 GPUs, emulated on ONE GPU: 8 clusters x (R+1 = 2) runs = 16 runs
(a rank of 8 holds 17).  Prints fit time / evaluations for the current environment (CGP_LANES, CGP_TILE_MODE, ...)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cgp_b200.engine import ClusterGPEngine  # noqa: E402
from tools import workloads as W  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1
X, y, lab, _ = W.cfg3(32768, 6, 8, 0)
best = None
for rep in range(3):
    eng = ClusterGPEngine(X, y, lab, "matern32", 1e-5)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.fit(R, seed=0)
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    if best is None or t < best[0]:
        best = (t, eng.fit_stats["evals"], eng.fit_stats["lockstep"])
env = {k: v for k, v in os.environ.items() if k.startswith("CGP_")}
print(json.dumps(dict(env=env, R=R, fit_s=round(best[0], 4), evals=best[1], steps=best[2], ms_per_eval=round(1e3 * best[0] / best[1], 3))))
