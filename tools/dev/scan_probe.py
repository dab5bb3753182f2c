"""Pruned-scan probe on ONE giant cluster (configs[3] shape at n = 32768): time of select_next for M candidates.
Env CGP_SCORE_CHUNK_MIN=0 restores the 1 GB chunks."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cgp_b200.engine import ClusterGPEngine  # noqa: E402
from tools import workloads as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1048576
X, y, lab, Q = W.cfg4(n, M)
eng = ClusterGPEngine(X, y, lab, "matern32", 1e-5, mem_fraction=0.9).set_theta(np.array([[-1.2, -0.9, 0.3, 0.1, -7.0]]))
Qd = torch.from_numpy(Q).cuda()
out = {}
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x, score, idx, cl = eng.select_next(Qd, k=3, sharded=True)
    torch.cuda.synchronize()
    out["scan_s"] = time.perf_counter() - t0
out.update(n=n, M=M, idx=idx, score=score, env={k: v for k, v in os.environ.items() if k.startswith("CGP_")})
print(json.dumps(out))
