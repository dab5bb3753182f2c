"""configs[3] memory logic: two engines of ONE 65 536-sample cluster in a row (the second must find room although the
first one's 69 GB of factors and 91 GB workspace sit in torch's allocator cache), each: one lock-step of 2 runs + final
factorisation + a small pruned scan."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cgp_b200 import CGPRegressor  # noqa: E402
from tools import workloads as W  # noqa: E402

X, y, lab, Q = W.cfg4(65536, 65536)
Qd = torch.from_numpy(Q).cuda()
model = None
for rep in range(2):
    model = None
    t0 = time.perf_counter()
    model = CGPRegressor("matern32", alpha=1e-5, n_restarts_optimizer=1, random_state=0, max_lockstep=1, mem_fraction=0.9)
    model.fit(X, y, lab)
    x, s, i = model.select_next(Qd, sharded=True)
    torch.cuda.synchronize()
    print(rep, round(time.perf_counter() - t0, 2), "s", i, s, "reserved GB", round(torch.cuda.memory_reserved() / 2**30, 1),
          "allocated GB", round(torch.cuda.memory_allocated() / 2**30, 1), flush=True)
