"""Where does a configs[1] iteration spend its host time?  (5 repetitions, then a cProfile of one.)"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from cgp_b200 import CGPRegressor  # noqa: E402

X, y, lab, Q = bench.make_workload("cfg2")
Qd = torch.from_numpy(Q).cuda()


def step():
    t0 = time.perf_counter()
    m = CGPRegressor("matern32", alpha=1e-5, n_restarts_optimizer=20, random_state=0, n_neighbors=3)
    m.fit(X, y, lab)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    m.select_next(Qd, sharded=True)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    fs = m.engine_.fit_stats
    return round((t1 - t0) * 1e3, 1), round((t2 - t1) * 1e3, 1), fs["lockstep"], round(fs["eval_s"] * 1e3, 1), round(fs["optimizer_host_s"] * 1e3, 1)


for i in range(6):
    print(step())
pr = cProfile.Profile()
pr.enable()
step()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
