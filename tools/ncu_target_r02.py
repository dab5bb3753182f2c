"""Bench-scale single-GPU target for ncu (round 2): the configs[2] workload, ONE batched LML+gradient evaluation of all
136 (cluster, restart) pairs (k_assemble grid = 136 pairs x 561 tiles), the k-NN routing of the 4 194 304 candidates
(k_knn_tc + exact finish) and one candidate chunk of the scan.
    python tools/ncu_target_r02.py [M]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgp_b200.engine import ClusterGPEngine, restart_thetas  # noqa: E402
from tools import workloads as W  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 4194304
X, y, lab, Q = W.cfg3(32768, 6, 8, M)
eng = ClusterGPEngine(X, y, lab, "matern32", 1e-5)
starts = restart_thetas(7, 16, 0)
starts[:, :] = np.clip(starts, -3.0, 2.0)  # PD at every start
starts[:, 6] = np.clip(starts[:, 6], -8.0, -2.0)
pc = np.repeat(np.arange(8, dtype=np.int32), 17)
th = np.tile(starts, (8, 1))
lml, grad, info = eng.lml_grad(pc, th, True)
print("pairs", len(pc), "non-PD", int(np.sum(info != 0)))
eng.set_theta(np.tile(np.array([-0.7] * 6 + [-4.0]), (8, 1)))
Qd = torch.from_numpy(Q).cuda()
ql = eng.route(Qd, 3)
(bs, bi, bc), _ = eng.ei_scan(Qd[:262144], 3, qlabel=ql[:262144])
torch.cuda.synchronize()
print("ok", float(lml[0]), int(bi.item()), np.bincount(ql.cpu().numpy()).tolist())
