"""Short single-GPU program for ncu: one batched LML+gradient evaluation (B pairs, n_c = 2048, d = 6), one final
factorisation, one exhaustive and one pruned EI scan — every kernel of the hot path launches at least once.
    python tools/ncu_target.py [B] [n_c] [M]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgp_b200.engine import ClusterGPEngine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
nc = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
M = int(sys.argv[3]) if len(sys.argv) > 3 else 32768
Cn, d = 2, 6
rng = np.random.default_rng(0)
X = rng.uniform(size=(Cn * nc, d))
labels = np.repeat(np.arange(Cn), nc)
y = labels + np.sin(5 * X.sum(1)) + 0.01 * rng.standard_normal(Cn * nc)
eng = ClusterGPEngine(X, y, labels, "matern32", 1e-5)
th = rng.uniform(-1.5, 0.5, size=(B, d + 1))
th[:, d] = rng.uniform(-6, -2, size=B)
lml, grad, info = eng.lml_grad(np.arange(B, dtype=np.int32) % Cn, th, True)
assert np.all(info == 0) and np.all(np.isfinite(lml))
eng.set_theta(np.tile(np.array([-0.7] * d + [-4.0]), (Cn, 1)))
Q = rng.uniform(size=(M, d))
(bs, bi, bc), _ = eng.ei_scan(Q, 3)
(ps, pi, pc), _ = eng.ei_scan(Q, 3, prune=True)  # exact bound-pruned scan: k_prune_select, k_predict_vnorm_sel, ...
torch.cuda.synchronize()
assert int(pi.item()) == int(bi.item()) and float(ps.item()) == float(bs.item())
print("ok", float(lml[0]), int(bi.item()), float(bs.item()))
