"""On-box measurements that feed DESIGN.md / bench.py roofline denominators:
fp64 peak (cuBLAS DGEMM through torch, test-only), our DMMA tile GEMM, HBM copy, and per-phase timings
of the hot path at config-3-like sizes.  Usage: python tools/microbench.py [--quick] > gpurun_out/microbench.json"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgp_b200 import _lib  # noqa: E402
from cgp_b200.engine import ClusterGPEngine  # noqa: E402


def timeit(fn, warm=2, rep=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(rep):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts), float(np.median(ts))


def main():
    quick = "--quick" in sys.argv
    out = {"gpu": torch.cuda.get_device_name(0)}
    _lib.handle()
    N = 4096 if quick else 8192
    A = torch.randn((N, N), dtype=torch.float64, device="cuda")
    B = torch.randn((N, N), dtype=torch.float64, device="cuda")
    C = torch.empty((N, N), dtype=torch.float64, device="cuda")
    fl = 2.0 * N ** 3
    t, tm = timeit(lambda: torch.matmul(A, B.T, out=C))
    out["cublas_dgemm_tflops"] = {"best": fl / t / 1e12, "median": fl / tm / 1e12, "N": N}
    t, tm = timeit(lambda: _lib.dgemm_nt(A, B, C))
    out["dmma_gemm_nt_tflops"] = {"best": fl / t / 1e12, "median": fl / tm / 1e12, "N": N}
    ref = torch.matmul(A[:256], B[:256].T)
    _lib.dgemm_nt(A, B, C)
    out["dmma_gemm_maxerr"] = float((C[:256, :256] - ref).abs().max())
    x = torch.empty(1 << 28, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    t, _ = timeit(lambda: y.copy_(x))
    out["hbm_copy_gbs"] = 2 * x.numel() * 8 / t / 1e9
    del x, y, A, B, C

    # ---- hot path phases at config-3-like shape (one GPU): C clusters of n_c, d = 6
    rng = np.random.default_rng(0)
    Cn, nc, d = (2, 2048, 6) if quick else (8, 4096, 6)
    n = Cn * nc
    X = rng.uniform(size=(n, d))
    r = (4 * (X[:, 0] > .5) + 2 * (X[:, 1] > .5) + (X[:, 2] > .5)).astype(np.int64) % Cn
    yv = r + np.sin(2 * np.pi * (r + 1) * X.sum(1) / 6) + 0.01 * rng.standard_normal(n)
    # equal-size clusters by construction: relabel by rank inside sorted order
    order = np.argsort(r, kind="stable")
    labels = np.empty(n, dtype=np.int64)
    labels[order] = np.repeat(np.arange(Cn), nc)
    eng = ClusterGPEngine(X, yv, labels, "matern32", 1e-5)
    for B in ([1, 4] if quick else [1, 8, 17 * 8]):
        pc = np.arange(B, dtype=np.int32) % Cn
        th = rng.uniform(-1.5, 0.5, size=(B, d + 1))
        th[:, d] = rng.uniform(-6, -2, size=B)
        t0 = time.perf_counter()
        eng.lml_grad(pc, th, True)
        torch.cuda.synchronize()
        t, tm = timeit(lambda: eng.lml_grad(pc, th, True), warm=1, rep=3)
        flops = B * float(nc) ** 3  # n^3/3 potrf + n^3/3 trtri + n^3/3 lauum
        _lib.profile_read(reset=True)
        _lib.profile_enable(True)
        eng.lml_grad(pc, th, True)
        prof = {k: round(v[0], 3) for k, v in _lib.profile_read(reset=True).items() if v[1]}
        _lib.profile_enable(False)
        out[f"lml_grad_B{B}_n{nc}"] = {"s": t, "tflops_alg": flops / t / 1e12, "first_call_s": time.perf_counter() - t0,
                                       "class_ms": prof}
    th = np.tile(np.array([-0.7] * d + [-4.0]), (Cn, 1))
    eng.set_theta(th)
    t, _ = timeit(lambda: eng._factorize(), warm=1, rep=3)
    out["factorize_s"] = t
    M = 1 << 16 if quick else 1 << 19
    Q = torch.from_numpy(rng.uniform(size=(M, d))).cuda()
    t, _ = timeit(lambda: eng.route(Q, 3), warm=1, rep=3)
    out["knn"] = {"s": t, "M": M, "n": n, "pairs_per_s": M * n / t, "fp64_flops": M * n * 3 * d / t / 1e12}
    ql = eng.route(Q, 3)
    t, _ = timeit(lambda: eng.ei_scan(Q, 3, qlabel=ql), warm=1, rep=3)
    out["ei_scan"] = {"s": t, "M": M, "cand_per_s": M / t, "tflops_alg": M * float(nc) ** 2 / t / 1e12}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
