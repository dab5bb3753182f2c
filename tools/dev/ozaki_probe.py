"""Feasibility probe for DESIGN.md §8 (NOT product code, nothing in cgp_b200/ uses it): how fast and how accurate would
the n_c^3 work of a fit be if the fp64 products ran as error-free int8 slice products (Ozaki scheme) on the int8
tensor cores instead of DMMA?  Library kernels only (torch.matmul fp64 = cuBLAS DGEMM, torch._int_mm = cuBLASLt int8
GEMM): this measures the ceiling a hand-written tcgen05 kind::i8 kernel could aim at, before anybody writes it.

    python tools/dev/ozaki_probe.py [n] > gpurun_out/ozaki_probe.json

C = A B^T with A, B [n, n] fp64.  Row i of A is scaled by 2^-e_i (e_i = exponent of its largest entry), then cut into S
signed slices of 6 bits (round to nearest, |slice| <= 64); likewise B.  Slice products are exact in int32
(64*64*n <= 2^29 for n <= 131072 per product; anti-diagonal sums of <= S products stay below 2^31 for n <= 4096 * 16 / S).
C ~ sum over anti-diagonals g = t + u <= S + 1 of 2^(-6 g) * (sum_{t+u=g} A_t B_u^T), times 2^(e_i + f_j).
"""
import json
import sys
import time
from fractions import Fraction

import numpy as np
import torch


def split(A, S):
    """-> (list of S int8 [n, k] slices, exponents e [n]) with A ~ 2^e * sum_t 2^(-6 t) slice_t."""
    amax = A.abs().amax(dim=1).clamp_min(1e-300)
    e = torch.ceil(torch.log2(amax))
    r = A * torch.exp2(-e)[:, None]  # in [-1, 1]
    out = []
    for _ in range(S):
        r = r * 64.0
        s = torch.round(r)
        r = r - s
        out.append(s.to(torch.int8))
    return out, e


def ozaki_mm(A, B, S, time_only=False):
    As, ea = split(A, S)
    Bs, eb = split(B, S)
    n = A.shape[0]
    C = torch.zeros((n, B.shape[0]), dtype=torch.float64, device=A.device)
    n_prod = 0
    for g in range(S + 1, 1, -1):  # smallest contributions first
        P = None
        for t in range(1, S + 1):
            u = g - t
            if 1 <= u <= S:
                p = torch._int_mm(As[t - 1], Bs[u - 1].t())
                n_prod += 1
                P = p if P is None else P + p
        C += P.to(torch.float64) * (2.0 ** (-6 * g))
    C *= torch.exp2(ea)[:, None] * torch.exp2(eb)[None, :]
    return C, n_prod


def timed(fn, rep=3):
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
    return min(ts)


def exact_entries(A, B, idx):
    """Exact dot products (rational arithmetic) of a few (i, j) entries."""
    A, B = A.cpu().numpy(), B.cpu().numpy()
    out = []
    for i, j in idx:
        s = Fraction(0)
        for x, y in zip(A[i], B[j]):
            s += Fraction(float(x)) * Fraction(float(y))
        out.append(float(s))
    return np.array(out)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    out = {"n": n, "gpu": torch.cuda.get_device_name(0)}
    # (1) raw library rates
    A = torch.randn((n, n), dtype=torch.float64, device=dev, generator=g)
    B = torch.randn((n, n), dtype=torch.float64, device=dev, generator=g)
    t64 = timed(lambda: torch.matmul(A, B.t()))
    a8 = torch.randint(-64, 65, (n, n), dtype=torch.int8, device=dev)
    b8 = torch.randint(-64, 65, (n, n), dtype=torch.int8, device=dev)
    t8 = timed(lambda: torch._int_mm(a8, b8.t()))
    out["dgemm_s"] = t64
    out["dgemm_tflops"] = 2.0 * n ** 3 / t64 / 1e12
    out["int8_gemm_s"] = t8
    out["int8_gemm_tops"] = 2.0 * n ** 3 / t8 / 1e12
    # (2) accuracy / time of the slice scheme on three operand classes
    rng = np.random.default_rng(1)
    idx = [(int(rng.integers(n)), int(rng.integers(n))) for _ in range(24)]
    cases = {}
    scale = torch.exp2(torch.randint(-20, 21, (n,), device=dev, generator=g).double())
    # a Cholesky-factor-like operand: L of a Matern-3/2 kernel matrix with jitter 1e-5 (what the trailing updates multiply)
    X = torch.rand((n, 6), dtype=torch.float64, device=dev, generator=g)
    D = torch.cdist(X / 0.5, X / 0.5)
    K = (1 + 3 ** 0.5 * D) * torch.exp(-(3 ** 0.5) * D) + 1.001e-3 * torch.eye(n, dtype=torch.float64, device=dev)
    L = torch.linalg.cholesky(K)
    for name, (P, Q) in {"gaussian": (A, B), "rows_scaled_2^+-20": (A * scale[:, None], B), "cholesky_factor": (L, L)}.items():
        ref = exact_entries(P, Q, idx)
        c64 = torch.matmul(P, Q.t())
        norm = np.array([float(P[i].abs() @ Q[j].abs()) for i, j in idx])  # |a_i|.|b_j|: the scale of the fp64 error bound
        res = {"dgemm_err_over_|a||b|": float(np.max(np.abs(np.array([float(c64[i, j]) for i, j in idx]) - ref) / norm))}
        for S in (6, 7, 8, 9, 10):
            c, n_prod = ozaki_mm(P, Q, S)
            err = float(np.max(np.abs(np.array([float(c[i, j]) for i, j in idx]) - ref) / norm))
            t = timed(lambda: ozaki_mm(P, Q, S), rep=2) if name == "gaussian" else None
            res[f"S={S}"] = {"int8_products": n_prod, "err_over_|a||b|": err, "s_total_torch_glue": t,
                             "s_int8_products_only": n_prod * t8, "speedup_products_only_vs_dgemm": t64 / (n_prod * t8)}
        cases[name] = res
    out["cases"] = cases
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
