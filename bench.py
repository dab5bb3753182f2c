"""Headline benchmark: one cGP fit + acquisition iteration (everything after clustering,
REF/cGP/cGP_constrained.py:287-379) on BASELINE.json configs[2]:
synthetic 6-D piecewise function, n = 32768 in 8 clusters, 16 L-BFGS restarts, 4M EI candidates.

    python bench.py --gpus 1 --steps 2 --warmup 3            # B200 arm (one process per GPU under torchrun for N>1)
    python bench.py --impl reference --steps 2 --warmup 1    # reference arm: scikit-learn/LAPACK on the host cores

A "step" = ClusterGPEngine(X, y, labels) -> fit (all clusters x 17 runs, lock-step L-BFGS-B) -> k-NN routing ->
predictive mean/variance -> EI -> argmax -> (N>1) all-gather of maxima + broadcast of the selected point.
metric value = candidates scanned per second of WHOLE iteration time (fit included); strong scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_REAL_STDOUT = None


def emit(line):
    """Print the ONE JSON line on the real stdout (library chatter such as NCCL's version banner is routed to stderr)."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


METRIC = "cGP fit+acq iteration (n=32k, 8 clusters, 16 restarts, 4M EI candidates): EI candidates/sec"
UNIT = "candidates/s"
# Total LML+grad evaluations of one config-3 fit (8 clusters x 17 L-BFGS-B runs), measured by the B200 arm
# (engine.fit_stats["evals"]); the optimiser is SciPy's own setulb fed by LML/gradients that agree with
# scikit-learn to ~1e-10, so the reference performs the same number.  Used only to extrapolate the CPU sample.
EVALS_PER_FIT_CFG3 = 4938

PRESETS = {
    "cfg3": dict(n=32768, d=6, C=8, R=16, M=4194304, k=3, name="configs[2]: 6-D piecewise, n=32768, 8 clusters, R=16, M=4194304"),
    "small": dict(n=4096, d=6, C=4, R=4, M=262144, k=3, name="smoke preset: n=4096, 4 clusters, R=4, M=262144"),
}


def make_workload(p):
    """X ~ default_rng(0).uniform[0,1]^6, r = 4[x0>.5]+2[x1>.5]+[x2>.5] (mod C),
    y = r + sin(2 pi (r+1) sum(x)/6) + 0.01 N(0,1), labels = r, candidates ~ default_rng(1).uniform."""
    rng = np.random.default_rng(0)
    X = rng.uniform(size=(p["n"], p["d"]))
    r = (4 * (X[:, 0] > .5) + 2 * (X[:, 1] > .5) + (X[:, 2] > .5)).astype(np.int64) % p["C"]
    y = r + np.sin(2 * np.pi * (r + 1) * X.sum(1) / 6) + 0.01 * rng.standard_normal(p["n"])
    Q = np.random.default_rng(1).uniform(size=(p["M"], p["d"]))
    return X, y, r, Q


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock / throttle reasons every 200 ms through NVML while the timed region runs."""

    def __init__(self, index=0):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.2)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU reference
def _ref_gpr(d, theta=None, R=0):
    """The reference's regressor, REF/cGP/cGP_constrained.py:169,315-316."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import Matern, WhiteKernel

    k = Matern(length_scale=np.ones(d), length_scale_bounds=(1e-05, 100000.0), nu=3 / 2) + \
        WhiteKernel(noise_level=1.0, noise_level_bounds=(1e-05, 100000.0))
    if theta is not None:
        k = k.clone_with_theta(theta)
    return GaussianProcessRegressor(kernel=k, random_state=0, normalize_y=True, alpha=1e-5,
                                    optimizer=None if theta is not None else "fmin_l_bfgs_b", n_restarts_optimizer=R)


def cpu_sample(p, X, y, labels, Q, evals_per_fit, m_sub=8192, n_eval=1):
    """Bounded sample of the reference CPU path on this host: `n_eval` LML+gradient evaluations of cluster 0 at
    full n_c through scikit-learn, plus k-NN routing + predict(return_std) + EI on `m_sub` candidates; extrapolated
    to the whole iteration (evals_per_fit evaluations spread evenly over clusters; M candidates)."""
    import warnings

    from sklearn.neighbors import KNeighborsClassifier
    from threadpoolctl import threadpool_info

    import oracle

    d = p["d"]
    theta = np.array([-0.7] * d + [-4.0])
    t_eval = []
    sizes = np.bincount(labels)
    c0 = 0
    Xt, yt = X[labels == c0], y[labels == c0].reshape(-1, 1)
    g = _ref_gpr(d, theta)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g.fit(Xt, yt)
        for _ in range(n_eval):
            t0 = time.perf_counter()
            g.log_marginal_likelihood(theta + 0.01, eval_gradient=True)
            t_eval.append(time.perf_counter() - t0)
    clf = KNeighborsClassifier(n_neighbors=p["k"]).fit(X, labels)
    Qs = Q[:m_sub]
    t0 = time.perf_counter()
    lab = clf.predict(Qs)
    t_knn = time.perf_counter() - t0
    t0 = time.perf_counter()
    mu, sd = g.predict(Qs, return_std=True)
    oracle.ei_scores(mu, sd, float(yt.max()), Xt.shape[0])
    t_pred = time.perf_counter() - t0
    # cost model: LML eval ~ n_c^3, predict ~ n_c^2 per candidate; clusters here are equal-sized to +-2 %
    s_eval = float(np.mean(t_eval))
    scale3 = float(np.mean((sizes / sizes[c0]) ** 3))
    scale2 = float(np.sum((sizes / sizes[c0]) ** 2 * (np.bincount(lab, minlength=len(sizes)) / max(len(lab), 1))))
    fit_s = s_eval * evals_per_fit * scale3
    acq_s = (t_knn + t_pred * max(scale2, 1e-9)) * (p["M"] / m_sub)
    threads = max([int(i.get("num_threads", 1)) for i in threadpool_info()] + [1])
    return dict(value=p["M"] / (fit_s + acq_s), unit=UNIT, cores=threads, host_cpus=os.cpu_count(),
                s_per_lml_grad_eval=s_eval, knn_s=t_knn, predict_ei_s=t_pred, fit_s_extrapolated=fit_s,
                acq_s_extrapolated=acq_s, iteration_s_extrapolated=fit_s + acq_s,
                sample=(f"{n_eval} sklearn log_marginal_likelihood(eval_gradient=True) at n_c={Xt.shape[0]} x "
                        f"{evals_per_fit} evaluations/fit (count measured by the B200 arm, same SciPy L-BFGS-B) + "
                        f"KNeighborsClassifier.predict / GPR.predict(return_std) / EI on {m_sub} of {p['M']} candidates, "
                        "extrapolated linearly"))


def run_reference(args, p):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    X, y, labels, Q = make_workload(p)
    vals = []
    last = None
    for i in range(args.warmup + args.steps):
        last = cpu_sample(p, X, y, labels, Q, EVALS_PER_FIT_CFG3 if args.preset == "cfg3" else 17 * p["C"] * 30,
                          m_sub=4096 if i < args.warmup else 8192)
        if i >= args.warmup:
            vals.append(last["iteration_s_extrapolated"])
    it = float(np.mean(vals))
    v = p["M"] / it
    last["kind"] = "reference"
    last["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": it * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": p["name"], "timing": "host wall clock of a bounded sample, extrapolated (see cpu_baseline.sample)"},
            "cpu_baseline": last,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args, p):
    import torch
    import torch.distributed as dist

    from cgp_b200 import _lib
    from cgp_b200.engine import ClusterGPEngine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    X, y, labels, Q = make_workload(p)
    lo, hi = (p["M"] * rank) // world, (p["M"] * (rank + 1)) // world
    Q_dev = torch.from_numpy(Q[lo:hi]).to(dev)
    Q_pin = torch.from_numpy(Q[lo:hi]).pin_memory()
    del Q
    _lib.handle(local)
    stats = {}

    def step(q_src, from_host):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        q = q_src.to(dev, non_blocking=True) if from_host else q_src
        eng = ClusterGPEngine(X, y, labels, "matern32", 1e-5)
        eng.fit(p["R"], seed=0)
        e[1].record()
        x_next, score, idx, cl = eng.select_next(q, k=p["k"], sharded=True)
        e[2].record()
        torch.cuda.synchronize()
        stats.update(fit_s=e[0].elapsed_time(e[1]) * 1e-3, acq_s=e[1].elapsed_time(e[2]) * 1e-3, evals=eng.fit_stats["evals"],
                     lockstep=eng.fit_stats["lockstep"], eval_s=eng.fit_stats["eval_s"], opt_host_s=eng.fit_stats["optimizer_host_s"], evals_per_cluster=eng.fit_stats["evals_per_cluster"].tolist(), nfev_sorted=sorted(int(v) for v in eng.fit_stats["nfev"]), x_next=x_next.tolist(), score=score, idx=idx, cluster=cl,
                     sizes=eng.counts.tolist(), lml=eng.model["lml_host"].tolist())
        stats["_engine"] = eng
        return x_next

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, q_src, from_host):
        barrier()
        t0 = time.perf_counter()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(nsteps):
            step(q_src, from_host)
        b.record()
        barrier()
        dt = torch.tensor([a.elapsed_time(b) * 1e-3, time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return float(dt[0]), float(dt[1])

    for i in range(args.warmup):
        # the last warm-up step runs with per-launch event timing on, so that the event pool of the C library is
        # created here and not inside the timed region
        _lib.profile_enable(2 if i == args.warmup - 1 else 0, local)
        step(Q_dev, False)
    _lib.profile_enable(False, local)
    # fp64 tensor peak measured here (cuBLAS DGEMM through torch; not in MEASURED_PEAKS.json, which has bf16 only)
    A = torch.randn((8192, 8192), dtype=torch.float64, device=dev)
    Bm = torch.randn((8192, 8192), dtype=torch.float64, device=dev)
    Cm = torch.empty_like(A)
    best = 1e9
    for _ in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        torch.matmul(A, Bm.T, out=Cm)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    fp64_peak = 2 * 8192.0 ** 3 / best / 1e12
    del A, Bm, Cm
    torch.cuda.empty_cache()

    sampler = ClockSampler(local)
    _lib.profile_read(reset=True, device=local)
    _lib.profile_enable(2, local)  # events around the large kernels only; every launch is still counted
    sampler.start()
    t_dev, t_wall = timed(args.steps, Q_dev, False)
    clocks = sampler.stop()
    prof = _lib.profile_read(reset=True, device=local)
    _lib.profile_enable(False, local)
    step_s = t_dev / args.steps
    res = dict(stats)
    t_e2e, _ = timed(args.steps, Q_pin, True)
    e2e_s = t_e2e / args.steps
    # outside every timed region: the exact bound-pruned scan on the same fitted model (reported beside the exhaustive one)
    eng = stats.pop("_engine")
    res.pop("_engine", None)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _xp, score_p, idx_p, cl_p = eng.select_next(Q_dev, k=p["k"], sharded=True, prune=True)
    b.record()
    barrier()
    tp = torch.tensor([a.elapsed_time(b) * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    pruned = {"acq_s": float(tp[0]), "acq_candidates_per_s": p["M"] / float(tp[0]),
              "same_selection": bool(idx_p == res["idx"] and score_p == res["score"] and cl_p == res["cluster"]),
              "note": "cgp_ei_argmax_pruned: mean-only EI bound, survivors scored exactly; NOT used in the timed region"}
    del eng

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel class (CUDA events on the launch stream, inside the timed region)
    sizes = np.asarray(res["sizes"], dtype=np.float64)
    ev_c = np.asarray(res["evals_per_cluster"], dtype=np.float64) * args.steps  # this rank's evaluations, timed region
    m_local = hi - lo
    PANEL, TILE = 4, 128
    alg = {"potrf_trail": 0.0, "trtri_trail": 0.0, "lauum": 0.0, "predict_vnorm": 0.0}
    for n_c, e_c in zip(sizes, ev_c):
        T = int(-(-n_c // TILE))
        for c0 in range(0, T, PANEL):
            c1 = min(T, c0 + PANEL)
            nt = max(0.0, n_c - TILE * c1)  # trailing order after this panel
            K = TILE * (c1 - c0)
            alg["potrf_trail"] += e_c * K * nt * (nt + 1)  # SYRK: 2 K nt(nt+1)/2
            ksum = sum(TILE * (c1 - max(j, c0)) for j in range(c1))
            alg["trtri_trail"] += e_c * 2.0 * nt * TILE * ksum
        alg["lauum"] += e_c * n_c ** 3 / 3.0
    alg["predict_vnorm"] = args.steps * m_local * float(np.mean(sizes ** 2))
    # k_lauum and k_predict_vnorm run alone on the caller's stream: their event brackets are exclusive kernel time.
    # The trailing-update kernels of the Cholesky / triangular inverse run on three streams at once (chain, bulk,
    # inverse: look-ahead schedule), so their brackets include the time they share the SMs with the other two streams;
    # they are listed (overlapped=true) but the roofline kernel is taken from the exclusive ones.
    exclusive = {"lauum", "predict_vnorm"}
    roof_all = {}
    for k, fl in alg.items():
        ms, cnt = prof[k]
        if ms > 0:
            roof_all[k] = {"ms": ms, "launches": cnt, "achieved": fl / (ms * 1e-3) / 1e12, "share_of_step": ms * 1e-3 / t_dev,
                           "overlapped": k not in exclusive}
    fit_wall = res["fit_s"] * args.steps
    fit_unit = {"s": fit_wall, "achieved": float(np.sum(ev_c * sizes ** 3)) / fit_wall / 1e12,
                "note": "whole LML+gradient unit: n_c^3 algorithmic flops per evaluation (potrf + inverse + K^-1) over the "
                        "fit wall time of the timed region, all kernels, host L-BFGS-B and final factorisation included"}
    top = max((k for k in roof_all if k in exclusive), key=lambda k: roof_all[k]["ms"])
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(top)
    roofline = {"bound": "tensor", "kernel": "k_" + top, "achieved": roof_all[top]["achieved"], "peak": fp64_peak,
                "unit": "TFLOP/s", "frac": roof_all[top]["achieved"] / fp64_peak, "traffic": traffic,
                "peak_source": "cuBLAS DGEMM 8192^3 fp64 measured in this run (MEASURED_PEAKS.json has no fp64 entry)",
                "algorithmic_flops": "potrf_trail: K n_t(n_t+1) per rank-K SYRK of the order-n_t trailing block; lauum: n_c^3/3; "
                                     "trtri_trail: 2 n_t 128 sum_j K_j; predict_vnorm: n_c^2 per candidate",
                "all": roof_all, "lml_grad_unit": fit_unit}
    gpu_launches = int(sum(c for _, c in prof.values()))
    kernel_ms = {k: round(v[0], 3) for k, v in prof.items() if v[1] and v[0] > 0}  # large kernels only (profiling level 2)
    kernel_launches = {k: v[1] for k, v in prof.items() if v[1]}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        Xw, yw, lw, Qw = make_workload(p)
        cpu = cpu_sample(p, Xw, yw, lw, Qw, res["evals"])
        cpu["kind"] = "port"
        cpu["note"] = ("scikit-learn 1.9.0 / SciPy / OpenBLAS objects configured exactly as the reference does "
                       "(the arithmetic the reference delegates to) + oracle EI restatement")
    line = {"metric": METRIC, "value": p["M"] / step_s, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": p["name"], "parallelism": f"pairs+candidates sharded over {world} GPU(s)",
                       "l2": "working set (36 GB of factor workspaces, 1 GB K* chunks) >> 126 MB L2; no flush needed",
                       "timing": "CUDA events around K whole steps, barrier+synchronize both sides, max over ranks"},
            "iteration_s": step_s, "fit_s": res["fit_s"], "acq_s": res["acq_s"],
            "acq_candidates_per_s": p["M"] / res["acq_s"], "lml_evals_per_fit_rank0": res["evals"],
            "lockstep_per_fit": res["lockstep"], "fit_eval_wall_s": res["eval_s"], "fit_optimizer_host_s": res["opt_host_s"], "nfev_per_pair_rank0_sorted": res["nfev_sorted"], "selected": {"idx": res["idx"], "score": res["score"], "cluster": res["cluster"],
                                                              "x": res["x_next"]},
            "wall_s_per_step": t_wall / args.steps,
            "e2e": {"value": p["M"] / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": int(Q_pin.numel() * 8 + X.nbytes + y.nbytes + labels.nbytes),
                    "d2h_bytes_per_step": int((p["d"] + 2) * 8), "ms_per_step": e2e_s * 1e3},
            "acq_pruned_scan": pruned,
            "gpu_launches": gpu_launches, "kernel_launches_timed_region": kernel_launches,
            "kernel_ms_timed_region": kernel_ms,
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--preset", default="cfg3", choices=sorted(PRESETS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    p = PRESETS[args.preset]
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, p)
    else:
        run_b200(args, p)


if __name__ == "__main__":
    main()
