"""Headline benchmark: one cGP fit + acquisition iteration (everything after clustering,
REF/cGP/cGP_constrained.py:287-379).  Default workload = BASELINE.json configs[2]: synthetic 6-D piecewise function,
n = 32768 in 8 clusters, 16 L-BFGS restarts, 4M EI candidates.

    python bench.py --gpus 1 --steps 2 --warmup 3            # B200 arm (one process per GPU under torchrun for N>1)
    python bench.py --impl reference --steps 2 --warmup 1    # reference arm: scikit-learn/LAPACK on the host cores
    python bench.py --preset cfg2|cfg4|cfg5|small            # the other BASELINE.json configs

A "step" = the reference-facing call sequence on one batch of inputs:
    CGPRegressor(...).fit(X, y, labels)   all clusters x (R+1) runs, lock-step L-BFGS-B, final factorisation
    .select_next(Q, sharded=True)         k-NN routing -> predictive mean/variance -> EI -> argmax
                                          -> (N>1) all-gather of maxima + broadcast of the selected point
The acquisition is the exact bound-pruned scan (the default of the public API: same selected candidate and score bits
as scoring every candidate); the exhaustive scan is timed beside it, outside the timed region.
metric value = candidates scanned per second of WHOLE iteration time (fit included); strong scaling.
`value`: candidate shard resident in HBM;  `e2e`: the same calls fed from pinned HOST buffers (H2D inside the region).
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

from tools import workloads as W  # noqa: E402

_REAL_STDOUT = None


def emit(line):
    """Print the ONE JSON line on the real stdout (library chatter such as NCCL's version banner is routed to stderr)."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


UNIT = "candidates/s"
PRESETS = {
    "cfg3": dict(d=6, R=16, M=4194304, k=3, metric="cGP fit+acq iteration (n=32k, 8 clusters, 16 restarts, 4M EI candidates): EI candidates/sec",
                 name="configs[2]: 6-D piecewise, n=32768, 8 clusters, R=16, M=4194304"),
    "cfg2": dict(d=2, R=20, M=65536, k=3, metric="cGP fit+acq iteration (Schaffer N2, N=2000, DGM4, 20 restarts, 64k EI candidates): EI candidates/sec",
                 name="configs[1]: Schaffer N2 2-D, N_PILOT=2000, 4 DGM clusters, R=20, M=65536"),
    "cfg4": dict(d=4, R=7, M=8388608, k=3, max_lockstep=1, mem_fraction=0.9,
                 metric="cGP iteration, ONE cluster n=65536 (one lock-step of 8 LML+grad evaluations + factorisation + 8M-candidate EI scan): EI candidates/sec",
                 name="configs[3]: single cluster n=65536, 4-D Matern-3/2, 8 (cluster, restart) pairs x ONE lock-step, M=8388608"),
    "cfg5": dict(d=8, R=32, M=16777216, k=3, metric="cGP fit+acq iteration (8-D, 64 clusters of ~512, 32 restarts, 16M EI candidates): EI candidates/sec",
                 name="configs[4]: 8-D, 64 clusters of n_c in [480,544], R=32, M=16777216"),
    "small": dict(d=6, R=4, M=262144, k=3, metric="cGP fit+acq iteration (smoke preset): EI candidates/sec",
                  name="smoke preset: n=4096, 4 clusters, R=4, M=262144"),
}
EVALS_FILE = os.path.join(ROOT, "profiles", "fit_evals.json")  # LML+grad evaluations per fit, measured by the B200 arm


def make_workload(preset, M=None):
    """(X, y, labels, Q) of a preset (tools/workloads.py)."""
    p = PRESETS[preset]
    M = p["M"] if M is None else M
    if preset == "cfg3":
        return W.cfg3(32768, 6, 8, M)
    if preset == "small":
        return W.cfg3(4096, 6, 4, M)
    if preset == "cfg4":
        return W.cfg4(65536, M)
    if preset == "cfg5":
        return W.cfg5(64, M)
    # cfg2: the DGM4 clustering of [X, y] is the step BEFORE the path (host, REF/cGP/cGP_constrained.py:214-232); its
    # labels are the committed fixture so that every run and the scikit-learn golden see the same clusters
    X, y = W.cfg2_samples()
    lab = np.load(os.path.join(ROOT, "tests", "golden", "cfg2_full.npz"))["labels"].astype(np.int64)
    return X, y, lab, W.cfg2_candidates(M)


def committed_evals(preset):
    try:
        return json.load(open(EVALS_FILE))[preset]["evals"]
    except Exception:  # noqa: BLE001
        return None


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


N_CPU_MAX = 8192  # largest cluster the scikit-learn sample is run at (K_gradient alone is n^2 (d+1) doubles)


def cpu_sample(preset, X, y, labels, Q, evals_per_fit, m_sub=8192, n_eval=3):
    """Bounded sample of the reference CPU path on this host with ALL host cores: `n_eval` LML+gradient evaluations of
    cluster 0 at full n_c through scikit-learn (N_eval = 3), plus k-NN routing + predict(return_std) + EI
    on `m_sub` candidates; extrapolated to the whole iteration (evals_per_fit evaluations spread over clusters ~ n_c^3;
    M candidates ~ n_c^2).  Returns (baseline dict, values for the parity block)."""
    import warnings

    from sklearn.neighbors import KNeighborsClassifier
    from threadpoolctl import threadpool_info, threadpool_limits

    import oracle

    p = PRESETS[preset]
    d = p["d"]
    theta = np.array([-0.7] * d + [-4.0])
    sizes = np.bincount(labels)
    c0 = 0
    Xt, yt = X[labels == c0], y[labels == c0].reshape(-1, 1)
    n_full = Xt.shape[0]
    scaled = n_full > N_CPU_MAX
    if scaled:  # configs[3]: scikit-learn cannot allocate its 172 GB K_gradient at n = 65536
        Xt, yt = Xt[:N_CPU_MAX], yt[:N_CPU_MAX]
    ncores = os.cpu_count() or 1
    # torchrun exports OMP_NUM_THREADS=1 to its workers: lift the BLAS pools to every host core explicitly
    with threadpool_limits(limits=ncores), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g = _ref_gpr(d, theta)
        g.fit(Xt, yt)
        t_eval, val = [], None
        for _ in range(n_eval):
            t0 = time.perf_counter()
            val = g.log_marginal_likelihood(theta + 0.01, eval_gradient=True)
            t_eval.append(time.perf_counter() - t0)
        clf = KNeighborsClassifier(n_neighbors=p["k"]).fit(X, labels)
        Qs = Q[:m_sub]
        t0 = time.perf_counter()
        lab = clf.predict(Qs)
        t_knn = time.perf_counter() - t0
        t0 = time.perf_counter()
        mu, sd = g.predict(Qs, return_std=True)
        ei = oracle.ei_scores(np.asarray(mu).reshape(-1), sd, float(yt.max()), Xt.shape[0])
        t_pred = time.perf_counter() - t0
        threads = max([int(i.get("num_threads", 1)) for i in threadpool_info()] + [1])
    n_s = Xt.shape[0]
    s_eval = float(np.mean(t_eval))
    scale3 = float(np.mean((sizes / n_s) ** 3))
    scale2 = float(np.sum((sizes / n_s) ** 2 * (np.bincount(lab, minlength=len(sizes)) / max(len(lab), 1))))
    fit_s = s_eval * evals_per_fit * scale3
    acq_s = (t_knn + t_pred * max(scale2, 1e-9)) * (p["M"] / m_sub)
    base = dict(value=p["M"] / (fit_s + acq_s), unit=UNIT, cores=threads, host_cpus=ncores, extrapolated=True,
                s_per_lml_grad_eval=s_eval, n_lml_grad_evals_timed=n_eval, knn_s=t_knn, predict_ei_s=t_pred,
                fit_s_extrapolated=fit_s, acq_s_extrapolated=acq_s, iteration_s_extrapolated=fit_s + acq_s,
                evals_per_fit=evals_per_fit,
                sample=(f"{n_eval} sklearn LML+grad evals at n_c={n_s}"
                        + (f" (of {n_full}: K_gradient does not fit; x (n/{n_s})^3)" if scaled else "")
                        + f" x {evals_per_fit} evals/fit (profiles/fit_evals.json) + kNN/predict/EI on {m_sub} of {p['M']} "
                        "candidates, extrapolated; CPU cost of the same dense-scan iteration (DESIGN.md 7)"))
    vals = dict(theta=theta + 0.01, theta_fit=theta, lml=float(val[0]), grad=np.asarray(val[1]), labels=lab,
                mean=np.asarray(mu).reshape(-1), std=np.asarray(sd), ei=ei, c0=c0, m_sub=m_sub, truncated=scaled, n_s=n_s)
    return base, vals


def cpu_full_iteration(preset, X, y, labels, Q):
    """The WHOLE iteration through scikit-learn, nothing extrapolated (feasible for configs[1] and the smoke preset):
    per cluster GaussianProcessRegressor.fit with R restarts (REF/cGP/cGP_constrained.py:315-327), k-NN routing of all
    M candidates, predict(return_std) + EI + the reference winner rule (:376-378).  -> (seconds, selected index)."""
    import warnings

    from sklearn.neighbors import KNeighborsClassifier
    from threadpoolctl import threadpool_limits

    import oracle

    p = PRESETS[preset]
    t0 = time.perf_counter()
    with threadpool_limits(limits=os.cpu_count() or 1), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clf = KNeighborsClassifier(n_neighbors=p["k"]).fit(X, labels)
        gprs = []
        for c in range(int(labels.max()) + 1):
            gprs.append(_ref_gpr(p["d"], None, p["R"]).fit(X[labels == c], y[labels == c].reshape(-1, 1)))
        t_fit = time.perf_counter() - t0
        qlab = clf.predict(Q)
        best = (-np.inf, -1)
        for c, g in enumerate(gprs):
            sel = np.nonzero(qlab == c)[0]
            if not len(sel):
                continue
            mu, sd = g.predict(Q[sel], return_std=True)
            sc = oracle.ei_scores(np.asarray(mu).reshape(-1), sd, float(y[labels == c].max()), int(np.sum(labels == c)))
            j = int(np.argmax(sc))
            if sc[j] > best[0]:
                best = (float(sc[j]), int(sel[j]))
    return time.perf_counter() - t0, t_fit, best


def run_reference(args, preset):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    p = PRESETS[preset]
    if args.measured:
        if preset not in ("cfg2", "small"):
            raise SystemExit("--measured runs the whole scikit-learn iteration: only feasible for cfg2 / small")
        X, y, labels, Q = make_workload(preset)
        ts = [cpu_full_iteration(preset, X, y, labels, Q) for _ in range(args.warmup + args.steps)][args.warmup:]
        it = float(np.mean([t[0] for t in ts]))
        v = p["M"] / it
        cpu = dict(value=v, unit=UNIT, cores=os.cpu_count(), kind="reference", extrapolated=False, iteration_s=it,
                   fit_s=float(np.mean([t[1] for t in ts])), selected_idx=ts[-1][2][1], selected_score=ts[-1][2][0],
                   sample="the WHOLE iteration measured: GaussianProcessRegressor.fit per cluster (R restarts) + "
                          "KNeighborsClassifier.predict + predict(return_std) + EI over all M candidates")
        emit({"impl": "reference", "metric": p["metric"], "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
              "warmup": args.warmup, "ms_per_step": it * 1e3, "higher_is_better": True, "scaling": "strong",
              "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": p["name"]},
              "timing": "host wall clock, whole iteration, all host cores", "cpu_baseline": cpu,
              "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return
    X, y, labels, Q = make_workload(preset, M=16384)
    evals = committed_evals(preset) or 17 * (int(labels.max()) + 1) * 30
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        last, _ = cpu_sample(preset, X, y, labels, Q, evals, m_sub=4096 if i < args.warmup else 8192,
                             n_eval=1 if i < args.warmup else 3)
        if i >= args.warmup:
            vals.append(last["iteration_s_extrapolated"])
    it = float(np.mean(vals))
    v = p["M"] / it
    last["kind"] = "reference"
    last["value"] = v
    line = {"impl": "reference", "metric": p["metric"], "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": it * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": p["name"]},
            "timing": "host wall clock of a bounded sample on all host cores, extrapolated (see cpu_baseline.sample)",
            "cpu_baseline": last,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------ parity block
def parity_block(preset, X, y, labels, Q, ref, dev):
    """GPU vs scikit-learn at FULL cluster size on the same theta (cluster 0, the `m_sub`-candidate subsample of the
    CPU leg): the numbers REF/cGP/cGP_constrained.py:315-327 / acquisition.py:12-27 would produce."""
    import torch

    from cgp_b200.engine import ClusterGPEngine

    p = PRESETS[preset]
    c0, m = ref["c0"], ref["m_sub"]
    Xt, yt = X[labels == c0][:ref["n_s"]], y[labels == c0][:ref["n_s"]]
    eng = ClusterGPEngine(Xt, yt, np.zeros(Xt.shape[0], dtype=np.int64), "matern32", 1e-5, device=dev.index,
                          mem_fraction=p.get("mem_fraction", 0.6))
    lml, grad, info = eng.lml_grad([0], ref["theta"][None, :], True)
    out = {"n_c": int(Xt.shape[0]), "candidates": int(m), "truncated_cluster": bool(ref["truncated"]),
           "lml_rel": abs(lml[0] - ref["lml"]) / abs(ref["lml"]),
           "grad_rel_scale": float(np.max(np.abs(grad[0] - ref["grad"])) / max(1.0, np.max(np.abs(ref["grad"]))))}
    eng.set_theta(ref["theta_fit"][None, :])
    Qs = torch.from_numpy(np.ascontiguousarray(Q[:m])).to(dev)
    zl = torch.zeros(m, dtype=torch.int64, device=dev)
    (bs, bi, _), (score, mean, std) = eng.ei_scan(Qs, p["k"], want_arrays=True, qlabel=zl)
    mean, std, score = mean.cpu().numpy(), std.cpu().numpy(), score.cpu().numpy()
    out["mean_rel"] = float(np.max(np.abs(mean - ref["mean"]) / np.maximum(np.abs(ref["mean"]), 1e-3)))  # |mean| floored at 1e-3
    out["var_rel"] = float(np.max(np.abs(std ** 2 - ref["std"] ** 2) / ref["std"] ** 2))
    out["ei_abs"] = float(np.max(np.abs(score - ref["ei"])))
    out["same_argmax_on_subsample"] = bool(int(bi.item()) == int(np.argmax(ref["ei"])))
    # routing: the whole model's k-NN (all clusters) on the same subsample, bit-exact labels
    X_dev = torch.from_numpy(np.ascontiguousarray(X)).to(dev)
    l_dev = torch.from_numpy(np.ascontiguousarray(labels.astype(np.int64))).to(dev)
    from cgp_b200 import _lib

    got = _lib.knn_labels(X_dev, l_dev, Qs, p["k"], int(labels.max()) + 1).cpu().numpy()
    out["knn_mismatches"] = int(np.sum(got != ref["labels"]))
    return {k: (float(f"{v:.3e}") if isinstance(v, float) else v) for k, v in out.items()}


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args, preset):
    import torch
    import torch.distributed as dist

    from cgp_b200 import CGPRegressor, _lib

    p = PRESETS[preset]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    X, y, labels, Q = make_workload(preset)
    M = Q.shape[0]
    lo, hi = (M * rank) // world, (M * (rank + 1)) // world
    Q_dev = torch.from_numpy(Q[lo:hi]).to(dev)
    Q_pin = torch.from_numpy(Q[lo:hi]).pin_memory()
    Q_head = Q[:8192].copy()
    del Q
    _lib.handle(local)
    stats = {}

    def step(q_src, from_host, prune=True):
        stats.pop("_model", None)  # the previous step's model (configs[3]: 69 GB of factors) is released first
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        q = q_src.to(dev, non_blocking=True) if from_host else q_src
        model = CGPRegressor("matern32", alpha=1e-5, n_restarts_optimizer=p["R"], random_state=0, n_neighbors=p["k"],
                             max_lockstep=p.get("max_lockstep"), mem_fraction=p.get("mem_fraction", 0.6))
        model.fit(X, y, labels)
        e[1].record()
        x_next, score, idx = model.select_next(q, sharded=True, prune=prune)
        e[2].record()
        torch.cuda.synchronize()
        eng = model.engine_
        fs = eng.fit_stats
        stats.update(fit_s=e[0].elapsed_time(e[1]) * 1e-3, acq_s=e[1].elapsed_time(e[2]) * 1e-3, evals=fs["evals"],
                     lockstep=fs["lockstep"], eval_s=fs["eval_s"], opt_host_s=fs["optimizer_host_s"],
                     evals_per_cluster=fs["evals_per_cluster"].tolist(), nfev_sorted=sorted(int(v) for v in fs["nfev"]),
                     x_next=x_next.tolist(), score=score, idx=idx, sizes=eng.counts.tolist(),
                     lml=eng.model["lml_host"].tolist(), shard_mode=fs["shard_mode"],
                     fit_async={k: (round(v, 4) if isinstance(v, float) else v) for k, v in fs.items()
                                if k in ("lanes", "stolen", "given", "idle_s", "pairs_local", "evals_per_rank",
                                         "steps_per_rank", "stolen_per_rank", "loop_s_per_rank")})
        stats["_model"] = model
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

    # fp64 tensor peak measured here (cuBLAS DGEMM through torch; MEASURED_PEAKS.json has bf16 only) BEFORE the warm-up
    # steps, so that the timed region follows the warm-up directly (the probe's buffers and empty_cache would otherwise
    # make the first timed step pay cudaMalloc for every workspace again)
    torch.cuda.empty_cache()
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

    for i in range(args.warmup):
        # the last warm-up step runs with per-launch event timing on, so that the event pool of the C library is
        # created here and not inside the timed region
        _lib.profile_enable(2 if i == args.warmup - 1 else 0, local)
        step(Q_dev, False)
    _lib.profile_enable(False, local)
    sampler = ClockSampler(local)
    _lib.profile_read(reset=True, device=local)
    _lib.profile_enable(2, local)  # events around the large kernels only; every launch is still counted
    sampler.start()
    t_dev, t_wall = timed(args.steps, Q_dev, False)
    clocks = sampler.stop()
    prof = _lib.profile_read(reset=True, device=local, detail=True)
    _lib.profile_enable(False, local)
    step_s = t_dev / args.steps
    res = {k: v for k, v in stats.items() if k != "_model"}  # no second reference to the model (configs[3]: 69 GB)
    t_e2e, _ = timed(args.steps, Q_pin, True)
    e2e_s = t_e2e / args.steps
    # outside every timed region: the EXHAUSTIVE scan on the same fitted model (every candidate scored)
    model = stats.pop("_model")
    exhaustive = None
    if not args.no_exhaustive and preset != "cfg4":  # cfg4: 8M x 65536^2 flops = 18 min on one GPU
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _x, score_f, idx_f = model.select_next(Q_dev, sharded=True, prune=False)
        b.record()
        barrier()
        tp = torch.tensor([a.elapsed_time(b) * 1e-3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        exhaustive = {"acq_s": float(tp[0]), "acq_candidates_per_s": M / float(tp[0]),
                      "same_selection": bool(idx_f == res["idx"] and score_f == res["score"])}
    del model
    torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline: per-class exclusive kernel time (events on the launch stream, timed region) and algorithmic flops
    # tallied by the library for exactly the launches that were timed (include/cgp_b200.h: cgp_profile_flops)
    sizes = np.asarray(res["sizes"], dtype=np.float64)
    roof_all = {}
    for k, v in prof.items():
        if v["ms"] > 0 and v["flops"] > 0:
            roof_all[k] = {"ms": round(v["ms"], 1), "timed": v["timed"],
                           "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2),
                           "frac": round(v["flops"] / (v["ms"] * 1e-3) / 1e12 / fp64_peak, 3),
                           "share_of_step": round(v["ms"] * 1e-3 / t_dev, 3)}
    ev_c = np.asarray(res["evals_per_cluster"], dtype=np.float64)  # this rank's evaluations in one step
    step_flops = float(np.sum(ev_c * sizes ** 3)) + float(np.sum(sizes ** 3)) * 2.0 / 3.0  # fit + final factorisation
    step_roof = {"algorithmic_flops": step_flops, "tflops": round(step_flops / step_s / 1e12, 2),
                 "frac": round(step_flops / step_s / 1e12 / fp64_peak, 3),
                 "fit_only_frac": round(step_flops / res["fit_s"] / 1e12 / fp64_peak, 3)}
    # small batches (<= CGP_FORK_MAX matrices per lane: every N > 1 run of this workload) run three streams and two lanes
    # at once, so no launch is timed exclusively: the per-kernel rates are those of the N = 1 line, only `step` is live
    roofline = {"bound": "tensor", "kernel": "k_lauum", "achieved": None, "peak": round(fp64_peak, 2), "unit": "TFLOP/s",
                "frac": None, "traffic": None, "note": "no exclusively timed launches at this batch size (see the N=1 line)",
                "step": step_roof}
    if roof_all:
        top = max(roof_all, key=lambda k: roof_all[k]["ms"])  # the class with the largest share of the step
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(top)
        roofline = {"bound": "tensor", "kernel": "k_" + top, "achieved": roof_all[top]["tflops"], "peak": round(fp64_peak, 2),
                    "unit": "TFLOP/s", "frac": roof_all[top]["frac"], "traffic": traffic,
                    "peak_source": "cuBLAS DGEMM 8192^3 measured in-run",
                    "classes": roof_all,
                    "step": step_roof}
    gpu_launches = int(sum(v["launches"] for v in prof.values()))
    cpu = parity = None
    detail_cpu = {}
    evals_file = committed_evals(preset)
    if world == 1 and not args.no_cpu_baseline:
        Xw, yw, lw, _ = make_workload(preset, M=0)
        cpu, refvals = cpu_sample(preset, Xw, yw, lw, Q_head, res["evals"])
        cpu["kind"] = "port"
        for k in ("host_cpus", "knn_s", "predict_ei_s", "n_lml_grad_evals_timed", "fit_s_extrapolated", "acq_s_extrapolated",
                  "evals_per_fit"):
            detail_cpu[k] = cpu.pop(k, None)  # kept in the detail file, not in the one-line summary
        parity = parity_block(preset, Xw, yw, lw, Q_head, refvals, dev)
    # fully MEASURED (nothing extrapolated) CPU comparison committed for configs[1]: the whole scikit-learn iteration on the
    # host cores against this arm's line for the same preset (profiles/, produced by `--impl reference --preset cfg2 --measured`)
    measured = None
    try:
        ref2 = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_reference_cfg2_measured.json")))
        gpu2 = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_cfg2_1gpu.json")))
        measured = {"workload": "configs[1]", "cpu_s": round(ref2["ms_per_step"] * 1e-3, 2),
                    "cpu_cores": ref2["cpu_baseline"]["cores"], "b200_s": round(gpu2["ms_per_step"] * 1e-3, 4),
                    "b200_e2e_s": round(gpu2["e2e"]["ms_per_step"] * 1e-3, 4),
                    "same_selected": ref2["cpu_baseline"]["selected_idx"] == gpu2["selected"]["idx"],
                    "source": "profiles/r02_bench_{reference_cfg2_measured,cfg2_1gpu}.json"}
    except Exception:  # noqa: BLE001
        pass
    line = {"metric": p["metric"], "value": M / step_s, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": p["name"]},
            "fit_s": round(res["fit_s"], 4), "acq_s": round(res["acq_s"], 4),
            "selected": {"idx": res["idx"], "score": res["score"], "x": [round(v, 6) for v in res["x_next"]]},
            "parity": parity,
            "e2e": {"value": M / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": int(Q_pin.numel() * 8 + X.nbytes + y.nbytes + labels.nbytes),
                    "d2h_bytes_per_step": int((p["d"] + 2) * 8), "ms_per_step": e2e_s * 1e3},
            "gpu_launches": gpu_launches, "roofline": roofline, "cpu_baseline": cpu,
            "cpu_comparison_fully_measured": measured, "clocks": clocks,
            "acq": "exact pruned scan (default)", "acq_exhaustive_scan": exhaustive,
            "acq_candidates_per_s": M / res["acq_s"], "lml_evals_per_fit_rank0": res["evals"],
            "lml_evals_committed": evals_file, "lockstep_per_fit": res["lockstep"], "shard_mode": res["shard_mode"], "fit_async_rank0": res["fit_async"],
            "fit_optimizer_host_s": round(res["opt_host_s"], 4),
            "l2": "working set >> 126 MB L2, no flush"}
    emit(line)
    detail = {"preset": preset, "n_gpus": world, "nfev_per_pair_rank0_sorted": res["nfev_sorted"],
              "evals_per_cluster_rank0": res["evals_per_cluster"], "sizes": res["sizes"], "lml": res["lml"],
              "kernel_classes": prof, "cpu_baseline_detail": detail_cpu, "fit_eval_wall_s": res["eval_s"],
              "wall_s_per_step": t_wall / args.steps}
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(detail, open(os.path.join(ROOT, "gpurun_out", f"bench_detail_{preset}_n{world}.json"), "w"))
    except OSError:
        pass
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
    ap.add_argument("--no-exhaustive", action="store_true")
    ap.add_argument("--measured", action="store_true",
                    help="reference arm: run the WHOLE scikit-learn iteration instead of an extrapolated sample (cfg2 / small)")
    args = ap.parse_args()
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, args.preset)
    else:
        run_b200(args, args.preset)


if __name__ == "__main__":
    main()
