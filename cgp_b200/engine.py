"""Host-side engine of the per-cluster GP hot path (everything after clustering in one cGP iteration,
REF/cGP/cGP_constrained.py:287-379): batched fit of all clusters x restarts, k-NN routing, routed
prediction and the EI candidate scan.  All arithmetic runs in the CUDA extension (cgp_b200/_lib.py);
this module only moves tensors, drives SciPy's L-BFGS-B state machines in lock-step and shards work
over ranks.  No CPU fallback.
"""
from __future__ import annotations

import gc
import math
import os
import time

import numpy as np
import torch

from . import _lib, dist as cdist
from .lbfgs import StealBoard, minimize_async, minimize_lockstep

_FIT_EPOCH = [0]  # per-process count of sharded fits: key prefix of the work-stealing board (same on every rank)

LOG_LO = math.log(1e-5)  # REF/cGP/cGP_constrained.py:169
LOG_HI = math.log(1e5)


def restart_thetas(P, R, seed=0, lo=None, hi=None):
    """Start points of runs 0..R (SK/gaussian_process/_gpr.py:312-333): run 0 = kernel theta (0 for the
    reference's template), runs 1..R = RandomState(seed).uniform(lo, hi) drawn sequentially.  The draws do
    not depend on results, so they are pre-drawn and shared by every cluster (random_state=0 per fit,
    REF/cGP/cGP_constrained.py:315)."""
    rng = np.random.RandomState(seed)
    lo = np.full(P, LOG_LO) if lo is None else np.asarray(lo, float)
    hi = np.full(P, LOG_HI) if hi is None else np.asarray(hi, float)
    out = np.zeros((R + 1, P))
    for r in range(R):
        out[r + 1] = rng.uniform(lo, hi)
    return out


class ClusterGPEngine:
    """Device-resident state of one clustered GP surrogate."""

    def __init__(self, X, y, labels, kind="matern32", alpha=1e-5, normalize_y=True, device=None,
                 mem_fraction=0.6):
        if not torch.cuda.is_available():
            raise _lib.CgpError("cgp_b200 needs a CUDA device (B200); there is no CPU fallback")
        _lib.lib()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
        y = np.asarray(y, dtype=np.float64).reshape(-1)
        labels = np.asarray(labels).astype(np.int64).reshape(-1)
        if X.ndim != 2 or X.shape[0] != y.shape[0] or labels.shape[0] != y.shape[0]:
            raise ValueError("X [n,d], y [n], labels [n] expected")
        self.kind = kind
        self.alpha = float(alpha)
        self._normalize_y = bool(normalize_y)
        self.n, self.d = X.shape
        self.P = self.d + 1
        self.C = int(labels.max()) + 1
        counts = np.bincount(labels, minlength=self.C)
        if np.any(counts == 0) or labels.min() < 0:
            raise ValueError("labels must be recoded to 0..C-1 with no empty cluster "
                             "(REF/cGP/cGP_constrained.py:249-255)")
        self.order = np.argsort(labels, kind="stable")  # cluster-sorted row order
        self.offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        self.counts = counts
        Xs = X[self.order]
        ys = y[self.order]
        # per-cluster target normalisation on the host, exactly as SK/_gpr.py:275-280
        self.y_mean = np.zeros(self.C)
        self.y_std = np.ones(self.C)
        self.y_max = np.zeros(self.C)
        yn = np.empty_like(ys)
        for c in range(self.C):
            seg = ys[self.offs[c]:self.offs[c + 1]]
            self.y_max[c] = np.max(seg)  # incumbent of REF/cGP/acquisition.py:23
            if normalize_y:
                m = np.mean(seg)
                s = np.std(seg)
                if s < 10 * np.finfo(np.float64).eps:  # _handle_zeros_in_scale
                    s = 1.0
                self.y_mean[c], self.y_std[c] = m, s
            yn[self.offs[c]:self.offs[c + 1]] = (seg - self.y_mean[c]) / self.y_std[c]
        dev = self.device
        self.X_host, self.y_host, self.labels_host = X, y, labels
        self.Xs = torch.from_numpy(Xs).to(dev)
        self.yn = torch.from_numpy(yn).to(dev)
        self.yn_host = yn
        self.X_orig = torch.from_numpy(X).to(dev)
        self.labels_dev = torch.from_numpy(labels).to(dev)
        self.y_mean_dev = torch.from_numpy(self.y_mean).to(dev)
        self.y_std_dev = torch.from_numpy(self.y_std).to(dev)
        self.y_max_dev = torch.from_numpy(self.y_max).to(dev)
        self.mem_fraction = mem_fraction
        self.theta = None
        self.model = None
        self._ws = None
        self._ws_lane = {}      # lane id (>= 1) -> workspace of that evaluation lane
        self._lane_streams = {}
        self._ws_capped = False
        self._slots = []
        self.fit_stats = {}

    # ------------------------------------------------------------------ workspace
    def _workspace(self, nbytes, partial_ok=False, lane=0):
        """Grow-only scratch buffer (one per evaluation lane; lane 0 also serves every non-fit call).
        partial_ok: the caller (batched LML) can work in chunks with less than nbytes."""
        if lane:
            ws = self._ws_lane.get(lane)
            have = 0 if ws is None else ws.numel()
            if have >= nbytes or (self._ws_capped and have > 0):
                return ws
            free, _total = torch.cuda.mem_get_info(self.device)
            want = min(nbytes, int((free + have) * self.mem_fraction * 0.5))
            if want > have:
                self._ws_lane[lane] = None
                self._ws_lane[lane] = torch.empty(want, dtype=torch.uint8, device=self.device)
            if want < nbytes:
                self._ws_capped = True
            return self._ws_lane[lane]
        have = 0 if self._ws is None else self._ws.numel()
        if have >= nbytes or (partial_ok and self._ws_capped and have > 0):
            return self._ws
        # cudaMemGetInfo is slow (measured: up to 59 ms per call on a busy B200) -> only when the workspace must grow.
        # Blocks cached by torch's allocator (e.g. the factors of a previous model) are as good as free: the allocator
        # reuses or releases them on demand (an explicit empty_cache here cost 9 ms + every later malloc, per engine).
        free, _total = torch.cuda.mem_get_info(self.device)
        free += max(0, torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device))
        budget = int((free + have) * self.mem_fraction)
        if partial_ok:
            # the fit's evaluation workspace must leave room for what follows it: the factors L and W of the final model
            # (configs[3]: 2 x 34 GB) next to the factorisation workspace, which reuses this buffer
            reserve = 16 * _lib.model_elems(self.offs)
            one_pair = max(_lib.lml_workspace_bytes(self.offs, np.asarray([c], dtype=np.int32), self.d) for c in range(self.C))
            budget = max(budget - reserve, min(one_pair, budget))
            if budget < min(one_pair, nbytes):
                raise _lib.CgpError(f"out of device memory: one LML evaluation needs a {one_pair / 2**30:.2f} GiB workspace, "
                                    f"{budget / 2**30:.2f} GiB are available (mem_fraction={self.mem_fraction})")
        want = min(nbytes, budget)
        if want < nbytes and not partial_ok:
            raise _lib.CgpError(f"out of device memory: this call needs a {nbytes / 2**30:.2f} GiB workspace, the budget is "
                                f"{budget / 2**30:.2f} GiB (mem_fraction={self.mem_fraction} of the free memory)")
        if want > have:
            self._ws = None
            self._ws = torch.empty(want, dtype=torch.uint8, device=self.device)
        self._ws_capped = partial_ok and want < nbytes  # at the memory budget: the C-ABI chunks the batch; do not ask again this fit
        return self._ws

    # ------------------------------------------------------------------ LML
    def _eval_slot(self, B):
        """A free set of persistent evaluation buffers (device theta/lml/grad/info + pinned host mirrors) holding at
        least B pairs.  Lock-steps reuse them, so the steady state of a fit allocates nothing (pinned allocations and
        device mallocs cost milliseconds and showed up as 5-50 ms spikes in the host part of small-matrix fits)."""
        for slot in self._slots:
            if not slot["busy"] and slot["cap"] >= B:
                return slot
        cap = 1
        while cap < B:
            cap *= 2
        dev, P = self.device, self.P
        slot = dict(cap=cap, busy=False, ev=torch.cuda.Event(),
                    th=torch.empty((cap, P), dtype=torch.float64, device=dev),
                    lml=torch.empty(cap, dtype=torch.float64, device=dev),
                    grad=torch.empty((cap, P), dtype=torch.float64, device=dev),
                    info=torch.empty(cap, dtype=torch.int32, device=dev),
                    th_h=torch.empty((cap, P), dtype=torch.float64, pin_memory=True),
                    lml_h=torch.empty(cap, dtype=torch.float64, pin_memory=True),
                    grad_h=torch.empty((cap, P), dtype=torch.float64, pin_memory=True),
                    info_h=torch.empty(cap, dtype=torch.int32, pin_memory=True))
        self._slots = [s for s in self._slots if s["busy"] or s["cap"] >= cap] + [slot]  # drop outgrown free slots
        return slot

    def _lane_stream(self, lane):
        """Lane 0 = the caller's current stream; every further lane owns a torch stream (and a library handle with its
        own internal streams), so that batches of different lanes run concurrently on the device."""
        if lane == 0:
            return torch.cuda.current_stream(self.device)
        st = self._lane_streams.get(lane)
        if st is None:
            st = self._lane_streams[lane] = torch.cuda.Stream(self.device)
        return st

    def lml_grad_submit(self, pair_cluster, thetas, want_grad=True, lane=0):
        """Enqueue a batched LML (+ gradient) evaluation and the device->pinned-host copies of its results.
        Returns a ticket for lml_grad_collect; nothing here waits for the device."""
        if lane:
            with torch.cuda.stream(self._lane_stream(lane)):
                return self._lml_grad_submit(pair_cluster, thetas, want_grad, lane)
        return self._lml_grad_submit(pair_cluster, thetas, want_grad, 0)

    def lml_grad_done(self, ticket):
        """Non-blocking: has the evaluation of `ticket` finished?"""
        return ticket[0]["ev"].query()

    def _lml_grad_submit(self, pair_cluster, thetas, want_grad, lane):
        pair_cluster = np.asarray(pair_cluster, dtype=np.int32)
        B = len(pair_cluster)
        thetas = np.asarray(thetas, dtype=np.float64).reshape(B, self.P)
        need = _lib.lml_workspace_bytes(self.offs, pair_cluster, self.d)
        ws = self._workspace(need, partial_ok=True, lane=lane)
        slot = self._eval_slot(B)
        slot["busy"] = True
        slot["th_h"].numpy()[:B] = thetas
        th = slot["th"][:B]
        th.copy_(slot["th_h"][:B], non_blocking=True)
        out = (slot["lml"][:B], slot["grad"][:B], slot["info"][:B])
        _lib.lml_grad_batched(self.Xs, self.yn, self.offs, pair_cluster, th, self.alpha, self.kind, ws, want_grad, out=out,
                              lane=lane)
        slot["lml_h"][:B].copy_(out[0], non_blocking=True)
        if want_grad:
            slot["grad_h"][:B].copy_(out[1], non_blocking=True)
        slot["info_h"][:B].copy_(out[2], non_blocking=True)
        slot["ev"].record(torch.cuda.current_stream(self.device))
        return slot, B, want_grad

    def lml_grad_collect(self, ticket):
        """-> (lml, grad | None, info) host arrays; non-PD -> (-inf, 0) as SK/_gpr.py:592-593."""
        slot, B, want_grad = ticket
        slot["ev"].synchronize()
        lml, info = slot["lml_h"].numpy()[:B].copy(), slot["info_h"].numpy()[:B].copy()
        grad = slot["grad_h"].numpy()[:B].copy() if want_grad else None
        slot["busy"] = False
        bad = info != 0
        lml[bad] = -np.inf
        if grad is not None:
            grad[bad] = 0.0
        return lml, grad, info

    def lml_grad(self, pair_cluster, thetas, want_grad=True):
        """Batched log-marginal likelihood (+ gradient) for pairs (cluster, theta).  Host arrays in/out."""
        return self.lml_grad_collect(self.lml_grad_submit(pair_cluster, thetas, want_grad))

    # ------------------------------------------------------------------ fit
    def fit(self, n_restarts, seed=0, theta0=None, bounds=None, optimize=True, shard_mode=None, max_lockstep=None):
        """All clusters x (R+1) L-BFGS-B runs (SK/_gpr.py:296-340) with batched GPU evaluations, sharded over ranks, then
        the final factorisation at each cluster's best theta (SK/_gpr.py:349-367).

        shard_mode (env CGP_SHARD_MODE):
        * "async" (default): every rank OWNS whole runs and advances them without waiting for any other rank; inside a
          rank up to CGP_LANES (2) evaluation lanes run concurrently (the latency-bound chain of one lane's
          factorisations overlaps the bulk of the other's, and the host advances one lane's state machines while the
          device works on the other); a rank that runs short of runs steals ready ones from the busiest rank over the
          c10d store (lbfgs.StealBoard).  One all-reduce of the (theta*, -lml) table at the end.  No per-step exchange,
          no ceil(active / world) quantisation, no rank waits for the slowest evaluation of another.
        * "balanced": every rank keeps all state machines; each global lock-step deals the waiting runs to the ranks and
          all-reduces (f, g) (round-1 default; 8 B200, config 3: 2.62 s / iteration).
        * "static": pairs dealt round-robin once, no exchange until the final all-gather (2.70 s; its slowest rank draws
          754 of the 4938 evaluations, mean 617).
        Every run sees the same sequence of (theta -> f, g) in every mode and on any number of ranks (the kernels are
        batch-invariant bit for bit), so the selected point and its score bits do not depend on the mode.
        max_lockstep: cut every run after that many evaluations (it keeps its best evaluated theta): the
        single-65536-sample-cluster configuration measures ONE evaluation of each of its R+1 runs (9 s each)."""
        if shard_mode is None:
            shard_mode = os.environ.get("CGP_SHARD_MODE", "async")
        rank, ws = cdist.world()
        P, C = self.P, self.C
        self._ws_capped = False
        lo = np.full(P, LOG_LO) if bounds is None else np.asarray(bounds, float)[:, 0]
        hi = np.full(P, LOG_HI) if bounds is None else np.asarray(bounds, float)[:, 1]
        starts = restart_thetas(P, n_restarts, seed, lo, hi)
        start0 = None  # per-cluster start of run 0 (warm start from a previous fit), else the same theta0 everywhere
        if theta0 is not None:
            theta0 = np.asarray(theta0, float)
            if theta0.ndim == 2:
                start0 = np.clip(theta0.reshape(C, P), lo, hi)
            else:
                starts[0] = theta0
        R1 = starts.shape[0]
        if not optimize:
            self.theta = np.tile(starts[0], (C, 1)) if start0 is None else start0.copy()
            self._factorize()
            return self
        n_pairs = C * R1  # pair p = (cluster p // R1, run p % R1)
        pair_c = np.arange(n_pairs, dtype=np.int32) // R1
        n_eval = [0]
        t_eval = [0.0]
        evals_c = np.zeros(C, dtype=np.int64)  # evaluations THIS rank performed, per cluster
        t_fit0 = time.perf_counter()
        x0_of = lambda p: starts[p % R1] if (start0 is None or p % R1) else start0[p // R1]  # noqa: E731
        gc_was_on = gc.isenabled()
        gc.disable()  # thousands of live L-BFGS-B machines: a generation-2 sweep inside a lock-step costs tens of ms
        extra = {}
        try:
            if shard_mode == "async":
                table, steps, extra = self._fit_async(n_pairs, pair_c, x0_of, lo, hi, rank, ws, max_lockstep, n_eval, t_eval,
                                                      evals_c)
                mode = "async"
            else:
                balanced = ws > 1 and shard_mode == "balanced"
                mine = list(range(n_pairs)) if balanced else cdist.my_pairs(n_pairs, rank, ws)
                pc_all = pair_c[mine]

                def eval_batch(idx, Xs):
                    t0 = time.perf_counter()
                    lml, grad, _ = self.lml_grad(pc_all[idx], Xs, True)
                    n_eval[0] += len(idx)
                    np.add.at(evals_c, pc_all[idx], 1)
                    t_eval[0] += time.perf_counter() - t0
                    return -lml, -grad

                def submit(idx, Xs):
                    n_eval[0] += len(idx)
                    np.add.at(evals_c, pc_all[idx], 1)
                    return self.lml_grad_submit(pc_all[idx], Xs, True)

                def collect(ticket):
                    t0 = time.perf_counter()
                    lml, grad, _ = self.lml_grad_collect(ticket)
                    t_eval[0] += time.perf_counter() - t0
                    return -lml, -grad

                shard = None
                if balanced:
                    def allreduce(buf):
                        t = torch.from_numpy(buf).to(self.device)
                        torch.distributed.all_reduce(t)
                        return t.cpu().numpy()

                    shard = dict(rank=rank, world=ws, cost=self.counts[pc_all].astype(np.float64) ** 3, allreduce=allreduce)
                xs, fs, nfev, nit, steps = minimize_lockstep(
                    eval_batch, [x0_of(p) for p in mine], lo, hi, submit=submit, collect=collect,
                    pipeline_min=int(os.environ.get("CGP_PIPELINE_MIN", "256")), shard=shard,
                    **({} if max_lockstep is None else {"max_steps": int(max_lockstep)}))
                rows = np.concatenate([xs, fs[:, None], nfev[:, None], nit[:, None]], axis=1).reshape(len(mine), P + 3)
                if balanced:  # every rank already holds every run's result
                    table = rows
                else:
                    table = cdist.gather_pair_rows(torch.from_numpy(rows).to(self.device), n_pairs, rank, ws).cpu().numpy()
                mode = "balanced" if balanced else "static"
        finally:
            if gc_was_on:
                gc.enable()
        t_opt = time.perf_counter() - t_fit0
        table = table.reshape(C, R1, P + 3)
        self.run_thetas = table[:, :, :P]
        self.run_values = table[:, :, P]
        best = np.argmin(self.run_values, axis=1)  # first minimum in run order, SK/_gpr.py:336-337
        self.theta = np.stack([self.run_thetas[c, best[c]] for c in range(C)])
        self.fit_stats = dict(lockstep=steps, evals=int(n_eval[0]), nfev=table[:, :, P + 1].reshape(-1).astype(np.int64),
                              nit=table[:, :, P + 2].reshape(-1).astype(np.int64), pairs_total=n_pairs, eval_s=t_eval[0],
                              optimizer_host_s=t_opt - t_eval[0], evals_per_cluster=evals_c, shard_mode=mode, **extra)
        self._factorize()
        return self

    def _fit_async(self, n_pairs, pair_c, x0_of, lo, hi, rank, ws, max_evals, n_eval, t_eval, evals_c):
        """Body of shard_mode "async" (see fit): -> (result table [n_pairs, P + 3], lane steps, extra stats)."""
        P = self.P
        n_lanes = max(1, int(os.environ.get("CGP_LANES", "2")))
        lane_max = int(os.environ.get("CGP_LANE_MAX", "48"))
        pipe_min = int(os.environ.get("CGP_PIPELINE_MIN", "256"))
        # a single pair that needs a large share of the device memory (configs[3]: 69 GB) leaves no room for a 2nd lane
        big = max(_lib.lml_workspace_bytes(self.offs, np.asarray([c], dtype=np.int32), self.d) for c in range(self.C))
        free, _ = torch.cuda.mem_get_info(self.device)
        if big * 4 > free * self.mem_fraction:
            n_lanes = 1
        # initial deal: heaviest runs first, round-robin (costs ~ n_c^3)
        order = np.argsort(-(self.counts[pair_c].astype(np.float64) ** 3), kind="stable")
        mine = np.sort(order[rank::ws])
        cur = torch.cuda.current_stream(self.device)
        for lane in range(1, n_lanes):
            self._lane_stream(lane).wait_stream(cur)  # the training arrays were uploaded on the caller's stream
        multi = [False]
        small = int(self.counts.max()) < 2048  # small matrices: the host's share of a step is never negligible

        def lane_policy(active):
            # many runs: two lanes hide the host's L-BFGS-B work behind the device; few runs: two lanes overlap the
            # latency-bound chain of one batch with the bulk of the other; in between one batch fills the device
            use = n_lanes if (active > pipe_min or active <= lane_max or small) else 1
            multi[0] = use > 1
            return use

        def submit(lane, gid, X):
            pc = pair_c[gid]
            n_eval[0] += len(gid)
            np.add.at(evals_c, pc, 1)
            if n_lanes > 1:
                _lib.profile_mute(multi[0], self.device.index, lane)  # concurrent lanes: event brackets not exclusive
            return self.lml_grad_submit(pc, X, True, lane=lane)

        def collect(ticket):
            t0 = time.perf_counter()
            lml, grad, _ = self.lml_grad_collect(ticket)
            t_eval[0] += time.perf_counter() - t0
            return -lml, -grad

        board = None
        if ws > 1:
            from torch.distributed.distributed_c10d import _get_default_store

            _FIT_EPOCH[0] += 1
            board = StealBoard(_get_default_store(), f"cgp_b200/fit{_FIT_EPOCH[0]}", rank, ws, n_pairs)
        finished, st = minimize_async([x0_of(int(p)) for p in mine], mine, lo, hi, submit, self.lml_grad_done, collect,
                                      n_lanes=n_lanes, lane_policy=lane_policy, board=board, max_evals=max_evals)
        for lane in range(1, n_lanes):
            if n_lanes > 1:
                _lib.profile_mute(False, self.device.index, lane)
            cur.wait_stream(self._lane_stream(lane))
        if n_lanes > 1:
            _lib.profile_mute(False, self.device.index, 0)
        table = np.zeros((n_pairs, P + 3))
        for gid, (x, f, nfev, nit) in finished.items():
            table[gid, :P], table[gid, P], table[gid, P + 1], table[gid, P + 2] = x, f, nfev, nit
        extra = dict(lanes=n_lanes, stolen=st["stolen"], given=st["given"], idle_s=st["idle_s"], pairs_local=len(mine))
        if ws > 1:  # every run finished on exactly one rank: the sum over ranks is the full table (exact: + 0.0)
            t = torch.from_numpy(table).to(self.device)
            torch.distributed.all_reduce(t)
            table = t.cpu().numpy()
            mine_stats = torch.tensor([float(n_eval[0]), float(st["steps"]), float(st["stolen"]), st["loop_s"]],
                                      dtype=torch.float64, device=self.device)
            allst = torch.empty((ws, 4), dtype=torch.float64, device=self.device)
            torch.distributed.all_gather_into_tensor(allst, mine_stats.reshape(1, 4))
            a = allst.cpu().numpy()
            extra.update(evals_per_rank=a[:, 0].astype(int).tolist(), steps_per_rank=a[:, 1].astype(int).tolist(),
                         stolen_per_rank=a[:, 2].astype(int).tolist(), loop_s_per_rank=[round(v, 3) for v in a[:, 3]])
        return table, st["steps"], extra

    def set_theta(self, theta):
        self.theta = np.asarray(theta, dtype=np.float64).reshape(self.C, self.P).copy()
        self._factorize()
        return self

    def _factorize(self):
        th = torch.from_numpy(np.ascontiguousarray(self.theta)).to(self.device)
        need = _lib.factorize_workspace_bytes(self.offs, self.d)
        ws = self._workspace(need)
        if ws.numel() < need:
            raise _lib.CgpError("not enough device memory for the final factorisation")
        m = _lib.factorize(self.Xs, self.yn, self.offs, th, self.alpha, self.kind, ws)
        info = m["info"].cpu().numpy()
        if np.any(info != 0):
            c = int(np.nonzero(info)[0][0])
            raise np.linalg.LinAlgError(
                f"The kernel of cluster {c} is not returning a positive definite matrix "
                f"({info[c]}-th leading minor). Try gradually increasing the 'alpha' parameter "
                "of your GaussianProcessRegressor estimator.")  # SK/_gpr.py:353-361
        m["theta_dev"] = th
        m["lml_host"] = m["lml"].cpu().numpy()
        self.model = m
        npad = [_lib.padded_n(int(c)) for c in self.counts]
        self.npad = np.asarray(npad, dtype=np.int64)
        self.moff = np.concatenate([[0], np.cumsum(self.npad * self.npad)]).astype(np.int64)

    # ------------------------------------------------------------------ incremental refit
    def append(self, x, y, label):
        """Add ONE sample to cluster `label` of the factorised model, theta unchanged.

        The reference refits every cluster from scratch in every sequential iteration
        (REF/cGP/cGP_constrained.py:276,315-327).  When the old samples keep their labels only one cluster changes,
        and at fixed theta its factor grows by one row: two triangular mat-vecs against the stored W = L^-1 (2 n^2
        flops) instead of a Cholesky (n^3/3) — ``cgp_append_sample``.  The cluster's targets are re-normalised
        (SK/_gpr.py:275-280) and alpha = K^-1 y, the LML value and the incumbent follow (``cgp_refresh_alpha``).
        Result == a from-scratch ``set_theta(self.theta)`` on the extended data (tests/test_gpu_append.py, 1e-10).
        Re-optimising theta afterwards is ``fit(R, theta0=self.theta)`` (run 0 of each cluster starts at its optimum)."""
        if self.model is None:
            raise _lib.CgpError("append() needs a factorised model (fit / set_theta first)")
        c = int(label)
        if not 0 <= c < self.C:
            raise ValueError("label must be an existing cluster id (a new cluster needs a full refit)")
        x = np.asarray(x, dtype=np.float64).reshape(1, self.d)
        # everything below replaces attributes instead of mutating them, except the new factor row (written in place
        # into the identity pad) -> a failure (non-PD, out of memory, ...) restores this snapshot: no half-updated engine
        keep = ("X_host", "y_host", "labels_host", "order", "n", "counts", "offs", "yn_host", "Xs", "yn", "X_orig",
                "labels_dev", "y_mean_dev", "y_std_dev", "y_max_dev", "npad", "moff")
        snap = {k: getattr(self, k) for k in keep}
        snap_stats = (self.y_mean.copy(), self.y_std.copy(), self.y_max.copy())
        snap_model = dict(self.model)
        try:
            return self._append(x, y, c)
        except BaseException:
            for k, v in snap.items():
                setattr(self, k, v)
            self.y_mean, self.y_std, self.y_max = snap_stats
            same_buffers = self.model["L"] is snap_model["L"]
            self.model.clear()
            self.model.update(snap_model)
            if same_buffers:  # the kernels may have written row n_c of the cluster's L / W block: back to identity pad
                npd, row = int(self.npad[c]), int(self.counts[c])
                if row < npd:
                    for key in ("L", "W"):
                        blk = self.model[key][self.moff[c]:self.moff[c + 1]].view(npd, npd)
                        blk[row].zero_()
                        blk[row, row] = 1.0
            raise

    def _append(self, x, y, c):
        dev = self.device
        m = self.model
        old_offs, old_npad, old_moff = self.offs.copy(), self.npad.copy(), self.moff.copy()
        pos = int(old_offs[c + 1])  # row of the new sample in cluster-sorted order (stable sort: last of its cluster)
        self.X_host = np.concatenate([self.X_host, x])
        self.y_host = np.append(self.y_host, float(y))
        self.labels_host = np.append(self.labels_host, np.int64(c))
        self.order = np.concatenate([self.order[:pos], [self.n], self.order[pos:]]).astype(self.order.dtype)
        self.n += 1
        self.counts = self.counts.copy()
        self.counts[c] += 1
        self.offs = np.concatenate([[0], np.cumsum(self.counts)]).astype(np.int64)
        # targets of cluster c: new statistics, re-normalised (same code path as __init__)
        seg = self.y_host[self.order[self.offs[c]:self.offs[c + 1]]]
        self.y_mean, self.y_std, self.y_max = self.y_mean.copy(), self.y_std.copy(), self.y_max.copy()
        self.y_max[c] = np.max(seg)
        if self._normalize_y:
            mu, sd = np.mean(seg), np.std(seg)
            if sd < 10 * np.finfo(np.float64).eps:
                sd = 1.0
            self.y_mean[c], self.y_std[c] = mu, sd
        yn = np.concatenate([self.yn_host[:pos], [0.0], self.yn_host[pos:]])
        yn[self.offs[c]:self.offs[c + 1]] = (seg - self.y_mean[c]) / self.y_std[c]
        self.yn_host = yn
        self.Xs = torch.cat([self.Xs[:pos], torch.from_numpy(x).to(dev), self.Xs[pos:]]).contiguous()
        self.yn = torch.from_numpy(yn).to(dev)
        self.X_orig = torch.cat([self.X_orig, torch.from_numpy(x).to(dev)]).contiguous()
        self.labels_dev = torch.from_numpy(self.labels_host).to(dev)
        self.y_mean_dev = torch.from_numpy(self.y_mean).to(dev)
        self.y_std_dev = torch.from_numpy(self.y_std).to(dev)
        self.y_max_dev = torch.from_numpy(self.y_max).to(dev)
        # factor buffers: in place unless the cluster crosses a 128-row padding boundary
        self.npad = np.asarray([_lib.padded_n(int(v)) for v in self.counts], dtype=np.int64)
        self.moff = np.concatenate([[0], np.cumsum(self.npad * self.npad)]).astype(np.int64)
        if self.npad[c] != old_npad[c]:
            for key in ("L", "W"):
                old = m[key]
                new = torch.zeros(int(self.moff[-1]), dtype=torch.float64, device=dev)
                for cc in range(self.C):
                    po, pn = int(old_npad[cc]), int(self.npad[cc])
                    src = old[old_moff[cc]:old_moff[cc + 1]].view(po, po)
                    dst = new[self.moff[cc]:self.moff[cc + 1]].view(pn, pn)
                    dst[:po, :po].copy_(src)
                    if pn > po:
                        dst.diagonal()[po:].fill_(1.0)  # identity pad
                m[key] = new
        av = torch.empty(self.n, dtype=torch.float64, device=dev)
        av[:pos] = m["alpha_vec"][:pos]
        av[pos + 1:] = m["alpha_vec"][pos:]
        m["alpha_vec"] = av
        pn, nc = int(self.npad[c]), int(self.counts[c])
        Lc = m["L"][self.moff[c]:self.moff[c + 1]].view(pn, pn)
        Wc = m["W"][self.moff[c]:self.moff[c + 1]].view(pn, pn)
        ws = self._workspace(_lib.append_workspace_bytes(nc))
        Xc = self.Xs[self.offs[c]:self.offs[c + 1]]
        info = _lib.append_sample(Xc, m["theta_dev"][c], self.alpha, self.kind, Lc, Wc, ws)
        lml = _lib.refresh_alpha(Lc, Wc, nc, self.yn[self.offs[c]:self.offs[c + 1]], av[self.offs[c]:self.offs[c + 1]], ws)
        if int(info.item()) != 0:
            raise np.linalg.LinAlgError(
                f"The kernel of cluster {c} is not returning a positive definite matrix after appending a sample "
                f"({int(info.item())}-th leading minor). Try gradually increasing the 'alpha' parameter "
                "of your GaussianProcessRegressor estimator.")
        lml_all = m["lml"].clone()
        lml_all[c] = lml[0]
        m["lml"] = lml_all
        m["lml_host"] = lml_all.cpu().numpy()
        m.pop("_mspe", None)
        return self

    # ------------------------------------------------------------------ accessors (host copies)
    def L_of(self, c):
        npd, n = int(self.npad[c]), int(self.counts[c])
        return self.model["L"][self.moff[c]:self.moff[c + 1]].view(npd, npd)[:n, :n].cpu().numpy()

    def W_of(self, c):
        npd, n = int(self.npad[c]), int(self.counts[c])
        return self.model["W"][self.moff[c]:self.moff[c + 1]].view(npd, npd)[:n, :n].cpu().numpy()

    def alpha_of(self, c):
        return self.model["alpha_vec"][self.offs[c]:self.offs[c + 1]].cpu().numpy()

    # ------------------------------------------------------------------ routing / prediction / acquisition
    def _to_dev(self, Q):
        if isinstance(Q, torch.Tensor):
            return Q.to(self.device, torch.float64).contiguous()
        return torch.from_numpy(np.ascontiguousarray(np.asarray(Q, dtype=np.float64))).to(self.device)

    def route(self, Q, k=3):
        """k-NN cluster labels of Q (REF/cGP/acquisition.py:7)."""
        Qd = self._to_dev(Q)
        return _lib.knn_labels(self.X_orig, self.labels_dev, Qd, min(k, self.n), self.C)

    def predict(self, Q, qlabel=None, k=3, return_std=True):
        """Routed posterior mean/std (REF/cGP/f_truth_dill.py:20-30)."""
        Qd = self._to_dev(Q)
        if qlabel is None:
            qlabel = self.route(Qd, k)
        m = self.model
        ws = self._workspace(_lib.score_workspace_bytes(self.offs, self.d, Qd.shape[0]))
        return _lib.predict(self.Xs, self.offs, m["theta_dev"], m["W"], m["alpha_vec"], self.y_mean_dev, self.y_std_dev,
                            Qd, qlabel, self.kind, return_std, ws)

    def ei_scan(self, Q, k=3, idx_base=0, want_arrays=False, qlabel=None, prune=False, penalty=None):
        """Dense EI scan of candidates Q -> ((score, idx, cluster) CUDA 1-elem tensors, arrays).
        prune=True: exact bound-pruned scan, identical selection (see cgp_ei_argmax_pruned).
        penalty: per-candidate additive soft-constraint term [M] (REF/cGP/acquisition.py:29) or None."""
        Qd = self._to_dev(Q)
        if qlabel is None:
            qlabel = self.route(Qd, k)
        if penalty is not None:
            penalty = self._to_dev(np.asarray(penalty, dtype=np.float64).reshape(-1) if not isinstance(penalty, torch.Tensor)
                                   else penalty)
        m = self.model
        ws = self._workspace(_lib.score_workspace_bytes(self.offs, self.d, Qd.shape[0]))
        return _lib.ei_argmax(self.Xs, self.offs, m["theta_dev"], m["W"], m["alpha_vec"], self.y_mean_dev,
                              self.y_std_dev, self.y_max_dev, Qd, qlabel, self.kind, idx_base, want_arrays, ws, prune,
                              penalty=penalty, alpha=self.alpha)

    def penalty_of(self, Q, qlabel, boundary_penalty):
        """Evaluate the reference's soft-constraint hook ``boundary_penalty(X, data_X)`` (REF/cGP/acquisition.py:29:
        called per point with the samples of the point's cluster) for every candidate at once: per routed cluster one
        vectorised call ``boundary_penalty(Q_c, X_c)``; hooks that only work point-wise are called per candidate.
        -> [M] float64 CUDA tensor, or None when the hook returns 0 everywhere."""
        Qh = Q.cpu().numpy() if isinstance(Q, torch.Tensor) else np.asarray(Q, dtype=np.float64)
        ql = qlabel.cpu().numpy()
        pen = np.zeros(Qh.shape[0])
        for c in range(self.C):
            sel = np.nonzero(ql == c)[0]
            if not len(sel):
                continue
            Xc = self.X_host[self.labels_host == c]
            v = None
            try:
                v = np.asarray(boundary_penalty(Qh[sel], Xc), dtype=np.float64)
                if v.ndim == 0:
                    v = np.full(len(sel), float(v))
                v = v.reshape(-1)
                if v.shape[0] != len(sel):
                    v = None
                else:  # spot-check the vectorised call against the reference's point-wise convention
                    for j in sel[:: max(1, len(sel) // 4)][:4]:
                        pj = float(np.asarray(boundary_penalty(Qh[j].reshape(1, -1), Xc), dtype=np.float64).reshape(-1)[0])
                        if not np.isclose(pj, v[np.nonzero(sel == j)[0][0]], rtol=1e-12, atol=0.0):
                            v = None
                            break
            except Exception:  # noqa: BLE001  (hook written for one point only)
                v = None
            if v is None:
                v = np.array([float(np.asarray(boundary_penalty(Qh[j].reshape(1, -1), Xc), dtype=np.float64).reshape(-1)[0])
                              for j in sel])
            pen[sel] = v
        if not np.any(pen):
            return None
        return torch.from_numpy(pen).to(self.device)

    # ------------------------------------------------------------------ MSPE (the reference's second acquisition)
    def _mspe_matrix(self):
        """B_c of the dense MSPE scan for every cluster, packed like W ([npad, npad] blocks), and c0[c] (host).
        REF/cGP/acquisition.py:52-68 takes the joint posterior covariance of [x; X_c] (predict(return_cov=True),
        SK/_gpr.py:464-478, every block scaled by y_std^2) and forms  mspe = s_xx - s_xo S_oo s_xo^T.  With
        Ko = Matern(X_c, X_c) (no White term: the cross kernel), K^-1 = W^T W, G = I - K^-1 Ko, S~ = Ko + s2 I - Ko K^-1 Ko:
            s_xx = ys^2 (1 + s2 - k*^T K^-1 k*),  s_xo = ys^2 k*^T G,  S_oo = ys^2 S~
            mspe = c0 - k*^T B k*,   c0 = ys^2 (1 + s2),   B = ys^2 K^-1 + ys^6 G S~ G^T        (symmetric, per cluster)
        Five [n_c, n_c] DMMA products per cluster, once per fitted model (cached until the model changes)."""
        cache = self.model.get("_mspe")
        if cache is not None:
            return cache
        B = torch.zeros(int(self.moff[-1]), dtype=torch.float64, device=self.device)
        c0 = np.zeros(self.C)
        for c in range(self.C):
            n, npad = int(self.counts[c]), int(self.npad[c])
            th = self.model["theta_dev"][c].contiguous()
            Xt = self.Xs[int(self.offs[c]):int(self.offs[c + 1])]
            Ko = torch.zeros((npad, npad), dtype=torch.float64, device=self.device)
            Ko[:n, :n] = _lib.kernel_cross(Xt, Xt, th, self.kind)
            W = self.model["W"][int(self.moff[c]):int(self.moff[c + 1])].view(npad, npad)
            Wt = W.t().contiguous()
            Kinv = _lib.dgemm_nt(Wt, Wt)          # W^T W (pad block: identity)
            KoKinv = _lib.dgemm_nt(Ko, Kinv)      # Ko K^-1 (both symmetric)
            G = (-KoKinv).t().contiguous()        # G = I - K^-1 Ko
            G.diagonal().add_(1.0)
            Soo = Ko - _lib.dgemm_nt(KoKinv, Ko)
            noise = float(np.exp(self.theta[c][self.d]))
            Soo.diagonal()[:n].add_(noise)
            H = _lib.dgemm_nt(_lib.dgemm_nt(G, Soo), G)  # G S~ G^T
            ys = float(self.y_std[c])
            Bc = B[int(self.moff[c]):int(self.moff[c + 1])].view(npad, npad)
            Bc[:n, :n] = (ys ** 2) * Kinv[:n, :n] + (ys ** 6) * H[:n, :n]
            c0[c] = ys * ys * (1.0 + noise)
        self.model["_mspe"] = (B, c0)
        return B, c0

    def mspe_scan(self, Q, k=3, idx_base=0, want_arrays=False, qlabel=None, penalty=None):
        """Dense scan of REF/cGP/acquisition.py:34-75 over candidates Q: for every candidate x routed to cluster c,
        ``value = (mspe(x) + penalty) / n_c`` with ``mspe = s_xx - s_xo S_oo s_xo^T`` (see _mspe_matrix); the reference
        MINIMISES it, so the selected candidate is the minimum, ties -> lowest cluster id, then lowest candidate index.
        One fused pass per candidate chunk in the C library (cgp_quadform_scan: cross kernel, one DMMA product, row dot).
        -> ((value, idx, cluster) 1-element CUDA tensors, value array in candidate order | None)."""
        Qd = self._to_dev(Q)
        if qlabel is None:
            qlabel = self.route(Qd, k)
        if penalty is not None and not isinstance(penalty, torch.Tensor):
            penalty = torch.from_numpy(np.ascontiguousarray(np.asarray(penalty, dtype=np.float64).reshape(-1))).to(self.device)
        B, c0 = self._mspe_matrix()
        ws = self._workspace(_lib.quadform_workspace_bytes(self.offs, self.d, Qd.shape[0]))
        return _lib.quadform_scan(self.Xs, self.offs, self.model["theta_dev"], B, c0, Qd, qlabel, self.kind, idx_base,
                                  penalty, want_arrays, ws)

    def select_next_mspe(self, candidates, k=3, sharded=False, boundary_penalty=None):
        """MSPE counterpart of select_next: candidates sharded over ranks, each rank scans its shard, all-gather of the
        per-rank minima with the reference's tie rule, broadcast of the selected point.
        -> (x_next np[d], value, global idx, cluster)."""
        rank, ws = cdist.world()
        local, base = self._local_shard(candidates, sharded, rank, ws)
        ql = self.route(local, k)
        pen = None if boundary_penalty is None else self.penalty_of(local, ql, boundary_penalty)
        (bv, bi, bc), _ = self.mspe_scan(local, k, idx_base=base, qlabel=ql, penalty=pen)
        li = int(bi.item()) - base
        x_local = local[li] if li >= 0 else torch.zeros(self.d, dtype=torch.float64, device=self.device)
        neg, idx, cl, x, _owner = cdist.reduce_best(-bv, bi, bc, x_local, rank, ws)
        return x.cpu().numpy(), -neg, idx, cl

    def _local_shard(self, candidates, sharded, rank, ws):
        """This rank's candidates and the global index of its first one (see select_next)."""
        if sharded:
            local = self._to_dev(candidates)
            sizes = torch.tensor([local.shape[0]], dtype=torch.int64, device=self.device)
            if ws > 1:
                alls = torch.empty(ws, dtype=torch.int64, device=self.device)
                torch.distributed.all_gather_into_tensor(alls, sizes)
                return local, int(alls[:rank].sum().item())
            return local, 0
        lo, hi = cdist.shard_range(candidates.shape[0], rank, ws)
        return self._to_dev(candidates[lo:hi]), lo

    def select_next(self, candidates, k=3, sharded=False, prune=True, boundary_penalty=None):
        """Pick the next sample from `candidates` [M,d].

        sharded=False: `candidates` is the full array (identical on every rank); each rank scores its
        contiguous shard.  sharded=True: `candidates` is already this rank's shard and idx_base is derived
        from the shard sizes (all-gathered).  Returns (x_next np[d], score, global idx, cluster).
        prune=True (default): the exact bound-pruned scan (cgp_ei_argmax_pruned) — same selected candidate and the same
        score bits as the exhaustive scan (tests/test_gpu_prune.py); prune=False scores every candidate.
        boundary_penalty: the reference's soft-constraint hook (X, data_X) -> additive term (see penalty_of); it is added
        to EI inside the GPU score, (EI + penalty) / n_c."""
        rank, ws = cdist.world()
        local, base = self._local_shard(candidates, sharded, rank, ws)
        ql = self.route(local, k)
        pen = None if boundary_penalty is None else self.penalty_of(local, ql, boundary_penalty)
        (bs, bi, bc), _ = self.ei_scan(local, k, idx_base=base, prune=prune, qlabel=ql, penalty=pen)
        li = int(bi.item()) - base
        x_local = local[li] if li >= 0 else torch.zeros(self.d, dtype=torch.float64, device=self.device)
        score, idx, cl, x, _owner = cdist.reduce_best(bs, bi, bc, x_local, rank, ws)
        return x.cpu().numpy(), score, idx, cl
