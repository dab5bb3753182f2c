"""Lock-step L-BFGS-B over many independent (cluster, restart) problems.

The optimiser arithmetic stays on the host and is SciPy's own reverse-communication routine
``_lbfgsb.setulb`` driven exactly like ``scipy.optimize._lbfgsb_py._minimize_lbfgsb`` (SP/optimize/
_lbfgsb_py.py:272-461: m=10, factr=ftol/eps with ftol=2.22e-9, pgtol=1e-5, maxls=20, maxiter=maxfun=15000),
which is what ``GaussianProcessRegressor._constrained_optimization`` calls (SK/gaussian_process/_gpr.py:658-674).
Instead of one objective call per run, every run that is waiting for (f, g) is evaluated in ONE batched GPU
launch per lock-step.  With many runs the set is split in two groups that alternate: while the GPU evaluates one
group the host advances the state machines of the other, so the host time disappears from the critical path.
Every run still sees exactly its own sequence of (x -> f, g) evaluations, so results do not depend on the grouping.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg.lapack import HAS_ILP64
from scipy.optimize import _lbfgsb

_INT = np.int64 if HAS_ILP64 else np.int32


def _check_scipy():
    """`_lbfgsb.setulb` is private SciPy API.  The positional signature and the integer task codes used below
    (task[0] == 3: "FG", 1: "NEW_X", ln_task, no iprint/csave) are those of the C port shipped since SciPy 1.15
    (SP/optimize/_lbfgsb_py.py:389-461); the Fortran wrapper of older releases takes different arguments.  Fail with a
    clear message instead of an opaque TypeError deep inside a fit.  Upper bound: the newest release this was run on."""
    import scipy

    from ._lib import CgpError

    try:
        major, minor = (int(v) for v in scipy.__version__.split(".")[:2])
    except ValueError as e:  # pragma: no cover
        raise CgpError(f"cannot parse the SciPy version {scipy.__version__!r}") from e
    if not ((1, 15) <= (major, minor) <= (1, 18)):
        raise CgpError(f"cgp_b200.lbfgs drives scipy.optimize._lbfgsb.setulb with the SciPy 1.15-1.18 calling convention; "
                       f"found SciPy {scipy.__version__} (pin scipy>=1.15,<1.19)")


_check_scipy()


class _Run:
    """One L-BFGS-B state machine (the body of _minimize_lbfgsb's while-loop, turned inside out)."""

    __slots__ = ("x", "f", "g", "lo", "hi", "nbd", "wa", "iwa", "task", "ln_task", "lsave", "isave", "dsave", "m",
                 "factr", "pgtol", "maxls", "maxiter", "maxfun", "nit", "nfev", "done", "best_f", "best_x")

    def __init__(self, x0, lo, hi, m=10, ftol=2.2204460492503131e-09, gtol=1e-5, maxfun=15000, maxiter=15000,
                 maxls=20):
        n = x0.shape[0]
        self.lo = np.array(lo, dtype=np.float64)
        self.hi = np.array(hi, dtype=np.float64)
        self.x = np.clip(np.array(x0, dtype=np.float64), self.lo, self.hi)  # _lbfgsb_py.py:360
        self.nbd = np.full(n, 2, dtype=_INT)  # both bounds finite
        self.f = np.array(0.0, dtype=np.float64)
        self.g = np.zeros(n, dtype=np.float64)
        self.m = m
        self.factr = ftol / np.finfo(float).eps
        self.pgtol = gtol
        self.maxls = maxls
        self.maxiter = maxiter
        self.maxfun = maxfun
        self.wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m, np.float64)
        self.iwa = np.zeros(3 * n, dtype=_INT)
        self.task = np.zeros(2, dtype=_INT)
        self.ln_task = np.zeros(2, dtype=_INT)
        self.lsave = np.zeros(4, dtype=_INT)
        self.isave = np.zeros(44, dtype=_INT)
        self.dsave = np.zeros(29, dtype=np.float64)
        self.nit = 0
        self.nfev = 0
        self.done = False
        self.best_f = np.inf   # best evaluated point so far: the result of a run cut off by max_steps
        self.best_x = self.x.copy()

    def advance(self):
        """Run setulb until it asks for (f, g) at self.x (returns True) or terminates (returns False)."""
        while True:
            self.g = self.g.astype(np.float64)
            _lbfgsb.setulb(self.m, self.x, self.lo, self.hi, self.nbd, self.f, self.g, self.factr, self.pgtol, self.wa,
                           self.iwa, self.task, self.lsave, self.isave, self.dsave, self.maxls, self.ln_task)
            if self.task[0] == 3:
                return True
            if self.task[0] == 1:
                self.nit += 1
                if self.nit >= self.maxiter:
                    self.task[0] = 5
                    self.task[1] = 504
                elif self.nfev > self.maxfun:
                    self.task[0] = 5
                    self.task[1] = 502
            else:
                self.done = True
                return False

    def feed(self, f, g):
        self.f = float(f)
        self.g = np.array(g, dtype=np.float64)
        self.nfev += 1
        if self.f < self.best_f:
            self.best_f, self.best_x = self.f, self.x.copy()


def minimize_lockstep(eval_batch, x0s, lo, hi, max_steps=100000, submit=None, collect=None, pipeline_min=256,
                      shard=None):
    """Minimise len(x0s) independent problems.

    eval_batch(idx: int array [b], X: float array [b, n]) -> (f [b], g [b, n]) evaluates problems `idx` at `X`.
    Optional asynchronous form: ``submit(idx, X) -> ticket`` enqueues the evaluation, ``collect(ticket) -> (f, g)``
    waits for it; with both given and at least `pipeline_min` runs waiting, two groups alternate so that host and
    device work overlap.
    Multi-rank form: ``shard = dict(rank, world, cost [R], allreduce)``.  Every rank holds ALL state machines; in each
    lock-step the waiting runs are dealt to the ranks (heaviest first, round-robin), each rank evaluates only its share
    and ``allreduce`` (sum over ranks of a [b, n+1] array whose foreign rows are zero) hands every rank all (f, g).
    The machines then advance identically everywhere, so the work stays balanced as runs finish at different times.
    max_steps: stop after that many lock-steps; unfinished runs return their best evaluated point.
    Returns (x [R, n], f [R], nfev [R], nit [R], n_lockstep)."""
    runs = [_Run(np.asarray(x0), lo, hi) for x0 in x0s]
    waiting = [i for i, r in enumerate(runs) if r.advance()]
    steps = 0

    def take(group):
        idx = np.asarray(group, dtype=np.int64)
        return idx, np.stack([runs[i].x.copy() for i in group])

    def absorb(group, f, g):
        nxt = []
        for k, i in enumerate(group):
            runs[i].feed(f[k], g[k])
            if runs[i].advance():
                nxt.append(i)
        return nxt

    if submit is not None and collect is not None and len(waiting) >= pipeline_min and not (
            shard is not None and shard["world"] > 1):
        # two alternating groups; falls back to one group once few runs remain
        half = len(waiting) // 2
        ga, gb = waiting[:half], waiting[half:]
        ta = submit(*take(ga))
        while steps < max_steps:
            tb = submit(*take(gb)) if gb else None
            ga = absorb(ga, *collect(ta))  # host works on A while the device runs B
            steps += 1
            if len(ga) + len(gb) < pipeline_min:
                if tb is not None:
                    gb = absorb(gb, *collect(tb))
                waiting = ga + gb
                break
            ta = submit(*take(ga)) if ga else None
            if tb is not None:
                gb = absorb(gb, *collect(tb))  # host works on B while the device runs A
            if ta is None:
                waiting = gb
                break
        else:
            waiting = []
    sharded = shard is not None and shard["world"] > 1
    while waiting and steps < max_steps:
        idx, X = take(waiting)
        if sharded:
            order = np.argsort(-np.asarray(shard["cost"])[idx], kind="stable")
            mine = np.sort(order[shard["rank"]::shard["world"]])
            buf = np.zeros((len(idx), X.shape[1] + 1))
            if len(mine):
                fm, gm = eval_batch(idx[mine], X[mine])
                buf[mine, 0] = fm
                buf[mine, 1:] = gm
            buf = shard["allreduce"](buf)
            f, g = buf[:, 0], buf[:, 1:]
        else:
            f, g = eval_batch(idx, X)
        waiting = absorb(waiting, f, g)
        steps += 1
    # finished runs: SciPy's result (x, f of the last accepted iterate); runs cut off by max_steps: best evaluated point
    xs = np.stack([r.x if r.done else r.best_x for r in runs])
    fs = np.array([float(r.f) if r.done else float(r.best_f) for r in runs])
    return xs, fs, np.array([r.nfev for r in runs]), np.array([r.nit for r in runs]), steps
