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
                 "factr", "pgtol", "maxls", "maxiter", "maxfun", "nit", "nfev", "done", "best_f", "best_x", "gid")

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
        self.gid = -1          # global run id (minimize_async: runs migrate between ranks)

    def advance(self):
        """Run setulb until it asks for (f, g) at self.x (returns True) or terminates (returns False)."""
        while True:  # self.g is always a private float64 array (feed copies), so SciPy's per-iteration astype is not needed
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


# ======================================================================================================================
# Asynchronous form: no global lock-step, several evaluation lanes per rank, work stealing between ranks
# ======================================================================================================================
class StealBoard:
    """Work stealing between ranks over the c10d key-value store (the TCPStore every torch.distributed job already has).

    Every rank owns whole L-BFGS-B runs and advances them without waiting for any other rank.  A rank whose number of
    active runs falls two or more below the busiest rank's posts a request to that rank; the victim answers at its next
    step boundary with the pickled state machines of some of its READY runs (an L-BFGS-B machine is ~15 KB), which
    then continue on the thief.  A run's sequence of (theta -> f, g) evaluations does not depend on where or in which
    batch it is evaluated (the kernels are batch-invariant bit for bit), so the result is independent of the schedule.
    Nobody leaves the loop before all runs of all ranks are finished (global counter), so every request is answered.

    All store traffic runs on a BACKGROUND THREAD (the c10d store calls release the GIL): it publishes this rank's load,
    mirrors the other ranks' loads, the global finished-runs counter and this rank's request queue into plain attributes,
    posts requests and fetches replies.  The optimiser loop only reads those attributes, so a store round trip (20-200 us,
    about ten per step boundary with 8 ranks) never sits between two evaluations.  (Measured on 8 B200, config 3: 2.12 s
    per iteration with the thread, 2.14 s with the same traffic on the optimiser thread — the store was not the limit.)

    Keys (under a per-fit prefix): load/<r> counter = active runs of rank r; done counter = finished runs (global);
    req/<v> counter = requests posted to victim v; who/<v>/<k> = "<thief>:<thief load>"; give/<v>/<k> = pickled runs."""

    def __init__(self, store, prefix, rank, world, total_runs, period=2e-4):
        import threading

        self.store, self.prefix, self.rank, self.world, self.total = store, prefix, rank, world, total_runs
        self.stolen = 0
        self.given = 0
        # main thread -> I/O thread
        self._load_target = 0      # active runs of this rank
        self._done_add = 0         # finished runs not yet added to the global counter
        self._want = None          # (victim, my_load): post a request
        self._outbox = []          # (k, pickled runs): replies to publish
        # I/O thread -> main thread
        self._loads = [0] * world  # mirror of the load counters
        self._done = 0             # mirror of the global counter
        self._inbox = []           # (k, thief_load): requests to answer
        self._reply = None         # runs received for the pending request
        self._pending = None       # (victim, k) of the one request in flight
        self._lock = threading.Lock()
        self._stop = False
        self._err = None
        self._period = period
        for r in range(world):     # make every key this thread reads exist
            store.add(self._k(f"load/{r}"), 0)
        store.add(self._k("done"), 0)
        store.add(self._k(f"req/{rank}"), 0)
        self._thread = threading.Thread(target=self._io_loop, daemon=True)
        self._thread.start()

    def _k(self, name):
        return f"{self.prefix}/{name}"

    # ------------------------------------------------------------------ I/O thread
    def _io_loop(self):
        import pickle
        import time

        store, published, served = self.store, 0, 0
        load_keys = [self._k(f"load/{r}") for r in range(self.world)]
        multi = hasattr(store, "multi_get")
        try:
            while not self._stop:
                with self._lock:
                    target, add, self._done_add = self._load_target, self._done_add, 0
                    out, self._outbox = self._outbox, []
                    want, self._want = self._want, None
                if target != published:
                    store.add(self._k(f"load/{self.rank}"), target - published)
                    published = target
                for k, blob in out:
                    store.set(self._k(f"give/{self.rank}/{k}"), blob)
                self._done = store.add(self._k("done"), add)
                cnt = store.add(self._k(f"req/{self.rank}"), 0)
                while served < cnt:
                    served += 1
                    who = store.get(self._k(f"who/{self.rank}/{served}")).decode()
                    self._drop(self._k(f"who/{self.rank}/{served}"))
                    with self._lock:
                        self._inbox.append((served, int(who.split(":")[1])))
                if self._pending is not None and self._reply is None:
                    v, k = self._pending
                    key = self._k(f"give/{v}/{k}")
                    if store.check([key]):
                        self._reply = pickle.loads(store.get(key))
                        self._drop(key)  # the blob (~15 KB per run) must not pile up in the store of a long BO loop
                elif want is not None and self._pending is None:
                    v, my_load = want
                    k = store.add(self._k(f"req/{v}"), 1)
                    store.set(self._k(f"who/{v}/{k}"), f"{self.rank}:{my_load}")
                    self._pending = (v, k)
                if multi:
                    try:  # one round trip for all counters (ASCII integers)
                        vals = [int(v) for v in store.multi_get(load_keys)]
                    except (RuntimeError, NotImplementedError, ValueError):  # store without multi_get (FileStore ...)
                        multi = False
                if not multi:
                    vals = [store.add(k, 0) for k in load_keys]
                self._loads = [v if r != self.rank else -1 for r, v in enumerate(vals)]
                if self._done >= self.total:
                    break
                time.sleep(self._period)
        except BaseException as e:  # noqa: BLE001  (surface store failures in the optimiser loop instead of hanging)
            self._err = e

    def _drop(self, key):
        try:
            self.store.delete_key(key)
        except Exception:  # noqa: BLE001  (stores without delete_key: the key simply stays)
            pass

    def _check(self):
        if self._err is not None:
            raise RuntimeError("work-stealing store thread failed") from self._err

    # ------------------------------------------------------------------ optimiser-loop side (never touches the store)
    def set_load(self, n):
        self._load_target = n

    def add_done(self, k):
        if k:
            with self._lock:
                self._done_add += k

    def all_done(self):
        self._check()
        return self._done >= self.total

    def maybe_request(self, my_load):
        """Ask the busiest rank for work if it holds at least two runs more than this rank."""
        if self._pending is not None or self._want is not None:
            return
        loads = self._loads
        v = int(np.argmax(loads))
        if loads[v] >= my_load + 2:
            self._want = (v, my_load)

    def poll_reply(self):
        """-> list of stolen runs (possibly empty) when the pending request has been answered, else None."""
        runs = self._reply
        if runs is None:
            return None
        self._reply = None
        self._pending = None
        self.stolen += len(runs)
        return runs

    def serve(self, ready, my_total):
        """Victim side, called at a step boundary: answer every queued request out of `ready` (list of runs, mutated)."""
        import pickle

        if not self._inbox:
            return my_total
        with self._lock:
            reqs, self._inbox = self._inbox, []
        for k, thief_load in reqs:
            give = max(0, min(len(ready), (my_total - thief_load) // 2))
            out = [ready.pop() for _ in range(give)]
            my_total -= give
            self.given += give
            blob = pickle.dumps(out, protocol=pickle.HIGHEST_PROTOCOL)
            with self._lock:
                self._outbox.append((k, blob))
        return my_total

    def close(self):
        self._stop = True
        self._thread.join(timeout=5)
        self._check()


def minimize_async(x0s, gids, lo, hi, submit, done, collect, n_lanes=2, lane_policy=None, board=None, max_evals=None,
                   poll_sleep=50e-6):
    """Minimise the runs this rank starts with (x0s[i] has global id gids[i]); with a StealBoard runs migrate between
    ranks.  Evaluation interface (asynchronous, per lane):
        ticket = submit(lane, gid [b] int64, X [b, n])   enqueue the evaluation of runs `gid` at `X` on lane `lane`
        done(ticket) -> bool                             non-blocking
        collect(ticket) -> (f [b], g [b, n])             blocks until finished
    Every lane is an independent lock-step over the runs it currently holds; lanes are refilled from the pool of ready
    runs as soon as they are free, so the device always has the next batch queued while the host advances the state
    machines of the previous one, and the latency-bound part of one lane's factorisations overlaps the bulk of another's.
    lane_policy(n_active) -> number of lanes to use at the moment (default: always n_lanes).
    max_evals: stop every run after that many evaluations (it returns its best evaluated point).
    Returns (results {gid: (x, f, nfev, nit)} of the runs that FINISHED on this rank, stats dict)."""
    import time

    t_start = time.perf_counter()
    runs_ready = []
    finished = {}
    n_steps = 0
    n_evals = 0

    def retire(r):
        if r.done:
            finished[r.gid] = (r.x.copy(), float(r.f), r.nfev, r.nit)
        else:
            finished[r.gid] = (r.best_x.copy(), float(r.best_f), r.nfev, r.nit)

    new_done = 0
    for x0, gid in zip(x0s, gids):
        r = _Run(np.asarray(x0), lo, hi)
        r.gid = int(gid)
        if r.advance():
            runs_ready.append(r)
        else:
            retire(r)
            new_done += 1
    lanes = [None] * n_lanes  # per lane: (ticket, [runs]) while busy
    if board is not None:
        board.set_load(len(runs_ready))
        board.add_done(new_done)
    t_idle = 0.0
    while True:
        progressed = False
        in_flight = sum(len(l[1]) for l in lanes if l is not None)
        # ---- absorb finished lanes
        for li in range(n_lanes):
            if lanes[li] is None or not done(lanes[li][0]):
                continue
            ticket, batch = lanes[li]
            f, g = collect(ticket)
            lanes[li] = None
            n_steps += 1
            new_done = 0
            for k, r in enumerate(batch):
                r.feed(f[k], g[k])
                if max_evals is not None and r.nfev >= max_evals:
                    retire(r)
                    new_done += 1
                elif r.advance():
                    runs_ready.append(r)
                else:
                    retire(r)
                    new_done += 1
            in_flight -= len(batch)
            progressed = True
            if board is not None:
                board.add_done(new_done)
        # ---- work stealing: answer requests out of the ready pool, collect a reply, ask when short of work
        if board is not None:
            total = len(runs_ready) + in_flight
            if runs_ready or total == 0:
                total = board.serve(runs_ready, total)
            got = board.poll_reply()
            if got:
                runs_ready.extend(got)
                total += len(got)
                progressed = True
            board.set_load(total)
            board.maybe_request(total)
        # ---- refill free lanes from the ready pool
        if runs_ready:
            active = len(runs_ready) + in_flight
            use = n_lanes if lane_policy is None else max(1, min(n_lanes, lane_policy(active)))
            busy = sum(1 for l in lanes if l is not None)
            free = [li for li in range(n_lanes) if lanes[li] is None][:max(0, use - busy)]
            if free:
                per = -(-len(runs_ready) // len(free))
                for li in free:
                    batch, runs_ready = runs_ready[:per], runs_ready[per:]
                    if not batch:
                        break
                    gid = np.asarray([r.gid for r in batch], dtype=np.int64)
                    X = np.stack([r.x.copy() for r in batch])
                    n_evals += len(batch)
                    lanes[li] = (submit(li, gid, X), batch)
                    progressed = True
        # ---- termination
        if not runs_ready and all(l is None for l in lanes):
            if board is None:
                break
            if board.all_done():
                break
        if not progressed:
            t0 = time.perf_counter()
            time.sleep(poll_sleep)
            t_idle += time.perf_counter() - t0
    if board is not None:
        board.close()
    return finished, dict(steps=n_steps, evals=n_evals, idle_s=t_idle, loop_s=time.perf_counter() - t_start,
                          stolen=0 if board is None else board.stolen, given=0 if board is None else board.given)
