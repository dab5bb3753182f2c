"""CPU, world_size 2 over gloo: the multi-GPU plumbing (pair dealing, theta* all-gather, EI-maxima all-gather with the
reference's tie rule, broadcast of the selected point) gives the same answer as a single rank."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cgp_b200 import dist as cdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- fit exchange: 3 clusters x 5 runs, P = 4
        n_pairs, P = 15, 4
        rng = np.random.default_rng(0)
        table = rng.normal(size=(n_pairs, P + 1))
        mine = cdist.my_pairs(n_pairs, rank, world)
        full = cdist.gather_pair_rows(torch.from_numpy(table[mine]), n_pairs, rank, world)
        ok_fit = bool(np.array_equal(full.numpy(), table))
        # ---- scoring exchange: 1001 candidates, scores with an exact tie across the two shards
        M, d = 1001, 3
        Q = rng.uniform(size=(M, d))
        score = rng.uniform(size=M)
        cluster = rng.integers(0, 4, size=M)
        score[[100, 900]] = 2.0  # tie: lower cluster id wins, then lower index
        cluster[100], cluster[900] = 3, 1
        lo, hi = cdist.shard_range(M, rank, world)
        j = lo + int(np.argmax(score[lo:hi]))
        s, i, c, x, owner = cdist.reduce_best(torch.tensor([score[j]]), torch.tensor([j]), torch.tensor([cluster[j]]),
                                              torch.from_numpy(Q[j].copy()), rank, world)
        ok_sel = (i == 900 and c == 1 and s == 2.0 and np.array_equal(x.numpy(), Q[900]) and owner == 1)
        # ---- balanced sharded lock-step: replicated L-BFGS-B state machines, each rank evaluates its share
        from cgp_b200.lbfgs import minimize_lockstep

        n_runs, dim = 13, 3
        A = rng.uniform(0.5, 2.0, size=(n_runs, dim))
        cen = rng.uniform(-1, 1, size=(n_runs, dim))
        calls = [0]

        def fg(idx, Xs):
            calls[0] += len(idx)
            r = Xs - cen[idx]
            return (A[idx] * r ** 2).sum(1) + 0.1 * np.sin(3 * Xs).sum(1), 2 * A[idx] * r + 0.3 * np.cos(3 * Xs)

        x0 = [np.full(dim, 0.3 * (i % 5) - 0.5) for i in range(n_runs)]
        lo3, hi3 = np.full(dim, -3.0), np.full(dim, 3.0)
        plain = minimize_lockstep(fg, x0, lo3, hi3)
        total_plain = calls[0]
        calls[0] = 0

        def allreduce(buf):
            t = torch.from_numpy(buf.copy())
            dist.all_reduce(t)
            return t.numpy()

        sh = minimize_lockstep(fg, x0, lo3, hi3, shard=dict(rank=rank, world=world, cost=np.arange(n_runs) % 3 + 1.0,
                                                            allreduce=allreduce))
        ok_bal = all(np.array_equal(a, b) for a, b in zip(plain[:4], sh[:4])) and calls[0] < 0.7 * total_plain
        # ---- asynchronous runs with work stealing over the c10d store: rank 0 starts with ALL runs, rank 1 with none
        import time

        from torch.distributed.distributed_c10d import _get_default_store

        from cgp_b200.lbfgs import StealBoard, minimize_async

        board = StealBoard(_get_default_store(), "test/fit1", rank, world, n_runs)
        gids = np.arange(n_runs) if rank == 0 else np.arange(0)
        fin, st = minimize_async([x0[i] for i in gids], gids, lo3, hi3,
                                 lambda lane, gid, X: (time.perf_counter() + 3e-3, fg(gid, X)),
                                 lambda t: time.perf_counter() >= t[0], lambda t: t[1], n_lanes=2, board=board)
        table = np.zeros((n_runs, dim + 3))
        for gid, (x, f, nfev, nit) in fin.items():
            table[gid] = np.concatenate([x, [f, nfev, nit]])
        t = torch.from_numpy(table)
        dist.all_reduce(t)
        counts = torch.tensor([float(len(fin)), float(st["stolen"]), float(st["given"])])
        allc = [torch.zeros(3) for _ in range(world)]
        dist.all_gather(allc, counts)
        tot = t.numpy()
        ok_async = (np.array_equal(tot[:, :dim], plain[0]) and np.array_equal(tot[:, dim], plain[1])
                    and np.array_equal(tot[:, dim + 1], plain[2]) and np.array_equal(tot[:, dim + 2], plain[3])
                    and sum(int(c[0]) for c in allc) == n_runs          # every run finished exactly once
                    and int(allc[1][1]) > 0 and int(allc[0][2]) == int(allc[1][1]))  # rank 1 stole what rank 0 gave
        q.put((rank, ok_fit, ok_sel and ok_bal and ok_async, (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][3] == (0, 500) and res[1][3] == (500, 1001)
    for rank, ok_fit, ok_sel, _ in res:
        assert ok_fit and ok_sel, (rank, ok_fit, ok_sel)


def test_single_rank_passthrough():
    assert cdist.world() == (0, 1)
    assert cdist.shard_range(10, 0, 1) == (0, 10)
    assert cdist.my_pairs(5, 1, 2) == [1, 3]
    t = torch.arange(6.0).reshape(2, 3)
    assert cdist.gather_pair_rows(t, 2, 0, 1) is t
