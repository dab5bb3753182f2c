"""Multi-GPU sharding of the hot path (one process per GPU, torch.distributed).

The path shards on its natural units (REF/cGP/cGP_parallel.py:396-407 is the precedent for
"clusters are independent"):
  * fit: the C*(R+1) (cluster, restart) L-BFGS-B runs are dealt to ranks; in the default "async" mode
    (engine.fit, lbfgs.minimize_async / StealBoard) every rank owns whole runs, idle ranks steal ready runs over
    the c10d store, and the only collective is one all-reduce of the per-run (theta*, -lml) table (a few KB);
    the helpers below serve the "static" mode (deal once, one all-gather of each rank's rows);
  * scoring: candidates are split into contiguous equal shards; the only exchange is one all-gather of each
    rank's (best score, candidate index, cluster) and one broadcast of the selected point.
All helpers work on CPU tensors with the gloo backend too (tests/test_dist_gloo.py).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(m, rank, world_size):
    """Contiguous candidate shard [lo, hi) of rank: [r*M/W, (r+1)*M/W)."""
    return (m * rank) // world_size, (m * (rank + 1)) // world_size


def my_pairs(n_pairs, rank, world_size):
    """Round-robin deal of (cluster, restart) pairs."""
    return list(range(rank, n_pairs, world_size))


def gather_pair_rows(local_rows, n_pairs, rank, world_size):
    """local_rows [len(my_pairs), P+1] (theta*, f) -> full table [n_pairs, P+1] on every rank."""
    if world_size == 1:
        return local_rows
    width = local_rows.shape[1]
    per = (n_pairs + world_size - 1) // world_size
    buf = torch.zeros((per, width), dtype=local_rows.dtype, device=local_rows.device)
    buf[: local_rows.shape[0]] = local_rows
    out = torch.empty((world_size * per, width), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(out, buf)
    out = out.view(world_size, per, width)
    full = torch.empty((n_pairs, width), dtype=local_rows.dtype, device=local_rows.device)
    for r in range(world_size):
        idx = my_pairs(n_pairs, r, world_size)
        full[idx] = out[r, : len(idx)]
    return full


def reduce_best(score, idx, cluster, x_local, rank, world_size):
    """All-gather each rank's (score, global candidate index, cluster) and pick the winner with the
    reference's tie rule (max score; then lowest cluster id; then lowest candidate index —
    REF/cGP/cGP_constrained.py:305,376-378), then broadcast the winning point from its owner.

    score/idx/cluster: 1-element tensors; x_local [d] the local winner's coordinates.
    Returns (score float, idx int, cluster int, x [d] tensor, owner_rank)."""
    if world_size == 1:
        return float(score.item()), int(idx.item()), int(cluster.item()), x_local, 0
    dev = x_local.device
    mine = torch.stack([score.to(torch.float64).reshape(()), idx.to(torch.float64).reshape(()),
                        cluster.to(torch.float64).reshape(())]).to(dev)
    # indices < 2^53 are exact in fp64, so one fp64 all-gather carries all three values
    allv = torch.empty((world_size, 3), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(allv, mine.reshape(1, 3))
    rows = allv.cpu().tolist()
    owner = -1
    best = None
    for r, (s, i, c) in enumerate(rows):
        if i < 0 or s != s:
            continue
        key = (-s, c, i)
        if best is None or key < best:
            best, owner = key, r
    if owner < 0:
        raise RuntimeError("no rank produced a valid candidate")
    x = x_local.clone()
    dist.broadcast(x, src=owner)
    return -best[0], int(best[2]), int(best[1]), x, owner
