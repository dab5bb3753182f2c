"""Multi-GPU comparison of the two fit sharding modes on config 3 inside ONE torchrun launch (GPU time is charged per
GPU, so the import / workload set-up is paid once):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/n8_modes.py [--reps 2]
Prints one JSON line per mode from rank 0: iteration / fit / acquisition seconds (max over ranks), evaluations per rank."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import PRESETS, make_workload  # noqa: E402
from cgp_b200 import _lib  # noqa: E402
from cgp_b200.engine import ClusterGPEngine  # noqa: E402


def main():
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 2
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    p = PRESETS["cfg3"]
    X, y, labels, Q = make_workload(p)
    lo, hi = (p["M"] * rank) // world, (p["M"] * (rank + 1)) // world
    Qd = torch.from_numpy(Q[lo:hi]).to(dev)
    del Q
    _lib.handle(local)

    def step(mode):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng = ClusterGPEngine(X, y, labels, "matern32", 1e-5)
        eng.fit(p["R"], seed=0, shard_mode=mode)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        x, score, idx, cl = eng.select_next(Qd, k=p["k"], sharded=True)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        v = torch.tensor([t2 - t0, t1 - t0, t2 - t1, eng.fit_stats["eval_s"]], dtype=torch.float64, device=dev)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        ev = torch.zeros(world, dtype=torch.float64, device=dev)
        ev[rank] = eng.fit_stats["evals"]
        dist.all_reduce(ev)
        return v.tolist(), ev.tolist(), (idx, score, eng.fit_stats["lockstep"])

    step("static")  # warm-up (allocations, NCCL channels)
    for mode in ["static", "balanced"] * reps:
        v, ev, sel = step(mode)
        if rank == 0:
            print(json.dumps(dict(mode=mode, n_gpus=world, iteration_s=v[0], fit_s=v[1], acq_s=v[2], eval_wall_s_max=v[3],
                                  evals_per_rank=ev, selected_idx=sel[0], score=sel[1], lockstep_rank0=sel[2])), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
