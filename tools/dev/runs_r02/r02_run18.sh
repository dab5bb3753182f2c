python -m pytest tests/test_gpu_multirank.py -m gpu -q 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --preset cfg5 --steps 2 --warmup 1 --no-exhaustive 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ('value','ms_per_step','fit_s','acq_s','fit_async_rank0','selected')})"
