mkdir -p gpurun_out
python bench.py --preset small --steps 1 --warmup 1 > gpurun_out/r02n_small.json 2> gpurun_out/r02n_small.err; echo "small rc=$?"; tail -2 gpurun_out/r02n_small.err
python -c "
s=open('gpurun_out/r02n_small.json').read(); print('line bytes', len(s)); print(s[:400])"
python -m pytest tests/test_gpu_multirank.py -m gpu -q 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02n_bench_cfg3_n2.json 2> gpurun_out/r02n_bench_cfg3_n2.err; echo "n2 rc=$?"
python -c "
import json
s=open('gpurun_out/r02n_bench_cfg3_n2.json').read(); d=json.loads(s)
print('line bytes', len(s), {k:d.get(k) for k in ('value','ms_per_step','fit_s','acq_s','fit_async_rank0','selected','acq_exhaustive_scan')})"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 2>/dev/null | cut -c1-300
