mkdir -p gpurun_out
python -m pytest tests/test_gpu_multirank.py -m gpu -q > gpurun_out/r02h_pytest_multirank.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02h_pytest_multirank.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 1 --no-exhaustive > gpurun_out/r02h_bench_cfg3_n2.json 2> gpurun_out/r02h_bench_cfg3_n2.err; echo "n2 rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02h_bench_cfg3_n2.json').read())
print({k:d.get(k) for k in ('value','ms_per_step','fit_s','acq_s','fit_async_rank0','selected')})"
