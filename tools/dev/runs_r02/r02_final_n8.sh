mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02_final_bench_cfg3_n8.json 2> gpurun_out/r02_final_bench_cfg3_n8.err; echo "n8 rc=$?"; tail -2 gpurun_out/r02_final_bench_cfg3_n8.err
python -c "
import json
s=open('gpurun_out/r02_final_bench_cfg3_n8.json').read(); d=json.loads(s)
print('line bytes', len(s), {k:d.get(k) for k in ('value','ms_per_step','fit_s','acq_s','fit_async_rank0','selected','acq_exhaustive_scan','clocks')}, d['e2e'])"
