mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest_gpu.log
tail -25 gpurun_out/r02c_pytest_gpu.log
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 "$@"; }
run2 --steps 2 --warmup 1 > gpurun_out/r02c_bench_cfg3_n2_async.json 2> gpurun_out/r02c_bench_cfg3_n2_async.err; echo "n2 async rc=$?"; tail -3 gpurun_out/r02c_bench_cfg3_n2_async.err
CGP_SHARD_MODE=balanced run2 --steps 2 --warmup 1 --no-exhaustive > gpurun_out/r02c_bench_cfg3_n2_balanced.json 2> gpurun_out/r02c_bench_cfg3_n2_balanced.err; echo "n2 bal rc=$?"
python bench.py --preset cfg5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02c_bench_cfg5.json 2> gpurun_out/r02c_bench_cfg5.err; echo "cfg5 rc=$?"
run2 --preset cfg5 --steps 1 --warmup 1 --no-exhaustive > gpurun_out/r02c_bench_cfg5_n2.json 2> gpurun_out/r02c_bench_cfg5_n2.err; echo "cfg5 n2 rc=$?"
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02c_bench_cfg3.json 2> gpurun_out/r02c_bench_cfg3.err
for f in gpurun_out/r02c_bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read() or '{}')
print({k:d.get(k) for k in ('value','ms_per_step','fit_s','acq_s','lml_evals_per_fit_rank0','lockstep_per_fit','shard_mode')})
print(d.get('acq_exhaustive_scan'), (d.get('e2e') or {}).get('ms_per_step'), d.get('selected'))
"; done
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_detail_cfg5_n1.json'))
for k,v in d['kernel_classes'].items():
    if v['launches']: print(k, v)
PY
