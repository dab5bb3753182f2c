mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
tail -15 gpurun_out/r02_pytest_gpu.log
python bench.py --steps 1 --warmup 1 > gpurun_out/r02b_bench_cfg3.json 2> gpurun_out/r02b_bench_cfg3.err; echo "bench rc=$?"; tail -3 gpurun_out/r02b_bench_cfg3.err
for fm in 48 96 1000000; do
CGP_FORK_MAX=$fm python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-exhaustive > gpurun_out/r02b_bench_cfg3_fork$fm.json 2> gpurun_out/r02b_bench_cfg3_fork$fm.err
done
CGP_LANES=1 CGP_FORK_MAX=1000000 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-exhaustive > gpurun_out/r02b_bench_cfg3_lanes1.json 2> gpurun_out/r02b_bench_cfg3_lanes1.err
python bench.py --preset cfg5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02b_bench_cfg5.json 2> gpurun_out/r02b_bench_cfg5.err; echo "cfg5 rc=$?"
python bench.py --preset cfg4 --steps 1 --warmup 1 > gpurun_out/r02b_bench_cfg4.json 2> gpurun_out/r02b_bench_cfg4.err; echo "cfg4 rc=$?"; tail -5 gpurun_out/r02b_bench_cfg4.err
for f in gpurun_out/r02b_bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read() or '{}')
print({k:d.get(k) for k in ('value','ms_per_step','fit_s','acq_s','parity','lml_evals_per_fit_rank0','lockstep_per_fit','shard_mode')})
r=d.get('roofline') or {}
print({k:r.get(k) for k in ('kernel','achieved','peak','frac')}, (r.get('step') or {}).get('frac'), (r.get('step') or {}).get('fit_only_frac')); print(r.get('classes'))
print(d.get('acq_exhaustive_scan'), (d.get('e2e') or {}).get('ms_per_step'), d.get('selected'))
"; done
