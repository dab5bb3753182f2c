mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=20 -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
tail -30 gpurun_out/r02_pytest_gpu.log
python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s > gpurun_out/r02_pytest_fullsize.log 2>&1; tail -25 gpurun_out/r02_pytest_fullsize.log
python bench.py --steps 1 --warmup 1 > gpurun_out/r02a_bench_cfg3.json 2> gpurun_out/r02a_bench_cfg3.err; echo "bench rc=$?"; tail -3 gpurun_out/r02a_bench_cfg3.err
CGP_FORK_MAX=1000000 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-exhaustive > gpurun_out/r02a_bench_cfg3_forkall.json 2> gpurun_out/r02a_bench_cfg3_forkall.err
CGP_FORK_MAX=0 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-exhaustive > gpurun_out/r02a_bench_cfg3_nofork.json 2> gpurun_out/r02a_bench_cfg3_nofork.err
python bench.py --preset cfg5 --steps 1 --warmup 1 > gpurun_out/r02a_bench_cfg5.json 2> gpurun_out/r02a_bench_cfg5.err; echo "cfg5 rc=$?"
python bench.py --preset cfg2 --steps 2 --warmup 2 > gpurun_out/r02a_bench_cfg2.json 2> gpurun_out/r02a_bench_cfg2.err; echo "cfg2 rc=$?"
for f in gpurun_out/r02a_bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read() or '{}')
print({k:d.get(k) for k in ('value','ms_per_step','fit_s','acq_s','parity','selected','lml_evals_per_fit_rank0','lockstep_per_fit')})
r=d.get('roofline') or {}
print({k:r.get(k) for k in ('kernel','achieved','peak','frac','step')}); print(r.get('classes'))
print(d.get('acq_exhaustive_scan'), (d.get('e2e') or {}).get('ms_per_step'))
"; done
