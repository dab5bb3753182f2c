mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02l_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02l_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --preset cfg2 > gpurun_out/r02l_bench_cfg2.json 2> gpurun_out/r02l_bench_cfg2.err; echo "cfg2 rc=$?"
python bench.py --preset cfg5 --no-exhaustive --no-cpu-baseline > gpurun_out/r02l_bench_cfg5.json 2> gpurun_out/r02l_bench_cfg5.err; echo "cfg5 rc=$?"
python bench.py --steps 1 --warmup 1 > gpurun_out/r02l_bench_cfg3.json 2> gpurun_out/r02l_bench_cfg3.err; echo "cfg3 rc=$?"
for f in gpurun_out/r02l_bench_*.json; do echo $f; python -c "
import json,sys
s=open('$f').read(); d=json.loads(s or '{}')
print(len(s), {k:d.get(k) for k in ('value','ms_per_step','fit_s','acq_s','fit_optimizer_host_s','selected')}, (d.get('e2e') or {}).get('ms_per_step'))
"; done
