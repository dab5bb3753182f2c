mkdir -p gpurun_out
python bench.py --preset cfg5 --no-exhaustive > gpurun_out/r02k_bench_cfg5.json 2> gpurun_out/r02k_bench_cfg5.err; echo "cfg5 rc=$?"
python bench.py --preset cfg2 > gpurun_out/r02k_bench_cfg2.json 2> gpurun_out/r02k_bench_cfg2.err; echo "cfg2 rc=$?"
python bench.py --preset small --steps 1 --warmup 1 > gpurun_out/r02k_bench_small.json 2> gpurun_out/r02k_bench_small.err; echo "small rc=$?"
for f in gpurun_out/r02k_bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read() or '{}')
print({k:d.get(k) for k in ('value','ms_per_step','fit_s','acq_s','fit_optimizer_host_s','fit_eval_wall_s','cpu_comparison_fully_measured')}, (d.get('e2e') or {}).get('ms_per_step'))
"; done
