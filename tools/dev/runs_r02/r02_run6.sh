mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_prune.py tests/test_gpu_fullsize.py tests/test_driver.py -m gpu -q -k "knn or prune or mspe or hygiene or cfg or penalty" > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02f_pytest.log
python bench.py --preset cfg5 --steps 1 --warmup 1 --no-cpu-baseline --no-exhaustive > gpurun_out/r02f_bench_cfg5.json 2> gpurun_out/r02f_bench_cfg5.err; echo "cfg5 rc=$?"
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-exhaustive > gpurun_out/r02f_bench_cfg3.json 2> gpurun_out/r02f_bench_cfg3.err; echo "cfg3 rc=$?"
python bench.py --preset cfg4 --steps 1 --warmup 1 > gpurun_out/r02f_bench_cfg4.json 2> gpurun_out/r02f_bench_cfg4.err; echo "cfg4 rc=$?"; tail -5 gpurun_out/r02f_bench_cfg4.err
for f in gpurun_out/r02f_bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read() or '{}')
print({k:d.get(k) for k in ('value','ms_per_step','fit_s','acq_s','lml_evals_per_fit_rank0','lockstep_per_fit','shard_mode','parity','fit_async_rank0')})
print(d.get('acq_exhaustive_scan'), (d.get('e2e') or {}).get('ms_per_step'), d.get('selected')); print(d.get('roofline')); print(d.get('cpu_baseline'))
"; done
python - <<'PY'
import json
for p in ('cfg5','cfg3','cfg4'):
    try:
        d=json.load(open(f'gpurun_out/bench_detail_{p}_n1.json'))
    except Exception as e:
        print(p, e); continue
    for k,v in d['kernel_classes'].items():
        if v['launches'] and v['ms']: print(p, k, v)
PY
