mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_driver.py tests/test_gpu_prune.py -m gpu -q -k "knn or mspe or penalty or constraint or prune" > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02d_pytest.log
python bench.py --preset cfg5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02d_bench_cfg5.json 2> gpurun_out/r02d_bench_cfg5.err; echo "cfg5 rc=$?"
python bench.py --preset cfg4 --steps 1 --warmup 1 > gpurun_out/r02d_bench_cfg4.json 2> gpurun_out/r02d_bench_cfg4.err; echo "cfg4 rc=$?"; tail -5 gpurun_out/r02d_bench_cfg4.err
python tools/ncu_target_r02.py > gpurun_out/r02d_target_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_assemble|k_knn_tc|k_knn|k_cross|k_grad_tiles" -c 8 -o gpurun_out/r02d_full_assemble_knn python tools/ncu_target_r02.py > gpurun_out/r02d_target_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02d_target_ncu.log
for f in gpurun_out/r02d_bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read() or '{}')
print({k:d.get(k) for k in ('value','ms_per_step','fit_s','acq_s','lml_evals_per_fit_rank0','lockstep_per_fit','shard_mode','parity','fit_async_rank0')})
print(d.get('acq_exhaustive_scan'), (d.get('e2e') or {}).get('ms_per_step'), d.get('selected')); print(d.get('roofline'))
"; done
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_detail_cfg5_n1.json'))
for k,v in d['kernel_classes'].items():
    if v['launches'] and v['ms']: print(k, v)
PY
